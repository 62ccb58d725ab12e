"""Where does a C1 training step go?  Times (a) the host sampler alone, (b) Trainer.train_epoch three times,
(c) graph replays alone without the sampler, on one B200.  usage: python profiles/scripts/r02_epoch_c1_probe.py"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gnn_recommendations_b200 as g  # noqa: E402
from gnn_recommendations_b200.synthetic import synth_pairs_device  # noqa: E402

dev = torch.device("cuda:0")
nu, ni, e, d, L = 6040, 3706, 1000209, 64, 3
u, i = synth_pairs_device(nu, ni, e, 42, dev)
empty = (np.zeros(0, np.int64), np.zeros(0, np.int64))
ds = g.InteractionDataset((u.cpu().numpy(), i.cpu().numpy()), empty, empty, nu, ni, device=dev, name="C1")
torch.manual_seed(42)
model = g.LightGCN(nu, ni, embedding_dim=d, n_layers=L, init_scale=0.1)
tr = g.Trainer(model, ds, {"batch_size": 512, "use_scheduler": False, "checkpoint_dir": "/tmp/gr_probe_ckpt"}, device=dev)
steps = len(u) // 512 + 1
s = tr._get_sampler()
t0 = time.perf_counter()
for _ in range(steps):
    s.sample(512)
print(f"sampler alone: {(time.perf_counter() - t0) / steps * 1e3:.4f} ms per batch", flush=True)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = tr.train_epoch()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    print(f"epoch {rep}: {sec:.3f} s, {sec / steps * 1e3:.4f} ms per step, loss {loss:.6f}, graph={tr._graph is not None}", flush=True)
gs = tr._graph
if gs is not None:
    with torch.cuda.stream(tr._train_stream):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            gs.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"graph replay alone (device): {e0.elapsed_time(e1) / 500:.4f} ms per step", flush=True)
        t0 = time.perf_counter()
        for _ in range(500):
            gs.run()
        gs.finish()
        print(f"graph.run incl. sampler + staging: {(time.perf_counter() - t0) / 500 * 1e3:.4f} ms per step", flush=True)
print("threads", torch.get_num_threads(), "cpus", os.cpu_count(), "load", os.getloadavg())
