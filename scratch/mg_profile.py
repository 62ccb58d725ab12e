"""Per-rank phase timing of the fused multi-GPU propagation (run under torchrun)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recommendations_b200 as g
from gnn_recommendations_b200 import _lib
from gnn_recommendations_b200.dist import PeerExchange, RowPartition
from gnn_recommendations_b200.synthetic import synth_pairs_device
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
nu, ni, e, d, L = bench.WORKLOADS[wl]
u, i = synth_pairs_device(nu, ni, e, 42, dev)
full = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=dev); del u, i
part = RowPartition(full.indptr, world)
csr = part.local_csr(full, rank); nloc = part.n_local(rank); del full; torch.cuda.empty_cache()
x0 = torch.randn(nloc, d, device=dev) * 0.1
ex = PeerExchange(part, d, dev)
acc = torch.empty_like(x0); out = torch.empty_like(x0)
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(3):
    marks = [ev()]; marks[0].record(); names = []
    def mark(n):
        m = ev(); m.record(); marks.append(m); names.append(n)
    ex.barrier(0); mark("barrier0")
    ex.scatter(0, x0); mark("scatter")
    ex.barrier(0); mark("barrier1")
    cur = 0
    for l in range(L):
        last = l == L - 1
        addend = x0 if l == 0 else acc
        if last:
            csr.spmm(ex.bufs[cur], addend=addend, out=out, scale=float(L + 1), scale_mode=_lib.GR_SCALE_DIV, want_y=False); mark(f"spmm{l}")
        else:
            nxt = cur ^ 1
            csr.spmm(ex.bufs[cur], addend=addend, out=acc, want_y=False, peers=ex.peers(nxt)); mark(f"spmm{l}")
            ex.barrier(nxt); mark(f"barrier_l{l}")
            cur = nxt
    torch.cuda.synchronize()
    if it == 2:
        ts = [marks[k].elapsed_time(marks[k + 1]) for k in range(len(names))]
        line = f"rank {rank} rows {nloc} nnz {csr.nnz} n_long {csr.n_long} items {csr.n_long_items} split {csr.n_split} | " + " ".join(f"{n}={t:.1f}" for n, t in zip(names, ts)) + f" | total={sum(ts):.1f}"
        for r in range(world):
            if r == rank: print(line, flush=True)
            dist.barrier()
dist.destroy_process_group()
