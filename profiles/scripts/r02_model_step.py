"""One model family's training steps at its BASELINE shape, eagerly (one launch per kernel), for ncu launch lists.
usage: python profiles/scripts/r02_model_step.py {lightgcn_c1|lightgcn_c4|gs_c2|ngcf_c3|gat_c3} [n_steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import gnn_recommendations_b200 as g  # noqa: E402
from gnn_recommendations_b200.synthetic import synth_split  # noqa: E402

case = sys.argv[1]
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
name, shape = case.split("_")
sp = synth_split(shape.upper(), 42)
nu, ni = sp["n_users"], sp["n_items"]
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = {"gs": lambda: g.OrthogonalBundleGNN(nu, ni, 64, 3, 8, 0.1, 0.0, 0.01),
         "ngcf": lambda: g.NGCF(nu, ni, 64, [64, 64, 64], 0.1, 0.1),
         "gat": lambda: g.GAT(nu, ni, 64, 3, 4, 0.1, 0.2, 0.1),
         "lightgcn": lambda: g.LightGCN(nu, ni, 64, 3, 0.1)}[name]()
ds = g.InteractionDataset(sp["train"], sp["valid"], sp["test"], nu, ni, device=dev, name=shape)
tr = g.Trainer(model, ds, {"batch_size": 512, "use_scheduler": False, "cuda_graph": False,
                           "checkpoint_dir": "/tmp/gr_prof_ckpt"}, device=dev)
torch.cuda.synchronize()
print("MARK setup done", flush=True)
loss = tr.train_steps(n_steps)
torch.cuda.synchronize()
print("loss", loss)
