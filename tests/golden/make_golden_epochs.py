"""Generates tests/golden/epochs.npz: one full training epoch of the UNMODIFIED reference Trainer
(src/training/trainer.py) for NGCF, GAT and OrthogonalBundleGNN on the tiny synthetic dataset, with
dropout 0 (the reference's dropout draws from the CPU generator and cannot be reproduced on a GPU).
Stores the initial and the post-epoch state_dict and the epoch loss.  Build container only.

    python tests/golden/make_golden_epochs.py
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import quiet, ref_dataset  # noqa: E402  (also puts the reference on sys.path)
from gnn_recommendations_b200.synthetic import SHAPES, synth_interactions  # noqa: E402
from src.models import GAT, NGCF, OrthogonalBundleGNN  # noqa: E402
from src.training.trainer import Trainer  # noqa: E402

nu, ni, e, _, _ = SHAPES["tiny"]
u, i, t = synth_interactions(nu, ni, e, 42)
ds, root = ref_dataset(u, i, t, nu, ni)
cfg = {"learning_rate": 1e-3, "weight_decay": 1e-4, "batch_size": 512, "epochs": 1, "eval_every": 1,
       "use_scheduler": False, "warmup_epochs": 0, "max_grad_norm": 1.0, "negative_samples": 1,
       "checkpoint_dir": os.path.join(root, "ckpt")}
out = {}
specs = {
    "ngcf": lambda: NGCF(nu, ni, embedding_dim=64, dropout=0.0, init_scale=0.1),
    "gat": lambda: GAT(nu, ni, embedding_dim=64, n_layers=2, n_heads=4, dropout=0.0, init_scale=0.1),
    "orthogonal_bundle": lambda: OrthogonalBundleGNN(nu, ni, embedding_dim=64, n_layers=3, block_size=8, dropout=0.0,
                                                      init_scale=0.1),
}
for name, make in specs.items():
    torch.manual_seed(42)
    m = make()
    for k, v in m.state_dict().items():
        out[f"{name}/init/{k}"] = v.detach().numpy().copy()
    tr = Trainer(m, ds, dict(cfg, model_name=None), device=torch.device("cpu"))
    torch.manual_seed(123)
    with quiet():
        loss = tr.train_epoch()
    out[f"{name}/loss"] = np.float64(loss)
    for k, v in m.state_dict().items():
        out[f"{name}/epoch/{k}"] = v.detach().numpy().copy()
    print(name, "epoch loss", loss, "params", sum(p.numel() for p in m.parameters()))
np.savez_compressed(os.path.join(HERE, "epochs.npz"), **out)
shutil.rmtree(root, ignore_errors=True)
print("epochs.npz written:", len(out), "arrays", os.path.getsize(os.path.join(HERE, "epochs.npz")), "bytes")
