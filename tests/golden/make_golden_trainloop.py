"""Generates tests/golden/trainloop.npz: the UNMODIFIED reference `Trainer.train()` (trainer.py:469-579)
for 4 epochs of LightGCN on the tiny dataset with a 2-epoch linear warm-up, the cosine schedule, validation
every epoch and early stopping — losses, learning rates, validation metrics, best epoch, final weights.
Build container only.

    python tests/golden/make_golden_trainloop.py
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import quiet, ref_dataset  # noqa: E402
from gnn_recommendations_b200.synthetic import SHAPES, synth_interactions  # noqa: E402
from src.models import LightGCN  # noqa: E402
from src.training.trainer import Trainer  # noqa: E402

nu, ni, e, _, _ = SHAPES["tiny"]
u, i, t = synth_interactions(nu, ni, e, 42)
ds, root = ref_dataset(u, i, t, nu, ni)
cfg = {"learning_rate": 5e-3, "weight_decay": 1e-4, "batch_size": 512, "epochs": 4, "eval_every": 1,
       "use_scheduler": True, "warmup_epochs": 2, "max_grad_norm": 1.0, "negative_samples": 1,
       "validation_metrics": ["recall@10", "ndcg@10", "recall@20"], "early_stopping_metric": "recall@10",
       "early_stopping": {"patience": 3, "min_delta": 0.0001}, "model_name": "lightgcn_golden",
       "checkpoint_dir": os.path.join(root, "ckpt")}
torch.manual_seed(42)
m = LightGCN(nu, ni, embedding_dim=64, n_layers=3, init_scale=0.1)
tr = Trainer(m, ds, cfg, device=torch.device("cpu"))
lrs = []
orig = tr.train_epoch


def spy():
    lrs.append(tr.optimizer.param_groups[0]["lr"])
    return orig()


tr.train_epoch = spy
torch.manual_seed(123)
with quiet():
    res = tr.train()
out = {"train_losses": np.asarray(res["train_losses"], dtype=np.float64), "lrs": np.asarray(lrs, dtype=np.float64),
       "best_metric": np.float64(res["best_metric"]), "best_epoch": np.int64(res["best_epoch"]),
       "user_w": m.user_embedding.weight.detach().numpy().copy(), "item_w": m.item_embedding.weight.detach().numpy().copy()}
for ep, vm in enumerate(res["valid_metrics"]):
    for k, v in vm.items():
        out[f"valid/{ep}/{k}"] = np.float64(v)
np.savez_compressed(os.path.join(HERE, "trainloop.npz"), **out)
shutil.rmtree(root, ignore_errors=True)
print("losses", out["train_losses"], "lrs", out["lrs"], "best", res["best_metric"], res["best_epoch"],
      {k: v for k, v in res["valid_metrics"][-1].items() if "recall" in k})
