"""Trainer — drop-in for src/training/trainer.py (Trainer :28-615).

Same constructor, config keys, defaults and public methods (train, train_epoch, validate,
save_checkpoint, load_checkpoint); the hot loop bodies run on the B200 kernels:

  _sample_batch   C sampler on torch's own mt19937 stream (bit-identical indices)   trainer.py:146-197
  train_epoch     per step: full-graph propagation (SpMM kernels) -> fused BPR kernel (gathers,
                  B x B loss, gradient scatter) -> backward SpMMs -> clip_grad_norm_ -> Adam
                  (the last two stay stock torch, as in the reference)                  trainer.py:199-281
  validate        fused score + mask(train) + top-K kernel, K = max of validation_metrics   trainer.py:283-347

The loss is accumulated on the device in float64 (the same double additions the reference does
with ``loss.item()``) and read back once per epoch instead of once per step.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.optim as optim

from .evaluator import _pairs, full_rank_topk, ground_truth_dict, seen_csr
from .layer_ops import device_dropout_seeds, new_dropout_seed
from .optim import fused_clip_adam_step, fused_clip_adam_supported
from .losses import BPRLoss, bpr_fused
from .metrics import compute_metrics_from_topk, topk_metrics_device
from .sampler import BprSampler


class _GraphStep:
    """One training step (trainer.py:249-276 after the sampling) captured in a CUDA graph."""

    RING = 16

    def __init__(self, tr: "Trainer", adj, total: torch.Tensor):
        from .optim import adam_step_scalars

        self.tr, self.total = tr, total
        self._scalars = adam_step_scalars
        self.key = (id(adj), id(tr.optimizer), id(tr.model), tr.batch_size, tr.max_grad_norm)
        sampler = tr._get_sampler()
        self.b = min(tr.batch_size, len(sampler))
        dev = tr.device
        # one staged int64 vector per step: users | positives | negatives | dropout seed of the step
        self.flat = torch.zeros(3 * self.b + 1, dtype=torch.int64, device=dev)
        self.idx = self.flat[:3 * self.b].view(3, self.b)
        self.seed = self.flat[3 * self.b:]
        self.sc = torch.ones(2, dtype=torch.float32, device=dev)
        self.stage = [(torch.zeros(3 * self.b + 1, dtype=torch.int64, pin_memory=True),
                       torch.zeros(2, dtype=torch.float32, pin_memory=True), torch.cuda.Event()) for _ in range(self.RING)]
        self.views = [tuple(st[0][r * self.b:(r + 1) * self.b].numpy() for r in range(3)) for st in self.stage]
        self.used = [False] * self.RING
        self.k = 0
        model, opt = tr.model, tr.optimizer
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=torch.cuda.current_stream(dev)), \
                device_dropout_seeds(self.seed) as seeds:
            x = tr._propagated(adj)
            loss = bpr_fused(x, model.n_users, self.idx[0], self.idx[1], self.idx[2].view(-1, 1))
            opt.zero_grad()
            loss.backward()
            fused_clip_adam_step(opt, tr.max_grad_norm, step_scalars_dev=self.sc)
            total += loss.detach().double()
            # did the step ask for dropout seeds?  Only then is the CPU generator consumed per step (a model
            # without active dropout must leave the sampler's mt19937 stream exactly as the reference does)
            self.uses_seed = seeds.drawn() > 0

    def valid_for(self, adj) -> bool:
        tr = self.tr
        return self.key == (id(adj), id(tr.optimizer), id(tr.model), tr.batch_size, tr.max_grad_norm) and \
            tr._graph_ok(adj, need_grads=False)

    def run(self) -> bool:
        tr = self.tr
        k = self.k
        self.k = (k + 1) % self.RING
        idx_h, sc_h, ev = self.stage[k]
        if self.used[k]:
            ev.synchronize()                         # the copy that last read this staging slot has run
        tr._get_sampler().sample(tr.batch_size, out=self.views[k])
        if self.uses_seed:
            idx_h[3 * self.b] = new_dropout_seed().value       # one draw of the CPU generator per step
        step_size, bc2 = self._scalars(tr.optimizer)          # advances the optimizer's step counters
        sc_h[0], sc_h[1] = step_size, bc2
        self.flat.copy_(idx_h, non_blocking=True)
        self.sc.copy_(sc_h, non_blocking=True)
        ev.record(torch.cuda.current_stream(tr.device))
        self.used[k] = True
        self.graph.replay()
        return True

    def finish(self):
        torch.cuda.current_stream(self.tr.device).synchronize()


class Trainer:
    def __init__(self, model, dataset, config: Dict, device: Optional[torch.device] = None):
        self.model, self.dataset, self.config = model, dataset, config
        self.model_name = config.get("model_name")
        self.dataset_name = config.get("dataset_name")
        self.device = torch.device("cuda") if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("gnn-recommendations_b200.Trainer needs a CUDA device (no CPU fallback)")
        self.model.to(self.device)
        self._maybe_warm_start_embeddings()

        lr = float(config.get("learning_rate", 0.001))
        wd = float(config.get("weight_decay", 1e-4))
        self.batch_size = int(config.get("batch_size", 2048))
        self.epochs = int(config.get("epochs", 300))
        self.eval_every = int(config.get("eval_every", 10))
        self.negative_samples = int(config.get("negative_samples", 1))
        self.optimizer = optim.Adam(self.model.parameters(), lr=lr, weight_decay=wd)
        self.use_scheduler = config.get("use_scheduler", True)
        self.warmup_epochs = int(config.get("warmup_epochs", 5))
        self.base_lr = lr
        self.scheduler = (optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=self.epochs - self.warmup_epochs,
                                                               eta_min=lr * 0.01) if self.use_scheduler else None)
        self.loss_fn = BPRLoss()
        self.max_grad_norm = float(config.get("max_grad_norm", 1.0))
        self.fused_optimizer = bool(config.get("fused_optimizer", True))   # gr_clip_adam_fused (same update)
        self.early_stopping = config.get("early_stopping", {})
        self.patience = int(self.early_stopping.get("patience", 20))
        self.min_delta = float(self.early_stopping.get("min_delta", 0.0001))
        self.validation_metrics = config.get("validation_metrics", ["recall@10", "ndcg@10"])
        self.early_stopping_metric = config.get("early_stopping_metric", "recall@10")
        self.current_epoch, self.best_metric, self.best_epoch, self.patience_counter = 0, 0.0, 0, 0
        self.train_losses: List[float] = []
        self.valid_metrics: List[Dict[str, float]] = []
        self.checkpoint_dir = Path(config.get("checkpoint_dir", "results/checkpoints"))
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self._sampler: Optional[BprSampler] = None
        self.step_times_ms: List[float] = []

    # ------------------------------------------------------------------ data access
    def _train_arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        train_data = self.dataset.train_data
        if train_data is None:
            train_file = self.dataset.processed_data_path / "train.txt"
            if not train_file.exists():
                raise ValueError("Train данные не найдены!")
            import pandas as pd

            train_data = pd.read_csv(train_file, sep="\t", header=None, names=["userId", "itemId"])
        return _pairs(train_data)

    def _get_sampler(self) -> BprSampler:
        if self._sampler is None:       # trainer.py:169-172 builds its positive sets once, too
            u, i = self._train_arrays()
            self._sampler = BprSampler(u, i, self.model.n_users, self.dataset.n_items)
        return self._sampler

    def _sample_batch(self, train_data=None, n_items: Optional[int] = None):
        """(users, pos_items, neg_items[B, 1]) on the device, bit-identical to the reference."""
        if self.negative_samples != 1:
            raise ValueError("negative_samples != 1 is shape-invalid in the reference's BPR loss "
                             "([B] - [B, n_neg] broadcast, losses.py:44)")
        u, p, n = self._get_sampler().sample(self.batch_size)
        dev = self.device
        return (torch.from_numpy(u).to(dev, non_blocking=True), torch.from_numpy(p).to(dev, non_blocking=True),
                torch.from_numpy(n).to(dev, non_blocking=True).view(-1, 1))

    # ------------------------------------------------------------------ training
    def _propagated(self, adj) -> torch.Tensor:
        if hasattr(self.model, "propagate"):
            return self.model.propagate(adj)
        user_emb, item_emb = self.model.get_all_embeddings(adj)
        return torch.cat([user_emb, item_emb], dim=0)

    def train_epoch(self) -> float:
        sampler = self._get_sampler()
        return self.train_steps(len(sampler) // self.batch_size + 1)          # trainer.py:237

    def train_steps(self, n_steps: int) -> float:
        """``n_steps`` bodies of the reference's epoch loop (trainer.py:237-279); ``train_epoch`` runs
        ``len(train) // batch_size + 1`` of them.  Returns the mean loss.

        The body is launch-bound at the dataset shapes (a 0.35 ms propagation under ~14 launches), so after
        three eager steps it is captured ONCE into a CUDA graph — propagation, fused BPR, backward kernels,
        fused clip+Adam — and replayed; per step the host only draws the batch (bit-exact sampler), stages the
        3B indices and Adam's two step-dependent scalars in pinned memory and launches the graph."""
        # All steps run on one dedicated non-default stream: autograd ties each parameter's AccumulateGrad node to
        # the stream it was created on, and a node created on the legacy default stream cannot be used inside
        # a stream capture ("would make the legacy stream depend on a capturing stream").
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_train_stream", None) is None:
            self._train_stream = torch.cuda.Stream(self.device)
        self._train_stream.wait_stream(cur)
        with torch.cuda.stream(self._train_stream):
            out = self._train_steps(n_steps)
        cur.wait_stream(self._train_stream)
        return out

    def _train_steps(self, n_steps: int) -> float:
        self.model.train()
        adj = self.dataset.get_torch_adjacency(normalized=True).to(self.device)
        if getattr(self, "_loss_acc", None) is None:
            self._loss_acc = torch.zeros((), dtype=torch.float64, device=self.device)
        total = self._loss_acc.zero_()
        n_batches = 0
        n_steps = int(n_steps)
        graph = getattr(self, "_graph", None)
        if graph is not None and not graph.valid_for(adj):
            graph = self._graph = None
        for it in range(n_steps):
            if graph is None and it >= 3 and n_steps - it >= 8 and self._graph_ok(adj):
                x = loss = users = pos = neg = None          # drop the last eager step's autograd graph
                graph = self._graph = self._capture_step(adj, total)
            if graph is not None:
                if graph.run():
                    n_batches += 1
                continue
            users, pos, neg = self._sample_batch()
            if users.numel() == 0:
                continue
            x = self._propagated(adj)
            loss = bpr_fused(x, self.model.n_users, users, pos, neg)
            if hasattr(self.model, "get_regularization_loss"):
                loss = loss + self.model.get_regularization_loss()
            self.optimizer.zero_grad()
            loss.backward()
            if self.fused_optimizer and fused_clip_adam_supported(self.optimizer):
                fused_clip_adam_step(self.optimizer, self.max_grad_norm)      # clip + Adam in two passes
            else:
                if self.max_grad_norm > 0:
                    torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
                self.optimizer.step()
            total += loss.detach().double()
            n_batches += 1
        if graph is not None:
            graph.finish()
        return float(total.item()) / n_batches if n_batches > 0 else 0.0

    # ------------------------------------------------------------------ CUDA-graph step
    def _graph_ok(self, adj, need_grads: bool = True) -> bool:
        """Capture needs: the fused optimizer, a device-resident CSR, one negative per sample and a model whose
        train-mode randomness (fused dropout) takes its seed from device memory (``_graph_safe``)."""
        from .graph_builder import NormAdjCSR

        if not self.config.get("cuda_graph", True) or not self.fused_optimizer:
            return False
        if not isinstance(adj, NormAdjCSR) or self.negative_samples != 1 or not hasattr(self.model, "propagate"):
            return False
        if not fused_clip_adam_supported(self.optimizer):
            return False
        if need_grads and any(p.grad is None for p in self.model.parameters()):
            return False
        if hasattr(self.model, "get_regularization_loss") or not getattr(self.model, "_graph_safe", False):
            return False
        return len(self._get_sampler()) > 0

    def _capture_step(self, adj, total):
        return _GraphStep(self, adj, total)

    # ------------------------------------------------------------------ validation
    @staticmethod
    def _parse_k_values(metrics_list: List[str]) -> List[int]:
        ks = set()
        for m in metrics_list:
            if "@" in m:
                try:
                    ks.add(int(m.split("@")[-1]))
                except ValueError:
                    continue
        return sorted(ks)

    def validate(self) -> Dict[str, float]:
        self.model.eval()
        with torch.no_grad():
            adj = self.dataset.get_torch_adjacency(normalized=True).to(self.device)
            user_emb, item_emb = self.model.get_all_embeddings(adj)
            valid_data = self.dataset.valid_data
            if valid_data is None:
                valid_file = self.dataset.processed_data_path / "valid.txt"
                if not valid_file.exists():
                    return {}
                import pandas as pd

                valid_data = pd.read_csv(valid_file, sep="\t", header=None, names=["userId", "itemId"])
            valid_pairs = _pairs(valid_data)
            eval_users = np.unique(valid_pairs[0])          # == sorted(ground_truth.keys()) (trainer.py:318)
            if len(eval_users) == 0:
                return {}
            k_values = self._parse_k_values(self.validation_metrics)
            max_k = max(k_values) if k_values else 10
            ip, it = seen_csr(eval_users, user_emb.shape[0], self._train_arrays())   # train only (:324)
            topk = full_rank_topk(user_emb, item_emb, eval_users, ip, it, max_k)
            gp, gi = seen_csr(eval_users, user_emb.shape[0], valid_pairs)           # ground truth rows
            return topk_metrics_device(topk, gp, gi, self.dataset.n_items, k_values if k_values else [10])

    # ------------------------------------------------------------------ warm start / export
    def _maybe_warm_start_embeddings(self):
        ws = self.config.get("warm_start", {})
        if not ws or not ws.get("enabled", False):
            return
        apply_to = ws.get("apply_to")
        if apply_to and self.model_name and apply_to != self.model_name:
            return
        path = ws.get("embeddings_path")
        if not path:
            raise ValueError("warm_start.enabled=true, but embeddings_path is not set")
        ckpt = torch.load(path, map_location=self.device)
        ue = ie = None
        if isinstance(ckpt, dict):
            if "user_embedding" in ckpt and "item_embedding" in ckpt:
                ue, ie = ckpt["user_embedding"], ckpt["item_embedding"]
            elif "model_state_dict" in ckpt:
                sd = ckpt["model_state_dict"]
                if "user_embedding.weight" in sd and "item_embedding.weight" in sd:
                    ue, ie = sd["user_embedding.weight"], sd["item_embedding.weight"]
        if ue is None or ie is None:
            raise ValueError("Warm-start embeddings not found in file. "
                             "Expected keys: user_embedding/item_embedding or model_state_dict.")
        if not hasattr(self.model, "user_embedding") or not hasattr(self.model, "item_embedding"):
            raise ValueError("Model does not expose user_embedding/item_embedding for warm-start.")
        for name, got, want in (("user_embedding", ue, self.model.user_embedding.weight),
                                ("item_embedding", ie, self.model.item_embedding.weight)):
            if got.shape != want.shape:
                raise ValueError(f"{name} shape mismatch: got {tuple(got.shape)}, expected {tuple(want.shape)}")
        with torch.no_grad():
            self.model.user_embedding.weight.copy_(ue.to(self.device))
            self.model.item_embedding.weight.copy_(ie.to(self.device))

    def _maybe_export_embeddings(self):
        cfg = self.config.get("export_embeddings", {})
        if not cfg or not cfg.get("enabled", False):
            return
        apply_to = cfg.get("apply_to")
        if apply_to and self.model_name and apply_to != self.model_name:
            return
        path = cfg.get("path")
        if not path:
            raise ValueError("export_embeddings.enabled=true, but path is not set")
        adj = self.dataset.get_torch_adjacency(normalized=True).to(self.device)
        with torch.no_grad():
            ue, ie = self.model.get_all_embeddings(adj)
        payload = {"user_embedding": ue.detach().cpu(), "item_embedding": ie.detach().cpu(),
                   "source_model": self.model_name, "dataset": self.dataset_name,
                   "embedding_dim": int(ue.size(1)), "n_users": int(ue.size(0)), "n_items": int(ie.size(0))}
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        torch.save(payload, path)

    def _get_orthogonality_metrics(self) -> Dict[str, float]:
        if not hasattr(self.model, "get_orthogonality_metrics"):
            return {}
        with torch.no_grad():
            m = self.model.get_orthogonality_metrics()
        return {k: float(v.detach().cpu()) for k, v in m.items()}

    # ------------------------------------------------------------------ outer loop
    def train(self) -> Dict:
        start = time.time()
        for epoch in range(1, self.epochs + 1):
            self.current_epoch = epoch
            if epoch <= self.warmup_epochs:                      # linear warm-up (trainer.py:492-496)
                for g in self.optimizer.param_groups:
                    g["lr"] = self.base_lr * (epoch / self.warmup_epochs)
            elif self.scheduler is not None:
                self.scheduler.step()
            train_loss = self.train_epoch()
            self.train_losses.append(train_loss)
            if epoch % self.eval_every == 0 or epoch == 1:
                vm = self.validate()
                self.valid_metrics.append(vm)
                cur = vm.get(self.early_stopping_metric, 0.0)
                print(f"Epoch {epoch:3d}/{self.epochs} | LR: {self.optimizer.param_groups[0]['lr']:.6f} | "
                      f"Train Loss: {train_loss:.4f} | {self.early_stopping_metric}: {cur:.4f} | "
                      f"NDCG@10: {vm.get('ndcg@10', 0.0):.4f}")
                orth = self._get_orthogonality_metrics()
                if orth:
                    print("Orthogonality: " + " | ".join(f"{k}={v:.2e}" for k, v in orth.items()))
                if cur > self.best_metric + self.min_delta:
                    self.best_metric, self.best_epoch, self.patience_counter = cur, epoch, 0
                    self.save_checkpoint(epoch, vm)
                else:
                    self.patience_counter += 1
                if self.patience_counter >= self.patience:
                    print(f"Early stopping at epoch {epoch}; best {self.best_metric:.4f} @ {self.best_epoch}")
                    break
        training_time = time.time() - start
        self._maybe_export_embeddings()
        return {"best_metric": self.best_metric, "best_epoch": self.best_epoch, "training_time": training_time,
                "train_losses": self.train_losses, "valid_metrics": self.valid_metrics}

    def save_checkpoint(self, epoch: int, metrics: Dict[str, float]):
        torch.save({"epoch": epoch, "model_state_dict": self.model.state_dict(),
                    "optimizer_state_dict": self.optimizer.state_dict(), "metrics": metrics,
                    "best_metric": self.best_metric}, self.checkpoint_dir / f"checkpoint_epoch_{epoch}.pt")

    def load_checkpoint(self, checkpoint_path: Path):
        ckpt = torch.load(checkpoint_path, map_location=self.device)
        self.model.load_state_dict(ckpt["model_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        self._graph = None          # the captured step points at the old optimizer-state tensors
        self.current_epoch = ckpt["epoch"]
        self.best_metric = ckpt.get("best_metric", 0.0)
