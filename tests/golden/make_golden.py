"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/gnn-recommendations, imported via sys.path) on seeded synthetic
inputs.  Runs only in the build container (the reference does not travel to the GPU
box); the fixtures it writes are committed.

    python tests/golden/make_golden.py            # tiny + C1 fixtures
"""
import contextlib
import hashlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/gnn-recommendations"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from gnn_recommendations_b200.synthetic import SHAPES, synth_interactions  # noqa: E402
from src.data.dataset import RecommendationDataset  # noqa: E402
from src.evaluation.evaluator import Evaluator  # noqa: E402
from src.models import GAT, NGCF, LightGCN, OrthogonalBundleGNN  # noqa: E402
from src.training.losses import BPRLoss  # noqa: E402
from src.training.metrics import compute_metrics_from_topk  # noqa: E402
from src.training.trainer import Trainer  # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ref_dataset(u, i, t, n_users, n_items):
    root = tempfile.mkdtemp(prefix="gr_golden_")
    shutil.copytree(os.path.join(REF, "config"), os.path.join(root, "config"))
    ds = RecommendationDataset("ml-1m", root_dir=root)
    ds.processed_data = pd.DataFrame({"userId": u, "itemId": i, "timestamp": t})
    ds.n_users, ds.n_items, ds.stats = n_users, n_items, {}
    with quiet():
        ds.split()
        ds.build_graph()
    return ds, root


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def np_init(model, seed, scale):
    """Platform-independent weights (numpy PCG64 ziggurat) for every parameter."""
    rng = np.random.default_rng(seed)
    sd = model.state_dict()
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            sd[k] = torch.from_numpy((rng.standard_normal(tuple(v.shape)) * scale).astype(np.float32))
    model.load_state_dict(sd)


def sd_np(model, prefix):
    return {f"{prefix}/{k}": v.detach().cpu().numpy() for k, v in model.state_dict().items()}


def canonical_topk_np(scores: np.ndarray, k: int) -> np.ndarray:
    ids = np.arange(scores.shape[1])
    return np.stack([np.lexsort((ids, -row.astype(np.float64)))[:k] for row in scores])


def pairs(df):
    return df["userId"].to_numpy().astype(np.int64), df["itemId"].to_numpy().astype(np.int64)


def eval_loop_scores(user_emb, item_emb, eval_users, seen):
    """evaluator.py:96-104 verbatim semantics, returning the masked score matrix."""
    scores = user_emb[torch.tensor(eval_users)] @ item_emb.T
    for r, u in enumerate(eval_users):
        if u in seen and len(seen[u]):
            scores[r, list(seen[u])] = float("-inf")
    return scores


def make_tiny():
    nu, ni, e, d, L = SHAPES["tiny"]
    u, i, t = synth_interactions(nu, ni, e, 42)
    ds, root = ref_dataset(u, i, t, nu, ni)
    out = {"user": u, "item": i, "ts": t, "n_users": nu, "n_items": ni}
    for name, df in (("train", ds.train_data), ("valid", ds.valid_data), ("test", ds.test_data)):
        out[f"{name}_u"], out[f"{name}_i"] = pairs(df)
    adj = ds.norm_adj_matrix
    out["adj_row"], out["adj_col"], out["adj_val"] = adj.row, adj.col, adj.data
    deg = np.asarray(ds.adj_matrix.tocsr().sum(axis=1)).ravel()
    out["deg"] = deg
    md = int(deg.max())
    out["dis_lut"] = np.power(np.maximum(np.arange(md + 1, dtype=np.float32), np.float32(1.0)), -0.5)
    with quiet():
        from src.data.graph_builder import normalize_adjacency_matrix
        rown = normalize_adjacency_matrix(ds.adj_matrix, "row")
    out["adjrow_val"] = rown.data
    out["dinv_lut"] = np.power(np.maximum(np.arange(md + 1, dtype=np.float32), np.float32(1.0)), -1.0)
    A = ds.get_torch_adjacency(normalized=True)

    # ---- same-seed construction (RNG order) + forward outputs of the four models ----
    torch.manual_seed(42)
    lg = LightGCN(nu, ni, embedding_dim=64, n_layers=3, init_scale=0.1)
    out.update(sd_np(lg, "lightgcn"))
    with torch.no_grad():
        ue, ie = lg(A)
        out["lightgcn/out_user"], out["lightgcn/out_item"] = ue.numpy(), ie.numpy()
        out["lightgcn/layers"] = torch.stack(lg.get_layer_embeddings(A)).numpy()
    torch.manual_seed(7)
    lg5 = LightGCN(nu, ni, embedding_dim=128, n_layers=4, init_scale=0.1)
    out.update(sd_np(lg5, "lightgcn_d128_l4"))
    with torch.no_grad():
        ue5, ie5 = lg5(A)
        out["lightgcn_d128_l4/out_user"], out["lightgcn_d128_l4/out_item"] = ue5.numpy(), ie5.numpy()

    torch.manual_seed(42)
    ng = NGCF(nu, ni, embedding_dim=64, layer_sizes=[64, 64, 64], dropout=0.1, init_scale=0.1).eval()
    out.update(sd_np(ng, "ngcf"))
    with torch.no_grad():
        a, b = ng(A)
        out["ngcf/out_user"], out["ngcf/out_item"] = a.numpy(), b.numpy()

    torch.manual_seed(42)
    ga = GAT(nu, ni, embedding_dim=64, n_layers=3, n_heads=4, dropout=0.1, alpha=0.2, init_scale=0.1).eval()
    out.update(sd_np(ga, "gat"))
    with torch.no_grad():
        a, b = ga(A)
        out["gat/out_user"], out["gat/out_item"] = a.numpy(), b.numpy()

    torch.manual_seed(42)
    ob = OrthogonalBundleGNN(nu, ni, embedding_dim=64, n_layers=3, block_size=8, residual_alpha=0.1,
                             dropout=0.0, init_scale=0.01).eval()
    # make the fixture non-trivial: larger skew params, non-uniform layer weights, visible embeddings
    with torch.no_grad():
        g = torch.Generator().manual_seed(5)
        for p in ob.parameters():
            if p.dim() == 2 and p.shape[0] == 8:
                p.copy_(torch.randn(8, 8, generator=g) * 0.3)
        ob.layer_weights.copy_(torch.tensor([0.3, -0.2, 0.5, 0.1]))
        ob.user_embedding.weight.mul_(10.0)
        ob.item_embedding.weight.mul_(10.0)
    out.update(sd_np(ob, "gs"))
    with torch.no_grad():
        a, b = ob(A)
        out["gs/out_user"], out["gs/out_item"] = a.numpy(), b.numpy()
    # gradient fixture for the Group-and-Shuffle model (all parameters)
    ob.train()
    users = torch.arange(0, 64)
    pos = torch.arange(0, 64) % ni
    neg = (torch.arange(0, 64) * 7 + 3) % ni
    a, b = ob(A)
    loss = BPRLoss()((a[users] * b[pos]).sum(1), (a[users].unsqueeze(1) * b[neg.view(-1, 1)]).sum(2))
    ob.zero_grad()
    loss.backward()
    out["gs/loss"] = loss.detach().numpy()
    for k, p in ob.named_parameters():
        out[f"gs_grad/{k}"] = p.grad.numpy()

    # gradient fixtures for NGCF / GAT in eval() mode (dropout = identity)
    for name, mdl in (("ngcf", ng), ("gat", ga)):
        a, b = mdl(A)
        loss = BPRLoss()((a[users] * b[pos]).sum(1), (a[users].unsqueeze(1) * b[neg.view(-1, 1)]).sum(2))
        mdl.zero_grad()
        loss.backward()
        out[f"{name}/loss"] = loss.detach().numpy()
        for k, p in mdl.named_parameters():
            out[f"{name}_grad/{k}"] = p.grad.numpy()
    out["grad_users"], out["grad_pos"], out["grad_neg"] = users.numpy(), pos.numpy(), neg.numpy()

    # ---- sampler + BPR step through the reference Trainer ----
    cfg = {"learning_rate": 1e-3, "weight_decay": 1e-4, "batch_size": 512, "epochs": 1, "eval_every": 1,
           "use_scheduler": False, "warmup_epochs": 0, "max_grad_norm": 1.0, "negative_samples": 1,
           "validation_metrics": ["recall@10", "recall@20", "recall@50", "ndcg@10", "ndcg@20", "ndcg@50"],
           "checkpoint_dir": os.path.join(root, "ckpt")}
    torch.manual_seed(42)
    lg = LightGCN(nu, ni, embedding_dim=64, n_layers=3, init_scale=0.1)
    tr = Trainer(lg, ds, cfg, device=torch.device("cpu"))
    train_pairs = list(zip(ds.train_data["userId"].astype(int), ds.train_data["itemId"].astype(int)))
    torch.manual_seed(123)
    for b in range(3):
        us, ps, ns = tr._sample_batch(train_pairs, ni)
        out[f"batch{b}/users"], out[f"batch{b}/pos"], out[f"batch{b}/neg"] = us.numpy(), ps.numpy(), ns.numpy()
    # loss + gradients of step 0 on batch0 with the initial weights (autograd)
    us, ps, ns = (torch.from_numpy(out["batch0/users"]), torch.from_numpy(out["batch0/pos"]),
                  torch.from_numpy(out["batch0/neg"]))
    ue, ie = lg.get_all_embeddings(A)
    loss = BPRLoss()((ue[us] * ie[ps]).sum(1), (ue[us].unsqueeze(1) * ie[ns]).sum(2))
    lg.zero_grad()
    loss.backward()
    out["step0/loss"] = loss.detach().numpy()
    out["step0/grad_user"] = lg.user_embedding.weight.grad.numpy().copy()
    out["step0/grad_item"] = lg.item_embedding.weight.grad.numpy().copy()
    # gradient of the same loss w.r.t. the PROPAGATED embeddings (spec of the fused BPR kernel)
    ue_d, ie_d = ue.detach().requires_grad_(True), ie.detach().requires_grad_(True)
    l2 = BPRLoss()((ue_d[us] * ie_d[ps]).sum(1), (ue_d[us].unsqueeze(1) * ie_d[ns]).sum(2))
    g1, g2 = torch.autograd.grad(l2, [ue_d, ie_d])
    out["step0/gprop_user"], out["step0/gprop_item"] = g1.numpy(), g2.numpy()
    # one full reference epoch (11 steps), same seed for the sampler
    lg.zero_grad()
    torch.manual_seed(123)
    with quiet():
        ep_loss = tr.train_epoch()
    out["epoch/loss"] = np.float64(ep_loss)
    out["epoch/user_w"] = lg.user_embedding.weight.detach().numpy().copy()
    out["epoch/item_w"] = lg.item_embedding.weight.detach().numpy().copy()
    with quiet():
        vm = tr.validate()
    for k, v in vm.items():
        out[f"validate/{k}"] = np.float64(v)

    # ---- full-ranking eval through the reference Evaluator ----
    ev = Evaluator(k_values=[10, 20], device=torch.device("cpu"))
    with quiet():
        em = ev.evaluate(lg, ds, ds.test_data)
    for k, v in em.items():
        out[f"evaluate/{k}"] = np.float64(v)
    with torch.no_grad():
        ue, ie = lg.get_all_embeddings(A)
    gt = ev._prepare_ground_truth(ds.test_data)
    eval_users = sorted(gt.keys())
    with quiet():
        seen = ev._get_train_items_by_user(ds)
    scores = eval_loop_scores(ue, ie, eval_users, seen)
    out["eval/user_emb"], out["eval/item_emb"] = ue.numpy(), ie.numpy()
    out["eval/users"] = np.asarray(eval_users, dtype=np.int64)
    out["eval/scores"] = scores.numpy()
    out["eval/topk20_canonical"] = canonical_topk_np(scores.numpy(), 20)
    out["eval/topk20_torch"] = torch.topk(scores, k=20, dim=1).indices.numpy()
    # fewer than K unmasked items: a user who has seen all but 5 items -> -inf ties inside the list
    seen2 = dict(seen)
    seen2[eval_users[0]] = set(range(ni)) - {3, 17, 42, 99, 150}
    sc2 = eval_loop_scores(ue, ie, eval_users[:4], seen2)
    out["eval/short_topk20_canonical"] = canonical_topk_np(sc2.numpy(), 20)
    # exact score ties: duplicate item rows
    ie_t = ie.clone()
    ie_t[10] = ie_t[5]
    ie_t[150] = ie_t[5]
    sc3 = eval_loop_scores(ue, ie_t, eval_users, seen)
    out["eval/tie_item_emb"] = ie_t.numpy()
    out["eval/tie_topk20_canonical"] = canonical_topk_np(sc3.numpy(), 20)
    np.savez_compressed(os.path.join(HERE, "tiny.npz"), **out)
    shutil.rmtree(root, ignore_errors=True)
    print("tiny.npz written:", len(out), "arrays")


def make_c1():
    nu, ni, e, d, L = SHAPES["C1"]
    u, i, t = synth_interactions(nu, ni, e, 42)
    ds, root = ref_dataset(u, i, t, nu, ni)
    out = {"n_users": nu, "n_items": ni,
           "sha_user": sha(u), "sha_item": sha(i), "sha_ts": sha(t)}
    tu, ti = pairs(ds.train_data)
    out["sha_train_u"], out["sha_train_i"] = sha(tu), sha(ti)
    out["n_train"] = len(tu)
    adj = ds.norm_adj_matrix
    out["nnz"] = adj.nnz
    out["sha_adj_row"], out["sha_adj_col"], out["sha_adj_val"] = (
        sha(adj.row.astype(np.int32)), sha(adj.col.astype(np.int32)), sha(adj.data.astype(np.float32)))
    deg = np.asarray(ds.adj_matrix.tocsr().sum(axis=1)).ravel()
    md = int(deg.max())
    out["dis_lut"] = np.power(np.maximum(np.arange(md + 1, dtype=np.float32), np.float32(1.0)), -0.5)
    A = ds.get_torch_adjacency(normalized=True)
    lg = LightGCN(nu, ni, embedding_dim=64, n_layers=3, init_scale=0.1)
    np_init(lg, 42, 0.1)
    with torch.no_grad():
        ue, ie = lg(A)
    out["sha_out_user"], out["sha_out_item"] = sha(ue.numpy()), sha(ie.numpy())
    out["sum_out_user"], out["sum_out_item"] = ue.double().sum().numpy(), ie.double().sum().numpy()
    out["out_user_head"], out["out_item_head"] = ue[:64].numpy(), ie[:64].numpy()
    ev = Evaluator(k_values=[10, 20], device=torch.device("cpu"))
    gt = ev._prepare_ground_truth(ds.test_data)
    eval_users = sorted(gt.keys())
    # seen = train ∪ valid (evaluator.py:126-156), built vectorised (iterrows takes minutes here)
    vu, vi = pairs(ds.valid_data)
    seen = {}
    for a, b in zip(np.concatenate([tu, vu]).tolist(), np.concatenate([ti, vi]).tolist()):
        seen.setdefault(a, set()).add(b)
    tk = []
    for s0 in range(0, len(eval_users), 2048):
        sc = eval_loop_scores(ue, ie, eval_users[s0:s0 + 2048], seen)
        tk.append(canonical_topk_np(sc.numpy(), 20))
    tk = np.concatenate(tk)
    out["topk20_canonical"] = tk.astype(np.int16)
    m = compute_metrics_from_topk(torch.from_numpy(tk), eval_users, gt, ni, [10, 20])
    for k, v in m.items():
        out[f"metrics/{k}"] = np.float64(v)
    # sampler: first two batches of an epoch under torch.manual_seed(2024)
    cfg = {"batch_size": 512, "checkpoint_dir": os.path.join(root, "ckpt"), "use_scheduler": False}
    tr = Trainer(lg, ds, cfg, device=torch.device("cpu"))
    train_pairs = list(zip(tu.tolist(), ti.tolist()))
    torch.manual_seed(2024)
    for b in range(2):
        us, ps, ns = tr._sample_batch(train_pairs, ni)
        out[f"batch{b}/users"], out[f"batch{b}/pos"], out[f"batch{b}/neg"] = us.numpy(), ps.numpy(), ns.numpy()
    us, ps, ns = (torch.from_numpy(out["batch0/users"]), torch.from_numpy(out["batch0/pos"]),
                  torch.from_numpy(out["batch0/neg"]))
    ue2, ie2 = lg.get_all_embeddings(A)
    loss = BPRLoss()((ue2[us] * ie2[ps]).sum(1), (ue2[us].unsqueeze(1) * ie2[ns]).sum(2))
    lg.zero_grad()
    loss.backward()
    out["step0/loss"] = loss.detach().numpy()
    out["step0/grad_user_head"] = lg.user_embedding.weight.grad[:64].numpy().copy()
    out["step0/grad_item_head"] = lg.item_embedding.weight.grad[:64].numpy().copy()
    out["step0/grad_user_sum"] = lg.user_embedding.weight.grad.double().sum().numpy()
    out["step0/grad_item_abs_sum"] = lg.item_embedding.weight.grad.double().abs().sum().numpy()
    np.savez_compressed(os.path.join(HERE, "c1.npz"), **out)
    shutil.rmtree(root, ignore_errors=True)
    print("c1.npz written:", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["tiny", "c1"]
    if "tiny" in which:
        make_tiny()
    if "c1" in which:
        make_c1()
