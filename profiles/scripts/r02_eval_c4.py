"""Full-ranking top-20 at the Amazon-Book shape (52 643 users x 91 599 items, d = 64) through the tensor-core
nomination path; with `breakdown` also with parts of the selection switched off (GR_TC_DEBUG bits), for ncu.
usage: python profiles/scripts/r02_eval_c4.py [reps] [breakdown]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gnn_recommendations_b200 as g  # noqa: E402
from gnn_recommendations_b200.evaluator import full_rank_topk, seen_csr  # noqa: E402
from gnn_recommendations_b200.synthetic import synth_pairs_device  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda:0")
nu, ni, e, d, L = 52643, 91599, 2984108, 64, 3
u, i = synth_pairs_device(nu, ni, e, 42, dev)
csr = g.NormAdjCSR.from_pairs(u, i, nu, ni, device=dev)
with torch.device(dev):
    model = g.LightGCN(nu, ni, embedding_dim=d, n_layers=L, init_scale=0.1)
with torch.no_grad():
    ue, ie = model.get_all_embeddings(csr)
eval_users = np.arange(nu)
ip, it = seen_csr(eval_users, nu, (u.cpu().numpy(), i.cpu().numpy()))
ip_d, it_d = torch.from_numpy(ip).to(dev), torch.from_numpy(it).to(dev)
eu_d = torch.from_numpy(eval_users).to(dev)


def timed(label, env):
    for k, v in env.items():
        os.environ[k] = v
    stats = {}
    for _ in range(2):
        out = full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=True, stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=True, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    for k in env:
        del os.environ[k]
    print(f"{label}: {e0.elapsed_time(e1) / reps:.3f} ms  rows re-ranked exactly: {stats['rows_reranked_exactly']}", flush=True)
    return out


exact = full_rank_topk(ue, ie, eu_d, ip_d, it_d, 20, tensor_cores=False)
a = timed("tensor-core nomination + exact re-scoring", {})
print("lists identical to the exact kernel:", bool(torch.equal(a, exact)), flush=True)
if len(sys.argv) > 2:       # kernel-time breakdown for an ncu launch list (lists are wrong in these modes: every row is re-ranked)
    timed("filter + appends, queues dropped (GR_TC_DEBUG=8)", {"GR_TC_DEBUG": "8"})
    timed("filter only (GR_TC_DEBUG=16)", {"GR_TC_DEBUG": "16"})
    timed("selection off (GR_TC_DEBUG=1)", {"GR_TC_DEBUG": "1"})
