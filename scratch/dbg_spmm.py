import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import gnn_recommendations_b200 as g
from gnn_recommendations_b200.synthetic import synth_split
from oracle import pyoracle as po, coracle
sp = synth_split("tiny", 42)
ref = po.build_norm_adj(*sp["train"], sp["n_users"], sp["n_items"])
d = int(sys.argv[1]); thr = int(sys.argv[2])
csr = g.NormAdjCSR(torch.from_numpy(ref["indptr"].astype(np.int32)).cuda(), torch.from_numpy(ref["indices"]).cuda(), torch.from_numpy(ref["vals"]).cuda(), ref["n"], ref["n"], long_threshold=thr)
print("n_long", csr.n_long, "groups", csr.n_groups, flush=True)
x = torch.randn(ref["n"], d, generator=torch.Generator().manual_seed(d))
want = coracle.spmm_fmaf(ref["indptr"], ref["indices"], ref["vals"], x.numpy())
for it in range(3):
    y, _ = csr.spmm(x.cuda()); torch.cuda.synchronize()
    print(it, "equal", np.array_equal(y.cpu().numpy().view(np.uint32), want.view(np.uint32)), flush=True)
