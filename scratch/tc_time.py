import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gnn_recommendations_b200 as g
nu, ni, d = 52643, 91599, 64
gen = torch.Generator(device="cuda").manual_seed(0)
ue = torch.randn(nu, d, device="cuda", generator=gen) * 0.1
ie = torch.randn(ni, d, device="cuda", generator=gen) * 0.1
eu = torch.arange(nu, device="cuda")
for _ in range(2):
    g.full_rank_topk(ue, ie, eu, None, None, 20, tensor_cores=True)
torch.cuda.synchronize()
print("done")
