import sys, os, torch
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    g.full_rank_topk(ue, ie, eu, None, None, 20, tensor_cores=True)
e1.record(); torch.cuda.synchronize()
print("debug", os.environ.get("GR_TC_DEBUG", "0"), "ms", e0.elapsed_time(e1) / 3)
