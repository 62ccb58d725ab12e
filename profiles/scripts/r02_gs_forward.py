"""OrthogonalBundleGNN eval forward at the C2 shape, fused (gr_spmm_csr_map_f32) and layer-wise (GR_GS_FUSED=0),
for an ncu launch list.  usage: python profiles/scripts/r02_gs_forward.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import gnn_recommendations_b200 as g  # noqa: E402
from gnn_recommendations_b200.synthetic import synth_split  # noqa: E402

sp = synth_split("C2", 42)
nu, ni = sp["n_users"], sp["n_items"]
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = g.OrthogonalBundleGNN(nu, ni, 64, 3, 8, 0.1, 0.0, 0.01).to(dev).eval()
ds = g.InteractionDataset(sp["train"], sp["valid"], sp["test"], nu, ni, device=dev, name="C2")
csr = ds.get_torch_adjacency(normalized=True)
for mode in ("1", "0"):
    os.environ["GR_GS_FUSED"] = mode
    with torch.no_grad():
        for _ in range(3):
            model.get_all_embeddings(csr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(20):
            model.get_all_embeddings(csr)
    e1.record()
    torch.cuda.synchronize()
    print(f"GR_GS_FUSED={mode}: {e0.elapsed_time(e1) / 20:.4f} ms per forward (warm L2)", flush=True)
