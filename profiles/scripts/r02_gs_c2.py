"""BASELINE.json configs[1] alone: OrthogonalBundleGNN at the Gowalla shape — eval forward and one training step
(the gs_c2 entry of bench.py's extras).  usage: python profiles/scripts/r02_gs_c2.py"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import torch  # noqa: E402

import bench  # noqa: E402
import gnn_recommendations_b200 as g  # noqa: E402

bench.MODEL_CASES = {"gs_c2": bench.MODEL_CASES["gs_c2"]}
peak, _ = bench.peaks()
out = bench.run_model_extras(g, torch.device("cuda:0"), peak)
print(json.dumps(out))
