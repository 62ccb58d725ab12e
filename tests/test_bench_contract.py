"""bench.py contract checks that need no GPU: the `--impl reference` arm (the reference's CPU path, oracle
port) prints ONE JSON line with the contract's keys; non-zero ranks print nothing; without a GPU the
product arm fails loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600)


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_contract_line():
    p = _run(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "lightgcn_propagation_edges_per_s" and d["unit"] == "edges/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.timeout(600)
def test_reference_arm_other_ranks_exit_silently():
    p = _run(["--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0", "--gpus", "2"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
@pytest.mark.timeout(600)
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run(["--workload", "C1", "--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu", "--no-extras"])
    assert p.returncode != 0 and p.stdout.strip() == ""
