"""bench.py contract pieces that run without a GPU: the reference arm's JSON line (`--impl reference`: the oracle port on the host
cores, rank 0 only) and the product arm's refusal to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env_extra=None, timeout=280):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "3", "--batch", "256"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "adapted_projector_fwd_bwd_samples_per_sec" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["rows_per_gpu"] == 256 and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "256 rows" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.timeout(120)
def test_reference_arm_other_ranks_exit_silently():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3", "--batch", "256"],
               {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=100)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout, out.stderr[-500:])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
@pytest.mark.timeout(120)
def test_product_arm_refuses_to_run_without_cuda():
    out = _run(["--gpus", "1", "--steps", "1", "--warmup", "3"], timeout=100)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
