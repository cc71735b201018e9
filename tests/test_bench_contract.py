"""CPU: bench.py's reference arm prints one JSON line with the contract's keys (tiny workload via
UPOL_BENCH_NMOL); the CUDA arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(args, env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          env=env, timeout=300, cwd=ROOT)


def test_reference_arm_json_contract(tmp_path):
    out = _run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"],
               {"UPOL_BENCH_NMOL": "150", "TMPDIR": str(tmp_path)})
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "uind_energy_grad_site_pairs_per_s"
    assert d["unit"] == "site-pairs/s" and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("imidazole_liquid_150mol")


def test_reference_arm_other_ranks_exit_quietly(tmp_path):
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
               {"UPOL_BENCH_NMOL": "150", "RANK": "1", "WORLD_SIZE": "2", "TMPDIR": str(tmp_path)})
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cuda_arm_has_no_cpu_fallback(tmp_path):
    out = _run(["--gpus", "1", "--steps", "1", "--warmup", "1"], {"UPOL_BENCH_NMOL": "150", "TMPDIR": str(tmp_path)})
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
