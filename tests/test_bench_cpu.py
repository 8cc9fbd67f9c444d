"""CPU: the bench's reference arm (`bench.py --impl reference`, the one place besides tests / smoke that may run the
oracle) prints the contract's JSON line, names ITS grid and says it is not the GPU arm's configuration; the config table
covers BASELINE.json's configs."""
import json
import os
import subprocess
import sys

from tests import helpers as H

ROOT = H.ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, JXF_REF_BUDGET_S="5", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MCUPS" and line["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["value"] > 0 and line["steps"] == 1 and line["dtype"] == "f64"
    cfg = line["config"]
    # the arm names the grid it really ran and refuses the claim of running the GPU arm's configuration
    assert cfg["same_config"] is False and "sample_of" in cfg and "512^3" in cfg["sample_of"]
    assert "512^3" not in cfg["workload"].split("(")[0]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]


def test_bench_configs_cover_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert len(base["configs"]) == 5
    assert set(bench.CONFIGS) >= {"sod1000", "riemann1024", "tgv256", "tgv512", "hit1024"}
