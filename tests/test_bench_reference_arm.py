"""`bench.py --impl reference` (the CPU arm the driver times beside the product arm): runs without a GPU on a small source and prints
ONE JSON line with the keys of the bench contract; under a multi-rank launch only rank 0 works and prints."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--rays", "40000", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ray-surface interactions/s" and d["unit"] == "interactions/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C2 ") and d["config"]["rays_total"] == 40000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
