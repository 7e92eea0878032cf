"""GPU: bench.py's own arm honours the driver's JSON contract (one line, the keys the judge reads, sane values) on a
small batch."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_arm_prints_one_contract_line():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "3", "--warmup", "3",
                          "--pairs", "4"], capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "GN-iters/s" and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 3 and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 4 * 2 * 3 * 480 * 640 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # host copies inside the timed region
    assert d["gpu_launches"] == 2 * 3                   # fused kernel + finalize/solve per step
    assert d["clocks"]["sm_mhz"] is None or d["clocks"]["sm_mhz"] > 0
    assert d["blob_segments"]["gn"]["roofline_frac"] > 0
