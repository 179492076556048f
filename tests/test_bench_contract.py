"""bench.py contract (CPU part): the reference arm prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import numpy as np

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--P", "20000", "--res", "256"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"] and d["vs_baseline"] is None


def test_mask_fragile_zeroes_exactly_the_flagged_pixels():
    o = {"fragile": np.array([[0, 1], [0, 0]], np.uint8)}
    g = np.arange(12, dtype=np.float32).reshape(3, 2, 2) + 1
    m = orc.mask_fragile(o, g)
    assert (m[:, 0, 1] == 0).all() and (m[:, 0, 0] != 0).all() and (m[:, 1, :] == g[:, 1, :]).all()
    assert g[0, 0, 1] == 2  # the input is not modified
