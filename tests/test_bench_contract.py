"""CPU test of the benchmark contract's reference arm: `bench.py --impl reference` prints exactly ONE JSON line on stdout
with the keys the driver reads (metric, value, unit, n_gpus, steps, warmup, ms_per_step, e2e, cpu_baseline, config ...),
and the product arm fails loudly without a GPU instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

import pytest

import freesasa_b200 as fs
from oracle import bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ob.ref_available(), reason="oracle/_ref not built")
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         check=True, capture_output=True, text=True, cwd=ROOT, timeout=600).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("atoms/sec") and d["unit"] == "atoms/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["atoms_per_gpu"] == 100000
    # both arms must describe the workload with the same keys and values (the driver compares `config` key by key)
    import bench

    assert d["config"] == bench.shared_config(bench.ALGS["lr"])


def test_product_arm_fails_loudly_without_a_gpu():
    if fs.available():
        pytest.skip("a B200 is visible")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode != 0
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines())  # no number without a GPU
