"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (CPU, the oracle port) prints ONE JSON
line with the keys the driver reads, and both arms describe a workload with the same `config` object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ttt2",
                          "--steps", "3", "--warmup", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 3 and d["warmup"] == 3 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_share_the_config_object():
    sys.path.insert(0, ROOT)
    import bench
    for name, wl in bench.WORKLOADS.items():
        cfg = bench.config_of(name, wl["B"])
        assert set(cfg) == {"workload", "batch_per_gpu", "policy"} and cfg["workload"] == wl["desc"]
    # BASELINE.json's configurations are all present: configs[0..3] as workloads, configs[4] = the same under torchrun
    assert bench.ORDER == ["tron", "blokus", "ttt4", "ttt2"]
