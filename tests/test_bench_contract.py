"""bench.py's reference arm runs on the host alone: check the JSON line it prints against the driver's contract
(the product arm needs a GPU and is exercised on the GPU box; its line shares the same builder for these keys)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Mpixels/s disparity->PointCloud2") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["steps"] == 1 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must fail loudly, not print a number."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_algorithmic_bytes_are_the_survey_figures():
    """20 N bytes per float frame (15.36 MB at 1280x720, 156.416 MB at 3840x2160), (W-70)(H-70) + 16 N
    for the mono8 callback; both arms describe a config with the same workload string."""
    import bench
    assert bench.algorithmic_bytes(bench.CONFIGS[3]) == 15_360_000
    assert bench.algorithmic_bytes(bench.CONFIGS[4]) == 156_416_000
    assert bench.algorithmic_bytes(bench.CONFIGS[2]) == (752 - 70) * (480 - 70) + 16 * 268_800
    assert bench.fusion_dims(1280, 720) == (705, 665, 665)
    assert bench.unit_pixels(bench.CONFIGS[5]) == 4 * 1280 * 720
    assert len({c["workload"] for c in bench.CONFIGS.values()}) == 4


def test_reference_arm_other_configs_share_the_workload_string():
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "3", "--steps", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["config"]["workload"] == bench.CONFIGS[3]["workload"] and d["cpu_baseline"]["kind"] == "port"
