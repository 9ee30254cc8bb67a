"""GPU: the driver's own entry points (__graft_entry__.smoke, a short bench.py run) exit cleanly.  The driver runs
both on a fresh B200 at round end; a stale assertion in either must show up in the GPU suite first."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_smoke_entry_point():
    import __graft_entry__ as g
    g.smoke()


def test_bench_default_line_is_complete():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "ms_per_step", "roofline", "cpu_baseline", "e2e", "gpu_launches",
                "clocks", "config"):
        assert key in line, key
    assert line["gpu_launches"] > 0 and line["value"] > 0
    assert 0.5 < line["roofline"]["frac"] < 1.2
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
