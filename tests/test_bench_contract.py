"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line (config identical to the b200
arm's `config_of`, physical cores used explicitly even when OMP_NUM_THREADS=1 as under torchrun, e2e/cpu_baseline blocks)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line_and_thread_count():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")          # what torchrun exports to every rank
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--grid", "24,24,12",
                        "--cpu-layers", "6", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "DoF*steps/s" and line["higher_is_better"] is True
    sys.path.insert(0, str(ROOT))
    import bench
    assert line["cpu_baseline"]["cores"] == bench.host_cores() and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": "DoF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    class A:
        grid, precond, bj_rows, cheb_degree, cheb_ratio = "24,24,12", "none", 64, 8, 100.0
    assert line["config"] == json.loads(json.dumps(bench.config_of(A, bench.WORKLOADS["c5"], (24, 24, 12))))
    assert "1/2 slab" in line["cpu_baseline"]["sample"]
    # rank != 0 prints nothing and exits 0
    r2 = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                        env=dict(env, RANK="1"), timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == ""
