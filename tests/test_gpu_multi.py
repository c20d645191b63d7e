"""-m gpu, needs >= 2 GPUs (skipped otherwise): multi-GPU path == single-GPU path (tests/dist_check.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("p2p,cut,fused,exact", [("1", "planes", "1", "0"), ("1", "planes", "0", "0"), ("0", "planes", "1", "0"),
                                                   ("1", "rows", "1", "0"), ("1", "rows", "1", "1"), ("1", "planes", "0", "1")])
def test_partitioned_step_matches_single_gpu(world, p2p, cut, fused, exact):
    """p2p = 1: halo of p and dot products through NVLink peer windows (CUDA IPC), with the collects and the push fused into
    the CG update kernels (fused = 1) or as separate tiny kernels (fused = 0); p2p = 0: the NCCL fallback.
    exact = 1: order-independent dot products -- iteration counts must then be EQUAL and phi agree to 1e-13."""
    import os
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, TB_P2P=p2p, DIST_CUT=cut, TB_P2P_FUSED=fused, TB_DOT_EXACT=exact)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world + 10 * int(p2p) + (20 if cut == "rows" else 0) + 40 * int(fused) + 80 * int(exact)),
                        str(ROOT / "tests" / "dist_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert ("peer path on" in r.stdout) == (p2p == "1"), r.stdout[-2000:]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_general_mesh_partition_matches_single_gpu(world, p2p):
    """unstructured LV mesh, RCB ownership + renumbering + host-side cut (dist.partition_host_mesh): send lists are not
    contiguous runs, so the ranks must AGREE on the unfused peer path (tb_csr_set_halo_fused); iterations +-1 vs one GPU"""
    import os
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, TB_P2P=p2p, DIST_MESH="lv")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29740 + world + 10 * int(p2p)),
                        str(ROOT / "tests" / "dist_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert ("peer path on" in r.stdout) == (p2p == "1"), r.stdout[-2000:]
