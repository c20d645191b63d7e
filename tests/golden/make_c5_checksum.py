"""Golden checksums of BASELINE config 5 (512x512x384 hexes, 101,320,065 dofs, FHN) for bench.py's `parity` block.

The assembled CSR image of this configuration (2.7 G nonzeros) does not fit the host, so the CPU side is the oracle's
matrix-free closed-form operator (oracle.StencilOracle, pinned against the assembled oracle on small grids by
tests/test_oracle_independent.py): same LieTrotterGodunov step, same CG recurrence and stopping rule, same cell sweep.
Runs in ~10 minutes on 8 cores and 12 GB:  python tests/golden/make_c5_checksum.py
Output: tests/golden/c5_checksum.json -- per step: CG iterations, sum(phi), sum(phi^2), sum(s) and phi at 1024 sample
nodes addressed by grid indices (a, b, c), so that a partitioned multi-GPU run can look its own nodes up by coordinate.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402
import bench  # noqa: E402

NSTEPS = 3


def sample_nodes(nel, nsamples=1024, seed=0):
    """Half uniformly random, half in a band around the initial wave front (x = L/2 or y = L/2)."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = nel
    half = nsamples // 2
    uni = np.stack([rng.integers(0, nx + 1, half), rng.integers(0, ny + 1, half), rng.integers(0, nz + 1, half)], axis=1)
    band = np.stack([nx // 2 + rng.integers(-6, 7, half), rng.integers(0, ny // 2 + 7, half), rng.integers(0, nz + 1, half)], axis=1)
    swap = rng.random(half) < 0.5
    band[swap, 0], band[swap, 1] = band[swap, 1].copy(), band[swap, 0].copy()
    corners = np.array([[a, b, c] for a in (0, nx) for b in (0, ny) for c in (0, nz)])
    return np.unique(np.concatenate([uni, band, corners]), axis=0)


def main(workload="c5", grid=None, out=None):
    W = bench.WORKLOADS[workload]
    nel = tuple(grid) if grid else W["nel"]
    h, dt = W["h"], W["dt"]
    kap = np.asarray(W["kappa"], dtype=np.float64)
    S = O.StencilOracle(nel, h, kap, O.FHN, O.default_params(O.FHN))
    lengths = tuple(n * h for n in nel)
    x = S.node_coords()
    u = bench.initial_state(x, "fhn", lengths, None)
    del x
    nodes = sample_nodes(nel)
    gid = S.node_id(nodes[:, 0], nodes[:, 1], nodes[:, 2])
    n = S.n
    steps = []
    t = 0.0
    for s in range(NSTEPS):
        t0 = time.time()
        it, rn, conv = S.step(u, t, dt)
        t += dt
        phi = u[:n]
        steps.append({"step": s + 1, "iters": it, "converged": bool(conv), "rnorm": rn, "sum_phi": float(np.sum(phi)),
                      "sum_phi2": float(np.dot(phi, phi)), "sum_s": float(np.sum(u[n:])), "phi_max": float(phi.max()),
                      "phi_min": float(phi.min()), "phi_samples": [float(v) for v in phi[gid]]})
        print(f"step {s + 1}: {it} iterations, {time.time() - t0:.1f} s, sum(phi) = {steps[-1]['sum_phi']:.15e}", flush=True)
    doc = {"workload": W["name"], "nel": list(nel), "h": h, "kappa": list(kap), "dt": dt, "dofs": n, "model": "fhn",
           "cg": {"atol": O.SQRT_EPS, "rtol": O.SQRT_EPS},
           "generator": "tests/golden/make_c5_checksum.py: oracle.StencilOracle (CPU, closed-form 27-point operator, matrix-free)",
           "sample_nodes": nodes.tolist(), "steps": steps}
    out = Path(out) if out else ROOT / "tests" / "golden" / f"{workload}_checksum.json"
    out.write_text(json.dumps(doc))
    print("wrote", out)


if __name__ == "__main__":
    if len(sys.argv) > 1:   # small grids for the tests: make_c5_checksum.py 64,64,16 out.json
        main(grid=[int(v) for v in sys.argv[1].split(",")], out=sys.argv[2])
    else:
        main()
