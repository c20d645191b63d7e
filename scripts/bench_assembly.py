#!/usr/bin/env python
"""BASELINE config 3: element assembly sweep (benchmarks-linear-form.jl / benchmarks-cuda-linear-form.jl shape):
Lagrange-1, QuadratureRuleCollection(2), generate_grid(ct, (500,100,100), (-1,-1,-1), (1,1,1)) for Hexahedron (5 M
elements) and Tetrahedron (30 M elements); mass, diffusion and the two linear forms, scattered into the fixed pattern.

    python scripts/bench_assembly.py [--grid 500,100,100] [--reps 5] [--modes 2,0]

One JSON line per (cell type, form, mode): elements/s and algorithmic GB/s (nv*dim*8 coordinates + nv*8 dof
ids per element + nv^2*8 matrix entries (or nv*8 vector entries)) against the measured copy bandwidth.
Timed on the library's stream with CUDA events (tb_timer_*), after one warm-up call; setup excluded.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import thunderbolt_jl_b200 as tb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="500,100,100")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--modes", default="2,0")
    ap.add_argument("--cells", default="hex,tet")
    ap.add_argument("--warm-s", type=float, default=0.4, help="seconds of untimed repetitions before each timed loop")
    args = ap.parse_args()
    nel = tuple(int(v) for v in args.grid.split(","))
    dev = tb.B200Device(0)
    peak = 6454.6
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    D = np.diag([0.17 * 0.62 / (0.17 + 0.62), 0.019 * 0.24 / (0.019 + 0.24), 0.019 * 0.24 / (0.019 + 0.24)])
    for cname in args.cells.split(","):
        ct = tb.Hexahedron if cname == "hex" else tb.Tetrahedron
        md = tb.generate_mesh(ct, nel, (-1, -1, -1), (1, 1, 1), device=dev)
        nv = md.nv
        M = tb.B200CSRMatrix.from_mesh(dev, md)
        b = tb.B200Vector(dev, md.ndofs, 1)
        forms = {
            "mass": (lambda: tb.core.assemble_mass(dev, md, M, 2, 1.0), nv * 3 * 8 + nv * 8 + nv * nv * 8),
            "diffusion_tensor": (lambda: tb.core.assemble_diffusion(dev, md, M, 2, tb._lib.D_TENSOR, D, 1.0),
                                 nv * 3 * 8 + nv * 8 + nv * nv * 8),
            "linear_cosexp": (lambda: tb.core.assemble_source(dev, md, b, 2, tb._lib.SRC_COSEXP, [0.0], 0.3), nv * 3 * 8 + nv * 8 + nv * 8),
            "linear_normt": (lambda: tb.core.assemble_source(dev, md, b, 2, tb._lib.SRC_NORMT, [0.0], 0.3), nv * 3 * 8 + nv * 8 + nv * 8),
        }
        for mode in (int(m) for m in args.modes.split(",")):
            dev.assembly_set_mode(mode)
            for fname, (fn, bytes_per_el) in forms.items():
                fn()                                   # warm-up (builds the adjacency cache in mode 2)
                dev.sync()
                t_w = time.perf_counter()              # keep the GPU busy until its clocks have left the idle state (a box
                while time.perf_counter() - t_w < args.warm_s:   # that sat idle during host-side setup ramps up over ~0.1 s)
                    for _ in range(args.reps):
                        fn()
                    dev.sync()
                dev.timer_start()
                for _ in range(args.reps):
                    fn()
                ms = dev.timer_stop() / args.reps
                info = dev.assembly_info()
                gbs = bytes_per_el * md.ncells / (ms * 1e-3) / 1e9
                print(json.dumps({"workload": f"C3 {cname} {'x'.join(map(str, nel))}", "form": fname, "cells": md.ncells,
                                  "dofs": md.ndofs, "mode_requested": mode, "mode_used": info["last_mode"],
                                  "chunks": info["last_chunks"], "ms": ms, "elements_per_s": md.ncells / (ms * 1e-3),
                                  "algorithmic_bytes_per_element": bytes_per_el, "achieved_gbs": gbs, "peak_gbs": peak,
                                  "frac": gbs / peak}), flush=True)
        for h in (M, b, md):
            h.free()


if __name__ == "__main__":
    main()
