// SELL-32 row kernel shared by tb_spmv (tb_csr.cu) and the CG iteration (tb_cg.cu).
// mul!(y, ::ThreadedSparseMatrixCSR, x), src/utils.jl:210-231: v = 0; for nz in row: v += A[nz]*x[col[nz]].
#pragma once
#include "tb_internal.cuh"

// Row (s*32 + lane) of y = A x.  Entry j of the row is at slice_ptr[s] + j*32 + lane, so one warp
// streams 256 B of values + 128 B of column ids per j, fully coalesced; the dependent x gathers hit
// L1/L2 (neighbouring rows share columns).  Loads are issued four entries ahead of their use; the
// additions stay strictly left to right (and unfused, -fmad=false) so the sum is bitwise the
// reference's.
__device__ __forceinline__ double tb_sell_row(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col,
                                              const double *__restrict__ val, const double *__restrict__ x, int64_t s,
                                              int lane) {
    const int64_t base = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - base) >> 5);
    const int *c = col + base + lane;
    const double *v = val + base + lane;
    double acc = 0.0;
    int j = 0;
    for (; j + 4 <= w; j += 4) {
        const int c0 = c[(j + 0) * 32], c1 = c[(j + 1) * 32], c2 = c[(j + 2) * 32], c3 = c[(j + 3) * 32];
        const double v0 = v[(j + 0) * 32], v1 = v[(j + 1) * 32], v2 = v[(j + 2) * 32], v3 = v[(j + 3) * 32];
        const double x0 = x[c0], x1 = x[c1], x2 = x[c2], x3 = x[c3];
        acc += v0 * x0;
        acc += v1 * x1;
        acc += v2 * x2;
        acc += v3 * x3;
    }
    for (; j < w; j++) acc += v[j * 32] * x[c[j * 32]];
    return acc;
}
