// Preconditioners for the inner CG beyond point Jacobi.
//
// The reference hands LinearSolve a `precs` function: KrylovJL_CG(precs = (A, p) -> (P, I), ldiv = false), where P is a
// KrylovPreconditioners object whose mul!(z, P, r) applies the INVERSE (bak/examples-gpu/spiral-wave.jl:95-105, tip at
// docs/src/literate-tutorials/ep01_spiral-wave.jl:129-131).  Two such P live here; both are fixed SPD linear operators,
// so Krylov's PCG recurrence (tb_cg.cu) stays valid:
//
//   TB_PRECOND_BLOCK_JACOBI   BlockJacobiPreconditioner(A, nblocks): rows are split into nblocks blocks (caller-given
//       block id per row -- the reference gets it from Metis -- or equal contiguous ranges of the dof numbering), each
//       diagonal block is inverted densely in fp64 at update! time and z = blockdiag(A)^-1 r is a batched dense
//       mat-vec: one CTA per block, r_b staged in shared memory, lane i accumulates row i reading the (symmetric)
//       inverse column-wise so that the 32 lanes of a warp read 256 contiguous bytes per term.
//   TB_PRECOND_CHEBYSHEV      z = q_d(D^-1 A) D^-1 r, the degree-d Chebyshev polynomial that minimises the residual of
//       A z = r over [lmax/ratio, lmax], D = diag(A), lmax = the Gershgorin bound max_i sum_j |a_ij| / a_ii
//       (deterministic, no power iteration).  d - 1 SpMVs and d fused vector kernels per application; cuts the CG
//       iteration count (dot products, host polls) by about d on stiff meshes.
#include "tb_internal.cuh"
#include "tb_spmv.cuh"
#include <math.h>

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

// ------------------------------------------------------------------------------------------------------------------
// Gershgorin bound of D^-1 A
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_gershgorin(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const double *__restrict__ val, int64_t nrows,
                 int64_t nslices, double *partials, unsigned *ticket, double *result) {
    __shared__ double sm[32];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double m = 0.0;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        const int64_t base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5;
        const int64_t row = s * TB_SLICE + lane;
        double sum = 0.0, diag = 0.0;
        for (int64_t j = 0; j < w; j++) {
            const double v = val[base + j * TB_SLICE + lane];
            sum += fabs(v);
            if (col[base + j * TB_SLICE + lane] == (int)row && v != 0.0) diag = v;   // padding has value 0 at the own column
        }
        if (row < nrows && diag > 0.0) m = fmax(m, sum / diag);
    }
    double b = tb_block_max(m, sm);
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = b;
        __threadfence();
        const unsigned tk = atomicInc(ticket, gridDim.x - 1);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double t = 0.0;
        for (unsigned k = threadIdx.x; k < gridDim.x; k += blockDim.x) t = fmax(t, ((volatile double *)partials)[k]);
        t = tb_block_max(t, sm);
        if (threadIdx.x == 0) *result = t;
    }
}

int32_t tb_pc_gershgorin(tb_ctx *ctx, const tb_csr *A, double *lmax) {
    const tb_pattern *pat = A->pat;
    TB_LAUNCH(ctx, k_gershgorin, tb_grid_for(ctx, pat->nslices * 32, 256, 4), 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val,
              pat->nrows, pat->nslices, ctx->d_partials + 2 * TB_MAX_PARTIALS, ctx->d_ticket + 2, ctx->d_scalar + 2);
    TB_CUDA(cudaMemcpyAsync(ctx->h_scalar + 2, ctx->d_scalar + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    double v = ctx->h_scalar[2];
    if (ctx->has_comm && ctx->nranks > 1) TB_TRY(tb_comm_allreduce_max(ctx, &v));
    *lmax = v;
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Chebyshev: vector kernels (the SpMV between them is the CG's own staged sweep, launched by tb_cg.cu)
//   first:  d = (1/theta) D^-1 r;  z = d;  res = r
//   step :  res -= w (= A d);  d = c1 d + c2 D^-1 res;  z += d
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cheb_first(const double *__restrict__ r, const double *__restrict__ dinv, double *__restrict__ d,
                                                    double *__restrict__ z, double *__restrict__ res, double inv_theta, int64_t n,
                                                    const CGState *st) {
    if (st->done) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double rv = r[i];
        const double dv = inv_theta * (dinv[i] * rv);
        d[i] = dv;
        z[i] = dv;
        res[i] = rv;
    }
}
__global__ void __launch_bounds__(256) k_cheb_step(const double *__restrict__ w, const double *__restrict__ dinv, double *__restrict__ d,
                                                   double *__restrict__ z, double *__restrict__ res, double c1, double c2, int64_t n,
                                                   const CGState *st) {
    if (st->done) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double rv = res[i] - w[i];
        const double dv = c1 * d[i] + c2 * (dinv[i] * rv);
        res[i] = rv;
        d[i] = dv;
        z[i] += dv;
    }
}

int32_t tb_pc_cheb_first(tb_ctx *ctx, const double *r, const double *dinv, double *d, double *z, double *res, double inv_theta,
                         int64_t n, const CGState *st) {
    TB_LAUNCH(ctx, k_cheb_first, TB_GRID(ctx, k_cheb_first, 256, 0, (n + 255) / 256), 256, 0, r, dinv, d, z, res, inv_theta, n, st);
    return TB_OK;
}
int32_t tb_pc_cheb_step(tb_ctx *ctx, const double *w, const double *dinv, double *d, double *z, double *res, double c1, double c2,
                        int64_t n, const CGState *st) {
    TB_LAUNCH(ctx, k_cheb_step, TB_GRID(ctx, k_cheb_step, 256, 0, (n + 255) / 256), 256, 0, w, dinv, d, z, res, c1, c2, n, st);
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Block Jacobi
// ------------------------------------------------------------------------------------------------------------------
struct tb_bj {
    int64_t nrows = 0, nblocks = 0;
    int max_bs = 0;
    int *d_perm = nullptr;        // rows grouped by block (ascending row id inside a block)
    int *d_pos = nullptr;         // per row: position inside its block
    int *d_blk = nullptr;         // per row: block id
    int64_t *d_bptr = nullptr;    // nblocks+1 offsets into perm
    int64_t *d_iptr = nullptr;    // nblocks+1 offsets into inv (doubles)
    double *d_inv = nullptr;      // dense inverses, column-major == row-major (symmetric)
    int64_t inv_len = 0;
};

void tb_bj_free(tb_bj *b) {
    if (!b) return;
    cudaFree(b->d_perm);
    cudaFree(b->d_pos);
    cudaFree(b->d_blk);
    cudaFree(b->d_bptr);
    cudaFree(b->d_iptr);
    cudaFree(b->d_inv);
    delete b;
}

extern "C" int32_t tb_cg_set_block_jacobi(tb_ctx *ctx, int64_t nrows, int64_t nblocks, const int32_t *row_block) {
    TB_REQUIRE(ctx && nrows > 0 && nblocks >= 1, "tb_cg_set_block_jacobi: bad argument");
    TB_REQUIRE(nrows < ((int64_t)1 << 31), "tb_cg_set_block_jacobi: too many rows");
    TB_DEV(ctx);
    if (ctx->bj) {
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        tb_bj_free(ctx->bj);
        ctx->bj = nullptr;
    }
    if (nblocks > nrows) nblocks = nrows;
    std::vector<int> blk((size_t)nrows);
    if (row_block) {
        for (int64_t i = 0; i < nrows; i++) {
            TB_REQUIRE(row_block[i] >= 0 && row_block[i] < nblocks, "tb_cg_set_block_jacobi: block id out of range at row %lld", (long long)i);
            blk[(size_t)i] = row_block[i];
        }
    } else {   // equal contiguous ranges of the dof numbering
        for (int64_t i = 0; i < nrows; i++) blk[(size_t)i] = (int)((i * nblocks) / nrows);
    }
    std::vector<int64_t> bptr((size_t)nblocks + 1, 0), iptr((size_t)nblocks + 1, 0);
    for (int64_t i = 0; i < nrows; i++) bptr[(size_t)blk[(size_t)i] + 1]++;
    int max_bs = 0;
    for (int64_t b = 0; b < nblocks; b++) {
        const int64_t bs = bptr[(size_t)b + 1];
        if (bs > max_bs) max_bs = (int)bs;
        iptr[(size_t)b + 1] = iptr[(size_t)b] + bs * bs;
        bptr[(size_t)b + 1] += bptr[(size_t)b];
    }
    TB_REQUIRE(max_bs <= 2048, "tb_cg_set_block_jacobi: largest block has %d rows (limit 2048); use more blocks", max_bs);
    std::vector<int> perm((size_t)nrows), pos((size_t)nrows);
    std::vector<int64_t> cur(bptr.begin(), bptr.end() - 1);
    for (int64_t i = 0; i < nrows; i++) {
        const int b = blk[(size_t)i];
        pos[(size_t)i] = (int)(cur[(size_t)b] - bptr[(size_t)b]);
        perm[(size_t)cur[(size_t)b]++] = (int)i;
    }
    tb_bj *B = new (std::nothrow) tb_bj();
    if (!B) return tb_fail(TB_ERR_NOMEM, "tb_cg_set_block_jacobi: host allocation failed");
    B->nrows = nrows;
    B->nblocks = nblocks;
    B->max_bs = max_bs;
    B->inv_len = iptr[(size_t)nblocks];
    cudaError_t e = cudaSuccess;
    auto up = [&](void **d, const void *h, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(d, bytes ? bytes : 8);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    up((void **)&B->d_perm, perm.data(), sizeof(int) * (size_t)nrows);
    up((void **)&B->d_pos, pos.data(), sizeof(int) * (size_t)nrows);
    up((void **)&B->d_blk, blk.data(), sizeof(int) * (size_t)nrows);
    up((void **)&B->d_bptr, bptr.data(), sizeof(int64_t) * (size_t)(nblocks + 1));
    up((void **)&B->d_iptr, iptr.data(), sizeof(int64_t) * (size_t)(nblocks + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&B->d_inv, sizeof(double) * (size_t)(B->inv_len ? B->inv_len : 1));
    if (e != cudaSuccess) {
        tb_bj_free(B);
        return tb_fail(e == cudaErrorMemoryAllocation ? TB_ERR_NOMEM : TB_ERR_CUDA, "tb_cg_set_block_jacobi: %s (dense inverses need %.1f MB)",
                       cudaGetErrorString(e), (double)B->inv_len * 8e-6);
    }
    ctx->bj = B;
    ctx->pc_version = ~0ull;   // the inverses have to be rebuilt for the new plan
    ctx->pc_uid = 0;
    return TB_OK;
}

// dense diagonal blocks out of the SELL image: one warp per slice, lane = row
__global__ void __launch_bounds__(256)
    k_bj_extract(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const double *__restrict__ val, int64_t nrows,
                 int64_t nslices, const int *__restrict__ blk, const int *__restrict__ pos, const int64_t *__restrict__ bptr,
                 const int64_t *__restrict__ iptr, double *__restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        const int64_t base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5;
        const int64_t row = s * TB_SLICE + lane;
        if (row >= nrows) continue;
        const int b = blk[row];
        const int64_t bs = bptr[b + 1] - bptr[b];
        double *dst = inv + iptr[b] + (int64_t)pos[row] * bs;
        for (int64_t j = 0; j < w; j++) {
            const int c = col[base + j * TB_SLICE + lane];
            const double v = val[base + j * TB_SLICE + lane];
            if (c < nrows && v != 0.0 && blk[c] == b) dst[pos[c]] = v;
        }
    }
}

// In-place Gauss-Jordan inversion without pivoting (the blocks are principal submatrices of an SPD matrix), one CTA per
// block, the block stays in global memory (L1/L2 resident for the sizes in use).  Fixed elimination order: deterministic.
__global__ void __launch_bounds__(512) k_bj_invert(const int64_t *__restrict__ bptr, const int64_t *__restrict__ iptr,
                                                   double *__restrict__ inv, int64_t nblocks, int *fail) {
    extern __shared__ double s_col[];   // column k of the current step (max_bs doubles) + row k
    for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int n = (int)(bptr[b + 1] - bptr[b]);
        double *a = inv + iptr[b];
        double *s_row = s_col + n;
        for (int k = 0; k < n; k++) {
            __syncthreads();
            const double piv = a[(int64_t)k * n + k];
            if (threadIdx.x == 0 && !(piv > 0.0)) *fail = 1;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                s_col[i] = a[(int64_t)i * n + k];
                s_row[i] = a[(int64_t)k * n + i];
            }
            __syncthreads();
            const double ip = 1.0 / piv;
            for (int64_t e = threadIdx.x; e < (int64_t)n * n; e += blockDim.x) {
                const int i = (int)(e / n), j = (int)(e % n);
                double v;
                if (i == k) v = (j == k) ? ip : s_row[j] * ip;
                else if (j == k) v = -s_col[i] * ip;
                else v = a[e] - s_col[i] * (s_row[j] * ip);
                a[e] = v;
            }
        }
        __syncthreads();
    }
}

// z_b = inv_b * r_b as a batched dense mat-vec.  One CTA per block, one warp per tile of 32 rows: lane i owns row i and walks
// the columns j = 0 .. n-1 left to right (fixed order: deterministic), reading column i of the (symmetric) inverse so that
// the 32 lanes fetch 256 contiguous bytes per term; r_b never touches shared memory -- a warp keeps 32 entries in
// registers and broadcasts them with shuffles.  No barriers, eight independent 256-byte loads in flight per warp: the
// kernel streams sum(n_b^2) * 8 bytes per application and is HBM bound.
// (Round-2 experiment, removed: stream only the lower triangle of the -- symmetrised -- inverse into a padded shared-memory
// image and walk that instead: half the DRAM bytes (350 vs 494 MB on C4 with 64-row blocks), but the load / barrier / compute
// phases of a CTA serialise and the index arithmetic fills the issue slots: 196 us against this kernel's 102 us.)
__global__ void __launch_bounds__(256) k_bj_apply(const double *__restrict__ r, double *__restrict__ z, const int *__restrict__ perm,
                                                  const int64_t *__restrict__ bptr, const int64_t *__restrict__ iptr,
                                                  const double *__restrict__ inv, int64_t nblocks, const CGState *st) {
    if (st && st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const int64_t p0 = bptr[b];
        const int n = (int)(bptr[b + 1] - p0);
        const double *a = inv + iptr[b];
        for (int i0 = warp * 32; i0 < n; i0 += nw * 32) {
            const int i = i0 + lane;
            const int ic = i < n ? i : n - 1;              // idle lanes read a valid column, their sum is dropped
            double acc = 0.0;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int jj = j0 + lane;
                const double xv = jj < n ? r[perm[p0 + jj]] : 0.0;
                const int jn = n - j0 < 32 ? n - j0 : 32;
                int j = 0;
                for (; j + 8 <= jn; j += 8) {
                    double av[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) av[q] = a[(int64_t)(j0 + j + q) * n + ic];
#pragma unroll
                    for (int q = 0; q < 8; q++) acc += av[q] * __shfl_sync(0xffffffffu, xv, j + q);
                }
                for (; j < jn; j++) acc += a[(int64_t)(j0 + j) * n + ic] * __shfl_sync(0xffffffffu, xv, j);
            }
            if (i < n) z[perm[p0 + i]] = acc;
        }
    }
}

// update!(P, A): rebuild the dense inverses from the current values of A
int32_t tb_pc_bj_update(tb_ctx *ctx, const tb_csr *A) {
    tb_bj *B = ctx->bj;
    TB_REQUIRE(B, "block-Jacobi preconditioner selected but tb_cg_set_block_jacobi was not called");
    const tb_pattern *pat = A->pat;
    TB_REQUIRE(B->nrows == pat->nrows, "block-Jacobi plan is for %lld rows, operator has %lld", (long long)B->nrows, (long long)pat->nrows);
    TB_CUDA(cudaMemsetAsync(B->d_inv, 0, sizeof(double) * (size_t)B->inv_len, ctx->stream));
    TB_LAUNCH(ctx, k_bj_extract, tb_grid_for(ctx, pat->nslices * 32, 256, 8), 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val, pat->nrows,
              pat->nslices, B->d_blk, B->d_pos, B->d_bptr, B->d_iptr, B->d_inv);
    int *d_fail = (int *)(ctx->d_scalar + 4);
    TB_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
    const size_t smem = sizeof(double) * 2 * (size_t)B->max_bs;
    const int grid = (int)(B->nblocks < (int64_t)ctx->sm_count * 2 ? B->nblocks : (int64_t)ctx->sm_count * 2);
    TB_LAUNCH(ctx, k_bj_invert, grid, 512, smem, B->d_bptr, B->d_iptr, B->d_inv, B->nblocks, d_fail);
    int fail = 0;
    TB_CUDA(cudaMemcpyAsync(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (fail) return tb_fail(TB_ERR_INVALID, "block-Jacobi: a diagonal block is not positive definite");
    return TB_OK;
}

int32_t tb_pc_bj_apply(tb_ctx *ctx, const double *r, double *z, const CGState *st) {
    tb_bj *B = ctx->bj;
    int threads = (B->max_bs + 31) / 32 * 32;          // one warp per 32-row tile of the largest block, at most 8 warps
    if (threads > 256) threads = 256;
    const int64_t cap = (int64_t)ctx->sm_count * (2048 / threads);
    const int grid = (int)(B->nblocks < cap ? B->nblocks : cap);
    TB_LAUNCH(ctx, k_bj_apply, grid, threads, 0, r, z, B->d_perm, B->d_bptr, B->d_iptr, B->d_inv, B->nblocks, st);
    return TB_OK;
}

extern "C" int32_t tb_cg_set_chebyshev(tb_ctx *ctx, int32_t degree, double ratio) {
    TB_REQUIRE(ctx, "tb_cg_set_chebyshev: ctx is NULL");
    TB_REQUIRE(degree >= 1 && degree <= 64 && ratio > 1.0, "tb_cg_set_chebyshev: need 1 <= degree <= 64 and ratio > 1");
    ctx->cheb_degree = degree;
    ctx->cheb_ratio = ratio;
    return TB_OK;
}
