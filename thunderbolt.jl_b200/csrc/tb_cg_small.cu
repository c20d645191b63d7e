// Conjugate gradients as ONE persistent cooperative kernel, for operators small enough that every lane of a
// one-CTA-per-SM grid owns at most 4 rows (C1: 66 k rows; limit 148*16*32*4 = 303 k rows -- above that the
// TMA-staged multi-kernel path is faster, measured on C2).
//
// Same recurrence as tb_cg.cu (Krylov.jl cg!, LinearSolve defaults, x0 = 0); what changes is the execution model:
//   * one launch per solve instead of 3 per iteration: at these sizes the multi-kernel path is launch-latency
//     bound (C1: 0.39 ms/step for 12 iterations of ~2 us of work each);
//   * x, r, p and Ap of a lane's rows live in REGISTERS for the whole solve; only p is also written to HBM/L2
//     because the SpMV of the next iteration gathers it across lanes;
//   * the two dot products and the visibility of p are grid-wide barriers (cooperative groups grid.sync());
//     every CTA adds the per-CTA partials in the same fixed order, so all CTAs hold bitwise identical scalars and
//     no scalar ever round-trips through the host;
//   * the matrix stream (values + compressed column stream) is re-read by the same warps every iteration and stays
//     L2 resident when it fits (C1 entirely, C2 mostly).
// Row sums keep the reference's left-to-right order (bitwise SpMV); the dot products are summed in a different
// tree than the multi-kernel path, so iterates agree with it (and with the oracle) to rounding, not bitwise.
#include <cooperative_groups.h>
#include "tb_internal.cuh"
#include "tb_spmv.cuh"

namespace cgp = cooperative_groups;

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))
#define PCG_THREADS 512
#define PCG_WARPS (PCG_THREADS / 32)
#define PCG_MAX_RPL 4   // measured: 1 row/lane 11.9 us/iteration (multi-kernel 33), 8 rows/lane 78 us (multi-kernel 61): crossover near 5

// Row (s*32 + lane) of y = A x with the compressed column stream (tb_csr.cu): slot j holds one offset for all
// 32 lanes (col = row + off) or TB_CCOL_EXPLICIT followed by an explicit 32-id block.
__device__ __forceinline__ double pcg_row_cc(const int64_t *__restrict__ slice_ptr, const int *__restrict__ ccol,
                                             const int64_t *__restrict__ cptr, const double *__restrict__ val,
                                             const double *__restrict__ x, int64_t s, int lane) {
    const int64_t base = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - base) >> 5);
    const int *hdr = ccol + cptr[s];
    const int *ex = hdr + ((w + 3) & ~3) + lane;
    const double *v = val + base + lane;
    const int row = (int)(s * TB_SLICE) + lane;
    double acc = 0.0;
    int e = 0, j = 0;
    for (; j + 4 <= w; j += 4) {
        int c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int h = hdr[j + k];
            c[k] = row + h;
            if (h == TB_CCOL_EXPLICIT) c[k] = ex[(e++) * 32];
        }
        const double v0 = v[(j + 0) * 32], v1 = v[(j + 1) * 32], v2 = v[(j + 2) * 32], v3 = v[(j + 3) * 32];
        const double x0 = x[c[0]], x1 = x[c[1]], x2 = x[c[2]], x3 = x[c[3]];
        acc += v0 * x0;
        acc += v1 * x1;
        acc += v2 * x2;
        acc += v3 * x3;
    }
    for (; j < w; j++) {
        const int h = hdr[j];
        int c = row + h;
        if (h == TB_CCOL_EXPLICIT) c = ex[(e++) * 32];
        acc += v[j * 32] * x[c];
    }
    return acc;
}

struct PcgMat {
    const int64_t *slice_ptr;
    const int *col;       // uncompressed SELL column ids
    const int *ccol;      // compressed stream (nullptr: use col)
    const int64_t *cptr;
};

__device__ __forceinline__ double pcg_row(const PcgMat &P, const double *__restrict__ val, const double *__restrict__ x, int64_t s,
                                          int lane) {
    return P.ccol ? pcg_row_cc(P.slice_ptr, P.ccol, P.cptr, val, x, s, lane) : tb_sell_row(P.slice_ptr, P.col, val, x, s, lane);
}

// block partial -> partials[blockIdx.x]; grid barrier; every CTA adds all partials in the same order
__device__ __forceinline__ double pcg_allsum(double v, double *partials, double *sm, cgp::grid_group &grid) {
    const double bs = tb_block_sum(v, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    grid.sync();
    double s = 0.0;
    if (threadIdx.x < 32) {
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) s += ((volatile double *)partials)[i];
        s = tb_warp_sum(s);
        if (threadIdx.x == 0) sm[32] = s;
    }
    __syncthreads();
    return sm[32];
}

// FROM_B: r = p = b.  Otherwise r = p = M*phi (+ bS)  ("b = M u_{n-1}" + add!(b, S), euler.jl:85-91).
template <int RPL, bool FROM_B>
__global__ void __launch_bounds__(PCG_THREADS, 1)
    k_cg_persistent(const PcgMat P, const double *__restrict__ Aval, const double *__restrict__ Mval,
                    const double *__restrict__ src, const double *__restrict__ bS, double *__restrict__ x_out,
                    double *__restrict__ pglob, int64_t nrows, int64_t nslices, CGState *st, double *partials,
                    const double *__restrict__ dinv) {
    cgp::grid_group grid = cgp::this_grid();
    __shared__ double sm[34];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * PCG_WARPS + warp, nw = (int64_t)gridDim.x * PCG_WARPS;
    double x[RPL], r[RPL], p[RPL], Ap[RPL], di[RPL];   // di = 1/a_ii (Jacobi) or 1
    bool own[RPL];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < RPL; k++) {
        const int64_t s = gw + k * nw;
        const int64_t row = s * TB_SLICE + lane;
        own[k] = s < nslices && row < nrows;
        x[k] = 0.0;
        r[k] = p[k] = Ap[k] = 0.0;
        di[k] = (dinv && own[k]) ? dinv[row] : 1.0;
        if (s < nslices) {
            double v;
            if (FROM_B) v = own[k] ? src[row] : 0.0;
            else {
                v = pcg_row(P, Mval, src, s, lane);
                if (own[k] && bS) v += bS[row];
            }
            if (own[k]) {
                const double z = dinv ? di[k] * v : v;         // z = M r, p = z, gamma = r.z
                r[k] = v;
                p[k] = z;
                pglob[row] = z;
                acc += v * z;
            }
        }
    }
    double gamma = pcg_allsum(acc, partials, sm, grid);        // also makes pglob visible grid-wide
    double rn = sqrt(gamma);
    const double eps = st->atol + st->rtol * rn;
    const long long itmax = st->itmax;
    bool solved = rn <= eps;
    long long iter = 0;
    while (!solved && iter < itmax) {
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < RPL; k++) {
            const int64_t s = gw + k * nw;
            if (s < nslices) {
                const double v = pcg_row(P, Aval, pglob, s, lane);
                if (own[k]) {
                    Ap[k] = v;
                    acc += p[k] * v;
                }
            }
        }
        const double pAp = pcg_allsum(acc, partials + gridDim.x, sm, grid);
        const double alpha = gamma / pAp;
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < RPL; k++)
            if (own[k]) {
                x[k] += alpha * p[k];
                r[k] -= alpha * Ap[k];
                acc += dinv ? r[k] * (di[k] * r[k]) : r[k] * r[k];
            }
        const double gnext = pcg_allsum(acc, partials + 2 * gridDim.x, sm, grid);
        rn = sqrt(gnext);
        solved = rn <= eps;
        iter++;
        if (rn != rn) iter = itmax;   // NaN never recovers: what the reference reports after grinding through itmax iterations
        if (!solved && iter < itmax) {
            const double beta = gnext / gamma;
            gamma = gnext;
#pragma unroll
            for (int k = 0; k < RPL; k++)
                if (own[k]) {
                    p[k] = (dinv ? di[k] * r[k] : r[k]) + beta * p[k];
                    pglob[(gw + k * nw) * TB_SLICE + lane] = p[k];
                }
            grid.sync();                                       // p complete before anyone gathers it
        }
    }
#pragma unroll
    for (int k = 0; k < RPL; k++)
        if (own[k]) x_out[(gw + k * nw) * TB_SLICE + lane] = x[k];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->gamma = gamma;
        st->rnorm = rn;
        st->eps = eps;
        st->iter = iter;
        st->solved = solved;
        st->done = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Two barriers per iteration instead of three, and cheaper ones (default; TB_PCG_V=1 selects the kernel above).
//
//   * The barrier is a ticket barrier that carries the reduction: a CTA stores its partial and takes a ticket; whoever draws
//     the epoch's last ticket adds the partials in slot order and publishes the total with a release store; one thread per
//     CTA spins on that one word with ld.acquire.gpu (which also drops the SM's stale L1 lines).  Against grid.sync()
//     followed by every CTA reading every partial this saves a round trip and 148 x 148 loads on ten cache lines.
//   * The third barrier ("p complete before anyone gathers it") is gone: the SpMV forms the direction at the gathered
//     columns itself, p_c = z_c + beta * p_old_c, from the z published before the r.z barrier and the PREVIOUS direction,
//     which has been globally visible for a whole iteration -- the same unfused multiply and add the owner of row c
//     performs, hence the same bits.  The owner writes its new p into the other half of a ping-pong pair; the two
//     barriers of the next iteration publish it before it becomes "previous".  On the 101 M-row operator this trade
//     (two gathers per entry for one kernel) LOSES because the gather is the bandwidth bottleneck (DESIGN, fused-p
//     experiment); here everything is L2 resident and the iteration is latency bound, so a barrier is worth more than
//     nine extra L2 hits per row.
//   Write-after-read hazards: z_c of iteration k+1 overwrites z_c of iteration k after the p.Ap barrier of iteration k+1,
//   which every CTA reaches only after its SpMV -- the last reader of the old z; the same barrier separates the last
//   read of p_{k-1} from the write of p_{k+1} into its slot.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pcg_st_release(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pcg_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pcg_ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
struct PcgBar {
    double *vals;                 // [gridDim.x] partial sums of the current epoch
    double *total;                // the sum, written by the CTA that arrived last
    unsigned long long *flag;     // epoch whose total is published (zeroed before the launch, like the ticket)
    unsigned long long *ticket;   // arrivals so far, all epochs
    unsigned long long epoch;
    bool dead;
};
// Arrive: publish the CTA's partial, take a ticket.  The CTA that draws the last ticket of the epoch adds all partials in slot
// order (one warp, independent loads, the same order whoever happens to be last: deterministic), stores the total and
// releases the epoch; everybody else has ONE thread spinning on ONE word.  (First attempt: every CTA polls every CTA's flag
// and adds the partials itself -- 148 x 148 loads per poll round on ten cache lines, measured SLOWER than the cooperative
// grid.sync() kernel: 17.7 vs 15.0 us per iteration on C1.)
__device__ __forceinline__ double pcg_allsum2(double v, PcgBar &B, double *sm) {
    const double bs = tb_block_sum(v, sm);          // valid in warp 0; its __syncthreads order the CTA's earlier global stores before thread 0
    B.epoch++;
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        B.vals[blockIdx.x] = bs;
        __threadfence();                            // release: partial, and the CTA's z / p stores, before the ticket
        const unsigned long long t = atomicAdd(B.ticket, 1ull);
        s_last = (t + 1 == B.epoch * gridDim.x);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();                            // acquire side of the ticket
        double s = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) s += __ldcg(B.vals + i);
        s = tb_warp_sum(s);
        if (threadIdx.x == 0) {
            *B.total = s;
            pcg_st_release(B.flag, B.epoch);
        }
    }
    if (threadIdx.x == 0) {
        // a CTA that never arrives must not hang the GPU: after 10 s the sum is poisoned (NaN ends the solve as "not
        // converged") and later waits fall through
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (!B.dead && pcg_ld_acquire(B.flag) < B.epoch) {
            if ((++spins & 255u) == 0) {
                const unsigned long long now = tb_globaltimer();
                if (t0 == 0) t0 = now;
                if (now - t0 > 10000000000ull) B.dead = true;
            }
        }
        sm[32] = B.dead ? __longlong_as_double(0x7ff8000000000000ll) : __ldcg(B.total);
    }
    __syncthreads();
    const double out = sm[32];
    __syncthreads();
    return out;
}

// gathered operand: the direction at column c (FUSED: formed on the fly, see above); L2 loads, never a stale L1 line
template <bool FUSED>
__device__ __forceinline__ double pcg_dir(const double *__restrict__ zg, const double *__restrict__ pold, double beta, int c) {
    if (FUSED) return __ldcg(zg + c) + beta * __ldcg(pold + c);
    return __ldcg(pold + c);
}
template <bool FUSED>
__device__ __forceinline__ double pcg_row2(const PcgMat &P, const double *__restrict__ val, const double *__restrict__ zg,
                                           const double *__restrict__ pold, double beta, int64_t s, int lane) {
    const int64_t base = P.slice_ptr[s];
    const int w = (int)((P.slice_ptr[s + 1] - base) >> 5);
    const double *v = val + base + lane;
    double acc = 0.0;
    if (P.ccol) {
        const int *hdr = P.ccol + P.cptr[s];
        const int *ex = hdr + ((w + 3) & ~3) + lane;
        const int row = (int)(s * TB_SLICE) + lane;
        int e = 0, j = 0;
        for (; j + 4 <= w; j += 4) {
            int c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int h = hdr[j + k];
                c[k] = row + h;
                if (h == TB_CCOL_EXPLICIT) c[k] = ex[(e++) * 32];
            }
            const double v0 = v[(j + 0) * 32], v1 = v[(j + 1) * 32], v2 = v[(j + 2) * 32], v3 = v[(j + 3) * 32];
            const double x0 = pcg_dir<FUSED>(zg, pold, beta, c[0]), x1 = pcg_dir<FUSED>(zg, pold, beta, c[1]),
                         x2 = pcg_dir<FUSED>(zg, pold, beta, c[2]), x3 = pcg_dir<FUSED>(zg, pold, beta, c[3]);
            acc += v0 * x0;
            acc += v1 * x1;
            acc += v2 * x2;
            acc += v3 * x3;
        }
        for (; j < w; j++) {
            const int h = hdr[j];
            int c = row + h;
            if (h == TB_CCOL_EXPLICIT) c = ex[(e++) * 32];
            acc += v[j * 32] * pcg_dir<FUSED>(zg, pold, beta, c);
        }
    } else {
        const int *c = P.col + base + lane;
        int j = 0;
        for (; j + 4 <= w; j += 4) {
            const int c0 = c[(j + 0) * 32], c1 = c[(j + 1) * 32], c2 = c[(j + 2) * 32], c3 = c[(j + 3) * 32];
            const double v0 = v[(j + 0) * 32], v1 = v[(j + 1) * 32], v2 = v[(j + 2) * 32], v3 = v[(j + 3) * 32];
            const double x0 = pcg_dir<FUSED>(zg, pold, beta, c0), x1 = pcg_dir<FUSED>(zg, pold, beta, c1),
                         x2 = pcg_dir<FUSED>(zg, pold, beta, c2), x3 = pcg_dir<FUSED>(zg, pold, beta, c3);
            acc += v0 * x0;
            acc += v1 * x1;
            acc += v2 * x2;
            acc += v3 * x3;
        }
        for (; j < w; j++) acc += v[j * 32] * pcg_dir<FUSED>(zg, pold, beta, c[j * 32]);
    }
    return acc;
}

template <int RPL, bool FROM_B>
__global__ void __launch_bounds__(PCG_THREADS, 1)
    k_cg_persistent2(const PcgMat P, const double *__restrict__ Aval, const double *__restrict__ Mval,
                     const double *__restrict__ src, const double *__restrict__ bS, double *__restrict__ x_out,
                     double *zglob, double *pg0, double *pg1, int64_t nrows, int64_t nslices, CGState *st, double *barvals,
                     unsigned long long *barflags, const double *__restrict__ dinv) {
    __shared__ double sm[34];
    PcgBar B{barvals, barvals + gridDim.x, barflags, barflags + 1, 0ull, false};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * PCG_WARPS + warp, nw = (int64_t)gridDim.x * PCG_WARPS;
    double x[RPL], r[RPL], p[RPL], Ap[RPL], di[RPL];   // di = 1/a_ii (Jacobi) or 1
    bool own[RPL];
    double *pg[2] = {pg0, pg1};
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < RPL; k++) {
        const int64_t s = gw + k * nw;
        const int64_t row = s * TB_SLICE + lane;
        own[k] = s < nslices && row < nrows;
        x[k] = 0.0;
        r[k] = p[k] = Ap[k] = 0.0;
        di[k] = (dinv && own[k]) ? dinv[row] : 1.0;
        if (s < nslices) {
            double v;
            if (FROM_B) v = own[k] ? src[row] : 0.0;
            else {
                v = pcg_row(P, Mval, src, s, lane);     // src is never written here: cached loads are fine
                if (own[k] && bS) v += bS[row];
            }
            if (own[k]) {
                const double z = dinv ? di[k] * v : v;         // z = M r, p = z, gamma = r.z
                r[k] = v;
                p[k] = z;
                pg0[row] = z;
                acc += v * z;
            }
        }
    }
    double gamma = pcg_allsum2(acc, B, sm);                    // also publishes pg0 = p_0
    double rn = sqrt(gamma);
    const double eps = st->atol + st->rtol * rn;
    const long long itmax = st->itmax;
    bool solved = rn <= eps;
    long long iter = 0;
    double beta = 0.0;
    int cur = 0;                                               // pg[cur]: the direction every CTA can see (p_k, or p_{k-1} when FUSED)
    while (!solved && iter < itmax) {
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < RPL; k++) {
            const int64_t s = gw + k * nw;
            if (s < nslices) {
                const double v = iter == 0 ? pcg_row2<false>(P, Aval, zglob, pg[cur], beta, s, lane)
                                           : pcg_row2<true>(P, Aval, zglob, pg[cur], beta, s, lane);
                if (own[k]) {
                    Ap[k] = v;
                    acc += p[k] * v;
                }
            }
        }
        // the direction used above is now "previous" for everybody; own rows' current p goes into the other slot -- nobody
        // reads that slot before the second barrier after this point
        if (iter > 0) cur ^= 1;
        const double pAp = pcg_allsum2(acc, B, sm);
        const double alpha = gamma / pAp;
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < RPL; k++)
            if (own[k]) {
                x[k] += alpha * p[k];
                r[k] -= alpha * Ap[k];
                const double z = dinv ? di[k] * r[k] : r[k];
                zglob[(gw + k * nw) * TB_SLICE + lane] = z;
                acc += r[k] * z;
            }
        const double gnext = pcg_allsum2(acc, B, sm);          // publishes z
        rn = sqrt(gnext);
        solved = rn <= eps;
        iter++;
        if (rn != rn) iter = itmax;   // NaN never recovers: what the reference reports after grinding through itmax iterations
        if (!solved && iter < itmax) {
            beta = gnext / gamma;
            gamma = gnext;
#pragma unroll
            for (int k = 0; k < RPL; k++)
                if (own[k]) {
                    p[k] = (dinv ? di[k] * r[k] : r[k]) + beta * p[k];
                    pg[cur ^ 1][(gw + k * nw) * TB_SLICE + lane] = p[k];
                }
        }
    }
#pragma unroll
    for (int k = 0; k < RPL; k++)
        if (own[k]) x_out[(gw + k * nw) * TB_SLICE + lane] = x[k];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->gamma = gamma;
        st->rnorm = rn;
        st->eps = eps;
        st->iter = iter;
        st->solved = solved;
        st->done = 1;
    }
}

template <int RPL, bool FROM_B>
static int32_t launch_pcg2(tb_ctx *ctx, int grid, const PcgMat &P, const double *Aval, const double *Mval, const double *src,
                           const double *bS, double *x, int64_t nrows, int64_t nslices, const double *dinv) {
    CGState *st = ctx->d_cg;
    double *zglob = ctx->d_cgwork, *pg0 = zglob + ctx->cgwork_ld, *pg1 = pg0 + ctx->cgwork_ld;
    double *barvals = ctx->d_partials;
    unsigned long long *barflags = reinterpret_cast<unsigned long long *>(ctx->d_partials + 2 * TB_MAX_PARTIALS);
    TB_CUDA(cudaMemsetAsync(barflags, 0, sizeof(unsigned long long) * 2, ctx->stream));   // published epoch | ticket
    void *args[] = {(void *)&P, (void *)&Aval, (void *)&Mval, (void *)&src, (void *)&bS, (void *)&x, (void *)&zglob, (void *)&pg0,
                    (void *)&pg1, (void *)&nrows, (void *)&nslices, (void *)&st, (void *)&barvals, (void *)&barflags, (void *)&dinv};
    // cooperative launch: the hand-made barrier needs every CTA resident, which only this launch mode guarantees
    TB_CUDA(cudaLaunchCooperativeKernel((void *)k_cg_persistent2<RPL, FROM_B>, dim3(grid), dim3(PCG_THREADS), args, 0, ctx->stream));
    ctx->launches++;
    return TB_OK;
}

template <int RPL, bool FROM_B>
static int32_t launch_pcg(tb_ctx *ctx, int grid, const PcgMat &P, const double *Aval, const double *Mval, const double *src,
                          const double *bS, double *x, double *pglob, int64_t nrows, int64_t nslices, const double *dinv) {
    CGState *st = ctx->d_cg;
    double *partials = ctx->d_partials;
    void *args[] = {(void *)&P, (void *)&Aval, (void *)&Mval, (void *)&src, (void *)&bS, (void *)&x, (void *)&pglob,
                    (void *)&nrows, (void *)&nslices, (void *)&st, (void *)&partials, (void *)&dinv};
    TB_CUDA(cudaLaunchCooperativeKernel((void *)k_cg_persistent<RPL, FROM_B>, dim3(grid), dim3(PCG_THREADS), args, 0, ctx->stream));
    ctx->launches++;
    return TB_OK;
}

// deferred read-back (tb_monodomain_run): totals of a run of solves stay on the device
__global__ void k_pcg_totals_reset(CGState *st) {
    st->iter_sum = 0;
    st->all_solved = 1;
}
__global__ void k_pcg_fold(CGState *st) {
    st->iter_sum += st->iter;
    st->all_solved &= st->solved;
}
int32_t tb_cg_deferred_begin(tb_ctx *ctx) {
    TB_LAUNCH(ctx, k_pcg_totals_reset, 1, 1, 0, ctx->d_cg);
    ctx->cg_deferred = true;
    return TB_OK;
}
int32_t tb_cg_deferred_end(tb_ctx *ctx, int64_t *iters_total, int32_t *all_solved) {
    ctx->cg_deferred = false;
    CGState *h = ctx->h_cg;
    TB_CUDA(cudaMemcpyAsync(h, ctx->d_cg, sizeof(CGState), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (iters_total) *iters_total = h->iter_sum;
    if (all_solved) *all_solved = h->all_solved;
    ctx->last_cg_iters = h->iter;
    return TB_OK;
}

__global__ void k_pcg_set_tol(CGState *st, double atol, double rtol, long long itmax) {
    st->atol = atol;
    st->rtol = rtol;
    st->itmax = itmax;
    st->done = 0;
    st->solved = 0;
    st->iter = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Mid-size operators (C2: 549 k rows): same idea, different residency.  The rows no longer fit the registers of one
// wave of lanes, so x, r, p, Ap stay in global memory -- they are L2-resident at these sizes (4 vectors of 4.4 MB) --
// and the SpMV is the bulk-async (TMA + mbarrier) sweep of tb_spmv.cuh, which is what makes the large-operator path
// fast.  What is saved is everything BETWEEN the phases: three launches, three last-block reductions and their
// scalar round trips per iteration (measured on C2: 36 + 13 + 6.5 us of kernels plus gaps = 61 us per iteration).
// ---------------------------------------------------------------------------------------------------------------
struct PcgTmaArgs {
    const int64_t *slice_ptr;
    const int *col;            // SELL column ids (wide slices, tb_spmv.cuh)
    const int *cstream;        // compressed column stream (CC) or SELL column ids
    const int64_t *cptr;
    unsigned val_bytes, col_bytes;
    tb_wide_list wide;
};

__device__ __forceinline__ double pcg_allsum_t(double v, double *partials, double *sm, cgp::grid_group &grid) {
    const double bs = tb_block_sum(v, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    grid.sync();
    double s = 0.0;
    if (threadIdx.x < 32) {
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) s += ((volatile double *)partials)[i];
        s = tb_warp_sum(s);
        if (threadIdx.x == 0) sm[32] = s;
    }
    __syncthreads();
    const double out = sm[32];
    __syncthreads();
    return out;
}

template <bool FROM_B, bool CC>
__global__ void __launch_bounds__(1024, 1)
    k_cg_persistent_tma(const PcgTmaArgs P, const double *__restrict__ Aval, const double *__restrict__ Mval,
                        const double *__restrict__ src, const double *__restrict__ bS, double *__restrict__ x,
                        double *__restrict__ r, double *__restrict__ p, double *__restrict__ Ap, int64_t nrows,
                        int64_t nslices, CGState *st, double *partials, const double *__restrict__ dinv) {
    cgp::grid_group grid = cgp::this_grid();
    extern __shared__ __align__(128) unsigned char tb_dyn_smem[];
    __shared__ double sm[34];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    tb_tma_ring ring;
    double acc = 0.0;
    if (FROM_B) {
        // Row i = tid + k * nth is the row this thread's lane gets in the sweep below (warp gw takes slices gw, gw + nw, ...),
        // so the partial sums of r.z -- and with them every bit of the solve -- are those of the fused "b = M u" start.  Rows of
        // wide slices are the exception in the sweep (one warp per row, summed by its lane 0), hence also here.
        const int lane = threadIdx.x & 31;
        const int64_t gw = tid >> 5, nw = nth >> 5;
        auto start = [&](int64_t i) {
            const double v = src[i];
            const double z = dinv ? dinv[i] * v : v;
            x[i] = 0.0;
            r[i] = v;
            p[i] = z;
            acc += v * z;
        };
        for (int64_t q = gw; q < (int64_t)P.wide.n * TB_SLICE; q += nw) {
            const int64_t i = P.wide.slices[q >> 5] * TB_SLICE + (q & 31);
            if (lane == 0 && i < nrows) start(i);
        }
        for (int64_t i = tid; i < nrows; i += nth) {
            const int64_t s = i >> 5;
            if (P.wide.n > 0 && P.slice_ptr[s + 1] - P.slice_ptr[s] > (int64_t)TB_TMA_WCAP * TB_SLICE) continue;
            start(i);
        }
    } else {
        tb_sell_sweep_tma<1, CC>(P.slice_ptr, Mval, P.cstream, P.cptr, src, nslices, P.val_bytes, P.col_bytes, tb_dyn_smem,
                                 [&](int64_t row, double v) {
                                     if (row < nrows) {
                                         if (bS) v += bS[row];
                                         const double z = dinv ? dinv[row] * v : v;
                                         x[row] = 0.0;
                                         r[row] = v;
                                         p[row] = z;
                                         acc += v * z;
                                     }
                                 }, &ring, P.col, P.wide);
    }
    double gamma = pcg_allsum_t(acc, partials, sm, grid);          // also publishes p grid-wide
    double rn = sqrt(gamma);
    const double eps = st->atol + st->rtol * rn;
    const long long itmax = st->itmax;
    bool solved = rn <= eps;
    long long iter = 0;
    const int64_t n2 = nrows >> 1;
    while (!solved && iter < itmax) {
        acc = 0.0;
        tb_sell_sweep_tma<1, CC>(P.slice_ptr, Aval, P.cstream, P.cptr, p, nslices, P.val_bytes, P.col_bytes, tb_dyn_smem,
                                 [&](int64_t row, double v) {
                                     if (row < nrows) {
                                         Ap[row] = v;
                                         acc += p[row] * v;
                                     }
                                 }, &ring, P.col, P.wide);
        const double pAp = pcg_allsum_t(acc, partials + gridDim.x, sm, grid);
        const double alpha = gamma / pAp;
        acc = 0.0;
        for (int64_t i = tid; i < n2; i += nth) {
            double2 xv = reinterpret_cast<double2 *>(x)[i], rv = reinterpret_cast<double2 *>(r)[i];
            const double2 pv = reinterpret_cast<const double2 *>(p)[i], av = reinterpret_cast<const double2 *>(Ap)[i];
            xv.x += alpha * pv.x;
            xv.y += alpha * pv.y;
            rv.x -= alpha * av.x;
            rv.y -= alpha * av.y;
            reinterpret_cast<double2 *>(x)[i] = xv;
            reinterpret_cast<double2 *>(r)[i] = rv;
            if (dinv) {
                const double2 dv = reinterpret_cast<const double2 *>(dinv)[i];
                acc += rv.x * (dv.x * rv.x);
                acc += rv.y * (dv.y * rv.y);
            } else {
                acc += rv.x * rv.x;
                acc += rv.y * rv.y;
            }
        }
        if ((nrows & 1) && tid == 0) {
            const int64_t i = nrows - 1;
            const double xn = x[i] + alpha * p[i], rr = r[i] - alpha * Ap[i];
            x[i] = xn;
            r[i] = rr;
            acc += dinv ? rr * (dinv[i] * rr) : rr * rr;
        }
        const double gnext = pcg_allsum_t(acc, partials + 2 * gridDim.x, sm, grid);
        rn = sqrt(gnext);
        solved = rn <= eps;
        iter++;
        if (rn != rn) iter = itmax;
        if (!solved && iter < itmax) {
            const double beta = gnext / gamma;
            gamma = gnext;
            for (int64_t i = tid; i < n2; i += nth) {
                double2 rv = reinterpret_cast<const double2 *>(r)[i];
                if (dinv) {
                    const double2 dv = reinterpret_cast<const double2 *>(dinv)[i];
                    rv.x = dv.x * rv.x;
                    rv.y = dv.y * rv.y;
                }
                double2 pv = reinterpret_cast<double2 *>(p)[i];
                pv.x = rv.x + beta * pv.x;
                pv.y = rv.y + beta * pv.y;
                reinterpret_cast<double2 *>(p)[i] = pv;
            }
            if ((nrows & 1) && tid == 0) p[nrows - 1] = (dinv ? dinv[nrows - 1] * r[nrows - 1] : r[nrows - 1]) + beta * p[nrows - 1];
            grid.sync();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->gamma = gamma;
        st->rnorm = rn;
        st->eps = eps;
        st->iter = iter;
        st->solved = solved;
        st->done = 1;
    }
}

template <bool FROM_B, bool CC>
static int32_t launch_pcg_tma(tb_ctx *ctx, int grid, int threads, size_t smem, const PcgTmaArgs &P, const double *Aval,
                              const double *Mval, const double *src, const double *bS, double *x, double *r, double *p, double *Ap,
                              int64_t nrows, int64_t nslices, const double *dinv) {
    TB_CUDA(tb_ensure_smem(ctx, (const void *)k_cg_persistent_tma<FROM_B, CC>, smem));
    CGState *st = ctx->d_cg;
    double *partials = ctx->d_partials;
    void *args[] = {(void *)&P, (void *)&Aval, (void *)&Mval, (void *)&src, (void *)&bS, (void *)&x, (void *)&r, (void *)&p,
                    (void *)&Ap, (void *)&nrows, (void *)&nslices, (void *)&st, (void *)&partials, (void *)&dinv};
    TB_CUDA(cudaLaunchCooperativeKernel((void *)k_cg_persistent_tma<FROM_B, CC>, dim3(grid), dim3(threads), args, smem, ctx->stream));
    ctx->launches++;
    return TB_OK;
}

// 0: not eligible; 1: register-resident kernel; 2: TMA-staged kernel with vectors in global memory
int tb_cg_persistent_kind(tb_ctx *ctx, const tb_pattern *pat, int *grid_out) {
    *grid_out = 0;
    if (ctx->cg_persistent != 2)
        if (const int g = tb_cg_persistent_grid(ctx, pat)) {
            *grid_out = g;
            return 1;
        }
    if (!ctx->cg_persistent || (ctx->has_comm && ctx->nranks > 1)) return 0;
    if (!(ctx->spmv_variant > 0 && pat->max_width_tma > 0)) return 0;
    if (pat->nrows > ctx->cg_persistent_max_rows) return 0;     // large operators: launch overhead is < 2 %, keep host-side polling
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    if (!coop || 3 * ctx->sm_count > 4 * TB_MAX_PARTIALS) return 0;
    *grid_out = ctx->sm_count;
    return 2;
}

int32_t tb_cg_run_persistent_tma(tb_ctx *ctx, int grid, const tb_csr *A, const double *b, const tb_csr *M, const double *phi,
                                 const double *bS, double *x, double atol, double rtol, int64_t itmax, int64_t *iters,
                                 double *rnorm, int32_t *converged, const double *dinv) {
    const tb_pattern *pat = A->pat;
    TB_TRY(tb_ctx_ensure_cgwork(ctx, pat->ncols));
    double *r = ctx->d_cgwork, *p = r + ctx->cgwork_ld, *Ap = p + ctx->cgwork_ld;
    const bool cc = ctx->spmv_compress && pat->d_ccol != nullptr;
    const tb_tma_geom g = tb_tma_geometry(pat->max_width_tma, cc ? tb_ccol_stage_ints(pat) : 32 * pat->max_width_tma, 1, 0);
    PcgTmaArgs P{pat->d_slice_ptr, pat->d_col, cc ? pat->d_ccol : pat->d_col, pat->d_cptr, g.val_bytes, g.col_bytes, tb_wide_list()};
    P.wide.slices = pat->d_wide_slices;
    P.wide.n = (int)pat->n_wide;
    const int64_t need = (pat->nslices + g.warps - 1) / g.warps;
    if (need < grid) grid = (int)(need < 1 ? 1 : need);
    TB_LAUNCH(ctx, k_pcg_set_tol, 1, 1, 0, ctx->d_cg, atol, rtol, (long long)itmax);
    if (ctx->profile) TB_CUDA(cudaEventRecord(ctx->prof_ev[0], ctx->stream));
    const double *src = b ? b : phi;
    const double *Mval = b ? nullptr : M->d_val;
    int32_t st;
    if (b) st = cc ? launch_pcg_tma<true, true>(ctx, grid, g.warps * 32, g.smem, P, A->d_val, Mval, src, bS, x, r, p, Ap, pat->nrows, pat->nslices, dinv)
                   : launch_pcg_tma<true, false>(ctx, grid, g.warps * 32, g.smem, P, A->d_val, Mval, src, bS, x, r, p, Ap, pat->nrows, pat->nslices, dinv);
    else st = cc ? launch_pcg_tma<false, true>(ctx, grid, g.warps * 32, g.smem, P, A->d_val, Mval, src, bS, x, r, p, Ap, pat->nrows, pat->nslices, dinv)
                 : launch_pcg_tma<false, false>(ctx, grid, g.warps * 32, g.smem, P, A->d_val, Mval, src, bS, x, r, p, Ap, pat->nrows, pat->nslices, dinv);
    if (st != TB_OK) return st;
    if (ctx->cg_deferred) {               // tb_monodomain_run: fold into the device-side totals, nothing comes back now
        TB_LAUNCH(ctx, k_pcg_fold, 1, 1, 0, ctx->d_cg);
        ctx->last_cg_persistent = 2;
        if (iters) *iters = 0;
        if (rnorm) *rnorm = 0.0;
        if (converged) *converged = 1;
        return TB_OK;
    }
    if (ctx->profile) TB_CUDA(cudaEventRecord(ctx->prof_ev[1], ctx->stream));
    CGState *h = ctx->h_cg;
    TB_CUDA(cudaMemcpyAsync(h, ctx->d_cg, sizeof(CGState), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->last_cg_iters = h->iter;
    ctx->last_cg_persistent = 2;
    if (ctx->profile && h->iter > 0) {
        float ms = 0.f;
        TB_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[0], ctx->prof_ev[1]));
        ctx->prof_spmv_ms += ms;
        ctx->prof_spmv_n += h->iter;
    }
    if (iters) *iters = h->iter;
    if (rnorm) *rnorm = h->rnorm;
    if (converged) *converged = h->solved;
    return TB_OK;
}

// Is the persistent path usable for this operator on this context?  Returns the grid size, or 0.
int tb_cg_persistent_grid(tb_ctx *ctx, const tb_pattern *pat) {
    if (!ctx->cg_persistent || (ctx->has_comm && ctx->nranks > 1)) return 0;
    if (pat->n_wide > 0) return 0;   // rows are bound to lanes here: a wide slice would be one warp's serial job (use the TMA kernel)
    static int coop = -1, per_sm = 0;
    if (coop < 0) {
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persistent<PCG_MAX_RPL, false>, PCG_THREADS, 0) != cudaSuccess)
            per_sm = 0;
    }
    if (!coop || per_sm < 1) return 0;
    int64_t grid = ctx->sm_count;                              // one CTA per SM
    if (3 * grid > 4 * TB_MAX_PARTIALS) return 0;
    const int64_t need_warps = pat->nslices;
    if (need_warps > grid * PCG_WARPS * PCG_MAX_RPL) return 0;
    const int64_t g2 = (need_warps + PCG_WARPS - 1) / PCG_WARPS;
    if (g2 < grid) grid = g2 < 1 ? 1 : g2;                     // tiny operators: fewer CTAs, cheaper barriers
    return (int)grid;
}

int32_t tb_cg_run_persistent(tb_ctx *ctx, int grid, const tb_csr *A, const double *b, const tb_csr *M, const double *phi,
                             const double *bS, double *x, double atol, double rtol, int64_t itmax, int64_t *iters,
                             double *rnorm, int32_t *converged, const double *dinv) {
    const tb_pattern *pat = A->pat;
    TB_TRY(tb_ctx_ensure_cgwork(ctx, pat->ncols));
    double *pglob = ctx->d_cgwork + ctx->cgwork_ld;
    const bool cc = ctx->spmv_compress && pat->d_ccol != nullptr;
    PcgMat P{pat->d_slice_ptr, pat->d_col, cc ? pat->d_ccol : nullptr, pat->d_cptr};
    const int64_t per_pass = (int64_t)grid * PCG_WARPS;
    const int64_t rpl = (pat->nslices + per_pass - 1) / per_pass;
    TB_LAUNCH(ctx, k_pcg_set_tol, 1, 1, 0, ctx->d_cg, atol, rtol, (long long)itmax);
    if (ctx->profile) TB_CUDA(cudaEventRecord(ctx->prof_ev[0], ctx->stream));
    const double *src = b ? b : phi;
    const double *Mval = b ? nullptr : M->d_val;
#define PCG_GO(R)                                                                                                              \
    (b ? launch_pcg<R, true>(ctx, grid, P, A->d_val, Mval, src, bS, x, pglob, pat->nrows, pat->nslices, dinv)                  \
       : launch_pcg<R, false>(ctx, grid, P, A->d_val, Mval, src, bS, x, pglob, pat->nrows, pat->nslices, dinv))
#define PCG_GO2(R)                                                                                                             \
    (b ? launch_pcg2<R, true>(ctx, grid, P, A->d_val, Mval, src, bS, x, pat->nrows, pat->nslices, dinv)                        \
       : launch_pcg2<R, false>(ctx, grid, P, A->d_val, Mval, src, bS, x, pat->nrows, pat->nslices, dinv))
    if (ctx->pcg_variant == 0) {
        const char *e = getenv("TB_PCG_V");
        ctx->pcg_variant = e && atoi(e) == 1 ? 1 : 2;
    }
    int32_t st = ctx->pcg_variant == 1 ? (rpl <= 1 ? PCG_GO(1) : rpl <= 2 ? PCG_GO(2) : PCG_GO(4))
                            : (rpl <= 1 ? PCG_GO2(1) : rpl <= 2 ? PCG_GO2(2) : PCG_GO2(4));
#undef PCG_GO
#undef PCG_GO2
    if (st != TB_OK) return st;
    if (ctx->cg_deferred) {
        TB_LAUNCH(ctx, k_pcg_fold, 1, 1, 0, ctx->d_cg);
        ctx->last_cg_persistent = 1;
        if (iters) *iters = 0;
        if (rnorm) *rnorm = 0.0;
        if (converged) *converged = 1;
        return TB_OK;
    }
    if (ctx->profile) TB_CUDA(cudaEventRecord(ctx->prof_ev[1], ctx->stream));
    CGState *h = ctx->h_cg;
    TB_CUDA(cudaMemcpyAsync(h, ctx->d_cg, sizeof(CGState), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->last_cg_iters = h->iter;
    ctx->last_cg_persistent = 1;
    if (ctx->profile && h->iter > 0) {
        // the whole solve is one launch: report it per iteration (SpMV + vector updates + barriers)
        float ms = 0.f;
        TB_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[0], ctx->prof_ev[1]));
        ctx->prof_spmv_ms += ms;
        ctx->prof_spmv_n += h->iter;
    }
    if (iters) *iters = h->iter;
    if (rnorm) *rnorm = h->rnorm;
    if (converged) *converged = h->solved;
    return TB_OK;
}

extern "C" int32_t tb_cg_set_persistent(tb_ctx *ctx, int32_t mode) {
    TB_REQUIRE(ctx, "tb_cg_set_persistent: ctx is NULL");
    TB_REQUIRE(mode >= 0 && mode <= 2, "tb_cg_set_persistent: mode must be 0 (off), 1 (auto) or 2 (TMA kernel whenever eligible)");
    ctx->cg_persistent = mode;
    return TB_OK;
}

extern "C" int32_t tb_cg_set_persistent_variant(tb_ctx *ctx, int32_t variant) {
    TB_REQUIRE(ctx && (variant == 1 || variant == 2), "tb_cg_set_persistent_variant: 1 (three grid.sync per iteration) or 2 (two flag barriers)");
    ctx->pcg_variant = variant;
    return TB_OK;
}

extern "C" int32_t tb_cg_last_path(tb_ctx *ctx, int32_t *persistent) {
    TB_REQUIRE(ctx && persistent, "tb_cg_last_path: NULL argument");
    *persistent = ctx->last_cg_persistent;
    return TB_OK;
}
