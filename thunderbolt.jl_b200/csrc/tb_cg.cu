// Conjugate gradients on the SELL-32 operator, all scalars device-resident.
// Replaces LinearSolve.solve!(cache) with KrylovJL_CG (src/solver/time/euler.jl:10,94,155-156), i.e.
// Krylov.jl 0.10.9 cg! with LinearSolve 3.87.0 defaults: x0 = 0, r = b, p = r, gamma = r.r,
// eps = atol + rtol*sqrt(gamma); loop { Ap; alpha = gamma/p.Ap; x += alpha p; r -= alpha Ap;
// gamma' = r.r; solved = sqrt(gamma') <= eps; if !solved { beta = gamma'/gamma; p = r + beta p };
// iter++ } until solved or iter >= itmax.
//
// One iteration = three kernels, no host round trip:
//   k_cg_spmv_dot : Ap = A p fused with the p.Ap reduction (warp shuffles -> block -> last-block)
//   k_cg_xr       : x += alpha p, r -= alpha Ap fused with the r.r reduction          (48 B/row)
//   k_cg_p        : p = r + beta p                                                   (24 B/row)
// alpha, beta, gamma, |r|, iter and the `done` flag live in a CGState in HBM; the last block of each
// reducing kernel (ticket counter) finishes the reduction in a fixed order -- deterministic run to
// run -- and advances the scalars.  Kernels launched after convergence see done = 1 and return at
// once, so the host enqueues iterations ahead and polls the flag only every few iterations.
// Algorithmic bytes per iteration and row: SpMV (nnzr*12 + 24) + 72.
#include "tb_internal.cuh"
#include "tb_spmv.cuh"

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

// ---- scalar recurrences (run by one thread) ------------------------------------------------------
__device__ __forceinline__ void cg_after_init(CGState *st, double gamma) {
    st->gamma = gamma;
    const double rn = sqrt(gamma);
    st->rnorm = rn;
    st->eps = st->atol + st->rtol * rn;
    st->solved = rn <= st->eps;
    st->iter = 0;
    st->done = st->solved || (st->iter >= st->itmax);
}
// A NaN scalar or a non-positive curvature can never satisfy the stopping rule; without an exit the host would keep
// enqueueing iterations up to itmax (= nrows by default).  Same exit as the persistent kernels: tired, not solved.
__device__ __forceinline__ void cg_give_up(CGState *st) {
    st->solved = 0;
    st->iter = st->itmax;
    st->done = 1;
}
__device__ __forceinline__ void cg_after_pAp(CGState *st, double pAp) {
    st->pAp = pAp;
    st->alpha = st->gamma / pAp;
    if (!(pAp > 0.0)) cg_give_up(st);
}
__device__ __forceinline__ void cg_after_rr(CGState *st, double gnext) {
    st->gamma_next = gnext;
    const double rn = sqrt(gnext);
    st->rnorm = rn;
    const int solved = rn <= st->eps;
    st->solved = solved;
    if (!solved) {
        st->beta = gnext / st->gamma;
        st->gamma = gnext;
    }
    st->iter += 1;
    st->done = solved || (st->iter >= st->itmax);
    if (gnext != gnext) cg_give_up(st);
}

// which: 0 init, 1 pAp, 2 rr.  Used on the multi-GPU path after the NCCL all-reduce of st->local[0].
// With a peer window (ar.wins != nullptr) this one thread IS the all-reduce: it waits for every rank's partial of this
// epoch and adds them in rank order, so all ranks advance with bitwise identical scalars.
__global__ void k_cg_scalar(CGState *st, int which, const tb_ar_args ar, bool exact) {
    if (which != 0 && st->done) return;
    // NCCL path in exact mode: local[0] / local[1] hold the all-reduced high and low words (summed separately)
    const double total = ar.wins ? (exact ? tb_ar_collect_acc<true>(ar) : tb_ar_collect(ar))
                                 : (exact ? st->local[0] + st->local[1] : st->local[0]);
    if (which == 0) cg_after_init(st, total);
    else if (which == 1) cg_after_pAp(st, total);
    else cg_after_rr(st, total);
}

__global__ void k_cg_set_tol(CGState *st, double atol, double rtol, long long itmax) {
    st->atol = atol;
    st->rtol = rtol;
    st->itmax = itmax;
    st->done = 0;
    st->solved = 0;
    st->iter = 0;
    st->beta = 0.0;   // the fused-p SpMV forms p = r + beta p_old: beta = 0 on the first iteration gives p0 = r0
}

// finishing step shared by the reducing kernels: single GPU -> advance the scalars here;
// multi GPU -> leave the rank-local sum for the all-reduce
template <int WHICH, bool X>
__device__ __forceinline__ void cg_finish(tb_acc<X> block_value, CGState *st, double *partials, unsigned *ticket, double *sm,
                                          bool dist, const tb_ar_args &ar) {
    tb_acc<X> total;
    if (tb_grid_sum_acc<X>(block_value, partials, ticket, sm, &total) && threadIdx.x == 0) {
        if (dist && ar.wins) tb_ar_publish_acc<X>(ar, total);
        else if (dist) {
            st->local[0] = total.hi;
            st->local[1] = total.low();
        } else if (WHICH == 0) cg_after_init(st, total.value());
        else if (WHICH == 1) cg_after_pAp(st, total.value());
        else cg_after_rr(st, total.value());
    }
}

// ---- init from a given right-hand side: x = 0, r = p = b, gamma = b.b ------------------------------
template <bool X>
__global__ void __launch_bounds__(256) k_cg_init_b(const double *__restrict__ b, double *__restrict__ x,
                                                   double *__restrict__ r, double *__restrict__ p, int64_t n, CGState *st,
                                                   double *partials, unsigned *ticket, bool dist, const tb_ar_args ar,
                                                   const double *__restrict__ dinv) {
    __shared__ double sm[64];
    tb_acc<X> acc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = b[i];
        const double z = dinv ? dinv[i] * v : v;      // z = M r (Jacobi), p = z, gamma = r.z
        x[i] = 0.0;
        r[i] = v;
        p[i] = z;
        acc.add_prod(v, z);
    }
    cg_finish<0, X>(tb_block_sum_acc<X>(acc, sm), st, partials, ticket, sm, dist, ar);
}

// ---- init fused with the backward-Euler right-hand side: r = p = M*phi (+ bS), x = 0 ----------------
// ("b = M u_{n-1}" + add!(b, source), src/solver/time/euler.jl:85-91)
template <bool X>
__global__ void __launch_bounds__(256)
    k_cg_init_Mphi(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const double *__restrict__ Mval,
                   const double *__restrict__ phi, const double *__restrict__ bS, double *__restrict__ x,
                   double *__restrict__ r, double *__restrict__ p, int64_t nrows, int64_t nslices, CGState *st,
                   double *partials, unsigned *ticket, bool dist, const tb_ar_args ar, const double *__restrict__ dinv) {
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    tb_acc<X> acc;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        double v = tb_sell_row(slice_ptr, col, Mval, phi, s, lane);
        const int64_t row = s * TB_SLICE + lane;
        if (row < nrows) {
            if (bS) v += bS[row];
            const double z = dinv ? dinv[row] * v : v;
            x[row] = 0.0;
            r[row] = v;
            p[row] = z;
            acc.add_prod(v, z);
        }
    }
    cg_finish<0, X>(tb_block_sum_acc<X>(acc, sm), st, partials, ticket, sm, dist, ar);
}

// ---- Ap = A p, p.Ap -------------------------------------------------------------------------------
template <bool X>
__global__ void __launch_bounds__(256)
    k_cg_spmv_dot(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const double *__restrict__ val,
                  const double *__restrict__ p, double *__restrict__ Ap, int64_t nrows, int64_t nslices, CGState *st,
                  double *partials, unsigned *ticket, bool dist, const tb_ar_args ar, const tb_hwait_args hw) {
    if (st->done) return;
    tb_halo_wait(hw);
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    tb_acc<X> acc;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        const double v = tb_sell_row(slice_ptr, col, val, p, s, lane);
        const int64_t row = s * TB_SLICE + lane;
        if (row < nrows) {
            Ap[row] = v;
            acc.add_prod(p[row], v);
        }
    }
    if (partials) cg_finish<1, X>(tb_block_sum_acc<X>(acc, sm), st, partials, ticket, sm, dist, ar);   // nullptr: plain y = A x
}

// ---- bulk-async (TMA) variants of the two SpMV-shaped kernels ------------------------------------------
// Same arithmetic and summation order as above; the matrix stream comes through cp.async.bulk + mbarrier
// (tb_spmv.cuh).  INIT = true: r = p = M*phi (+bS), x = 0, gamma; false: Ap = A p, p.Ap.
template <int STAGES, bool INIT, bool CC, bool X>
__global__ void __launch_bounds__(1024, 1)
    k_cg_spmv_tma(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const int *__restrict__ cstream,
                  const int64_t *__restrict__ cptr,
                  const double *__restrict__ val, const double *__restrict__ xin, const double *__restrict__ bS,
                  double *__restrict__ xout, double *__restrict__ r, double *__restrict__ pout, int64_t nrows,
                  int64_t nslices, unsigned val_bytes, unsigned col_bytes, CGState *st, double *partials, unsigned *ticket,
                  bool dist, const tb_ar_args ar, const tb_hwait_args hw, const double *__restrict__ dinv, const tb_wide_list wide) {
    if (!INIT && st->done) return;
    tb_halo_wait(hw);
    extern __shared__ __align__(128) unsigned char tb_dyn_smem[];
    __shared__ double sm[64];
    tb_acc<X> acc;
    auto epi = [&](int64_t row, double v) {
        if (row < nrows) {
            if (INIT) {
                if (bS) v += bS[row];
                const double z = dinv ? dinv[row] * v : v;
                xout[row] = 0.0;
                r[row] = v;
                pout[row] = z;
                acc.add_prod(v, z);
            } else {
                r[row] = v;              // r aliases Ap here
                acc.add_prod(xin[row], v);
            }
        }
    };
    tb_sell_sweep_tma<STAGES, CC>(slice_ptr, val, cstream, cptr, xin, nslices, val_bytes, col_bytes, tb_dyn_smem, epi, nullptr, col, wide);
    if (!partials) return;   // plain y = A x (inner sweeps of the Chebyshev preconditioner): no reduction, scalars untouched
    const tb_acc<X> bs = tb_block_sum_acc<X>(acc, sm);
    if (INIT) cg_finish<0, X>(bs, st, partials, ticket, sm, dist, ar);
    else cg_finish<1, X>(bs, st, partials, ticket, sm, dist, ar);
}

// ---- experiment: the p update inside the SpMV's gather (TB_SPMV_FUSEP=1; single GPU, unpreconditioned) ----------------
// Ap = A (r + beta p_old), p_new = r + beta p_old for the own rows, p_new.Ap.  One launch less and 16 B/row less vector traffic
// per iteration, paid with a second gather per matrix entry.  Same bits as k_cg_spmv_tma + k_cg_p.
template <bool CC>
__global__ void __launch_bounds__(1024, 1)
    k_cg_spmv_fp(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const int *__restrict__ cstream,
                 const int64_t *__restrict__ cptr, const double *__restrict__ val, const double *__restrict__ r,
                 const double *__restrict__ p_old, double *__restrict__ p_new, double *__restrict__ Ap, int64_t nrows, int64_t nslices,
                 unsigned val_bytes, unsigned col_bytes, CGState *st, double *partials, unsigned *ticket, const tb_wide_list wide) {
    if (st->done) return;
    extern __shared__ __align__(128) unsigned char tb_dyn_smem[];
    __shared__ double sm[64];
    const double beta = st->beta;
    tb_acc<false> acc;
    auto epi = [&](int64_t row, double v) {
        if (row < nrows) {
            const double pn = r[row] + beta * p_old[row];
            p_new[row] = pn;
            Ap[row] = v;
            acc.add_prod(pn, v);
        }
    };
    tb_sell_sweep_tma<1, CC, decltype(epi), true>(slice_ptr, val, cstream, cptr, r, nslices, val_bytes, col_bytes, tb_dyn_smem, epi, nullptr,
                                                  col, wide, p_old, beta);
    tb_ar_args ar;
    ar.wins = nullptr;
    cg_finish<1, false>(tb_block_sum_acc<false>(acc, sm), st, partials, ticket, sm, false, ar);
}

static int32_t launch_spmv_fp(tb_ctx *ctx, const tb_pattern *pat, const double *val, const double *r, const double *p_old, double *p_new,
                              double *Ap, CGState *st, double *part, unsigned *tick) {
    const bool cc = ctx->spmv_compress && pat->d_ccol != nullptr;
    const tb_tma_geom g = tb_tma_geometry(pat->max_width_tma, cc ? tb_ccol_stage_ints(pat) : 32 * pat->max_width_tma, 1, 0);
    const int64_t need = (pat->nslices + g.warps - 1) / g.warps;
    const int grid = (int)(need < ctx->sm_count ? need : ctx->sm_count);
    tb_wide_list wide;
    wide.slices = pat->d_wide_slices;
    wide.n = (int)pat->n_wide;
    if (cc) {
        TB_CUDA(tb_ensure_smem(ctx, (const void *)k_cg_spmv_fp<true>, g.smem));
        TB_LAUNCH(ctx, k_cg_spmv_fp<true>, grid, g.warps * 32, g.smem, pat->d_slice_ptr, pat->d_col, pat->d_ccol, pat->d_cptr, val, r, p_old,
                  p_new, Ap, pat->nrows, pat->nslices, g.val_bytes, g.col_bytes, st, part, tick, wide);
    } else {
        TB_CUDA(tb_ensure_smem(ctx, (const void *)k_cg_spmv_fp<false>, g.smem));
        TB_LAUNCH(ctx, k_cg_spmv_fp<false>, grid, g.warps * 32, g.smem, pat->d_slice_ptr, pat->d_col, pat->d_col, nullptr, val, r, p_old,
                  p_new, Ap, pat->nrows, pat->nslices, g.val_bytes, g.col_bytes, st, part, tick, wide);
    }
    return TB_OK;
}

template <int STAGES, bool INIT>
static int32_t launch_spmv_tma(tb_ctx *ctx, int warps_override, const tb_pattern *pat, const double *val, const double *xin,
                               const double *bS, double *xout, double *r, double *pout, CGState *st, double *part,
                               unsigned *tick, bool dist, const tb_ar_args &ar, const tb_hwait_args &hw, const double *dinv) {
    const bool cc = ctx->spmv_compress && pat->d_ccol != nullptr;
    const tb_tma_geom g = tb_tma_geometry(pat->max_width_tma, cc ? tb_ccol_stage_ints(pat) : 32 * pat->max_width_tma, STAGES, warps_override);
    const int64_t need = (pat->nslices + g.warps - 1) / g.warps;
    const int grid = (int)(need < ctx->sm_count ? need : ctx->sm_count);   // one CTA per SM, one balanced wave
    tb_wide_list wide;
    wide.slices = pat->d_wide_slices;
    wide.n = (int)pat->n_wide;
#define TB_SPMV_TMA_LAUNCH(CCV, XV, colstream, cptrv)                                                                          \
    do {                                                                                                                      \
        TB_CUDA(tb_ensure_smem(ctx, (const void *)k_cg_spmv_tma<STAGES, INIT, CCV, XV>, g.smem));                             \
        TB_LAUNCH(ctx, (k_cg_spmv_tma<STAGES, INIT, CCV, XV>), grid, g.warps * 32, g.smem, pat->d_slice_ptr, pat->d_col, colstream, \
                  cptrv, val, xin, bS, xout, r, pout, pat->nrows, pat->nslices, g.val_bytes, g.col_bytes, st, part, tick, dist, ar, \
                  hw, dinv, wide);                                                                                            \
    } while (0)
    if (cc) {
        if (ctx->exact_dot) TB_SPMV_TMA_LAUNCH(true, true, pat->d_ccol, pat->d_cptr);
        else TB_SPMV_TMA_LAUNCH(true, false, pat->d_ccol, pat->d_cptr);
    } else {
        if (ctx->exact_dot) TB_SPMV_TMA_LAUNCH(false, true, pat->d_col, nullptr);
        else TB_SPMV_TMA_LAUNCH(false, false, pat->d_col, nullptr);
    }
#undef TB_SPMV_TMA_LAUNCH
    return TB_OK;
}

// variant table (ctx->spmv_variant, env TB_SPMV_VARIANT), chosen by measurement (profiles/):
//   0 LDG kernel | 1 staged, 1 stage, as many warps as fit (default) | 2 staged, 1 stage, 16 warps
//   3 staged, 2 stages, as many warps as fit | 4 staged, 1 stage, 24 warps
template <bool INIT>
static int32_t dispatch_spmv_tma(tb_ctx *ctx, const tb_pattern *pat, const double *val, const double *xin, const double *bS,
                                 double *xout, double *r, double *pout, CGState *st, double *part, unsigned *tick, bool dist,
                                 const tb_ar_args &ar, const tb_hwait_args &hw, const double *dinv) {
    switch (ctx->spmv_variant) {
    case 2: return launch_spmv_tma<1, INIT>(ctx, 16, pat, val, xin, bS, xout, r, pout, st, part, tick, dist, ar, hw, dinv);
    case 3: return launch_spmv_tma<2, INIT>(ctx, 0, pat, val, xin, bS, xout, r, pout, st, part, tick, dist, ar, hw, dinv);
    case 4: return launch_spmv_tma<1, INIT>(ctx, 24, pat, val, xin, bS, xout, r, pout, st, part, tick, dist, ar, hw, dinv);
    default: return launch_spmv_tma<1, INIT>(ctx, 0, pat, val, xin, bS, xout, r, pout, st, part, tick, dist, ar, hw, dinv);
    }
}

// ---- x += alpha p; r -= alpha Ap; r.r  (128-bit loads/stores) --------------------------------------
template <bool X>
__global__ void __launch_bounds__(256) k_cg_xr(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                                               const double *__restrict__ Ap, int64_t n, CGState *st, double *partials,
                                               unsigned *ticket, bool dist, const tb_ar_args ar,
                                               const double *__restrict__ dinv) {
    // partials == nullptr: update only (general preconditioners form r.z in k_cg_dot after z = P r)
    if (st->done) return;
    __shared__ double sm[64];
    const double alpha = st->alpha;
    const int64_t n2 = n >> 1;
    tb_acc<X> acc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        double2 xv = reinterpret_cast<double2 *>(x)[i], rv = reinterpret_cast<double2 *>(r)[i];
        const double2 pv = reinterpret_cast<const double2 *>(p)[i], av = reinterpret_cast<const double2 *>(Ap)[i];
        xv.x += alpha * pv.x;
        xv.y += alpha * pv.y;
        rv.x -= alpha * av.x;
        rv.y -= alpha * av.y;
        reinterpret_cast<double2 *>(x)[i] = xv;
        reinterpret_cast<double2 *>(r)[i] = rv;
        if (dinv) {   // gamma' = r.z, z = M r
            const double2 dv = reinterpret_cast<const double2 *>(dinv)[i];
            acc.add_prod(rv.x, dv.x * rv.x);
            acc.add_prod(rv.y, dv.y * rv.y);
        } else {
            acc.add_prod(rv.x, rv.x);
            acc.add_prod(rv.y, rv.y);
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        const double xn = x[i] + alpha * p[i], rn = r[i] - alpha * Ap[i];
        x[i] = xn;
        r[i] = rn;
        acc.add_prod(rn, dinv ? dinv[i] * rn : rn);
    }
    if (partials) cg_finish<2, X>(tb_block_sum_acc<X>(acc, sm), st, partials, ticket, sm, dist, ar);
}

// ---- gamma = a.b for an explicit preconditioned residual (WHICH 0: initial, 2: per iteration) --------------------------
template <bool X>
__global__ void __launch_bounds__(256) k_cg_dot(const double *__restrict__ a, const double *__restrict__ b, int64_t n, CGState *st,
                                                double *partials, unsigned *ticket, bool dist, const tb_ar_args ar, int which) {
    if (st->done) return;
    __shared__ double sm[64];
    tb_acc<X> acc;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc.add_prod(a[i], b[i]);
    const tb_acc<X> bs = tb_block_sum_acc<X>(acc, sm);
    if (which == 0) cg_finish<0, X>(bs, st, partials, ticket, sm, dist, ar);
    else cg_finish<2, X>(bs, st, partials, ticket, sm, dist, ar);
}

// ---- p = r + beta p ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cg_p(const double *__restrict__ r, double *__restrict__ p, int64_t n,
                                              const CGState *st, const double *__restrict__ dinv) {
    if (st->done) return;
    const double beta = st->beta;
    const int64_t n2 = n >> 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        double2 rv = reinterpret_cast<const double2 *>(r)[i];
        if (dinv) {   // p = z + beta p
            const double2 dv = reinterpret_cast<const double2 *>(dinv)[i];
            rv.x = dv.x * rv.x;
            rv.y = dv.y * rv.y;
        }
        double2 pv = reinterpret_cast<double2 *>(p)[i];
        pv.x = rv.x + beta * pv.x;
        pv.y = rv.y + beta * pv.y;
        reinterpret_cast<double2 *>(p)[i] = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = (dinv ? dinv[n - 1] * r[n - 1] : r[n - 1]) + beta * p[n - 1];
}

// ---- fused multi-GPU variants (peer path): the all-reduce collects and the halo push live INSIDE the CG kernels --------
// k_cg_xr_fused : every CTA collects p.Ap from the window (all ranks' partials, rank order), forms alpha itself, updates
//                 x and r, and the last CTA publishes this rank's r.z partial.
// k_cg_p_fused  : every CTA collects r.z, forms beta / the stopping decision itself, updates p and -- for the rows a
//                 neighbour needs -- stores the new value straight into that neighbour's ghost block; the last CTA
//                 raises the neighbours' halo flags.  CTA 0 writes the advanced scalars into the OTHER CGState (ping-pong:
//                 a CTA that starts late must still read the old gamma), which the next iteration's kernels then read.
// Three launches per iteration, like the single-GPU path; no NCCL call, no helper kernel.
template <bool X>
__global__ void __launch_bounds__(256) k_cg_xr_fused(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                                                     const double *__restrict__ Ap, int64_t n, const CGState *sin, double *partials,
                                                     unsigned *ticket, const tb_ar_args ar_in, const tb_ar_args ar_out,
                                                     const double *__restrict__ dinv) {
    if (sin->done) return;
    __shared__ double sm[64];
    __shared__ double s_alpha;
    if (threadIdx.x == 0) {
        const double pAp = tb_ar_collect_acc<X>(ar_in);
        s_alpha = pAp > 0.0 ? sin->gamma / pAp : NAN;   // NaN / non-positive curvature: leave x and r alone, poison r.z so that
    }                                                   // k_cg_p_fused gives up (same outcome as cg_after_pAp on one GPU)
    __syncthreads();
    const double alpha = s_alpha;
    const bool bad = alpha != alpha;
    const int64_t n2 = bad ? 0 : n >> 1;
    tb_acc<X> acc;
    if (bad) acc.set(NAN, 0.0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
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
            acc.add_prod(rv.x, dv.x * rv.x);
            acc.add_prod(rv.y, dv.y * rv.y);
        } else {
            acc.add_prod(rv.x, rv.x);
            acc.add_prod(rv.y, rv.y);
        }
    }
    if (!bad && (n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = n - 1;
        const double xn = x[i] + alpha * p[i], rn = r[i] - alpha * Ap[i];
        x[i] = xn;
        r[i] = rn;
        acc.add_prod(rn, dinv ? dinv[i] * rn : rn);
    }
    tb_acc<X> total;
    if (tb_grid_sum_acc<X>(tb_block_sum_acc<X>(acc, sm), partials, ticket, sm, &total) && threadIdx.x == 0)
        tb_ar_publish_acc<X>(ar_out, total);
}

__device__ __forceinline__ void tb_push_row(const tb_push_args &pa, int64_t row, double v) {
    for (int k = 0; k < pa.n; k++) {
        const long long o = row - pa.lo[k];
        if (o >= 0 && o < pa.len[k]) pa.dst[k][o] = v;
    }
}

template <bool X>
__global__ void __launch_bounds__(256) k_cg_p_fused(const double *__restrict__ r, double *__restrict__ p, int64_t n,
                                                    const CGState *sin, CGState *sout, const tb_ar_args ar_in,
                                                    const double *__restrict__ dinv, const tb_push_args pa, unsigned *ticket) {
    if (sin->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *sout = *sin;
        return;
    }
    __shared__ double s_beta, s_gnext;
    __shared__ int s_solved, s_last;
    if (threadIdx.x == 0) {
        const double gnext = tb_ar_collect_acc<X>(ar_in);
        s_gnext = gnext;
        s_solved = sqrt(gnext) <= sin->eps;
        s_beta = gnext / sin->gamma;
    }
    __syncthreads();
    const bool solved = s_solved != 0;
    if (!solved) {
        const double beta = s_beta;
        const int64_t n2 = n >> 1;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
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
            tb_push_row(pa, 2 * i, pv.x);
            tb_push_row(pa, 2 * i + 1, pv.y);
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            const double v = (dinv ? dinv[n - 1] * r[n - 1] : r[n - 1]) + beta * p[n - 1];
            p[n - 1] = v;
            tb_push_row(pa, n - 1, v);
        }
        // all pushes of this CTA are out before its ticket; the last CTA raises the neighbours' flags
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicInc(ticket, gridDim.x - 1);
            s_last = (t == gridDim.x - 1);
            if (s_last) {
                __threadfence_system();
                for (int k = 0; k < pa.n; k++) *(volatile unsigned long long *)pa.flag[k] = pa.epoch;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CGState s = *sin;
        const double gnext = s_gnext;
        s.gamma_next = gnext;
        s.rnorm = sqrt(gnext);
        s.solved = solved;
        if (!solved) {
            s.beta = gnext / s.gamma;
            s.gamma = gnext;
        }
        s.iter += 1;
        s.done = solved || (s.iter >= s.itmax);
        if (gnext != gnext) {   // NaN right-hand side or pAp <= 0 upstream (alpha = inf/NaN poisons r): tired, not solved
            s.solved = 0;
            s.iter = s.itmax;
            s.done = 1;
        }
        *sout = s;
    }
}

// ---- Jacobi preconditioner: dinv = 1 / diag(A) ------------------------------------------------------------------
__global__ void k_diag_slot(const int64_t *__restrict__ rowptr, const int64_t *__restrict__ slice_ptr,
                            const int *__restrict__ col, int64_t nrows, int *__restrict__ slot) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = slice_ptr[r >> 5] + (r & 31);
        const int len = (int)(rowptr[r + 1] - rowptr[r]);
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (col[base + (int64_t)mid * TB_SLICE] < (int)r) lo = mid + 1; else hi = mid;
        }
        slot[r] = (lo < len && col[base + (int64_t)lo * TB_SLICE] == (int)r) ? lo : -1;
    }
}
__global__ void k_dinv(const int64_t *__restrict__ slice_ptr, const int *__restrict__ slot, const double *__restrict__ val,
                       int64_t nrows, int64_t len, double *__restrict__ dinv) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < len; r += (int64_t)gridDim.x * blockDim.x) {
        double d = 0.0;
        if (r < nrows && slot[r] >= 0) d = 1.0 / val[slice_ptr[r >> 5] + (r & 31) + (int64_t)slot[r] * TB_SLICE];
        dinv[r] = d;
    }
}
static int32_t cg_build_dinv(tb_ctx *ctx, const tb_csr *A, const double **out) {
    tb_pattern *pat = A->pat;
    if (!pat->d_diag_slot) {
        TB_CUDA(cudaMalloc(&pat->d_diag_slot, sizeof(int) * (size_t)(pat->nrows + 1)));
        TB_LAUNCH(ctx, k_diag_slot, tb_grid_for(ctx, pat->nrows, 256, 8), 256, 0, pat->d_rowptr, pat->d_slice_ptr, pat->d_col,
                  pat->nrows, pat->d_diag_slot);
    }
    const int64_t len = tb_round_up(pat->nrows, 32);
    if (ctx->dinv_len < len) {
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_dinv);
        ctx->d_dinv = nullptr;
        ctx->dinv_len = 0;
        TB_CUDA(cudaMalloc(&ctx->d_dinv, sizeof(double) * (size_t)len));
        ctx->dinv_len = len;
    }
    TB_LAUNCH(ctx, k_dinv, tb_grid_for(ctx, len, 256, 8), 256, 0, pat->d_slice_ptr, pat->d_diag_slot, A->d_val, pat->nrows, len,
              ctx->d_dinv);
    *out = ctx->d_dinv;
    return TB_OK;
}

// next all-reduce epoch on the window path (all ranks call this in the same order), or the NCCL marker
static tb_ar_args cg_next_ar(tb_ctx *ctx, bool peer) {
    tb_ar_args a;
    a.wins = nullptr;
    a.rank = ctx->rank;
    a.nranks = ctx->nranks;
    a.slot = 0;
    a.epoch = 0;
    if (peer) {
        a.wins = ctx->peer.d_peer_win;
        a.epoch = ++ctx->peer.ar_epoch;
        a.slot = (int)(a.epoch % TB_AR_SLOTS);
    }
    return a;
}

// finish a distributed dot product: window path = one single-thread kernel that collects; NCCL path = all-reduce + scalar kernel
static int32_t cg_allreduce_then(tb_ctx *ctx, int which, const tb_ar_args &ar) {
    // NCCL path: in exact mode the high and low words are all-reduced separately (the high words' rounding error is not
    // recovered, so this path is deterministic but not order-independent; the peer-window path is both)
    if (!ar.wins)
        TB_NCCL(ncclAllReduce(ctx->d_cg->local, ctx->d_cg->local, ctx->exact_dot ? 2 : 1, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    TB_LAUNCH(ctx, k_cg_scalar, 1, 1, 0, ctx->d_cg, which, ar, ctx->exact_dot != 0);
    return TB_OK;
}

// b != NULL: solve A x = b.  b == NULL: right-hand side is M*phi (+ bS), built inside the init kernel.
int32_t tb_cg_run_impl(tb_ctx *ctx, const tb_csr *A, const double *b, const tb_csr *M, double *phi, const double *bS,
                       double *x, double atol, double rtol, int64_t itmax, int64_t *iters, double *rnorm,
                       int32_t *converged, int precond) {
    const tb_pattern *pat = A->pat;
    const int64_t n = pat->nrows;
    const double *dinv = nullptr;
    const bool gen_pc = precond == TB_PRECOND_BLOCK_JACOBI || precond == TB_PRECOND_CHEBYSHEV;   // z = P r is an explicit vector
    if (precond == TB_PRECOND_JACOBI || precond == TB_PRECOND_CHEBYSHEV) TB_TRY(cg_build_dinv(ctx, A, &dinv));
    const double *pc_dinv = dinv;                   // Chebyshev's D^-1
    if (gen_pc) dinv = nullptr;                     // the CG kernels themselves run the unpreconditioned forms on (r, z)
    double cheb_c1[64], cheb_c2[64], cheb_inv_theta = 0.0;
    const int cheb_d = ctx->cheb_degree;
    // update!(P, A) only when the operator's values changed since the setup was built (A = M - dt K is rebuilt only when dt moves)
    const bool pc_stale = !(ctx->pc_uid == A->uid && ctx->pc_version == A->version && ctx->pc_kind == precond);
    if (precond == TB_PRECOND_BLOCK_JACOBI && pc_stale) TB_TRY(tb_pc_bj_update(ctx, A));
    if (precond == TB_PRECOND_CHEBYSHEV) {
        if (pc_stale) TB_TRY(tb_pc_gershgorin(ctx, A, &ctx->pc_lmax));
        const double lmax = ctx->pc_lmax;
        TB_REQUIRE(lmax > 0.0, "Chebyshev preconditioner: operator has no positive diagonal");
        const double hi = lmax, lo = lmax / ctx->cheb_ratio;
        const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma1 = theta / delta;
        cheb_inv_theta = 1.0 / theta;
        double rho = 1.0 / sigma1;
        for (int k = 1; k < cheb_d; k++) {
            const double rho_new = 1.0 / (2.0 * sigma1 - rho);
            cheb_c1[k] = rho_new * rho;
            cheb_c2[k] = 2.0 * rho_new / delta;
            rho = rho_new;
        }
    }
    if (gen_pc) {
        ctx->pc_uid = A->uid;
        ctx->pc_version = A->version;
        ctx->pc_kind = precond;
        const int64_t ld = tb_round_up(pat->ncols, 32);
        if (ctx->pcwork_ld < ld) {
            TB_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->d_pcwork);
            ctx->d_pcwork = nullptr;
            ctx->pcwork_ld = 0;
            TB_CUDA(cudaMalloc(&ctx->d_pcwork, sizeof(double) * 4 * (size_t)ld));
            TB_CUDA(cudaMemsetAsync(ctx->d_pcwork, 0, sizeof(double) * 4 * (size_t)ld, ctx->stream));
            ctx->pcwork_ld = ld;
        }
    }
    int pgrid = 0;
    // exact-dot mode and the general preconditioners live in the multi-kernel path only
    const int pkind = (ctx->exact_dot || gen_pc) ? 0 : tb_cg_persistent_kind(ctx, pat, &pgrid);   // small / mid-size operator: one persistent cooperative kernel
    if (pkind == 1) return tb_cg_run_persistent(ctx, pgrid, A, b, M, phi, bS, x, atol, rtol, itmax, iters, rnorm, converged, dinv);
    if (pkind == 2) return tb_cg_run_persistent_tma(ctx, pgrid, A, b, M, phi, bS, x, atol, rtol, itmax, iters, rnorm, converged, dinv);
    ctx->last_cg_persistent = 0;
    TB_TRY(tb_ctx_ensure_cgwork(ctx, pat->ncols));
    double *r = ctx->d_cgwork, *p = r + ctx->cgwork_ld, *Ap = p + ctx->cgwork_ld;
    CGState *st = ctx->d_cg;
    const bool dist = ctx->has_comm && ctx->nranks > 1;
    // peer-memory path: dot products through the windows; halo of p by direct stores when the plan has peer targets
    const bool peer_ar = dist && ctx->peer.on;
    const bool peer_halo = peer_ar && pat->halo.nneigh > 0 && pat->halo.peer_ready;
    double *part = ctx->d_partials;
    unsigned *tick = ctx->d_ticket;
    const int64_t need_v = (n / 2 + 256) / 256, need_s = (pat->nslices + 7) / 8;
    const int grid_xr = TB_GRID(ctx, k_cg_xr<true>, 256, 0, need_v);
    const int grid_p = TB_GRID(ctx, k_cg_p, 256, 0, need_v);
    const int grid_ib = TB_GRID(ctx, k_cg_init_b<true>, 256, 0, (n + 255) / 256);
    const int grid_s = TB_GRID(ctx, k_cg_spmv_dot<true>, 256, 0, need_s);
    const int grid_im = TB_GRID(ctx, k_cg_init_Mphi<true>, 256, 0, need_s);
    const bool tma = ctx->spmv_variant > 0 && pat->max_width_tma > 0;   // slices above TB_TMA_WCAP take the LDG row kernel inside the sweep
    const tb_hwait_args nowait{nullptr, 0, 0, nullptr, nullptr};
    // fused peer path: collects and halo push inside k_cg_xr_fused / k_cg_p_fused, scalars ping-pong between st[0] and st[1]
    const bool fused = peer_ar && pat->halo.peer_ready && pat->halo.fused && !gen_pc;   // agreed by all ranks (tb_csr_set_halo_fused)
    const int grid_xrf = fused ? TB_GRID(ctx, k_cg_xr_fused<true>, 256, 0, need_v) : 0;
    const int grid_pf = fused ? TB_GRID(ctx, k_cg_p_fused<true>, 256, 0, need_v) : 0;
    const bool X = ctx->exact_dot != 0;
    double *pc_z = ctx->d_pcwork, *pc_d = pc_z + ctx->pcwork_ld, *pc_res = pc_d + ctx->pcwork_ld, *pc_w = pc_res + ctx->pcwork_ld;
    const int grid_dot = TB_GRID(ctx, k_cg_dot<true>, 256, 0, (n + 255) / 256);
    // z = P r (every kernel in here returns at once when the solve is already done)
    auto apply_pc = [&](const double *rr, double *zz, CGState *state) -> int32_t {
        if (precond == TB_PRECOND_BLOCK_JACOBI) return tb_pc_bj_apply(ctx, rr, zz, state);
        TB_TRY(tb_pc_cheb_first(ctx, rr, pc_dinv, pc_d, zz, pc_res, cheb_inv_theta, n, state));
        for (int k = 1; k < cheb_d; k++) {
            if (pat->halo.nneigh > 0) TB_TRY(tb_halo_exchange(ctx, pat, pc_d));   // the inner direction lives outside the peer-mapped vectors: NCCL
            const tb_ar_args ar0 = cg_next_ar(ctx, false);
            if (tma)
                TB_TRY(dispatch_spmv_tma<false>(ctx, pat, A->d_val, pc_d, nullptr, nullptr, pc_w, nullptr, state, nullptr, tick, false, ar0, nowait, nullptr));
            else
                TB_LAUNCH(ctx, k_cg_spmv_dot<false>, grid_s, 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val, pc_d, pc_w, n, pat->nslices,
                          state, nullptr, tick, false, ar0, nowait);
            TB_TRY(tb_pc_cheb_step(ctx, pc_w, pc_dinv, pc_d, zz, pc_res, cheb_c1[k], cheb_c2[k], n, state));
        }
        return TB_OK;
    };

    TB_LAUNCH(ctx, k_cg_set_tol, 1, 1, 0, st, atol, rtol, (long long)itmax);
    {
        const tb_ar_args ar = cg_next_ar(ctx, peer_ar);
        if (b) {
            if (X) TB_LAUNCH(ctx, k_cg_init_b<true>, grid_ib, 256, 0, b, x, r, p, n, st, part, tick, dist, ar, dinv);
            else TB_LAUNCH(ctx, k_cg_init_b<false>, grid_ib, 256, 0, b, x, r, p, n, st, part, tick, dist, ar, dinv);
        } else {
            if (M->pat->halo.nneigh > 0) TB_TRY(tb_halo_exchange(ctx, M->pat, phi));   // phi lives in the caller's vector: NCCL
            if (tma)
                TB_TRY(dispatch_spmv_tma<true>(ctx, pat, M->d_val, phi, bS, x, r, p, st, part, tick, dist, ar, nowait, dinv));
            else if (X)
                TB_LAUNCH(ctx, k_cg_init_Mphi<true>, grid_im, 256, 0, pat->d_slice_ptr, pat->d_col, M->d_val, phi, bS, x, r, p, n,
                          pat->nslices, st, part, tick, dist, ar, dinv);
            else
                TB_LAUNCH(ctx, k_cg_init_Mphi<false>, grid_im, 256, 0, pat->d_slice_ptr, pat->d_col, M->d_val, phi, bS, x, r, p, n,
                          pat->nslices, st, part, tick, dist, ar, dinv);
        }
        if (dist) TB_TRY(cg_allreduce_then(ctx, 0, ar));
        if (gen_pc) {   // z0 = P r0, gamma = r0.z0 (stopping rule on sqrt(r.z) like Krylov's PCG), p0 = z0
            TB_TRY(apply_pc(r, pc_z, st));
            const tb_ar_args arz = cg_next_ar(ctx, peer_ar);
            if (X) TB_LAUNCH(ctx, k_cg_dot<true>, grid_dot, 256, 0, r, pc_z, n, st, part, tick, dist, arz, 0);
            else TB_LAUNCH(ctx, k_cg_dot<false>, grid_dot, 256, 0, r, pc_z, n, st, part, tick, dist, arz, 0);
            if (dist) TB_TRY(cg_allreduce_then(ctx, 0, arz));
            TB_CUDA(cudaMemcpyAsync(p, pc_z, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }

    int64_t enq = 0;
    int64_t chunk = ctx->last_cg_iters + 1;
    if (chunk < 2) chunk = 2;
    CGState *h = ctx->h_cg;
    for (;;) {
        if (chunk > itmax - enq) chunk = itmax - enq;
        for (int64_t k = 0; k < chunk && fused; k++) {
            const int64_t it = enq + k;
            CGState *sin = st + (it & 1), *sout = st + ((it + 1) & 1);
            tb_hwait_args hw = nowait;
            if (it == 0) TB_TRY(tb_halo_push(ctx, pat, p, sin, &hw));   // p of iteration 0 comes from the init kernel
            else hw = ctx->peer.hw_next;                                 // ... later ones were pushed by k_cg_p_fused
            const bool prof = ctx->profile && it < TB_PROF_MAX;
            if (prof) TB_CUDA(cudaEventRecord(ctx->prof_ev[2 * it], ctx->stream));
            const tb_ar_args ar1 = cg_next_ar(ctx, true);
            if (tma)
                TB_TRY(dispatch_spmv_tma<false>(ctx, pat, A->d_val, p, nullptr, nullptr, Ap, nullptr, sin, part, tick, dist, ar1, hw, nullptr));
            else if (X)
                TB_LAUNCH(ctx, k_cg_spmv_dot<true>, grid_s, 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val, p, Ap, n,
                          pat->nslices, sin, part, tick, dist, ar1, hw);
            else
                TB_LAUNCH(ctx, k_cg_spmv_dot<false>, grid_s, 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val, p, Ap, n,
                          pat->nslices, sin, part, tick, dist, ar1, hw);
            if (prof) TB_CUDA(cudaEventRecord(ctx->prof_ev[2 * it + 1], ctx->stream));
            const tb_ar_args ar2 = cg_next_ar(ctx, true);
            if (X) TB_LAUNCH(ctx, k_cg_xr_fused<true>, grid_xrf, 256, 0, x, r, p, Ap, n, sin, part + TB_MAX_PARTIALS, tick + 1, ar1, ar2, dinv);
            else TB_LAUNCH(ctx, k_cg_xr_fused<false>, grid_xrf, 256, 0, x, r, p, Ap, n, sin, part + TB_MAX_PARTIALS, tick + 1, ar1, ar2, dinv);
            tb_push_args pa;
            TB_TRY(tb_halo_push_args(ctx, pat, &pa, &ctx->peer.hw_next));
            if (X) TB_LAUNCH(ctx, k_cg_p_fused<true>, grid_pf, 256, 0, r, p, n, sin, sout, ar2, dinv, pa, ctx->d_ticket + 6);
            else TB_LAUNCH(ctx, k_cg_p_fused<false>, grid_pf, 256, 0, r, p, n, sin, sout, ar2, dinv, pa, ctx->d_ticket + 6);
        }
        const bool fusep = ctx->spmv_fusep && !dist && !gen_pc && !dinv && !X && tma;
        for (int64_t k = 0; k < chunk && fusep; k++) {
            const int64_t it = enq + k;
            double *pp[2] = {p, ctx->d_cgwork + 3 * ctx->cgwork_ld};
            const bool prof = ctx->profile && it < TB_PROF_MAX;
            if (prof) TB_CUDA(cudaEventRecord(ctx->prof_ev[2 * it], ctx->stream));
            TB_TRY(launch_spmv_fp(ctx, pat, A->d_val, r, pp[it & 1], pp[(it + 1) & 1], Ap, st, part, tick));
            if (prof) TB_CUDA(cudaEventRecord(ctx->prof_ev[2 * it + 1], ctx->stream));
            const tb_ar_args ar2 = cg_next_ar(ctx, false);
            TB_LAUNCH(ctx, k_cg_xr<false>, grid_xr, 256, 0, x, r, pp[(it + 1) & 1], Ap, n, st, part + TB_MAX_PARTIALS, tick + 1, false, ar2, nullptr);
        }
        for (int64_t k = 0; k < chunk && !fused && !fusep; k++) {
            tb_hwait_args hw = nowait;
            if (peer_halo) TB_TRY(tb_halo_push(ctx, pat, p, st, &hw));
            else if (pat->halo.nneigh > 0) TB_TRY(tb_halo_exchange(ctx, pat, p));
            const bool prof = ctx->profile && enq + k < TB_PROF_MAX;
            if (prof) TB_CUDA(cudaEventRecord(ctx->prof_ev[2 * (enq + k)], ctx->stream));
            const tb_ar_args ar1 = cg_next_ar(ctx, peer_ar);
            if (tma)
                TB_TRY(dispatch_spmv_tma<false>(ctx, pat, A->d_val, p, nullptr, nullptr, Ap, nullptr, st, part, tick, dist, ar1, hw, nullptr));
            else if (X)
                TB_LAUNCH(ctx, k_cg_spmv_dot<true>, grid_s, 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val, p, Ap, n,
                          pat->nslices, st, part, tick, dist, ar1, hw);
            else
                TB_LAUNCH(ctx, k_cg_spmv_dot<false>, grid_s, 256, 0, pat->d_slice_ptr, pat->d_col, A->d_val, p, Ap, n,
                          pat->nslices, st, part, tick, dist, ar1, hw);
            if (prof) TB_CUDA(cudaEventRecord(ctx->prof_ev[2 * (enq + k) + 1], ctx->stream));
            if (dist) TB_TRY(cg_allreduce_then(ctx, 1, ar1));
            if (gen_pc) {   // x, r update; z = P r; gamma' = r.z; p = z + beta p
                const tb_ar_args arn = cg_next_ar(ctx, false);
                TB_LAUNCH(ctx, k_cg_xr<false>, grid_xr, 256, 0, x, r, p, Ap, n, st, nullptr, tick + 1, false, arn, nullptr);
                TB_TRY(apply_pc(r, pc_z, st));
                const tb_ar_args ar2 = cg_next_ar(ctx, peer_ar);
                if (X) TB_LAUNCH(ctx, k_cg_dot<true>, grid_dot, 256, 0, r, pc_z, n, st, part + TB_MAX_PARTIALS, tick + 1, dist, ar2, 2);
                else TB_LAUNCH(ctx, k_cg_dot<false>, grid_dot, 256, 0, r, pc_z, n, st, part + TB_MAX_PARTIALS, tick + 1, dist, ar2, 2);
                if (dist) TB_TRY(cg_allreduce_then(ctx, 2, ar2));
                TB_LAUNCH(ctx, k_cg_p, grid_p, 256, 0, pc_z, p, n, st, nullptr);
                continue;
            }
            const tb_ar_args ar2 = cg_next_ar(ctx, peer_ar);
            if (X) TB_LAUNCH(ctx, k_cg_xr<true>, grid_xr, 256, 0, x, r, p, Ap, n, st, part + TB_MAX_PARTIALS, tick + 1, dist, ar2, dinv);
            else TB_LAUNCH(ctx, k_cg_xr<false>, grid_xr, 256, 0, x, r, p, Ap, n, st, part + TB_MAX_PARTIALS, tick + 1, dist, ar2, dinv);
            if (dist) TB_TRY(cg_allreduce_then(ctx, 2, ar2));
            TB_LAUNCH(ctx, k_cg_p, grid_p, 256, 0, r, p, n, st, dinv);
        }
        enq += chunk;
        TB_CUDA(cudaMemcpyAsync(h, fused ? st + (enq & 1) : st, sizeof(CGState), cudaMemcpyDeviceToHost, ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h->done || enq >= itmax) break;
        chunk = 4;
    }
    ctx->last_cg_iters = h->iter;
    if (peer_ar) {
        unsigned long long err = 0;
        TB_CUDA(cudaMemcpy(&err, &ctx->peer.win->err, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) return tb_fail(TB_ERR_COMM, "CG: a wait on a peer rank's data timed out (peer dead or out of step)");
    }
    if (ctx->profile) {
        // only launches that did work: iteration k ran its SpMV iff k < iter (later ones saw done = 1)
        const int64_t nreal = h->iter < TB_PROF_MAX ? h->iter : TB_PROF_MAX;
        for (int64_t k = 0; k < nreal && k < enq; k++) {
            float ms = 0.f;
            TB_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[2 * k], ctx->prof_ev[2 * k + 1]));
            ctx->prof_spmv_ms += ms;
            ctx->prof_spmv_n += 1;
        }
    }
    if (iters) *iters = h->iter;
    if (rnorm) *rnorm = h->rnorm;
    if (converged) *converged = h->solved;
    return TB_OK;
}

extern "C" int32_t tb_cg_solve_pc(tb_ctx *ctx, const tb_csr *A, const tb_vec *b, int32_t bcol, tb_vec *x, int32_t xcol,
                                  int32_t precond, double atol, double rtol, int64_t itmax, int64_t *iters, double *rnorm,
                                  int32_t *converged);

extern "C" int32_t tb_cg_solve(tb_ctx *ctx, const tb_csr *A, const tb_vec *b, int32_t bcol, tb_vec *x, int32_t xcol,
                               double atol, double rtol, int64_t itmax, int64_t *iters, double *rnorm,
                               int32_t *converged) {
    return tb_cg_solve_pc(ctx, A, b, bcol, x, xcol, TB_PRECOND_NONE, atol, rtol, itmax, iters, rnorm, converged);
}

extern "C" int32_t tb_cg_solve_pc(tb_ctx *ctx, const tb_csr *A, const tb_vec *b, int32_t bcol, tb_vec *x, int32_t xcol,
                                  int32_t precond, double atol, double rtol, int64_t itmax, int64_t *iters, double *rnorm,
                                  int32_t *converged) {
    TB_REQUIRE(ctx && A && b && x, "tb_cg_solve: NULL argument");
    TB_REQUIRE(precond >= TB_PRECOND_NONE && precond <= TB_PRECOND_CHEBYSHEV, "tb_cg_solve: unknown preconditioner %d", precond);
    TB_REQUIRE(bcol >= 0 && bcol < b->ncols && xcol >= 0 && xcol < x->ncols, "tb_cg_solve: column out of range");
    TB_REQUIRE(b->n >= A->pat->nrows && x->n >= A->pat->nrows, "tb_cg_solve: vector shorter than the operator");
    TB_REQUIRE(itmax >= 0, "tb_cg_solve: itmax must be >= 0");
    TB_DEV(ctx);
    return tb_cg_run_impl(ctx, A, b->d + (size_t)bcol * b->ld, nullptr, nullptr, nullptr, x->d + (size_t)xcol * x->ld, atol,
                          rtol, itmax, iters, rnorm, converged, precond);
}
