// Pointwise ionic-model sweep: the method of _pointwise_step_outer_kernel!
// (src/solver/time/partitioned_solver.jl:38-52) that ext/CuThunderboltExt.jl:103-124 provided for
// CuVector, rebuilt for sm_100a.
//
// Data movement: state-blocked SoA, every state column padded to a multiple of 32 doubles, so each
// thread moves TWO neighbouring nodes per state with one 128-bit load and one 128-bit store
// (a warp touches 512 contiguous bytes per state).  The reference writes du to memory and re-reads it
// (4*nstates*8 B/node); here the state lives in registers across all sub-steps, so the sweep costs
// the algorithmic minimum 2*nstates*8 B/node: FHN 32 B, PCG2019 112 B.
#include "tb_internal.cuh"
#include "tb_cells.cuh"

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

template <int MODEL, bool ADAPTIVE, bool MAXD>
__global__ void __launch_bounds__(256)
    k_cell_step(double *__restrict__ u, int64_t ld, int64_t n, const double *__restrict__ phi_src,
                const tb_cell_params prm, double t, double dt, int substeps, double thr, double *partials,
                unsigned *ticket, double *result) {
    constexpr int NS = tb_cell_traits<MODEL>::NS;
    constexpr int PHI = tb_cell_traits<MODEL>::PHI;
    __shared__ double sm[32];
    const int64_t npairs = (n + 1) >> 1;
    double dmax = -INFINITY;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = q << 1;
        const bool two = i + 1 < n;
        double ua[NS], ub[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const double2 v = *reinterpret_cast<const double2 *>(u + (int64_t)s * ld + i);
            ua[s] = v.x;
            ub[s] = v.y;
        }
        if (phi_src) {   // fused "phi <- x" of the backward-Euler step (euler.jl:94 copies x into the u view)
            const double2 v = *reinterpret_cast<const double2 *>(phi_src + i);
            ua[PHI] = v.x;
            ub[PHI] = v.y;
        }
        const double da = tb_cell_node_step<MODEL, ADAPTIVE>(prm, ua, t, dt, substeps, thr);
        if (MAXD) dmax = fmax(dmax, da);
        if (two) {
            const double db = tb_cell_node_step<MODEL, ADAPTIVE>(prm, ub, t, dt, substeps, thr);
            if (MAXD) dmax = fmax(dmax, db);
#pragma unroll
            for (int s = 0; s < NS; s++) *reinterpret_cast<double2 *>(u + (int64_t)s * ld + i) = make_double2(ua[s], ub[s]);
        } else {
#pragma unroll
            for (int s = 0; s < NS; s++) u[(int64_t)s * ld + i] = ua[s];
        }
    }
    if (MAXD) {
        double b = tb_block_max(dmax, sm);
        __shared__ int s_last;
        if (threadIdx.x == 0) {
            partials[blockIdx.x] = b;
            __threadfence();
            unsigned tk = atomicInc(ticket, gridDim.x - 1);
            s_last = (tk == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double m = -INFINITY;
            for (unsigned k = threadIdx.x; k < gridDim.x; k += blockDim.x) m = fmax(m, ((volatile double *)partials)[k]);
            m = tb_block_max(m, sm);
            if (threadIdx.x == 0) *result = m;
        }
    }
}

template <int MODEL>
static int32_t launch_cell(tb_ctx *ctx, const tb_cell_params &prm, double *u, int64_t n, int64_t ld, const double *phi_src,
                           double t, double dt, int substeps, double thr, bool want_max) {
    const bool adaptive = substeps > 1;
    const int64_t need = ((n + 1) / 2 + 255) / 256;
    double *res = ctx->d_scalar;
    double *part = ctx->d_partials + 2 * TB_MAX_PARTIALS;
    unsigned *tick = ctx->d_ticket + 2;
#define TB_CELL_LAUNCH(AD, MX)                                                                                        \
    TB_LAUNCH(ctx, (k_cell_step<MODEL, AD, MX>), TB_GRID(ctx, (k_cell_step<MODEL, AD, MX>), 256, 0, need), 256, 0, u, ld, n, \
              phi_src, prm, t, dt, substeps, thr, part, tick, res)
    if (adaptive) {
        if (want_max) TB_CELL_LAUNCH(true, true); else TB_CELL_LAUNCH(true, false);
    } else {
        if (want_max) TB_CELL_LAUNCH(false, true); else TB_CELL_LAUNCH(false, false);
    }
#undef TB_CELL_LAUNCH
    return TB_OK;
}

int32_t tb_cell_step_raw(tb_ctx *ctx, int model, const double *params, int nparams, double *u, int64_t n, int64_t ld,
                         int phi_idx, const double *phi_src, double t, double dt, int substeps, double thr,
                         double *max_dphi) {
    TB_REQUIRE(tb_model_known(model), "tb_cell_step: unknown ionic model %d", model);
    const int np = tb_model_nparams(model);
    TB_REQUIRE(params && nparams == np, "tb_cell_step: model %d takes %d parameters, got %d", model, np, nparams);
    TB_REQUIRE(phi_idx == tb_model_phi(model), "tb_cell_step: model %d keeps the transmembrane potential in state %d, not %d", model,
               tb_model_phi(model), phi_idx);
    TB_REQUIRE((ld & 1) == 0, "tb_cell_step: column stride must be even");
    if (n == 0) {   // a rank that owns no points contributes the identity of the max-reduction
        if (max_dphi) *max_dphi = -INFINITY;
        return TB_OK;
    }
    tb_cell_params prm;
    for (int i = 0; i < 36; i++) prm.p[i] = i < np ? params[i] : 0.0;
    if (model == TB_FHN) TB_TRY((launch_cell<0>(ctx, prm, u, n, ld, phi_src, t, dt, substeps, thr, max_dphi != nullptr)));
    else if (model == TB_ALIEV_PANFILOV) TB_TRY((launch_cell<2>(ctx, prm, u, n, ld, phi_src, t, dt, substeps, thr, max_dphi != nullptr)));
    else TB_TRY((launch_cell<1>(ctx, prm, u, n, ld, phi_src, t, dt, substeps, thr, max_dphi != nullptr)));
    if (max_dphi) {
        TB_CUDA(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        *max_dphi = ctx->h_scalar[0];
    }
    return TB_OK;
}

extern "C" int32_t tb_cell_step(tb_ctx *ctx, int32_t model, const double *params, int32_t nparams, tb_vec *u,
                                int32_t phi_idx, double t, double dt, int32_t substeps, double reaction_threshold,
                                double *max_dphi) {
    TB_REQUIRE(ctx && u, "tb_cell_step: NULL argument");
    TB_REQUIRE(tb_model_known(model), "tb_cell_step: unknown ionic model %d", model);
    const int ns = tb_model_nstates(model);
    TB_REQUIRE(u->ncols == ns, "tb_cell_step: state vector has %d columns, model needs %d", u->ncols, ns);
    TB_DEV(ctx);
    return tb_cell_step_raw(ctx, model, params, nparams, u->d, u->n, u->ld, phi_idx, nullptr, t, dt, substeps,
                            reaction_threshold, max_dphi);
}
