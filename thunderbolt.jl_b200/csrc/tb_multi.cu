// Multi-subdomain reaction-diffusion splits:
//   * PointwiseMultiODEFunction: several ionic models on disjoint point sets, each block of the solution vector in
//     PointBlockedLayout (all states of a point consecutive) or StateBlockedLayout
//     (src/solver/time/partitioned_solver.jl:23-35,126-155; src/modeling/solution_variables.jl:41-68);
//   * the heat sub-problem's view u[heat_dofrange] with a scattered index set (src/discretization/fem.jl:472-521): gather /
//     scatter between the blocked state vector and the contiguous phi_m vector the CG works on;
//   * BilinearInterfaceDiffusionIntegrator: K_e[i,j] -= [[N_i]] D [[N_j]] dGamma over interface cells
//     (src/modeling/core/diffusion.jl:81-140).
#include "tb_internal.cuh"
#include "tb_cells.cuh"
#include <math.h>

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

// ---- index sets ----------------------------------------------------------------------------------------------------
struct tb_index {
    tb_ctx *ctx;
    int64_t n;
    int64_t *d;
};

extern "C" int32_t tb_index_create(tb_ctx *ctx, const int64_t *idx, int64_t n, int32_t index_base, tb_index **out) {
    TB_REQUIRE(ctx && out && (idx || n == 0) && n >= 0, "tb_index_create: bad argument");
    TB_DEV(ctx);
    *out = nullptr;
    tb_index *ix = new (std::nothrow) tb_index();
    if (!ix) return tb_fail(TB_ERR_NOMEM, "tb_index_create: host allocation failed");
    ix->ctx = ctx;
    ix->n = n;
    ix->d = nullptr;
    std::vector<int64_t> h((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        h[(size_t)i] = idx[i] - index_base;
        if (h[(size_t)i] < 0) {
            delete ix;
            return tb_fail(TB_ERR_INVALID, "tb_index_create: negative index at position %lld", (long long)i);
        }
    }
    cudaError_t e = cudaMalloc(&ix->d, sizeof(int64_t) * (size_t)(n + 1));
    if (e == cudaSuccess && n) e = cudaMemcpy(ix->d, h.data(), sizeof(int64_t) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(ix->d);
        delete ix;
        return tb_fail(TB_ERR_CUDA, "tb_index_create: %s", cudaGetErrorString(e));
    }
    *out = ix;
    return TB_OK;
}

extern "C" int32_t tb_index_destroy(tb_index *ix) {
    if (!ix) return TB_OK;
    cudaSetDevice(ix->ctx->device);
    cudaStreamSynchronize(ix->ctx->stream);
    cudaFree(ix->d);
    delete ix;
    return TB_OK;
}

__global__ void k_gather(double *__restrict__ dst, const double *__restrict__ src, const int64_t *__restrict__ idx, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void k_scatter(double *__restrict__ dst, const int64_t *__restrict__ idx, const double *__restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[idx[i]] = src[i];
}

// dst[i, dcol] = src[idx[i], scol]   (u_heat = view(u, heat_dofrange))
extern "C" int32_t tb_vec_gather(tb_vec *dst, int32_t dcol, const tb_vec *src, int32_t scol, const tb_index *ix) {
    TB_REQUIRE(dst && src && ix, "tb_vec_gather: NULL argument");
    TB_REQUIRE(dcol >= 0 && dcol < dst->ncols && scol >= 0 && scol < src->ncols && dst->n >= ix->n, "tb_vec_gather: shape mismatch");
    tb_ctx *ctx = dst->ctx;
    TB_DEV(ctx);
    if (ix->n == 0) return TB_OK;
    TB_LAUNCH(ctx, k_gather, tb_grid_for(ctx, ix->n, 256, 8), 256, 0, dst->d + (size_t)dcol * dst->ld, src->d + (size_t)scol * src->ld,
              ix->d, ix->n);
    return TB_OK;
}
// dst[idx[i], dcol] = src[i, scol]
extern "C" int32_t tb_vec_scatter(tb_vec *dst, int32_t dcol, const tb_index *ix, const tb_vec *src, int32_t scol) {
    TB_REQUIRE(dst && src && ix, "tb_vec_scatter: NULL argument");
    TB_REQUIRE(dcol >= 0 && dcol < dst->ncols && scol >= 0 && scol < src->ncols && src->n >= ix->n, "tb_vec_scatter: shape mismatch");
    tb_ctx *ctx = dst->ctx;
    TB_DEV(ctx);
    if (ix->n == 0) return TB_OK;
    TB_LAUNCH(ctx, k_scatter, tb_grid_for(ctx, ix->n, 256, 8), 256, 0, dst->d + (size_t)dcol * dst->ld, ix->d,
              src->d + (size_t)scol * src->ld, ix->n);
    return TB_OK;
}

// ---- blocked cell sweep ----------------------------------------------------------------------------------------------
// One point per thread; state s of point k at base[k*pstride + s*sstride]: PointBlockedLayout = (nstates, 1),
// StateBlockedLayout = (1, npoints).  A warp of a point-blocked block touches 32*nstates consecutive doubles; the
// per-state loads of a thread hit the lines its neighbours just brought in (L1), so DRAM traffic stays 2*nstates*8 B/point.
template <int MODEL, bool ADAPTIVE>
__global__ void __launch_bounds__(256)
    k_cell_step_block(double *__restrict__ base, int64_t npoints, int64_t pstride, int64_t sstride, const tb_cell_params prm, double t,
                      double dt, int substeps, double thr, double *partials, unsigned *ticket, double *result) {
    constexpr int NS = tb_cell_traits<MODEL>::NS;
    __shared__ double sm[32];
    double dmax = -INFINITY;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < npoints; k += (int64_t)gridDim.x * blockDim.x) {
        double u[NS];
        double *p = base + k * pstride;
#pragma unroll
        for (int s = 0; s < NS; s++) u[s] = p[s * sstride];
        dmax = fmax(dmax, tb_cell_node_step<MODEL, ADAPTIVE>(prm, u, t, dt, substeps, thr));
#pragma unroll
        for (int s = 0; s < NS; s++) p[s * sstride] = u[s];
    }
    if (result) {
        double b = tb_block_max(dmax, sm);
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
            double m = -INFINITY;
            for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) m = fmax(m, ((volatile double *)partials)[q]);
            m = tb_block_max(m, sm);
            if (threadIdx.x == 0) *result = fmax(*result, m);   // running maximum over the blocks of one call
        }
    }
}

template <int MODEL>
static int32_t launch_block(tb_ctx *ctx, const tb_cell_block &b, double *base, int64_t ps, int64_t ss, double t, double dt,
                            int substeps, double thr, double *res) {
    tb_cell_params prm;
    for (int i = 0; i < 36; i++) prm.p[i] = i < b.nparams ? b.params[i] : 0.0;
    const int64_t need = (b.npoints + 255) / 256;
    double *part = ctx->d_partials + 2 * TB_MAX_PARTIALS;
    unsigned *tick = ctx->d_ticket + 2;
    if (substeps > 1)
        TB_LAUNCH(ctx, (k_cell_step_block<MODEL, true>), TB_GRID(ctx, (k_cell_step_block<MODEL, true>), 256, 0, need), 256, 0, base,
                  b.npoints, ps, ss, prm, t, dt, substeps, thr, part, tick, res);
    else
        TB_LAUNCH(ctx, (k_cell_step_block<MODEL, false>), TB_GRID(ctx, (k_cell_step_block<MODEL, false>), 256, 0, need), 256, 0, base,
                  b.npoints, ps, ss, prm, t, dt, substeps, thr, part, tick, res);
    return TB_OK;
}

__global__ void k_set_scalar(double *p, double v) { *p = v; }

extern "C" int32_t tb_cell_step_blocks(tb_ctx *ctx, const tb_cell_block *blocks, int32_t nblocks, tb_vec *u, double t, double dt,
                                       int32_t substeps, double reaction_threshold, double *max_dphi) {
    TB_REQUIRE(ctx && blocks && u && nblocks >= 0, "tb_cell_step_blocks: bad argument");
    TB_REQUIRE(u->ncols == 1, "tb_cell_step_blocks: the blocked state vector is one flat column");
    TB_DEV(ctx);
    for (int32_t i = 0; i < nblocks; i++) {
        const tb_cell_block &b = blocks[i];
        TB_REQUIRE(tb_model_known(b.model), "tb_cell_step_blocks: block %d: unknown ionic model %d", i, b.model);
        TB_REQUIRE(b.nparams == tb_model_nparams(b.model), "tb_cell_step_blocks: block %d: model %d takes %d parameters, got %d", i,
                   b.model, tb_model_nparams(b.model), b.nparams);
        TB_REQUIRE(b.layout == TB_LAYOUT_STATE_BLOCKED || b.layout == TB_LAYOUT_POINT_BLOCKED, "tb_cell_step_blocks: block %d: unknown layout", i);
        TB_REQUIRE(b.offset >= 0 && b.npoints >= 0 && b.offset + b.npoints * tb_model_nstates(b.model) <= u->n,
                   "tb_cell_step_blocks: block %d reaches beyond the state vector", i);
    }
    double *res = nullptr;
    if (max_dphi) {
        res = ctx->d_scalar + 6;
        TB_LAUNCH(ctx, k_set_scalar, 1, 1, 0, res, -INFINITY);
    }
    for (int32_t i = 0; i < nblocks; i++) {
        const tb_cell_block &b = blocks[i];
        if (b.npoints == 0) continue;
        const int ns = tb_model_nstates(b.model);
        const int64_t ps = b.layout == TB_LAYOUT_POINT_BLOCKED ? ns : 1, ss = b.layout == TB_LAYOUT_POINT_BLOCKED ? 1 : b.npoints;
        double *base = u->d + b.offset;
        if (b.model == TB_FHN) TB_TRY((launch_block<0>(ctx, b, base, ps, ss, t, dt, substeps, reaction_threshold, res)));
        else if (b.model == TB_ALIEV_PANFILOV) TB_TRY((launch_block<2>(ctx, b, base, ps, ss, t, dt, substeps, reaction_threshold, res)));
        else TB_TRY((launch_block<1>(ctx, b, base, ps, ss, t, dt, substeps, reaction_threshold, res)));
    }
    if (max_dphi) {
        TB_CUDA(cudaMemcpyAsync(ctx->h_scalar + 6, res, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        *max_dphi = ctx->h_scalar[6];
    }
    return TB_OK;
}

// ---- interface diffusion -----------------------------------------------------------------------------------------------
// An interface cell couples two coincident facets ("here" / "there") of neighbouring subdomains whose nodes were
// duplicated (FerriteInterfaceElements' InterfaceCell; the test case is test/integration/test_electrophysiology.jl:124-195).
// Per cell: dofs = (here dofs, there dofs), jump of basis i = -N_i on the here side, +N_i on the there side,
// dGamma = the average of the two sides' detJ*w (getdetJdV_average), K_e[i,j] -= jump_i * D * jump_j * dGamma.
struct tb_facet_tables {
    int k, nq;              // nodes per side, quadrature points
    double N[16][4];        // N[q][a]
    double dN[16][4][2];    // dN[q][a][d], d < facet dimension
    double w[16];
};

static int facet_gauss(int order, double *p, double *w) {
    switch (order) {
    case 1: p[0] = 0.0; w[0] = 2.0; return 1;
    case 2: p[0] = -0.5773502691896257; p[1] = 0.5773502691896257; w[0] = w[1] = 1.0; return 2;
    case 3: p[0] = -0.7745966692414834; p[1] = 0.0; p[2] = 0.7745966692414834; w[0] = w[2] = 0.5555555555555556; w[1] = 0.8888888888888888; return 3;
    default: p[0] = -0.8611363115940526; p[1] = -0.3399810435848563; p[2] = 0.3399810435848563; p[3] = 0.8611363115940526;
             w[0] = w[3] = 0.3478548451374538; w[1] = w[2] = 0.6521451548625461; return 4;
    }
}

__global__ void k_interface_elements(const double *__restrict__ xh, const double *__restrict__ xt, int64_t nif, int sdim, double D,
                                     const tb_facet_tables T, double *__restrict__ EA) {
    const int k = T.k, fd = sdim - 1;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nif; c += (int64_t)gridDim.x * blockDim.x) {
        double Ke[64];
        for (int i = 0; i < 4 * k * k; i++) Ke[i] = 0.0;
        for (int q = 0; q < T.nq; q++) {
            double dO = 0.0;
            for (int side = 0; side < 2; side++) {
                const double *X = (side ? xt : xh) + c * k * sdim;
                double t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0};
                for (int a = 0; a < k; a++)
                    for (int d = 0; d < sdim; d++) {
                        t0[d] += X[a * sdim + d] * T.dN[q][a][0];
                        if (fd == 2) t1[d] += X[a * sdim + d] * T.dN[q][a][1];
                    }
                double dj;
                if (fd == 1) dj = sqrt(t0[0] * t0[0] + t0[1] * t0[1]);
                else {
                    const double n0 = t0[1] * t1[2] - t0[2] * t1[1], n1 = t0[2] * t1[0] - t0[0] * t1[2], n2 = t0[0] * t1[1] - t0[1] * t1[0];
                    dj = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
                }
                dO += dj * T.w[q];
            }
            dO = dO / 2.0;
            for (int i = 0; i < 2 * k; i++) {
                const double ji = i < k ? -T.N[q][i] : T.N[q][i - k];
                for (int j = 0; j < 2 * k; j++) {
                    const double jj = j < k ? -T.N[q][j] : T.N[q][j - k];
                    Ke[i * 2 * k + j] -= (ji * D * jj) * dO;
                }
            }
        }
        for (int i = 0; i < 4 * k * k; i++) EA[c * 4 * k * k + i] = Ke[i];
    }
}

// sequential scatter in interface-cell order: the reference's element loop, bit for bit (interfaces are lower-dimensional:
// their count is negligible next to the bulk cells)
__global__ void k_interface_scatter(const double *__restrict__ EA, const int64_t *__restrict__ dofs, int64_t nif, int nd,
                                    const int64_t *__restrict__ rowptr, const int64_t *__restrict__ slice_ptr, const int *__restrict__ col,
                                    double *__restrict__ val, int *fail) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int64_t c = 0; c < nif; c++)
        for (int i = 0; i < nd; i++) {
            const int64_t r = dofs[c * nd + i];
            const int64_t base = slice_ptr[r >> 5] + (r & 31);
            const int len = (int)(rowptr[r + 1] - rowptr[r]);
            for (int j = 0; j < nd; j++) {
                const int target = (int)dofs[c * nd + j];
                int lo = 0, hi = len;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (col[base + (int64_t)mid * TB_SLICE] < target) lo = mid + 1; else hi = mid;
                }
                if (lo < len && col[base + (int64_t)lo * TB_SLICE] == target) val[base + (int64_t)lo * TB_SLICE] += EA[(c * nd + i) * nd + j];
                else *fail = 1;
            }
        }
}

extern "C" int32_t tb_assemble_interface_diffusion(tb_ctx *ctx, int32_t facet_type, int32_t sdim, int64_t nif, const int64_t *dofs,
                                                   int32_t index_base, const double *coords_here, const double *coords_there,
                                                   int32_t qorder, double D, tb_csr *K) {
    TB_REQUIRE(ctx && K && nif >= 0 && (nif == 0 || (dofs && coords_here && coords_there)), "tb_assemble_interface_diffusion: bad argument");
    TB_REQUIRE((facet_type == TB_FACET_LINE2 && sdim == 2) || (facet_type == TB_FACET_QUAD4 && sdim == 3),
               "tb_assemble_interface_diffusion: supported interfaces are LINE2 facets in 2D and QUAD4 facets in 3D");
    TB_REQUIRE(qorder >= 1 && qorder <= 4, "tb_assemble_interface_diffusion: Gauss order 1..4");
    TB_DEV(ctx);
    const tb_pattern *pat = K->pat;
    K->version++;
    TB_CUDA(cudaMemsetAsync(K->d_val, 0, sizeof(double) * (size_t)pat->sell_len, ctx->stream));
    if (nif == 0) return TB_OK;
    tb_facet_tables T;
    memset(&T, 0, sizeof(T));
    double gp[4], gw[4];
    const int ng = facet_gauss(qorder, gp, gw);
    if (facet_type == TB_FACET_LINE2) {
        T.k = 2;
        T.nq = ng;
        for (int q = 0; q < ng; q++) {
            T.N[q][0] = 0.5 * (1.0 - gp[q]);
            T.N[q][1] = 0.5 * (1.0 + gp[q]);
            T.dN[q][0][0] = -0.5;
            T.dN[q][1][0] = 0.5;
            T.w[q] = gw[q];
        }
    } else {
        T.k = 4;
        T.nq = ng * ng;
        const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
        for (int b = 0; b < ng; b++)
            for (int a = 0; a < ng; a++) {   // first coordinate fastest, like the cell rules
                const int q = b * ng + a;
                for (int v = 0; v < 4; v++) {
                    T.N[q][v] = 0.25 * (1.0 + sx[v] * gp[a]) * (1.0 + sy[v] * gp[b]);
                    T.dN[q][v][0] = 0.25 * sx[v] * (1.0 + sy[v] * gp[b]);
                    T.dN[q][v][1] = 0.25 * (1.0 + sx[v] * gp[a]) * sy[v];
                }
                T.w[q] = gw[a] * gw[b];
            }
    }
    const int k = T.k, nd = 2 * k;
    std::vector<int64_t> hd((size_t)(nif * nd));
    for (int64_t i = 0; i < nif * nd; i++) {
        hd[(size_t)i] = dofs[i] - index_base;
        TB_REQUIRE(hd[(size_t)i] >= 0 && hd[(size_t)i] < pat->nrows, "tb_assemble_interface_diffusion: dof id out of range");
    }
    double *d_xh = nullptr, *d_xt = nullptr, *d_EA = nullptr;
    int64_t *d_dofs = nullptr;
    int *d_fail = (int *)(ctx->d_scalar + 4);
    const size_t xb = sizeof(double) * (size_t)(nif * k * sdim);
    TB_CUDA(cudaMalloc(&d_xh, xb));
    TB_CUDA(cudaMalloc(&d_xt, xb));
    TB_CUDA(cudaMalloc(&d_EA, sizeof(double) * (size_t)(nif * nd * nd)));
    TB_CUDA(cudaMalloc(&d_dofs, sizeof(int64_t) * (size_t)(nif * nd)));
    TB_CUDA(cudaMemcpyAsync(d_xh, coords_here, xb, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaMemcpyAsync(d_xt, coords_there, xb, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaMemcpyAsync(d_dofs, hd.data(), sizeof(int64_t) * (size_t)(nif * nd), cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
    TB_LAUNCH(ctx, k_interface_elements, tb_grid_for(ctx, nif, 128, 8), 128, 0, d_xh, d_xt, nif, sdim, D, T, d_EA);
    TB_LAUNCH(ctx, k_interface_scatter, 1, 32, 0, d_EA, d_dofs, nif, nd, pat->d_rowptr, pat->d_slice_ptr, pat->d_col, K->d_val, d_fail);
    int fail = 0;
    TB_CUDA(cudaMemcpyAsync(&fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_xh);
    cudaFree(d_xt);
    cudaFree(d_EA);
    cudaFree(d_dofs);
    if (fail) return tb_fail(TB_ERR_INVALID, "tb_assemble_interface_diffusion: the pattern has no entry for an interface dof pair");
    return TB_OK;
}

// ---- pieces of the lead-field / Poisson ECG reconstructions (src/modeling/electrophysiology/ecg.jl:166-619) -----------------
// out[c] = sum_i Z[i, c] * v[i] for every column c of Z: `-cache.Z * cache.κ∇φₘ_t` (ecg.jl:617-619) with the lead fields stored
// as the columns of one device vector.  One column per blockIdx.y, deterministic two-stage sums.
__global__ void __launch_bounds__(256) k_vec_dots(const double *__restrict__ Z, int64_t ld, const double *__restrict__ v, int64_t n,
                                                  double *partials) {
    __shared__ double sm[32];
    const double *z = Z + (int64_t)blockIdx.y * ld;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += z[i] * v[i];
    const double b = tb_block_sum(acc, sm);
    if (threadIdx.x == 0) partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = b;
}
__global__ void k_vec_dots_finish(const double *partials, int nb, double *out) {
    double s = 0.0;
    for (int i = 0; i < nb; i++) s += partials[(int64_t)blockIdx.x * nb + i];
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

extern "C" int32_t tb_vec_dots(tb_ctx *ctx, const tb_vec *Z, const tb_vec *v, int32_t vcol, double *out) {
    TB_REQUIRE(ctx && Z && v && out, "tb_vec_dots: NULL argument");
    TB_REQUIRE(vcol >= 0 && vcol < v->ncols && v->n >= Z->n, "tb_vec_dots: shape mismatch");
    TB_DEV(ctx);
    const int nb = 64;
    double *part = nullptr, *d_out = nullptr;
    TB_CUDA(cudaMalloc(&part, sizeof(double) * (size_t)(nb * Z->ncols)));
    TB_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)Z->ncols));
    TB_LAUNCH(ctx, k_vec_dots, dim3(nb, Z->ncols), 256, 0, Z->d, Z->ld, v->d + (size_t)vcol * v->ld, Z->n, part);
    TB_LAUNCH(ctx, k_vec_dots_finish, Z->ncols, 1, 0, part, nb, d_out);
    TB_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)Z->ncols, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(part);
    cudaFree(d_out);
    return TB_OK;
}

// d[r] = A[r, r]
__global__ void k_csr_diag(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, const double *__restrict__ val, int64_t nrows,
                           int64_t nslices, double *__restrict__ d) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        const int64_t base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5, row = s * TB_SLICE + lane;
        double dv = 0.0;
        for (int64_t j = 0; j < w; j++)
            if (col[base + j * TB_SLICE + lane] == (int)row) dv += val[base + j * TB_SLICE + lane];   // padding adds 0
        if (row < nrows) d[row] = dv;
    }
}
extern "C" int32_t tb_csr_diagonal(const tb_csr *A, tb_vec *d, int32_t col) {
    TB_REQUIRE(A && d && col >= 0 && col < d->ncols && d->n >= A->pat->nrows, "tb_csr_diagonal: bad argument");
    tb_ctx *ctx = A->pat->ctx;
    TB_DEV(ctx);
    const tb_pattern *p = A->pat;
    TB_LAUNCH(ctx, k_csr_diag, tb_grid_for(ctx, p->nslices * 32, 256, 8), 256, 0, p->d_slice_ptr, p->d_col, A->d_val, p->nrows, p->nslices,
              d->d + (size_t)col * d->ld);
    return TB_OK;
}

// Ferrite's apply_zero!(K, f, ch) on the matrix (ecg.jl:338): rows and columns of the constrained dofs are zeroed and their
// diagonal entries set to `diag_value` (Ferrite uses the mean of |diag(K)|, which the caller computes from tb_csr_diagonal)
__global__ void k_mark(const int64_t *__restrict__ idx, int64_t n, unsigned char *flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) flag[idx[i]] = 1;
}
__global__ void __launch_bounds__(256)
    k_apply_zero(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col, double *__restrict__ val, int64_t nrows, int64_t nslices,
                 const unsigned char *__restrict__ flag, double diag_value) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        const int64_t base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5, row = s * TB_SLICE + lane;
        if (row >= nrows) continue;
        const bool fr = flag[row] != 0;
        bool diag_done = false;
        for (int64_t j = 0; j < w; j++) {
            const int c = col[base + j * TB_SLICE + lane];
            if (fr || flag[c]) {
                double nv = 0.0;
                if (fr && c == (int)row && !diag_done) {   // the first stored (row, row) entry is the real diagonal; padding repeats the own column
                    nv = diag_value;
                    diag_done = true;
                }
                val[base + j * TB_SLICE + lane] = nv;
            } else if (c == (int)row) diag_done = true;
        }
    }
}
extern "C" int32_t tb_csr_apply_zero(tb_csr *A, const tb_index *constrained, double diag_value) {
    TB_REQUIRE(A && constrained, "tb_csr_apply_zero: NULL argument");
    tb_ctx *ctx = A->pat->ctx;
    TB_DEV(ctx);
    const tb_pattern *p = A->pat;
    A->version++;
    unsigned char *flag = nullptr;
    TB_CUDA(cudaMalloc(&flag, (size_t)p->ncols + 1));
    TB_CUDA(cudaMemsetAsync(flag, 0, (size_t)p->ncols + 1, ctx->stream));
    if (constrained->n) TB_LAUNCH(ctx, k_mark, tb_grid_for(ctx, constrained->n, 256, 4), 256, 0, constrained->d, constrained->n, flag);
    TB_LAUNCH(ctx, k_apply_zero, tb_grid_for(ctx, p->nslices * 32, 256, 8), 256, 0, p->d_slice_ptr, p->d_col, A->d_val, p->nrows, p->nslices,
              flag, diag_value);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(flag);
    return TB_OK;
}

__global__ void k_fill_at(double *__restrict__ v, const int64_t *__restrict__ idx, int64_t n, double value) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[idx[i]] = value;
}
// v[idx[i], col] = value   (the right-hand-side half of apply_zero!)
extern "C" int32_t tb_vec_fill_at(tb_vec *v, int32_t col, const tb_index *ix, double value) {
    TB_REQUIRE(v && ix && col >= 0 && col < v->ncols, "tb_vec_fill_at: bad argument");
    tb_ctx *ctx = v->ctx;
    TB_DEV(ctx);
    if (ix->n) TB_LAUNCH(ctx, k_fill_at, tb_grid_for(ctx, ix->n, 256, 4), 256, 0, v->d + (size_t)col * v->ld, ix->d, ix->n, value);
    return TB_OK;
}

// ---- output staging (store_timestep_field!, src/ferrite-addons/io.jl:18-93) ----------------------------------------------------
// A snapshot of one state column leaves the device on its own copy stream, ordered after the work already queued on the
// compute stream, into caller-provided PINNED host memory (tb_host_alloc); stepping continues while the copy and the
// writer (host) run.  tb_stage_wait blocks until every staged snapshot has landed.
extern "C" int32_t tb_host_alloc(int64_t bytes, void **out) {
    TB_REQUIRE(out && bytes > 0, "tb_host_alloc: bad argument");
    *out = nullptr;
    TB_CUDA(cudaMallocHost(out, (size_t)bytes));
    return TB_OK;
}
extern "C" int32_t tb_host_free(void *p) {
    if (p) TB_CUDA(cudaFreeHost(p));
    return TB_OK;
}
extern "C" int32_t tb_vec_stage_col(const tb_vec *v, int32_t col, double *pinned_host) {
    TB_REQUIRE(v && pinned_host && col >= 0 && col < v->ncols, "tb_vec_stage_col: bad argument");
    tb_ctx *ctx = v->ctx;
    TB_DEV(ctx);
    if (!ctx->stage_stream) {
        TB_CUDA(cudaStreamCreateWithFlags(&ctx->stage_stream, cudaStreamNonBlocking));
        TB_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev, cudaEventDisableTiming));
    }
    TB_CUDA(cudaEventRecord(ctx->stage_ev, ctx->stream));                 // the state as of the work queued so far
    TB_CUDA(cudaStreamWaitEvent(ctx->stage_stream, ctx->stage_ev, 0));
    TB_CUDA(cudaMemcpyAsync(pinned_host, v->d + (size_t)col * v->ld, sizeof(double) * (size_t)v->n, cudaMemcpyDeviceToHost, ctx->stage_stream));
    // the compute stream must not overwrite the column before the copy has read it
    TB_CUDA(cudaEventRecord(ctx->stage_ev, ctx->stage_stream));
    TB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->stage_ev, 0));
    return TB_OK;
}
extern "C" int32_t tb_stage_wait(tb_ctx *ctx) {
    TB_REQUIRE(ctx, "tb_stage_wait: ctx is NULL");
    if (ctx->stage_stream) TB_CUDA(cudaStreamSynchronize(ctx->stage_stream));
    return TB_OK;
}
