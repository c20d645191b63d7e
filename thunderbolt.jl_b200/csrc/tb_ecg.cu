// Pseudo-ECG after Plonsey (1964), Gauss form: Plonsey1964ECGGaussCache of the reference
// (src/modeling/electrophysiology/ecg.jl:1-160; user story docs/src/literate-tutorials/ep04_geselowitz-ecg.jl:96-125).
//   update_ecg!    flux_q = sum_i (D(x_q) . gradN_i) phi_i                at every quadrature point     (:14-38)
//   evaluate_ecg   phi_e(x) = -1/(4 pi kappa_t) sum_cells sum_q ((flux_q . (x_q - x)) / |x_q - x|^3) dOmega   (:86-148)
// The reference stores the fluxes (ncells*nq vectors) between the two calls; here both happen in ONE element sweep:
// the flux of a quadrature point lives in registers and is consumed immediately by up to ECG_NE electrodes, so the
// sweep reads the mesh once (hex: 192 + 32 + 32 B/element + 8 gathered phi values) and writes ECG_NE scalars.
// Same tile staging as the assembly kernels; per-electrode sums are reduced warp -> block -> last block in a fixed
// order (deterministic; the reference adds cell by cell, so values agree to rounding, not bitwise).
#include "tb_internal.cuh"
#include "tb_elements.cuh"

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))
#define ECG_BLOCK 128
#define ECG_NE 8

struct EcgElectrodes {
    double x[ECG_NE * 3];
    int n;
};

template <int NV, int DIM>
__global__ void __launch_bounds__(ECG_BLOCK)
    k_ecg_plonsey(const int *__restrict__ conn, const int *__restrict__ celldofs, const double *__restrict__ coords, int64_t ncells,
                  const tb_elem_tables *__restrict__ gT, int nq, int kind, const double *__restrict__ ddata, double cmchi,
                  const double *__restrict__ phi, const EcgElectrodes el, double *partials, unsigned *ticket, double *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sW = reinterpret_cast<double *>(smem_raw);
    double *sN = sW + nq;
    double *sdN = sN + nq * NV;
    double *sX = sdN + nq * NV * DIM;
    __shared__ double sm[32];
    for (int i = threadIdx.x; i < nq; i += ECG_BLOCK) sW[i] = gT->w[i];
    for (int i = threadIdx.x; i < nq * NV; i += ECG_BLOCK) sN[i] = gT->N[i];
    for (int i = threadIdx.x; i < nq * NV * DIM; i += ECG_BLOCK) sdN[i] = gT->dN[i];
    __shared__ double sD[9];                       // constant coefficient: evaluated once per CTA
    if (kind < 2 && threadIdx.x == 0) tb_eval_D<NV, DIM>(kind, ddata, cmchi, 0, nullptr, sD);
    double acc[ECG_NE];
#pragma unroll
    for (int k = 0; k < ECG_NE; k++) acc[k] = 0.0;
    const int64_t ntiles = (ncells + ECG_BLOCK - 1) / ECG_BLOCK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = tile * ECG_BLOCK;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ECG_BLOCK * NV; idx += ECG_BLOCK) {
            const int e = idx / NV, a = idx - e * NV;
            if (e0 + e < ncells) {
                const int64_t node = conn[e0 * NV + idx];
#pragma unroll
                for (int d = 0; d < DIM; d++) sX[(a * DIM + d) * ECG_BLOCK + e] = coords[node * DIM + d];
            }
        }
        __syncthreads();
        const int64_t e = e0 + threadIdx.x;
        if (e < ncells) {
            const double *X = sX + threadIdx.x;
            double u[NV];
#pragma unroll
            for (int a = 0; a < NV; a++) u[a] = phi[celldofs[e * NV + a]];
            double local[ECG_NE];
#pragma unroll
            for (int k = 0; k < ECG_NE; k++) local[k] = 0.0;
            for (int q = 0; q < nq; q++) {
                double G[NV * DIM], Dloc[DIM * DIM], f[DIM], xq[DIM];
                const double *Nq = sN + q * NV;
                const double dO = tb_map_qp<NV, DIM, ECG_BLOCK, true>(X, sdN + q * NV * DIM, G) * sW[q];
                if (kind >= 2) tb_eval_D<NV, DIM>(kind, ddata, cmchi, e, Nq, Dloc);
                const double *D = kind >= 2 ? Dloc : sD;
#pragma unroll
                for (int r = 0; r < DIM; r++) f[r] = 0.0;
#pragma unroll
                for (int i = 0; i < NV; i++)
#pragma unroll
                    for (int r = 0; r < DIM; r++) {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; k++) s += D[r * DIM + k] * G[i * DIM + k];
                        f[r] += s * u[i];
                    }
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < NV; a++) s += Nq[a] * X[(a * DIM + d) * ECG_BLOCK];
                    xq[d] = s;
                }
#pragma unroll
                for (int k = 0; k < ECG_NE; k++) {
                    if (k < el.n) {
                        double n2 = 0.0, fd = 0.0;
#pragma unroll
                        for (int d = 0; d < DIM; d++) {
                            const double dv = xq[d] - el.x[k * 3 + d];
                            n2 += dv * dv;
                            fd += f[d] * dv;
                        }
                        const double n = sqrt(n2);
                        local[k] += fd / (n * n * n) * dO;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < ECG_NE; k++) acc[k] += local[k];
        }
    }
    // per-electrode block sums, then the last block adds the block partials in block order
    for (int k = 0; k < el.n; k++) {
        const double bs = tb_block_sum(acc[k], sm);
        if (threadIdx.x == 0) partials[(size_t)k * gridDim.x + blockIdx.x] = bs;
    }
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int k = 0; k < el.n; k++) {
            double s = 0.0;
            for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += ((volatile double *)partials)[(size_t)k * gridDim.x + i];
            s = tb_block_sum(s, sm);
            if (threadIdx.x == 0) out[k] = s;
        }
    }
}

template <int NV, int DIM>
static int32_t launch_ecg(tb_ctx *ctx, const tb_mesh *m, const tb_elem_tables *d_T, int nq, int kind, const double *d_data,
                          double cmchi, const double *phi, const EcgElectrodes &el, double *d_out) {
    const size_t smem = sizeof(double) * nq * (1 + NV + NV * DIM) + sizeof(double) * NV * DIM * ECG_BLOCK;
    TB_CUDA(cudaFuncSetAttribute(k_ecg_plonsey<NV, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ecg_plonsey<NV, DIM>, ECG_BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t ntiles = (m->ncells + ECG_BLOCK - 1) / ECG_BLOCK;
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (cap * ECG_NE > 4 * TB_MAX_PARTIALS) cap = 4 * TB_MAX_PARTIALS / ECG_NE;
    const int grid = (int)(ntiles < cap ? (ntiles < 1 ? 1 : ntiles) : cap);
    TB_LAUNCH(ctx, (k_ecg_plonsey<NV, DIM>), grid, ECG_BLOCK, smem, m->d_conn, m->d_celldofs, m->d_coords, m->ncells, d_T, nq, kind,
              d_data, cmchi, phi, el, ctx->d_partials, ctx->d_ticket + 5, d_out);
    return TB_OK;
}

// phi_e[k] for `ne` electrodes (rows of `electrodes`, dim doubles each).  On a partitioned mesh every rank sweeps the
// cells it owns... (not distributed yet: the caller passes the whole mesh; multi-GPU ECG = sum over ranks of disjoint cell sets)
extern "C" int32_t tb_ecg_plonsey(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, int32_t kind, const double *data, int64_t ndata,
                                  double cm_chi, const tb_vec *phi, int32_t phicol, const double *electrodes, int32_t ne,
                                  double kappa_t, double *phi_e) {
    TB_REQUIRE(ctx && mesh && phi && electrodes && phi_e, "tb_ecg_plonsey: NULL argument");
    TB_REQUIRE(ne >= 0, "tb_ecg_plonsey: negative electrode count");
    TB_REQUIRE(phicol >= 0 && phicol < phi->ncols && phi->n >= mesh->ndofs, "tb_ecg_plonsey: phi is smaller than the mesh's dof count");
    TB_REQUIRE(kind >= TB_D_SCALAR && kind <= TB_D_CELL_TENSOR, "tb_ecg_plonsey: unknown coefficient kind %d", kind);
    TB_REQUIRE(kind != TB_D_SPECTRAL || mesh->dim == 3, "tb_ecg_plonsey: spectral coefficient needs a 3D mesh");
    const int64_t need = kind == TB_D_SCALAR ? 1 : kind == TB_D_TENSOR ? mesh->dim * mesh->dim
                         : kind == TB_D_CELL_TENSOR ? mesh->ncells * mesh->dim * mesh->dim : 3 + mesh->ncells * mesh->nv * 9;
    TB_REQUIRE(data && ndata == need, "tb_ecg_plonsey: coefficient kind %d needs %lld doubles, got %lld", kind, (long long)need, (long long)ndata);
    TB_REQUIRE(cm_chi != 0.0 && kappa_t != 0.0, "tb_ecg_plonsey: Cm*chi and kappa_t must be non-zero");
    TB_DEV(ctx);
    const tb_elem_tables *d_T = nullptr;
    int nq_tab = 0;
    TB_TRY(tb_get_tables(ctx, mesh->celltype, qorder, &d_T, &nq_tab));
    double *d_data = nullptr, *d_out = nullptr;
    TB_CUDA(cudaMalloc(&d_data, sizeof(double) * (size_t)ndata));
    TB_CUDA(cudaMalloc(&d_out, sizeof(double) * ECG_NE));
    TB_CUDA(cudaMemcpyAsync(d_data, data, sizeof(double) * (size_t)ndata, cudaMemcpyHostToDevice, ctx->stream));
    const double *ph = phi->d + (size_t)phicol * phi->ld;
    int32_t st = TB_OK;
    const double scale = 4 * 3.14159265358979323846 * kappa_t;
    for (int k0 = 0; k0 < ne && st == TB_OK; k0 += ECG_NE) {
        EcgElectrodes el;
        memset(&el, 0, sizeof(el));
        el.n = ne - k0 < ECG_NE ? ne - k0 : ECG_NE;
        for (int k = 0; k < el.n; k++)
            for (int d = 0; d < mesh->dim; d++) el.x[k * 3 + d] = electrodes[(size_t)(k0 + k) * mesh->dim + d];
        switch (mesh->celltype) {
        case TB_QUAD4: st = launch_ecg<4, 2>(ctx, mesh, d_T, nq_tab, kind, d_data, cm_chi, ph, el, d_out); break;
        case TB_HEX8: st = launch_ecg<8, 3>(ctx, mesh, d_T, nq_tab, kind, d_data, cm_chi, ph, el, d_out); break;
        case TB_TRI3: st = launch_ecg<3, 2>(ctx, mesh, d_T, nq_tab, kind, d_data, cm_chi, ph, el, d_out); break;
        default: st = launch_ecg<4, 3>(ctx, mesh, d_T, nq_tab, kind, d_data, cm_chi, ph, el, d_out); break;
        }
        if (st == TB_OK) {
            double h[ECG_NE];
            if (cudaMemcpyAsync(h, d_out, sizeof(double) * el.n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                st = tb_fail(TB_ERR_CUDA, "tb_ecg_plonsey: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            else
                for (int k = 0; k < el.n; k++) phi_e[k0 + k] = -h[k] / scale;
        }
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_data);
    cudaFree(d_out);
    return st;
}
