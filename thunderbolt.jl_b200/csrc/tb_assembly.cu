// Element-loop assembly of M, K and the stimulus vector into the fixed pattern.
// Replaces update_operator!(op, t) of FerriteOperators' element loop as driven by
// src/solver/time/euler.jl:143-153,173-175 and src/solver/interface.jl:66-94.
//
// Kernel shape (one element per thread, 128 elements per block iteration):
//   1. the block copies the quadrature/shape tables into shared memory once;
//   2. it reads 128 elements' connectivity + celldofs coalesced, and gathers their vertex coordinates
//      with 128*nv independent loads into a transposed shared-memory tile X[c][e] (bank-conflict free
//      when thread e later reads its own element);
//   3. each thread integrates its element matrix in registers (packed upper triangle) and scatters it
//      with fp64 atomics (RED.ADD.F64) into the SELL-32 image of the reference's CSR pattern; the
//      position of (row, col) is found by binary search in the row's sorted columns, so no
//      element->nonzero map (51 GB for the 100 M element slab) is ever stored.
// Atomic scatter order is not deterministic, so values agree with the oracle to rounding (1e-14 rel);
// the pattern itself is bit exact.  Rows >= nrows (ghost rows of a partitioned mesh) are skipped.
#include "tb_internal.cuh"
#include "tb_elements.cuh"

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))
#define AS_BLOCK 128

static int nv_of(int ct) { return ct == TB_QUAD4 ? 4 : ct == TB_HEX8 ? 8 : ct == TB_TRI3 ? 3 : 4; }

extern "C" int32_t tb_quadrature(int32_t celltype, int32_t qorder, int32_t *nq, double *pts, double *weights) {
    TB_REQUIRE(nq, "tb_quadrature: nq is NULL");
    TB_REQUIRE(celltype >= TB_QUAD4 && celltype <= TB_TET4, "tb_quadrature: unknown cell type %d", celltype);
    tb_elem_tables T;
    if (tb_build_tables(celltype, qorder, &T))
        return tb_fail(TB_ERR_UNSUPPORTED, "tb_quadrature: order %d not available for cell type %d", qorder, celltype);
    *nq = T.nq;
    if (pts)
        for (int i = 0; i < T.nq * T.dim; i++) pts[i] = T.xi[i];
    if (weights)
        for (int i = 0; i < T.nq; i++) weights[i] = T.w[i];
    return TB_OK;
}

struct SellView {
    const int64_t *rowptr;
    const int64_t *slice_ptr;
    const int *col;
    double *val;
    int64_t nrows;
};

// position of (row, c) in the SELL arrays, or -1
__device__ __forceinline__ int64_t sell_find(const SellView &S, int row, int c) {
    const int64_t base = S.slice_ptr[row >> 5] + (row & 31);
    int lo = 0, hi = (int)(S.rowptr[row + 1] - S.rowptr[row]);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (S.col[base + (int64_t)mid * TB_SLICE] < c) lo = mid + 1; else hi = mid;
    }
    return base + (int64_t)lo * TB_SLICE;
}

// stage tables + one tile of elements; returns through shared pointers
template <int NV, int DIM>
__device__ __forceinline__ void stage_tile(const int *__restrict__ conn, const int *__restrict__ celldofs,
                                           const double *__restrict__ coords, int64_t e0, int64_t ncells, int *sDof,
                                           double *sX) {
    // celldofs tile, coalesced
    for (int idx = threadIdx.x; idx < AS_BLOCK * NV; idx += AS_BLOCK) {
        const int64_t g = e0 * NV + idx;
        sDof[idx] = g < ncells * NV ? celldofs[g] : 0;
    }
    // coordinates: one (element, vertex) pair per load group
    for (int idx = threadIdx.x; idx < AS_BLOCK * NV; idx += AS_BLOCK) {
        const int e = idx / NV, a = idx - e * NV;
        const int64_t g = e0 * NV + idx;
        if (g < ncells * NV) {
            const int64_t node = conn[g];
#pragma unroll
            for (int d = 0; d < DIM; d++) sX[(a * DIM + d) * AS_BLOCK + e] = coords[node * DIM + d];
        }
    }
}

template <int NV, int DIM, int OP>
__global__ void __launch_bounds__(AS_BLOCK)
    k_assemble_bilinear(const int *__restrict__ conn, const int *__restrict__ celldofs, const double *__restrict__ coords,
                        int64_t ncells, const tb_elem_tables *__restrict__ gT, int nq, double rho, int kind,
                        const double *__restrict__ ddata, double cmchi, SellView S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared memory: the nq used rows of the quadrature/shape tables, then the coordinate tile, then the dof tile
    double *sW = reinterpret_cast<double *>(smem_raw);
    double *sN = sW + nq;
    double *sdN = sN + nq * NV;
    double *sX = sdN + nq * NV * DIM;
    int *sDof = reinterpret_cast<int *>(sX + NV * DIM * AS_BLOCK);
    for (int i = threadIdx.x; i < nq; i += AS_BLOCK) sW[i] = gT->w[i];
    for (int i = threadIdx.x; i < nq * NV; i += AS_BLOCK) sN[i] = gT->N[i];
    for (int i = threadIdx.x; i < nq * NV * DIM; i += AS_BLOCK) sdN[i] = gT->dN[i];
    const tb_tables_view sT{nq, sW, sN, sdN};
    const int64_t ntiles = (ncells + AS_BLOCK - 1) / AS_BLOCK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = tile * AS_BLOCK;
        __syncthreads();
        stage_tile<NV, DIM>(conn, celldofs, coords, e0, ncells, sDof, sX);
        __syncthreads();
        const int64_t e = e0 + threadIdx.x;
        if (e < ncells) {
            double acc[NV * (NV + 1) / 2];
            if (OP == 0) tb_element_mass<NV, DIM, AS_BLOCK>(sT, sX + threadIdx.x, rho, acc);
            else tb_element_diffusion<NV, DIM, AS_BLOCK>(sT, sX + threadIdx.x, kind, ddata, cmchi, e, acc);
            int dof[NV];
#pragma unroll
            for (int a = 0; a < NV; a++) dof[a] = sDof[threadIdx.x * NV + a];
#pragma unroll
            for (int i = 0; i < NV; i++) {
                if (dof[i] >= S.nrows) continue;
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    const double v = i <= j ? acc[tb_sym<NV>(i, j)] : acc[tb_sym<NV>(j, i)];
                    atomicAdd(S.val + sell_find(S, dof[i], dof[j]), v);
                }
            }
        }
    }
}

template <int NV, int DIM>
__global__ void __launch_bounds__(AS_BLOCK)
    k_assemble_source(const int *__restrict__ conn, const int *__restrict__ celldofs, const double *__restrict__ coords,
                      int64_t ncells, const tb_elem_tables *__restrict__ gT, int nq, int kind, const double *__restrict__ prm,
                      double t, const double *__restrict__ fq, double *__restrict__ b, int64_t nrows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared memory: the nq used rows of the quadrature/shape tables, then the coordinate tile, then the dof tile
    double *sW = reinterpret_cast<double *>(smem_raw);
    double *sN = sW + nq;
    double *sdN = sN + nq * NV;
    double *sX = sdN + nq * NV * DIM;
    int *sDof = reinterpret_cast<int *>(sX + NV * DIM * AS_BLOCK);
    for (int i = threadIdx.x; i < nq; i += AS_BLOCK) sW[i] = gT->w[i];
    for (int i = threadIdx.x; i < nq * NV; i += AS_BLOCK) sN[i] = gT->N[i];
    for (int i = threadIdx.x; i < nq * NV * DIM; i += AS_BLOCK) sdN[i] = gT->dN[i];
    const tb_tables_view sT{nq, sW, sN, sdN};
    __shared__ double sprm[8];
    if (threadIdx.x < 8) sprm[threadIdx.x] = prm ? prm[threadIdx.x] : 0.0;
    const int64_t ntiles = (ncells + AS_BLOCK - 1) / AS_BLOCK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = tile * AS_BLOCK;
        __syncthreads();
        stage_tile<NV, DIM>(conn, celldofs, coords, e0, ncells, sDof, sX);
        __syncthreads();
        const int64_t e = e0 + threadIdx.x;
        if (e < ncells) {
            double be[NV];
            tb_element_source<NV, DIM, AS_BLOCK>(sT, sX + threadIdx.x, kind, sprm, t, fq ? fq + e * nq : nullptr, be);
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const int d = sDof[threadIdx.x * NV + j];
                if (d < nrows) atomicAdd(b + d, be[j]);
            }
        }
    }
}

static size_t assembly_smem(int nv, int dim, int nq) {
    return sizeof(double) * nq * (1 + nv + nv * dim) + sizeof(double) * nv * dim * AS_BLOCK + sizeof(int) * nv * AS_BLOCK;
}

static int32_t upload_tables(tb_ctx *ctx, int celltype, int qorder, tb_elem_tables **d_T, int *nq) {
    tb_elem_tables T;
    memset(&T, 0, sizeof(T));
    if (tb_build_tables(celltype, qorder, &T))
        return tb_fail(TB_ERR_UNSUPPORTED, "assembly: quadrature order %d not available for cell type %d", qorder, celltype);
    TB_CUDA(cudaMalloc(d_T, sizeof(T)));
    TB_CUDA(cudaMemcpy(*d_T, &T, sizeof(T), cudaMemcpyHostToDevice));
    if (nq) *nq = T.nq;
    return TB_OK;
}

template <int NV, int DIM, int OP>
static int32_t launch_bilinear(tb_ctx *ctx, const tb_mesh *m, const tb_elem_tables *d_T, int nq, double rho, int kind,
                               const double *d_data, double cmchi, const SellView &S) {
    const size_t smem = assembly_smem(NV, DIM, nq);
    TB_CUDA(cudaFuncSetAttribute(k_assemble_bilinear<NV, DIM, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (m->ncells + AS_BLOCK - 1) / AS_BLOCK;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_bilinear<NV, DIM, OP>, AS_BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    int grid = (int)(ntiles < (int64_t)ctx->sm_count * per_sm ? ntiles : (int64_t)ctx->sm_count * per_sm);
    TB_LAUNCH(ctx, (k_assemble_bilinear<NV, DIM, OP>), grid, AS_BLOCK, smem, m->d_conn, m->d_celldofs, m->d_coords,
              m->ncells, d_T, nq, rho, kind, d_data, cmchi, S);
    return TB_OK;
}

static int32_t assemble_bilinear(tb_ctx *ctx, const tb_mesh *mesh, int qorder, int op, double rho, int kind,
                                 const double *data, int64_t ndata, double cmchi, tb_csr *A) {
    TB_REQUIRE(ctx && mesh && A, "assemble: NULL argument");
    TB_REQUIRE(A->pat->nrows == mesh->ndofs_owned && A->pat->ncols == mesh->ndofs,
               "assemble: operator is %lld x %lld but the mesh has %lld owned / %lld total dofs", (long long)A->pat->nrows,
               (long long)A->pat->ncols, (long long)mesh->ndofs_owned, (long long)mesh->ndofs);
    TB_REQUIRE(ctx->assembly_mode == 0, "assemble: colouring mode is not implemented yet (use mode 0, atomics)");
    TB_DEV(ctx);
    tb_elem_tables *d_T = nullptr;
    int nq = 0;
    TB_TRY(upload_tables(ctx, mesh->celltype, qorder, &d_T, &nq));
    double *d_data = nullptr;
    if (op == 1) {
        const int64_t need = kind == TB_D_SCALAR ? 1 : kind == TB_D_TENSOR ? mesh->dim * mesh->dim
                                                                          : 3 + mesh->ncells * mesh->nv * 9;
        TB_REQUIRE(kind >= TB_D_SCALAR && kind <= TB_D_SPECTRAL, "tb_assemble_diffusion: unknown coefficient kind %d", kind);
        TB_REQUIRE(kind != TB_D_SPECTRAL || mesh->dim == 3, "tb_assemble_diffusion: spectral coefficient needs a 3D mesh");
        TB_REQUIRE(data && ndata == need, "tb_assemble_diffusion: coefficient kind %d needs %lld doubles, got %lld", kind,
                   (long long)need, (long long)ndata);
        TB_REQUIRE(cmchi != 0.0, "tb_assemble_diffusion: Cm*chi must be non-zero");
        TB_CUDA(cudaMalloc(&d_data, sizeof(double) * (size_t)ndata));
        TB_CUDA(cudaMemcpyAsync(d_data, data, sizeof(double) * (size_t)ndata, cudaMemcpyHostToDevice, ctx->stream));
    }
    TB_CUDA(cudaMemsetAsync(A->d_val, 0, sizeof(double) * (size_t)A->pat->sell_len, ctx->stream));
    SellView S{A->pat->d_rowptr, A->pat->d_slice_ptr, A->pat->d_col, A->d_val, A->pat->nrows};
    int32_t st = TB_OK;
#define DISPATCH(NV, DIM)                                                                                   \
    st = op == 0 ? launch_bilinear<NV, DIM, 0>(ctx, mesh, d_T, nq, rho, kind, d_data, cmchi, S)            \
                 : launch_bilinear<NV, DIM, 1>(ctx, mesh, d_T, nq, rho, kind, d_data, cmchi, S)
    switch (mesh->celltype) {
    case TB_QUAD4: DISPATCH(4, 2); break;
    case TB_HEX8: DISPATCH(8, 3); break;
    case TB_TRI3: DISPATCH(3, 2); break;
    default: DISPATCH(4, 3); break;
    }
#undef DISPATCH
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_T);
    cudaFree(d_data);
    if (st != TB_OK) return st;
    if (e != cudaSuccess) return tb_fail(TB_ERR_CUDA, "assemble: kernel failed: %s", cudaGetErrorString(e));
    return TB_OK;
}

extern "C" int32_t tb_assemble_mass(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, double rho, tb_csr *M) {
    return assemble_bilinear(ctx, mesh, qorder, 0, rho, 0, nullptr, 0, 1.0, M);
}

extern "C" int32_t tb_assemble_diffusion(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, int32_t kind,
                                         const double *data, int64_t ndata, double cm_chi, tb_csr *K) {
    return assemble_bilinear(ctx, mesh, qorder, 1, 1.0, kind, data, ndata, cm_chi, K);
}

template <int NV, int DIM>
static int32_t launch_source(tb_ctx *ctx, const tb_mesh *m, const tb_elem_tables *d_T, int nq, int kind, const double *d_prm,
                             double t, const double *d_fq, double *b) {
    const size_t smem = assembly_smem(NV, DIM, nq);
    TB_CUDA(cudaFuncSetAttribute(k_assemble_source<NV, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (m->ncells + AS_BLOCK - 1) / AS_BLOCK;
    int grid = (int)(ntiles < (int64_t)ctx->sm_count * 4 ? ntiles : (int64_t)ctx->sm_count * 4);
    TB_LAUNCH(ctx, (k_assemble_source<NV, DIM>), grid, AS_BLOCK, smem, m->d_conn, m->d_celldofs, m->d_coords, m->ncells,
              d_T, nq, kind, d_prm, t, d_fq, b, m->ndofs_owned);
    return TB_OK;
}

static int32_t assemble_source(tb_ctx *ctx, const tb_mesh *mesh, int qorder, int kind, const double *prm, int nprm,
                               double t, const double *fq, tb_vec *b, int bcol) {
    TB_REQUIRE(ctx && mesh && b, "tb_assemble_source: NULL argument");
    TB_REQUIRE(bcol >= 0 && bcol < b->ncols && b->n >= mesh->ndofs_owned, "tb_assemble_source: vector too small");
    TB_REQUIRE(fq || (kind >= TB_SRC_NONE && kind <= TB_SRC_ENDO), "tb_assemble_source: unknown source kind %d", kind);
    TB_REQUIRE(nprm >= 0 && nprm <= 8, "tb_assemble_source: at most 8 parameters");
    TB_DEV(ctx);
    tb_elem_tables *d_T = nullptr;
    int nq = 0;
    TB_TRY(upload_tables(ctx, mesh->celltype, qorder, &d_T, &nq));
    double hp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nprm; i++) hp[i] = prm[i];
    double *d_prm = nullptr, *d_fq = nullptr;
    TB_CUDA(cudaMalloc(&d_prm, sizeof(hp)));
    TB_CUDA(cudaMemcpyAsync(d_prm, hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
    if (fq) {
        const size_t bytes = sizeof(double) * (size_t)(mesh->ncells * nq);
        TB_CUDA(cudaMalloc(&d_fq, bytes));
        TB_CUDA(cudaMemcpyAsync(d_fq, fq, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    double *bp = b->d + (size_t)bcol * b->ld;
    TB_CUDA(cudaMemsetAsync(bp, 0, sizeof(double) * (size_t)b->n, ctx->stream));
    int32_t st = TB_OK;
    if (fq || kind != TB_SRC_NONE) {
        switch (mesh->celltype) {
        case TB_QUAD4: st = launch_source<4, 2>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
        case TB_HEX8: st = launch_source<8, 3>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
        case TB_TRI3: st = launch_source<3, 2>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
        default: st = launch_source<4, 3>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_T);
    cudaFree(d_prm);
    cudaFree(d_fq);
    if (st != TB_OK) return st;
    if (e != cudaSuccess) return tb_fail(TB_ERR_CUDA, "tb_assemble_source: kernel failed: %s", cudaGetErrorString(e));
    return TB_OK;
}

extern "C" int32_t tb_assemble_source(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, int32_t kind, const double *prm,
                                      int32_t nprm, double t, tb_vec *b, int32_t bcol) {
    TB_REQUIRE(kind == TB_SRC_NONE || prm || nprm == 0, "tb_assemble_source: prm is NULL");
    return assemble_source(ctx, mesh, qorder, kind, prm, nprm, t, nullptr, b, bcol);
}

extern "C" int32_t tb_assemble_source_qp(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, const double *fq, tb_vec *b,
                                         int32_t bcol) {
    TB_REQUIRE(fq, "tb_assemble_source_qp: fq is NULL");
    return assemble_source(ctx, mesh, qorder, 0, nullptr, 0, 0.0, fq, b, bcol);
}
