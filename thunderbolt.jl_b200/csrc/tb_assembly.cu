// Element-loop assembly of M, K and the stimulus vector into the fixed pattern.
// Replaces update_operator!(op, t) of FerriteOperators' element loop as driven by
// src/solver/time/euler.jl:143-153,173-175 and src/solver/interface.jl:66-94.
//
// Two strategies (tb_assembly_set_mode / env TB_ASSEMBLY_MODE):
//   mode 2 (default) "element assembly + ordered gather" -- the B200 take on the reference's
//     ElementAssemblyStrategy (per-element results, then a reduction):
//       phase 1  k_element_matrices / k_element_vectors: one element per thread, full element matrix
//                (vector) written to a scratch buffer EA[cell][i][j] (EAb[cell][j]), no atomics;
//       phase 2  k_gather_rows / k_gather_vec: one matrix row per lane (one SELL slice per warp); the
//                row walks its dof -> (cell, local index) adjacency in ASCENDING cell order, adds row
//                `a` of each adjacent element matrix into a shared-memory image of the row and stores
//                the slice once, fully coalesced.  Every entry is the sum of its element contributions
//                in the reference's sequential element order, starting from 0.0: the assembled values
//                are BITWISE those of the CPU path, run to run and across GPU counts.
//     Rows are processed in chunks so the scratch stays under ctx->ea_budget_bytes (default 8 GB; a
//     chunk covers the contiguous cell range its rows touch, so a locality-preserving numbering wastes
//     nothing); if some chunk's cell range cannot fit, the call falls back to mode 0 and says so
//     (tb_assembly_info).
//   mode 0 "atomic scatter": element matrix in registers, RED.ADD.F64 into the pattern.
//
// Mode-0 kernel shape (one element per thread, 128 elements per block iteration):
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

struct SrcParams {      // travels by value as a kernel argument
    double p[8];
    tb_src_program prog;   // used when kind == TB_SRC_PROGRAM
};
__device__ __forceinline__ void stage_program(const SrcParams &prm, int kind, tb_src_program *sprog) {
    if (kind != TB_SRC_PROGRAM) return;
    const int *src = reinterpret_cast<const int *>(&prm.prog);
    int *dst = reinterpret_cast<int *>(sprog);
    for (int i = threadIdx.x; i < (int)(sizeof(tb_src_program) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
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
                      int64_t ncells, const tb_elem_tables *__restrict__ gT, int nq, int kind, const SrcParams prm,
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
    __shared__ tb_src_program sprog;
    if (threadIdx.x < 8) sprm[threadIdx.x] = prm.p[threadIdx.x];
    stage_program(prm, kind, &sprog);
    const tb_src_program *prog = kind == TB_SRC_PROGRAM ? &sprog : nullptr;
    const int64_t ntiles = (ncells + AS_BLOCK - 1) / AS_BLOCK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = tile * AS_BLOCK;
        __syncthreads();
        stage_tile<NV, DIM>(conn, celldofs, coords, e0, ncells, sDof, sX);
        __syncthreads();
        const int64_t e = e0 + threadIdx.x;
        if (e < ncells) {
            double be[NV];
            tb_element_source<NV, DIM, AS_BLOCK>(sT, sX + threadIdx.x, kind, sprm, t, fq ? fq + e * nq : nullptr, be, prog);
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const int d = sDof[threadIdx.x * NV + j];
                if (d < nrows) atomicAdd(b + d, be[j]);
            }
        }
    }
}


// =====================================================================================================
// mode 2: element matrices + ordered row gather
// =====================================================================================================
__global__ void k_adjg_count(const int *__restrict__ celldofs, int64_t npos, int nrows, int *count) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (int64_t)gridDim.x * blockDim.x) {
        const int d = celldofs[p];
        if (d < nrows) atomicAdd(&count[d], 1);
    }
}
__global__ void k_adjg_widen(const int *__restrict__ src, int64_t *__restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_adjg_fill(const int *__restrict__ celldofs, int64_t npos, int nrows, const int64_t *__restrict__ adjptr,
                            int *cursor, unsigned *__restrict__ adj) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (int64_t)gridDim.x * blockDim.x) {
        const int d = celldofs[p];
        if (d < nrows) adj[adjptr[d] + atomicAdd(&cursor[d], 1)] = (unsigned)p;
    }
}
// the fill order above is arbitrary (atomics); sorting each short list makes the adjacency -- and with it the
// summation order of the gather -- deterministic and equal to the reference's element order
__global__ void k_adjg_sort(int nrows, const int64_t *__restrict__ adjptr, unsigned *__restrict__ adj) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        unsigned *a = adj + adjptr[r];
        const int n = (int)(adjptr[r + 1] - adjptr[r]);
        for (int i = 1; i < n; i++) {
            const unsigned v = a[i];
            int k = i - 1;
            while (k >= 0 && a[k] > v) {
                a[k + 1] = a[k];
                k--;
            }
            a[k + 1] = v;
        }
    }
}

static int32_t mesh_ensure_adjacency(tb_ctx *ctx, const tb_mesh *m) {
    if (m->d_adjptr) return TB_OK;
    const int nrows = (int)m->ndofs_owned;
    const int64_t npos = m->ncells * m->nv;
    TB_REQUIRE(npos < ((int64_t)1 << 32), "assembly: mesh has %lld (cell, vertex) pairs, the adjacency holds at most 2^32",
               (long long)npos);
    const int grid = ctx->sm_count * 8;
    int *count = nullptr;
    int64_t *count64 = nullptr, *adjptr = nullptr;
    unsigned *adj = nullptr;
    TB_CUDA(cudaMalloc(&count, sizeof(int) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&count64, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&adjptr, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(nrows + 1), ctx->stream));
    TB_LAUNCH(ctx, k_adjg_count, grid, 256, 0, m->d_celldofs, npos, nrows, count);
    TB_LAUNCH(ctx, k_adjg_widen, grid, 256, 0, count, count64, (int64_t)nrows + 1);
    TB_TRY(tb_exclusive_scan_i64(ctx, count64, adjptr, (int64_t)nrows + 1));
    int64_t nadj = 0;
    TB_CUDA(cudaMemcpy(&nadj, adjptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMalloc(&adj, sizeof(unsigned) * (size_t)(nadj + 1)));
    TB_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(nrows + 1), ctx->stream));
    TB_LAUNCH(ctx, k_adjg_fill, grid, 256, 0, m->d_celldofs, npos, nrows, adjptr, count, adj);
    TB_LAUNCH(ctx, k_adjg_sort, tb_grid_for(ctx, nrows, 128, 16), 128, 0, nrows, adjptr, adj);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(count);
    cudaFree(count64);
    m->d_adjptr = adjptr;
    m->d_adj = adj;
    m->nadj = nadj;
    return TB_OK;
}

// per 32-row slice: smallest and largest adjacent cell (lists are sorted: first entry = smallest, last = largest)
__global__ void k_slice_cell_range(int nrows, int64_t nslices, const int64_t *__restrict__ adjptr,
                                   const unsigned *__restrict__ adj, int nv, int *__restrict__ smin, int *__restrict__ smax) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += nw) {
        const int64_t r = s * TB_SLICE + lane;
        int mn = INT32_MAX, mx = -1;
        if (r < nrows) {
            const int64_t a = adjptr[r], b = adjptr[r + 1];
            if (b > a) {
                mn = (int)(adj[a] / (unsigned)nv);
                mx = (int)(adj[b - 1] / (unsigned)nv);
            }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) {
            smin[s] = mn;
            smax[s] = mx;
        }
    }
}

typedef tb_mesh::GatherPlan GatherPlan;

// Largest row chunk whose contiguous cell range fits `budget_cells` (host arithmetic on the cached per-slice
// ranges); plan.ok = false if even a single slice does not fit.  which: 0 matrices, 1 vectors.
static int32_t plan_gather(tb_ctx *ctx, const tb_mesh *m, int64_t budget_cells, int which, const GatherPlan **out) {
    GatherPlan &plan = m->plan[which];
    *out = &plan;
    if (plan.budget_cells == budget_cells) return TB_OK;
    const int64_t nrows = m->ndofs_owned;
    const int64_t nslices = (nrows + TB_SLICE - 1) / TB_SLICE;
    if ((int64_t)m->slice_cmin.size() != nslices) {
        int *d = nullptr;
        TB_CUDA(cudaMalloc(&d, sizeof(int) * 2 * (size_t)(nslices + 1)));
        TB_LAUNCH(ctx, k_slice_cell_range, tb_grid_for(ctx, nslices * 32, 256, 8), 256, 0, (int)nrows, nslices, m->d_adjptr,
                  m->d_adj, m->nv, d, d + nslices);
        m->slice_cmin.resize((size_t)nslices);
        m->slice_cmax.resize((size_t)nslices);
        TB_CUDA(cudaMemcpyAsync(m->slice_cmin.data(), d, sizeof(int) * (size_t)nslices, cudaMemcpyDeviceToHost, ctx->stream));
        TB_CUDA(cudaMemcpyAsync(m->slice_cmax.data(), d + nslices, sizeof(int) * (size_t)nslices, cudaMemcpyDeviceToHost, ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d);
    }
    plan = GatherPlan();
    plan.budget_cells = budget_cells;
    if (budget_cells < 1 || nslices == 0) return TB_OK;
    const double cells_per_row = (double)m->ncells / (double)(nrows > 0 ? nrows : 1);
    int64_t spc = (int64_t)((double)budget_cells / (cells_per_row > 1e-9 ? cells_per_row : 1e-9) * 0.9) / TB_SLICE;   // slices per chunk
    if (spc > nslices) spc = nslices;
    if (spc < 1) spc = 1;
    for (;;) {
        const int64_t nch = (nslices + spc - 1) / spc;
        std::vector<int> cmin((size_t)nch, INT32_MAX), cmax((size_t)nch, -1);
        int64_t worst = 0;
        for (int64_t ch = 0; ch < nch; ch++) {
            int mn = INT32_MAX, mx = -1;
            const int64_t s1 = (ch + 1) * spc < nslices ? (ch + 1) * spc : nslices;
            for (int64_t s = ch * spc; s < s1; s++) {
                if (m->slice_cmin[(size_t)s] < mn) mn = m->slice_cmin[(size_t)s];
                if (m->slice_cmax[(size_t)s] > mx) mx = m->slice_cmax[(size_t)s];
            }
            cmin[(size_t)ch] = mn;
            cmax[(size_t)ch] = mx;
            if (mx >= mn && (int64_t)mx - mn + 1 > worst) worst = (int64_t)mx - mn + 1;
        }
        if (worst <= budget_cells) {
            plan.ok = true;
            plan.rows_per_chunk = spc * TB_SLICE;
            plan.cmin.swap(cmin);
            plan.cmax.swap(cmax);
            plan.max_cells = worst;
            return TB_OK;
        }
        if (spc == 1) return TB_OK;
        spc = (spc + 1) / 2;
    }
}

// ---- phase 1: element matrices / vectors into the scratch buffer ------------------------------------------
// Each thread integrates one element matrix in registers; the warp then transposes its 32 matrices through a
// padded shared-memory tile so that EA is written with fully coalesced 256 B stores (a thread writing its own
// 8*NV*NV bytes would touch 32 different lines per store instruction).
#define EA_PAD 33
template <int NV, int DIM, int OP>
__global__ void __launch_bounds__(AS_BLOCK)
    k_element_matrices(const int *__restrict__ conn, const double *__restrict__ coords, int64_t c0, int64_t c1,
                       const tb_elem_tables *__restrict__ gT, int nq, double rho, int kind, const double *__restrict__ ddata,
                       double cmchi, double *__restrict__ EA) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NE = NV * NV;
    double *sW = reinterpret_cast<double *>(smem_raw);
    double *sN = sW + nq;
    double *sdN = sN + nq * NV;
    double *sX = sdN + nq * NV * DIM;
    double *sT = sX + NV * DIM * AS_BLOCK + (threadIdx.x >> 5) * (NE * EA_PAD);   // this warp's transpose tile
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nq; i += AS_BLOCK) sW[i] = gT->w[i];
    for (int i = threadIdx.x; i < nq * NV; i += AS_BLOCK) sN[i] = gT->N[i];
    for (int i = threadIdx.x; i < nq * NV * DIM; i += AS_BLOCK) sdN[i] = gT->dN[i];
    const tb_tables_view sT_tab{nq, sW, sN, sdN};
    __shared__ double sD[9];                       // constant diffusion tensor, evaluated once per CTA
    if (OP == 1 && kind < 2 && threadIdx.x == 0) tb_eval_D<NV, DIM>(kind, ddata, cmchi, 0, nullptr, sD);
    const int64_t ncl = c1 - c0;
    const int64_t ntiles = (ncl + AS_BLOCK - 1) / AS_BLOCK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = c0 + tile * AS_BLOCK;
        __syncthreads();
        for (int idx = threadIdx.x; idx < AS_BLOCK * NV; idx += AS_BLOCK) {
            const int e = idx / NV, a = idx - e * NV;
            if (e0 + e < c1) {
                const int64_t node = conn[e0 * NV + idx];
#pragma unroll
                for (int d = 0; d < DIM; d++) sX[(a * DIM + d) * AS_BLOCK + e] = coords[node * DIM + d];
            }
        }
        __syncthreads();
        const int64_t e = e0 + threadIdx.x;
        if (e < c1) {
            if (OP == 0) {
                double acc[NV * (NV + 1) / 2];
                tb_element_mass<NV, DIM, AS_BLOCK>(sT_tab, sX + threadIdx.x, rho, acc);
#pragma unroll
                for (int i = 0; i < NV; i++)
#pragma unroll
                    for (int j = 0; j < NV; j++)
                        sT[(i * NV + j) * EA_PAD + lane] = i <= j ? acc[tb_sym<NV>(i, j)] : acc[tb_sym<NV>(j, i)];
            } else {
                double Ke[NE];
                tb_element_diffusion_full<NV, DIM, AS_BLOCK>(sT_tab, sX + threadIdx.x, kind, ddata, cmchi, e, Ke, kind < 2 ? sD : nullptr);
#pragma unroll
                for (int i = 0; i < NE; i++) sT[i * EA_PAD + lane] = Ke[i];
            }
        }
        __syncwarp();
        // the warp's elements are EA[(ew0 - c0) .. +32) * NE, one contiguous run of 32*NE doubles
        const int64_t ew0 = e0 + (threadIdx.x & ~31);
        const int nvalid = (int)(c1 - ew0 < 32 ? (c1 - ew0 > 0 ? c1 - ew0 : 0) : 32);
        double *out = EA + (ew0 - c0) * NE;
        for (int k = lane; k < nvalid * NE; k += 32) {
            const int el = k / NE, i = k - el * NE;
            out[k] = sT[i * EA_PAD + el];
        }
        __syncwarp();
    }
}

// ---- phase 1, split form (default): geometry per (element, point), then one thread per element-matrix column group ----
// One thread per element keeps the whole NV x NV matrix in registers (hex: 64 accumulators + 24 gradients = the 255
// register limit, 8 warps per SM, fp64 pipe 62 % busy waiting on its own dependent instructions).  Here a CTA works on a
// tile of TE elements in two alternating phases over batches of qb quadrature points:
//   A  one thread per (element, point): J, det J, J^-1, the NV spatial gradients and dOmega (and the diffusion tensor when
//      it varies per point) -> shared memory, element index fastest (bank-conflict free);
//   B  NV / CPT threads per element, each owning CPT adjacent columns j of Ke (NV * CPT accumulators): for every point of
//      the batch it reads the element's gradients (broadcast within the element's threads), forms gradN_j . D once per
//      column and subtracts ((gradN_j . D) . gradN_i) dOmega for all i -- the same operations in the same order as the
//      one-thread kernel (tb_element_diffusion_full), so Ke is bit for bit the same; only the leading "0.0 +" of each
//      short dot product is dropped (it can only change the sign of an exact zero, which no accumulator can observe:
//      they start at +0.0 and x - (+-0) = x, (+0) - (+-0) = +0).
// The element's threads hold adjacent columns, so EA[cell][i][j0..j0+CPT) leaves as 16 B (CPT = 2) stores that fill whole
// sectors; no transposition tile is needed.
template <int NV, int DIM, int OP, int CPT, int TE>
__global__ void __launch_bounds__(TE *(NV / CPT), 2)
    k_element_matrices_split(const int *__restrict__ conn, const double *__restrict__ coords, int64_t c0, int64_t c1,
                             const tb_elem_tables *__restrict__ gT, int nq, int qb, double rho, int kind,
                             const double *__restrict__ ddata, double cmchi, double *__restrict__ EA, int64_t ea_si, int ea_sc) {
    static_assert(NV % CPT == 0, "columns per thread must divide the element size");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TPE = NV / CPT, BLOCK = TE * TPE, NG = NV * DIM;
    constexpr int GS = OP == 1 ? NG + 1 : 1;            // doubles per (element, point): gradients + dOmega | dOmega
    double *sW = reinterpret_cast<double *>(smem_raw);
    double *sN = sW + nq;
    double *sdN = sN + nq * NV;
    double *sX = sdN + nq * NV * DIM;
    double *sG = sX + NG * TE;
    double *sDq = sG + (size_t)qb * GS * TE;            // per-point tensors (kind >= 2 only)
    for (int i = threadIdx.x; i < nq; i += BLOCK) sW[i] = gT->w[i];
    for (int i = threadIdx.x; i < nq * NV; i += BLOCK) sN[i] = gT->N[i];
    for (int i = threadIdx.x; i < nq * NV * DIM; i += BLOCK) sdN[i] = gT->dN[i];
    __shared__ double sD[9];                            // constant diffusion tensor, evaluated once per CTA
    if (OP == 1 && kind < 2 && threadIdx.x == 0) tb_eval_D<NV, DIM>(kind, ddata, cmchi, 0, nullptr, sD);
    const int jj = threadIdx.x % TPE, el = threadIdx.x / TPE;
    const int64_t ncl = c1 - c0;
    const int64_t ntiles = (ncl + TE - 1) / TE;
    // vertex coordinates of a tile: TE * NV (element, vertex) pairs, CPT per thread; the NEXT tile's are fetched into
    // registers while this tile is integrated, so the dependent conn -> coords loads never sit in front of a barrier.
    // (Tried instead: node ids two tiles ahead + cp.async of the coordinates into a second shared-memory buffer, no
    // registers held -- hex K unchanged at 3.51 ms, mass and tetrahedra 6-9 % slower: kept the register version.)
    double pre[CPT * DIM];
    auto fetch = [&](int64_t tile) {
        const int64_t e0 = c0 + tile * TE;
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int idx = threadIdx.x + k * BLOCK;
            if (tile < ntiles && e0 + idx / NV < c1) {
                const int64_t node = conn[e0 * NV + idx];
#pragma unroll
                for (int d = 0; d < DIM; d++) pre[k * DIM + d] = coords[node * DIM + d];
            }
        }
    };
    fetch(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = c0 + tile * TE;
        // (no barrier here: phase B of the previous tile reads sG only, and every thread passed the barrier that followed
        // its last phase A -- the only reader of sX -- before it can arrive here)
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int idx = threadIdx.x + k * BLOCK;
            const int e = idx / NV, a = idx - e * NV;
#pragma unroll
            for (int d = 0; d < DIM; d++) sX[(a * DIM + d) * TE + e] = pre[k * DIM + d];
        }
        __syncthreads();
        fetch(tile + gridDim.x);
        double acc[NV * CPT];
#pragma unroll
        for (int i = 0; i < NV * CPT; i++) acc[i] = 0.0;
        for (int q0 = 0; q0 < nq; q0 += qb) {
            const int nqb = nq - q0 < qb ? nq - q0 : qb;
            if (q0 > 0) __syncthreads();                // phase B of the previous batch is done with sG
            for (int task = threadIdx.x; task < TE * nqb; task += BLOCK) {
                const int ql = task / TE, e = task - ql * TE, q = q0 + ql;
                if (e0 + e >= c1) continue;
                double *g = sG + (size_t)ql * GS * TE + e;
                if constexpr (OP == 1) {
                    double G[NG];
                    const double dO = tb_map_qp<NV, DIM, TE, true>(sX + e, sdN + q * NG, G) * sW[q];
#pragma unroll
                    for (int c = 0; c < NG; c++) g[c * TE] = G[c];
                    g[NG * TE] = dO;
                    if (kind >= 2) {
                        double Dq[DIM * DIM];
                        tb_eval_D<NV, DIM>(kind, ddata, cmchi, e0 + e, sN + q * NV, Dq);
#pragma unroll
                        for (int c = 0; c < DIM * DIM; c++) sDq[((size_t)ql * DIM * DIM + c) * TE + e] = Dq[c];
                    }
                } else {
                    g[0] = tb_map_qp<NV, DIM, TE, false>(sX + e, sdN + q * NG, nullptr) * sW[q];
                }
            }
            __syncthreads();
            if (e0 + el < c1) {
                for (int ql = 0; ql < nqb; ql++) {
                    const double *g = sG + (size_t)ql * GS * TE + el;
                    if constexpr (OP == 0) {
                        const double dO = g[0];
                        const double *Nq = sN + (q0 + ql) * NV;
#pragma unroll
                        for (int c = 0; c < CPT; c++) {
                            const double Nj = Nq[jj * CPT + c];
#pragma unroll
                            for (int i = 0; i < NV; i++) acc[i * CPT + c] += rho * (Nq[i] * Nj) * dO;
                        }
                    } else {
                        // gradN_j . D for this thread's columns first (the tensor is only live here), then row by row so
                        // that only one gradN_i is in registers at a time
                        const double dO = g[NG * TE];
                        double gD[CPT][DIM];
                        {
                            double Dq[DIM * DIM];
                            if (kind >= 2) {
#pragma unroll
                                for (int c = 0; c < DIM * DIM; c++) Dq[c] = sDq[((size_t)ql * DIM * DIM + c) * TE + el];
                            } else {
#pragma unroll
                                for (int c = 0; c < DIM * DIM; c++) Dq[c] = sD[c];
                            }
#pragma unroll
                            for (int c = 0; c < CPT; c++) {
                                const int j = jj * CPT + c;
                                double Gj[DIM];
#pragma unroll
                                for (int d = 0; d < DIM; d++) Gj[d] = g[(j * DIM + d) * TE];
                                if (kind == 0) {
                                    // _inner_product_helper(a, B::AbstractFloat, c) = (a . c) * B: keep gradN_j, scale later
#pragma unroll
                                    for (int d = 0; d < DIM; d++) gD[c][d] = Gj[d];
                                } else {
#pragma unroll
                                    for (int l = 0; l < DIM; l++) {
                                        double s = Gj[0] * Dq[l];
#pragma unroll
                                        for (int k = 1; k < DIM; k++) s += Gj[k] * Dq[k * DIM + l];
                                        gD[c][l] = s;
                                    }
                                }
                            }
                        }
                        const double D0 = sD[0];
                        double Gi[NG];      // all loads issued before the arithmetic that consumes them
#pragma unroll
                        for (int c = 0; c < NG; c++) Gi[c] = g[c * TE];
#pragma unroll
                        for (int i = 0; i < NV; i++) {
#pragma unroll
                            for (int c = 0; c < CPT; c++) {
                                double s = gD[c][0] * Gi[i * DIM];
#pragma unroll
                                for (int l = 1; l < DIM; l++) s += gD[c][l] * Gi[i * DIM + l];
                                if (kind == 0) acc[i * CPT + c] -= (s * D0) * dO;
                                else acc[i * CPT + c] -= s * dO;
                            }
                        }
                    }
                }
            }
        }
        // EA[i][cell][j] (row-of-the-element major, ea_si = cells of the chunk * NV): the rows that gather local row i of
        // neighbouring cells -- neighbouring matrix rows on a structured grid -- read one contiguous run
        if (e0 + el < c1) {
            double *out = EA + (e0 - c0 + el) * ea_sc + jj * CPT;
#pragma unroll
            for (int i = 0; i < NV; i++) {
                if constexpr (CPT == 2) {
                    *reinterpret_cast<double2 *>(out + i * ea_si) = make_double2(acc[i * 2], acc[i * 2 + 1]);
                } else {
#pragma unroll
                    for (int c = 0; c < CPT; c++) out[i * ea_si + c] = acc[i * CPT + c];
                }
            }
        }
    }
}

template <int NV, int DIM>
__global__ void __launch_bounds__(AS_BLOCK)
    k_element_vectors(const int *__restrict__ conn, const double *__restrict__ coords, int64_t c0, int64_t c1,
                      const tb_elem_tables *__restrict__ gT, int nq, int kind, const SrcParams prm, double t,
                      const double *__restrict__ fq, double *__restrict__ EAb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sW = reinterpret_cast<double *>(smem_raw);
    double *sN = sW + nq;
    double *sdN = sN + nq * NV;
    double *sX = sdN + nq * NV * DIM;
    for (int i = threadIdx.x; i < nq; i += AS_BLOCK) sW[i] = gT->w[i];
    for (int i = threadIdx.x; i < nq * NV; i += AS_BLOCK) sN[i] = gT->N[i];
    for (int i = threadIdx.x; i < nq * NV * DIM; i += AS_BLOCK) sdN[i] = gT->dN[i];
    const tb_tables_view sT{nq, sW, sN, sdN};
    __shared__ double sprm[8];
    __shared__ tb_src_program sprog;
    if (threadIdx.x < 8) sprm[threadIdx.x] = prm.p[threadIdx.x];
    stage_program(prm, kind, &sprog);
    const tb_src_program *prog = kind == TB_SRC_PROGRAM ? &sprog : nullptr;
    const int64_t ncl = c1 - c0;
    const int64_t ntiles = (ncl + AS_BLOCK - 1) / AS_BLOCK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t e0 = c0 + tile * AS_BLOCK;
        __syncthreads();
        for (int idx = threadIdx.x; idx < AS_BLOCK * NV; idx += AS_BLOCK) {
            const int e = idx / NV, a = idx - e * NV;
            if (e0 + e < c1) {
                const int64_t node = conn[e0 * NV + idx];
#pragma unroll
                for (int d = 0; d < DIM; d++) sX[(a * DIM + d) * AS_BLOCK + e] = coords[node * DIM + d];
            }
        }
        __syncthreads();
        const int64_t e = e0 + threadIdx.x;
        if (e < c1) {
            double be[NV];
            tb_element_source<NV, DIM, AS_BLOCK>(sT, sX + threadIdx.x, kind, sprm, t, fq ? fq + e * nq : nullptr, be, prog);
#pragma unroll
            for (int j = 0; j < NV; j++) EAb[(e - c0) * NV + j] = be[j];
        }
    }
}

// ---- phase 2: ordered gather ---------------------------------------------------------------------------------
// One warp per SELL slice, one row per lane.  sacc[j*32 + lane] is the shared-memory image of entry j of the
// lane's row (bank-conflict free); the lane walks its adjacency in ascending (cell, a) order and adds row a of
// EA[cell] at the positions of the cell's dofs (binary search in the row's sorted columns, which the 32 lanes
// read from the same 128 B lines).  The slice is then stored once, 256 B per entry slot.
template <int NV> struct GatherCell {
    int dof[NV];
    double kv[NV];
};
template <int NV>
__device__ __forceinline__ void gather_load(GatherCell<NV> &g, unsigned p, const int *__restrict__ celldofs,
                                            const double *__restrict__ EA, int64_t c0, int64_t ea_si, int ea_sc) {
    const unsigned c = p / (unsigned)NV;
    const int a = (int)(p - c * (unsigned)NV);
    const int *cd = celldofs + (int64_t)c * NV;
    const double *ke = EA + ((int64_t)c - c0) * ea_sc + a * ea_si;   // row a of the cell's matrix: (NV*NV, NV) or (NV, plane)
    if constexpr (NV % 4 == 0) {
#pragma unroll
        for (int b = 0; b < NV; b += 4) {
            const int4 d4 = *reinterpret_cast<const int4 *>(cd + b);
            g.dof[b] = d4.x; g.dof[b + 1] = d4.y; g.dof[b + 2] = d4.z; g.dof[b + 3] = d4.w;
        }
#pragma unroll
        for (int b = 0; b < NV; b += 2) {
            const double2 k2 = *reinterpret_cast<const double2 *>(ke + b);
            g.kv[b] = k2.x; g.kv[b + 1] = k2.y;
        }
    } else {
#pragma unroll
        for (int b = 0; b < NV; b++) {
            g.dof[b] = cd[b];
            g.kv[b] = ke[b];
        }
    }
}

// top_step = largest power of two <= maxw.  The NV lower-bound searches of one cell run in lock step (uniform
// trip count, no divergence), probing the slice's column ids staged in shared memory; the next cell's dof ids and
// matrix row are in flight while the current one is searched.
template <int NV>
__global__ void __launch_bounds__(256)
    k_gather_rows(const int64_t *__restrict__ adjptr, const unsigned *__restrict__ adj, const int *__restrict__ celldofs,
                  const double *__restrict__ EA, int64_t c0, int64_t ea_si, int ea_sc, SellView S, int64_t slice0, int64_t slice1,
                  int maxw, int top_step, int wlo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    double *acc = reinterpret_cast<double *>(smem_raw) + (size_t)warp * maxw * TB_SLICE + lane;
    int *scol = reinterpret_cast<int *>(smem_raw + (size_t)nwb * maxw * TB_SLICE * sizeof(double)) + (size_t)warp * maxw * TB_SLICE + lane;
    for (int64_t s = slice0 + (int64_t)blockIdx.x * nwb + warp; s < slice1; s += (int64_t)gridDim.x * nwb) {
        const int64_t base = S.slice_ptr[s];
        const int w = (int)((S.slice_ptr[s + 1] - base) >> 5);
        if (w <= wlo || w > maxw) continue;      // this launch handles slices with wlo < width <= maxw (warp-uniform)
        const int *gcol = S.col + base + lane;
        for (int j = 0; j < w; j++) {
            acc[j * TB_SLICE] = 0.0;
            scol[j * TB_SLICE] = gcol[(int64_t)j * TB_SLICE];
        }
        const int64_t r = s * TB_SLICE + lane;
        if (r < S.nrows) {
            const int len = (int)(S.rowptr[r + 1] - S.rowptr[r]);
            int64_t q = adjptr[r];
            const int64_t q1 = adjptr[r + 1];
            GatherCell<NV> cur, nxt;
            if (q < q1) gather_load<NV>(nxt, adj[q], celldofs, EA, c0, ea_si, ea_sc);
            for (; q < q1; q++) {
                cur = nxt;
                if (q + 1 < q1) gather_load<NV>(nxt, adj[q + 1], celldofs, EA, c0, ea_si, ea_sc);
                int lo[NV];
#pragma unroll
                for (int b = 0; b < NV; b++) lo[b] = 0;
                for (int step = top_step; step > 0; step >>= 1) {
#pragma unroll
                    for (int b = 0; b < NV; b++) {
                        const int t = lo[b] + step;
                        if (t <= len && scol[(t - 1) * TB_SLICE] < cur.dof[b]) lo[b] = t;
                    }
                }
#pragma unroll
                for (int b = 0; b < NV; b++) acc[lo[b] * TB_SLICE] += cur.kv[b];
            }
        }
        double *dst = S.val + base + lane;
        for (int j = 0; j < w; j++) dst[(int64_t)j * TB_SLICE] = acc[j * TB_SLICE];
    }
}

// ---- phase 2 without searches: slot table ---------------------------------------------------------------------
// Where a cell's dofs sit in a row never changes: computed once per (mesh, pattern) -- one byte per (adjacency entry, local
// dof), nv * nv bytes per cell -- the NV lower-bound searches per (row, cell) and the staging of the slice's column ids
// disappear from every later sweep, and so do the celldofs loads.  The row image alone needs half the shared memory, so
// twice as many warps are resident.  Same additions in the same order: bitwise the searching kernel.  Not built when a
// row is wider than 255 entries or the table would not fit the budget (then the searching kernel runs).
template <int NV>
__global__ void __launch_bounds__(256)
    k_adjpos_build(const int64_t *__restrict__ adjptr, const unsigned *__restrict__ adj, const int *__restrict__ celldofs,
                   SellView S, unsigned char *__restrict__ adjpos) {
    constexpr int PS = NV == 3 ? 4 : NV;               // bytes per adjacency entry
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < S.nrows; r += (int64_t)gridDim.x * blockDim.x) {
        const int len = (int)(S.rowptr[r + 1] - S.rowptr[r]);
        const int *rc = S.col + S.slice_ptr[r >> 5] + (r & 31);
        const int64_t q1 = adjptr[r + 1];
        for (int64_t q = adjptr[r]; q < q1; q++) {
            const int *cd = celldofs + (int64_t)(adj[q] / (unsigned)NV) * NV;
#pragma unroll
            for (int b = 0; b < NV; b++) {
                const int d = cd[b];
                int lo = 0, hi = len;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (rc[(int64_t)mid * TB_SLICE] < d) lo = mid + 1; else hi = mid;
                }
                adjpos[q * PS + b] = (unsigned char)lo;
            }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(256)
    k_gather_rows_pos(const int64_t *__restrict__ adjptr, const unsigned *__restrict__ adj, const unsigned char *__restrict__ adjpos,
                      const double *__restrict__ EA, int64_t c0, int64_t ea_si, int ea_sc, SellView S, int64_t slice0,
                      int64_t slice1, int maxw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PS = NV == 3 ? 4 : NV;
    constexpr int GB = 4;                              // adjacency entries in flight per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    double *acc = reinterpret_cast<double *>(smem_raw) + (size_t)warp * maxw * TB_SLICE + lane;
    struct Entry {
        double kv[NV];
        unsigned char pos[PS];
    };
    auto load = [&](Entry &e, int64_t q) {
        const unsigned p = adj[q];
        const unsigned c = p / (unsigned)NV;
        const int a = (int)(p - c * (unsigned)NV);
        const double *ke = EA + ((int64_t)c - c0) * ea_sc + a * ea_si;
        if constexpr (NV % 2 == 0) {
#pragma unroll
            for (int b = 0; b < NV; b += 2) {
                const double2 k2 = *reinterpret_cast<const double2 *>(ke + b);
                e.kv[b] = k2.x; e.kv[b + 1] = k2.y;
            }
        } else {
#pragma unroll
            for (int b = 0; b < NV; b++) e.kv[b] = ke[b];
        }
        if constexpr (PS == 8) *reinterpret_cast<uint2 *>(e.pos) = *reinterpret_cast<const uint2 *>(adjpos + q * 8);
        else *reinterpret_cast<unsigned *>(e.pos) = *reinterpret_cast<const unsigned *>(adjpos + q * 4);
    };
    for (int64_t s = slice0 + (int64_t)blockIdx.x * nwb + warp; s < slice1; s += (int64_t)gridDim.x * nwb) {
        const int64_t base = S.slice_ptr[s];
        const int w = (int)((S.slice_ptr[s + 1] - base) >> 5);
        for (int j = 0; j < w; j++) acc[j * TB_SLICE] = 0.0;
        const int64_t r = s * TB_SLICE + lane;
        if (r < S.nrows) {
            // GB adjacency entries at a time: their (cell, a) words first, then all GB matrix rows and slot words -- two
            // round trips per batch instead of two per entry -- and the additions in adjacency order as always
            const int64_t q1 = adjptr[r + 1];
            for (int64_t q = adjptr[r]; q < q1; q += GB) {
                Entry e[GB];
#pragma unroll
                for (int k = 0; k < GB; k++)
                    if (q + k < q1) load(e[k], q + k);
#pragma unroll
                for (int k = 0; k < GB; k++)
                    if (q + k < q1) {
#pragma unroll
                        for (int b = 0; b < NV; b++) acc[e[k].pos[b] * TB_SLICE] += e[k].kv[b];
                    }
            }
        }
        double *dst = S.val + base + lane;
        for (int j = 0; j < w; j++) dst[(int64_t)j * TB_SLICE] = acc[j * TB_SLICE];
    }
}

// b[r] = sum over the row's adjacency (ascending cell order) of EAb[cell][a]
__global__ void __launch_bounds__(256) k_gather_vec(const int64_t *__restrict__ adjptr, const unsigned *__restrict__ adj,
                                                    const double *__restrict__ EAb, int64_t p0, double *__restrict__ b,
                                                    int64_t r0, int64_t r1) {
    for (int64_t r = r0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += (int64_t)gridDim.x * blockDim.x) {
        double sum = 0.0;
        const int64_t q1 = adjptr[r + 1];
        for (int64_t q = adjptr[r]; q < q1; q++) sum += EAb[(int64_t)adj[q] - p0];
        b[r] = sum;
    }
}

static size_t element_smem(int nv, int dim, int nq) {
    return sizeof(double) * nq * (1 + nv + nv * dim) + sizeof(double) * nv * dim * AS_BLOCK;
}
static size_t element_matrix_smem(int nv, int dim, int nq) {
    return element_smem(nv, dim, nq) + sizeof(double) * (size_t)(AS_BLOCK / 32) * nv * nv * EA_PAD;
}

// Launch geometry of the split element kernel: columns per thread and elements per tile by element size, the batch of
// quadrature points sized so that two CTAs fit an SM (TB_ELEMENT_QB overrides; TB_ELEMENT_SPLIT=0 selects the
// one-thread-per-element kernel).
template <int NV> struct SplitGeom {
    static constexpr int CPT = NV % 2 == 0 ? 2 : 1;
    static constexpr int TE = NV == 8 ? 64 : NV == 4 ? 128 : 64;
    static constexpr int BLOCK = TE * (NV / CPT);
};
static int element_split_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TB_ELEMENT_SPLIT");
        v = e ? atoi(e) != 0 : 1;
    }
    return v;
}
static int ea_planes_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("TB_EA_PLANES");
        v = e ? atoi(e) != 0 : 1;
    }
    return v;
}
template <int NV, int DIM, int OP>
static size_t element_split_smem(int nq, int kind, int *qb_out) {
    constexpr int TE = SplitGeom<NV>::TE;
    const size_t fixed = sizeof(double) * ((size_t)nq * (1 + NV + NV * DIM) + (size_t)NV * DIM * TE);
    const size_t per_q = sizeof(double) * (size_t)TE * ((OP == 1 ? NV * DIM + 1 : 1) + (OP == 1 && kind >= 2 ? DIM * DIM : 0));
    int qb = (int)(((size_t)100 * 1024 - fixed) / per_q);   // two CTAs per SM (the kernel is compiled for <= 128 registers)
    if (qb > nq) qb = nq;
    if (qb < 1) qb = 1;
    qb = (nq + (nq + qb - 1) / qb - 1) / ((nq + qb - 1) / qb);   // balanced batches
    const char *e = getenv("TB_ELEMENT_QB");
    if (e && atoi(e) > 0) qb = atoi(e) < nq ? atoi(e) : nq;
    while (qb > 1 && fixed + per_q * qb > (size_t)200 * 1024) qb--;
    *qb_out = qb;
    return fixed + per_q * qb;
}

// slot table of the search-free gather (k_adjpos_build), cached on the mesh per pattern; adjpos_state says whether it applies
template <int NV>
static int32_t mesh_ensure_adjpos(tb_ctx *ctx, const tb_mesh *m, const tb_pattern *pat, const SellView &S) {
    static int enabled = -1;
    if (enabled < 0) {
        const char *e = getenv("TB_GATHER_POS");
        enabled = e ? atoi(e) != 0 : 1;
    }
    if (m->adjpos_state == 1 && m->adjpos_pat_uid == pat->uid) return TB_OK;
    if (m->adjpos_state == -1 && m->adjpos_pat_uid == pat->uid) return TB_OK;
    cudaFree(m->d_adjpos);
    m->d_adjpos = nullptr;
    m->adjpos_pat_uid = pat->uid;
    m->adjpos_state = -1;
    constexpr int PS = NV == 3 ? 4 : NV;
    const size_t bytes = (size_t)m->nadj * PS;
    const size_t cap = ctx->ea_budget_bytes / 4;          // C5 (6.4 GB of slots) keeps the searching kernel: setup runs once there
    if (!enabled || pat->max_width > 255 || bytes > cap || m->nadj == 0) return TB_OK;
    if (cudaMalloc(&m->d_adjpos, bytes + 16) != cudaSuccess) {
        cudaGetLastError();
        m->d_adjpos = nullptr;
        return TB_OK;
    }
    TB_LAUNCH(ctx, k_adjpos_build<NV>, tb_grid_for(ctx, S.nrows, 256, 8), 256, 0, m->d_adjptr, m->d_adj, m->d_celldofs, S, m->d_adjpos);
    m->adjpos_state = 1;
    return TB_OK;
}

template <int NV, int DIM, int OP>
static int32_t gather_bilinear_t(tb_ctx *ctx, const tb_mesh *m, const tb_elem_tables *d_T, int nq, double rho, int kind,
                                 const double *d_data, double cmchi, const tb_pattern *pat, const SellView &S,
                                 const GatherPlan &plan, double *EA) {
    using SG = SplitGeom<NV>;
    const bool split = element_split_enabled();
    int qb = nq;
    const size_t smem1 = split ? element_split_smem<NV, DIM, OP>(nq, kind, &qb) : element_matrix_smem(NV, DIM, nq);
    const int block1 = split ? SG::BLOCK : AS_BLOCK, tile1 = split ? SG::TE : AS_BLOCK;
    int per_sm = 1;
    if (split) {
        TB_CUDA(cudaFuncSetAttribute((k_element_matrices_split<NV, DIM, OP, SG::CPT, SG::TE>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k_element_matrices_split<NV, DIM, OP, SG::CPT, SG::TE>), block1, smem1);
    } else {
        TB_CUDA(cudaFuncSetAttribute(k_element_matrices<NV, DIM, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_element_matrices<NV, DIM, OP>, AS_BLOCK, smem1);
    }
    if (per_sm < 1) per_sm = 1;
    // Two width classes: the bulk of the slices (<= TB_TMA_WCAP entries per row) and the few wide ones a high-valence
    // vertex produces (LV apex: 2*nc + 3).  Sizing the shared-memory row image for the widest slice would leave one
    // warp per CTA for the whole mesh, so each class gets its own launch geometry.
    struct GatherGeom {
        int maxw, wlo, warps, top_step, per_sm;
        size_t smem;
    } gg[2];
    int ngeom = 0;
    auto make_geom = [&](int maxw, int wlo) -> int32_t {
        GatherGeom &G = gg[ngeom++];
        G.maxw = maxw < 1 ? 1 : maxw;
        G.wlo = wlo;
        const size_t per_warp = (size_t)G.maxw * TB_SLICE * (sizeof(double) + sizeof(int));
        int warps = (int)((100 * 1024) / per_warp);
        G.warps = warps > 8 ? 8 : warps < 1 ? 1 : warps;
        G.smem = (size_t)G.warps * per_warp;
        G.top_step = 1;
        while (G.top_step * 2 <= G.maxw) G.top_step *= 2;
        G.per_sm = 1;
        return TB_OK;
    };
    if (pat->n_wide > 0 && pat->max_width_tma > 0) {
        make_geom(pat->max_width_tma, 0);
        make_geom(pat->max_width, pat->max_width_tma);
    } else {
        make_geom(pat->max_width, 0);
    }
    size_t smem_max = 0;
    for (int k = 0; k < ngeom; k++) smem_max = gg[k].smem > smem_max ? gg[k].smem : smem_max;
    TB_CUDA(cudaFuncSetAttribute(k_gather_rows<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    for (int k = 0; k < ngeom; k++) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&gg[k].per_sm, k_gather_rows<NV>, gg[k].warps * 32, gg[k].smem);
        if (gg[k].per_sm < 1) gg[k].per_sm = 1;
    }
    TB_TRY(mesh_ensure_adjpos<NV>(ctx, m, pat, S));
    const bool use_pos = m->adjpos_state == 1 && m->adjpos_pat_uid == pat->uid;
    int pos_warps = 8, pos_per_sm = 1;
    size_t pos_smem = 0;
    if (use_pos) {
        const size_t per_warp = (size_t)(pat->max_width < 1 ? 1 : pat->max_width) * TB_SLICE * sizeof(double);
        const int fit = (int)((100 * 1024) / per_warp);
        pos_warps = fit > 8 ? 8 : fit < 1 ? 1 : fit;
        pos_smem = (size_t)pos_warps * per_warp;
        TB_CUDA(cudaFuncSetAttribute(k_gather_rows_pos<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pos_smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pos_per_sm, k_gather_rows_pos<NV>, pos_warps * 32, pos_smem);
        if (pos_per_sm < 1) pos_per_sm = 1;
    }
    const int nch = (int)plan.cmin.size();
    for (int ch = 0; ch < nch; ch++) {
        const int64_t r0 = (int64_t)ch * plan.rows_per_chunk;
        const int64_t r1 = r0 + plan.rows_per_chunk < pat->nrows ? r0 + plan.rows_per_chunk : pat->nrows;
        const int64_t s0 = r0 / TB_SLICE, s1 = (r1 + TB_SLICE - 1) / TB_SLICE;
        int64_t c0 = 0, ea_si = NV;
        int ea_sc = NV * NV;
        if (plan.cmax[ch] >= plan.cmin[ch]) {
            c0 = plan.cmin[ch];
            const int64_t c1 = (int64_t)plan.cmax[ch] + 1;
            if (split && ea_planes_enabled()) {
                ea_si = (c1 - c0) * NV;
                ea_sc = NV;
            }
            const int64_t ntiles = (c1 - c0 + tile1 - 1) / tile1;
            const int grid = (int)(ntiles < (int64_t)ctx->sm_count * per_sm ? ntiles : (int64_t)ctx->sm_count * per_sm);
            if (split)
                TB_LAUNCH(ctx, (k_element_matrices_split<NV, DIM, OP, SG::CPT, SG::TE>), grid, block1, smem1, m->d_conn, m->d_coords,
                          c0, c1, d_T, nq, qb, rho, kind, d_data, cmchi, EA, ea_si, ea_sc);
            else
                TB_LAUNCH(ctx, (k_element_matrices<NV, DIM, OP>), grid, AS_BLOCK, smem1, m->d_conn, m->d_coords, c0, c1, d_T, nq,
                          rho, kind, d_data, cmchi, EA);
        }
        if (use_pos) {
            const int64_t need = (s1 - s0 + pos_warps - 1) / pos_warps;
            const int grid2 = (int)(need < (int64_t)ctx->sm_count * pos_per_sm ? need : (int64_t)ctx->sm_count * pos_per_sm);
            if (grid2 > 0)
                TB_LAUNCH(ctx, k_gather_rows_pos<NV>, grid2, pos_warps * 32, pos_smem, m->d_adjptr, m->d_adj, m->d_adjpos, EA, c0, ea_si,
                          ea_sc, S, s0, s1, pat->max_width < 1 ? 1 : pat->max_width);
            continue;
        }
        for (int k = 0; k < ngeom; k++) {
            const GatherGeom &G = gg[k];
            const int64_t need = (s1 - s0 + G.warps - 1) / G.warps;
            const int grid2 = (int)(need < (int64_t)ctx->sm_count * G.per_sm ? need : (int64_t)ctx->sm_count * G.per_sm);
            if (grid2 > 0)
                TB_LAUNCH(ctx, k_gather_rows<NV>, grid2, G.warps * 32, G.smem, m->d_adjptr, m->d_adj, m->d_celldofs, EA, c0, ea_si,
                          ea_sc, S, s0, s1, G.maxw, G.top_step, G.wlo);
        }
    }
    return TB_OK;
}

static size_t assembly_smem(int nv, int dim, int nq) {
    return sizeof(double) * nq * (1 + nv + nv * dim) + sizeof(double) * nv * dim * AS_BLOCK + sizeof(int) * nv * AS_BLOCK;
}

// Quadrature/shape tables live on the device for the lifetime of the context: one upload per (cell type, order).
int32_t tb_get_tables(tb_ctx *ctx, int celltype, int qorder, const tb_elem_tables **d_T, int *nq) {
    if (celltype < 0 || celltype > 3 || qorder < 1 || qorder > 4)
        return tb_fail(TB_ERR_UNSUPPORTED, "assembly: quadrature order %d not available for cell type %d", qorder, celltype);
    if (!ctx->d_tables[celltype][qorder]) {
        tb_elem_tables T;
        memset(&T, 0, sizeof(T));
        if (tb_build_tables(celltype, qorder, &T))
            return tb_fail(TB_ERR_UNSUPPORTED, "assembly: quadrature order %d not available for cell type %d", qorder, celltype);
        tb_elem_tables *d = nullptr;
        TB_CUDA(cudaMalloc(&d, sizeof(T)));
        // stream-ordered: the context stream is non-blocking, a plain cudaMemcpy is not ordered against its kernels
        if (cudaMemcpyAsync(d, &T, sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            cudaFree(d);
            return tb_fail(TB_ERR_CUDA, "assembly: table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        ctx->d_tables[celltype][qorder] = d;
        ctx->tables_nq[celltype][qorder] = T.nq;
    }
    *d_T = ctx->d_tables[celltype][qorder];
    if (nq) *nq = ctx->tables_nq[celltype][qorder];
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
    A->version++;
    TB_REQUIRE(A->pat->nrows == mesh->ndofs_owned && A->pat->ncols == mesh->ndofs,
               "assemble: operator is %lld x %lld but the mesh has %lld owned / %lld total dofs", (long long)A->pat->nrows,
               (long long)A->pat->ncols, (long long)mesh->ndofs_owned, (long long)mesh->ndofs);
    if (op == 1) {   // argument checks before anything is allocated
        const int64_t need = kind == TB_D_SCALAR ? 1 : kind == TB_D_TENSOR ? mesh->dim * mesh->dim
                             : kind == TB_D_CELL_TENSOR ? mesh->ncells * mesh->dim * mesh->dim : 3 + mesh->ncells * mesh->nv * 9;
        TB_REQUIRE(kind >= TB_D_SCALAR && kind <= TB_D_CELL_TENSOR, "tb_assemble_diffusion: unknown coefficient kind %d", kind);
        TB_REQUIRE(kind != TB_D_SPECTRAL || mesh->dim == 3, "tb_assemble_diffusion: spectral coefficient needs a 3D mesh");
        TB_REQUIRE(data && ndata == need, "tb_assemble_diffusion: coefficient kind %d needs %lld doubles, got %lld", kind,
                   (long long)need, (long long)ndata);
        TB_REQUIRE(cmchi != 0.0, "tb_assemble_diffusion: Cm*chi must be non-zero");
    }
    TB_DEV(ctx);
    const tb_elem_tables *d_T = nullptr;
    int nq = 0;
    TB_TRY(tb_get_tables(ctx, mesh->celltype, qorder, &d_T, &nq));
    // Constant coefficients (a scalar or one tensor) go through a small buffer owned by the context: the upload is
    // stream-ordered (the driver stages pageable host memory before cudaMemcpyAsync returns), nothing is allocated or freed
    // and the call does not block.  Coefficient FIELDS (per-cell data) get a buffer for the duration of the call, which then
    // has to wait for its kernels.
    double *d_data = nullptr;
    const bool own_data = op == 1 && ndata > 16;
    if (op == 1) {
        if (own_data) {
            if (cudaMalloc(&d_data, sizeof(double) * (size_t)ndata) != cudaSuccess) {
                cudaGetLastError();
                return tb_fail(TB_ERR_NOMEM, "tb_assemble_diffusion: cannot allocate %lld coefficient values", (long long)ndata);
            }
        } else {
            d_data = ctx->d_dconst;
        }
        if (cudaMemcpyAsync(d_data, data, sizeof(double) * (size_t)ndata, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            if (own_data) cudaFree(d_data);
            return tb_fail(TB_ERR_CUDA, "tb_assemble_diffusion: coefficient upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    SellView S{A->pat->d_rowptr, A->pat->d_slice_ptr, A->pat->d_col, A->d_val, A->pat->nrows};
    int32_t st = TB_OK;
    // ---- mode 2: element matrices + ordered gather, when the scratch fits ----
    bool gathered = false;
    if (ctx->assembly_mode == 2 && A->pat->max_width <= TB_MAXROW) {
        const GatherPlan *planp = nullptr;
        bool ok = false;
        double *EA = nullptr;
        st = mesh_ensure_adjacency(ctx, mesh);
        if (st == TB_OK) {
            const int64_t per_cell = (int64_t)mesh->nv * mesh->nv * (int64_t)sizeof(double);
            size_t budget = ctx->ea_budget_bytes;
            // the scratch of an earlier call that is large enough for the whole mesh settles it; otherwise ask the driver how
            // much memory there is (cudaMemGetInfo costs a fraction of a millisecond: not on every call)
            if (ctx->ea_bytes < (size_t)mesh->ncells * (size_t)per_cell || budget < ctx->ea_bytes) {
                size_t free_b = 0, total_b = 0;
                cudaMemGetInfo(&free_b, &total_b);
                free_b += ctx->ea_bytes;                 // what the cached scratch holds would be reused
                if (budget > free_b / 2) budget = free_b / 2;
            }
            st = plan_gather(ctx, mesh, (int64_t)(budget / (size_t)per_cell), 0, &planp);
            ok = st == TB_OK && planp->ok;
            if (ok) {
                // the scratch is cached on the context (cudaMalloc/cudaFree of GBs costs 10-80 ms per call, measured);
                // tb_assembly_release_scratch gives it back
                const size_t bytes = (size_t)(planp->max_cells > 0 ? planp->max_cells : 1) * (size_t)per_cell;
                if (ctx->ea_bytes < bytes) {
                    cudaFree(ctx->d_ea);
                    ctx->d_ea = nullptr;
                    ctx->ea_bytes = 0;
                    if (cudaMalloc(&ctx->d_ea, bytes) == cudaSuccess) ctx->ea_bytes = bytes;
                    else {
                        cudaGetLastError();
                        ok = false;
                    }
                }
                EA = static_cast<double *>(ctx->d_ea);
            }
        }
        if (st == TB_OK && ok) {
            const GatherPlan &plan = *planp;
#define DISPATCH_G(NV, DIM)                                                                                         \
    st = op == 0 ? gather_bilinear_t<NV, DIM, 0>(ctx, mesh, d_T, nq, rho, kind, d_data, cmchi, A->pat, S, plan, EA) \
                 : gather_bilinear_t<NV, DIM, 1>(ctx, mesh, d_T, nq, rho, kind, d_data, cmchi, A->pat, S, plan, EA)
            switch (mesh->celltype) {
            case TB_QUAD4: DISPATCH_G(4, 2); break;
            case TB_HEX8: DISPATCH_G(8, 3); break;
            case TB_TRI3: DISPATCH_G(3, 2); break;
            default: DISPATCH_G(4, 3); break;
            }
#undef DISPATCH_G
            gathered = true;
            ctx->assembly_last_mode = 2;
            ctx->assembly_last_chunks = (int)plan.cmin.size();
        }
        if (gathered || st != TB_OK) {
            if (!own_data) return st;         // stream-ordered: a failing kernel surfaces at the next blocking call on this stream
            cudaError_t e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_data);
            if (st != TB_OK) return st;
            if (e != cudaSuccess) return tb_fail(TB_ERR_CUDA, "assemble (gather): kernel failed: %s", cudaGetErrorString(e));
            return TB_OK;
        }
    }
    // ---- mode 0: atomic scatter ----
    ctx->assembly_last_mode = 0;
    ctx->assembly_last_chunks = 0;
    TB_CUDA(cudaMemsetAsync(A->d_val, 0, sizeof(double) * (size_t)A->pat->sell_len, ctx->stream));
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
    if (!own_data) return st;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
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
static int32_t launch_source(tb_ctx *ctx, const tb_mesh *m, const tb_elem_tables *d_T, int nq, int kind, const SrcParams &d_prm,
                             double t, const double *d_fq, double *b) {
    const size_t smem = assembly_smem(NV, DIM, nq);
    TB_CUDA(cudaFuncSetAttribute(k_assemble_source<NV, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (m->ncells + AS_BLOCK - 1) / AS_BLOCK;
    int grid = (int)(ntiles < (int64_t)ctx->sm_count * 4 ? ntiles : (int64_t)ctx->sm_count * 4);
    TB_LAUNCH(ctx, (k_assemble_source<NV, DIM>), grid, AS_BLOCK, smem, m->d_conn, m->d_celldofs, m->d_coords, m->ncells,
              d_T, nq, kind, d_prm, t, d_fq, b, m->ndofs_owned);
    return TB_OK;
}

template <int NV, int DIM>
static int32_t launch_element_vectors(tb_ctx *ctx, const tb_mesh *m, const tb_elem_tables *d_T, int nq, int kind,
                                      const SrcParams &d_prm, double t, const double *d_fq, int64_t c0, int64_t c1, double *EAb) {
    const size_t smem = element_smem(NV, DIM, nq);
    TB_CUDA(cudaFuncSetAttribute(k_element_vectors<NV, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (c1 - c0 + AS_BLOCK - 1) / AS_BLOCK;
    const int grid = (int)(ntiles < (int64_t)ctx->sm_count * 4 ? ntiles : (int64_t)ctx->sm_count * 4);
    TB_LAUNCH(ctx, (k_element_vectors<NV, DIM>), grid, AS_BLOCK, smem, m->d_conn, m->d_coords, c0, c1, d_T, nq, kind, d_prm, t,
              d_fq, EAb);
    return TB_OK;
}

static int32_t assemble_source(tb_ctx *ctx, const tb_mesh *mesh, int qorder, int kind, const double *prm, int nprm,
                               double t, const double *fq, tb_vec *b, int bcol, const tb_src_program *prog = nullptr) {
    TB_REQUIRE(ctx && mesh && b, "tb_assemble_source: NULL argument");
    TB_REQUIRE(bcol >= 0 && bcol < b->ncols && b->n >= mesh->ndofs_owned, "tb_assemble_source: vector too small");
    TB_REQUIRE(fq || prog || (kind >= TB_SRC_NONE && kind <= TB_SRC_ENDO), "tb_assemble_source: unknown source kind %d", kind);
    TB_REQUIRE(nprm >= 0 && nprm <= 8, "tb_assemble_source: at most 8 parameters");
    TB_DEV(ctx);
    const tb_elem_tables *d_T = nullptr;
    int nq = 0;
    TB_TRY(tb_get_tables(ctx, mesh->celltype, qorder, &d_T, &nq));
    SrcParams d_prm;                      // travels as a kernel argument: no allocation, no copy, nothing to free
    memset(&d_prm, 0, sizeof(d_prm));
    for (int i = 0; i < 8; i++) d_prm.p[i] = i < nprm ? prm[i] : 0.0;
    if (prog) {
        d_prm.prog = *prog;
        kind = TB_SRC_PROGRAM;
    }
    double *d_fq = nullptr;
    if (fq) {
        const size_t bytes = sizeof(double) * (size_t)(mesh->ncells * nq);
        TB_CUDA(cudaMalloc(&d_fq, bytes));
        TB_CUDA(cudaMemcpyAsync(d_fq, fq, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    double *bp = b->d + (size_t)bcol * b->ld;
    int32_t st = TB_OK;
    bool gathered = false;
    if (ctx->assembly_mode == 2 && (fq || kind != TB_SRC_NONE)) {
        // ordered gather: element vectors into the cached scratch, then b[r] = sum over the row's adjacency
        const GatherPlan *planp = nullptr;
        bool ok = false;
        st = mesh_ensure_adjacency(ctx, mesh);
        if (st == TB_OK) {
            size_t budget = ctx->ea_budget_bytes < ((size_t)2 << 30) ? ctx->ea_budget_bytes : ((size_t)2 << 30);
            const int64_t per_cell = (int64_t)mesh->nv * (int64_t)sizeof(double);
            st = plan_gather(ctx, mesh, (int64_t)(budget / (size_t)per_cell), 1, &planp);
            ok = st == TB_OK && planp->ok;
            if (ok) {
                const size_t bytes = (size_t)(planp->max_cells > 0 ? planp->max_cells : 1) * (size_t)per_cell;
                if (ctx->ea_bytes < bytes) {
                    cudaFree(ctx->d_ea);
                    ctx->d_ea = nullptr;
                    ctx->ea_bytes = 0;
                    if (cudaMalloc(&ctx->d_ea, bytes) == cudaSuccess) ctx->ea_bytes = bytes;
                    else {
                        cudaGetLastError();
                        ok = false;
                    }
                }
            }
        }
        if (st == TB_OK && ok) {
            const GatherPlan &plan = *planp;
            double *EAb = static_cast<double *>(ctx->d_ea);
            const int nch = (int)plan.cmin.size();
            const int64_t nrows = mesh->ndofs_owned;
            for (int ch = 0; ch < nch && st == TB_OK; ch++) {
                const int64_t r0 = (int64_t)ch * plan.rows_per_chunk;
                const int64_t r1 = r0 + plan.rows_per_chunk < nrows ? r0 + plan.rows_per_chunk : nrows;
                int64_t c0 = 0;
                if (plan.cmax[ch] >= plan.cmin[ch]) {
                    c0 = plan.cmin[ch];
                    const int64_t c1 = (int64_t)plan.cmax[ch] + 1;
                    switch (mesh->celltype) {
                    case TB_QUAD4: st = launch_element_vectors<4, 2>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, c0, c1, EAb); break;
                    case TB_HEX8: st = launch_element_vectors<8, 3>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, c0, c1, EAb); break;
                    case TB_TRI3: st = launch_element_vectors<3, 2>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, c0, c1, EAb); break;
                    default: st = launch_element_vectors<4, 3>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, c0, c1, EAb); break;
                    }
                }
                if (st == TB_OK && r1 > r0)
                    TB_LAUNCH(ctx, k_gather_vec, tb_grid_for(ctx, r1 - r0, 256, 8), 256, 0, mesh->d_adjptr, mesh->d_adj, EAb,
                              c0 * mesh->nv, bp, r0, r1);
            }
            // rows past the owned block (ghost slots of a partitioned vector) stay zero like in mode 0
            if (b->n > nrows) TB_CUDA(cudaMemsetAsync(bp + nrows, 0, sizeof(double) * (size_t)(b->n - nrows), ctx->stream));
            gathered = true;
            ctx->assembly_last_mode = 2;
            ctx->assembly_last_chunks = nch;
        }
    }
    if (!gathered && st == TB_OK) {
        ctx->assembly_last_mode = 0;
        ctx->assembly_last_chunks = 0;
        TB_CUDA(cudaMemsetAsync(bp, 0, sizeof(double) * (size_t)b->n, ctx->stream));
        if (fq || kind != TB_SRC_NONE) {
            switch (mesh->celltype) {
            case TB_QUAD4: st = launch_source<4, 2>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
            case TB_HEX8: st = launch_source<8, 3>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
            case TB_TRI3: st = launch_source<3, 2>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
            default: st = launch_source<4, 3>(ctx, mesh, d_T, nq, kind, d_prm, t, d_fq, bp); break;
            }
        }
    }
    if (!d_fq) return st;                 // built-in families: fully stream-ordered, the step that consumes b follows on the same stream
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
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

extern "C" int32_t tb_assemble_source_program(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, const int32_t *code,
                                              int32_t ncode, const double *consts, int32_t nconsts, double t, tb_vec *b,
                                              int32_t bcol) {
    TB_REQUIRE(mesh && code && (consts || nconsts == 0), "tb_assemble_source_program: NULL argument");
    tb_src_program P;
    memset(&P, 0, sizeof(P));
    const int rc = tb_program_build(code, ncode, consts, nconsts, mesh->dim, &P);
    static const char *why[] = {"", "too long (or too many constants)", "unknown opcode", "operand out of range",
                                "stack underflow or overflow", "must leave exactly one value"};
    if (rc) return tb_fail(TB_ERR_INVALID, "tb_assemble_source_program: invalid program: %s", why[-rc]);
    return assemble_source(ctx, mesh, qorder, TB_SRC_PROGRAM, nullptr, 0, t, nullptr, b, bcol, &P);
}

extern "C" int32_t tb_assemble_source_qp(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, const double *fq, tb_vec *b,
                                         int32_t bcol) {
    TB_REQUIRE(fq, "tb_assemble_source_qp: fq is NULL");
    return assemble_source(ctx, mesh, qorder, 0, nullptr, 0, 0.0, fq, b, bcol);
}

extern "C" int32_t tb_assembly_release_scratch(tb_ctx *ctx) {
    TB_REQUIRE(ctx, "tb_assembly_release_scratch: ctx is NULL");
    TB_DEV(ctx);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_ea);
    ctx->d_ea = nullptr;
    ctx->ea_bytes = 0;
    return TB_OK;
}
