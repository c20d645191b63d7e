// CSR operators stored as a sliced-ELL (SELL-32) image of the reference's CSR pattern.
//   create_system_matrix            src/solver/interface.jl:159-173 (Ferrite allocate_matrix)
//   mul!(y, A, x)                   src/utils.jl:210-231
//   nz(A) = nz(M) - dt nz(K)        src/solver/time/euler.jl:104-116
//
// HBM layout: rows are grouped in slices of 32 consecutive rows (one warp).  Slice s stores
// w_s = max row length in the slice entries per row, column-major inside the slice: entry j of row r
// sits at slice_ptr[s] + j*32 + (r & 31).  A warp therefore reads 256 contiguous bytes of values and
// 128 contiguous bytes of column ids per j, and each lane accumulates its row strictly left to right
// -- the same summation order as the reference's row loop, so SpMV results are bitwise identical.
// Padding entries carry value 0 and the row's own index as column.
#include <algorithm>
#include "tb_internal.cuh"
#include "tb_spmv.cuh"
#include <cub/cub.cuh>

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

// ------------------------------------------------------------------------------------------------
// pattern construction
// ------------------------------------------------------------------------------------------------
// maxw[0] widest slice, maxw[1] widest slice not above TB_TMA_WCAP, maxw[2] number of slices above it
__global__ void k_slice_width(const int64_t *rowptr, int64_t nrows, int64_t nslices, int64_t *width32, int *maxw) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslices; s += (int64_t)gridDim.x * blockDim.x) {
        int64_t r0 = s * TB_SLICE, r1 = r0 + TB_SLICE < nrows ? r0 + TB_SLICE : nrows;
        int64_t w = 0;
        for (int64_t r = r0; r < r1; r++) {
            int64_t l = rowptr[r + 1] - rowptr[r];
            w = l > w ? l : w;
        }
        width32[s] = w * TB_SLICE;
        atomicMax(maxw, (int)w);
        if (w <= TB_TMA_WCAP) atomicMax(maxw + 1, (int)w);
        else atomicAdd(maxw + 2, 1);
    }
}

__global__ void k_collect_wide(const int64_t *slice_ptr, int64_t nslices, int64_t *out, int *count) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslices; s += (int64_t)gridDim.x * blockDim.x)
        if (((slice_ptr[s + 1] - slice_ptr[s]) >> 5) > TB_TMA_WCAP) out[atomicAdd(count, 1)] = s;
}

// one warp per slice: copies CSR column ids into the slice, pads with the row index
__global__ void k_sell_fill_cols(const int64_t *rowptr, const int *colidx, int64_t nrows, int64_t ncols, int64_t nslices,
                                 const int64_t *slice_ptr, int *sell_col) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        int64_t r = s * TB_SLICE + lane;
        int64_t base = slice_ptr[s];
        int w = (int)((slice_ptr[s + 1] - base) / TB_SLICE);
        int64_t p0 = 0;
        int len = 0;
        if (r < nrows) {
            p0 = rowptr[r];
            len = (int)(rowptr[r + 1] - p0);
        }
        int padcol = (int)(r < ncols ? r : 0);
        for (int j = 0; j < w; j++) sell_col[base + (int64_t)j * TB_SLICE + lane] = j < len ? colidx[p0 + j] : padcol;
    }
}

int32_t tb_exclusive_scan_i64(tb_ctx *ctx, const int64_t *in, int64_t *out, int64_t n) {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, ctx->stream));
    void *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, tmp_bytes));
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, ctx->stream));
    ctx->launches++;
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    return TB_OK;
}

// ---- compressed column stream -----------------------------------------------------------------------
// In a slice of 32 consecutive rows of a (locally) structured mesh most entry slots j have
// col = row + off_j with ONE offset for all 32 lanes.  Such a slot is stored as a single int32 (off_j)
// instead of 32 column ids; the others keep their explicit 32 ids.  Lossless, pattern-agnostic (an
// unstructured matrix simply ends up all-explicit), and it removes both a third of the SpMV's HBM
// traffic (C5: 348 -> ~244 B/row) and the col -> x load dependency.
// Stream of slice s, in ints, at ccol[cptr[s]]: hdr[w] (off_j, or TB_CCOL_EXPLICIT), padded to a multiple
// of 4 ints, then one 32-int block per explicit slot in slot order.
#define TB_CCOL_EXPLICIT INT32_MIN

// pass 0 (write == false): cnt[s] = ints of the slice's stream; pass 1: emit.  One warp per slice.
// Padding lanes (j >= row length, or row >= nrows) are wildcards: their value is 0, so their column may
// be anything valid; they are REWRITTEN in the uncompressed array to row + off_j so both images agree.
template <bool WRITE>
__global__ void k_ccol_build(const int64_t *rowptr, int64_t nrows, int64_t ncols, int64_t nslices, const int64_t *slice_ptr,
                             int *sell_col, int64_t *cnt, const int64_t *cptr, int *ccol, int *max_ints) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        const int64_t base = slice_ptr[s];
        const int w = (int)((slice_ptr[s + 1] - base) >> 5);
        const int64_t r = s * TB_SLICE + lane;
        const int len = r < nrows ? (int)(rowptr[r + 1] - rowptr[r]) : 0;
        const int hdr_ints = (w + 3) & ~3;
        int nexp = 0;
        int *out = WRITE ? ccol + cptr[s] : nullptr;
        for (int j = 0; j < w; j++) {
            const int c = sell_col[base + (int64_t)j * TB_SLICE + lane];
            const bool real = j < len;
            const unsigned realmask = __ballot_sync(0xffffffffu, real);
            const int first = realmask ? __ffs(realmask) - 1 : 0;        // w is the max row length, so realmask != 0
            const int off = realmask ? __shfl_sync(0xffffffffu, c - (int)r, first) : 0;
            const int64_t cand = r + off;
            const bool ok = real ? (c - (int)r == off) : (cand >= 0 && cand < ncols);
            const bool uniform = __all_sync(0xffffffffu, ok);
            if (WRITE) {
                if (uniform) {
                    if (lane == 0) out[j] = off;
                    if (!real) sell_col[base + (int64_t)j * TB_SLICE + lane] = (int)cand;
                } else {
                    if (lane == 0) out[j] = TB_CCOL_EXPLICIT;
                    out[hdr_ints + nexp * 32 + lane] = c;
                }
            }
            nexp += uniform ? 0 : 1;
        }
        if (WRITE) {
            if (lane < hdr_ints - w) out[w + lane] = 0;
        } else if (lane == 0) {
            const int ints = hdr_ints + nexp * 32;
            cnt[s] = ints;
            if (w <= TB_TMA_WCAP) atomicMax(max_ints, ints);   // wide slices are never staged (tb_spmv.cuh)
        }
    }
}

static int32_t build_ccol(tb_ctx *ctx, tb_pattern *p) {
    int64_t *cnt = nullptr;
    int *d_max = nullptr;
    TB_CUDA(cudaMalloc(&cnt, sizeof(int64_t) * (size_t)(p->nslices + 1)));
    TB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int64_t) * (size_t)(p->nslices + 1), ctx->stream));
    TB_CUDA(cudaMalloc(&d_max, sizeof(int)));
    TB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), ctx->stream));
    TB_CUDA(cudaMalloc(&p->d_cptr, sizeof(int64_t) * (size_t)(p->nslices + 1)));
    const int grid = ctx->sm_count * 8;
    TB_LAUNCH(ctx, (k_ccol_build<false>), grid, 256, 0, p->d_rowptr, p->nrows, p->ncols, p->nslices, p->d_slice_ptr, p->d_col,
              cnt, nullptr, nullptr, d_max);
    TB_TRY(tb_exclusive_scan_i64(ctx, cnt, p->d_cptr, p->nslices + 1));
    TB_CUDA(cudaMemcpy(&p->ccol_len, p->d_cptr + p->nslices, sizeof(int64_t), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMemcpy(&p->max_ccol_ints, d_max, sizeof(int), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMalloc(&p->d_ccol, sizeof(int) * (size_t)(p->ccol_len + 32)));
    TB_LAUNCH(ctx, (k_ccol_build<true>), grid, 256, 0, p->d_rowptr, p->nrows, p->ncols, p->nslices, p->d_slice_ptr, p->d_col,
              nullptr, p->d_cptr, p->d_ccol, nullptr);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(cnt);
    cudaFree(d_max);
    return TB_OK;
}

// builds slice_ptr + sell columns from device CSR (rowptr int64, colidx int32); takes ownership of rowptr
static int32_t pattern_from_device_csr(tb_ctx *ctx, int64_t nrows, int64_t ncols, int64_t nnz, int64_t *d_rowptr,
                                       const int *d_colidx, tb_pattern **out) {
    tb_pattern *p = new (std::nothrow) tb_pattern();
    if (!p) return tb_fail(TB_ERR_NOMEM, "tb_pattern: host allocation failed");
    p->uid = tb_next_uid();
    p->ctx = ctx;
    p->nrows = nrows;
    p->ncols = ncols;
    p->nnz = nnz;
    p->d_rowptr = d_rowptr;
    p->nslices = (nrows + TB_SLICE - 1) / TB_SLICE;
    *out = p;
    int64_t *width = nullptr;
    TB_CUDA(cudaMalloc(&width, sizeof(int64_t) * (size_t)(p->nslices + 1)));
    TB_CUDA(cudaMemsetAsync(width, 0, sizeof(int64_t) * (size_t)(p->nslices + 1), ctx->stream));
    TB_CUDA(cudaMalloc(&p->d_slice_ptr, sizeof(int64_t) * (size_t)(p->nslices + 1)));
    int *d_maxw = nullptr;
    int h_maxw[3] = {0, 0, 0};
    TB_CUDA(cudaMalloc(&d_maxw, sizeof(int) * 3));
    TB_CUDA(cudaMemsetAsync(d_maxw, 0, sizeof(int) * 3, ctx->stream));
    TB_LAUNCH(ctx, k_slice_width, tb_grid_for(ctx, p->nslices, 256, 8), 256, 0, d_rowptr, nrows, p->nslices, width, d_maxw);
    TB_CUDA(cudaMemcpyAsync(h_maxw, d_maxw, sizeof(int) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));   // the context stream is non-blocking: a plain cudaMemcpy would not wait for the kernel
    p->max_width = h_maxw[0];
    p->max_width_tma = h_maxw[1];
    p->n_wide = h_maxw[2];
    TB_TRY(tb_exclusive_scan_i64(ctx, width, p->d_slice_ptr, p->nslices + 1));
    TB_CUDA(cudaMemcpy(&p->sell_len, p->d_slice_ptr + p->nslices, sizeof(int64_t), cudaMemcpyDeviceToHost));
    cudaFree(width);
    if (p->n_wide > 0) {   // ids of the wide slices, ascending (collected in arbitrary order, sorted on the host: there are few)
        TB_CUDA(cudaMalloc(&p->d_wide_slices, sizeof(int64_t) * (size_t)p->n_wide));
        TB_CUDA(cudaMemsetAsync(d_maxw, 0, sizeof(int), ctx->stream));
        TB_LAUNCH(ctx, k_collect_wide, tb_grid_for(ctx, p->nslices, 256, 8), 256, 0, p->d_slice_ptr, p->nslices, p->d_wide_slices, d_maxw);
        std::vector<int64_t> hw((size_t)p->n_wide);
        TB_CUDA(cudaMemcpyAsync(hw.data(), p->d_wide_slices, sizeof(int64_t) * hw.size(), cudaMemcpyDeviceToHost, ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        std::sort(hw.begin(), hw.end());
        TB_CUDA(cudaMemcpyAsync(p->d_wide_slices, hw.data(), sizeof(int64_t) * hw.size(), cudaMemcpyHostToDevice, ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(d_maxw);
    TB_CUDA(cudaMalloc(&p->d_col, sizeof(int) * (size_t)(p->sell_len + 32)));
    TB_LAUNCH(ctx, k_sell_fill_cols, ctx->sm_count * 8, 256, 0, d_rowptr, d_colidx, nrows, ncols, p->nslices,
              p->d_slice_ptr, p->d_col);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    return build_ccol(ctx, p);
}

int32_t tb_pattern_release(tb_pattern *p) {
    if (!p) return TB_OK;
    if (--p->refcount > 0) return TB_OK;
    cudaFree(p->d_ccol);
    cudaFree(p->d_cptr);
    cudaFree(p->d_rowptr);
    cudaFree(p->d_slice_ptr);
    cudaFree(p->d_col);
    cudaFree(p->d_diag_slot);
    cudaFree(p->d_wide_slices);
    cudaFree(p->halo.d_send_rows);
    cudaFree(p->halo.d_sendbuf);
    delete p;
    return TB_OK;
}

uint64_t tb_next_uid() {
    static uint64_t next = 0;
    return ++next;
}

static int32_t csr_from_pattern(tb_pattern *p, tb_csr **out) {
    tb_csr *a = new (std::nothrow) tb_csr();
    if (!a) return tb_fail(TB_ERR_NOMEM, "tb_csr: host allocation failed");
    a->pat = p;
    a->uid = tb_next_uid();
    a->d_val = nullptr;
    cudaError_t e = cudaMalloc(&a->d_val, sizeof(double) * (size_t)(p->sell_len + 32));
    if (e != cudaSuccess) {
        delete a;
        return tb_fail(TB_ERR_NOMEM, "tb_csr: cudaMalloc of %lld values failed: %s", (long long)p->sell_len,
                       cudaGetErrorString(e));
    }
    TB_CUDA(cudaMemsetAsync(a->d_val, 0, sizeof(double) * (size_t)(p->sell_len + 32), p->ctx->stream));
    *out = a;
    return TB_OK;
}

__global__ void k_i64_to_i32_b(const int64_t *src, int *dst, int64_t n, int base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (int)(src[i] - base);
}
__global__ void k_sub_base(int64_t *a, int64_t n, int64_t base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        a[i] -= base;
}
// verifies strictly increasing columns per row and the column range
__global__ void k_check_csr(const int64_t *rowptr, const int *colidx, int64_t nrows, int64_t ncols, int *bad) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t a = rowptr[r], b = rowptr[r + 1];
        if (b < a) *bad = 1;
        for (int64_t k = a; k < b; k++) {
            int c = colidx[k];
            if (c < 0 || c >= ncols) *bad = 1;
            if (k > a && colidx[k - 1] >= c) *bad = 1;
        }
    }
}

extern "C" int32_t tb_csr_create(tb_ctx *ctx, int64_t nrows, int64_t ncols, const int64_t *rowptr,
                                 const int64_t *colidx, int32_t index_base, tb_csr **out) {
    TB_REQUIRE(ctx && rowptr && colidx && out, "tb_csr_create: NULL argument");
    TB_REQUIRE(nrows > 0 && ncols > 0 && ncols < INT32_MAX, "tb_csr_create: bad shape");
    TB_REQUIRE(index_base == 0 || index_base == 1, "tb_csr_create: index_base must be 0 or 1");
    TB_DEV(ctx);
    *out = nullptr;
    int64_t nnz = rowptr[nrows] - index_base;
    TB_REQUIRE(nnz > 0 && rowptr[0] == index_base, "tb_csr_create: rowptr does not start at index_base or nnz == 0");
    int64_t *d_rowptr = nullptr, *d_col64 = nullptr;
    int *d_col = nullptr, *d_bad = nullptr;
    TB_CUDA(cudaMalloc(&d_rowptr, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&d_col64, sizeof(int64_t) * (size_t)nnz));
    TB_CUDA(cudaMalloc(&d_col, sizeof(int) * (size_t)nnz));
    TB_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    TB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    TB_CUDA(cudaMemcpyAsync(d_rowptr, rowptr, sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaMemcpyAsync(d_col64, colidx, sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    if (index_base) TB_LAUNCH(ctx, k_sub_base, tb_grid_for(ctx, nrows + 1, 256, 8), 256, 0, d_rowptr, nrows + 1, (int64_t)index_base);
    TB_LAUNCH(ctx, k_i64_to_i32_b, tb_grid_for(ctx, nnz, 256, 8), 256, 0, d_col64, d_col, nnz, index_base);
    TB_LAUNCH(ctx, k_check_csr, tb_grid_for(ctx, nrows, 256, 8), 256, 0, d_rowptr, d_col, nrows, ncols, d_bad);
    int bad = 0;
    TB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_col64);
    cudaFree(d_bad);
    if (bad) {
        cudaFree(d_rowptr);
        cudaFree(d_col);
        return tb_fail(TB_ERR_INVALID, "tb_csr_create: columns must be strictly increasing per row and within [0,ncols)");
    }
    tb_pattern *p = nullptr;
    int32_t st = pattern_from_device_csr(ctx, nrows, ncols, nnz, d_rowptr, d_col, &p);
    cudaFree(d_col);
    if (st != TB_OK) {
        tb_pattern_release(p);
        return st;
    }
    st = csr_from_pattern(p, out);
    if (st != TB_OK) tb_pattern_release(p);
    return st;
}

// ---- device-side allocate_matrix(dh) ---------------------------------------------------------------
__global__ void k_adj_count(const int *celldofs, int64_t npos, int nrows, int *count) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (int64_t)gridDim.x * blockDim.x) {
        int d = celldofs[p];
        if (d < nrows) atomicAdd(&count[d], 1);
    }
}
__global__ void k_adj_fill(const int *celldofs, int64_t npos, int nv, int nrows, const int64_t *adjptr, int *cursor,
                           int *adj) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (int64_t)gridDim.x * blockDim.x) {
        int d = celldofs[p];
        if (d < nrows) adj[adjptr[d] + atomicAdd(&cursor[d], 1)] = (int)(p / nv);
    }
}
__global__ void k_i32_to_i64_c(const int *src, int64_t *dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// sorted-unique union of the dofs of all cells adjacent to row r; returns its length.
// cols lives in local memory; TB_MAXROW bounds it (overflow sets *err).
__device__ int row_union(int r, const int64_t *adjptr, const int *adj, const int *celldofs, int nv, int *cols, int *err) {
    int m = 0;
    for (int64_t q = adjptr[r]; q < adjptr[r + 1]; q++) {
        const int *cd = celldofs + (int64_t)adj[q] * nv;
        for (int a = 0; a < nv; a++) {
            int c = cd[a];
            int lo = 0, hi = m;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cols[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < m && cols[lo] == c) continue;
            if (m >= TB_MAXROW) {
                *err = 1;
                continue;
            }
            for (int k = m; k > lo; k--) cols[k] = cols[k - 1];
            cols[lo] = c;
            m++;
        }
    }
    return m;
}

__global__ void k_row_len(int nrows, const int64_t *adjptr, const int *adj, const int *celldofs, int nv, int64_t *rowlen,
                          int *err) {
    int cols[TB_MAXROW];
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x)
        rowlen[r] = row_union((int)r, adjptr, adj, celldofs, nv, cols, err);
}
__global__ void k_row_fill(int nrows, const int64_t *adjptr, const int *adj, const int *celldofs, int nv,
                           const int64_t *rowptr, int *colidx, int *err) {
    int cols[TB_MAXROW];
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        int m = row_union((int)r, adjptr, adj, celldofs, nv, cols, err);
        int64_t p0 = rowptr[r];
        for (int k = 0; k < m; k++) colidx[p0 + k] = cols[k];
    }
}

extern "C" int32_t tb_csr_create_from_mesh(tb_ctx *ctx, const tb_mesh *mesh, tb_csr **out) {
    TB_REQUIRE(ctx && mesh && out, "tb_csr_create_from_mesh: NULL argument");
    TB_DEV(ctx);
    *out = nullptr;
    const int nrows = (int)mesh->ndofs_owned;
    const int64_t npos = mesh->ncells * mesh->nv;
    int grid = ctx->sm_count * 8;
    int *count = nullptr, *adj = nullptr, *d_err = nullptr;
    int64_t *count64 = nullptr, *adjptr = nullptr, *rowlen = nullptr, *d_rowptr = nullptr;
    TB_CUDA(cudaMalloc(&count, sizeof(int) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&count64, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&adjptr, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&d_err, sizeof(int)));
    TB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    TB_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(nrows + 1), ctx->stream));
    TB_LAUNCH(ctx, k_adj_count, grid, 256, 0, mesh->d_celldofs, npos, nrows, count);
    TB_LAUNCH(ctx, k_i32_to_i64_c, grid, 256, 0, count, count64, (int64_t)nrows + 1);
    TB_TRY(tb_exclusive_scan_i64(ctx, count64, adjptr, (int64_t)nrows + 1));
    int64_t nadj = 0;
    TB_CUDA(cudaMemcpy(&nadj, adjptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMalloc(&adj, sizeof(int) * (size_t)(nadj + 1)));
    TB_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(nrows + 1), ctx->stream));
    TB_LAUNCH(ctx, k_adj_fill, grid, 256, 0, mesh->d_celldofs, npos, mesh->nv, nrows, adjptr, count, adj);
    // row lengths -> rowptr
    TB_CUDA(cudaMalloc(&rowlen, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMalloc(&d_rowptr, sizeof(int64_t) * (size_t)(nrows + 1)));
    TB_CUDA(cudaMemsetAsync(rowlen, 0, sizeof(int64_t) * (size_t)(nrows + 1), ctx->stream));
    TB_LAUNCH(ctx, k_row_len, tb_grid_for(ctx, nrows, 128, 16), 128, 0, nrows, adjptr, adj, mesh->d_celldofs, mesh->nv,
              rowlen, d_err);
    TB_TRY(tb_exclusive_scan_i64(ctx, rowlen, d_rowptr, (int64_t)nrows + 1));
    int64_t nnz = 0;
    int err = 0;
    TB_CUDA(cudaMemcpy(&nnz, d_rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost));
    TB_CUDA(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(rowlen);
    cudaFree(count);
    cudaFree(count64);
    if (err) {
        cudaFree(adj); cudaFree(adjptr); cudaFree(d_rowptr); cudaFree(d_err);
        return tb_fail(TB_ERR_UNSUPPORTED, "tb_csr_create_from_mesh: a row has more than %d nonzeros", TB_MAXROW);
    }
    int *d_col = nullptr;
    TB_CUDA(cudaMalloc(&d_col, sizeof(int) * (size_t)(nnz + 1)));
    TB_LAUNCH(ctx, k_row_fill, tb_grid_for(ctx, nrows, 128, 16), 128, 0, nrows, adjptr, adj, mesh->d_celldofs, mesh->nv,
              d_rowptr, d_col, d_err);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(adj);
    cudaFree(adjptr);
    cudaFree(d_err);
    tb_pattern *p = nullptr;
    int32_t st = pattern_from_device_csr(ctx, nrows, mesh->ndofs, nnz, d_rowptr, d_col, &p);
    cudaFree(d_col);
    if (st != TB_OK) {
        tb_pattern_release(p);
        return st;
    }
    st = csr_from_pattern(p, out);
    if (st != TB_OK) tb_pattern_release(p);
    return st;
}

extern "C" int32_t tb_csr_create_like(const tb_csr *pattern_of, tb_csr **out) {
    TB_REQUIRE(pattern_of && out, "tb_csr_create_like: NULL argument");
    TB_DEV(pattern_of->pat->ctx);
    *out = nullptr;
    pattern_of->pat->refcount++;
    int32_t st = csr_from_pattern(pattern_of->pat, out);
    if (st != TB_OK) pattern_of->pat->refcount--;
    return st;
}

extern "C" int32_t tb_csr_destroy(tb_csr *a) {
    if (!a) return TB_OK;
    cudaSetDevice(a->pat->ctx->device);
    cudaStreamSynchronize(a->pat->ctx->stream);
    cudaFree(a->d_val);
    tb_pattern_release(a->pat);
    delete a;
    return TB_OK;
}

extern "C" int32_t tb_csr_storage(const tb_csr *a, int64_t *stored_entries, int64_t *column_stream_bytes,
                                  int32_t *max_width) {
    TB_REQUIRE(a, "tb_csr_storage: matrix is NULL");
    if (stored_entries) *stored_entries = a->pat->sell_len;
    if (column_stream_bytes)
        *column_stream_bytes = (a->pat->ctx->spmv_compress && a->pat->d_ccol) ? a->pat->ccol_len * 4 : a->pat->sell_len * 4;
    if (max_width) *max_width = a->pat->max_width;
    return TB_OK;
}

extern "C" int32_t tb_csr_sizes(const tb_csr *a, int64_t *nrows, int64_t *ncols, int64_t *nnz) {
    TB_REQUIRE(a, "tb_csr_sizes: matrix is NULL");
    if (nrows) *nrows = a->pat->nrows;
    if (ncols) *ncols = a->pat->ncols;
    if (nnz) *nnz = a->pat->nnz;
    return TB_OK;
}

// ---- CSR <-> SELL value/pattern transfer -------------------------------------------------------------
// DIR 0: csr <- sell, 1: sell <- csr   (T = double values or int columns widened to int64)
template <typename TS, typename TC, int DIR>
__global__ void k_sell_csr_copy(const int64_t *rowptr, const int64_t *slice_ptr, int64_t nrows, TS *sell, TC *csr, TC add) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t nslices = (nrows + TB_SLICE - 1) / TB_SLICE;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        int64_t r = s * TB_SLICE + lane;
        if (r >= nrows) continue;
        int64_t base = slice_ptr[s] + lane, p0 = rowptr[r];
        int len = (int)(rowptr[r + 1] - p0);
        for (int j = 0; j < len; j++) {
            if (DIR == 0) csr[p0 + j] = (TC)sell[base + (int64_t)j * TB_SLICE] + add;
            else sell[base + (int64_t)j * TB_SLICE] = (TS)csr[p0 + j];
        }
    }
}

extern "C" int32_t tb_csr_download_pattern(const tb_csr *a, int64_t *rowptr, int64_t *colidx, int32_t index_base) {
    TB_REQUIRE(a && rowptr && colidx, "tb_csr_download_pattern: NULL argument");
    tb_pattern *p = a->pat;
    tb_ctx *ctx = p->ctx;
    TB_DEV(ctx);
    int64_t *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, sizeof(int64_t) * (size_t)p->nnz));
    TB_LAUNCH(ctx, (k_sell_csr_copy<int, int64_t, 0>), ctx->sm_count * 8, 256, 0, p->d_rowptr, p->d_slice_ptr, p->nrows,
              p->d_col, tmp, (int64_t)index_base);
    TB_CUDA(cudaMemcpyAsync(colidx, tmp, sizeof(int64_t) * p->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaMemcpyAsync(rowptr, p->d_rowptr, sizeof(int64_t) * (p->nrows + 1), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    if (index_base)
        for (int64_t i = 0; i <= p->nrows; i++) rowptr[i] += index_base;
    return TB_OK;
}

extern "C" int32_t tb_csr_values_download(const tb_csr *a, double *vals) {
    TB_REQUIRE(a && vals, "tb_csr_values_download: NULL argument");
    tb_pattern *p = a->pat;
    tb_ctx *ctx = p->ctx;
    TB_DEV(ctx);
    double *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, sizeof(double) * (size_t)p->nnz));
    TB_LAUNCH(ctx, (k_sell_csr_copy<double, double, 0>), ctx->sm_count * 8, 256, 0, p->d_rowptr, p->d_slice_ptr, p->nrows,
              a->d_val, tmp, 0.0);
    TB_CUDA(cudaMemcpyAsync(vals, tmp, sizeof(double) * p->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    return TB_OK;
}

extern "C" int32_t tb_csr_values_upload(tb_csr *a, const double *vals) {
    TB_REQUIRE(a && vals, "tb_csr_values_upload: NULL argument");
    a->version++;
    tb_pattern *p = a->pat;
    tb_ctx *ctx = p->ctx;
    TB_DEV(ctx);
    double *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, sizeof(double) * (size_t)p->nnz));
    TB_CUDA(cudaMemcpyAsync(tmp, vals, sizeof(double) * p->nnz, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaMemsetAsync(a->d_val, 0, sizeof(double) * (size_t)p->sell_len, ctx->stream));
    TB_LAUNCH(ctx, (k_sell_csr_copy<double, double, 1>), ctx->sm_count * 8, 256, 0, p->d_rowptr, p->d_slice_ptr, p->nrows,
              a->d_val, tmp, 0.0);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    return TB_OK;
}

extern "C" int32_t tb_csr_zero(tb_csr *a) {
    TB_REQUIRE(a, "tb_csr_zero: matrix is NULL");
    a->version++;
    TB_CUDA(cudaMemsetAsync(a->d_val, 0, sizeof(double) * (size_t)a->pat->sell_len, a->pat->ctx->stream));
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------
// nz(A) = nz(M) - dt*nz(K): one pass over the padded value arrays (padding stays 0 - dt*0 = 0).
// 24 B per stored entry; 128-bit loads/stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_axpby_values(const double2 *__restrict__ M, const double2 *__restrict__ K,
                                                      double2 *__restrict__ A, double dt, int64_t n2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        double2 m = M[i], k = K[i], a;
        a.x = m.x - dt * k.x;
        a.y = m.y - dt * k.y;
        A[i] = a;
    }
}

extern "C" int32_t tb_csr_axpby_values(tb_csr *A, const tb_csr *M, const tb_csr *K, double dt) {
    TB_REQUIRE(A && M && K, "tb_csr_axpby_values: NULL argument");
    TB_REQUIRE(A->pat == M->pat && A->pat == K->pat, "tb_csr_axpby_values: A, M, K must share one pattern");
    A->version++;
    tb_ctx *ctx = A->pat->ctx;
    TB_DEV(ctx);
    int64_t n2 = A->pat->sell_len / 2;   // sell_len is a multiple of 32
    TB_LAUNCH(ctx, k_axpby_values, TB_GRID(ctx, k_axpby_values, 256, 0, (n2 + 255) / 256), 256, 0, (const double2 *)M->d_val,
              (const double2 *)K->d_val, (double2 *)A->d_val, dt, n2);
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------
// SpMV.  One warp per slice, lane = row; optional fused dot(x_row, y_row) for CG's p.Ap.
// Bytes per row (hex, 27 nnz): 27*(8+4) matrix + 8 x (gathers hit L1/L2) + 8 y = 340 (+8 rowptr in the
// CSR accounting of SURVEY 8d = 348).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sell_spmv(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col,
                                                   const double *__restrict__ val, const double *__restrict__ x,
                                                   double *__restrict__ y, int64_t nrows, int64_t nslices) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t s = warp; s < nslices; s += nwarps) {
        double acc = tb_sell_row(slice_ptr, col, val, x, s, lane);
        const int64_t r = s * TB_SLICE + lane;
        if (r < nrows) y[r] = acc;
    }
}

int32_t tb_spmv_raw(tb_ctx *ctx, const tb_csr *A, double *x, double *y) {
    const tb_pattern *p = A->pat;
    if (p->halo.nneigh > 0) TB_TRY(tb_halo_exchange(ctx, p, x));
    int grid = TB_GRID(ctx, k_sell_spmv, 256, 0, (p->nslices + 7) / 8);
    TB_LAUNCH(ctx, k_sell_spmv, grid, 256, 0, p->d_slice_ptr, p->d_col, A->d_val, x, y, p->nrows, p->nslices);
    return TB_OK;
}

extern "C" int32_t tb_spmv(tb_ctx *ctx, const tb_csr *A, const tb_vec *x, int32_t xcol, tb_vec *y, int32_t ycol) {
    TB_REQUIRE(ctx && A && x && y, "tb_spmv: NULL argument");
    TB_REQUIRE(xcol >= 0 && xcol < x->ncols && ycol >= 0 && ycol < y->ncols, "tb_spmv: column out of range");
    TB_REQUIRE(x->n >= A->pat->ncols && y->n >= A->pat->nrows, "tb_spmv: vector shorter than the operator");
    TB_REQUIRE(x->d + (size_t)xcol * x->ld != y->d + (size_t)ycol * y->ld, "tb_spmv: x and y must not alias");
    TB_DEV(ctx);
    return tb_spmv_raw(ctx, A, x->d + (size_t)xcol * x->ld, y->d + (size_t)ycol * y->ld);
}

