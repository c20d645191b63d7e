// Mesh handles: upload of a Ferrite grid + closed DofHandler, and the device-side restatement of
// generate_grid (src/mesh/generators.jl:942) with first-touch DoF numbering
// (DofHandler close!, src/discretization/fem.jl:180-182) for meshes too large to build on a host.
#include "tb_internal.cuh"
#include <cub/cub.cuh>

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

static int nv_of(int ct) { return ct == TB_QUAD4 ? 4 : ct == TB_HEX8 ? 8 : ct == TB_TRI3 ? 3 : 4; }
static int dim_of(int ct) { return (ct == TB_QUAD4 || ct == TB_TRI3) ? 2 : 3; }

// ---- small conversion kernels -----------------------------------------------------------------
__global__ void k_i64_to_i32(const int64_t *src, int *dst, int64_t n, int base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (int)(src[i] - base);
}
__global__ void k_i32_to_i64(const int *src, int64_t *dst, int64_t n, int base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (int64_t)src[i] + base;
}
__global__ void k_node2dof_from_cells(const int *conn, const int *celldofs, int *node2dof, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        node2dof[conn[i]] = celldofs[i];
}
__global__ void k_dof_coords(const int *node2dof, const double *coords, double *out, int64_t nnodes, int dim) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        int d = node2dof[n];
        if (d >= 0)
            for (int c = 0; c < dim; c++) out[(int64_t)d * dim + c] = coords[n * dim + c];
    }
}

// ---- structured generators (Ferrite node / cell order) -----------------------------------------
__device__ __forceinline__ double grid_coord(double l, double r, int64_t i, int64_t n) {
    if (i == n) return r;
    return l + ((double)i * (r - l)) / (double)n;
}

#include "tb_grid.cuh"

__global__ void k_gen_nodes(GridDesc g, double *coords, int64_t nnodes) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = n % g.nn[0], j = (n / g.nn[0]) % g.nn[1], k = n / (g.nn[0] * g.nn[1]);
        coords[n * g.dim + 0] = grid_coord(g.left[0], g.right[0], i, g.nel[0]);
        coords[n * g.dim + 1] = grid_coord(g.left[1], g.right[1], j, g.nel[1]);
        if (g.dim == 3) coords[n * g.dim + 2] = grid_coord(g.left[2], g.right[2], k, g.nel[2]);
    }
}

__global__ void k_gen_cells(GridDesc g, int *conn, int64_t nboxes) {
    const int tets[6][4] = {{0, 1, 3, 7}, {0, 4, 1, 7}, {1, 2, 3, 7}, {1, 6, 2, 7}, {1, 4, 5, 7}, {1, 5, 6, 7}};
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nboxes; b += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = b % g.nel[0], j = (b / g.nel[0]) % g.nel[1], k = b / (g.nel[0] * g.nel[1]);
        int64_t nx = g.nn[0], ny = g.nn[1];
#define ND(ii, jj, kk) ((int)((((int64_t)(kk)) * ny + (jj)) * nx + (ii)))
        if (g.dim == 2) {
            int q[4] = {ND(i, j, 0), ND(i + 1, j, 0), ND(i + 1, j + 1, 0), ND(i, j + 1, 0)};
            if (g.celltype == TB_QUAD4) {
                for (int a = 0; a < 4; a++) conn[b * 4 + a] = q[a];
            } else {
                int *c = conn + b * 6;
                c[0] = q[0]; c[1] = q[1]; c[2] = q[3];
                c[3] = q[1]; c[4] = q[2]; c[5] = q[3];
            }
        } else {
            int h[8] = {ND(i, j, k),         ND(i + 1, j, k),         ND(i + 1, j + 1, k),     ND(i, j + 1, k),
                        ND(i, j, k + 1),     ND(i + 1, j, k + 1),     ND(i + 1, j + 1, k + 1), ND(i, j + 1, k + 1)};
            if (g.celltype == TB_HEX8) {
                for (int a = 0; a < 8; a++) conn[b * 8 + a] = h[a];
            } else {
                for (int s = 0; s < 6; s++)
                    for (int a = 0; a < 4; a++) conn[(b * 6 + s) * 4 + a] = h[tets[s][a]];
            }
        }
#undef ND
    }
}

// ---- first-touch numbering ----------------------------------------------------------------------
__global__ void k_first_touch(const int *conn, unsigned long long *firstpos, int64_t npos) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (int64_t)gridDim.x * blockDim.x)
        atomicMin(&firstpos[conn[p]], (unsigned long long)p);
}

#define FT_BLOCK 256
#define FT_ITEMS 8
// one block owns FT_BLOCK*FT_ITEMS consecutive positions
__global__ void __launch_bounds__(FT_BLOCK) k_touch_count(const int *conn, const unsigned long long *firstpos,
                                                          int64_t npos, int64_t *blockcount) {
    int64_t base = (int64_t)blockIdx.x * FT_BLOCK * FT_ITEMS;
    int cnt = 0;
    for (int it = 0; it < FT_ITEMS; it++) {
        int64_t p = base + (int64_t)threadIdx.x * FT_ITEMS + it;
        if (p < npos && firstpos[conn[p]] == (unsigned long long)p) cnt++;
    }
    typedef cub::BlockReduce<int, FT_BLOCK> BR;
    __shared__ typename BR::TempStorage tmp;
    int total = BR(tmp).Sum(cnt);
    if (threadIdx.x == 0) blockcount[blockIdx.x] = total;
}

__global__ void __launch_bounds__(FT_BLOCK) k_touch_assign(const int *conn, const unsigned long long *firstpos,
                                                           int64_t npos, const int64_t *blockoff, int *node2dof) {
    int64_t base = (int64_t)blockIdx.x * FT_BLOCK * FT_ITEMS;
    int flags[FT_ITEMS], cnt = 0;
    for (int it = 0; it < FT_ITEMS; it++) {
        int64_t p = base + (int64_t)threadIdx.x * FT_ITEMS + it;
        flags[it] = (p < npos && firstpos[conn[p]] == (unsigned long long)p) ? 1 : 0;
        cnt += flags[it];
    }
    typedef cub::BlockScan<int, FT_BLOCK> BS;
    __shared__ typename BS::TempStorage tmp;
    int excl;
    BS(tmp).ExclusiveSum(cnt, excl);
    int64_t id = blockoff[blockIdx.x] + excl;
    for (int it = 0; it < FT_ITEMS; it++) {
        int64_t p = base + (int64_t)threadIdx.x * FT_ITEMS + it;
        if (flags[it]) node2dof[conn[p]] = (int)(id++);
    }
}

__global__ void k_celldofs(const int *conn, const int *node2dof, int *celldofs, int64_t npos) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npos; p += (int64_t)gridDim.x * blockDim.x)
        celldofs[p] = node2dof[conn[p]];
}

// numbering from conn (device); fills m->d_node2dof, m->d_celldofs, m->ndofs
static int32_t close_dofs_device(tb_mesh *m) {
    tb_ctx *ctx = m->ctx;
    int64_t npos = m->ncells * m->nv;
    unsigned long long *firstpos = nullptr;
    TB_CUDA(cudaMalloc(&firstpos, sizeof(unsigned long long) * (size_t)m->nnodes));
    TB_CUDA(cudaMemsetAsync(firstpos, 0xFF, sizeof(unsigned long long) * (size_t)m->nnodes, ctx->stream));
    TB_CUDA(cudaMemsetAsync(m->d_node2dof, 0xFF, sizeof(int) * (size_t)m->nnodes, ctx->stream));
    int grid = ctx->sm_count * 8;
    TB_LAUNCH(ctx, k_first_touch, grid, 256, 0, m->d_conn, firstpos, npos);
    int64_t nblocks = (npos + FT_BLOCK * FT_ITEMS - 1) / (FT_BLOCK * FT_ITEMS);
    int64_t *blockcount = nullptr, *blockoff = nullptr;
    TB_CUDA(cudaMalloc(&blockcount, sizeof(int64_t) * (size_t)(nblocks + 1)));
    TB_CUDA(cudaMalloc(&blockoff, sizeof(int64_t) * (size_t)(nblocks + 1)));
    TB_CUDA(cudaMemsetAsync(blockcount, 0, sizeof(int64_t) * (size_t)(nblocks + 1), ctx->stream));
    TB_LAUNCH(ctx, k_touch_count, (unsigned)nblocks, FT_BLOCK, 0, m->d_conn, firstpos, npos, blockcount);
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, blockcount, blockoff, (int)(nblocks + 1), ctx->stream));
    void *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, tmp_bytes));
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, blockcount, blockoff, (int)(nblocks + 1), ctx->stream));
    ctx->launches++;
    TB_LAUNCH(ctx, k_touch_assign, (unsigned)nblocks, FT_BLOCK, 0, m->d_conn, firstpos, npos, blockoff, m->d_node2dof);
    TB_LAUNCH(ctx, k_celldofs, grid, 256, 0, m->d_conn, m->d_node2dof, m->d_celldofs, npos);
    int64_t total = 0;
    TB_CUDA(cudaMemcpyAsync(&total, blockoff + nblocks, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    m->ndofs = m->ndofs_owned = total;
    cudaFree(tmp);
    cudaFree(blockcount);
    cudaFree(blockoff);
    cudaFree(firstpos);
    return TB_OK;
}

static int32_t mesh_alloc(tb_ctx *ctx, int celltype, int64_t ncells, int64_t nnodes, tb_mesh **out) {
    tb_mesh *m = new (std::nothrow) tb_mesh();
    if (!m) return tb_fail(TB_ERR_NOMEM, "tb_mesh: host allocation failed");
    m->ctx = ctx;
    m->celltype = celltype;
    m->nv = nv_of(celltype);
    m->dim = dim_of(celltype);
    m->ncells = ncells;
    m->nnodes = nnodes;
    *out = m;
    TB_CUDA(cudaMalloc(&m->d_conn, sizeof(int) * (size_t)(ncells * m->nv + 1)));
    TB_CUDA(cudaMalloc(&m->d_celldofs, sizeof(int) * (size_t)(ncells * m->nv + 1)));
    TB_CUDA(cudaMalloc(&m->d_coords, sizeof(double) * (size_t)(nnodes * m->dim + 1)));
    TB_CUDA(cudaMalloc(&m->d_node2dof, sizeof(int) * (size_t)(nnodes + 1)));
    return TB_OK;
}

extern "C" int32_t tb_mesh_destroy(tb_mesh *m) {
    if (!m) return TB_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_conn);
    cudaFree(m->d_celldofs);
    cudaFree(m->d_coords);
    cudaFree(m->d_node2dof);
    cudaFree(m->d_ghost_global);
    cudaFree(m->d_adjptr);
    cudaFree(m->d_adj);
    cudaFree(m->d_adjpos);
    delete m;
    return TB_OK;
}

extern "C" int32_t tb_mesh_create(tb_ctx *ctx, int32_t celltype, int64_t ncells, int64_t nnodes, const int64_t *conn,
                                  const double *coords, const int64_t *celldofs, int64_t ndofs, int32_t index_base,
                                  tb_mesh **out) {
    TB_REQUIRE(ctx && conn && coords && celldofs && out, "tb_mesh_create: NULL argument");
    TB_REQUIRE(celltype >= TB_QUAD4 && celltype <= TB_TET4, "tb_mesh_create: unknown cell type %d", celltype);
    TB_REQUIRE(ncells > 0 && nnodes > 0 && ndofs > 0, "tb_mesh_create: empty mesh");
    TB_REQUIRE(nnodes < INT32_MAX && ndofs < INT32_MAX, "tb_mesh_create: more than 2^31 nodes per GPU");
    TB_DEV(ctx);
    *out = nullptr;
    tb_mesh *m = nullptr;
    int32_t st = mesh_alloc(ctx, celltype, ncells, nnodes, &m);
    if (st != TB_OK) {
        tb_mesh_destroy(m);
        return st;
    }
    int64_t npos = ncells * m->nv;
    int64_t *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, sizeof(int64_t) * (size_t)npos));
    int grid = tb_grid_for(ctx, npos, 256, 8);
    TB_CUDA(cudaMemcpyAsync(tmp, conn, sizeof(int64_t) * npos, cudaMemcpyHostToDevice, ctx->stream));
    TB_LAUNCH(ctx, k_i64_to_i32, grid, 256, 0, tmp, m->d_conn, npos, index_base);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    TB_CUDA(cudaMemcpyAsync(tmp, celldofs, sizeof(int64_t) * npos, cudaMemcpyHostToDevice, ctx->stream));
    TB_LAUNCH(ctx, k_i64_to_i32, grid, 256, 0, tmp, m->d_celldofs, npos, index_base);
    TB_CUDA(cudaMemcpyAsync(m->d_coords, coords, sizeof(double) * nnodes * m->dim, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaMemsetAsync(m->d_node2dof, 0xFF, sizeof(int) * (size_t)nnodes, ctx->stream));
    TB_LAUNCH(ctx, k_node2dof_from_cells, grid, 256, 0, m->d_conn, m->d_celldofs, m->d_node2dof, npos);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    m->ndofs = m->ndofs_owned = ndofs;
    *out = m;
    return TB_OK;
}

extern "C" int32_t tb_mesh_generate_grid(tb_ctx *ctx, int32_t celltype, const int64_t *nel3, const double *left3,
                                         const double *right3, tb_mesh **out) {
    TB_REQUIRE(ctx && nel3 && left3 && right3 && out, "tb_mesh_generate_grid: NULL argument");
    TB_REQUIRE(celltype >= TB_QUAD4 && celltype <= TB_TET4, "tb_mesh_generate_grid: unknown cell type %d", celltype);
    TB_DEV(ctx);
    *out = nullptr;
    GridDesc g;
    g.celltype = celltype;
    g.dim = dim_of(celltype);
    for (int d = 0; d < 3; d++) {
        g.nel[d] = d < g.dim ? nel3[d] : 1;
        g.nn[d] = d < g.dim ? nel3[d] + 1 : 1;
        g.left[d] = left3[d];
        g.right[d] = right3[d];
        TB_REQUIRE(g.nel[d] >= 1, "tb_mesh_generate_grid: nel[%d] must be >= 1", d);
    }
    int64_t nboxes = g.nel[0] * g.nel[1] * g.nel[2];
    int64_t nnodes = g.nn[0] * g.nn[1] * g.nn[2];
    int64_t ncells = nboxes * (celltype == TB_TRI3 ? 2 : celltype == TB_TET4 ? 6 : 1);
    TB_REQUIRE(nnodes < INT32_MAX && ncells * nv_of(celltype) < ((int64_t)1 << 40), "tb_mesh_generate_grid: grid too large");
    tb_mesh *m = nullptr;
    int32_t st = mesh_alloc(ctx, celltype, ncells, nnodes, &m);
    if (st != TB_OK) {
        tb_mesh_destroy(m);
        return st;
    }
    TB_LAUNCH(ctx, k_gen_nodes, tb_grid_for(ctx, nnodes, 256, 8), 256, 0, g, m->d_coords, nnodes);
    TB_LAUNCH(ctx, k_gen_cells, tb_grid_for(ctx, nboxes, 256, 8), 256, 0, g, m->d_conn, nboxes);
    st = close_dofs_device(m);
    if (st != TB_OK) {
        tb_mesh_destroy(m);
        return st;
    }
    *out = m;
    return TB_OK;
}

extern "C" int32_t tb_mesh_sizes(const tb_mesh *m, int64_t *ncells, int64_t *nnodes, int64_t *ndofs, int32_t *nv,
                                 int32_t *dim) {
    TB_REQUIRE(m, "tb_mesh_sizes: mesh is NULL");
    if (ncells) *ncells = m->ncells;
    if (nnodes) *nnodes = m->nnodes;
    if (ndofs) *ndofs = m->ndofs;
    if (nv) *nv = m->nv;
    if (dim) *dim = m->dim;
    return TB_OK;
}

static int32_t download_i32_as_i64(tb_ctx *ctx, const int *d_src, int64_t n, int64_t *host) {
    int64_t *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, sizeof(int64_t) * (size_t)n));
    TB_LAUNCH(ctx, k_i32_to_i64, tb_grid_for(ctx, n, 256, 8), 256, 0, d_src, tmp, n, 0);
    TB_CUDA(cudaMemcpyAsync(host, tmp, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    return TB_OK;
}

extern "C" int32_t tb_mesh_download(const tb_mesh *m, int64_t *conn, double *coords, int64_t *celldofs) {
    TB_REQUIRE(m, "tb_mesh_download: mesh is NULL");
    tb_ctx *ctx = m->ctx;
    TB_DEV(ctx);
    if (conn) TB_TRY(download_i32_as_i64(ctx, m->d_conn, m->ncells * m->nv, conn));
    if (celldofs) TB_TRY(download_i32_as_i64(ctx, m->d_celldofs, m->ncells * m->nv, celldofs));
    if (coords) {
        TB_CUDA(cudaMemcpyAsync(coords, m->d_coords, sizeof(double) * m->nnodes * m->dim, cudaMemcpyDeviceToHost,
                                ctx->stream));
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return TB_OK;
}

extern "C" int32_t tb_mesh_dof_coords(const tb_mesh *m, double *host) {
    TB_REQUIRE(m && host, "tb_mesh_dof_coords: NULL argument");
    tb_ctx *ctx = m->ctx;
    TB_DEV(ctx);
    double *tmp = nullptr;
    size_t bytes = sizeof(double) * (size_t)(m->ndofs * m->dim);
    TB_CUDA(cudaMalloc(&tmp, bytes));
    TB_CUDA(cudaMemsetAsync(tmp, 0, bytes, ctx->stream));
    TB_LAUNCH(ctx, k_dof_coords, tb_grid_for(ctx, m->nnodes, 256, 8), 256, 0, m->d_node2dof, m->d_coords, tmp, m->nnodes,
              m->dim);
    TB_CUDA(cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    return TB_OK;
}

// ---- multi-GPU: local sub-mesh of a contiguous dof range --------------------------------------------
// A cell is local if it touches at least one dof in [lo, hi).  Local dof ids: owned = global - lo,
// ghosts = nowned + rank of the global id among the (sorted, unique) ghosts.
__global__ void k_mark_local_cells(const int *celldofs, int64_t ncells, int nv, int lo, int hi, int *flag) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
        int f = 0;
        for (int a = 0; a < nv; a++) {
            int d = celldofs[c * nv + a];
            f |= (d >= lo && d < hi);
        }
        flag[c] = f;
    }
}
__global__ void k_mark_ghost_dofs(const int *celldofs, const int *cellflag, int64_t ncells, int nv, int lo, int hi,
                                  int *dofflag) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
        if (!cellflag[c]) continue;
        for (int a = 0; a < nv; a++) {
            int d = celldofs[c * nv + a];
            if (d < lo || d >= hi) dofflag[d] = 1;
        }
    }
}
__global__ void k_mark_local_nodes(const int *conn, const int *cellflag, int64_t ncells, int nv, int *nodeflag) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
        if (!cellflag[c]) continue;
        for (int a = 0; a < nv; a++) nodeflag[conn[c * nv + a]] = 1;
    }
}
__global__ void k_ghost_list(const int *dofflag, const int *dofscan, int64_t ndofs, int64_t *ghost_global) {
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < ndofs; d += (int64_t)gridDim.x * blockDim.x)
        if (dofflag[d]) ghost_global[dofscan[d]] = d;
}
__global__ void k_compact_cells(const int *conn, const int *celldofs, const int *cellflag, const int *cellscan,
                                const int *nodescan, const int *dofflag, const int *dofscan, int64_t ncells, int nv,
                                int lo, int hi, int nowned, int *lconn, int *ldofs) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
        if (!cellflag[c]) continue;
        int64_t lc = cellscan[c];
        for (int a = 0; a < nv; a++) {
            lconn[lc * nv + a] = nodescan[conn[c * nv + a]];
            int d = celldofs[c * nv + a];
            ldofs[lc * nv + a] = (d >= lo && d < hi) ? d - lo : nowned + dofscan[d];
        }
    }
}
__global__ void k_compact_nodes(const double *coords, const int *node2dof, const int *nodeflag, const int *nodescan,
                                const int *dofscan, int64_t nnodes, int dim, int lo, int hi, int nowned, double *lcoords,
                                int *lnode2dof) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        if (!nodeflag[n]) continue;
        int ln = nodescan[n];
        for (int c = 0; c < dim; c++) lcoords[(int64_t)ln * dim + c] = coords[n * dim + c];
        int d = node2dof[n];
        lnode2dof[ln] = d < 0 ? -1 : ((d >= lo && d < hi) ? d - lo : nowned + dofscan[d]);
    }
}

static int32_t scan_flags(tb_ctx *ctx, const int *flag, int *scan, int64_t n, int64_t *total) {
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag, scan, (int)n, ctx->stream));
    void *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, tmp_bytes));
    TB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, scan, (int)n, ctx->stream));
    ctx->launches++;
    int last_scan = 0, last_flag = 0;
    TB_CUDA(cudaMemcpyAsync(&last_scan, scan + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaMemcpyAsync(&last_flag, flag + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
    *total = (int64_t)last_scan + last_flag;
    return TB_OK;
}

extern "C" int32_t tb_mesh_extract_local(const tb_mesh *g, int64_t dof_lo, int64_t dof_hi, tb_mesh **out,
                                         int64_t *nghost) {
    TB_REQUIRE(g && out, "tb_mesh_extract_local: NULL argument");
    TB_REQUIRE(0 <= dof_lo && dof_lo < dof_hi && dof_hi <= g->ndofs, "tb_mesh_extract_local: bad dof range");
    tb_ctx *ctx = g->ctx;
    TB_DEV(ctx);
    *out = nullptr;
    int grid = ctx->sm_count * 8;
    int *cellflag, *cellscan, *dofflag, *dofscan, *nodeflag, *nodescan;
    TB_CUDA(cudaMalloc(&cellflag, sizeof(int) * (size_t)g->ncells));
    TB_CUDA(cudaMalloc(&cellscan, sizeof(int) * (size_t)g->ncells));
    TB_CUDA(cudaMalloc(&dofflag, sizeof(int) * (size_t)g->ndofs));
    TB_CUDA(cudaMalloc(&dofscan, sizeof(int) * (size_t)g->ndofs));
    TB_CUDA(cudaMalloc(&nodeflag, sizeof(int) * (size_t)g->nnodes));
    TB_CUDA(cudaMalloc(&nodescan, sizeof(int) * (size_t)g->nnodes));
    TB_CUDA(cudaMemsetAsync(dofflag, 0, sizeof(int) * (size_t)g->ndofs, ctx->stream));
    TB_CUDA(cudaMemsetAsync(nodeflag, 0, sizeof(int) * (size_t)g->nnodes, ctx->stream));
    TB_LAUNCH(ctx, k_mark_local_cells, grid, 256, 0, g->d_celldofs, g->ncells, g->nv, (int)dof_lo, (int)dof_hi, cellflag);
    TB_LAUNCH(ctx, k_mark_ghost_dofs, grid, 256, 0, g->d_celldofs, cellflag, g->ncells, g->nv, (int)dof_lo, (int)dof_hi,
              dofflag);
    TB_LAUNCH(ctx, k_mark_local_nodes, grid, 256, 0, g->d_conn, cellflag, g->ncells, g->nv, nodeflag);
    int64_t lcells = 0, lghost = 0, lnodes = 0;
    TB_TRY(scan_flags(ctx, cellflag, cellscan, g->ncells, &lcells));
    TB_TRY(scan_flags(ctx, dofflag, dofscan, g->ndofs, &lghost));
    TB_TRY(scan_flags(ctx, nodeflag, nodescan, g->nnodes, &lnodes));
    TB_REQUIRE(lcells > 0, "tb_mesh_extract_local: no cell touches dofs [%lld,%lld)", (long long)dof_lo, (long long)dof_hi);
    tb_mesh *m = nullptr;
    int32_t st = mesh_alloc(ctx, g->celltype, lcells, lnodes, &m);
    if (st != TB_OK) {
        tb_mesh_destroy(m);
        return st;
    }
    int nowned = (int)(dof_hi - dof_lo);
    m->ndofs_owned = nowned;
    m->ndofs = nowned + lghost;
    m->nghost = lghost;
    m->dof_lo = dof_lo;
    TB_CUDA(cudaMalloc(&m->d_ghost_global, sizeof(int64_t) * (size_t)(lghost + 1)));
    TB_LAUNCH(ctx, k_ghost_list, grid, 256, 0, dofflag, dofscan, g->ndofs, m->d_ghost_global);
    TB_LAUNCH(ctx, k_compact_cells, grid, 256, 0, g->d_conn, g->d_celldofs, cellflag, cellscan, nodescan, dofflag, dofscan,
              g->ncells, g->nv, (int)dof_lo, (int)dof_hi, nowned, m->d_conn, m->d_celldofs);
    TB_LAUNCH(ctx, k_compact_nodes, grid, 256, 0, g->d_coords, g->d_node2dof, nodeflag, nodescan, dofscan, g->nnodes,
              g->dim, (int)dof_lo, (int)dof_hi, nowned, m->d_coords, m->d_node2dof);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(cellflag); cudaFree(cellscan); cudaFree(dofflag); cudaFree(dofscan); cudaFree(nodeflag); cudaFree(nodescan);
    if (nghost) *nghost = lghost;
    *out = m;
    return TB_OK;
}

// Marks a mesh uploaded with tb_mesh_create as the LOCAL part of a partitioned mesh: its dofs are numbered owned-first
// (0 .. ndofs_owned-1 = global ids dof_lo ..), ghosts after (ascending global id).  What tb_mesh_extract_local produces on
// the device, for callers that cut the mesh on the host (general partitions, no global replica in HBM).
extern "C" int32_t tb_mesh_set_ownership(tb_mesh *m, int64_t ndofs_owned, int64_t dof_lo, const int64_t *ghost_global,
                                         int64_t nghost) {
    TB_REQUIRE(m, "tb_mesh_set_ownership: mesh is NULL");
    TB_REQUIRE(ndofs_owned > 0 && nghost >= 0 && ndofs_owned + nghost == m->ndofs,
               "tb_mesh_set_ownership: %lld owned + %lld ghosts != %lld dofs of the mesh", (long long)ndofs_owned, (long long)nghost,
               (long long)m->ndofs);
    TB_REQUIRE(nghost == 0 || ghost_global, "tb_mesh_set_ownership: ghost list is NULL");
    for (int64_t i = 1; i < nghost; i++)
        TB_REQUIRE(ghost_global[i - 1] < ghost_global[i], "tb_mesh_set_ownership: ghost ids must be strictly ascending");
    tb_ctx *ctx = m->ctx;
    TB_DEV(ctx);
    TB_REQUIRE(!m->d_adjptr, "tb_mesh_set_ownership: call before the first assembly");
    cudaFree(m->d_ghost_global);
    m->d_ghost_global = nullptr;
    TB_CUDA(cudaMalloc(&m->d_ghost_global, sizeof(int64_t) * (size_t)(nghost + 1)));
    if (nghost) TB_CUDA(cudaMemcpy(m->d_ghost_global, ghost_global, sizeof(int64_t) * (size_t)nghost, cudaMemcpyHostToDevice));
    m->ndofs_owned = ndofs_owned;
    m->nghost = nghost;
    m->dof_lo = dof_lo;
    return TB_OK;
}

// =====================================================================================================
// Structured grids, local part only: what tb_mesh_extract_local(tb_mesh_generate_grid(...), lo, hi) returns, built
// WITHOUT ever holding the global grid in HBM.  Ferrite's first-touch numbering of a
// Quadrilateral / Hexahedron grid has a closed form: node (a, b, c) is first touched by cell (max(a-1,0), max(b-1,0),
// max(c-1,0)) -- the lexicographically first cell that contains it -- and its id is the number of nodes introduced by the
// cells before that one plus its rank among the cell's new vertices in local vertex order.  All temporaries are sized by
// the slab of cell layers that can touch [lo, hi).
// =====================================================================================================
struct SlabDesc {
    GridDesc g;
    int64_t k0, k1;          // cell layers of the outermost direction
    int64_t cells_per_layer, nodes_per_layer;
    int64_t ncells, nnodes;  // of the slab
    int64_t node_base, dof_base, ndofs_slab;
    int64_t lo, hi;
};
// slab cell id -> its vertices' global (a, b, c); vertex v
__device__ __forceinline__ void slab_cell_vertex(const SlabDesc &S, int64_t sc, int v, int64_t &a, int64_t &b, int64_t &c) {
    const GridDesc &g = S.g;
    const int oa[8] = {0, 1, 1, 0, 0, 1, 1, 0}, ob[8] = {0, 0, 1, 1, 0, 0, 1, 1}, oc[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    if (g.dim == 3) {
        const int64_t i = sc % g.nel[0], j = (sc / g.nel[0]) % g.nel[1], k = S.k0 + sc / (g.nel[0] * g.nel[1]);
        a = i + oa[v]; b = j + ob[v]; c = k + oc[v];
    } else {
        const int64_t i = sc % g.nel[0], j = S.k0 + sc / g.nel[0];
        a = i + oa[v]; b = j + ob[v]; c = 0;
    }
}
__global__ void k_slab_mark_cells(SlabDesc S, int nv, int *cellflag) {
    for (int64_t sc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sc < S.ncells; sc += (int64_t)gridDim.x * blockDim.x) {
        int f = 0;
        for (int v = 0; v < nv; v++) {
            int64_t a, b, c;
            slab_cell_vertex(S, sc, v, a, b, c);
            const int64_t d = tb_grid_dof(S.g, a, b, c);
            f |= (d >= S.lo && d < S.hi);
        }
        cellflag[sc] = f;
    }
}
__global__ void k_slab_mark_nodes_dofs(SlabDesc S, int nv, const int *cellflag, int *nodeflag, int *dofflag) {
    for (int64_t sc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sc < S.ncells; sc += (int64_t)gridDim.x * blockDim.x) {
        if (!cellflag[sc]) continue;
        for (int v = 0; v < nv; v++) {
            int64_t a, b, c;
            slab_cell_vertex(S, sc, v, a, b, c);
            const int64_t node = (c * S.g.nn[1] + b) * S.g.nn[0] + a;
            nodeflag[node - S.node_base] = 1;
            const int64_t d = tb_grid_dof(S.g, a, b, c);
            if (d < S.lo || d >= S.hi) dofflag[d - S.dof_base] = 1;
        }
    }
}
__global__ void k_slab_ghost_list(SlabDesc S, const int *dofflag, const int *dofscan, int64_t *ghost_global) {
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < S.ndofs_slab; d += (int64_t)gridDim.x * blockDim.x)
        if (dofflag[d]) ghost_global[dofscan[d]] = d + S.dof_base;
}
__global__ void k_slab_compact_cells(SlabDesc S, int nv, const int *cellflag, const int *cellscan, const int *nodescan,
                                     const int *dofscan, int nowned, int *lconn, int *ldofs) {
    for (int64_t sc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sc < S.ncells; sc += (int64_t)gridDim.x * blockDim.x) {
        if (!cellflag[sc]) continue;
        const int64_t lc = cellscan[sc];
        for (int v = 0; v < nv; v++) {
            int64_t a, b, c;
            slab_cell_vertex(S, sc, v, a, b, c);
            const int64_t node = (c * S.g.nn[1] + b) * S.g.nn[0] + a;
            lconn[lc * nv + v] = nodescan[node - S.node_base];
            const int64_t d = tb_grid_dof(S.g, a, b, c);
            ldofs[lc * nv + v] = (d >= S.lo && d < S.hi) ? (int)(d - S.lo) : nowned + dofscan[d - S.dof_base];
        }
    }
}
__global__ void k_slab_compact_nodes(SlabDesc S, const int *nodeflag, const int *nodescan, const int *dofscan, int nowned,
                                     double *lcoords, int *lnode2dof) {
    const GridDesc &g = S.g;
    for (int64_t sn = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sn < S.nnodes; sn += (int64_t)gridDim.x * blockDim.x) {
        if (!nodeflag[sn]) continue;
        const int64_t n = sn + S.node_base;
        const int64_t a = n % g.nn[0], b = (n / g.nn[0]) % g.nn[1], c = n / (g.nn[0] * g.nn[1]);
        const int ln = nodescan[sn];
        lcoords[(int64_t)ln * g.dim + 0] = grid_coord(g.left[0], g.right[0], a, g.nel[0]);
        lcoords[(int64_t)ln * g.dim + 1] = grid_coord(g.left[1], g.right[1], b, g.nel[1]);
        if (g.dim == 3) lcoords[(int64_t)ln * g.dim + 2] = grid_coord(g.left[2], g.right[2], c, g.nel[2]);
        const int64_t d = tb_grid_dof(g, a, b, c);
        lnode2dof[ln] = (d >= S.lo && d < S.hi) ? (int)(d - S.lo) : nowned + dofscan[d - S.dof_base];
    }
}

extern "C" int32_t tb_mesh_generate_grid_local(tb_ctx *ctx, int32_t celltype, const int64_t *nel3, const double *left3,
                                               const double *right3, int64_t dof_lo, int64_t dof_hi, tb_mesh **out, int64_t *nghost) {
    TB_REQUIRE(ctx && nel3 && left3 && right3 && out, "tb_mesh_generate_grid_local: NULL argument");
    TB_REQUIRE(celltype == TB_QUAD4 || celltype == TB_HEX8, "tb_mesh_generate_grid_local: closed-form numbering exists for Quadrilateral and Hexahedron grids only");
    TB_DEV(ctx);
    *out = nullptr;
    SlabDesc S;
    GridDesc &g = S.g;
    g.celltype = celltype;
    g.dim = dim_of(celltype);
    for (int d = 0; d < 3; d++) {
        g.nel[d] = d < g.dim ? nel3[d] : 1;
        g.nn[d] = d < g.dim ? nel3[d] + 1 : 1;
        g.left[d] = left3[d];
        g.right[d] = right3[d];
        TB_REQUIRE(g.nel[d] >= 1, "tb_mesh_generate_grid_local: nel[%d] must be >= 1", d);
    }
    const int od = g.dim - 1;                                     // outermost loop direction
    const int64_t L = od == 2 ? g.nn[0] * g.nn[1] : g.nn[0];      // nodes per outer layer
    const int64_t CL = od == 2 ? g.nel[0] * g.nel[1] : g.nel[0];  // cells per outer layer
    const int64_t ndofs_global = g.nn[0] * g.nn[1] * g.nn[2];
    TB_REQUIRE(0 <= dof_lo && dof_lo < dof_hi && dof_hi <= ndofs_global, "tb_mesh_generate_grid_local: bad dof range");
    // node layers 0 and 1 share the ids [0, 2L); layer c >= 2 holds [cL, (c+1)L)
    const int64_t c_lo = dof_lo < 2 * L ? 0 : dof_lo / L, c_hi = (dof_hi - 1) < 2 * L ? 1 : (dof_hi - 1) / L;
    S.k0 = c_lo > 0 ? c_lo - 1 : 0;
    S.k1 = c_hi < g.nel[od] ? c_hi : g.nel[od] - 1;
    S.cells_per_layer = CL;
    S.nodes_per_layer = L;
    S.ncells = (S.k1 - S.k0 + 1) * CL;
    S.nnodes = (S.k1 - S.k0 + 2) * L;
    S.node_base = S.k0 * L;
    S.dof_base = S.k0 <= 1 ? 0 : S.k0 * L;
    S.ndofs_slab = (S.k1 + 2) * L - S.dof_base;
    S.lo = dof_lo;
    S.hi = dof_hi;
    TB_REQUIRE(S.ncells < INT32_MAX && S.nnodes < INT32_MAX, "tb_mesh_generate_grid_local: slab too large for one GPU");
    const int nv = nv_of(celltype);
    const int grid = ctx->sm_count * 8;
    int *cellflag, *cellscan, *dofflag, *dofscan, *nodeflag, *nodescan;
    TB_CUDA(cudaMalloc(&cellflag, sizeof(int) * (size_t)S.ncells));
    TB_CUDA(cudaMalloc(&cellscan, sizeof(int) * (size_t)S.ncells));
    TB_CUDA(cudaMalloc(&dofflag, sizeof(int) * (size_t)S.ndofs_slab));
    TB_CUDA(cudaMalloc(&dofscan, sizeof(int) * (size_t)S.ndofs_slab));
    TB_CUDA(cudaMalloc(&nodeflag, sizeof(int) * (size_t)S.nnodes));
    TB_CUDA(cudaMalloc(&nodescan, sizeof(int) * (size_t)S.nnodes));
    TB_CUDA(cudaMemsetAsync(dofflag, 0, sizeof(int) * (size_t)S.ndofs_slab, ctx->stream));
    TB_CUDA(cudaMemsetAsync(nodeflag, 0, sizeof(int) * (size_t)S.nnodes, ctx->stream));
    TB_LAUNCH(ctx, k_slab_mark_cells, grid, 256, 0, S, nv, cellflag);
    TB_LAUNCH(ctx, k_slab_mark_nodes_dofs, grid, 256, 0, S, nv, cellflag, nodeflag, dofflag);
    int64_t lcells = 0, lghost = 0, lnodes = 0;
    TB_TRY(scan_flags(ctx, cellflag, cellscan, S.ncells, &lcells));
    TB_TRY(scan_flags(ctx, dofflag, dofscan, S.ndofs_slab, &lghost));
    TB_TRY(scan_flags(ctx, nodeflag, nodescan, S.nnodes, &lnodes));
    TB_REQUIRE(lcells > 0, "tb_mesh_generate_grid_local: no cell touches dofs [%lld,%lld)", (long long)dof_lo, (long long)dof_hi);
    tb_mesh *m = nullptr;
    int32_t st = mesh_alloc(ctx, celltype, lcells, lnodes, &m);
    if (st != TB_OK) {
        tb_mesh_destroy(m);
        return st;
    }
    const int nowned = (int)(dof_hi - dof_lo);
    m->ndofs_owned = nowned;
    m->ndofs = nowned + lghost;
    m->nghost = lghost;
    m->dof_lo = dof_lo;
    TB_CUDA(cudaMalloc(&m->d_ghost_global, sizeof(int64_t) * (size_t)(lghost + 1)));
    TB_LAUNCH(ctx, k_slab_ghost_list, grid, 256, 0, S, dofflag, dofscan, m->d_ghost_global);
    TB_LAUNCH(ctx, k_slab_compact_cells, grid, 256, 0, S, nv, cellflag, cellscan, nodescan, dofscan, nowned, m->d_conn, m->d_celldofs);
    TB_LAUNCH(ctx, k_slab_compact_nodes, grid, 256, 0, S, nodeflag, nodescan, dofscan, nowned, m->d_coords, m->d_node2dof);
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(cellflag); cudaFree(cellscan); cudaFree(dofflag); cudaFree(dofscan); cudaFree(nodeflag); cudaFree(nodescan);
    if (nghost) *nghost = lghost;
    *out = m;
    return TB_OK;
}

extern "C" int32_t tb_mesh_ghosts(const tb_mesh *m, int64_t *ghost_global) {
    TB_REQUIRE(m && ghost_global, "tb_mesh_ghosts: NULL argument");
    if (m->nghost == 0) return TB_OK;
    TB_CUDA(cudaMemcpyAsync(ghost_global, m->d_ghost_global, sizeof(int64_t) * m->nghost, cudaMemcpyDeviceToHost,
                            m->ctx->stream));
    TB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return TB_OK;
}
