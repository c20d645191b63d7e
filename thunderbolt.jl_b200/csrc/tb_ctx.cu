// Context, device vectors, timers.  Replaces the CuVector factories of
// ext/CuThunderboltExt.jl:126-127,144-146 and the device selection of src/devices.jl:1-4.
#include "tb_internal.cuh"
#include <string.h>

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

static thread_local char g_last_error[1024] = "";

int32_t tb_fail(int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *tb_last_error(void) { return g_last_error; }
extern "C" int32_t tb_version(void) { return 100; }

int tb_grid_for(tb_ctx *ctx, int64_t work_items, int block, int blocks_per_sm) {
    int64_t need = (work_items + block - 1) / block;
    int64_t cap = (int64_t)ctx->sm_count * blocks_per_sm;
    if (cap > TB_MAX_PARTIALS) cap = TB_MAX_PARTIALS;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static int32_t ctx_init(tb_ctx *ctx, int32_t device, void *stream);

extern "C" int32_t tb_ctx_create(int32_t device, void *stream, tb_ctx **out) {
    TB_REQUIRE(out != nullptr, "tb_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    TB_CUDA(cudaGetDeviceCount(&ndev));
    TB_REQUIRE(device >= 0 && device < ndev, "tb_ctx_create: device %d out of range (%d visible)", device, ndev);
    TB_CUDA(cudaSetDevice(device));
    tb_ctx *ctx = new (std::nothrow) tb_ctx();
    if (!ctx) return tb_fail(TB_ERR_NOMEM, "tb_ctx_create: host allocation failed");
    ctx->device = device;
    const int32_t st = ctx_init(ctx, device, stream);
    if (st != TB_OK) {   // free whatever was built (tb_ctx_destroy tolerates a partially initialised context)
        tb_ctx_destroy(ctx);
        return st;
    }
    *out = ctx;
    return TB_OK;
}

static int32_t ctx_init(tb_ctx *ctx, int32_t device, void *stream) {
    cudaDeviceProp prop;
    TB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = prop.totalGlobalMem;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    if (prop.major != 10)
        return tb_fail(TB_ERR_UNSUPPORTED, "tb_ctx_create: libtbolt_b200 is built for sm_100a only, device is sm_%d%d",
                       prop.major, prop.minor);
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        TB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    TB_CUDA(cudaEventCreate(&ctx->ev0));
    TB_CUDA(cudaEventCreate(&ctx->ev1));
    TB_CUDA(cudaMalloc(&ctx->d_partials, sizeof(double) * 8 * TB_MAX_PARTIALS));   // high words | low words (exact-dot mode)
    TB_CUDA(cudaMalloc(&ctx->d_ticket, sizeof(unsigned) * 8));
    TB_CUDA(cudaMemsetAsync(ctx->d_ticket, 0, sizeof(unsigned) * 8, ctx->stream));
    TB_CUDA(cudaMalloc(&ctx->d_cg, 2 * sizeof(CGState)));   // [1] is the ping-pong partner of the fused multi-GPU path
    TB_CUDA(cudaMemsetAsync(ctx->d_cg, 0, 2 * sizeof(CGState), ctx->stream));
    TB_CUDA(cudaMallocHost(&ctx->h_cg, sizeof(CGState)));
    TB_CUDA(cudaMalloc(&ctx->d_scalar, sizeof(double) * 16));
    TB_CUDA(cudaMalloc(&ctx->d_dconst, sizeof(double) * 16));
    TB_CUDA(cudaMallocHost(&ctx->h_scalar, sizeof(double) * 16));
    if (const char *v = getenv("TB_SPMV_VARIANT")) ctx->spmv_variant = atoi(v);
    if (const char *v = getenv("TB_SPMV_COMPRESS")) ctx->spmv_compress = atoi(v);
    if (const char *v = getenv("TB_CG_PERSISTENT")) ctx->cg_persistent = atoi(v) < 0 ? 0 : atoi(v) > 2 ? 2 : atoi(v);
    if (const char *v = getenv("TB_P2P_FUSED")) ctx->p2p_fused = atoi(v);
    if (const char *v = getenv("TB_DOT_EXACT")) ctx->exact_dot = atoi(v) != 0;
    if (const char *v = getenv("TB_SPMV_FUSEP")) ctx->spmv_fusep = atoi(v) != 0;
    if (const char *v = getenv("TB_CG_PERSISTENT_MAX_ROWS")) ctx->cg_persistent_max_rows = atoll(v);
    if (const char *v = getenv("TB_ASSEMBLY_MODE")) ctx->assembly_mode = atoi(v) == 0 ? 0 : 2;
    if (const char *v = getenv("TB_EA_BUDGET_MB")) ctx->ea_budget_bytes = (size_t)(atof(v) * 1048576.0);
    return TB_OK;
}

extern "C" int32_t tb_ctx_destroy(tb_ctx *ctx) {
    if (!ctx) return TB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    tb_peer_release(ctx);
    if (ctx->has_comm) ncclCommDestroy(ctx->comm);
    cudaFree(ctx->d_partials);
    cudaFree(ctx->d_ticket);
    cudaFree(ctx->d_cgwork);
    cudaFree(ctx->d_cg);
    cudaFreeHost(ctx->h_cg);
    cudaFree(ctx->d_scalar);
    cudaFree(ctx->d_dconst);
    cudaFreeHost(ctx->h_scalar);
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 5; b++) cudaFree(ctx->d_tables[a][b]);
    cudaFree(ctx->d_flush);
    cudaFree(ctx->d_dinv);
    cudaFree(ctx->d_pcwork);
    tb_bj_free(ctx->bj);
    cudaFree(ctx->d_ea);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 2 * TB_PROF_MAX; i++)
        if (ctx->prof_ev[i]) cudaEventDestroy(ctx->prof_ev[i]);
    if (ctx->stage_stream) {
        cudaStreamDestroy(ctx->stage_stream);
        cudaEventDestroy(ctx->stage_ev);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TB_OK;
}

extern "C" int32_t tb_sync(tb_ctx *ctx) {
    TB_REQUIRE(ctx, "tb_sync: ctx is NULL");
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_device_info(tb_ctx *ctx, int32_t *sm_count, int64_t *total_mem_bytes, int32_t *cc_major,
                                  int32_t *cc_minor) {
    TB_REQUIRE(ctx, "tb_device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)ctx->total_mem;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    return TB_OK;
}

extern "C" int32_t tb_timer_start(tb_ctx *ctx) {
    TB_REQUIRE(ctx, "tb_timer_start: ctx is NULL");
    TB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_timer_stop(tb_ctx *ctx, double *elapsed_ms) {
    TB_REQUIRE(ctx && elapsed_ms, "tb_timer_stop: NULL argument");
    TB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    TB_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    TB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *elapsed_ms = (double)ms;
    return TB_OK;
}

extern "C" int32_t tb_launch_count(tb_ctx *ctx, int64_t *count) {
    TB_REQUIRE(ctx && count, "tb_launch_count: NULL argument");
    *count = ctx->launches;
    return TB_OK;
}

__global__ void tb_flush_kernel(double *buf, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = v;
}

extern "C" int32_t tb_l2_flush(tb_ctx *ctx) {
    TB_REQUIRE(ctx, "tb_l2_flush: ctx is NULL");
    if (!ctx->d_flush) {
        ctx->flush_bytes = (size_t)256 << 20;   // 256 MiB > 126 MB L2
        TB_CUDA(cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    }
    int64_t n = (int64_t)(ctx->flush_bytes / sizeof(double));
    TB_LAUNCH(ctx, tb_flush_kernel, ctx->sm_count * 8, 256, 0, (double *)ctx->d_flush, n, 1.0);
    return TB_OK;
}

int32_t tb_ctx_ensure_cgwork(tb_ctx *ctx, int64_t n) {
    int64_t ld = tb_round_up(n, 32);
    if (ld <= ctx->cgwork_ld && ctx->d_cgwork) return TB_OK;
    if (ctx->peer.win)
        return tb_fail(TB_ERR_INVALID, "CG work vectors of %lld rows are mapped by the peer ranks (tb_peer_export); an operator with %lld "
                       "columns cannot be solved on this context", (long long)ctx->cgwork_ld, (long long)n);
    if (ctx->d_cgwork) {
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        TB_CUDA(cudaFree(ctx->d_cgwork));
        ctx->d_cgwork = nullptr;
    }
    TB_CUDA(cudaMalloc(&ctx->d_cgwork, sizeof(double) * 4 * (size_t)ld));   // r | p | Ap | p' (ping-pong partner of the fused-p experiment)
    TB_CUDA(cudaMemsetAsync(ctx->d_cgwork, 0, sizeof(double) * 4 * (size_t)ld, ctx->stream));
    ctx->cgwork_ld = ld;
    return TB_OK;
}

// ---- vectors ----------------------------------------------------------------------------------

extern "C" int32_t tb_vec_create(tb_ctx *ctx, int64_t n, int32_t ncols, tb_vec **out) {
    TB_REQUIRE(ctx && out, "tb_vec_create: NULL argument");
    TB_REQUIRE(n >= 0 && ncols >= 1, "tb_vec_create: bad size n=%lld ncols=%d", (long long)n, ncols);
    *out = nullptr;
    TB_CUDA(cudaSetDevice(ctx->device));
    tb_vec *v = new (std::nothrow) tb_vec();
    if (!v) return tb_fail(TB_ERR_NOMEM, "tb_vec_create: host allocation failed");
    v->ctx = ctx;
    v->n = n;
    v->ncols = ncols;
    v->ld = tb_round_up(n > 0 ? n : 1, 32);
    size_t bytes = sizeof(double) * (size_t)v->ld * (size_t)ncols;
    cudaError_t e = cudaMalloc(&v->d, bytes);
    if (e != cudaSuccess) {
        delete v;
        return tb_fail(TB_ERR_NOMEM, "tb_vec_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    TB_CUDA(cudaMemsetAsync(v->d, 0, bytes, ctx->stream));
    *out = v;
    return TB_OK;
}

extern "C" int32_t tb_vec_destroy(tb_vec *v) {
    if (!v) return TB_OK;
    cudaSetDevice(v->ctx->device);
    cudaStreamSynchronize(v->ctx->stream);
    cudaFree(v->d);
    delete v;
    return TB_OK;
}

extern "C" int32_t tb_vec_sizes(const tb_vec *v, int64_t *n, int32_t *ncols) {
    TB_REQUIRE(v, "tb_vec_sizes: v is NULL");
    if (n) *n = v->n;
    if (ncols) *ncols = v->ncols;
    return TB_OK;
}

extern "C" int32_t tb_vec_upload(tb_vec *v, const double *host) {
    TB_REQUIRE(v && host, "tb_vec_upload: NULL argument");
    TB_DEV(v->ctx);
    TB_CUDA(cudaMemcpy2DAsync(v->d, sizeof(double) * v->ld, host, sizeof(double) * v->n, sizeof(double) * v->n,
                              v->ncols, cudaMemcpyHostToDevice, v->ctx->stream));
    TB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_vec_download(const tb_vec *v, double *host) {
    TB_REQUIRE(v && host, "tb_vec_download: NULL argument");
    TB_DEV(v->ctx);
    TB_CUDA(cudaMemcpy2DAsync(host, sizeof(double) * v->n, v->d, sizeof(double) * v->ld, sizeof(double) * v->n,
                              v->ncols, cudaMemcpyDeviceToHost, v->ctx->stream));
    TB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_vec_upload_col(tb_vec *v, int32_t col, const double *host, int64_t offset, int64_t count) {
    TB_REQUIRE(v && host, "tb_vec_upload_col: NULL argument");
    TB_REQUIRE(col >= 0 && col < v->ncols && offset >= 0 && count >= 0 && offset + count <= v->n,
               "tb_vec_upload_col: range out of bounds");
    TB_DEV(v->ctx);
    TB_CUDA(cudaMemcpyAsync(v->d + (size_t)col * v->ld + offset, host, sizeof(double) * count, cudaMemcpyHostToDevice,
                            v->ctx->stream));
    TB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_vec_download_col(const tb_vec *v, int32_t col, double *host, int64_t offset, int64_t count) {
    TB_REQUIRE(v && host, "tb_vec_download_col: NULL argument");
    TB_REQUIRE(col >= 0 && col < v->ncols && offset >= 0 && count >= 0 && offset + count <= v->n,
               "tb_vec_download_col: range out of bounds");
    TB_DEV(v->ctx);
    TB_CUDA(cudaMemcpyAsync(host, v->d + (size_t)col * v->ld + offset, sizeof(double) * count, cudaMemcpyDeviceToHost,
                            v->ctx->stream));
    TB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_vec_fill(tb_vec *v, int32_t col, double value) {
    TB_REQUIRE(v, "tb_vec_fill: v is NULL");
    TB_REQUIRE(col >= 0 && col < v->ncols, "tb_vec_fill: column out of range");
    TB_DEV(v->ctx);
    if (v->n == 0) return TB_OK;
    tb_ctx *ctx = v->ctx;
    TB_LAUNCH(ctx, tb_flush_kernel, tb_grid_for(ctx, v->n, 256, 8), 256, 0, v->d + (size_t)col * v->ld, v->n, value);
    return TB_OK;
}

extern "C" int32_t tb_vec_copy(tb_vec *dst, int32_t dcol, const tb_vec *src, int32_t scol) {
    TB_REQUIRE(dst && src, "tb_vec_copy: NULL argument");
    TB_REQUIRE(dst->n == src->n && dcol >= 0 && dcol < dst->ncols && scol >= 0 && scol < src->ncols,
               "tb_vec_copy: shape mismatch");
    TB_DEV(dst->ctx);
    TB_CUDA(cudaMemcpyAsync(dst->d + (size_t)dcol * dst->ld, src->d + (size_t)scol * src->ld, sizeof(double) * src->n,
                            cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return TB_OK;
}

__global__ void tb_axpy_kernel(double *__restrict__ y, const double *__restrict__ x, double a, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] += a * x[i];
}

extern "C" int32_t tb_vec_axpy(tb_vec *y, int32_t ycol, double a, const tb_vec *x, int32_t xcol) {
    TB_REQUIRE(y && x, "tb_vec_axpy: NULL argument");
    TB_REQUIRE(ycol >= 0 && ycol < y->ncols && xcol >= 0 && xcol < x->ncols && x->n == y->n, "tb_vec_axpy: shape mismatch");
    TB_DEV(y->ctx);
    if (y->n == 0) return TB_OK;
    tb_ctx *ctx = y->ctx;
    TB_LAUNCH(ctx, tb_axpy_kernel, tb_grid_for(ctx, y->n, 256, 8), 256, 0, y->d + (size_t)ycol * y->ld,
              x->d + (size_t)xcol * x->ld, a, y->n);
    return TB_OK;
}

extern "C" int32_t tb_vec_devptr(const tb_vec *v, int32_t col, void **ptr, int64_t *ld) {
    TB_REQUIRE(v && ptr, "tb_vec_devptr: NULL argument");
    TB_REQUIRE(col >= 0 && col < v->ncols, "tb_vec_devptr: column out of range");
    *ptr = v->d + (size_t)col * v->ld;
    if (ld) *ld = v->ld;
    return TB_OK;
}

extern "C" int32_t tb_profile_enable(tb_ctx *ctx, int32_t on) {
    TB_REQUIRE(ctx, "tb_profile_enable: ctx is NULL");
    if (on && !ctx->prof_ev[0])
        for (int i = 0; i < 2 * TB_PROF_MAX; i++) TB_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
    ctx->profile = on != 0;
    ctx->prof_spmv_ms = 0.0;
    ctx->prof_spmv_n = 0;
    return TB_OK;
}

extern "C" int32_t tb_profile_get(tb_ctx *ctx, double *spmv_ms_total, int64_t *spmv_launches) {
    TB_REQUIRE(ctx && spmv_ms_total && spmv_launches, "tb_profile_get: NULL argument");
    *spmv_ms_total = ctx->prof_spmv_ms;
    *spmv_launches = ctx->prof_spmv_n;
    return TB_OK;
}

extern "C" int32_t tb_cg_set_exact_dot(tb_ctx *ctx, int32_t on) {
    TB_REQUIRE(ctx, "tb_cg_set_exact_dot: ctx is NULL");
    ctx->exact_dot = on != 0;
    return TB_OK;
}

extern "C" int32_t tb_assembly_set_mode(tb_ctx *ctx, int32_t mode) {
    TB_REQUIRE(ctx, "tb_assembly_set_mode: ctx is NULL");
    TB_REQUIRE(mode == 0 || mode == 2, "tb_assembly_set_mode: mode must be 0 (atomic scatter) or 2 (ordered gather)");
    ctx->assembly_mode = mode;
    return TB_OK;
}

extern "C" int32_t tb_assembly_info(tb_ctx *ctx, int32_t *mode_requested, int32_t *mode_last_used, int32_t *chunks_last) {
    TB_REQUIRE(ctx, "tb_assembly_info: ctx is NULL");
    if (mode_requested) *mode_requested = ctx->assembly_mode;
    if (mode_last_used) *mode_last_used = ctx->assembly_last_mode;
    if (chunks_last) *chunks_last = ctx->assembly_last_chunks;
    return TB_OK;
}

extern "C" int32_t tb_assembly_set_scratch_budget(tb_ctx *ctx, int64_t bytes) {
    TB_REQUIRE(ctx && bytes >= 0, "tb_assembly_set_scratch_budget: bad argument");
    ctx->ea_budget_bytes = (size_t)bytes;
    return TB_OK;
}
