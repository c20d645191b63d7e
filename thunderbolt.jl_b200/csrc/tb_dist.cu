// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference has no distributed path at all (shared-memory threads only); this is the
// B200-native addition: rows are partitioned by dof ownership (contiguous dof ranges = z-slabs of a
// structured grid), each SpMV is preceded by a halo exchange of the boundary entries of x, and every
// CG dot product is completed by an 8-byte all-reduce (tb_cg.cu).
#include "tb_internal.cuh"

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

extern "C" int32_t tb_comm_unique_id(void *out128) {
    TB_REQUIRE(out128, "tb_comm_unique_id: out is NULL");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    TB_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return TB_OK;
}

extern "C" int32_t tb_ctx_comm_init(tb_ctx *ctx, int32_t rank, int32_t nranks, const void *nccl_unique_id) {
    TB_REQUIRE(ctx && nccl_unique_id, "tb_ctx_comm_init: NULL argument");
    TB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "tb_ctx_comm_init: bad rank %d of %d", rank, nranks);
    TB_REQUIRE(!ctx->has_comm, "tb_ctx_comm_init: communicator already initialised");
    TB_DEV(ctx);
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id, sizeof(id));
    TB_NCCL(ncclCommInitRank(&ctx->comm, nranks, id, rank));
    ctx->has_comm = true;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return TB_OK;
}

extern "C" int32_t tb_comm_barrier(tb_ctx *ctx) {
    TB_REQUIRE(ctx, "tb_comm_barrier: ctx is NULL");
    TB_DEV(ctx);
    if (ctx->has_comm && ctx->nranks > 1)
        TB_NCCL(ncclAllReduce(ctx->d_scalar + 8, ctx->d_scalar + 8, 1, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

extern "C" int32_t tb_comm_allreduce_max(tb_ctx *ctx, double *value_inout) {
    TB_REQUIRE(ctx && value_inout, "tb_comm_allreduce_max: NULL argument");
    TB_DEV(ctx);
    if (!(ctx->has_comm && ctx->nranks > 1)) return TB_OK;
    ctx->h_scalar[1] = *value_inout;
    TB_CUDA(cudaMemcpyAsync(ctx->d_scalar + 1, ctx->h_scalar + 1, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TB_NCCL(ncclAllReduce(ctx->d_scalar + 1, ctx->d_scalar + 1, 1, ncclDouble, ncclMax, ctx->comm, ctx->stream));
    TB_CUDA(cudaMemcpyAsync(ctx->h_scalar + 1, ctx->d_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    *value_inout = ctx->h_scalar[1];
    return TB_OK;
}

extern "C" int32_t tb_csr_set_halo(tb_csr *A, int32_t nneigh, const int32_t *neigh_ranks, const int64_t *send_ptr,
                                   const int64_t *send_rows, const int64_t *recv_ptr) {
    TB_REQUIRE(A, "tb_csr_set_halo: matrix is NULL");
    tb_pattern *p = A->pat;
    tb_ctx *ctx = p->ctx;
    TB_DEV(ctx);
    tb_halo &h = p->halo;
    cudaFree(h.d_send_rows);
    cudaFree(h.d_sendbuf);
    h = tb_halo();
    if (nneigh == 0) return TB_OK;
    TB_REQUIRE(neigh_ranks && send_ptr && send_rows && recv_ptr, "tb_csr_set_halo: NULL argument");
    TB_REQUIRE(ctx->has_comm, "tb_csr_set_halo: context has no communicator");
    h.nneigh = nneigh;
    h.ranks.assign(neigh_ranks, neigh_ranks + nneigh);
    h.send_ptr.assign(send_ptr, send_ptr + nneigh + 1);
    h.recv_ptr.assign(recv_ptr, recv_ptr + nneigh + 1);
    h.nsend = send_ptr[nneigh];
    h.nrecv = recv_ptr[nneigh];
    TB_REQUIRE(h.nrecv == p->ncols - p->nrows, "tb_csr_set_halo: recv count %lld != ghost columns %lld",
               (long long)h.nrecv, (long long)(p->ncols - p->nrows));
    std::vector<int> rows((size_t)h.nsend);
    for (int64_t i = 0; i < h.nsend; i++) {
        TB_REQUIRE(send_rows[i] >= 0 && send_rows[i] < p->nrows, "tb_csr_set_halo: send row out of range");
        rows[(size_t)i] = (int)send_rows[i];
    }
    TB_CUDA(cudaMalloc(&h.d_send_rows, sizeof(int) * (size_t)(h.nsend + 1)));
    TB_CUDA(cudaMalloc(&h.d_sendbuf, sizeof(double) * (size_t)(h.nsend + 1)));
    TB_CUDA(cudaMemcpy(h.d_send_rows, rows.data(), sizeof(int) * (size_t)h.nsend, cudaMemcpyHostToDevice));
    return TB_OK;
}

__global__ void k_halo_pack(const double *__restrict__ x, const int *__restrict__ rows, double *__restrict__ buf, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        buf[i] = x[rows[i]];
}

// Refreshes the ghost block x[nrows .. ncols) from the owning ranks.  Every rank calls this the same
// number of times (the CG `done` flag is identical on all ranks), so the send/recv pairs always match.
int32_t tb_halo_exchange(tb_ctx *ctx, const tb_pattern *pat, double *x) {
    const tb_halo &h = pat->halo;
    if (h.nneigh == 0) return TB_OK;
    TB_LAUNCH(ctx, k_halo_pack, tb_grid_for(ctx, h.nsend, 256, 4), 256, 0, x, h.d_send_rows, h.d_sendbuf, h.nsend);
    TB_NCCL(ncclGroupStart());
    for (int i = 0; i < h.nneigh; i++) {
        int64_t ns = h.send_ptr[i + 1] - h.send_ptr[i], nr = h.recv_ptr[i + 1] - h.recv_ptr[i];
        if (ns > 0) TB_NCCL(ncclSend(h.d_sendbuf + h.send_ptr[i], (size_t)ns, ncclDouble, h.ranks[i], ctx->comm, ctx->stream));
        if (nr > 0) TB_NCCL(ncclRecv(x + pat->nrows + h.recv_ptr[i], (size_t)nr, ncclDouble, h.ranks[i], ctx->comm, ctx->stream));
    }
    TB_NCCL(ncclGroupEnd());
    return TB_OK;
}
