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
    TB_CUDA(cudaMemcpyAsync(h.d_send_rows, rows.data(), sizeof(int) * (size_t)h.nsend, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
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

// =====================================================================================================
// Peer-memory path (NVLink stores between the ranks of one box), see tb_internal.cuh "tb_peer_window".
// =====================================================================================================
struct tb_peer_blob {   // what every rank publishes to the others (TB_PEER_BLOB_BYTES)
    cudaIpcMemHandle_t win;      // 64 B
    cudaIpcMemHandle_t cgwork;   // 64 B
    int64_t cgwork_ld;
    int64_t reserved[3];
};
static_assert(sizeof(tb_peer_blob) == TB_PEER_BLOB_BYTES, "tb_peer_blob layout");

extern "C" int32_t tb_peer_export(tb_ctx *ctx, int64_t ncols, void *blob_out) {
    TB_REQUIRE(ctx && blob_out && ncols > 0, "tb_peer_export: bad argument");
    TB_REQUIRE(ctx->has_comm && ctx->nranks > 1 && ctx->nranks <= TB_MAX_RANKS, "tb_peer_export: needs a communicator of 2..%d ranks", TB_MAX_RANKS);
    TB_REQUIRE(!ctx->peer.on && !ctx->peer.win, "tb_peer_export: already exported");
    TB_DEV(ctx);
    TB_TRY(tb_ctx_ensure_cgwork(ctx, ncols));
    TB_CUDA(cudaMalloc(&ctx->peer.win, sizeof(tb_peer_window)));
    TB_CUDA(cudaMemsetAsync(ctx->peer.win, 0, sizeof(tb_peer_window), ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    tb_peer_blob b;
    memset(&b, 0, sizeof(b));
    TB_CUDA(cudaIpcGetMemHandle(&b.win, ctx->peer.win));
    TB_CUDA(cudaIpcGetMemHandle(&b.cgwork, ctx->d_cgwork));
    b.cgwork_ld = ctx->cgwork_ld;
    memcpy(blob_out, &b, sizeof(b));
    return TB_OK;
}

extern "C" int32_t tb_peer_attach(tb_ctx *ctx, const void *blobs, int32_t nranks) {
    TB_REQUIRE(ctx && blobs, "tb_peer_attach: NULL argument");
    TB_REQUIRE(ctx->peer.win && !ctx->peer.on, "tb_peer_attach: call tb_peer_export first (once)");
    TB_REQUIRE(nranks == ctx->nranks, "tb_peer_attach: %d blobs for %d ranks", nranks, ctx->nranks);
    TB_DEV(ctx);
    const tb_peer_blob *b = static_cast<const tb_peer_blob *>(blobs);
    tb_peer &P = ctx->peer;
    for (int q = 0; q < nranks; q++) {
        if (q == ctx->rank) {
            P.peer_win[q] = P.win;
            P.peer_cgwork[q] = ctx->d_cgwork;
        } else {
            void *w = nullptr, *c = nullptr;
            TB_CUDA(cudaIpcOpenMemHandle(&w, b[q].win, cudaIpcMemLazyEnablePeerAccess));
            TB_CUDA(cudaIpcOpenMemHandle(&c, b[q].cgwork, cudaIpcMemLazyEnablePeerAccess));
            P.peer_win[q] = static_cast<tb_peer_window *>(w);
            P.peer_cgwork[q] = static_cast<double *>(c);
        }
        P.peer_ld[q] = b[q].cgwork_ld;
    }
    TB_CUDA(cudaMalloc(&P.d_peer_win, sizeof(tb_peer_window *) * TB_MAX_RANKS));
    TB_CUDA(cudaMemcpyAsync(P.d_peer_win, P.peer_win, sizeof(tb_peer_window *) * TB_MAX_RANKS, cudaMemcpyHostToDevice, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    P.on = true;
    return TB_OK;
}

extern "C" int32_t tb_peer_enabled(tb_ctx *ctx, int32_t *on) {
    TB_REQUIRE(ctx && on, "tb_peer_enabled: NULL argument");
    *on = ctx->peer.on ? 1 : 0;
    return TB_OK;
}

// must run on every rank before the context goes away (peers may still hold mappings of our memory)
int32_t tb_peer_release(tb_ctx *ctx) {
    tb_peer &P = ctx->peer;
    for (int q = 0; q < TB_MAX_RANKS; q++) {
        if (P.on && q != ctx->rank && q < ctx->nranks) {
            if (P.peer_win[q]) cudaIpcCloseMemHandle(P.peer_win[q]);
            if (P.peer_cgwork[q]) cudaIpcCloseMemHandle(P.peer_cgwork[q]);
        }
        P.peer_win[q] = nullptr;
        P.peer_cgwork[q] = nullptr;
    }
    cudaFree(P.d_peer_win);
    cudaFree(P.win);
    P = tb_peer();
    return TB_OK;
}

extern "C" int32_t tb_csr_set_halo_peer(tb_csr *A, const int64_t *dst_off, const int32_t *dst_slot) {
    TB_REQUIRE(A && dst_off && dst_slot, "tb_csr_set_halo_peer: NULL argument");
    tb_halo &h = A->pat->halo;
    tb_ctx *ctx = A->pat->ctx;
    TB_REQUIRE(ctx->peer.on, "tb_csr_set_halo_peer: peers are not attached");
    TB_REQUIRE(h.nneigh <= TB_MAX_RANKS, "tb_csr_set_halo_peer: too many neighbours");
    if (h.nneigh == 0) {   // a rank whose rows reference no foreign dof: nothing to push or to wait for
        h.contiguous = true;
        h.peer_ready = true;
        return TB_OK;
    }
    TB_REQUIRE(A->pat->ncols <= ctx->cgwork_ld, "tb_csr_set_halo_peer: operator has more columns than the exported work vectors");
    h.dst_off.assign(dst_off, dst_off + h.nneigh);
    h.dst_slot.assign(dst_slot, dst_slot + h.nneigh);
    for (int i = 0; i < h.nneigh; i++) {
        const int64_t ns = h.send_ptr[i + 1] - h.send_ptr[i];
        TB_REQUIRE(h.dst_slot[i] >= 0 && h.dst_slot[i] < TB_MAX_RANKS, "tb_csr_set_halo_peer: bad flag slot");
        TB_REQUIRE(h.dst_off[i] >= 0 && h.dst_off[i] + ns <= ctx->peer.peer_ld[h.ranks[i]],
                   "tb_csr_set_halo_peer: destination range outside the neighbour's work vector");
    }
    // slab partitions send one run of consecutive rows to each neighbour: then the p-update kernel can push in line
    h.contiguous = true;
    h.range_lo.assign((size_t)h.nneigh, 0);
    {
        std::vector<int> rows((size_t)h.nsend);
        TB_CUDA(cudaMemcpy(rows.data(), h.d_send_rows, sizeof(int) * (size_t)h.nsend, cudaMemcpyDeviceToHost));
        for (int i = 0; i < h.nneigh; i++) {
            const int64_t a = h.send_ptr[i], b = h.send_ptr[i + 1];
            h.range_lo[(size_t)i] = b > a ? rows[(size_t)a] : 0;
            for (int64_t k = a; k < b; k++)
                if (rows[(size_t)k] != rows[(size_t)a] + (k - a)) h.contiguous = false;
        }
    }
    h.peer_ready = true;
    return TB_OK;
}

extern "C" int32_t tb_csr_halo_fused_capable(const tb_csr *A, int32_t *capable) {
    TB_REQUIRE(A && capable, "tb_csr_halo_fused_capable: NULL argument");
    const tb_halo &h = A->pat->halo;
    // a rank without neighbours never pushes or waits: it can follow either path
    *capable = (A->pat->ctx->p2p_fused && (h.nneigh == 0 || (h.peer_ready && h.contiguous))) ? 1 : 0;
    return TB_OK;
}

extern "C" int32_t tb_csr_set_halo_fused(tb_csr *A, int32_t on) {
    TB_REQUIRE(A, "tb_csr_set_halo_fused: matrix is NULL");
    tb_halo &h = A->pat->halo;
    if (on) TB_REQUIRE(h.nneigh == 0 || (h.peer_ready && h.contiguous), "tb_csr_set_halo_fused: this rank's send lists are not contiguous runs");
    h.fused = on != 0;
    return TB_OK;
}

// arguments of the in-line push (k_cg_p_fused) for the next halo epoch
int32_t tb_halo_push_args(tb_ctx *ctx, const tb_pattern *pat, tb_push_args *out, tb_hwait_args *wait_out) {
    const tb_halo &h = pat->halo;
    tb_peer &P = ctx->peer;
    const unsigned long long epoch = ++P.halo_epoch;
    out->n = h.nneigh;
    out->epoch = epoch;
    for (int i = 0; i < h.nneigh; i++) {
        const int q = h.ranks[i];
        out->lo[i] = h.range_lo[(size_t)i];
        out->len[i] = h.send_ptr[i + 1] - h.send_ptr[i];
        out->dst[i] = P.peer_cgwork[q] + P.peer_ld[q] + h.dst_off[i];
        out->flag[i] = &P.peer_win[q]->hflag[h.dst_slot[i]];
    }
    wait_out->hflag = P.win->hflag;
    wait_out->n = h.nneigh;
    wait_out->epoch = epoch;
    wait_out->err = &P.win->err;
    wait_out->stat = &P.win->halo_wait_ns;
    return TB_OK;
}

struct HaloPush {
    int n;
    long long begin[TB_MAX_RANKS + 1];
    double *dst[TB_MAX_RANKS];
    unsigned long long *flag[TB_MAX_RANKS];
};

// dst_k[i - begin_k] = p[rows[i]] for every neighbour k, then (last block) raise the neighbours' flags
__global__ void __launch_bounds__(256) k_halo_push(const double *__restrict__ p, const int *__restrict__ rows, const HaloPush hp,
                                                   unsigned long long epoch, unsigned *ticket, const CGState *st) {
    if (st && st->done) return;
    const long long n = hp.begin[hp.n];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int k = 0;
        while (i >= hp.begin[k + 1]) k++;
        hp.dst[k][i - hp.begin[k]] = p[rows[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicInc(ticket, gridDim.x - 1);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            for (int k = 0; k < hp.n; k++) *(volatile unsigned long long *)hp.flag[k] = epoch;
        }
    }
}

// Pushes the boundary entries of p (the CG direction vector inside ctx->d_cgwork) into the neighbours' ghost blocks.
// wait_out: what the following SpMV on this rank has to wait for.
int32_t tb_halo_push(tb_ctx *ctx, const tb_pattern *pat, const double *p, const CGState *st, tb_hwait_args *wait_out) {
    const tb_halo &h = pat->halo;
    tb_peer &P = ctx->peer;
    const unsigned long long epoch = ++P.halo_epoch;
    HaloPush hp;
    hp.n = h.nneigh;
    for (int i = 0; i < h.nneigh; i++) {
        const int q = h.ranks[i];
        hp.begin[i] = h.send_ptr[i];
        hp.dst[i] = P.peer_cgwork[q] + P.peer_ld[q] + h.dst_off[i];       // p is the second work vector
        hp.flag[i] = &P.peer_win[q]->hflag[h.dst_slot[i]];
    }
    hp.begin[h.nneigh] = h.nsend;
    if (h.nneigh > 0)
        TB_LAUNCH(ctx, k_halo_push, tb_grid_for(ctx, h.nsend, 256, 4), 256, 0, p, h.d_send_rows, hp, epoch, ctx->d_ticket + 4, st);
    wait_out->hflag = P.win->hflag;
    wait_out->n = h.nneigh;
    wait_out->epoch = epoch;
    wait_out->err = &P.win->err;
    wait_out->stat = &P.win->halo_wait_ns;
    return TB_OK;
}

// time CTA 0 spent waiting for peers since the last reset: {all-reduce collects, halo flags} in ms, and the wait counts
extern "C" int32_t tb_peer_stats(tb_ctx *ctx, double *ar_wait_ms, int64_t *ar_waits, double *halo_wait_ms, int64_t *halo_waits, int32_t reset) {
    TB_REQUIRE(ctx, "tb_peer_stats: ctx is NULL");
    unsigned long long v[4] = {0, 0, 0, 0};
    if (ctx->peer.on) {
        TB_DEV(ctx);
        TB_CUDA(cudaStreamSynchronize(ctx->stream));
        TB_CUDA(cudaMemcpy(v, &ctx->peer.win->ar_wait_ns, sizeof(v), cudaMemcpyDeviceToHost));
        if (reset) TB_CUDA(cudaMemset(&ctx->peer.win->ar_wait_ns, 0, sizeof(v)));
    }
    if (ar_wait_ms) *ar_wait_ms = (double)v[0] * 1e-6;
    if (ar_waits) *ar_waits = (int64_t)v[1];
    if (halo_wait_ms) *halo_wait_ms = (double)v[2] * 1e-6;
    if (halo_waits) *halo_waits = (int64_t)v[3];
    return TB_OK;
}
