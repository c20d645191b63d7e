// Internal declarations shared by the translation units of libtbolt_b200.so.
// Nothing in here is part of the ABI; the ABI is include/tbolt_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>
#include <utility>
#include <new>

#include "../../include/tbolt_b200.h"

#define TB_SLICE 32            // SELL slice height = warp size
#define TB_MAX_PARTIALS 4096   // upper bound on blocks of any reducing kernel
#define TB_MAXROW 512          // max nonzeros per row the device pattern builder supports (apex rows of an LV mesh reach 2*nc + 3)
#define TB_TMA_WCAP 48         // widest slice (entries per row) that is staged through shared memory by the bulk-async SpMV;
                               // wider slices (a handful of high-valence rows) take the LDG row kernel inside the same sweep
#define TB_PROF_MAX 256        // SpMV launches per solve that the profiler brackets with events

int32_t tb_fail(int32_t code, const char *fmt, ...);

#define TB_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return tb_fail(e__ == cudaErrorMemoryAllocation ? TB_ERR_NOMEM : TB_ERR_CUDA, "%s:%d: %s -> %s", \
                           __FILE__, __LINE__, #call, cudaGetErrorString(e__));                         \
    } while (0)

#define TB_NCCL(call)                                                                                   \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess)                                                                         \
            return tb_fail(TB_ERR_COMM, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
    } while (0)

#define TB_REQUIRE(cond, ...)                                  \
    do {                                                       \
        if (!(cond)) return tb_fail(TB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define TB_TRY(call)                   \
    do {                               \
        int32_t s__ = (call);          \
        if (s__ != TB_OK) return s__;  \
    } while (0)

// launch bookkeeping: every kernel launch of the library goes through TB_LAUNCH so that
// tb_launch_count() is an honest count and launch errors surface immediately.
#define TB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
    do {                                                                                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                \
        (ctx)->launches++;                                                              \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return tb_fail(TB_ERR_CUDA, "%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel, \
                           cudaGetErrorString(e__));                                    \
    } while (0)

// Grid of a persistent grid-stride kernel: exactly the number of CTAs that are resident at once
// (occupancy query, cached per call site), so the sweep is ONE balanced wave -- a grid larger than the
// resident capacity runs a ragged second wave at a fraction of the bandwidth (measured: 1184 CTAs with
// 888 resident cost the SpMV 25 % of its time).  `need` = CTAs that have work at all.
#define TB_GRID(ctx, kernel, block, smem, need)                                                     \
    ([&]() -> int {                                                                                 \
        static int per_sm__ = 0;                                                                    \
        if (!per_sm__) {                                                                            \
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm__, kernel, block, smem) != cudaSuccess || per_sm__ < 1) \
                per_sm__ = 1;                                                                       \
        }                                                                                           \
        int64_t cap__ = (int64_t)(ctx)->sm_count * per_sm__;                                        \
        if (cap__ > TB_MAX_PARTIALS) cap__ = TB_MAX_PARTIALS;                                       \
        int64_t n__ = (need);                                                                       \
        if (n__ < 1) n__ = 1;                                                                       \
        return (int)(n__ < cap__ ? n__ : cap__);                                                    \
    }())

// ---- peer-memory (NVLink) communication between the one-process-per-GPU ranks of a box ----------------
// Every rank owns one tb_peer_window in its HBM and maps the windows (and CG work vectors) of all other
// ranks through CUDA IPC.  Kernels then communicate with plain stores over NVLink:
//   * dot products: the last block of a reducing kernel stores its rank-local sum into slot [e % SLOTS][rank]
//     of EVERY rank's window, fences, then stores the epoch e into the matching flag; the consumer (one
//     thread) waits for all flags of epoch e and adds the values in rank order -- identical bits on all ranks;
//   * halo: the owner stores its boundary entries of p straight into the neighbour's ghost block of p and then
//     raises hflag[slot] = epoch in the neighbour's window; the neighbour's SpMV waits for it at kernel start.
#define TB_MAX_RANKS 16
#define TB_AR_SLOTS 8
struct tb_peer_window {
    double val[TB_AR_SLOTS][TB_MAX_RANKS];
    double val_lo[TB_AR_SLOTS][TB_MAX_RANKS];   // low words of the partial sums (exact-dot mode, see tb_acc)
    unsigned long long flag[TB_AR_SLOTS][TB_MAX_RANKS];
    unsigned long long hflag[TB_MAX_RANKS];
    unsigned long long err;          // set by a waiter that timed out
    // instrumentation (CTA 0 only, so the numbers are one CTA's view of the critical path): time spent waiting for the
    // other ranks' partial sums / halo flags, and how often
    unsigned long long ar_wait_ns, ar_waits, halo_wait_ns, halo_waits;
    unsigned long long pad[11];
};
// halo wait of an SpMV kernel (n == 0: nothing to wait for)
struct tb_hwait_args {
    const unsigned long long *hflag;
    int n;
    unsigned long long epoch;
    unsigned long long *err;
    unsigned long long *stat;   // -> {halo_wait_ns, halo_waits} of this rank's window (nullable)
};
struct tb_peer {
    bool on = false;
    tb_peer_window *win = nullptr;                     // this rank's window
    tb_peer_window *peer_win[TB_MAX_RANKS] = {};       // windows of all ranks as seen from this GPU
    double *peer_cgwork[TB_MAX_RANKS] = {};            // base of every rank's CG work vectors (r | p | Ap)
    int64_t peer_ld[TB_MAX_RANKS] = {};
    tb_peer_window **d_peer_win = nullptr;             // device copy of peer_win
    unsigned long long ar_epoch = 0, halo_epoch = 0;
    tb_hwait_args hw_next = {nullptr, 0, 0, nullptr, nullptr};   // fused path: what the next SpMV waits for (set when k_cg_p_fused is enqueued)
};
// what a reducing / consuming kernel needs to take part in the window all-reduce (wins == nullptr: NCCL path)
struct tb_ar_args {
    tb_peer_window *const *wins;
    int rank, nranks, slot;
    unsigned long long epoch;
};
// in-line halo push of k_cg_p_fused: neighbour k wants rows [lo, lo+len) of p at dst[0..len)
struct tb_push_args {
    int n;
    long long lo[TB_MAX_RANKS], len[TB_MAX_RANKS];
    double *dst[TB_MAX_RANKS];
    unsigned long long *flag[TB_MAX_RANKS];
    unsigned long long epoch;
};

struct tb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    size_t total_mem = 0;
    int cc_major = 0, cc_minor = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    // scratch for deterministic two-stage reductions (block partials + last-block ticket)
    double *d_partials = nullptr;   // 4 * TB_MAX_PARTIALS, followed by the same layout again for the low words (exact-dot mode)
    unsigned *d_ticket = nullptr;   // 8 tickets
    // CG workspace (grown on demand): r, p, Ap as one tb_vec-like allocation
    double *d_cgwork = nullptr;
    int64_t cgwork_ld = 0;
    int pcg_variant = 0;              // register-resident persistent CG: 1 = three grid.sync per iteration, 2 = two flag barriers (0: env TB_PCG_V, default 2)
    struct CGState *d_cg = nullptr;   // device scalars of the running solve
    struct CGState *h_cg = nullptr;   // pinned mirror
    double *d_scalar = nullptr;       // small device scratch (16 doubles)
    double *d_dconst = nullptr;       // constant diffusion coefficient of the running assembly (<= 16 doubles, stream-ordered uploads)
    double *h_scalar = nullptr;       // pinned
    double *d_dinv = nullptr;         // Jacobi preconditioner: 1/diag(A) of the operator being solved
    int64_t dinv_len = 0;
    struct tb_bj *bj = nullptr;       // block-Jacobi plan + dense inverses (tb_precond.cu)
    uint64_t pc_uid = 0, pc_version = ~0ull;   // operator the cached Chebyshev bound / block inverses were built from
    int pc_kind = -1;
    double pc_lmax = 0.0;
    int cheb_degree = 8;              // Chebyshev preconditioner: polynomial degree and lmax/lmin of the target interval
    double cheb_ratio = 30.0;
    double *d_pcwork = nullptr;       // z | d | res | w of the general preconditioners
    int64_t pcwork_ld = 0;
    void *d_flush = nullptr;
    size_t flush_bytes = 0;
    int assembly_mode = 2;            // 0: fp64-atomic scatter, 2: element matrices + ordered row gather (deterministic; env TB_ASSEMBLY_MODE)
    int assembly_last_mode = -1;      // what the last assembly call really ran (2 falls back to 0 when the scratch cannot fit)
    int assembly_last_chunks = 0;
    size_t ea_budget_bytes = (size_t)8 << 30;   // scratch for element matrices per chunk (env TB_EA_BUDGET_MB)
    struct tb_elem_tables *d_tables[4][5] = {};   // quadrature/shape tables per (cell type, order), uploaded once (tb_assembly.cu)
    int tables_nq[4][5] = {};
    cudaStream_t stage_stream = nullptr;   // output staging (tb_vec_stage_col)
    cudaEvent_t stage_ev = nullptr;
    void *d_ea = nullptr;             // cached scratch of the vector (per-step source) assembly
    size_t ea_bytes = 0;
    int spmv_variant = 1;             // 0: LDG kernel, 1..: bulk-async (TMA) staged kernel configurations (env TB_SPMV_VARIANT)
    int spmv_compress = 1;            // use the compressed column stream in the staged kernels (env TB_SPMV_COMPRESS)
    int64_t last_cg_iters = 4;        // launch-ahead hint for the next solve
    int cg_persistent = 1;            // 0: never; 1: auto (small: register-resident kernel, mid-size: TMA kernel); 2: TMA kernel whenever eligible (tests); env TB_CG_PERSISTENT
    int last_cg_persistent = 0;       // path of the last solve: 0 multi-kernel, 1 persistent (registers), 2 persistent (TMA sweep, vectors in L2)
    int spmv_fusep = 0;               // experiment (env TB_SPMV_FUSEP): p = r + beta p formed inside the SpMV's gather, no k_cg_p launch
    int exact_dot = 0;                // 1: CG dot products accumulated in double-double (order-independent after rounding; env TB_DOT_EXACT)
    int p2p_fused = 1;                // multi-GPU peer path: 1 = collects and halo push inside the CG kernels (3 launches per iteration), 0 = separate tiny kernels (env TB_P2P_FUSED)
    int64_t cg_persistent_max_rows = 4000000;   // above this the multi-kernel path is used (env TB_CG_PERSISTENT_MAX_ROWS)
    // per-kernel profiling of the dominant kernel (SpMV inside CG): CUDA events around each launch
    bool profile = false;
    bool cg_deferred = false;         // tb_monodomain_run: the persistent CG paths leave their scalars on the device (no read-back per solve)
    cudaEvent_t prof_ev[2 * TB_PROF_MAX] = {};
    double prof_spmv_ms = 0.0;
    int64_t prof_spmv_n = 0;
    // communicator
    bool has_comm = false;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    tb_peer peer;
    // per-context record of cudaFuncAttributeMaxDynamicSharedMemorySize settings (the attribute is per device, and a
    // process may hold contexts on several devices)
    std::vector<std::pair<const void *, size_t>> smem_set;
};

// device-resident scalars of a CG solve; mirrors the locals of Krylov.jl's cg!
struct CGState {
    double gamma;       // r.r of the current iterate
    double pAp;
    double alpha;
    double beta;
    double gamma_next;
    double rnorm;
    double eps;         // atol + rtol*|r0|
    double atol, rtol;
    double local[2];    // per-rank partial sums awaiting the allreduce (multi-GPU); [1] = low word in exact-dot mode
    long long iter;
    long long itmax;
    int done;           // solved || tired
    int solved;
    int all_solved;     // tb_monodomain_run (deferred read-back): AND of `solved` over the run's solves
    int pad;
    long long iter_sum; // ... and the sum of their iteration counts (k_pcg_fold)
};

struct tb_vec {
    tb_ctx *ctx;
    int64_t n;
    int ncols;
    int64_t ld;   // padded column stride (multiple of 32 doubles)
    double *d;
};

struct tb_mesh {
    tb_ctx *ctx;
    int celltype, nv, dim;
    int64_t ncells, nnodes;
    int64_t ndofs;        // all dofs referenced by celldofs (owned + ghosts)
    int64_t ndofs_owned;  // rows this rank owns (== ndofs on one GPU)
    int *d_conn = nullptr;      // ncells*nv node ids
    int *d_celldofs = nullptr;  // ncells*nv dof ids
    double *d_coords = nullptr; // nnodes*dim
    int *d_node2dof = nullptr;  // nnodes, -1 if the node carries no dof
    int64_t *d_ghost_global = nullptr;
    int64_t nghost = 0;
    int64_t dof_lo = 0;         // global id of local dof 0
    // dof -> (cell, local index) adjacency of the owned dofs, built on first use by the gather assembly
    // (tb_assembly.cu): entry = cell*nv + a, each list sorted ascending = the reference's element order
    mutable int64_t *d_adjptr = nullptr;   // ndofs_owned + 1
    mutable unsigned *d_adj = nullptr;
    mutable unsigned char *d_adjpos = nullptr;   // per adjacency entry (row, cell, a): slot of each of the cell's nv dofs in the row (gather assembly)
    mutable uint64_t adjpos_pat_uid = 0;         // pattern the slots were computed for
    mutable int adjpos_state = 0;                // 0 not tried, 1 built, -1 not applicable (rows wider than 255, or over budget)
    mutable int64_t nadj = 0;
    // per 32-row slice: smallest / largest adjacent cell (host copy), and the chunk plans derived from it
    mutable std::vector<int> slice_cmin, slice_cmax;
    struct GatherPlan {
        int64_t budget_cells = -1;     // the budget this plan was made for
        bool ok = false;
        int64_t rows_per_chunk = 0;    // multiple of 32
        std::vector<int> cmin, cmax;   // per chunk, inclusive; cmin > cmax: no cells
        int64_t max_cells = 0;
    };
    mutable GatherPlan plan[2];        // [0] matrices, [1] vectors
};

struct tb_halo {
    int nneigh = 0;
    std::vector<int> ranks;
    std::vector<int64_t> send_ptr, recv_ptr;
    int *d_send_rows = nullptr;   // concatenated local row ids to pack
    double *d_sendbuf = nullptr;
    int64_t nsend = 0, nrecv = 0;
    // peer-memory push (tb_csr_set_halo_peer): where each neighbour wants our entries inside ITS p vector,
    // and which of its halo flags is ours
    bool peer_ready = false;
    std::vector<int64_t> dst_off;
    std::vector<int> dst_slot;
    bool contiguous = false;          // every neighbour's send list is one run of consecutive rows (slab partitions):
    std::vector<int64_t> range_lo;    // first row of that run -- lets k_cg_p_fused push while it updates
    bool fused = false;               // decision agreed by ALL ranks (tb_csr_set_halo_fused): the fused and the unfused path
                                      // consume different numbers of halo epochs, a per-rank choice would desynchronise them
};

// sparsity pattern shared by M, K and A (sliced ELL image of the reference's CSR pattern)
struct tb_pattern {
    tb_ctx *ctx;
    int64_t nrows, ncols, nnz;
    int64_t nslices;
    int64_t sell_len;             // padded number of stored entries
    int max_width = 0;            // widest slice (entries per row)
    int max_width_tma = 0;        // widest slice among those <= TB_TMA_WCAP (sizes the staged kernel's shared-memory stage)
    int64_t n_wide = 0;           // slices wider than TB_TMA_WCAP
    int64_t *d_wide_slices = nullptr;   // their ids, ascending (tb_spmv.cuh: tb_wide_list)
    int64_t *d_rowptr = nullptr;  // nrows+1 (CSR row pointers, 0-based)
    int64_t *d_slice_ptr = nullptr; // nslices+1 offsets into col/val, multiples of 32
    int *d_col = nullptr;         // sell_len column ids (padding: a valid column, value 0)
    // compressed column stream (tb_csr.cu: one int32 offset per slot when col = row + off for all lanes)
    int *d_ccol = nullptr;
    int64_t *d_cptr = nullptr;    // nslices+1 offsets into d_ccol, in ints, multiples of 4
    int64_t ccol_len = 0;
    int max_ccol_ints = 0;        // largest per-slice stream (sizes the shared-memory stage)
    int refcount = 1;
    uint64_t uid = 0;             // identity of the pattern (the mesh caches data computed against it)
    int *d_diag_slot = nullptr;   // per row: entry slot of the diagonal (-1: none); built on first use by the Jacobi preconditioner
    tb_halo halo;
    // element colouring cache (assembly mode 1)
};

struct tb_csr {
    tb_pattern *pat;
    double *d_val;   // sell_len
    uint64_t version = 0;   // bumped by every entry point that writes values: lets the preconditioners cache their setup
    uint64_t uid = 0;       // distinguishes handles that reuse an address
};
uint64_t tb_next_uid();

struct tb_monodomain {
    tb_ctx *ctx;
    const tb_csr *M, *K;
    tb_csr *A;
    int model;
    double params[40];
    int nparams;
    int phi_idx;
    double atol, rtol;
    int64_t itmax;
    int precond = 0;
    int substeps;
    double threshold;
    const tb_vec *bS;
    int bS_col;
    double dt_last;
    tb_vec *b, *x;
    bool timing;
    cudaEvent_t ev[4];
    double section_ms[3];
    // tb_monodomain_run_host: copy streams and events (created on first use)
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e_phi = nullptr, e_s = nullptr, e_done = nullptr, o_s = nullptr, e_chunk[64] = {};
    int rh_chunks = 16;            // pieces of the phi column in tb_monodomain_run_host (download / chasing upload)
};

// ionic model tables (tb_cells.cuh traits, host side)
static inline bool tb_model_known(int m) { return m == TB_FHN || m == TB_PCG2019 || m == TB_ALIEV_PANFILOV; }
static inline int tb_model_nstates(int m) { return m == TB_PCG2019 ? 7 : 2; }
static inline int tb_model_nparams(int m) { return m == TB_PCG2019 ? 36 : 6; }
static inline int tb_model_phi(int m) { return m == TB_ALIEV_PANFILOV ? 1 : 0; }

static inline int64_t tb_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ double tb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double tb_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- dot-product accumulators ------------------------------------------------------------------------------------
// tb_acc<false>: plain fp64 accumulation (the default; fastest).  tb_acc<true>: double-double accumulation with exact
// products (TwoProd by fma) and error-free additions (TwoSum), i.e. the sum is carried to ~2^-104 relative accuracy and
// only the FINAL value is rounded to fp64.  Two different summation orders then round to the same double except when
// the exact sum lies within ~1e-24 (relative) of a rounding boundary, so the CG scalars -- and with them every
// iterate, the stopping decision and the iteration count -- no longer depend on the grid size, on the number of GPUs
// or on whether the sum was formed sequentially on a CPU (the oracle's "exact" mode).  SURVEY 7 hard part 1.
__device__ __forceinline__ void tb_two_sum(double a, double b, double &s, double &e) {
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
template <bool X> struct tb_acc;
template <> struct tb_acc<false> {
    double hi = 0.0;
    __device__ __forceinline__ void add_prod(double a, double b) { hi += a * b; }
    __device__ __forceinline__ void add(const tb_acc &o) { hi += o.hi; }
    __device__ __forceinline__ void set(double h, double) { hi = h; }
    __device__ __forceinline__ double low() const { return 0.0; }
    __device__ __forceinline__ double value() const { return hi; }
};
template <> struct tb_acc<true> {
    double hi = 0.0, lo = 0.0;
    __device__ __forceinline__ void add_pair(double h, double l) {
        double s, e;
        tb_two_sum(hi, h, s, e);
        e += lo + l;
        hi = s + e;               // renormalise (fast two-sum)
        lo = e - (hi - s);
    }
    __device__ __forceinline__ void add_prod(double a, double b) {
        const double p = a * b;
        add_pair(p, __fma_rn(a, b, -p));
    }
    __device__ __forceinline__ void add(const tb_acc &o) { add_pair(o.hi, o.lo); }
    __device__ __forceinline__ void set(double h, double l) { hi = h; lo = l; }
    __device__ __forceinline__ double low() const { return lo; }
    __device__ __forceinline__ double value() const { return hi + lo; }
};
template <bool X> __device__ __forceinline__ tb_acc<X> tb_warp_sum_acc(tb_acc<X> v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tb_acc<X> w;
        w.set(__shfl_xor_sync(0xffffffffu, v.hi, o), X ? __shfl_xor_sync(0xffffffffu, v.low(), o) : 0.0);
        v.add(w);
    }
    return v;
}
// Sum over the block; result valid in every thread of warp 0.  `sm` has >= 64 doubles.
template <bool X> __device__ __forceinline__ tb_acc<X> tb_block_sum_acc(tb_acc<X> v, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = tb_warp_sum_acc<X>(v);
    __syncthreads();
    if (lane == 0) {
        sm[warp] = v.hi;
        if (X) sm[32 + warp] = v.low();
    }
    __syncthreads();
    tb_acc<X> r;
    if (warp == 0) {
        if (lane < nw) r.set(sm[lane], X ? sm[32 + lane] : 0.0);
        r = tb_warp_sum_acc<X>(r);
    }
    return r;
}

// Sum over the block; result valid in every thread of warp 0.  `sm` has >= 32 doubles.
__device__ __forceinline__ double tb_block_sum(double v, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = tb_warp_sum(v);
    __syncthreads();   // protect sm against a previous use
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    v = (warp == 0 && lane < nw) ? sm[lane] : 0.0;
    if (warp == 0) v = tb_warp_sum(v);
    return v;
}
__device__ __forceinline__ double tb_block_max(double v, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = tb_warp_max(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    v = (warp == 0 && lane < nw) ? sm[lane] : -INFINITY;
    if (warp == 0) v = tb_warp_max(v);
    return v;
}

// Deterministic grid reduction: every block deposits its partial, the last block to arrive
// (ticket) sums all partials in a fixed order.  Returns true in ALL threads of the last block,
// with *total valid in thread 0 of that block.  The ticket resets itself for the next launch.
__device__ __forceinline__ bool tb_grid_sum(double block_value /* valid in thread 0 */, double *partials,
                                            unsigned *ticket, double *sm, double *total) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_value;
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += ((volatile double *)partials)[i];
    s = tb_block_sum(s, sm);
    if (threadIdx.x == 0) *total = s;
    return true;
}

// tb_grid_sum for accumulators: partials[] holds the high words, partials[4*TB_MAX_PARTIALS + ...] the low words.
template <bool X>
__device__ __forceinline__ bool tb_grid_sum_acc(tb_acc<X> block_value /* valid in thread 0 */, double *partials, unsigned *ticket,
                                                double *sm, tb_acc<X> *total) {
    __shared__ int s_last_acc;
    double *plo = partials + 4 * TB_MAX_PARTIALS;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_value.hi;
        if (X) plo[blockIdx.x] = block_value.low();
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);
        s_last_acc = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last_acc) return false;
    __threadfence();
    tb_acc<X> s;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        tb_acc<X> w;
        w.set(((volatile double *)partials)[i], X ? ((volatile double *)plo)[i] : 0.0);
        s.add(w);
    }
    s = tb_block_sum_acc<X>(s, sm);
    if (threadIdx.x == 0) *total = s;
    return true;
}

// raise a kernel's dynamic shared-memory limit once per context
static inline cudaError_t tb_ensure_smem(tb_ctx *ctx, const void *func, size_t smem) {
    for (auto &e : ctx->smem_set)
        if (e.first == func) {
            if (e.second >= smem) return cudaSuccess;
            cudaError_t r = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (r == cudaSuccess) e.second = smem;
            return r;
        }
    cudaError_t r = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (r == cudaSuccess) ctx->smem_set.emplace_back(func, smem);
    return r;
}

// internal cross-TU helpers
int32_t tb_pattern_release(tb_pattern *p);
int32_t tb_ctx_ensure_cgwork(tb_ctx *ctx, int64_t n);
int32_t tb_halo_exchange(tb_ctx *ctx, const tb_pattern *pat, double *x);
int32_t tb_peer_release(tb_ctx *ctx);
int32_t tb_halo_push(tb_ctx *ctx, const tb_pattern *pat, const double *p, const struct CGState *st, tb_hwait_args *wait_out);
int32_t tb_halo_push_args(tb_ctx *ctx, const tb_pattern *pat, tb_push_args *out, tb_hwait_args *wait_out);

// ---- device side of the window protocol -------------------------------------------------------------------
__device__ __forceinline__ unsigned long long tb_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long tb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= epoch; gives up after 20 s (a dead peer must not hang the GPU) and records it in *err
__device__ __forceinline__ void tb_wait_flag(const unsigned long long *flag, unsigned long long epoch, unsigned long long *err) {
    if (tb_ld_acquire_sys(flag) >= epoch) return;
    const unsigned long long t0 = tb_globaltimer();
    while (tb_ld_acquire_sys(flag) < epoch) {
        if (tb_globaltimer() - t0 > 20000000000ull) {
            if (err) *err = 1;
            return;
        }
    }
}
// one thread: publish this rank's partial sum to every rank's window
__device__ __forceinline__ void tb_ar_publish(const tb_ar_args &a, double v) {
    for (int q = 0; q < a.nranks; q++) a.wins[q]->val[a.slot][a.rank] = v;
    __threadfence_system();
    for (int q = 0; q < a.nranks; q++) *(volatile unsigned long long *)&a.wins[q]->flag[a.slot][a.rank] = a.epoch;
}
// the same two steps for (hi, lo) pairs
template <bool X> __device__ __forceinline__ void tb_ar_publish_acc(const tb_ar_args &a, const tb_acc<X> &v) {
    for (int q = 0; q < a.nranks; q++) {
        a.wins[q]->val[a.slot][a.rank] = v.hi;
        if (X) a.wins[q]->val_lo[a.slot][a.rank] = v.low();
    }
    __threadfence_system();
    for (int q = 0; q < a.nranks; q++) *(volatile unsigned long long *)&a.wins[q]->flag[a.slot][a.rank] = a.epoch;
}
template <bool X> __device__ __forceinline__ double tb_ar_collect_acc(const tb_ar_args &a) {
    tb_peer_window *w = a.wins[a.rank];
    tb_acc<X> s;
    const unsigned long long t0 = blockIdx.x == 0 ? tb_globaltimer() : 0ull;
    for (int q = 0; q < a.nranks; q++) {
        tb_wait_flag(&w->flag[a.slot][q], a.epoch, &w->err);
        tb_acc<X> t;
        t.set(*(volatile double *)&w->val[a.slot][q], X ? *(volatile double *)&w->val_lo[a.slot][q] : 0.0);
        s.add(t);
    }
    if (blockIdx.x == 0) {
        w->ar_wait_ns += tb_globaltimer() - t0;
        w->ar_waits += 1;
    }
    return s.value();
}
// one thread: wait for all ranks' partials of this epoch and add them in rank order
__device__ __forceinline__ double tb_ar_collect(const tb_ar_args &a) {
    tb_peer_window *w = a.wins[a.rank];
    double s = 0.0;
    const unsigned long long t0 = blockIdx.x == 0 ? tb_globaltimer() : 0ull;
    for (int q = 0; q < a.nranks; q++) {
        tb_wait_flag(&w->flag[a.slot][q], a.epoch, &w->err);
        s += *(volatile double *)&w->val[a.slot][q];
    }
    if (blockIdx.x == 0) {
        w->ar_wait_ns += tb_globaltimer() - t0;
        w->ar_waits += 1;
    }
    return s;
}
// kernel prologue of an SpMV that reads ghost entries pushed by the neighbours
__device__ __forceinline__ void tb_halo_wait(const tb_hwait_args &h) {
    if (h.n > 0) {
        const unsigned long long t0 = (blockIdx.x == 0 && threadIdx.x == 0 && h.stat) ? tb_globaltimer() : 0ull;
        if ((int)threadIdx.x < h.n) tb_wait_flag(h.hflag + threadIdx.x, h.epoch, h.err);
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0 && h.stat) {
            h.stat[0] += tb_globaltimer() - t0;
            h.stat[1] += 1;
        }
    }
}
void tb_bj_free(struct tb_bj *b);
int32_t tb_pc_gershgorin(tb_ctx *ctx, const tb_csr *A, double *lmax);
int32_t tb_pc_cheb_first(tb_ctx *ctx, const double *r, const double *dinv, double *d, double *z, double *res, double inv_theta,
                         int64_t n, const struct CGState *st);
int32_t tb_pc_cheb_step(tb_ctx *ctx, const double *w, const double *dinv, double *d, double *z, double *res, double c1, double c2,
                        int64_t n, const struct CGState *st);
int32_t tb_pc_bj_update(tb_ctx *ctx, const tb_csr *A);
int32_t tb_pc_bj_apply(tb_ctx *ctx, const double *r, double *z, const struct CGState *st);
int32_t tb_cg_run_impl(tb_ctx *ctx, const tb_csr *A, const double *b, const tb_csr *M, double *phi, const double *bS,
                       double *x, double atol, double rtol, int64_t itmax, int64_t *iters, double *rnorm,
                       int32_t *converged, int precond = 0);
int32_t tb_cg_deferred_begin(tb_ctx *ctx);
int32_t tb_cg_deferred_end(tb_ctx *ctx, int64_t *iters_total, int32_t *all_solved);
int tb_cg_persistent_grid(tb_ctx *ctx, const tb_pattern *pat);
int tb_cg_persistent_kind(tb_ctx *ctx, const tb_pattern *pat, int *grid_out);
int32_t tb_cg_run_persistent_tma(tb_ctx *ctx, int grid, const tb_csr *A, const double *b, const tb_csr *M, const double *phi,
                                 const double *bS, double *x, double atol, double rtol, int64_t itmax, int64_t *iters,
                                 double *rnorm, int32_t *converged, const double *dinv);
int32_t tb_cg_run_persistent(tb_ctx *ctx, int grid, const tb_csr *A, const double *b, const tb_csr *M, const double *phi,
                             const double *bS, double *x, double atol, double rtol, int64_t itmax, int64_t *iters,
                             double *rnorm, int32_t *converged, const double *dinv);
int32_t tb_spmv_raw(tb_ctx *ctx, const tb_csr *A, double *x, double *y);
int32_t tb_cell_step_raw(tb_ctx *ctx, int model, const double *params, int nparams, double *u, int64_t n, int64_t ld,
                         int phi_idx, const double *phi_src, double t, double dt, int substeps, double thr,
                         double *max_dphi);
int32_t tb_get_tables(tb_ctx *ctx, int celltype, int qorder, const struct tb_elem_tables **d_T, int *nq);
int32_t tb_exclusive_scan_i64(tb_ctx *ctx, const int64_t *in, int64_t *out, int64_t n);
int tb_grid_for(tb_ctx *ctx, int64_t work_items, int block, int blocks_per_sm);
