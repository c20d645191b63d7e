// SELL-32 row kernels shared by tb_spmv (tb_csr.cu) and the CG iteration (tb_cg.cu).
// mul!(y, ::ThreadedSparseMatrixCSR, x), src/utils.jl:210-231: v = 0; for nz in row: v += A[nz]*x[col[nz]].
//
// Two implementations of the same arithmetic (identical summation order, bitwise equal results):
//   tb_sell_row        plain LDG: every lane loads its own values/column ids (any slice width)
//   tb_sell_sweep_tma  sm_100a bulk-async pipeline: one elected lane per warp streams the slice's
//                      contiguous value block and its (compressed) column stream into shared memory with
//                      cp.async.bulk (TMA, SASS UBLKCP) completing on an mbarrier; the lanes then read
//                      shared memory and only the x gathers go through the LSU, all of a row's gathers
//                      in flight at once.  Slices wider than TB_TMA_WCAP (a few high-valence rows of an
//                      unstructured mesh) are not staged: the same warp runs tb_sell_row on them.
//                      Measured on the 101 M-row hex operator (uncompressed columns): 5.3 ms vs 6.25 ms
//                      for the LDG kernel (1.02 vs 0.87 of the measured copy bandwidth).
#pragma once
#include "tb_internal.cuh"

#define TB_CCOL_EXPLICIT INT32_MIN   // header value of a slot whose 32 column ids are stored explicitly

// Gathered operand of the SpMV.  FP = false: x[c].  FP = true ("fused p update", experiment of VERDICT r1 item 10): the CG
// direction p = r + beta p_old is formed on the fly from r (= x) and p_old (= x2) -- the same unfused multiply and add
// k_cg_p performs, hence the same bits -- so that the separate p-update kernel and its 24 B/row disappear.
template <bool FP>
__device__ __forceinline__ double tb_gx(const double *__restrict__ x, const double *__restrict__ x2, double beta, int64_t c) {
    if (FP) return x[c] + beta * x2[c];
    return x[c];
}

// Row (s*32 + lane) of y = A x.  Entry j of the row is at slice_ptr[s] + j*32 + lane, so one warp
// streams 256 B of values + 128 B of column ids per j, fully coalesced; the dependent x gathers hit
// L1/L2 (neighbouring rows share columns).  The additions stay strictly left to right (and unfused,
// -fmad=false) so the sum is bitwise the reference's.
template <bool FP = false>
__device__ __forceinline__ double tb_sell_row(const int64_t *__restrict__ slice_ptr, const int *__restrict__ col,
                                              const double *__restrict__ val, const double *__restrict__ x, int64_t s,
                                              int lane, const double *__restrict__ x2 = nullptr, double beta = 0.0) {
    const int64_t base = slice_ptr[s];
    const int w = (int)((slice_ptr[s + 1] - base) >> 5);
    const int *c = col + base + lane;
    const double *v = val + base + lane;
    double acc = 0.0;
    int j = 0;
    for (; j + 4 <= w; j += 4) {
        const int c0 = c[(j + 0) * 32], c1 = c[(j + 1) * 32], c2 = c[(j + 2) * 32], c3 = c[(j + 3) * 32];
        const double v0 = v[(j + 0) * 32], v1 = v[(j + 1) * 32], v2 = v[(j + 2) * 32], v3 = v[(j + 3) * 32];
        const double g0 = tb_gx<FP>(x, x2, beta, c0), g1 = tb_gx<FP>(x, x2, beta, c1), g2 = tb_gx<FP>(x, x2, beta, c2), g3 = tb_gx<FP>(x, x2, beta, c3);
        acc += v0 * g0;
        acc += v1 * g1;
        acc += v2 * g2;
        acc += v3 * g3;
    }
    for (; j < w; j++) acc += v[j * 32] * tb_gx<FP>(x, x2, beta, c[j * 32]);
    return acc;
}

// ---------------------------------------------------------------------------------------------------
// bulk-async (TMA) pipeline
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned tb_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tb_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tb_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tb_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TB_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TB_WAIT_DONE;\n"
        "bra TB_WAIT_LOOP;\n"
        "TB_WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier (src/dst 16 B aligned, size % 16 == 0)
__device__ __forceinline__ void tb_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Launch geometry of the staged kernels: as many warps per CTA as fit the shared memory (the gather
// phase is latency bound: more warps = more gathers in flight; measured 21 x 1 stage > 16 x 1 > 8 x 2 >
// 4 x 3), one CTA per SM.
struct tb_tma_geom {
    int warps;
    unsigned val_bytes;    // bytes reserved per stage for the value block
    unsigned col_bytes;    // bytes reserved per stage for the column stream
    size_t smem;
};
// col_ints: largest per-slice column stream in ints (compressed), or 32*max_width (uncompressed)
static inline tb_tma_geom tb_tma_geometry(int max_width, int col_ints, int stages, int warps_override) {
    tb_tma_geom g;
    const int w = max_width < 1 ? 1 : max_width;
    g.val_bytes = (unsigned)w * 256u;
    g.col_bytes = (unsigned)((col_ints + 3) & ~3) * 4u;
    const size_t per_warp = (size_t)stages * (g.val_bytes + g.col_bytes) + (size_t)stages * 8;
    int n = (int)((220 * 1024) / per_warp);
    if (n > 32) n = 32;
    if (n < 1) n = 1;
    if (warps_override > 0 && warps_override < n) n = warps_override;
    g.warps = n;
    g.smem = (size_t)n * per_warp;
    return g;
}

// Shared-memory column stage of the compressed stream, in ints.  Default: the slot headers only (one int per slot, padded to
// 4) -- slices that also carry explicit 32-id blocks read their ids from the uncompressed SELL array, see tb_sell_sweep_tma.
// TB_SPMV_HDRONLY=0 sizes the stage for the largest stream of the operator instead (round-1 behaviour: 21 instead of 31 warps
// per SM on the 27-point operator).
// Only where the compression works (stream at most half of the uncompressed ids: structured grids) -- on an unstructured
// mesh nearly every slice carries explicit blocks and staging them through TMA is the faster way (C4: 79.9 vs 81.5 us).
static inline int tb_ccol_stage_ints(const tb_pattern *pat) {
    static int hdr_only = -1;
    if (hdr_only < 0) {
        const char *e = getenv("TB_SPMV_HDRONLY");
        hdr_only = e ? atoi(e) != 0 : 1;
    }
    const int hdr = ((pat->max_width_tma < 1 ? 1 : pat->max_width_tma) + 3) & ~3;
    const bool compresses = pat->ccol_len * 2 <= pat->sell_len;
    return hdr_only && compresses && hdr < pat->max_ccol_ints ? hdr : pat->max_ccol_ints;
}

// One warp = one private ring of STAGES slices.  `epi(row, acc)` is called by every lane with the
// finished row sum (rows >= nrows included: the caller masks).
// CC = false: cstream = SELL column ids, cptr unused.  CC = true: cstream/cptr = compressed column stream.
// Dynamic shared memory: nwarps*STAGES*(val_bytes+col_bytes) + nwarps*STAGES*8 bytes.
// Ring state of one warp, carried across several sweeps of ONE kernel (persistent CG): the mbarriers are initialised
// once (re-initialising a live mbarrier is undefined) and the phase parity simply keeps counting.
struct tb_tma_ring {
    int stage = 0;
    unsigned parity = 0;
    bool ready = false;
};

// Slices wider than TB_TMA_WCAP (ids ascending).  A lane-per-row sweep over such a slice is one warp walking hundreds of
// dependent gathers while the rest of the grid waits at the next barrier (LV, 483-wide apex slice: ~60 us of a 124 us
// iteration).  With the list, every (wide slice, row) pair becomes ONE WARP's job instead: the lanes fetch 32 entries of
// the row at once and the products are added in entry order through shuffles -- the same left-to-right sum, bit for bit.
struct tb_wide_list {
    const int64_t *slices = nullptr;
    int n = 0;
};

template <int STAGES, bool CC, class Epilogue, bool FP = false>
__device__ __forceinline__ void tb_sell_sweep_tma(const int64_t *__restrict__ slice_ptr, const double *__restrict__ val,
                                                  const int *__restrict__ cstream, const int64_t *__restrict__ cptr,
                                                  const double *__restrict__ x, int64_t nslices, unsigned val_bytes,
                                                  unsigned col_bytes, unsigned char *smem, Epilogue epi,
                                                  tb_tma_ring *ring = nullptr, const int *__restrict__ col = nullptr,
                                                  const tb_wide_list wide = tb_wide_list(), const double *__restrict__ x2 = nullptr,
                                                  double beta = 0.0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const size_t stage_bytes = (size_t)val_bytes + col_bytes;
    unsigned char *wbase = smem + (size_t)warp * STAGES * stage_bytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)nwarp * STAGES * stage_bytes) + warp * STAGES;
    if (!ring || !ring->ready) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; s++) tb_mbar_init(tb_smem_addr(bars + s), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (ring) ring->ready = true;
    }

    const int64_t gw = (int64_t)blockIdx.x * nwarp + warp, nw = (int64_t)gridDim.x * nwarp;

    // Where a slice lives: loaded by every lane (one broadcast transaction each) ONE ITERATION AHEAD of its use, so that neither
    // the bulk copies of the next slice nor the loop over the current one start behind an L2 round trip (ncu source view of
    // the round-1 form: 15 % of the stall samples on `slice_ptr[s + 1] - slice_ptr[s]`, and the empty stage waited as long
    // again before its copy was even issued).
    struct Meta {
        int64_t base, cb;
        unsigned n, cn;      // stored entries (32 * width); ints of the compressed column stream
    };
    auto load_meta = [&](int64_t s) {
        Meta m{0, 0, 0u, 0u};
        if (s < nslices) {
            m.base = slice_ptr[s];
            m.n = (unsigned)(slice_ptr[s + 1] - m.base);
            if (CC) {
                m.cb = cptr[s];
                m.cn = (unsigned)(cptr[s + 1] - m.cb);
            }
        }
        return m;
    };
    auto issue = [&](int stage, const Meta &m) {   // lane 0 only
        const int64_t base = m.base;
        const unsigned n = m.n;
        const unsigned bar = tb_smem_addr(bars + stage);
        unsigned char *dst = wbase + (size_t)stage * stage_bytes;
        if (n > TB_TMA_WCAP * 32u) {            // wide slice: nothing is staged, the phase completes at once
            tb_mbar_expect_tx(bar, 0u);
            return;
        }
        if (CC) {
            const int64_t cb = m.cb;
            unsigned cn = m.cn;                                    // ints, multiple of 4
            if (cn * 4u > col_bytes) cn = 0;                       // stream larger than the stage: its ids are read from `col` (below)
            tb_mbar_expect_tx(bar, n * 8u + cn * 4u);
            if (n) {
                tb_bulk_g2s(tb_smem_addr(dst), val + base, n * 8u, bar);
                if (cn) tb_bulk_g2s(tb_smem_addr(dst + val_bytes), cstream + cb, cn * 4u, bar);
            }
        } else {
            tb_mbar_expect_tx(bar, n * 12u);
            if (n) {
                tb_bulk_g2s(tb_smem_addr(dst), val + base, n * 8u, bar);
                tb_bulk_g2s(tb_smem_addr(dst + val_bytes), cstream + base, n * 4u, bar);
            }
        }
    };

    // a sweep always leaves every stage consumed, so the next one (same kernel, `ring`) starts filling at ring->stage
    int stage = ring ? ring->stage : 0;
    unsigned parity = ring ? ring->parity : 0;
    int64_t s_issue = gw;
    Meta held[STAGES];                     // metadata of the slices in flight, by ring position (consumption order)
    {
        int st = stage;
#pragma unroll
        for (int k = 0; k < STAGES; k++) {
            held[k] = load_meta(s_issue);
            if (lane == 0 && s_issue < nslices) issue(st, held[k]);
            s_issue += nw;
            if (++st == STAGES) st = 0;
        }
    }
    s_issue = gw + (int64_t)STAGES * nw;   // same value in every lane
    Meta ahead = load_meta(s_issue);       // the slice that will be issued when the first stage is free

    // wide rows first (the first staged slices are already in flight): one warp per row
    for (int64_t q = gw; q < (int64_t)wide.n * TB_SLICE; q += nw) {
        const int64_t s = wide.slices[q >> 5];
        const int rl = (int)(q & 31);
        const int64_t base = slice_ptr[s];
        const int w = (int)((slice_ptr[s + 1] - base) >> 5);
        const int *c = (CC ? col : cstream) + base + rl;
        const double *v = val + base + rl;
        double acc = 0.0;
        // eight chunks of 32 entries are fetched before the first of them is added: the (column id -> x) load chains of a
        // chunk are two dependent L2 round trips, and a 483-entry apex row walked one chunk at a time kept its warp -- and with
        // it the end of the sweep -- busy for ~27 us.  The additions stay in entry order.
        constexpr int WU = 8;
        for (int j0 = 0; j0 < w; j0 += 32 * WU) {
            double prod[WU];
#pragma unroll
            for (int u = 0; u < WU; u++) {
                const int j = j0 + u * 32 + lane;
                prod[u] = j < w ? v[(int64_t)j * 32] * tb_gx<FP>(x, x2, beta, c[(int64_t)j * 32]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < WU; u++) {
                const int left = w - (j0 + u * 32);
                const int m = left < 32 ? left : 32;
                for (int k = 0; k < m; k++) acc += __shfl_sync(0xffffffffu, prod[u], k);
            }
        }
        if (lane == 0) epi(s * TB_SLICE + rl, acc);
    }

    for (int64_t s = gw; s < nslices; s += nw) {
        const Meta cur = held[0];
        const int w = (int)(cur.n >> 5);
        const int row = (int)(s * TB_SLICE) + lane;
        tb_mbar_wait(tb_smem_addr(bars + stage), parity);
        const double *sv = reinterpret_cast<const double *>(wbase + (size_t)stage * stage_bytes) + lane;
        const int *sc = reinterpret_cast<const int *>(wbase + (size_t)stage * stage_bytes + val_bytes);
        double acc = 0.0;
        const bool wide_done = w > TB_TMA_WCAP && wide.n > 0;   // handled row by row above
        if (w > TB_TMA_WCAP) {
            if (!wide_done) acc = tb_sell_row<FP>(slice_ptr, CC ? col : cstream, val, x, s, lane, x2, beta);   // not staged (see `issue`)
        } else if (CC) {
            // header: one int per slot (uniform offset or EXPLICIT); explicit blocks follow the padded header
            const int hdr_ints = (w + 3) & ~3;
            const int *sexp = sc + hdr_ints + lane;
            int e = 0, j = 0;
            const int cn = (int)cur.cn;
            if ((unsigned)cn * 4u > col_bytes) {
                // The column stage holds headers only (tb_ccol_stage_ints): a slice with explicit blocks -- on a structured grid
                // the one slice in sixteen that straddles a grid line -- takes its ids from the SELL array instead, 128
                // coalesced bytes per slot; its values were staged like everybody's.  Sizing the stage for these slices costs
                // a third of the warps an SM can hold.
                const int *gc = col + cur.base + lane;
                for (; j + 9 <= w; j += 9) {
                    double xv[9], vv[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) xv[k] = tb_gx<FP>(x, x2, beta, gc[(int64_t)(j + k) * 32]);
#pragma unroll
                    for (int k = 0; k < 9; k++) vv[k] = sv[(j + k) * 32];
#pragma unroll
                    for (int k = 0; k < 9; k++) acc += vv[k] * xv[k];
                }
                for (; j < w; j++) acc += sv[j * 32] * tb_gx<FP>(x, x2, beta, gc[(int64_t)j * 32]);
            } else if (cn == hdr_ints) {
                // fast path (most slices of a structured mesh): every slot is a uniform offset
                for (; j + 9 <= w; j += 9) {
                    double xv[9], vv[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) xv[k] = tb_gx<FP>(x, x2, beta, row + sc[j + k]);
#pragma unroll
                    for (int k = 0; k < 9; k++) vv[k] = sv[(j + k) * 32];
#pragma unroll
                    for (int k = 0; k < 9; k++) acc += vv[k] * xv[k];
                }
                for (; j < w; j++) acc += sv[j * 32] * tb_gx<FP>(x, x2, beta, row + sc[j]);
            }
            for (; j + 9 <= w; j += 9) {
                double xv[9], vv[9];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const int h = sc[j + k];                                  // broadcast read
                    int c = row + h;
                    if (h == TB_CCOL_EXPLICIT) c = sexp[(e++) * 32];          // warp-uniform branch
                    xv[k] = tb_gx<FP>(x, x2, beta, c);
                }
#pragma unroll
                for (int k = 0; k < 9; k++) vv[k] = sv[(j + k) * 32];
#pragma unroll
                for (int k = 0; k < 9; k++) acc += vv[k] * xv[k];
            }
            for (; j < w; j++) {
                const int h = sc[j];
                int c = row + h;
                if (h == TB_CCOL_EXPLICIT) c = sexp[(e++) * 32];
                acc += sv[j * 32] * tb_gx<FP>(x, x2, beta, c);
            }
        } else {
            const int *scl = sc + lane;
            int j = 0;
            for (; j + 9 <= w; j += 9) {
                double xv[9], vv[9];
#pragma unroll
                for (int k = 0; k < 9; k++) xv[k] = tb_gx<FP>(x, x2, beta, scl[(j + k) * 32]);
#pragma unroll
                for (int k = 0; k < 9; k++) vv[k] = sv[(j + k) * 32];
#pragma unroll
                for (int k = 0; k < 9; k++) acc += vv[k] * xv[k];
            }
            for (; j < w; j++) acc += sv[j * 32] * tb_gx<FP>(x, x2, beta, scl[j * 32]);
        }
        if (!wide_done) epi(s * TB_SLICE + lane, acc);
        // every lane has consumed its shared-memory operands (acc depends on all of them): the stage may be refilled
        __syncwarp();
        if (lane == 0 && s_issue < nslices) issue(stage, ahead);
#pragma unroll
        for (int k = 0; k + 1 < STAGES; k++) held[k] = held[k + 1];
        held[STAGES - 1] = ahead;
        s_issue += nw;
        ahead = load_meta(s_issue);        // in flight while the next slice is waited for and consumed
        if (++stage == STAGES) {
            stage = 0;
            parity ^= 1u;
        }
    }
    if (ring) {
        ring->stage = stage;
        ring->parity = parity;
    }
}
