// Fused LieTrotterGodunov step over (BackwardEulerSolver, ForwardEulerCellSolver | AdaptiveForwardEulerSubstepper).
//   OS outer step, children in tuple order      src/solver/time/integrator/operatorsplitting-interface.jl:23-232
//   perform_backward_euler_step!                src/solver/time/euler.jl:71-101
//   perform_step!(::PointwiseODEFunction, ...)  src/solver/time/partitioned_solver.jl:14-21
// Stream-ordered sequence per step (no host round trip except the CG convergence poll):
//   [A = M - dt K             only when !(dt ~ dt_last), euler.jl:82,104-116]
//   r = p = M*phi (+ bS), x = 0, gamma = r.r     one SELL sweep, fused init of CG ("b = M u_{n-1}")
//   CG iterations                                 tb_cg.cu ("inner solve")
//   cell sweep reading phi from x                 tb_cell.cu ("reaction solve"; fuses the copy x -> u view)
#include "tb_internal.cuh"
#include <math.h>

#define TB_DEV(ctx) TB_CUDA(cudaSetDevice((ctx)->device))

extern "C" int32_t tb_monodomain_create(tb_ctx *ctx, const tb_csr *M, const tb_csr *K, int32_t model, const double *params,
                                        int32_t nparams, int32_t phi_idx, tb_monodomain **out) {
    TB_REQUIRE(ctx && M && K && params && out, "tb_monodomain_create: NULL argument");
    TB_REQUIRE(M->pat == K->pat, "tb_monodomain_create: M and K must share one pattern");
    TB_REQUIRE(tb_model_known(model), "tb_monodomain_create: unknown ionic model %d", model);
    TB_REQUIRE(nparams == tb_model_nparams(model), "tb_monodomain_create: wrong parameter count %d", nparams);
    TB_REQUIRE(phi_idx == tb_model_phi(model), "tb_monodomain_create: model %d keeps the transmembrane potential in state %d", model,
               tb_model_phi(model));
    TB_DEV(ctx);
    *out = nullptr;
    tb_monodomain *md = new (std::nothrow) tb_monodomain();
    if (!md) return tb_fail(TB_ERR_NOMEM, "tb_monodomain_create: host allocation failed");
    md->ctx = ctx;
    md->M = M;
    md->K = K;
    md->A = nullptr;
    md->model = model;
    md->nparams = nparams;
    for (int i = 0; i < 40; i++) md->params[i] = i < nparams ? params[i] : 0.0;
    md->phi_idx = phi_idx;
    // LinearSolve defaults: abstol = reltol = sqrt(eps(Float64)), maxiters = length(b)
    md->atol = md->rtol = sqrt(2.220446049250313e-16);
    md->itmax = M->pat->nrows;
    md->substeps = 1;
    md->threshold = 0.1;
    md->bS = nullptr;
    md->bS_col = 0;
    md->dt_last = 0.0;   // euler.jl:172: first step always builds A
    md->b = md->x = nullptr;
    md->timing = false;
    for (int i = 0; i < 3; i++) md->section_ms[i] = 0.0;
    int32_t st = tb_csr_create_like(M, &md->A);
    if (st == TB_OK) st = tb_vec_create(ctx, M->pat->ncols, 1, &md->x);
    if (st == TB_OK)
        for (int i = 0; i < 4; i++)
            if (cudaEventCreate(&md->ev[i]) != cudaSuccess) st = tb_fail(TB_ERR_CUDA, "tb_monodomain_create: event creation failed");
    if (st != TB_OK) {
        tb_csr_destroy(md->A);
        tb_vec_destroy(md->x);
        delete md;
        return st;
    }
    *out = md;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_destroy(tb_monodomain *md) {
    if (!md) return TB_OK;
    cudaSetDevice(md->ctx->device);
    cudaStreamSynchronize(md->ctx->stream);
    tb_csr_destroy(md->A);
    tb_vec_destroy(md->x);
    for (int i = 0; i < 4; i++) cudaEventDestroy(md->ev[i]);
    if (md->s_in) cudaStreamDestroy(md->s_in);
    if (md->s_out) cudaStreamDestroy(md->s_out);
    for (cudaEvent_t e : {md->e_phi, md->e_s, md->e_done, md->o_s})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : md->e_chunk)
        if (e) cudaEventDestroy(e);
    delete md;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_set_cg(tb_monodomain *md, double atol, double rtol, int64_t itmax) {
    TB_REQUIRE(md, "tb_monodomain_set_cg: handle is NULL");
    TB_REQUIRE(atol >= 0 && rtol >= 0 && itmax >= 0, "tb_monodomain_set_cg: negative tolerance or itmax");
    md->atol = atol;
    md->rtol = rtol;
    md->itmax = itmax;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_set_preconditioner(tb_monodomain *md, int32_t precond) {
    TB_REQUIRE(md, "tb_monodomain_set_preconditioner: handle is NULL");
    TB_REQUIRE(precond >= TB_PRECOND_NONE && precond <= TB_PRECOND_CHEBYSHEV, "tb_monodomain_set_preconditioner: unknown preconditioner %d", precond);
    md->precond = precond;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_set_cell_solver(tb_monodomain *md, int32_t substeps, double reaction_threshold) {
    TB_REQUIRE(md, "tb_monodomain_set_cell_solver: handle is NULL");
    TB_REQUIRE(substeps >= 1, "tb_monodomain_set_cell_solver: substeps must be >= 1");
    md->substeps = substeps;
    md->threshold = reaction_threshold;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_set_source(tb_monodomain *md, const tb_vec *bS, int32_t col) {
    TB_REQUIRE(md, "tb_monodomain_set_source: handle is NULL");
    if (bS) TB_REQUIRE(col >= 0 && col < bS->ncols && bS->n >= md->M->pat->nrows, "tb_monodomain_set_source: vector too small");
    md->bS = bS;
    md->bS_col = col;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_enable_timing(tb_monodomain *md, int32_t on) {
    TB_REQUIRE(md, "tb_monodomain_enable_timing: handle is NULL");
    md->timing = on != 0;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_section_ms(tb_monodomain *md, double *ms3) {
    TB_REQUIRE(md && ms3, "tb_monodomain_section_ms: NULL argument");
    for (int i = 0; i < 3; i++) ms3[i] = md->section_ms[i];
    return TB_OK;
}

// Julia's isapprox(a, b) default: |a-b| <= sqrt(eps)*max(|a|,|b|)   (euler.jl:82 `dt ≈ dt_last`)
static bool approx_equal(double a, double b) {
    const double rt = sqrt(2.220446049250313e-16);
    return a == b || fabs(a - b) <= rt * fmax(fabs(a), fabs(b));
}

static int32_t monodomain_step_impl(tb_monodomain *md, tb_vec *u, double t, double dt, int64_t *iters, double *rnorm,
                                    int32_t *converged, double *max_dphi, cudaEvent_t before_cells = nullptr);

extern "C" int32_t tb_monodomain_step(tb_monodomain *md, tb_vec *u, double t, double dt, int64_t *iters, double *rnorm,
                                      int32_t *converged) {
    return monodomain_step_impl(md, u, t, dt, iters, rnorm, converged, nullptr);
}

// Same step, additionally returning the reaction tangent R = max_i du[i, phi] of the cell sweep's last rhs
// evaluation (over ALL ranks), which is what ReactionTangentController reads from cache.dumat (rtc.jl:51-78).
extern "C" int32_t tb_monodomain_step_rt(tb_monodomain *md, tb_vec *u, double t, double dt, int64_t *iters, double *rnorm,
                                         int32_t *converged, double *reaction_tangent) {
    TB_REQUIRE(reaction_tangent, "tb_monodomain_step_rt: reaction_tangent is NULL");
    TB_TRY(monodomain_step_impl(md, u, t, dt, iters, rnorm, converged, reaction_tangent));
    TB_TRY(tb_comm_allreduce_max(md->ctx, reaction_tangent));
    // _get_reaction_tangent starts from R = 0.0 and takes max(R, maximum(dumat[:, phi])) (rtc.jl:57-66): never negative
    if (!(*reaction_tangent > 0.0)) *reaction_tangent = 0.0;
    return TB_OK;
}

static int32_t monodomain_step_impl(tb_monodomain *md, tb_vec *u, double t, double dt, int64_t *iters, double *rnorm,
                                    int32_t *converged, double *max_dphi, cudaEvent_t before_cells) {
    TB_REQUIRE(md && u, "tb_monodomain_step: NULL argument");
    tb_ctx *ctx = md->ctx;
    const tb_pattern *pat = md->M->pat;
    const int ns = tb_model_nstates(md->model);
    TB_REQUIRE(u->ncols == ns, "tb_monodomain_step: state vector has %d columns, model needs %d", u->ncols, ns);
    TB_REQUIRE(u->n >= pat->ncols, "tb_monodomain_step: state vector has %lld rows, operator needs %lld", (long long)u->n,
               (long long)pat->ncols);
    TB_DEV(ctx);
    if (!approx_equal(dt, md->dt_last)) {
        TB_TRY(tb_csr_axpby_values(md->A, md->M, md->K, dt));
        md->dt_last = dt;
    }
    double *phi = u->d + (size_t)md->phi_idx * u->ld;
    const double *bS = md->bS ? md->bS->d + (size_t)md->bS_col * md->bS->ld : nullptr;
    if (md->timing) TB_CUDA(cudaEventRecord(md->ev[0], ctx->stream));
    // the fused init kernel is "b = M u_{n-1}"; its time is reported together with the solve and split
    // out by the bench through ncu launch lists
    int64_t it = 0;
    double rn = 0.0;
    int32_t conv = 0;
    TB_TRY(tb_cg_run_impl(ctx, md->A, nullptr, md->M, phi, bS, md->x->d, md->atol, md->rtol, md->itmax, &it, &rn, &conv, md->precond));
    if (md->timing) TB_CUDA(cudaEventRecord(md->ev[1], ctx->stream));
    if (before_cells) TB_CUDA(cudaStreamWaitEvent(ctx->stream, before_cells, 0));   // run_host: the non-phi columns are still in flight
    // reaction step on the owned points, phi taken from the CG solution (the reference copies x into the
    // u view even when the solve failed; the caller then rolls back, type.jl:510-532)
    TB_TRY(tb_cell_step_raw(ctx, md->model, md->params, md->nparams, u->d, pat->nrows, u->ld, md->phi_idx, md->x->d, t, dt,
                            md->substeps, md->threshold, max_dphi));
    if (md->timing) {
        TB_CUDA(cudaEventRecord(md->ev[2], ctx->stream));
        TB_CUDA(cudaEventSynchronize(md->ev[2]));
        float a = 0.f, b = 0.f;
        TB_CUDA(cudaEventElapsedTime(&a, md->ev[0], md->ev[1]));
        TB_CUDA(cudaEventElapsedTime(&b, md->ev[1], md->ev[2]));
        md->section_ms[0] = 0.0;
        md->section_ms[1] = a;
        md->section_ms[2] = b;
    }
    if (iters) *iters = it;
    if (rnorm) *rnorm = rn;
    if (converged) *converged = conv;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_run(tb_monodomain *md, tb_vec *u, double t0, double dt, int64_t nsteps,
                                     int64_t *iters_total, int32_t *all_converged) {
    TB_REQUIRE(md && u && nsteps >= 0, "tb_monodomain_run: bad argument");
    int64_t total = 0;
    int32_t all = 1;
    // Small and mid-size operators solve in ONE persistent kernel that needs nothing from the host: then the whole run is
    // enqueued without a single read-back -- iteration counts and convergence flags are folded into device-side totals
    // (k_pcg_fold) and fetched once at the end.  At C1 size a third of a step was the blocking scalar read-back and the
    // launch gaps behind it.  (The multi-kernel path polls the host by design and keeps the per-step loop below.)
    {
        tb_ctx *ctx = md->ctx;
        int pgrid = 0;
        const bool gen_pc = md->precond == TB_PRECOND_BLOCK_JACOBI || md->precond == TB_PRECOND_CHEBYSHEV;
        if (nsteps > 1 && !gen_pc && !ctx->exact_dot && !md->timing && !ctx->profile &&
            tb_cg_persistent_kind(ctx, md->M->pat, &pgrid) != 0) {
            TB_DEV(ctx);
            TB_TRY(tb_cg_deferred_begin(ctx));
            int32_t st = TB_OK;
            for (int64_t s = 0; s < nsteps && st == TB_OK; s++) {
                st = tb_monodomain_step(md, u, t0, dt, nullptr, nullptr, nullptr);
                t0 += dt;
            }
            const int32_t st2 = tb_cg_deferred_end(ctx, &total, &all);
            if (st != TB_OK) return st;
            if (st2 != TB_OK) return st2;
            if (iters_total) *iters_total = total;
            if (all_converged) *all_converged = all;
            return TB_OK;
        }
    }
    for (int64_t s = 0; s < nsteps; s++) {
        int64_t it = 0;
        int32_t conv = 0;
        // t_n = t0 + n*dt, evaluated like the integrator's `t += dt` bookkeeping is NOT: the reference
        // accumulates t (diffeq-interface.jl:374-406), so do the same
        TB_TRY(tb_monodomain_step(md, u, t0, dt, &it, nullptr, &conv));
        t0 += dt;
        total += it;
        all &= conv;
    }
    TB_CUDA(cudaStreamSynchronize(md->ctx->stream));
    if (iters_total) *iters_total = total;
    if (all_converged) *all_converged = all;
    return TB_OK;
}

extern "C" int32_t tb_monodomain_step_host(tb_monodomain *md, tb_vec *u_dev, const double *u_in, double *u_out, double t,
                                           double dt, int64_t *iters, double *rnorm, int32_t *converged) {
    TB_REQUIRE(md && u_dev && u_in && u_out, "tb_monodomain_step_host: NULL argument");
    tb_ctx *ctx = md->ctx;
    TB_DEV(ctx);
    // host -> device (asynchronous on the context stream when the host buffer is pinned)
    TB_CUDA(cudaMemcpy2DAsync(u_dev->d, sizeof(double) * u_dev->ld, u_in, sizeof(double) * u_dev->n,
                              sizeof(double) * u_dev->n, u_dev->ncols, cudaMemcpyHostToDevice, ctx->stream));
    TB_TRY(tb_monodomain_step(md, u_dev, t, dt, iters, rnorm, converged));
    TB_CUDA(cudaMemcpy2DAsync(u_out, sizeof(double) * u_dev->n, u_dev->d, sizeof(double) * u_dev->ld,
                              sizeof(double) * u_dev->n, u_dev->ncols, cudaMemcpyDeviceToHost, ctx->stream));
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    return TB_OK;
}

#define RH_CHUNKS_MAX 64
extern "C" int32_t tb_monodomain_set_host_chunks(tb_monodomain *md, int32_t nchunks) {
    TB_REQUIRE(md && nchunks >= 1 && nchunks <= RH_CHUNKS_MAX, "tb_monodomain_set_host_chunks: 1 <= nchunks <= 64");
    md->rh_chunks = nchunks;
    return TB_OK;
}
#define RH_TRACE_MAX 64
extern "C" int32_t tb_monodomain_run_host(tb_monodomain *md, tb_vec *u, double *buf0, double *buf1, double t0, double dt,
                                          int64_t nsteps, int64_t *iters_total, int32_t *all_converged) {
    TB_REQUIRE(md && u && buf0 && buf1 && nsteps >= 0, "tb_monodomain_run_host: bad argument");
    TB_REQUIRE(buf0 != buf1, "tb_monodomain_run_host: the two host buffers must be distinct");
    tb_ctx *ctx = md->ctx;
    TB_DEV(ctx);
    if (!md->s_in) {
        TB_CUDA(cudaStreamCreateWithFlags(&md->s_in, cudaStreamNonBlocking));
        TB_CUDA(cudaStreamCreateWithFlags(&md->s_out, cudaStreamNonBlocking));
        for (cudaEvent_t *e : {&md->e_phi, &md->e_s, &md->e_done, &md->o_s}) TB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (int i = 0; i < RH_CHUNKS_MAX; i++) TB_CUDA(cudaEventCreateWithFlags(&md->e_chunk[i], cudaEventDisableTiming));
        if (const char *e = getenv("TB_RUNHOST_CHUNKS")) md->rh_chunks = atoi(e);
    }
    const int64_t n = u->n, ld = u->ld;
    const int nc = u->ncols, phic = md->phi_idx;
    const int nchunks = md->rh_chunks < 1 ? 1 : md->rh_chunks > RH_CHUNKS_MAX ? RH_CHUNKS_MAX : md->rh_chunks;
    const int64_t csz = tb_round_up((n + nchunks - 1) / nchunks, 32);
    // TB_RUNHOST_TRACE=<file>: per-step timeline from CUDA events on the three streams (ms since the first step started):
    // compute start | CG + cells done | phi download done | phi upload (next step's input) done | other columns down | up
    const char *trace_path = getenv("TB_RUNHOST_TRACE");
    std::vector<cudaEvent_t> tev;
    const int64_t ntrace = trace_path ? (nsteps < RH_TRACE_MAX ? nsteps : RH_TRACE_MAX) : 0;
    if (ntrace > 0) {
        tev.resize((size_t)ntrace * 6);
        for (auto &e : tev) TB_CUDA(cudaEventCreate(&e));
    }
    auto trace = [&](int64_t step, int k, cudaStream_t st) {
        if (step < ntrace) cudaEventRecord(tev[(size_t)step * 6 + k], st);
    };
    double *buf[2] = {buf0, buf1};
    int64_t total = 0;
    int32_t all = 1;
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    // step 0: nothing to wait for
    TB_CUDA(cudaMemcpyAsync(u->d + (size_t)phic * ld, buf[0] + (size_t)phic * n, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, md->s_in));
    TB_CUDA(cudaEventRecord(md->e_phi, md->s_in));
    for (int c = 0; c < nc; c++)
        if (c != phic)
            TB_CUDA(cudaMemcpyAsync(u->d + (size_t)c * ld, buf[0] + (size_t)c * n, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, md->s_in));
    TB_CUDA(cudaEventRecord(md->e_s, md->s_in));
    for (int64_t s = 0; s < nsteps; s++) {
        double *out = buf[(s + 1) & 1];
        int64_t it = 0;
        int32_t conv = 0;
        TB_CUDA(cudaStreamWaitEvent(ctx->stream, md->e_phi, 0));
        trace(s, 0, ctx->stream);
        TB_TRY(monodomain_step_impl(md, u, t0, dt, &it, nullptr, &conv, nullptr, md->e_s));
        TB_CUDA(cudaEventRecord(md->e_done, ctx->stream));
        trace(s, 1, ctx->stream);
        t0 += dt;
        total += it;
        all &= conv;
        const bool more = s + 1 < nsteps;
        // download of this step's result; the upload of the next step's input chases it chunk by chunk (full duplex)
        TB_CUDA(cudaStreamWaitEvent(md->s_out, md->e_done, 0));
        for (int k = 0; k < nchunks; k++) {
            const int64_t o = k * csz, len = o + csz <= n ? csz : n - o;
            if (len <= 0) break;
            TB_CUDA(cudaMemcpyAsync(out + (size_t)phic * n + o, u->d + (size_t)phic * ld + o, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, md->s_out));
            TB_CUDA(cudaEventRecord(md->e_chunk[k], md->s_out));
            if (more) {
                TB_CUDA(cudaStreamWaitEvent(md->s_in, md->e_chunk[k], 0));
                TB_CUDA(cudaMemcpyAsync(u->d + (size_t)phic * ld + o, out + (size_t)phic * n + o, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, md->s_in));
            }
        }
        trace(s, 2, md->s_out);
        if (more) TB_CUDA(cudaEventRecord(md->e_phi, md->s_in));
        trace(s, 3, md->s_in);
        for (int c = 0; c < nc; c++)
            if (c != phic)
                TB_CUDA(cudaMemcpyAsync(out + (size_t)c * n, u->d + (size_t)c * ld, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, md->s_out));
        TB_CUDA(cudaEventRecord(md->o_s, md->s_out));
        trace(s, 4, md->s_out);
        if (more) {
            TB_CUDA(cudaStreamWaitEvent(md->s_in, md->o_s, 0));
            for (int c = 0; c < nc; c++)
                if (c != phic)
                    TB_CUDA(cudaMemcpyAsync(u->d + (size_t)c * ld, out + (size_t)c * n, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, md->s_in));
            TB_CUDA(cudaEventRecord(md->e_s, md->s_in));
        }
        trace(s, 5, md->s_in);
    }
    TB_CUDA(cudaStreamSynchronize(ctx->stream));
    TB_CUDA(cudaStreamSynchronize(md->s_out));
    TB_CUDA(cudaStreamSynchronize(md->s_in));
    if (ntrace > 0) {
        if (FILE *f = fopen(trace_path, ctx->rank == 0 ? "w" : "a")) {
            fprintf(f, "# rank %d: step, compute_start, compute_done, phi_d2h_done, phi_h2d_done, rest_d2h_done, rest_h2d_done [ms]\n", ctx->rank);
            for (int64_t k = 0; k < ntrace; k++) {
                fprintf(f, "%d,%lld", ctx->rank, (long long)k);
                for (int j = 0; j < 6; j++) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, tev[0], tev[(size_t)k * 6 + j]);
                    fprintf(f, ",%.3f", ms);
                }
                fprintf(f, "\n");
            }
            fclose(f);
        }
        for (auto &e : tev) cudaEventDestroy(e);
    }
    if (iters_total) *iters_total = total;
    if (all_converged) *all_converged = all;
    return TB_OK;
}
