/*
 * tbolt_b200.h -- C ABI of libtbolt_b200.so: the B200-native (sm_100a) monodomain hot path that
 * replaces Thunderbolt.jl's ext/CuThunderboltExt.jl behind the solver/operator dispatch points.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / opaque handles, returns an
 * int32 status (TB_OK = 0) and never throws.  The message of the last failure is available from
 * tb_last_error().  Host pointers are borrowed for the duration of the call only; device memory is
 * owned by the library behind the handles and released by the matching *_destroy.  A tb_ctx is
 * used from one host thread at a time; all work is ordered on the context's CUDA stream (a non-blocking
 * stream unless the caller supplies one); only *_download, the solves and steps that return scalars
 * (tb_cg_solve*, tb_monodomain_step*, tb_monodomain_run*, tb_ecg_plonsey), assemblies that ship per-cell
 * or per-point host data (coefficient fields, tb_assemble_source_qp) and tb_sync block the host -- tb_assemble_mass,
 * tb_assemble_diffusion with a constant coefficient and tb_assemble_source[_program] are fully stream-ordered (an
 * asynchronous kernel failure then surfaces at the next blocking call on the context).
 *
 * Each declaration cites the reference interface (file:line under JuliaHealth/Thunderbolt.jl
 * v0.0.4) that a Julia `ccall` of it replaces; INTEGRATION.md shows those bindings.
 *
 * Layout contracts shared with the reference:
 *   - state vectors are state-blocked (SoA): state s of point i is u[s*N + i] on the host side
 *     (src/modeling/solution_variables.jl:60-63); on the device each state column is padded to a
 *     multiple of 32 doubles so 128-bit loads stay aligned -- invisible through this API.
 *   - matrices are CSR with the reference's pattern (Ferrite allocate_matrix, sorted columns,
 *     diagonal included; src/solver/interface.jl:159-168).  Values are exchanged in CSR nonzero
 *     order (`nonzeros(A)`); on the device they live in a sliced-ELL (SELL-32) image of that
 *     pattern, again invisible through this API.
 *   - K is assembled NEGATIVE semi-definite exactly like the reference (diffusion.jl:28-50), so the
 *     backward-Euler operator is A = M - dt*K (euler.jl:104-116).
 */
#ifndef TBOLT_B200_H
#define TBOLT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_OK 0
#define TB_ERR_INVALID 1     /* bad argument (null handle, size mismatch, unsupported enum) */
#define TB_ERR_CUDA 2        /* a CUDA runtime call or kernel failed */
#define TB_ERR_NOMEM 3       /* device or host allocation failed */
#define TB_ERR_COMM 4        /* NCCL failure, or a wait on a peer rank's data timed out (peer-memory path) */
#define TB_ERR_UNSUPPORTED 5 /* valid request the library does not implement */

/* cell types (Ferrite Quadrilateral / Hexahedron / Triangle / Tetrahedron, Lagrange order 1) */
#define TB_QUAD4 0
#define TB_HEX8 1
#define TB_TRI3 2
#define TB_TET4 3

/* ionic models: src/modeling/cells/fhn.jl:6-13 (6 parameters a,b,c,d,e,f) and
 * src/modeling/cells/pcg2019.jl:4-48 (36 parameters in declaration order g_Na .. E_Ca) */
#define TB_FHN 0
#define TB_PCG2019 1
#define TB_ALIEV_PANFILOV 2   /* src/modeling/cells/aliev-panfilov.jl:1-34: 6 parameters c_t,k,a,eps0,mu1,mu2; states (s, phi_m): phi_idx = 1 */

/* diffusion coefficient kinds, D = kappa/(Cm*chi)  (src/modeling/core/coefficients.jl:152-162)
 *   SCALAR   data[0]                                   ConstantCoefficient(number)
 *   TENSOR   data[dim*dim] row-major symmetric tensor  ConstantCoefficient(SymmetricTensor) :106-120
 *   SPECTRAL data = lambda[3], then for every cell and local node a: f[3], s[3], n[3]
 *            SpectralTensorCoefficient over an OrthotropicMicrostructureModel of FieldCoefficients
 *            (coefficients.jl:36-99,451-488; microstructure.jl:136-187) */
#define TB_D_SCALAR 0
#define TB_D_TENSOR 1
#define TB_D_SPECTRAL 2
#define TB_D_CELL_TENSOR 3 /* data = one dim x dim tensor per cell (piecewise-constant coefficient fields) */

/* built-in stimulus families f(x,t); a Julia closure cannot cross a C ABI, so either one of these
 * or host-evaluated values per quadrature point (tb_assemble_source_qp).  prm[]:
 *   BOX    max_d x_d < p0 && t < p1 ? p2 : 0   bak/examples/conduction-velocity-benchmark.jl:47-50
 *   BALL   |x| < p0 && t < p1 ? p2 : 0         test/integration/test_electrophysiology.jl:83
 *   COSEXP cos(2 pi t) exp(-|x|^2)             benchmarks/benchmarks-cuda-linear-form.jl:4-18
 *   NORMT  |x| + t                             benchmarks/benchmarks-linear-form.jl:16-21
 *   ENDO   t <= p1 && x0 < p0 ? p2/p3*exp(t/p3) : 0   docs/src/literate-tutorials/ep04_geselowitz-ecg.jl:15-26 */
#define TB_SRC_NONE 0
#define TB_SRC_BOX 1
#define TB_SRC_BALL 2
#define TB_SRC_COSEXP 3
#define TB_SRC_NORMT 4
#define TB_SRC_ENDO 5
#define TB_SRC_PROGRAM 6 /* a traced closure, see tb_assemble_source_program */

typedef struct tb_ctx tb_ctx;
typedef struct tb_vec tb_vec;
typedef struct tb_mesh tb_mesh;
typedef struct tb_csr tb_csr;
typedef struct tb_monodomain tb_monodomain;

/* ---- context --------------------------------------------------------------------------------
 * Replaces the device selection of src/devices.jl:1-4 / ext/CuThunderboltExt.jl (CudaDevice).
 * `stream` may be NULL (library creates its own non-blocking stream) or an existing cudaStream_t. */
int32_t tb_version(void);
const char *tb_last_error(void);
int32_t tb_ctx_create(int32_t device, void *stream, tb_ctx **out);
int32_t tb_ctx_destroy(tb_ctx *ctx);
int32_t tb_sync(tb_ctx *ctx);
int32_t tb_device_info(tb_ctx *ctx, int32_t *sm_count, int64_t *total_mem_bytes, int32_t *cc_major, int32_t *cc_minor);
/* CUDA-event timer on the context's stream (the stream every kernel of this library launches on) */
int32_t tb_timer_start(tb_ctx *ctx);
int32_t tb_timer_stop(tb_ctx *ctx, double *elapsed_ms);
/* number of kernels this library has launched on ctx so far (bench.py's gpu_launches) */
int32_t tb_launch_count(tb_ctx *ctx, int64_t *count);
/* per-kernel timing of the dominant kernel (the SpMV inside CG): CUDA events around every launch that
 * did work, accumulated since the last tb_profile_enable(ctx, 1).  bench.py's roofline.achieved. */
int32_t tb_profile_enable(tb_ctx *ctx, int32_t on);
int32_t tb_profile_get(tb_ctx *ctx, double *spmv_ms_total, int64_t *spmv_launches);
/* overwrite a buffer larger than L2 so that the next timed kernel starts cold */
int32_t tb_l2_flush(tb_ctx *ctx);
/* multi-GPU: join a communicator of `nranks` one-process-per-GPU peers.  `nccl_unique_id` is the
 * 128-byte ncclUniqueId created by tb_comm_unique_id on rank 0 and shipped by the host (e.g.
 * torch.distributed).  The reference has no counterpart (shared-memory only). */
int32_t tb_comm_unique_id(void *out128);
int32_t tb_ctx_comm_init(tb_ctx *ctx, int32_t rank, int32_t nranks, const void *nccl_unique_id);
int32_t tb_comm_barrier(tb_ctx *ctx);
int32_t tb_comm_allreduce_max(tb_ctx *ctx, double *value_inout);

/* ---- vectors --------------------------------------------------------------------------------
 * create_system_vector / adapt_vector_type (src/solver/interface.jl:175-181, src/utils.jl:425-427;
 * ext/CuThunderboltExt.jl:126-127,144-146).  A tb_vec has n rows and ncols state columns. */
int32_t tb_vec_create(tb_ctx *ctx, int64_t n, int32_t ncols, tb_vec **out);
int32_t tb_vec_destroy(tb_vec *v);
int32_t tb_vec_sizes(const tb_vec *v, int64_t *n, int32_t *ncols);
/* host layout: column c occupies host[c*n .. c*n+n) (the reference's flat SoA vector) */
int32_t tb_vec_upload(tb_vec *v, const double *host);
int32_t tb_vec_download(const tb_vec *v, double *host);
int32_t tb_vec_upload_col(tb_vec *v, int32_t col, const double *host, int64_t offset, int64_t count);
int32_t tb_vec_download_col(const tb_vec *v, int32_t col, double *host, int64_t offset, int64_t count);
int32_t tb_vec_fill(tb_vec *v, int32_t col, double value);
int32_t tb_vec_copy(tb_vec *dst, int32_t dcol, const tb_vec *src, int32_t scol);
/* y[:,ycol] += a * x[:,xcol]   (add!(b, S) of euler.jl:88-91 is the a = 1 case) */
int32_t tb_vec_axpy(tb_vec *y, int32_t ycol, double a, const tb_vec *x, int32_t xcol);
/* raw device pointer of a column and the padded column stride (for interop with torch/CuArray) */
int32_t tb_vec_devptr(const tb_vec *v, int32_t col, void **ptr, int64_t *ld);

/* ---- mesh -----------------------------------------------------------------------------------
 * What Ferrite hands the operators: cells (node ids), node coordinates, celldofs of the closed
 * DofHandler (src/discretization/fem.jl:180-182).  index_base is 1 for Julia callers. */
int32_t tb_mesh_create(tb_ctx *ctx, int32_t celltype, int64_t ncells, int64_t nnodes, const int64_t *conn,
                       const double *coords, const int64_t *celldofs, int64_t ndofs, int32_t index_base,
                       tb_mesh **out);
/* device-side generate_grid + DofHandler close! for synthetic benchmarks (src/mesh/generators.jl:942):
 * Ferrite's node/cell order and first-touch DoF numbering, built in HBM without a host mesh. */
int32_t tb_mesh_generate_grid(tb_ctx *ctx, int32_t celltype, const int64_t *nel3, const double *left3,
                              const double *right3, tb_mesh **out);
int32_t tb_mesh_destroy(tb_mesh *m);
int32_t tb_mesh_sizes(const tb_mesh *m, int64_t *ncells, int64_t *nnodes, int64_t *ndofs, int32_t *nv, int32_t *dim);
/* any of the three may be NULL; ids come back 0-based */
int32_t tb_mesh_download(const tb_mesh *m, int64_t *conn, double *coords, int64_t *celldofs);
/* coordinates at the dof locations, ndofs x dim (evaluate_coefficient_at_dof_locations of a
 * CartesianCoordinateSystem, src/modeling/core/coefficients.jl:199-245) */
int32_t tb_mesh_dof_coords(const tb_mesh *m, double *host);
/* multi-GPU: keep the cells that touch dofs [dof_lo, dof_hi) of `global` and renumber: owned dofs
 * first (global id - dof_lo), ghosts after them sorted by global id.  ghost_global (nullable) gets
 * the ghosts' global ids; *nghost their count. */
int32_t tb_mesh_extract_local(const tb_mesh *global, int64_t dof_lo, int64_t dof_hi, tb_mesh **out, int64_t *nghost);
/* the same local mesh for a STRUCTURED Quadrilateral / Hexahedron grid, generated directly from Ferrite's closed-form first-touch
 * numbering: identical arrays to tb_mesh_extract_local(tb_mesh_generate_grid(...), lo, hi), but no rank ever holds the global
 * grid -- all temporaries are sized by the slab of cell layers that can touch [dof_lo, dof_hi) */
int32_t tb_mesh_generate_grid_local(tb_ctx *ctx, int32_t celltype, const int64_t *nel3, const double *left3, const double *right3,
                                    int64_t dof_lo, int64_t dof_hi, tb_mesh **out, int64_t *nghost);
int32_t tb_mesh_ghosts(const tb_mesh *local, int64_t *ghost_global);
/* the same ownership marking for a LOCAL mesh that was cut on the host and uploaded with tb_mesh_create (general
 * partitions: no global replica in HBM): dofs 0 .. ndofs_owned-1 are owned (global ids dof_lo ..), the rest are ghosts with
 * the given ascending global ids.  Call before the first assembly on this mesh. */
int32_t tb_mesh_set_ownership(tb_mesh *m, int64_t ndofs_owned, int64_t dof_lo, const int64_t *ghost_global, int64_t nghost);

/* ---- CSR operators ----------------------------------------------------------------------------
 * create_system_matrix (src/solver/interface.jl:159-173; ext/CuThunderboltExt.jl:129-139). */
int32_t tb_csr_create(tb_ctx *ctx, int64_t nrows, int64_t ncols, const int64_t *rowptr, const int64_t *colidx,
                      int32_t index_base, tb_csr **out);
/* device-side allocate_matrix(dh): all dof pairs sharing a cell, diagonal included, sorted.  Rows are
 * the mesh's owned dofs, columns all its dofs. */
int32_t tb_csr_create_from_mesh(tb_ctx *ctx, const tb_mesh *mesh, tb_csr **out);
/* a second operator on the same pattern (M, K and A share one pattern in the reference: euler.jl:134-146) */
int32_t tb_csr_create_like(const tb_csr *pattern_of, tb_csr **out);
int32_t tb_csr_destroy(tb_csr *a);
int32_t tb_csr_sizes(const tb_csr *a, int64_t *nrows, int64_t *ncols, int64_t *nnz);
/* HBM footprint of the operator's streams: stored (padded) entries, bytes of the column stream the SpMV
 * actually reads (lossless compression: one offset per slot where col = row + off for a whole slice),
 * widest slice.  Used by bench.py to report real bytes next to the algorithmic ones. */
int32_t tb_csr_storage(const tb_csr *a, int64_t *stored_entries, int64_t *column_stream_bytes, int32_t *max_width);
int32_t tb_csr_download_pattern(const tb_csr *a, int64_t *rowptr, int64_t *colidx, int32_t index_base);
int32_t tb_csr_values_download(const tb_csr *a, double *vals); /* nonzeros(A), CSR order */
int32_t tb_csr_values_upload(tb_csr *a, const double *vals);
int32_t tb_csr_zero(tb_csr *a);
/* nz(A) = nz(M) - dt*nz(K)   (_implicit_euler_heat_solver_update_system_matrix!, euler.jl:104-116) */
int32_t tb_csr_axpby_values(tb_csr *A, const tb_csr *M, const tb_csr *K, double dt);
/* y = A x   (mul!(y, ::ThreadedSparseMatrixCSR, x), src/utils.jl:210-231): sequential left-to-right
 * row sums, so the result is bitwise the reference's */
int32_t tb_spmv(tb_ctx *ctx, const tb_csr *A, const tb_vec *x, int32_t xcol, tb_vec *y, int32_t ycol);
/* multi-GPU halo plan of a row-partitioned operator: for each neighbour rank the local rows whose x
 * entries it needs, and where its entries land in our ghost block.  send_ptr/recv_ptr have nneigh+1
 * entries; recv offsets are relative to the first ghost. */
int32_t tb_csr_set_halo(tb_csr *A, int32_t nneigh, const int32_t *neigh_ranks, const int64_t *send_ptr,
                        const int64_t *send_rows, const int64_t *recv_ptr);

/* NVLink peer-memory path between the one-process-per-GPU ranks of ONE box (optional; NCCL is the fallback):
 * every rank exports a TB_PEER_BLOB_BYTES blob (CUDA IPC handles of its mailbox window and its CG work vectors,
 * sized for operators with up to `ncols` columns), the host ships all blobs to all ranks, tb_peer_attach maps
 * them.  From then on the CG kernels finish their dot products by storing partial sums into every rank's window
 * (summed in rank order: identical bits everywhere) and, once tb_csr_set_halo_peer told them where each
 * neighbour keeps our ghost entries (offset inside the neighbour's vector, and which of its halo flags is ours),
 * push the boundary of the direction vector straight into the neighbours' ghost blocks before each SpMV. */
#define TB_PEER_BLOB_BYTES 160
int32_t tb_peer_export(tb_ctx *ctx, int64_t ncols, void *blob_out);
int32_t tb_peer_attach(tb_ctx *ctx, const void *blobs, int32_t nranks);
int32_t tb_peer_enabled(tb_ctx *ctx, int32_t *on);
int32_t tb_csr_set_halo_peer(tb_csr *A, const int64_t *dst_off, const int32_t *dst_slot);
/* The fused communication path (collects and halo push inside the CG update kernels, 3 launches per iteration) consumes
 * a different number of halo epochs per solve than the unfused one, so ALL ranks must take the same path.
 * tb_csr_halo_fused_capable reports whether THIS rank could (every send list one run of consecutive rows, fusion not
 * disabled); the host reduces that with MIN over the ranks and hands the agreed value to tb_csr_set_halo_fused.
 * Until that call the unfused path is used. */
/* instrumentation of the peer path: time (ms) one CTA of this rank spent waiting for the other ranks' partial sums and
 * halo flags since the last reset, and the number of waits -- what the slowest rank costs per iteration */
int32_t tb_peer_stats(tb_ctx *ctx, double *ar_wait_ms, int64_t *ar_waits, double *halo_wait_ms, int64_t *halo_waits, int32_t reset);
int32_t tb_csr_halo_fused_capable(const tb_csr *A, int32_t *capable);
int32_t tb_csr_set_halo_fused(tb_csr *A, int32_t on);

/* ---- assembly ---------------------------------------------------------------------------------
 * update_operator!(op, t) of the FerriteOperators element loop (called at euler.jl:173-175) with the
 * element kernels of src/modeling/core/mass.jl:28-43, diffusion.jl:28-50,
 * analytical_coefficient.jl:80-101.  Values are zeroed, then every element matrix/vector is
 * scattered into the fixed pattern. */
int32_t tb_quadrature(int32_t celltype, int32_t qorder, int32_t *nq, double *pts, double *weights);
int32_t tb_assemble_mass(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, double rho, tb_csr *M);
int32_t tb_assemble_diffusion(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, int32_t kind, const double *data,
                              int64_t ndata, double cm_chi, tb_csr *K);
int32_t tb_assemble_source(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, int32_t kind, const double *prm,
                           int32_t nprm, double t, tb_vec *b, int32_t bcol);
/* Traced closure: the AnalyticalCoefficient's f(x, t) (analytical_coefficient.jl:80-101 evaluates it at every quadrature
 * point) is called ONCE on the host with tracing numbers and shipped as a postfix program, code[i] = op | arg << 8 with
 * op one of (in this order, from 0): X(arg = coordinate) T CONST(arg = index into consts) | ADD SUB MUL DIV MIN MAX POW
 * LT LE GT GE EQ NE AND OR (binary; comparisons give 1.0 / 0.0) | NEG ABS SQRT EXP LOG SIN COS TANH NOT (unary) |
 * SELECT (cond a b -> cond != 0 ? a : b).  At most 96 instructions, 24 constants, stack depth 16.  Evaluated on the
 * device in plain IEEE fp64 in program order: nothing but the 400-byte program crosses the host link, per call. */
int32_t tb_assemble_source_program(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, const int32_t *code, int32_t ncode,
                                   const double *consts, int32_t nconsts, double t, tb_vec *b, int32_t bcol);
/* general closure path (anything the tracer cannot express): fq[cell*nq + q] = f(x_q, t) evaluated by the host */
int32_t tb_assemble_source_qp(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, const double *fq, tb_vec *b,
                              int32_t bcol);
/* Assembly strategy (FerriteOperators' strategies: sequential / per-colour / per-element, SURVEY a-10):
 *   2 (default)  per-element results + ordered row gather: element matrices (vectors) go to a scratch buffer,
 *                then every matrix row sums its contributions in ascending element order -- deterministic and
 *                bitwise equal to the reference's sequential CPU assembly; rows are processed in chunks so the
 *                scratch stays under the budget below.  Falls back to 0 if a chunk's cell range cannot fit.
 *   0            fp64 atomic scatter (order of additions not deterministic, values equal to rounding).
 * tb_assembly_info reports the requested mode, the one the LAST assembly call really ran and its chunk count. */
int32_t tb_assembly_set_mode(tb_ctx *ctx, int32_t mode);
int32_t tb_assembly_info(tb_ctx *ctx, int32_t *mode_requested, int32_t *mode_last_used, int32_t *chunks_last);
int32_t tb_assembly_set_scratch_budget(tb_ctx *ctx, int64_t bytes);
/* the element scratch is cached on the context between assembly calls; this frees it (e.g. after setup) */
int32_t tb_assembly_release_scratch(tb_ctx *ctx);

/* ---- ECG post-processing ---------------------------------------------------------------------------
 * Plonsey1964ECGGaussCache (src/modeling/electrophysiology/ecg.jl:55-160): update_ecg!(cache, phi_m) followed by
 * evaluate_ecg(cache, x, kappa_t) for `ne` electrodes, fused into one element sweep:
 *   phi_e(x) = -1/(4 pi kappa_t) * sum_cells sum_q ((flux_q . (x_q - x)) / |x_q - x|^3) dOmega,
 *   flux_q = sum_i (D(x_q) . gradN_i) phi_i,  D = the coefficient of the diffusion operator (kind/data/cm_chi as in
 * tb_assemble_diffusion).  electrodes: ne x dim doubles; phi_e: ne doubles. */
int32_t tb_ecg_plonsey(tb_ctx *ctx, const tb_mesh *mesh, int32_t qorder, int32_t kind, const double *data, int64_t ndata,
                       double cm_chi, const tb_vec *phi, int32_t phicol, const double *electrodes, int32_t ne, double kappa_t,
                       double *phi_e);

/* ---- linear solve -----------------------------------------------------------------------------
 * LinearSolve.solve!(cache) with KrylovJL_CG (euler.jl:10,94,155-156): unpreconditioned CG from
 * x0 = 0, stop when |r| <= atol + rtol*|r0|, at most itmax iterations.  Non-convergence is NOT an
 * error status: it comes back as *converged = 0 so the caller can report ReturnCode.MaxIters and
 * euler.jl:95-100 returns false. */
int32_t tb_cg_solve(tb_ctx *ctx, const tb_csr *A, const tb_vec *b, int32_t bcol, tb_vec *x, int32_t xcol, double atol,
                    double rtol, int64_t itmax, int64_t *iters, double *rnorm, int32_t *converged);

/* Preconditioned variant: LinearSolve.KrylovJL_CG(precs = ..., ldiv = false) as in bak/examples-gpu/spiral-wave.jl:95-105
 * and the tip at ep01_spiral-wave.jl:129-131.  z = M r with M the stored inverse of the preconditioner, gamma = r.z,
 * stop on sqrt(r.z) <= atol + rtol*sqrt(r0.z0), p = z + beta p (Krylov.jl cg!).  TB_PRECOND_JACOBI: M = diag(A)^-1
 * (KrylovPreconditioners' BlockJacobi with blocks of one row), rebuilt from the operator at every solve. */
#define TB_PRECOND_NONE 0
#define TB_PRECOND_JACOBI 1
/* BlockJacobiPreconditioner(A, nblocks) as used by the reference's GPU example (bak/examples-gpu/spiral-wave.jl:95-105):
 * dense fp64 inverses of the diagonal blocks, z = blockdiag(A)^-1 r as a batched dense mat-vec.  Configure the row
 * partition once with tb_cg_set_block_jacobi; the inverses are rebuilt from the operator at every solve (update!(P, A)). */
#define TB_PRECOND_BLOCK_JACOBI 2
/* z = q_d(D^-1 A) D^-1 r: degree-d Chebyshev polynomial over [lmax/ratio, lmax], lmax = Gershgorin bound of D^-1 A.
 * Configure with tb_cg_set_chebyshev (default degree 8, ratio 30). */
#define TB_PRECOND_CHEBYSHEV 3
int32_t tb_cg_solve_pc(tb_ctx *ctx, const tb_csr *A, const tb_vec *b, int32_t bcol, tb_vec *x, int32_t xcol, int32_t precond,
                       double atol, double rtol, int64_t itmax, int64_t *iters, double *rnorm, int32_t *converged);

/* row_block: nrows block ids in [0, nblocks) (the reference takes them from Metis), or NULL for equal contiguous ranges of
 * the dof numbering.  Blocks may have different sizes (<= 2048 rows); memory = sum of squares of the block sizes. */
int32_t tb_cg_set_block_jacobi(tb_ctx *ctx, int64_t nrows, int64_t nblocks, const int32_t *row_block);
int32_t tb_cg_set_chebyshev(tb_ctx *ctx, int32_t degree, double ratio);

/* Single-GPU solves of small and mid-size operators run as ONE persistent cooperative kernel per solve (grid-wide
 * barriers instead of kernel boundaries; same recurrence, same stopping rule):
 *   path 1: <= 148*16*32*4 rows -- x, r, p, Ap of a lane's rows live in registers for the whole solve;
 *   path 2: <= 4 M rows (env TB_CG_PERSISTENT_MAX_ROWS) -- vectors stay in L2-resident global memory, the SpMV is the
 *           TMA-staged sweep of the large-operator path;
 *   path 0: everything else (and every multi-GPU solve): three kernels per iteration, scalars polled by the host.
 * mode 0 = never, 1 = auto (default), 2 = path 2 whenever it is eligible (lets the tests reach it on tiny systems).
 * tb_cg_last_path reports the path of the last solve. */
int32_t tb_cg_set_persistent(tb_ctx *ctx, int32_t mode);
/* register-resident persistent kernel (operators up to ~300 k rows): 2 (default) = two flag barriers per iteration, the
 * direction of the gathered columns formed on the fly; 1 = three cooperative grid barriers.  Bitwise equal results. */
int32_t tb_cg_set_persistent_variant(tb_ctx *ctx, int32_t variant);
int32_t tb_cg_last_path(tb_ctx *ctx, int32_t *path);
/* Order-independent dot products (off by default; env TB_DOT_EXACT=1).  Krylov's cg! forms r.r and p.Ap with BLAS `dot`
 * (euler.jl:94 -> KrylovJL_CG); any two implementations differ in summation order, and on ill-conditioned operators
 * (LV meshes) that rounding noise decides on which side of `atol + rtol |r0|` an iterate falls -- hence the +-1 in the
 * iteration rule.  With exact = 1 every CG dot product is accumulated in double-double (exact products by fma, error-free
 * additions) and rounded once at the end, across blocks and across GPUs, so alpha, beta, the residual norms, the stopping
 * decision and therefore every iterate are the same bits for any grid size and any number of GPUs -- and equal to a CPU
 * reference that rounds its dot products the same way.  Costs ~20 extra fp64 operations per row in kernels that are
 * bandwidth bound; forces solver path 0. */
int32_t tb_cg_set_exact_dot(tb_ctx *ctx, int32_t on);

/* ---- cell sweep -------------------------------------------------------------------------------
 * _pointwise_step_outer_kernel! (src/solver/time/partitioned_solver.jl:38-52; the method
 * ext/CuThunderboltExt.jl:111-124 provided for CuVector).  substeps <= 1: ForwardEulerCellSolver
 * (:80-99); substeps > 1: AdaptiveForwardEulerSubstepper (:196-234) with `reaction_threshold`.
 * `u` holds num_states columns; phi_idx is the 0-based state column of the transmembrane potential.
 * max_dphi (nullable) receives the signed max_i du[i, phi] of the LAST rhs evaluation of the sweep, i.e. what the
 * reference's cache.dumat holds afterwards (rtc.jl:64-67); -inf when the vector has no rows. */
int32_t tb_cell_step(tb_ctx *ctx, int32_t model, const double *params, int32_t nparams, tb_vec *u, int32_t phi_idx,
                     double t, double dt, int32_t substeps, double reaction_threshold, double *max_dphi);

/* ---- multi-subdomain splits (PointwiseMultiODEFunction, heat_dofrange views, interface diffusion) -------------------
 * semidiscretize(::ReactionDiffusionSplit{<:Dict}) (src/discretization/fem.jl:434-542) packs the solution vector
 * subdomain by subdomain, every block in PointBlockedLayout (src/modeling/solution_variables.jl:41-68), and gives the
 * heat sub-problem the scattered index set heat_dofrange; perform_step!(::PointwiseMultiODEFunction, ...) sweeps the
 * blocks one after the other (src/solver/time/partitioned_solver.jl:23-35,126-155). */
typedef struct tb_index tb_index;
int32_t tb_index_create(tb_ctx *ctx, const int64_t *idx, int64_t n, int32_t index_base, tb_index **out);
int32_t tb_index_destroy(tb_index *ix);
/* dst[i, dcol] = src[idx[i], scol]  -- the copy behind `view(u, heat_dofrange)` handed to the heat solver */
int32_t tb_vec_gather(tb_vec *dst, int32_t dcol, const tb_vec *src, int32_t scol, const tb_index *ix);
/* dst[idx[i], dcol] = src[i, scol] */
int32_t tb_vec_scatter(tb_vec *dst, int32_t dcol, const tb_index *ix, const tb_vec *src, int32_t scol);

#define TB_LAYOUT_STATE_BLOCKED 0 /* StateBlockedLayout: state s of point k at offset + s*npoints + k */
#define TB_LAYOUT_POINT_BLOCKED 1 /* PointBlockedLayout: state s of point k at offset + k*nstates + s */
typedef struct tb_cell_block {   /* StateBlock(offset, npoints, nstates, layout) + the PointwiseODEFunction's model */
    int64_t offset;              /* 0-based, in doubles, into the flat state vector */
    int64_t npoints;
    int32_t model;               /* TB_FHN | TB_PCG2019 | TB_ALIEV_PANFILOV */
    int32_t layout;
    int32_t nparams;
    int32_t reserved;
    double params[36];
} tb_cell_block;
/* one cell sweep per block, in order; `u` is ONE flat column.  max_dphi (nullable): max over all blocks of the phi_m
 * component of the last rhs evaluation (rtc.jl:68-73 loops the children the same way). */
int32_t tb_cell_step_blocks(tb_ctx *ctx, const tb_cell_block *blocks, int32_t nblocks, tb_vec *u, double t, double dt,
                            int32_t substeps, double reaction_threshold, double *max_dphi);

/* BilinearInterfaceDiffusionIntegrator (src/modeling/core/diffusion.jl:81-140): K_e[i,j] -= [[N_i]] D [[N_j]] dGamma over
 * interface cells, each a pair of coincident facets ("here", "there") whose nodes were duplicated.  dofs: nif x 2k ids
 * (here side first), coords_*: nif x k x sdim.  K is zeroed first like every update_operator!; combine it with the bulk
 * operator through tb_csr_axpby_values.  The scatter runs in interface-cell order (bitwise the sequential loop). */
#define TB_FACET_LINE2 0
#define TB_FACET_QUAD4 1
int32_t tb_assemble_interface_diffusion(tb_ctx *ctx, int32_t facet_type, int32_t sdim, int64_t nif, const int64_t *dofs,
                                        int32_t index_base, const double *coords_here, const double *coords_there,
                                        int32_t qorder, double D, tb_csr *K);

/* ---- building blocks of the Poisson / Geselowitz ECG reconstructions (src/modeling/electrophysiology/ecg.jl:166-619) ----
 * Everything else they need already exists: rectangular operators for the heart -> torso transfer and the electrode
 * evaluation (tb_csr_create with ncols != nrows + tb_csr_values_upload + tb_spmv), tb_assemble_diffusion on the torso
 * mesh (TB_D_CELL_TENSOR for the heart conductivity extended by zero), tb_cg_solve for the lead fields and for phi_e. */
/* out[c] = sum_i Z[i, c] * v[i, vcol] for every column c of Z: `-Z * kappa_grad_phi` (ecg.jl:617-619); blocks until done */
int32_t tb_vec_dots(tb_ctx *ctx, const tb_vec *Z, const tb_vec *v, int32_t vcol, double *out);
/* d[r, col] = A[r, r] */
int32_t tb_csr_diagonal(const tb_csr *A, tb_vec *d, int32_t col);
/* Ferrite apply_zero!(K, f, ch), matrix half: zero the rows and columns of the constrained dofs, put diag_value on their
 * diagonal (Ferrite: mean |K_ii|); right-hand-side half: tb_vec_fill_at(f, col, constrained, 0.0) */
int32_t tb_csr_apply_zero(tb_csr *A, const tb_index *constrained, double diag_value);
int32_t tb_vec_fill_at(tb_vec *v, int32_t col, const tb_index *ix, double value);

/* ---- output staging for store_timestep_field! (src/ferrite-addons/io.jl:18-93) -----------------------------------------
 * A snapshot of one state column is copied to PINNED host memory on a dedicated copy stream, ordered after the work already
 * queued on the compute stream; the call returns at once, later kernels that overwrite the column wait for the copy, the
 * host writer reads the buffer after tb_stage_wait.  The hot path never blocks on I/O. */
int32_t tb_host_alloc(int64_t bytes, void **out);   /* cudaMallocHost */
int32_t tb_host_free(void *p);
int32_t tb_vec_stage_col(const tb_vec *v, int32_t col, double *pinned_host);
int32_t tb_stage_wait(tb_ctx *ctx);

/* ---- fused LieTrotterGodunov step -------------------------------------------------------------
 * One OS.LieTrotterGodunov((BackwardEulerSolver, cell solver)) step (operatorsplitting-interface.jl:23-232;
 * perform_backward_euler_step!, euler.jl:71-101; partitioned_solver.jl:14-21): refresh A when dt
 * changed, b = M*phi (+ bS), CG, phi <- x, cell sweep -- without returning to the host in between. */
int32_t tb_monodomain_create(tb_ctx *ctx, const tb_csr *M, const tb_csr *K, int32_t model, const double *params,
                             int32_t nparams, int32_t phi_idx, tb_monodomain **out);
int32_t tb_monodomain_destroy(tb_monodomain *md);
int32_t tb_monodomain_set_cg(tb_monodomain *md, double atol, double rtol, int64_t itmax);
int32_t tb_monodomain_set_preconditioner(tb_monodomain *md, int32_t precond);   /* TB_PRECOND_* for the inner CG */
int32_t tb_monodomain_set_cell_solver(tb_monodomain *md, int32_t substeps, double reaction_threshold);
/* bS (nullable): the source operator's vector, added to b as is (euler.jl:88-91).  Borrowed. */
int32_t tb_monodomain_set_source(tb_monodomain *md, const tb_vec *bS, int32_t col);
/* after the system matrix is formed K's values may be released to save HBM when dt is fixed */
int32_t tb_monodomain_step(tb_monodomain *md, tb_vec *u, double t, double dt, int64_t *iters, double *rnorm,
                           int32_t *converged);
/* the same step, also returning the reaction tangent R = max_i du[i, phi_m] left by the cell sweep's last rhs
 * evaluation, maximum over all ranks and clamped at 0 like _get_reaction_tangent (R = 0.0; R = max(R, maximum(...))):
 * the input of ReactionTangentController (src/solver/time/rtc.jl:51-78,121-133) */
int32_t tb_monodomain_step_rt(tb_monodomain *md, tb_vec *u, double t, double dt, int64_t *iters, double *rnorm,
                              int32_t *converged, double *reaction_tangent);
/* `nsteps` steps back to back with one host round trip at the end (iters = total, converged = all) */
/* (tb_monodomain_run: when the inner solve is one persistent kernel -- operators up to ~4 M rows, no block-Jacobi / Chebyshev --
 * the whole run is enqueued without a host read-back; iteration counts and convergence flags are accumulated on the device) */
int32_t tb_monodomain_run(tb_monodomain *md, tb_vec *u, double t0, double dt, int64_t nsteps, int64_t *iters_total,
                          int32_t *all_converged);
/* end-to-end variant on HOST buffers: uploads u_in, steps, downloads into u_out (may alias u_in) */
int32_t tb_monodomain_step_host(tb_monodomain *md, tb_vec *u_dev, const double *u_in, double *u_out, double t, double dt,
                                int64_t *iters, double *rnorm, int32_t *converged);
/* `nsteps` steps with the state living in HOST memory between steps (what a host-side integrator that owns `u` sees):
 * step n reads buf[n & 1] and writes buf[(n+1) & 1]; layout of a buffer = tb_vec_download's (column c at c*n).  Every
 * step uploads the whole state and downloads the whole result, but the transfers are pipelined on two copy streams:
 * the download of step n's phi (in chunks) runs full duplex with the upload of step n+1's phi, and the other state
 * columns travel while CG runs (the diffusion solve only needs phi).  Pinned buffers are required for the overlap. */
int32_t tb_monodomain_run_host(tb_monodomain *md, tb_vec *u_dev, double *buf0, double *buf1, double t0, double dt,
                               int64_t nsteps, int64_t *iters_total, int32_t *all_converged);
/* number of pieces (1..64, default 16 or env TB_RUNHOST_CHUNKS) the phi column is cut into for the chunk-chased
 * download / upload of tb_monodomain_run_host */
int32_t tb_monodomain_set_host_chunks(tb_monodomain *md, int32_t nchunks);
/* per-section CUDA-event timings of the last step, ms: [0] "b = M u", [1] "inner solve",
 * [2] "reaction solve" (the reference's TimerOutputs labels, euler.jl:85,94; partitioned_solver.jl:20) */
int32_t tb_monodomain_section_ms(tb_monodomain *md, double *ms3);
int32_t tb_monodomain_enable_timing(tb_monodomain *md, int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* TBOLT_B200_H */
