# ThunderboltB200Ext.jl -- the reference-side binding of libtbolt_b200.so.
#
# Replaces ext/CuThunderboltExt.jl (which is stale in v0.0.4, SURVEY 0-3) at exactly the dispatch points that
# extension hooks (SURVEY 8b).  NOT EXECUTED in this repository's CI: there is no Julia toolchain in the build
# image; the same C ABI is exercised through ctypes by thunderbolt.jl_b200/api.py and tests/.
#
# Load as a package extension of Thunderbolt (Project.toml: [weakdeps]/[extensions]) or `include` it after
# `using Thunderbolt`.
module ThunderboltB200Ext

using Thunderbolt, LinearSolve, SciMLBase, LinearAlgebra, SparseArrays
import Thunderbolt: create_system_vector, create_system_matrix, adapt_vector_type, setup_operator, update_operator!,
    needs_update, _pointwise_step_outer_kernel!, PointwiseODEFunction, AbstractPointwiseSolverCache,
    AdaptiveForwardEulerSubstepperCache, ForwardEulerCellSolverCache, BilinearMassIntegrator, BilinearDiffusionIntegrator,
    LinearIntegrator, AnalyticalTransmembraneStimulationProtocol, ParametrizedFHNModel, ParametrizedPCG2019Model,
    ConductivityToDiffusivityCoefficient, ConstantCoefficient, ElementAssemblyStrategy, AbstractGPUDevice, num_states
import Ferrite, Tensors

const LIB = Ref{String}("libtbolt_b200.so")

struct TBError <: Exception
    status::Int32
    msg::String
end
check(status::Int32) = status == 0 ? nothing : throw(TBError(status, unsafe_string(ccall((:tb_last_error, LIB[]), Cstring, ()))))
macro tb(f, argtypes, args...)
    esc(:(check(ccall(($(QuoteNode(f)), LIB[]), Int32, $argtypes, $(args...)))))
end

# ---- device (src/devices.jl:1-4) -------------------------------------------------------------------
mutable struct B200Device <: AbstractGPUDevice
    h::Ptr{Cvoid}
    function B200Device(id::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        @tb tb_ctx_create (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}) Int32(id) C_NULL r
        finalizer(d -> ccall((:tb_ctx_destroy, LIB[]), Int32, (Ptr{Cvoid},), d.h), new(r[]))
    end
end
const DEFAULT_DEVICE = Ref{Union{Nothing, B200Device}}(nothing)
default_device() = something(DEFAULT_DEVICE[], (DEFAULT_DEVICE[] = B200Device(0)))

# ---- vectors (src/solver/interface.jl:175-181, src/utils.jl:425-427) ---------------------------------
mutable struct B200Vector{T} <: AbstractVector{T}
    h::Ptr{Cvoid}
    n::Int
    ncols::Int
    dev::B200Device
end
function B200Vector{Float64}(dev::B200Device, n::Integer, ncols::Integer = 1)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    @tb tb_vec_create (Ptr{Cvoid}, Int64, Int32, Ref{Ptr{Cvoid}}) dev.h n ncols r
    finalizer(v -> ccall((:tb_vec_destroy, LIB[]), Int32, (Ptr{Cvoid},), v.h), B200Vector{Float64}(r[], n, ncols, dev))
end
Base.size(v::B200Vector) = (v.n * v.ncols,)
function B200Vector(u::Vector{Float64}; nstates = 1, dev = default_device())
    v = B200Vector{Float64}(dev, length(u) ÷ nstates, nstates)
    @tb tb_vec_upload (Ptr{Cvoid}, Ptr{Float64}) v.h u
    v
end
function Base.Array(v::B200Vector{Float64})
    out = Vector{Float64}(undef, v.n * v.ncols)
    @tb tb_vec_download (Ptr{Cvoid}, Ptr{Float64}) v.h out
    out
end
create_system_vector(::Type{<:B200Vector{T}}, f::Thunderbolt.AbstractSemidiscreteFunction) where {T} =
    B200Vector{T}(default_device(), Thunderbolt.solution_size(f))
create_system_vector(::Type{<:B200Vector{T}}, dh::Ferrite.DofHandler) where {T} = B200Vector{T}(default_device(), Ferrite.ndofs(dh))
adapt_vector_type(::Type{<:B200Vector}, v::Vector) = v   # coordinates stay on the host: FHN/PCG2019 ignore x (SURVEY a-5)

# ---- mesh handle: what Ferrite hands the operators (fem.jl:180-182) ----------------------------------
const CELLTYPE = Dict(Ferrite.Quadrilateral => 0, Ferrite.Hexahedron => 1, Ferrite.Triangle => 2, Ferrite.Tetrahedron => 3)
mutable struct B200Mesh
    h::Ptr{Cvoid}
    dev::B200Device
end
function B200Mesh(dev::B200Device, dh::Ferrite.DofHandler)
    grid = Ferrite.get_grid(dh)
    CT = typeof(first(grid.cells))
    nv = length(first(grid.cells).nodes)
    conn = Int64[n for c in grid.cells for n in c.nodes]                       # 1-based node ids
    coords = Float64[x for n in grid.nodes for x in n.x]
    cdofs = Int64[d for c in 1:Ferrite.getncells(grid) for d in Ferrite.celldofs(dh, c)]
    r = Ref{Ptr{Cvoid}}(C_NULL)
    @tb tb_mesh_create (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Int64}, Int64, Int32, Ref{Ptr{Cvoid}}) dev.h Int32(CELLTYPE[CT]) length(grid.cells) length(grid.nodes) conn coords cdofs Ferrite.ndofs(dh) Int32(1) r
    finalizer(m -> ccall((:tb_mesh_destroy, LIB[]), Int32, (Ptr{Cvoid},), m.h), B200Mesh(r[], dev))
end

# ---- matrices (src/solver/interface.jl:159-173) --------------------------------------------------------
mutable struct B200CSRMatrix{Tv, Ti} <: AbstractSparseMatrix{Tv, Ti}
    h::Ptr{Cvoid}
    n::Int
    dev::B200Device
end
function create_system_matrix(::Type{<:B200CSRMatrix{Tv, Ti}}, dh::Ferrite.AbstractDofHandler) where {Tv, Ti}
    dev = default_device()
    # Ferrite's pattern, transposed to CSR exactly like the ThreadedSparseMatrixCSR method: bit-exact pattern parity
    Acsc = convert(SparseMatrixCSC{Tv, Int64}, Ferrite.allocate_matrix(dh))
    r = Ref{Ptr{Cvoid}}(C_NULL)
    n = size(Acsc, 1)
    @tb tb_csr_create (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ref{Ptr{Cvoid}}) dev.h n n Acsc.colptr Acsc.rowval Int32(1) r
    finalizer(A -> ccall((:tb_csr_destroy, LIB[]), Int32, (Ptr{Cvoid},), A.h), B200CSRMatrix{Tv, Ti}(r[], n, dev))
end
Base.size(A::B200CSRMatrix) = (A.n, A.n)
function SparseArrays.nonzeros(A::B200CSRMatrix{Tv}) where {Tv}
    nnz = Ref{Int64}(0)
    @tb tb_csr_sizes (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Int64}) A.h C_NULL C_NULL nnz
    v = Vector{Tv}(undef, nnz[])
    @tb tb_csr_values_download (Ptr{Cvoid}, Ptr{Float64}) A.h v
    v
end
LinearAlgebra.mul!(y::B200Vector, A::B200CSRMatrix, x::B200Vector) =
    (@tb tb_spmv (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32) A.dev.h A.h x.h Int32(0) y.h Int32(0); y)
# euler.jl:104-116
Thunderbolt._implicit_euler_heat_solver_update_system_matrix!(A::B200CSRMatrix, M, K, Δt) =
    @tb tb_csr_axpby_values (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64) A.h M.A.h K.A.h Float64(Δt)

# ---- operators (src/solver/interface.jl:17-94; discretization/operator.jl:2-32) ---------------------------
struct B200BilinearOperator{I}
    A::B200CSRMatrix{Float64, Int32}
    integrator::I
    dh::Ferrite.DofHandler
    mesh::B200Mesh
end
struct B200LinearOperator{I}
    b::B200Vector{Float64}
    integrator::I
    dh::Ferrite.DofHandler
    mesh::B200Mesh
end
const MESHES = IdDict{Any, B200Mesh}()
mesh_of(dev, dh) = get!(() -> B200Mesh(dev, dh), MESHES, dh)

function setup_operator(s::ElementAssemblyStrategy{<:B200Device}, i::Thunderbolt.AbstractBilinearIntegrator, solver::Thunderbolt.AbstractSolver, dh)
    B200BilinearOperator(create_system_matrix(B200CSRMatrix{Float64, Int32}, dh), i, dh, mesh_of(s.device, dh))
end
function setup_operator(s::ElementAssemblyStrategy{<:B200Device}, i::LinearIntegrator, dh)
    B200LinearOperator(B200Vector{Float64}(s.device, Ferrite.ndofs(dh)), i, dh, mesh_of(s.device, dh))
end
# FerriteOperators' QuadratureRuleCollection carries its order as the type parameter (`QuadratureRuleCollection(2)` ==
# `QuadratureRuleCollection{2}()`, fem.jl:52-55) -- no upstream helper needed
qorder_of(::Thunderbolt.QuadratureRuleCollection{order}) where {order} = order
qorder(i) = qorder_of(i.qrc)

function update_operator!(op::B200BilinearOperator{<:BilinearMassIntegrator}, t)
    @tb tb_assemble_mass (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, Ptr{Cvoid}) op.A.dev.h op.mesh.h Int32(qorder(op.integrator)) Float64(op.integrator.ρ.val) op.A.h
end
# lower the coefficient tree of a diffusion integrator to (kind, data, Cₘχ) of the C ABI
function diffusion_data(D, mesh)
    cmχ = 1.0
    if D isa ConductivityToDiffusivityCoefficient                                  # coefficients.jl:122-162
        cmχ = D.capacitance_coefficient.val * D.χ_coefficient.val
        D = D.conductivity_tensor_coefficient
    end
    κ = D.val                                                                       # ConstantCoefficient
    data = κ isa Number ? Float64[κ] : Float64[κ[i, j] for i in 1:size(κ, 1) for j in 1:size(κ, 2)]
    kind = κ isa Number ? Int32(0) : Int32(1)             # TB_D_SCALAR / TB_D_TENSOR; SpectralTensorCoefficient: kind 2 with
    return kind, data, Float64(cmχ)                       # λ[3] + per-cell nodal f,s,n (see tbolt_b200.h), same call
end
function update_operator!(op::B200BilinearOperator{<:BilinearDiffusionIntegrator}, t)
    kind, data, cmχ = diffusion_data(op.integrator.D, op.mesh)
    @tb tb_assemble_diffusion (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64, Float64, Ptr{Cvoid}) op.A.dev.h op.mesh.h Int32(qorder(op.integrator)) Int32(kind) data length(data) Float64(cmχ) op.A.h
end
# A Julia closure cannot cross the C ABI: evaluate f at the quadrature points on the host, ship the values.
# x_q = Σ_a M_a(ξ_q) x_a exactly as AnalyticalCoefficientElementCache does (coefficients.jl:279-292), through Ferrite's
# own CellValues / spatial_coordinate -- nothing here is missing upstream.  The result is ncells × nq doubles in
# cell-major order; it is re-evaluated only while t lies in a nonzero interval (needs_update below).
function evaluate_at_quadrature_points(f, dh::Ferrite.DofHandler, qorder::Integer, t)
    grid = Ferrite.get_grid(dh)
    sdh = first(dh.subdofhandlers)
    CT = typeof(Ferrite.getcells(grid, first(sdh.cellset)))
    RS = Ferrite.getrefshape(CT)
    ip = Ferrite.geometric_interpolation(CT)
    qr = Ferrite.QuadratureRule{RS}(qorder)
    cv = Ferrite.CellValues(qr, Ferrite.Lagrange{RS, 1}(), ip)
    nq = Ferrite.getnquadpoints(cv)
    fq = Matrix{Float64}(undef, nq, Ferrite.getncells(grid))
    for cell in Ferrite.CellIterator(dh)
        Ferrite.reinit!(cv, cell)
        x = Ferrite.getcoordinates(cell)
        for q in 1:nq
            fq[q, Ferrite.cellid(cell)] = f(Ferrite.spatial_coordinate(cv, q, x), t)
        end
    end
    return fq
end
# ... but its EXPRESSION can.  `f(x, t)` is called once with tracing numbers; the recorded expression tree becomes the postfix
# program of tb_assemble_source_program (tbolt_b200.h, TB_SRC_PROGRAM: the opcode order below is the header's) that the
# element kernel evaluates at every quadrature point -- per update only the ~400-byte program crosses the host link, not
# ncells × nq doubles.  `&&` / `||` / `if` on x or t need a Bool and throw a TypeError while tracing (use `ifelse`, `&`, `|`,
# `min`, `max`): then, or when the program would exceed 96 instructions, the closure keeps the host-evaluated path above.
const TB_OPS = (:X, :T, :CONST, :ADD, :SUB, :MUL, :DIV, :MIN, :MAX, :POW, :LT, :LE, :GT, :GE, :EQ, :NE, :AND, :OR,
                :NEG, :ABS, :SQRT, :EXP, :LOG, :SIN, :COS, :TANH, :NOT, :SELECT)
const TB_OPCODE = Dict(op => Int32(i - 1) for (i, op) in enumerate(TB_OPS))
struct TracedReal <: Real
    op::Symbol
    args::Tuple
end
TracedReal(v::Real) = v isa TracedReal ? v : TracedReal(:CONST, (Float64(v),))
Base.promote_rule(::Type{TracedReal}, ::Type{<:Real}) = TracedReal
Base.convert(::Type{TracedReal}, v::Real) = TracedReal(v)
Base.convert(::Type{TracedReal}, v::TracedReal) = v
for (fn, op) in ((:+, :ADD), (:-, :SUB), (:*, :MUL), (:/, :DIV), (:min, :MIN), (:max, :MAX), (:^, :POW),
                 (:<, :LT), (:<=, :LE), (:(==), :EQ), (:&, :AND), (:|, :OR))
    @eval Base.$fn(a::TracedReal, b::TracedReal) = TracedReal($(QuoteNode(op)), (a, b))
end
Base.:>(a::TracedReal, b::TracedReal) = TracedReal(:GT, (a, b))
Base.:>=(a::TracedReal, b::TracedReal) = TracedReal(:GE, (a, b))
Base.:!=(a::TracedReal, b::TracedReal) = TracedReal(:NE, (a, b))
Base.literal_pow(::typeof(^), a::TracedReal, ::Val{2}) = a * a               # what Julia does for Float64
Base.literal_pow(::typeof(^), a::TracedReal, ::Val{3}) = a * a * a
for (fn, op) in ((:-, :NEG), (:abs, :ABS), (:sqrt, :SQRT), (:exp, :EXP), (:log, :LOG), (:sin, :SIN), (:cos, :COS),
                 (:tanh, :TANH), (:!, :NOT))
    @eval Base.$fn(a::TracedReal) = TracedReal($(QuoteNode(op)), (a,))
end
Base.ifelse(c::TracedReal, a::Real, b::Real) = TracedReal(:SELECT, (c, TracedReal(a), TracedReal(b)))
Base.abs2(a::TracedReal) = a * a
Base.zero(::Type{TracedReal}) = TracedReal(0.0); Base.one(::Type{TracedReal}) = TracedReal(1.0)
function emit!(code::Vector{Int32}, consts::Vector{Float64}, n::TracedReal)
    if n.op === :X
        push!(code, TB_OPCODE[:X] | Int32(n.args[1]) << 8)
    elseif n.op === :T
        push!(code, TB_OPCODE[:T])
    elseif n.op === :CONST
        k = findfirst(c -> c === n.args[1], consts)
        k === nothing && (push!(consts, n.args[1]); k = length(consts))
        push!(code, TB_OPCODE[:CONST] | Int32(k - 1) << 8)
    else
        foreach(a -> emit!(code, consts, a), n.args)
        push!(code, TB_OPCODE[n.op])
    end
    return code
end
const TRACED_SOURCES = IdDict{Any, Any}()        # closure => (code, consts) | nothing
"postfix program of `f(x::Vec{dim}, t)`, or `nothing` if f cannot be traced"
function trace_source(f, dim::Integer)
    try
        x = Tensors.Vec{dim, TracedReal}(ntuple(d -> TracedReal(:X, (d - 1,)), dim))
        root = TracedReal(f(x, TracedReal(:T, ())))
        code, consts = Int32[], Float64[]
        emit!(code, consts, root)
        (length(code) > 96 || length(consts) > 24) && return nothing
        return code, consts            # the library validates the stack depth (tb_assemble_source_program)
    catch err
        err isa Union{TypeError, MethodError} || rethrow()
        return nothing
    end
end
function update_operator!(op::B200LinearOperator, t)
    proto = op.integrator.integrand::AnalyticalTransmembraneStimulationProtocol
    prog = get!(() -> trace_source(proto.f.f, Ferrite.getspatialdim(Ferrite.get_grid(op.dh))), TRACED_SOURCES, proto.f.f)
    if prog !== nothing
        code, consts = prog
        rc = ccall((:tb_assemble_source_program, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int32}, Int32, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}, Int32),
                   op.b.dev.h, op.mesh.h, Int32(qorder(op.integrator)), code, length(code), consts, length(consts), Float64(t), op.b.h, Int32(0))
        rc == 0 && return
        TRACED_SOURCES[proto.f.f] = nothing          # e.g. deeper than the 16-entry stack: host-evaluated from now on
    end
    fq = evaluate_at_quadrature_points(proto.f.f, op.dh, qorder(op.integrator), t)
    @tb tb_assemble_source_qp (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Cvoid}, Int32) op.b.dev.h op.mesh.h Int32(qorder(op.integrator)) fq op.b.h Int32(0)
end
needs_update(op::B200LinearOperator, t) = any(iv -> iv[1] ≤ t ≤ iv[2], op.integrator.integrand.nonzero_intervals)
LinearAlgebra.mul!(b::B200Vector, M::B200BilinearOperator, u::B200Vector) = mul!(b, M.A, u)
Thunderbolt.add!(b::B200Vector, S::B200LinearOperator) =
    @tb tb_vec_axpy (Ptr{Cvoid}, Int32, Float64, Ptr{Cvoid}, Int32) b.h Int32(0) 1.0 S.b.h Int32(0)

# ---- linear solve (euler.jl:94,155-156) ----------------------------------------------------------------
Base.@kwdef struct B200CG <: LinearSolve.SciMLLinearSolveAlgorithm
    atol::Float64 = sqrt(eps(Float64))
    rtol::Float64 = sqrt(eps(Float64))
    maxiters::Int = 0            # 0 -> length(b), LinearSolve's default
    precs::Symbol = :none        # :none | :jacobi | :block_jacobi | :chebyshev -- where KrylovJL_CG takes `precs = ..., ldiv = false`
                                 # (bak/examples-gpu/spiral-wave.jl:95-105: BlockJacobiPreconditioner(A, 1000, CUDABackend()))
    nblocks::Int = 1000          # :block_jacobi -- number of diagonal blocks (equal contiguous dof ranges unless row_block is given)
    row_block::Union{Nothing, Vector{Int32}} = nothing   # block id per row, e.g. a Metis partition (0-based)
    degree::Int = 8              # :chebyshev -- polynomial degree and lmax / lmin of the target interval
    ratio::Float64 = 30.0
end
precond_id(alg::B200CG) = Int32(Dict(:none => 0, :jacobi => 1, :block_jacobi => 2, :chebyshev => 3)[alg.precs])   # TB_PRECOND_*
# what `precs(A, p)` does in the reference: hand the preconditioner's parameters over (the inverses / the Gershgorin bound
# are rebuilt by the library whenever the operator's values change -- update!(P, A))
function configure!(alg::B200CG, A::B200CSRMatrix)
    if alg.precs === :block_jacobi
        rb = alg.row_block === nothing ? C_NULL : pointer(alg.row_block)
        GC.@preserve alg @tb tb_cg_set_block_jacobi (Ptr{Cvoid}, Int64, Int64, Ptr{Int32}) A.dev.h Int64(A.n) Int64(alg.nblocks) rb
    elseif alg.precs === :chebyshev
        @tb tb_cg_set_chebyshev (Ptr{Cvoid}, Int32, Float64) A.dev.h Int32(alg.degree) alg.ratio
    end
end
# order-independent CG dot products (double-double accumulation, rounded once): iterates independent of grid size / GPU count
exact_dot!(dev::B200Device, on::Bool = true) = @tb tb_cg_set_exact_dot (Ptr{Cvoid}, Int32) dev.h Int32(on)
mutable struct B200CGCache
    iters::Int64
    resid::Float64
end
LinearSolve.init_cacheval(::B200CG, A, b, u, Pl, Pr, maxiters, abstol, reltol, verbose, assumptions) = B200CGCache(0, 0.0)
function SciMLBase.solve!(cache::LinearSolve.LinearCache, alg::B200CG; kwargs...)
    A, b, u = cache.A::B200CSRMatrix, cache.b::B200Vector, cache.u::B200Vector
    configure!(alg, A)
    it, rn, conv = Ref{Int64}(0), Ref{Float64}(0.0), Ref{Int32}(0)
    itmax = alg.maxiters == 0 ? length(b) : alg.maxiters
    @tb tb_cg_solve_pc (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int32, Float64, Float64, Int64, Ref{Int64}, Ref{Float64}, Ref{Int32}) A.dev.h A.h b.h Int32(0) u.h Int32(0) precond_id(alg) alg.atol alg.rtol itmax it rn conv
    cache.cacheval.iters, cache.cacheval.resid = it[], rn[]
    # non-convergence is a retcode, not an exception: euler.jl:95-100 then returns false and the integrator rolls back
    SciMLBase.build_linear_solution(alg, u, nothing, cache; retcode = conv[] == 1 ? ReturnCode.Success : ReturnCode.MaxIters, iters = it[])
end

# ---- cell sweep (partitioned_solver.jl:38-52; the method CuThunderboltExt.jl:111-124 had for CuVector) -----------
model_id(::ParametrizedFHNModel) = Int32(0)
model_id(::ParametrizedPCG2019Model) = Int32(1)
model_id(::Thunderbolt.ParametrizedAlievPanfilovModel) = Int32(2)   # states (s, φₘ): transmembranepotential_index == 2
params(m) = Float64[getfield(m, f) for f in fieldnames(typeof(m))]
substeps(c::ForwardEulerCellSolverCache) = (Int32(1), 0.1)
substeps(c::AdaptiveForwardEulerSubstepperCache) = (Int32(c.substeps), Float64(c.reaction_threshold))
function _pointwise_step_outer_kernel!(f::PointwiseODEFunction, t::Real, Δt::Real, cache::AbstractPointwiseSolverCache, u::B200Vector)
    p = params(f.ode)
    ns, thr = substeps(cache)
    φidx = Thunderbolt.transmembranepotential_index(f.ode) - 1
    # The reaction tangent R = maximum(dumat[:, φₘidx]) that ReactionTangentController reads (src/solver/time/rtc.jl:51-78)
    # is reduced inside the sweep and kept on the cache instead of a dumat that never exists on the device.
    R = Ref{Float64}(0.0)
    @tb tb_cell_step (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Ptr{Cvoid}, Int32, Float64, Float64, Int32, Float64, Ptr{Float64}) u.dev.h model_id(f.ode) p Int32(length(p)) u.h Int32(φidx) Float64(t) Float64(Δt) ns thr R
    REACTION_TANGENT[objectid(cache)] = R[]
    return true
end

# ---- ReactionTangentController (src/solver/time/rtc.jl:51-78) ------------------------------------------------
# get_reaction_tangent walks the sub-integrators and takes maximum(cache.dumat[:, φₘidx]); for a B200Vector-typed
# pointwise cache the value was already reduced on the device by the sweep above.
const REACTION_TANGENT = Dict{UInt, Float64}()
reaction_tangent(cache::AbstractPointwiseSolverCache) = get(REACTION_TANGENT, objectid(cache), 0.0)
# REQUIRES AN UPSTREAM PATCH (get_reaction_tangent indexes cache.dumat directly and offers no dispatch point); the one
# changed expression, rtc.jl:64-67:
#   R = max(R, subintegrator.cache.uₙ isa B200Vector ? reaction_tangent(subintegrator.cache) :
#                                                       maximum(@view subintegrator.cache.dumat[:, φₘidx]))

# ---- pseudo-ECG (src/modeling/electrophysiology/ecg.jl:55-160) -------------------------------------------------------
# Plonsey1964ECGGaussCache keeps κ∇φₘ at every quadrature point between update_ecg! and evaluate_ecg; for a
# B200-backed diffusion operator the cache keeps a device copy of φₘ instead and the flux is formed in registers inside
# the one element sweep tb_ecg_plonsey launches.
struct B200PlonseyECGCache{O}
    op::O                       # B200BilinearOperator{<:BilinearDiffusionIntegrator}
    φₘ::B200Vector{Float64}
end
function Thunderbolt.Plonsey1964ECGGaussCache(op::B200BilinearOperator{<:BilinearDiffusionIntegrator}, φₘ::B200Vector)
    c = B200PlonseyECGCache(op, B200Vector{Float64}(φₘ.dev, φₘ.n, 1))
    Thunderbolt.update_ecg!(c, φₘ)
    c
end
Thunderbolt.update_ecg!(c::B200PlonseyECGCache, φₘ::B200Vector) =
    @tb tb_vec_copy (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32) c.φₘ.h Int32(0) φₘ.h Int32(0)
function Thunderbolt.evaluate_ecg(c::B200PlonseyECGCache, x::AbstractVector{<:Ferrite.Vec{sdim}}, κₜ::Real) where {sdim}
    kind, data, cmchi = diffusion_data(c.op.integrator.D, c.op.mesh)      # same lowering as update_operator! uses
    pts = Float64[xi[d] for xi in x for d in 1:sdim]
    out = zeros(length(x))
    @tb tb_ecg_plonsey (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64, Float64, Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Float64, Ptr{Float64}) c.op.mesh.dev.h c.op.mesh.h Int32(qorder(c.op.integrator)) kind data Int64(length(data)) cmchi c.φₘ.h Int32(0) pts Int32(length(x)) Float64(κₜ) out
    out
end
Thunderbolt.evaluate_ecg(c::B200PlonseyECGCache, x::Ferrite.Vec, κₜ::Real) = Thunderbolt.evaluate_ecg(c, [x], κₜ)[1]

# ---- multi-subdomain splits (fem.jl:434-542; partitioned_solver.jl:23-35,126-155) -----------------------------------------
# A child of a PointwiseMultiODEFunction owns the slots `f.associated_states` of the flat solution vector in
# PointBlockedLayout; perform_step!(::PointwiseMultiODEFunction) calls the outer kernel once per child.
Base.@kwdef struct TBCellBlock            # tb_cell_block (include/tbolt_b200.h)
    offset::Int64
    npoints::Int64
    model::Int32
    layout::Int32                         # 0 StateBlockedLayout, 1 PointBlockedLayout
    nparams::Int32
    reserved::Int32 = 0
    params::NTuple{36, Float64}
end
function cell_block(f::PointwiseODEFunction)
    p = params(f.ode)
    ns = num_states(f.ode)
    TBCellBlock(offset = first(f.associated_states) - 1, npoints = length(f.associated_states) ÷ ns, model = model_id(f.ode),
                layout = Int32(f.layout isa Thunderbolt.PointBlockedLayout), nparams = Int32(length(p)),
                params = ntuple(i -> i <= length(p) ? p[i] : 0.0, 36))
end
function step_block!(f::PointwiseODEFunction, t::Real, Δt::Real, cache::AbstractPointwiseSolverCache, u::B200Vector)
    ns, thr = substeps(cache)
    blk = Ref(cell_block(f))
    R = Ref{Float64}(0.0)
    @tb tb_cell_step_blocks (Ptr{Cvoid}, Ptr{TBCellBlock}, Int32, Ptr{Cvoid}, Float64, Float64, Int32, Float64, Ptr{Float64}) u.dev.h blk Int32(1) u.h Float64(t) Float64(Δt) ns thr R
    REACTION_TANGENT[objectid(cache)] = max(get(REACTION_TANGENT, objectid(cache), 0.0), R[])
    return true
end
# the heat child's `view(u, heat_dofrange)` with a scattered index set: forward / backward sync of the OS integrator
mutable struct B200Index
    h::Ptr{Cvoid}
end
function B200Index(dev::B200Device, idx::Vector{Int})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    @tb tb_index_create (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Ref{Ptr{Cvoid}}) dev.h idx Int64(length(idx)) Int32(1) r
    finalizer(i -> ccall((:tb_index_destroy, LIB[]), Int32, (Ptr{Cvoid},), i.h), B200Index(r[]))
end
gather!(dst::B200Vector, src::B200Vector, ix::B200Index) =
    @tb tb_vec_gather (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Ptr{Cvoid}) dst.h Int32(0) src.h Int32(0) ix.h
scatter!(dst::B200Vector, ix::B200Index, src::B200Vector) =
    @tb tb_vec_scatter (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int32) dst.h Int32(0) ix.h src.h Int32(0)
# BilinearInterfaceDiffusionIntegrator (diffusion.jl:81-140): dofs = ninterfaces x 2k (here side first, 1-based), coordinates of the
# two facets; K is zeroed first -- add it to the bulk operator with tb_csr_axpby_values(K, K_bulk, K_if, -1.0)
function assemble_interface_diffusion!(K::B200CSRMatrix, facet::Symbol, sdim::Integer, dofs::Matrix{Int64}, xh::Array{Float64}, xt::Array{Float64}, qorder::Integer, D::Real)
    @tb tb_assemble_interface_diffusion (Ptr{Cvoid}, Int32, Int32, Int64, Ptr{Int64}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Cvoid}) K.dev.h Int32(facet === :line ? 0 : 1) Int32(sdim) Int64(size(dofs, 2)) dofs Int32(1) xh xt Int32(qorder) Float64(D) K.h
end

# ---- lead-field / Poisson ECG building blocks (ecg.jl:166-619) ------------------------------------------------------------------
# transfer and electrode evaluation are rectangular B200CSRMatrix (tb_csr_create with ncols != nrows + tb_csr_values_upload);
# `-Z * κ∇φₘ_t` with the lead fields as the columns of one B200Vector:
function lead_potentials(Z::B200Vector, src::B200Vector)
    out = zeros(Z.ncols)
    @tb tb_vec_dots (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}) Z.dev.h Z.h src.h Int32(0) out
    return -out
end
# Ferrite.apply_zero!(K, f, ch): matrix half (diagonal of the constrained dofs := meandiag(K)) and right-hand-side half
apply_zero!(K::B200CSRMatrix, ix::B200Index, meandiag::Real) = @tb tb_csr_apply_zero (Ptr{Cvoid}, Ptr{Cvoid}, Float64) K.h ix.h Float64(meandiag)
apply_zero!(f::B200Vector, ix::B200Index) = @tb tb_vec_fill_at (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Float64) f.h Int32(0) ix.h 0.0

# ---- assembly strategy knobs ------------------------------------------------------------------------------------
# 2 = per-element results + ordered row gather (default: deterministic, bitwise the sequential CPU assembly),
# 0 = fp64 atomic scatter.  FerriteOperators' ElementAssemblyStrategy maps to 2.
set_assembly_mode!(dev::B200Device, mode::Integer) = @tb tb_assembly_set_mode (Ptr{Cvoid}, Int32) dev.h Int32(mode)
release_assembly_scratch!(dev::B200Device) = @tb tb_assembly_release_scratch (Ptr{Cvoid},) dev.h

# ---- multi-GPU (one Julia process per GPU, e.g. under MPI.jl; the reference itself is shared-memory only) -----------
# comm_init! joins the NCCL communicator; peer_attach! additionally maps every rank's mailbox window and CG work
# vectors through CUDA IPC so that the CG kernels exchange halos and dot products with plain NVLink stores.
function comm_init!(dev::B200Device, rank::Integer, nranks::Integer, unique_id::Vector{UInt8})
    @tb tb_ctx_comm_init (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}) dev.h Int32(rank) Int32(nranks) unique_id
end
function peer_export(dev::B200Device, ncols::Integer)
    blob = zeros(UInt8, 160)                       # TB_PEER_BLOB_BYTES
    @tb tb_peer_export (Ptr{Cvoid}, Int64, Ptr{UInt8}) dev.h Int64(ncols) blob
    blob
end
peer_attach!(dev::B200Device, blobs::Vector{UInt8}, nranks::Integer) =
    @tb tb_peer_attach (Ptr{Cvoid}, Ptr{UInt8}, Int32) dev.h blobs Int32(nranks)

end # module
