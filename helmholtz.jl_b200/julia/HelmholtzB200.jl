# HelmholtzB200.jl -- Julia shim over libhelmholtz_b200.so (include/helmholtz_b200.h).
#
# Drop-in for the acoustic solve path of JuliaInv/Helmholtz.jl: same exported names, argument orders and
# return shapes (HelmholtzParam, GetHelmholtzOperator, GetHelmholtzShiftOP, getABL, getMaximalFrequency,
# getMGparam, getShiftedLaplacianMultigridSolver, solveLinearSystem, solveLinearSystem!, copySolver, clear!).
# Every solve is one `ccall` into the shared library; no arithmetic happens in Julia.
#
# NOTE: no Julia runtime exists in the build image, so this file is written against the C header and kept
# trivially thin; the identical logic is exercised through the Python ctypes mirror (../api.py) by tests/.
#
# Reference lines mirrored: src/Helmholtz.jl:13-34; src/GetHelmholtz.jl:14-50,75-83,97-220;
# src/ShiftedLaplacianMultigridSolver.jl:4-30,33-109; src/getPointSource.jl:63-112.
module HelmholtzB200

using LinearAlgebra
using SparseArrays

# ---- jInv plug-in conformance -----------------------------------------------------------------------------------
# The reference's solver is a jInv linear-solver plug-in: it subtypes jInv.LinearSolvers.AbstractSolver and EXTENDS
# jInv's generic functions (src/Helmholtz.jl:4-8; src/ShiftedLaplacianMultigridSolver.jl:4,17,32,104), so that jInv's
# forward-modelling / inversion code dispatches into it (`solveLinearSystem(A, B, param::AbstractSolver, doTranspose)`,
# `copySolver(param)` per worker, `clear!(param)`).  When jInv is installed this module does exactly the same -- the
# import lines below are the reference's own -- and takes RegularMesh / getRegularMesh from jInv.Mesh.  Without jInv
# (stand-alone use, CI) it defines stand-ins with the same names so that the rest of the file is identical.
const HAVE_JINV = Base.find_package("jInv") !== nothing
@static if HAVE_JINV
    using jInv.Mesh                                   # RegularMesh, getRegularMesh        (src/Helmholtz.jl:4)
    using jInv.LinearSolvers                          #                                    (src/Helmholtz.jl:5)
    import jInv.Utils.clear!                          #                                    (src/Helmholtz.jl:6)
    import jInv.LinearSolvers.AbstractSolver          #                                    (src/Helmholtz.jl:7)
    import jInv.LinearSolvers.solveLinearSystem       #                                    (src/Helmholtz.jl:8)
    import jInv.LinearSolvers.solveLinearSystem!      # in-place variant jInv callers use (test/Elastic/...DDElasticHelmholtz.jl:57-58)
    import jInv.LinearSolvers.copySolver              #                                    (src/ShiftedLaplacianMultigridSolver.jl:17)
else
    abstract type AbstractSolver end
    function clear! end
    function solveLinearSystem end
    function solveLinearSystem! end
    function copySolver end
end

export HelmholtzParam, getShiftedHelmholtzParam, GetHelmholtzOperator, GetHelmholtzShiftOP, getABL,
       getMaximalFrequency, getAcousticPointSource, loc2cs, getTopPointSrc, getMidPointSrc,
       MGparam, getMGparam, hierarchyExists, ShiftedLaplacianMultigridSolver,
       getShiftedLaplacianMultigridSolver, copySolver, solveLinearSystem, solveLinearSystem!, clear!,
       RegularMesh, getRegularMesh,
       GetHelmholtzMatrix, GetHelmholtzOperatorHOStencil, setOperatorHO!,
       SlabHandle, SlabHandleNCCL, slabUniqueId, slabPartition, setFrequencyABL!, getGamma

const LIB = get(ENV, "HELMHOLTZ_B200_LIB", joinpath(@__DIR__, "..", "lib", "libhelmholtz_b200.so"))

const HH_OK = 0
const HH_NOT_CONVERGED = 1
const HH_C64, HH_C32 = 0, 1
const HH_MAX_LEVELS = 12

struct HHError <: Exception
    code::Int
    msg::String
end
Base.showerror(io::IO, e::HHError) = print(io, "libhelmholtz_b200 error ", e.code, ": ", e.msg)

function check(rc::Integer, h::Ptr{Cvoid} = C_NULL)
    if rc < 0
        msg = unsafe_string(ccall((:hh_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
        throw(HHError(rc, msg))
    end
    return rc
end

# ---- jInv.Mesh.RegularMesh: only domain / n / h / dim cross the ABI.  With jInv loaded this is jInv's own type. ----
@static if !HAVE_JINV
    struct RegularMesh
        domain::Vector{Float64}
        n::Vector{Int64}
        h::Vector{Float64}
        dim::Int
    end
    getRegularMesh(domain, n) = (d = vec(Float64.(domain)); nn = vec(Int64.(n));
                                 RegularMesh(d, nn, (d[2:2:end] .- d[1:2:end]) ./ nn, length(nn)))
    clear!(M::RegularMesh) = nothing                  # jInv clears the mesh's cached operators; the stand-in has none
end

# ---- src/Helmholtz.jl:13-20 ----
mutable struct HelmholtzParam
    Mesh::RegularMesh
    gamma::Array{Float64}
    m::Array{Float64}
    omega::Union{Float64,ComplexF64}
    NeumannOnTop::Bool
    Sommerfeld::Bool
end
function clear!(HP::HelmholtzParam)                   # src/Helmholtz.jl:25-30
    clear!(HP.Mesh)
    return
end
getShiftedHelmholtzParam(p::HelmholtzParam, s::Float64) =
    HelmholtzParam(p.Mesh, p.gamma .+ s * real(p.omega), p.m, p.omega, p.NeumannOnTop, p.Sommerfeld)

# ---- device handle (HelmholtzParam + device state); freed by a finalizer ----
mutable struct Handle
    ptr::Ptr{Cvoid}
    N::Int
    VAL::DataType
end
function Handle(Mesh, m, omega, gamma, NeumannOnTop::Bool, Sommerfeld::Bool, orderNeumannBC::Int = 2;
                VAL::DataType = ComplexF64, devices::Vector{Int32} = Int32[0])
    nodes = Int64.(Mesh.n .+ 1)
    h = Float64.(Mesh.h)
    mm = vec(Float64.(m)); gg = vec(Float64.(gamma))
    length(mm) == prod(nodes) == length(gg) || error("m and gamma must have prod(n+1) entries")
    out = Ref{Ptr{Cvoid}}(C_NULL)
    w = ComplexF64(omega)
    rc = ccall((:hh_create_multi, LIB), Cint,
               (Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Cint, Cint,
                Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}),
               Mesh.dim, nodes, h, mm, gg, real(w), imag(w), NeumannOnTop, Sommerfeld, orderNeumannBC,
               VAL == ComplexF64 ? HH_C64 : HH_C32, devices, length(devices), out)
    check(rc)
    hd = Handle(out[], prod(nodes), VAL)
    finalizer(x -> (x.ptr != C_NULL && ccall((:hh_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), hd)
    return hd
end

# ---- one grid split into slabs along the last dimension over several GPUs (include/helmholtz_b200.h) ----
# All slabs in this process (one host thread per slab; B, X stay whole-grid arrays): a drop-in for Handle.
function SlabHandle(Mesh, m, omega, gamma, NeumannOnTop::Bool, Sommerfeld::Bool, levels::Int, orderNeumannBC::Int = 2;
                    VAL::DataType = ComplexF64, devices::Vector{Int32} = Int32[0, 1])
    nodes = Int64.(Mesh.n .+ 1)
    h = Float64.(Mesh.h)
    mm = vec(Float64.(m)); gg = vec(Float64.(gamma))
    length(mm) == prod(nodes) == length(gg) || error("m and gamma must have prod(n+1) entries")
    out = Ref{Ptr{Cvoid}}(C_NULL)
    w = ComplexF64(omega)
    rc = ccall((:hh_create_slab_local, LIB), Cint,
               (Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Cint, Cint,
                Ptr{Cint}, Cint, Cint, Ref{Ptr{Cvoid}}),
               Mesh.dim, nodes, h, mm, gg, real(w), imag(w), NeumannOnTop, Sommerfeld, orderNeumannBC,
               VAL == ComplexF64 ? HH_C64 : HH_C32, devices, length(devices), levels, out)
    check(rc)
    hd = Handle(out[], prod(nodes), VAL)
    finalizer(x -> (x.ptr != C_NULL && ccall((:hh_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), hd)
    return hd
end
# One process (Distributed worker) per GPU.  `id` = slabUniqueId() of one worker, sent to the others; m, gamma hold the
# planes plane0+1 : plane0+nplanes of the last dimension (at least koff+1 : koff+nloc of slabPartition).  B and X of
# this worker hold its owned planes (slabPlanes).
function slabUniqueId()
    id = zeros(UInt8, 128)
    check(ccall((:hh_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
function slabPartition(n3::Integer, levels::Integer, nranks::Integer, rank::Integer)
    out = zeros(Int64, 7, levels)  # own0, own1, koff, nloc, zb, ze, n2g per level (0-based planes)
    rc = ccall((:hh_slab_partition, LIB), Cint, (Int64, Cint, Cint, Cint, Ptr{Int64}), n3, levels, nranks, rank, out)
    rc == 0 || error("no slab partition for these sizes")
    return out
end
function SlabHandleNCCL(Mesh, m, omega, gamma, NeumannOnTop::Bool, Sommerfeld::Bool, levels::Int, rank::Int, nranks::Int,
                        id::Vector{UInt8}, plane0::Int, nplanes::Int, orderNeumannBC::Int = 2;
                        VAL::DataType = ComplexF64, device::Int = 0)
    nodes = Int64.(Mesh.n .+ 1)
    h = Float64.(Mesh.h)
    mm = vec(Float64.(m)); gg = vec(Float64.(gamma))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    w = ComplexF64(omega)
    rc = ccall((:hh_create_slab_nccl, LIB), Cint,
               (Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Cint, Cint,
                Cint, Cint, Cint, Cint, Ptr{UInt8}, Int64, Int64, Ref{Ptr{Cvoid}}),
               Mesh.dim, nodes, h, mm, gg, real(w), imag(w), NeumannOnTop, Sommerfeld, orderNeumannBC,
               VAL == ComplexF64 ? HH_C64 : HH_C32, device, levels, rank, nranks, id, plane0, nplanes, out)
    check(rc)
    o0 = Ref{Int64}(0); o1 = Ref{Int64}(0)
    ccall((:hh_slab_info, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}, Ref{Int64}, Ref{Int64}),
          out[], C_NULL, C_NULL, C_NULL, o0, o1)
    hd = Handle(out[], Int(nodes[1] * nodes[2] * (o1[] - o0[])), VAL)  # N = nodes of the owned planes
    finalizer(x -> (x.ptr != C_NULL && ccall((:hh_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), hd)
    return hd
end

# ---- device-side set-up for frequency sweeps: gamma <- gamma_const + getABL(...) on the device, omega replaced ----
function setFrequencyABL!(hd::Handle, omega::Float64, gamma_const::Float64, ABLpad::Array{Int64}, ABLamp::Float64)
    check(ccall((:hh_set_frequency_abl, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Int64}, Cdouble),
                hd.ptr, omega, 0.0, gamma_const, ABLpad, ABLamp), hd.ptr)
end
function getGamma(hd::Handle)
    g = zeros(Float64, hd.N)
    check(ccall((:hh_get_gamma, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), hd.ptr, g), hd.ptr)
    return g
end
function getMaximalFrequency(hd::Handle)
    out = Ref{Cdouble}(0.0)
    check(ccall((:hh_get_maximal_frequency_device, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), hd.ptr, out), hd.ptr)
    return out[]
end

# ---- GetHelmholtzOperatorHO (src/GetHelmholtz.jl:54-72) ----
# the explicit stencil coef[node, s] (s = offset index), ComplexF64; beta as in getSpreadNodalLaplacianAndMass
function GetHelmholtzOperatorHOStencil(Mesh, m, omega, gamma, NeumannOnTop::Bool, Sommerfeld::Bool, beta = 1.0)
    nodes = Int64.(Mesh.n .+ 1)
    bb = Mesh.dim == 3 ? (beta == 1 ? [1.0, 1.0] : Float64.(beta)) : [Float64(beta), Float64(beta)]
    coef = zeros(ComplexF64, prod(nodes), 3^Mesh.dim)
    w = ComplexF64(omega)
    check(ccall((:hh_ho_stencil, LIB), Cint,
                (Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Ptr{Float64}, Ptr{ComplexF64}),
                Mesh.dim, nodes, Float64.(Mesh.h), vec(Float64.(m)), vec(Float64.(gamma)), real(w), imag(w), NeumannOnTop, Sommerfeld,
                bb, coef))
    return coef
end
# make GetHelmholtzOperatorHO the operator of a handle (before the first solve / hh_setup); enable = false goes back
function setOperatorHO!(hd::Handle, m, gamma, beta; enable::Bool = true)
    bb = length(beta) == 2 ? Float64.(beta) : [Float64(beta), Float64(beta)]
    check(ccall((:hh_set_operator_ho, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                hd.ptr, enable, vec(Float64.(m)), vec(Float64.(gamma)), bb))
end

# ---- the explicit SparseMatrixCSC of GetHelmholtzOperator / GetHelmholtzOperatorHO (for H \ q and the like) ----
function GetHelmholtzMatrix(Mesh, m, omega, gamma, NeumannOnTop::Bool, Sommerfeld::Bool, orderNeumannBC::Int = 2;
                            shift::Float64 = 0.0, betaHO = nothing)
    nodes = Int64.(Mesh.n .+ 1)
    N = prod(nodes)
    w = ComplexF64(omega)
    bb = betaHO === nothing ? Ptr{Float64}(C_NULL) : (length(betaHO) == 2 ? Float64.(betaHO) : [Float64(betaHO), Float64(betaHO)])
    colptr = zeros(Int64, N + 1)
    call(rowval, nzval) = check(ccall((:hh_assemble_csc, LIB), Cint,
        (Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Cint, Cdouble, Ptr{Float64},
         Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}),
        Mesh.dim, nodes, Float64.(Mesh.h), vec(Float64.(m)), vec(Float64.(gamma)), real(w), imag(w), NeumannOnTop, Sommerfeld,
        orderNeumannBC, shift, bb, colptr, rowval, nzval))
    call(Ptr{Int64}(C_NULL), Ptr{ComplexF64}(C_NULL))
    nnz = colptr[end]
    rowval = zeros(Int64, nnz); nzval = zeros(ComplexF64, nnz)
    call(rowval, nzval)
    return SparseMatrixCSC(N, N, colptr .+ 1, rowval .+ 1, nzval)   # the library's indices are 0-based
end

# ---- operator objects: matrix-free counterpart of the sparse H (src/GetHelmholtz.jl:33-50) ----
struct HelmholtzShiftOP
    shift::Float64
    omega::Float64
end
GetHelmholtzShiftOP(mNodal::Array{Float64}, omega::Float64, shift::Float64) = HelmholtzShiftOP(shift, omega)

struct HelmholtzOperator
    hd::Handle
    shift::Float64
    adjoint::Bool
end
Base.:+(H::HelmholtzOperator, S::HelmholtzShiftOP) = HelmholtzOperator(H.hd, H.shift + S.shift, H.adjoint)
Base.adjoint(H::HelmholtzOperator) = HelmholtzOperator(H.hd, H.shift, !H.adjoint)
Base.size(H::HelmholtzOperator) = (H.hd.N, H.hd.N)
function Base.:*(H::HelmholtzOperator, x::AbstractVecOrMat)
    X = Array{H.hd.VAL}(reshape(x, H.hd.N, :))
    Y = similar(X)
    check(ccall((:hh_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint, Cdouble, Cint),
                H.hd.ptr, X, Y, size(X, 2), H.shift != 0.0, H.shift, H.adjoint), H.hd.ptr)
    return ndims(x) == 1 ? vec(Y) : Y
end

function getABL(n::Array{Int64}, NeumannAtFirstDim::Bool, ABLpad::Array{Int64}, ABLamp::Float64)
    gamma = zeros(Float64, tuple(n...))
    check(ccall((:hh_get_abl, LIB), Cint, (Cint, Ptr{Int64}, Cint, Ptr{Int64}, Cdouble, Ptr{Float64}),
                length(n), n, NeumannAtFirstDim, ABLpad, ABLamp, gamma))
    return gamma
end

function getMaximalFrequency(m::Union{Array{Float64},Array{Float32},Float64}, M)
    mm = vec(Float64.(m)); out = Ref{Cdouble}(0.0)
    check(ccall((:hh_get_maximal_frequency, LIB), Cint, (Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ref{Cdouble}),
                mm, length(mm), M.dim, Float64.(M.h), out))
    return out[]
end

# the three methods of src/GetHelmholtz.jl:14-16, 22-31, 33-50
GetHelmholtzOperator(Hparam::HelmholtzParam, orderNeumannBC::Int64 = 2) =
    GetHelmholtzOperator(Hparam.Mesh, Hparam.m, Hparam.omega, Hparam.gamma, Hparam.NeumannOnTop, Hparam.Sommerfeld, orderNeumannBC)
function GetHelmholtzOperator(Msh, mNodal::Array{Float64}, omega::Union{Float64,ComplexF64}, gamma::Array,
                              NeumannAtFirstDim::Bool, ABLpad::Array{Int64}, ABLamp::Float64, Sommerfeld::Bool,
                              orderNeumannBC::Int64 = 2)
    abl = getABL(Msh.n .+ 1, NeumannAtFirstDim, ABLpad, ABLamp)
    gamma = isempty(gamma) ? abl : reshape(gamma, size(abl)) .+ abl
    H = GetHelmholtzOperator(Msh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, orderNeumannBC)
    return H, gamma
end
GetHelmholtzOperator(Msh, mNodal::Array{Float64}, omega::Union{Float64,ComplexF64}, gamma::Array{Float64},
                     NeumannAtFirstDim::Bool, Sommerfeld::Bool, orderNeumannBC::Int64 = 2) =
    HelmholtzOperator(Handle(Msh, mNodal, omega, gamma, NeumannAtFirstDim, Sommerfeld, orderNeumannBC), 0.0, false)

# ---- src/getPointSource.jl:63-112 ----
loc2cs(n::Array{Int64}, sub::Array{Int64}) =
    Int(ccall((:hh_point_source_index, LIB), Int64, (Cint, Ptr{Int64}, Ptr{Int64}), length(sub), n, sub))
getTopPointSrc(Minv) = Minv.dim == 3 ? [div(Minv.n[1] + 1, 2); div(Minv.n[2] + 1, 2); 1] : [div(Minv.n[1] + 1, 2); 1]
getMidPointSrc(Minv) = [div(Minv.n[d] + 1, 2) for d in 1:Minv.dim]
function getAcousticPointSource(Minv, TYPE, src = getTopPointSrc(Minv))
    n_nodes = Minv.n .+ 1
    q = zeros(TYPE, tuple(n_nodes...))
    q[loc2cs(n_nodes, src)] = 1.0 ./ (norm(Minv.h)^2)
    return q, src
end

# ---- Multigrid.MGparam: the fields the reference sets / mutates; the hierarchy lives behind `hd` ----
mutable struct MGparam
    VAL::DataType
    levels::Int64
    numCores::Int64
    maxOuterIter::Int64
    relativeTol::Float64
    relaxType::String
    relaxParam::Float64
    relaxPre::Union{Int64,Function}
    relaxPost::Union{Int64,Function}
    cycleType::Char
    coarseSolveType::String
    coarseIters::Int64
    doTranspose::Int64
    hd::Union{Nothing,Handle}
    builtFor::Any
end
getMGparam(VAL::DataType, IND::DataType, levels, numCores, maxIter, relativeTol, relaxType, relaxParam, relaxPre, relaxPost,
           cycleType, coarseSolveType, strongConnParam = 0.5, FilteringParam = 0.0, transferOperatorType = "FullWeighting") =
    MGparam(VAL, levels, numCores, maxIter, relativeTol, relaxType, relaxParam, relaxPre, relaxPost, cycleType,
            coarseSolveType, 10, 0, nothing, nothing)
getMGparam(levels::Int64, args...) = getMGparam(ComplexF64, Int64, levels, args...)
hierarchyExists(MG::MGparam) = MG.hd !== nothing && ccall((:hh_hierarchy_exists, LIB), Cint, (Ptr{Cvoid},), MG.hd.ptr) == 1

struct hh_mg_options
    levels::Int32; relax_type::Int32; cycle_type::Int32; coarse_type::Int32; coarse_iters::Int32; do_transpose::Int32
    relax_pre::NTuple{HH_MAX_LEVELS,Int32}; relax_post::NTuple{HH_MAX_LEVELS,Int32}
    relax_param::Float64; shift::NTuple{HH_MAX_LEVELS,Float64}
end
struct hh_solve_options
    krylov::Int32; inner::Int32; max_iter::Int32; do_transpose::Int32; rel_tol::Float64
end
sweeps(v, l) = v isa Function ? Int32(v(l)) : Int32(v)
function mg_options(MG::MGparam, shift::Vector{Float64}, doTranspose::Int)
    relax = Dict("Jac" => 0, "Jac-GMRES" => 1)[MG.relaxType]
    cyc = Dict('V' => 0, 'W' => 1, 'K' => 2)[MG.cycleType]
    coarse = Dict("NoMUMPS" => 0, "Julia" => 0, "GMRES" => 1)[MG.coarseSolveType]
    hh_mg_options(MG.levels, relax, cyc, coarse, MG.coarseIters, doTranspose,
                  ntuple(l -> sweeps(MG.relaxPre, l), HH_MAX_LEVELS), ntuple(l -> sweeps(MG.relaxPost, l), HH_MAX_LEVELS),
                  MG.relaxParam, ntuple(l -> shift[min(l, length(shift))], HH_MAX_LEVELS))
end

# ---- src/ShiftedLaplacianMultigridSolver.jl:4-30 ----
mutable struct ShiftedLaplacianMultigridSolver <: AbstractSolver   # jInv.LinearSolvers.AbstractSolver when jInv is loaded
    helmParam::HelmholtzParam
    MG::MGparam
    shift::Array{Float64}
    Krylov::String
    inner::Int64
    doClear::Int64
    verbose::Bool
    setupTime::Real
    nPrec::Int
    solveTime::Real
end
getShiftedLaplacianMultigridSolver(helmParam::HelmholtzParam, MG::MGparam, shift::Array{Float64}, Krylov::String = "BiCGSTAB",
                                   inner::Int64 = 5, verbose::Bool = false) =
    ShiftedLaplacianMultigridSolver(helmParam, MG, shift, Krylov, inner, 0, verbose, 0.0, 0, 0.0)
getShiftedLaplacianMultigridSolver(helmParam::HelmholtzParam, MG::MGparam, shift::Float64, Krylov::String = "BiCGSTAB",
                                   inner::Int64 = 5, verbose::Bool = false) =
    getShiftedLaplacianMultigridSolver(helmParam, MG, ones(MG.levels) * shift, Krylov, inner, verbose)

function copySolver(s::ShiftedLaplacianMultigridSolver)  # :18-22 -- settings only, no hierarchy (a method of jInv's copySolver)
    clear!(s.helmParam.Mesh)                             # :20
    MG = s.MG
    MG2 = MGparam(MG.VAL, MG.levels, MG.numCores, MG.maxOuterIter, MG.relativeTol, MG.relaxType, MG.relaxParam, MG.relaxPre,
                  MG.relaxPost, MG.cycleType, MG.coarseSolveType, MG.coarseIters, 0, nothing, nothing)
    return getShiftedLaplacianMultigridSolver(s.helmParam, MG2, s.shift, s.Krylov, s.inner, s.verbose)
end

function clear!(MG::MGparam)
    if MG.hd !== nothing
        ccall((:hh_clear, LIB), Cint, (Ptr{Cvoid},), MG.hd.ptr)
        finalize(MG.hd)
    end
    MG.hd = nothing; MG.builtFor = nothing
end
function clear!(s::ShiftedLaplacianMultigridSolver)  # :105-109 (a method of jInv.Utils.clear!)
    clear!(s.MG)
    clear!(s.helmParam)
    s.doClear = 0
end

function ensureHierarchy(param::ShiftedLaplacianMultigridSolver, doTranspose::Int)
    MG = param.MG; hp = param.helmParam
    sig = (MG.levels, MG.relaxType, MG.relaxParam, [sweeps(MG.relaxPre, l) for l in 1:MG.levels],
           [sweeps(MG.relaxPost, l) for l in 1:MG.levels], MG.cycleType, MG.coarseSolveType, MG.coarseIters, param.shift[1], doTranspose)
    if MG.hd === nothing
        MG.hd = Handle(hp.Mesh, hp.m, hp.omega, hp.gamma, hp.NeumannOnTop, hp.Sommerfeld, 2; VAL = MG.VAL)
    end
    if !hierarchyExists(MG) || MG.builtFor != sig   # MGsetup (:50-66) / transposeHierarchy (:68-70)
        o = Ref(mg_options(MG, vec(param.shift), doTranspose))
        check(ccall((:hh_setup, LIB), Cint, (Ptr{Cvoid}, Ref{hh_mg_options}), MG.hd.ptr, o), MG.hd.ptr)
        MG.builtFor = sig; MG.doTranspose = doTranspose
    end
    return MG.hd
end

# in-place variant: a method of jInv.LinearSolvers.solveLinearSystem!(A, B, X, param::AbstractSolver, doTranspose)
function solveLinearSystem!(ShiftedHT, B, X, param::ShiftedLaplacianMultigridSolver, doTranspose::Int64 = 0)
    param.helmParam.omega isa ComplexF64 && imag(param.helmParam.omega) != 0 &&
        throw(MethodError(GetHelmholtzShiftOP, (param.helmParam.m, param.helmParam.omega, param.shift[1])))  # :77
    if param.doClear == 1
        clear!(param.MG)
    end
    if norm(B) == 0.0                                  # :40-43
        X .= 0
        return X, param
    end
    tt = time_ns()
    hd = ensureHierarchy(param, doTranspose)
    param.setupTime += (time_ns() - tt) / 1e9
    MG = param.MG
    Bm = Array{MG.VAL}(reshape(B, hd.N, :)); nrhs = size(Bm, 2)
    Xm = (X isa Array{MG.VAL} && length(X) == length(Bm)) ? X : similar(Bm)
    iters = zeros(Int32, nrhs); relres = zeros(Float64, nrhs)
    so = Ref(hh_solve_options(param.Krylov == "GMRES" ? 0 : 1, max(param.inner, 1), MG.maxOuterIter, doTranspose, MG.relativeTol))
    tt = time_ns()
    rc = check(ccall((:hh_solve, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{hh_solve_options}, Ptr{Int32}, Ptr{Float64}),
                     hd.ptr, Bm, Xm, nrhs, so, iters, relres), hd.ptr)
    param.solveTime += (time_ns() - tt) / 1e9
    param.nPrec += sum(iters)
    Xm === X || (X .= reshape(Xm, size(X)))
    if rc == HH_NOT_CONVERGED
        println("WARNING: MG solver reached maximum iterations without convergence")   # :97-99
    end
    return X, param
end

# a method of jInv.LinearSolvers.solveLinearSystem (src/ShiftedLaplacianMultigridSolver.jl:32-33): jInv's callers reach it
# by dispatch on the solver type; ShiftedHT (the adjoint of the shifted matrix in the reference) is accepted and unused
function solveLinearSystem(ShiftedHT, B, param::ShiftedLaplacianMultigridSolver, doTranspose::Int64 = 0)
    if size(B, 2) == 1
        B = vec(B)                                     # :34-36
    end
    X = zeros(param.MG.VAL, size(B))
    return solveLinearSystem!(ShiftedHT, B, X, param, doTranspose)
end

end # module
