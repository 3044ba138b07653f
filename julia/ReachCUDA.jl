# ReachCUDA.jl -- the thin `ccall` wrapper a ReachabilityAnalysis.jl maintainer adds so that the
# discrete `post` of LGG09 / GLGM06 runs on libreach_cuda.so (C ABI: include/reach.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary, so this file is an UNTESTED SKETCH until
# it is run under Julia.  What is checked here: every `ccall` names a symbol of include/reach.h, passes as many
# argument types and values as the C prototype has parameters and layout-compatible types (Cint / Int64 / Ptr{Float64} /
# handles / mirrored structs), blocks and brackets balance (tests/test_abi.py).  The same ABI is exercised call for
# call by the Python ctypes binding (reachabilityanalysis.jl_b200/_lib.py) and from plain C (tests/c_consumer).
#
# Intended placement: src/Algorithms/LGG09/reach_cuda.jl and src/Algorithms/GLGM06/reach_cuda.jl,
# loaded through `Requires`/a package extension like the other optional back-ends
# (src/Initialization/init.jl:119-132), selected by a new flag `LGG09(...; device=:cuda)`.
module ReachCUDA

using LazySets, LinearAlgebra
using LazySets: set
using IntervalArithmetic.Symbols: (..)                       # as src/Initialization/init.jl:40, LGG09Module.jl:11
import LazySets: ρ, dim
import ..ReachabilityAnalysis: TemplateReachSet, ReachSet, Flowpipe, MixedFlowpipe, AbstractFlowpipe, TimeInterval, zeroT,
                               _collect, _isdisjoint, support_function_matrix, tspan, array

const lib = "libreach_cuda"

# --- mirror of the C structs -------------------------------------------------------------------
struct ReachTerm
    kind::Int32; which::Int32; q::Int32; p::Int32
    M::Ptr{Float64}; c::Ptr{Float64}; r::Ptr{Float64}; G::Ptr{Float64}
end
struct ReachSetExpr
    nalt::Int32
    alt_len::Ptr{Int32}
    terms::Ptr{ReachTerm}
end
struct ReachLGG09Opts
    time_block::Int32; share_negated::Int32; tile_m::Int32; tile_n::Int32; path::Int32
    reserved::NTuple{3,Int32}
end

struct ReachInvariant
    kind::Int32; m::Int32
    a::Ptr{Float64}; b::Ptr{Float64}
end

mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(device::Integer=0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:reach_create, lib), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
        rc == 0 || error(unsafe_string(ccall((:reach_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(ref[])
        finalizer(c -> ccall((:reach_destroy, lib), Cvoid, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

function check(ctx::Ctx, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:reach_last_error, lib), Cstring, (Ptr{Cvoid},), ctx.h))
    rc == -1 ? throw(ArgumentError(msg)) : error("libreach_cuda ($rc): $msg")
end

# --- flattening of the lazy set tree (the LazySets ρ rules, as data) -----------------------------
# Returns Vector{Vector{NamedTuple}}: alternatives of terms (kind, which, M, c, r, G).
# `T` is the accumulated pre-map (n × q, or `nothing` = identity), `α` the accumulated scalar, `w` = which (0: r_k,
# 1: r_{k+1}).  Same rules as reachabilityanalysis.jl_b200/support_program.py, which is the tested copy.
flatten(X::AbstractHyperrectangle, T, α, w, Φ) = [[(kind=0, which=w, M=T, c=α .* center(X), r=abs(α) .* radius_hyperrectangle(X), G=nothing)]]
flatten(X::BallInf, T, α, w, Φ) = [[(kind=0, which=w, M=T, c=α .* center(X), r=fill(abs(α) * X.radius, dim(X)), G=nothing)]]
flatten(X::Interval, T, α, w, Φ) = [[(kind=0, which=w, M=T, c=[α * (min(X) + max(X)) / 2], r=[abs(α) * (max(X) - min(X)) / 2], G=nothing)]]
flatten(X::AbstractSingleton, T, α, w, Φ) = [[(kind=0, which=w, M=T, c=α .* element(X), r=nothing, G=nothing)]]
flatten(X::ZeroSet, T, α, w, Φ) = [NamedTuple[]]
flatten(X::AbstractZonotope, T, α, w, Φ) = [[(kind=1, which=w, M=T, c=α .* center(X), r=nothing, G=α .* Matrix(genmat(X)))]]
function flatten(X::LinearMap, T, α, w, Φ)
    M = matrix(X)
    if T === nothing && α == 1 && w == 0 && Φ !== nothing && M == Φ
        return flatten(X.X, nothing, 1.0, 1, Φ)                 # evaluate at r_{k+1} = Φᵀ r_k
    end
    return flatten(X.X, T === nothing ? Matrix(M) : T * M, α, w, Φ)
end
flatten(X::ScaledSet, T, α, w, Φ) = flatten(X.X, T, α * X.α, w, Φ)     # δ * U  (ForwardModule.jl:153)
flatten(X::MinkowskiSum, T, α, w, Φ) = [vcat(a, b) for a in flatten(X.X, T, α, w, Φ) for b in flatten(X.Y, T, α, w, Φ)]
flatten(X::ConvexHull, T, α, w, Φ) = vcat(flatten(X.X, T, α, w, Φ), flatten(X.Y, T, α, w, Φ))
function flatten(X::CartesianProduct, T, α, w, Φ)               # ρ(d, X × Y) = ρ(d_X, X) + ρ(d_Y, Y)
    nx, ny = dim(X.X), dim(X.Y)
    base = T === nothing ? Matrix{Float64}(I, nx + ny, nx + ny) : T
    return [vcat(a, b) for a in flatten(X.X, base[:, 1:nx], α, w, Φ) for b in flatten(X.Y, base[:, (nx + 1):end], α, w, Φ)]
end
flatten(X::LazySet, T, α, w, Φ) = throw(ArgumentError("$(typeof(X)) cannot be flattened for the device"))
# Ω₀ may use the r_{k+1} shortcut for `LinearMap(Φ, ·)`; the input set never does (ρ(r_k, V) is evaluated at r_k)
flatten_Ω₀(Ω₀, Φ) = flatten(Ω₀, nothing, 1.0, 0, Φ)
flatten_V(V) = flatten(V, nothing, 1.0, 0, nothing)

# Pack into C structs; `keep` roots every array for the duration of the ccall (GC.@preserve).
function setexpr(alts, keep)
    terms = ReachTerm[]
    for alt in alts, t in alt
        ptr(x) = x === nothing ? Ptr{Float64}(C_NULL) : (a = Array{Float64}(x); push!(keep, a); pointer(a))
        q = t.M === nothing ? 0 : size(t.M, 2)
        p = t.G === nothing ? 0 : size(t.G, 2)
        push!(terms, ReachTerm(t.kind, t.which, q, p, ptr(t.M), ptr(t.c), ptr(t.r), ptr(t.G)))
    end
    lens = Int32[length(a) for a in alts]
    push!(keep, terms, lens)
    return Ref(ReachSetExpr(length(alts), pointer(lens), pointer(terms)))
end

# The invariant `X` of reach_*_LGG09! (reach_homog.jl:382-424, reach_inhomog.jl:174-218) as a `reach_invariant`.
invariant(::Universe, keep) = C_NULL
function invariant(X::HalfSpace, keep)
    a, b = Vector{Float64}(X.a), Float64[X.b]
    push!(keep, a, b)
    return Ref(ReachInvariant(1, 1, pointer(a), pointer(b)))
end
function invariant(X::AbstractHyperrectangle, keep)
    lo, hi = Vector{Float64}(low(X)), Vector{Float64}(high(X))
    push!(keep, lo, hi)
    return Ref(ReachInvariant(2, 0, pointer(lo), pointer(hi)))
end
function invariant(X::AbstractPolyhedron, keep)
    cs = constraints_list(X)
    A = Matrix{Float64}(reduce(hcat, [c.a for c in cs]))        # n × m column-major == m × n row-major, face i at a + i*n
    b = Float64[c.b for c in cs]
    push!(keep, A, b)
    return Ref(ReachInvariant(3, length(cs), pointer(A), pointer(b)))
end

# --- reach_homog_LGG09! / reach_inhomog_LGG09! (LGG09/reach_homog.jl:6-33, reach_inhomog.jl:5-33) ---
# `X` may be Universe, a HalfSpace, a hyperrectangle or a polyhedron; the library stops launching work once a step is
# certified disjoint from `X` and reports the number of valid steps, and `F` is resized like `resize!(F, k - 1)`.
function reach_inhomog_LGG09_cuda!(F::Vector{RT}, dirs, Ω₀::LazySet{N}, Φ::AbstractMatrix{N}, NSTEPS::Integer,
                                   δ::N, X::LazySet, U::Union{LazySet,Nothing}, Δt0::TimeInterval;
                                   ctx::Ctx=Ctx()) where {N,RT<:TemplateReachSet}
    ℓ = Matrix{Float64}(reduce(hcat, _collect(dirs)))          # n × ndirs, column-major, dense
    n, ndirs = size(ℓ)
    ρℓ = Matrix{N}(undef, ndirs, NSTEPS)                       # same layout the reference allocates (:21)
    keep = Any[]
    Φd = Matrix{Float64}(Φ)
    om = setexpr(flatten_Ω₀(Ω₀, Φd), keep)
    v = U === nothing ? C_NULL : setexpr(flatten_V(U), keep)
    inv = invariant(X, keep)
    done = Ref{Int64}(NSTEPS)
    GC.@preserve keep Φd ℓ ρℓ begin
        rc = ccall((:reach_lgg09, lib), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{ReachSetExpr}, Ptr{ReachSetExpr},
                    Ptr{ReachInvariant}, Ptr{ReachLGG09Opts}, Ptr{Float64}, Ptr{Int64}),
                   ctx.h, n, ndirs, NSTEPS, Φd, ℓ, om, v, inv, C_NULL, ρℓ, done)
        check(ctx, rc)
    end
    Δt = (zero(N) .. δ) + Δt0                                  # time stamps stay in Julia (:26-30)
    @inbounds for k in 1:done[]
        F[k] = TemplateReachSet(dirs, view(ρℓ, :, k), Δt)
        Δt += δ
    end
    done[] < NSTEPS && resize!(F, done[])                      # reach_inhomog.jl:213-215
    return ρℓ
end

# the homogeneous twin (LGG09/reach_homog.jl:6-33): same call without an input set
reach_homog_LGG09_cuda!(F, dirs, Ω₀, Φ, NSTEPS, δ, X, Δt0; ctx::Ctx=Ctx()) =
    reach_inhomog_LGG09_cuda!(F, dirs, Ω₀, Φ, NSTEPS, δ, X, nothing, Δt0; ctx=ctx)

# --- device-resident flowpipe: the plan API (Flowpipes/Flowpipe.jl:22-25 "gains device-resident storage") ----------
# The ρℓ matrix stays in HBM; `ρ(d, fp)` over a template direction is one device reduction (AbstractFlowpipe.jl:35-37),
# `support_function_matrix(fp)` fetches lazily (Flowpipe.jl:304-306), `fp[k]` materialises one TemplateReachSet.
mutable struct LGG09Plan
    h::Ptr{Cvoid}
    ctx::Ctx
    function LGG09Plan(ctx::Ctx, n::Integer, ndirs::Integer, NSTEPS::Integer)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx, ccall((:reach_lgg09_plan_create, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{ReachLGG09Opts}, Ref{Ptr{Cvoid}}),
                         ctx.h, n, ndirs, NSTEPS, C_NULL, ref))
        p = new(ref[], ctx)
        finalizer(q -> ccall((:reach_lgg09_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), q.h), p)
        return p
    end
end

mutable struct DeviceTemplateFlowpipe{N,VN,TN} <: AbstractFlowpipe
    plan::LGG09Plan
    dirs::TN
    ℓ::Matrix{Float64}
    nsets::Int
    δ::N
    Δt0::TimeInterval
    sfmat::Union{Nothing,Matrix{N}}          # filled by the first support_function_matrix call
    Xk::Union{Nothing,Vector}                # filled by the first `array(fp)` call
    ext::Dict{Symbol,Any}
end

function post_LGG09_cuda(dirs, Ω₀::LazySet{N}, Φ::AbstractMatrix{N}, NSTEPS::Integer, δ::N, X::LazySet,
                         U::Union{LazySet,Nothing}, Δt0::TimeInterval=zeroT; ctx::Ctx=Ctx()) where {N}
    ℓ = Matrix{Float64}(reduce(hcat, _collect(dirs)))
    n, ndirs = size(ℓ)
    plan = LGG09Plan(ctx, n, ndirs, NSTEPS)
    keep = Any[]
    Φd = Matrix{Float64}(Φ)
    om = setexpr(flatten_Ω₀(Ω₀, Φd), keep)
    v = U === nothing ? C_NULL : setexpr(flatten_V(U), keep)
    inv = invariant(X, keep)
    done = Ref{Int64}(0)
    GC.@preserve keep Φd ℓ begin
        check(ctx, ccall((:reach_lgg09_plan_upload, lib), Cint,
                         (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{ReachSetExpr}, Ptr{ReachSetExpr}, Ptr{ReachInvariant}),
                         plan.h, Φd, ℓ, om, v, inv))
    end
    check(ctx, ccall((:reach_lgg09_plan_run, lib), Cint, (Ptr{Cvoid},), plan.h))
    check(ctx, ccall((:reach_lgg09_plan_sync, lib), Cint, (Ptr{Cvoid},), plan.h))
    check(ctx, ccall((:reach_lgg09_plan_nsteps_done, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}), plan.h, done))
    return DeviceTemplateFlowpipe{N,eltype(dirs),typeof(dirs)}(plan, dirs, ℓ, done[], δ, Δt0, nothing, nothing, Dict{Symbol,Any}())
end

Base.length(fp::DeviceTemplateFlowpipe) = fp.nsets
tspan(fp::DeviceTemplateFlowpipe{N}) where {N} = (zero(N) .. (fp.nsets * fp.δ)) + fp.Δt0

function support_function_matrix(fp::DeviceTemplateFlowpipe{N}) where {N}
    if fp.sfmat === nothing
        M = Matrix{N}(undef, size(fp.ℓ, 2), fp.nsets)
        GC.@preserve M check(fp.plan.ctx, ccall((:reach_lgg09_plan_fetch, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}),
                                                fp.plan.h, 0, fp.nsets, M))
        fp.sfmat = M
    end
    return fp.sfmat
end

function Base.getindex(fp::DeviceTemplateFlowpipe{N}, k::Integer) where {N}
    M = support_function_matrix(fp)
    Δt = (zero(N) .. fp.δ) + fp.Δt0
    for _ in 2:k
        Δt += fp.δ                                              # repeated addition, as the fill loop does
    end
    return TemplateReachSet(fp.dirs, view(M, :, k), Δt)
end

# The generic AbstractFlowpipe interface (iterate, first / last, eachindex, set, tstart / tend, vars, dim, σ:
# AbstractFlowpipe.jl:56-216) goes through `array(fp)`: the reach-sets are materialised once, on demand, as views
# into the fetched matrix with the time stamps of the fill loop (reach_homog.jl:26-30).
function array(fp::DeviceTemplateFlowpipe{N}) where {N}
    if fp.Xk === nothing
        M = support_function_matrix(fp)
        Δt = (zero(N) .. fp.δ) + fp.Δt0
        Xk = Vector{TemplateReachSet}(undef, fp.nsets)
        for k in 1:fp.nsets
            Xk[k] = TemplateReachSet(fp.dirs, view(M, :, k), Δt)
            Δt += fp.δ
        end
        fp.Xk = Xk
    end
    return fp.Xk
end

# ρ(d, fp) = max_k ρ(d, F[k]) (AbstractFlowpipe.jl:35-37); TemplateReachSet.jl:97-115: exact template direction, then a
# positive multiple of one, else the LP over the template polyhedron (host, on the fetched matrix)
function ρ(d::AbstractVector, fp::DeviceTemplateFlowpipe{N}) where {N}
    for (i, ℓi) in enumerate(eachcol(fp.ℓ))
        k = samedir_factor(d, ℓi)
        k === nothing && continue
        vmax = Vector{Float64}(undef, size(fp.ℓ, 2))
        GC.@preserve vmax check(fp.plan.ctx, ccall((:reach_lgg09_plan_max_over_steps, lib), Cint,
                                                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}), fp.plan.h, vmax, C_NULL))
        return vmax[i] * k
    end
    return maximum(ρ(d, set(fp[k])) for k in 1:length(fp))
end
function samedir_factor(d, ℓ)
    j = findfirst(!iszero, ℓ)
    j === nothing && return nothing
    k = d[j] / ℓ[j]
    return (k > 0 && isapprox(d, k .* ℓ; rtol=1e-12, atol=0)) ? k : nothing
end

# --- the split caller `_solve_distributed` (Continuous/solve.jl:84-117) for LGG09 -------------------------------
# Ω₀s: the discretised initial sets of the splits (same Φ, U and template).  One device batch instead of one `post`
# per split; returns the MixedFlowpipe the reference builds at solve.jl:116 plus the ρℓ matrices.
function solve_distributed_LGG09_cuda(dirs, Ω₀s::Vector{<:LazySet{N}}, Φ::AbstractMatrix{N}, NSTEPS::Integer, δ::N,
                                      U::Union{LazySet,Nothing}, Δt0::TimeInterval; ctx::Ctx=Ctx()) where {N}
    ℓ = Matrix{Float64}(reduce(hcat, _collect(dirs)))
    n, ndirs = size(ℓ)
    m = length(Ω₀s)
    keep = Any[]
    Φd = Matrix{Float64}(Φ)
    oms = ReachSetExpr[setexpr(flatten_Ω₀(Ω₀, Φd), keep)[] for Ω₀ in Ω₀s]
    v = U === nothing ? C_NULL : setexpr(flatten_V(U), keep)
    ρs = Array{N}(undef, ndirs, NSTEPS, m)                      # split i: ρs[:, :, i], the reference's ρℓ layout
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep Φd ℓ oms ρs begin
        check(ctx, ccall((:reach_lgg09_batch_create, lib), Cint,
                         (Ptr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{ReachSetExpr},
                          Ptr{ReachSetExpr}, Ref{Ptr{Cvoid}}), ctx.h, n, ndirs, NSTEPS, m, Φd, ℓ, oms, v, h))
        try
            check(ctx, ccall((:reach_lgg09_batch_run, lib), Cint, (Ptr{Cvoid},), h[]))
            for i in 1:m
                check(ctx, ccall((:reach_lgg09_batch_fetch, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}),
                                 h[], i - 1, 0, NSTEPS, pointer(ρs, (i - 1) * ndirs * NSTEPS + 1)))
            end
        finally
            ccall((:reach_lgg09_batch_destroy, lib), Cvoid, (Ptr{Cvoid},), h[])
        end
    end
    RT = TemplateReachSet{N,eltype(dirs),typeof(dirs),typeof(view(ρs, :, 1, 1))}
    fps = Vector{Flowpipe{N,RT,Vector{RT}}}(undef, m)
    for i in 1:m
        F = Vector{RT}(undef, NSTEPS)
        Δt = (zero(N) .. δ) + Δt0
        @inbounds for k in 1:NSTEPS
            F[k] = TemplateReachSet(dirs, view(ρs, :, k, i), Δt)
            Δt += δ
        end
        fps[i] = Flowpipe(F, Dict{Symbol,Any}(:sfmat => view(ρs, :, :, i)))
    end
    return MixedFlowpipe(fps)
end

# --- reach_inhomog_GLGM06! (GLGM06/reach_inhomog.jl:6-43) -----------------------------------------
function reach_inhomog_GLGM06_cuda!(F::Vector{ReachSet{N,Zonotope{N,Vector{N},Matrix{N}}}}, Ω0::Zonotope{N},
                                    Φ::AbstractMatrix, NSTEPS::Integer, δ::Float64, max_order::Integer, ::Universe,
                                    U::Union{Zonotope,Nothing}, Δt0::TimeInterval; ctx::Ctx=Ctx()) where {N}
    n, p0 = size(genmat(Ω0))
    gcap = max(p0, max_order * n)
    C = Matrix{N}(undef, n, NSTEPS)
    G = Array{N}(undef, n, gcap, NSTEPS)
    ng = Vector{Int32}(undef, NSTEPS)
    Φd, c0, G0 = Matrix{Float64}(Φ), Vector{Float64}(center(Ω0)), Matrix{Float64}(genmat(Ω0))
    cU = U === nothing ? C_NULL : Vector{Float64}(center(U))
    GU = U === nothing ? C_NULL : Matrix{Float64}(genmat(U))
    pU = U === nothing ? -1 : size(genmat(U), 2)
    GC.@preserve Φd c0 G0 cU GU C G ng begin
        rc = ccall((:reach_glgm06, lib), Cint,
                   (Ptr{Cvoid}, Cint, Int64, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64},
                    Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Int32}),
                   ctx.h, n, NSTEPS, 1, max_order, Φd, c0, G0, p0, cU, GU, pU, C, G, gcap, ng)
        check(ctx, rc)
    end
    Δt = (zero(N) .. δ) + Δt0
    @inbounds for k in 1:NSTEPS
        F[k] = ReachSet(Zonotope(C[:, k], G[:, 1:ng[k], k]), Δt)
        Δt += δ
    end
    return F
end

# --- device-resident zonotope flowpipe (reach_glgm06_plan_*): nothing is brought to the host until asked for ---------
mutable struct GLGM06Plan
    h::Ptr{Cvoid}
    ctx::Ctx
    n::Int; nsteps::Int; nsplits::Int; gcap::Int
end

function post_GLGM06_cuda(Ω0s::Vector{<:Zonotope{N}}, Φ::AbstractMatrix, NSTEPS::Integer, max_order::Integer,
                          U::Union{Zonotope,Nothing}; ctx::Ctx=Ctx()) where {N}
    n, p0 = size(genmat(Ω0s[1]))
    m = length(Ω0s)
    pU = U === nothing ? -1 : size(genmat(U), 2)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:reach_glgm06_plan_create, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                     ctx.h, n, NSTEPS, m, max_order, p0, pU, ref))
    plan = GLGM06Plan(ref[], ctx, n, NSTEPS, m, max(p0, max_order * n))
    finalizer(q -> ccall((:reach_glgm06_plan_destroy, lib), Cvoid, (Ptr{Cvoid},), q.h), plan)
    Φd = Matrix{Float64}(Φ)
    c0 = reduce(hcat, [Vector{Float64}(center(Z)) for Z in Ω0s])            # n × nsplits
    G0 = reduce(hcat, [Matrix{Float64}(genmat(Z)) for Z in Ω0s])            # n × (p0·nsplits) == n × p0 × nsplits
    cU = U === nothing ? C_NULL : Vector{Float64}(center(U))
    GU = U === nothing ? C_NULL : Matrix{Float64}(genmat(U))
    GC.@preserve Φd c0 G0 cU GU begin
        check(ctx, ccall((:reach_glgm06_plan_upload, lib), Cint,
                         (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                         plan.h, Φd, c0, G0, cU, GU))
    end
    check(ctx, ccall((:reach_glgm06_plan_run, lib), Cint, (Ptr{Cvoid},), plan.h))
    check(ctx, ccall((:reach_glgm06_plan_sync, lib), Cint, (Ptr{Cvoid},), plan.h))
    return plan
end

# one zonotope of the flowpipe, exactly as the reference would hold it ([kept generators in sorted order | diagonal])
function Base.getindex(plan::GLGM06Plan, split::Integer, step::Integer)
    ng = Ref{Cint}(0)
    check(plan.ctx, ccall((:reach_glgm06_plan_ngens, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cint}), plan.h, split - 1, step - 1, ng))
    c = Vector{Float64}(undef, plan.n)
    G = Matrix{Float64}(undef, plan.n, ng[])
    GC.@preserve c G check(plan.ctx, ccall((:reach_glgm06_plan_fetch, lib), Cint,
                                           (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                                           plan.h, split - 1, step - 1, c, G, C_NULL))
    return Zonotope(c, G)
end

# ρ(d, Z_{split,step}) for every split and step on the device; ρ(d, MixedFlowpipe) is `maximum` of the result
function ρ(d::AbstractVector, plan::GLGM06Plan)
    out = Matrix{Float64}(undef, plan.nsplits, plan.nsteps)
    dd = Vector{Float64}(d)
    GC.@preserve dd out check(plan.ctx, ccall((:reach_glgm06_plan_support, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                                              plan.h, dd, out))
    return out
end

# --- several GPUs in one Julia process (reach_multi_*): directions sharded inside the library ------------------------
mutable struct MultiCtx
    h::Ptr{Cvoid}
    function MultiCtx(devices::Vector{<:Integer})
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        ids = Vector{Cint}(devices)
        rc = ccall((:reach_multi_create, lib), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), ref, ids, length(ids))
        rc == 0 || error(unsafe_string(ccall((:reach_multi_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(ref[])
        finalizer(x -> ccall((:reach_multi_destroy, lib), Cvoid, (Ptr{Cvoid},), x.h), m)
        return m
    end
end

function reach_inhomog_LGG09_cuda!(F::Vector{RT}, dirs, Ω₀::LazySet{N}, Φ::AbstractMatrix{N}, NSTEPS::Integer, δ::N,
                                   ::Universe, U::Union{LazySet,Nothing}, Δt0::TimeInterval, multi::MultiCtx) where {N,RT<:TemplateReachSet}
    ℓ = Matrix{Float64}(reduce(hcat, _collect(dirs)))
    n, ndirs = size(ℓ)
    ρℓ = Matrix{N}(undef, ndirs, NSTEPS)
    keep = Any[]
    Φd = Matrix{Float64}(Φ)
    om = setexpr(flatten_Ω₀(Ω₀, Φd), keep)
    v = U === nothing ? C_NULL : setexpr(flatten_V(U), keep)
    GC.@preserve keep Φd ℓ ρℓ begin
        rc = ccall((:reach_multi_lgg09, lib), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{ReachSetExpr}, Ptr{ReachSetExpr},
                    Ptr{ReachLGG09Opts}, Ptr{Float64}, Ptr{Ptr{Cvoid}}, Ptr{Float64}),
                   multi.h, n, ndirs, NSTEPS, Φd, ℓ, om, v, C_NULL, ρℓ, C_NULL, C_NULL)
        rc == 0 || error(unsafe_string(ccall((:reach_multi_last_error, lib), Cstring, (Ptr{Cvoid},), multi.h)))
    end
    Δt = (zero(N) .. δ) + Δt0
    @inbounds for k in 1:NSTEPS
        F[k] = TemplateReachSet(dirs, view(ρℓ, :, k), Δt)
        Δt += δ
    end
    return ρℓ
end

# --- reach_homog_BOX! / reach_inhomog_BOX! (BOX/reach_homog.jl:8-47,50-69, reach_inhomog.jl:8-60) -----------
# Ω0 and U are already hyperrectangles here (BOX/post.jl:29,40).  `recursive` only exists without inputs.
function reach_BOX_cuda!(F::Vector{ReachSet{N,Hyperrectangle{N,Vector{N},Vector{N}}}}, Ω0::Hyperrectangle{N},
                         Φ::AbstractMatrix, NSTEPS::Integer, δ::N, X::LazySet, U::Union{Hyperrectangle,Nothing},
                         recursive::Bool, Δt0::TimeInterval; ctx::Ctx=Ctx()) where {N}
    n = dim(Ω0)
    C = Matrix{N}(undef, n, NSTEPS)
    R = Matrix{N}(undef, n, NSTEPS)
    Φd, c0, r0 = Matrix{Float64}(Φ), Vector{Float64}(center(Ω0)), Vector{Float64}(radius_hyperrectangle(Ω0))
    cU = U === nothing ? C_NULL : Vector{Float64}(center(U))
    rU = U === nothing ? C_NULL : Vector{Float64}(radius_hyperrectangle(U))
    GC.@preserve Φd c0 r0 cU rU C R begin
        rc = ccall((:reach_box, lib), Cint,
                   (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint,
                    Ptr{ReachLGG09Opts}, Ptr{Float64}, Ptr{Float64}),
                   ctx.h, n, NSTEPS, Φd, c0, r0, cU, rU, recursive, C_NULL, C, R)
        check(ctx, rc)
    end
    Δt = (zero(N) .. δ) + Δt0
    F[1] = ReachSet(Ω0, Δt)
    k = 2
    @inbounds while k <= NSTEPS
        Hₖ = Hyperrectangle(C[:, k], R[:, k]; check_bounds=false)
        !(X isa Universe) && _isdisjoint(X, Hₖ) && break      # the invariant check stays in Julia (reach_homog.jl:101)
        Δt += δ
        F[k] = ReachSet(Hₖ, Δt)
        k += 1
    end
    k < NSTEPS + 1 && resize!(F, k - 1)
    return F
end

end # module
