# ReachCUDA.jl -- the thin `ccall` wrapper a ReachabilityAnalysis.jl maintainer adds so that the
# discrete `post` of LGG09 / GLGM06 runs on libreach_cuda.so (C ABI: include/reach.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia` binary.  The same ABI is
# exercised by the Python ctypes binding (reachabilityanalysis.jl_b200/_lib.py), call for call.
#
# Intended placement: src/Algorithms/LGG09/reach_cuda.jl and src/Algorithms/GLGM06/reach_cuda.jl,
# loaded through `Requires`/a package extension like the other optional back-ends
# (src/Initialization/init.jl:119-132), selected by a new flag `LGG09(...; device=:cuda)`.
module ReachCUDA

using LazySets, LinearAlgebra
import ..ReachabilityAnalysis: TemplateReachSet, ReachSet, Flowpipe, TimeInterval, zeroT, _collect

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
flatten(X::AbstractHyperrectangle, T, α, w, Φ) = [[(kind=0, which=w, M=T, c=α .* center(X), r=abs(α) .* radius_hyperrectangle(X), G=nothing)]]
flatten(X::AbstractSingleton, T, α, w, Φ) = [[(kind=0, which=w, M=T, c=α .* element(X), r=nothing, G=nothing)]]
flatten(X::ZeroSet, T, α, w, Φ) = [NamedTuple[]]
flatten(X::AbstractZonotope, T, α, w, Φ) = [[(kind=1, which=w, M=T, c=α .* center(X), r=nothing, G=α .* Matrix(genmat(X)))]]
function flatten(X::LinearMap, T, α, w, Φ)
    M = matrix(X)
    if T === nothing && α == 1 && w == 0 && M == Φ
        return flatten(X.X, nothing, 1.0, 1, Φ)                 # evaluate at r_{k+1} = Φᵀ r_k
    end
    return flatten(X.X, T === nothing ? Matrix(M) : T * M, α, w, Φ)
end
flatten(X::ScaledSet, T, α, w, Φ) = flatten(X.X, T, α * X.α, w, Φ)     # δ * U  (ForwardModule.jl:153)
flatten(X::MinkowskiSum, T, α, w, Φ) = [vcat(a, b) for a in flatten(X.X, T, α, w, Φ) for b in flatten(X.Y, T, α, w, Φ)]
flatten(X::ConvexHull, T, α, w, Φ) = vcat(flatten(X.X, T, α, w, Φ), flatten(X.Y, T, α, w, Φ))
flatten(X::LazySet, T, α, w, Φ) = throw(ArgumentError("$(typeof(X)) cannot be flattened for the device"))

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

# --- reach_homog_LGG09! / reach_inhomog_LGG09! (LGG09/reach_homog.jl:6-33, reach_inhomog.jl:5-33) ---
function reach_inhomog_LGG09_cuda!(F::Vector{RT}, dirs, Ω₀::LazySet{N}, Φ::AbstractMatrix{N}, NSTEPS::Integer,
                                   δ::N, ::Universe, U::Union{LazySet,Nothing}, Δt0::TimeInterval;
                                   ctx::Ctx=Ctx()) where {N,RT<:TemplateReachSet}
    ℓ = reduce(hcat, _collect(dirs))                           # n × ndirs, column-major
    n, ndirs = size(ℓ)
    ρℓ = Matrix{N}(undef, ndirs, NSTEPS)                       # same layout the reference allocates (:21)
    keep = Any[]
    Φd = Matrix{Float64}(Φ)
    om = setexpr(flatten(Ω₀, nothing, 1.0, 0, Φd), keep)
    v = U === nothing ? C_NULL : setexpr(flatten(U, nothing, 1.0, 0, Φd), keep)
    GC.@preserve keep Φd ℓ ρℓ begin
        rc = ccall((:reach_lgg09, lib), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{ReachSetExpr}, Ptr{ReachSetExpr},
                    Ptr{Cvoid}, Ptr{ReachLGG09Opts}, Ptr{Float64}, Ptr{Int64}),
                   ctx.h, n, ndirs, NSTEPS, Φd, ℓ, om, v, C_NULL, C_NULL, ρℓ, C_NULL)
        check(ctx, rc)
    end
    Δt = (zero(N) .. δ) + Δt0                                  # time stamps stay in Julia (:26-30)
    @inbounds for k in 1:NSTEPS
        F[k] = TemplateReachSet(dirs, view(ρℓ, :, k), Δt)
        Δt += δ
    end
    return ρℓ
end

# --- the split caller `_solve_distributed` (Continuous/solve.jl:84-117) for LGG09 -------------------------------
# Ω₀s: the discretised initial sets of the splits (same Φ, U and template).  One device batch instead of one `post`
# per split; returns the MixedFlowpipe the reference builds at solve.jl:116 plus the ρℓ matrices.
function solve_distributed_LGG09_cuda(dirs, Ω₀s::Vector{<:LazySet{N}}, Φ::AbstractMatrix{N}, NSTEPS::Integer, δ::N,
                                      U::Union{LazySet,Nothing}, Δt0::TimeInterval; ctx::Ctx=Ctx()) where {N}
    ℓ = reduce(hcat, _collect(dirs))
    n, ndirs = size(ℓ)
    m = length(Ω₀s)
    keep = Any[]
    Φd = Matrix{Float64}(Φ)
    oms = ReachSetExpr[setexpr(flatten(Ω₀, nothing, 1.0, 0, Φd), keep)[] for Ω₀ in Ω₀s]
    v = U === nothing ? C_NULL : setexpr(flatten(U, nothing, 1.0, 0, Φd), keep)
    ρ = Array{N}(undef, ndirs, NSTEPS, m)                      # split i: ρ[:, :, i], the reference's ρℓ layout
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep Φd ℓ oms ρ begin
        check(ctx, ccall((:reach_lgg09_batch_create, lib), Cint,
                         (Ptr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{ReachSetExpr},
                          Ptr{ReachSetExpr}, Ref{Ptr{Cvoid}}), ctx.h, n, ndirs, NSTEPS, m, Φd, ℓ, oms, v, h))
        try
            check(ctx, ccall((:reach_lgg09_batch_run, lib), Cint, (Ptr{Cvoid},), h[]))
            for i in 1:m
                check(ctx, ccall((:reach_lgg09_batch_fetch, lib), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Float64}),
                                 h[], i - 1, 0, NSTEPS, pointer(ρ, (i - 1) * ndirs * NSTEPS + 1)))
            end
        finally
            ccall((:reach_lgg09_batch_destroy, lib), Cvoid, (Ptr{Cvoid},), h[])
        end
    end
    RT = TemplateReachSet{N,eltype(dirs),typeof(dirs),typeof(view(ρ, :, 1, 1))}
    fps = Vector{Flowpipe{N,RT,Vector{RT}}}(undef, m)
    for i in 1:m
        F = Vector{RT}(undef, NSTEPS)
        Δt = (zero(N) .. δ) + Δt0
        @inbounds for k in 1:NSTEPS
            F[k] = TemplateReachSet(dirs, view(ρ, :, k, i), Δt)
            Δt += δ
        end
        fps[i] = Flowpipe(F, Dict{Symbol,Any}(:sfmat => view(ρ, :, :, i)))
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
