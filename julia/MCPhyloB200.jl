# MCPhyloB200.jl — the reference-side binding a maintainer adds to MCPhylo.jl so that
#   logpdf(d::PhyloDist, x), gradlogpdf(d::PhyloDist, x)            (src/distributions/Phylodist.jl:107-138)
#   logpdf(d::MultiplePhyloDist, x), __logpdf(d::MultiplePhyloDist, x)   (Phylodist.jl:281-297)
# run on a B200 through libmcphylo_b200.so (include/mcphylo_b200.h).  Samplers, the model graph
# and everything else keep calling the same generic functions.
#
# NOT EXECUTED IN THIS REPO'S CI: the build container has no Julia.  The Python ctypes binding
# mcphylo.jl_b200/capi.py makes exactly these calls with exactly these argument orders and IS
# tested on the GPU (tests/test_gpu_parity.py), see INTEGRATION.md.
#
# Usage:  using MCPhylo; include("MCPhyloB200.jl"); MCPhyloB200.enable!("/path/to/libmcphylo_b200.so")
#         ENV["MCPHYLO_B200_DEVICES"] = "0,1,2,3,4,5,6,7" before the first evaluation shards every alignment
#         over the 8 GPUs of the box (one context, mcp_create_multi); nothing else changes.
module MCPhyloB200

using MCPhylo
using LinearAlgebra: Diagonal
import MCPhylo: PhyloDist, MultiplePhyloDist, logpdf, gradlogpdf, __logpdf,
                post_order, get_leaves, get_branchlength_vector, get_mother

const LIB = Ref{String}("libmcphylo_b200")
# The handle is process-local and created lazily; it is never stored in a PhyloDist / Model, so
# serialised chains (src/output/fileio.jl:15-35) stay loadable.
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
# One resident device alignment per data array OBJECT (the observed node's value is the same Array for
# the life of a chain, src/model/dependent.jl:344-358).  Keyed by objectid and NOT holding the array:
# `mcmc` deep-copies the model -- and with it the data -- on every call (src/model/mcmc.jl:115), so a
# table with strong references would pin one more host array and one more device alignment per run.
# A finalizer on the array queues its device alignment for destruction; the queue is drained at the
# start of the next evaluation (finalizers must not re-enter the library in the middle of a call).
const ALIGNMENTS = Dict{UInt,Ptr{Cvoid}}()
const SLABS = Dict{Tuple{UInt,Int},Array{Float64,3}}()      # MultiplePhyloDist: per-tree slabs of a 4-d array
const RELEASED = Ptr{Cvoid}[]
const RELEASED_LOCK = Threads.SpinLock()

last_error(ctx) = unsafe_string(ccall((:mcp_last_error, LIB[]), Cstring, (Ptr{Cvoid},), ctx))
check(rc::Cint, ctx = CTX[]) = rc == 0 ? nothing : error("libmcphylo_b200 ($rc): " * last_error(ctx))

"""
    context()

The process-wide library context, created on first use.  One GPU: `MCPHYLO_B200_DEVICE=3` (default 0).
Several GPUs of the box behind the SAME calls: `MCPHYLO_B200_DEVICES=0,1,2,3,4,5,6,7` -- the alignment's
site axis is split across them and every `logpdf` / `gradlogpdf` ends in one all-reduce of
`[logL, gradient]` (`mcp_create_multi`; `MCPHYLO_B200_REDUCE` = auto | nccl | peer | host).
"""
function context()
    if CTX[] == C_NULL
        out = Ref{Ptr{Cvoid}}(C_NULL)
        if haskey(ENV, "MCPHYLO_B200_DEVICES")
            ids = Cint[parse(Cint, strip(t)) for t in split(ENV["MCPHYLO_B200_DEVICES"], ",")]
            mode = Cint(findfirst(==(lowercase(get(ENV, "MCPHYLO_B200_REDUCE", "auto"))), ["auto", "nccl", "peer", "host"]) - 1)
            rc = ccall((:mcp_create_multi, LIB[]), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}, Cint), out, length(ids), ids, mode)
        else
            rc = ccall((:mcp_create, LIB[]), Cint, (Ref{Ptr{Cvoid}}, Cint), out, parse(Cint, get(ENV, "MCPHYLO_B200_DEVICE", "0")))
        end
        rc == 0 || error("libmcphylo_b200 ($rc): " * last_error(C_NULL))   # no CPU fallback
        CTX[] = out[]
        # no per-evaluation timing events (they only feed mcp_get_stats): 5-13 us of a 36-58 us MCMC-sized call
        get(ENV, "MCPHYLO_B200_TIMING", "0") == "1" || ccall((:mcp_set_timing, LIB[]), Cint, (Ptr{Cvoid}, Cint), CTX[], 0)
    end
    CTX[]
end

# device alignments whose host arrays have been collected
function drain_released()
    isempty(RELEASED) && return
    lock(RELEASED_LOCK)
    handles = copy(RELEASED); empty!(RELEASED)
    unlock(RELEASED_LOCK)
    CTX[] == C_NULL && return
    for h in handles
        ccall((:mcp_alignment_destroy, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], h)
    end
end

function release_later(x)
    key = objectid(x)
    h = pop!(ALIGNMENTS, key, C_NULL)
    for k in [k for k in keys(SLABS) if k[1] == key]       # slabs cut from a 4-d array die with it
        release_later(pop!(SLABS, k))
    end
    h == C_NULL && return
    lock(RELEASED_LOCK); push!(RELEASED, h); unlock(RELEASED_LOCK)
end

function alignment(x::Array{Float64,3}, leaf_nums::Vector{Int32})
    drain_released()
    h = get(ALIGNMENTS, objectid(x), C_NULL)
    h != C_NULL && return h
    K, S, NN = size(x)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mcp_alignment_from_dense, LIB[]), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Cint, Int64, Cint, Ptr{Int32}, Cint, Ref{Ptr{Cvoid}}),
                context(), x, K, S, NN, leaf_nums, length(leaf_nums), out))
    ALIGNMENTS[objectid(x)] = out[]
    finalizer(release_later, x)
    out[]
end

# Tree -> flat arrays, using MCPhyloTree's own accessors so the numbering rule is never re-derived.
function flatten(tree)
    po = post_order(tree)
    NN = length(po)
    postorder_num = Int32[n.num for n in po]
    parent_num = zeros(Int32, NN)
    for n in po
        n.root || (parent_num[n.num] = get_mother(n).num)
    end
    leaf_nums = Int32[l.num for l in get_leaves(tree)]
    NN, postorder_num, parent_num, get_branchlength_vector(tree), leaf_nums
end

function evaluate(d::PhyloDist, x::Array{Float64,3}, want_grad::Bool)
    NN, po, pa, blv, leaf_nums = flatten(d.tree)
    U, D, Uinv, mu = d.substitution_model(d.base_freq, d.substitution_rates)
    ll = Ref{Float64}(0.0)
    grad = want_grad ? Vector{Float64}(undef, NN - 1) : Float64[]
    check(ccall((:mcp_eval, LIB[]), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Cint, Ptr{Float64},
                 Cint, Ref{Float64}, Ptr{Float64}),
                context(), alignment(x, leaf_nums), NN, po, pa, Vector{Float64}(blv),
                Matrix{Float64}(U), Vector{Float64}(D), Matrix{Float64}(Uinv), Float64(mu),
                d.rates, length(d.rates), d.base_freq,
                want_grad, ll, want_grad ? pointer(grad) : C_NULL))
    ll[], grad
end

# (logL, d logL / d branch length, d logL / d rates[r]) in one call (mcp_eval_rate_gradient).  The reference samples the
# Gamma shape behind `rates` gradient-free (src/Likelihood/Rates.jl:11-38); with this a gradient-based sampler can take
# alpha along:  d logL / d alpha = dot(rate_grad, d discrete_gamma_rates(alpha, alpha, k) / d alpha).
function rate_gradient(d::PhyloDist, x::Array{Float64,3})
    NN, po, pa, blv, leaf_nums = flatten(d.tree)
    U, D, Uinv, mu = d.substitution_model(d.base_freq, d.substitution_rates)
    ll = Ref{Float64}(0.0)
    grad = Vector{Float64}(undef, NN - 1)
    rgrad = Vector{Float64}(undef, length(d.rates))
    check(ccall((:mcp_eval_rate_gradient, LIB[]), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Cint, Ptr{Float64},
                 Ref{Float64}, Ptr{Float64}, Ptr{Float64}),
                context(), alignment(x, leaf_nums), NN, po, pa, Vector{Float64}(blv),
                Matrix{Float64}(U), Vector{Float64}(D), Matrix{Float64}(Uinv), Float64(mu),
                d.rates, length(d.rates), d.base_freq, ll, grad, rgrad))
    ll[], grad, rgrad
end

# Derivatives of the normalised rate matrix A = mu * U * Diagonal(D) * Uinv with respect to theta =
# (base_freq..., substitution_rates...), as a K x K x n_par array, for ANY substitution_model function: central
# differences of the K x K matrix, Richardson-extrapolated twice (O(h^6); mcphylo.jl_b200/substitution_models.py holds
# the closed forms for Restriction / JC / GTR / freeK and checks them against exactly this quotient).
function rate_matrix_derivatives(model, base_freq::Vector{Float64}, subst::Vector{Float64})
    K = length(base_freq)
    theta = vcat(base_freq, subst)
    A(th) = begin
        U, D, Uinv, mu = model(th[1:K], th[K+1:end])
        mu .* (Matrix{Float64}(U) * Diagonal(Vector{Float64}(D)) * Matrix{Float64}(Uinv))
    end
    dA = zeros(K, K, length(theta))
    for p in eachindex(theta)
        h = 1e-3 * max(abs(theta[p]), 1e-2)
        cd(s) = begin
            tp = copy(theta); tm = copy(theta); tp[p] += s; tm[p] -= s
            (A(tp) .- A(tm)) ./ (2s)
        end
        d1, d2, d3 = cd(h), cd(h / 2), cd(h / 4)
        r1 = (4 .* d2 .- d1) ./ 3; r2 = (4 .* d3 .- d2) ./ 3
        dA[:, :, p] = (16 .* r2 .- r1) ./ 15
    end
    dA
end

# (logL, d logL / d branch length, d logL / d base_freq, d logL / d substitution_rates, d logL / d rates) in one call
# (mcp_eval_model_gradient): the gradient pass accumulates per-branch moment matrices, the library contracts them
# with d P / d theta.  The reference samples these parameters gradient-free (SliceSimplex(:mypi),
# src/samplers/tree_samplers.jl:49); base_freq entries are independent coordinates here (root term + rate matrix).
function model_gradient(d::PhyloDist, x::Array{Float64,3})
    NN, po, pa, blv, leaf_nums = flatten(d.tree)
    U, D, Uinv, mu = d.substitution_model(d.base_freq, d.substitution_rates)
    K = length(d.base_freq)
    subst = Vector{Float64}(vec(d.substitution_rates))
    dA = rate_matrix_derivatives(d.substitution_model, Vector{Float64}(d.base_freq), subst)
    n_par = size(dA, 3)
    dpi = zeros(K, n_par); for s in 1:K; dpi[s, s] = 1.0; end
    ll = Ref{Float64}(0.0)
    grad = Vector{Float64}(undef, NN - 1)
    pgrad = Vector{Float64}(undef, n_par)
    rgrad = Vector{Float64}(undef, length(d.rates))          # d logL / d rates[r], from the same moments
    check(ccall((:mcp_eval_model_gradient, LIB[]), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Cint, Ptr{Float64},
                 Cint, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                context(), alignment(x, leaf_nums), NN, po, pa, Vector{Float64}(blv),
                Matrix{Float64}(U), Vector{Float64}(D), Matrix{Float64}(Uinv), Float64(mu),
                d.rates, length(d.rates), d.base_freq, n_par, dA, dpi, ll, grad, pgrad, rgrad, C_NULL))
    ll[], grad, pgrad[1:K], pgrad[K+1:end], rgrad
end

# Likelihood + branch-length prior in one device call: what logpdfgrad!(::Type{provided}, ...)
# (src/samplers/sampler.jl:172-190) assembles from gradlogpdf(m, target) and the Zygote-differentiated
# prior (src/Likelihood/Prior.jl:39-57).  Topology priors contribute (0, zeros) (Prior.jl:59-66).
prior_spec(::MCPhylo.UniformBranchLength) = (Cint(0), Float64[0.0])
prior_spec(p::MCPhylo.exponentialBL) = (Cint(1), Float64[p.scale])
prior_spec(p::MCPhylo.CompoundDirichlet) = (Cint(2), Float64[p.alpha, p.a, p.beta, p.c])

function posterior(d::PhyloDist, x::Array{Float64,3}, prior)
    NN, po, pa, blv, leaf_nums = flatten(d.tree)
    U, D, Uinv, mu = d.substitution_model(d.base_freq, d.substitution_rates)
    kind, pp = prior_spec(prior)
    lp = Ref{Float64}(0.0)
    grad = Vector{Float64}(undef, NN - 1)
    check(ccall((:mcp_eval_posterior, LIB[]), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}, Cint, Ptr{Float64},
                 Cint, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                context(), alignment(x, leaf_nums), NN, po, pa, Vector{Float64}(blv),
                Matrix{Float64}(U), Vector{Float64}(D), Matrix{Float64}(Uinv), Float64(mu),
                d.rates, length(d.rates), d.base_freq, kind, pp, lp, grad))
    lp[], grad
end

function evaluate(d::MultiplePhyloDist, x::Array{Float64,4}, want_grad::Bool)
    T = length(d.DistCollector)
    flat = [flatten(pd.tree) for pd in d.DistCollector]
    models = [pd.substitution_model(pd.base_freq, pd.substitution_rates) for pd in d.DistCollector]
    slabs = [get!(() -> x[:, :, 1:d.size_array[t], t], SLABS, (objectid(x), t)) for t in 1:T]
    isempty(SLABS) || haskey(ALIGNMENTS, objectid(x)) || (ALIGNMENTS[objectid(x)] = C_NULL; finalizer(release_later, x))
    alns = Ptr{Cvoid}[alignment(slabs[t], flat[t][5]) for t in 1:T]
    NN = Int32[f[1] for f in flat]
    blv = [Vector{Float64}(f[4]) for f in flat]
    U = [Matrix{Float64}(m[1]) for m in models]; D = [Vector{Float64}(m[2]) for m in models]
    Uinv = [Matrix{Float64}(m[3]) for m in models]; mu = Float64[m[4] for m in models]
    rates = [pd.rates for pd in d.DistCollector]; pis = [pd.base_freq for pd in d.DistCollector]
    ll = zeros(Float64, T)
    grads = [Vector{Float64}(undef, NN[t] - 1) for t in 1:T]
    ptrs(v) = Ptr{Cvoid}[pointer(a) for a in v]
    GC.@preserve flat blv U D Uinv rates pis grads begin
        check(ccall((:mcp_eval_batch, LIB[]), Cint,
                    (Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
                     Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Float64},
                     Ptr{Ptr{Cvoid}}, Cint, Ptr{Ptr{Cvoid}}, Cint, Ptr{Float64}, Ptr{Ptr{Cvoid}}),
                    context(), T, alns, NN, ptrs([f[2] for f in flat]), ptrs([f[3] for f in flat]),
                    ptrs(blv), ptrs(U), ptrs(D), ptrs(Uinv), mu, ptrs(rates),
                    length(rates[1]), ptrs(pis), want_grad, ll, want_grad ? ptrs(grads) : C_NULL))
    end
    ll, grads
end

"""Route the four PhyloDist methods through the GPU library (method redefinition)."""
function enable!(libpath::AbstractString = LIB[])
    LIB[] = libpath
    @eval MCPhylo begin
        logpdf(d::PhyloDist, x::Array{Float64,3})::Float64 = $(evaluate)(d, x, false)[1]
        gradlogpdf(d::PhyloDist, x::Array{Float64,3}) = $(evaluate)(d, x, true)
        logpdf(d::MultiplePhyloDist, x::Array{Float64,4})::Float64 = sum($(evaluate)(d, x, false)[1])
        function __logpdf(d::MultiplePhyloDist, x::Array{Float64,4})
            ll, g = $(evaluate)(d, x, true)
            Tuple[(ll[i], g[i]) for i in eachindex(ll)]
        end
        # tree-space HMC target: one PhyloDist-distributed data node + the tree's own prior
        function logpdfgrad!(::Type{provided}, m::Model, x::T, params::ElementOrVector{Symbol},
                             target::ElementOrVector{Symbol}, transform::Bool) where {T<:GeneralNode}
            m[params] = relist(m, x, params, transform)
            tgt = asvec(target)
            t_node = m[asvec(params)[1]]
            if length(tgt) == 1 && t_node.distr isa TreeDistribution
                m[tgt[1]] = update!(m[tgt[1]], m)
                node = m[tgt[1]]
                if node.distr isa PhyloDist && node.value isa Array{Float64,3}
                    return $(posterior)(node.distr, node.value, t_node.distr.length_distr)
                end
            end
            v, grad = gradlogpdf(m, tgt)                    # anything else: the stock composition
            vp, gradp = gradlogpdf(t_node, x)
            vp + v, gradp .+ grad
        end
    end
    nothing
end

"""Frees every device alignment and the context (also safe to call between `mcmc` runs: everything is
re-created on demand)."""
function shutdown!()
    drain_released()
    for (_, a) in ALIGNMENTS
        a == C_NULL || ccall((:mcp_alignment_destroy, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], a)
    end
    empty!(ALIGNMENTS); empty!(SLABS)
    CTX[] == C_NULL || ccall((:mcp_destroy, LIB[]), Cint, (Ptr{Cvoid},), CTX[])
    CTX[] = C_NULL
    nothing
end

end # module
