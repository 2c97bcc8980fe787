# julia/TNQSCuda.jl — the reference-side host code of the CUDA engine: every entry point of include/tnqs_cuda.h bound with
# `ccall`, and a cache type that plugs into TensorNetworkQuantumSimulator's own API (update, apply_gates, expect, rescale,
# symmetric_gauge, truncate, inner, entanglement, network, message ...).
#
# NOT EXECUTED in this repository's build image or on its GPU box (no Julia there); it is the binding the C header was designed
# for.  What keeps it honest meanwhile: tests/test_julia_binding.py checks that every symbol of the header is bound here with the
# header's argument count and types, and tensornetworkquantumsimulator_b200/{lib,engine}.py is the executable ctypes twin of
# this file, function by function (same names, same packing, same order of calls).
#
# Use:   include("julia/TNQSCuda.jl"); using .TNQSCuda
#        ψ = random_tensornetworkstate(ComplexF64, g; bond_dimension = 16)
#        c = CudaBeliefPropagationCache(ψ; chi_cap = 16)            # uploads the state, identity messages
#        c = update(c); c, errs = apply_gates(layer, c; apply_kwargs = (; maxdim = 16, cutoff = 1e-10))
#        expect(c, ("Z", [(16, 16)])); ψ′ = network(c)             # back to ITensors
# ENV["TNQS_CUDA_LIB"] overrides the library path.
#
# Layout contract (include/tnqs_cuda.h): ComplexF64 interleaved = Julia ComplexF64; column-major = Julia Array; vertex id =
# position in collect(vertices(g)) - 1; undirected edge id = position in collect(edges(g)) - 1; directed edge 2e + dir
# (dir 0 = src -> dst); site tensor (s, bonds by ascending edge id); message (b, b').
module TNQSCuda

using ITensors, NamedGraphs, Dictionaries, LinearAlgebra
using NamedGraphs: NamedEdge, NamedGraph
using NamedGraphs.GraphsExtensions: post_order_dfs_edges
using Graphs: vertices, edges, src, dst, a_star, has_edge
import TensorNetworkQuantumSimulator as TN
import TensorNetworkQuantumSimulator: update, apply_gates, expect, rescale, rescale!, symmetric_gauge, symmetric_gauge!,
    network, message, messages, partitionfunction, entanglement, inner, maxvirtualdim, graph
using TensorNetworkQuantumSimulator: TensorNetworkState, TensorNetwork, AbstractBeliefPropagationCache, siteinds, virtualinds,
    forest_cover_edge_sequence, toitensor, edge_color, is_tree

export CudaBeliefPropagationCache, jacobi_update, reduced_density_matrix_1site, bond_dims, kernel_launches

const LIB = get(ENV, "TNQS_CUDA_LIB", joinpath(@__DIR__, "..", "tensornetworkquantumsimulator_b200", "libtnqs_cuda.so"))
const Ctx = Ptr{Cvoid}                         # tnqs_ctx*

# ---- structs of the header (isbits, C layout) -----------------------------------------------------------------------------
struct BpOpts                                   # tnqs_bp_opts
    maxiter::Int32
    tolerance::Float64                          # < 0 = `nothing`
    normalize::Int32
    hermitize::Int32
    mode::Int32                                 # 0 REFERENCE_ORDER, 1 JACOBI
end
struct ApplyOpts                                # tnqs_apply_opts
    maxdim::Int32                               # <= 0 = no limit
    cutoff::Float64                             # NaN = `nothing`
    normalize_tensors::Int32
end
const BP_REFERENCE_ORDER = Int32(0)
const BP_JACOBI = Int32(1)
const KC_COUNT = 8

# ---- raw bindings: one function per symbol of include/tnqs_cuda.h, same argument order -----------------------------------
tnqs_version() = unsafe_string(ccall((:tnqs_version, LIB), Cstring, ()))
tnqs_status_string(st) = unsafe_string(ccall((:tnqs_status_string, LIB), Cstring, (Int32,), st))
tnqs_last_error(ctx) = unsafe_string(ccall((:tnqs_last_error, LIB), Cstring, (Ctx,), ctx))
tnqs_create(out, nv, ne, esrc, edst, d, chi, chi_cap, device) =
    ccall((:tnqs_create, LIB), Int32, (Ref{Ctx}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Int32),
          out, nv, ne, esrc, edst, d, chi, chi_cap, device)
tnqs_destroy(ctx) = ccall((:tnqs_destroy, LIB), Int32, (Ctx,), ctx)
tnqs_clone(ctx, out) = ccall((:tnqs_clone, LIB), Int32, (Ctx, Ref{Ctx}), ctx, out)
tnqs_sync(ctx) = ccall((:tnqs_sync, LIB), Int32, (Ctx,), ctx)
tnqs_num_vertices(ctx) = ccall((:tnqs_num_vertices, LIB), Int32, (Ctx,), ctx)
tnqs_num_edges(ctx) = ccall((:tnqs_num_edges, LIB), Int32, (Ctx,), ctx)
tnqs_get_bond_dims(ctx, chi) = ccall((:tnqs_get_bond_dims, LIB), Int32, (Ctx, Ptr{Int32}), ctx, chi)
tnqs_site_size(ctx, v) = ccall((:tnqs_site_size, LIB), Int64, (Ctx, Int32), ctx, v)
tnqs_set_site(ctx, v, data, n) = ccall((:tnqs_set_site, LIB), Int32, (Ctx, Int32, Ptr{ComplexF64}, Int64), ctx, v, data, n)
tnqs_get_site(ctx, v, data, n) = ccall((:tnqs_get_site, LIB), Int32, (Ctx, Int32, Ptr{ComplexF64}, Int64), ctx, v, data, n)
tnqs_set_sites(ctx, count, v, data, n) =
    ccall((:tnqs_set_sites, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Ptr{ComplexF64}}, Ptr{Int64}), ctx, count, v, data, n)
tnqs_get_sites(ctx, count, v, data, n) =
    ccall((:tnqs_get_sites, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Ptr{ComplexF64}}, Ptr{Int64}), ctx, count, v, data, n)
tnqs_fill_random(ctx, seed) = ccall((:tnqs_fill_random, LIB), Int32, (Ctx, UInt64), ctx, seed)
tnqs_set_message(ctx, k, data, n) = ccall((:tnqs_set_message, LIB), Int32, (Ctx, Int32, Ptr{ComplexF64}, Int64), ctx, k, data, n)
tnqs_get_message(ctx, k, data, n) = ccall((:tnqs_get_message, LIB), Int32, (Ctx, Int32, Ptr{ComplexF64}, Int64), ctx, k, data, n)
tnqs_reset_messages(ctx) = ccall((:tnqs_reset_messages, LIB), Int32, (Ctx,), ctx)
tnqs_bp_update(ctx, edge_seq, nseq, opts, iters, mean_diff) =
    ccall((:tnqs_bp_update, LIB), Int32, (Ctx, Ptr{Int32}, Int32, Ref{BpOpts}, Ref{Int32}, Ref{Float64}),
          ctx, edge_seq, nseq, opts, iters, mean_diff)
tnqs_apply_1site(ctx, n, v, gates, normalize_tensors) =
    ccall((:tnqs_apply_1site, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{ComplexF64}, Int32), ctx, n, v, gates, normalize_tensors)
tnqs_apply_2site(ctx, n, e, flip, gates, opts, ranks, truncerr, svals) =
    ccall((:tnqs_apply_2site, LIB), Int32,
          (Ctx, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{ComplexF64}, Ref{ApplyOpts}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
          ctx, n, e, flip, gates, opts, ranks, truncerr, svals)
tnqs_apply_2site_begin(ctx, n, e, flip, gates, opts, remote, send_buf, n_cut, slot_complex) =
    ccall((:tnqs_apply_2site_begin, LIB), Int32,
          (Ctx, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{ComplexF64}, Ref{ApplyOpts}, Ptr{Int32}, Ref{Ptr{Cvoid}}, Ref{Int64}, Ref{Int64}),
          ctx, n, e, flip, gates, opts, remote, send_buf, n_cut, slot_complex)
tnqs_apply_2site_end(ctx, recv_buf, ranks, truncerr, svals) =
    ccall((:tnqs_apply_2site_end, LIB), Int32, (Ctx, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
          ctx, recv_buf, ranks, truncerr, svals)
tnqs_apply_circuit(ctx, ngates, gate_v, gate_off, gate_data, aopts, edge_seq, nseq, bopts, update_cache, truncerr, refreshes) =
    ccall((:tnqs_apply_circuit, LIB), Int32,
          (Ctx, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{ComplexF64}, Ref{ApplyOpts}, Ptr{Int32}, Int32, Ref{BpOpts}, Int32, Ptr{Float64}, Ref{Int32}),
          ctx, ngates, gate_v, gate_off, gate_data, aopts, edge_seq, nseq, bopts, update_cache, truncerr, refreshes)
tnqs_apply_circuit_fetch(ctx, ngates, gate_v, gate_off, gate_data, aopts, edge_seq, nseq, bopts, update_cache, truncerr, refreshes,
                         fetch_count, fetch_v, fetch_data, fetch_n) =
    ccall((:tnqs_apply_circuit_fetch, LIB), Int32,
          (Ctx, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{ComplexF64}, Ref{ApplyOpts}, Ptr{Int32}, Int32, Ref{BpOpts}, Int32, Ptr{Float64}, Ref{Int32},
           Int32, Ptr{Int32}, Ptr{Ptr{ComplexF64}}, Ptr{Int64}),
          ctx, ngates, gate_v, gate_off, gate_data, aopts, edge_seq, nseq, bopts, update_cache, truncerr, refreshes,
          fetch_count, fetch_v, fetch_data, fetch_n)
tnqs_rdm_1site(ctx, n, v, rho) = ccall((:tnqs_rdm_1site, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{ComplexF64}), ctx, n, v, rho)
tnqs_expect_2site_nn(ctx, n, e, ops, out) =
    ccall((:tnqs_expect_2site_nn, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{ComplexF64}, Ptr{ComplexF64}), ctx, n, e, ops, out)
tnqs_expect_path(ctx, n, path_off, path_v, ops, out) =
    ccall((:tnqs_expect_path, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{ComplexF64}, Ptr{ComplexF64}),
          ctx, n, path_off, path_v, ops, out)
tnqs_loop_weights(ctx, nsub, sub_off, sub_edges, weights) =
    ccall((:tnqs_loop_weights, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{ComplexF64}), ctx, nsub, sub_off, sub_edges, weights)
tnqs_loop_transfer(ctx, len, walk_v, out) =
    ccall((:tnqs_loop_transfer, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{ComplexF64}), ctx, len, walk_v, out)
tnqs_vertex_scalars(ctx, out) = ccall((:tnqs_vertex_scalars, LIB), Int32, (Ctx, Ptr{ComplexF64}), ctx, out)
tnqs_edge_scalars(ctx, out) = ccall((:tnqs_edge_scalars, LIB), Int32, (Ctx, Ptr{ComplexF64}), ctx, out)
tnqs_rescale(ctx) = ccall((:tnqs_rescale, LIB), Int32, (Ctx,), ctx)
tnqs_symmetric_gauge(ctx, regularization) = ccall((:tnqs_symmetric_gauge, LIB), Int32, (Ctx, Float64), ctx, regularization)
tnqs_truncate(ctx, ngroups, group_off, group_edges, aopts, edge_seq, nseq, bopts) =
    ccall((:tnqs_truncate, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Int32}, Ref{ApplyOpts}, Ptr{Int32}, Int32, Ref{BpOpts}),
          ctx, ngroups, group_off, group_edges, aopts, edge_seq, nseq, bopts)
tnqs_set_bra(ctx, bra) = ccall((:tnqs_set_bra, LIB), Int32, (Ctx, Ctx), ctx, bra)
tnqs_message_slab(ctx, dev_ptr, n_complex, slot_stride) =
    ccall((:tnqs_message_slab, LIB), Int32, (Ctx, Ref{Ptr{Cvoid}}, Ref{Int64}, Ref{Int64}), ctx, dev_ptr, n_complex, slot_stride)
tnqs_halo_pack(ctx, n, k, dev_buf) = ccall((:tnqs_halo_pack, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Cvoid}), ctx, n, k, dev_buf)
tnqs_halo_unpack(ctx, n, k, dev_buf) = ccall((:tnqs_halo_unpack, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Cvoid}), ctx, n, k, dev_buf)
tnqs_halo_alloc(ctx, world, nrecv, recv_k, ipc_handle, dev_ptr) =
    ccall((:tnqs_halo_alloc, LIB), Int32, (Ctx, Int32, Int32, Ptr{Int32}, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
          ctx, world, nrecv, recv_k, ipc_handle, dev_ptr)
tnqs_halo_peer(ctx, peer, ipc_handle, raw_dev_ptr, peer_nrecv) =
    ccall((:tnqs_halo_peer, LIB), Int32, (Ctx, Int32, Ptr{UInt8}, Ptr{Cvoid}, Int32), ctx, peer, ipc_handle, raw_dev_ptr, peer_nrecv)
tnqs_halo_route(ctx, n, k, peer, peer_slot) =
    ccall((:tnqs_halo_route, LIB), Int32, (Ctx, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), ctx, n, k, peer, peer_slot)
tnqs_halo_close(ctx) = ccall((:tnqs_halo_close, LIB), Int32, (Ctx,), ctx)
tnqs_set_partition(ctx, owner, rank) = ccall((:tnqs_set_partition, LIB), Int32, (Ctx, Ptr{Int32}, Int32), ctx, owner, rank)
tnqs_kernel_launches(ctx) = ccall((:tnqs_kernel_launches, LIB), Int64, (Ctx,), ctx)
tnqs_bp_sweeps_total(ctx) = ccall((:tnqs_bp_sweeps_total, LIB), Int64, (Ctx,), ctx)
tnqs_bp_message_updates_total(ctx) = ccall((:tnqs_bp_message_updates_total, LIB), Int64, (Ctx,), ctx)
tnqs_profile_enable(ctx, on) = ccall((:tnqs_profile_enable, LIB), Int32, (Ctx, Int32), ctx, on)
tnqs_profile_read(ctx, ms, launches, reset) =
    ccall((:tnqs_profile_read, LIB), Int32, (Ctx, Ptr{Float64}, Ptr{Int64}, Int32), ctx, ms, launches, reset)
tnqs_stream(ctx) = ccall((:tnqs_stream, LIB), Ptr{Cvoid}, (Ctx,), ctx)

# ---- errors: a failed call becomes a Julia exception with the library's message, like the reference's error(...) calls
# (src/utils.jl:38-71); TNQS_ERR_NEGATIVE_EIGENVALUE (-4) keeps the reference's DomainError of sqrt(negative) (src/utils.jl:18-26)
function check(ctx, st)
    st == 0 && return nothing
    msg = "$(tnqs_status_string(st)): $(tnqs_last_error(ctx))"
    st == -4 && throw(DomainError(msg))
    st == -8 && throw(OutOfMemoryError())
    error(msg)
end

# ---- the cache type (extension point: AbstractBeliefPropagationCache, src/MessagePassing/abstractbeliefpropagationcache.jl:4-36)
mutable struct CudaBeliefPropagationCache{V} <: AbstractBeliefPropagationCache{V}
    ctx::Ctx
    g::NamedGraph{V}
    vs::Vector{V}
    es::Vector{NamedEdge{V}}
    vid::Dict{V, Int32}
    eid::Dict{Tuple{V, V}, Int32}               # (src, dst) -> directed edge id 2e + dir
    sinds::Dictionary                            # site indices, to rebuild ITensors on the way out
    binds::Vector{Index}                         # bond indices by edge id; replaced when a gate changes a bond dimension
    chi_cap::Int
    bra::Union{Nothing, CudaBeliefPropagationCache{V}}   # attached bra layer (BilinearForm), kept alive here
end

function _wrap(ctx::Ctx, g::NamedGraph{V}, vs, es, sinds, binds, chi_cap) where {V}
    vid = Dict{V, Int32}(v => Int32(i - 1) for (i, v) in enumerate(vs))
    eid = Dict{Tuple{V, V}, Int32}()
    for (k, e) in enumerate(es)
        eid[(src(e), dst(e))] = Int32(2 * (k - 1))
        eid[(dst(e), src(e))] = Int32(2 * (k - 1) + 1)
    end
    c = CudaBeliefPropagationCache{V}(ctx, g, vs, es, vid, eid, sinds, binds, chi_cap, nothing)
    finalizer(x -> (x.ctx == C_NULL || tnqs_destroy(x.ctx); x.ctx = C_NULL), c)
    return c
end

incident(c::CudaBeliefPropagationCache, v) = _incident(c.es, v)
vertex_id(c::CudaBeliefPropagationCache, v) = c.vid[v]
directed_id(c::CudaBeliefPropagationCache, e) = c.eid[(src(e), dst(e))]
graph(c::CudaBeliefPropagationCache) = c.g
TN.siteinds(c::CudaBeliefPropagationCache) = c.sinds

"""
    CudaBeliefPropagationCache(ψ::TensorNetworkState; chi_cap = maxvirtualdim(ψ), device = 0)

BeliefPropagationCache(ψ) (src/MessagePassing/beliefpropagationcache.jl:9-13,21-25) resident on one GPU: uploads every site
tensor in the ABI's canonical layout with one `tnqs_set_sites` call; messages start as the identity default
(src/TensorNetworks/tensornetworkstate.jl:64-67).  `chi_cap` = the largest bond dimension gates may grow the state to.
"""
function CudaBeliefPropagationCache(ψ::TensorNetworkState; chi_cap::Integer = maxvirtualdim(ψ), device::Integer = 0,
                                    pad_to::Union{Nothing, Vector{<:Integer}} = nothing)
    g = graph(ψ); vs = collect(vertices(g)); es = collect(edges(g))
    binds = Index[only(virtualinds(ψ, e)) for e in es]
    bufs = _canonical_arrays(ψ, vs, es, binds)
    if pad_to !== nothing                         # embed in zero-padded tensors with the given bond dimensions (see `inner`)
        binds = Index[dim(binds[k]) == pad_to[k] ? binds[k] : Index(Int(pad_to[k]); tags = tags(binds[k])) for k in 1:length(es)]
        bufs = [_zero_pad(bufs[i], (size(bufs[i], 1), (dim(binds[k]) for k in _incident(es, v))...)) for (i, v) in enumerate(vs)]
    end
    vid = Dict(v => Int32(i - 1) for (i, v) in enumerate(vs))
    esrc = Int32[vid[src(e)] for e in es]; edst = Int32[vid[dst(e)] for e in es]
    d = Int32[dim(only(siteinds(ψ, v))) for v in vs]
    chi = Int32.(dim.(binds))
    out = Ref{Ctx}(C_NULL)
    check(C_NULL, tnqs_create(out, length(vs), length(es), esrc, edst, d, chi, max(chi_cap, maximum(chi; init = Int32(1))), device))
    c = _wrap(out[], g, vs, es, copy(siteinds(ψ)), binds, Int(max(chi_cap, maximum(chi; init = Int32(1)))))
    _set_sites!(c, bufs)
    return c
end

_incident(es, v) = [k for k in 1:length(es) if src(es[k]) == v || dst(es[k]) == v]

# site tensors in the ABI's canonical layout: (s, bonds by ascending edge id), promoted to ComplexF64
function _canonical_arrays(ψ::TensorNetworkState, vs, es, binds)
    bufs = Vector{Array{ComplexF64}}(undef, length(vs))
    for (i, v) in enumerate(vs)
        legs = [binds[k] for k in _incident(es, v)]
        bufs[i] = ComplexF64.(ITensors.array(ψ[v], only(siteinds(ψ, v)), legs...))
    end
    return bufs
end

function _zero_pad(A::Array{ComplexF64}, dims::Tuple)
    size(A) == dims && return A
    B = zeros(ComplexF64, dims)
    B[(1:n for n in size(A))...] = A
    return B
end

function _set_sites!(c::CudaBeliefPropagationCache, bufs::Vector{<:Array{ComplexF64}})
    ids = Int32.(0:(length(bufs) - 1)); ns = Int64.(length.(bufs))
    GC.@preserve bufs begin
        ptrs = Ptr{ComplexF64}[pointer(b) for b in bufs]
        check(c.ctx, tnqs_set_sites(c.ctx, length(bufs), ids, ptrs, ns))
    end
    return c
end

bond_dims(c::CudaBeliefPropagationCache) = (chi = zeros(Int32, length(c.es)); check(c.ctx, tnqs_get_bond_dims(c.ctx, chi)); chi)
maxvirtualdim(c::CudaBeliefPropagationCache) = isempty(c.es) ? 1 : Int(maximum(bond_dims(c)))
kernel_launches(c::CudaBeliefPropagationCache) = tnqs_kernel_launches(c.ctx)

# bond indices follow the device's bond dimensions (a two-site gate replaces the bond index, as ITensors' svd does)
function _refresh_binds!(c::CudaBeliefPropagationCache)
    for (k, χ) in enumerate(bond_dims(c))
        dim(c.binds[k]) == χ || (c.binds[k] = Index(Int(χ); tags = tags(c.binds[k])))
    end
    return c
end

"""
    network(c) -> TensorNetworkState

Back to ITensors (`network`, src/MessagePassing/beliefpropagationcache.jl:21): one `tnqs_get_sites` call, then
`itensor(A, s, legs...)` per vertex.
"""
function network(c::CudaBeliefPropagationCache{V}) where {V}
    _refresh_binds!(c)
    n = Int64[tnqs_site_size(c.ctx, Int32(i - 1)) for i in 1:length(c.vs)]
    bufs = [Vector{ComplexF64}(undef, n[i]) for i in 1:length(c.vs)]
    GC.@preserve bufs begin
        ptrs = Ptr{ComplexF64}[pointer(b) for b in bufs]
        check(c.ctx, tnqs_get_sites(c.ctx, length(bufs), Int32.(0:(length(bufs) - 1)), ptrs, n))
    end
    tensors = Dictionary{V, ITensor}()
    for (i, v) in enumerate(c.vs)
        s = only(c.sinds[v]); legs = [c.binds[k] for k in incident(c, v)]
        set!(tensors, v, itensor(reshape(bufs[i], (dim(s), dim.(legs)...)), s, legs...))
    end
    return TensorNetworkState(TensorNetwork(tensors, c.g), copy(c.sinds))
end

function Base.copy(c::CudaBeliefPropagationCache)        # src/MessagePassing/beliefpropagationcache.jl:27-29
    out = Ref{Ctx}(C_NULL)
    check(c.ctx, tnqs_clone(c.ctx, out))
    c2 = _wrap(out[], c.g, c.vs, c.es, copy(c.sinds), copy(c.binds), c.chi_cap)
    if c.bra !== nothing
        check(c2.ctx, tnqs_set_bra(c2.ctx, c.bra.ctx)); c2.bra = c.bra
    end
    return c2
end

# message / setmessage! (src/MessagePassing/abstractbeliefpropagationcache.jl:82-91): ITensor with indices (b, b')
function message(c::CudaBeliefPropagationCache, e)
    _refresh_binds!(c)
    k = directed_id(c, e); b = c.binds[k ÷ 2 + 1]; χ = dim(b)
    m = Matrix{ComplexF64}(undef, χ, χ)
    check(c.ctx, tnqs_get_message(c.ctx, k, m, χ * χ))
    return itensor(m, b, b')
end
messages(c::CudaBeliefPropagationCache) = Dictionary([NamedEdge(a => b) for (a, b) in keys(c.eid)], [message(c, NamedEdge(a => b)) for (a, b) in keys(c.eid)])
function TN.setmessage!(c::CudaBeliefPropagationCache, e, m::ITensor)
    k = directed_id(c, e); b = c.binds[k ÷ 2 + 1]
    A = ComplexF64.(ITensors.array(m, b, b'))
    check(c.ctx, tnqs_set_message(c.ctx, k, A, length(A)))
    return c
end
reset_messages!(c::CudaBeliefPropagationCache) = (check(c.ctx, tnqs_reset_messages(c.ctx)); c)

# ---- update (src/MessagePassing/abstractbeliefpropagationcache.jl:201-222; defaults beliefpropagationcache.jl:31,55-65,108-121)
TN.default_bp_maxiter(c::CudaBeliefPropagationCache) = is_tree(c.g) ? 1 : 25
TN.default_bp_update_kwargs(c::CudaBeliefPropagationCache) = (; maxiter = TN.default_bp_maxiter(c), tolerance = 1.0e-8, verbose = false)

_bp_opts(; maxiter, tolerance, normalize = true, enforce_hermiticity = false, mode = BP_REFERENCE_ORDER) =
    BpOpts(maxiter, isnothing(tolerance) ? -1.0 : Float64(tolerance), normalize, enforce_hermiticity, mode)
_edge_seq(c, edge_sequence) = Int32[directed_id(c, e) for e in edge_sequence]

function update!(c::CudaBeliefPropagationCache; maxiter = TN.default_bp_maxiter(c), tolerance = 1.0e-8,
                 edge_sequence = forest_cover_edge_sequence(c.g), verbose = false, normalize = true,
                 enforce_hermiticity = false, mode = BP_REFERENCE_ORDER, kwargs...)
    seq = _edge_seq(c, edge_sequence)
    opts = _bp_opts(; maxiter, tolerance, normalize, enforce_hermiticity, mode)
    iters = Ref{Int32}(0); diff = Ref{Float64}(NaN)
    check(c.ctx, tnqs_bp_update(c.ctx, seq, length(seq), opts, iters, diff))
    verbose && !isnothing(tolerance) && diff[] <= tolerance && println("BP converged to desired precision after $(iters[]) iterations.")
    return c
end
update(c::CudaBeliefPropagationCache; kwargs...) = update!(copy(c); kwargs...)          # non-! functions copy on entry (:206)
jacobi_update(c::CudaBeliefPropagationCache; kwargs...) = update!(copy(c); mode = BP_JACOBI, kwargs...)

# ---- apply_gates (src/Apply/apply_gates.jl:13-23 for a state, :42-92 for a cache)
function _pack_circuit(c::CudaBeliefPropagationCache, circuit::Vector)
    gates = toitensor(circuit, c.g, c.sinds)                                  # gate_definitions.jl:4-6 -> (ITensor, vertices)
    gate_v = fill(Int32(-1), 2 * length(gates)); off = Int64[]; data = ComplexF64[]
    for (i, (o, verts)) in enumerate(gates)
        length(verts) in (1, 2) || error("gates act on one or two vertices")
        for (j, v) in enumerate(verts)
            gate_v[2 * (i - 1) + j] = vertex_id(c, v)
        end
        s = [only(c.sinds[v]) for v in verts]
        push!(off, length(data))
        append!(data, vec(ComplexF64.(ITensors.array(o, prime.(s)..., s...))))   # o[s1', s2', s1, s2]
    end
    return gate_v, off, data
end

_apply_opts(kw) = ApplyOpts(something(get(kw, :maxdim, nothing), 0), something(get(kw, :cutoff, nothing), NaN),
                            get(kw, :normalize_tensors, true))

function apply_gates!(circuit::Vector, c::CudaBeliefPropagationCache; apply_kwargs = (;), bp_update_kwargs = TN.default_bp_update_kwargs(c),
                      update_cache = true, verbose = false)
    gate_v, off, data = _pack_circuit(c, circuit)
    a = _apply_opts(apply_kwargs)
    b = _bp_opts(; maxiter = get(bp_update_kwargs, :maxiter, TN.default_bp_maxiter(c)), tolerance = get(bp_update_kwargs, :tolerance, 1.0e-8))
    seq = _edge_seq(c, get(bp_update_kwargs, :edge_sequence, forest_cover_edge_sequence(c.g)))
    errs = zeros(Float64, max(length(circuit), 1)); nrefresh = Ref{Int32}(0)
    check(c.ctx, tnqs_apply_circuit(c.ctx, length(circuit), gate_v, off, isempty(data) ? ComplexF64[0] : data, a, seq, length(seq), b,
                                    update_cache, errs, nrefresh))
    _refresh_binds!(c)
    return c, errs[1:length(circuit)]
end
apply_gates(circuit::Vector, c::CudaBeliefPropagationCache; kwargs...) = apply_gates!(circuit, copy(c); kwargs...)   # :51

# ---- expect (src/expect.jl:58-83,115-122): 1-site, 2-site on an edge, 2-site on two vertices joined by their shortest path
function reduced_density_matrix_1site(c::CudaBeliefPropagationCache, verts::Vector)
    dmax = maximum(dim(only(c.sinds[v])) for v in c.vs)
    ρ = zeros(ComplexF64, dmax, dmax, length(verts))
    check(c.ctx, tnqs_rdm_1site(c.ctx, length(verts), Int32[vertex_id(c, v) for v in verts], ρ))
    return ρ                                                                   # un-normalised, ρ[s, s', i]
end

_opmatrix(c, name::String, v) = ComplexF64.(ITensors.array(ITensors.op(name, only(c.sinds[v])), only(c.sinds[v])', only(c.sinds[v])))  # O[s', s]

function expect(c::CudaBeliefPropagationCache, obs::Tuple)
    ops, verts, coeff = TN.collectobservable(obs, c.g)                         # src/expect.jl:158-175
    iszero(coeff) && return 0                                                  # :63
    if length(verts) == 1
        v = only(verts); d = dim(only(c.sinds[v]))
        ρ = reduced_density_matrix_1site(c, [v])[1:d, 1:d, 1]
        O = _opmatrix(c, only(ops), v)
        return coeff * sum(transpose(O) .* ρ) / tr(ρ)
    elseif length(verts) == 2
        v1, v2 = verts
        opbuf = vcat(vec(_opmatrix(c, ops[1], v1)), vec(_opmatrix(c, ops[2], v2)))
        out = zeros(ComplexF64, 2)
        path = Int32[vertex_id(c, v1)]
        for e in a_star(c.g, v1, v2)
            push!(path, vertex_id(c, dst(e)))
        end
        check(c.ctx, tnqs_expect_path(c.ctx, 1, Int32[0, length(path)], path, opbuf, out))
        return coeff * out[1] / out[2]
    end
    return coeff * _expect_tree(c, verts, ops)
end

# three or more vertices: the region is the Steiner tree of the support (src/expect.jl:66); a tree region is contracted by passing
# un-normalised messages from its leaves to a root, i.e. with existing entry points on two clones of the handle (engine.py
# `_expect_tree` is the executable twin): numerator = clone with the operators applied and the original state attached as the bra,
# one un-normalised sweep over the child -> parent edges, the root's vertex scalar; denominator = the same without operators.
function _expect_tree(c::CudaBeliefPropagationCache, verts::Vector, ops::Vector{String})
    t = NamedGraphs.steiner_tree(c.g, verts)
    tvs = collect(vertices(t)); root = first(tvs)
    for e in edges(c.g)                                           # a chord would make the region loopy
        (src(e) in tvs && dst(e) in tvs && !(has_edge(t, src(e), dst(e)))) && error("the region of this observable has a loop")
    end
    # post_order_dfs_edges lists child -> parent edges, children first (the first half of forest_cover_edge_sequence,
    # src/MessagePassing/beliefpropagationcache.jl:80-81)
    seq = Int32[directed_id(c, e) for e in post_order_dfs_edges(t, root)]
    vals = ComplexF64[]
    for with_ops in (true, false)
        w = copy(c)
        if with_ops
            vids = Int32[vertex_id(c, v) for v in verts]
            gates = reduce(vcat, [vec(_opmatrix(c, o, v)) for (o, v) in zip(ops, verts)])
            check(w.ctx, tnqs_apply_1site(w.ctx, length(vids), vids, gates, 0))
            check(w.ctx, tnqs_set_bra(w.ctx, c.ctx)); w.bra = c
        end
        opts = BpOpts(1, -1.0, 0, 0, BP_REFERENCE_ORDER)
        iters = Ref{Int32}(0); diff = Ref{Float64}(NaN)
        check(w.ctx, tnqs_bp_update(w.ctx, seq, length(seq), opts, iters, diff))
        num = zeros(ComplexF64, length(c.vs)); check(w.ctx, tnqs_vertex_scalars(w.ctx, num))
        push!(vals, num[vertex_id(c, root) + 1])
        with_ops && check(w.ctx, tnqs_set_bra(w.ctx, C_NULL))
    end
    return vals[1] / vals[2]
end

expect(c::CudaBeliefPropagationCache, obs::Vector{<:Tuple}) = [expect(c, o) for o in obs]

# ---- scalars, rescale, gauge, truncate, entanglement
function partitionfunction(c::CudaBeliefPropagationCache)                      # abstractbeliefpropagationcache.jl:252-267
    num = zeros(ComplexF64, length(c.vs)); den = zeros(ComplexF64, length(c.es))
    check(c.ctx, tnqs_vertex_scalars(c.ctx, num)); check(c.ctx, tnqs_edge_scalars(c.ctx, den))
    any(iszero, den) && return zero(ComplexF64)
    return exp(sum(log, num) - sum(log, den; init = zero(ComplexF64)))
end
rescale!(c::CudaBeliefPropagationCache) = (check(c.ctx, tnqs_rescale(c.ctx)); c)                      # :281-291
rescale(c::CudaBeliefPropagationCache) = rescale!(copy(c))
symmetric_gauge!(c::CudaBeliefPropagationCache; regularization = 10 * eps(Float64)) =
    (check(c.ctx, tnqs_symmetric_gauge(c.ctx, regularization)); c)                                     # src/symmetric_gauge.jl:1-56
symmetric_gauge(c::CudaBeliefPropagationCache; kwargs...) = symmetric_gauge!(copy(c); kwargs...)

function Base.truncate(c::CudaBeliefPropagationCache; maxdim, cutoff = nothing, normalize_tensors = true,
                       bp_update_kwargs = TN.default_bp_update_kwargs(c))                             # src/truncate.jl:12-38
    c = copy(c)
    z = maximum(length(incident(c, v)) for v in c.vs)
    groups = edge_color(c.g, z)                                                # :20
    undirected = Dict((src(e), dst(e)) => Int32(k - 1) for (k, e) in enumerate(c.es))
    eidx(p) = haskey(undirected, (first(p), last(p))) ? undirected[(first(p), last(p))] : undirected[(last(p), first(p))]
    group_off = Int32[0]; group_edges = Int32[]
    for grp in groups
        append!(group_edges, Int32[eidx(p) for p in grp]); push!(group_off, length(group_edges))
    end
    a = ApplyOpts(maxdim, something(cutoff, NaN), normalize_tensors)
    b = _bp_opts(; maxiter = bp_update_kwargs.maxiter, tolerance = bp_update_kwargs.tolerance)
    seq = _edge_seq(c, forest_cover_edge_sequence(c.g))
    check(c.ctx, tnqs_truncate(c.ctx, length(groups), group_off, group_edges, a, seq, length(seq), b))
    return _refresh_binds!(c)
end

# ---- loop corrections (src/MessagePassing/loopcorrection.jl): the generalised loops come from NamedGraphs as in the reference
# (:11-12), their weights from the device (one call), the series is summed here (:13-15)
function loop_weights(c::CudaBeliefPropagationCache, egs)                      # weights(bpc, egs), loopcorrection.jl:81-96
    undirected = Dict{Any, Int32}()
    for (k, e) in enumerate(c.es)
        undirected[(src(e), dst(e))] = Int32(k - 1); undirected[(dst(e), src(e))] = Int32(k - 1)
    end
    sub_off = Int32[0]; sub_edges = Int32[]
    for eg in egs
        append!(sub_edges, Int32[undirected[(src(e), dst(e))] for e in edges(eg)]); push!(sub_off, length(sub_edges))
    end
    ws = zeros(ComplexF64, length(egs))
    check(c.ctx, tnqs_loop_weights(c.ctx, length(egs), sub_off, sub_edges, ws))
    return ws
end

function TN.loopcorrected_partitionfunction(c::CudaBeliefPropagationCache, max_configuration_size::Integer)    # :3-16
    zbp = partitionfunction(c)
    egs = NamedGraphs.edgeinduced_subgraphs_no_leaves(c.g, max_configuration_size)
    isempty(egs) && return zbp
    return zbp * (1 + sum(loop_weights(rescale(c), egs)))
end

# loop_correlation (src/MessagePassing/beliefpropagationcache.jl:145-179): the transfer matrix of the closed walk comes from the
# device, its eigenvalues are taken here
function TN.loop_correlation(c::CudaBeliefPropagationCache, loop::Vector{<:NamedEdge}, target_e::NamedEdge)
    is_tree(c.g) && return 0
    walk = Int32[vertex_id(c, src(e)) for e in vcat(loop, [target_e])]
    χ = Int(bond_dims(c)[directed_id(c, target_e) ÷ 2 + 1])
    t = Matrix{ComplexF64}(undef, χ^2, χ^2)
    check(c.ctx, tnqs_loop_transfer(c.ctx, length(walk), walk, t))
    λs = reverse(sort(eigvals(t); by = abs))
    return 1 - abs(λs[1]) / sum(abs.(λs))
end

function entanglement(c::CudaBeliefPropagationCache, e)                        # src/symmetric_gauge.jl:85-114
    k = directed_id(c, e); χ = Int(bond_dims(c)[k ÷ 2 + 1])
    m1 = Matrix{ComplexF64}(undef, χ, χ); m2 = similar(m1)
    check(c.ctx, tnqs_get_message(c.ctx, k, m1, χ * χ)); check(c.ctx, tnqs_get_message(c.ctx, k ⊻ Int32(1), m2, χ * χ))
    x = sqrt(Hermitian((m1 + m1') / 2 + 10 * eps() * I)); y = sqrt(Hermitian((m2 + m2') / 2 + 10 * eps() * I))
    p = svdvals(transpose(x) * y) .^ 2
    p ./= sum(p)
    return -sum(q -> q > 0 ? q * log(q) : 0.0, p)
end

# ---- inner(ψ, ϕ; alg = "bp") (src/Forms/bilinearform.jl:28-54, src/inner.jl:82-86): BP on <ϕ|ψ> with ϕ attached as the bra layer.
# The engine wants equal bond dimensions on both layers: the thinner layer is embedded in zero-padded tensors (the value of the
# contraction is unchanged; engine.py `_pad_bonds` is the executable twin).
function inner(ket::CudaBeliefPropagationCache, bra::CudaBeliefPropagationCache; maxiter = TN.default_bp_maxiter(ket), tolerance = 1.0e-8)
    bond_dims(ket) == bond_dims(bra) || error("inner: both layers need the same bond dimensions (use the TensorNetworkState method)")
    c = copy(ket)
    check(c.ctx, tnqs_set_bra(c.ctx, bra.ctx)); c.bra = bra
    reset_messages!(c)
    update!(c; maxiter, tolerance)
    return partitionfunction(c)
end

function inner(ψ::TensorNetworkState, ϕ::TensorNetworkState, ::Val{:cuda}; device::Integer = 0, kwargs...)
    es = collect(edges(graph(ψ)))
    χ = [max(dim(only(virtualinds(ψ, e))), dim(only(virtualinds(ϕ, e)))) for e in es]
    ket = CudaBeliefPropagationCache(ψ; pad_to = χ, device)
    bra = CudaBeliefPropagationCache(ϕ; pad_to = χ, device)
    return inner(ket, bra; kwargs...)
end

# ---- TensorNetworkState front doors: upload, run, download (what `apply_gates(circuit, ψ::TensorNetworkState)` does on the CPU,
# src/Apply/apply_gates.jl:13-23)
function apply_gates(circuit::Vector, ψ::TensorNetworkState, ::Val{:cuda}; apply_kwargs = (;), bp_update_kwargs = nothing,
                     chi_cap::Integer = max(maxvirtualdim(ψ), something(get(apply_kwargs, :maxdim, nothing), 64)), device::Integer = 0, kwargs...)
    c = CudaBeliefPropagationCache(ψ; chi_cap, device)
    kw = isnothing(bp_update_kwargs) ? TN.default_bp_update_kwargs(c) : bp_update_kwargs
    update!(c; kw...)                                                          # apply_gates.jl:19
    c, errs = apply_gates!(circuit, c; apply_kwargs, bp_update_kwargs = kw, kwargs...)
    return network(c), errs
end

function expect(ψ::TensorNetworkState, obs, ::Val{:cuda}; device::Integer = 0, cache_update_kwargs = nothing)   # src/expect.jl:124-136
    c = CudaBeliefPropagationCache(ψ; device)
    kw = isnothing(cache_update_kwargs) ? TN.default_bp_update_kwargs(c) : cache_update_kwargs
    return expect(update!(c; kw...), obs)
end

end # module
