/*
 * tnqs_cuda.h — C-ABI of libtnqs_cuda.so: the B200-native (sm_100a) engine for the two data-parallel hot
 * paths of haimeng-zhang/TensorNetworkQuantumSimulator:
 *   (1) the belief-propagation (BP) message sweep over every edge of a TensorNetworkState, and
 *   (2) the batched, non-overlapping two-site gate application with SVD truncation (simple update),
 * plus the BP expectation values / scalars / rescaling / symmetric gauge / BP truncation on the same path.
 *
 * The reference is pure Julia (no FFI exists); each entry point below names the Julia method it replaces
 * (paths relative to the reference repo).  A Julia caller binds these with `ccall`;
 * this repo drives them through Python ctypes.
 *
 * Conventions
 *   - ComplexF64 = interleaved double[2] (Julia ComplexF64 / C double _Complex / cuDoubleComplex); all
 *     arrays column-major (first index fastest); ids are 0-based int32.
 *   - Undirected edge e = (esrc[e], edst[e]).  Directed edge id k = 2*e + dir, dir 0 = esrc->edst.
 *   - Site tensor of vertex v: T[s, b_e1, .., b_ez], incident edges e1 < e2 < .. < ez (ascending edge id).
 *   - Message of directed edge k: m[b, b'] (chi_e x chi_e); b contracts the ket leg, b' the bra leg.
 *   - 1-site gate o[s', s]; 2-site gate o[s1', s2', s1, s2] (first vertex = the edge's `flip ? edst : esrc`).
 *   - Host buffers are caller-owned and only touched during the call.  Every function returns a status
 *     (0 = OK, negative = TNQS_ERR_*); no exceptions, callbacks or longjmp cross the boundary.
 *   - One handle = one network (+ messages) resident on one GPU; not thread-safe per handle.  Multi-GPU runs
 *     use one handle per rank (vertex partition + halo, see tnqs_halo_* and DESIGN.md section (e)).
 *   - There is NO CPU fallback: creating a handle without a CUDA device fails with TNQS_ERR_CUDA.
 */
#ifndef TNQS_CUDA_H
#define TNQS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnqs_ctx tnqs_ctx;

enum {
  TNQS_OK = 0,
  TNQS_ERR_INVALID_ARG = -1,
  TNQS_ERR_SHAPE = -2,
  TNQS_ERR_OVERLAPPING_BATCH = -3,
  TNQS_ERR_NEGATIVE_EIGENVALUE = -4,   /* reference: sqrt of a negative real throws (src/utils.jl:18-26) */
  TNQS_ERR_SVD_NO_CONVERGENCE = -5,
  TNQS_ERR_CUDA = -6,
  TNQS_ERR_NCCL = -7,
  TNQS_ERR_OOM = -8,
  TNQS_ERR_UNSUPPORTED = -9
};

/* BP sweep scheduling. REFERENCE_ORDER = bit-equivalent to the sequential Gauss-Seidel loop over the given
 * edge sequence (dependency-level schedule); JACOBI = every message of a sweep from the previous sweep's
 * values (same fixed point, different iterates; used for throughput and multi-GPU runs). */
enum { TNQS_BP_REFERENCE_ORDER = 0, TNQS_BP_JACOBI = 1 };

/* Options of `update` (src/MessagePassing/beliefpropagationcache.jl:43-65,108-121). */
typedef struct {
  int32_t maxiter;        /* >= 1 (reference default 25, 1 on trees) */
  double tolerance;       /* < 0  => `nothing`: run all maxiter sweeps, no diffs computed */
  int32_t normalize;      /* message /= sum of all entries (default 1) */
  int32_t hermitize;      /* enforce_hermiticity (default 0) */
  int32_t mode;           /* TNQS_BP_* */
} tnqs_bp_opts;

/* apply_kwargs of apply_gates / simple_update (src/Apply/simple_update.jl:17-19,46-52). */
typedef struct {
  int32_t maxdim;         /* <= 0 => no limit */
  double cutoff;          /* NaN => `nothing` (no cutoff-based truncation); relative, on S^2 */
  int32_t normalize_tensors;
} tnqs_apply_opts;

const char* tnqs_version(void);
const char* tnqs_status_string(int32_t status);
/* Last error message of this handle (or of the failed tnqs_create when ctx == NULL). */
const char* tnqs_last_error(const tnqs_ctx* ctx);

/* ---- state container: TensorNetworkState + BeliefPropagationCache ------------------------------------------
 * replaces: struct TensorNetworkState (src/TensorNetworks/tensornetworkstate.jl:4-7),
 *           struct BeliefPropagationCache (src/MessagePassing/beliefpropagationcache.jl:9-13).
 * d[nv] site dimensions, chi[ne] initial bond dimensions, chi_cap = largest bond dimension any edge may reach
 * (messages are stored in chi_cap^2 slots).  Sites start zeroed, messages start as the identity default
 * (src/TensorNetworks/tensornetworkstate.jl:64-67).  device = CUDA device ordinal. */
int32_t tnqs_create(tnqs_ctx** out, int32_t nv, int32_t ne, const int32_t* esrc, const int32_t* edst,
                    const int32_t* d, const int32_t* chi, int32_t chi_cap, int32_t device);
int32_t tnqs_destroy(tnqs_ctx* ctx);
/* Deep copy on the same device (Base.copy of the cache, src/MessagePassing/beliefpropagationcache.jl:27-29). */
int32_t tnqs_clone(const tnqs_ctx* ctx, tnqs_ctx** out);
int32_t tnqs_sync(tnqs_ctx* ctx);

int32_t tnqs_num_vertices(const tnqs_ctx* ctx);
int32_t tnqs_num_edges(const tnqs_ctx* ctx);
int32_t tnqs_get_bond_dims(const tnqs_ctx* ctx, int32_t* chi /* [ne] */);
int64_t tnqs_site_size(const tnqs_ctx* ctx, int32_t v);      /* number of ComplexF64 elements */

/* Site tensors (setindex_preserve!/getindex, src/TensorNetworks/abstracttensornetwork.jl:41-44).
 * `n` = element count of the host buffer, must equal tnqs_site_size. */
int32_t tnqs_set_site(tnqs_ctx* ctx, int32_t v, const double* data, int64_t n);
int32_t tnqs_get_site(const tnqs_ctx* ctx, int32_t v, double* data, int64_t n);
/* The same for a list of sites in one call (a whole state): data[i] / n[i] belong to vertex v[i].  Every copy is queued on
 * the handle's stream and the call synchronises once, so pinned host buffers move at PCIe speed instead of paying one
 * round trip per site (what a Julia caller does for `ψ[v] = T` over all v, tensornetworkstate.jl:124-144). */
int32_t tnqs_set_sites(tnqs_ctx* ctx, int32_t count, const int32_t* v, const double* const* data, const int64_t* n);
int32_t tnqs_get_sites(const tnqs_ctx* ctx, int32_t count, const int32_t* v, double* const* data, const int64_t* n);
/* Fill every site with N(0,1)+iN(0,1) from the counter-based generator (stream = vertex id, element =
 * column-major offset); random_tensornetworkstate, src/TensorNetworks/tensornetworkstate.jl:82-92. */
int32_t tnqs_fill_random(tnqs_ctx* ctx, uint64_t seed);

/* Messages (message/setmessage!, src/MessagePassing/abstractbeliefpropagationcache.jl:82-91). n = chi_e^2. */
int32_t tnqs_set_message(tnqs_ctx* ctx, int32_t k, const double* data, int64_t n);
int32_t tnqs_get_message(const tnqs_ctx* ctx, int32_t k, double* data, int64_t n);
int32_t tnqs_reset_messages(tnqs_ctx* ctx);                  /* back to the identity default */

/* ---- BP sweep -------------------------------------------------------------------------------------------------
 * replaces: update(bpc; maxiter, tolerance, edge_sequence) (src/MessagePassing/abstractbeliefpropagationcache.jl:
 * 201-222), update_iteration! (:182-196), updated_message (:144-168), message_diff
 * (src/MessagePassing/beliefpropagationcache.jl:15-19).  edge_seq = directed edge ids in update order
 * (the reference's forest_cover_edge_sequence is an input, :74-85).  Outputs: sweeps run, last mean diff
 * (NaN when tolerance < 0). */
int32_t tnqs_bp_update(tnqs_ctx* ctx, const int32_t* edge_seq, int32_t nseq, const tnqs_bp_opts* opts,
                       int32_t* iters, double* mean_diff);

/* ---- gates ----------------------------------------------------------------------------------------------------
 * replaces: simple_update 1-site branch (src/Apply/simple_update.jl:21-23,63-65).  Vertices must be distinct. */
int32_t tnqs_apply_1site(tnqs_ctx* ctx, int32_t n, const int32_t* v, const double* gates /* n*d*d cplx */,
                         int32_t normalize_tensors);
/* replaces: apply_gate! + simple_update 2-site branch (src/Apply/apply_gates.jl:95-122,
 * src/Apply/simple_update.jl:24-66, src/utils.jl:18-26).  Edges must be vertex-disjoint
 * (TNQS_ERR_OVERLAPPING_BATCH otherwise).  Outputs per gate: kept rank, truncation error, singular values
 * (svals[i*chi_cap + j], j < rank; normalised if normalize_tensors).  Any output pointer may be NULL. */
int32_t tnqs_apply_2site(tnqs_ctx* ctx, int32_t n, const int32_t* e, const int32_t* flip,
                         const double* gates /* n*d^4 cplx */, const tnqs_apply_opts* opts,
                         int32_t* ranks, double* truncerr, double* svals);
/* Two-stage form of tnqs_apply_2site for vertex-partitioned runs.  `remote[i]` = 0 (both vertices
 * local), 1 (the gate's first vertex lives on another rank) or 2 (the second one does).  _begin runs everything that needs
 * one site only (message roots, absorption, QR) and exposes a device buffer with the local R factors of the n_cut
 * cut-crossing gates (batch order; slot_complex ComplexF64 per slot: a (q, nc) header, then R).  The host exchanges the slots
 * with the partner ranks (NCCL) and passes the received buffer (same order and layout) to _end, which runs the identical
 * theta / SVD / truncation on both sides and finishes the local halves.  No other call may touch the handle in between. */
int32_t tnqs_apply_2site_begin(tnqs_ctx* ctx, int32_t n, const int32_t* e, const int32_t* flip, const double* gates,
                               const tnqs_apply_opts* opts, const int32_t* remote, void** send_buf, int64_t* n_cut,
                               int64_t* slot_complex);
int32_t tnqs_apply_2site_end(tnqs_ctx* ctx, const void* recv_buf, int32_t* ranks, double* truncerr, double* svals);
/* replaces: apply_gates(circuit, bpc; apply_kwargs, bp_update_kwargs, update_cache)
 * (src/Apply/apply_gates.jl:42-92): ordered gate list, BP refresh before a 2-site gate that touches a vertex
 * touched since the last refresh (1-site gates mark vertices too, :84), final refresh if update_cache.
 * Gate i acts on gate_v[2i] (and gate_v[2i+1], or -1 for 1-site gates); its array starts at
 * gate_data[gate_off[i]] (complex elements).  Consecutive independent gates run as one batch.
 * truncerr[ngates]; refreshes = number of BP updates performed. */
int32_t tnqs_apply_circuit(tnqs_ctx* ctx, int32_t ngates, const int32_t* gate_v, const int64_t* gate_off,
                           const double* gate_data, const tnqs_apply_opts* aopts,
                           const int32_t* edge_seq, int32_t nseq, const tnqs_bp_opts* bopts,
                           int32_t update_cache, double* truncerr, int32_t* refreshes);
/* The same, and every listed site tensor is copied to the host as soon as the last gate is done, on a second stream that runs
 * beside the closing BP refresh (what a caller does next anyway: `ψ[v]` for all v after apply_gates).  fetch_n[i]: capacity
 * of fetch_data[i] in ComplexF64 on entry, element count written on return (bond dimensions may have changed:
 * tnqs_get_bond_dims).  Pinned host buffers make the copies asynchronous; pageable ones work but serialise. */
int32_t tnqs_apply_circuit_fetch(tnqs_ctx* ctx, int32_t ngates, const int32_t* gate_v, const int64_t* gate_off,
                                 const double* gate_data, const tnqs_apply_opts* aopts,
                                 const int32_t* edge_seq, int32_t nseq, const tnqs_bp_opts* bopts,
                                 int32_t update_cache, double* truncerr, int32_t* refreshes,
                                 int32_t fetch_count, const int32_t* fetch_v, double* const* fetch_data, int64_t* fetch_n);

/* ---- BP measurements ------------------------------------------------------------------------------------------
 * replaces: expect(::Algorithm"bp", cache, obs) for one-site observables (src/expect.jl:58-83): returns the
 * un-normalised BP reduced density matrices rho[i][s, s'] = sum T[s..] conj(T[s'..]) prod m_in (d*d cplx each);
 * <O> = sum O[s',s] rho[s,s'] / tr(rho).  Also serves rdm (src/rdm.jl:49-70) for single vertices. */
int32_t tnqs_rdm_1site(tnqs_ctx* ctx, int32_t n, const int32_t* v, double* rho /* n*dmax*dmax cplx, dmax=max d */);
/* Two-site nearest-neighbour observables on edges e[i]: out[i] = <O1 (x) O2> numerator and denominator
 * (out[4i..4i+1] = numer re,im; out[4i+2..4i+3] = denom re,im); ops[i] = O1[s',s] then O2[s',s] (2*d*d cplx). */
int32_t tnqs_expect_2site_nn(tnqs_ctx* ctx, int32_t n, const int32_t* e, const double* ops, double* out);
/* Two-site observables on vertices joined by a path of the network (the Steiner tree of the support, src/expect.jl:66):
 * path i = path_v[path_off[i] .. path_off[i+1]) (consecutive vertices share an edge), O1 acts on its first vertex, O2 on
 * its last; ops and out as in tnqs_expect_2site_nn. */
int32_t tnqs_expect_path(tnqs_ctx* ctx, int32_t n, const int32_t* path_off, const int32_t* path_v, const double* ops, double* out);
/* replaces: weights(bpc, egs) / weight(bpc, eg) (src/MessagePassing/loopcorrection.jl:81-96), the terms of
 * loopcorrected_partitionfunction (:3-16).  Subgraph i is induced by the edges sub_edges[sub_off[i] .. sub_off[i+1]) (no
 * duplicates; it may be disconnected: the weight is then the product over its components).  weights[i] (cplx) = the
 * contraction of the factors of its vertices with the messages that enter them along every other edge and, on each of its
 * own edges, the antiprojector 1 - message(e) (x) message(reverse e) (:49-60).  The handle must hold the RESCALED cache
 * (tnqs_rescale: edge and vertex scalars 1), as the reference's caller does (:8).  Evaluated as message updates with
 * substituted messages over a spanning tree of the subgraph; cost grows as prod over its independent cycles of (chi^2 + 1):
 * TNQS_ERR_UNSUPPORTED beyond 2^22 tree contractions per subgraph.  Not available on a partitioned handle. */
int32_t tnqs_loop_weights(tnqs_ctx* ctx, int32_t nsub, const int32_t* sub_off, const int32_t* sub_edges, double* weights);
/* replaces: the contraction inside loop_correlation (src/MessagePassing/beliefpropagationcache.jl:145-174).  walk_v =
 * the len >= 3 distinct vertices of a closed walk v_0 -> v_1 -> .. -> v_{len-1} -> v_0; the bond of the last step is cut open.
 * out = chi^2 x chi^2 cplx, column-major: column (b, b') = the environment that arrives at the open leg of v_{len-1} when the
 * leg of v_0 carries the matrix unit |b)(b'| (ket index fastest in rows and columns).  loop_correlation is
 * 1 - |lambda_1| / sum |lambda_i| of its eigenvalues (:175-178), which the caller takes on the host. */
int32_t tnqs_loop_transfer(tnqs_ctx* ctx, int32_t len, const int32_t* walk_v, double* out);
/* vertex_scalar (src/MessagePassing/abstractbeliefpropagationcache.jl:21-27) for all vertices, [nv] cplx;
 * edge_scalar (src/MessagePassing/beliefpropagationcache.jl:38-40) for all edges, [ne] cplx. */
int32_t tnqs_vertex_scalars(tnqs_ctx* ctx, double* out);
int32_t tnqs_edge_scalars(tnqs_ctx* ctx, double* out);
/* rescale! (src/MessagePassing/abstractbeliefpropagationcache.jl:281-285; rescale_messages!
 * src/MessagePassing/beliefpropagationcache.jl:129-142, rescale_vertices! :87-106). */
int32_t tnqs_rescale(tnqs_ctx* ctx);

/* ---- gauge / truncation ---------------------------------------------------------------------------------------
 * replaces: symmetric_gauge! (src/symmetric_gauge.jl:1-56). */
int32_t tnqs_symmetric_gauge(tnqs_ctx* ctx, double regularization);
/* replaces: truncate(bpc; maxdim, cutoff, normalize_tensors, bp_update_kwargs) (src/truncate.jl:12-38): identity
 * gates per colour group (groups given as group_off[ngroups+1] offsets into group_edges), BP update after each. */
int32_t tnqs_truncate(tnqs_ctx* ctx, int32_t ngroups, const int32_t* group_off, const int32_t* group_edges,
                      const tnqs_apply_opts* aopts, const int32_t* edge_seq, int32_t nseq,
                      const tnqs_bp_opts* bopts);

/* replaces: BilinearForm(ket, bra) + BeliefPropagationCache over it (src/Forms/bilinearform.jl:28-54, src/inner.jl:82-86).
 * With a bra attached, every BP contraction of `ctx` (tnqs_bp_update, tnqs_vertex_scalars, ...) closes the ket tensors of `ctx`
 * against the conjugated site tensors of `bra` (same device, graph, site and bond dimensions; it must stay alive and unchanged
 * while attached; NULL detaches).  Messages are chi x chi but no longer Hermitian; inner(psi, phi; alg = "bp") is the product of
 * the vertex scalars over the product of the edge scalars after an update.  Gates and gauges are meaningless on such a handle. */
int32_t tnqs_set_bra(tnqs_ctx* ctx, tnqs_ctx* bra);

/* ---- multi-GPU halo (vertex partition; one handle per rank holds the whole graph's topology but only its own
 * vertices' tensors are meaningful) -------------------------------------------------------------------------
 * Device pointer + element count of the contiguous message slab; slot k starts at k*chi_cap^2 complex elements.
 * The host side exchanges boundary slots with NCCL (torch.distributed) between sweeps. */
int32_t tnqs_message_slab(tnqs_ctx* ctx, void** dev_ptr, int64_t* n_complex, int64_t* slot_stride);
/* Gather / scatter the listed directed-edge slots to / from a contiguous device buffer (halo pack / unpack). */
int32_t tnqs_halo_pack(tnqs_ctx* ctx, int32_t n, const int32_t* k, void* dev_buf);
int32_t tnqs_halo_unpack(tnqs_ctx* ctx, int32_t n, const int32_t* k, const void* dev_buf);
/* Peer-memory halo: the boundary exchange and the convergence all-reduce of vertex-partitioned BP, fused into the kernels.
 * Every handle owns one block (cudaMalloc, exportable over CUDA IPC) = a signal area + a receive buffer of 2 x nrecv message
 * slots (two sweep parities); slot j belongs to the ghost -> owned message recv_k[j].  A rank opens the blocks of ALL other
 * ranks (CUDA IPC handle for another process of the node, plain device pointer inside one process) and routes each
 * owned -> ghost message k to (peer, slot).  From then on tnqs_bp_update (JACOBI) needs no host-side exchange:
 *   - the message epilogue stores a boundary message both locally and, with a P2P store over NVLink, in the peer's buffer;
 *   - every sweep ends with a device-side barrier: each rank stores {epoch, diff sum, message count} into every rank's signal
 *     area and waits for the others' (k_halo_barrier), sums in rank order (the same bits everywhere) and commits the half
 *     of its receive buffer written in that sweep; the next sweep writes the other half;
 *   - the "stop when the global mean diff <= tolerance or after maxiter sweeps" test runs on the device (WHILE graph node), so a
 *     whole refresh is ONE graph launch per rank.  All ranks must call tnqs_bp_update together with the same options; iters and
 *     mean_diff are global.  A rank that never arrives makes the others fail with TNQS_ERR_NCCL after a time-out (no hang).
 * ipc_handle: 64 bytes (cudaIpcMemHandle_t), may be NULL; dev_ptr: the block's device address, may be NULL. */
int32_t tnqs_halo_alloc(tnqs_ctx* ctx, int32_t world, int32_t nrecv, const int32_t* recv_k, void* ipc_handle, void** dev_ptr);
int32_t tnqs_halo_peer(tnqs_ctx* ctx, int32_t peer, const void* ipc_handle, void* raw_dev_ptr, int32_t peer_nrecv);
int32_t tnqs_halo_route(tnqs_ctx* ctx, int32_t n, const int32_t* k, const int32_t* peer, const int32_t* peer_slot);
int32_t tnqs_halo_close(tnqs_ctx* ctx);
/* Restrict BP updates / diffs to messages whose source vertex has owner[v] == rank (NULL => own everything). */
int32_t tnqs_set_partition(tnqs_ctx* ctx, const int32_t* owner, int32_t rank);

/* ---- instrumentation ------------------------------------------------------------------------------------------
 * Kernel launches issued by this handle since creation (the `gpu_launches` figure of bench.py). */
int64_t tnqs_kernel_launches(const tnqs_ctx* ctx);
/* BP sweeps / directed-edge message updates performed by this handle since creation (all entry points). */
int64_t tnqs_bp_sweeps_total(const tnqs_ctx* ctx);
int64_t tnqs_bp_message_updates_total(const tnqs_ctx* ctx);
/* Device time (ms) and launch count per kernel class since the last reset, measured with CUDA events on the
 * handle's stream when profiling is enabled: classes are TNQS_KC_*.  */
enum { TNQS_KC_MODE_APPLY = 0, TNQS_KC_MODE_CLOSE = 1, TNQS_KC_BP_FINALIZE = 2, TNQS_KC_PERMUTE = 3,
       TNQS_KC_EIGH = 4, TNQS_KC_QR = 5, TNQS_KC_SVD = 6, TNQS_KC_MISC = 7, TNQS_KC_COUNT = 8 };
int32_t tnqs_profile_enable(tnqs_ctx* ctx, int32_t on);
int32_t tnqs_profile_read(tnqs_ctx* ctx, double* ms /* [TNQS_KC_COUNT] */, int64_t* launches /* [TNQS_KC_COUNT] */,
                          int32_t reset);
/* The handle's CUDA stream (cudaStream_t) so a host can bracket calls with its own events. */
void* tnqs_stream(tnqs_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TNQS_CUDA_H */
