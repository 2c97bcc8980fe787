"""Vertex-partitioned multi-GPU execution of the two hot paths.

One process per GPU (`torch.distributed`, NCCL over NVLink; gloo on CPU for the host-logic tests).  Every rank owns a
set of vertices (contiguous row strips for grids) and builds a *local* graph = owned vertices + ghost vertices (the
neighbours across a cut) + every edge with an owned endpoint, keeping the global relative order of vertices and edges so
the canonical leg order of owned site tensors is unchanged.  Ghost sites carry no data; only messages cross a cut:

  * BP runs in JACOBI mode: all owned-source messages from the previous sweep's values.  With the peer-memory halo
    (`halo_setup_p2p`, the default on GPUs) a whole refresh is ONE C-ABI call = one CUDA-graph launch per rank: the message
    epilogues store every boundary message straight into the receive buffer of the neighbouring GPU (P2P store over NVLink,
    blocks opened through CUDA IPC), every sweep ends with a device-side barrier that also all-reduces the diff sum
    (`k_halo_barrier`: each rank stores its sum into every rank's signal area and waits for the others'), the received half
    is committed and the convergence test runs on the device.  Without it (gloo tests, `p2p_bytes=None`): per sweep one
    C-ABI call, one grouped send/recv of the boundary slots (`halo_exchange`) and one all-reduce.
    The reference's sequential order is a diagonal wavefront that would idle most strips, so distributed
    results are compared with the single-GPU engine in JACOBI mode (same kernels, batch-independent reduction splits).
  * Two-site gates whose vertices are both owned run locally.  A gate across a cut is listed by BOTH ranks: each absorbs
    and QR-reduces its own site (`tnqs_apply_2site_begin`), the two small R factors are exchanged (one grouped send/recv
    per colour class), and both run the identical deterministic theta / SVD / truncation and finish their own half
    (`tnqs_apply_2site_end`).

The compute on each rank is a callback (`LocalCompute`), so the same driver runs the CUDA engine (tests -m gpu, bench)
or the CPU oracle (gloo tests of the host logic).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .graphs import Graph


def strip_owner(g: Graph, world: int, axis: int = 1) -> np.ndarray:
    """Contiguous strips along `axis` of a grid graph with tuple names (1-based coordinates)."""
    coords = np.array([n[axis] for n in g.names])
    lo, hi = coords.min(), coords.max()
    n = hi - lo + 1
    return np.minimum(((coords - lo) * world) // n, world - 1).astype(np.int32)


@dataclass
class LocalGraph:
    g: Graph                      # local graph (owned first? no: global order kept)
    l2g_v: np.ndarray             # local vertex -> global vertex
    l2g_e: np.ndarray             # local edge -> global edge
    g2l_v: Dict[int, int]
    g2l_e: Dict[int, int]
    owned: np.ndarray             # bool per local vertex
    owner_local: np.ndarray       # int32 per local vertex (global owner rank)
    send: Dict[int, np.ndarray]   # neighbour rank -> local directed edge ids (owned -> ghost), ordered by global directed id
    recv: Dict[int, np.ndarray]   # neighbour rank -> local directed edge ids (ghost -> owned), ordered by global directed id

    def to_local_dedge(self, kg: int, gglobal: Graph) -> int:
        e, d = divmod(int(kg), 2)
        le = self.g2l_e[e]
        # local edge keeps (esrc, edst) orientation of the global edge
        return 2 * le + d


def build_local_graph(gg: Graph, owner: np.ndarray, rank: int) -> LocalGraph:
    owned_g = [v for v in range(gg.nv) if owner[v] == rank]
    ghosts = sorted({gg.other(e, v) for v in owned_g for e in gg.incident[v]} - set(owned_g))
    verts = sorted(owned_g + ghosts)
    g2l_v = {v: i for i, v in enumerate(verts)}
    edges = [e for e in range(gg.ne) if owner[int(gg.esrc[e])] == rank or owner[int(gg.edst[e])] == rank]
    g2l_e = {e: i for i, e in enumerate(edges)}
    lg = Graph([gg.names[v] for v in verts], [g2l_v[int(gg.esrc[e])] for e in edges], [g2l_v[int(gg.edst[e])] for e in edges])
    owned = np.array([owner[v] == rank for v in verts])
    send: Dict[int, List[Tuple[int, int]]] = {}
    recv: Dict[int, List[Tuple[int, int]]] = {}
    for e in edges:
        a, b = int(gg.esrc[e]), int(gg.edst[e])
        for d, (u, w) in enumerate(((a, b), (b, a))):            # directed u -> w, global id 2e + d
            kg = 2 * e + d
            if owner[u] == rank and owner[w] != rank:
                send.setdefault(int(owner[w]), []).append((kg, 2 * g2l_e[e] + d))
            elif owner[u] != rank and owner[w] == rank:
                recv.setdefault(int(owner[u]), []).append((kg, 2 * g2l_e[e] + d))
    fin = lambda m: {r: np.array([l for _, l in sorted(v)], dtype=np.int32) for r, v in m.items()}
    return LocalGraph(lg, np.array(verts, dtype=np.int64), np.array(edges, dtype=np.int64), g2l_v, g2l_e, owned,
                      np.array([owner[v] for v in verts], dtype=np.int32), fin(send), fin(recv))


class LocalCompute:
    """What a rank must be able to do; implemented by EngineCompute (CUDA) and, in the gloo tests, by an oracle adapter."""

    def sweep(self, seq: np.ndarray, want_diff: bool) -> float:          # one JACOBI sweep over `seq`; returns sum of diffs
        raise NotImplementedError

    def pack(self, ks: np.ndarray):                                      # -> torch tensor [len(ks), slot] complex128
        raise NotImplementedError

    def unpack(self, ks: np.ndarray, buf) -> None:
        raise NotImplementedError


class EngineCompute(LocalCompute):
    def __init__(self, bpc, device):
        import torch
        from . import engine as E
        self.E, self.torch, self.bpc, self.device = E, torch, bpc, device
        self.slot = bpc.psi.chi_cap ** 2

    def sweep(self, seq, want_diff):
        info = {}
        self.E.update(self.bpc, maxiter=1, tolerance=0.0 if want_diff else None, edge_sequence=seq, mode="jacobi", inplace=True, info=info)
        return float(info["mean_diff"]) * len(seq) if want_diff else 0.0

    def pack(self, ks):
        from . import lib
        buf = self.torch.empty((len(ks), self.slot), dtype=self.torch.complex128, device=self.device)
        k = lib.i32(ks)
        lib.check(self.bpc.psi._h.ptr, lib.load().tnqs_halo_pack(self.bpc.psi._h.ptr, len(k), lib.ptr_i32(k), buf.data_ptr()))
        return buf

    def unpack(self, ks, buf):
        from . import lib
        k = lib.i32(ks)
        self.torch.cuda.current_stream().synchronize()
        lib.check(self.bpc.psi._h.ptr, lib.load().tnqs_halo_unpack(self.bpc.psi._h.ptr, len(k), lib.ptr_i32(k), buf.data_ptr()))

    # ---- peer-memory halo (tnqs_halo_*) ----
    def halo_alloc(self, world, recv_k):
        """This rank's halo block for the ghost -> owned messages `recv_k`; returns (64-byte CUDA IPC handle, device pointer)."""
        import ctypes as C
        from . import lib
        k = lib.i32(recv_k)
        handle = (C.c_char * 64)()
        ptr = C.c_void_p()
        lib.check(self.bpc.psi._h.ptr, lib.load().tnqs_halo_alloc(self.bpc.psi._h.ptr, int(world), len(k), lib.ptr_i32(k), C.cast(handle, C.c_void_p), C.byref(ptr)))
        return bytes(handle), int(ptr.value or 0)

    def halo_peer(self, peer, nrecv, handle=None, ptr=None):
        import ctypes as C
        from . import lib
        hb = C.create_string_buffer(handle, 64) if handle is not None else None
        lib.check(self.bpc.psi._h.ptr, lib.load().tnqs_halo_peer(self.bpc.psi._h.ptr, int(peer), C.cast(hb, C.c_void_p) if hb is not None else None,
                                                                  C.c_void_p(ptr) if ptr is not None else None, int(nrecv)))

    def halo_route(self, ks, peers, slots):
        from . import lib
        k, p, s = lib.i32(ks), lib.i32(peers), lib.i32(slots)
        lib.check(self.bpc.psi._h.ptr, lib.load().tnqs_halo_route(self.bpc.psi._h.ptr, len(k), lib.ptr_i32(k), lib.ptr_i32(p), lib.ptr_i32(s)))

    def refresh(self, seq, maxiter, tolerance):
        """A whole partitioned BP refresh in one call (needs the peer-memory halo): returns (sweeps, global mean diff)."""
        info = {}
        self.E.update(self.bpc, maxiter=maxiter, tolerance=tolerance, edge_sequence=seq, mode="jacobi", inplace=True, info=info)
        return info["iters"], info["mean_diff"]


def halo_exchange(lgraph: LocalGraph, comp: LocalCompute, dist, slot: int, torch, device) -> int:
    """Send every owned->ghost message to the ghost's owner, receive every ghost->owned message.  Returns bytes sent."""
    ops, recv_bufs, sent = [], {}, 0
    for r in sorted(set(lgraph.send) | set(lgraph.recv)):
        if r in lgraph.recv:
            recv_bufs[r] = torch.empty((len(lgraph.recv[r]), slot), dtype=torch.complex128, device=device)
            ops.append(dist.P2POp(dist.irecv, recv_bufs[r], r))
        if r in lgraph.send:
            sb = comp.pack(lgraph.send[r])
            sent += sb.numel() * 16
            ops.append(dist.P2POp(dist.isend, sb, r))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for r, buf in recv_bufs.items():
        comp.unpack(lgraph.recv[r], buf)
    return sent


def halo_layout(lgraph: LocalGraph):
    """Order of this rank's receive buffer: neighbours ascending, inside a neighbour ascending global directed edge id (the
    order of lgraph.recv[r]).  Returns (local directed edges per slot, {neighbour: first slot})."""
    recv_k, first = [], {}
    for r in sorted(lgraph.recv):
        first[r] = len(recv_k)
        recv_k += [int(k) for k in lgraph.recv[r]]
    return np.array(recv_k, dtype=np.int32), first


def halo_routes(lgraph: LocalGraph, rank: int, first_of_peer: Dict[int, Dict[int, int]]):
    """Where each owned -> ghost message goes: (local directed edges, peer ranks, slots in the peer's receive buffer).
    first_of_peer[p][rank] = first slot of this rank's block in p's buffer; both sides order a block by global directed id."""
    ks, peers, slots = [], [], []
    for p in sorted(lgraph.send):
        base = first_of_peer[p][rank]
        for j, k in enumerate(lgraph.send[p]):
            ks.append(int(k)); peers.append(int(p)); slots.append(base + j)
    return ks, peers, slots


def halo_setup_p2p(lgraph: LocalGraph, comp, dist, rank: int, world: int) -> int:
    """Open the peer-memory halo between the processes of one node: every rank allocates its block, the CUDA IPC handles and
    buffer layouts are all-gathered, every rank opens every other rank's block (the barrier signals go all-to-all) and routes
    its boundary messages into its neighbours' receive buffers.  Returns the bytes this rank stores in peers per sweep."""
    recv_k, first = halo_layout(lgraph)
    handle, _ = comp.halo_alloc(world, recv_k)
    mine = {"handle": handle, "nrecv": int(len(recv_k)), "first": {int(r): int(o) for r, o in first.items()}}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    for p in range(world):
        if p != rank:
            comp.halo_peer(p, everyone[p]["nrecv"], handle=everyone[p]["handle"])
    ks, peers, slots = halo_routes(lgraph, rank, {p: everyone[p]["first"] for p in lgraph.send})
    comp.halo_route(ks, peers, slots)
    dist.barrier()
    return 16 * comp.slot * len(ks)


def update_distributed(lgraph: LocalGraph, comp: LocalCompute, dist, slot: int, torch, device, maxiter: int, tolerance=None,
                       info: dict | None = None, p2p_bytes: int | None = None) -> None:
    """JACOBI BP over the partitioned network.  p2p_bytes = None: per sweep one local C-ABI call, one halo exchange
    (send/recv), one all-reduce.  p2p_bytes = result of halo_setup_p2p: the whole refresh is one C-ABI call (one graph launch
    per rank; boundary stores, barrier, all-reduce and convergence test on the device)."""
    g = lgraph.g
    seq = getattr(lgraph, "_owned_seq", None)           # directed edges leaving an owned vertex, ascending: built once per local graph
    if seq is None:
        seq = np.array([k for k in range(2 * g.ne) if lgraph.owned[g.dsrc(k)]], dtype=np.int32)
        lgraph._owned_seq = seq
    iters, mean, halo_bytes = 0, float("nan"), 0
    if p2p_bytes is not None:
        iters, mean = comp.refresh(seq, maxiter, tolerance)
        if info is not None:
            info.update(iters=iters, mean_diff=mean, halo_bytes=iters * p2p_bytes, messages_per_sweep=len(seq))
        return
    for it in range(1, maxiter + 1):
        iters = it
        s = comp.sweep(seq, tolerance is not None)
        halo_bytes += halo_exchange(lgraph, comp, dist, slot, torch, device)
        if tolerance is not None:
            t = torch.tensor([s, float(len(seq))], dtype=torch.float64, device=device)
            dist.all_reduce(t)
            mean = float(t[0] / t[1])
            if mean <= tolerance:
                break
    if info is not None:
        info.update(iters=iters, mean_diff=mean, halo_bytes=halo_bytes, messages_per_sweep=len(seq))


def class_gates_local(gg: Graph, lgraph: LocalGraph, owner: np.ndarray, rank: int, edges_global: Sequence[int]):
    """Split one colour class (global edge ids, first vertex = esrc) into this rank's batch: local edge ids, remote flags
    (0 both owned, 1 first vertex remote, 2 second vertex remote) and, per neighbour rank, the positions of the shared
    cut gates in batch order (both sides enumerate them by ascending global edge id)."""
    es, rm, cut_with, ncut = [], [], {}, 0
    for e in sorted(int(x) for x in edges_global):
        a, b = int(gg.esrc[e]), int(gg.edst[e])
        oa, ob = int(owner[a]), int(owner[b])
        if oa != rank and ob != rank:
            continue
        es.append(lgraph.g2l_e[e])
        flag = 0 if (oa == rank and ob == rank) else (1 if oa != rank else 2)
        rm.append(flag)
        if flag:
            cut_with.setdefault(ob if flag == 2 else oa, []).append(ncut)
            ncut += 1
    return np.array(es, dtype=np.int32), np.array(rm, dtype=np.int32), cut_with


def apply_class_distributed(E, bpc, gg: Graph, lgraph: LocalGraph, owner, rank, edges_global, gate, apply_kwargs, dist, torch, device,
                            batch=None):
    """One colour class of identical two-site gates over the partitioned network (tnqs_apply_2site_begin / _end + one
    grouped exchange of R factors).  `batch` = a precomputed class_gates_local(...) result.  Returns (ranks, truncation
    errors) of this rank's batch."""
    es_all, rm_all, cut_with = batch if batch is not None else class_gates_local(gg, lgraph, owner, rank, edges_global)
    if len(es_all) == 0:
        return np.zeros(0, dtype=np.int32), np.zeros(0)
    ranks_all, errs_all = np.zeros(len(es_all), dtype=np.int32), np.zeros(len(es_all))
    # gates with both vertices on this rank need no exchange: they go through tnqs_apply_2site, which processes a batch in
    # memory-bounded chunks (a whole class of chi = 32 gates holds ~5 copies of 2 x 32 MiB per gate); only the cut-crossing gates
    # take the two-stage path, all of them in one exchange (their positions in `cut_with` count cut gates only)
    inner = np.flatnonzero(rm_all == 0)
    if len(inner):
        r_in, e_in, _ = E.apply_2site_batch(bpc, es_all[inner], [gate] * len(inner), apply_kwargs=apply_kwargs)
        ranks_all[inner], errs_all[inner] = r_in, e_in
    cutpos = np.flatnonzero(rm_all != 0)
    if len(cutpos) == 0:
        return ranks_all, errs_all
    es, rm = es_all[cutpos], rm_all[cutpos]
    send_ptr, ncut, slot = E.apply_2site_begin(bpc, es, [gate] * len(es), remote=rm, apply_kwargs=apply_kwargs)
    recv = None
    if ncut:
        send = _wrap_device(torch, send_ptr, ncut * slot, device)
        recv = torch.empty(ncut * slot, dtype=torch.complex128, device=device)
        ops = []
        for r, pos in sorted(cut_with.items()):
            idx = torch.as_tensor(pos, device=device)
            sbuf = send.view(ncut, slot)[idx].contiguous()
            rbuf = torch.empty_like(sbuf)
            ops.append((r, idx, sbuf, rbuf))
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, sb, r) for r, _, sb, _ in ops] + [dist.P2POp(dist.irecv, rb, r) for r, _, _, rb in ops])
        for q in reqs:
            q.wait()
        for r, idx, _, rb in ops:
            recv.view(ncut, slot)[idx] = rb
        torch.cuda.synchronize()
    ranks, errs, _ = E.apply_2site_end(bpc, len(es), recv.data_ptr() if recv is not None else None)
    ranks_all[cutpos], errs_all[cutpos] = ranks, errs
    return ranks_all, errs_all


def _wrap_device(torch, ptr: int, n_complex: int, device):
    """View n_complex ComplexF64 at a raw device pointer as a torch tensor (no copy)."""
    class _Arr:
        __cuda_array_interface__ = {"shape": (2 * n_complex,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    return torch.view_as_complex(torch.as_tensor(_Arr(), device=device).view(n_complex, 2))
