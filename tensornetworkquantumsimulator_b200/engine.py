"""Host-side mirror of the reference interface for the hot path, over libtnqs_cuda.so.

Same names, argument meaning and error behaviour as the Julia API it stands in for:
`tensornetworkstate`, `random_tensornetworkstate` (src/TensorNetworks/tensornetworkstate.jl:82-144),
`BeliefPropagationCache`, `update`, `message(s)`, `partitionfunction`, `rescale`
(src/MessagePassing/*.jl), `apply_gates` / `apply_circuit` (src/Apply/apply_gates.jl:13-133),
`expect(...; alg="bp")` (src/expect.jl:53-136), `truncate` (src/truncate.jl:12-38,74-79),
`symmetric_gauge`, `gauge_and_scale`, `entanglement` (src/symmetric_gauge.jl),
`loopcorrected_partitionfunction`, `norm_sqr` / `inner` with alg="loopcorrections" (src/MessagePassing/loopcorrection.jl,
src/norm_sqr.jl:56-76, src/inner.jl:67-86), `loop_correlation(s)` (src/MessagePassing/beliefpropagationcache.jl:144-197).

Non-mutating functions copy on entry and return new objects like the reference (`copy(bpc)`,
apply_gates.jl:51, abstractbeliefpropagationcache.jl:206); pass `inplace=True` to skip the device copy.
All arithmetic runs in the CUDA library; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
from typing import List, Optional, Sequence

import numpy as np

from . import gates as _gates
from . import lib as _lib
from .graphs import (Graph, edge_color as _edge_color, forest_cover_edge_sequence, steiner_tree as _steiner_tree, tree_schedule as _tree_schedule,
                     edgeinduced_subgraphs_no_leaves as _subgraphs_no_leaves, cycle_walks as _cycle_walks)
from .lib import ApplyOpts, BpOpts, TnqsError, check, i32, ptr_f64, ptr_i32, ptr_i64

DEFAULT_BP_MAXITER = 25      # src/MessagePassing/beliefpropagationcache.jl:108
DEFAULT_TOLERANCE = 1.0e-8   # ComplexF64, src/MessagePassing/beliefpropagationcache.jl:109-112


class _Handle:
    """Owns one tnqs_ctx."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr:
                _lib.load().tnqs_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _create(g: Graph, d: Sequence[int], chi: Sequence[int], chi_cap: int, device: int) -> _Handle:
    lib = _lib.load()
    out = C.c_void_p()
    esrc, edst, dd, cc = i32(g.esrc), i32(g.edst), i32(d), i32(chi if len(chi) else [1])
    st = lib.tnqs_create(C.byref(out), g.nv, g.ne, ptr_i32(esrc), ptr_i32(edst), ptr_i32(dd), ptr_i32(cc), int(chi_cap), int(device))
    if st != 0:
        raise TnqsError(st, (lib.tnqs_last_error(None) or b"").decode())
    return _Handle(out)


class TensorNetworkState:
    """Site tensors on the device (struct TensorNetworkState, src/TensorNetworks/tensornetworkstate.jl:4-7)."""

    def __init__(self, g: Graph, handle: _Handle, chi_cap: int, device: int = 0):
        self.g, self._h, self.chi_cap, self.device = g, handle, chi_cap, device
        if not hasattr(g, "_default_seq"):
            g._default_seq = None

    # ---- construction helpers ----
    @staticmethod
    def from_arrays(g: Graph, tensors: Sequence[np.ndarray], chi_cap: Optional[int] = None, device: int = 0) -> "TensorNetworkState":
        chi = np.ones(max(g.ne, 1), dtype=np.int32)
        for v, t in enumerate(tensors):
            if t.ndim != 1 + g.degree(v):
                raise ValueError(f"tensor of vertex {v} has rank {t.ndim}, expected {1 + g.degree(v)}")
            for i, e in enumerate(g.incident[v]):
                chi[e] = t.shape[1 + i]
        cap = int(chi_cap if chi_cap is not None else max(1, chi.max()))
        h = _create(g, [t.shape[0] for t in tensors], chi[:g.ne], cap, device)
        psi = TensorNetworkState(g, h, cap, device)
        psi.set_sites(range(g.nv), tensors)
        return psi

    def set_sites(self, vs, tensors) -> None:
        """Several site tensors in one call (tnqs_set_sites: copies queued back to back, one synchronisation)."""
        vs = i32(list(vs))
        flats = [np.asfortranarray(t, dtype=np.complex128).ravel(order="F").view(np.float64) for t in tensors]
        ptrs = (C.POINTER(C.c_double) * len(flats))(*[ptr_f64(f) for f in flats])
        ns = np.asarray([f.size // 2 for f in flats], dtype=np.int64)
        check(self._h.ptr, _lib.load().tnqs_set_sites(self._h.ptr, len(flats), ptr_i32(vs), ptrs, _lib.ptr_i64(ns)))

    def set_site(self, v: int, t: np.ndarray) -> None:
        a = np.asfortranarray(t, dtype=np.complex128)
        flat = a.ravel(order="F").view(np.float64)
        check(self._h.ptr, _lib.load().tnqs_set_site(self._h.ptr, int(v), ptr_f64(flat), a.size))

    def __getitem__(self, v) -> np.ndarray:
        """Site tensor as ndarray [s, b_e1..b_ez] (getindex); v is a vertex name (see _vid)."""
        v = _vid(self.g, v)
        lib = _lib.load()
        n = lib.tnqs_site_size(self._h.ptr, v)
        buf = np.empty(2 * n, dtype=np.float64)
        check(self._h.ptr, lib.tnqs_get_site(self._h.ptr, v, ptr_f64(buf), n))
        chi = self.bond_dims()
        shape = (n // int(np.prod([chi[e] for e in self.g.incident[v]], dtype=np.int64)),) + tuple(int(chi[e]) for e in self.g.incident[v])
        return buf.view(np.complex128).reshape(shape, order="F")

    def tensors(self) -> List[np.ndarray]:
        """Every site tensor (tnqs_get_sites: one call, one synchronisation)."""
        lib = _lib.load()
        nv = self.g.nv
        ns = np.asarray([lib.tnqs_site_size(self._h.ptr, v) for v in range(nv)], dtype=np.int64)
        bufs = [np.empty(2 * int(n), dtype=np.float64) for n in ns]
        ptrs = (C.POINTER(C.c_double) * nv)(*[ptr_f64(b) for b in bufs])
        vs = i32(np.arange(nv))
        check(self._h.ptr, lib.tnqs_get_sites(self._h.ptr, nv, ptr_i32(vs), ptrs, _lib.ptr_i64(ns)))
        chi = self.bond_dims()
        out = []
        for v in range(nv):
            dims = tuple(int(chi[e]) for e in self.g.incident[v])
            out.append(bufs[v].view(np.complex128).reshape((int(ns[v]) // int(np.prod(dims, dtype=np.int64)),) + dims, order="F"))
        return out

    def bond_dims(self) -> np.ndarray:
        chi = np.ones(max(self.g.ne, 1), dtype=np.int32)
        check(self._h.ptr, _lib.load().tnqs_get_bond_dims(self._h.ptr, ptr_i32(chi)))
        return chi[:self.g.ne]

    def maxvirtualdim(self) -> int:
        return int(self.bond_dims().max()) if self.g.ne else 1

    def site_dims(self) -> np.ndarray:
        """Site dimension of every vertex (2 for qubits, 4 for "Pauli" sites); fixed for the life of the handle."""
        if not hasattr(self, "_site_dims"):
            lib = _lib.load()
            chi = self.bond_dims()
            self._site_dims = np.asarray(
                [lib.tnqs_site_size(self._h.ptr, v) // int(np.prod([chi[e] for e in self.g.incident[v]], dtype=np.int64))
                 for v in range(self.g.nv)], dtype=np.int32)
        return self._site_dims

    def copy(self) -> "TensorNetworkState":
        out = C.c_void_p()
        st = _lib.load().tnqs_clone(self._h.ptr, C.byref(out))
        if st != 0:
            raise TnqsError(st, (_lib.load().tnqs_last_error(None) or b"").decode())
        new = TensorNetworkState(self.g, _Handle(out), self.chi_cap, self.device)
        if hasattr(self, "_site_dims"):
            new._site_dims = self._site_dims
        return new

    def kernel_launches(self) -> int:
        return int(_lib.load().tnqs_kernel_launches(self._h.ptr))


def maxvirtualdim(x) -> int:
    return network(x).maxvirtualdim()


def tensornetworkstate(f, g: Graph, chi_cap: int = 1, device: int = 0) -> TensorNetworkState:
    """Product state with all bonds of dimension 1 (src/TensorNetworks/tensornetworkstate.jl:124-144).
    `chi_cap` = the largest bond dimension the state may grow to (message slot size on the device)."""
    ts = []
    for v in range(g.nv):
        vec = _gates.state_vector(f(g.names[v]))
        ts.append(vec.reshape((vec.shape[0],) + (1,) * g.degree(v)))
    return TensorNetworkState.from_arrays(g, ts, chi_cap=chi_cap, device=device)


def zerostate(g: Graph, chi_cap: int = 1, device: int = 0) -> TensorNetworkState:
    """All spins up (src/TensorNetworks/tensornetworkstate_constructors.jl:8-12)."""
    return tensornetworkstate(lambda v: "↑", g, chi_cap=chi_cap, device=device)


def paulitensornetworkstate(f, g: Graph, chi_cap: int = 1, device: int = 0) -> TensorNetworkState:
    """Operator in the Heisenberg picture as a product state over "Pauli" sites (d = 4, basis I, X, Y, Z): f maps a vertex to
    "I", "X", "Y" or "Z" (src/TensorNetworks/tensornetworkstate_constructors.jl:19-24).  Gates on such a state are Pauli
    transfer matrices (gates.heisenberg_gate_array)."""
    return tensornetworkstate(lambda v: _gates.PAULI_STATES[f(v)], g, chi_cap=chi_cap, device=device)


def identitytensornetworkstate(g: Graph, chi_cap: int = 1, device: int = 0) -> TensorNetworkState:
    """The identity operator on Pauli sites (src/TensorNetworks/tensornetworkstate_constructors.jl:31-35)."""
    return paulitensornetworkstate(lambda v: "I", g, chi_cap=chi_cap, device=device)


def random_tensornetworkstate(g: Graph, bond_dimension: int = 1, seed: int = 0, d: int = 2,
                              chi_cap: Optional[int] = None, device: int = 0) -> TensorNetworkState:
    """N(0,1)+iN(0,1) entries generated on the device with the counter-based generator (rng.py has the same
    formula); src/TensorNetworks/tensornetworkstate.jl:82-92."""
    cap = int(chi_cap if chi_cap is not None else bond_dimension)
    h = _create(g, [d] * g.nv, [bond_dimension] * g.ne, cap, device)
    psi = TensorNetworkState(g, h, cap, device)
    check(h.ptr, _lib.load().tnqs_fill_random(h.ptr, int(seed)))
    return psi


class BeliefPropagationCache:
    """Network + one message per directed edge (src/MessagePassing/beliefpropagationcache.jl:9-13), messages at their default.
    On the device the messages live in the same handle as the site tensors, so the public constructor wraps a CLONE of psi:
    a second cache over the same state does not clobber the messages of the first, and in-place operations on the cache never
    reach the caller's state (the reference's constructor does not alias either).  `_share=True` (internal; callers that
    have just made their own copy, or that want the state to follow the cache) wraps psi's handle itself."""

    def __init__(self, psi: TensorNetworkState, _keep_messages: bool = False, _share: bool = False):
        self.psi = psi if _share else psi.copy()
        if not _keep_messages:
            check(self.psi._h.ptr, _lib.load().tnqs_reset_messages(self.psi._h.ptr))

    @property
    def g(self) -> Graph:
        return self.psi.g

    def copy(self) -> "BeliefPropagationCache":
        out = BeliefPropagationCache(self.psi.copy(), _keep_messages=True, _share=True)
        bra = getattr(self.psi, "_bra", None)
        if bra is not None:
            _attach_bra(out, bra)
        return out


def network(x) -> TensorNetworkState:
    return x.psi if isinstance(x, BeliefPropagationCache) else x


def _int_named(g: Graph) -> bool:
    flag = getattr(g, "_int_named", None)
    if flag is None:
        flag = any(isinstance(n, (int, np.integer)) and not isinstance(n, bool) for n in g.names)
        g._int_named = flag
    return flag


def _vid(g: Graph, x) -> int:
    """Vertex NAME -> dense 0-based id.  The reference resolves vertices by name only (graph_ops.jl:53-67 keeps whatever labels
    the circuit uses, e.g. [("Rx", [3], ..), ("CPHASE", [3, 1], ..)]): on a graph with integer names an integer is therefore
    always a name, never an id.  Graphs whose names are tuples / strings additionally accept a dense id (internal callers and
    the benchmarks address vertices that way)."""
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        if _int_named(g):
            try:
                return g.index[int(x)]
            except KeyError:
                raise KeyError(f"vertex {int(x)} is not in the graph (integer-named graphs are addressed by name)") from None
        return int(x)
    return g.v(x)


def message(bpc: BeliefPropagationCache, edge) -> np.ndarray:
    """Message of the directed edge (src => dst names, or a directed edge id) as ndarray m[b, b']."""
    g = bpc.g
    k = int(edge) if isinstance(edge, (int, np.integer)) else g.dedge(_vid(g, edge[0]), _vid(g, edge[1]))
    chi = int(bpc.psi.bond_dims()[k // 2])
    buf = np.empty(2 * chi * chi, dtype=np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_get_message(bpc.psi._h.ptr, k, ptr_f64(buf), chi * chi))
    return buf.view(np.complex128).reshape((chi, chi), order="F")


def setmessage(bpc: BeliefPropagationCache, k: int, m: np.ndarray) -> None:
    a = np.asfortranarray(m, dtype=np.complex128).ravel(order="F").view(np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_set_message(bpc.psi._h.ptr, int(k), ptr_f64(a), m.size))


def messages(bpc: BeliefPropagationCache) -> List[np.ndarray]:
    return [message(bpc, k) for k in range(2 * bpc.g.ne)]


def default_edge_sequence(g: Graph) -> np.ndarray:
    if getattr(g, "_default_seq", None) is None:
        g._default_seq = forest_cover_edge_sequence(g)
    return g._default_seq


def default_bp_update_kwargs(x) -> dict:
    """src/MessagePassing/beliefpropagationcache.jl:114-121."""
    g = network(x).g
    return dict(maxiter=1 if g.is_tree() else DEFAULT_BP_MAXITER, tolerance=DEFAULT_TOLERANCE)


def _bp_opts(g: Graph, maxiter=None, tolerance=None, normalize=True, enforce_hermiticity=False, mode="reference") -> BpOpts:
    if maxiter is None:
        maxiter = 1 if g.is_tree() else DEFAULT_BP_MAXITER           # default_bp_maxiter, beliefpropagationcache.jl:31
    m = {"reference": _lib.TNQS_BP_REFERENCE_ORDER, "jacobi": _lib.TNQS_BP_JACOBI}[mode]
    return BpOpts(int(maxiter), -1.0 if tolerance is None else float(tolerance), int(normalize), int(enforce_hermiticity), m)


def update(bpc: BeliefPropagationCache, maxiter: Optional[int] = None, tolerance: Optional[float] = None,
           edge_sequence: Optional[Sequence[int]] = None, normalize: bool = True, enforce_hermiticity: bool = False,
           mode: str = "reference", verbose: bool = False, inplace: bool = False, info: Optional[dict] = None) -> BeliefPropagationCache:
    """BP sweeps until `maxiter` or mean(1 - fidelity) <= tolerance
    (src/MessagePassing/abstractbeliefpropagationcache.jl:201-222; a bare update has tolerance=None)."""
    out = bpc if inplace else bpc.copy()
    g = out.g
    seq = i32(default_edge_sequence(g) if edge_sequence is None else edge_sequence)
    opts = _bp_opts(g, maxiter, tolerance, normalize, enforce_hermiticity, mode)
    iters, mean = C.c_int32(0), C.c_double(float("nan"))
    check(out.psi._h.ptr, _lib.load().tnqs_bp_update(out.psi._h.ptr, ptr_i32(seq), len(seq), C.byref(opts), C.byref(iters), C.byref(mean)))
    if verbose and tolerance is not None and mean.value <= tolerance:
        print(f"BP converged to desired precision after {iters.value} iterations.")
    if info is not None:
        info["iters"], info["mean_diff"] = iters.value, mean.value
    return out


def _apply_opts(apply_kwargs: Optional[dict]) -> ApplyOpts:
    kw = dict(apply_kwargs or {})
    maxdim = kw.pop("maxdim", None)
    cutoff = kw.pop("cutoff", None)
    normalize_tensors = kw.pop("normalize_tensors", True)
    if kw:
        raise TypeError(f"unknown apply_kwargs {sorted(kw)}")
    return ApplyOpts(0 if maxdim is None else int(maxdim), float("nan") if cutoff is None else float(cutoff), int(normalize_tensors))


def _pack_circuit(g: Graph, circuit: Sequence[tuple], site_dims=None):
    """Gate tuples -> the flat arrays of tnqs_apply_circuit.  site_dims (per vertex, or a callable returning them) selects the Heisenberg branch of
    `toitensor` for gates on Pauli sites (gate_definitions.jl:41-42)."""
    gv = np.full(2 * len(circuit), -1, dtype=np.int32)
    off = np.zeros(len(circuit), dtype=np.int64)
    chunks, pos = [], 0
    for i, gate in enumerate(circuit):
        verts = gate[1] if isinstance(gate[1], list) else [gate[1]]
        if len(verts) not in (1, 2):
            raise ValueError("gates act on one or two vertices")
        if isinstance(gate[0], np.ndarray):
            arr = gate[0]
        else:
            if callable(site_dims):
                site_dims = site_dims()                      # looked up only when a gate is given by name
            arr = _gates.gate_array(gate, None if site_dims is None else [site_dims[_vid(g, x)] for x in verts])
        for j, x in enumerate(verts):
            gv[2 * i + j] = _vid(g, x)
        flat = np.asarray(arr, dtype=np.complex128).ravel(order="F")
        off[i] = pos
        pos += flat.size
        chunks.append(flat)
    data = np.concatenate(chunks) if chunks else np.zeros(1, dtype=np.complex128)
    return gv, off, data.view(np.float64)


def apply_gates(circuit: Sequence[tuple], psi, apply_kwargs: Optional[dict] = None, bp_update_kwargs: Optional[dict] = None,
                update_cache: bool = True, verbose: bool = False, inplace: bool = False, info: Optional[dict] = None, fetch=None):
    """Apply an ordered list of one/two-site gates by simple update with lazy BP refreshes
    (src/Apply/apply_gates.jl:13-23 for a TensorNetworkState, :42-92 for a cache).  Returns
    (state or cache, truncation errors).  fetch = (vertex ids int32, ctypes array of double*, capacities int64): the listed site
    tensors are copied into those host buffers while the closing BP refresh runs (tnqs_apply_circuit_fetch); the int64 array
    holds the element counts afterwards."""
    if isinstance(psi, TensorNetworkState):
        kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(psi)
        bpc = update(BeliefPropagationCache(psi if inplace else psi.copy(), _share=True), inplace=True, **kw)
        bpc, errs = apply_gates(circuit, bpc, apply_kwargs, kw, update_cache, verbose, True, info, fetch)
        return bpc.psi, errs
    out = psi if inplace else psi.copy()
    g = out.g
    kw = dict(bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(out))
    kw.pop("verbose", None)
    seq = i32(kw.pop("edge_sequence", None) if kw.get("edge_sequence") is not None else default_edge_sequence(g))
    kw.pop("edge_sequence", None)
    bopts = _bp_opts(g, **kw)
    aopts = _apply_opts(apply_kwargs)
    gv, off, data = _pack_circuit(g, circuit, out.psi.site_dims)
    errs = np.zeros(max(len(circuit), 1), dtype=np.float64)
    refreshes = C.c_int32(0)
    if fetch is None:
        check(out.psi._h.ptr, _lib.load().tnqs_apply_circuit(
            out.psi._h.ptr, len(circuit), ptr_i32(gv), ptr_i64(off), ptr_f64(data), C.byref(aopts), ptr_i32(seq), len(seq),
            C.byref(bopts), int(update_cache), ptr_f64(errs), C.byref(refreshes)))
    else:
        fv, fptrs, fns = fetch
        check(out.psi._h.ptr, _lib.load().tnqs_apply_circuit_fetch(
            out.psi._h.ptr, len(circuit), ptr_i32(gv), ptr_i64(off), ptr_f64(data), C.byref(aopts), ptr_i32(seq), len(seq),
            C.byref(bopts), int(update_cache), ptr_f64(errs), C.byref(refreshes), len(fv), ptr_i32(fv), fptrs, ptr_i64(fns)))
    if info is not None:
        info["refreshes"] = refreshes.value
    return out, errs[:len(circuit)]


apply_circuit = apply_gates      # src/Apply/apply_gates.jl:133


def circuit_schedule(g: Graph, circuit: Sequence[tuple], update_cache: bool = True) -> List[tuple]:
    """The batches and BP refreshes `apply_gates` / tnqs_apply_circuit execute for an ordered gate list, by the reference's
    rule (src/Apply/apply_gates.jl:56-89): a refresh before a 2-site gate that touches a vertex touched since the last
    refresh (1-site gates mark vertices too), consecutive independent gates of equal arity form one batch, final refresh.
    Returns [("refresh",) | ("gates1", [gate indices], folds) | ("gates2", [gate indices])].  Runs are defined by the rule, not by
    colour: gates of a new colour class that precede its first overlapping gate are applied before the refresh.  Consecutive
    one-site gates on a vertex are applied as ONE gate, their product (`fold_gates`): folds[j] lists the later gates multiplied
    onto the j-th gate of the batch, in program order (empty for most)."""
    verts = []
    for gate in circuit:
        v = gate[1] if isinstance(gate[1], list) else [gate[1]]
        verts.append([_vid(g, x) for x in v])
    out: List[tuple] = []
    affected, inbatch, batch, arity_b = set(), set(), [], 0
    folds, pos1 = [], {}

    def flush():
        nonlocal batch, inbatch, folds, pos1
        if batch:
            out.append(("gates1", batch, folds) if arity_b == 1 else ("gates2", batch))
        batch, inbatch, folds, pos1 = [], set(), [], {}

    for i, vs in enumerate(verts):
        arity = len(vs)
        if update_cache and arity == 2 and any(v in affected for v in vs):
            flush(); out.append(("refresh",)); affected = set()
        if arity == 1 and batch and arity_b == 1 and vs[0] in pos1:
            folds[pos1[vs[0]]].append(i)
            continue
        if batch and (arity != arity_b or any(v in inbatch for v in vs)):
            flush()
        arity_b = arity
        batch.append(i); inbatch |= set(vs); affected |= set(vs)
        if arity == 1:
            folds.append([]); pos1[vs[0]] = len(batch) - 1
    flush()
    if update_cache:
        out.append(("refresh",))
    return out


def fold_gates(first: np.ndarray, later: Sequence[np.ndarray]) -> np.ndarray:
    """Product of consecutive one-site gates on a vertex, later ... first, as tnqs_apply_circuit forms it before applying them in
    one pass: the same scalar operations in the same order (o[s', s] <- sum_t g[s', t] o[t, s], t ascending), so that a schedule
    executed call by call reproduces apply_gates bit for bit."""
    o = np.array(first, dtype=np.complex128)
    d = o.shape[0]
    for g in later:
        g = np.asarray(g, dtype=np.complex128)
        new = np.zeros((d, d), dtype=np.complex128)
        for s2 in range(d):
            for t in range(d):
                for s1 in range(d):
                    a, b = complex(g[s1, t]), complex(o[t, s2])
                    p = complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)
                    new[s1, s2] = complex(new[s1, s2].real + p.real, new[s1, s2].imag + p.imag)
        o = new
    return o


def _pack_gate_arrays(gates: Sequence[np.ndarray]) -> np.ndarray:
    """Gate arrays o[s1', s2', s1, s2] (or o[s', s]) -> one flat float64 buffer, each gate column-major (the layout of the C-ABI).
    Vectorised: a colour class is typically one gate repeated, or gates of one shape."""
    n = len(gates)
    if n == 0:
        return np.zeros(2, dtype=np.float64)
    first = gates[0]
    if all(a is first for a in gates):
        flat = np.asarray(first, dtype=np.complex128).ravel(order="F")
        return np.tile(flat, n).view(np.float64)
    shapes = {np.shape(a) for a in gates}
    if len(shapes) == 1:
        arr = np.asarray(gates, dtype=np.complex128)                       # [n, ...]
        arr = arr.transpose((0,) + tuple(range(arr.ndim - 1, 0, -1)))      # reversed axes per gate = column-major order when flattened C-style
        return np.ascontiguousarray(arr).reshape(-1).view(np.float64)
    return np.concatenate([np.asarray(a, dtype=np.complex128).ravel(order="F") for a in gates]).view(np.float64)


def apply_2site_batch(bpc: BeliefPropagationCache, edges: Sequence[int], gates: Sequence[np.ndarray], flip: Optional[Sequence[int]] = None,
                      apply_kwargs: Optional[dict] = None):
    """In place: one vertex-disjoint batch of two-site gates (apply_gate!, src/Apply/apply_gates.jl:95-122).
    Returns (ranks, truncation errors, singular values [n, chi_cap])."""
    n = len(edges)
    es = i32(edges)
    fl = i32(flip if flip is not None else [0] * n)
    data = _pack_gate_arrays(gates)
    ranks = np.zeros(n, dtype=np.int32)
    errs = np.zeros(n, dtype=np.float64)
    sv = np.zeros((n, bpc.psi.chi_cap), dtype=np.float64)
    aopts = _apply_opts(apply_kwargs)
    check(bpc.psi._h.ptr, _lib.load().tnqs_apply_2site(bpc.psi._h.ptr, n, ptr_i32(es), ptr_i32(fl), ptr_f64(data), C.byref(aopts),
                                                        ptr_i32(ranks), ptr_f64(errs), ptr_f64(sv.reshape(-1))))
    return ranks, errs, sv


def apply_2site_begin(bpc: BeliefPropagationCache, edges: Sequence[int], gates: Sequence[np.ndarray], flip=None, remote=None,
                      apply_kwargs: Optional[dict] = None):
    """First stage of a batch with cut-crossing gates (tnqs_apply_2site_begin).  Returns (device pointer of the send buffer,
    number of cut gates, slot size in ComplexF64)."""
    n = len(edges)
    es, fl = i32(edges), i32(flip if flip is not None else [0] * n)
    rm = i32(remote if remote is not None else [0] * n)
    data = _pack_gate_arrays(gates)
    aopts = _apply_opts(apply_kwargs)
    send, ncut, slot = C.c_void_p(), C.c_int64(0), C.c_int64(0)
    check(bpc.psi._h.ptr, _lib.load().tnqs_apply_2site_begin(bpc.psi._h.ptr, n, ptr_i32(es), ptr_i32(fl), ptr_f64(data), C.byref(aopts),
                                                              ptr_i32(rm), C.byref(send), C.byref(ncut), C.byref(slot)))
    return send.value, ncut.value, slot.value


def apply_2site_end(bpc: BeliefPropagationCache, n: int, recv_ptr):
    ranks = np.zeros(n, dtype=np.int32)
    errs = np.zeros(n, dtype=np.float64)
    sv = np.zeros((n, bpc.psi.chi_cap), dtype=np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_apply_2site_end(bpc.psi._h.ptr, C.c_void_p(recv_ptr) if recv_ptr else None,
                                                            ptr_i32(ranks), ptr_f64(errs), ptr_f64(sv.reshape(-1))))
    return ranks, errs, sv


def apply_1site_batch(x, vertices: Sequence[int], gates: Sequence[np.ndarray], normalize_tensors: bool = True) -> None:
    psi = network(x)
    vs = i32(vertices)
    data = _pack_gate_arrays(gates)
    check(psi._h.ptr, _lib.load().tnqs_apply_1site(psi._h.ptr, len(vs), ptr_i32(vs), ptr_f64(data), int(normalize_tensors)))


# ---- measurements -------------------------------------------------------------------------------------------
def rdm_1site(bpc: BeliefPropagationCache, vertices: Sequence[int]) -> np.ndarray:
    """Un-normalised BP one-site reduced density matrices rho[i][s, s'] (src/rdm.jl:49-70 for single vertices)."""
    vs = i32(vertices)
    lib = _lib.load()
    # dmax is the largest site dimension of the network; the library pads every block to dmax x dmax
    dmax = _site_dmax(bpc.psi)
    out = np.zeros(len(vs) * dmax * dmax * 2, dtype=np.float64)
    check(bpc.psi._h.ptr, lib.tnqs_rdm_1site(bpc.psi._h.ptr, len(vs), ptr_i32(vs), ptr_f64(out)))
    return out.view(np.complex128).reshape((len(vs), dmax, dmax)).transpose(0, 2, 1).copy()   # [i, s, s'] (column-major blocks)


def _site_dmax(psi: TensorNetworkState) -> int:
    if not hasattr(psi, "_dmax"):
        lib = _lib.load()
        chi = psi.bond_dims()
        dm = 1
        for v in range(psi.g.nv):
            n = lib.tnqs_site_size(psi._h.ptr, v)
            dm = max(dm, n // int(np.prod([chi[e] for e in psi.g.incident[v]], dtype=np.int64)))
        psi._dmax = int(dm)
    return psi._dmax


def _collect_observable(g: Graph, obs: tuple):
    """src/expect.jl:160-176."""
    op, verts = obs[0], obs[1]
    coeff = obs[2] if len(obs) > 2 else 1
    verts = verts if isinstance(verts, list) else [verts]
    names = list(op) if isinstance(op, str) else list(op)
    if len(names) != len(verts):
        raise ValueError("Invalid observable: need as many operators as vertices passed.")
    return names, [_vid(g, x) for x in verts], coeff


def _expect_tree(bpc: BeliefPropagationCache, vs: Sequence[int], ops: Sequence[np.ndarray], tree=None) -> complex:
    """<O_1 .. O_n> on three or more vertices: numerator / denominator of the contraction over their Steiner tree with the
    converged messages entering the region (src/expect.jl:66-81).  A tree region is contracted by passing un-normalised messages
    from its leaves to a root, so this is a composition of existing entry points on two clones of the handle:
    numerator = clone with the operators applied (tnqs_apply_1site, no normalisation) and the original state attached as the bra
    (tnqs_set_bra), one un-normalised tnqs_bp_update over the tree's child -> parent edges, then the root's vertex scalar;
    denominator = the same without operators."""
    g = bpc.g
    tv, te = tree if tree is not None else _steiner_tree(g, vs)
    root = min(int(v) for v in tv)
    seq = _tree_schedule(g, tv, te, root)
    vals = []
    for with_ops in (True, False):
        work = bpc.copy()                                       # state + converged messages
        if with_ops:
            apply_1site_batch(work, list(vs), list(ops), normalize_tensors=False)
            _attach_bra(work, bpc.psi)
        update(work, maxiter=1, tolerance=None, edge_sequence=seq, normalize=False, inplace=True)
        vals.append(complex(vertex_scalars(work)[root]))
        if with_ops:
            _attach_bra(work, None)
    return vals[0] / vals[1]


def expect(x, obs, alg: str = "bp", cache_update_kwargs: Optional[dict] = None, tree=None):
    """<O> by belief propagation (src/expect.jl:53-83,115-136): one-site observables, two-site observables (contracted
    along the path joining the two vertices) and observables on three or more vertices (contracted over their Steiner tree;
    tree = (vertices, edge ids) overrides graphs.steiner_tree); a TensorNetworkState input builds and updates a fresh cache first."""
    if alg != "bp":
        raise ValueError("Only alg='bp' is on the CUDA path (Expected BeliefPropagationCache or TensorNetworkState for 'bp' algorithm)")
    if isinstance(x, TensorNetworkState):
        kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(x)
        x = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True, **kw)
    single = isinstance(obs, tuple)
    obs_list = [obs] if single else list(obs)
    g = x.g
    parsed = [_collect_observable(g, o) for o in obs_list]
    out: List[complex] = [0j] * len(parsed)
    one = [(i, p) for i, p in enumerate(parsed) if len(p[1]) == 1 and p[2] != 0]
    two = [(i, p) for i, p in enumerate(parsed) if len(p[1]) == 2 and p[2] != 0]
    for i, (names, vs, coeff) in enumerate(parsed):
        if len(vs) > 2 and coeff != 0:
            out[i] = coeff * _expect_tree(x, vs, [_gates.PAULI[n] for n in names], tree)
    if one:
        rho = rdm_1site(x, [p[1][0] for _, p in one])
        for (i, (names, vs, coeff)), r in zip(one, rho):
            o = _gates.PAULI[names[0]]
            d = o.shape[0]
            r = r[:d, :d]
            out[i] = coeff * np.sum(o.T * r) / np.trace(r)        # sum_{s,s'} O[s',s] rho[s,s']
    if two:
        # the Steiner tree of two vertices is a shortest path between them (src/expect.jl:66); BFS, lowest ids first
        offs, pv, ops = [0], [], []
        for _, (names, vs, coeff) in two:
            path = _shortest_path(g, vs[0], vs[1])
            pv += path
            offs.append(len(pv))
            ops += [_gates.PAULI[names[0]].ravel(order="F"), _gates.PAULI[names[1]].ravel(order="F")]
        offs_a, pv_a = i32(offs), i32(pv)
        ops_a = np.concatenate(ops).astype(np.complex128).view(np.float64)
        res = np.zeros(4 * len(two), dtype=np.float64)
        check(x.psi._h.ptr, _lib.load().tnqs_expect_path(x.psi._h.ptr, len(two), ptr_i32(offs_a), ptr_i32(pv_a), ptr_f64(ops_a), ptr_f64(res)))
        for j, (i, (names, vs, coeff)) in enumerate(two):
            out[i] = coeff * complex(res[4 * j], res[4 * j + 1]) / complex(res[4 * j + 2], res[4 * j + 3])
    return out[0] if single else out


def reduced_density_matrix(x, vertices, alg: str = "bp", normalize: bool = True, cache_update_kwargs: Optional[dict] = None, tree=None) -> np.ndarray:
    """BP reduced density matrix of one or more qubits (src/rdm.jl:49-70) as a 2^n x 2^n matrix rho[(s_1..s_n), (s_1'..s_n')],
    first vertex slow.  One vertex: tnqs_rdm_1site.  Several: rho = 2^-n sum_a <P_a> P_a over the 4^n Pauli strings, every <P_a>
    measured on the same region as the reference's single contraction (shortest path / Steiner tree), so only the normalised
    matrix is available."""
    if alg != "bp":
        raise ValueError("Only alg='bp' is on the CUDA path")
    if isinstance(x, TensorNetworkState):
        kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(x)
        x = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True, **kw)
    vs = [_vid(x.g, v) for v in (vertices if isinstance(vertices, list) else [vertices])]
    if len(vs) == 1:
        r = rdm_1site(x, vs)[0]
        d = int(x.psi.site_dims()[vs[0]])
        r = r[:d, :d]
        return r / np.trace(r) if normalize else r
    if not normalize:
        raise NotImplementedError("multi-site reduced density matrices are assembled from expectation values: normalised only")
    if any(int(x.psi.site_dims()[v]) != 2 for v in vs):
        raise NotImplementedError("multi-site reduced density matrices need qubit sites")
    n = len(vs)
    strings = ["".join(p) for p in itertools.product("IXYZ", repeat=n)]
    vals = expect(x, [(s, [x.g.names[v] for v in vs]) for s in strings[1:]], tree=tree)
    rho = np.eye(2 ** n, dtype=np.complex128)
    for s, val in zip(strings[1:], vals):
        p = np.array([[1.0 + 0j]])
        for ch in s:
            p = np.kron(p, _gates.PAULI[ch])
        rho += val * p
    return rho / 2 ** n


def _shortest_path(g: Graph, a: int, b: int) -> List[int]:
    if a == b:
        raise ValueError("Repeated vertex in collection")
    prev = {a: -1}
    queue = [a]
    while queue:
        u = queue.pop(0)
        if u == b:
            break
        for w in g.neighbors(u):
            if w not in prev:
                prev[w] = u
                queue.append(w)
    if b not in prev:
        raise ValueError("vertices are not connected")
    path = [b]
    while path[-1] != a:
        path.append(prev[path[-1]])
    return path[::-1]


def vertex_scalars(bpc: BeliefPropagationCache) -> np.ndarray:
    out = np.zeros(2 * bpc.g.nv, dtype=np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_vertex_scalars(bpc.psi._h.ptr, ptr_f64(out)))
    return out.view(np.complex128)


def edge_scalars(bpc: BeliefPropagationCache) -> np.ndarray:
    out = np.zeros(2 * max(bpc.g.ne, 1), dtype=np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_edge_scalars(bpc.psi._h.ptr, ptr_f64(out)))
    return out.view(np.complex128)[:bpc.g.ne]


def partitionfunction(bpc: BeliefPropagationCache) -> complex:
    """exp(sum log vertex scalars - sum log edge scalars) (src/MessagePassing/abstractbeliefpropagationcache.jl:252-267)."""
    num, den = vertex_scalars(bpc), edge_scalars(bpc)
    if np.any(den == 0):
        return 0.0
    return complex(np.exp(np.sum(np.log(num)) - np.sum(np.log(den))))


def _attach_bra(bpc: "BeliefPropagationCache", bra: Optional[TensorNetworkState]) -> None:
    check(bpc.psi._h.ptr, _lib.load().tnqs_set_bra(bpc.psi._h.ptr, bra._h.ptr if bra is not None else None))
    # the reference lives on the state that owns the device handle (the raw bra pointer sits in that handle), not on the cache
    # object: network(bilinear_cache(..)) stays valid after the cache is dropped.  The library also detaches on tnqs_destroy.
    bpc.psi._bra = bra


def _pad_bonds(psi: TensorNetworkState, chi: np.ndarray) -> TensorNetworkState:
    """psi embedded in zero-padded site tensors with bond dimensions chi[e] >= its own (the state is unchanged as a vector)."""
    own = psi.bond_dims()
    if np.array_equal(own, chi):
        return psi
    if np.any(own > chi):
        raise ValueError("cannot pad to smaller bond dimensions")
    ts = []
    for v, t in enumerate(psi.tensors()):
        shape = (t.shape[0],) + tuple(int(chi[e]) for e in psi.g.incident[v])
        big = np.zeros(shape, dtype=np.complex128, order="F")
        big[tuple(slice(0, n) for n in t.shape)] = t
        ts.append(big)
    return TensorNetworkState.from_arrays(psi.g, ts, chi_cap=int(max(1, chi.max(initial=1))), device=psi.device)


def bilinear_cache(psi: TensorNetworkState, phi: TensorNetworkState, delta_default: bool = False) -> "BeliefPropagationCache":
    """BeliefPropagationCache(BilinearForm(psi, phi)) (src/Forms/bilinearform.jl:28-37): BP messages of <phi|psi>.  phi must share
    psi's graph and site dimensions and stay unchanged while the cache lives.  The device contracts layers of equal bond
    dimensions: where they differ (an evolved operator against a product operator, examples/2dIsing_dynamics_Heisenbergpicture.jl:
    55-56) the thinner layer is embedded in zero-padded tensors, which changes no message entry that is ever multiplied by a
    non-zero number.  delta_default: start from the reference's default message of a bilinear form, delta over (ket bond,
    operator bond of dimension 1, bra bond) = the single entry [1, 1] (bilinearform.jl:39-42), instead of the identity."""
    if psi.g is not phi.g and (psi.g.nv != phi.g.nv or not np.array_equal(psi.g.esrc, phi.g.esrc) or not np.array_equal(psi.g.edst, phi.g.edst)):
        raise ValueError("inner: both states must live on the same graph")
    chi = np.maximum(psi.bond_dims(), phi.bond_dims())
    ket = _pad_bonds(psi, chi)
    bra = _pad_bonds(phi, chi)
    bpc = BeliefPropagationCache(ket.copy() if ket is psi else ket, _share=True)
    _attach_bra(bpc, bra)
    if delta_default:
        for e in range(psi.g.ne):
            m = np.zeros((int(chi[e]), int(chi[e])), dtype=np.complex128)
            m[0, 0] = 1.0
            setmessage(bpc, 2 * e, m)
            setmessage(bpc, 2 * e + 1, m)
    return bpc


def inner(psi: TensorNetworkState, phi: TensorNetworkState, alg: str = "bp", cache_update_kwargs: Optional[dict] = None,
          max_configuration_size: Optional[int] = None) -> complex:
    """inner(psi, phi; alg = "bp" | "loopcorrections") = <phi|psi> by BP on the bilinear form (src/inner.jl:67-86): exact on trees;
    with loop corrections the partition function is multiplied by 1 + the weights of the generalised loops of at most
    max_configuration_size edges."""
    if alg not in ("bp", "loopcorrections"):
        raise ValueError("Only alg='bp' and alg='loopcorrections' are on the CUDA path")
    kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(psi)
    bpc = update(bilinear_cache(psi, phi, delta_default=True), inplace=True, **kw)
    if alg == "loopcorrections":
        return loopcorrected_partitionfunction(bpc, _need_size(max_configuration_size))
    return partitionfunction(bpc)


def contract(tensors: Sequence[np.ndarray], g: Graph, alg: str = "bp", bp_update_kwargs: Optional[dict] = None, device: int = 0,
             info: Optional[dict] = None) -> complex:
    """contract(tn; alg = "bp") for a plain TensorNetwork without site indices (src/contract.jl:7-9): tensors[v] = A_v[b_e1 .. b_ez].
    The device sees it as the bilinear form of a ket with site dimension 1 against the all-ones product bra (bond dimension 1,
    zero-padded): its messages m[b, b'] carry the reference's vector message in column b' = 0 and zeros elsewhere, the
    normalisation, the diff and the partition function are then the reference's, and the start is its all-ones default
    (src/TensorNetworks/tensornetwork.jl:62-64)."""
    if alg != "bp":
        raise ValueError("Only alg='bp' is on the CUDA path")
    ket = TensorNetworkState.from_arrays(g, [np.asarray(t, dtype=np.complex128)[None, ...] for t in tensors], device=device)
    ones = TensorNetworkState.from_arrays(g, [np.ones((1,) * (1 + g.degree(v)), dtype=np.complex128) for v in range(g.nv)], device=device)
    chi = ket.bond_dims()
    bpc = BeliefPropagationCache(ket, _share=True)
    _attach_bra(bpc, _pad_bonds(ones, chi))
    for e in range(g.ne):
        m = np.zeros((int(chi[e]), int(chi[e])), dtype=np.complex128)
        m[:, 0] = 1.0
        setmessage(bpc, 2 * e, m)
        setmessage(bpc, 2 * e + 1, m)
    kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(ket)
    update(bpc, inplace=True, info=info, **kw)
    return partitionfunction(bpc)


def norm_sqr(x, alg: str = "bp", cache_update_kwargs: Optional[dict] = None, max_configuration_size: Optional[int] = None) -> complex:
    """norm_sqr(x; alg = "bp" | "loopcorrections") (src/norm_sqr.jl:56-76): partition function of the updated cache, with loop
    corrections up to max_configuration_size edges for the latter."""
    if alg not in ("bp", "loopcorrections"):
        raise ValueError("Only alg='bp' and alg='loopcorrections' are on the CUDA path")
    if isinstance(x, TensorNetworkState):
        kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(x)
        x = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True, **kw)
    if alg == "loopcorrections":
        return loopcorrected_partitionfunction(x, _need_size(max_configuration_size))
    return partitionfunction(x)


def norm(x, alg: str = "bp", **kwargs) -> complex:
    """LinearAlgebra.norm(x; alg, ..) = sqrt(norm_sqr) (src/norm_sqr.jl:90-91)."""
    return complex(np.sqrt(complex(norm_sqr(x, alg=alg, **kwargs))))


def _need_size(max_configuration_size: Optional[int]) -> int:
    if max_configuration_size is None:
        raise ValueError("alg='loopcorrections' needs max_configuration_size")
    return int(max_configuration_size)


def loop_weights(bpc: BeliefPropagationCache, egs: Sequence[Sequence[int]]) -> np.ndarray:
    """weights(bpc, egs) (src/MessagePassing/loopcorrection.jl:81-96): one complex number per edge-induced subgraph (tuples of
    edge ids) of the RESCALED cache.  On the device each weight is a signed sum of tree contractions of BP message updates with
    substituted messages (csrc/loop_corrections.cuh)."""
    egs = [tuple(int(e) for e in eg) for eg in egs]
    if not egs:
        return np.zeros(0, dtype=np.complex128)
    off = np.zeros(len(egs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(eg) for eg in egs])
    edges = np.array([e for eg in egs for e in eg], dtype=np.int32)
    out = np.zeros(2 * len(egs), dtype=np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_loop_weights(bpc.psi._h.ptr, len(egs), ptr_i32(off), ptr_i32(edges), ptr_f64(out)))
    return out.view(np.complex128)


def loopcorrected_partitionfunction(bpc: BeliefPropagationCache, max_configuration_size: int) -> complex:
    """loopcorrected_partitionfunction(bp_cache, max_configuration_size) (src/MessagePassing/loopcorrection.jl:3-16):
    Z_bp (1 + sum of the weights of every leafless edge-induced subgraph of at most that many edges), the weights taken on a
    rescaled copy of the cache."""
    zbp = partitionfunction(bpc)
    egs = _subgraphs_no_leaves(bpc.g, int(max_configuration_size))
    if not egs:
        return zbp
    ws = loop_weights(rescale(bpc), egs)
    return zbp * (1 + complex(np.sum(ws)))


def loop_correlation(bpc: BeliefPropagationCache, loop: Sequence[int], target_e: int) -> float:
    """loop_correlation(bpc, loop, target_e) (src/MessagePassing/beliefpropagationcache.jl:145-179): `loop` + [target_e] are the
    directed edge ids (k = 2e + dir) of a simple closed walk that target_e closes; 1 - |lambda_1| / sum |lambda_i| over the
    spectrum of the walk's transfer matrix (contracted on the device; the eigenvalues of the chi^2 x chi^2 matrix are taken on the
    host, as the reference does with LinearAlgebra.eigvals)."""
    g = bpc.g
    if g.is_tree():
        return 0.0
    ks = [int(k) for k in loop] + [int(target_e)]
    vs = [g.dsrc(k) for k in ks]
    if [g.ddst(k) for k in ks] != vs[1:] + vs[:1]:
        raise ValueError("loop + [target_e] must be a closed walk")
    walk = np.array(vs, dtype=np.int32)
    n2 = int(bpc.psi.bond_dims()[ks[-1] // 2]) ** 2
    out = np.zeros(2 * n2 * n2, dtype=np.float64)
    check(bpc.psi._h.ptr, _lib.load().tnqs_loop_transfer(bpc.psi._h.ptr, len(vs), ptr_i32(walk), ptr_f64(out)))
    lam = np.abs(np.linalg.eigvals(out.view(np.complex128).reshape(n2, n2, order="F")))
    return float(1.0 - lam.max() / lam.sum())


def loop_correlations(x, smallest_loop_size: int, bp_update_kwargs: Optional[dict] = None) -> List[float]:
    """loop_correlations (beliefpropagationcache.jl:182-197): one value per simple cycle of at most smallest_loop_size edges; a
    state is wrapped in an updated cache first."""
    if isinstance(x, TensorNetworkState):
        kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(x)
        x = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True, **kw)
    return [loop_correlation(x, loop, target) for loop, target in _cycle_walks(x.g, int(smallest_loop_size))]


def rescale(bpc: BeliefPropagationCache, inplace: bool = False) -> BeliefPropagationCache:
    """src/MessagePassing/abstractbeliefpropagationcache.jl:281-291."""
    out = bpc if inplace else bpc.copy()
    check(out.psi._h.ptr, _lib.load().tnqs_rescale(out.psi._h.ptr))
    return out


def symmetric_gauge(x, regularization: float = 10 * np.finfo(np.float64).eps, cache_update_kwargs: Optional[dict] = None,
                    inplace: bool = False):
    """src/symmetric_gauge.jl:1-68.  A TensorNetworkState is updated with maxiter=40 first and a state is returned."""
    if isinstance(x, TensorNetworkState):
        kw = cache_update_kwargs if cache_update_kwargs is not None else dict(maxiter=40)
        bpc = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True, **kw)
        return symmetric_gauge(bpc, regularization, inplace=True).psi
    out = x if inplace else x.copy()
    check(out.psi._h.ptr, _lib.load().tnqs_symmetric_gauge(out.psi._h.ptr, float(regularization)))
    return out


def gauge_and_scale(psi: TensorNetworkState, cache_update_kwargs: Optional[dict] = None) -> TensorNetworkState:
    """symmetrize_and_bpnormalize (src/symmetric_gauge.jl:76-83)."""
    kw = cache_update_kwargs if cache_update_kwargs is not None else dict(maxiter=40)
    bpc = update(BeliefPropagationCache(psi.copy(), _share=True), inplace=True, **kw)
    rescale(bpc, inplace=True)
    return symmetric_gauge(bpc, inplace=True).psi


def truncate(x, maxdim: int, cutoff: Optional[float] = None, normalize_tensors: bool = True, alg: str = "bp",
             edge_groups: Optional[List[List[int]]] = None, bp_update_kwargs: Optional[dict] = None, inplace: bool = False):
    """src/truncate.jl:12-38 (cache) and :74-79 (state: bare update first)."""
    if alg != "bp":
        raise ValueError("Only alg='bp' is on the CUDA path")
    if isinstance(x, TensorNetworkState):
        bpc = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True)
        return truncate(bpc, maxdim, cutoff, normalize_tensors, alg, edge_groups, bp_update_kwargs, True).psi
    out = x if inplace else x.copy()
    g = out.g
    groups = edge_groups if edge_groups is not None else _edge_color(g)
    goff = np.zeros(len(groups) + 1, dtype=np.int32)
    goff[1:] = np.cumsum([len(grp) for grp in groups])
    gedges = i32([e for grp in groups for e in grp])
    kw = dict(bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(out))
    kw.pop("verbose", None)
    seq = i32(default_edge_sequence(g))
    bopts = _bp_opts(g, **kw)
    aopts = _apply_opts(dict(maxdim=maxdim, cutoff=cutoff, normalize_tensors=normalize_tensors))
    check(out.psi._h.ptr, _lib.load().tnqs_truncate(out.psi._h.ptr, len(groups), ptr_i32(goff), ptr_i32(gedges), C.byref(aopts),
                                                     ptr_i32(seq), len(seq), C.byref(bopts)))
    return out


def entanglement(x, edge, alg: str = "bp") -> float:
    """Von Neumann entropy across a bond from its two messages (src/symmetric_gauge.jl:85-114).  The two
    chi x chi messages are read back and the scalar is formed on the host, like the reference's own scalar
    post-processing."""
    if isinstance(x, TensorNetworkState):
        x = update(BeliefPropagationCache(x.copy(), _share=True), inplace=True)
    g = x.g
    e = int(edge) if isinstance(edge, (int, np.integer)) else g.edge_of[(_vid(g, edge[0]), _vid(g, edge[1]))]
    roots = []
    for k in (2 * e, 2 * e + 1):
        w, u = np.linalg.eigh(message(x, k))
        if np.any(w < -10 * np.finfo(np.float64).eps):
            raise TnqsError(-4, "negative message eigenvalue")
        w = np.where(np.abs(w) < 10 * np.finfo(np.float64).eps, 0.0, w)
        roots.append((u * np.sqrt(np.maximum(w, 0.0))) @ u.conj().T)
    s = np.linalg.svd(roots[0] @ roots[1].T, compute_uv=False)
    s = s / np.linalg.norm(s)
    ee = 0.0
    for d in s:
        if abs(d) >= np.finfo(np.float64).eps:
            ee -= d * d * math.log(d * d)
    return abs(ee)


def normalize(psi: TensorNetworkState, alg: str = "bp", cache_update_kwargs: Optional[dict] = None) -> TensorNetworkState:
    """BP normalisation (src/normalize.jl:1-6): update, then rescale so every vertex and edge scalar is 1."""
    if alg != "bp":
        raise ValueError("Only alg='bp' is on the CUDA path")
    kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(psi)
    bpc = update(BeliefPropagationCache(psi.copy(), _share=True), inplace=True, **kw)
    return rescale(bpc, inplace=True).psi


def sample(psi: TensorNetworkState, nsamples: int, alg: str = "bp", bp_update_kwargs: Optional[dict] = None,
           gauge_state: bool = True, rng=None):
    """Site-by-site BP sampling (src/sampling.jl:3-46): per vertex the BP reduced density matrix,
    a draw from its diagonal, projection of the site onto the outcome, and a BP update of the projected cache.
    Every tensor operation is an existing C-ABI call; only the draw (inverse CDF on `rng.random()`) happens on the host.
    Returns ([{vertex name: outcome}], psi) like the reference."""
    if alg != "bp":
        raise ValueError("Only alg='bp' is on the CUDA path")
    rng = rng if rng is not None else np.random.default_rng()
    kw = dict(bp_update_kwargs or {})
    g = psi.g
    bpc = update(BeliefPropagationCache(psi.copy(), _share=True), inplace=True, **kw)
    if gauge_state:                                        # symmetrize_and_normalize (src/symmetric_gauge.jl:70-74)
        rescale(bpc, inplace=True)
        symmetric_gauge(bpc, inplace=True)
    out = []
    for _ in range(nsamples):
        proj = bpc.copy()
        bits = {}
        for v in range(g.nv):
            rho = rdm_1site(proj, [v])[0]
            p = np.real(np.diag(rho))
            p = np.clip(p / p.sum(), 0.0, None)
            config = int(min(np.searchsorted(np.cumsum(p), rng.random() * p.sum(), side="right"), len(p) - 1))
            bits[g.names[v]] = config
            d = len(p)
            proj_op = np.zeros((d, d), dtype=np.complex128)
            proj_op[config, config] = 1.0
            apply_1site_batch(proj, [v], [proj_op], normalize_tensors=False)
            if v != g.nv - 1:
                update(proj, inplace=True, **kw)
        out.append(bits)
    return out, psi
