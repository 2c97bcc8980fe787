"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

NumPy complex128 (+ LAPACK through numpy.linalg: zheevd, zgeqrf/zungqr, zgesdd) restatement of the
two hot paths of haimeng-zhang/TensorNetworkQuantumSimulator: the belief-propagation message sweep
and the two-site gate application with SVD truncation, plus the BP expectation values, scalars,
rescaling, symmetric gauge and BP truncation that sit on the same path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker.  The product (`tensornetworkquantumsimulator_b200`,
`libtnqs_cuda.so`) never imports it and has no CPU fallback.

"Parity unpinned": the reference is pure Julia on ITensors; Julia is not installed here or on the GPU
box, the numerics live in un-vendored packages (ITensors 0.9, NamedGraphs 0.7, TensorOperations 5.2,
SimpleGraphAlgorithms 0.6; `Project.toml:30-51`, no Manifest), and the reference's tests hold no golden
vectors.  What pins this restatement instead (tests/test_oracle_*.py):
an independent exact statevector simulator, BP exactness on trees, the reference's own property tests
(test/test_apply.jl:11-53, test/test_beliefpropagation.jl:26-53, test/test_constructors.jl:69-73,
test/test_truncate.jl:16-33) and the physical invariants of the symmetric gauge.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Conventions: site tensor of v = ndarray [s, b_e1 .. b_ez], incident edges ascending; message of the
directed edge k=2e+dir = ndarray m[b, b'] (b contracts the ket leg, b' the bra leg).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from tensornetworkquantumsimulator_b200 import gates as _gates
from tensornetworkquantumsimulator_b200 import rng as _rng
from tensornetworkquantumsimulator_b200.graphs import (Graph, forest_cover_edge_sequence, edge_color, steiner_tree,
                                                       edgeinduced_subgraphs_no_leaves, cycle_walks)

EPS = float(np.finfo(np.float64).eps)
DEFAULT_BP_MAXITER = 25          # src/MessagePassing/beliefpropagationcache.jl:108
DEFAULT_TOL_F64 = 1.0e-8         # src/MessagePassing/beliefpropagationcache.jl:109-112


# ---------------------------------------------------------------------------------------------
# containers  (src/TensorNetworks/tensornetworkstate.jl:4-18, src/MessagePassing/beliefpropagationcache.jl:9-13)
# ---------------------------------------------------------------------------------------------
@dataclass
class TensorNetworkState:
    g: Graph
    tensors: List[np.ndarray]

    def copy(self) -> "TensorNetworkState":
        return TensorNetworkState(self.g, [t.copy() for t in self.tensors])

    def bond_dim(self, e: int) -> int:
        v = int(self.g.esrc[e])
        return self.tensors[v].shape[1 + self.g.leg(v, e)]

    def bond_dims(self) -> np.ndarray:
        return np.array([self.bond_dim(e) for e in range(self.g.ne)], dtype=np.int32)

    def maxvirtualdim(self) -> int:
        return int(self.bond_dims().max()) if self.g.ne else 1


@dataclass
class BeliefPropagationCache:
    network: TensorNetworkState
    messages: Dict[int, np.ndarray] = field(default_factory=dict)
    bra: Optional[TensorNetworkState] = None        # BilinearForm(ket = network, bra): src/Forms/bilinearform.jl:28-54
    delta_default: bool = False                     # bilinear forms: default message = delta over (ket, operator, bra) bonds

    def copy(self) -> "BeliefPropagationCache":
        return BeliefPropagationCache(self.network.copy(), {k: m.copy() for k, m in self.messages.items()}, self.bra,
                                      self.delta_default)

    def bra_tensor(self, v: int) -> np.ndarray:
        """The tensor whose conjugate closes a contraction at v: the state itself, or the bra of a bilinear form."""
        return (self.bra if self.bra is not None else self.network).tensors[v]

    @property
    def g(self) -> Graph:
        return self.network.g


def tensornetworkstate(f, g: Graph) -> TensorNetworkState:
    """Product state, all bonds of dimension 1 (src/TensorNetworks/tensornetworkstate.jl:124-144)."""
    ts = []
    for v in range(g.nv):
        vec = _gates.state_vector(f(g.names[v]))
        ts.append(vec.reshape((vec.shape[0],) + (1,) * g.degree(v)).copy())
    return TensorNetworkState(g, ts)


def paulitensornetworkstate(f, g: Graph) -> TensorNetworkState:
    """Heisenberg-picture operator over Pauli sites, basis (I, X, Y, Z) (src/TensorNetworks/tensornetworkstate_constructors.jl:1,19-24)."""
    return tensornetworkstate(lambda v: _gates.PAULI_STATES[f(v)], g)


def identitytensornetworkstate(g: Graph) -> TensorNetworkState:
    """src/TensorNetworks/tensornetworkstate_constructors.jl:31-35."""
    return paulitensornetworkstate(lambda v: "I", g)


def random_tensornetworkstate(g: Graph, bond_dimension: int = 1, seed: int = 0, d: int = 2) -> TensorNetworkState:
    """Random N(0,1)+iN(0,1) entries (src/TensorNetworks/tensornetworkstate.jl:82-92), drawn from the
    counter-based generator: stream = vertex id, element index = column-major offset."""
    ts = []
    for v in range(g.nv):
        shape = (d,) + (bond_dimension,) * g.degree(v)
        flat = _rng.complex_normal(seed, v, int(np.prod(shape)))
        ts.append(flat.reshape(shape, order="F").copy())
    return TensorNetworkState(g, ts)


# ---------------------------------------------------------------------------------------------
# messages
# ---------------------------------------------------------------------------------------------
def message(bpc: BeliefPropagationCache, k: int) -> np.ndarray:
    """Stored message or the un-normalised identity default
    (src/MessagePassing/abstractbeliefpropagationcache.jl:88-91, src/TensorNetworks/tensornetworkstate.jl:64-67)."""
    m = bpc.messages.get(k)
    if m is None:
        chi = bpc.network.bond_dim(k // 2)
        if bpc.bra is None:
            return np.eye(chi, dtype=np.complex128)
        # bilinear form: m[b_ket, b_bra], the layers may have different bond dimensions.  The reference's default is
        # delta(ket bond, operator bond, bra bond) with an operator bond of dimension 1 = the single entry [1, 1]
        # (src/Forms/bilinearform.jl:39-42); the identity default is what BeliefPropagationCache(ket, {}, bra) used before
        # `inner` learnt the former (golden fixtures were made with it).
        chib = bpc.bra.bond_dim(k // 2)
        if bpc.delta_default:
            m = np.zeros((chi, chib), dtype=np.complex128)
            m[0, 0] = 1.0
            return m
        return np.eye(chi, chib, dtype=np.complex128)
    return m


def _absorb(t: np.ndarray, axis: int, m: np.ndarray) -> np.ndarray:
    """out[.., c, ..] = sum_b t[.., b, ..] m[b, c] on `axis`."""
    return np.moveaxis(np.tensordot(t, m, axes=([axis], [0])), -1, axis)


def incoming(bpc: BeliefPropagationCache, v: int, ignore: Sequence[int] = ()) -> List[Tuple[int, np.ndarray]]:
    """(leg axis, message) for every message into v except those from vertices in `ignore`
    (src/MessagePassing/abstractbeliefpropagationcache.jl:132-142)."""
    g = bpc.g
    out = []
    for i, e in enumerate(g.incident[v]):
        u = g.other(e, v)
        if u in ignore:
            continue
        out.append((1 + i, message(bpc, g.dedge(u, v))))
    return out


def updated_message(bpc: BeliefPropagationCache, k: int, normalize: bool = True,
                    enforce_hermiticity: bool = False) -> np.ndarray:
    """src/MessagePassing/abstractbeliefpropagationcache.jl:144-168: absorb the incoming messages of
    src(k) (all but the one from dst(k)) into the ket, close with the bra, optionally hermitise,
    divide by the sum of ALL entries."""
    g = bpc.g
    v, w = g.dsrc(k), g.ddst(k)
    t = bpc.network.tensors[v]
    a = t
    for axis, m in incoming(bpc, v, ignore=(w,)):
        a = _absorb(a, axis, m)
    ax_w = 1 + g.leg(v, k // 2)
    others = [i for i in range(t.ndim) if i != ax_w]
    out = np.tensordot(a, bpc.bra_tensor(v).conj(), axes=(others, others))          # [b_w, b_w']
    if enforce_hermiticity:                                          # beliefpropagationcache.jl:123-127
        out = (out + out.conj().T) / 2
    if normalize:
        s = out.sum()
        if s != 0:
            out = out / s
    return out


def message_diff(a: np.ndarray, b: np.ndarray) -> float:
    """1 - |<a,b>|^2 / (|a|^2 |b|^2)  (src/MessagePassing/beliefpropagationcache.jl:15-19)."""
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    f = abs(np.vdot(a, b) / (na * nb)) ** 2
    return 1.0 - f


def default_bp_update_kwargs(tns: TensorNetworkState) -> dict:
    """src/MessagePassing/beliefpropagationcache.jl:114-119."""
    return dict(maxiter=1 if tns.g.is_tree() else DEFAULT_BP_MAXITER, tolerance=DEFAULT_TOL_F64)


def update(bpc: BeliefPropagationCache, maxiter: Optional[int] = None, tolerance: Optional[float] = None,
           edge_sequence: Optional[Sequence[int]] = None, normalize: bool = True,
           enforce_hermiticity: bool = False, info: Optional[dict] = None) -> BeliefPropagationCache:
    """Sequential Gauss-Seidel BP (src/MessagePassing/abstractbeliefpropagationcache.jl:182-218, defaults
    src/MessagePassing/beliefpropagationcache.jl:56-65): bare `update(bpc)` has tolerance=None and runs all
    `maxiter` sweeps (25, or 1 on trees)."""
    g = bpc.g
    if maxiter is None:
        maxiter = 1 if g.is_tree() else DEFAULT_BP_MAXITER
    if edge_sequence is None:
        edge_sequence = forest_cover_edge_sequence(g)
    bpc = bpc.copy()
    iters, mean = 0, float("nan")
    for it in range(1, maxiter + 1):
        iters = it
        diff = 0.0
        for k in edge_sequence:
            k = int(k)
            prev = message(bpc, k) if tolerance is not None else None
            new = updated_message(bpc, k, normalize, enforce_hermiticity)
            bpc.messages[k] = new
            if tolerance is not None:
                diff += message_diff(new, prev)
        if tolerance is not None:
            mean = diff / len(edge_sequence)
            if mean <= tolerance:
                break
    if info is not None:
        info["iters"] = iters
        info["mean_diff"] = mean
    return bpc


def jacobi_sweep(bpc: BeliefPropagationCache, normalize: bool = True, shared: bool = False) -> BeliefPropagationCache:
    """One JACOBI sweep (every message from the previous sweep's values; the engine's TNQS_BP_JACOBI mode, used for the multi-GPU
    runs - not a reference schedule, same fixed point).  shared = the vertex-centric evaluation planned for round 2 (DESIGN.md
    section 8): the z outgoing messages of a vertex share partial absorptions through prefix / suffix products, e.g. z = 4:
    A12 = T m1 m2, A34 = T m3 m4, out4 = A12 m3, out3 = A12 m4, out1 = A34 m2, out2 = A34 m1 - the same numbers up to the
    order of the floating-point products."""
    g = bpc.g
    out = bpc.copy()
    if not shared:
        out.messages = {k: updated_message(bpc, k, normalize) for k in range(2 * g.ne)}
        return out
    for v in range(g.nv):
        t = bpc.network.tensors[v]
        inc = incoming(bpc, v)                                    # (axis, message) per leg, ascending
        z = len(inc)
        # prefix[i] = T with legs 0..i-1 absorbed, suffix[i] = T with legs i..z-1 absorbed (each one absorption further)
        prefix = [t]
        for axis, m in inc[:-1]:
            prefix.append(_absorb(prefix[-1], axis, m))
        for i in range(z):                                        # outgoing message on leg i: prefix[i] with legs i+1.. absorbed
            a = prefix[i]
            for axis, m in inc[i + 1:]:
                a = _absorb(a, axis, m)
            ax = inc[i][0]
            others = [j for j in range(t.ndim) if j != ax]
            mout = np.tensordot(a, bpc.bra_tensor(v).conj(), axes=(others, others))
            if normalize and mout.sum() != 0:
                mout = mout / mout.sum()
            e = g.incident[v][i]
            out.messages[g.dedge(v, g.other(e, v))] = mout
    return out


def jacobi_absorptions_per_vertex(z: int, shared: bool) -> int:
    """Leg absorptions (each 8 d chi^(z+1) flops at uniform chi) one vertex costs per JACOBI sweep: z (z - 1) message by message;
    with shared prefixes (z - 1) + sum_i (z - 1 - i) = (z - 1)(z + 2) / 2 as written above, and 3z - 6 with prefixes AND suffixes
    (z >= 3: 6 instead of 12 at z = 4 would need both halves stored; the two-halves scheme of DESIGN.md section 8 does 8)."""
    return z * (z - 1) if not shared else (z - 1) * (z + 2) // 2


# ---------------------------------------------------------------------------------------------
# scalars, partition function, rescale
# ---------------------------------------------------------------------------------------------
def vertex_scalar(bpc: BeliefPropagationCache, v: int) -> complex:
    """src/MessagePassing/abstractbeliefpropagationcache.jl:21-27."""
    t = bpc.network.tensors[v]
    a = t
    for axis, m in incoming(bpc, v):
        a = _absorb(a, axis, m)
    return complex(np.vdot(bpc.bra_tensor(v), a))            # sum conj(t) * a


def edge_scalar(bpc: BeliefPropagationCache, e: int) -> complex:
    """src/MessagePassing/beliefpropagationcache.jl:38-40: sum_{b,b'} m_e[b,b'] m_rev[b,b']."""
    return complex(np.sum(message(bpc, 2 * e) * message(bpc, 2 * e + 1)))


def freenergy(bpc: BeliefPropagationCache) -> complex:
    """src/MessagePassing/abstractbeliefpropagationcache.jl:252-263."""
    num = np.array([vertex_scalar(bpc, v) for v in range(bpc.g.nv)])
    den = np.array([edge_scalar(bpc, e) for e in range(bpc.g.ne)])
    if np.any(den == 0):
        return -math.inf
    return np.sum(np.log(num.astype(np.complex128))) - np.sum(np.log(den.astype(np.complex128)))


def partitionfunction(bpc: BeliefPropagationCache) -> complex:
    return complex(np.exp(freenergy(bpc)))


def inner(psi: TensorNetworkState, phi: TensorNetworkState, maxiter: Optional[int] = None, tolerance: Optional[float] = None,
          edge_sequence: Optional[Sequence[int]] = None) -> complex:
    """inner(psi, phi; alg = "bp") (src/inner.jl:82-86): BP on BilinearForm(psi, phi) - the bra layer is dag(phi), the operator
    layer the identity (src/Forms/bilinearform.jl:28-37) - then the partition function.  <phi|psi> exactly on trees."""
    kw = default_bp_update_kwargs(psi)
    bpc = BeliefPropagationCache(psi.copy(), {}, phi, delta_default=True)
    bpc = update(bpc, maxiter=maxiter if maxiter is not None else kw["maxiter"],
                 tolerance=tolerance if tolerance is not None else kw["tolerance"], edge_sequence=edge_sequence)
    return partitionfunction(bpc)


def _csign(x: complex) -> complex:
    return 0.0 if x == 0 else x / abs(x)


def rescale_messages(bpc: BeliefPropagationCache, edges: Optional[Sequence[int]] = None) -> None:
    """In place; src/MessagePassing/beliefpropagationcache.jl:129-142."""
    for e in (range(bpc.g.ne) if edges is None else edges):
        me, mer = message(bpc, 2 * e), message(bpc, 2 * e + 1)
        me, mer = me / np.linalg.norm(me), mer / np.linalg.norm(mer)
        n = complex(np.sum(me * mer))
        if n.imag == 0:
            s = _csign(n)
            me = me * s
            n = n * s
        r = 1.0 / np.sqrt(n)
        bpc.messages[2 * e] = me * r
        bpc.messages[2 * e + 1] = mer * r


def rescale_vertices(bpc: BeliefPropagationCache, vertices: Optional[Sequence[int]] = None) -> None:
    """In place; src/MessagePassing/beliefpropagationcache.jl:87-106 (TensorNetworkState branch).  A bilinear form has no rule
    there (the reference raises, :102-103): its ket layer is divided by the vertex scalar, which makes the scalar 1 - what the
    TensorNetwork branch (:100-101) does up to its sign factor, and what the loop series needs."""
    for v in (range(bpc.g.nv) if vertices is None else vertices):
        vn = vertex_scalar(bpc, v)
        if bpc.bra is not None:
            bpc.network.tensors[v] = bpc.network.tensors[v] * (1.0 / complex(vn))
            continue
        s = _csign(vn) if vn.imag == 0 else 1.0
        bpc.network.tensors[v] = bpc.network.tensors[v] * (s * (1.0 / np.sqrt(complex(vn))))


def rescale(bpc: BeliefPropagationCache) -> BeliefPropagationCache:
    """src/MessagePassing/abstractbeliefpropagationcache.jl:281-291."""
    bpc = bpc.copy()
    rescale_messages(bpc)
    rescale_vertices(bpc)
    return bpc


# ---------------------------------------------------------------------------------------------
# matrix functions and truncation
# ---------------------------------------------------------------------------------------------
class NegativeEigenvalueError(ValueError):
    pass


def pseudo_sqrt_inv_sqrt(m: np.ndarray, cutoff: float = 10 * EPS) -> Tuple[np.ndarray, np.ndarray]:
    """src/utils.jl:18-26: Hermitian eigh, f(λ)=0 if λ==0 or |λ|<cutoff else sqrt(λ) (1/sqrt(λ));
    a negative λ beyond the cutoff throws in the reference (sqrt of a negative real)."""
    w, u = np.linalg.eigh(m)
    keep = ~((w == 0) | (np.abs(w) < cutoff))
    if np.any(w[keep] < 0):
        raise NegativeEigenvalueError(f"message eigenvalue {w[keep].min()} < 0")
    rs = np.zeros_like(w)
    ri = np.zeros_like(w)
    rs[keep] = np.sqrt(w[keep])
    ri[keep] = 1.0 / np.sqrt(w[keep])
    return (u * rs) @ u.conj().T, (u * ri) @ u.conj().T


def truncate_spectrum(p: np.ndarray, maxdim: Optional[int] = None, cutoff: Optional[float] = None,
                      mindim: int = 1) -> Tuple[int, float]:
    """NDTensors `truncate!` as reached through `factorize_svd(...; maxdim, cutoff)` at
    src/Apply/simple_update.jl:46-52.  p = S^2 sorted descending.  Returns
    (kept rank, truncerr = discarded weight / total weight); sequential FP64 tail sums."""
    p = np.asarray(p, dtype=np.float64)
    n = p.shape[0]
    if n == 1:
        return 1, 0.0
    maxdim = n if maxdim is None else min(int(maxdim), n)
    t = 0.0
    while n > maxdim:
        t += float(p[n - 1])
        n -= 1
    scale = 0.0
    for x in p:                     # fixed ascending order; the device kernel sums in the same order
        scale += float(x)
    if scale == 0.0:
        scale = 1.0
    if cutoff is not None:
        while n > mindim and t + float(p[n - 1]) <= cutoff * scale:
            t += float(p[n - 1])
            n -= 1
    return n, t / scale


# ---------------------------------------------------------------------------------------------
# gate application
# ---------------------------------------------------------------------------------------------
def simple_update_2site(bpc: BeliefPropagationCache, gate: np.ndarray, v1: int, v2: int,
                        maxdim: Optional[int] = None, cutoff: Optional[float] = None,
                        normalize_tensors: bool = True, detail: Optional[dict] = None):
    """src/Apply/simple_update.jl:24-66 with envs from src/Apply/apply_gates.jl:101.
    gate[s1', s2', s1, s2].  Returns (T1', T2', S, truncerr) in canonical leg order."""
    g = bpc.g
    e = g.edge_of[(v1, v2)]
    t1, t2 = bpc.network.tensors[v1], bpc.network.tensors[v2]

    def reduce_site(v, other, t):
        envs = [(ax, *pseudo_sqrt_inv_sqrt(m)) for ax, m in incoming(bpc, v, ignore=(other,))]
        a = t
        for ax, rs, _ in envs:                                   # simple_update.jl:35-36
            a = _absorb(a, ax, rs)
        ax_b = 1 + g.leg(v, e)
        outer = [i for i in range(t.ndim) if i not in (0, ax_b)]
        am = np.transpose(a, outer + [0, ax_b])
        oshape = am.shape[:len(outer)]
        d, chi = am.shape[-2], am.shape[-1]
        q, r = np.linalg.qr(am.reshape(int(np.prod(oshape, dtype=np.int64)), d * chi))   # :39-40
        return q.reshape(oshape + (q.shape[1],)), r.reshape(r.shape[0], d, chi), envs, outer, ax_b

    q1, r1, envs1, outer1, axb1 = reduce_site(v1, v2, t1)
    q2, r2, envs2, outer2, axb2 = reduce_site(v2, v1, t2)
    theta = np.einsum("asb,ctb->asct", r1, r2)                   # R1 * R2        (:43)
    theta = np.einsum("xyst,asct->axcy", gate, theta)            # apply(o, .)    (:43)
    n1, d1, n2, d2 = theta.shape
    u, s, vh = np.linalg.svd(theta.reshape(n1 * d1, n2 * d2), full_matrices=False)    # :46-52
    rank, err = truncate_spectrum(s * s, maxdim, cutoff)
    u, s, vh = u[:, :rank], s[:rank], vh[:rank]
    rs = np.sqrt(s)
    r1n = (u * rs).reshape(n1, d1, rank)                          # ortho="none": U sqrt(S)
    r2n = (rs[:, None] * vh).reshape(rank, n2, d2).transpose(1, 2, 0)

    def rebuild(q, rn, envs, outer, ax_b, ndim):
        # Q * dag(M^-1/2): Q'[.., b, ..] = sum_c Q[.., c, ..] conj(Minv[b, c])      (:55-56)
        for pos, (ax, _, ri) in enumerate(envs):
            q = _absorb(q, outer.index(ax), ri.conj().T)
        tn = np.tensordot(q, rn, axes=([q.ndim - 1], [0]))        # [outer.., s, u]  (:57)
        perm_src = outer + [0, ax_b]
        inv = [perm_src.index(i) for i in range(ndim)]
        return np.ascontiguousarray(np.transpose(tn, inv))

    t1n = rebuild(q1, r1n, envs1, outer1, axb1, t1.ndim)
    t2n = rebuild(q2, r2n, envs2, outer2, axb2, t2.ndim)
    if normalize_tensors:                                          # :58-65
        s = s / np.linalg.norm(s)
        t1n = t1n / np.linalg.norm(t1n)
        t2n = t2n / np.linalg.norm(t2n)
    if detail is not None:
        detail.update(rank=rank, svals=s.copy(), truncerr=err)
    return t1n, t2n, s, err


def apply_gate(bpc: BeliefPropagationCache, gate: np.ndarray, vs: Sequence[int], maxdim=None, cutoff=None,
               normalize_tensors: bool = True, detail: Optional[dict] = None) -> float:
    """In place on `bpc`; src/Apply/apply_gates.jl:95-122 (+ 1-site branch src/Apply/simple_update.jl:21-23,63-65)."""
    if len(vs) == 1:
        v = vs[0]
        t = np.tensordot(gate, bpc.network.tensors[v], axes=([1], [0]))
        if normalize_tensors:
            t = t / np.linalg.norm(t)
        bpc.network.tensors[v] = t
        return 0.0
    v1, v2 = vs
    t1, t2, s, err = simple_update_2site(bpc, gate, v1, v2, maxdim, cutoff, normalize_tensors, detail)
    k = bpc.g.dedge(v1, v2)
    # message <- S * sign(S) on the diagonal (apply_gates.jl:105-115)
    bpc.messages[k] = np.diag(s).astype(np.complex128)
    bpc.messages[k ^ 1] = np.diag(s).astype(np.complex128)
    bpc.network.tensors[v1] = t1
    bpc.network.tensors[v2] = t2
    return float(err)


def _vid(g: Graph, x) -> int:
    """Vertex NAME -> dense id; the reference resolves by name only (e.g. the integer labels of build_graph_from_gates,
    graph_ops.jl:53-67).  On a graph whose names are integers an integer is a name; elsewhere it may also be a dense id."""
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
        if any(isinstance(n, (int, np.integer)) for n in g.names):
            return g.index[int(x)]
        return int(x)
    return g.v(x)


def _resolve_circuit(g: Graph, circuit: Sequence[tuple], tensors=None):
    out = []
    for gate in circuit:
        if isinstance(gate[0], np.ndarray):
            arr, verts = gate[0], gate[1]
        else:
            vl = gate[1] if isinstance(gate[1], list) else [gate[1]]
            dims = None if tensors is None else [tensors[_vid(g, x)].shape[0] for x in vl]
            arr, verts = _gates.gate_array(gate, dims), gate[1]      # Pauli sites (d = 4): Heisenberg branch, gate_definitions.jl:41-42
        verts = verts if isinstance(verts, list) else [verts]
        out.append((np.asarray(arr, dtype=np.complex128), [_vid(g, x) for x in verts]))
    return out


def apply_gates(circuit: Sequence[tuple], psi, apply_kwargs: Optional[dict] = None,
                bp_update_kwargs: Optional[dict] = None, update_cache: bool = True,
                info: Optional[dict] = None):
    """src/Apply/apply_gates.jl:13-23 (TensorNetworkState: build cache, initial update, return the network)
    and :42-92 (ordered loop; a BP refresh before a 2-site gate that touches a vertex touched since the last
    refresh -- 1-site gates mark vertices too, :84; final refresh)."""
    apply_kwargs = dict(apply_kwargs or {})
    if isinstance(psi, TensorNetworkState):
        kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(psi)
        bpc = update(BeliefPropagationCache(psi), **kw)
        bpc, errs = apply_gates(circuit, bpc, apply_kwargs, kw, update_cache, info)
        return bpc.network, errs
    bpc = psi.copy()
    kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(bpc.network)
    gates = _resolve_circuit(bpc.g, circuit, bpc.network.tensors)
    affected: set = set()
    errs = np.zeros(len(gates))
    refreshes = 0
    for i, (arr, vs) in enumerate(gates):
        if update_cache and len(vs) >= 2 and any(v in affected for v in vs):
            bpc = update(bpc, **kw)
            refreshes += 1
            affected = set()
        errs[i] = apply_gate(bpc, arr, vs, **apply_kwargs)
        affected |= set(vs)
    if update_cache:
        bpc = update(bpc, **kw)
        refreshes += 1
    if info is not None:
        info["refreshes"] = refreshes
    return bpc, errs


# ---------------------------------------------------------------------------------------------
# BP expectation values (src/expect.jl:58-83)
# ---------------------------------------------------------------------------------------------
def _tree_path(g: Graph, a: int, b: int) -> List[int]:
    """Shortest path a..b (the Steiner tree of two vertices)."""
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
    path = [b]
    while path[-1] != a:
        path.append(prev[path[-1]])
    return path[::-1]


def _region_contract(bpc: BeliefPropagationCache, region: List[int], ops: Dict[int, np.ndarray]) -> complex:
    """Contract [T, conj(T) (+ op)] over a path region with all incoming messages."""
    g = bpc.g
    env = None            # matrix on the bond towards the next vertex of the path: env[b, b']
    for i, v in enumerate(region):
        t = bpc.network.tensors[v]
        ket = t
        if v in ops:
            ket = np.tensordot(ops[v], t, axes=([1], [0]))       # O[s', s] T[s] -> index s' (paired with bra)
        prev_v = region[i - 1] if i > 0 else None
        next_v = region[i + 1] if i + 1 < len(region) else None
        for axis, m in incoming(bpc, v, ignore=tuple(x for x in (prev_v, next_v) if x is not None)):
            ket = _absorb(ket, axis, m)
        if prev_v is not None:
            ket = _absorb(ket, 1 + g.leg(v, g.edge_of[(prev_v, v)]), env)
        if next_v is None:
            return complex(np.vdot(t, ket))
        ax = 1 + g.leg(v, g.edge_of[(v, next_v)])
        others = [j for j in range(t.ndim) if j != ax]
        env = np.tensordot(ket, t.conj(), axes=(others, others))
    raise AssertionError


def _tree_region_contract(bpc: BeliefPropagationCache, vertices: Sequence[int], edges: Sequence[int], ops: Dict[int, np.ndarray]) -> complex:
    """Contract [T, conj(T) (+ op)] over the vertices of a tree region with every message that enters the region from outside
    (src/expect.jl:66-81 with steiner_vs = the tree's vertices).  Recursive: the environment a subtree passes to its parent is a
    matrix on their shared bond.  The region must not contain an edge of the graph outside the tree."""
    g = bpc.g
    vs, es = set(int(v) for v in vertices), set(int(e) for e in edges)
    for e in range(g.ne):
        if e not in es and int(g.esrc[e]) in vs and int(g.edst[e]) in vs:
            raise NotImplementedError("region with a loop")
    tree_nb = {v: [] for v in vs}
    for e in es:
        a, b = int(g.esrc[e]), int(g.edst[e])
        tree_nb[a].append(b); tree_nb[b].append(a)

    def absorbed(v: int, parent: Optional[int]) -> np.ndarray:
        t = bpc.network.tensors[v]
        ket = np.tensordot(ops[v], t, axes=([1], [0])) if v in ops else t
        for axis, m in incoming(bpc, v, ignore=tuple(tree_nb[v])):           # messages from outside the region
            ket = _absorb(ket, axis, m)
        for c in tree_nb[v]:
            if c != parent:
                ket = _absorb(ket, 1 + g.leg(v, g.edge_of[(c, v)]), env(c, v))
        return ket

    def env(v: int, parent: int) -> np.ndarray:
        ket = absorbed(v, parent)
        ax = 1 + g.leg(v, g.edge_of[(v, parent)])
        others = [j for j in range(ket.ndim) if j != ax]
        return np.tensordot(ket, bpc.bra_tensor(v).conj(), axes=(others, others))      # [b, b']

    root = min(vs)
    return complex(np.vdot(bpc.bra_tensor(root), absorbed(root, None)))


def expect(bpc: BeliefPropagationCache, obs: tuple, tree=None) -> complex:
    """obs = (op string or list of op names, vertex names/ids[, coeff]); src/expect.jl:58-83,160-176.  tree = (vertices, edge ids)
    of the Steiner tree to contract over for three or more vertices (default graphs.steiner_tree)."""
    g = bpc.g
    op, verts = obs[0], obs[1]
    coeff = obs[2] if len(obs) > 2 else 1
    verts = verts if isinstance(verts, list) else [verts]
    vs = [_vid(g, x) for x in verts]
    names = list(op) if isinstance(op, str) else list(op)
    if len(names) != len(vs):
        raise ValueError("Invalid observable: need as many operators as vertices passed.")
    if coeff == 0:
        return 0
    if len(vs) == 1:
        region = vs
    elif len(vs) == 2:
        region = _tree_path(g, vs[0], vs[1])
    else:
        tv, te = tree if tree is not None else steiner_tree(g, vs)
        ops = {v: _gates.PAULI[n] for v, n in zip(vs, names)}
        return coeff * _tree_region_contract(bpc, tv, te, ops) / _tree_region_contract(bpc, tv, te, {})
    ops = {v: _gates.PAULI[n] for v, n in zip(vs, names)}
    denom = _region_contract(bpc, region, {})
    numer = _region_contract(bpc, region, ops)
    return coeff * numer / denom


def expect_tns(psi: TensorNetworkState, obs, cache_update_kwargs: Optional[dict] = None):
    """src/expect.jl:124-136: build a cache, update with default_bp_update_kwargs, measure."""
    kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(psi)
    bpc = update(BeliefPropagationCache(psi), **kw)
    if isinstance(obs, list):
        return [expect(bpc, o) for o in obs]
    return expect(bpc, obs)


def reduced_density_matrix(bpc: BeliefPropagationCache, vs: Sequence[int], tree=None) -> np.ndarray:
    """Normalised BP reduced density matrix of the vertices vs (src/rdm.jl:49-70) as rho[(s_1..s_n), (s_1'..s_n')], first vertex
    slow: the region (shortest path / Steiner tree) is contracted once per matrix element with |s><s'| inserted on the sites
    (the reference leaves those site legs open in a single contraction)."""
    g = bpc.g
    n = len(vs)
    if n == 1:
        return rdm_1site(bpc, vs[0])
    if n == 2:
        tv = _tree_path(g, vs[0], vs[1])
        te = [g.edge_of[(tv[i], tv[i + 1])] for i in range(len(tv) - 1)]
    else:
        tv, te = tree if tree is not None else steiner_tree(g, vs)
    dims = [bpc.network.tensors[v].shape[0] for v in vs]
    big = int(np.prod(dims))
    rho = np.zeros((big, big), dtype=np.complex128)
    for i, si in enumerate(np.ndindex(*dims)):
        for j, sj in enumerate(np.ndindex(*dims)):
            ops = {}
            for v, d, a, b in zip(vs, dims, si, sj):
                o = np.zeros((d, d), dtype=np.complex128)
                o[b, a] = 1.0                               # |s'><s|: tr(rho |s'><s|) = rho[s, s']
                ops[v] = o
            rho[i, j] = _tree_region_contract(bpc, tv, te, ops)
    return rho / np.trace(rho)


def rdm_1site(bpc: BeliefPropagationCache, v: int) -> np.ndarray:
    """BP one-site reduced density matrix rho[s, s'] (src/rdm.jl:49-70 restricted to one vertex), trace 1."""
    t = bpc.network.tensors[v]
    a = t
    for axis, m in incoming(bpc, v):
        a = _absorb(a, axis, m)
    others = list(range(1, t.ndim))
    rho = np.tensordot(a, t.conj(), axes=(others, others))
    return rho / np.trace(rho)


# ---------------------------------------------------------------------------------------------
# symmetric gauge, entanglement, BP truncation
# ---------------------------------------------------------------------------------------------
def symmetric_gauge(bpc: BeliefPropagationCache, regularization: float = 10 * EPS) -> BeliefPropagationCache:
    """src/symmetric_gauge.jl:1-56 (SURVEY.md Appendix A.8 for the index roles)."""
    bpc = bpc.copy()
    g = bpc.g
    for e in range(g.ne):
        a, c = int(g.esrc[e]), int(g.edst[e])
        xw, xu = np.linalg.eigh(message(bpc, 2 * e))
        yw, yu = np.linalg.eigh(message(bpc, 2 * e + 1))
        xw, yw = xw + regularization, yw + regularization                   # :14-15 (a shift)
        if np.any(xw < 0) or np.any(yw < 0):
            raise NegativeEigenvalueError("negative message eigenvalue in symmetric_gauge")
        root_x = (xu * np.sqrt(xw)) @ xu.conj().T
        root_y = (yu * np.sqrt(yw)) @ yu.conj().T
        inv_x = (xu / np.sqrt(xw)) @ xu.conj().T
        inv_y = (yu / np.sqrt(yw)) @ yu.conj().T
        ce = root_x.T @ root_y                                              # :27-28
        u, s, vh = np.linalg.svd(ce)                                        # :30
        rs = np.sqrt(s)
        ax_a, ax_c = 1 + g.leg(a, e), 1 + g.leg(c, e)
        ta = _absorb(bpc.network.tensors[a], ax_a, inv_x.T)                 # T~[.., l] = sum_b T[.., b] X^-1/2[l, b]
        tc = _absorb(bpc.network.tensors[c], ax_c, inv_y.T)
        bpc.network.tensors[a] = _absorb(ta, ax_a, u * rs)                  # :34, :45-46
        bpc.network.tensors[c] = _absorb(tc, ax_c, vh.T * rs)               # :35-36, :47
        bpc.messages[2 * e] = np.diag(s).astype(np.complex128)              # :51-52
        bpc.messages[2 * e + 1] = np.diag(s).astype(np.complex128)
    return bpc


def gauge_and_scale(psi: TensorNetworkState, maxiter: int = 40, edge_sequence=None) -> TensorNetworkState:
    """symmetrize_and_bpnormalize (src/symmetric_gauge.jl:76-83): update(maxiter=40), rescale, symmetric gauge."""
    bpc = update(BeliefPropagationCache(psi), maxiter=maxiter, edge_sequence=edge_sequence)
    bpc = rescale(bpc)
    return symmetric_gauge(bpc).network


def entanglement(bpc: BeliefPropagationCache, e: int) -> float:
    """src/symmetric_gauge.jl:85-103."""
    r1, _ = pseudo_sqrt_inv_sqrt(message(bpc, 2 * e))
    r2, _ = pseudo_sqrt_inv_sqrt(message(bpc, 2 * e + 1))
    s = np.linalg.svd(r1 @ r2.T, compute_uv=False)
    s = s / np.linalg.norm(s)
    ee = 0.0
    for d in s:
        if abs(d) >= EPS:
            ee -= d * d * math.log(d * d)
    return abs(ee)


def truncate_bp(bpc: BeliefPropagationCache, maxdim: int, cutoff: Optional[float] = None,
                normalize_tensors: bool = True, edge_groups: Optional[List[List[int]]] = None,
                bp_update_kwargs: Optional[dict] = None) -> BeliefPropagationCache:
    """src/truncate.jl:12-38: identity gates per colour group, a BP update after each group."""
    bpc = bpc.copy()
    g = bpc.g
    kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(bpc.network)
    if edge_groups is None:
        edge_groups = edge_color(g)
    ident = np.eye(4, dtype=np.complex128).reshape(2, 2, 2, 2)
    for grp in edge_groups:
        for e in grp:
            if bpc.network.bond_dim(e) == 1:                       # truncatable_edge, truncate.jl:5-10
                continue
            apply_gate(bpc, ident, [int(g.esrc[e]), int(g.edst[e])], maxdim=maxdim, cutoff=cutoff,
                       normalize_tensors=normalize_tensors)
        bpc = update(bpc, **kw)
    return bpc


def truncate_tns(psi: TensorNetworkState, **kwargs) -> TensorNetworkState:
    """src/truncate.jl:74-79: bare update (25 sweeps / 1 on trees, no tolerance) then truncate."""
    bpc = update(BeliefPropagationCache(psi))
    return truncate_bp(bpc, **kwargs).network


def sample(psi: TensorNetworkState, nsamples: int, bp_update_kwargs: Optional[dict] = None, gauge_state: bool = True, rng=None):
    """src/sampling.jl:3-46 (BP branch).  The projection keeps the site index and zeroes the other components, which gives
    the same BP contractions as the reference's removal of the index (norm_factors, tensornetworkstate.jl:48-49)."""
    rng = rng if rng is not None else np.random.default_rng()
    kw = dict(bp_update_kwargs or {})
    g = psi.g
    bpc = update(BeliefPropagationCache(psi), **kw)
    if gauge_state:
        bpc = symmetric_gauge(rescale(bpc))
    out = []
    for _ in range(nsamples):
        proj = bpc.copy()
        bits = {}
        for v in range(g.nv):
            rho = rdm_1site(proj, v)
            p = np.real(np.diag(rho))
            p = np.clip(p / p.sum(), 0.0, None)
            config = int(min(np.searchsorted(np.cumsum(p), rng.random() * p.sum(), side="right"), len(p) - 1))
            bits[g.names[v]] = config
            op = np.zeros((len(p), len(p)), dtype=np.complex128)
            op[config, config] = 1.0
            apply_gate(proj, op, [v], normalize_tensors=False)
            if v != g.nv - 1:
                proj = update(proj, **kw)
        out.append(bits)
    return out, psi


# ---------------------------------------------------------------------------------------------
# single-layer networks: contract(tn; alg = "bp") (src/contract.jl:7-9)
# ---------------------------------------------------------------------------------------------
def contract_bp(g: Graph, tensors: Sequence[np.ndarray], maxiter: Optional[int] = None, tolerance: Optional[float] = DEFAULT_TOL_F64,
                edge_sequence: Optional[Sequence[int]] = None, info: Optional[dict] = None) -> complex:
    """BP contraction of a plain TensorNetwork (no site indices; tensors[v] = A_v[b_e1 .. b_ez]): messages are vectors on the
    bonds, default = all ones (`delta` of a single index, src/TensorNetworks/tensornetwork.jl:62-64), factors = the tensors
    themselves (:66-68); update / normalisation / diff / partition function as for states
    (src/MessagePassing/abstractbeliefpropagationcache.jl:144-218,252-267).  Written directly with vector messages, independent of
    the bilinear-form embedding the engine uses for the same quantity."""
    if maxiter is None:
        maxiter = 1 if g.is_tree() else DEFAULT_BP_MAXITER
    seq = [int(k) for k in (edge_sequence if edge_sequence is not None else forest_cover_edge_sequence(g))]
    msg = {k: np.ones(tensors[g.dsrc(k)].shape[g.leg(g.dsrc(k), k // 2)], dtype=np.complex128) for k in range(2 * g.ne)}

    def absorbed(v: int, skip: Optional[int]) -> np.ndarray:
        a = tensors[v]
        for i in reversed(range(a.ndim)):                        # contract from the last leg so that the axes keep their numbers
            e = g.incident[v][i]
            u = g.other(e, v)
            if u == skip:
                continue
            a = np.tensordot(a, msg[g.dedge(u, v)], axes=([i], [0]))
        return a

    iters, mean = 0, float("nan")
    for it in range(1, maxiter + 1):
        iters, diff = it, 0.0
        for k in seq:
            new = absorbed(g.dsrc(k), g.ddst(k))
            s = new.sum()
            if s != 0:
                new = new / s
            if tolerance is not None:
                diff += message_diff(new, msg[k])
            msg[k] = new
        if tolerance is not None:
            mean = diff / len(seq)
            if mean <= tolerance:
                break
    if info is not None:
        info["iters"], info["mean_diff"], info["messages"] = iters, mean, msg
    num = np.array([complex(absorbed(v, None)) for v in range(g.nv)])
    den = np.array([np.sum(msg[2 * e] * msg[2 * e + 1]) for e in range(g.ne)])
    if np.any(den == 0):
        return 0.0
    return complex(np.exp(np.sum(np.log(num.astype(np.complex128))) - np.sum(np.log(den.astype(np.complex128)))))


def contract_exact(g: Graph, tensors: Sequence[np.ndarray]) -> complex:
    """Full contraction of a small single-layer network (independent ground truth)."""
    return complex(to_statevector(TensorNetworkState(g, [t[None, ...] for t in tensors]))[0])


# ---------------------------------------------------------------------------------------------
# loop corrections (src/MessagePassing/loopcorrection.jl, loop_correlation(s): beliefpropagationcache.jl:144-197)
# ---------------------------------------------------------------------------------------------
def _contract_network(tensors: List[np.ndarray], labels: List[List[int]]) -> np.ndarray:
    """Greedy pairwise contraction of a closed or open network: labels[i] names the axes of tensors[i]; a label that appears
    twice is summed.  The pair with the smallest result goes first (the reference asks for a greedy sequence too:
    loopcorrection.jl:88).  Returns the tensor over the labels that are left, in ascending label order."""
    ts = [(np.asarray(t), list(l)) for t, l in zip(tensors, labels)]
    while len(ts) > 1:
        best = None
        for i in range(len(ts)):
            for j in range(i + 1, len(ts)):
                shared = set(ts[i][1]) & set(ts[j][1])
                if not shared and best is not None:
                    continue
                size = 1
                for t, l in (ts[i], ts[j]):
                    for ax, lab in enumerate(l):
                        if lab not in shared:
                            size *= t.shape[ax]
                key = (0 if shared else 1, size)
                if best is None or key < best[0]:
                    best = (key, i, j, shared)
        _, i, j, shared = best
        (a, la), (b, lb) = ts[i], ts[j]
        shared = sorted(shared)
        out = np.tensordot(a, b, axes=([la.index(x) for x in shared], [lb.index(x) for x in shared]))
        lout = [x for x in la if x not in shared] + [x for x in lb if x not in shared]
        ts = [ts[k] for k in range(len(ts)) if k not in (i, j)] + [(out, lout)]
    t, l = ts[0]
    return np.transpose(t, np.argsort(l)) if l else t


def loop_weight(bpc: BeliefPropagationCache, edges: Sequence[int]) -> complex:
    """weight(bpc, eg) (loopcorrection.jl:81-91) for the subgraph induced by `edges`: the factors of its vertices, the messages
    entering them along every edge that is not in the subgraph, and on every edge of the subgraph the antiprojector
    delta - message(e) (x) message(reverse e) over the doubled (ket, bra) bond (loopcorrection.jl:49-60).  The cache is expected to
    be rescaled (edge and vertex scalars 1)."""
    g = bpc.g
    es = sorted(set(int(e) for e in edges))
    vs = sorted(set(int(x) for e in es for x in (g.esrc[e], g.edst[e])))
    nlab = [0]

    def fresh() -> int:
        nlab[0] += 1
        return nlab[0] - 1

    ket_lab = {(v, e): fresh() for v in vs for e in g.incident[v]}
    bra_lab = {(v, e): fresh() for v in vs for e in g.incident[v]}
    site_lab = {v: fresh() for v in vs}
    ts, ls = [], []
    for v in vs:
        ts.append(bpc.network.tensors[v]); ls.append([site_lab[v]] + [ket_lab[(v, e)] for e in g.incident[v]])
        ts.append(bpc.bra_tensor(v).conj()); ls.append([site_lab[v]] + [bra_lab[(v, e)] for e in g.incident[v]])
        for e in g.incident[v]:
            if e not in es:
                ts.append(message(bpc, g.dedge(g.other(e, v), v))); ls.append([ket_lab[(v, e)], bra_lab[(v, e)]])
    for e in es:
        a, b = int(g.esrc[e]), int(g.edst[e])
        m_ab, m_ba = message(bpc, 2 * e), message(bpc, 2 * e + 1)          # a -> b enters b; b -> a enters a
        n, nb = m_ab.shape
        ap = np.einsum("ik,jl->ijkl", np.eye(n), np.eye(nb)) - np.einsum("ij,kl->ijkl", m_ab, m_ba)
        ts.append(ap); ls.append([ket_lab[(b, e)], bra_lab[(b, e)], ket_lab[(a, e)], bra_lab[(a, e)]])
    return complex(_contract_network(ts, ls))


def loop_weights(bpc: BeliefPropagationCache, egs: Sequence[Sequence[int]]) -> np.ndarray:
    """loopcorrection.jl:94-96."""
    return np.array([loop_weight(bpc, eg) for eg in egs], dtype=np.complex128)


def loopcorrected_partitionfunction(bpc: BeliefPropagationCache, max_configuration_size: int) -> complex:
    """loopcorrection.jl:3-16: Z_bp * (1 + sum of the weights of every leafless edge-induced subgraph of at most
    max_configuration_size edges), weights taken on the rescaled cache."""
    zbp = partitionfunction(bpc)
    bpc = rescale(bpc)
    egs = edgeinduced_subgraphs_no_leaves(bpc.g, int(max_configuration_size))
    if not egs:
        return zbp
    return zbp * (1 + complex(np.sum(loop_weights(bpc, egs))))


def norm_sqr_loopcorrected(psi: TensorNetworkState, max_configuration_size: int, maxiter: Optional[int] = None,
                           tolerance: Optional[float] = None) -> complex:
    """norm_sqr(psi; alg = "loopcorrections", max_configuration_size) (src/norm_sqr.jl:56-76)."""
    kw = default_bp_update_kwargs(psi)
    bpc = update(BeliefPropagationCache(psi.copy()), maxiter=maxiter if maxiter is not None else kw["maxiter"],
                 tolerance=tolerance if tolerance is not None else kw["tolerance"])
    return loopcorrected_partitionfunction(bpc, max_configuration_size)


def loop_transfer_matrix(bpc: BeliefPropagationCache, loop: Sequence[int], target: int) -> np.ndarray:
    """The matrix whose spectrum loop_correlation looks at (beliefpropagationcache.jl:145-179): `loop` + [`target`] are the
    directed edges (ids k = 2e + dir) of a closed walk v0 -> v1 -> .. -> v0 whose last step is `target`; the factors of its
    vertices are contracted with every message that enters from outside, with the bond of `target` cut open at its source:
    rows = the (ket, bra) pair the source keeps, columns = the pair on the far side (ket index fastest in both)."""
    g = bpc.g
    ks = [int(k) for k in loop] + [int(target)]
    es = [k // 2 for k in ks]
    vs = [g.dsrc(k) for k in ks]
    if [g.ddst(k) for k in ks] != vs[1:] + vs[:1] or len(set(vs)) != len(vs):
        raise ValueError("loop + target must be a simple closed walk")
    nlab = [0]

    def fresh() -> int:
        nlab[0] += 1
        return nlab[0] - 1

    et = es[-1]
    src = vs[-1]
    chi, chib = message(bpc, target).shape
    # open labels first so that they come out as the leading axes: (far ket, far bra, src ket, src bra)
    far_k, far_b, src_k, src_b = fresh(), fresh(), fresh(), fresh()
    bond_k = {e: fresh() for e in es[:-1]}
    bond_b = {e: fresh() for e in es[:-1]}
    ts, ls = [], []
    for v in vs:
        sl = fresh()
        lk, lb = [sl], [sl]
        for e in g.incident[v]:
            if e == et:
                lk.append(src_k if v == src else far_k); lb.append(src_b if v == src else far_b)
            elif e in bond_k:
                lk.append(bond_k[e]); lb.append(bond_b[e])
            else:
                a, b = fresh(), fresh()
                lk.append(a); lb.append(b)
                ts.append(message(bpc, g.dedge(g.other(e, v), v))); ls.append([a, b])
        ts.append(bpc.network.tensors[v]); ls.append(lk)
        ts.append(bpc.bra_tensor(v).conj()); ls.append(lb)
    t = _contract_network(ts, ls)                       # axes (far ket, far bra, src ket, src bra)
    t = np.transpose(t, (2, 3, 0, 1))
    return t.reshape(chi * chib, chi * chib, order="F")


def loop_correlation(bpc: BeliefPropagationCache, loop: Sequence[int], target: int) -> float:
    """beliefpropagationcache.jl:145-179: 1 - |lambda_1| / sum |lambda_i| of the loop's transfer matrix; 0 on a tree."""
    if bpc.g.is_tree():
        return 0.0
    lam = np.abs(np.linalg.eigvals(loop_transfer_matrix(bpc, loop, target)))
    return float(1.0 - lam.max() / lam.sum())


def loop_correlations(bpc: BeliefPropagationCache, smallest_loop_size: int) -> List[float]:
    """beliefpropagationcache.jl:182-192: one value per simple cycle of at most `smallest_loop_size` edges."""
    out = []
    for loop, target in cycle_walks(bpc.g, smallest_loop_size):
        out.append(loop_correlation(bpc, loop, target))
    return out


# ---------------------------------------------------------------------------------------------
# exact helpers for tiny systems (independent ground truth, not part of the restatement)
# ---------------------------------------------------------------------------------------------
def to_statevector(psi: TensorNetworkState) -> np.ndarray:
    """Contract all bonds; returns psi[s_0, s_1, ..] flattened with vertex 0 slowest."""
    g = psi.g
    done: List[int] = []
    labels: List = []
    big = None
    order = list(range(g.nv))
    for v in order:
        t = psi.tensors[v]
        tl = [("s", v)] + [("e", e) for e in g.incident[v]]
        if big is None:
            big, labels = t, tl
        else:
            shared = [l for l in tl if l in labels]
            big = np.tensordot(big, t, axes=([labels.index(l) for l in shared], [tl.index(l) for l in shared]))
            labels = [l for l in labels if l not in shared] + [l for l in tl if l not in shared]
        done.append(v)
    perm = [labels.index(("s", v)) for v in range(g.nv)]
    assert len(labels) == g.nv
    return np.transpose(big, perm).reshape(-1)


def statevector_circuit(g: Graph, f, circuit: Sequence[tuple]) -> np.ndarray:
    """Exact evolution of a product state; vertex 0 slowest."""
    n = g.nv
    psi = np.array([1.0 + 0j])
    for v in range(n):
        psi = np.kron(psi, _gates.state_vector(f(g.names[v])))
    psi = psi.reshape((2,) * n)
    for arr, vs in _resolve_circuit(g, circuit):
        if len(vs) == 1:
            psi = np.moveaxis(np.tensordot(arr, psi, axes=([1], [vs[0]])), 0, vs[0])
        else:
            psi = np.moveaxis(np.tensordot(arr, psi, axes=([2, 3], [vs[0], vs[1]])), [0, 1], [vs[0], vs[1]])
    return psi.reshape(-1)


def statevector_expect(psi: np.ndarray, n: int, ops: Dict[int, np.ndarray]) -> complex:
    phi = psi.reshape((2,) * n)
    for v, o in ops.items():
        phi = np.moveaxis(np.tensordot(o, phi, axes=([1], [v])), 0, v)
    return complex(np.vdot(psi, phi.reshape(-1)) / np.vdot(psi, psi))
