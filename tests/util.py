"""Shared helpers for the parity tests: seeded inputs, oracle <-> engine conversion, gauge-invariant comparisons."""
import numpy as np

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, rng


def ragged_state(g, dims, seed, d=2):
    """Random oracle state with per-edge bond dimensions `dims` (counter-based generator, stream = vertex)."""
    ts = []
    for v in range(g.nv):
        shape = (d,) + tuple(int(dims[e]) for e in g.incident[v])
        ts.append(rng.complex_normal(seed, v, int(np.prod(shape))).reshape(shape, order="F").copy())
    return O.TensorNetworkState(g, ts)


def to_engine(psi_o, chi_cap=None, with_messages=None):
    from tensornetworkquantumsimulator_b200 import engine as E
    psi = E.TensorNetworkState.from_arrays(psi_o.g, psi_o.tensors, chi_cap=chi_cap)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    if with_messages is not None:
        for k, m in with_messages.items():
            E.setmessage(bpc, k, m)
    return bpc


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def two_site(g, tensors, e):
    """Gauge-invariant two-site tensor of edge e: contract the shared bond of T[esrc] and T[edst]."""
    a, c = int(g.esrc[e]), int(g.edst[e])
    return np.tensordot(tensors[a], tensors[c], axes=([1 + g.leg(a, e)], [1 + g.leg(c, e)]))


def ising_layer(g, dt=0.25, hx=1.0, hz=0.8, J=0.5, groups=None):
    layer = [("Rx", [g.names[v]], 2 * hx * dt) for v in range(g.nv)]
    layer += [("Rz", [g.names[v]], 2 * hz * dt) for v in range(g.nv)]
    for grp in (groups or G.edge_color(g)):
        layer += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 2 * J * dt) for e in grp]
    return layer


def statevector_fidelity(a, b):
    return abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)
