"""Randomised pins of the oracle (hypothesis): on random trees with ragged bond dimensions one BP sweep in forest-cover order is
exact (test/test_beliefpropagation.jl:26-53 generalised), and random circuits without truncation reproduce the statevector
(test/test_apply.jl:20,53: norm preservation, generalised to the full state).  Sized to run in seconds."""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates
from util import ragged_state, statevector_fidelity


@st.composite
def random_tree(draw):
    n = draw(st.integers(2, 7))
    parents = [draw(st.integers(0, i - 1)) for i in range(1, n)]
    esrc, edst = parents, list(range(1, n))
    return G.Graph(list(range(n)), esrc, edst)


@settings(max_examples=25, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
@given(random_tree(), st.integers(0, 10 ** 6), st.sampled_from([2, 3]))
def test_bp_is_exact_on_random_ragged_trees(g, seed, d):
    rs = np.random.default_rng(seed)
    dims = rs.integers(1, 4, size=g.ne)
    psi = ragged_state(g, dims, seed=seed % 1000, d=d)
    sv = O.to_statevector(psi)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=1)
    n2 = np.vdot(sv, sv).real
    assert abs(O.partitionfunction(bpc) - n2) < 1e-11 * n2
    v = int(rs.integers(0, g.nv))
    rho = O.rdm_1site(bpc, v)
    t = sv.reshape((d,) * g.nv)
    a = np.moveaxis(t, v, 0).reshape(d, -1)
    exact = a @ a.conj().T
    assert np.max(np.abs(rho - exact / np.trace(exact))) < 1e-11
    # rescale: every vertex and edge scalar becomes one, the state is the same ray
    r = O.rescale(bpc)
    assert np.max(np.abs(np.array([O.vertex_scalar(r, u) for u in range(g.nv)]) - 1)) < 1e-11
    assert np.max(np.abs(np.array([O.edge_scalar(r, e) for e in range(g.ne)]) - 1)) < 1e-11
    assert abs(statevector_fidelity(O.to_statevector(r.network), sv) - 1) < 1e-11


GATES1 = ["Rx", "Ry", "Rz"]
GATES2 = ["Rxx", "Ryy", "Rzz", "CRx", "Rxxyy", "CPHASE"]


@settings(max_examples=15, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
@given(random_tree(), st.integers(0, 10 ** 6))
def test_untruncated_circuits_on_random_trees_match_the_statevector(g, seed):
    rs = np.random.default_rng(seed)
    circ = []
    for _ in range(3 * g.nv):
        if rs.random() < 0.4:
            circ.append((GATES1[rs.integers(len(GATES1))], [int(rs.integers(g.nv))], float(rs.normal())))
        else:
            e = int(rs.integers(g.ne))
            pair = [int(g.esrc[e]), int(g.edst[e])]
            if rs.random() < 0.5:
                pair = pair[::-1]
            circ.append((GATES2[rs.integers(len(GATES2))], pair, float(rs.normal())))
    psi, errs = O.apply_gates(circ, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=dict(maxdim=64, cutoff=None, normalize_tensors=True))
    sv = O.statevector_circuit(g, lambda v: "↑", circ)
    assert abs(statevector_fidelity(O.to_statevector(psi), sv) - 1) < 1e-11
    assert np.max(errs) < 1e-12


def test_vertex_centric_jacobi_sweep_gives_the_same_messages():
    # the shared-absorption evaluation of a JACOBI sweep (round-2 plan) is the same map on messages
    for g, chi in ((G.named_grid((3, 4)), 3), (G.named_grid((2, 2, 2)), 2), (G.named_hexagonal_lattice_graph(1, 2), 4)):
        psi = O.random_tensornetworkstate(g, bond_dimension=chi, seed=3)
        b = O.update(O.BeliefPropagationCache(psi), maxiter=2)
        a1, a2 = O.jacobi_sweep(b), O.jacobi_sweep(b, shared=True)
        for k in range(2 * g.ne):
            assert np.max(np.abs(a1.messages[k] - a2.messages[k])) < 1e-13 * np.max(np.abs(a1.messages[k]))
    assert [O.jacobi_absorptions_per_vertex(z, False) for z in (2, 3, 4, 6)] == [2, 6, 12, 30]
    assert [O.jacobi_absorptions_per_vertex(z, True) for z in (2, 3, 4, 6)] == [2, 5, 9, 20]
