"""Heisenberg-picture path (second half): "Pauli" sites (d = 4), Pauli-transfer-matrix gates
(src/Apply/gate_definitions.jl:60-86), paulitensornetworkstate / identitytensornetworkstate
(src/TensorNetworks/tensornetworkstate_constructors.jl:1,19-35) and inner(psi, phi; alg = "bp") between layers of different bond
dimensions (examples/2dIsing_dynamics_Heisenbergpicture.jl:55-56).  CPU: the host-side gate definitions and the oracle against
dense operator algebra and the Schroedinger-picture statevector."""
import numpy as np
import pytest

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates

PAULIS = [gates.I2, gates.X, gates.Y, gates.Z]


def pauli_coefficients(op, n):
    """c[a_1..a_n] = tr(P_a op) / 2^n, vertex 0 = first (slow) Kronecker factor."""
    c = np.empty((4,) * n, dtype=np.complex128)
    for idx in np.ndindex(*c.shape):
        p = np.array([[1.0 + 0j]])
        for a in idx:
            p = np.kron(p, PAULIS[a])
        c[idx] = np.trace(p @ op) / 2 ** n
    return c


def dense_unitary(n, circuit):
    """Product of the qubit gates of `circuit` (first gate applied first) as a 2^n x 2^n matrix, vertex 0 slow."""
    u = np.eye(2 ** n, dtype=np.complex128)
    for gate in circuit:
        arr = gates.gate_array(gate)
        vs = gate[1]
        t = u.reshape((2,) * n + (2 ** n,))
        if len(vs) == 1:
            t = np.moveaxis(np.tensordot(arr, t, axes=([1], [vs[0]])), 0, vs[0])
        else:
            t = np.moveaxis(np.tensordot(arr, t, axes=([2, 3], [vs[0], vs[1]])), [0, 1], [vs[0], vs[1]])
        u = t.reshape(2 ** n, 2 ** n)
    return u


def test_pauli_transfer_matrix_is_the_heisenberg_map():
    rs = np.random.default_rng(3)
    for name, nq, theta in (("Rz", 1, 0.37), ("Ry", 1, -0.8), ("Rxx", 2, 0.61), ("Rzz", 2, 1.3), ("Rxy", 2, 0.45)):
        m = gates.heisenberg_gate_array((name, list(range(nq)), theta))
        assert m.shape == (4,) * (2 * nq) and np.abs(m.imag).max() == 0
        u = gates.pauli_rotation_matrix(name[1:], theta)
        c = rs.normal(size=(4,) * nq)
        op = sum(c[idx] * (np.kron(PAULIS[idx[0]], PAULIS[idx[1]]) if nq == 2 else PAULIS[idx[0]]) for idx in np.ndindex(*c.shape))
        want = pauli_coefficients(u.conj().T @ op @ u, nq)
        got = np.tensordot(m, c, axes=(list(range(nq, 2 * nq)), list(range(nq))))
        assert np.max(np.abs(got - want)) < 1e-14
        mm = m.reshape(4 ** nq, 4 ** nq).real
        assert np.max(np.abs(mm @ mm.T - np.eye(4 ** nq))) < 1e-14                      # orthogonal: the Frobenius norm of O is kept
    # the qubit gate of the same tuple is the same unitary (Rxx/Ryy/Rzz tuples are rescaled on the qubit branch only)
    assert np.allclose(gates.gate_array(("Rxx", [0, 1], 0.61)).reshape(4, 4), gates.pauli_rotation_matrix("XX", 0.61))
    assert np.allclose(gates.gate_array(("Rz", [0], 0.37)), gates.pauli_rotation_matrix("Z", 0.37))
    with pytest.raises(ValueError):
        gates.gate_array(("CNOT", [0, 1]), site_dims=[4, 4])


def ising_layers(g, nlayers, dt=0.3, h=-1.0, J=-1.0):
    """The Heisenberg example's layer (examples/2dIsing_dynamics_Heisenbergpicture.jl:29-35) as vertex-id tuples."""
    layer = [("Rz", [v], h * dt) for v in range(g.nv)]
    for grp in G.edge_color(g):
        layer += [("Rxx", [int(g.esrc[e]), int(g.edst[e])], 2 * J * dt) for e in grp]
    layer += [("Rz", [v], h * dt) for v in range(g.nv)]
    return layer * nlayers


def test_heisenberg_evolution_on_a_tree_matches_dense_operator_algebra():
    # comb tree, no truncation: the evolved operator network holds exactly the Pauli coefficients of U^dag Z_c U (gates applied
    # in the list's order: the LAST gate of the list is the innermost conjugation), and inner(.,.; alg = "bp") is exact
    g = G.named_comb_tree((2, 3))
    n, c = g.nv, 2
    circ = ising_layers(g, 2)
    o0 = O.paulitensornetworkstate(lambda v: "Z" if g.v(v) == c else "I", g)
    ot, errs = O.apply_gates(circ, o0, apply_kwargs=dict(maxdim=64, cutoff=1e-14, normalize_tensors=False))
    assert np.max(errs) < 1e-12 and all(t.shape[0] == 4 for t in ot.tensors)
    got = O.to_statevector(ot).reshape((4,) * n)
    u = dense_unitary(n, circ[::-1])                      # M_L ... M_1 c  <->  U_1^dag ... U_L^dag O U_L ... U_1 has U = U_1 U_2 ... U_L
    zc = np.array([[1.0 + 0j]])
    for v in range(n):
        zc = np.kron(zc, gates.Z if v == c else gates.I2)
    ut = u.conj().T @ zc @ u
    want = pauli_coefficients(ut, n)
    assert np.max(np.abs(got - want)) < 1e-12
    # tr(O(t) O(0)) / 2^n and tr(O(t)) / 2^n through BP on the bilinear form, layers of different bond dimensions
    assert ot.maxvirtualdim() > 1 and o0.maxvirtualdim() == 1
    tr_t0 = O.inner(ot, o0)
    assert abs(tr_t0 - np.trace(ut @ zc) / 2 ** n) < 1e-12
    ident = O.identitytensornetworkstate(g)
    assert abs(O.inner(ot, ident) - np.trace(ut) / 2 ** n) < 1e-12
    # <0|O(t)|0> = the Schroedinger-picture <Z_c> of the circuit run the other way round
    proj = O.tensornetworkstate(lambda v: [1, 0, 0, 1], g)                         # <0|P|0> = 1 for I and Z
    psi = O.statevector_circuit(g, lambda v: "↑", circ[::-1])
    want_z = O.statevector_expect(psi, n, {c: gates.Z})
    assert abs(O.inner(ot, proj) - want_z) < 1e-12
    assert abs(want_z - 1) > 1e-3                                                   # the circuit does something


def test_bilinear_default_message_and_zero_padding():
    # reference default of a bilinear form: a single [1, 1] entry (src/Forms/bilinearform.jl:39-42); embedding the thinner layer
    # in zero-padded tensors (what the engine does before tnqs_set_bra) leaves every sweep's partition function unchanged
    g = G.named_grid((3, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=5)
    phi = O.random_tensornetworkstate(g, bond_dimension=2, seed=6)
    b = O.BeliefPropagationCache(psi, {}, phi, delta_default=True)
    m = O.message(b, 0)
    assert m.shape == (3, 2) and m[0, 0] == 1 and np.count_nonzero(m) == 1
    padded = O.TensorNetworkState(g, [np.pad(t, [(0, 0)] + [(0, 1)] * (t.ndim - 1)) for t in phi.tensors])
    seq = G.forest_cover_edge_sequence(g)
    for sweeps in (1, 4):
        z0 = O.partitionfunction(O.update(O.BeliefPropagationCache(psi, {}, phi, delta_default=True), maxiter=sweeps, edge_sequence=seq))
        z1 = O.partitionfunction(O.update(O.BeliefPropagationCache(psi, {}, padded, delta_default=True), maxiter=sweeps, edge_sequence=seq))
        assert abs(z0 - z1) < 1e-13 * abs(z0)
    b3 = O.update(O.BeliefPropagationCache(psi, {}, padded, delta_default=True), maxiter=2, edge_sequence=seq)
    b2 = O.update(O.BeliefPropagationCache(psi, {}, phi, delta_default=True), maxiter=2, edge_sequence=seq)
    for k in range(2 * g.ne):
        assert np.max(np.abs(O.message(b3, k)[:, :2] - O.message(b2, k))) < 1e-12 and np.all(O.message(b3, k)[:, 2:] == 0)


def test_engine_packs_heisenberg_gates_for_pauli_sites():
    # host logic only (no device): the circuit packer picks the transfer matrices when the sites have dimension 4
    from tensornetworkquantumsimulator_b200 import engine as E
    g = G.named_grid((2, 2))
    circ = [("Rz", [(1, 1)], 0.3), ("Rxx", [(1, 1), (1, 2)], 0.5)]
    gv, off, data = E._pack_circuit(g, circ, lambda: np.full(g.nv, 4, dtype=np.int32))
    assert list(off) == [0, 16] and data.size == 2 * (16 + 256)
    m = gates.heisenberg_gate_array(("Rxx", [0, 1], 0.5))
    assert np.array_equal(data.view(np.complex128)[16:], m.ravel(order="F"))
    gv2, off2, data2 = E._pack_circuit(g, circ, lambda: np.full(g.nv, 2, dtype=np.int32))
    assert list(off2) == [0, 4] and data2.size == 2 * (4 + 16)


# ---- observables on three or more vertices: contraction over the Steiner tree of the support -----------------
def _composed_tree_expect(bpc, vs, ops, tv, te):
    """The engine's composition (engine._expect_tree) replayed with oracle primitives: operators applied to a copy of the ket, the
    original state as the bra, one un-normalised sweep over the tree's child -> parent edges, the root's vertex scalar."""
    g = bpc.g
    root = min(tv)
    seq = G.tree_schedule(g, tv, te, root)
    vals = []
    for with_ops in (True, False):
        ket = bpc.network.copy()
        if with_ops:
            for v, o in zip(vs, ops):
                ket.tensors[v] = np.tensordot(o, ket.tensors[v], axes=([1], [0]))
        work = O.BeliefPropagationCache(ket, {k: m.copy() for k, m in bpc.messages.items()}, bpc.network if with_ops else None)
        work = O.update(work, maxiter=1, tolerance=None, edge_sequence=seq, normalize=False)
        vals.append(O.vertex_scalar(work, root))
    return vals[0] / vals[1]


def test_multi_site_observables_over_the_steiner_tree():
    from util import ising_layer
    # (1) on a tree BP is exact: three- and four-site observables against the statevector
    g = G.named_comb_tree((3, 3))
    circ = ising_layer(g, dt=0.2) * 2
    psi, _ = O.apply_gates(circ, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=dict(maxdim=16, cutoff=None, normalize_tensors=True))
    sv = O.statevector_circuit(g, lambda v: "↑", circ)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=1)
    for names, vs in (("ZXZ", [0, 4, 8]), ("XYZX", [2, 3, 6, 7]), ("ZZZ", [0, 1, 2])):
        want = O.statevector_expect(sv, g.nv, {v: gates.PAULI[n] for v, n in zip(vs, names)})
        got = O.expect(bpc, (names, vs))
        assert abs(got - want) < 1e-12, (names, got, want)
        tv, te = G.steiner_tree(g, vs)
        assert set(vs) <= set(tv) and len(te) == len(tv) - 1
        assert abs(_composed_tree_expect(bpc, vs, [gates.PAULI[n] for n in names], tv, te) - want) < 1e-12
    # (2) loopy grid: the message-passing composition equals the recursive region contraction, the two-vertex case equals the
    # path contraction, and a region with a chord is refused
    g = G.named_grid((4, 4))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=9)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=30, tolerance=1e-12)
    vs = [g.v((1, 1)), g.v((1, 3)), g.v((3, 2))]
    tv, te = G.steiner_tree(g, vs)
    a = O.expect(bpc, ("ZXZ", vs), tree=(tv, te))
    b = _composed_tree_expect(bpc, vs, [gates.Z, gates.X, gates.Z], tv, te)
    assert abs(a - b) < 1e-12 and abs(a.imag) < 1e-10
    pv = O._tree_path(g, vs[0], vs[2])
    pe = [g.edge_of[(pv[i], pv[i + 1])] for i in range(len(pv) - 1)]
    two = O._tree_region_contract(bpc, pv, pe, {vs[0]: gates.Z, vs[2]: gates.X}) / O._tree_region_contract(bpc, pv, pe, {})
    assert abs(two - O.expect(bpc, ("ZX", [vs[0], vs[2]]))) < 1e-12
    plaquette = [g.v((1, 1)), g.v((1, 2)), g.v((2, 2)), g.v((2, 1))]
    edges3 = [g.edge_of[(plaquette[i], plaquette[i + 1])] for i in range(3)]
    with pytest.raises(NotImplementedError):
        G.tree_schedule(g, plaquette, edges3, plaquette[0])
    with pytest.raises(NotImplementedError):
        O.expect(bpc, ("ZZZZ", plaquette), tree=(plaquette, edges3))
