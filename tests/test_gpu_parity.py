"""GPU parity tests: the CUDA path (through the C-ABI, via the ctypes host layer) against the CPU oracle on the
same seeded inputs.  Tolerances are BASELINE.json's: ranks / schedules exact, singular values and BP messages
1e-10 relative, expectation values 1e-8 absolute."""
import math
import numpy as np
import pytest

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates
from util import ragged_state, to_engine, rel, two_site, ising_layer, statevector_fidelity

pytestmark = pytest.mark.gpu

MSG_TOL = 1e-10
EXP_TOL = 1e-8


@pytest.fixture(scope="module")
def E():
    from tensornetworkquantumsimulator_b200 import engine
    return engine


GRAPHS = {
    "grid4x4": lambda: G.named_grid((4, 4)),
    "grid3x3x2": lambda: G.named_grid((3, 3, 2)),
    "hex2x2": lambda: G.named_hexagonal_lattice_graph(2, 2),
    "comb3x3": lambda: G.named_comb_tree((3, 3)),
    "heavyhex": lambda: G.heavy_hexagonal_lattice(1, 2),
}


def check_messages(E, bpc_e, bpc_o, tol=MSG_TOL):
    for k in range(2 * bpc_o.g.ne):
        assert rel(E.message(bpc_e, k), O.message(bpc_o, k)) < tol, f"message {k}"


@pytest.mark.parametrize("name", list(GRAPHS))
@pytest.mark.parametrize("chi", [1, 3, 4])
def test_bp_sweeps_match_oracle(E, name, chi):
    g = GRAPHS[name]()
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=11)
    seq = G.forest_cover_edge_sequence(g)
    for sweeps in (1, 3):
        bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=sweeps, edge_sequence=seq)
        be = E.update(to_engine(psi_o), maxiter=sweeps, edge_sequence=seq)
        check_messages(E, be, bo)


@pytest.mark.parametrize("name,chi", [("grid3x3", 8), ("grid3x3", 16), ("hex", 8), ("hex", 16), ("grid4x3", 8)])
def test_bp_dmma_fast_path(E, name, chi):
    # uniform chi in {8, 16}, d = 2, degree 3..4: the FP64 tensor-core kernels (kernels_dmma.cuh) carry the sweep
    g = {"grid3x3": lambda: G.named_grid((3, 3)), "grid4x3": lambda: G.named_grid((4, 3)),
         "hex": lambda: G.named_hexagonal_lattice_graph(2, 1)}[name]()
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=41)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    seq = G.forest_cover_edge_sequence(g)
    io, ie = {}, {}
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=2, tolerance=0.0, edge_sequence=seq, info=io)
    be = E.update(to_engine(psi_o), maxiter=2, tolerance=0.0, edge_sequence=seq, info=ie)
    check_messages(E, be, bo)
    assert ie["mean_diff"] == pytest.approx(io["mean_diff"], rel=1e-8)
    v = g.nv // 2
    assert abs(E.expect(be, ("Z", [g.names[v]])) - O.expect(bo, ("Z", [g.names[v]]))) < EXP_TOL


def test_bp_ragged_bonds_and_tolerance(E):
    g = G.named_grid((4, 3))
    rs = np.random.default_rng(5)
    dims = rs.integers(1, 6, size=g.ne)
    psi_o = ragged_state(g, dims, seed=3)
    io, ie = {}, {}
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=25, tolerance=1e-8, info=io)
    be = E.update(to_engine(psi_o), maxiter=25, tolerance=1e-8, info=ie)
    assert ie["iters"] == io["iters"]
    assert ie["mean_diff"] == pytest.approx(io["mean_diff"], rel=1e-6, abs=1e-14)
    check_messages(E, be, bo)
    # first-sweep diffs are measured against the identity default
    io, ie = {}, {}
    O.update(O.BeliefPropagationCache(psi_o), maxiter=1, tolerance=0.0, info=io)
    E.update(to_engine(psi_o), maxiter=1, tolerance=0.0, info=ie)
    assert ie["mean_diff"] == pytest.approx(io["mean_diff"], rel=1e-10)


def test_bp_options_hermitize_and_unnormalised(E):
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=2)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=2, enforce_hermiticity=True)
    be = E.update(to_engine(psi_o), maxiter=2, enforce_hermiticity=True)
    check_messages(E, be, bo)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=1, normalize=False)
    be = E.update(to_engine(psi_o), maxiter=1, normalize=False)
    check_messages(E, be, bo)


def test_bp_jacobi_mode(E):
    g = G.named_grid((3, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=8)
    bo = O.BeliefPropagationCache(psi_o)
    for _ in range(2):                      # Jacobi: every message from the previous sweep's values
        new = {k: O.updated_message(bo, k) for k in range(2 * g.ne)}
        bo.messages = new
    be = E.update(to_engine(psi_o), maxiter=2, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    check_messages(E, be, bo)
    # same fixed point as the reference order
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=200, tolerance=1e-28)
    be = E.update(to_engine(psi_o), maxiter=400, tolerance=1e-28, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    v = g.v((2, 2))
    assert abs(E.expect(be, ("Z", [(2, 2)])) - O.expect(bo, ("Z", [(2, 2)]))) < EXP_TOL
    del v


def test_bp_exact_on_tree_and_partition_function(E):
    # /root/reference/test/test_beliefpropagation.jl:32-54
    g = G.named_comb_tree((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=123)
    be = E.update(to_engine(psi_o))
    sv = O.to_statevector(psi_o)
    assert E.partitionfunction(be) == pytest.approx(np.vdot(sv, sv).real, rel=1e-11)
    for v in range(g.nv):
        assert abs(E.expect(be, ("Z", [g.names[v]])) - O.statevector_expect(sv, g.nv, {v: gates.Z})) < 1e-12


@pytest.mark.parametrize("name", ["grid4x4", "hex2x2", "grid3x3x2"])
def test_expectation_values(E, name):
    g = GRAPHS[name]()
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=21)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=10)
    be = E.update(to_engine(psi_o), maxiter=10)
    obs = [(p, [g.names[v]]) for v in range(g.nv) for p in "XYZ"]
    got, want = E.expect(be, obs), [O.expect(bo, o) for o in obs]
    assert np.max(np.abs(np.array(got) - np.array(want))) < EXP_TOL
    obs2 = []
    for e in range(0, g.ne, 2):
        a, c = g.names[int(g.esrc[e])], g.names[int(g.edst[e])]
        obs2 += [("ZZ", [a, c]), ("XY", [c, a], 0.5)]
    got, want = E.expect(be, obs2), [O.expect(bo, o) for o in obs2]
    assert np.max(np.abs(np.array(got) - np.array(want))) < EXP_TOL


def test_scalars_and_rescale(E):
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=4)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=5)
    be = E.update(to_engine(psi_o), maxiter=5)
    vs = np.array([O.vertex_scalar(bo, v) for v in range(g.nv)])
    es = np.array([O.edge_scalar(bo, e) for e in range(g.ne)])
    assert rel(E.vertex_scalars(be), vs) < MSG_TOL
    assert rel(E.edge_scalars(be), es) < MSG_TOL
    ro, re_ = O.rescale(bo), E.rescale(be)
    check_messages(E, re_, ro)
    for v in range(g.nv):
        assert rel(re_.psi[v], ro.network.tensors[v]) < MSG_TOL
    assert np.allclose(E.vertex_scalars(re_), 1.0, atol=1e-10)
    assert np.allclose(E.edge_scalars(re_), 1.0, atol=1e-10)


def test_one_site_gates(E):
    g = G.named_grid((3, 2))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=6)
    bo = O.BeliefPropagationCache(psi_o.copy())
    be = to_engine(psi_o)
    arrs = [gates.op_matrix("Rx", 0.3 + 0.1 * v) @ gates.op_matrix("H") for v in range(g.nv)]
    for norm in (True, False):
        for v in range(g.nv):
            O.apply_gate(bo, arrs[v], [v], normalize_tensors=norm)
        E.apply_1site_batch(be, list(range(g.nv)), arrs, normalize_tensors=norm)
        for v in range(g.nv):
            assert rel(be.psi[v], bo.network.tensors[v]) < 1e-13


def gate_case(E, g, psi_o, edges, gate_arrs, kw, sweeps=6, flip=None):
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=sweeps)
    be = E.update(to_engine(psi_o, chi_cap=max(16, psi_o.maxvirtualdim())), maxiter=sweeps)
    details = []
    for i, e in enumerate(edges):
        a, c = int(g.esrc[e]), int(g.edst[e])
        vs = [c, a] if flip and flip[i] else [a, c]
        det = {}
        O.apply_gate(bo, gate_arrs[i], vs, detail=det, **kw)
        details.append(det)
    ranks, errs, sv = E.apply_2site_batch(be, edges, gate_arrs, flip=flip, apply_kwargs=kw)
    for i, e in enumerate(edges):
        det = details[i]
        assert ranks[i] == det["rank"], f"rank of gate {i}"
        assert rel(sv[i, :ranks[i]], det["svals"]) < MSG_TOL
        assert errs[i] == pytest.approx(det["truncerr"], rel=1e-8, abs=1e-16)
        for k in (2 * e, 2 * e + 1):
            assert rel(E.message(be, k), O.message(bo, k)) < MSG_TOL
    te = be.psi.tensors()
    for e in edges:
        assert rel(two_site(g, te, e), two_site(g, bo.network.tensors, e)) < 1e-9
    assert list(be.psi.bond_dims()) == list(bo.network.bond_dims())
    return be, bo


@pytest.mark.parametrize("kw", [dict(maxdim=4, cutoff=1e-10, normalize_tensors=True),
                                dict(maxdim=None, cutoff=None, normalize_tensors=False),
                                dict(maxdim=2, cutoff=None, normalize_tensors=True)])
def test_two_site_gate_batch(E, kw):
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=4, seed=31)
    groups = G.grid_edge_color(g)
    rs = np.random.default_rng(0)
    for grp in groups[:2]:
        arrs = []
        for e in grp:
            h = rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4))
            u, _ = np.linalg.qr(h)
            arrs.append(u.reshape(2, 2, 2, 2))
        gate_case(E, g, psi_o, grp, arrs, kw, flip=[i % 2 for i in range(len(grp))])


@pytest.mark.parametrize("chi", [8, 16])
def test_two_site_gate_uniform_chi_tensor_core_paths(E, chi):
    # uniform chi in {8, 16}: fused slab absorption on the matricised tensors, CholeskyQR2, tensor-core leg kernels
    g = G.named_grid((3, 4)) if chi == 16 else G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=61)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    grp = G.grid_edge_color(g)[1]
    rs = np.random.default_rng(3)
    arrs = []
    for e in grp:
        h = rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4))
        u, _ = np.linalg.qr(h)
        arrs.append(u.reshape(2, 2, 2, 2))
    gate_case(E, g, psi_o, grp, arrs, dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True), sweeps=3)


@pytest.mark.parametrize("shape,chi", [((3, 3), 20), ((3, 2), 32), ((4, 3), 24)])
def test_two_site_gate_large_bond(E, shape, chi):
    # d*chi in 33..64: tensor-core leg / Gram kernels with 3-4 column groups, CholeskyQR2 at nc = 64, 128 x 128 SVD (global W)
    g = G.named_grid(shape)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=83)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    grp = G.grid_edge_color(g)[0]
    rs = np.random.default_rng(4)
    arrs = []
    for e in grp:
        h = rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4))
        u, _ = np.linalg.qr(h)
        arrs.append(u.reshape(2, 2, 2, 2))
    gate_case(E, g, psi_o, grp, arrs, dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True), sweeps=2)


def test_two_site_gate_ragged_and_degree3(E):
    g = G.named_hexagonal_lattice_graph(2, 2)
    dims = np.random.default_rng(9).integers(1, 5, size=g.ne)
    psi_o = ragged_state(g, dims, seed=12)
    grp = G.edge_color(g)[0]
    arrs = [gates.gate_array(("Rxx", [0, 1], 0.8)) for _ in grp]
    gate_case(E, g, psi_o, grp, arrs, dict(maxdim=8, cutoff=1e-12, normalize_tensors=True))


def test_two_site_gate_tall_qr_tree(E):
    # degree-4 vertices with chi=6: P_o = 216 > rows of one QR block for nc = 12 -> exercises the TSQR tree
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=6, seed=77)
    e = g.edge_of[(g.v((2, 2)), g.v((2, 3)))]
    gate_case(E, g, psi_o, [e], [gates.gate_array(("CNOT", [0, 1]))], dict(maxdim=6, cutoff=1e-10, normalize_tensors=True), sweeps=4)


def test_two_site_gate_3d_chi8_deep_tree(E):
    g = G.named_grid((3, 3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=5)
    c = g.v((2, 2, 2))
    e = g.incident[c][0]
    gate_case(E, g, psi_o, [e], [gates.gate_array(("Rzz", [0, 1], 1.1))], dict(maxdim=3, cutoff=None, normalize_tensors=True), sweeps=3)


def test_circuit_known_answer_two_qubits(E):
    # /root/reference/test/test_apply.jl:11-20
    circuit = [("Rx", [(1, 1)], 0.5), ("Rx", [(2, 1)], 0.2), ("CPHASE", [(1, 1), (2, 1)], -0.3)]
    g = G.build_graph_from_circuit(circuit)
    psi0 = E.tensornetworkstate(lambda v: "↓", g, chi_cap=2)
    psi, errs = E.apply_circuit(circuit, psi0, apply_kwargs=dict(maxdim=2, cutoff=1e-10, normalize_tensors=False))
    assert isinstance(psi, E.TensorNetworkState)
    assert psi.maxvirtualdim() <= 2
    got = O.to_statevector(O.TensorNetworkState(g, psi.tensors()))
    assert np.vdot(got, got).real == pytest.approx(1.0, abs=1e-12)
    assert statevector_fidelity(got, O.statevector_circuit(g, lambda v: "↓", circuit)) == pytest.approx(1.0, abs=1e-12)


def test_ising_layer_grid_matches_oracle(E):
    # /root/reference/test/test_apply.jl:23-53 + refresh schedule
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=1, seed=123)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    layer = ising_layer(g, groups=G.grid_edge_color(g))
    kw = dict(cutoff=1e-10, normalize_tensors=False)
    io, ie = {}, {}
    bo = O.update(O.BeliefPropagationCache(psi_o), **O.default_bp_update_kwargs(psi_o))
    bo, errs_o = O.apply_gates(layer, bo, apply_kwargs=kw, info=io)
    be = E.update(to_engine(psi_o, chi_cap=4), **E.default_bp_update_kwargs(to_engine(psi_o)))
    be, errs_e = E.apply_gates(layer, be, apply_kwargs=kw, info=ie)
    assert ie["refreshes"] == io["refreshes"] == 5
    assert be.psi.maxvirtualdim() <= 2
    assert list(be.psi.bond_dims()) == list(bo.network.bond_dims())
    assert np.allclose(errs_e, errs_o, atol=1e-14)
    sv = O.to_statevector(O.TensorNetworkState(g, be.psi.tensors()))
    assert np.vdot(sv, sv).real == pytest.approx(1.0, abs=1e-10)
    for v in range(g.nv):
        assert abs(E.expect(be, ("Z", [g.names[v]])) - O.expect(bo, ("Z", [g.names[v]]))) < EXP_TOL


def test_trotter_evolution_config_c1_small(E):
    # BASELINE config C1 scaled down: 4x4 grid, 6 Trotter layers in ONE apply_gates call, maxdim=4, cutoff=1e-10
    g = G.named_grid((4, 4))
    layer = ising_layer(g, dt=0.01, hx=2.5, hz=0.2, J=1.0, groups=G.grid_edge_color(g))
    circuit = layer * 6
    kw = dict(maxdim=4, cutoff=1e-10, normalize_tensors=True)
    psi_o, _ = O.apply_gates(circuit, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=kw)
    psi_e, _ = E.apply_gates(circuit, E.tensornetworkstate(lambda v: "↑", g, chi_cap=4), apply_kwargs=kw)
    assert list(psi_e.bond_dims()) == list(psi_o.bond_dims())
    want = O.expect_tns(psi_o, ("Z", [(2, 2)]))
    got = E.expect(psi_e, ("Z", [(2, 2)]), alg="bp")
    assert abs(got - want) < EXP_TOL
    zz_w = O.expect_tns(psi_o, ("ZZ", [(2, 2), (2, 3)]))
    zz_g = E.expect(psi_e, ("ZZ", [(2, 2), (2, 3)]))
    assert abs(zz_g - zz_w) < EXP_TOL


def test_tree_circuit_matches_statevector(E):
    g = G.named_comb_tree((3, 2))
    rs = np.random.default_rng(0)
    layer = []
    for _ in range(3):
        layer += [("Rx", [g.names[v]], float(rs.uniform(0, 2))) for v in range(g.nv)]
        layer += [("Ry", [g.names[v]], float(rs.uniform(0, 2))) for v in range(g.nv)]
        for grp in G.edge_color(g):
            layer += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 0.7) for e in grp]
    psi, _ = E.apply_gates(layer, E.tensornetworkstate(lambda v: "↑", g, chi_cap=8), apply_kwargs=dict(cutoff=1e-14, normalize_tensors=True))
    sv = O.statevector_circuit(g, lambda v: "↑", layer)
    got = O.to_statevector(O.TensorNetworkState(g, psi.tensors()))
    assert statevector_fidelity(got, sv) == pytest.approx(1.0, abs=1e-11)


def test_symmetric_gauge(E):
    g = G.named_grid((3, 2))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=7)
    bo = O.rescale(O.update(O.BeliefPropagationCache(psi_o), maxiter=30))
    be = E.rescale(E.update(to_engine(psi_o), maxiter=30))
    so, se = O.symmetric_gauge(bo), E.symmetric_gauge(be)
    check_messages(E, se, so, tol=1e-9)
    a = O.to_statevector(psi_o)
    b = O.to_statevector(O.TensorNetworkState(g, se.psi.tensors()))
    assert statevector_fidelity(a, b) == pytest.approx(1.0, abs=1e-10)
    # the new diag(S) messages are a BP fixed point of the gauged state (up to BP convergence)
    for k in range(2 * g.ne):
        m = E.message(se, k)
        new = O.updated_message(O.BeliefPropagationCache(O.TensorNetworkState(g, se.psi.tensors()),
                                                         {j: E.message(se, j) for j in range(2 * g.ne)}), k)
        assert rel(new, m / m.sum()) < 1e-6


def test_ghz_entanglement_log2(E):
    # /root/reference/test/test_constructors.jl:69-73
    g = G.named_hexagonal_lattice_graph(3, 3)
    ts = []
    for v in range(g.nv):
        z = g.degree(v)
        t = np.zeros((2,) + (2,) * z, dtype=np.complex128)
        t[(0,) * (z + 1)] = 1
        t[(1,) * (z + 1)] = 1
        ts.append(t)
    psi = E.TensorNetworkState.from_arrays(g, ts)
    assert E.entanglement(psi, 0) == pytest.approx(math.log(2), abs=1e-12)


def test_truncate_bp(E):
    # /root/reference/test/test_truncate.jl:16-33
    g = G.named_hexagonal_lattice_graph(2, 2)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=123)
    groups = G.edge_color(g)
    bo = O.update(O.BeliefPropagationCache(psi_o))
    bo = O.truncate_bp(bo, maxdim=2, cutoff=1e-10, normalize_tensors=False, edge_groups=groups)
    pe = E.truncate(E.TensorNetworkState.from_arrays(g, psi_o.tensors), maxdim=2, cutoff=1e-10, normalize_tensors=False, edge_groups=groups)
    assert pe.maxvirtualdim() <= 2
    a = O.to_statevector(bo.network)
    b = O.to_statevector(O.TensorNetworkState(g, pe.tensors()))
    assert statevector_fidelity(a, b) == pytest.approx(1.0, abs=1e-8)


def test_error_codes(E):
    from tensornetworkquantumsimulator_b200.lib import TnqsError
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=1)
    be = to_engine(psi_o)
    e1, e2 = g.incident[g.v((2, 2))][:2]
    cn = gates.gate_array(("CNOT", [0, 1]))
    with pytest.raises(TnqsError) as ei:
        E.apply_2site_batch(be, [e1, e2], [cn, cn])
    assert ei.value.status == -3                     # TNQS_ERR_OVERLAPPING_BATCH
    with pytest.raises(TnqsError) as ei:
        be.psi.set_site(0, np.zeros((2, 3, 3)))
    assert ei.value.status == -2                     # TNQS_ERR_SHAPE
    # a message with a negative eigenvalue beyond the cutoff: the reference throws, the engine returns an error
    k = g.dedge(g.v((1, 2)), g.v((2, 2)))
    E.setmessage(be, k, np.diag([1.0, -0.5]).astype(np.complex128))
    e = g.edge_of[(g.v((2, 2)), g.v((2, 3)))]
    with pytest.raises(TnqsError) as ei:
        E.apply_2site_batch(be, [e], [cn])
    assert ei.value.status == -4                     # TNQS_ERR_NEGATIVE_EIGENVALUE
    with pytest.raises(O.NegativeEigenvalueError):
        bo = O.BeliefPropagationCache(psi_o.copy(), {k: np.diag([1.0, -0.5]).astype(np.complex128)})
        O.apply_gate(bo, cn, [g.v((2, 2)), g.v((2, 3))])


# ---- vertex partition on ONE GPU: two handles play two ranks, buffers are swapped by pointer ---------------------------------
def _two_rank_setup(E, g, psi_o, chi_cap):
    from tensornetworkquantumsimulator_b200 import parallel as P
    owner = P.strip_owner(g, 2, axis=1)
    lgs = [P.build_local_graph(g, owner, r) for r in range(2)]
    caches = []
    for r, lg in enumerate(lgs):
        ts = []
        for lv, gv in enumerate(lg.l2g_v):
            if lg.owned[lv]:
                ts.append(psi_o.tensors[int(gv)])
            else:                                     # ghost placeholder: only the cut legs exist locally
                dims = [psi_o.bond_dim(int(lg.l2g_e[e])) for e in lg.g.incident[lv]]
                ts.append(np.zeros([2] + dims, dtype=np.complex128))
        psi = E.TensorNetworkState.from_arrays(lg.g, ts, chi_cap=chi_cap)
        bpc = E.BeliefPropagationCache(psi, _share=True)
        from tensornetworkquantumsimulator_b200 import lib
        ow = lib.i32(lg.owner_local)
        lib.check(psi._h.ptr, lib.load().tnqs_set_partition(psi._h.ptr, lib.ptr_i32(ow), r))
        caches.append(bpc)
    return owner, lgs, caches


def test_partitioned_bp_matches_single_handle_jacobi(E):
    import torch
    from tensornetworkquantumsimulator_b200 import parallel as P
    g = G.named_grid((4, 6))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=17)
    owner, lgs, caches = _two_rank_setup(E, g, psi_o, 3)
    ref = E.update(to_engine(psi_o), maxiter=3, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    comps = [P.EngineCompute(c, torch.device("cuda", 0)) for c in caches]
    for sweep in range(3):
        for r in range(2):
            lg = lgs[r]
            seq = np.array([k for k in range(2 * lg.g.ne) if lg.owned[lg.g.dsrc(k)]], dtype=np.int32)
            comps[r].sweep(seq, True)
        bufs = {r: comps[r].pack(lgs[r].send[1 - r]) for r in range(2)}
        for r in range(2):
            comps[r].unpack(lgs[r].recv[1 - r], bufs[1 - r])
    for r in range(2):
        lg = lgs[r]
        for lk in range(2 * lg.g.ne):
            if not lg.owned[lg.g.dsrc(lk)] and not lg.owned[lg.g.ddst(lk)]:
                continue
            kg = 2 * int(lg.l2g_e[lk // 2]) + (lk & 1)
            assert np.array_equal(E.message(caches[r], lk), E.message(ref, kg)), "partitioned JACOBI must be bitwise the single-handle result"


def test_partitioned_gates_match_single_handle(E):
    import torch
    from tensornetworkquantumsimulator_b200 import parallel as P
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=4, seed=23)
    owner, lgs, caches = _two_rank_setup(E, g, psi_o, 8)
    seq = np.arange(2 * g.ne)
    ref = E.update(to_engine(psi_o, chi_cap=8), maxiter=4, edge_sequence=seq, mode="jacobi")
    comps = [P.EngineCompute(c, torch.device("cuda", 0)) for c in caches]
    for sweep in range(4):
        for r in range(2):
            lg = lgs[r]
            comps[r].sweep(np.array([k for k in range(2 * lg.g.ne) if lg.owned[lg.g.dsrc(k)]], dtype=np.int32), False)
        bufs = {r: comps[r].pack(lgs[r].send[1 - r]) for r in range(2)}
        for r in range(2):
            comps[r].unpack(lgs[r].recv[1 - r], bufs[1 - r])
    kw = dict(maxdim=3, cutoff=1e-12, normalize_tensors=True)
    gate = gates.gate_array(("Rxx", [0, 1], 0.7))
    for grp in G.grid_edge_color(g):                      # includes the class whose vertical edges cross the cut
        rr, re_, _ = E.apply_2site_batch(ref, sorted(grp), [gate] * len(grp), apply_kwargs=kw)
        state = []
        for r in range(2):
            es, rm, cut_with = P.class_gates_local(g, lgs[r], owner, r, grp)
            state.append((es, rm, E.apply_2site_begin(caches[r], es, [gate] * len(es), remote=rm, apply_kwargs=kw)))
        for r in range(2):
            es, rm, (ptr, ncut, slot) = state[r]
            other_ptr = state[1 - r][2][0]
            assert ncut == state[1 - r][2][1]
            ranks, errs, _ = E.apply_2site_end(caches[r], len(es), other_ptr if ncut else None)
            glob = [int(lgs[r].l2g_e[e]) for e in es]
            for i, e in enumerate(glob):
                j = sorted(grp).index(e)
                assert ranks[i] == rr[j]
                assert errs[i] == pytest.approx(re_[j], rel=1e-9, abs=1e-18)
    tr = ref.psi.tensors()
    for r in range(2):
        tl = caches[r].psi.tensors()
        for lv, gv in enumerate(lgs[r].l2g_v):
            if lgs[r].owned[lv]:
                assert rel(tl[lv], tr[int(gv)]) < 1e-12, f"site {gv} on rank {r}"


def test_bench_phase_sequence_equals_apply_gates(E):
    # bench.py times one Trotter layer as the explicit run / refresh schedule of engine.circuit_schedule; executing that
    # schedule call by call must be bit-identical to apply_gates (tnqs_apply_circuit with the affected-vertices rule)
    g = G.named_grid((4, 4))
    groups = G.grid_edge_color(g)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=4, seed=9)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    seq = G.forest_cover_edge_sequence(g)
    kw = dict(maxdim=4, cutoff=1e-10, normalize_tensors=True)
    bp_kw = dict(maxiter=25, tolerance=1e-8)
    rx, rz, rzz = gates.gate_array(("Rx", [0], 0.5)), gates.gate_array(("Rz", [0], 0.4)), gates.gate_array(("Rzz", [0, 1], 0.25))
    a = E.update(to_engine(psi_o, chi_cap=4), edge_sequence=seq, **bp_kw)
    b = a.copy()
    layer = [(rx, [v]) for v in range(g.nv)] + [(rz, [v]) for v in range(g.nv)]
    for grp in groups:
        layer += [(rzz, [int(g.esrc[e]), int(g.edst[e])]) for e in grp]
    info = {}
    a, errs = E.apply_gates(layer, a, apply_kwargs=kw, bp_update_kwargs=bp_kw, inplace=True, info=info)
    sched = E.circuit_schedule(g, layer)
    for item in sched:
        if item[0] == "refresh":
            E.update(b, edge_sequence=seq, inplace=True, **bp_kw)
        elif item[0] == "gates1":
            E.apply_1site_batch(b, [layer[i][1][0] for i in item[1]], [E.fold_gates(layer[i][0], [layer[k][0] for k in f]) for i, f in zip(item[1], item[2])])
        else:
            es = [g.edge_of[(layer[i][1][0], layer[i][1][1])] for i in item[1]]
            E.apply_2site_batch(b, es, [layer[i][0] for i in item[1]], flip=[0 if int(g.esrc[e]) == layer[i][1][0] else 1 for e, i in zip(es, item[1])], apply_kwargs=kw)
    assert info["refreshes"] == 5 == sum(1 for it in sched if it[0] == "refresh")
    assert [(len(it[1]), sum(len(f) for f in it[2])) for it in sched if it[0] == "gates1"] == [(g.nv, g.nv)]      # Rz folded onto Rx: one pass
    # runs are defined by the rule, not by colour: the first gates of the third class precede its refresh
    assert [len(it[1]) for it in sched if it[0] == "gates2"] != [len(grp) for grp in groups]
    assert list(a.psi.bond_dims()) == list(b.psi.bond_dims())
    for v in range(g.nv):
        assert np.array_equal(a.psi[v], b.psi[v])
    for k in range(2 * g.ne):
        assert np.array_equal(E.message(a, k), E.message(b, k))


def test_chunked_levels_and_gate_batches(E, monkeypatch):
    # a 1 MB scratch budget forces BP levels into sub-levels (staged messages + commit pass) and gate batches into chunks
    monkeypatch.setenv("TNQS_SCRATCH_MB", "1")
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=71)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    seq = G.forest_cover_edge_sequence(g)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=2, edge_sequence=seq)
    be = E.update(to_engine(psi_o, chi_cap=8), maxiter=2, edge_sequence=seq)
    check_messages(E, be, bo)
    bj = O.BeliefPropagationCache(psi_o)
    bj.messages = {k: O.updated_message(bj, k) for k in range(2 * g.ne)}
    ej = E.update(to_engine(psi_o, chi_cap=8), maxiter=1, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    check_messages(E, ej, bj)
    # generic (leg-by-leg) chains under the same budget: 40 KB of scratch per chain at chi = 5, so the JACOBI level is cut into
    # sub-levels and the cuts fall inside the group of updates of one vertex - those fall back from shared to separate absorptions
    psi5 = O.random_tensornetworkstate(g, bond_dimension=5, seed=72)
    b5 = O.BeliefPropagationCache(psi5)
    for _ in range(2):
        b5.messages = {k: O.updated_message(b5, k) for k in range(2 * g.ne)}
    e5 = E.update(to_engine(psi5, chi_cap=5), maxiter=2, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    check_messages(E, e5, b5)
    grp = G.grid_edge_color(g)[2]
    arr = gates.gate_array(("Rxx", [0, 1], 0.6))
    dets = []
    for e in grp:
        det = {}
        O.apply_gate(bo, arr, [int(g.esrc[e]), int(g.edst[e])], maxdim=8, cutoff=1e-12, detail=det)
        dets.append(det)
    ranks, errs, sv = E.apply_2site_batch(be, grp, [arr] * len(grp), apply_kwargs=dict(maxdim=8, cutoff=1e-12))
    for i, det in enumerate(dets):
        assert ranks[i] == det["rank"]
        assert rel(sv[i, :ranks[i]], det["svals"]) < MSG_TOL
    te = be.psi.tensors()
    for e in grp:
        assert rel(two_site(g, te, e), two_site(g, bo.network.tensors, e)) < 1e-9


def test_config_c3_eagle_kicked_ising_small(E):
    # BASELINE config C3 scaled down: 127-qubit Eagle heavy-hex, kicked Ising (Rx(0.4) on all, Rzz(pi/2) in 3 colour classes),
    # 2 steps, maxdim=4, cutoff=1e-12, <Z> on all sites  (examples/heavyhexIsing_dynamics.jl:15-29)
    g = G.eagle_heavy_hex()
    groups = G.edge_color(g)
    assert len(groups) == 3
    layer = [("Rx", [g.names[v]], 0.4) for v in range(g.nv)]
    for grp in groups:
        layer += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], math.pi / 2) for e in grp]
    kw = dict(maxdim=4, cutoff=1e-12, normalize_tensors=True)
    psi_o, errs_o = O.apply_gates(layer * 2, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=kw)
    psi_e, errs_e = E.apply_gates(layer * 2, E.tensornetworkstate(lambda v: "↑", g, chi_cap=4), apply_kwargs=kw)
    assert list(psi_e.bond_dims()) == list(psi_o.bond_dims())
    assert np.allclose(errs_e, errs_o, atol=1e-12)
    obs = [("Z", [g.names[v]]) for v in range(g.nv)]
    want = np.array(O.expect_tns(psi_o, obs))
    got = np.array(E.expect(psi_e, obs))
    assert np.max(np.abs(got - want)) < EXP_TOL


def test_config_c4_cubic_small(E):
    # BASELINE config C4 scaled down: 3x3x3 open cubic lattice (coordination up to 6), [Rz, Rxx in 6 colour classes, Rz],
    # maxdim=3, BP run to tolerance 1e-12  (examples/3dIsing_dynamics.jl:22-28)
    g = G.named_grid((3, 3, 3))
    groups = G.grid_edge_color(g)
    assert len(groups) == 6
    layer = [("Rz", [g.names[v]], 0.3) for v in range(g.nv)]
    for grp in groups:
        layer += [("Rxx", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 0.4) for e in grp]
    layer += [("Rz", [g.names[v]], 0.3) for v in range(g.nv)]
    kw = dict(maxdim=3, cutoff=1e-10, normalize_tensors=True)
    bp_kw = dict(maxiter=60, tolerance=1e-12)
    psi_o, _ = O.apply_gates(layer * 2, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=kw, bp_update_kwargs=bp_kw)
    psi_e, _ = E.apply_gates(layer * 2, E.tensornetworkstate(lambda v: "↑", g, chi_cap=3), apply_kwargs=kw, bp_update_kwargs=bp_kw)
    assert list(psi_e.bond_dims()) == list(psi_o.bond_dims())
    c = (2, 2, 2)
    want = O.expect_tns(psi_o, ("Z", [c]), cache_update_kwargs=bp_kw)
    got = E.expect(psi_e, ("Z", [c]), cache_update_kwargs=bp_kw)
    assert abs(got - want) < EXP_TOL


@pytest.mark.parametrize("d", [3, 4])
def test_site_dimension_3_and_4(E, d):
    # spin-1 (d=3) and Pauli-basis (d=4) sites (tensornetworkstate_constructors.jl:19-35): generic kernels, d^2 x d^2 gates
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=29, d=d)
    seq = G.forest_cover_edge_sequence(g)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=4, edge_sequence=seq)
    be = E.update(to_engine(psi_o, chi_cap=9), maxiter=4, edge_sequence=seq)
    check_messages(E, be, bo)
    rho = E.rdm_1site(be, list(range(g.nv)))
    for v in range(g.nv):
        want = O.rdm_1site(bo, v)
        got = rho[v][:d, :d] / np.trace(rho[v][:d, :d])
        assert np.max(np.abs(got - want)) < EXP_TOL
    rs = np.random.default_rng(6)
    grp = G.grid_edge_color(g)[0]
    arrs = []
    for e in grp:
        h = rs.normal(size=(d * d, d * d)) + 1j * rs.normal(size=(d * d, d * d))
        u, _ = np.linalg.qr(h)
        arrs.append(u.reshape(d, d, d, d))
    dets = []
    for e, a in zip(grp, arrs):
        det = {}
        O.apply_gate(bo, a, [int(g.esrc[e]), int(g.edst[e])], maxdim=6, cutoff=1e-12, detail=det)
        dets.append(det)
    ranks, errs, sv = E.apply_2site_batch(be, grp, arrs, apply_kwargs=dict(maxdim=6, cutoff=1e-12))
    for i, det in enumerate(dets):
        assert ranks[i] == det["rank"]
        assert rel(sv[i, :ranks[i]], det["svals"]) < MSG_TOL
    te = be.psi.tensors()
    for e in grp:
        assert rel(two_site(g, te, e), two_site(g, bo.network.tensors, e)) < 1e-9


def test_empty_and_degenerate_inputs(E):
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=1)
    be = to_engine(psi_o)
    before = be.psi.tensors()
    out, errs = E.apply_gates([], be, apply_kwargs=dict(maxdim=2), update_cache=False)        # empty circuit: a copy
    assert len(errs) == 0
    for a, b in zip(before, out.psi.tensors()):
        assert np.array_equal(a, b)
    assert E.expect(be, []) == []                                                              # no observables
    assert E.expect(be, ("Z", [(2, 2)], 0)) == 0                                               # zero coefficient (expect.jl:63)
    # a single isolated vertex: no edges, no messages, expectation from the site alone
    g1 = G.Graph([(1,)], [], [])
    psi1 = E.TensorNetworkState.from_arrays(g1, [np.array([0.6, 0.8j])])
    b1 = E.update(E.BeliefPropagationCache(psi1), edge_sequence=[])
    assert E.expect(b1, ("Z", [(1,)])) == pytest.approx(0.36 - 0.64, abs=1e-14)
    # tree: default maxiter is 1 and BP is exact (beliefpropagationcache.jl:31)
    gt = G.path_graph(5)
    psit = O.random_tensornetworkstate(gt, bond_dimension=3, seed=3)
    info = {}
    bt = E.update(to_engine(psit), info=info)
    assert info["iters"] == 1
    sv_ = O.to_statevector(psit)
    assert abs(E.expect(bt, ("Z", [(3,)])) - O.statevector_expect(sv_, 5, {2: gates.Z})) < 1e-12


def test_bp_sampling_matches_oracle_draw_for_draw(E):
    # SURVEY.md 8f item 2 (src/sampling.jl:3-46): same uniform draws -> same bit strings; product state -> its own configuration
    g = G.named_comb_tree((3, 2))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=13)
    want, _ = O.sample(psi_o, 4, rng=np.random.default_rng(7))
    got, _ = E.sample(E.TensorNetworkState.from_arrays(g, psi_o.tensors), 4, rng=np.random.default_rng(7))
    assert got == want
    g2 = G.named_grid((3, 3))
    bits, _ = E.sample(E.tensornetworkstate(lambda v: "↓", g2), 2, bp_update_kwargs=dict(maxiter=3), rng=np.random.default_rng(0))
    assert all(b == 1 for s in bits for b in s.values())


def test_two_site_expectation_along_paths(E):
    # src/expect.jl:58-83 with the Steiner tree (= shortest path) of two non-adjacent vertices
    g = G.named_grid((4, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=37)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=12)
    be = E.update(to_engine(psi_o), maxiter=12)
    obs = [("ZZ", [(1, 1), (3, 2)]), ("XY", [(4, 3), (1, 1)], 2.0), ("YZ", [(2, 2), (2, 3)]), ("ZX", [(1, 3), (4, 1)])]
    got, want = E.expect(be, obs), [O.expect(bo, o) for o in obs]
    assert np.max(np.abs(np.array(got) - np.array(want))) < EXP_TOL
    # exact on a tree
    gt = G.named_comb_tree((3, 3))
    psit = O.random_tensornetworkstate(gt, bond_dimension=2, seed=5)
    bt = E.update(to_engine(psit))
    sv = O.to_statevector(psit)
    a, b = gt.v((1, 3)), gt.v((3, 3))
    ex = O.statevector_expect(sv, gt.nv, {a: gates.X, b: gates.Z})
    assert abs(E.expect(bt, ("XZ", [(1, 3), (3, 3)])) - ex) < 1e-12


def test_full_size_c2_properties(E):
    # BASELINE config C2 at full size (32x32, chi=16, 1.77 GiB): size-independent properties instead of an oracle run
    import bench
    g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
    psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    info = {}
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=40, tolerance=1e-10, info=info)
    assert info["iters"] < 40 and info["mean_diff"] <= 1e-10
    # idempotence: one more sweep on converged messages changes nothing beyond the tolerance
    m0 = E.message(bpc, 1234)
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=1, tolerance=0.0, info=info)
    assert info["mean_diff"] < 1e-9 and rel(E.message(bpc, 1234), m0) < 1e-4
    # messages are Hermitian, positive and normalised to unit entry sum
    w = np.linalg.eigvalsh(m0)
    assert np.allclose(m0, m0.conj().T, atol=1e-12) and w.min() > -1e-12 and abs(m0.sum() - 1) < 1e-12
    # rescale: every vertex and edge scalar becomes 1, the BP partition function 1
    E.rescale(bpc, inplace=True)
    assert np.allclose(E.vertex_scalars(bpc), 1.0, atol=1e-9) and np.allclose(E.edge_scalars(bpc), 1.0, atol=1e-9)
    # the sequential order and the JACOBI schedule reach the same fixed point
    z_ref = E.expect(bpc, ("Z", [(16, 16)]))
    bj = bpc.copy()
    E.update(bj, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=60, tolerance=1e-12, mode="jacobi")
    assert abs(E.expect(bj, ("Z", [(16, 16)])) - z_ref) < 1e-6
    # one colour class of Rzz with maxdim = chi: ranks <= chi, 0 <= truncation error < 1, unit-norm singular values that
    # become the bond messages; a second identical run is bit-identical (no atomics, fixed reduction orders)
    kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
    b2 = bpc.copy()
    r1, e1, s1 = E.apply_2site_batch(bpc, groups[0], [rzz] * len(groups[0]), apply_kwargs=kw)
    r2, e2, s2 = E.apply_2site_batch(b2, groups[0], [rzz] * len(groups[0]), apply_kwargs=kw)
    assert np.array_equal(r1, r2) and np.array_equal(e1, e2) and np.array_equal(s1, s2)
    assert r1.max() <= 16 and r1.min() >= 1 and e1.min() >= 0 and e1.max() < 1
    assert np.allclose(np.sum(s1 ** 2, axis=1), 1.0, atol=1e-12)
    e = groups[0][17]
    assert np.allclose(np.diag(E.message(bpc, 2 * e)).real, s1[17, :r1[17]], atol=1e-14)
    for v in (int(g.esrc[e]), int(g.edst[e])):
        assert abs(np.linalg.norm(bpc.psi[v]) - 1.0) < 1e-12
    # unitary one-site gates preserve <Z>-independent quantities: the BP norm stays 1 after Rx on every site + refresh
    E.apply_1site_batch(bpc, list(range(g.nv)), [rx] * g.nv)
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=25, tolerance=1e-8)
    zs = E.expect(bpc, [("Z", [g.names[v]]) for v in (0, 500, 1023)])
    assert all(abs(z.imag) < 1e-10 and abs(z.real) <= 1 + 1e-10 for z in zs)


def test_full_size_c4_properties(E):
    # BASELINE config C4 at full size (examples/3dIsing_dynamics.jl:8-28 scaled to 8x8x8, open boundaries: 512 sites of coordination
    # up to 6, 1344 edges in 6 colour classes, chi = 8: 8 MiB per interior site, 3.2 GiB in all), BP run to tolerance 1e-12:
    # size-independent properties instead of an oracle run (one CPU sweep of this state is 0.5 TFLOP)
    g = G.named_grid((8, 8, 8))
    groups = G.grid_edge_color(g)
    assert g.nv == 512 and g.ne == 1344 and len(groups) == 6 and max(g.degree(v) for v in range(g.nv)) == 6
    psi = E.random_tensornetworkstate(g, bond_dimension=8, seed=2)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    seq = G.forest_cover_edge_sequence(g)
    info = {}
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=100, tolerance=1e-12, info=info)
    assert info["iters"] < 100 and info["mean_diff"] <= 1e-12
    k = 2 * g.edge_of[(g.v((4, 4, 4)), g.v((4, 4, 5)))]
    m0 = E.message(bpc, k)
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=1, tolerance=0.0, info=info)            # idempotence at the fixed point
    assert info["mean_diff"] < 1e-11 and rel(E.message(bpc, k), m0) < 1e-5
    w = np.linalg.eigvalsh(m0)
    assert np.allclose(m0, m0.conj().T, atol=1e-12) and w.min() > -1e-12 and abs(m0.sum() - 1) < 1e-12
    E.rescale(bpc, inplace=True)
    assert np.allclose(E.vertex_scalars(bpc), 1.0, atol=1e-9) and np.allclose(E.edge_scalars(bpc), 1.0, atol=1e-9)
    c = (4, 4, 4)
    z_ref = E.expect(bpc, ("Z", [c]))
    bj = bpc.copy()
    E.update(bj, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=200, tolerance=1e-14, mode="jacobi")
    assert abs(E.expect(bj, ("Z", [c])) - z_ref) < 1e-6                                             # same fixed point in both orders
    # linearity of the vertex scalar in |T_v|^2 (a scalar multiple of one site tensor)
    vs0 = E.vertex_scalars(bpc)
    v = g.v(c)
    t = bpc.psi[v]
    bpc.psi.set_site(v, 3.0 * t)
    assert abs(E.vertex_scalars(bpc)[v] / vs0[v] - 9.0) < 1e-10
    bpc.psi.set_site(v, t)
    # one colour class of Rxx with maxdim = chi (QR of 32768 x 16 matrices, 32 x 32 theta'): reproducible bit for bit, unit-norm
    # singular values that become the bond messages, unit-norm site tensors
    rxx = gates.gate_array(("Rxx", [0, 1], 0.4))
    kw = dict(maxdim=8, cutoff=1e-10, normalize_tensors=True)
    b2 = bpc.copy()
    r1, e1, s1 = E.apply_2site_batch(bpc, groups[0], [rxx] * len(groups[0]), apply_kwargs=kw)
    r2, e2, s2 = E.apply_2site_batch(b2, groups[0], [rxx] * len(groups[0]), apply_kwargs=kw)
    assert np.array_equal(r1, r2) and np.array_equal(e1, e2) and np.array_equal(s1, s2)
    assert r1.max() <= 8 and r1.min() >= 1 and e1.min() >= 0 and e1.max() < 1
    assert np.allclose(np.sum(s1 ** 2, axis=1), 1.0, atol=1e-12) and np.all(np.diff(s1, axis=1) <= 1e-15)
    e = groups[0][5]
    assert np.allclose(np.diag(E.message(bpc, 2 * e)).real, s1[5, :r1[5]], atol=1e-14)
    for u in (int(g.esrc[e]), int(g.edst[e])):
        assert abs(np.linalg.norm(bpc.psi[u]) - 1.0) < 1e-12


def test_c5_strip_chi32_properties(E):
    # BASELINE config C5 is a 128x128 grid at chi = 32 (497 GiB, >= 4 GPUs); one 16x16 strip of it (7.3 GiB) exercises the same
    # kernels on one GPU: chi = 32 message updates (generic tensor-core leg / closing kernels, chunked levels), JACOBI sweeps as the
    # partitioned runs use them, 128 x 128 theta' with the Jacobi iteration in global memory
    g = G.named_grid((16, 16))
    groups = G.grid_edge_color(g)
    psi = E.random_tensornetworkstate(g, bond_dimension=32, seed=3)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    info = {}
    E.update(bpc, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=60, tolerance=1e-10, mode="jacobi", info=info)
    assert info["iters"] < 60 and info["mean_diff"] <= 1e-10
    E.rescale(bpc, inplace=True)
    assert np.allclose(E.vertex_scalars(bpc), 1.0, atol=1e-9) and np.allclose(E.edge_scalars(bpc), 1.0, atol=1e-9)
    z_j = E.expect(bpc, ("Z", [(8, 8)]))
    br = bpc.copy()
    E.update(br, inplace=True, maxiter=25, tolerance=1e-12)                                          # sequential order from the Jacobi fixed point
    assert abs(E.expect(br, ("Z", [(8, 8)])) - z_j) < 1e-6
    rzz = gates.gate_array(("Rzz", [0, 1], 0.25))
    kw = dict(maxdim=32, cutoff=1e-10, normalize_tensors=True)
    grp = groups[1][:24]
    r1, e1, s1 = E.apply_2site_batch(bpc, grp, [rzz] * len(grp), apply_kwargs=kw)
    assert r1.max() <= 32 and r1.min() >= 1 and e1.min() >= 0 and e1.max() < 1
    assert np.allclose(np.sum(s1 ** 2, axis=1), 1.0, atol=1e-12) and np.all(np.diff(s1, axis=1) <= 1e-15)
    e = grp[3]
    assert np.allclose(np.diag(E.message(bpc, 2 * e)).real, s1[3, :r1[3]], atol=1e-14)
    for u in (int(g.esrc[e]), int(g.edst[e])):
        assert abs(np.linalg.norm(bpc.psi[u]) - 1.0) < 1e-12


def _oracle_jacobi(psi_o, sweeps):
    bo = O.BeliefPropagationCache(psi_o)
    for _ in range(sweeps):                 # Jacobi: every message from the previous sweep's values
        bo.messages = {k: O.updated_message(bo, k) for k in range(2 * bo.g.ne)}
    return bo


@pytest.mark.parametrize("shape,chi", [((6, 6), 8), ((5, 4), 16)])
def test_bp_whole_sweep_kernel_matches_level_launches_and_oracle(E, monkeypatch, shape, chi):
    # REFERENCE_ORDER sweeps as ONE persistent data-flow launch (k_bp_sweep, the default on uniform chi in {8, 16} grids)
    # against the level-by-level launches (TNQS_SWEEP=0) and the oracle; the ring of 4 intermediates forces slot reuse
    g = G.named_grid(shape)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=5)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    seq = G.forest_cover_edge_sequence(g)
    io = {}
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=3, tolerance=0.0, edge_sequence=seq, info=io)
    res = {}
    for name, env in (("level", {"TNQS_SWEEP": "0"}), ("sweep", {"TNQS_SWEEP": "1"}), ("sweep_ring4", {"TNQS_SWEEP": "1", "TNQS_SWEEP_RING": "4"})):
        for k, val in env.items():
            monkeypatch.setenv(k, val)
        ie = {}
        be = E.update(to_engine(psi_o, chi_cap=chi), maxiter=3, tolerance=0.0, edge_sequence=seq, info=ie)
        launches = be.psi.kernel_launches()
        for k in env:
            monkeypatch.delenv(k)
        check_messages(E, be, bo)
        assert ie["iters"] == io["iters"] == 3 and abs(ie["mean_diff"] - io["mean_diff"]) < 1e-10 * max(1.0, abs(io["mean_diff"]))
        res[name] = ([E.message(be, k) for k in range(2 * g.ne)], launches)
    assert res["sweep"][1] < res["level"][1]          # one launch per sweep instead of two per level
    for k in range(2 * g.ne):                          # the fast-path arithmetic is the same; degree-2 vertices differ in rounding only
        assert rel(res["sweep"][0][k], res["level"][0][k]) < 1e-13
        assert np.array_equal(res["sweep"][0][k], res["sweep_ring4"][0][k])     # queue shape never changes a result


@pytest.mark.parametrize("env", [{}, {"TNQS_SWEEP_JACOBI": "1"}, {"TNQS_SWEEP_JACOBI": "1", "TNQS_SWEEP_RING": "3"}])
def test_bp_jacobi_two_sweeps_uniform_chi(E, monkeypatch, env):
    # JACOBI on a lattice that mixes tensor-core chains (degree 3, 4) and generic ones (corners): every message of sweep 2
    # must be computed from the values of sweep 1 (a level that mixes both kinds stages its messages)
    for k, val in env.items():
        monkeypatch.setenv(k, val)
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=9)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    for sweeps in (1, 2):
        bo = _oracle_jacobi(psi_o, sweeps)
        be = E.update(to_engine(psi_o, chi_cap=8), maxiter=sweeps, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
        check_messages(E, be, bo)


def test_set_get_sites_batched_round_trip(E):
    g = G.named_grid((3, 3))
    psi_o = ragged_state(g, [1 + (e % 3) for e in range(g.ne)], seed=3)
    psi = to_engine(psi_o).psi
    ts = psi.tensors()                                       # tnqs_get_sites
    for v in range(g.nv):
        assert np.array_equal(ts[v], psi_o.tensors[v]) and np.array_equal(psi[v], ts[v])
    psi.set_sites([4, 0], [2 * ts[4], 3 * ts[0]])            # tnqs_set_sites
    assert np.array_equal(psi[4], 2 * ts[4]) and np.array_equal(psi[0], 3 * ts[0]) and np.array_equal(psi[1], ts[1])
    with pytest.raises(Exception):
        psi.set_sites([1], [ts[4]])                          # wrong element count


@pytest.mark.parametrize("shape,chi,world", [((4, 6), 3, 2), ((4, 6), 8, 2), ((3, 8), 4, 4), ((3, 4, 3), 8, 2)])
def test_partitioned_bp_peer_memory_halo_matches_single_handle(E, shape, chi, world):
    # (the 3 x 4 x 3 case at chi = 8 puts a degree-6 vertex and its degree-5 neighbours on either side of the cut: deep path + halo)
    # `world` "ranks" inside one process (one host thread per handle, all on cuda:0): each handle's BP kernels store their
    # boundary messages straight into the neighbours' receive buffers, every sweep ends with the device-side barrier /
    # all-reduce, and a whole refresh (fixed sweeps, then run to tolerance) is one call per rank - no host-side exchange
    import threading
    import torch
    from tensornetworkquantumsimulator_b200 import parallel as P, lib
    g = G.named_grid(shape)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=17)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    owner = P.strip_owner(g, world, axis=1)
    lgs = [P.build_local_graph(g, owner, r) for r in range(world)]
    caches = []
    for r, lg in enumerate(lgs):
        ts = []
        for lv, gv in enumerate(lg.l2g_v):
            if lg.owned[lv]:
                ts.append(psi_o.tensors[int(gv)])
            else:
                ts.append(np.zeros([2] + [psi_o.bond_dim(int(lg.l2g_e[e])) for e in lg.g.incident[lv]], dtype=np.complex128))
        psi = E.TensorNetworkState.from_arrays(lg.g, ts, chi_cap=chi)
        ow = lib.i32(lg.owner_local)
        lib.check(psi._h.ptr, lib.load().tnqs_set_partition(psi._h.ptr, lib.ptr_i32(ow), r))
        caches.append(E.BeliefPropagationCache(psi, _share=True))
    comps = [P.EngineCompute(c, torch.device("cuda", 0)) for c in caches]
    layouts = [P.halo_layout(lg) for lg in lgs]
    ptrs = [comps[r].halo_alloc(world, layouts[r][0])[1] for r in range(world)]
    for r in range(world):
        for p in range(world):
            if p != r:
                comps[r].halo_peer(p, len(layouts[p][0]), ptr=ptrs[p])
        comps[r].halo_route(*P.halo_routes(lgs[r], r, {p: layouts[p][1] for p in lgs[r].send}))
    seqs = [np.array([k for k in range(2 * lg.g.ne) if lg.owned[lg.g.dsrc(k)]], dtype=np.int32) for lg in lgs]

    def run(maxiter, tolerance):
        out, errs = [None] * world, []
        def work(r):
            try:
                out[r] = comps[r].refresh(seqs[r], maxiter, tolerance)
            except Exception as ex:                              # pragma: no cover
                errs.append(ex)
        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        [t.start() for t in th]; [t.join() for t in th]
        assert not errs, errs
        return out

    def compare(ref, bitwise):
        for r, lg in enumerate(lgs):
            for lk in range(2 * lg.g.ne):
                if not lg.owned[lg.g.dsrc(lk)] and not lg.owned[lg.g.ddst(lk)]:
                    continue
                kg = 2 * int(lg.l2g_e[lk // 2]) + (lk & 1)
                if bitwise:
                    assert np.array_equal(E.message(caches[r], lk), E.message(ref, kg)), "peer-memory halo must reproduce the single-handle JACOBI result bitwise"
                else:
                    assert rel(E.message(caches[r], lk), E.message(ref, kg)) < 1e-12

    run(3, None)
    compare(E.update(to_engine(psi_o, chi_cap=chi), maxiter=3, edge_sequence=np.arange(2 * g.ne), mode="jacobi"), True)
    # convergence test on the device with the global mean: same sweep count on every rank and as the single handle
    res = run(200, 1e-10)
    info = {}
    ref = E.update(to_engine(psi_o, chi_cap=chi), maxiter=203, tolerance=1e-10, edge_sequence=np.arange(2 * g.ne), mode="jacobi", info=info)
    assert len({it for it, _ in res}) == 1 and res[0][0] + 3 == info["iters"]
    assert all(abs(m - info["mean_diff"]) <= 1e-12 * max(1.0, abs(info["mean_diff"])) for _, m in res)
    compare(ref, False)


def test_apply_gates_fetch_returns_final_sites(E):
    # tnqs_apply_circuit_fetch: the site tensors copied out beside the closing refresh are the final ones, sizes reported
    import ctypes as C
    from tensornetworkquantumsimulator_b200 import lib
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=4)
    circuit = [("Rx", [v], 0.3) for v in g.names] + [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 0.4) for e in range(g.ne)]
    kw = dict(maxdim=4, cutoff=1e-12, normalize_tensors=True)
    a = to_engine(psi_o, chi_cap=4)
    b = to_engine(psi_o, chi_cap=4)
    E.apply_gates(circuit, a, apply_kwargs=kw, inplace=True)
    cap = 2 * 4 ** 4
    bufs = [np.zeros(2 * cap, dtype=np.float64) for _ in range(g.nv)]
    ptrs = (C.POINTER(C.c_double) * g.nv)(*[lib.ptr_f64(x) for x in bufs])
    ns = np.full(g.nv, cap, dtype=np.int64)
    E.apply_gates(circuit, b, apply_kwargs=kw, inplace=True, fetch=(lib.i32(np.arange(g.nv)), ptrs, ns))
    ta = a.psi.tensors()
    for v in range(g.nv):
        assert ns[v] == ta[v].size
        assert np.array_equal(bufs[v][:2 * ns[v]].view(np.complex128), ta[v].ravel(order="F"))
    for k in range(2 * g.ne):
        assert np.array_equal(E.message(a, k), E.message(b, k))


@pytest.mark.parametrize("name,chi", [("comb3x3", 3), ("grid4x4", 4), ("grid4x4", 8)])
def test_bilinear_form_bp_inner(E, name, chi):
    # BilinearForm BP (tnqs_set_bra): messages of <phi|psi> after fixed sweeps against the oracle (1e-10), the partition function
    # = inner(psi, phi; alg = "bp") (src/inner.jl:82-86), exact on the tree; chi = 8 runs the tensor-core sweep kernel with a bra
    g = GRAPHS[name]()
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=31)
    phi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=32)
    for s in (psi_o, phi_o):
        for v in range(g.nv):
            s.tensors[v] /= np.linalg.norm(s.tensors[v])
    seq = G.forest_cover_edge_sequence(g)
    psi_e = E.TensorNetworkState.from_arrays(g, psi_o.tensors, chi_cap=chi)
    phi_e = E.TensorNetworkState.from_arrays(g, phi_o.tensors, chi_cap=chi)
    bo = O.update(O.BeliefPropagationCache(psi_o.copy(), {}, phi_o), maxiter=3, edge_sequence=seq)
    be = E.update(E.bilinear_cache(psi_e, phi_e), maxiter=3, edge_sequence=seq, inplace=True)
    check_messages(E, be, bo)
    zo, ze = O.partitionfunction(bo), E.partitionfunction(be)
    assert abs(ze - zo) < 1e-9 * abs(zo)
    bc = be.copy()                                       # a copy stays a bilinear form
    E.update(bc, maxiter=1, edge_sequence=seq, inplace=True)
    bo1 = O.update(bo, maxiter=1, edge_sequence=seq)
    check_messages(E, bc, bo1)
    if name == "comb3x3":
        exact = np.vdot(O.to_statevector(phi_o), O.to_statevector(psi_o))
        assert abs(E.inner(psi_e, phi_e) - exact) < 1e-10 * abs(exact)
    # detaching the bra gives the norm network back
    from tensornetworkquantumsimulator_b200 import engine as _E
    _E._attach_bra(be, None)
    bn = E.update(E.BeliefPropagationCache(be.psi), maxiter=3, edge_sequence=seq, inplace=True)      # (the constructor resets the messages)
    check_messages(E, bn, O.update(O.BeliefPropagationCache(psi_o), maxiter=3, edge_sequence=seq))


@pytest.mark.parametrize("d,shape,chi,rows", [(4, (3, 2), 16, 256), (4, (3, 3), 9, 144), (3, (3, 3), 15, 135)])
def test_two_site_gate_theta_longer_than_128_rows(E, d, shape, chi, rows):
    # Pauli (d = 4) and spin-1 (d = 3) sites with d * chi close to 64: theta' = (q d) x (q d) with q = d chi, i.e. up to 256 rows -
    # the Jacobi SVD must rotate every row (a row loop that stops at 128 returns garbage singular values with status 0)
    g = G.named_grid(shape)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=97, d=d)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    deg = [g.degree(v) for v in range(g.nv)]
    e = max(range(g.ne), key=lambda e: deg[int(g.esrc[e])] + deg[int(g.edst[e])])       # the edge between the best-connected vertices
    q = [min(chi ** (deg[v] - 1), d * chi) for v in (int(g.esrc[e]), int(g.edst[e]))]
    assert max(q) * d == rows
    rs = np.random.default_rng(11)
    h = rs.normal(size=(d * d, d * d)) + 1j * rs.normal(size=(d * d, d * d))
    u, _ = np.linalg.qr(h)
    gate_case(E, g, psi_o, [e], [u.reshape(d, d, d, d)], dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True), sweeps=3)


@pytest.mark.parametrize("mode", ["jacobi", "reference"])
def test_bp_regions_of_bond_dimension_8_and_16_in_one_sweep(E, mode):
    # an inhomogeneous state (cutoff-driven truncation with maxdim = 16): vertices with all bonds 8 and vertices with all bonds 16
    # in the same BP level -> one slab / closing launch pair per bond dimension, vertices in between on the generic kernels
    g = G.named_grid((6, 3))
    dims = np.array([16 if min(g.names[int(a)][0], g.names[int(b)][0]) >= 4 else 8 for a, b in zip(g.esrc, g.edst)])
    psi_o = ragged_state(g, dims, seed=23)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    kinds = {tuple(sorted(set(int(dims[e]) for e in g.incident[v]))) for v in range(g.nv)}
    assert {(8,), (16,), (8, 16)} <= kinds
    if mode == "jacobi":
        bo = O.BeliefPropagationCache(psi_o)
        for _ in range(2):
            bo.messages = {k: O.updated_message(bo, k) for k in range(2 * g.ne)}
        be = E.update(to_engine(psi_o, chi_cap=16), maxiter=2, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    else:
        seq = G.forest_cover_edge_sequence(g)
        bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=2, edge_sequence=seq)
        be = E.update(to_engine(psi_o, chi_cap=16), maxiter=2, edge_sequence=seq)
    check_messages(E, be, bo)


@pytest.mark.parametrize("mode", ["jacobi", "reference"])
def test_bp_high_coordination_vertices_chi8(E, mode):
    # 3x3x3 cubic lattice at chi = 8 (the C4 geometry in small): corners / edges of degree 3 / 4 on the slab kernels, faces and the
    # centre of degree 5 / 6 on the deep path (kernels_deep.cuh: three legs through shared memory, the last one or two fused with the
    # closing).  Two sweeps against the oracle; the same sweeps with TNQS_DEEP=0 (leg-by-leg generic kernels) are covered by
    # test_deep_path_agrees_with_the_generic_kernels.
    g = G.named_grid((3, 3, 3))
    assert sorted(set(g.degree(v) for v in range(g.nv))) == [3, 4, 5, 6]
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=29)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    if mode == "jacobi":
        bo = O.BeliefPropagationCache(psi_o)
        for _ in range(2):
            bo.messages = {k: O.updated_message(bo, k) for k in range(2 * g.ne)}
        be = E.update(to_engine(psi_o, chi_cap=8), maxiter=2, edge_sequence=np.arange(2 * g.ne), mode="jacobi")
    else:
        seq = G.forest_cover_edge_sequence(g)
        bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=2, edge_sequence=seq)
        be = E.update(to_engine(psi_o, chi_cap=8), maxiter=2, edge_sequence=seq)
    check_messages(E, be, bo)
    # <Z> at the centre (degree 6) and on a face (degree 5) after convergence
    seq = G.forest_cover_edge_sequence(g)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=30, tolerance=1e-12, edge_sequence=seq)
    be = E.update(to_engine(psi_o, chi_cap=8), maxiter=30, tolerance=1e-12, edge_sequence=seq)
    for name in [(2, 2, 2), (2, 2, 1)]:
        assert abs(E.expect(be, ("Z", [name])) - O.expect(bo, ("Z", [name]))) < EXP_TOL


def _projected_two_site(g, tensors, e, seed):
    # gauge-invariant probe of the two-site tensor of edge e that never forms it: every other leg of both sites is contracted with a
    # fixed random vector, the shared bond is summed - a d x d matrix
    out = []
    for v in (int(g.esrc[e]), int(g.edst[e])):
        t = tensors[v]
        keep = 1 + g.leg(v, e)
        for ax in range(t.ndim - 1, 0, -1):
            if ax == keep:
                continue
            rs = np.random.default_rng(seed + 17 * ax + 1000 * v)
            vec = rs.normal(size=t.shape[ax]) + 1j * rs.normal(size=t.shape[ax])
            t = np.tensordot(t, vec, axes=([ax], [0]))
        out.append(t)                                   # [s, bond]
    return out[0] @ out[1].T


@pytest.mark.parametrize("kw", [dict(maxdim=8, cutoff=1e-10, normalize_tensors=True), dict(maxdim=6, cutoff=None, normalize_tensors=True)])
def test_two_site_gates_on_high_coordination_vertices(E, kw):
    # 3x3x3 lattice at chi = 8: gates on centre-face (degree 6 / 5), face-edge (5 / 4) and edge-corner (4 / 3) bonds.  Sites with 4 or 5
    # outer legs take their sqrt(M) absorptions through k_slab_batch (two passes, the first one reading the canonical site tensor);
    # the 8 MiB centre tensor makes the full two-site tensor too large to compare, so the gauge-invariant probe above stands in
    g = G.named_grid((3, 3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=37)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    edges = [g.edge_of[(g.v(a), g.v(b))] for a, b in [((2, 2, 2), (2, 2, 3)), ((1, 2, 2), (1, 2, 3)), ((3, 3, 1), (3, 3, 2))]]
    degs = [sorted((g.degree(int(g.esrc[e])), g.degree(int(g.edst[e])))) for e in edges]
    assert degs == [[5, 6], [4, 5], [3, 4]]
    rs = np.random.default_rng(5)
    arrs = []
    for _ in edges:
        u, _r = np.linalg.qr(rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4)))
        arrs.append(u.reshape(2, 2, 2, 2))
    seq = G.forest_cover_edge_sequence(g)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=4, edge_sequence=seq)
    be = E.update(to_engine(psi_o, chi_cap=16), maxiter=4, edge_sequence=seq)
    dets = []
    for i, e in enumerate(edges):
        det = {}
        O.apply_gate(bo, arrs[i], [int(g.esrc[e]), int(g.edst[e])], detail=det, **kw)
        dets.append(det)
    ranks, errs, sv = E.apply_2site_batch(be, edges, arrs, apply_kwargs=kw)
    te = be.psi.tensors()
    for i, e in enumerate(edges):
        assert ranks[i] == dets[i]["rank"]
        assert rel(sv[i, :ranks[i]], dets[i]["svals"]) < MSG_TOL
        assert errs[i] == pytest.approx(dets[i]["truncerr"], rel=1e-8, abs=1e-16)
        for k in (2 * e, 2 * e + 1):
            assert rel(E.message(be, k), O.message(bo, k)) < MSG_TOL
        assert rel(_projected_two_site(g, te, e, 3), _projected_two_site(g, bo.network.tensors, e, 3)) < 1e-9
    assert list(be.psi.bond_dims()) == list(bo.network.bond_dims())


def test_deep_path_with_a_bra_layer(E):
    # <phi|psi> on the 3x3x3 lattice: the closing of the deep path contracts against the bra's site tensors
    g = G.named_grid((3, 3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=31)
    phi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=32)
    seq = G.forest_cover_edge_sequence(g)
    psi = E.TensorNetworkState.from_arrays(g, psi_o.tensors, chi_cap=8)
    phi = E.TensorNetworkState.from_arrays(g, phi_o.tensors, chi_cap=8)
    bn = E.update(E.bilinear_cache(psi, phi), maxiter=2, edge_sequence=seq, inplace=True)
    bo = O.update(O.BeliefPropagationCache(psi_o.copy(), {}, phi_o), maxiter=2, edge_sequence=seq)
    check_messages(E, bn, bo)
    # JACOBI order: the six updates of the centre vertex share their three-leg products of the KET, each closes against the bra
    bj = O.BeliefPropagationCache(psi_o.copy(), {}, phi_o)
    for _ in range(2):
        bj.messages = {k: O.updated_message(bj, k) for k in range(2 * g.ne)}
    ej = E.update(E.bilinear_cache(psi, phi), maxiter=2, edge_sequence=np.arange(2 * g.ne), mode="jacobi", inplace=True)
    check_messages(E, ej, bj)


def test_bilinear_form_outlives_its_cache_and_checks_shapes(E):
    # the bra pointer lives in the ket's device handle: dropping the cache object (or the bra) must not leave it dangling, and a
    # gate that changes the ket's bond dimensions while a bra is attached must be refused by the next BP update
    import gc
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=41)
    phi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=42)
    seq = G.forest_cover_edge_sequence(g)
    psi = E.TensorNetworkState.from_arrays(g, psi_o.tensors, chi_cap=4)
    phi = E.TensorNetworkState.from_arrays(g, phi_o.tensors, chi_cap=4)
    ket = E.network(E.bilinear_cache(psi, phi))           # the cache is garbage from here on
    del phi
    gc.collect()
    bn = E.update(E.BeliefPropagationCache(ket, _share=True), maxiter=3, edge_sequence=seq, inplace=True)
    bo = O.update(O.BeliefPropagationCache(psi_o.copy(), {}, phi_o), maxiter=3, edge_sequence=seq)
    check_messages(E, bn, bo)
    # grow one bond of the ket: 2 -> 4 under a generic two-site gate; BP must now fail loudly instead of reading the bra out of bounds
    rs = np.random.default_rng(2)
    u, _ = np.linalg.qr(rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4)))
    bn = E.BeliefPropagationCache(ket, _share=True)        # identity messages again (those of a bilinear form are not positive)
    ranks, _, _ = E.apply_2site_batch(bn, [0], [u.reshape(2, 2, 2, 2)], apply_kwargs=dict(maxdim=4, cutoff=None))
    assert ranks[0] == 4
    with pytest.raises(E.TnqsError) as ei:
        E.update(bn, maxiter=1, edge_sequence=seq, inplace=True)
    assert ei.value.status == -2                           # TNQS_ERR_SHAPE
    # destroying the bra detaches it: the ket is a plain state again
    ket._bra = None
    gc.collect()
    plain = E.update(E.BeliefPropagationCache(ket, _share=True), maxiter=1, edge_sequence=seq, inplace=True)
    assert np.isfinite(E.partitionfunction(plain))


def test_config_c1_full_size(E):
    # BASELINE config C1 as the README runs it (README.md:40-66): 5x5 grid, J = 1, hx = 2.5, hz = 0.2, dt = 0.01, 50 Trotter layers in
    # ONE apply_gates call (one lazy-refresh schedule over all 5750 gates), maxdim = 10, cutoff = 1e-10, then <Z> at (2, 2) by BP
    g = G.named_grid((5, 5))
    layer = ising_layer(g, dt=0.01, hx=2.5, hz=0.2, J=1.0, groups=G.grid_edge_color(g))
    circuit = layer * 50
    kw = dict(maxdim=10, cutoff=1e-10, normalize_tensors=True)
    io, ie = {}, {}
    psi_o, errs_o = O.apply_gates(circuit, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=kw, info=io)
    psi_e, errs_e = E.apply_gates(circuit, E.tensornetworkstate(lambda v: "↑", g, chi_cap=10), apply_kwargs=kw, info=ie)
    assert ie["refreshes"] == io["refreshes"]
    assert list(psi_e.bond_dims()) == list(psi_o.bond_dims())                      # truncation ranks exact
    assert np.allclose(errs_e, errs_o, rtol=1e-6, atol=1e-13)
    for obs in [("Z", [(2, 2)]), ("Z", [(1, 1)]), ("ZZ", [(2, 2), (2, 3)]), ("X", [(3, 3)])]:
        assert abs(E.expect(psi_e, obs, alg="bp") - O.expect_tns(psi_o, obs)) < EXP_TOL


def test_config_c3_full_size_eagle(E):
    # BASELINE config C3 at full size (examples/heavyhexIsing_dynamics.jl:15-29): 127-qubit Eagle heavy-hex, kicked Ising, 20 steps of
    # [Rx(0.4) on all, Rzz(pi/2) in 3 colour classes], maxdim = 32, cutoff = 1e-12, <Z> on all 127 sites after every 5th step.
    # Bonds grow 1 -> 32 (ragged dimensions, Householder / CholeskyQR2 mix, theta' up to 128 x 128) and saturate around step 7.
    g = G.eagle_heavy_hex()
    groups = G.edge_color(g)
    layer = [("Rx", [g.names[v]], 0.4) for v in range(g.nv)]
    for grp in groups:
        layer += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], math.pi / 2) for e in grp]
    kw = dict(maxdim=32, cutoff=1e-12, normalize_tensors=True)
    psi_o = O.tensornetworkstate(lambda v: "↑", g)
    psi_e = E.tensornetworkstate(lambda v: "↑", g, chi_cap=32)
    obs = [("Z", [g.names[v]]) for v in range(g.nv)]
    for step in range(20):
        psi_o, errs_o = O.apply_gates(layer, psi_o, apply_kwargs=kw)
        psi_e, errs_e = E.apply_gates(layer, psi_e, apply_kwargs=kw)
        assert list(psi_e.bond_dims()) == list(psi_o.bond_dims()), f"bond dimensions after step {step + 1}"
        assert np.allclose(errs_e, errs_o, rtol=1e-6, atol=1e-12), f"truncation errors of step {step + 1}"
        if (step + 1) % 5 == 0:
            want = np.array(O.expect_tns(psi_o, obs))
            got = np.array(E.expect(psi_e, obs))
            assert np.max(np.abs(got - want)) < EXP_TOL, f"<Z> after step {step + 1}"
    assert psi_e.maxvirtualdim() == 32


@pytest.mark.parametrize("cutoff", [0.2, 1e-10])
def test_optimistic_gate_batches_fall_back_when_ranks_drop(E, cutoff):
    # saturated bonds (chi = maxdim) run optimistically: the device commits a gate only if it kept its bond dimension.  A coarse
    # cutoff makes every gate of the class truncate below chi, so all of them take the conventional second pass; the fine cutoff keeps
    # the ranks at chi (all committed by the first pass).  Either way the result is the oracle's.
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=67)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    grp = G.grid_edge_color(g)[0]
    arrs = [gates.gate_array(("Rzz", [0, 1], 0.3 + 0.05 * i)) for i in range(len(grp))]
    be, bo = gate_case(E, g, psi_o, grp, arrs, dict(maxdim=8, cutoff=cutoff, normalize_tensors=True), sweeps=3)
    dims = be.psi.bond_dims()
    assert (min(dims[e] for e in grp) < 8) == (cutoff == 0.2)
    # a second class on the now ragged state (mixed optimistic / conventional chunks are decided per batch)
    grp2 = G.grid_edge_color(g)[2]
    bo2 = O.update(bo, maxiter=3)
    be2 = E.update(be, maxiter=3)
    dets = []
    for e in grp2:
        det = {}
        O.apply_gate(bo2, arrs[0], [int(g.esrc[e]), int(g.edst[e])], maxdim=8, cutoff=1e-10, detail=det)
        dets.append(det)
    ranks, errs, sv = E.apply_2site_batch(be2, grp2, [arrs[0]] * len(grp2), apply_kwargs=dict(maxdim=8, cutoff=1e-10))
    for i, det in enumerate(dets):
        assert ranks[i] == det["rank"] and rel(sv[i, :ranks[i]], det["svals"]) < MSG_TOL


def test_graph_replayed_gate_classes_follow_the_oracle_over_layers(E):
    # saturated bonds (random chi = 8 state, maxdim = 8): from the second layer on a colour class is captured into a CUDA graph and
    # from the third on replayed with only the gate matrices refreshed.  Four layers with a different angle per layer (the refresh
    # matters) against the oracle, layer by layer; then the same with the replay switched off must give bit-identical tensors.
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=8, seed=91)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    groups = G.grid_edge_color(g)
    kw = dict(maxdim=8, cutoff=1e-12, normalize_tensors=True)

    def layer(l):
        out = [("Rx", [g.names[v]], 0.3 + 0.1 * l) for v in range(g.nv)]
        for grp in groups:
            out += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 0.2 + 0.07 * l) for e in grp]
        return out

    def run(engine_state):
        bo = O.update(O.BeliefPropagationCache(psi_o.copy()), **O.default_bp_update_kwargs(psi_o))
        be = E.update(E.BeliefPropagationCache(engine_state, _share=True), inplace=True, **E.default_bp_update_kwargs(engine_state))
        for l in range(4):
            bo, errs_o = O.apply_gates(layer(l), bo, apply_kwargs=kw)
            be, errs_e = E.apply_gates(layer(l), be, apply_kwargs=kw, inplace=True)
            assert list(be.psi.bond_dims()) == list(bo.network.bond_dims()), f"layer {l}"
            assert np.allclose(errs_e, errs_o, rtol=1e-6, atol=1e-13), f"layer {l}"
            for v in (0, 5, 10, 15):
                assert abs(E.expect(be, ("Z", [g.names[v]])) - O.expect(bo, ("Z", [g.names[v]]))) < EXP_TOL, f"layer {l}, vertex {v}"
        return be

    b1 = run(E.TensorNetworkState.from_arrays(g, psi_o.tensors, chi_cap=8))
    assert b1.psi.maxvirtualdim() == 8
    import os
    os.environ["TNQS_GATE_GRAPH_OFF_FOR_TEST"] = "1"
    try:
        b2 = run(E.TensorNetworkState.from_arrays(g, psi_o.tensors, chi_cap=8))
    finally:
        os.environ.pop("TNQS_GATE_GRAPH_OFF_FOR_TEST")
    for a, b in zip(b1.psi.tensors(), b2.psi.tensors()):
        assert np.array_equal(a, b)


def test_optimistic_gates_with_rank_deficient_messages_take_the_conventional_pass(E):
    # a chi = 2 state embedded in zero-padded chi = 4 tensors, maxdim = 4: bonds are "saturated", so the batch starts optimistically, but
    # every message has two vanishing eigenvalues (sqrt(M) M^(-1/2) is a projector, not 1): k_eigh_sqrt's flag fails every commit
    # predicate and the gates run conventionally.  Gauge-invariant results against the oracle on the same padded tensors.
    g = G.named_grid((3, 3))
    small = O.random_tensornetworkstate(g, bond_dimension=2, seed=19)
    ts = []
    for v, t in enumerate(small.tensors):
        big = np.zeros((2,) + (4,) * g.degree(v), dtype=np.complex128)
        big[tuple(slice(0, n) for n in t.shape)] = t / np.linalg.norm(t)
        ts.append(big)
    psi_o = O.TensorNetworkState(g, ts)
    grp = G.grid_edge_color(g)[0]
    arr = gates.gate_array(("Rxx", [0, 1], 0.7))
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=8)
    be = E.update(to_engine(psi_o, chi_cap=4), maxiter=8)
    dets = []
    for e in grp:
        det = {}
        O.apply_gate(bo, arr, [int(g.esrc[e]), int(g.edst[e])], maxdim=4, cutoff=1e-12, detail=det)
        dets.append(det)
    ranks, errs, sv = E.apply_2site_batch(be, grp, [arr] * len(grp), apply_kwargs=dict(maxdim=4, cutoff=1e-12))
    for i, det in enumerate(dets):
        assert ranks[i] == det["rank"]
        assert rel(sv[i, :ranks[i]], det["svals"]) < 1e-8
    te = be.psi.tensors()
    for e in grp:
        assert rel(two_site(g, te, e), two_site(g, bo.network.tensors, e)) < 1e-7


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_circuits_on_random_states_follow_the_oracle(E, seed):
    # randomised end-to-end check of every gate path at once (conventional, optimistic, optimistic + W, graph capture and replay, the
    # fallbacks between them): a random lattice, ragged bond dimensions with some bonds already at maxdim, three layers of random
    # one- and two-site gates with a cutoff that lets some ranks drop; bond dimensions exact, <Z> within 1e-8 after every layer
    rs = np.random.default_rng(1000 + seed)
    shapes = [G.named_grid((3, 3)), G.named_grid((4, 3)), G.named_hexagonal_lattice_graph(2, 1), G.named_grid((2, 2, 2)), G.named_comb_tree((3, 3)), G.named_grid((5, 2))]
    g = shapes[seed % len(shapes)]
    maxdim = int(rs.choice([4, 6, 8]))
    dims = rs.integers(max(2, maxdim - 3), maxdim + 1, size=g.ne)
    dims[rs.random(g.ne) < 0.5] = maxdim                                   # half of the bonds saturated
    psi_o = ragged_state(g, dims, seed=seed)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    cutoff = float(rs.choice([1e-12, 1e-6, 1e-3]))
    kw = dict(maxdim=maxdim, cutoff=cutoff, normalize_tensors=True)
    groups = G.edge_color(g)
    bo = O.update(O.BeliefPropagationCache(psi_o.copy()), **O.default_bp_update_kwargs(psi_o))
    be = E.update(to_engine(psi_o, chi_cap=maxdim), inplace=True, **E.default_bp_update_kwargs(network_of(E, psi_o, maxdim)))
    names = ["Rxx", "Ryy", "Rzz"]
    for layer in range(3):
        circ = [("Rx", [g.names[v]], float(rs.uniform(0, 1))) for v in range(g.nv)]
        for grp in groups:
            nm = names[int(rs.integers(3))]
            th = float(rs.uniform(0.1, 0.6))
            for e in grp:
                a, b = int(g.esrc[e]), int(g.edst[e])
                vs = [g.names[b], g.names[a]] if rs.random() < 0.3 else [g.names[a], g.names[b]]
                circ.append((nm, vs, th))
        bo, errs_o = O.apply_gates(circ, bo, apply_kwargs=kw)
        be, errs_e = E.apply_gates(circ, be, apply_kwargs=kw, inplace=True)
        assert list(be.psi.bond_dims()) == list(bo.network.bond_dims()), f"seed {seed} layer {layer}"
        assert np.allclose(errs_e, errs_o, rtol=1e-5, atol=1e-12), f"seed {seed} layer {layer}"
        for v in range(g.nv):
            assert abs(E.expect(be, ("Z", [g.names[v]])) - O.expect(bo, ("Z", [g.names[v]]))) < EXP_TOL, f"seed {seed} layer {layer} vertex {v}"


def network_of(E, psi_o, cap):
    return E.TensorNetworkState.from_arrays(psi_o.g, psi_o.tensors, chi_cap=cap)


def _gate_of_operator_rank(rs, d, r):
    """A d^2 x d^2 two-site gate o[s1', s2', s1, s2] = sum_k A_k[s1', s1] B_k[s2', s2] with r generic terms (operator-Schmidt rank r)."""
    o = np.zeros((d, d, d, d), dtype=np.complex128)
    for _ in range(r):
        a = rs.normal(size=(d, d)) + 1j * rs.normal(size=(d, d))
        b = rs.normal(size=(d, d)) + 1j * rs.normal(size=(d, d))
        o += np.einsum("ab,cd->acbd", a, b)
    return o / np.linalg.norm(o) * d


@pytest.mark.parametrize("kw", [dict(maxdim=4, cutoff=1e-10, normalize_tensors=True),
                                dict(maxdim=8, cutoff=1e-12, normalize_tensors=False),
                                dict(maxdim=None, cutoff=None, normalize_tensors=True),     # no cutoff: the noise ranks of the full theta' are the answer
                                dict(maxdim=3, cutoff=None, normalize_tensors=True)])
def test_two_site_gates_of_every_operator_rank(E, kw):
    # the Jacobi SVD of a gate runs on the row space of theta' when (operator-Schmidt rank of the gate) x chi <= half its columns
    # (SvdDesc in kernels_small.cuh): ranks 1 and 2 take that path at chi = 4 (theta' 16 x 16 -> 16 x 4 / 16 x 8), ranks 3 and 4 and
    # every request that would have to report noise ranks take the full one - all against the oracle's LAPACK SVD of theta'
    g = G.named_grid((4, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=4, seed=71)
    rs = np.random.default_rng(7)
    named = [gates.gate_array(("Rzz", [0, 1], 0.37)), gates.gate_array(("Rxx", [0, 1], 1.1)), gates.gate_array(("CNOT", [0, 1])),
             gates.gate_array(("Rzz", [0, 1], 0.0)), gates.gate_array(("SWAP", [0, 1]))]
    for gi, grp in enumerate(G.grid_edge_color(g)[:2]):
        arrs = []
        for i, e in enumerate(grp):
            arrs.append(named[(i + gi) % len(named)] if i % 2 else _gate_of_operator_rank(rs, 2, 1 + (i // 2 + gi) % 4))
        gate_case(E, g, psi_o, grp, arrs, kw, flip=[(i // 2) % 2 for i in range(len(grp))])


def test_two_site_gates_of_low_operator_rank_on_qutrits_and_ragged_bonds(E):
    # d = 3 (a 9 x 9 operator matrix), ragged bonds (theta' not square: the reduction needs m >= n, the other gates run in full)
    g = G.named_grid((3, 3))
    rs = np.random.default_rng(8)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=72, d=3)
    grp = G.grid_edge_color(g)[0]
    arrs = [_gate_of_operator_rank(rs, 3, 1 + i % 3) for i in range(len(grp))]
    gate_case(E, g, psi_o, grp, arrs, dict(maxdim=3, cutoff=1e-10, normalize_tensors=True))
    psi_r = ragged_state(g, [2, 4, 3, 2, 4, 2, 3, 4, 2, 3, 4, 2], 73)
    for grp in G.grid_edge_color(g)[:2]:
        arrs = [_gate_of_operator_rank(rs, 2, 1 + i % 2) for i in range(len(grp))]
        gate_case(E, g, psi_r, grp, arrs, dict(maxdim=4, cutoff=1e-10, normalize_tensors=True), flip=[i % 2 for i in range(len(grp))])


def test_row_space_reduction_agrees_with_the_full_svd(E, monkeypatch):
    # same batch with and without the reduction: ranks equal, singular values to 1e-12, two-site tensors to 1e-10
    g = G.named_grid((3, 4))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=16, seed=74)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    grp = G.grid_edge_color(g)[0]
    arrs = [gates.gate_array(("Rzz", [0, 1], 0.1 + 0.2 * i)) for i in range(len(grp))]
    kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
    out = []
    for flag in ("1", "0"):
        monkeypatch.setenv("TNQS_SVD_ROWSPACE", flag)
        be = E.update(to_engine(psi_o, chi_cap=16), maxiter=4)
        ranks, errs, sv = E.apply_2site_batch(be, grp, arrs, apply_kwargs=kw)
        out.append((ranks, errs, sv, be.psi.tensors()))
    (r1, e1, s1, t1), (r0, e0, s0, t0) = out
    assert list(r1) == list(r0) and np.allclose(e1, e0, rtol=1e-8, atol=1e-18)
    assert rel(s1, s0) < 1e-12
    for e in grp:
        assert rel(two_site(g, t1, e), two_site(g, t0, e)) < 1e-10


def test_adaptive_second_cholesky_pass(E, monkeypatch):
    # CholeskyQR2 of the optimistic W path drops its second pass for matrices whose first factor is provably well-conditioned
    # (bound on kappa_2 <= 32): same ranks, singular values and two-site tensors as the two-pass run for a random (well-conditioned)
    # state, and an ill-conditioned state (site tensors with a graded leg) keeps the second pass and stays on the oracle
    g = G.named_grid((3, 4))
    grp = G.grid_edge_color(g)[0]
    arrs = [gates.gate_array(("Rzz", [0, 1], 0.1 + 0.2 * i)) for i in range(len(grp))]
    kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
    for graded in (False, True):
        psi_o = O.random_tensornetworkstate(g, bond_dimension=16, seed=75)
        for v in range(g.nv):
            if graded:                                         # columns of the matricised tensor spread over two and a half decades
                scale = np.logspace(0, -2.5, 16)
                shape = [1] * psi_o.tensors[v].ndim
                shape[1] = 16
                psi_o.tensors[v] = psi_o.tensors[v] * scale.reshape(shape)
            psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
        out = []
        for flag in ("1", "0"):
            monkeypatch.setenv("TNQS_CHOL_ADAPTIVE", flag)
            be = E.update(to_engine(psi_o, chi_cap=16), maxiter=4)
            n0 = be.psi.kernel_launches()
            ranks, errs, sv = E.apply_2site_batch(be, grp, arrs, apply_kwargs=kw)
            out.append((ranks, errs, sv, be.psi.tensors(), be.psi.kernel_launches() - n0))
        (r1, e1, s1, t1, l1), (r0, e0, s0, t0, l0) = out
        assert list(r1) == list(r0)
        assert rel(s1, s0) < 1e-11
        for e in grp:
            assert rel(two_site(g, t1, e), two_site(g, t0, e)) < 1e-9
        monkeypatch.setenv("TNQS_CHOL_ADAPTIVE", "1")
        gate_case(E, g, psi_o, grp, arrs, kw, sweeps=4)


@pytest.mark.parametrize("shape,chi", [((3, 2), 32), ((3, 3), 24)])
def test_low_rank_gates_at_large_bond(E, shape, chi):
    # theta' (128 x 128 at chi = 32) does not fit shared memory, its row-space matrix (128 x 64 for an Rzz) does: the Jacobi
    # iteration then runs in shared memory on W = theta' conj(Q) staged through the global work buffer; mixed with full-rank gates
    # (global iteration on theta') in one batch
    g = G.named_grid(shape)
    psi_o = O.random_tensornetworkstate(g, bond_dimension=chi, seed=84)
    for v in range(g.nv):
        psi_o.tensors[v] /= np.linalg.norm(psi_o.tensors[v])
    rs = np.random.default_rng(5)
    for gi, grp in enumerate(G.grid_edge_color(g)[:2]):
        arrs = []
        for i, e in enumerate(grp):
            if (i + gi) % 3 == 2:
                h = rs.normal(size=(4, 4)) + 1j * rs.normal(size=(4, 4))
                u, _ = np.linalg.qr(h)
                arrs.append(u.reshape(2, 2, 2, 2))
            else:
                arrs.append(gates.gate_array((("Rzz", "Rxx")[i % 2], [0, 1], 0.3 + 0.1 * i)))
        gate_case(E, g, psi_o, grp, arrs, dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True), sweeps=2, flip=[i % 2 for i in range(len(grp))])
