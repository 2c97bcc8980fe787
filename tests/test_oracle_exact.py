"""Pins the CPU oracle (oracle/tnqs_oracle.py) against independent ground truth: an exact statevector
simulator, BP exactness on trees and the reference's own property tests.  CPU only."""
import math
import numpy as np
import pytest

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates


def exact_norm_sqr(psi):
    sv = O.to_statevector(psi)
    return float(np.vdot(sv, sv).real)


def ising_layer(g, dt=0.25, hx=1.0, hz=0.8, J=0.5, groups=None):
    layer = [("Rx", [g.names[v]], 2 * hx * dt) for v in range(g.nv)]
    layer += [("Rz", [g.names[v]], 2 * hz * dt) for v in range(g.nv)]
    for grp in (groups or G.edge_color(g)):
        layer += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 2 * J * dt) for e in grp]
    return layer


def test_bp_exact_on_tree():
    # /root/reference/test/test_beliefpropagation.jl:32-54
    g = G.named_comb_tree((3, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=2, seed=123)
    bpc = O.BeliefPropagationCache(psi)
    assert not bpc.messages
    bpc = O.update(bpc)
    assert len(bpc.messages) == 2 * g.ne
    z = O.partitionfunction(bpc)
    assert z == pytest.approx(exact_norm_sqr(psi), rel=1e-12)
    sv = O.to_statevector(psi)
    for v in range(g.nv):
        ex = O.statevector_expect(sv, g.nv, {v: gates.Z})
        assert abs(O.expect(bpc, ("Z", [g.names[v]])) - ex) < 1e-13
    v = g.v((2, 1))
    rho = O.rdm_1site(bpc, v)
    full = sv.reshape((2,) * g.nv)
    full = np.moveaxis(full, v, 0).reshape(2, -1)
    rho_ex = full @ full.conj().T
    rho_ex /= np.trace(rho_ex)
    assert np.linalg.norm(rho - rho_ex) <= 1e-13


def test_two_site_expect_on_tree():
    g = G.named_comb_tree((3, 2))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=5)
    bpc = O.update(O.BeliefPropagationCache(psi))
    sv = O.to_statevector(psi)
    a, b = g.v((1, 1)), g.v((2, 1))
    ex = O.statevector_expect(sv, g.nv, {a: gates.Z, b: gates.Z})
    assert abs(O.expect(bpc, ("ZZ", [(1, 1), (2, 1)])) - ex) < 1e-13
    a, b = g.v((1, 2)), g.v((3, 2))          # non-adjacent: path region
    ex = O.statevector_expect(sv, g.nv, {a: gates.X, b: gates.Y})
    assert abs(O.expect(bpc, ("XY", [(1, 2), (3, 2)], 0.5)) - 0.5 * ex) < 1e-13


def test_two_qubit_circuit_known_answer():
    # /root/reference/test/test_apply.jl:11-20
    circuit = [("Rx", [(1, 1)], 0.5), ("Rx", [(2, 1)], 0.2), ("CPHASE", [(1, 1), (2, 1)], -0.3)]
    g = G.build_graph_from_circuit(circuit)
    psi0 = O.tensornetworkstate(lambda v: "↓", g)
    psi, errs = O.apply_gates(circuit, psi0, apply_kwargs=dict(maxdim=2, cutoff=1e-10, normalize_tensors=False))
    assert psi.maxvirtualdim() <= 2
    assert exact_norm_sqr(psi) == pytest.approx(1.0, abs=1e-13)
    sv = O.statevector_circuit(g, lambda v: "↓", circuit)
    got = O.to_statevector(psi)
    assert abs(np.vdot(sv, got)) == pytest.approx(1.0, abs=1e-13)


def test_ising_layer_grid_rank_and_norm():
    # /root/reference/test/test_apply.jl:23-53: Rzz has operator-Schmidt rank 2, cutoff removes the rest
    g = G.named_grid((3, 3))
    psi0 = O.random_tensornetworkstate(g, bond_dimension=1, seed=123)
    for v in range(g.nv):
        psi0.tensors[v] /= np.linalg.norm(psi0.tensors[v])
    info = {}
    bpc, errs = O.apply_gates(ising_layer(g, groups=G.grid_edge_color(g)), O.update(O.BeliefPropagationCache(psi0), **O.default_bp_update_kwargs(psi0)),
                              apply_kwargs=dict(cutoff=1e-10, normalize_tensors=False), info=info)
    assert bpc.network.maxvirtualdim() <= 2
    assert exact_norm_sqr(bpc.network) == pytest.approx(1.0, abs=1e-10)
    assert info["refreshes"] == 5            # 1 (first Rzz) + 3 (later classes) + 1 (final)
    assert np.all(errs < 1e-10)


def test_tree_circuit_matches_statevector():
    g = G.named_comb_tree((3, 2))
    layer = []
    rng = np.random.default_rng(0)
    for _ in range(3):
        layer += [("Rx", [g.names[v]], float(rng.uniform(0, 2))) for v in range(g.nv)]
        layer += [("Ry", [g.names[v]], float(rng.uniform(0, 2))) for v in range(g.nv)]
        for grp in G.edge_color(g):
            layer += [("Rzz", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 0.7) for e in grp]
            layer += [("CNOT", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]]) for e in grp[:1]]
    psi0 = O.tensornetworkstate(lambda v: "↑", g)
    psi, errs = O.apply_gates(layer, psi0, apply_kwargs=dict(cutoff=1e-14, normalize_tensors=True))
    sv = O.statevector_circuit(g, lambda v: "↑", layer)
    got = O.to_statevector(psi)
    fid = abs(np.vdot(sv, got)) ** 2 / (np.vdot(got, got).real * np.vdot(sv, sv).real)
    assert fid == pytest.approx(1.0, abs=1e-12)
    bpc = O.update(O.BeliefPropagationCache(psi))
    for v in range(g.nv):
        assert abs(O.expect(bpc, ("Z", [g.names[v]])) - O.statevector_expect(sv, g.nv, {v: gates.Z})) < 1e-11


def ghz(g):
    ts = []
    for v in range(g.nv):
        z = g.degree(v)
        t = np.zeros((2,) + (2,) * z, dtype=np.complex128)
        t[(0,) * (z + 1)] = 1
        t[(1,) * (z + 1)] = 1
        ts.append(t)
    return O.TensorNetworkState(g, ts)


def test_ghz_entanglement_log2():
    # /root/reference/test/test_constructors.jl:69-73
    g = G.named_hexagonal_lattice_graph(3, 3)
    psi = ghz(g)
    bpc = O.update(O.BeliefPropagationCache(psi))
    assert O.entanglement(bpc, 0) == pytest.approx(math.log(2), abs=1e-12)


@pytest.mark.parametrize("kind,tol", [("tree", 1e-12), ("grid", 1e-7)])
def test_symmetric_gauge_invariants(kind, tol):
    # tree: BP is exact after one sweep, so the fixed-point check is sharp; grid: limited by BP convergence
    g = G.named_comb_tree((3, 2)) if kind == "tree" else G.named_grid((3, 2))
    psi = O.random_tensornetworkstate(g, bond_dimension=2, seed=7)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=3 if kind == "tree" else 200)
    bpc = O.rescale(bpc)
    for v in range(g.nv):
        assert O.vertex_scalar(bpc, v) == pytest.approx(1.0, abs=1e-12)
    for e in range(g.ne):
        assert O.edge_scalar(bpc, e) == pytest.approx(1.0, abs=1e-12)
    sg = O.symmetric_gauge(bpc)
    a, b = O.to_statevector(bpc.network), O.to_statevector(sg.network)
    fid = abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)
    assert fid == pytest.approx(1.0, abs=1e-12)
    # diag(S) messages are a BP fixed point (up to scale)
    for k in range(2 * g.ne):
        new = O.updated_message(sg, k)
        old = sg.messages[k] / sg.messages[k].sum()
        assert np.linalg.norm(new - old) < tol


def test_truncate_bp():
    # /root/reference/test/test_truncate.jl:16-33
    g = G.named_hexagonal_lattice_graph(2, 2)
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=123)
    psi = O.gauge_and_scale(psi)
    tr = O.truncate_tns(psi, maxdim=2, cutoff=1e-10, normalize_tensors=False)
    assert tr.maxvirtualdim() <= 2
    a, b = O.to_statevector(psi), O.to_statevector(tr)
    f = abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)
    assert 0 <= f <= 1 + 1e-12
    same = O.truncate_tns(psi, maxdim=3, cutoff=None, normalize_tensors=False)     # no truncation: same state
    b = O.to_statevector(same)
    f = abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real)
    assert f == pytest.approx(1.0, abs=1e-9)


def test_truncation_rule():
    # SURVEY.md Appendix B.3
    p = np.array([1.0, 1e-3, 1e-6, 1e-12, 0.0])
    assert O.truncate_spectrum(p, None, None) == (5, 0.0)
    n, err = O.truncate_spectrum(p, None, 0.0)
    assert n == 4 and err == 0.0
    n, err = O.truncate_spectrum(p, None, 1e-10)
    assert n == 3 and err == pytest.approx(1e-12 / p.sum())
    n, err = O.truncate_spectrum(p, 2, 1e-10)
    assert n == 2 and err == pytest.approx((1e-6 + 1e-12) / p.sum())
    assert O.truncate_spectrum(np.array([0.3]), 1, 1e-3) == (1, 0.0)
    n, err = O.truncate_spectrum(np.array([1.0, 1e-20]), None, 1e-10)
    assert n == 1
    n, err = O.truncate_spectrum(np.array([0.0, 0.0]), None, 1e-10)
    assert n == 1 and err == 0.0


def test_message_normalisation_and_diff():
    g = G.named_grid((2, 2))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=3)
    bpc = O.BeliefPropagationCache(psi)
    m = O.updated_message(bpc, 0)
    assert m.sum() == pytest.approx(1.0)
    assert np.allclose(m, m.conj().T)
    info = {}
    out = O.update(bpc, maxiter=100, tolerance=1e-14, info=info)
    assert info["iters"] < 100 and info["mean_diff"] <= 1e-14
    assert O.message_diff(m, 3.0 * m) == pytest.approx(0.0, abs=1e-15)


def test_circuit_schedule_follows_the_reference_rule():
    # runs are defined by the affected-vertices rule, not by colour (apply_gates.jl:56-89): the host-side
    # schedule (engine.circuit_schedule, mirrored by tnqs_apply_circuit) must agree with the oracle's gate-by-gate loop
    from tensornetworkquantumsimulator_b200.engine import circuit_schedule
    for shape in [(3, 3), (4, 4), (5, 3)]:
        g = G.named_grid(shape)
        layer = ising_layer(g, groups=G.grid_edge_color(g))
        sched = circuit_schedule(g, layer * 2)
        info = {}
        psi0 = O.tensornetworkstate(lambda v: "↑", g)
        O.apply_gates(layer * 2, O.BeliefPropagationCache(psi0), apply_kwargs=dict(maxdim=2, cutoff=1e-10), info=info)
        assert sum(1 for it in sched if it[0] == "refresh") == info["refreshes"]
        # every gate exactly once; program order is preserved per vertex (one-site gates folded onto an earlier one of the same
        # vertex and batch keep their order inside the fold)
        ids = [i for it in sched if it[0] != "refresh" for i in it[1]] + [k for it in sched if it[0] == "gates1" for f in it[2] for k in f]
        assert sorted(ids) == list(range(len(layer) * 2))
        last = {}
        for it in sched:
            if it[0] == "refresh":
                continue
            groups_ = [[i] + (it[2][j] if it[0] == "gates1" else []) for j, i in enumerate(it[1])]
            for grp in groups_:
                assert grp == sorted(grp)
                for x in (layer * 2)[grp[0]][1]:
                    assert last.get(g.v(x), -1) < grp[0]
                    last[g.v(x)] = grp[-1]
        for it in sched:
            if it[0] == "gates2":                                        # a batch is vertex-disjoint
                vs = [g.v(x) for i in it[1] for x in (layer * 2)[i][1]]
                assert len(vs) == len(set(vs))


def test_bp_sampling_known_answers():
    # /root/reference/test/test_sampling.jl:15-21: a product state samples to its own configuration; GHZ samples are all-equal
    g = G.named_grid((3, 2))
    bits, _ = O.sample(O.tensornetworkstate(lambda v: "↑", g), 3, rng=np.random.default_rng(0))
    assert all(b == 0 for s in bits for b in s.values())
    gt = G.named_comb_tree((3, 2))
    outs, _ = O.sample(ghz(gt), 8, rng=np.random.default_rng(1))
    assert all(len(set(s.values())) == 1 for s in outs)
    assert {next(iter(s.values())) for s in outs} == {0, 1}


def test_bilinear_form_bp_is_exact_on_trees_and_reduces_to_the_norm():
    # inner(psi, phi; alg = "bp") (src/inner.jl:82-86 over src/Forms/bilinearform.jl:28-54): <phi|psi> exactly on a tree;
    # BilinearForm(psi, psi) is the norm network
    g = G.named_comb_tree((3, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=1)
    phi = O.random_tensornetworkstate(g, bond_dimension=3, seed=2)
    exact = np.vdot(O.to_statevector(phi), O.to_statevector(psi))
    assert abs(O.inner(psi, phi) - exact) < 1e-12 * abs(exact)
    assert abs(O.inner(phi, psi) - np.conj(exact)) < 1e-12 * abs(exact)
    n2 = O.partitionfunction(O.update(O.BeliefPropagationCache(psi)))
    assert abs(O.inner(psi, psi) - n2) < 1e-13 * abs(n2)
