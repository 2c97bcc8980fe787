"""The HOST logic of engine.py executed without a GPU: tests/fake_lib.py answers the C-ABI calls with the oracle, so what is tested
here is the Python side only - argument packing, column-major flattening, shapes, and the ORDER OF CALLS of the composed operations
(inner between different bond dimensions, observables over a Steiner tree, multi-site density matrices, the Heisenberg-picture
circuit).  The device side of the same calls is tests/test_gpu_parity.py / tests/test_gpu_compositions.py."""
import gc
import importlib.util
import os

import numpy as np
import pytest

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates, lib
from fake_lib import FakeLib
from util import to_engine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture
def E(monkeypatch):
    fake = FakeLib()
    monkeypatch.setattr(lib, "_lib", fake)
    from tensornetworkquantumsimulator_b200 import engine
    engine._fake = fake
    yield engine
    gc.collect()
    fake.close()


def test_basic_round_trips(E):
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=4)
    be = to_engine(psi_o)
    for a, b in zip(be.psi.tensors(), psi_o.tensors):
        assert np.array_equal(a, b)
    assert list(be.psi.site_dims()) == [2] * g.nv and be.psi.maxvirtualdim() == 3
    seq = G.forest_cover_edge_sequence(g)
    bo = O.update(O.BeliefPropagationCache(psi_o), maxiter=3, edge_sequence=seq)
    be = E.update(be, maxiter=3, edge_sequence=seq)
    for k in range(2 * g.ne):
        assert np.array_equal(E.message(be, k), O.message(bo, k))
    assert abs(E.expect(be, ("Z", [(2, 2)])) - O.expect(bo, ("Z", [(2, 2)]))) < 1e-14
    assert abs(E.expect(be, ("ZX", [(1, 1), (3, 2)])) - O.expect(bo, ("ZX", [(1, 1), (3, 2)]))) < 1e-14
    assert abs(E.partitionfunction(be) - O.partitionfunction(bo)) < 1e-12 * abs(O.partitionfunction(bo))


def test_heisenberg_example_through_the_host_api(E):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    d = np.load(os.path.join(GOLD, "heisenberg_grid4x4.npz"))
    g = G.named_grid((4, 4))
    o0 = E.paulitensornetworkstate(lambda v: "Z" if v == (2, 2) else "I", g, chi_cap=4)
    assert list(o0.site_dims()) == [4] * g.nv
    bpc = E.BeliefPropagationCache(o0.copy(), _share=True)
    for l in range(2):
        bpc, errs = E.apply_gates(mg.heisenberg_layer(g), bpc, apply_kwargs=dict(maxdim=4, cutoff=1e-14, normalize_tensors=False))
        bpc = E.rescale(bpc)
        assert list(bpc.psi.bond_dims()) == list(d["bond_dims"][l])
        assert abs(E.partitionfunction(bpc) - d["frobenius"][l]) < 1e-12
        n0 = len(E._fake.calls)
        assert abs(E.inner(bpc.psi, o0) - d["trace_ot_o0"][l]) < 1e-12           # ragged chi against chi = 1: zero-padded bra
        calls = E._fake.calls[n0:]
        assert calls.count("tnqs_set_bra") == 1 and calls.count("tnqs_set_message") == 2 * g.ne and calls[-1] == "tnqs_bp_update"


def test_inner_pads_the_thinner_layer(E):
    g = G.named_comb_tree((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=4, seed=71, d=4)
    phi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=72, d=4)
    exact = np.vdot(O.to_statevector(phi_o), O.to_statevector(psi_o))
    psi_e = E.TensorNetworkState.from_arrays(g, psi_o.tensors)
    phi_e = E.TensorNetworkState.from_arrays(g, phi_o.tensors)
    assert abs(E.inner(psi_e, phi_e) - exact) < 1e-12 * abs(exact)
    assert abs(E.inner(phi_e, psi_e) - np.conj(exact)) < 1e-12 * abs(exact)
    assert list(phi_e.bond_dims()) == [2] * g.ne                                   # the inputs are left alone
    with pytest.raises(ValueError):
        E.inner(psi_e, E.TensorNetworkState.from_arrays(G.named_grid((3, 3)), O.random_tensornetworkstate(G.named_grid((3, 3)), 2, 1).tensors))


def test_multi_site_observables_and_density_matrices(E):
    for g in (G.named_comb_tree((3, 3)), G.named_grid((4, 4))):
        psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=9)
        kw = dict(maxiter=30, tolerance=1e-12)
        bo = O.update(O.BeliefPropagationCache(psi_o), **kw)
        be = E.update(to_engine(psi_o), **kw)
        cases = [("ZXZ", [0, 4, 8]), ("XYZX", [2, 3, 6, 7])] if g.is_tree() else \
                [("ZXZ", [g.v((1, 1)), g.v((1, 3)), g.v((3, 2))]), ("ZZZ", [g.v((2, 1)), g.v((2, 2)), g.v((2, 3))])]
        for names, vs in cases:
            tree = G.steiner_tree(g, vs)
            n0 = len(E._fake.calls)
            got = E.expect(be, (names, vs), tree=tree)
            assert abs(got - O.expect(bo, (names, vs), tree=tree)) < 1e-12
            calls = E._fake.calls[n0:]
            # numerator clone: operators, then the bra, then one partial sweep; denominator clone: the sweep alone
            assert calls[:4] == ["tnqs_clone", "tnqs_apply_1site", "tnqs_set_bra", "tnqs_bp_update"] and calls.count("tnqs_clone") == 2
        assert abs(E.expect(be, ("ZXZ", cases[0][1], 0)) - 0) == 0
        before = [m.copy() for m in E.messages(be)]
        for vs in ([cases[0][1][0]], cases[0][1][:2], cases[0][1]):
            tree = G.steiner_tree(g, vs) if len(vs) > 2 else None
            assert np.max(np.abs(E.reduced_density_matrix(be, vs, tree=tree) - O.reduced_density_matrix(bo, vs, tree=tree))) < 1e-12
        for a, b in zip(before, E.messages(be)):                                   # measuring leaves the cache untouched
            assert np.array_equal(a, b)
    with pytest.raises(NotImplementedError):
        g = G.named_grid((3, 3))
        be = E.update(to_engine(O.random_tensornetworkstate(g, bond_dimension=2, seed=1)), maxiter=2)
        plaquette = [g.v((1, 1)), g.v((1, 2)), g.v((2, 2)), g.v((2, 1))]
        E.expect(be, ("ZZZZ", plaquette), tree=(plaquette, [g.edge_of[(plaquette[i], plaquette[i + 1])] for i in range(3)]))


def test_gate_batches_truncate_gauge_sampling_through_the_host_api(E):
    from util import ising_layer, two_site
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=3, seed=21)
    kw = dict(maxiter=20, tolerance=1e-10)
    bo = O.update(O.BeliefPropagationCache(psi_o), **kw)
    be = E.update(to_engine(psi_o, chi_cap=6), **kw)
    # one vertex-disjoint batch, second half of it flipped
    grp = G.grid_edge_color(g)[0]
    arr = gates.gate_array(("CRx", [0, 1], 0.7))
    flip = [i % 2 for i in range(len(grp))]
    dets = []
    for e, f in zip(grp, flip):
        a, c = int(g.esrc[e]), int(g.edst[e])
        det = {}
        O.apply_gate(bo, arr, [c, a] if f else [a, c], maxdim=4, cutoff=1e-12, detail=det)
        dets.append(det)
    ranks, errs, sv = E.apply_2site_batch(be, grp, [arr] * len(grp), flip=flip, apply_kwargs=dict(maxdim=4, cutoff=1e-12))
    assert list(ranks) == [d["rank"] for d in dets]
    for i, d in enumerate(dets):
        assert np.allclose(sv[i, :ranks[i]], d["svals"], rtol=0, atol=1e-14) and np.all(sv[i, ranks[i]:] == 0)
    for e in grp:
        assert np.array_equal(two_site(g, be.psi.tensors(), e), two_site(g, bo.network.tensors, e))
    with pytest.raises(lib.TnqsError) as ei:                       # two edges sharing a vertex in one batch
        E.apply_2site_batch(be, [0, [e for e in g.incident[int(g.esrc[0])] if e != 0][0]], [arr, arr])
    assert ei.value.status == -3
    # the circuit entry point and the run / refresh schedule agree on the number of refreshes
    layer = ising_layer(g, groups=G.grid_edge_color(g))
    info_e, info_o = {}, {}
    ce, errs_e = E.apply_gates(layer, be, apply_kwargs=dict(maxdim=4, cutoff=1e-10), info=info_e)
    co, errs_o = O.apply_gates(layer, bo, apply_kwargs=dict(maxdim=4, cutoff=1e-10), info=info_o)
    assert info_e["refreshes"] == info_o["refreshes"] == sum(1 for s in E.circuit_schedule(g, layer) if s[0] == "refresh")
    assert np.array_equal(errs_e, errs_o) and list(ce.psi.bond_dims()) == list(co.network.bond_dims())
    # truncate / symmetric gauge / entanglement / normalize
    te = E.truncate(ce, maxdim=2)
    to = O.truncate_bp(co, maxdim=2)
    assert te.psi.maxvirtualdim() == 2 and abs(E.expect(te, ("Z", [(2, 2)])) - O.expect(to, ("Z", [(2, 2)]))) < 1e-13
    ge, go = E.symmetric_gauge(ce), O.symmetric_gauge(co)
    for e in range(g.ne):
        assert abs(E.entanglement(ge, e) - O.entanglement(go, e)) < 1e-12
    ne = E.normalize(ce.psi)
    assert abs(E.norm_sqr(ne) - 1) < 1e-10
    # sampling: the same draws as the oracle with the same generator
    small = O.random_tensornetworkstate(G.named_comb_tree((2, 3)), bond_dimension=2, seed=5)
    se, _ = E.sample(E.TensorNetworkState.from_arrays(small.g, small.tensors), 4, rng=np.random.default_rng(7))
    so, _ = O.sample(small, 4, rng=np.random.default_rng(7))
    assert se == so


def test_single_layer_network_contraction(E):
    # contract(tn; alg = "bp") (src/contract.jl:7-9, test/test_beliefpropagation.jl:12-28): exact on a tree; on a loopy graph the
    # engine's bilinear embedding follows the oracle's vector-message BP sweep by sweep
    rs = np.random.default_rng(12)
    g = G.named_comb_tree((3, 3))
    ts = [rs.normal(size=(2,) * g.degree(v)) + 1j * rs.normal(size=(2,) * g.degree(v)) for v in range(g.nv)]
    exact = O.contract_exact(g, ts)
    assert abs(O.contract_bp(g, ts) - exact) < 1e-12 * abs(exact)
    assert abs(E.contract(ts, g) - exact) < 1e-12 * abs(exact)
    g = G.named_grid((3, 3))
    ts = [rs.uniform(0.5, 1.5, size=(3,) * g.degree(v)) + 0.1j * rs.normal(size=(3,) * g.degree(v)) for v in range(g.nv)]
    for sweeps in (1, 2, 5):
        kw = dict(maxiter=sweeps, tolerance=1e-14)
        io, ie = {}, {}
        zo = O.contract_bp(g, ts, info=io, **kw)
        ze = E.contract(ts, g, bp_update_kwargs=kw, info=ie)
        assert abs(ze - zo) < 1e-12 * abs(zo) and ie["iters"] == io["iters"] and abs(ie["mean_diff"] - io["mean_diff"]) < 1e-13
    zo = O.contract_bp(g, ts)
    assert abs(E.contract(ts, g) - zo) < 1e-10 * abs(zo)
    assert abs(zo - O.contract_exact(g, ts)) < 0.05 * abs(zo)               # BP is approximate on loops, close for a near-positive network


def test_integer_vertex_names_resolve_by_name(E):
    # circuits with integer vertex labels (graph_ops.jl:53-67 numbers nothing: [("Rx", [3], ..), ("CNOT", [3, 10]), ..]): an integer
    # is a NAME there, never a dense id.  Truth = a hand-written 3-qubit statevector, independent of any name resolution.
    circuit = [("Rx", [3], 0.7), ("Rx", [10], 0.3), ("Rx", [1], 1.1), ("CNOT", [3, 10]), ("CNOT", [10, 1]), ("Rx", [1], 0.4)]
    g = G.build_graph_from_gates(circuit)
    assert g.names == [3, 10, 1] and g.ne == 2                      # ids 0, 1, 2 by first appearance: name 1 is id 2, name 3 is id 0
    psi = E.tensornetworkstate(lambda v: "↑", g, chi_cap=4)
    out, errs = E.apply_gates(circuit, psi, apply_kwargs=dict(maxdim=4, cutoff=1e-14))
    rx = lambda t: np.array([[np.cos(t / 2), -1j * np.sin(t / 2)], [-1j * np.sin(t / 2), np.cos(t / 2)]])
    cnot = np.eye(4)[[0, 1, 3, 2]].astype(complex)
    I2 = np.eye(2)
    sv = np.zeros(8, complex); sv[0] = 1                            # qubit order (3, 10, 1), first slowest
    sv = np.kron(np.kron(rx(0.7), rx(0.3)), rx(1.1)) @ sv
    sv = np.kron(cnot, I2) @ sv
    sv = np.kron(I2, cnot) @ sv
    sv = np.kron(np.kron(I2, I2), rx(0.4)) @ sv
    Z = np.diag([1.0, -1.0])
    want = {3: np.vdot(sv, np.kron(np.kron(Z, I2), I2) @ sv).real, 10: np.vdot(sv, np.kron(np.kron(I2, Z), I2) @ sv).real,
            1: np.vdot(sv, np.kron(np.kron(I2, I2), Z) @ sv).real}
    for name, w in want.items():
        assert abs(E.expect(out, ("Z", [name])) - w) < 1e-12        # a path is a tree: BP is exact
    assert abs(E.expect(out, ("ZZ", [3, 1])) - np.vdot(sv, np.kron(np.kron(Z, I2), Z) @ sv).real) < 1e-12
    assert out[10].shape[0] == 2 and out[10].ndim == 3              # getindex by name: the middle vertex has two bonds
    with pytest.raises(KeyError):
        E.expect(out, ("Z", [0]))                                   # 0 is not a vertex of this graph
    assert E.circuit_schedule(g, circuit)[-1] == ("refresh",)
    # the oracle resolves the same way
    bo, _ = O.apply_gates(circuit, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=dict(maxdim=4, cutoff=1e-14))
    assert abs(O.expect(O.update(O.BeliefPropagationCache(bo)), ("Z", [1])) - want[1]) < 1e-12


def test_cache_constructor_does_not_alias_the_state(E):
    # BeliefPropagationCache(psi) wraps a clone: a second cache over the same state keeps the first one's messages intact and
    # in-place work on a cache never reaches psi (the reference's constructor does not alias either)
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=8)
    psi = E.TensorNetworkState.from_arrays(g, psi_o.tensors, chi_cap=4)
    b1 = E.update(E.BeliefPropagationCache(psi), maxiter=4, inplace=True)
    m1 = [m.copy() for m in E.messages(b1)]
    b2 = E.BeliefPropagationCache(psi)
    for a, b in zip(E.messages(b1), m1):
        assert np.array_equal(a, b)
    E.apply_1site_batch(b2, [0], [gates.gate_array(("X", [0]))], normalize_tensors=False)
    assert np.array_equal(psi[(1, 1)], psi_o.tensors[0]) and not np.array_equal(b2.psi[(1, 1)], psi_o.tensors[0])
