"""Loop corrections (src/MessagePassing/loopcorrection.jl, loop_correlation(s) in
src/MessagePassing/beliefpropagationcache.jl:144-197) on the CPU: what pins the oracle twin, the host enumeration of the generalised
loops, and the host logic of engine.py over the oracle-backed fake library.  The device side is tests/test_gpu_compositions.py.

What pins the oracle: the loop series is EXACT when every leafless edge-induced subgraph is included and BP has converged -
Z_exact = Z_bp (1 + sum of all weights) - so weights, antiprojector orientation and the enumeration (disconnected unions included)
are checked together against the exact statevector norm / the exact contraction."""
import gc
import itertools

import numpy as np
import pytest

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, lib
from fake_lib import FakeLib
from util import to_engine


def _brute_force_no_leaves(g, mx):
    ends = list(zip(g.esrc.tolist(), g.edst.tolist()))
    out = []
    for k in range(1, mx + 1):
        for S in itertools.combinations(range(g.ne), k):
            deg = {}
            for e in S:
                for x in ends[e]:
                    deg[x] = deg.get(x, 0) + 1
            if all(d >= 2 for d in deg.values()):
                out.append(S)
    return out


@pytest.mark.parametrize("dims,mx", [((3, 3), 12), ((2, 4), 10), ((2, 5), 9), ((3, 4), 8)])
def test_generalised_loops_match_the_definition(dims, mx):
    g = G.named_grid(dims)
    got = G.edgeinduced_subgraphs_no_leaves(g, mx)
    assert got == sorted(_brute_force_no_leaves(g, mx), key=lambda c: (len(c), c))
    assert len(set(got)) == len(got)


def test_generalised_loops_of_the_reference_example():
    # examples/loopcorrections.jl:15-28: max_configuration_size = 2 * smallest loop - 1
    g = G.named_grid((4, 4))
    egs = G.edgeinduced_subgraphs_no_leaves(g, 7)
    assert [len(x) for x in egs] == [4] * 9 + [6] * 12 + [7] * 12          # plaquettes, 2x1 rectangles, the same with their middle edge
    assert G.edgeinduced_subgraphs_no_leaves(G.named_grid((4, 1)), 5) == []
    h = G.named_hexagonal_lattice_graph(2, 2)
    cyc = G.simple_cycles_limited_length(h, 6)
    assert all(len(c) == 6 for c in cyc) and len(cyc) == 4
    assert G.edgeinduced_subgraphs_no_leaves(h, 6) == cyc
    # a periodic ring has exactly one loop
    ring = G.named_grid((5, 1), periodic=True)
    assert G.edgeinduced_subgraphs_no_leaves(ring, 5) == [tuple(range(5))]
    assert [len(w[0]) + 1 for w in G.cycle_walks(ring, 5)] == [5]


def _converged(psi):
    return O.update(O.BeliefPropagationCache(psi.copy()), maxiter=3000, tolerance=1e-16)


@pytest.mark.parametrize("dims,chi,seed", [((2, 2), 3, 5), ((2, 3), 2, 6), ((3, 3), 2, 7), ((2, 4), 2, 8)])
def test_full_loop_series_is_exact(dims, chi, seed):
    g = G.named_grid(dims)
    psi = O.random_tensornetworkstate(g, bond_dimension=chi, seed=seed)
    bpc = _converged(psi)
    sv = O.to_statevector(psi)
    exact = complex(np.vdot(sv, sv))
    zbp = O.partitionfunction(bpc)
    zall = O.loopcorrected_partitionfunction(bpc, g.ne)
    assert abs(zall - exact) < 1e-8 * abs(exact)
    assert abs(zbp - exact) > 1e-3 * abs(exact)                         # BP alone is visibly off: the corrections carry the difference


def _k4():
    return G._from_named_edges(list(range(4)), [(a, b) for a in range(4) for b in range(a + 1, 4)])


@pytest.mark.parametrize("which", ["k4", "cube"])
def test_full_loop_series_is_exact_with_triangles_and_high_degree(which):
    # K4 (triangles, every vertex of loop degree 3 in the largest subgraphs) and the 2 x 2 x 2 cube (five independent cycles):
    # enumeration against brute force and the series against the exact norm
    g = _k4() if which == "k4" else G.named_grid((2, 2, 2))
    egs = G.edgeinduced_subgraphs_no_leaves(g, g.ne)
    assert egs == sorted(_brute_force_no_leaves(g, g.ne), key=lambda c: (len(c), c))
    psi = O.random_tensornetworkstate(g, bond_dimension=2, seed=3 if which == "k4" else 4)
    bpc = O.update(O.BeliefPropagationCache(psi.copy()), maxiter=5000, tolerance=1e-16)
    sv = O.to_statevector(psi)
    exact = complex(np.vdot(sv, sv))
    assert abs(O.loopcorrected_partitionfunction(bpc, g.ne) - exact) < 1e-8 * abs(exact)


def test_bilinear_and_single_layer_series_are_exact():
    # inner(psi, phi; alg = "loopcorrections") (src/inner.jl:67-86): the same weights on the messages of the bilinear form
    g = G.named_grid((2, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=2, seed=11)
    phi = O.random_tensornetworkstate(g, bond_dimension=2, seed=12)
    for v in range(g.nv):                                              # phi close to psi: BP on <phi|psi> converges
        phi.tensors[v] = psi.tensors[v] + 0.2 * phi.tensors[v]
    bpc = O.update(O.BeliefPropagationCache(psi.copy(), {}, phi, delta_default=True), maxiter=3000, tolerance=1e-16)
    exact = complex(np.vdot(O.to_statevector(phi), O.to_statevector(psi)))
    assert abs(O.loopcorrected_partitionfunction(bpc, g.ne) - exact) < 1e-8 * abs(exact)


def test_weights_vanish_where_bp_is_exact():
    # a product state has rank-one messages: every antiprojector annihilates them
    g = G.named_grid((3, 3))
    psi = O.tensornetworkstate(lambda v: "X+", g)
    bpc = O.rescale(_converged(psi))
    ws = O.loop_weights(bpc, G.edgeinduced_subgraphs_no_leaves(g, 6))
    assert np.max(np.abs(ws)) < 1e-14
    # a subgraph with a leaf has weight 0 at the BP fixed point (why only leafless ones enter the series)
    psi = O.random_tensornetworkstate(g, bond_dimension=2, seed=3)
    bpc = O.rescale(_converged(psi))
    plaq = G.edgeinduced_subgraphs_no_leaves(g, 4)[0]
    tail = next(e for e in range(g.ne) if e not in plaq and (int(g.esrc[e]) in {int(g.esrc[f]) for f in plaq} | {int(g.edst[f]) for f in plaq}))
    assert abs(O.loop_weight(bpc, plaq + (tail,))) < 1e-9
    assert abs(O.loop_weight(bpc, plaq)) > 1e-6


def test_loop_transfer_matrix_and_correlation():
    # on a ring the trace of the transfer matrix is the exact norm; loop_correlation is 0 for a product state and on a tree
    ring = G.named_grid((5, 1), periodic=True)
    psi = O.random_tensornetworkstate(ring, bond_dimension=3, seed=9)
    bpc = _converged(psi)
    (loop, target), = G.cycle_walks(ring, 5)
    m = O.loop_transfer_matrix(bpc, loop, target)
    sv = O.to_statevector(psi)
    assert abs(np.trace(m) - np.vdot(sv, sv)) < 1e-10 * abs(np.vdot(sv, sv))
    c = O.loop_correlation(bpc, loop, target)
    assert 0 < c < 1
    # the leading eigenvector of the transfer matrix is the BP fixed point (message into the far vertex), the others carry the
    # correlation around the loop: with them removed by the antiprojector the loop weight is the sum of the sub-leading ones
    lam = np.linalg.eigvals(m)
    lam = lam[np.argsort(-np.abs(lam))]
    w = O.loop_weight(O.rescale(bpc), tuple(range(ring.ne)))
    assert abs(w - np.sum(lam[1:]) / lam[0]) < 1e-8
    prod = O.tensornetworkstate(lambda v: "Z+", G.named_grid((3, 3)))
    assert max(O.loop_correlations(_converged(prod), 4)) < 1e-14
    tree = G.named_comb_tree((3, 2))
    assert O.loop_correlations(_converged(O.random_tensornetworkstate(tree, bond_dimension=2, seed=1)), 6) == []


# ---- host logic of engine.py over the fake library ----
@pytest.fixture
def E(monkeypatch):
    fake = FakeLib()
    monkeypatch.setattr(lib, "_lib", fake)
    from tensornetworkquantumsimulator_b200 import engine
    engine._fake = fake
    yield engine
    gc.collect()
    fake.close()


def test_engine_host_logic_of_loop_corrections(E):
    g = G.named_grid((3, 3))
    psi_o = O.random_tensornetworkstate(g, bond_dimension=2, seed=21)
    kw = dict(maxiter=60, tolerance=1e-14)
    bo = O.update(O.BeliefPropagationCache(psi_o.copy()), **kw)
    be = E.update(to_engine(psi_o), **kw)
    msgs_before = [E.message(be, k).copy() for k in range(2 * g.ne)]
    for size in (4, 7):
        n0 = len(E._fake.calls)
        z = E.loopcorrected_partitionfunction(be, size)
        assert E._fake.calls[n0:] == ["tnqs_clone", "tnqs_rescale", "tnqs_loop_weights"]     # rescale works on a copy
        zo = O.loopcorrected_partitionfunction(bo, size)
        assert abs(z - zo) < 1e-12 * abs(zo)
    # the caller's cache is not rescaled (loopcorrection.jl:8 works on a copy)
    for k in range(2 * g.ne):
        assert np.array_equal(E.message(be, k), msgs_before[k])
    psi_e = be.psi
    assert abs(E.norm_sqr(psi_e, alg="loopcorrections", max_configuration_size=4, cache_update_kwargs=kw) - O.norm_sqr_loopcorrected(psi_o, 4, **kw)) < 1e-12
    assert abs(E.norm(be, alg="loopcorrections", max_configuration_size=4) ** 2 - O.loopcorrected_partitionfunction(bo, 4)) < 1e-12
    with pytest.raises(ValueError):
        E.norm_sqr(be, alg="loopcorrections")
    with pytest.raises(ValueError):
        E.norm_sqr(be, alg="boundarymps")
    # a tree has no loops: the BP value comes back and the device is not asked
    tree = G.named_comb_tree((2, 3))
    pt = O.random_tensornetworkstate(tree, bond_dimension=2, seed=2)
    bt = E.update(to_engine(pt))
    n0 = len(E._fake.calls)
    assert E.loopcorrected_partitionfunction(bt, 6) == E.partitionfunction(bt)
    assert "tnqs_loop_weights" not in E._fake.calls[n0:]
    # loop correlations: walk packing and matrix layout
    co = O.loop_correlations(bo, 4)
    ce = E.loop_correlations(be, 4)
    assert len(ce) == 4 and np.allclose(ce, co, atol=1e-12)
    assert E.loop_correlations(bt, 6) == []
    # inner with loop corrections runs the bilinear cache through the same path
    phi_o = psi_o.copy()
    for v in range(g.nv):
        phi_o.tensors[v] = phi_o.tensors[v] + 0.1 * O.random_tensornetworkstate(g, bond_dimension=2, seed=22).tensors[v]
    ze = E.inner(to_engine(psi_o).psi, to_engine(phi_o).psi, alg="loopcorrections", max_configuration_size=4, cache_update_kwargs=kw)
    bb = O.update(O.BeliefPropagationCache(psi_o.copy(), {}, phi_o, delta_default=True), **kw)
    assert abs(ze - O.loopcorrected_partitionfunction(bb, 4)) < 1e-12 * abs(ze)
