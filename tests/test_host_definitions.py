"""Host-side definitions shared by the engine and the oracle checked against independent derivations,
so that a slip there cannot hide behind "engine == oracle": gate matrices from their generators / projector formulas
(src/Apply/gate_definitions.jl:28-57,98-128; ITensors "Qubit" ops), the BP edge sequence
(src/MessagePassing/beliefpropagationcache.jl:74-85) and the edge colourings (src/truncate.jl:20) from their defining properties."""
import numpy as np
import pytest
from scipy.linalg import expm

from tensornetworkquantumsimulator_b200 import graphs as G, gates

I2, X, Y, Z = np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0])
P0, P1 = np.diag([1.0, 0.0]), np.diag([0.0, 1.0])
kron = np.kron


def test_gate_definitions_from_generators():
    th = 0.731
    cases = {
        ("Rx", th): expm(-0.5j * th * X), ("Ry", th): expm(-0.5j * th * Y), ("Rz", th): expm(-0.5j * th * Z),
        # tuples carry qiskit's angle; the ITensors op takes half of it: exp(-i (theta/2) PP)
        ("Rxx", th): expm(-0.5j * th * kron(X, X)), ("Ryy", th): expm(-0.5j * th * kron(Y, Y)), ("Rzz", th): expm(-0.5j * th * kron(Z, Z)),
        ("CRx", th): kron(P0, I2) + kron(P1, expm(-0.5j * th * X)), ("CRy", th): kron(P0, I2) + kron(P1, expm(-0.5j * th * Y)),
        ("CRz", th): kron(P0, I2) + kron(P1, expm(-0.5j * th * Z)),
        ("P", th): P0 + np.exp(1j * th) * P1, ("CPHASE", th): kron(P0, I2) + kron(P1, P0 + np.exp(1j * th) * P1),
        ("Rxxyy", th): expm(-0.5j * th * (kron(X, X) + kron(Y, Y))),
        ("Rxxyyzz", th): expm(-0.5j * th * (kron(X, X) + kron(Y, Y) + kron(Z, Z))),
        ("CNOT", None): kron(P0, I2) + kron(P1, X), ("CX", None): kron(P0, I2) + kron(P1, X), ("CY", None): kron(P0, I2) + kron(P1, Y),
        ("CZ", None): kron(P0, I2) + kron(P1, Z),
        ("SWAP", None): 0.5 * (kron(I2, I2) + kron(X, X) + kron(Y, Y) + kron(Z, Z)),
        ("iSWAP", None): expm(0.25j * np.pi * (kron(X, X) + kron(Y, Y))),
        ("H", None): (X + Z) / np.sqrt(2), ("S", None): P0 + 1j * P1, ("T", None): P0 + np.exp(0.25j * np.pi) * P1,
        ("X", None): X, ("Y", None): Y, ("Z", None): Z,
    }
    for (name, par), want in cases.items():
        nq = int(np.log2(want.shape[0]))
        tup = (name, list(range(nq))) if par is None else (name, list(range(nq)), par)
        got = gates.gate_array(tup).reshape(want.shape)            # o[s1', s2', s1, s2] row-major = the Kronecker matrix, first site slow
        assert np.max(np.abs(got - want)) < 1e-14, name
        assert np.max(np.abs(got @ got.conj().T - np.eye(2 ** nq))) < 1e-14
    # xx_plus_yy(theta, beta) (gate_definitions.jl:98-105): exp(-i theta/4 (XX + YY)) conjugated by Rz(beta) on the first site
    th, be = 0.9, 0.4
    rz = lambda a: expm(-0.5j * a * Z)
    want = kron(rz(be), I2) @ expm(-0.25j * th * (kron(X, X) + kron(Y, Y))) @ kron(rz(-be), I2)
    assert np.max(np.abs(gates.gate_array(("xx_plus_yy", [0, 1], (th, be))).reshape(4, 4) - want)) < 1e-14
    # Pauli strings: one letter per vertex
    assert np.array_equal(gates.gate_array(("XZ", [0, 1])).reshape(4, 4), kron(X, Z))
    with pytest.raises(ValueError):
        gates.gate_array(("NoSuchGate", [0]))
    # states
    for name, vec in (("↑", [1, 0]), ("↓", [0, 1]), ("X+", [1, 1]), ("X-", [1, -1]), ("Y+", [1, 1j]), ("Y-", [1, -1j])):
        v = gates.state_vector(name)
        assert abs(abs(np.vdot(v, np.array(vec) / np.linalg.norm(vec))) - 1) < 1e-15


GRAPHS = [G.named_grid((4, 5)), G.named_grid((3, 3, 2)), G.named_comb_tree((3, 4)), G.named_hexagonal_lattice_graph(2, 3),
          G.heavy_hexagonal_lattice(2, 2), G.eagle_heavy_hex(), G.named_grid((4, 4), periodic=True)]


@pytest.mark.parametrize("g", GRAPHS, ids=lambda g: f"nv{g.nv}ne{g.ne}")
def test_forest_cover_edge_sequence_properties(g):
    seq = G.forest_cover_edge_sequence(g).tolist()
    assert sorted(seq) == list(range(2 * g.ne))                        # every directed edge exactly once
    # the sequence is a concatenation of blocks [up edges of a tree] + [the same reversed and flipped]; inside the first half an edge
    # u -> p comes after every tree edge c -> u of u's children (post-order), so on a tree one sweep is exact
    pos = {k: i for i, k in enumerate(seq)}
    for k in seq:
        assert abs(pos[k] - pos[k ^ 1]) % 2 == 1                        # mirror positions inside one block
    if g.is_tree():
        up = seq[:g.ne]
        assert seq[g.ne:] == [k ^ 1 for k in reversed(up)]
        done = set()
        for k in up:
            u, p = g.dsrc(k), g.ddst(k)
            for e in g.incident[u]:
                w = g.other(e, u)
                if w != p:
                    assert g.dedge(w, u) in done, "a child's message must be updated before its parent's"
            done.add(k)


@pytest.mark.parametrize("g", GRAPHS, ids=lambda g: f"nv{g.nv}ne{g.ne}")
def test_edge_colouring_is_proper(g):
    groups = G.edge_color(g)
    flat = [e for grp in groups for e in grp]
    assert sorted(flat) == list(range(g.ne))
    maxdeg = max(g.degree(v) for v in range(g.nv))
    assert len(groups) <= maxdeg + 1                                    # Vizing bound; bipartite lattices reach maxdeg
    for grp in groups:
        touched = [int(g.esrc[e]) for e in grp] + [int(g.edst[e]) for e in grp]
        assert len(set(touched)) == len(touched), "two edges of one colour class share a vertex"


def test_steiner_tree_properties():
    g = G.named_grid((5, 5))
    for terms in ([0, 24], [0, 4, 20], [12, 0, 24, 4], [7]):
        vs, es = G.steiner_tree(g, terms)
        assert set(terms) <= set(vs) and len(es) == len(vs) - 1
        assert len(G.connected_components(g, allowed_edges=set(es))[0]) >= 1
        sub = set()
        for e in es:
            sub |= {int(g.esrc[e]), int(g.edst[e])}
        assert sub == set(vs) or len(vs) == 1
    vs, es = G.steiner_tree(g, [0, 24])
    assert len(es) == 8                                                  # a shortest path for two terminals
    with pytest.raises(ValueError):
        G.steiner_tree(g, [3, 3])


def test_hexagonal_lattice_is_the_networkx_generator():
    # NamedGraphs.named_hexagonal_lattice_graph follows networkx.hexagonal_lattice_graph (vertex names shifted to 1-based): same
    # vertices, same edges, no dangling corner; heavy_hexagonal_lattice(5, 5) of examples/heavyhexIsing_dynamics.jl:9 then has the
    # 164 vertices / 188 edges SURVEY section 8d quotes
    nx = pytest.importorskip("networkx")
    for m, n in ((1, 1), (2, 1), (1, 2), (2, 2), (2, 3), (3, 3), (4, 4)):
        h = nx.hexagonal_lattice_graph(m, n)
        g = G.named_hexagonal_lattice_graph(m, n)
        want = sorted(tuple(sorted(((a[0] + 1, a[1] + 1), (b[0] + 1, b[1] + 1)))) for a, b in h.edges())
        got = sorted(tuple(sorted((g.names[s], g.names[t]))) for s, t in zip(g.esrc.tolist(), g.edst.tolist()))
        assert got == want and g.nv == h.number_of_nodes()
        assert min(g.degree(v) for v in range(g.nv)) == 2
    hh = G.heavy_hexagonal_lattice(5, 5)
    assert (hh.nv, hh.ne) == (164, 188)


def _dependency_levels(g, seq):
    """Levels of a sequential (Gauss-Seidel) sweep as the engine builds them: an update is placed strictly after
    the last write of every message it reads (RAW), after the last read of the message it writes (WAR) and after its last write (WAW)."""
    into = {}
    for k in range(2 * g.ne):
        into.setdefault(g.ddst(k), []).append(k)
    last_write = np.full(2 * g.ne, -1)
    last_read = np.full(2 * g.ne, -1)
    level = {}
    for k in (int(x) for x in seq):
        ins = [m for m in into[g.dsrc(k)] if m != (k ^ 1)]
        lv = max([0] + [last_write[m] + 1 for m in ins] + [last_read[k] + 1, last_write[k] + 1])
        level[k] = lv
        for m in ins:
            last_read[m] = max(last_read[m], lv)
        last_write[k] = lv
    return level


@pytest.mark.parametrize("dims", [(16, 16), (8, 8, 8)])
def test_reference_order_never_updates_two_messages_of_an_interior_vertex_in_one_level(dims):
    # why the shared partial products of the JACOBI sweeps (DESIGN.md sections 4 and 6) have no counterpart in the reference's order:
    # in the forest-cover sequence of an open grid the updates that leave one interior vertex always sit in different dependency
    # levels, i.e. some message they read changes in between
    g = G.named_grid(dims)
    level = _dependency_levels(g, G.forest_cover_edge_sequence(g))
    assert len(level) == 2 * g.ne
    zmax = 2 * len(dims)
    seen = set()
    for k, lv in level.items():
        v = g.dsrc(k)
        if g.degree(v) == zmax:
            assert (v, lv) not in seen
            seen.add((v, lv))
    # levels are what makes the order parallel at all: far fewer levels than messages
    assert max(level.values()) + 1 < g.ne // 4
