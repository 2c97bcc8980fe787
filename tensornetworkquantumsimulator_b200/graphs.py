"""Lattice builders, edge colourings and BP edge sequences (host side, integer work only).

Mirrors the graph helpers the reference takes from NamedGraphs / SimpleGraphAlgorithms and
`/root/reference/src/graph_ops.jl` (heavy_hexagonal_lattice :6-18, build_graph_from_gates :53-67).
The reference names vertices by tuples; here a `Graph` keeps the tuple names next to dense 0-based
ids, because the C-ABI (include/tnqs_cuda.h) speaks int32 ids.

Conventions shared by the oracle, the engine and the C-ABI:
  * undirected edge e = (esrc[e], edst[e]); directed edge id k = 2*e + dir, dir 0 = esrc->edst;
  * the legs of the site tensor of v are ordered by ascending incident edge id (`incident[v]`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Hashable, Iterable, List, Sequence, Tuple

import numpy as np


@dataclass
class Graph:
    names: List[Hashable]                      # vertex id -> name (tuple for lattices)
    esrc: np.ndarray                           # int32 [ne]
    edst: np.ndarray                           # int32 [ne]
    index: Dict[Hashable, int] = field(default_factory=dict)
    incident: List[List[int]] = field(default_factory=list)   # v -> ascending edge ids
    edge_of: Dict[Tuple[int, int], int] = field(default_factory=dict)

    def __post_init__(self):
        self.esrc = np.asarray(self.esrc, dtype=np.int32)
        self.edst = np.asarray(self.edst, dtype=np.int32)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.incident = [[] for _ in self.names]
        self.edge_of = {}
        for e, (a, b) in enumerate(zip(self.esrc.tolist(), self.edst.tolist())):
            if a == b or (a, b) in self.edge_of:
                raise ValueError(f"self loop or duplicate edge {a}-{b}")
            self.incident[a].append(e)
            self.incident[b].append(e)
            self.edge_of[(a, b)] = e
            self.edge_of[(b, a)] = e

    # ---- sizes ----
    @property
    def nv(self) -> int:
        return len(self.names)

    @property
    def ne(self) -> int:
        return int(self.esrc.shape[0])

    def degree(self, v: int) -> int:
        return len(self.incident[v])

    def other(self, e: int, v: int) -> int:
        a, b = int(self.esrc[e]), int(self.edst[e])
        return b if v == a else a

    def neighbors(self, v: int) -> List[int]:
        return [self.other(e, v) for e in self.incident[v]]

    def leg(self, v: int, e: int) -> int:
        """Position (0-based, among bond legs) of edge e in the site tensor of v."""
        return self.incident[v].index(e)

    # ---- directed edges ----
    def dedge(self, u: int, v: int) -> int:
        """Directed edge id of u->v."""
        e = self.edge_of[(u, v)]
        return 2 * e + (0 if int(self.esrc[e]) == u else 1)

    def dsrc(self, k: int) -> int:
        e, d = divmod(k, 2)
        return int(self.esrc[e]) if d == 0 else int(self.edst[e])

    def ddst(self, k: int) -> int:
        e, d = divmod(k, 2)
        return int(self.edst[e]) if d == 0 else int(self.esrc[e])

    def v(self, name) -> int:
        return self.index[name]

    def is_tree(self) -> bool:
        return self.ne == self.nv - 1 and len(connected_components(self)) == 1


def _from_named_edges(names: Sequence[Hashable], named_edges: Iterable[Tuple[Hashable, Hashable]]) -> Graph:
    idx = {n: i for i, n in enumerate(names)}
    es = [(idx[a], idx[b]) for a, b in named_edges]
    return Graph(list(names), [a for a, _ in es], [b for _, b in es])


def named_grid(dims: Sequence[int], periodic: bool = False) -> Graph:
    """Hypercubic grid with 1-based tuple names, first coordinate fastest (NamedGraphs `named_grid`)."""
    dims = tuple(int(x) for x in dims)
    nd = len(dims)
    names = []
    for lin in range(int(np.prod(dims))):
        c, r = [], lin
        for n in dims:
            c.append(r % n + 1)
            r //= n
        names.append(tuple(c))
    edges = []
    for name in names:
        for ax in range(nd):
            if name[ax] < dims[ax]:
                nb = list(name); nb[ax] += 1
                edges.append((name, tuple(nb)))
            elif periodic and dims[ax] > 2:
                nb = list(name); nb[ax] = 1
                edges.append((name, tuple(nb)))
    return _from_named_edges(names, edges)


def named_comb_tree(dims: Tuple[int, int]) -> Graph:
    """Comb tree: a backbone path of dims[0] vertices (i,1), each with a tooth (i,2..dims[1])."""
    nx, ny = dims
    names = [(i, j) for j in range(1, ny + 1) for i in range(1, nx + 1)]
    edges = [((i, 1), (i + 1, 1)) for i in range(1, nx)]
    edges += [((i, j), (i, j + 1)) for i in range(1, nx + 1) for j in range(1, ny)]
    return _from_named_edges(names, edges)


def path_graph(n: int) -> Graph:
    return _from_named_edges([(i,) for i in range(1, n + 1)], [((i,), (i + 1,)) for i in range(1, n)])


def named_hexagonal_lattice_graph(m: int, n: int) -> Graph:
    """Honeycomb patch of m x n hexagons (same construction as networkx `hexagonal_lattice_graph`)."""
    M = 2 * m + 1
    rows, cols = range(M + 1), range(n + 1)
    col_edges = [((i, j), (i, j + 1)) for i in cols for j in rows[:M]]
    row_edges = [((i, j), (i + 1, j)) for i in cols[:n] for j in rows if i % 2 == j % 2]
    nodes = {(i, j) for i in cols for j in rows}
    nodes.discard((0, M))
    nodes.discard((n, M * (n % 2)))
    names = sorted(nodes)
    edges = [(a, b) for a, b in col_edges + row_edges if a in nodes and b in nodes]
    names = [(a + 1, b + 1) for a, b in names]
    edges = [((a[0] + 1, a[1] + 1), (b[0] + 1, b[1] + 1)) for a, b in edges]
    return _from_named_edges(names, edges)


def heavy_hexagonal_lattice(nx: int, ny: int) -> Graph:
    """Hexagonal lattice with one extra vertex on every edge (/root/reference/src/graph_ops.jl:6-18)."""
    g = named_hexagonal_lattice_graph(nx, ny)
    names = [(2 * a - 1, 2 * b - 1) for a, b in g.names]
    edges = []
    for e in range(g.ne):
        a, b = names[int(g.esrc[e])], names[int(g.edst[e])]
        mid = ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2)
        names.append(mid)
        edges += [(a, mid), (mid, b)]
    return _from_named_edges(names, edges)


def eagle_heavy_hex() -> Graph:
    """IBM Eagle 127-qubit heavy-hex coupling map: 7 rows (14,15,15,15,15,15,14 qubits) joined by
    6 x 4 bridge qubits; V=127, E=144, max degree 3."""
    row_len = [14, 15, 15, 15, 15, 15, 14]
    names, edges = [], []
    for r, n in enumerate(row_len):
        off = 0 if r < 6 else 1               # last row is shifted by one column
        cols = list(range(off, off + n))
        for c in cols:
            names.append(("q", r, c))
        edges += [(("q", r, c), ("q", r, c + 1)) for c in cols[:-1]]
    for r in range(6):
        cols = [0, 4, 8, 12] if r % 2 == 0 else [2, 6, 10, 14]
        for c in cols:
            b = ("b", r, c)
            names.append(b)
            edges += [(("q", r, c), b), (b, ("q", r + 1, c))]
    g = _from_named_edges(names, edges)
    assert g.nv == 127 and g.ne == 144
    return g


def build_graph_from_gates(circuit: Sequence[tuple]) -> Graph:
    """Graph induced by a circuit of (name, vertices[, param]) tuples (/root/reference/src/graph_ops.jl:53-67)."""
    names: List[Hashable] = []
    for gate in circuit:
        for v in _as_vertex_list(gate[1]):
            if v not in names:
                names.append(v)
    edges = []
    for gate in circuit:
        vs = _as_vertex_list(gate[1])
        if len(vs) == 2 and (vs[0], vs[1]) not in edges and (vs[1], vs[0]) not in edges:
            edges.append((vs[0], vs[1]))
    g = _from_named_edges(names, edges)
    if len(connected_components(g)) != 1:
        raise ValueError("The circuit graph is not connected")
    return g


build_graph_from_circuit = build_graph_from_gates


def _as_vertex_list(vs) -> list:
    if isinstance(vs, list):
        return vs
    return [vs]


# ---------------------------------------------------------------------------------------------
# connected components / spanning forests / edge sequences
# ---------------------------------------------------------------------------------------------
def connected_components(g: Graph, allowed_edges=None) -> List[List[int]]:
    seen = [False] * g.nv
    comps = []
    for s in range(g.nv):
        if seen[s]:
            continue
        comp, stack = [], [s]
        seen[s] = True
        while stack:
            u = stack.pop()
            comp.append(u)
            for e in g.incident[u]:
                if allowed_edges is not None and e not in allowed_edges:
                    continue
                w = g.other(e, u)
                if not seen[w]:
                    seen[w] = True
                    stack.append(w)
        comps.append(sorted(comp))
    return comps


def _bfs_spanning_forest(g: Graph, remaining: set) -> List[Tuple[int, Dict[int, List[int]]]]:
    """BFS spanning tree of every connected component of the subgraph made of `remaining` edges.
    Returns [(root, children_map)] for components with at least one edge."""
    seen = [False] * g.nv
    out = []
    for root in range(g.nv):
        if seen[root] or not any(e in remaining for e in g.incident[root]):
            continue
        seen[root] = True
        children: Dict[int, List[int]] = {root: []}
        queue = [root]
        while queue:
            u = queue.pop(0)
            for e in g.incident[u]:
                if e not in remaining:
                    continue
                w = g.other(e, u)
                if not seen[w]:
                    seen[w] = True
                    children[u].append(w)
                    children[w] = []
                    queue.append(w)
        out.append((root, children))
    return out


def forest_cover_edge_sequence(g: Graph) -> np.ndarray:
    """Default BP edge order (/root/reference/src/MessagePassing/beliefpropagationcache.jl:74-85):
    for each forest of a forest cover, for each tree, the post-order DFS edges (child->parent) followed
    by the same list reversed and flipped.  Returns directed edge ids (2e+dir), length 2*ne.

    `forest_cover` / `post_order_dfs_edges` live in NamedGraphs (not in /root/reference); this is a
    restatement (repeated BFS spanning forests of the uncovered edges, roots = lowest vertex id).
    Tie-breaking upstream is unpinned, which is why the C-ABI takes the edge sequence as an input."""
    remaining = set(range(g.ne))
    seq: List[int] = []
    while remaining:
        forest = _bfs_spanning_forest(g, remaining)
        covered = set()
        for root, children in forest:
            tree_edges: List[int] = []
            # iterative post-order DFS emitting child->parent directed edges
            stack = [(root, -1, 0)]
            while stack:
                u, parent, i = stack.pop()
                ch = children[u]
                if i < len(ch):
                    stack.append((u, parent, i + 1))
                    stack.append((ch[i], u, 0))
                elif parent >= 0:
                    tree_edges.append(g.dedge(u, parent))
                    covered.add(g.edge_of[(u, parent)])
            seq += tree_edges + [k ^ 1 for k in reversed(tree_edges)]
        remaining -= covered
    return np.asarray(seq, dtype=np.int32)


# ---------------------------------------------------------------------------------------------
# edge colouring
# ---------------------------------------------------------------------------------------------
def steiner_tree(g: Graph, terminals: Sequence[int]) -> Tuple[List[int], List[int]]:
    """A tree of g that spans `terminals`: (vertices, undirected edge ids).  Greedy shortest-path growth (start from the first
    terminal, repeatedly join the terminal nearest to the tree by a BFS shortest path, lowest ids first) - exact for two
    terminals and on trees.  The reference calls `steiner_tree(network, obs_vs)` of Graphs.jl (src/expect.jl:66, src/rdm.jl:55),
    which is not in /root/reference; on a loopy graph equally short trees give different BP regions, so the tree is an input of
    the measurement functions (`tree=`) and this is only the default."""
    terms = [int(t) for t in terminals]
    if len(set(terms)) != len(terms):
        raise ValueError("Repeated vertex in collection")
    in_tree = {terms[0]}
    edges: List[int] = []
    todo = [t for t in terms[1:]]
    while todo:
        prev = {v: -1 for v in sorted(in_tree)}              # multi-source BFS from the tree
        queue = sorted(in_tree)
        hit = None
        while queue and hit is None:
            u = queue.pop(0)
            for w in g.neighbors(u):
                if w not in prev:
                    prev[w] = u
                    queue.append(w)
                    if w in todo:
                        hit = w
                        break
        if hit is None:
            raise ValueError("vertices are not connected")
        w = hit
        while prev[w] != -1:
            edges.append(int(g.edge_of[(prev[w], w)]))
            in_tree.add(w)
            w = prev[w]
        todo = [t for t in todo if t not in in_tree]
    return sorted(in_tree), sorted(edges)


def tree_schedule(g: Graph, vertices: Sequence[int], edges: Sequence[int], root: int) -> List[int]:
    """Directed edge ids child -> parent of the tree (vertices, edges), children before parents (post-order towards `root`);
    raises if the vertex set has an edge of g that is not in the tree (the induced region would have a loop)."""
    vs, es = set(int(v) for v in vertices), set(int(e) for e in edges)
    for e in range(g.ne):
        if e not in es and int(g.esrc[e]) in vs and int(g.edst[e]) in vs:
            raise NotImplementedError("the region of this observable has a loop (an edge between two tree vertices that is not in the tree)")
    adj: Dict[int, List[int]] = {v: [] for v in vs}
    for e in es:
        a, b = int(g.esrc[e]), int(g.edst[e])
        adj[a].append(b); adj[b].append(a)
    order, parent, stack = [], {root: -1}, [root]
    while stack:
        u = stack.pop()
        order.append(u)
        for w in sorted(adj[u]):
            if w not in parent:
                parent[w] = u
                stack.append(w)
    if len(order) != len(vs) or len(es) != len(vs) - 1:
        raise ValueError("not a tree")
    return [g.dedge(u, parent[u]) for u in reversed(order) if parent[u] != -1]


def grid_edge_color(g: Graph) -> List[List[int]]:
    """Deterministic closed-form colouring of an open hypercubic grid: per axis, edges starting at an odd
    coordinate, then at an even one (2*ndim colours).  Stands in for the solver-dependent
    `edge_color(g, k)` of SimpleGraphAlgorithms.  Returns groups of undirected edge ids."""
    nd = len(g.names[0])
    groups: List[List[int]] = [[] for _ in range(2 * nd)]
    for e in range(g.ne):
        a, b = g.names[int(g.esrc[e])], g.names[int(g.edst[e])]
        ax = next(i for i in range(nd) if a[i] != b[i])
        lo = min(a[ax], b[ax])
        wrap = abs(a[ax] - b[ax]) != 1
        if wrap:
            raise ValueError("grid_edge_color handles open grids only; use edge_color")
        groups[2 * ax + (0 if lo % 2 == 1 else 1)].append(e)
    return [grp for grp in groups if grp]


def edge_color(g: Graph, k: int | None = None) -> List[List[int]]:
    """Proper edge colouring with max-degree colours for bipartite graphs (alternating-path
    algorithm), greedy otherwise.  Deterministic.  Returns groups of undirected edge ids."""
    delta = max(g.degree(v) for v in range(g.nv))
    ncol = max(delta, k or 0)
    col_at: List[Dict[int, int]] = [dict() for _ in range(g.nv)]   # v -> {colour: edge}
    colour = [-1] * g.ne
    bip = _bipartition(g)
    for e in range(g.ne):
        u, v = int(g.esrc[e]), int(g.edst[e])
        fu = next(c for c in range(ncol + 1) if c not in col_at[u])
        fv = next(c for c in range(ncol + 1) if c not in col_at[v])
        if bip is not None and fu != fv and (fu >= ncol or fv >= ncol or True):
            # flip the fu/fv alternating path that starts at v with colour fu
            if fu in col_at[v]:
                path, x, c = [], v, fu
                while c in col_at[x]:
                    pe = col_at[x][c]
                    path.append(pe)
                    x = g.other(pe, x)
                    c = fv if c == fu else fu
                for pe in path:
                    a, b = int(g.esrc[pe]), int(g.edst[pe])
                    del col_at[a][colour[pe]]
                    del col_at[b][colour[pe]]
                for pe in path:
                    a, b = int(g.esrc[pe]), int(g.edst[pe])
                    colour[pe] = fv if colour[pe] == fu else fu
                    col_at[a][colour[pe]] = pe
                    col_at[b][colour[pe]] = pe
            c = fu
        else:
            c = next(c for c in range(2 * ncol + 1) if c not in col_at[u] and c not in col_at[v])
        colour[e] = c
        col_at[u][c] = e
        col_at[v][c] = e
    ncols = max(colour) + 1
    groups = [[e for e in range(g.ne) if colour[e] == c] for c in range(ncols)]
    for grp in groups:
        vs = [int(g.esrc[e]) for e in grp] + [int(g.edst[e]) for e in grp]
        assert len(vs) == len(set(vs)), "edge colouring is not proper"
    return [grp for grp in groups if grp]


def _bipartition(g: Graph):
    side = [-1] * g.nv
    for s in range(g.nv):
        if side[s] >= 0:
            continue
        side[s] = 0
        stack = [s]
        while stack:
            u = stack.pop()
            for w in g.neighbors(u):
                if side[w] < 0:
                    side[w] = 1 - side[u]
                    stack.append(w)
                elif side[w] == side[u]:
                    return None
    return side


# ---------------------------------------------------------------------------------------------
# generalised loops (loop corrections)
# ---------------------------------------------------------------------------------------------
def simple_cycles_limited_length(g: Graph, max_length: int) -> List[Tuple[int, ...]]:
    """Every simple cycle of g with at most `max_length` edges, each once, as the sorted tuple of its edge ids
    (NamedGraphs.unique_simplecycles_limited_length, used by src/MessagePassing/beliefpropagationcache.jl:185-187).
    Depth-limited DFS from the smallest vertex of the cycle."""
    out = set()
    for root in range(g.nv):
        stack = [(root, (), (root,))]
        while stack:
            v, es, vs = stack.pop()
            for e in g.incident[v]:
                u = g.other(e, v)
                if u == root and len(es) >= 2 and e not in es:
                    out.add(tuple(sorted(es + (e,))))
                elif u > root and u not in vs and len(es) + 1 < max_length:
                    stack.append((u, es + (e,), vs + (u,)))
    return sorted(out, key=lambda c: (len(c), c))


def edgeinduced_subgraphs_no_leaves(g: Graph, max_number_of_edges: int) -> List[Tuple[int, ...]]:
    """Every non-empty set of at most `max_number_of_edges` edges whose edge-induced subgraph has no vertex of degree one - the
    "generalised loops" of the loop series (NamedGraphs.edgeinduced_subgraphs_no_leaves as called by
    src/MessagePassing/loopcorrection.jl:11-12; upstream code is not vendored: this is the definition its name and its use state,
    and the one that makes the series exact when every such subgraph is included - tests/test_loopcorrections.py).  Disconnected
    subgraphs (vertex-disjoint unions of smaller ones) are part of it once the size allows two loops.  Sorted tuples of edge ids,
    ordered by size, then lexicographically.

    Connected ones: ESU enumeration of the connected vertex sets of the line graph, anchored at the smallest edge id, pruned when
    the leaves of the partial set cannot all be cured with the edges that are left (one new edge cures at most two)."""
    ne = g.ne
    if max_number_of_edges < 3 or ne == 0:
        return []
    ends = [(int(a), int(b)) for a, b in zip(g.esrc.tolist(), g.edst.tolist())]
    adj = [sorted(set(f for x in ends[e] for f in g.incident[x]) - {e}) for e in range(ne)]
    found: List[Tuple[int, ...]] = []
    deg = [0] * g.nv

    def extend(sub: List[int], subset: set, ext: List[int], nbhood: set, anchor: int, leaves: int):
        if leaves == 0:
            found.append(tuple(sorted(sub)))
        room = max_number_of_edges - len(sub)
        if room == 0 or (leaves + 1) // 2 > room:
            return
        if leaves > room:
            # every leaf needs a new edge at it, and one edge serves two leaves only if it joins them: with p such edges
            # available at most min(p, leaves // 2) edges do double duty
            p = 0
            for e in sub:
                for x in ends[e]:
                    if deg[x] == 1:
                        for f in g.incident[x]:
                            if f not in subset:
                                y = ends[f][0] if ends[f][1] == x else ends[f][1]
                                if deg[y] == 1:
                                    p += 1
            if leaves - min(p // 2, leaves // 2) > room:
                return
        ext = list(ext)
        while ext:
            w = ext.pop()
            # exclusive neighbours of w: beyond the anchor, not in the set, not adjacent to the set already
            new = [u for u in adj[w] if u > anchor and u not in subset and u not in nbhood]
            sub.append(w); subset.add(w)
            lv = leaves
            for x in ends[w]:                       # a vertex of degree 0 becomes a leaf, a leaf stops being one
                lv += 1 if deg[x] == 0 else (-1 if deg[x] == 1 else 0)
                deg[x] += 1
            extend(sub, subset, ext + new, nbhood | set(new), anchor, lv)
            for x in ends[w]:
                deg[x] -= 1
            sub.pop(); subset.discard(w)

    for a in range(ne):
        for x in ends[a]:
            deg[x] += 1
        first = [u for u in adj[a] if u > a]
        extend([a], {a}, first, set(first) | {a}, a, 2)
        for x in ends[a]:
            deg[x] -= 1
    connected = sorted(set(found), key=lambda c: (len(c), c))
    # vertex-disjoint unions
    vsets = [frozenset(x for e in c for x in ends[e]) for c in connected]
    smallest = len(connected[0]) if connected else max_number_of_edges + 1
    out = list(connected)
    if 2 * smallest <= max_number_of_edges:
        def combine(start: int, edges: Tuple[int, ...], verts: frozenset):
            for j in range(start, len(connected)):
                if len(edges) + len(connected[j]) > max_number_of_edges:
                    break                      # `connected` is sorted by size
                if verts & vsets[j]:
                    continue
                union = tuple(sorted(edges + connected[j]))
                out.append(union)
                combine(j + 1, union, verts | vsets[j])
        for i in range(len(connected)):
            combine(i + 1, connected[i], vsets[i])
    return sorted(set(out), key=lambda c: (len(c), c))


def cycle_walks(g: Graph, max_length: int) -> List[Tuple[List[int], int]]:
    """Every simple cycle of at most `max_length` edges as a closed walk of directed edge ids k = 2e + dir: (all steps but the
    last, the last step).  The walk starts at the smallest vertex of the cycle and leaves it along its smaller edge; the second
    element is what loop_correlations hands to loop_correlation as the target (beliefpropagationcache.jl:185-190: the reverse of
    the last edge of the cycle's path, i.e. the step that closes the walk, traversed towards the start)."""
    out = []
    for cyc in simple_cycles_limited_length(g, max_length):
        cs = set(cyc)
        start = min(int(x) for e in cyc for x in (g.esrc[e], g.edst[e]))
        e0 = min(e for e in g.incident[start] if e in cs)
        walk, v, e = [], start, e0
        while True:
            u = g.other(e, v)
            walk.append(g.dedge(v, u))
            if u == start:
                break
            e = next(f for f in g.incident[u] if f in cs and f != e)
            v = u
        out.append((walk[:-1], walk[-1]))
    return out
