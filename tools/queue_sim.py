"""Host-side model of the k_bp_sweep queue: P CTAs draw items in order, an item blocks its CTA until its dependencies
are done.  Compares queue orders for the C2 sweep (32x32 grid, forest-cover sequence).  CPU only."""
import sys, os, heapq
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tensornetworkquantumsimulator_b200 import graphs as G

def build(nx=32):
    g = G.named_grid((nx, nx))
    seq = [int(k) for k in G.forest_cover_edge_sequence(g)]
    # levels exactly as bp_update_impl
    ne = g.ne
    lastwrite = [-1] * (2 * ne); lastread = [-1] * (2 * ne)
    level = []
    ins = []
    for k in seq:
        v, w = g.dsrc(k), g.ddst(k)
        rs = [g.dedge(g.other(e, v), v) for e in g.incident[v] if g.other(e, v) != w]
        lvl = 0
        for r in rs:
            if lastwrite[r] >= 0: lvl = max(lvl, lastwrite[r] + 1)
        lvl = max(lvl, lastread[k] + 1, lastwrite[k] + 1)
        level.append(lvl); lastwrite[k] = lvl
        for r in rs: lastread[r] = max(lastread[r], lvl)
        ins.append(rs)
    return g, seq, level, ins

def chain_order_level_major(seq, level, ins, sort_by_producer=True):
    nl = max(level) + 1
    order = []
    wpos = {}
    for L in range(nl):
        idx = [i for i in range(len(seq)) if level[i] == L]
        if sort_by_producer:
            idx.sort(key=lambda i: max([wpos.get(r, -1) for r in ins[i]] + [-1]))
        for i in idx:
            wpos[seq[i]] = len(order); order.append(i)
    return order

def deps_for(order, seq, ins, R):
    """dep (RAW + ring) and war lists in chain-position space."""
    pos_of_writer = {}; readers = {}
    dep = []; war = []
    for p, i in enumerate(order):
        d = set()
        if p - R >= 0: d.add(p - R)
        for r in ins[i]:
            if r in pos_of_writer: d.add(pos_of_writer[r])
        k = seq[i]
        w = set(readers.get(k, []))
        if k in pos_of_writer: w.add(pos_of_writer[k])
        readers[k] = []; pos_of_writer[k] = p
        for r in ins[i]: readers.setdefault(r, []).append(p)
        dep.append(sorted(d)); war.append(sorted(w))
    return dep, war

def emit(order, dep, nslab, nsplit, D, small):
    items = []; pending = []; head = 0; closed = [False] * len(order)
    def close(m):
        for j in range(nsplit[m]): items.append(("c", m, j))
        closed[m] = True
    for p in range(len(order)):
        upto = max([q for q in dep[p] if not closed[q]] + [-1])
        while upto >= 0 and head < len(pending) and pending[head] <= upto:
            close(pending[head]); head += 1
        if small[p]:
            items.append(("s", p, 0)); closed[p] = True
        else:
            for y in range(nslab): items.append(("a", p, y))
            pending.append(p)
            while len(pending) - head > D:
                close(pending[head]); head += 1
    while head < len(pending):
        close(pending[head]); head += 1
    return items

def simulate(items, dep, war, nslab, nsplit, P=296, ta=18.0, tc=4.5, ts=3.0, tover=0.0):
    n = len(dep)
    slabs_done_t = [0.0] * n; slabs_cnt = [0] * n
    close_cnt = [0] * n; close_t = [0.0] * n
    done_t = [None] * n
    free = [(0.0, w) for w in range(P)]
    heapq.heapify(free)
    wait = 0.0
    for kind, m, y in items:
        t, w = heapq.heappop(free)
        if kind == "a":
            ready = max([done_t[q] for q in dep[m]] + [0.0])
            st = max(t, ready); wait += st - t
            en = st + ta + tover
            slabs_cnt[m] += 1; slabs_done_t[m] = max(slabs_done_t[m], en)
        elif kind == "s":
            ready = max([done_t[q] for q in dep[m]] + [done_t[q] for q in war[m]] + [0.0])
            st = max(t, ready); wait += st - t
            en = st + ts + tover; done_t[m] = en
        else:
            assert slabs_cnt[m] == nslab
            st = max(t, slabs_done_t[m]); wait += st - t
            en = st + tc + tover
            close_cnt[m] += 1; close_t[m] = max(close_t[m], en)
            if close_cnt[m] == nsplit[m]:
                ready = max([done_t[q] for q in war[m]] + [0.0])
                if ready > close_t[m]: wait += ready - close_t[m]; en = ready + 1.0
                done_t[m] = max(close_t[m], en)
        heapq.heappush(free, (en, w))
    return max(t for t, _ in free), wait / P

def list_schedule(seq, ins, level, R, P=296, ta=18.0, tc=4.5, nslab=32, nsplit_v=8):
    """Greedy list scheduling on the simulated machine; returns chain order (by absorb start) - priority: lowest level first
    among ready chains (ready = producers done), which keeps the wavefront moving while the queue stays topological."""
    n = len(seq)
    writer = {}
    prod = []
    for i in range(n):
        prod.append([writer[r] for r in ins[i] if r in writer])
        writer[seq[i]] = i
    return prod

if __name__ == "__main__":
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    g, seq, level, ins = build(nx)
    n = len(seq)
    print("chains", n, "levels", max(level) + 1)
    deg = [g.degree(g.dsrc(k)) for k in seq]
    for name, sortp in (("level-major", False), ("level-major sorted by producer", True)):
        order = chain_order_level_major(seq, level, ins, sortp)
        small = [deg[i] == 2 for i in order]
        nsplit = [1 if deg[i] == 2 else (8 if deg[i] == 4 else 8) for i in order]
        for R, D in ((48, 12), (48, 4), (96, 12), (96, 24), (400, 12)):
            dep, war = deps_for(order, seq, ins, R)
            items = emit(order, dep, 32, nsplit, D, small)
            mk, wt = simulate(items, dep, war, 32, nsplit)
            work = sum(18.0 if it[0] == "a" else (4.5 if it[0] == "c" else 3.0) for it in items) / 296
            print(f"{name}: R={R} D={D}: makespan {mk/1e3:.2f} ms, ideal {work/1e3:.2f} ms, mean wait per CTA {wt/1e3:.2f} ms")
