"""CPU model: how often could a message update in the reference's sequential order reuse a partial product of its vertex?

In a JACOBI sweep the updates that leave one vertex share their pair / triple products (DESIGN.md sections 4 and 6).  In the reference's
order the messages change between the updates of a vertex, so a product can only be reused while none of its input messages has been
rewritten.  This script replays the forest-cover sequence of an open grid with a version counter per message and a cache of the
latest product per (vertex, set of legs):

    python tools/ro_reuse_sim.py [nx] [ny]

Result on 32 x 32 (and 16 x 16): with the pairing {down, up} / {left, right} - legs (0, 3) and (1, 2) of an interior vertex of
named_grid - every second update of an interior vertex finds the pair product it needs unchanged (1800 of 3600 per sweep); the pairings
(0, 1) / (2, 3) and (0, 2) / (1, 3) never do, and no triple partition of the cubic lattice does.  No policy that may pick any pair does
better than 50 %."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tensornetworkquantumsimulator_b200 import graphs as G


def simulate(g, z, partitions, sweeps=3):
    seq = [int(k) for k in G.forest_cover_edge_sequence(g)]

    def kin(v, l):                                   # the message that enters v on leg l
        k = 2 * int(g.incident[v][l])
        return k if g.ddst(k) == v else k ^ 1

    for part in partitions:
        ver = np.zeros(2 * g.ne, dtype=np.int64)
        cache = {}
        for sw in range(sweeps):
            hits = uses = 0
            for k in seq:
                v = g.dsrc(k)
                if g.degree(v) == z:
                    j = list(g.incident[v]).index(k // 2)
                    for grp in part:
                        if j in grp:
                            continue
                        key, now = (v, grp), tuple(int(ver[kin(v, l)]) for l in grp)
                        uses += 1
                        hits += cache.get(key) == now
                        cache[key] = now
                ver[k] += 1
        print(f"  legs {part}: {hits} of {uses} products reusable in a steady sweep")


if __name__ == "__main__":
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else nx
    print(f"{nx} x {ny} grid, interior vertices (degree 4), pair products:")
    simulate(G.named_grid((nx, ny)), 4, [((0, 1), (2, 3)), ((0, 2), (1, 3)), ((0, 3), (1, 2))])
    print("8 x 8 x 8 cubic lattice, interior vertices (degree 6), triple products:")
    simulate(G.named_grid((8, 8, 8)), 6, [((0, 1, 2), (3, 4, 5)), ((0, 2, 4), (1, 3, 5)), ((0, 1, 3), (2, 4, 5)), ((0, 5, 1), (2, 3, 4))])
