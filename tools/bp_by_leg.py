"""JACOBI sweeps over the messages of C2 grouped by the position of their closing leg in the site tensor (0 = fastest bond leg):
how much the strided slab patterns of the low legs cost."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
L = lib.load(); h = psi._h.ptr
by_leg = {0: [], 1: [], 2: [], 3: []}
for k in range(2 * g.ne):
    v = g.dsrc(k)
    if g.degree(v) == 4:
        by_leg[g.leg(v, k // 2)].append(k)
for leg, ks in by_leg.items():
    s = np.asarray(ks, dtype=np.int32)
    E.update(bpc, edge_sequence=s, inplace=True, maxiter=2, mode="jacobi")
    ts = []
    for _ in range(3):
        L.tnqs_sync(h); t0 = time.perf_counter()
        E.update(bpc, edge_sequence=s, inplace=True, maxiter=4, mode="jacobi")
        L.tnqs_sync(h); ts.append(1e3 * (time.perf_counter() - t0) / 4)
    print(f"closing leg {leg}: {len(ks)} messages, {min(ts):.3f} ms per sweep = {1e3 * min(ts) / len(ks):.3f} us per message", flush=True)
