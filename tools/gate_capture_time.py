"""Wall time of the first five applications of a colour class on C2: eager, capture + first replay, replays (TNQS_DEBUG_QUEUE=1 prints the staged bytes)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=3); E.rescale(bpc, inplace=True)
L = lib.load(); h = psi._h.ptr
kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
for grp in groups[:2]:
    ts = []
    for i in range(5):
        L.tnqs_sync(h); t0 = time.perf_counter()
        E.apply_2site_batch(bpc, grp, [rzz] * len(grp), apply_kwargs=kw)
        L.tnqs_sync(h); ts.append(round(1e3 * (time.perf_counter() - t0), 2))
    print(len(grp), ts)
