"""Host-side overhead of the BP update call on C2: cached plan vs rebuilt plan (after a gate batch invalidates it)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=3); E.rescale(bpc, inplace=True)
L = lib.load(); h = psi._h.ptr
def t(fn, n=3):
    out = []
    for _ in range(n):
        L.tnqs_sync(h); t0 = time.perf_counter(); fn(); L.tnqs_sync(h); out.append(1e3 * (time.perf_counter() - t0))
    return [round(x, 2) for x in out]
kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
print("update 1 sweep, cached plan      :", t(lambda: E.update(bpc, edge_sequence=seq, inplace=True, maxiter=1)))
print("update 2 sweeps, cached plan     :", t(lambda: E.update(bpc, edge_sequence=seq, inplace=True, maxiter=2)))
def inval():
    E.apply_2site_batch(bpc, groups[0][:1], [rzz], apply_kwargs=kw)
for i in range(3):
    inval(); print("update 1 sweep after invalidation:", t(lambda: E.update(bpc, edge_sequence=seq, inplace=True, maxiter=1), 1))
print("gate class                        :", t(lambda: E.apply_2site_batch(bpc, groups[1], [rzz] * len(groups[1]), apply_kwargs=kw)))
vs = list(range(g.nv))
print("1-site all                        :", t(lambda: E.apply_1site_batch(bpc, vs, [rx] * g.nv)))
print("expect one                        :", t(lambda: E.expect(bpc, ("Z", [g.v((16, 16))]))))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); E.apply_1site_batch(bpc, vs, [rx] * g.nv); E.expect(bpc, ("Z", [g.v((16, 16))])); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
