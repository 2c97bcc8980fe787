"""One warm-up + one measured class of two-site gates on C2 (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tensornetworkquantumsimulator_b200 import engine as E
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g, groups, rx, rz, rzz, seq = bench.workload(nx, nx)
psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=3); E.rescale(bpc, inplace=True)
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
for grp in groups[:2]:
    E.apply_2site_batch(bpc, grp, [rzz] * len(grp), apply_kwargs=kw)
print("done", psi.kernel_launches())
