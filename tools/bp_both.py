"""Two reference-order sweeps (k_bp_sweep) and two JACOBI sweeps (k_absorb_slab + k_close_dmma) on the C2 state, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=2); E.rescale(bpc, inplace=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=2)
E.update(bpc, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=2, mode="jacobi")
print("ok", psi.kernel_launches())
