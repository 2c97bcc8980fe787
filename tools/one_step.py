"""Exactly one warm-up layer + one measured layer of the C2 bench stepper (for ncu launch lists: the measured layer is
everything after the marker launch of k_fill_random into a 1-element scratch... simpler: the LAST `nstep` launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tensornetworkquantumsimulator_b200 import engine as E
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10); E.rescale(bpc, inplace=True)
st = bench.Stepper(E, torch, bpc, g, groups, rx, rz, rzz, seq, 16, obs_v=g.v((16, 16)))
st.step()
n0 = psi.kernel_launches()
st.step()
print("launches_in_measured_step", psi.kernel_launches() - n0)
