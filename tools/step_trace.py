"""Per-step phase times of the bench stepper (steady-state check)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tensornetworkquantumsimulator_b200 import engine as E
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10); E.rescale(bpc, inplace=True)
st = bench.Stepper(E, torch, bpc, g, groups, rx, rz, rzz, seq, 16, obs_v=g.v((16, 16)))
for i in range(8):
    st.reset(); t0 = time.perf_counter(); st.step(); torch.cuda.synchronize(); w = time.perf_counter() - t0
    print(i, "wall %.1f" % (1e3 * w), "bp %.1f gates %.1f other %.1f sweeps %d" % (st.bp_ms, st.gate_ms, st.other_ms, st.bp_sweeps), "chi", set(psi.bond_dims().tolist()))
