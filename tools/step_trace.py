"""Per-step phase times of the bench stepper (steady-state check)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tensornetworkquantumsimulator_b200 import engine as E
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10); E.rescale(bpc, inplace=True)
st = bench.Stepper(E, torch, bpc, g, groups, rx, rz, rzz, seq, 16, obs_v=g.v((16, 16)))
import ctypes as C, json
from tensornetworkquantumsimulator_b200 import lib
L = lib.load(); h = psi._h.ptr
prof = len(sys.argv) > 1 and sys.argv[1] == "prof"
for i in range(8):
    if prof: L.tnqs_profile_read(h, None, None, 1); L.tnqs_profile_enable(h, 1)
    st.reset(); t0 = time.perf_counter(); st.step(); torch.cuda.synchronize(); w = time.perf_counter() - t0
    print(i, "wall %.1f" % (1e3 * w), "bp %.1f gates %.1f other %.1f sweeps %d" % (st.bp_ms, st.gate_ms, st.other_ms, st.bp_sweeps), "chi", set(psi.bond_dims().tolist()))
    if prof:
        ms = (C.c_double * 8)(); ln = (C.c_int64 * 8)()
        L.tnqs_profile_read(h, ms, ln, 1); L.tnqs_profile_enable(h, 0)
        print("   ", json.dumps({k: [round(ms[j], 1), ln[j]] for j, k in enumerate(lib.KERNEL_CLASSES)}))
# kernel-class profile of one more step (CUDA events per launch; graphs off while profiling)
import ctypes as C, json
from tensornetworkquantumsimulator_b200 import lib
L = lib.load(); h = psi._h.ptr
L.tnqs_profile_read(h, None, None, 1); L.tnqs_profile_enable(h, 1)
st.reset(); t0 = time.perf_counter(); st.step(); torch.cuda.synchronize(); w = time.perf_counter() - t0
ms = (C.c_double * 8)(); ln = (C.c_int64 * 8)()
L.tnqs_profile_read(h, ms, ln, 1); L.tnqs_profile_enable(h, 0)
print("profiled step: wall %.1f bp %.1f gates %.1f other %.1f" % (1e3 * w, st.bp_ms, st.gate_ms, st.other_ms))
print(json.dumps({k: [round(ms[i], 2), ln[i]] for i, k in enumerate(lib.KERNEL_CLASSES)}))
print("schedule:", [(it[0], len(it[1]) if len(it) > 1 else 0) for it in st.sched])
