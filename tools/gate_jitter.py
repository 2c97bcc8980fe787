"""Wall vs CUDA-event time of every gate phase of the bench stepper over several steps (host jitter hunt)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10); E.rescale(bpc, inplace=True)
st = bench.Stepper(E, torch, bpc, g, groups, rx, rz, rzz, seq, 16, obs_v=g.v((16, 16)))
orig = E.apply_2site_batch
def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(*a, **k)
    torch.cuda.synchronize(); w = 1e3 * (time.perf_counter() - t0)
    timed.log.append(round(w, 1))
    return r
timed.log = []
E.apply_2site_batch = timed
import gc
if os.environ.get('NOGC') == '1':
    gc.collect(); gc.freeze(); gc.disable()
gc.callbacks.append(lambda phase, info: print('   [gc]', phase, info) if phase == 'stop' and info['generation'] == 2 else None)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    timed.log = []
    st.reset(); t0 = time.perf_counter(); st.step(); torch.cuda.synchronize(); w = time.perf_counter() - t0
    print(i, "wall %.1f" % (1e3 * w), "bp %.1f gates %.1f other %.1f" % (st.bp_ms, st.gate_ms, st.other_ms), "classes", timed.log, flush=True)
