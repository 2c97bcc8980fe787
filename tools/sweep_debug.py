import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
nx, chi = 32, 16
g, groups, rx, rz, rzz, seq = bench.workload(nx, nx)
def run(sweep, iters, env=None):
    os.environ["TNQS_SWEEP"] = "1" if sweep else "0"
    for k, v in (env or {}).items(): os.environ[k] = v
    psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    E.update(bpc, edge_sequence=np.arange(2 * g.ne, dtype=np.int32), inplace=True, maxiter=iters, mode="jacobi")
    for k in (env or {}): os.environ.pop(k)
    return [E.message(bpc, k) for k in range(2 * g.ne)]
for iters in (1, 2):
    a = run(False, iters)
    for env in ({}, {"TNQS_SWEEP_RING": "5000"}, {"TNQS_SWEEP_LAG": "1"}, {"TNQS_NO_GRAPH": "1"}):
        b = run(True, iters, env)
        d = np.array([np.linalg.norm(x - y) / np.linalg.norm(x) for x, y in zip(a, b)])
        bad = np.nonzero(d > 1e-12)[0]
        print(f"iters {iters} env {env}: worst {d.max():.2e}, bad {len(bad)} of {len(d)}, first bad {bad[:12].tolist()}, degrees {[g.degree(g.dsrc(int(k))) for k in bad[:12]]}", flush=True)
