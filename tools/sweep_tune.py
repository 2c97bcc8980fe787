"""Timing of k_bp_sweep under queue settings (TNQS_SWEEP_LAG / _RING): python tools/sweep_tune.py [nx] [chi]."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g, groups, rx, rz, rzz, seq = bench.workload(nx, nx)
L = lib.load()
def run(mode, env):
    for k, v in env.items(): os.environ[k] = v
    psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    h = psi._h.ptr
    s = seq if mode == "reference" else np.arange(2 * g.ne, dtype=np.int32)
    kw = {} if mode == "reference" else {"mode": "jacobi"}
    E.update(bpc, edge_sequence=s, inplace=True, maxiter=2, **kw)
    ts = []
    for _ in range(3):
        L.tnqs_sync(h); t0 = time.perf_counter()
        E.update(bpc, edge_sequence=s, inplace=True, maxiter=4, **kw)
        L.tnqs_sync(h); ts.append(1e3 * (time.perf_counter() - t0) / 4)
    for k in env: os.environ.pop(k)
    return min(ts)
combos = [{"TNQS_SWEEP": "0"}, {}, {"TNQS_SWEEP_LAG": "2"}, {"TNQS_SWEEP_LAG": "4"}, {"TNQS_SWEEP_LAG": "6"}, {"TNQS_SWEEP_LAG": "8"}, {"TNQS_SWEEP_LAG": "6", "TNQS_SWEEP_RING": "32"}, {"TNQS_SWEEP_LAG": "6", "TNQS_SWEEP_RING": "96"}]
os.environ["TNQS_SWEEP_JACOBI"] = "1"
for mode in ("reference", "jacobi"):
    for env in combos:
        sys.stderr.write(f"--- {mode} {env}\n"); sys.stderr.flush()
        print(f"{mode} {env}: {run(mode, env):.2f} ms/sweep", flush=True)
