"""k_bp_sweep (one persistent launch per sweep) against the level launches (TNQS_SWEEP=0): message agreement and time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g, groups, rx, rz, rzz, seq = bench.workload(nx, nx)
L = lib.load()
def run(mode, sweep, iters=3):
    os.environ["TNQS_SWEEP"] = "1" if sweep else "0"
    psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    h = psi._h.ptr
    s = seq if mode == "reference" else np.arange(2 * g.ne, dtype=np.int32)
    kw = {} if mode == "reference" else {"mode": "jacobi"}
    info = {}
    E.update(bpc, edge_sequence=s, inplace=True, maxiter=iters, tolerance=1e-30, info=info, **kw)
    msgs = [E.message(bpc, k) for k in range(2 * g.ne)]
    ts = []
    for _ in range(3):
        L.tnqs_sync(h); t0 = time.perf_counter()
        E.update(bpc, edge_sequence=s, inplace=True, maxiter=4, **kw)
        L.tnqs_sync(h); ts.append(1e3 * (time.perf_counter() - t0) / 4)
    return msgs, min(ts), info
for mode in ("reference", "jacobi"):
    a, ta, ia = run(mode, False)
    b, tb, ib = run(mode, True)
    worst = max(np.linalg.norm(x - y) / np.linalg.norm(x) for x, y in zip(a, b))
    print(f"{mode}: level launches {ta:.2f} ms/sweep, k_bp_sweep {tb:.2f} ms/sweep, worst relative message difference {worst:.2e}, mean diff {ia.get('mean_diff')} vs {ib.get('mean_diff')}", flush=True)
