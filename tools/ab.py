"""A/B timing of BP sweeps under environment switches, with an agreement check against the first combination:
    python tools/ab.py [--nx 32] [--chi 16] [--mode reference|jacobi] "" "TNQS_SWEEP_PF=0" "TNQS_SWEEP_LAG=4,TNQS_SWEEP_RING=64"
Each argument is one combination (comma-separated VAR=value; "" = defaults).  Prints ms per sweep (best of 3 x 4 sweeps, host clock
around synchronised calls) and the largest relative difference of any message after 2 sweeps from the same start."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=32); ap.add_argument("--chi", type=int, default=16)
ap.add_argument("--mode", default="reference"); ap.add_argument("combos", nargs="*", default=[""])
a = ap.parse_args()
g, groups, rx, rz, rzz, seq = bench.workload(a.nx, a.nx)
L = lib.load()
s = seq if a.mode == "reference" else np.arange(2 * g.ne, dtype=np.int32)
kw = {} if a.mode == "reference" else {"mode": "jacobi"}
ref = None
probe = list(range(0, 2 * g.ne, max(1, 2 * g.ne // 64)))
for combo in a.combos:
    env = dict(kv.split("=") for kv in combo.split(",") if kv)
    for k, v in env.items(): os.environ[k] = v
    psi = E.random_tensornetworkstate(g, bond_dimension=a.chi, seed=1)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    h = psi._h.ptr
    E.update(bpc, edge_sequence=s, inplace=True, maxiter=2, **kw)
    msgs = [E.message(bpc, k) for k in probe]
    if ref is None: ref = msgs
    worst = max(float(np.linalg.norm(x - y) / np.linalg.norm(y)) for x, y in zip(msgs, ref))
    ts = []
    for _ in range(3):
        L.tnqs_sync(h); t0 = time.perf_counter()
        E.update(bpc, edge_sequence=s, inplace=True, maxiter=4, **kw)
        L.tnqs_sync(h); ts.append(1e3 * (time.perf_counter() - t0) / 4)
    for k in env: os.environ.pop(k)
    print(f"{a.mode} [{combo or 'defaults'}]: {min(ts):.3f} ms/sweep  ({2 * g.ne / min(ts) * 1e3:.0f} msg/s)  max rel diff vs first {worst:.2e}", flush=True)
    del bpc, psi
