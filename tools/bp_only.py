"""BP sweeps only on the C2 state (for ncu captures of the sweep kernels): `python tools/bp_only.py [nx] [chi] [mode] [sweeps]`."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mode = sys.argv[3] if len(sys.argv) > 3 else "reference"
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
g, groups, rx, rz, rzz, seq = bench.workload(nx, nx)
psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=3); E.rescale(bpc, inplace=True)
s = seq if mode == "reference" else np.arange(2 * g.ne)
kw = dict(mode="jacobi") if mode != "reference" else {}
E.update(bpc, edge_sequence=s, inplace=True, maxiter=1, **kw)
t0 = time.perf_counter()
E.update(bpc, edge_sequence=s, inplace=True, maxiter=sweeps, **kw)
dt = time.perf_counter() - t0
print(f"{mode}: {1e3 * dt / sweeps:.3f} ms per sweep ({2 * g.ne * sweeps / dt:.0f} msg/s), launches {psi.kernel_launches()}")
