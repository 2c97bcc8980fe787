"""Time of the loop-weight evaluation (tnqs_loop_weights) for the plaquettes of an n x n grid at bond dimension chi, and of the
first-order series (plaquettes, 2 x 1 rectangles, theta graphs: max_configuration_size = 7) at small chi.
usage: python tools/loop_bench.py [n=32] [chi=8] [size=4]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G, lib as L


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    size = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    g = G.named_grid((n, n))
    psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=3)
    bpc = E.update(E.BeliefPropagationCache(psi, _share=True), inplace=True, maxiter=20, tolerance=1e-10)
    t0 = time.perf_counter()
    egs = G.edgeinduced_subgraphs_no_leaves(g, size)
    t_enum = time.perf_counter() - t0
    res = E.rescale(bpc)
    lib = L.load()
    times = []
    for rep in range(3):
        l0 = lib.tnqs_kernel_launches(res.psi._h.ptr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ws = E.loop_weights(res, egs)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = lib.tnqs_kernel_launches(res.psi._h.ptr) - l0
    # message updates executed: per connected subgraph sum over vertices of prod_{cuts touching the subtree} (chi^2 + 1); for the
    # plaquettes every vertex sees the one cut
    nplaq = sum(1 for eg in egs if len(eg) == 4)
    z = 4
    flops_update = 8.0 * z * 2 * chi ** (z + 1)
    print(f"{n}x{n} chi={chi} size<={size}: {len(egs)} generalised loops ({nplaq} plaquettes), enumeration {t_enum:.2f} s (host)")
    print(f"  loop_weights: {min(times) * 1e3:.1f} ms (best of 3; {[round(t * 1e3, 1) for t in times]}), {launches} launches, "
          f"{len(egs) / min(times):.0f} weights/s")
    if size == 4:
        updates = nplaq * 4 * (chi * chi + 1)
        print(f"  plaquettes: {updates} message updates, {updates * flops_update / min(times) / 1e12:.2f} TFLOP/s on the interior count "
              f"(8 z d chi^(z+1) per update)")
    zbp = E.partitionfunction(bpc)
    print(f"  sum of weights {complex(np.sum(ws)):.6e}; Z_bp {zbp:.6e}")


if __name__ == "__main__":
    main()
