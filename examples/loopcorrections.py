"""examples/loopcorrections.jl of the reference on the CUDA engine: BP norm, first-order loop-corrected norm and (for these small
lattices) the exact norm of a random chi = 3 state on a line, a hexagonal and a square lattice.  max_configuration_size =
2 * (smallest loop) - 1 keeps every connected generalised loop built from one or two adjacent smallest loops and no disconnected
pair.  Needs a B200 (there is no CPU fallback):  python examples/loopcorrections.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G


def exact_norm(psi):
    """Dense contraction (examples only: 16 qubits at most here)."""
    g, ts = psi.g, psi.tensors()
    labels, nxt = {}, g.nv
    operands = []
    for v in range(g.nv):
        sub = [v]
        for e in g.incident[v]:
            if e not in labels:
                labels[e] = nxt
                nxt += 1
            sub.append(labels[e])
        operands += [ts[v], sub]
    sv = np.einsum(*operands, list(range(g.nv)), optimize="greedy")
    return float(np.sqrt(np.vdot(sv, sv).real))


def main():
    nx, ny, chi = 4, 4, 3
    lattices = [(G.named_grid((nx, 1)), "line", 0), (G.named_hexagonal_lattice_graph(2, 2), "hexagonal", 6),
                (G.named_grid((nx, ny)), "square", 4)]
    for g, name, smallest_loop_size in lattices:
        print(f"Testing for {name} lattice with {g.nv} vertices")
        psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1634)
        psi = E.normalize(psi, alg="bp")
        norm_bp = E.norm(psi, alg="bp")
        size = max(2 * smallest_loop_size - 1, 0)
        norm_lc = E.norm(psi, alg="loopcorrections", max_configuration_size=size)
        print(f"Bp Value for norm is {norm_bp.real:.10f}")
        print(f"1st Order Loop Corrected Value for norm is {norm_lc.real:.10f}   ({len(G.edgeinduced_subgraphs_no_leaves(g, size))} generalised loops)")
        print(f"Exact Value for norm is {exact_norm(psi):.10f}")
        if smallest_loop_size:
            corr = E.loop_correlations(psi, smallest_loop_size)
            print(f"Loop correlations of the {len(corr)} smallest loops: max {max(corr):.4f}")


if __name__ == "__main__":
    main()
