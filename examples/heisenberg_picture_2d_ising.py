"""examples/2dIsing_dynamics_Heisenbergpicture.jl of the reference on the CUDA engine: Z at the centre of a 4x4 grid evolved in the
Heisenberg picture as a tensor network over Pauli sites (d = 4, basis I, X, Y, Z), gates = Pauli transfer matrices, traces through
BP on the bilinear form.  Needs a B200 (there is no CPU fallback):  python examples/heisenberg_picture_2d_ising.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G


def main():
    nx, ny = 4, 4
    g = G.named_grid((nx, ny))
    vz = (2, 2)
    maxdim, cutoff = 4, 1.0e-14
    psi0 = E.paulitensornetworkstate(lambda v: "Z" if v == vz else "I", g, chi_cap=maxdim)
    apply_kwargs = dict(maxdim=maxdim, cutoff=cutoff, normalize_tensors=False)
    bpc = E.BeliefPropagationCache(psi0.copy(), _share=True)
    h, J, dt, steps = -1.0, -1.0, 0.04, 10
    layer = [("Rz", [v], h * dt) for v in g.names]
    for grp in G.edge_color(g, 4):
        layer += [("Rxx", [g.names[g.esrc[e]], g.names[g.edst[e]]], 2 * J * dt) for e in grp]
    layer += [("Rz", [v], h * dt) for v in g.names]
    print(f"Initial bond dimension of the Heisenberg operator is {E.maxvirtualdim(bpc)}")
    ident = E.identitytensornetworkstate(g)
    for l in range(1, steps + 1):
        bpc, errors = E.apply_gates(layer, bpc, apply_kwargs=apply_kwargs)
        bpc = E.rescale(bpc)                                   # Frobenius norm of O(t) back to one
        print(f"Layer {l}: Frobenius norm {E.partitionfunction(bpc).real:.12f}, max bond dimension {E.maxvirtualdim(bpc)}, "
              f"max gate error {errors.max():.2e}")
        psi = E.network(bpc)
        # Trace(O(t)) is exactly zero for a traceless O(0): BP then divides numerically vanishing scalars, so the printed number is
        # rounding noise (the reference prints the same kind of noise)
        print(f"  Trace(O(t) O(0)) = {E.inner(psi, psi0):.10f}   Trace(O(t)) = {E.inner(psi, ident):.3e}")


if __name__ == "__main__":
    main()
