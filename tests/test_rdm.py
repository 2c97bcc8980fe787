"""Multi-site BP reduced density matrices (src/rdm.jl:49-70): the oracle's region contraction against the exact
statevector on a tree, and the Pauli-string assembly the engine uses (engine.reduced_density_matrix) against it."""
import itertools

import numpy as np

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates


def exact_rdm(sv, n, vs):
    psi = sv.reshape((2,) * n)
    rest = [i for i in range(n) if i not in vs]
    a = np.transpose(psi, list(vs) + rest).reshape(2 ** len(vs), -1)
    rho = a @ a.conj().T
    return rho / np.trace(rho)


def pauli_assembled(expect, vs):
    n = len(vs)
    rho = np.zeros((2 ** n, 2 ** n), dtype=np.complex128)
    for s in itertools.product("IXYZ", repeat=n):
        p = np.array([[1.0 + 0j]])
        for ch in s:
            p = np.kron(p, gates.PAULI[ch])
        rho += (1.0 if set(s) == {"I"} else expect("".join(s), list(vs))) * p
    return rho / 2 ** n


def test_reduced_density_matrices_on_a_tree_and_by_pauli_assembly():
    g = G.named_comb_tree((3, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=2, seed=17)
    sv = O.to_statevector(psi)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=1)
    for vs in ([4], [0, 8], [3, 1], [0, 4, 8]):
        want = exact_rdm(sv, g.nv, vs)
        got = O.reduced_density_matrix(bpc, vs)
        assert np.max(np.abs(got - want)) < 1e-12
        assert np.max(np.abs(pauli_assembled(lambda s, v: O.expect(bpc, (s, v)), vs) - want)) < 1e-12
    # loopy graph: the assembly from expectation values is the same matrix as the direct contraction of the region
    g = G.named_grid((3, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=18)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=40, tolerance=1e-13)
    for vs in ([0, 8], [0, 2, 7]):
        tree = G.steiner_tree(g, vs) if len(vs) > 2 else None
        direct = O.reduced_density_matrix(bpc, vs, tree=tree)
        assert abs(np.trace(direct) - 1) < 1e-13 and np.max(np.abs(direct - direct.conj().T)) < 1e-7     # Hermitian at the BP fixed point
        assert np.max(np.abs(pauli_assembled(lambda s, v: O.expect(bpc, (s, v), tree=tree), vs) - direct)) < 1e-12
