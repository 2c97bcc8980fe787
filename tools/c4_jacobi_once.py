import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G
g = G.named_grid((8, 8, 8))
psi = E.random_tensornetworkstate(g, bond_dimension=8, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=2, mode="jacobi")
