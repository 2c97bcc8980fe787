import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G, gates
g = G.named_grid((8, 8, 8))
psi = E.random_tensornetworkstate(g, bond_dimension=8, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=2, mode="jacobi")
grp = list(G.grid_edge_color(g)[0])
rzz = gates.gate_array(("Rzz", [0, 1], 0.25))
E.apply_2site_batch(bpc, grp, [rzz] * len(grp), apply_kwargs=dict(maxdim=8, cutoff=1e-10, normalize_tensors=True))
