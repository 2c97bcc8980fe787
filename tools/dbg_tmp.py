import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates, engine as E
from util import to_engine
chi=3
g=G.named_grid((3,3)); psi=O.random_tensornetworkstate(g,bond_dimension=chi,seed=77)
e=g.edge_of[(g.v((2,2)),g.v((2,3)))]
be=to_engine(psi,chi_cap=16)
try:
    r=E.apply_2site_batch(be,[e],[gates.gate_array(("CNOT",[0,1]))],apply_kwargs=dict(maxdim=chi,cutoff=1e-10))
    print("ok",r[0])
except Exception as ex: print("FAIL",str(ex)[-60:])
