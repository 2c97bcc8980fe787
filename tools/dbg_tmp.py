import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates, engine as E
from util import to_engine
g = G.named_grid((4, 4)); groups = G.grid_edge_color(g)
psi_o = O.random_tensornetworkstate(g, bond_dimension=4, seed=9)
seq = G.forest_cover_edge_sequence(g)
kw = dict(maxdim=4, cutoff=1e-10, normalize_tensors=True); bp_kw = dict(maxiter=25, tolerance=1e-8)
rzz = gates.gate_array(("Rzz", [0, 1], 0.25)); rx = gates.gate_array(("Rx", [0], 0.5)); rz = gates.gate_array(("Rz", [0], 0.4))
a0 = E.update(to_engine(psi_o, chi_cap=4), edge_sequence=seq, **bp_kw)
def diff(x, y, what):
    d = max(np.max(np.abs(x.psi[v] - y.psi[v])) for v in range(g.nv)); m = max(np.max(np.abs(E.message(x,k)-E.message(y,k))) for k in range(2*g.ne))
    print(what, "site diff", d, "msg diff", m)
vs = list(range(g.nv))
one = [(rx, [v]) for v in vs] + [(rz, [v]) for v in vs]
a, b = a0.copy(), a0.copy()
E.apply_gates(one, a, apply_kwargs=kw, bp_update_kwargs=bp_kw, inplace=True, update_cache=False)
E.apply_1site_batch(b, vs, [rx]*g.nv); E.apply_1site_batch(b, vs, [rz]*g.nv); diff(a, b, "1-site only, no cache update")
a, b = a0.copy(), a0.copy()
info={}
E.apply_gates(one, a, apply_kwargs=kw, bp_update_kwargs=bp_kw, inplace=True, info=info)
E.apply_1site_batch(b, vs, [rx]*g.nv); E.apply_1site_batch(b, vs, [rz]*g.nv); E.update(b, edge_sequence=seq, inplace=True, **bp_kw); diff(a, b, "1-site + final update (refreshes %d)" % info["refreshes"])
for nc in (3, 4):
    a, b = a0.copy(), a0.copy()
    layer = list(one)
    for grp in groups[:nc]: layer += [(rzz, [int(g.esrc[e]), int(g.edst[e])]) for e in grp]
    E.apply_gates(layer, a, apply_kwargs=kw, bp_update_kwargs=bp_kw, inplace=True, info=info)
    E.apply_1site_batch(b, vs, [rx]*g.nv); E.apply_1site_batch(b, vs, [rz]*g.nv)
    for grp in groups[:nc]:
        E.update(b, edge_sequence=seq, inplace=True, **bp_kw); E.apply_2site_batch(b, grp, [rzz]*len(grp), apply_kwargs=kw)
    E.update(b, edge_sequence=seq, inplace=True, **bp_kw)
    diff(a, b, "%d classes (refreshes %d)" % (nc, info["refreshes"]))
