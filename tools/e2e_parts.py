"""Pieces of the bench's e2e leg, timed one by one (host clock, synchronised): set_sites, apply_gates with / without the fetch of all
sites, expect, get_sites."""
import sys, os, time, gc, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10); E.rescale(bpc, inplace=True)
L = lib.load(); h = psi._h.ptr
layer = [(rx, [v]) for v in range(g.nv)] + [(rz, [v]) for v in range(g.nv)]
for grp in groups:
    layer += [(rzz, [int(g.esrc[e]), int(g.edst[e])]) for e in grp]
kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
sizes = [int(L.tnqs_site_size(h, v)) for v in range(g.nv)]
host = [torch.empty(2 * n, dtype=torch.float64).pin_memory() for n in sizes]
ptrs = [C.cast(x.data_ptr(), C.POINTER(C.c_double)) for x in host]
vs_arr = lib.i32(np.arange(g.nv)); ns_arr = np.asarray(sizes, dtype=np.int64); parr = (C.POINTER(C.c_double) * g.nv)(*ptrs)
lib.check(h, L.tnqs_get_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))
obs = g.v((16, 16))
gc.collect(); gc.freeze(); gc.disable()
def clock(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return round(1e3 * (time.perf_counter() - t0), 1)
for rep in range(5):
    a = clock(lambda: lib.check(h, L.tnqs_set_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr))))
    fns = ns_arr.copy()
    b = clock(lambda: E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True, fetch=(vs_arr, parr, fns)))
    c = clock(lambda: E.expect(bpc, ("Z", [obs])))
    d = clock(lambda: E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True))
    e = clock(lambda: E.expect(bpc, ("Z", [obs])))
    f = clock(lambda: lib.check(h, L.tnqs_get_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr))))
    print(f"set_sites {a}  apply+fetch {b}  expect {c}  apply {d}  expect {e}  get_sites {f}", flush=True)
