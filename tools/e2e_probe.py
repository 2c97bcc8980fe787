"""Where the host-API step (apply_gates + expect) spends its time compared with the bench stepper."""
import sys, os, time, gc, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
g, groups, rx, rz, rzz, seq = bench.workload(32, 32)
psi = E.random_tensornetworkstate(g, bond_dimension=16, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10); E.rescale(bpc, inplace=True)
st = bench.Stepper(E, torch, bpc, g, groups, rx, rz, rzz, seq, 16, obs_v=g.v((16, 16)))
L = lib.load(); h = psi._h.ptr
gc.collect(); gc.freeze(); gc.disable()
def t(name, fn, n=4):
    out = []
    for _ in range(n):
        L.tnqs_sync(h); t0 = time.perf_counter(); fn(); L.tnqs_sync(h); out.append(round(1e3 * (time.perf_counter() - t0), 1))
    print(name, out, flush=True)
t("stepper.step", st.step, 6)
layer = [(rx, [v]) for v in range(g.nv)] + [(rz, [v]) for v in range(g.nv)]
for grp in groups:
    layer += [(rzz, [int(g.esrc[e]), int(g.edst[e])]) for e in grp]
kw = dict(maxdim=16, cutoff=1e-10, normalize_tensors=True)
info = {}
t("apply_gates", lambda: E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True, info=info))
t("expect", lambda: E.expect(bpc, ("Z", [st.obs_v])))
t("apply_gates+expect", lambda: (E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True), E.expect(bpc, ("Z", [st.obs_v]))))
t("pack_circuit", lambda: E._pack_circuit(g, layer))
t("default_edge_sequence", lambda: E.default_edge_sequence(g))
sizes = [int(L.tnqs_site_size(h, v)) for v in range(g.nv)]
host = [torch.empty(2 * n, dtype=torch.float64).pin_memory() for n in sizes]
ptrs = [C.cast(x.data_ptr(), C.POINTER(C.c_double)) for x in host]
vs_arr = lib.i32(np.arange(g.nv)); ns_arr = np.asarray(sizes, dtype=np.int64); parr = (C.POINTER(C.c_double) * g.nv)(*ptrs)
t("get_sites", lambda: lib.check(h, L.tnqs_get_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr))))
t("set_sites", lambda: lib.check(h, L.tnqs_set_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr))))
t("apply_gates after pinned alloc", lambda: E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True, info=info))
t("set+apply+expect+get", lambda: (lib.check(h, L.tnqs_set_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr))), E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True), E.expect(bpc, ("Z", [st.obs_v])),
                                   lib.check(h, L.tnqs_get_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
