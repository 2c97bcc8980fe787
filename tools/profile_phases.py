"""Per-kernel-class device time (CUDA events per launch, engine profiling) of one BP update and one gate class on C2."""
import sys, os, json, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, lib
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g, groups, rx, rz, rzz, seq = bench.workload(nx, nx)
psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=5); E.rescale(bpc, inplace=True)
L = lib.load(); h = psi._h.ptr
def prof(name, fn, reps=1):
    fn()                                   # warm
    L.tnqs_profile_read(h, None, None, 1); L.tnqs_profile_enable(h, 1)
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    L.tnqs_sync(h); wall = (time.perf_counter() - t0) / reps
    ms = (C.c_double * 8)(); ln = (C.c_int64 * 8)()
    L.tnqs_profile_read(h, ms, ln, 1); L.tnqs_profile_enable(h, 0)
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    L.tnqs_sync(h); wall2 = (time.perf_counter() - t0) / reps
    print(json.dumps({"phase": name, "wall_ms_profiled": 1e3 * wall, "wall_ms_plain": 1e3 * wall2,
                      "class_ms": {k: round(ms[i] / reps, 3) for i, k in enumerate(lib.KERNEL_CLASSES)},
                      "class_launches": {k: ln[i] // reps for i, k in enumerate(lib.KERNEL_CLASSES)}}))
prof("bp_2_sweeps", lambda: E.update(bpc, edge_sequence=seq, inplace=True, maxiter=2))
prof("bp_2_sweeps_jacobi", lambda: E.update(bpc, edge_sequence=np.arange(2 * g.ne), inplace=True, maxiter=2, mode="jacobi"))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
for grp in groups[:2]:
    prof(f"gate_class_{len(grp)}", lambda: E.apply_2site_batch(bpc, grp, [rzz] * len(grp), apply_kwargs=kw))
vs = list(range(g.nv))
prof("one_site_all", lambda: E.apply_1site_batch(bpc, vs, [rx] * g.nv))
prof("expect_all_Z", lambda: E.expect(bpc, [("Z", [v]) for v in range(g.nv)]))
