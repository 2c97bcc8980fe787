"""Per-kernel-class device time of JACOBI sweeps on the C4 lattice (8x8x8 cubic, chi = 8) for the library selected by TNQS_LIB_PATH:
    TNQS_LIB_PATH=variants/lib_x.so python tools/c4_variant.py [label]"""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G, lib
g = G.named_grid((8, 8, 8))
psi = E.random_tensornetworkstate(g, bond_dimension=8, seed=1)
bpc = E.BeliefPropagationCache(psi, _share=True)
L = lib.load(); h = psi._h.ptr
seq = np.arange(2 * g.ne)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=2, mode="jacobi")
L.tnqs_sync(h); L.tnqs_profile_read(h, None, None, 1); L.tnqs_profile_enable(h, 1)
E.update(bpc, edge_sequence=seq, inplace=True, maxiter=4, mode="jacobi"); L.tnqs_sync(h)
ms = (C.c_double * 8)(); ln = (C.c_int64 * 8)()
L.tnqs_profile_read(h, ms, ln, 1); L.tnqs_profile_enable(h, 0)
print(json.dumps({"label": sys.argv[1] if len(sys.argv) > 1 else os.environ.get("TNQS_LIB_PATH", "default"),
                  "ms_per_sweep": {k: round(ms[i] / 4, 3) for i, k in enumerate(lib.KERNEL_CLASSES) if ln[i]}, "total": round(sum(ms) / 4, 3)}))
