"""ctypes binding of libtnqs_cuda.so (the C-ABI declared in include/tnqs_cuda.h).

This is the binding a Julia maintainer would write with `ccall` (INTEGRATION.md shows that stub); Python
plays the host language here because Julia is not installed in the build image.  There is no CPU
fallback: if the shared library is missing the import fails loudly, and creating a handle without a
CUDA device raises `TnqsError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TNQS_LIB_PATH") or os.path.join(_HERE, "libtnqs_cuda.so")      # the override selects another BUILD of the same
                                                                                              # library (kernel experiments, tools/ab.py)

# every symbol include/tnqs_cuda.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "tnqs_version", "tnqs_status_string", "tnqs_last_error", "tnqs_create", "tnqs_destroy", "tnqs_clone",
    "tnqs_sync", "tnqs_num_vertices", "tnqs_num_edges", "tnqs_get_bond_dims", "tnqs_site_size", "tnqs_set_site",
    "tnqs_get_site", "tnqs_set_sites", "tnqs_get_sites", "tnqs_fill_random", "tnqs_set_message", "tnqs_get_message", "tnqs_reset_messages",
    "tnqs_bp_update", "tnqs_apply_1site", "tnqs_apply_2site", "tnqs_apply_2site_begin", "tnqs_apply_2site_end", "tnqs_apply_circuit", "tnqs_apply_circuit_fetch", "tnqs_rdm_1site",
    "tnqs_expect_2site_nn", "tnqs_expect_path", "tnqs_loop_weights", "tnqs_loop_transfer", "tnqs_vertex_scalars", "tnqs_edge_scalars", "tnqs_rescale", "tnqs_symmetric_gauge",
    "tnqs_truncate", "tnqs_message_slab", "tnqs_halo_pack", "tnqs_halo_unpack", "tnqs_set_partition", "tnqs_set_bra", "tnqs_halo_alloc", "tnqs_halo_peer", "tnqs_halo_route", "tnqs_halo_close",
    "tnqs_kernel_launches", "tnqs_bp_sweeps_total", "tnqs_bp_message_updates_total", "tnqs_profile_enable", "tnqs_profile_read", "tnqs_stream",
]

KERNEL_CLASSES = ["mode_apply", "mode_close", "bp_finalize", "permute", "eigh", "qr", "svd", "misc"]

TNQS_BP_REFERENCE_ORDER, TNQS_BP_JACOBI = 0, 1


class BpOpts(C.Structure):
    _fields_ = [("maxiter", C.c_int32), ("tolerance", C.c_double), ("normalize", C.c_int32),
                ("hermitize", C.c_int32), ("mode", C.c_int32)]


class ApplyOpts(C.Structure):
    _fields_ = [("maxdim", C.c_int32), ("cutoff", C.c_double), ("normalize_tensors", C.c_int32)]


class TnqsError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libtnqs_cuda status {status}: {message}")
        self.status = status


_lib = None


def load() -> C.CDLL:
    """Load libtnqs_cuda.so (built in-tree by `make -C tensornetworkquantumsimulator_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                          "(make -C tensornetworkquantumsimulator_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    p, i32, i64, f64p, i32p, i64p, vp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p
    lib.tnqs_version.restype = C.c_char_p
    lib.tnqs_status_string.restype = C.c_char_p; lib.tnqs_status_string.argtypes = [i32]
    lib.tnqs_last_error.restype = C.c_char_p; lib.tnqs_last_error.argtypes = [p]
    lib.tnqs_create.argtypes = [C.POINTER(p), i32, i32, i32p, i32p, i32p, i32p, i32, i32]
    lib.tnqs_destroy.argtypes = [p]
    lib.tnqs_clone.argtypes = [p, C.POINTER(p)]
    lib.tnqs_sync.argtypes = [p]
    lib.tnqs_num_vertices.argtypes = [p]; lib.tnqs_num_edges.argtypes = [p]
    lib.tnqs_get_bond_dims.argtypes = [p, i32p]
    lib.tnqs_site_size.argtypes = [p, i32]; lib.tnqs_site_size.restype = i64
    lib.tnqs_set_site.argtypes = [p, i32, f64p, i64]; lib.tnqs_get_site.argtypes = [p, i32, f64p, i64]
    lib.tnqs_set_sites.argtypes = [p, i32, i32p, C.POINTER(f64p), i64p]; lib.tnqs_get_sites.argtypes = [p, i32, i32p, C.POINTER(f64p), i64p]
    lib.tnqs_fill_random.argtypes = [p, C.c_uint64]
    lib.tnqs_set_message.argtypes = [p, i32, f64p, i64]; lib.tnqs_get_message.argtypes = [p, i32, f64p, i64]
    lib.tnqs_reset_messages.argtypes = [p]
    lib.tnqs_bp_update.argtypes = [p, i32p, i32, C.POINTER(BpOpts), i32p, f64p]
    lib.tnqs_apply_1site.argtypes = [p, i32, i32p, f64p, i32]
    lib.tnqs_apply_2site.argtypes = [p, i32, i32p, i32p, f64p, C.POINTER(ApplyOpts), i32p, f64p, f64p]
    lib.tnqs_apply_2site_begin.argtypes = [p, i32, i32p, i32p, f64p, C.POINTER(ApplyOpts), i32p, C.POINTER(vp), i64p, i64p]
    lib.tnqs_apply_2site_end.argtypes = [p, vp, i32p, f64p, f64p]
    lib.tnqs_apply_circuit.argtypes = [p, i32, i32p, i64p, f64p, C.POINTER(ApplyOpts), i32p, i32, C.POINTER(BpOpts), i32, f64p, i32p]
    lib.tnqs_apply_circuit_fetch.argtypes = [p, i32, i32p, i64p, f64p, C.POINTER(ApplyOpts), i32p, i32, C.POINTER(BpOpts), i32, f64p, i32p, i32, i32p, C.POINTER(f64p), i64p]
    lib.tnqs_rdm_1site.argtypes = [p, i32, i32p, f64p]
    lib.tnqs_expect_2site_nn.argtypes = [p, i32, i32p, f64p, f64p]
    lib.tnqs_expect_path.argtypes = [p, i32, i32p, i32p, f64p, f64p]
    lib.tnqs_loop_weights.argtypes = [p, i32, i32p, i32p, f64p]
    lib.tnqs_loop_transfer.argtypes = [p, i32, i32p, f64p]
    lib.tnqs_vertex_scalars.argtypes = [p, f64p]; lib.tnqs_edge_scalars.argtypes = [p, f64p]
    lib.tnqs_rescale.argtypes = [p]
    lib.tnqs_symmetric_gauge.argtypes = [p, C.c_double]
    lib.tnqs_truncate.argtypes = [p, i32, i32p, i32p, C.POINTER(ApplyOpts), i32p, i32, C.POINTER(BpOpts)]
    lib.tnqs_message_slab.argtypes = [p, C.POINTER(vp), i64p, i64p]
    lib.tnqs_halo_pack.argtypes = [p, i32, i32p, vp]; lib.tnqs_halo_unpack.argtypes = [p, i32, i32p, vp]
    lib.tnqs_set_partition.argtypes = [p, i32p, i32]
    lib.tnqs_set_bra.argtypes = [p, p]
    lib.tnqs_halo_alloc.argtypes = [p, i32, i32, i32p, vp, C.POINTER(vp)]
    lib.tnqs_halo_peer.argtypes = [p, i32, vp, vp, i32]
    lib.tnqs_halo_route.argtypes = [p, i32, i32p, i32p, i32p]
    lib.tnqs_halo_close.argtypes = [p]
    lib.tnqs_kernel_launches.argtypes = [p]; lib.tnqs_kernel_launches.restype = i64
    lib.tnqs_bp_sweeps_total.argtypes = [p]; lib.tnqs_bp_sweeps_total.restype = i64
    lib.tnqs_bp_message_updates_total.argtypes = [p]; lib.tnqs_bp_message_updates_total.restype = i64
    lib.tnqs_profile_enable.argtypes = [p, i32]
    lib.tnqs_profile_read.argtypes = [p, f64p, i64p, i32]
    lib.tnqs_stream.argtypes = [p]; lib.tnqs_stream.restype = vp
    _lib = lib
    return lib


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr_i32(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def ptr_i64(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def ptr_f64(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def check(handle, status: int):
    if status != 0:
        lib = load()
        msg = lib.tnqs_last_error(handle)
        raise TnqsError(status, (msg or b"").decode() or lib.tnqs_status_string(status).decode())
