"""Rates of the two hot paths on the lattices and bond dimensions of every BASELINE config (one GPU, random TNS, saturated bonds).

For each shape: a BP refresh of `SWEEPS` sweeps in the reference's forest-cover order and in JACOBI order, and every colour class of
Rzz gates once (maxdim = chi, cutoff 1e-10), all timed with the engine's per-launch CUDA events (tnqs_profile_*) plus the host wall
clock around the call.  Flop counts are the algorithmic ones of SURVEY 8d (bench.flops_per_message / bench.gate_flops); the peak is
the measured ZGEMM rate of profiles/r01_fp64_peaks.json.  One JSON line per shape.

    python tools/config_rates.py [C1 C2 C3 C4 C5s ...]
"""
import sys, os, json, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from tensornetworkquantumsimulator_b200 import engine as E, graphs as G, gates, lib

SWEEPS = 2
SHAPES = {
    # name: (graph factory, chi, what BASELINE.json calls it)
    "C1": (lambda: G.named_grid((5, 5)), 10, "5x5 grid, chi 10"),
    "C2": (lambda: G.named_grid((32, 32)), 16, "32x32 grid, chi 16"),
    "C3": (G.eagle_heavy_hex, 32, "127-qubit heavy-hex, chi 32"),
    "C4": (lambda: G.named_grid((8, 8, 8)), 8, "8x8x8 cubic, chi 8"),
    "C5s": (lambda: G.named_grid((16, 16)), 32, "16x16 strip of the 128x128 grid, chi 32"),
}


def colouring(g):
    try:
        return G.grid_edge_color(g)
    except Exception:
        return G.edge_color(g)


def timed(L, h, fn):
    fn()                                                        # warm: plans, graphs, pools
    L.tnqs_sync(h)
    L.tnqs_profile_read(h, None, None, 1); L.tnqs_profile_enable(h, 1)
    fn(); L.tnqs_sync(h)
    ms = (C.c_double * 8)(); ln = (C.c_int64 * 8)()
    L.tnqs_profile_read(h, ms, ln, 1); L.tnqs_profile_enable(h, 0)
    t0 = time.perf_counter(); fn(); L.tnqs_sync(h)
    timed.classes = {k: round(ms[i], 3) for i, k in enumerate(lib.KERNEL_CLASSES) if ln[i]}
    return sum(ms), int(sum(ln)), 1e3 * (time.perf_counter() - t0)


def run(name):
    mk, chi, label = SHAPES[name]
    g = mk()
    peak = bench.peaks()["fp64_tflops_sustained"]
    psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    seq = G.forest_cover_edge_sequence(g)
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=5); E.rescale(bpc, inplace=True)
    L = lib.load(); h = psi._h.ptr
    fl = bench.flops_per_message(g, chi)
    out = {"shape": name, "what": label, "vertices": g.nv, "edges": g.ne, "chi": chi, "max_degree": max(g.degree(v) for v in range(g.nv)),
           "state_gib": round(bench.psi_bytes(g, chi) / 2 ** 30, 3), "peak_tflops": peak}
    for mode, kw in (("reference_order", dict(edge_sequence=seq)), ("jacobi", dict(edge_sequence=np.arange(2 * g.ne), mode="jacobi"))):
        dev, launches, wall = timed(L, h, lambda: E.update(bpc, inplace=True, maxiter=SWEEPS, **kw))
        t = max(dev, 1e-9) if launches else wall
        out[mode] = {"ms_per_sweep_device": round(dev / SWEEPS, 4), "ms_per_sweep_wall": round(wall / SWEEPS, 4), "launches_per_sweep": launches // SWEEPS,
                     "messages_per_s": round(2 * g.ne * SWEEPS / (wall * 1e-3)), "tflops": round(fl.sum() * SWEEPS / (wall * 1e-3) / 1e12, 3),
                     "frac_of_peak": round(fl.sum() * SWEEPS / (wall * 1e-3) / 1e12 / peak, 4),
                     "frac_of_peak_device_time": round(fl.sum() * SWEEPS / (t * 1e-3) / 1e12 / peak, 4), "class_ms": dict(timed.classes)}
    rzz = gates.gate_array(("Rzz", [0, 1], 0.25))
    akw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
    n_g = 0; f_g = 0.0; dev_g = 0.0; wall_g = 0.0; per = []
    for grp in colouring(g):
        grp = list(grp)
        dev, launches, wall = timed(L, h, lambda: E.apply_2site_batch(bpc, grp, [rzz] * len(grp), apply_kwargs=akw))
        f = sum(bench.gate_flops(g, e, chi) for e in grp)
        n_g += len(grp); f_g += f; dev_g += dev; wall_g += wall
        per.append({"gates": len(grp), "ms_device": round(dev, 3), "ms_wall": round(wall, 3), "launches": launches, "class_ms": dict(timed.classes)})
    out["two_site_gates"] = {"classes": per, "gates_per_s": round(n_g / (wall_g * 1e-3)), "nominal_tflops": round(f_g / (wall_g * 1e-3) / 1e12, 3),
                             "frac_of_peak_nominal": round(f_g / (wall_g * 1e-3) / 1e12 / peak, 4), "bond_dims_after": sorted(set(int(x) for x in psi.bond_dims()))}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(SHAPES)):
        run(nm)
