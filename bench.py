#!/usr/bin/env python
"""bench.py — headline benchmark of the tnqs-cuda engine (contract: task statement; definitions: DESIGN.md section 5).

N = 1 (BASELINE.json configs[1], "C2"): 32x32 open grid, random TensorNetworkState chi=16 ComplexF64 (counter-based generator,
seed 1, BP-rescaled).  One step = one Trotter layer: Rx, Rz on every vertex + Rzz in 4 deterministic colour classes with maxdim=16,
cutoff=1e-10, executed as the exact run / refresh schedule of the reference's apply_gates (lazy BP refreshes with the default
bp_update_kwargs: <=25 sweeps, tolerance 1e-8, forest-cover edge order, REFERENCE_ORDER level schedule) and <Z> at the centre.
N > 1: weak scaling, one 32x32 row strip per GPU of a 32x(32 N) lattice; BP in JACOBI mode, a refresh = one graph launch per rank
(boundary messages stored into the neighbour GPU's buffer by the message epilogue, device-side barrier + all-reduce over NVLink
peer memory; TNQS_BENCH_HALO=nccl: pack / send / recv / unpack per sweep instead), gates across a cut through
tnqs_apply_2site_begin/_end with one R-factor exchange per colour class.

One JSON line on stdout (rank 0).  value = BP edge-message updates / s over the BP phases (device-resident state, CUDA events on
the engine's stream); `gates` = two-site gate applications / s over the gate phases of the same steps; `e2e` = the same step
through the host API with the state in pinned HOST memory (every site tensor H2D, the layer, every site tensor D2H inside the
timed region); `roofline` = algorithmic flops of the dominant kernel / its CUDA-event time against the measured cuBLAS ZGEMM peak;
`cpu_baseline` / `--impl reference` = the CPU oracle (NumPy + OpenBLAS, all host threads) on a bounded sample of the same tensors.
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def peaks():
    out = {"fp64_tflops": 36.995, "fp64_tflops_sustained": 36.936, "hbm_gbs": 6541.1, "source": "fallback"}
    p = os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")
    if os.path.exists(p):
        d = json.load(open(p))["roofline_denominator"]
        out.update(fp64_tflops=d["fp64_tflops"], fp64_tflops_sustained=d["fp64_tflops_sustained"],
                   source="profiles/r01_fp64_peaks.json (cuBLAS ZGEMM 8192^3 measured on this pool's B200)")
    m = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(m):
        out["hbm_gbs"] = json.load(open(m)).get("hbm_gbs", out["hbm_gbs"])
    return out


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe).  Sampled through NVML inside this
    process (a thread, one cheap query every TNQS_BENCH_SMI_MS ms): a concurrent `nvidia-smi -lms` loop takes driver locks
    for milliseconds per query and stalls the host-driven gate phases (measured: +100 ms outliers per step).
    TNQS_BENCH_SMI=nvidia-smi selects the external sampler, TNQS_BENCH_NO_SMI=1 none."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
    ACTIVE = None          # the sampler of the running timed region; the stepper opens its window around GPU-bound calls

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.thread, self.stop_flag = index, None, [], None, False
        self.sm, self.mx, self.reasons = [], [], set()
        self.window = threading.Event()        # set while the main thread sits in a GPU-bound call (a BP refresh = one graph launch):
        self.windowed = False                  # queries issued then cannot stall host-driven phases

    @classmethod
    def gpu_bound(cls, on: bool):
        s = cls.ACTIVE
        if s is not None:
            s.windowed = True
            (s.window.set if on else s.window.clear)()

    def start(self):
        ClockSampler.ACTIVE = self
        if os.environ.get("TNQS_BENCH_NO_SMI") == "1":
            return
        if os.environ.get("TNQS_BENCH_SMI") != "nvidia-smi":
            try:
                import pynvml
                pynvml.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
                self.nvml, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
                self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)))
                self.thread = threading.Thread(target=self._poll, daemon=True)
                self.thread.start()
                return
            except Exception:
                self.thread = None
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", os.environ.get("TNQS_BENCH_SMI_MS", "200"), "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        period = float(os.environ.get("TNQS_BENCH_SMI_MS", "100")) * 1e-3
        n = self.nvml
        last = 0.0

        def sample():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass

        while not self.stop_flag:
            if self.windowed:
                # one query at the OPENING of a GPU-bound window (a BP refresh lasts >= 10 ms), never late in it: the query must be
                # over before the host-driven phase that follows starts launching
                if not self.window.wait(0.02):
                    continue
                if time.monotonic() - last >= period:
                    sample()
                    last = time.monotonic()
                while self.window.is_set() and not self.stop_flag:
                    time.sleep(0.001)
            else:
                sample()
                time.sleep(period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        ClockSampler.ACTIVE = None
        if self.proc is None and self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        sm, mx, reasons = self.sm, self.mx, self.reasons
        if self.proc is None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            how = "nvml"
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            how = "nvidia-smi"
            for ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * max(mx)] if mx else []
        return {"sm_mhz": float(np.median(busy)) if busy else (float(np.median(sm)) if sm else None),
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "sampler": how + (" (queried while a BP refresh occupies the GPU)" if self.windowed else "")}


def workload(nx: int, ny: int):
    from tensornetworkquantumsimulator_b200 import graphs as G, gates
    g = G.named_grid((nx, ny))
    groups = G.grid_edge_color(g)
    dt, hx, hz, J = 0.25, 1.0, 0.8, 0.5            # examples/2dIsing_dynamics.jl:13-17
    rx = gates.gate_array(("Rx", [0], 2 * hx * dt)); rz = gates.gate_array(("Rz", [0], 2 * hz * dt))
    rzz = gates.gate_array(("Rzz", [0, 1], 2 * J * dt))
    seq = G.forest_cover_edge_sequence(g)
    return g, groups, rx, rz, rzz, seq


def flops_per_message(g, chi, d=2):
    """Algorithmic flops of one message update in the reference's contraction order: 8*d*P*sum_k chi_k."""
    f = np.zeros(2 * g.ne)
    for k in range(2 * g.ne):
        v = g.dsrc(k)
        z = g.degree(v)
        f[k] = 8.0 * d * chi ** z * z * chi
    return f


def gate_flops(g, e, chi, d=2):
    """Nominal flops of one two-site gate: absorb + QR(explicit Q) + reabsorb + Q R' per site, theta, SVD."""
    tot = 0.0
    n = d * chi
    qs = []
    for v in (int(g.esrc[e]), int(g.edst[e])):
        z = g.degree(v)
        P = chi ** z
        Po = P // chi
        q = min(Po, n)
        qs.append(q)
        tot += 8.0 * d * P * (z - 1) * chi                      # absorb sqrt(M)
        tot += (z - 1) * 52.0 * chi ** 3                        # eigh + roots (nominal)
        tot += 2 * (8.0 * Po * n * n - 8.0 * n ** 3 / 3) if Po >= n else 2 * (8.0 * n * Po * Po - 8.0 * Po ** 3 / 3)
        tot += 8.0 * Po * q * (z - 1) * chi                     # reabsorb M^-1/2 into Q
        tot += 8.0 * Po * q * d * chi                           # Q R'
    q1, q2 = qs
    tot += 8.0 * (q1 * q2 * d * d * chi + d ** 4 * q1 * q2)
    m, nn = max(q1, q2) * d, min(q1, q2) * d
    tot += 4.0 * (4 * m * m * nn + 8 * m * nn * nn + 9 * nn ** 3)
    return tot


class Stepper:
    """One Trotter layer as the phase sequence tnqs_apply_circuit executes (tests/test_gpu_parity.py checks the
    equivalence), with CUDA events on the engine's stream around the BP and the gate phases."""

    def __init__(self, E, torch, bpc, g, groups, rx, rz, rzz, seq, chi, obs_v):
        self.E, self.torch, self.bpc, self.g, self.groups, self.seq, self.chi, self.obs_v = E, torch, bpc, g, groups, seq, chi, obs_v
        self.rx, self.rz, self.rzz = rx, rz, rzz
        from tensornetworkquantumsimulator_b200 import lib
        self.stream = torch.cuda.ExternalStream(lib.load().tnqs_stream(bpc.psi._h.ptr))
        self.kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
        self.bp_kw = E.default_bp_update_kwargs(bpc)
        self.reset()

    def reset(self):
        self.bp_ms = self.gate_ms = self.other_ms = 0.0
        self.msg_updates = self.gates2 = self.gates1 = 0
        self.bp_sweeps = 0

    def _timed(self, fn):
        t = self.torch
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        r = fn()
        e1.record(self.stream)
        e1.synchronize()
        return r, e0.elapsed_time(e1)

    def bp(self):
        info = {}
        ClockSampler.gpu_bound(True)
        _, ms = self._timed(lambda: self.E.update(self.bpc, edge_sequence=self.seq, inplace=True, info=info, **self.bp_kw))
        ClockSampler.gpu_bound(False)
        self.bp_ms += ms
        self.bp_sweeps += info["iters"]
        self.msg_updates += info["iters"] * len(self.seq)

    def step(self):
        """One Trotter layer as the exact run/refresh schedule of apply_gates (engine.circuit_schedule mirrors the rule of
        tnqs_apply_circuit; tests/test_gpu_parity.py checks bitwise equality with apply_gates)."""
        E, g = self.E, self.g
        if not hasattr(self, "sched"):
            self.layer = [(self.rx, [v]) for v in range(g.nv)] + [(self.rz, [v]) for v in range(g.nv)]
            for grp in self.groups:
                self.layer += [(self.rzz, [int(g.esrc[e]), int(g.edst[e])]) for e in grp]
            # the batches of the schedule with their arguments resolved once (vertex ids, edge ids, orientation, gate arrays): what a
            # compiled host would do per call in microseconds is kept out of the timed regions of an interpreted one
            self.sched = []
            for item in E.circuit_schedule(g, self.layer):
                if item[0] == "refresh":
                    self.sched.append(item)
                elif item[0] == "gates1":
                    ids, folds = item[1], item[2]
                    # consecutive one-site gates of a vertex (Rx, Rz) are one gate, as in tnqs_apply_circuit
                    arrs = [E.fold_gates(self.layer[i][0], [self.layer[k][0] for k in f]) for i, f in zip(ids, folds)]
                    self.sched.append(("gates1", [self.layer[i][1][0] for i in ids], arrs, len(ids) + sum(len(f) for f in folds)))
                else:
                    ids = item[1]
                    es = [g.edge_of[(self.layer[i][1][0], self.layer[i][1][1])] for i in ids]
                    fl = [0 if int(g.esrc[e]) == self.layer[i][1][0] else 1 for e, i in zip(es, ids)]
                    self.sched.append(("gates2", es, [self.layer[i][0] for i in ids], fl))
        for item in self.sched:
            if item[0] == "refresh":
                self.bp()
            elif item[0] == "gates1":
                _, ms = self._timed(lambda: E.apply_1site_batch(self.bpc, item[1], item[2]))
                self.other_ms += ms; self.gates1 += item[3]
            else:
                _, ms = self._timed(lambda: E.apply_2site_batch(self.bpc, item[1], item[2], flip=item[3], apply_kwargs=self.kw))
                self.gate_ms += ms; self.gates2 += len(item[1])
        (z,), ms = self._timed(lambda: (E.expect(self.bpc, ("Z", [self.obs_v])),))
        self.other_ms += ms
        return z


def run_distributed(args, torch, dist, rank, world, local):
    """N > 1: weak scaling.  Global lattice nx x (ny*N) cut into N row strips (one per GPU); BP in JACOBI mode with one
    NCCL halo exchange + one all-reduce per sweep; gates across a cut through tnqs_apply_2site_begin/_end with one
    R-factor exchange per colour class (tensornetworkquantumsimulator_b200/parallel.py)."""
    from tensornetworkquantumsimulator_b200 import engine as E, lib, graphs as G, gates, parallel as P
    chi, nx, ny = args.chi, args.nx, args.ny
    dev = torch.device("cuda", local)
    gg = G.named_grid((nx, ny * world))
    groups = G.grid_edge_color(gg)
    owner = P.strip_owner(gg, world, axis=1)
    lg = P.build_local_graph(gg, owner, rank)
    dt, hx, hz, J = 0.25, 1.0, 0.8, 0.5
    rx = gates.gate_array(("Rx", [0], 2 * hx * dt)); rz = gates.gate_array(("Rz", [0], 2 * hz * dt)); rzz = gates.gate_array(("Rzz", [0, 1], 2 * J * dt))
    rzrx = E.fold_gates(rx, [rz])
    psi = E.random_tensornetworkstate(lg.g, bond_dimension=chi, seed=1 + rank, device=local)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    ow = lib.i32(lg.owner_local)
    lib.check(psi._h.ptr, lib.load().tnqs_set_partition(psi._h.ptr, lib.ptr_i32(ow), rank))
    comp = P.EngineCompute(bpc, dev)
    slot = chi * chi
    # boundary messages: written by the BP kernels straight into the neighbour's buffer (peer memory over NVLink, CUDA IPC);
    # TNQS_BENCH_HALO=nccl selects the pack / send / recv / unpack exchange instead
    p2p = None if os.environ.get("TNQS_BENCH_HALO") == "nccl" else P.halo_setup_p2p(lg, comp, dist, rank, world)
    kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
    owned_v = [v for v in range(lg.g.nv) if lg.owned[v]]
    P.update_distributed(lg, comp, dist, slot, torch, dev, maxiter=10, p2p_bytes=p2p)
    E.rescale(bpc, inplace=True)
    P.halo_exchange(lg, comp, dist, slot, torch, dev)
    batches = [P.class_gates_local(gg, lg, owner, rank, grp) for grp in groups]
    stats = {"bp_s": 0.0, "gate_s": 0.0, "other_s": 0.0, "msg": 0, "gates2": 0, "sweeps": 0, "halo": 0}

    estream = torch.cuda.ExternalStream(lib.load().tnqs_stream(psi._h.ptr))
    # TNQS_BENCH_TRACE=1: every rank prints, for each timed call of its last step, (host start in ms since the step began, host ms
    # inside the call, device ms) - inter-rank skew and host stalls show up here
    trace = [] if os.environ.get("TNQS_BENCH_TRACE") == "1" else None
    trace_t0 = [time.perf_counter()]

    def dev_timed(fn):
        """Device time of fn: CUDA events on the engine's stream and on torch's current stream (NCCL), after a full
        synchronisation; the later end event counts."""
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        w0 = time.perf_counter()
        e0.record(estream)
        r = fn()
        w1 = time.perf_counter()
        e1.record(estream); e2.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        dt = 1e-3 * max(e0.elapsed_time(e1), e0.elapsed_time(e2))
        if trace is not None:
            trace.append((round(1e3 * (w0 - trace_t0[0]), 2), round(1e3 * (w1 - w0), 2), round(1e3 * dt, 2)))
        return r, dt

    def bp():
        info = {}
        ClockSampler.gpu_bound(True)
        _, dt = dev_timed(lambda: P.update_distributed(lg, comp, dist, slot, torch, dev, maxiter=25, tolerance=1e-8, info=info, p2p_bytes=p2p))
        ClockSampler.gpu_bound(False)
        stats["bp_s"] += dt
        stats["msg"] += info["iters"] * info["messages_per_sweep"]; stats["sweeps"] += info["iters"]; stats["halo"] += info["halo_bytes"]

    def step():
        if trace is not None:
            trace.clear(); trace_t0[0] = time.perf_counter()
        # Rx then Rz on a vertex = one gate, their product, as tnqs_apply_circuit applies them (engine.fold_gates)
        _, dt = dev_timed(lambda: E.apply_1site_batch(bpc, owned_v, [rzrx] * len(owned_v)))
        stats["other_s"] += dt
        for grp, batch in zip(groups, batches):
            bp()
            _, dt = dev_timed(lambda: P.apply_class_distributed(E, bpc, gg, lg, owner, rank, grp, rzz, kw, dist, torch, dev, batch=batch))
            stats["gate_s"] += dt
            es, rm, _ = batch
            stats["gates2"] += int(np.sum(rm == 0)) + 0.5 * int(np.sum(rm != 0))          # a cut gate counts half on each side
        bp()
        return E.expect(bpc, ("Z", [owned_v[len(owned_v) // 2]]))

    for _ in range(args.warmup):
        step()
    for k in stats:
        stats[k] = 0
    launches0 = psi.kernel_launches()
    gc.collect(); gc.freeze(); gc.disable()                       # no collector pauses inside the timed region (as timeit does)
    dist.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    zs = [step() for _ in range(args.steps)]
    torch.cuda.synchronize()
    ev1.record(); torch.cuda.synchronize()
    dist.barrier()
    total_ms = ev0.elapsed_time(ev1)
    gc.enable()
    if trace is not None:
        print(f"[trace rank {rank}] (start, host, device) ms per call of the last step: {trace}", file=sys.stderr, flush=True)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, stats["bp_s"], stats["gate_s"], stats["other_s"], float(stats["msg"]), float(stats["gates2"]), float(stats["sweeps"]),
                      float(stats["halo"]), float(psi.kernel_launches() - launches0)], dtype=torch.float64, device=dev)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    total_ms, bp_s, gate_s, other_s = mx[:4].tolist()
    msg, gates2, halo, launches = sm[4].item(), sm[5].item(), sm[7].item(), sm[8].item()
    pk = peaks()
    fm_all = flops_per_message(lg.g, chi)
    owned_k = [k for k in range(2 * lg.g.ne) if lg.owned[lg.g.dsrc(k)]]
    fl = torch.tensor([float(fm_all[owned_k].sum()), float(len(owned_k))], dtype=torch.float64, device=dev)
    dist.all_reduce(fl)
    fm = fl[0].item() / fl[1].item()                                    # algorithmic flops per message, averaged over the whole lattice
    out = {
        "metric": "bp_edge_message_updates_per_sec", "value": msg / bp_s, "unit": "messages/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
        "config": bench_config(nx, ny, chi, world, p2p is not None),
        "gates": {"metric": "two_site_gate_applications_per_sec", "value": gates2 / gate_s, "unit": "gates/s"},
        "bp": {"sweeps_per_step": mx[6].item() / args.steps, "approx_tflops": msg / bp_s * fm / 1e12, "halo_bytes_per_step": halo / args.steps},
        "phase_ms_per_step": {"bp": 1e3 * bp_s / args.steps, "two_site_gates": 1e3 * gate_s / args.steps, "one_site": 1e3 * other_s / args.steps},
        "roofline": {"bound": "tensor", "kernel": (f"k_absorb_slab<{chi}> + k_close_dmma<{chi}> (the two launches of a JACOBI sweep)" if chi in (8, 16) else "k_leg_dmma + k_close_leg_dmma (generic tensor-core leg / closing kernels of a JACOBI sweep, chunked levels; the four updates leaving a vertex share their absorptions: 8 leg passes instead of the 12 the algorithmic flop count stands for)") + " incl. the device-side barrier; whole BP phases, per GPU",
                     "achieved": msg / bp_s * fm / 1e12 / world, "peak": pk["fp64_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": msg / bp_s * fm / 1e12 / world / pk["fp64_tflops_sustained"], "traffic": None, "peak_source": pk["source"],
                     "note": "per-GPU rate over the BP phases (max over ranks, so the skew between ranks at the first sweep barrier is inside); single-kernel counters: profiles/README.md"},
        "gpu_launches": int(launches), "clocks": clocks, "expect_z": [float(np.real(z)) for z in zs],
    }
    if args.no_e2e:
        return out
    # ---- e2e: the same distributed step with this rank's site tensors in pinned HOST memory (H2D before, D2H after) ----
    L = lib.load(); h = psi._h.ptr
    sizes = [int(L.tnqs_site_size(h, v)) for v in owned_v]
    host = [torch.empty(2 * n, dtype=torch.float64).pin_memory() for n in sizes]
    ptrs = [C.cast(t_.data_ptr(), C.POINTER(C.c_double)) for t_ in host]
    vs_arr = lib.i32(owned_v); ns_arr = np.asarray(sizes, dtype=np.int64)
    parr = (C.POINTER(C.c_double) * len(ptrs))(*ptrs)
    lib.check(h, L.tnqs_get_sites(h, len(ptrs), lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))
    e2e_t, e2e_msg = 0.0, 0
    nrep = max(1, min(args.steps, 2))
    for rep in range(nrep + 1):                                   # first pass is a warm-up
        dist.barrier(); torch.cuda.synchronize()
        m0 = stats["msg"]; t0 = time.perf_counter()
        lib.check(h, L.tnqs_set_sites(h, len(ptrs), lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))
        step()
        lib.check(h, L.tnqs_get_sites(h, len(ptrs), lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))
        torch.cuda.synchronize(); dist.barrier()
        if rep > 0:
            e2e_t += time.perf_counter() - t0; e2e_msg += stats["msg"] - m0
    te = torch.tensor([e2e_t, float(e2e_msg), float(sum(16 * n for n in sizes))], dtype=torch.float64, device=dev)
    tmx = te.clone(); dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
    tsm = te.clone(); dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
    out["e2e"] = {"value": tsm[1].item() / tmx[0].item(), "unit": "messages/s", "h2d_bytes_per_step": int(tsm[2].item()),
                  "d2h_bytes_per_step": int(tsm[2].item()), "ms_per_step": 1e3 * tmx[0].item() / nrep,
                  "note": "every owned site tensor H2D from pinned host memory (tnqs_set_sites), the distributed layer, every owned site tensor D2H (tnqs_get_sites); host wall clock around the public calls, max over ranks"}
    return out


def run_engine(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if world > 1:
        out = run_distributed(args, torch, dist, rank, world, local)
        dist.destroy_process_group()
        return out, None
    from tensornetworkquantumsimulator_b200 import engine as E, lib
    chi, nx, ny = args.chi, args.nx, args.ny
    g, groups, rx, rz, rzz, seq = workload(nx, ny)
    t0 = time.time()
    psi = E.random_tensornetworkstate(g, bond_dimension=chi, seed=1 + rank, device=local)
    bpc = E.BeliefPropagationCache(psi, _share=True)
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=10)
    E.rescale(bpc, inplace=True)
    setup_s = time.time() - t0
    st = Stepper(E, torch, bpc, g, groups, rx, rz, rzz, seq, chi, obs_v=g.v((nx // 2, ny // 2)))
    for _ in range(args.warmup):
        st.step()
    st.reset()
    launches0 = psi.kernel_launches()
    gc.collect(); gc.freeze(); gc.disable()                       # no collector pauses inside the timed region (as timeit does)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(st.stream)
    zs, step_wall = [], []
    for _ in range(args.steps):
        t0 = time.perf_counter(); zs.append(st.step()); step_wall.append(1e3 * (time.perf_counter() - t0))
    ev1.record(st.stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev0.elapsed_time(ev1)
    gc.enable()
    clocks = sampler.stop() if rank == 0 else None
    launches = psi.kernel_launches() - launches0
    stats = torch.tensor([total_ms, st.bp_ms, st.gate_ms, st.other_ms, float(st.msg_updates), float(st.gates2), float(st.bp_sweeps)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, bp_ms, gate_ms, other_ms = mx[:4].tolist()
        msg_updates, gates2 = sm[4].item(), sm[5].item()
        bp_sweeps = mx[6].item()
    else:
        total_ms, bp_ms, gate_ms, other_ms, msg_updates, gates2, bp_sweeps = stats.tolist()

    # ---- roofline of the dominant kernel (k_mode_apply inside the BP sweeps): one profiled sweep, CUDA events per launch ----
    pk = peaks()
    L = lib.load()
    L.tnqs_profile_read(psi._h.ptr, None, None, 1)
    L.tnqs_profile_enable(psi._h.ptr, 1)
    E.update(bpc, edge_sequence=seq, inplace=True, maxiter=2)
    ms = (C.c_double * 8)(); ln = (C.c_int64 * 8)()
    L.tnqs_profile_read(psi._h.ptr, ms, ln, 1)
    L.tnqs_profile_enable(psi._h.ptr, 0)
    fm = flops_per_message(g, chi)
    sweep_flops = float(fm[np.asarray(seq)].sum())
    zsum = sum(g.degree(g.dsrc(int(k))) for k in seq)
    apply_flops = sweep_flops - 8.0 * 2 * sum(chi ** g.degree(g.dsrc(int(k))) * chi for k in seq)   # sweeps minus the closings
    one_launch_per_sweep = ln[0] == 2 and ln[1] == 0            # k_bp_sweep: absorptions + closings + epilogues of a whole sweep in one launch
    if one_launch_per_sweep:
        apply_tf = 2 * sweep_flops / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else None
        close_tf = None
    else:
        apply_tf = 2 * apply_flops / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else None
        close_tf = 2 * (sweep_flops - apply_flops) / (ms[1] * 1e-3) / 1e12 if ms[1] > 0 else None
    jac_seq = np.arange(2 * g.ne, dtype=np.int32)
    E.update(bpc, edge_sequence=jac_seq, inplace=True, maxiter=2, mode="jacobi")
    j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    j0.record(st.stream); E.update(bpc, edge_sequence=jac_seq, inplace=True, maxiter=6, mode="jacobi"); j1.record(st.stream); j1.synchronize()
    jacobi_value = 6 * len(jac_seq) / (j0.elapsed_time(j1) * 1e-3)
    bp_value = msg_updates / (bp_ms * 1e-3)
    bp_tf = bp_value * sweep_flops / len(seq) / 1e12
    gflops = sum(gate_flops(g, e, chi) for grp in groups for e in grp) / sum(len(grp) for grp in groups)
    gate_value = gates2 / (gate_ms * 1e-3)
    out = {
        "metric": "bp_edge_message_updates_per_sec", "value": bp_value, "unit": "messages/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
        "config": bench_config(nx, ny, chi, 1),
        "gates": {"metric": "two_site_gate_applications_per_sec", "value": gate_value, "unit": "gates/s",
                  "gflop_per_gate_nominal": gflops / 1e9, "achieved_tflops": gate_value * gflops / 1e12,
                  "flop_count": "nominal count of the reference's algorithm (SURVEY 8d: explicit Q, inverse absorption, full Golub-Reinsch SVD of theta'); "
                                "the engine executes less of it (R only, T' = psi R^-1 R', SVD on the row space of theta', 3M products)",
                  "frac_of_fp64_peak": gate_value * gflops / 1e12 / pk["fp64_tflops_sustained"]},
        "bp": {"sweeps_per_step": bp_sweeps / args.steps, "gflop_per_message": sweep_flops / len(seq) / 1e9, "achieved_tflops": bp_tf,
               "jacobi_mode_messages_per_s": jacobi_value, "jacobi_mode_frac_of_fp64_peak": jacobi_value * sweep_flops / len(seq) / 1e12 / pk["fp64_tflops_sustained"],
               "frac_of_fp64_peak": bp_tf / pk["fp64_tflops_sustained"]},
        "phase_ms_per_step": {"bp": bp_ms / args.steps, "two_site_gates": gate_ms / args.steps, "one_site_and_expect": other_ms / args.steps},
        "roofline": {"bound": "tensor", "kernel": "k_bp_sweep<16> (one persistent launch per BP sweep: DMMA leg absorptions, closings and message epilogues of all 3968 updates)" if one_launch_per_sweep
                     else "k_absorb_slab<16> (fused BP leg absorption, DMMA)", "achieved": apply_tf, "peak": pk["fp64_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": (apply_tf / pk["fp64_tflops_sustained"]) if apply_tf else None,
                     "flop_count": "algorithmic: 8 flops per complex multiply-add of the reference's contraction (SURVEY 8d); the kernels form a complex "
                                   "product from three real DMMA products (3M), i.e. execute 3/4 of the multiplications that count stands for",
                     "traffic": (20.63e9 if one_launch_per_sweep else 206.0e6),
                     "traffic_detail": ({"dram_bytes_read": 11.98e9, "dram_bytes_written": 8.64e9, "algorithmic_bytes_per_launch": float(sum(16.0 * 2 * chi ** g.degree(g.dsrc(int(k))) for k in seq)),
                                         "note": "one launch = one sweep of 3968 updates (ncu --set full, profiles/r02g_bp_sweep_ncu_full_raw.csv); algorithmic = every site tensor read once per update; the rest is the "
                                                 "slab-major intermediate X (written once, read once by the closing) and the closing's second read of the site tensor, partly served by L2 (hit rate 40 %)"}
                                        if one_launch_per_sweep else {"dram_bytes_per_launch": 206.0e6, "algorithmic_bytes_per_launch": 260.0e6, "launch": "62 messages x 32 slabs (profiles/r01_final_dmma_ncu_full_raw.csv)"}),
                     "peak_source": pk["source"], "launches_profiled": int(ln[0]), "avg_launch_ms": ms[0] / max(1, ln[0]),
                     "close_kernel_tflops": close_tf},
        "gpu_launches": int(launches), "clocks": clocks, "step_wall_ms": [round(x, 1) for x in step_wall],
        "expect_z_center": [float(np.real(z)) for z in zs], "setup_s": setup_s,
    }
    del zsum
    return out, (E, torch, bpc, g, groups, rx, rz, rzz, seq, chi, st)


def psi_bytes(g, chi, d=2):
    return sum(16 * d * chi ** g.degree(v) for v in range(g.nv))


def run_e2e(ctx, steps):
    """The same step through the host API with the state in pinned HOST memory: every site tensor H2D, the layer
    (apply_gates -> tnqs_apply_circuit), <Z>, every site tensor D2H, inside the timed region."""
    E, torch, bpc, g, groups, rx, rz, rzz, seq, chi, st = ctx
    from tensornetworkquantumsimulator_b200 import lib
    L = lib.load()
    h = bpc.psi._h.ptr
    sizes = [int(L.tnqs_site_size(h, v)) for v in range(g.nv)]
    host = [torch.empty(2 * n, dtype=torch.float64).pin_memory() for n in sizes]
    ptrs = [C.cast(t.data_ptr(), C.POINTER(C.c_double)) for t in host]
    vs_arr = lib.i32(np.arange(g.nv)); ns_arr = np.asarray(sizes, dtype=np.int64)
    parr = (C.POINTER(C.c_double) * g.nv)(*ptrs)
    lib.check(h, L.tnqs_get_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))
    layer = [(rx, [v]) for v in range(g.nv)] + [(rz, [v]) for v in range(g.nv)]
    for grp in groups:
        layer += [(rzz, [int(g.esrc[e]), int(g.edst[e])]) for e in grp]
    kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
    h2d = sum(16 * n for n in sizes) + 16 * (2 * 4 * g.nv + 16 * sum(len(x) for x in groups)) + 4 * len(seq)
    d2h = sum(16 * n for n in sizes) + 8 * len(layer) + 16 * 4
    t_tot = 0.0
    upd = 0
    parts = np.zeros(4)
    per_step = []
    gc.collect(); gc.freeze(); gc.disable()
    for s in range(steps + 1):                       # first pass is a warm-up
        torch.cuda.synchronize()
        u0 = L.tnqs_bp_message_updates_total(h)
        t0 = time.perf_counter()
        lib.check(h, L.tnqs_set_sites(h, g.nv, lib.ptr_i32(vs_arr), parr, lib.ptr_i64(ns_arr)))
        t1 = time.perf_counter()
        info = {}
        fns = ns_arr.copy()
        _, errs = E.apply_gates(layer, bpc, apply_kwargs=kw, inplace=True, info=info, fetch=(vs_arr, parr, fns))      # D2H of every site beside the closing refresh
        t2 = time.perf_counter()
        z = E.expect(bpc, ("Z", [st.obs_v]))
        torch.cuda.synchronize()
        t3 = t4 = time.perf_counter()
        dt = t4 - t0
        if s > 0:
            t_tot += dt
            per_step.append(round(1e3 * dt, 1))
            upd += L.tnqs_bp_message_updates_total(h) - u0
            parts += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
    gc.enable()
    return {"value": upd / t_tot, "message_updates_per_step": upd / steps, "unit": "messages/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": 1e3 * t_tot / steps, "steps": steps, "step_ms": per_step,
            "parts_ms": {"h2d_set_sites": 1e3 * parts[0] / steps, "apply_gates_with_d2h_of_all_sites": 1e3 * parts[1] / steps, "expect": 1e3 * parts[2] / steps},
            "note": "host wall clock around tnqs_set_sites (all V sites from pinned host memory) + apply_gates(fetch=...) = tnqs_apply_circuit_fetch (the layer; every site tensor copied back to pinned host memory beside the closing BP refresh) + expect"}


def cpu_sample(nx, ny, chi, n_msgs, n_gates, seed=1):
    """Oracle (NumPy + OpenBLAS, all host threads) on a bounded sample of the same workload: n_msgs message updates in
    the reference's edge order and n_gates two-site gates, on tensors drawn from the same counter-based generator."""
    from oracle import tnqs_oracle as O
    from tensornetworkquantumsimulator_b200 import graphs as G, gates, rng
    g, groups, rx, rz, rzz, seq = workload(nx, ny)
    need = set()
    ks = [int(k) for k in seq[len(seq) // 2: len(seq) // 2 + n_msgs]]
    for k in ks:
        need.add(g.dsrc(k))
    es = groups[0][len(groups[0]) // 2: len(groups[0]) // 2 + n_gates]
    for e in es:
        need |= {int(g.esrc[e]), int(g.edst[e])}
    ts = [None] * g.nv
    for v in range(g.nv):
        shape = (2,) + (chi,) * g.degree(v)
        if v in need:
            ts[v] = rng.complex_normal(seed, v, int(np.prod(shape))).reshape(shape, order="F") / chi ** (g.degree(v) / 2)
        else:
            ts[v] = np.zeros((2,) + (1,) * g.degree(v), dtype=np.complex128)     # placeholder, never touched
    psi = O.TensorNetworkState(g, ts)
    psi.bond_dim = lambda e: chi
    bpc = O.BeliefPropagationCache(psi)
    rs = np.random.default_rng(0)
    for v in need:                                   # Hermitian PSD messages into the sampled vertices
        for e in g.incident[v]:
            a = rs.normal(size=(chi, chi)) + 1j * rs.normal(size=(chi, chi))
            m = a @ a.conj().T
            bpc.messages[g.dedge(g.other(e, v), v)] = m / m.sum()
    t0 = time.perf_counter()
    for k in ks:
        new = O.updated_message(bpc, k)
        O.message_diff(new, O.message(bpc, k))
    t_msg = time.perf_counter() - t0
    t0 = time.perf_counter()
    for e in es:
        O.simple_update_2site(bpc, rzz, int(g.esrc[e]), int(g.edst[e]), maxdim=chi, cutoff=1e-10)
    t_gate = time.perf_counter() - t0
    return len(ks) / t_msg, (len(es) / t_gate) if es else None, t_msg, t_gate


def bench_config(nx, ny, chi, world, p2p=True):
    """The `config` object of the JSON line - ONE definition for the engine arm and the `--impl reference` arm, so that the driver
    compares like with like (its `same_config` check)."""
    from tensornetworkquantumsimulator_b200 import graphs as G
    if world > 1:
        name = "C5" if (nx, ny * world, chi) == (128, 128, 32) else "C2 strips"
        return {"workload": f"{name}: {nx}x{ny * world} open grid cut into {world} row strips of {nx}x{ny} (one per GPU), random TNS chi={chi} "
                            "ComplexF64; step = one Trotter layer (Rx,Rz + 4 Rzz classes, maxdim=chi, cutoff=1e-10), lazy BP refreshes "
                            "(<=25 sweeps, tol 1e-8) + <Z>", "chi": chi,
                "bp_mode": ("JACOBI; a refresh is one graph launch per rank: boundary messages stored by the message epilogue straight into the neighbour GPU's buffer (P2P over NVLink, CUDA IPC), device-side barrier + all-reduce + convergence test per sweep"
                            if p2p else "JACOBI + NCCL halo exchange (pack / send / recv / unpack) and all-reduce per sweep"),
                "gates": "cut-crossing gates via R-factor exchange (tnqs_apply_2site_begin/_end)", "per_rank_lattice": f"{nx}x{ny}",
                "l2": f"inputs larger than L2: state {psi_bytes(G.named_grid((nx, ny)), chi) / 2**30:.2f} GiB per GPU vs 126 MB"}
    g = G.named_grid((nx, ny))
    return {"workload": f"C2: {nx}x{ny} open grid, random TNS chi={chi} ComplexF64 seed 1 (BP-rescaled); step = one Trotter layer "
                        "(Rx,Rz all sites + 4 Rzz colour classes, maxdim=chi, cutoff=1e-10) with the reference's lazy BP refreshes "
                        "(<=25 sweeps, tol 1e-8, forest-cover order, REFERENCE_ORDER) + <Z> at the centre",
            "chi": chi, "messages_per_sweep": 2 * g.ne, "bp_mode": "REFERENCE_ORDER", "per_rank_lattice": f"{nx}x{ny}",
            "l2": f"inputs larger than L2: state {psi_bytes(g, chi) / 2**30:.2f} GiB per GPU vs 126 MB"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm as restated by the oracle (Julia cannot run here),
    all host threads, bounded sample of the same config per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count()
    # bounded sample per step so that the whole --steps/--warmup run stays within a few minutes (~15 s of CPU work per 3000 messages)
    budget = max(1, args.warmup + args.steps)
    n_msgs = int(max(200, min(args.cpu_msgs, args.cpu_msgs * 8 // budget)))
    n_gates = int(max(10, min(args.cpu_gates, args.cpu_gates * 8 // budget)))
    vals, gvals = [], []
    for s in range(args.warmup + args.steps):
        mps, gps, t_msg, t_gate = cpu_sample(args.nx, args.ny, args.chi, n_msgs, n_gates)
        if s >= args.warmup:
            vals.append(mps); gvals.append(gps)
    v = float(np.mean(vals))
    sample = f"{n_msgs} message updates (middle of the forest-cover sweep) + {n_gates} two-site gates per step on the C2 tensors"
    return {"impl": "reference", "metric": "bp_edge_message_updates_per_sec", "value": v, "unit": "messages/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_msgs / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": bench_config(args.nx, args.ny, args.chi, int(os.environ.get("WORLD_SIZE", "1")), os.environ.get("TNQS_BENCH_HALO") != "nccl"),
            "gates": {"metric": "two_site_gate_applications_per_sec", "value": float(np.mean(gvals)), "unit": "gates/s"},
            "cpu_baseline": {"value": v, "unit": "messages/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "restated CPU oracle (NumPy + OpenBLAS, all host threads) - not Julia"},
            "e2e": {"value": v, "unit": "messages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--ny", type=int, default=32)
    ap.add_argument("--chi", type=int, default=16)
    ap.add_argument("--cpu-msgs", type=int, default=3000)
    ap.add_argument("--cpu-gates", type=int, default=200)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        out = run_reference(args)
        if out is not None:
            print(json.dumps(out))
        return
    out, ctx = run_engine(args)
    rank = int(os.environ.get("RANK", "0"))
    if not args.no_e2e and ctx is not None:
        out["e2e"] = run_e2e(ctx, max(1, min(args.steps, 6)))
    if rank == 0:
        if not args.no_cpu and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            mps, gps, t_msg, t_gate = cpu_sample(args.nx, args.ny, args.chi, args.cpu_msgs, args.cpu_gates)
            out["cpu_baseline"] = {"value": mps, "unit": "messages/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{args.cpu_msgs} message updates + {args.cpu_gates} two-site gates of the same C2 tensors ({t_msg + t_gate:.1f} s of CPU work)",
                                   "gates_per_s": gps, "note": "restated CPU oracle (NumPy + OpenBLAS, all host threads) - not Julia"}
        print(json.dumps(out))


if __name__ == "__main__":
    main()
