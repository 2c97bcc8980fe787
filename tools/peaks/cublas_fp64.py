"""cuBLAS DGEMM / ZGEMM throughput on this B200 (roofline denominator for the FP64-bound paths).
Burst = best of 10, sustained = back to back for ~3 s.  CUDA events, 3 warm-ups.  Prints one JSON line."""
import json, torch
def run(dtype, n, flop_per_mac):
    a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3): a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = flop_per_mac * n ** 3
    reps = max(1, int(3000 / best))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): a @ b
    e1.record(); torch.cuda.synchronize()
    return fl / best * 1e-9, fl * reps / e0.elapsed_time(e1) * 1e-9
out = {}
for name, dt, n, f in [("dgemm", torch.float64, 8192, 2), ("zgemm", torch.complex128, 4096, 8), ("zgemm_8192", torch.complex128, 8192, 8)]:
    b, s = run(dt, n, f); out[name + "_tflops"] = round(b, 3); out[name + "_tflops_sustained"] = round(s, 3)
print(json.dumps(out))
