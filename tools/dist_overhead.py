"""Per-sweep breakdown of the distributed JACOBI BP driver, NCCL halo exchange against the peer-memory halo, and the
cost of a gate class afterwards (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local)); torch.cuda.set_device(local)
from tensornetworkquantumsimulator_b200 import engine as E, lib, graphs as G, gates, parallel as P
chi = 16
gg = G.named_grid((32, 32 * world)); owner = P.strip_owner(gg, world, axis=1); lg = P.build_local_graph(gg, owner, rank)
groups = G.grid_edge_color(gg)
psi = E.random_tensornetworkstate(lg.g, bond_dimension=chi, seed=1 + rank, device=local)
bpc = E.BeliefPropagationCache(psi, _share=True)
ow = lib.i32(lg.owner_local); lib.check(psi._h.ptr, lib.load().tnqs_set_partition(psi._h.ptr, lib.ptr_i32(ow), rank))
dev = torch.device("cuda", local); comp = P.EngineCompute(bpc, dev); slot = chi * chi
seq = np.array([k for k in range(2 * lg.g.ne) if lg.owned[lg.g.dsrc(k)]], dtype=np.int32)
rzz = gates.gate_array(("Rzz", [0, 1], 0.25))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
batches = [P.class_gates_local(gg, lg, owner, rank, grp) for grp in groups]
def sync(): torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
def report(name, t):
    if rank == 0: print(name, {k: round(1e3 * v / 10, 3) for k, v in t.items()}, flush=True)
def gate_class():
    sync(); t0 = time.perf_counter()
    P.apply_class_distributed(E, bpc, gg, lg, owner, rank, groups[0], rzz, kw, dist, torch, dev, batch=batches[0])
    torch.cuda.synchronize(); return time.perf_counter() - t0
# ---- NCCL halo ----
t = {"sweep": 0.0, "halo": 0.0, "allreduce": 0.0}
for it in range(12):
    sync(); t0 = time.perf_counter()
    s = comp.sweep(seq, True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    P.halo_exchange(lg, comp, dist, slot, torch, dev)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    x = torch.tensor([s, float(len(seq))], dtype=torch.float64, device=dev); dist.all_reduce(x); m = float(x[0] / x[1])
    t3 = time.perf_counter()
    if it >= 2:
        t["sweep"] += t1 - t0; t["halo"] += t2 - t1; t["allreduce"] += t3 - t2
report("nccl halo, ms per sweep:", t)
g1 = [gate_class() for _ in range(4)]
if rank == 0: print("gate class before IPC (ms):", [round(1e3 * x, 1) for x in g1], flush=True)
# ---- peer-memory halo ----
t0 = time.perf_counter(); p2p = P.halo_setup_p2p(lg, comp, dist, rank, world); torch.cuda.synchronize()
if rank == 0: print("halo_setup_p2p %.1f ms, bytes to peers per sweep %d" % (1e3 * (time.perf_counter() - t0), p2p), flush=True)
ts = []
for it in range(4):
    sync(); t0 = time.perf_counter()
    comp.refresh(seq, 10, None)                      # one call = one graph launch: 10 sweeps incl. boundary stores, barrier, commit
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) / 10)
if rank == 0: print("peer-memory halo, device-driven refresh, ms per sweep:", [round(1e3 * x, 3) for x in ts], flush=True)
# short refreshes as the bench issues them (1-2 sweeps, convergence test on): fixed cost per call, then the bench's own sequence
# gate class -> refresh without a barrier in between, wall clock per rank
ts2 = []
for it in range(6):
    sync(); t0 = time.perf_counter()
    comp.refresh(seq, 2, None)
    torch.cuda.synchronize(); ts2.append(time.perf_counter() - t0)
print(f"rank {rank}: 2-sweep refresh, ms per call:", [round(1e3 * x, 2) for x in ts2], flush=True)
for cls in range(4):
    sync(); t0 = time.perf_counter()
    P.apply_class_distributed(E, bpc, gg, lg, owner, rank, groups[cls], rzz, kw, dist, torch, dev, batch=batches[cls])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    info = {}
    P.update_distributed(lg, comp, dist, slot, torch, dev, maxiter=25, tolerance=1e-8, info=info, p2p_bytes=p2p)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rank {rank}: class {cls}: gates {1e3 * (t1 - t0):.2f} ms, refresh {1e3 * (t2 - t1):.2f} ms ({info['iters']} sweeps)", flush=True)
it_, mean_ = comp.refresh(seq, 25, 1e-8)
if rank == 0: print("refresh to 1e-8:", it_, "sweeps, global mean diff", mean_, flush=True)
g2 = [gate_class() for _ in range(4)]
if rank == 0: print("gate class after IPC (ms):", [round(1e3 * x, 1) for x in g2], flush=True)
dist.destroy_process_group()
