"""Per-sweep breakdown of the distributed JACOBI BP driver (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local)); torch.cuda.set_device(local)
from tensornetworkquantumsimulator_b200 import engine as E, lib, graphs as G, parallel as P
chi = 16
gg = G.named_grid((32, 32 * world)); owner = P.strip_owner(gg, world, axis=1); lg = P.build_local_graph(gg, owner, rank)
psi = E.random_tensornetworkstate(lg.g, bond_dimension=chi, seed=1 + rank, device=local)
bpc = E.BeliefPropagationCache(psi)
ow = lib.i32(lg.owner_local); lib.check(psi._h.ptr, lib.load().tnqs_set_partition(psi._h.ptr, lib.ptr_i32(ow), rank))
dev = torch.device("cuda", local); comp = P.EngineCompute(bpc, dev); slot = chi * chi
seq = np.array([k for k in range(2 * lg.g.ne) if lg.owned[lg.g.dsrc(k)]], dtype=np.int32)
t = {"sweep": 0.0, "halo": 0.0, "allreduce": 0.0}
for it in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = comp.sweep(seq, True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    P.halo_exchange(lg, comp, dist, slot, torch, dev)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    x = torch.tensor([s, float(len(seq))], dtype=torch.float64, device=dev); dist.all_reduce(x); m = float(x[0] / x[1])
    t3 = time.perf_counter()
    if it >= 2:
        t["sweep"] += t1 - t0; t["halo"] += t2 - t1; t["allreduce"] += t3 - t2
if rank == 0:
    print({k: round(1e3 * v / 10, 3) for k, v in t.items()}, "ms per sweep; messages", len(seq))
dist.destroy_process_group()
