"""world_size-2 gloo test of the multi-rank host logic (partition, halo lists, sweep/exchange/all-reduce driver) with
the CPU oracle as the per-rank compute.  The CUDA engine takes the oracle's place on the GPU box (tests -m gpu, bench)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import tnqs_oracle as O
    from tensornetworkquantumsimulator_b200 import graphs as G, parallel as P
    g = G.named_grid((4, 6))
    chi = 3
    psi = O.random_tensornetworkstate(g, bond_dimension=chi, seed=5)
    owner = P.strip_owner(g, world, axis=1)
    lg = P.build_local_graph(g, owner, rank)
    ts = [psi.tensors[int(gv)] if lg.owned[lv] else np.zeros([2] + [chi] * lg.g.degree(lv), dtype=np.complex128) for lv, gv in enumerate(lg.l2g_v)]
    bpc = O.BeliefPropagationCache(O.TensorNetworkState(lg.g, ts))

    class OracleCompute(P.LocalCompute):
        def sweep(self, seq, want_diff):
            new = {int(k): O.updated_message(bpc, int(k)) for k in seq}
            s = sum(O.message_diff(new[k], O.message(bpc, k)) for k in new) if want_diff else 0.0
            bpc.messages.update(new)
            return s
        def pack(self, ks):
            return torch.from_numpy(np.stack([O.message(bpc, int(k)).ravel(order="F") for k in ks]))
        def unpack(self, ks, buf):
            for k, row in zip(ks, buf.numpy()):
                bpc.messages[int(k)] = row.reshape((chi, chi), order="F").copy()

    info = {}
    P.update_distributed(lg, OracleCompute(), dist, chi * chi, torch, torch.device("cpu"), maxiter=50, tolerance=1e-12, info=info)
    # global single-process Jacobi with the same stopping rule
    ref = O.BeliefPropagationCache(psi)
    for it in range(1, 51):
        new = {k: O.updated_message(ref, k) for k in range(2 * g.ne)}
        d = sum(O.message_diff(new[k], O.message(ref, k)) for k in new) / (2 * g.ne)
        ref.messages = new
        if d <= 1e-12:
            break
    err = 0.0
    for lk in range(2 * lg.g.ne):
        if lg.owned[lg.g.dsrc(lk)] or lg.owned[lg.g.ddst(lk)]:
            kg = 2 * int(lg.l2g_e[lk // 2]) + (lk & 1)
            err = max(err, float(np.linalg.norm(O.message(bpc, lk) - ref.messages[kg])))
    out[rank] = (info["iters"], it, err, info["halo_bytes"])
    dist.destroy_process_group()


def test_two_rank_halo_bp_matches_global_jacobi():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29541, out), nprocs=world, join=True)
    for r in range(world):
        iters, ref_iters, err, halo = out[r]
        assert iters == ref_iters
        assert err < 1e-13
        assert halo > 0


def test_partition_halo_lists_are_consistent():
    sys.path.insert(0, ROOT)
    from tensornetworkquantumsimulator_b200 import graphs as G, parallel as P
    g = G.named_grid((5, 8))
    owner = P.strip_owner(g, 4, axis=1)
    lgs = [P.build_local_graph(g, owner, r) for r in range(4)]
    for r, lg in enumerate(lgs):
        assert sorted(int(v) for lv, v in enumerate(lg.l2g_v) if lg.owned[lv]) == [v for v in range(g.nv) if owner[v] == r]
        for v in range(lg.g.nv):                        # canonical leg order of owned vertices is the global one
            if lg.owned[v]:
                assert [int(lg.l2g_e[e]) for e in lg.g.incident[v]] == g.incident[int(lg.l2g_v[v])]
        for nb, ks in lg.send.items():
            mine = [2 * int(lg.l2g_e[k // 2]) + (k & 1) for k in ks]
            theirs = [2 * int(lgs[nb].l2g_e[k // 2]) + (k & 1) for k in lgs[nb].recv[r]]
            assert mine == theirs and len(mine) > 0


@pytest.mark.parametrize("shape,world", [((4, 6), 2), ((3, 8), 4), ((5, 16), 8)])
def test_peer_memory_halo_layout_and_routes_are_consistent(shape, world):
    # host logic of the peer-memory halo (parallel.halo_layout / halo_routes): every owned -> ghost message of a rank is
    # routed to the slot of its neighbour's receive buffer that the neighbour reads back as the same global directed edge,
    # every slot is written exactly once, and nothing is routed to a rank that is not a neighbour
    sys.path.insert(0, ROOT)
    from tensornetworkquantumsimulator_b200 import graphs as G, parallel as P
    g = G.named_grid(shape)
    owner = P.strip_owner(g, world, axis=1)
    lgs = [P.build_local_graph(g, owner, r) for r in range(world)]
    layouts = [P.halo_layout(lg) for lg in lgs]

    def global_dedge(lg, lk):
        return 2 * int(lg.l2g_e[lk // 2]) + (lk & 1)

    written = [np.zeros(len(layouts[r][0]), dtype=int) for r in range(world)]
    for r, lg in enumerate(lgs):
        ks, peers, slots = P.halo_routes(lg, r, {p: layouts[p][1] for p in lg.send})
        assert len(ks) == sum(len(v) for v in lg.send.values())
        for k, p, s in zip(ks, peers, slots):
            assert p != r and p in lgs[r].send and r in lgs[p].recv
            assert lg.owned[lg.g.dsrc(k)] and not lg.owned[lg.g.ddst(k)]
            peer_k = int(layouts[p][0][s])                       # what the peer commits from that slot
            assert global_dedge(lg, k) == global_dedge(lgs[p], peer_k)
            assert lgs[p].owned[lgs[p].g.ddst(peer_k)] and not lgs[p].owned[lgs[p].g.dsrc(peer_k)]
            written[p][s] += 1
    for r in range(world):
        assert np.all(written[r] == 1), "every receive slot is written by exactly one peer message"
