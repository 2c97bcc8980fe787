"""Generates tests/golden/*.npz from the CPU oracle (run in the build container: python tests/golden/make_golden.py).
The fixtures pin the oracle against drift and give the GPU tests reference-independent targets.  They are NOT Julia
outputs (parity unpinned); julia/dump_golden.jl produces the same files from the real reference."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200 import graphs as G, gates
from util import ising_layer

HERE = os.path.dirname(os.path.abspath(__file__))


def bp_case():
    g = G.named_grid((3, 3))
    psi = O.random_tensornetworkstate(g, bond_dimension=3, seed=11)
    seq = G.forest_cover_edge_sequence(g)
    info = {}
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=3, tolerance=0.0, edge_sequence=seq, info=info)
    msgs = np.stack([O.message(bpc, k) for k in range(2 * g.ne)])
    z = np.array([O.expect(bpc, ("Z", [g.names[v]])) for v in range(g.nv)])
    np.savez(os.path.join(HERE, "bp_grid3x3_chi3.npz"), seed=11, chi=3, seq=seq, messages=msgs, expect_z=z, mean_diff=info["mean_diff"])


def gate_case():
    g = G.named_grid((4, 4))
    psi = O.random_tensornetworkstate(g, bond_dimension=4, seed=31)
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=6)
    grp = G.grid_edge_color(g)[0]
    arr = gates.gate_array(("Rzz", [0, 1], 0.9))
    ranks, svals, errs = [], [], []
    for e in grp:
        det = {}
        O.apply_gate(bpc, arr, [int(g.esrc[e]), int(g.edst[e])], maxdim=4, cutoff=1e-10, detail=det)
        ranks.append(det["rank"]); errs.append(det["truncerr"]); svals.append(np.pad(det["svals"], (0, 4 - det["rank"])))
    np.savez(os.path.join(HERE, "gates_grid4x4_chi4.npz"), seed=31, chi=4, edges=np.array(grp), gate=arr, ranks=np.array(ranks),
             svals=np.array(svals), truncerr=np.array(errs))


def circuit_case():
    g = G.named_grid((4, 4))
    layer = ising_layer(g, dt=0.01, hx=2.5, hz=0.2, J=1.0, groups=G.grid_edge_color(g))
    psi, errs = O.apply_gates(layer * 6, O.tensornetworkstate(lambda v: "↑", g), apply_kwargs=dict(maxdim=4, cutoff=1e-10, normalize_tensors=True))
    z = O.expect_tns(psi, ("Z", [(2, 2)]))
    zz = O.expect_tns(psi, ("ZZ", [(2, 2), (2, 3)]))
    np.savez(os.path.join(HERE, "circuit_c1_4x4.npz"), bond_dims=psi.bond_dims(), expect_z=z, expect_zz=zz, errs=errs)


def bp_tensor_core_case():
    # uniform chi = 8 grid: the shape the FP64 tensor-core kernels (k_bp_sweep / k_absorb_slab / k_close_dmma) carry, plus
    # the BP inner product of two states on it (bilinear form)
    g = G.named_grid((4, 4))
    psi = O.random_tensornetworkstate(g, bond_dimension=8, seed=51)
    phi = O.random_tensornetworkstate(g, bond_dimension=8, seed=52)
    for s in (psi, phi):
        for v in range(g.nv):
            s.tensors[v] /= np.linalg.norm(s.tensors[v])
    seq = G.forest_cover_edge_sequence(g)
    info = {}
    bpc = O.update(O.BeliefPropagationCache(psi), maxiter=4, tolerance=0.0, edge_sequence=seq, info=info)
    msgs = np.stack([O.message(bpc, k) for k in range(2 * g.ne)])
    z = np.array([O.expect(bpc, ("Z", [g.names[v]])) for v in range(g.nv)])
    bj = O.BeliefPropagationCache(psi)
    for _ in range(2):
        bj.messages = {k: O.updated_message(bj, k) for k in range(2 * g.ne)}
    bb = O.update(O.BeliefPropagationCache(psi.copy(), {}, phi), maxiter=4, edge_sequence=seq)
    np.savez(os.path.join(HERE, "bp_grid4x4_chi8.npz"), seed_psi=51, seed_phi=52, chi=8, seq=seq, messages=msgs, expect_z=z, mean_diff=info["mean_diff"],
             jacobi_messages=np.stack([O.message(bj, k) for k in range(2 * g.ne)]),
             bilinear_messages=np.stack([O.message(bb, k) for k in range(2 * g.ne)]), bilinear_z=O.partitionfunction(bb))


def heisenberg_layer(g, dt=0.04, h=-1.0, J=-1.0):
    """examples/2dIsing_dynamics_Heisenbergpicture.jl:26-35."""
    layer = [("Rz", [g.names[v]], h * dt) for v in range(g.nv)]
    for grp in G.grid_edge_color(g):
        layer += [("Rxx", [g.names[int(g.esrc[e])], g.names[int(g.edst[e])]], 2 * J * dt) for e in grp]
    layer += [("Rz", [g.names[v]], h * dt) for v in range(g.nv)]
    return layer


def heisenberg_case():
    # the Heisenberg-picture example on its own 4x4 grid (Pauli sites d = 4, transfer-matrix gates, maxdim 4, cutoff 1e-14, no
    # tensor normalisation, rescale after every layer, tr(O(t) O(0)) / 2^n through BP on the bilinear form)
    g = G.named_grid((4, 4))
    o0 = O.paulitensornetworkstate(lambda v: "Z" if v == (2, 2) else "I", g)
    layer = heisenberg_layer(g)
    bpc = O.BeliefPropagationCache(o0)
    dims, trs, zs, errs = [], [], [], []
    for _ in range(3):
        bpc, e = O.apply_gates(layer, bpc, apply_kwargs=dict(maxdim=4, cutoff=1e-14, normalize_tensors=False))
        bpc = O.rescale(bpc)
        dims.append(bpc.network.bond_dims()); errs.append(e)
        zs.append(O.partitionfunction(bpc)); trs.append(O.inner(bpc.network, o0))
    np.savez(os.path.join(HERE, "heisenberg_grid4x4.npz"), bond_dims=np.array(dims), trace_ot_o0=np.array(trs), frobenius=np.array(zs),
             errs=np.array(errs))


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (bp_case, gate_case, circuit_case, bp_tensor_core_case, heisenberg_case):
        if not only or fn.__name__ in only:
            fn()
    print(sorted(os.listdir(HERE)))
