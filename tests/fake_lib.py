"""TEST INFRASTRUCTURE ONLY: a stand-in for the ctypes handle of libtnqs_cuda.so whose entry points are answered by the CPU oracle,
so that the HOST logic of tensornetworkquantumsimulator_b200/engine.py (argument packing, column-major flattening, the order of
calls in composed operations such as inner / _expect_tree / reduced_density_matrix) can be executed without a GPU.  It implements
only the entry points those host paths use, with the semantics include/tnqs_cuda.h documents.  It is installed by a pytest
fixture (monkeypatching lib._lib) and never ships: the product has no CPU fallback."""
import ctypes as C

import numpy as np

from oracle import tnqs_oracle as O
from tensornetworkquantumsimulator_b200.graphs import Graph


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).view(dtype) if dtype != np.complex128 else \
        np.ctypeslib.as_array(ptr, shape=(2 * int(n),)).view(np.complex128)


def _hid(h):
    if h is None:
        return None
    return h.value if isinstance(h, C.c_void_p) else int(h)


class _State:
    def __init__(self, g, tensors, chi_cap):
        self.g, self.chi_cap = g, chi_cap
        self.bpc = O.BeliefPropagationCache(O.TensorNetworkState(g, tensors))
        self.bra = None


class FakeLib:
    def __init__(self):
        self.h = {}
        self.next = 1
        self.calls = []
        self.issued = []          # the c_void_p objects handed to the host: nulled by close() so that no handle of the fake
                                  # ever reaches the real tnqs_destroy after the fixture is gone

    def close(self):
        for p in self.issued:
            p.value = None
        self.h.clear()

    # ---- helpers ----
    def _s(self, h) -> _State:
        return self.h[_hid(h)]

    def _bpc(self, st: _State) -> O.BeliefPropagationCache:
        st.bpc.bra = self._s(st.bra).bpc.network if st.bra is not None else None
        return st.bpc

    def __getattr__(self, name):
        raise AttributeError(f"FakeLib does not implement {name} (tests/fake_lib.py covers the host paths of the CPU tests only)")

    # ---- container ----
    def tnqs_version(self):
        return b"tnqs-cuda fake (oracle-backed, tests only)"

    def tnqs_status_string(self, st):
        return b"fake status"

    def tnqs_last_error(self, h):
        return b"fake error"

    def tnqs_create(self, out, nv, ne, esrc, edst, d, chi, chi_cap, device):
        self.calls.append("tnqs_create")
        es, ed = _arr(esrc, ne, np.int32).copy(), _arr(edst, ne, np.int32).copy()
        dd, cc = _arr(d, nv, np.int32).copy(), _arr(chi, max(ne, 1), np.int32).copy()
        g = Graph(list(range(nv)), es, ed)
        ts = [np.zeros((int(dd[v]),) + tuple(int(cc[e]) for e in g.incident[v]), dtype=np.complex128) for v in range(nv)]
        assert all(int(c) <= chi_cap for c in cc[:ne]), "bond dimension above chi_cap"
        self.h[self.next] = _State(g, ts, chi_cap)
        out._obj.value = self.next
        self.issued.append(out._obj)
        self.next += 1
        return 0

    def tnqs_destroy(self, h):
        self.h.pop(_hid(h), None)
        return 0

    def tnqs_clone(self, h, out):
        self.calls.append("tnqs_clone")
        st = self._s(h)
        new = _State(st.g, [t.copy() for t in st.bpc.network.tensors], st.chi_cap)
        new.bpc.messages = {k: m.copy() for k, m in st.bpc.messages.items()}
        new.bra = st.bra
        self.h[self.next] = new
        out._obj.value = self.next
        self.issued.append(out._obj)
        self.next += 1
        return 0

    def tnqs_site_size(self, h, v):
        return int(self._s(h).bpc.network.tensors[int(v)].size)

    def tnqs_get_bond_dims(self, h, chi):
        st = self._s(h)
        if st.g.ne:
            _arr(chi, st.g.ne, np.int32)[:] = st.bpc.network.bond_dims()
        return 0

    def tnqs_set_site(self, h, v, data, n):
        st = self._s(h)
        t = st.bpc.network.tensors[int(v)]
        assert int(n) == t.size
        st.bpc.network.tensors[int(v)] = _arr(data, n, np.complex128).reshape(t.shape, order="F").copy()
        return 0

    def tnqs_get_site(self, h, v, data, n):
        t = self._s(h).bpc.network.tensors[int(v)]
        assert int(n) == t.size
        _arr(data, n, np.complex128)[:] = t.ravel(order="F")
        return 0

    def tnqs_set_sites(self, h, count, v, data, n):
        self.calls.append("tnqs_set_sites")
        vs, ns = _arr(v, count, np.int32), _arr(n, count, np.int64)
        for i in range(count):
            self.tnqs_set_site(h, int(vs[i]), data[i], int(ns[i]))
        return 0

    def tnqs_get_sites(self, h, count, v, data, n):
        vs, ns = _arr(v, count, np.int32), _arr(n, count, np.int64)
        for i in range(count):
            self.tnqs_get_site(h, int(vs[i]), data[i], int(ns[i]))
        return 0

    # ---- messages ----
    def tnqs_set_message(self, h, k, data, n):
        self.calls.append("tnqs_set_message")
        st = self._s(h)
        chi = st.bpc.network.bond_dim(int(k) // 2)
        assert int(n) == chi * chi
        st.bpc.messages[int(k)] = _arr(data, n, np.complex128).reshape((chi, chi), order="F").copy()
        return 0

    def tnqs_get_message(self, h, k, data, n):
        st = self._s(h)
        _arr(data, n, np.complex128)[:] = O.message(self._bpc(st), int(k)).ravel(order="F")
        return 0

    def tnqs_reset_messages(self, h):
        self._s(h).bpc.messages = {}
        return 0

    # ---- BP ----
    def tnqs_bp_update(self, h, seq, nseq, opts, iters, mean):
        self.calls.append("tnqs_bp_update")
        st, o = self._s(h), opts._obj
        assert o.mode == 0, "the fake runs REFERENCE_ORDER only"
        info = {}
        new = O.update(self._bpc(st), maxiter=int(o.maxiter), tolerance=None if o.tolerance < 0 else float(o.tolerance),
                       edge_sequence=_arr(seq, nseq, np.int32).tolist(), normalize=bool(o.normalize),
                       enforce_hermiticity=bool(o.hermitize), info=info)
        st.bpc.messages = new.messages
        iters._obj.value = int(info["iters"])
        mean._obj.value = float(info["mean_diff"])
        return 0

    def tnqs_set_bra(self, h, bra):
        self.calls.append("tnqs_set_bra")
        st = self._s(h)
        b = _hid(bra)
        if b is not None:
            other = self._s(bra)
            assert b != _hid(h) and list(other.bpc.network.bond_dims()) == list(st.bpc.network.bond_dims()), "TNQS_ERR_SHAPE"
            assert [t.shape[0] for t in other.bpc.network.tensors] == [t.shape[0] for t in st.bpc.network.tensors]
        st.bra = b
        return 0

    def tnqs_vertex_scalars(self, h, out):
        st = self._s(h)
        _arr(out, st.g.nv, np.complex128)[:] = [O.vertex_scalar(self._bpc(st), v) for v in range(st.g.nv)]
        return 0

    def tnqs_edge_scalars(self, h, out):
        st = self._s(h)
        if st.g.ne:
            _arr(out, st.g.ne, np.complex128)[:] = [O.edge_scalar(self._bpc(st), e) for e in range(st.g.ne)]
        return 0

    def tnqs_rescale(self, h):
        self.calls.append("tnqs_rescale")
        st = self._s(h)
        new = O.rescale(self._bpc(st))
        st.bpc.network.tensors, st.bpc.messages = new.network.tensors, new.messages
        return 0

    # ---- gates ----
    def tnqs_apply_1site(self, h, n, v, gates, normalize):
        self.calls.append("tnqs_apply_1site")
        st = self._s(h)
        vs = _arr(v, n, np.int32)
        sizes = [st.bpc.network.tensors[int(x)].shape[0] ** 2 for x in vs]
        flat = _arr(gates, sum(sizes), np.complex128)
        pos = 0
        for x, sz in zip(vs, sizes):
            d = st.bpc.network.tensors[int(x)].shape[0]
            O.apply_gate(st.bpc, flat[pos:pos + sz].reshape((d, d), order="F"), [int(x)], normalize_tensors=bool(normalize))
            pos += sz
        return 0

    def tnqs_apply_circuit(self, h, ngates, gate_v, gate_off, gate_data, aopts, seq, nseq, bopts, update_cache, truncerr, refreshes):
        self.calls.append("tnqs_apply_circuit")
        st, a, b = self._s(h), aopts._obj, bopts._obj
        gv, off = _arr(gate_v, 2 * ngates, np.int32), _arr(gate_off, ngates, np.int64)
        circuit = []
        for i in range(ngates):
            vs = [int(gv[2 * i])] + ([int(gv[2 * i + 1])] if gv[2 * i + 1] >= 0 else [])
            dims = [st.bpc.network.tensors[v].shape[0] for v in vs]
            n = int(np.prod(dims)) ** 2
            base = np.ctypeslib.as_array(gate_data, shape=(2 * (int(off[i]) + n),)).view(np.complex128)
            circuit.append((base[int(off[i]):int(off[i]) + n].reshape(tuple(dims) * 2, order="F").copy(), vs))
        info = {}
        kw = dict(maxiter=int(b.maxiter), tolerance=None if b.tolerance < 0 else float(b.tolerance),
                  edge_sequence=_arr(seq, nseq, np.int32).tolist())
        new, errs = O.apply_gates(circuit, self._bpc(st), dict(maxdim=None if a.maxdim <= 0 else int(a.maxdim),
                                                                cutoff=None if np.isnan(a.cutoff) else float(a.cutoff),
                                                                normalize_tensors=bool(a.normalize_tensors)), kw, bool(update_cache), info)
        st.bpc.network.tensors, st.bpc.messages = new.network.tensors, new.messages
        assert st.bpc.network.maxvirtualdim() <= st.chi_cap, "TNQS_ERR_SHAPE: bond dimension above chi_cap"
        if ngates:
            _arr(truncerr, ngates, np.float64)[:] = errs
        refreshes._obj.value = int(info["refreshes"])
        return 0

    def tnqs_apply_2site(self, h, n, e, flip, gates, opts, ranks, truncerr, svals):
        self.calls.append("tnqs_apply_2site")
        st, a = self._s(h), opts._obj
        es, fl = _arr(e, n, np.int32), _arr(flip, n, np.int32)
        touched = [int(st.g.esrc[x]) for x in es] + [int(st.g.edst[x]) for x in es]
        if len(set(touched)) != len(touched):
            return -3                                              # TNQS_ERR_OVERLAPPING_BATCH
        pos = 0
        for i in range(n):
            v1, v2 = int(st.g.esrc[es[i]]), int(st.g.edst[es[i]])
            if fl[i]:
                v1, v2 = v2, v1
            d1, d2 = st.bpc.network.tensors[v1].shape[0], st.bpc.network.tensors[v2].shape[0]
            sz = (d1 * d2) ** 2
            flat = np.ctypeslib.as_array(gates, shape=(2 * (pos + sz),)).view(np.complex128)[pos:pos + sz]
            pos += sz
            det = {}
            err = O.apply_gate(st.bpc, flat.reshape((d1, d2, d1, d2), order="F").copy(), [v1, v2],
                               maxdim=None if a.maxdim <= 0 else int(a.maxdim), cutoff=None if np.isnan(a.cutoff) else float(a.cutoff),
                               normalize_tensors=bool(a.normalize_tensors), detail=det)
            if ranks:
                _arr(ranks, n, np.int32)[i] = det["rank"]
            if truncerr:
                _arr(truncerr, n, np.float64)[i] = err
            if svals:
                _arr(svals, n * st.chi_cap, np.float64)[i * st.chi_cap:i * st.chi_cap + det["rank"]] = det["svals"]
        return 0

    def tnqs_truncate(self, h, ngroups, group_off, group_edges, aopts, seq, nseq, bopts):
        self.calls.append("tnqs_truncate")
        st, a, b = self._s(h), aopts._obj, bopts._obj
        off = _arr(group_off, ngroups + 1, np.int32)
        ge = _arr(group_edges, int(off[ngroups]), np.int32)
        groups = [[int(x) for x in ge[off[i]:off[i + 1]]] for i in range(ngroups)]
        kw = dict(maxiter=int(b.maxiter), tolerance=None if b.tolerance < 0 else float(b.tolerance),
                  edge_sequence=_arr(seq, nseq, np.int32).tolist())
        new = O.truncate_bp(self._bpc(st), maxdim=int(a.maxdim), cutoff=None if np.isnan(a.cutoff) else float(a.cutoff),
                            normalize_tensors=bool(a.normalize_tensors), edge_groups=groups, bp_update_kwargs=kw)
        st.bpc.network.tensors, st.bpc.messages = new.network.tensors, new.messages
        return 0

    def tnqs_symmetric_gauge(self, h, regularization):
        self.calls.append("tnqs_symmetric_gauge")
        st = self._s(h)
        new = O.symmetric_gauge(self._bpc(st), float(regularization))
        st.bpc.network.tensors, st.bpc.messages = new.network.tensors, new.messages
        return 0

    # ---- measurements ----
    def tnqs_rdm_1site(self, h, n, v, rho):
        st = self._s(h)
        dmax = max(t.shape[0] for t in st.bpc.network.tensors)
        out = _arr(rho, n * dmax * dmax, np.complex128).reshape((n, dmax, dmax))
        out[:] = 0
        for i, x in enumerate(_arr(v, n, np.int32)):
            r = O.rdm_1site(self._bpc(st), int(x))                    # [s, s'] -> column-major block: element (s, s') at s + dmax s'
            out[i, :r.shape[1], :r.shape[0]] = r.T
        return 0

    def tnqs_expect_path(self, h, n, path_off, path_v, ops, out):
        self.calls.append("tnqs_expect_path")
        st = self._s(h)
        offs = _arr(path_off, n + 1, np.int32)
        pv = _arr(path_v, int(offs[n]), np.int32)
        res = _arr(out, 4 * n, np.float64)
        pos = 0
        for i in range(n):
            path = [int(x) for x in pv[offs[i]:offs[i + 1]]]
            d1, d2 = st.bpc.network.tensors[path[0]].shape[0], st.bpc.network.tensors[path[-1]].shape[0]
            flat = np.ctypeslib.as_array(ops, shape=(2 * (pos + d1 * d1 + d2 * d2),)).view(np.complex128)
            o1 = flat[pos:pos + d1 * d1].reshape((d1, d1), order="F")
            o2 = flat[pos + d1 * d1:pos + d1 * d1 + d2 * d2].reshape((d2, d2), order="F")
            pos += d1 * d1 + d2 * d2
            num = O._region_contract(self._bpc(st), path, {path[0]: o1, path[-1]: o2})
            den = O._region_contract(self._bpc(st), path, {})
            res[4 * i:4 * i + 4] = [num.real, num.imag, den.real, den.imag]
        return 0

    def tnqs_loop_weights(self, h, nsub, sub_off, sub_edges, weights):
        self.calls.append("tnqs_loop_weights")
        st = self._s(h)
        offs = _arr(sub_off, nsub + 1, np.int32)
        es = _arr(sub_edges, int(offs[nsub]), np.int32)
        res = _arr(weights, nsub, np.complex128)
        for i in range(nsub):
            res[i] = O.loop_weight(self._bpc(st), [int(e) for e in es[offs[i]:offs[i + 1]]])
        return 0

    def tnqs_loop_transfer(self, h, length, walk_v, out):
        self.calls.append("tnqs_loop_transfer")
        st = self._s(h)
        g = st.g
        vs = [int(x) for x in _arr(walk_v, length, np.int32)]
        ks = [g.dedge(vs[i], vs[(i + 1) % length]) for i in range(length)]
        m = O.loop_transfer_matrix(self._bpc(st), ks[:-1], ks[-1])
        _arr(out, m.size, np.complex128)[:] = m.reshape(-1, order="F")
        return 0

    def tnqs_kernel_launches(self, h):
        return 0
