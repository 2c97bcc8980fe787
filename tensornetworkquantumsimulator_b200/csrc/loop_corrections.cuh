// Loop corrections to the BP partition function (src/MessagePassing/loopcorrection.jl) and the loop transfer matrices of
// loop_correlation (src/MessagePassing/beliefpropagationcache.jl:144-179) on the BP message-update machinery.  Included by
// tnqs_cuda.cu (needs build_plan / run_plan / scratch_budget).
//
// weight(eg) (loopcorrection.jl:81-91) is the contraction of the factors of the subgraph eg with the messages that enter it from
// outside and, on each of its edges e = (a, b), the antiprojector  P_e = 1 - |m_{a->b}) (m_{b->a}|  over the doubled bond.  Here it
// is evaluated as a sum of TREE contractions, each of which is a run of ordinary message updates with substituted messages:
//   * a spanning tree of eg (BFS from its smallest vertex) is contracted from the leaves to the root; the environment a vertex
//     hands to its parent is its message update with the environments of its children in place of their messages, followed by
//     P_e:  X <- X - m_{child->parent} (m_{parent->child} | X)                                              (k_antiproject);
//   * each of the c = |E| - |V| + 1 edges left over is cut:  P_e = sum_rho |rho)(rho|  -  |m_{a->b})(m_{b->a}|  with rho running over
//     the chi^2 matrix units - chi^2 terms that put the same matrix unit on both ends, and one term (sign -1) that puts the BP
//     messages there, i.e. treats the edge as if it were outside eg.
// A vertex is evaluated once per assignment of the cut edges that touch its subtree, not once per assignment of all of them.  The
// cost is prod_cuts (chi^2 + 1) root contractions: meant for the small bond dimensions loop corrections are used at (the
// reference's example runs chi = 3); a request that would need more than LOOP_MAX_TERMS evaluations is refused.
#pragma once

static const long long LOOP_MAX_TERMS = 1ll << 22;

struct LoopTree {
  std::vector<int> verts, parent, parent_edge, depth;      // BFS order; parent = index into verts (-1 at the root)
  std::vector<int> cuts;                                   // edge ids left out of the spanning tree
  std::vector<std::vector<int>> cutset;                    // per vertex: ascending indices into `cuts` that touch its subtree
  std::vector<long long> ninst, base;                      // per vertex: evaluations and offset of its first output slot
  int sub = -1;
};

// option of cut j: 0 .. chi^2 - 1 = matrix unit, chi^2 = the BP-message term
static long long loop_radix(const tnqs_ctx* c, int e) { return (long long)c->chi[e] * c->chi[e] + 1; }

static void loop_trees_of(const tnqs_ctx* c, const int32_t* edges, int n, int sub, std::vector<LoopTree>& out) {
  std::set<int> es(edges, edges + n);
  REQUIRE((int)es.size() == n, TNQS_ERR_INVALID_ARG, "duplicate edge in a subgraph");
  for (int e : es) REQUIRE(e >= 0 && e < c->ne, TNQS_ERR_INVALID_ARG, "subgraph edge out of range");
  std::set<int> vs;
  for (int e : es) { vs.insert(c->esrc[e]); vs.insert(c->edst[e]); }
  std::set<int> seen;
  for (int root : vs) {
    if (seen.count(root)) continue;
    LoopTree t; t.sub = sub;
    std::map<int, int> idx;
    std::set<int> tree_edges, comp_edges;
    t.verts.push_back(root); t.parent.push_back(-1); t.parent_edge.push_back(-1); t.depth.push_back(0);
    idx[root] = 0; seen.insert(root);
    for (size_t h = 0; h < t.verts.size(); ++h) {
      const int v = t.verts[h];
      for (int e : c->inc[v]) {
        if (!es.count(e)) continue;
        comp_edges.insert(e);
        const int u = c->other(e, v);
        if (idx.count(u)) continue;
        idx[u] = (int)t.verts.size(); seen.insert(u);
        t.verts.push_back(u); t.parent.push_back((int)h); t.parent_edge.push_back(e); t.depth.push_back(t.depth[h] + 1);
        tree_edges.insert(e);
      }
    }
    for (int e : comp_edges) if (!tree_edges.count(e)) t.cuts.push_back(e);
    const int nvt = (int)t.verts.size();
    t.cutset.assign(nvt, {});
    for (size_t j = 0; j < t.cuts.size(); ++j)
      for (int end : {c->esrc[t.cuts[j]], c->edst[t.cuts[j]]})
        for (int i = idx[end]; i >= 0; i = t.parent[i])
          if (std::find(t.cutset[i].begin(), t.cutset[i].end(), (int)j) == t.cutset[i].end()) t.cutset[i].push_back((int)j);
    t.ninst.assign(nvt, 1); t.base.assign(nvt, 0);
    for (int i = 0; i < nvt; ++i) {
      std::sort(t.cutset[i].begin(), t.cutset[i].end());
      for (int j : t.cutset[i]) {
        t.ninst[i] *= loop_radix(c, t.cuts[j]);
        REQUIRE(t.ninst[i] <= LOOP_MAX_TERMS, TNQS_ERR_UNSUPPORTED,
                "loop weights: a subgraph needs more than 2^22 tree contractions (prod over its independent cycles of chi^2 + 1)");
      }
    }
    out.push_back(std::move(t));
  }
}

// One batch of trees: every (vertex, assignment) is a chain; level = depth from the deepest vertex of the batch.
static void loop_run_batch(tnqs_ctx* c, const std::vector<LoopTree*>& trees, const std::map<int, const cplx*>& units,
                           std::vector<std::vector<cplx>>& root_values) {
  // output slots: the roots of all trees first (one read-back), then everything else
  long long total = 0; int maxdepth = 0;
  for (LoopTree* t : trees) { t->base[0] = total; total += t->ninst[0]; }
  const long long nroot = total;
  for (LoopTree* t : trees) for (size_t i = 1; i < t->verts.size(); ++i) { t->base[i] = total; total += t->ninst[i]; maxdepth = std::max(maxdepth, t->depth[i]); }
  const int dd = c->dmax * c->dmax;
  const size_t oslot = (size_t)std::max<long long>(c->slot, dd);
  cplx* outbuf = c->talloc<cplx>((size_t)total * oslot);
  std::vector<std::vector<Chain>> levels(maxdepth + 1);
  std::vector<std::vector<AntiProjDesc>> aps(maxdepth + 1);
  for (LoopTree* t : trees) {
    const int nvt = (int)t->verts.size(), ncut = (int)t->cuts.size();
    std::vector<std::vector<int>> children(nvt);
    for (int i = 1; i < nvt; ++i) children[t->parent[i]].push_back(i);
    std::vector<long long> opt(ncut, 0);
    for (int i = 0; i < nvt; ++i) {
      const int v = t->verts[i];
      for (long long inst = 0; inst < t->ninst[i]; ++inst) {
        long long r = inst;
        for (int j : t->cutset[i]) { const long long rad = loop_radix(c, t->cuts[j]); opt[j] = r % rad; r /= rad; }
        Chain ch; ch.v = v;
        for (size_t l = 0; l < c->inc[v].size(); ++l) {
          const int e = c->inc[v][l], u = c->other(e, v);
          if (e == t->parent_edge[i]) continue;
          Absorb ab; ab.leg = (int)l; ab.ldm = c->chi[e]; ab.flags = 0; ab.mat = c->msg(c->dedge(u, e));
          for (int ci : children[i]) if (t->parent_edge[ci] == e) {
            long long cinst = 0, mul = 1;
            for (int j : t->cutset[ci]) { cinst += opt[j] * mul; mul *= loop_radix(c, t->cuts[j]); }
            ab.mat = outbuf + (size_t)(t->base[ci] + cinst) * oslot;
          }
          for (int j = 0; j < ncut; ++j) if (t->cuts[j] == e) {
            const long long n2 = (long long)c->chi[e] * c->chi[e];
            if (opt[j] < n2) ab.mat = units.at(c->chi[e]) + (size_t)opt[j] * n2;
          }
          ch.absorb.push_back(ab);
        }
        cplx* dst = outbuf + (size_t)(t->base[i] + inst) * oslot;
        ch.dst = dst;
        const int lvl = maxdepth - t->depth[i];
        if (i == 0) ch.close_leg = -1;
        else {
          const int e = t->parent_edge[i], p = t->verts[t->parent[i]];
          ch.close_leg = c->leg(v, e);
          AntiProjDesc ap; ap.x = dst; ap.mf = c->msg(c->dedge(v, e)); ap.mb = c->msg(c->dedge(p, e)); ap.nn = c->chi[e] * c->chi[e];
          aps[lvl].push_back(ap);
        }
        levels[lvl].push_back(std::move(ch));
      }
    }
  }
  const size_t budget = std::max<size_t>(scratch_budget(c) / (size_t)(maxdepth + 1), (size_t)64 << 20);
  for (int lvl = 0; lvl <= maxdepth; ++lvl) {
    if (levels[lvl].empty()) continue;
    std::vector<std::vector<Chain>> one(1);
    one[0].swap(levels[lvl]);
    Plan p = build_plan(c, one, budget);
    run_plan(c, p);
    if (!aps[lvl].empty()) {
      AntiProjDesc* d = c->upload(aps[lvl]);
      LAUNCH(c, TNQS_KC_MISC, k_antiproject, (unsigned)aps[lvl].size(), 128, 0, d);
    }
  }
  // root values: trace of the d x d block of every root evaluation
  std::vector<cplx> h((size_t)nroot * oslot);
  CU(cudaMemcpyAsync(h.data(), outbuf, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  for (size_t ti = 0; ti < trees.size(); ++ti) {
    const LoopTree* t = trees[ti];
    const int d = c->d[t->verts[0]];
    std::vector<cplx> vals((size_t)t->ninst[0]);
    for (long long inst = 0; inst < t->ninst[0]; ++inst) {
      double re = 0, im = 0;
      for (int s = 0; s < d; ++s) { const cplx x = h[(size_t)(t->base[0] + inst) * oslot + s + s * d]; re += x.x; im += x.y; }
      vals[(size_t)inst] = cmake(re, im);
    }
    root_values[ti] = std::move(vals);
  }
}

// chi x chi matrix units of one bond dimension, stacked = the identity of size chi^2
static const cplx* loop_units(tnqs_ctx* c, int chi, std::map<int, const cplx*>& bank) {
  auto it = bank.find(chi);
  if (it != bank.end()) return it->second;
  const size_t n2 = (size_t)chi * chi;
  cplx* u = c->talloc<cplx>(n2 * n2);
  CU(cudaMemsetAsync(u, 0, n2 * n2 * sizeof(cplx), c->stream));
  LAUNCH(c, TNQS_KC_MISC, k_unit_diagonal, (unsigned)((n2 + 255) / 256), 256, 0, u, (long long)n2);
  bank[chi] = u;
  return u;
}

static void loop_weights_impl(tnqs_ctx* c, int nsub, const int32_t* sub_off, const int32_t* sub_edges, double* out) {
  REQUIRE(!c->partitioned, TNQS_ERR_UNSUPPORTED, "loop weights on a partitioned handle");
  std::vector<LoopTree> trees;
  for (int i = 0; i < nsub; ++i) {
    const int n = sub_off[i + 1] - sub_off[i];
    REQUIRE(n >= 1, TNQS_ERR_INVALID_ARG, "empty subgraph");
    loop_trees_of(c, sub_edges + sub_off[i], n, i, trees);
  }
  // batches bounded by the memory of the environments (a quarter of the scratch budget) and by the descriptor count
  const size_t oslot = (size_t)std::max<long long>(c->slot, (long long)c->dmax * c->dmax);
  const long long cap = std::max<long long>(1, std::min<long long>(1 << 20, (long long)(scratch_budget(c) / 4 / (oslot * sizeof(cplx)))));
  std::vector<cplx> w((size_t)nsub, cmake(1, 0));
  size_t i0 = 0;
  while (i0 < trees.size()) {
    std::vector<LoopTree*> batch; long long evals = 0;
    size_t i1 = i0;
    while (i1 < trees.size()) {
      long long n = 0; for (long long x : trees[i1].ninst) n += x;
      if (!batch.empty() && evals + n > cap) break;
      REQUIRE(n <= cap, TNQS_ERR_OOM, "loop weights: one subgraph needs more environments than fit the scratch budget");
      batch.push_back(&trees[i1]); evals += n; ++i1;
    }
    std::map<int, const cplx*> units;
    for (LoopTree* t : batch) for (int e : t->cuts) loop_units(c, c->chi[e], units);
    std::vector<std::vector<cplx>> roots(batch.size());
    loop_run_batch(c, batch, units, roots);
    for (size_t bi = 0; bi < batch.size(); ++bi) {
      const LoopTree& t = *batch[bi];
      // signed sum over the assignments of all cuts (the root's cut set): fixed order, compensated
      double sre = 0, sim = 0, cre = 0, cim = 0;
      for (long long inst = 0; inst < t.ninst[0]; ++inst) {
        long long r = inst; int sign = 1;
        for (int j : t.cutset[0]) { const long long rad = loop_radix(c, t.cuts[j]); if (r % rad == rad - 1) sign = -sign; r /= rad; }
        const cplx v = roots[bi][(size_t)inst];
        double y = sign * v.x - cre, tt = sre + y; cre = (tt - sre) - y; sre = tt;
        y = sign * v.y - cim; tt = sim + y; cim = (tt - sim) - y; sim = tt;
      }
      const cplx a = w[t.sub];
      w[t.sub] = cmake(a.x * sre - a.y * sim, a.x * sim + a.y * sre);
    }
    c->release_temps();
    i0 = i1;
  }
  for (int i = 0; i < nsub; ++i) { out[2 * i] = w[i].x; out[2 * i + 1] = w[i].y; }
}

// Transfer matrix of a closed walk v_0 -> v_1 -> .. -> v_{L-1} -> v_0 with the bond of its last step cut open
// (beliefpropagationcache.jl:145-179): column rho = the environment that arrives at the open leg of v_{L-1} when the matrix unit
// rho sits on the leg of v_0, i.e. L levels of chi^2 message updates.  out: chi^2 x chi^2 column-major, ket index fastest.
static void loop_transfer_impl(tnqs_ctx* c, int len, const int32_t* walk, double* out) {
  REQUIRE(!c->partitioned, TNQS_ERR_UNSUPPORTED, "loop transfer matrix on a partitioned handle");
  REQUIRE(len >= 3, TNQS_ERR_INVALID_ARG, "a loop needs at least three vertices");
  std::set<int> seen;
  for (int i = 0; i < len; ++i) { REQUIRE(walk[i] >= 0 && walk[i] < c->nv, TNQS_ERR_INVALID_ARG, "walk vertex out of range"); REQUIRE(seen.insert(walk[i]).second, TNQS_ERR_INVALID_ARG, "walk visits a vertex twice"); }
  auto edge_between = [&](int a, int b) { for (int e : c->inc[a]) if (c->other(e, a) == b) return e; return -1; };
  std::vector<int> step(len);                                   // edge from walk[i] to walk[i + 1]
  for (int i = 0; i < len; ++i) { step[i] = edge_between(walk[i], walk[(i + 1) % len]); REQUIRE(step[i] >= 0, TNQS_ERR_INVALID_ARG, "consecutive walk vertices do not share an edge"); }
  const int ecut = step[len - 1], chi = c->chi[ecut];
  const long long n2 = (long long)chi * chi;
  REQUIRE(n2 * len <= LOOP_MAX_TERMS, TNQS_ERR_UNSUPPORTED, "loop transfer matrix: bond dimension too large");
  std::map<int, const cplx*> bank;
  const cplx* units = loop_units(c, chi, bank);
  cplx* buf = c->talloc<cplx>((size_t)len * n2 * c->slot);
  std::vector<std::vector<Chain>> levels(len);
  for (int i = 0; i < len; ++i) {
    const int v = walk[i], ein = step[(i + len - 1) % len], eout = step[i];
    for (long long rho = 0; rho < n2; ++rho) {
      Chain ch; ch.v = v;
      for (size_t l = 0; l < c->inc[v].size(); ++l) {
        const int e = c->inc[v][l], u = c->other(e, v);
        if (e == eout) continue;
        Absorb ab; ab.leg = (int)l; ab.ldm = c->chi[e]; ab.flags = 0; ab.mat = c->msg(c->dedge(u, e));
        if (e == ein) ab.mat = i == 0 ? units + (size_t)rho * n2 : buf + ((size_t)(i - 1) * n2 + rho) * c->slot;
        ch.absorb.push_back(ab);
      }
      ch.close_leg = c->leg(v, eout); ch.dst = buf + ((size_t)i * n2 + rho) * c->slot;
      levels[i].push_back(std::move(ch));
    }
  }
  Plan p = build_plan(c, levels, scratch_budget(c));
  run_plan(c, p);
  std::vector<cplx> h((size_t)n2 * c->slot);
  CU(cudaMemcpyAsync(h.data(), buf + (size_t)(len - 1) * n2 * c->slot, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  cplx* o = (cplx*)out;
  for (long long rho = 0; rho < n2; ++rho) for (long long i = 0; i < n2; ++i) o[i + rho * n2] = h[(size_t)rho * c->slot + i];
}
