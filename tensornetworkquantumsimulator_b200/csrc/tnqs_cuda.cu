// libtnqs_cuda.so — C-ABI entry points (include/tnqs_cuda.h) over the engine in engine.cuh.
// Build: see csrc/Makefile (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared).
#include "engine.cuh"
#include <set>
#include <map>

static thread_local std::string g_create_error;
static void bp_cache_drop(tnqs_ctx* c);
static void halo_release(tnqs_ctx* c);
static void drop_pending_gate(tnqs_ctx* c);
static void gate_graphs_drop(tnqs_ctx* c);
static bool gate_wpath();
// TNQS_SVD_ROWSPACE=0: the Jacobi SVD of a gate always iterates the full theta' (no row-space reduction, see SvdDesc)
// TNQS_CHOL_ADAPTIVE=0: CholeskyQR2 always runs both passes
static bool chol_adaptive() { const char* e = getenv("TNQS_CHOL_ADAPTIVE"); return !(e && e[0] == '0'); }
// Operator-Schmidt rank of a two-site gate o[s1', s2', s1, s2] over the (site 1 | site 2) cut: the same pivoted Gram-Schmidt and
// threshold as k_svd_trunc (which decides for itself; the host only needs the number to size shared memory)
static int gate_operator_rank(const cplx* g, int d) {
  const int d2 = d * d;
  std::vector<cplx> rows((size_t)d2 * d2), basis((size_t)d2);
  for (int e1 = 0; e1 < d2; ++e1) for (int e2 = 0; e2 < d2; ++e2) {
    const int s1p = e1 % d, s1 = e1 / d, s2p = e2 % d, s2 = e2 / d;
    rows[(size_t)e1 * d2 + e2] = g[s1p + d * (s2p + d * (s1 + d * s2))];
  }
  std::vector<double> nrm(d2); double maxn2 = 0;
  for (int e = 0; e < d2; ++e) { double s = 0; for (int k = 0; k < d2; ++k) { const cplx x = rows[(size_t)e * d2 + k]; s += x.x * x.x + x.y * x.y; } nrm[e] = s; maxn2 = std::max(maxn2, s); }
  int ro = 0;
  for (int it = 0; it < d2; ++it) {
    int best = 0; double bn = -1;
    for (int e = 0; e < d2; ++e) if (nrm[e] > bn) { bn = nrm[e]; best = e; }
    if (!(bn > 1e-26 * maxn2) || bn <= 0.0) break;
    const double inv = 1.0 / std::sqrt(bn);
    for (int k = 0; k < d2; ++k) { const cplx x = rows[(size_t)best * d2 + k]; basis[k] = cmake(x.x * inv, x.y * inv); }
    for (int e = 0; e < d2; ++e) {
      if (e == best) { nrm[e] = -1; continue; }
      if (nrm[e] < 0) continue;
      double pr = 0, pi = 0;
      for (int k = 0; k < d2; ++k) { const cplx b = basis[k], x = rows[(size_t)e * d2 + k]; pr += b.x * x.x + b.y * x.y; pi += b.x * x.y - b.y * x.x; }
      double s = 0;
      for (int k = 0; k < d2; ++k) {
        cplx& x = rows[(size_t)e * d2 + k]; const cplx b = basis[k];
        x.x -= pr * b.x - pi * b.y; x.y -= pr * b.y + pi * b.x; s += x.x * x.x + x.y * x.y;
      }
      nrm[e] = s;
    }
    ++ro;
  }
  return ro;
}
static bool svd_rowspace() { const char* e = getenv("TNQS_SVD_ROWSPACE"); return !(e && e[0] == '0'); }

#define API_BEGIN(ctx) if (!(ctx)) return TNQS_ERR_INVALID_ARG; try { CU(cudaSetDevice((ctx)->device));
#define API_END(ctx) } catch (const TnqsError& e) { (ctx)->err = e.msg; (ctx)->release_temps(); return e.code; } \
  catch (const std::exception& e) { (ctx)->err = e.what(); (ctx)->release_temps(); return TNQS_ERR_INVALID_ARG; } return TNQS_OK;

extern "C" {

const char* tnqs_version(void) { return "tnqs-cuda 0.1 (sm_100a)"; }

const char* tnqs_status_string(int32_t s) {
  switch (s) {
    case TNQS_OK: return "ok";
    case TNQS_ERR_INVALID_ARG: return "invalid argument";
    case TNQS_ERR_SHAPE: return "shape mismatch";
    case TNQS_ERR_OVERLAPPING_BATCH: return "gates of a batch overlap";
    case TNQS_ERR_NEGATIVE_EIGENVALUE: return "message has a negative eigenvalue beyond the cutoff";
    case TNQS_ERR_SVD_NO_CONVERGENCE: return "Jacobi iteration did not converge";
    case TNQS_ERR_CUDA: return "CUDA error";
    case TNQS_ERR_NCCL: return "NCCL error";
    case TNQS_ERR_OOM: return "out of device memory";
    case TNQS_ERR_UNSUPPORTED: return "unsupported configuration";
  }
  return "unknown";
}

const char* tnqs_last_error(const tnqs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
static void set_smem_attrs() {   // per device: called at every handle creation (cheap)
  cudaFuncSetAttribute(k_eigh_sqrt, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_house_qr, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_svd_trunc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_close_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(k_absorb_slab<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_absorb_slab<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_close_dmma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_close_dmma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_bp_level<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_bp_level<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute((k_bp_sweep<16, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  cudaFuncSetAttribute(k_bp_sweep<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  // NOT the maximal carve-out here: the sweep kernel needs the L1 that two 72 KB CTAs leave (measured: 11.6 ms per sweep with the
  // default split, 14.3 ms with the L1 carved down to the minimum).  TNQS_SWEEP_CARVEOUT=<percent> overrides (experiments).
  if (const char* e = getenv("TNQS_SWEEP_CARVEOUT")) {
    cudaFuncSetAttribute(k_bp_sweep<16>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    cudaFuncSetAttribute((k_bp_sweep<16, true>), cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
  }
  cudaFuncSetAttribute(k_bp_sweep<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_close_leg_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_leg_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_leg_canon_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_tri_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  // three 72 KB CTAs per SM need the largest shared-memory carve-out
  cudaFuncSetAttribute(k_absorb_slab<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_chol_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_deep_absorb<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_deep_close<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_slab_batch<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_slab_batch<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_deep_absorb<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_deep_close<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static void alloc_site(tnqs_ctx* c, int v) {
  long long n = c->site_elems(v);
  c->site[v] = (cplx*)c->dalloc((size_t)n * sizeof(cplx));
  c->site_n[v] = n;
}

static void upload_chi(tnqs_ctx* c) {
  if (c->ne == 0) return;
  c->h2d(c->d_chi, c->chi.data(), sizeof(int) * std::max(c->ne, 1));
}

static void reset_messages(tnqs_ctx* c) {
  if (c->ne == 0) return;
  LAUNCH(c, TNQS_KC_MISC, k_msg_identity, 2 * c->ne, 128, 0, c->msg_cur, c->slot, c->d_chi);
}

static size_t scratch_budget(tnqs_ctx* c) {
  if (const char* e = getenv("TNQS_SCRATCH_MB")) { long mb = atol(e); if (mb > 0) return (size_t)mb << 20; }   // tests: force chunking
  // cudaMemGetInfo takes driver locks and is not cheap on a busy host: the answer is kept for 64 calls (or until the state changes
  // shape), which is as current as the quarter-of-free-memory rule needs
  if (c->budget_calls > 0 && c->budget_version == c->state_version) { --c->budget_calls; return c->budget_bytes; }
  size_t fr = 0, tot = 0;
  cudaMemGetInfo(&fr, &tot);
  size_t b = fr / 4;
  const size_t cap = (size_t)16 << 30, floor_ = (size_t)64 << 20;
  c->budget_bytes = std::max(floor_, std::min(b, cap)); c->budget_calls = 64; c->budget_version = c->state_version;
  return c->budget_bytes;
}

extern "C" int32_t tnqs_create(tnqs_ctx** out, int32_t nv, int32_t ne, const int32_t* esrc, const int32_t* edst,
                               const int32_t* d, const int32_t* chi, int32_t chi_cap, int32_t device) {
  if (!out || nv <= 0 || ne < 0 || !d || (ne > 0 && (!esrc || !edst || !chi)) || chi_cap < 1) {
    g_create_error = "tnqs_create: invalid argument"; return TNQS_ERR_INVALID_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_create_error = "tnqs_create: no usable CUDA device (this engine has no CPU fallback)"; return TNQS_ERR_CUDA;
  }
  tnqs_ctx* c = new tnqs_ctx();
  try {
    c->device = device;
    CU(cudaSetDevice(device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount; c->smem_optin = prop.sharedMemPerBlockOptin;
    set_smem_attrs();
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // keep freed stream-ordered allocations in the pool across synchronisations (default: trimmed to 0 at every sync)
      cudaMemPool_t pool; CU(cudaDeviceGetDefaultMemPool(&pool, device));
      unsigned long long keep = ~0ull;
      CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    c->nv = nv; c->ne = ne; c->chi_cap = chi_cap;
    c->esrc.assign(esrc, esrc + ne); c->edst.assign(edst, edst + ne);
    c->d.assign(d, d + nv); c->chi.assign(chi, chi + ne);
    c->inc.assign(nv, {});
    std::set<std::pair<int, int>> seen;
    for (int e = 0; e < ne; ++e) {
      int a = esrc[e], b = edst[e];
      REQUIRE(a >= 0 && a < nv && b >= 0 && b < nv && a != b, TNQS_ERR_INVALID_ARG, "bad edge endpoints");
      REQUIRE(seen.insert({std::min(a, b), std::max(a, b)}).second, TNQS_ERR_INVALID_ARG, "duplicate edge");
      REQUIRE(chi[e] >= 1 && chi[e] <= chi_cap, TNQS_ERR_SHAPE, "bond dimension outside [1, chi_cap]");
      c->inc[a].push_back(e); c->inc[b].push_back(e);
    }
    c->dmax = 1;
    for (int v = 0; v < nv; ++v) {
      REQUIRE(d[v] >= 1 && d[v] <= 4, TNQS_ERR_UNSUPPORTED, "site dimension must be in [1, 4]");
      REQUIRE((int)c->inc[v].size() <= TNQS_MAX_LEGS, TNQS_ERR_UNSUPPORTED, "vertex degree > 8");
      c->dmax = std::max(c->dmax, d[v]);
    }
    REQUIRE(chi_cap <= 64, TNQS_ERR_UNSUPPORTED, "chi_cap > 64 is not supported");
    c->slot = (long long)chi_cap * chi_cap;
    c->site.assign(nv, nullptr); c->site_n.assign(nv, 0);
    for (int v = 0; v < nv; ++v) { alloc_site(c, v); CU(cudaMemsetAsync(c->site[v], 0, c->site_n[v] * sizeof(cplx), c->stream)); }
    size_t mbytes = (size_t)std::max(1, 2 * ne) * c->slot * sizeof(cplx);
    c->msg_cur = (cplx*)c->dalloc(mbytes); c->msg_stage = (cplx*)c->dalloc(mbytes);
    CU(cudaMemsetAsync(c->msg_cur, 0, mbytes, c->stream)); CU(cudaMemsetAsync(c->msg_stage, 0, mbytes, c->stream));
    c->d_chi = (int*)c->dalloc(sizeof(int) * std::max(ne, 1));
    c->d_status = (int*)c->dalloc(sizeof(int));
    CU(cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream));
    upload_chi(c);
    reset_messages(c);
    c->sync();
  } catch (const TnqsError& e) { g_create_error = e.msg; int code = e.code; delete c; return code; }
  *out = c;
  return TNQS_OK;
}

extern "C" int32_t tnqs_destroy(tnqs_ctx* c) {
  if (!c) return TNQS_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->collect_profile();
  drop_pending_gate(c);
  bp_cache_drop(c);
  gate_graphs_drop(c);
  // bilinear forms: no handle may keep a pointer to this one (its cached BP plan holds our site pointers), nor we to another
  for (tnqs_ctx* u : c->bra_users) { cudaStreamSynchronize(u->stream); bp_cache_drop(u); u->bra = nullptr; ++u->state_version; }
  c->bra_users.clear();
  if (c->bra) { auto& us = c->bra->bra_users; us.erase(std::remove(us.begin(), us.end(), c), us.end()); c->bra = nullptr; }
  c->arena_destroy();
  c->hstage_destroy();
  halo_release(c);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); c->copy_stream = nullptr; }
  if (c->d_halo_parity) cudaFreeAsync(c->d_halo_parity, c->stream);
  if (c->d_halo_out) cudaFreeAsync(c->d_halo_out, c->stream);
  if (c->d_halo_recv_k) cudaFreeAsync(c->d_halo_recv_k, c->stream);
  for (cplx* p : c->site) if (p) cudaFreeAsync(p, c->stream);
  if (c->msg_cur) cudaFreeAsync(c->msg_cur, c->stream);
  if (c->msg_stage) cudaFreeAsync(c->msg_stage, c->stream);
  if (c->d_chi) cudaFreeAsync(c->d_chi, c->stream);
  if (c->d_status) cudaFreeAsync(c->d_status, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaStreamDestroy(c->stream);
  delete c;
  return TNQS_OK;
}

extern "C" int32_t tnqs_clone(const tnqs_ctx* src, tnqs_ctx** out) {
  if (!src || !out) return TNQS_ERR_INVALID_ARG;
  tnqs_ctx* c = nullptr;
  int32_t st = tnqs_create(&c, src->nv, src->ne, src->esrc.data(), src->edst.data(), src->d.data(), src->chi.data(),
                           src->chi_cap, src->device);
  if (st != TNQS_OK) return st;
  try {
    CU(cudaStreamSynchronize(src->stream));
    for (int v = 0; v < c->nv; ++v)
      CU(cudaMemcpyAsync(c->site[v], src->site[v], src->site_n[v] * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->msg_cur, src->msg_cur, (size_t)std::max(1, 2 * c->ne) * c->slot * sizeof(cplx),
                       cudaMemcpyDeviceToDevice, c->stream));
    c->owner = src->owner; c->rank = src->rank; c->partitioned = src->partitioned;
    c->sync();
  } catch (const TnqsError& e) { g_create_error = e.msg; int code = e.code; tnqs_destroy(c); return code; }
  *out = c;
  return TNQS_OK;
}

extern "C" int32_t tnqs_sync(tnqs_ctx* c) { API_BEGIN(c) c->sync(); API_END(c) }
extern "C" int32_t tnqs_num_vertices(const tnqs_ctx* c) { return c ? c->nv : 0; }
extern "C" int32_t tnqs_num_edges(const tnqs_ctx* c) { return c ? c->ne : 0; }
extern "C" int32_t tnqs_get_bond_dims(const tnqs_ctx* c, int32_t* chi) {
  if (!c || !chi) return TNQS_ERR_INVALID_ARG;
  for (int e = 0; e < c->ne; ++e) chi[e] = c->chi[e];
  return TNQS_OK;
}
extern "C" int64_t tnqs_site_size(const tnqs_ctx* c, int32_t v) { return (c && v >= 0 && v < c->nv) ? c->site_n[v] : -1; }

extern "C" int32_t tnqs_set_site(tnqs_ctx* c, int32_t v, const double* data, int64_t n) {
  API_BEGIN(c)
  REQUIRE(v >= 0 && v < c->nv && data, TNQS_ERR_INVALID_ARG, "tnqs_set_site: bad vertex or null data");
  REQUIRE(n == c->site_n[v], TNQS_ERR_SHAPE, "tnqs_set_site: element count does not match the site tensor");
  CU(cudaMemcpyAsync(c->site[v], data, n * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_get_site(const tnqs_ctx* cc, int32_t v, double* data, int64_t n) {
  tnqs_ctx* c = const_cast<tnqs_ctx*>(cc);
  API_BEGIN(c)
  REQUIRE(v >= 0 && v < c->nv && data, TNQS_ERR_INVALID_ARG, "tnqs_get_site: bad vertex or null data");
  REQUIRE(n == c->site_n[v], TNQS_ERR_SHAPE, "tnqs_get_site: element count does not match the site tensor");
  CU(cudaMemcpyAsync(data, c->site[v], n * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_set_sites(tnqs_ctx* c, int32_t count, const int32_t* v, const double* const* data, const int64_t* n) {
  API_BEGIN(c)
  REQUIRE(count >= 0 && (count == 0 || (v && data && n)), TNQS_ERR_INVALID_ARG, "tnqs_set_sites: null argument");
  for (int i = 0; i < count; ++i) {
    REQUIRE(v[i] >= 0 && v[i] < c->nv && data[i], TNQS_ERR_INVALID_ARG, "tnqs_set_sites: bad vertex or null data");
    REQUIRE(n[i] == c->site_n[v[i]], TNQS_ERR_SHAPE, "tnqs_set_sites: element count does not match the site tensor");
  }
  for (int i = 0; i < count; ++i) CU(cudaMemcpyAsync(c->site[v[i]], data[i], n[i] * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_get_sites(const tnqs_ctx* cc, int32_t count, const int32_t* v, double* const* data, const int64_t* n) {
  tnqs_ctx* c = const_cast<tnqs_ctx*>(cc);
  API_BEGIN(c)
  REQUIRE(count >= 0 && (count == 0 || (v && data && n)), TNQS_ERR_INVALID_ARG, "tnqs_get_sites: null argument");
  for (int i = 0; i < count; ++i) {
    REQUIRE(v[i] >= 0 && v[i] < c->nv && data[i], TNQS_ERR_INVALID_ARG, "tnqs_get_sites: bad vertex or null data");
    REQUIRE(n[i] == c->site_n[v[i]], TNQS_ERR_SHAPE, "tnqs_get_sites: element count does not match the site tensor");
  }
  for (int i = 0; i < count; ++i) CU(cudaMemcpyAsync(data[i], c->site[v[i]], n[i] * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_fill_random(tnqs_ctx* c, uint64_t seed) {
  API_BEGIN(c)
  for (int v = 0; v < c->nv; ++v) {
    long long n = c->site_n[v];
    int blocks = (int)std::min<long long>((n + 255) / 256, 4096);
    LAUNCH(c, TNQS_KC_MISC, k_fill_random, blocks, 256, 0, c->site[v], n, (unsigned long long)seed, (unsigned long long)v);
  }
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_set_message(tnqs_ctx* c, int32_t k, const double* data, int64_t n) {
  API_BEGIN(c)
  REQUIRE(k >= 0 && k < 2 * c->ne && data, TNQS_ERR_INVALID_ARG, "tnqs_set_message: bad edge or null data");
  long long x = c->chi[k >> 1];
  REQUIRE(n == x * x, TNQS_ERR_SHAPE, "tnqs_set_message: element count != chi^2");
  CU(cudaMemcpyAsync(c->msg(k), data, n * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_get_message(const tnqs_ctx* cc, int32_t k, double* data, int64_t n) {
  tnqs_ctx* c = const_cast<tnqs_ctx*>(cc);
  API_BEGIN(c)
  REQUIRE(k >= 0 && k < 2 * c->ne && data, TNQS_ERR_INVALID_ARG, "tnqs_get_message: bad edge or null data");
  long long x = c->chi[k >> 1];
  REQUIRE(n == x * x, TNQS_ERR_SHAPE, "tnqs_get_message: element count != chi^2");
  CU(cudaMemcpyAsync(data, c->msg(k), n * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_reset_messages(tnqs_ctx* c) { API_BEGIN(c) reset_messages(c); c->sync(); API_END(c) }

// ---------------------------------------------------------------------------------------------------------------
// BP sweep
// ---------------------------------------------------------------------------------------------------------------
// The last BP plan (level schedule, descriptors, scratch) is kept on the handle and reused by later updates with the same
// edge sequence and options as long as no site tensor or bond dimension changed: a refresh of 25 sweeps, or the
// sweep-by-sweep calls of a partitioned run, build it once.
// Loop state of a refresh that runs as ONE graph launch: the sweep body sits in a conditional WHILE node, so the
// "stop when the mean diff <= tolerance or after maxiter sweeps" test (abstractbeliefpropagationcache.jl:201-218) runs on
// the device and no host round trip separates two sweeps.
struct BpLoopState { double tol; double mean; int maxiter; int iter; int has_tol; int pad; };
// d_cnt: device denominator (the global message count of a partitioned run) or null (nseq)
__global__ void k_bp_loop_check(cudaGraphConditionalHandle h, BpLoopState* st, const double* d_sum, int nseq, const double* d_cnt, const int* status) {
  const int it = ++st->iter;
  bool more = it < st->maxiter;
  if (st->has_tol) { const double mean = *d_sum / (d_cnt ? *d_cnt : (double)nseq); st->mean = mean; if (mean <= st->tol) more = false; }
  if (status && (*status & ST_HALO_TIMEOUT)) more = false;
  cudaGraphSetConditional(h, more ? 1u : 0u);
}
struct BpPlanCache {
  Plan plan; std::vector<int> seq; int mode = -1, normalize = -1, hermitize = -1, want_diff = -1; long long version = -1, bra_version = -1;
  double* d_diff = nullptr; std::vector<void*> owned;
  BpLoopState* d_loop = nullptr;
  cudaGraphExec_t graph = nullptr;     // the whole refresh: WHILE(sweep: all levels + diff sum + loop test), built once per launch structure
};
static void bp_cache_drop(tnqs_ctx* c) {
  BpPlanCache* pc = (BpPlanCache*)c->bp_cache;
  if (!pc) return;
  if (pc->graph) cudaGraphExecDestroy(pc->graph);
  for (void* p : pc->owned) c->dfree(p);
  delete pc; c->bp_cache = nullptr;
}
static int bp_level_split_cap() { const char* e = getenv("TNQS_BP_SPLIT"); return e ? std::max(1, atoi(e)) : 8; }      // read at plan-build time only
static bool graphs_disabled() {
  static int d = -1;
  if (d < 0) { const char* e = getenv("TNQS_NO_GRAPH"); d = (e && e[0] == '1') ? 1 : 0; }
  return d == 1;
}

static void bp_update_impl(tnqs_ctx* c, const int32_t* seq_in, int nseq_in, const tnqs_bp_opts& o, int* iters_out,
                           double* mean_out) {
  REQUIRE(o.maxiter >= 1, TNQS_ERR_INVALID_ARG, "You need to specify a number of iterations for BP!");
  // a gate applied to either layer since tnqs_set_bra may have changed its bond dimensions: the closings would read the
  // bra's site tensors with the ket's extents
  REQUIRE(!c->bra || (c->bra->chi == c->chi && c->bra->d == c->d), TNQS_ERR_SHAPE,
          "bilinear form: bond dimensions of ket and bra differ (a layer changed after tnqs_set_bra)");
  std::vector<int> seq;
  for (int i = 0; i < nseq_in; ++i) {
    int k = seq_in[i];
    REQUIRE(k >= 0 && k < 2 * c->ne, TNQS_ERR_INVALID_ARG, "edge sequence entry out of range");
    if (c->partitioned && c->owner[c->dsrc(k)] != c->rank) continue;
    seq.push_back(k);
  }
  const int nseq = (int)seq.size();
  if (iters_out) *iters_out = 0;
  if (mean_out) *mean_out = NAN;
  if (nseq == 0) return;
  const bool want_diff = o.tolerance >= 0;
  // dependency levels of the sequential (Gauss-Seidel) order; no message is read and written in the same level
  std::vector<int> level;
  int nlevels = 1;
  auto compute_levels = [&]() {                    // only when the plan is (re)built: a cached plan already holds its schedule
  level.assign(nseq, 0);
  if (o.mode == TNQS_BP_REFERENCE_ORDER) {
    std::vector<int> lastwrite(2 * c->ne, -1), lastread(2 * c->ne, -1);
    for (int i = 0; i < nseq; ++i) {
      int k = seq[i], v = c->dsrc(k), w = c->ddst(k), lvl = 0;
      for (int e : c->inc[v]) {
        int u = c->other(e, v);
        if (u == w) continue;
        int r = c->dedge(u, e);
        if (lastwrite[r] >= 0) lvl = std::max(lvl, lastwrite[r] + 1);     // RAW
      }
      lvl = std::max(lvl, lastread[k] + 1);                                 // WAR: strictly later, so a fused level (absorb and
                                                                            // epilogue items of different messages in flight
                                                                            // together) may update its messages in place
      lvl = std::max(lvl, lastwrite[k] + 1);                                // WAW
      level[i] = lvl; lastwrite[k] = lvl;
      for (int e : c->inc[v]) {
        int u = c->other(e, v);
        if (u == w) continue;
        int r = c->dedge(u, e);
        lastread[r] = std::max(lastread[r], lvl);
      }
      nlevels = std::max(nlevels, lvl + 1);
    }
  } else {
    std::vector<char> seen(2 * c->ne, 0);
    for (int k : seq) { REQUIRE(!seen[k], TNQS_ERR_INVALID_ARG, "JACOBI mode needs each directed edge at most once"); seen[k] = 1; }
  }
  };
  BpPlanCache* pc = (BpPlanCache*)c->bp_cache;
  const bool same_request = pc && pc->mode == o.mode && pc->normalize == o.normalize && pc->hermitize == o.hermitize &&
                            pc->want_diff == (int)want_diff && pc->seq == seq;
  const long long bra_version = c->bra ? c->bra->state_version : -1;      // the plan holds the bra's site pointers too
  if (!(same_request && pc->version == c->state_version && pc->bra_version == bra_version)) {
    // (re)build; when only site pointers / values changed the new plan lands in the old one's buffers and keeps its graph
    BpPlanCache* old = same_request ? pc : nullptr;
    if (pc && !old) { bp_cache_drop(c); pc = nullptr; }
    compute_levels();
    BpPlanCache* nc = old ? old : new BpPlanCache();
    std::vector<void*> saved; saved.swap(c->temp);            // allocations made while building belong to the cache
    c->arena_bypass = true;
    try {
      if (!nc->d_diff) nc->d_diff = c->talloc<double>(nseq + 1);
      if (!nc->d_loop) nc->d_loop = (BpLoopState*)c->talloc<double>(sizeof(BpLoopState) / sizeof(double));
      std::vector<std::vector<Chain>> levels(nlevels);
      std::vector<int> level_size(nlevels, 0);
      for (int i = 0; i < nseq; ++i) ++level_size[level[i]];
      int narrow_level = 20;
      if (const char* e = getenv("TNQS_BP_NARROW")) narrow_level = atoi(e);
      for (int i = 0; i < nseq; ++i) {
        int k = seq[i], v = c->dsrc(k), w = c->ddst(k);
        Chain ch; ch.v = v;
        for (size_t j = 0; j < c->inc[v].size(); ++j) {
          int ee = c->inc[v][j], u = c->other(ee, v);
          if (u == w) { ch.close_leg = (int)j; continue; }
          Absorb ab; ab.leg = (int)j; ab.mat = c->msg(c->dedge(u, ee)); ab.ldm = c->chi[ee]; ab.flags = 0;
          ch.absorb.push_back(ab);
        }
        ch.dst = c->msg_st(k); ch.prev = c->msg(k); ch.diff = want_diff ? nc->d_diff + i : nullptr;
        ch.bp_mode = 1; ch.normalize = o.normalize; ch.hermitize = o.hermitize; ch.commit_k = k;
        if (c->halo_active() && !c->halo_route.empty() && c->halo_route[k].peer >= 0) {
          const tnqs_ctx::HaloRoute& rt = c->halo_route[k];
          const tnqs_ctx::HaloPeer& pr = c->halo_peers[rt.peer];
          ch.dst2 = pr.base + (size_t)rt.slot * c->slot; ch.dst2_pstride = (long long)pr.nrecv * c->slot; ch.parity = c->d_halo_parity;
        }
        ch.split_cap = o.mode == TNQS_BP_JACOBI ? 4 : bp_level_split_cap();      // whole-sweep launches want long CTAs, level launches more
        // Narrow dependency levels (the spine and the tips of the wavefront: 35 of the 124 levels of a 32 x 32 grid hold fewer than
        // 22 messages) are pure critical path - the device is mostly idle while their closings run - so those closings are cut
        // into four times as many, four times shorter splits.  A function of the edge sequence only: reruns are bitwise reproducible.
        if (o.mode == TNQS_BP_REFERENCE_ORDER && level_size[level[i]] <= narrow_level) { ch.split_cap = 4 * ch.split_cap; ch.split_rows = 128; }
        levels[level[i]].push_back(ch);
      }
      Plan np;
      if (!build_sweep_plan(c, levels, /*in_place=*/o.mode == TNQS_BP_REFERENCE_ORDER, old ? &old->plan : nullptr, np))
        np = build_plan(c, levels, scratch_budget(c), (old && !old->plan.sweep) ? &old->plan : nullptr, /*strict_levels=*/o.mode == TNQS_BP_REFERENCE_ORDER);
      if (nc->graph && !same_launches(np, nc->plan)) { cudaGraphExecDestroy(nc->graph); nc->graph = nullptr; }
      nc->plan = np;
    } catch (...) { c->arena_bypass = false; for (void* p : c->temp) c->dfree(p); c->temp.swap(saved); if (!old) delete nc; else { c->bp_cache = nullptr; } throw; }
    c->arena_bypass = false;
    for (void* p : c->temp) nc->owned.push_back(p);
    c->temp.clear(); c->temp.swap(saved);
    nc->seq = seq; nc->mode = o.mode; nc->normalize = o.normalize; nc->hermitize = o.hermitize; nc->want_diff = want_diff;
    nc->version = c->state_version; nc->bra_version = bra_version;
    c->bp_cache = nc; pc = nc;
  }
  const Plan& plan = pc->plan;
  // vertex-partitioned run with an open peer-memory halo: every sweep ends with the device-side barrier / all-reduce and the
  // commit of what the peers stored, inside the same graph - a whole refresh is one launch per rank, no host round trip
  const bool halo = c->halo_active();
  REQUIRE(!halo || o.mode == TNQS_BP_JACOBI, TNQS_ERR_INVALID_ARG, "peer-memory halo: partitioned BP runs in JACOBI mode");
  HaloBarrierArgs hb;
  if (halo) {
    memset(&hb, 0, sizeof(hb));
    hb.world = c->halo_world; hb.rank = c->rank;
    for (int r = 0; r < c->halo_world; ++r) {
      hb.sig[r] = r == c->rank ? (double*)c->halo_block : (double*)c->halo_peers[r].block;
      REQUIRE(hb.sig[r], TNQS_ERR_INVALID_ARG, "peer-memory halo: the block of every other rank must be opened (tnqs_halo_peer)");
    }
    hb.epoch = c->d_halo_parity + 1; hb.next_parity = c->d_halo_parity;
    hb.local_sum = want_diff ? pc->d_diff + nseq : nullptr; hb.local_cnt = (double)nseq;
    hb.out_sum = c->d_halo_out; hb.out_cnt = c->d_halo_out + 1;
    hb.status = c->d_status; hb.timeout_bit = ST_HALO_TIMEOUT;
  }
  auto end_of_sweep = [&]() {                                  // after the sweep's kernels and k_sum
    if (!halo) return;
    LAUNCH(c, TNQS_KC_BP_FINALIZE, k_halo_barrier, 1, HALO_MAX_WORLD, 0, hb);
    if (c->halo_nrecv > 0)
      LAUNCH(c, TNQS_KC_BP_FINALIZE, k_halo_commit, c->halo_nrecv, 128, 0, c->msg_cur, c->halo_in, c->slot, c->halo_nrecv, c->d_halo_recv_k, c->d_chi, c->d_halo_parity);
  };
  double* d_diff = pc->d_diff;
  double* d_sum = d_diff + nseq;
  int it = 0; double mean = NAN;
  const bool use_graph = !c->profiling && !graphs_disabled() && (plan.nlaunch >= 8 || plan.sweep || halo);
  const int per_sweep = plan.nlaunch + (want_diff ? 1 : 0);
  if (use_graph) {
    if (!pc->graph) {
      cudaGraph_t g = nullptr;
      CU(cudaGraphCreate(&g, 0));
      try {
        cudaGraphConditionalHandle h;
        CU(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));     // every launch enters the body once
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = h; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
        cudaGraphNode_t node;
        CU(cudaGraphAddNode(&node, g, nullptr, 0, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        CU(cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        cudaGraph_t cap = nullptr;
        try {
          run_plan(c, plan);
          if (want_diff) LAUNCH(c, TNQS_KC_BP_FINALIZE, k_sum, 1, 1024, 0, d_diff, nseq, d_sum);
          end_of_sweep();
          LAUNCH(c, TNQS_KC_BP_FINALIZE, k_bp_loop_check, 1, 1, 0, h, pc->d_loop, halo ? (const double*)c->d_halo_out : (const double*)d_sum, nseq,
                 halo ? (const double*)(c->d_halo_out + 1) : (const double*)nullptr, (const int*)c->d_status);
        } catch (...) { cudaStreamEndCapture(c->stream, &cap); throw; }
        CU(cudaStreamEndCapture(c->stream, &cap));
        cudaError_t ge = cudaGraphInstantiate(&pc->graph, g, 0);
        if (ge != cudaSuccess) throw TnqsError(TNQS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ge));
      } catch (...) { cudaGraphDestroy(g); throw; }
      cudaGraphDestroy(g);
      c->launches -= per_sweep + 1 + (halo ? 1 + (c->halo_nrecv > 0 ? 1 : 0) : 0);       // the capture pass launched nothing
    }
    BpLoopState hs; hs.tol = o.tolerance; hs.mean = NAN; hs.maxiter = o.maxiter; hs.iter = 0; hs.has_tol = want_diff; hs.pad = 0;
    c->h2d(pc->d_loop, &hs, sizeof(hs));
    CU(cudaGraphLaunch(pc->graph, c->stream));
    c->d2h_small(&hs, pc->d_loop, sizeof(hs));
    it = hs.iter; mean = want_diff ? hs.mean : NAN;
    c->launches += (long long)it * (per_sweep + 1 + (halo ? 1 + (c->halo_nrecv > 0 ? 1 : 0) : 0));
    c->bp_sweeps += it; c->bp_msg_updates += (long long)it * nseq;
  } else {
    for (it = 1; it <= o.maxiter; ++it) {
      run_plan(c, plan);
      if (want_diff) LAUNCH(c, TNQS_KC_BP_FINALIZE, k_sum, 1, 1024, 0, d_diff, nseq, d_sum);
      end_of_sweep();
      ++c->bp_sweeps; c->bp_msg_updates += nseq;
      if (want_diff) {
        double s[2] = {0, (double)nseq};
        c->d2h_small(s, halo ? c->d_halo_out : d_sum, (halo ? 2 : 1) * sizeof(double));
        mean = s[0] / s[1];
        if (mean <= o.tolerance) break;
      }
    }
  }
  if (plan.sweep && plan.d_prof) {
    long long t[32];
    c->d2h_small(t, plan.d_prof, sizeof(t));
    CU(cudaMemsetAsync(plan.d_prof, 0, sizeof(t), c->stream));
    const double na = (double)std::max(1ll, t[0]), nc = (double)std::max(1ll, t[16]);
    fprintf(stderr, "[tnqs trace] %lld absorb items, cycles each: producer wait %.0f, body %.0f, publish %.0f | %lld closing items: slab wait %.0f, body %.0f, tail %.0f\n",
            t[0], t[2] / na, t[8] / na, t[9] / na, t[16], t[17] / nc, t[18] / nc, t[19] / nc);
  }
  if (plan.d_qctr) {
    int q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    c->d2h_small(q, plan.d_qctr, (plan.sweep ? 8 : 4) * sizeof(int));
    if (q[3] || q[5] || q[6]) {
      static const char* dbg = getenv("TNQS_DEBUG_QUEUE");
      if (dbg) fprintf(stderr, "[tnqs] %s wait polls in %d sweeps: producers/ring %d, slabs %d, readers %d\n", plan.sweep ? "k_bp_sweep" : "k_bp_level", it, q[3], q[5], q[6]);
      CU(cudaMemsetAsync(plan.d_qctr + 3, 0, sizeof(int), c->stream));
      if (plan.sweep) CU(cudaMemsetAsync(plan.d_qctr + 5, 0, 2 * sizeof(int), c->stream));
    }
    if (q[2]) { bp_cache_drop(c); throw TnqsError(TNQS_ERR_CUDA, "internal: a BP queue wait timed out"); }
  }
  if (halo) { const int st = c->read_status(); if (st & ST_HALO_TIMEOUT) { bp_cache_drop(c); throw TnqsError(TNQS_ERR_NCCL, "peer-memory halo: a rank did not reach the sweep barrier (time-out)"); } }
  if (it > o.maxiter) it = o.maxiter;
  if (iters_out) *iters_out = it;
  if (mean_out) *mean_out = mean;
}

extern "C" int32_t tnqs_bp_update(tnqs_ctx* c, const int32_t* edge_seq, int32_t nseq, const tnqs_bp_opts* opts,
                                  int32_t* iters, double* mean_diff) {
  API_BEGIN(c)
  REQUIRE(opts && (edge_seq || nseq == 0) && nseq >= 0, TNQS_ERR_INVALID_ARG, "tnqs_bp_update: null argument");
  bp_update_impl(c, edge_seq, nseq, *opts, iters, mean_diff);
  c->sync();
  c->release_temps();
  API_END(c)
}

// ---------------------------------------------------------------------------------------------------------------
// gates
// ---------------------------------------------------------------------------------------------------------------
static void apply_1site_impl(tnqs_ctx* c, int n, const int32_t* vs, const double* gates, int normalize) {
  if (n == 0) return;
  std::set<int> seen;
  std::vector<cplx> hg; std::vector<size_t> goff(n);
  size_t base = 0;
  for (int i = 0; i < n; ++i) {
    int v = vs[i];
    REQUIRE(v >= 0 && v < c->nv, TNQS_ERR_INVALID_ARG, "tnqs_apply_1site: vertex out of range");
    REQUIRE(seen.insert(v).second, TNQS_ERR_OVERLAPPING_BATCH, "tnqs_apply_1site: repeated vertex in batch");
    goff[i] = hg.size();
    int d = c->d[v];
    const cplx* g = (const cplx*)gates + base;
    hg.insert(hg.end(), g, g + d * d);
    base += (size_t)d * d;
  }
  cplx* d_g = c->upload(hg);
  const int NB = 64;                       // norm partials per site
  double* d_part = c->talloc<double>((size_t)n * NB);
  double* d_n2 = c->talloc<double>(n);
  std::vector<Gate1Desc> g1(n); std::vector<Norm2Desc> nd(n); std::vector<ScaleDesc> sd(n);
  for (int i = 0; i < n; ++i) {
    int v = vs[i];
    // in place: every thread owns one fibre T[:, rest] (reads its d values, writes them back), so no allocation and the
    // site pointers (and any cached BP plan) stay valid
    g1[i].in = c->site[v]; g1[i].out = c->site[v]; g1[i].gate = d_g + goff[i]; g1[i].d = c->d[v];
    g1[i].rest = c->site_n[v] / c->d[v]; g1[i].norm2_partial = d_part + (size_t)i * NB;
    nd[i].x = nullptr; nd[i].n = 0; nd[i].out = d_n2 + i; nd[i].partial = d_part + (size_t)i * NB; nd[i].npartial = NB;
    sd[i].x = c->site[v]; sd[i].n = c->site_n[v]; sd[i].scalar = d_n2 + i; sd[i].mode = 0;
  }
  Gate1Desc* dg1 = c->upload(g1);
  LAUNCH(c, TNQS_KC_MISC, k_gate1, dim3(NB, n), 256, 0, dg1);
  if (normalize) {
    Norm2Desc* dnd = c->upload(nd); ScaleDesc* dsd = c->upload(sd);
    LAUNCH(c, TNQS_KC_MISC, k_norm2, n, 64, 0, dnd);
    LAUNCH(c, TNQS_KC_MISC, k_scale, dim3(NB, n), 256, 0, dsd);
  }
}

extern "C" int32_t tnqs_apply_1site(tnqs_ctx* c, int32_t n, const int32_t* v, const double* gates, int32_t normalize) {
  API_BEGIN(c)
  REQUIRE(n >= 0 && (n == 0 || (v && gates)), TNQS_ERR_INVALID_ARG, "tnqs_apply_1site: null argument");
  apply_1site_impl(c, n, v, gates, normalize);
  c->sync();
  c->release_temps();
  API_END(c)
}

// both directed messages of edge list[i] <- diag(svals[i*cap .. i*cap + rank[i]))
__global__ void k_msg_diag(cplx* slab, long long slot, const int* list, const double* svals, const int* rank, int cap, const int* commit = nullptr) {
  int i = blockIdx.x, e = list[i], r = rank[i];
  if (commit && commit[i] == 0) return;
  cplx* a = slab + (size_t)(2 * e) * slot; cplx* b = a + slot;
  for (int idx = threadIdx.x; idx < r * r; idx += blockDim.x) {
    int x = idx % r, y = idx / r;
    cplx v = cmake(x == y ? svals[(size_t)i * cap + x] : 0.0, 0.0);
    a[idx] = v; b[idx] = v;
  }
}


// Optimistic gate batches: gate i may commit (overwrite its two site tensors in place, reset its messages) iff its kept rank is the
// one the host assumed, neither of its QR factorisations raised the CholeskyQR2 flag and no kernel of the batch reported an error.
// flag_idx[2i], flag_idx[2i+1]: positions of the two sites in the flag array, or -1 (Householder path / remote half).
__global__ void k_gate_ok(int n, const int* rank, const int* expect, const int* flags, const int* flag_idx, const int* cut, const int* status, int* ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int fa = flag_idx[2 * i], fb = flag_idx[2 * i + 1];
  ok[i] = rank[i] == expect[i] && (fa < 0 || flags[fa] == 0) && (fb < 0 || flags[fb] == 0) && (!cut || (cut[2 * i] == 0 && cut[2 * i + 1] == 0)) && *status == 0;
}

// Debug aid (TNQS_DEBUG=1): Frobenius norm^2 and finiteness of a device buffer, printed to stderr.
static void dbg_norm(tnqs_ctx* c, const char* name, int idx, const cplx* p, size_t n) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("TNQS_DEBUG"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on) return;
  std::vector<cplx> h(n);
  cudaMemcpyAsync(h.data(), p, n * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream);
  cudaStreamSynchronize(c->stream);
  double s = 0; int bad = 0;
  for (size_t i = 0; i < n; ++i) { s += h[i].x * h[i].x + h[i].y * h[i].y; if (!std::isfinite(h[i].x) || !std::isfinite(h[i].y)) ++bad; }
  fprintf(stderr, "[tnqs dbg] %-10s[%d] n=%zu norm2=%.15g nonfinite=%d\n", name, idx, n, s, bad);
  const char* dir = getenv("TNQS_DEBUG_DIR");
  if (dir) {
    char path[512]; snprintf(path, sizeof(path), "%s/%s_%d.bin", dir, name, idx);
    FILE* f = fopen(path, "wb"); if (f) { fwrite(h.data(), sizeof(cplx), n, f); fclose(f); }
  }
}


// Plan of the TSQR tree of one matrix (Po x nc, column-major, ld = Po), see DESIGN.md "QR".
struct QrLevel { cplx* base; long long ld; long long rows; std::vector<long long> blk_begin; };   // blocks = row ranges

// pack / unpack the R factors of cut-crossing gates: slot = [ (q, nc) header | R (q x nc, ld = q) ]
struct XchgDesc { cplx* slot; cplx* r; int q, nc; };
__global__ void k_xchg_pack(const XchgDesc* __restrict__ descs) {
  const XchgDesc d = descs[blockIdx.x];
  if (threadIdx.x == 0) d.slot[0] = cmake((double)d.q, (double)d.nc);
  for (int i = threadIdx.x; i < d.q * d.nc; i += blockDim.x) d.slot[1 + i] = d.r[i];
}
__global__ void k_xchg_headers(const cplx* buf, long long slot, int n, cplx* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = buf[(size_t)i * slot];
}
__global__ void k_xchg_unpack(const XchgDesc* __restrict__ descs) {
  const XchgDesc d = descs[blockIdx.x];
  for (int i = threadIdx.x; i < d.q * d.nc; i += blockDim.x) d.r[i] = d.slot[1 + i];
}

// One vertex-disjoint batch of two-site gates (simple update).  stage1 = roots, matricise, absorb, QR (everything that
// needs only one site); stage2 = theta, gate, SVD + truncation, reabsorb, T' = Q R', normalise, commit.  Between the two
// a vertex-partitioned run exchanges the R factors of gates that cross a cut (`remote`), so each side finishes its own
// half with an identical, deterministic theta/SVD.  All temporaries live until stage2 returns.
// Three ways through it (apply_2site_impl chooses per chunk): conventional (the host reads the CholeskyQR2 flags and the kept ranks,
// then allocates and commits); optimistic (saturated bonds: nothing is read back until the end, the device commits per gate - and with
// square R factors the W path replaces reabsorption, Q R' and the permute back by one in-place product); and the optimistic chain
// replayed from a CUDA graph (GateGraph below).  Gates that fail the optimistic assumptions are redone conventionally.
struct GateBatch {
  struct SiteJob {
    int v, e, jb, d, chib; long long Po; int nc, q; std::vector<int> outer;   // outer: bond leg indices ascending
    cplx* A; cplx* Q; cplx* Rbuf; std::vector<cplx*> root, invroot;             // Q: where the thin Q factor lives
    cplx* colfix = nullptr;     // CholeskyQR2: Q holds Q1 = A R1^-1 and the true factor is Q1 * colfix (colfix = R2^-1, nc x nc); the column
                                // operation is folded into T' = Q R' (it commutes with the outer-leg absorptions on the rows)
    bool remote;
  };
  tnqs_ctx* c; int n = 0, NJ = 0, cap = 1; tnqs_apply_opts o;
  std::vector<int> es; std::vector<cplx> hgates; std::vector<SiteJob> jobs; std::vector<int> LJ;
  std::vector<PermDesc> pd; std::vector<int> ncols;
  std::vector<ThetaDesc> td; std::vector<SvdDesc> sv;
  int* d_rank = nullptr; double* d_err = nullptr; double* d_sv = nullptr;
  std::vector<int> cut;                 // gate indices that have a remote vertex, in batch order
  cplx* xsend = nullptr; long long xslot = 0;
  void absorb_outer(bool inverse);
  std::vector<char> from_site;          // forward absorption reads the canonical site tensor directly (no matricise pass)
  bool slab_ok(int j) const {           // fused slab kernel: 2 or 3 outer legs, all of the same dimension N in {8, 16}
    const SiteJob& J = jobs[j];
    const int F = (int)J.outer.size();
    if (dmma_disabled() || (F != 2 && F != 3)) return false;
    const int N = c->chi[c->inc[J.v][J.outer[0]]];
    if (N != 8 && N != 16) return false;
    for (int l : J.outer) if (c->chi[c->inc[J.v][l]] != N) return false;
    return true;
  }
  bool deep_ok(int j) const {           // 4 or 5 outer legs of dimension 8 on a qubit site: two k_slab_batch passes (kernels_deep.cuh)
    const SiteJob& J = jobs[j];
    const int F = (int)J.outer.size();
    if (dmma_disabled() || deep_disabled() || J.d != 2 || (F != 4 && F != 5)) return false;
    for (int l : J.outer) if (c->chi[c->inc[J.v][l]] != 8) return false;
    return true;
  }
  // CholeskyQR2 runs optimistically for every tall matrix; its per-matrix flags (a pivot below the tolerance) are read as late as
  // possible - after the host has built the next stage's descriptors, while the device finishes the QR kernels - and the flagged
  // matrices are then redone by Householder TSQR
  std::vector<int> chol_list; std::vector<cplx*> chol_Qb, chol_Rinv; int* d_chol_flag = nullptr; bool chol_pending = false;
  // Optimistic mode (saturated bonds: every gate of the batch is expected to keep its bond dimension, the steady state of a
  // truncated evolution): the host assumes rank = old chi and no CholeskyQR2 flag, enqueues the whole chain without reading
  // anything back, and the device commits a gate only if the assumptions held (k_gate_ok).  One synchronisation per batch; the
  // gates whose assumptions failed are left untouched and listed in `failed` for a second, conventional pass.
  bool optimistic = false; int* d_ok = nullptr; std::vector<int> failed;
  // W path (optimistic mode, square R): with messages of full rank sqrt(M) M^(-1/2) = 1, so T' = Q' R' = psi (R^-1 R') - the new site
  // tensor is the OLD one times a small matrix over its (site, gate-bond) legs, computed in place in the canonical layout
  // (k_leg_canon_dmma).  No inverse absorption, no Q R' product, no permute back; a message with a cut eigenvalue (k_eigh_sqrt's
  // flag) fails the gate's commit predicate and sends it to the conventional second pass.
  std::vector<char> wpath; int* d_jobcut = nullptr;
  bool defer_collect = false;            // graph capture: stage2 stops after the last enqueue, the caller collects after the launch
  char* gates_staged = nullptr;          // where the gate matrices were staged for upload (refreshed before a graph replay)
  void collect_optimistic(int32_t* ranks_out, double* err_out, double* svals_out);
  void householder_qr(const std::vector<int>& list);
  void resolve_chol_flags();
  void stage1(int n_, const int32_t* es_, const int32_t* flips, const double* gates, const tnqs_apply_opts& o_, const int32_t* remote);
  void stage2(const cplx* xrecv, int32_t* ranks_out, double* err_out, double* svals_out);
};

void GateBatch::absorb_outer(bool inverse) {   // forward: on A (psi~); inverse: on the Q factor
  // Fast path: every outer leg of the same dimension N in {8, 16}, 2 or 3 of them: each column of the matricised tensor
  // is a contiguous N^F slab, so the fused slab kernel absorbs all the legs in one pass, in place (k_absorb_slab).
  std::vector<int> generic, deep;
  std::vector<SlabDesc> slab[2];                 // [0]: N = 8, [1]: N = 16
  int maxcols[2] = {0, 0};
  for (int j : LJ) {
    if (inverse && !wpath.empty() && wpath[j]) continue;
    SiteJob& J = jobs[j];
    const int F = (int)J.outer.size();
    int N = F ? c->chi[c->inc[J.v][J.outer[0]]] : 0;
    const bool ok = slab_ok(j) && ncols[j] >= 1;
    if (!ok && !inverse && from_site[j] && deep_ok(j)) { deep.push_back(j); continue; }
    if (!ok) { if (F) generic.push_back(j); continue; }
    cplx* buf = inverse ? J.Q : J.A;
    SlabDesc d; d.t = buf; d.x = buf; d.F = F;
    long long st = 1;
    for (int i = 0; i < 3; ++i) { d.g[i] = i < F ? st : 0; d.m[i] = i < F ? (inverse ? J.invroot[i] : J.root[i]) : nullptr; st *= N; }
    d.gs = J.Po; d.gw = 2 * J.Po; d.nslabs = ncols[j]; d.flags = inverse ? 3 : 0;
    if (!inverse && from_site[j]) {     // column (s, b) of the matrix = slab T[s, :, :, :, b] of the canonical tensor: matricise on the fly
      d.t = c->site[J.v];
      for (int i = 0; i < F; ++i) d.g[i] = pd[j].ostride[i];
      d.gs = 1; d.gw = pd[j].bstride;
    }
    slab[N == 16].push_back(d); maxcols[N == 16] = std::max(maxcols[N == 16], ncols[j]);
  }
  for (int w = 0; w < 2; ++w) {
    if (slab[w].empty()) continue;
    const int N = w ? 16 : 8;
    SlabDesc* dsd = c->upload(slab[w]);
    dim3 grid((unsigned)((maxcols[w] + 1) / 2 * 2), (unsigned)slab[w].size(), 1);
    size_t smem = N == 16 ? SlabCfg<16>::SMEM : SlabCfg<8>::SMEM;
    if (w) LAUNCH(c, TNQS_KC_MODE_APPLY, k_absorb_slab<16>, grid, SLAB_THREADS, smem, dsd);
    else LAUNCH(c, TNQS_KC_MODE_APPLY, k_absorb_slab<8>, grid, SLAB_THREADS, smem, dsd);
  }
  if (!deep.empty()) {
    // 4 or 5 outer legs of dimension 8: legs 1-3 from the canonical site tensor into the matrix layout (warp = o4, CTA = (o5, s, b)),
    // then legs 4 (and 5) in place (F = 5: slab (o1, o4, o5), warp = o2, CTA = (o3, column); F = 4: slab (o1, o2, o4), warp = o3,
    // CTA = column)
    std::vector<BatchSlabDesc> p1, p2; int g1 = 1, g2 = 1;
    for (int j : deep) {
      const SiteJob& J = jobs[j]; const PermDesc& p = pd[j];
      const int F = (int)J.outer.size();
      BatchSlabDesc a; a.in = c->site[J.v]; a.out = J.A; a.flags = 0; a.n3 = 8;
      const long long mst[5] = {1, 8, 64, 512, 4096};
      for (int i = 0; i < 3; ++i) { a.m[i] = J.root[i]; a.gi[i] = p.ostride[i]; a.go[i] = mst[i]; }
      a.gi[3] = p.ostride[3]; a.go[3] = mst[3];
      a.nc[0] = F == 5 ? 8 : 1; a.gi[4] = F == 5 ? p.ostride[4] : 0; a.go[4] = mst[4];
      a.nc[1] = 2; a.gi[5] = 1; a.go[5] = J.Po;
      a.nc[2] = J.chib; a.gi[6] = p.bstride; a.go[6] = 2 * J.Po;
      p1.push_back(a); g1 = std::max(g1, a.nc[0] * a.nc[1] * a.nc[2]);
      BatchSlabDesc b; b.in = J.A; b.out = J.A; b.flags = 0; b.n3 = 8;
      if (F == 5) {
        const long long st[3] = {1, 512, 4096};
        b.m[0] = nullptr; b.m[1] = J.root[3]; b.m[2] = J.root[4];
        for (int i = 0; i < 3; ++i) b.gi[i] = b.go[i] = st[i];
        b.gi[3] = b.go[3] = 8;
        b.nc[0] = 8; b.gi[4] = b.go[4] = 64;
        b.nc[1] = ncols[j]; b.gi[5] = b.go[5] = J.Po;
      } else {
        const long long st[3] = {1, 8, 512};
        b.m[0] = nullptr; b.m[1] = nullptr; b.m[2] = J.root[3];
        for (int i = 0; i < 3; ++i) b.gi[i] = b.go[i] = st[i];
        b.gi[3] = b.go[3] = 64;
        b.nc[0] = ncols[j]; b.gi[4] = b.go[4] = J.Po;
        b.nc[1] = 1; b.gi[5] = b.go[5] = 0;
      }
      b.nc[2] = 1; b.gi[6] = b.go[6] = 0;
      p2.push_back(b); g2 = std::max(g2, b.nc[0] * b.nc[1] * b.nc[2]);
    }
    BatchSlabDesc* d1 = c->upload(p1); BatchSlabDesc* d2 = c->upload(p2);
    LAUNCH(c, TNQS_KC_MODE_APPLY, k_slab_batch<8>, dim3((unsigned)g1, (unsigned)p1.size(), 1), DeepCfg<8>::THREADS, DeepCfg<8>::SMEM_A, d1);
    LAUNCH(c, TNQS_KC_MODE_APPLY, k_slab_batch<8>, dim3((unsigned)g2, (unsigned)p2.size(), 1), DeepCfg<8>::THREADS, DeepCfg<8>::SMEM_A, d2);
  }
  size_t maxouter = 0;
  for (int jj : generic) { SiteJob& J = jobs[jj]; maxouter = std::max(maxouter, J.outer.size()); }
  for (size_t s = 0; s < maxouter; ++s) {
    std::vector<ModeApplyDesc> ma;
    for (int j : generic) {
      SiteJob& J = jobs[j];
      if (s >= J.outer.size()) continue;
      long long Lp = 1;
      for (size_t i = 0; i < s; ++i) Lp *= c->chi[c->inc[J.v][J.outer[i]]];
      int x = c->chi[c->inc[J.v][J.outer[s]]];
      long long R = J.Po / (Lp * x) * ncols[j];
      cplx* buf = inverse ? J.Q : J.A;
      ModeApplyDesc d; d.in = buf; d.out = buf; d.m = inverse ? J.invroot[s] : J.root[s];
      d.L = (int)Lp; d.n = x; d.np = x; d.R = (int)R; d.in_sb = Lp; d.in_sr = Lp * x; d.out_sb = Lp; d.out_sr = Lp * x;
      d.ldm = x; d.flags = inverse ? 3 : 0;      // Q * dag(M^-1/2): element (c, b) = conj(Minv[b, c])
      ma.push_back(d);
    }
    launch_mode_apply(c, TNQS_KC_MODE_APPLY, ma);
  }
}

void GateBatch::householder_qr(const std::vector<int>& list) {
    if (list.empty()) return;
    for (int j : list) jobs[j].Q = jobs[j].A;
    std::vector<std::vector<QrLevel>> trees(NJ);
    size_t maxlev = 0;
    for (int j : list) {
      SiteJob& J = jobs[j];
      const long long hmax = 2 * J.nc;      // small blocks: many resident CTAs hide the latency of the column-by-column sweep
      QrLevel lv; lv.base = J.A; lv.ld = J.Po; lv.rows = J.Po;
      while (true) {
        long long nb = lv.rows <= hmax ? 1 : (lv.rows + hmax - 1) / hmax;
        lv.blk_begin.clear();
        for (long long b = 0; b <= nb; ++b) lv.blk_begin.push_back(lv.rows * b / nb);
        trees[j].push_back(lv);
        if (nb == 1) break;
        QrLevel nx; nx.rows = nb * J.nc; nx.ld = nx.rows; nx.base = c->talloc<cplx>((size_t)nx.rows * J.nc);
        lv = nx;
      }
      maxlev = std::max(maxlev, trees[j].size());
    }
    // upward sweep: factor every block, R factors stack into the next level
    for (size_t l = 0; l < maxlev; ++l) {
      std::vector<QrDesc> qd; size_t smem = 0;
      for (int j : list) {
        if (l >= trees[j].size()) continue;
        SiteJob& J = jobs[j]; QrLevel& lv = trees[j][l];
        bool top = l + 1 == trees[j].size();
        size_t nb = lv.blk_begin.size() - 1;
        for (size_t b = 0; b < nb; ++b) {
          QrDesc d; d.a = lv.base + lv.blk_begin[b]; d.lda = lv.ld; d.h = (int)(lv.blk_begin[b + 1] - lv.blk_begin[b]); d.nc = J.nc;
          if (top) { d.rout = J.Rbuf; d.ldr = std::min(d.h, J.nc); }
          else { d.rout = trees[j][l + 1].base + (long long)b * J.nc; d.ldr = (int)trees[j][l + 1].ld; }
          qd.push_back(d);
          smem = std::max(smem, ((size_t)d.h * d.nc + d.nc + 2) * sizeof(cplx));
        }
      }
      REQUIRE(smem <= 200 * 1024, TNQS_ERR_UNSUPPORTED, "QR block does not fit shared memory");
      QrDesc* dqd = c->upload(qd);
      LAUNCH(c, TNQS_KC_QR, k_house_qr, (unsigned)qd.size(), 256, smem, dqd);
    }
    // downward sweep: Q_block <- Q_block * (its nc x nc slice of the parent's explicit Q)
    for (int l = (int)maxlev - 2; l >= 0; --l) {
      std::vector<ModeApplyDesc> ma; long long maxtiles = 1;
      for (int j : list) {
        if ((size_t)l + 1 >= trees[j].size()) continue;
        SiteJob& J = jobs[j]; QrLevel& lv = trees[j][l]; QrLevel& up = trees[j][l + 1];
        size_t nb = lv.blk_begin.size() - 1;
        for (size_t b = 0; b < nb; ++b) {
          ModeApplyDesc d; d.in = lv.base + lv.blk_begin[b]; d.out = const_cast<cplx*>(d.in);
          d.m = up.base + (long long)b * J.nc; d.ldm = (int)up.ld; d.flags = 0;
          d.L = (int)(lv.blk_begin[b + 1] - lv.blk_begin[b]); d.n = J.nc; d.np = J.nc; d.R = 1;
          d.in_sb = lv.ld; d.in_sr = 0; d.out_sb = lv.ld; d.out_sr = 0;
          ma.push_back(d);
          maxtiles = std::max(maxtiles, ((long long)d.L + MA_BM - 1) / MA_BM);
        }
      }
      (void)maxtiles;
      launch_mode_apply(c, TNQS_KC_QR, ma);
    }
  }

void GateBatch::resolve_chol_flags() {
  if (!chol_pending) return;
  chol_pending = false;
  const int nj = (int)chol_list.size();
  std::vector<int> flags(nj);
  c->d2h_small(flags.data(), d_chol_flag, nj * sizeof(int));
  std::vector<int> redo;
  for (int i = 0; i < nj; ++i) if (flags[i]) { redo.push_back(chol_list[i]); jobs[chol_list[i]].colfix = nullptr; }
  householder_qr(redo);                       // overwrites Q (= A, in place) and Rbuf of the flagged matrices
}

void GateBatch::stage1(int n_, const int32_t* es_, const int32_t* flips, const double* gates, const tnqs_apply_opts& o_, const int32_t* remote) {
  n = n_; o = o_; cap = c->chi_cap; NJ = 2 * n;
  es.assign(es_, es_ + n);
  std::set<int> seen;
  jobs.assign(NJ, SiteJob());
  size_t goff = 0;
  for (int i = 0; i < n; ++i) {
    int e = es[i];
    REQUIRE(e >= 0 && e < c->ne, TNQS_ERR_INVALID_ARG, "tnqs_apply_2site: edge out of range");
    int v1 = flips && flips[i] ? c->edst[e] : c->esrc[e], v2 = c->other(e, v1);
    REQUIRE(seen.insert(v1).second && seen.insert(v2).second, TNQS_ERR_OVERLAPPING_BATCH,
            "tnqs_apply_2site: gates of a batch share a vertex");
    REQUIRE(c->d[v1] == c->d[v2], TNQS_ERR_UNSUPPORTED, "two-site gate on sites of different dimension");
    int rm = remote ? remote[i] : 0;
    REQUIRE(rm >= 0 && rm <= 2, TNQS_ERR_INVALID_ARG, "remote flag must be 0, 1 or 2");
    if (rm) cut.push_back(i);
    for (int s = 0; s < 2; ++s) {
      SiteJob& J = jobs[2 * i + s];
      J.v = s == 0 ? v1 : v2; J.e = e; J.jb = c->leg(J.v, e); J.d = c->d[J.v]; J.chib = c->chi[e];
      J.remote = rm == s + 1; J.A = J.Q = J.Rbuf = nullptr;
      J.Po = 1;
      for (size_t j = 0; j < c->inc[J.v].size(); ++j) if ((int)j != J.jb) { J.outer.push_back((int)j); J.Po *= c->chi[c->inc[J.v][j]]; }
      J.nc = J.d * J.chib; J.q = (int)std::min<long long>(J.Po, J.nc);
      REQUIRE(J.nc <= 64, TNQS_ERR_UNSUPPORTED, "d * chi > 64 is not supported by the two-site gate path");
      if (J.remote) { J.outer.clear(); J.Rbuf = c->talloc<cplx>((size_t)J.nc * J.nc); } else LJ.push_back(2 * i + s);
    }
    int d4 = c->d[v1] * c->d[v1] * c->d[v1] * c->d[v1];
    const cplx* g = (const cplx*)gates + goff; goff += d4;
    hgates.insert(hgates.end(), g, g + d4);
  }
  // ---- 1. sqrt / inverse sqrt of the incoming messages (src/utils.jl:18-26) ----
  std::vector<EighDesc> eg;
  int maxchi = 1;
  d_jobcut = c->talloc<int>(NJ);
  CU(cudaMemsetAsync(d_jobcut, 0, NJ * sizeof(int), c->stream));
  for (int jj : LJ) { SiteJob& J = jobs[jj];
    for (int j : J.outer) {
      int ee = c->inc[J.v][j], u = c->other(ee, J.v), x = c->chi[ee];
      EighDesc d; d.m = c->msg(c->dedge(u, ee)); d.n = x; d.ldm = x;
      d.root = c->talloc<cplx>((size_t)x * x); d.invroot = c->talloc<cplx>((size_t)x * x);
      d.cutoff = 10 * TNQS_EPS; d.shift = 0; d.mode = 0; d.cutflag = d_jobcut + jj;
      J.root.push_back(d.root); J.invroot.push_back(d.invroot);
      eg.push_back(d); maxchi = std::max(maxchi, x);
    } }
  if (!eg.empty()) {
    EighDesc* deg = c->upload(eg);
    int threads = 32 * std::max(1, std::min(32, (maxchi + 1) / 2));
    size_t smem = (size_t)2 * maxchi * maxchi * sizeof(cplx) + 3 * maxchi * sizeof(double);
    LAUNCH(c, TNQS_KC_EIGH, k_eigh_sqrt, (unsigned)eg.size(), threads, smem, deg, c->d_status);
  }
  // ---- 2. matricise: A[outer.., (s, b)] ----
  pd.assign(NJ, PermDesc());
  long long maxtot = 1;
  for (int j : LJ) {
    SiteJob& J = jobs[j];
    J.A = c->talloc<cplx>((size_t)c->site_n[J.v]);
    J.Rbuf = c->talloc<cplx>((size_t)J.nc * J.nc);
    PermDesc& p = pd[j];
    p.src = c->site[J.v]; p.dst = J.A; p.d = J.d; p.chib = J.chib; p.nouter = (int)J.outer.size();
    long long stride = J.d;
    for (size_t l = 0, oi = 0; l < c->inc[J.v].size(); ++l) {
      int x = c->chi[c->inc[J.v][l]];
      if ((int)l == J.jb) p.bstride = stride; else { p.odim[oi] = x; p.ostride[oi] = stride; ++oi; }
      stride *= x;
    }
    p.Po = J.Po; p.to_matrix = 1; p.norm2 = nullptr;
    maxtot = std::max(maxtot, (long long)c->site_n[J.v]);
  }
  from_site.assign(NJ, 0);
  {
    // sites whose outer legs go through the fused slab kernel are matricised by that kernel (it reads the canonical layout)
    std::vector<PermDesc> lp;
    for (int j : LJ) { if (jobs[j].d == 2 && (slab_ok(j) || deep_ok(j))) from_site[j] = 1; else lp.push_back(pd[j]); }
    PermDesc* dpd = c->upload(lp);
    long long maxPo = 1; for (int j : LJ) maxPo = std::max(maxPo, jobs[j].Po);
    unsigned gx = (unsigned)((maxPo + 255) / 256); (void)maxtot;
    if (!lp.empty()) LAUNCH(c, TNQS_KC_PERMUTE, k_permute, dim3(gx, (unsigned)lp.size()), 256, 0, dpd);
  }
  for (int j : LJ) if (!from_site[j]) dbg_norm(c, "matricised", j, jobs[j].A, (size_t)c->site_n[jobs[j].v]);
  // ---- 3. absorb sqrt(M_k) on the outer legs, in place in the matrix layout (simple_update.jl:35-36) ----
  ncols.assign(NJ, 0);
  for (int j = 0; j < NJ; ++j) ncols[j] = jobs[j].nc;
  absorb_outer(false);
  for (int j : LJ) dbg_norm(c, "absorbed", j, jobs[j].A, (size_t)c->site_n[jobs[j].v]);
  // ---- 4. thin QR with explicit Q (simple_update.jl:39-40) ----
  // Tall, well-conditioned matrices: CholeskyQR2 on the tensor cores (Gram -> Cholesky -> Q = A R^-1, twice; as
  // accurate as Householder while cond(A) < ~1e6, which the Cholesky pivots check on the device).  Everything else
  // (short, wide, rank-deficient or ill-conditioned - the norm while bonds are still growing): Householder TSQR.
  {
    std::vector<int> house_list;
    static int no_chol = -1;
    if (no_chol < 0) { const char* e = getenv("TNQS_NO_CHOLQR"); no_chol = (e && e[0] == '1') ? 1 : 0; }
    for (int j : LJ) {
      SiteJob& J = jobs[j];
      if (!no_chol && !dmma_disabled() && J.nc <= 64 && J.Po >= 4 * (long long)J.nc) chol_list.push_back(j); else house_list.push_back(j);
    }
    if (!chol_list.empty()) {
      const int nj = (int)chol_list.size();
      int* d_flag = c->talloc<int>(nj);
      CU(cudaMemsetAsync(d_flag, 0, nj * sizeof(int), c->stream));
      std::vector<cplx*> Qb(nj), G(nj), Rinv(nj), part(nj);
      int nsplit = std::max(1, std::min(16, (4 * c->num_sms + nj - 1) / nj));
      for (int i = 0; i < nj; ++i) {
        SiteJob& J = jobs[chol_list[i]];
        Qb[i] = c->talloc<cplx>((size_t)J.Po * J.nc); G[i] = c->talloc<cplx>((size_t)J.nc * J.nc);
        Rinv[i] = c->talloc<cplx>((size_t)J.nc * J.nc); part[i] = c->talloc<cplx>((size_t)nsplit * J.nc * J.nc);
      }
      // W path (optimistic batches, square R): nothing but R is used afterwards, so a matrix whose first factor is provably
      // well-conditioned skips its second pass on the device (CholDesc in kernels_dmma.cuh)
      int* d_need = nullptr;
      std::vector<char> adaptive(nj, 0);
      if (optimistic && chol_adaptive() && gate_wpath() && !dmma_disabled()) {
        bool any = false;
        for (int i = 0; i < nj; ++i) { const SiteJob& J = jobs[chol_list[i]]; adaptive[i] = J.q == J.nc && J.nc <= 64; any = any || adaptive[i]; }
        if (any) { d_need = c->talloc<int>(nj); CU(cudaMemsetAsync(d_need, 1, nj * sizeof(int), c->stream)); }
      }
      for (int pass = 0; pass < 2; ++pass) {
        std::vector<ModeCloseDesc> mc(nj); std::vector<CloseReduceDesc> rd(nj); std::vector<CholDesc> cd(nj); std::vector<ModeApplyDesc> ma(nj);
        size_t maxnn = 1; int maxn = 1;
        for (int i = 0; i < nj; ++i) {
          SiteJob& J = jobs[chol_list[i]];
          const cplx* src = pass == 0 ? J.A : Qb[i];
          int ns = (int)std::min<long long>(nsplit, std::max<long long>(1, J.Po / 256));
          ModeCloseDesc& m = mc[i];
          m.a = src; m.b = src; m.partial = part[i]; m.L = (int)J.Po; m.R = 1; m.n = J.nc; m.np = J.nc;
          m.a_sb = J.Po; m.a_sr = 0; m.b_sb = J.Po; m.b_sr = 0; m.nsplit = ns; m.rows_per_split = ((J.Po + ns - 1) / ns + 3) / 4 * 4;
          CloseReduceDesc& r = rd[i];
          r.partial = part[i]; r.nsplit = ns; r.n = J.nc; r.np = J.nc; r.dst = G[i]; r.prev = nullptr; r.diff = nullptr; r.mode = 0; r.normalize = 0; r.hermitize = 0;
          CholDesc& k = cd[i];
          k.g = G[i]; k.n = J.nc; k.rinv = Rinv[i]; k.rtot = J.Rbuf; k.first = pass == 0; k.flag = d_flag + i; k.tol = 1e-11;
          if (d_need && adaptive[i]) { if (pass == 0) k.need_out = d_need + i; else { k.need_in = d_need + i; m.need = d_need + i; } }
          ModeApplyDesc& a = ma[i];
          a.in = src; a.out = Qb[i]; a.m = Rinv[i]; a.L = (int)J.Po; a.n = J.nc; a.np = J.nc; a.R = 1;
          a.in_sb = J.Po; a.in_sr = 0; a.out_sb = J.Po; a.out_sr = 0; a.ldm = J.nc; a.flags = 0;
          if (d_need && adaptive[i]) a.commit = d_need + i;               // Q1 feeds the second Gram matrix only
          maxnn = std::max(maxnn, (size_t)J.nc * J.nc); maxn = std::max(maxn, J.nc);
        }
        ModeCloseDesc* dmc = c->upload(mc);
        LAUNCH(c, TNQS_KC_QR, k_gram_dmma, dim3(nsplit, nj), 256, 8 * 512 * sizeof(double), dmc);
        CloseReduceDesc* drd = c->upload(rd);
        LAUNCH(c, TNQS_KC_QR, k_close_reduce, nj, 256, maxnn * sizeof(cplx), drd);
        CholDesc* dcd = c->upload(cd);
        LAUNCH(c, TNQS_KC_QR, k_chol_inv, nj, 64, 2 * (size_t)maxn * maxn * sizeof(cplx), dcd);
        if (pass == 0) {                                            // Q1 = A R1^-1; the second factor R2^-1 stays pending (colfix)
          if (d_need) {                                             // the two-leg instantiation honours the commit word
            ModeApplyDesc* dma = c->upload(ma);
            ApplyLaunch al = plan_apply_launch(ma.data(), ma.size());
            REQUIRE(al.dmma, TNQS_ERR_UNSUPPORTED, "adaptive CholeskyQR2 needs the tensor-core leg kernel");
            LAUNCH(c, TNQS_KC_QR, k_leg_canon_dmma, al.grid, 256, al.smem, dma, al.tpc);
          } else launch_mode_apply(c, TNQS_KC_QR, ma);
        }
      }
      for (int i = 0; i < nj; ++i) { jobs[chol_list[i]].Q = Qb[i]; jobs[chol_list[i]].colfix = Rinv[i]; }
      d_chol_flag = d_flag; chol_pending = true;
    }
    householder_qr(house_list);
    if (!cut.empty()) resolve_chol_flags();      // the R factors of cut-crossing gates leave the device at the end of this stage
  }

  for (int j : LJ) { dbg_norm(c, "Q", j, jobs[j].Q, (size_t)jobs[j].Po * jobs[j].q); dbg_norm(c, "R", j, jobs[j].Rbuf, (size_t)jobs[j].q * jobs[j].nc); }
  // R factors of the cut-crossing gates -> exchange buffer
  if (!cut.empty()) {
    int ncmax = 1; for (const SiteJob& J : jobs) ncmax = std::max(ncmax, J.nc);
    xslot = 1 + (long long)ncmax * ncmax;
    xsend = c->talloc<cplx>((size_t)cut.size() * xslot);
    std::vector<XchgDesc> xd;
    for (size_t i = 0; i < cut.size(); ++i) {
      const SiteJob& J = jobs[2 * cut[i]].remote ? jobs[2 * cut[i] + 1] : jobs[2 * cut[i]];
      XchgDesc x; x.slot = xsend + i * xslot; x.r = J.Rbuf; x.q = J.q; x.nc = J.nc; xd.push_back(x);
    }
    XchgDesc* dxd = c->upload(xd);
    LAUNCH(c, TNQS_KC_MISC, k_xchg_pack, (unsigned)xd.size(), 256, 0, dxd);
  }
}

void GateBatch::stage2(const cplx* xrecv, int32_t* ranks_out, double* err_out, double* svals_out) {
  if (!cut.empty()) {
    REQUIRE(xrecv, TNQS_ERR_INVALID_ARG, "tnqs_apply_2site_end: the batch has remote halves but no receive buffer was given");
    std::vector<cplx> hdr(cut.size());
    cplx* d_hdr = c->talloc<cplx>(cut.size());
    LAUNCH(c, TNQS_KC_MISC, k_xchg_headers, (unsigned)((cut.size() + 127) / 128), 128, 0, xrecv, xslot, (int)cut.size(), d_hdr);
    CU(cudaMemcpyAsync(hdr.data(), d_hdr, cut.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    std::vector<XchgDesc> xd;
    for (size_t i = 0; i < cut.size(); ++i) {
      SiteJob& J = jobs[2 * cut[i]].remote ? jobs[2 * cut[i]] : jobs[2 * cut[i] + 1];
      int q = (int)hdr[i].x, nc = (int)hdr[i].y;
      REQUIRE(nc == J.nc && q >= 1 && q <= nc, TNQS_ERR_SHAPE, "remote R factor does not match the gate bond");
      J.q = q;
      XchgDesc x; x.slot = const_cast<cplx*>(xrecv) + i * xslot; x.r = J.Rbuf; x.q = q; x.nc = nc; xd.push_back(x);
    }
    XchgDesc* dxd = c->upload(xd);
    LAUNCH(c, TNQS_KC_MISC, k_xchg_unpack, (unsigned)xd.size(), 256, 0, dxd);
  }
  // ---- 5. theta = R1 R2 over the bond, gate (simple_update.jl:43) ----
  td.assign(n, ThetaDesc()); sv.assign(n, SvdDesc());
  d_rank = c->talloc<int>(n); d_err = c->talloc<double>(n); d_sv = c->talloc<double>((size_t)n * cap);
  cplx* d_gates = c->talloc<cplx>((size_t)n * 256);
  std::vector<cplx*> r1n(n), r2n(n);
  size_t svd_smem_wv = 0; int svd_maxn = 2; bool any_global = false; int last_ro = 0, last_d = 0; const cplx* last_g = nullptr;
  {
    size_t goff = 0; std::vector<cplx> hg;
    for (int i = 0; i < n; ++i) {
      SiteJob& A = jobs[2 * i]; SiteJob& B = jobs[2 * i + 1];
      int d = A.d, d4 = d * d * d * d;
      const cplx* g = hgates.data() + goff; goff += d4;
      size_t hoff = hg.size(); hg.insert(hg.end(), g, g + d4);
      ThetaDesc& t = td[i];
      t.r1 = A.Rbuf; t.q1 = A.q; t.ld1 = A.q; t.r2 = B.Rbuf; t.q2 = B.q; t.ld2 = B.q; t.gate = d_gates + hoff; t.d = d; t.chi = A.chib;
      int m = A.q * d, nn = B.q * d;
      REQUIRE(std::max(m, nn) <= TNQS_JACOBI_MAX_ROWS, TNQS_ERR_UNSUPPORTED, "two-site gate: theta has more than 256 rows (q * d) - not supported by the Jacobi SVD");
      t.theta = c->talloc<cplx>((size_t)m * nn);
      r1n[i] = c->talloc<cplx>((size_t)m * std::min(std::min(m, nn), cap));
      r2n[i] = c->talloc<cplx>((size_t)nn * std::min(std::min(m, nn), cap));
      SvdDesc& s = sv[i];
      s.theta = t.theta; s.m = m; s.n = nn; s.r1out = r1n[i]; s.r2out = r2n[i];
      s.svals = d_sv + (size_t)i * cap; s.rank = d_rank + i; s.truncerr = d_err + i; s.msg_a = nullptr; s.msg_b = nullptr;
      s.maxdim = o.maxdim; s.has_cutoff = !std::isnan(o.cutoff); s.cutoff = s.has_cutoff ? o.cutoff : 0.0;
      s.normalize = o.normalize_tensors; s.cap = cap; s.mindim = 1;
      if (svd_rowspace()) { s.rs_gate = t.gate; s.rs_r2 = B.Rbuf; s.rs_q2 = B.q; s.rs_ld2 = B.q; s.rs_d = d; s.rs_chi = A.chib; }
      int mm = std::max(m, nn), mn = std::min(m, nn);
      size_t wv = (size_t)mm * mn * sizeof(cplx);                      // only W = theta' V is iterated (kernels_small.cuh)
      if (wv + mn * 12 + 64 <= 100 * 1024) { s.work = nullptr; svd_smem_wv = std::max(svd_smem_wv, wv); }
      else {
        s.work = c->talloc<cplx>(wv / sizeof(cplx)); any_global = true;
        if (s.rs_gate && m >= nn) {              // the row-space matrix of theta' may fit shared memory although theta' does not
          if (!last_g || last_d != d || memcmp(g, last_g, (size_t)d4 * sizeof(cplx)) != 0) { last_ro = gate_operator_rank(g, d); last_g = g; last_d = d; }
          const size_t rho = (size_t)last_ro * A.chib, need = (size_t)m * rho * sizeof(cplx);
          if (2 * rho <= (size_t)nn && (s.has_cutoff || (s.maxdim > 0 && (size_t)s.maxdim <= rho)) && need + mn * 12 + 64 <= 190 * 1024)
            svd_smem_wv = std::max(svd_smem_wv, need);
        }
      }
      svd_maxn = std::max(svd_maxn, mn);
    }
    c->h2d(d_gates, hg.data(), hg.size() * sizeof(cplx));
    gates_staged = c->cap_last;
  }
  (void)any_global;
  if (!optimistic) resolve_chol_flags();        // the only host read between the stages; by now the device has (nearly) drained the QR kernels
  {
    ThetaDesc* dtd = c->upload(td);
    LAUNCH(c, TNQS_KC_SVD, k_theta, n, 256, 0, dtd);
    // ---- 6. Jacobi SVD + truncation (simple_update.jl:46-61) ----
    SvdDesc* dsv = c->upload(sv);
    int threads = 32 * std::max(1, std::min(16, (svd_maxn + 1) / 2));
    size_t smem = svd_smem_wv + (size_t)svd_maxn * 12 + 64;
    LAUNCH(c, TNQS_KC_SVD, k_svd_trunc, n, threads, smem, dsv, c->d_status, svd_smem_wv);
  }
  for (int i = 0; i < n; ++i) dbg_norm(c, "theta", i, td[i].theta, (size_t)sv[i].m * sv[i].n);
  // The kept ranks decide every shape from here on, so the host has to see them - but step 7 does not depend on them: the ranks
  // and the status word travel to a pinned page on the copy stream as soon as the SVD is done, while the main stream already
  // runs the inverse absorption; the host builds the remaining descriptors behind that kernel instead of an idle device.
  std::vector<int> ranks(n);
  int st = 0;
  if (optimistic) {
    std::vector<int> fidx(2 * n, -1);
    for (size_t i = 0; i < chol_list.size(); ++i) fidx[chol_list[i]] = (int)i;      // job index 2 * gate + side
    for (int i = 0; i < n; ++i) ranks[i] = c->chi[es[i]];
    int* d_expect = c->upload(ranks); int* d_fidx = c->upload(fidx);
    d_ok = c->talloc<int>(n);
    int* no_flags = d_chol_flag ? d_chol_flag : c->d_status;                           // never read when every index is -1
    wpath.assign(NJ, 0);
    bool any_w = false;
    if (gate_wpath() && !dmma_disabled())
      for (int j : LJ) { wpath[j] = jobs[j].q == jobs[j].nc && jobs[j].nc <= 64; any_w = any_w || wpath[j]; }
    LAUNCH(c, TNQS_KC_MISC, k_gate_ok, (n + 127) / 128, 128, 0, n, d_rank, d_expect, no_flags, d_fidx, any_w ? d_jobcut : (int*)nullptr, c->d_status, d_ok);
    for (int j = 0; j < NJ; ++j) ncols[j] = jobs[j].q;
    absorb_outer(true);
  } else {
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    int* hr = (int*)c->hret_side((size_t)(n + 1) * sizeof(int));
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, ev, 0));
    cudaEventDestroy(ev);
    CU(cudaMemcpyAsync(hr, d_rank, n * sizeof(int), cudaMemcpyDeviceToHost, c->copy_stream));
    CU(cudaMemcpyAsync(hr + n, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->copy_stream));
    // ---- 7. Q <- Q * dag(M^-1/2) on the outer legs (simple_update.jl:55-56) ----
    for (int j = 0; j < NJ; ++j) ncols[j] = jobs[j].q;
    absorb_outer(true);
    CU(cudaStreamSynchronize(c->copy_stream));
    memcpy(ranks.data(), hr, n * sizeof(int));
    st = hr[n];
    if (st) { int z = 0; c->h2d(c->d_status, &z, sizeof(int)); }
  }
  if (st & ST_NEG_EIG) throw TnqsError(TNQS_ERR_NEGATIVE_EIGENVALUE, "message eigenvalue < -10 eps in simple update");
  if (st & ST_NO_CONV) throw TnqsError(TNQS_ERR_SVD_NO_CONVERGENCE, "Jacobi eigen/SVD iteration did not converge");
  if (st & ST_RANK_CAP) throw TnqsError(TNQS_ERR_SHAPE, "kept rank exceeds chi_cap (raise chi_cap or set maxdim)");
  // ---- 8. T' = Q R' (simple_update.jl:57), normalise, back to the canonical layout ----
  std::vector<cplx*> Tm(NJ), newsite(NJ);
  std::vector<double*> w_partial(NJ, nullptr); int w_npartial = 0;      // per-CTA |T'|^2 of the in-place W products
  {
    // pending column factors of CholeskyQR2: R'' = R2^-1 R' (one small product per site instead of a pass over Q)
    std::vector<cplx*> rfin(NJ, nullptr);
    std::vector<SmallMatmulDesc> sm;
    for (int j : LJ) {
      SiteJob& J = jobs[j]; int r = ranks[j / 2];
      rfin[j] = (j & 1) ? r2n[j / 2] : r1n[j / 2];
      if (!J.colfix || (!wpath.empty() && wpath[j])) continue;
      SmallMatmulDesc x; x.a = J.colfix; x.b = rfin[j]; x.m = J.nc; x.k = J.q; x.n = J.d * r; x.lda = J.nc; x.ldb = J.q; x.ldc = J.nc; x.transA = 0;
      x.c = c->talloc<cplx>((size_t)J.nc * J.d * r);
      rfin[j] = x.c; sm.push_back(x);
    }
    if (!sm.empty()) { SmallMatmulDesc* dsm = c->upload(sm); LAUNCH(c, TNQS_KC_MISC, k_small_matmul, (unsigned)sm.size(), 256, 0, dsm); }
    if (!wpath.empty()) {
      std::vector<TriSolveDesc> ts; std::vector<ModeApplyDesc> mw; size_t tsmem = 0;
      for (int j : LJ) {
        if (!wpath[j]) continue;
        SiteJob& J = jobs[j]; const int r = ranks[j / 2], v = J.v;
        TriSolveDesc t; t.r = J.Rbuf; t.ldr = J.q; t.b = (j & 1) ? r2n[j / 2] : r1n[j / 2]; t.ldb = J.q; t.n = J.nc; t.m = J.d * r;
        t.w = c->talloc<cplx>((size_t)J.nc * J.d * r); t.ldw = J.nc;
        ts.push_back(t); tsmem = std::max(tsmem, ((size_t)t.n * t.n + (size_t)t.n * t.m) * sizeof(cplx));
        long long Llo = 1, hi = 1;
        for (int l = 0; l < (int)c->inc[v].size(); ++l) { const int x = c->chi[c->inc[v][l]]; if (l < J.jb) Llo *= x; else if (l > J.jb) hi *= x; }
        ModeApplyDesc d; d.in = c->site[v]; d.out = c->site[v]; d.m = t.w; d.ldm = J.nc; d.flags = 0;
        d.L = (int)Llo; d.R = (int)hi; d.n = J.nc; d.np = J.d * r; d.kd = J.d;
        d.in_sl = J.d; d.in_sb = J.d * Llo; d.in_sr = (long long)J.d * Llo * J.chib;
        d.out_sl = J.d; d.out_sb = J.d * Llo; d.out_sr = (long long)J.d * Llo * r;
        d.commit = d_ok + j / 2;
        mw.push_back(d);
      }
      if (!ts.empty()) {
        REQUIRE(tsmem <= 160 * 1024, TNQS_ERR_UNSUPPORTED, "triangular solve does not fit shared memory");
        TriSolveDesc* dts = c->upload(ts);
        LAUNCH(c, TNQS_KC_MISC, k_tri_solve, (unsigned)ts.size(), 256, tsmem, dts);
        ApplyLaunch al = plan_apply_launch(mw.data(), mw.size());
        REQUIRE(al.dmma, TNQS_ERR_UNSUPPORTED, "internal: the in-place two-leg product needs the tensor-core leg kernel");
        if (o.normalize_tensors) {                    // the kernel leaves |T'|^2 per CTA: no separate pass over the tensor for the norm
          w_npartial = (int)al.grid.x;
          double* part = c->talloc<double>((size_t)w_npartial * mw.size());
          size_t k = 0;
          for (int j : LJ) if (wpath[j]) { w_partial[j] = part + k * w_npartial; mw[k].norm2_partial = w_partial[j]; ++k; }
        }
        ModeApplyDesc* dmw = c->upload(mw);
        LAUNCH(c, TNQS_KC_MODE_APPLY, k_leg_canon_dmma, al.grid, 256, al.smem, dmw, al.tpc);
      }
    }
    std::vector<ModeApplyDesc> ma; long long maxtiles = 1; int maxcol = 1;
    for (int j : LJ) {
      if (!wpath.empty() && wpath[j]) continue;
      SiteJob& J = jobs[j]; int r = ranks[j / 2];
      Tm[j] = c->talloc<cplx>((size_t)J.Po * J.d * r);
      ma.emplace_back(); ModeApplyDesc& d = ma.back();
      d.in = J.Q; d.out = Tm[j]; d.m = rfin[j];
      d.L = (int)J.Po; d.n = J.q; d.np = J.d * r; d.R = 1; d.in_sb = J.Po; d.in_sr = 0; d.out_sb = J.Po; d.out_sr = 0;
      d.ldm = J.q; d.flags = 0;
      maxtiles = std::max(maxtiles, (J.Po + MA_BM - 1) / MA_BM); maxcol = std::max(maxcol, (d.np + MA_BN - 1) / MA_BN);
    }
    (void)maxtiles; (void)maxcol;
    launch_mode_apply(c, TNQS_KC_MODE_APPLY, ma);
  }
  double* d_n2 = c->talloc<double>(NJ);
  if (o.normalize_tensors) {
    std::vector<Norm2Desc> nd;
    std::vector<ScaleDesc> sw;
    for (int j : LJ) {
      const bool w = !wpath.empty() && wpath[j];
      Norm2Desc x; x.x = w ? c->site[jobs[j].v] : Tm[j]; x.n = jobs[j].Po * jobs[j].d * ranks[j / 2]; x.out = d_n2 + j; x.partial = nullptr; x.npartial = 0;
      if (w && w_partial[j]) { x.partial = w_partial[j]; x.npartial = w_npartial; }
      nd.push_back(x);
      if (w) { ScaleDesc sc; sc.x = c->site[jobs[j].v]; sc.n = x.n; sc.scalar = d_n2 + j; sc.mode = 0; sc.commit = d_ok + j / 2; sw.push_back(sc); }
    }
    Norm2Desc* dnd = c->upload(nd);
    if (!nd.empty()) LAUNCH(c, TNQS_KC_MISC, k_norm2, (unsigned)nd.size(), 1024, 0, dnd);
    if (!sw.empty()) { ScaleDesc* dsw = c->upload(sw); LAUNCH(c, TNQS_KC_MISC, k_scale, dim3(64, (unsigned)sw.size()), 256, 0, dsw); }
  }
  {
    long long maxtot2 = 1;
    for (int j : LJ) {
      if (!wpath.empty() && wpath[j]) { newsite[j] = c->site[jobs[j].v]; continue; }
      SiteJob& J = jobs[j]; int r = ranks[j / 2];
      long long tot = J.Po * J.d * r;
      // same size as before (saturated bonds): the old tensor is dead by now, write over it and keep its address, so
      // cached BP plans (which hold site pointers) stay valid
      REQUIRE(!optimistic || tot == c->site_n[J.v], TNQS_ERR_INVALID_ARG, "internal: optimistic gate batch on an unsaturated bond");
      newsite[j] = tot == c->site_n[J.v] ? c->site[J.v] : (cplx*)c->dalloc((size_t)tot * sizeof(cplx));
      PermDesc& p = pd[j];
      p.src = Tm[j]; p.dst = newsite[j]; p.chib = r; p.to_matrix = 0; p.norm2 = o.normalize_tensors ? d_n2 + j : nullptr;
      p.commit = optimistic ? d_ok + j / 2 : nullptr;
      long long stride = J.d;
      for (size_t l = 0, oi = 0; l < c->inc[J.v].size(); ++l) {
        int x = (int)l == J.jb ? r : c->chi[c->inc[J.v][l]];
        if ((int)l == J.jb) p.bstride = stride; else { p.odim[oi] = x; p.ostride[oi] = stride; ++oi; }
        stride *= x;
      }
      maxtot2 = std::max(maxtot2, tot);
    }
    std::vector<PermDesc> lp; for (int j : LJ) if (wpath.empty() || !wpath[j]) lp.push_back(pd[j]);
    PermDesc* dpd = c->upload(lp);
    long long maxPo = 1; for (int j : LJ) maxPo = std::max(maxPo, jobs[j].Po);
    unsigned gx = (unsigned)((maxPo + 255) / 256); (void)maxtot2;
    if (!lp.empty()) LAUNCH(c, TNQS_KC_PERMUTE, k_permute, dim3(gx, (unsigned)lp.size()), 256, 0, dpd);
  }
  if (optimistic) {
    // nothing changes shape or address in this mode (rank = old chi, every tensor rewritten in place); the device decides per gate
    std::vector<int> ks(es.begin(), es.end());
    int* dks = c->upload(ks);
    LAUNCH(c, TNQS_KC_MISC, k_msg_diag, n, 128, 0, c->msg_cur, c->slot, dks, d_sv, d_rank, cap, d_ok);
    if (!defer_collect) collect_optimistic(ranks_out, err_out, svals_out);
    return;
  }
  // ---- 9. commit: bond dimensions, site pointers, messages <- diag(S) (apply_gates.jl:105-119) ----
  bool layout_changed = false;
  for (int i = 0; i < n; ++i) { if (c->chi[es[i]] != ranks[i]) layout_changed = true; c->chi[es[i]] = ranks[i]; }
  if (layout_changed) upload_chi(c);
  for (int j : LJ) {
    const int v = jobs[j].v;
    if (newsite[j] != c->site[v]) { c->dfree(c->site[v]); c->site[v] = newsite[j]; layout_changed = true; }
    c->site_n[v] = jobs[j].Po * jobs[j].d * ranks[j / 2];
  }
  for (int j = 0; j < NJ; ++j) {                      // ghost placeholders follow the new bond dimension
    if (!jobs[j].remote) continue;
    int v = jobs[j].v;
    if (c->site_n[v] == c->site_elems(v)) continue;
    c->dfree(c->site[v]); c->site_n[v] = c->site_elems(v);
    c->site[v] = (cplx*)c->dalloc((size_t)c->site_n[v] * sizeof(cplx));
    CU(cudaMemsetAsync(c->site[v], 0, (size_t)c->site_n[v] * sizeof(cplx), c->stream));
    layout_changed = true;
  }
  if (layout_changed) ++c->state_version;
  {
    std::vector<int> ks(es.begin(), es.end());
    int* dks = c->upload(ks);
    LAUNCH(c, TNQS_KC_MISC, k_msg_diag, n, 128, 0, c->msg_cur, c->slot, dks, d_sv, d_rank, cap);
  }

  if (ranks_out) for (int i = 0; i < n; ++i) ranks_out[i] = ranks[i];
  // truncation errors and singular values: one pinned staging copy each, one synchronisation for everything
  const size_t eb = err_out ? n * sizeof(double) : 0, sb = svals_out ? (size_t)n * cap * sizeof(double) : 0;
  char* hb = (eb + sb) ? (char*)c->hret_side(eb + sb) : nullptr;
  if (eb) CU(cudaMemcpyAsync(hb, d_err, eb, cudaMemcpyDeviceToHost, c->stream));
  if (sb) CU(cudaMemcpyAsync(hb + eb, d_sv, sb, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  if (eb) memcpy(err_out, hb, eb);
  if (sb) memcpy(svals_out, hb + eb, sb);
}

// Results of an optimistic batch (after its kernels were enqueued, eagerly or by a graph launch): which gates committed, truncation
// errors, singular values - one staged copy, one synchronisation.
static void collect_optimistic_results(tnqs_ctx* c, int n, int cap, const int* es, const int* d_ok, const double* d_err, const double* d_sv,
                                       int32_t* ranks_out, double* err_out, double* svals_out, std::vector<int>& failed) {
  const size_t ob = n * sizeof(int), eb = err_out ? n * sizeof(double) : 0, sb = svals_out ? (size_t)n * cap * sizeof(double) : 0;
  char* hb = (char*)c->hret_side(ob + sizeof(int) + eb + sb + 16);
  CU(cudaMemcpyAsync(hb, d_ok, ob, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(hb + ob, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  const size_t off = (ob + sizeof(int) + 15) / 16 * 16;
  if (eb) CU(cudaMemcpyAsync(hb + off, d_err, eb, cudaMemcpyDeviceToHost, c->stream));
  if (sb) CU(cudaMemcpyAsync(hb + off + eb, d_sv, sb, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  const int st = *(int*)(hb + ob);
  if (st) { int z = 0; c->h2d(c->d_status, &z, sizeof(int)); }
  if (st & ST_NEG_EIG) throw TnqsError(TNQS_ERR_NEGATIVE_EIGENVALUE, "message eigenvalue < -10 eps in simple update");
  if (st & ST_NO_CONV) throw TnqsError(TNQS_ERR_SVD_NO_CONVERGENCE, "Jacobi eigen/SVD iteration did not converge");
  if (st & ST_RANK_CAP) throw TnqsError(TNQS_ERR_SHAPE, "kept rank exceeds chi_cap (raise chi_cap or set maxdim)");
  const int* ok = (const int*)hb;
  for (int i = 0; i < n; ++i) if (!ok[i]) failed.push_back(i);
  if (ranks_out) for (int i = 0; i < n; ++i) ranks_out[i] = c->chi[es[i]];
  if (eb) memcpy(err_out, hb + off, eb);
  if (sb) memcpy(svals_out, hb + off + eb, sb);
}
void GateBatch::collect_optimistic(int32_t* ranks_out, double* err_out, double* svals_out) {
  collect_optimistic_results(c, n, cap, es.data(), d_ok, d_err, d_sv, ranks_out, err_out, svals_out, failed);
}

// A captured optimistic batch.  In the steady state of a truncated evolution a colour class is the same launch sequence over the
// same addresses at every layer (site tensors rewritten in place, temporaries from the same positions of the bump arena, bond
// dimensions unchanged), so after it has run once eagerly with every gate committed it is captured - descriptor uploads included,
// from a pinned buffer the graph owns - and replayed: one launch and one read-back per class instead of ~25 launches and ~20
// uploads, and nothing for a preempted host thread to stall between them.  Only the gate matrices are refreshed before a replay.
struct GateGraph {
  std::vector<int> es, fl; tnqs_apply_opts o; long long version = -1; int n = 0, cap = 0;
  cudaGraphExec_t exec = nullptr; char* stage = nullptr; size_t stage_size = 0; char* gates_staged = nullptr; size_t gates_bytes = 0;
  int* d_ok = nullptr; double* d_err = nullptr; double* d_sv = nullptr; long long nlaunch = 0; int seen_ok = 0; bool no_graph = false;
  size_t arena_chunk0 = 0, arena_off0 = 0; char* arena_base0 = nullptr;      // where the bump arena stood when the batch was captured
};
// The temporaries of a captured batch are baked into it as addresses of the bump arena: a replay is only valid from the same arena
// position (every entry point releases its temporaries before it returns, so this is the start of the arena in practice).
static bool arena_at(const tnqs_ctx* c, const GateGraph* gg) {
  return c->arena_chunk == gg->arena_chunk0 && c->arena_off == gg->arena_off0 && !c->arena.empty() && c->arena[0].base == gg->arena_base0;
}
// The upload stages of all captured batches of a handle are slices of ONE pinned pool (page-locking memory costs tens of
// milliseconds per call on these hosts): bump-allocated, rewound when the last graph of the handle is gone.
static void gate_graph_free(GateGraph* gg) {
  if (gg->exec) cudaGraphExecDestroy(gg->exec);
  delete gg;
}
static void gate_graphs_drop(tnqs_ctx* c) {
  for (void* p : c->gate_graphs) gate_graph_free((GateGraph*)p);
  c->gate_graphs.clear(); c->gg_pool_off = 0;
}
static char* gate_graph_stage(tnqs_ctx* c, size_t bytes) {
  if (!c->gg_pool) {
    const size_t sz = (size_t)48 << 20;
    if (cudaHostAlloc((void**)&c->gg_pool, sz, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); c->gg_pool = nullptr; return nullptr; }
    c->gg_pool_size = sz; c->gg_pool_off = 0;
  }
  if (c->gg_pool_off + bytes > c->gg_pool_size) return nullptr;
  char* p = c->gg_pool + c->gg_pool_off; c->gg_pool_off += bytes;
  return p;
}
static bool same_opts(const tnqs_apply_opts& a, const tnqs_apply_opts& b) {
  return a.maxdim == b.maxdim && a.normalize_tensors == b.normalize_tensors && memcmp(&a.cutoff, &b.cutoff, sizeof(double)) == 0;
}
static GateGraph* gate_graph_find(tnqs_ctx* c, int n, const int32_t* es, const int32_t* fl, const tnqs_apply_opts& o) {
  GateGraph* hit = nullptr;
  for (size_t i = 0; i < c->gate_graphs.size();) {
    GateGraph* gg = (GateGraph*)c->gate_graphs[i];
    if (gg->version != c->state_version) {                 // a shape or an address changed
      gate_graph_free(gg); c->gate_graphs.erase(c->gate_graphs.begin() + i);
      if (c->gate_graphs.empty()) c->gg_pool_off = 0;
      continue;
    }
    if (!hit && gg->n == n && same_opts(gg->o, o) && memcmp(gg->es.data(), es, n * sizeof(int)) == 0) {
      bool same = true;
      for (int k = 0; k < n && same; ++k) same = gg->fl[k] == (fl ? fl[k] : 0);
      if (same) hit = gg;
    }
    ++i;
  }
  return hit;
}
static bool gate_graphs_enabled(tnqs_ctx* c) {
  if (c->profiling || graphs_disabled()) return false;
  if (getenv("TNQS_GATE_GRAPH_OFF_FOR_TEST")) return false;          // read on every call (the other switches are latched at first use)
  static int off = -1;
  if (off < 0) { const char* e = getenv("TNQS_GATE_GRAPH"); const char* d = getenv("TNQS_DEBUG"); off = ((e && e[0] == '0') || (d && d[0] == '1')) ? 1 : 0; }
  return off == 0;
}

static bool gate_wpath() { const char* e = getenv("TNQS_GATE_WPATH"); return !(e && e[0] == '0'); }
static bool gate_optimistic() { const char* e = getenv("TNQS_GATE_OPTIMISTIC"); return !(e && e[0] == '0'); }

static void apply_2site_impl(tnqs_ctx* c, int n, const int32_t* es, const int32_t* flips, const double* gates,
                             const tnqs_apply_opts& o, int32_t* ranks_out, double* err_out, double* svals_out) {
  if (n == 0) return;
  // The gates of a batch are independent (vertex-disjoint; none reads a message another one writes), so a batch whose
  // temporaries (about four copies of the touched site tensors) would not fit is processed in chunks.
  std::set<int> seen;
  for (int i = 0; i < n; ++i) {
    REQUIRE(es[i] >= 0 && es[i] < c->ne, TNQS_ERR_INVALID_ARG, "tnqs_apply_2site: edge out of range");
    REQUIRE(seen.insert(c->esrc[es[i]]).second && seen.insert(c->edst[es[i]]).second, TNQS_ERR_OVERLAPPING_BATCH,
            "tnqs_apply_2site: gates of a batch share a vertex");
  }
  const size_t budget = scratch_budget(c);
  const cplx* g = (const cplx*)gates;
  int i0 = 0; size_t goff = 0;
  while (i0 < n) {
    size_t bytes = 0, gcount = 0; int i1 = i0;
    while (i1 < n) {
      int e = es[i1], d = c->d[c->esrc[e]];
      size_t need = (size_t)5 * (c->site_n[c->esrc[e]] + c->site_n[c->edst[e]]) * sizeof(cplx);
      if (i1 > i0 && bytes + need > budget) break;
      bytes += need; gcount += (size_t)d * d * d * d; ++i1;
    }
    const int nb = i1 - i0;
    // saturated bonds (chi = maxdim on every edge of the chunk): the steady state of a truncated evolution, where a gate keeps its
    // bond dimension - run the chunk optimistically (one synchronisation, GateBatch::optimistic) and redo what did not hold
    bool optimistic = gate_optimistic() && o.maxdim > 0 && o.maxdim <= c->chi_cap;
    for (int i = i0; i < i1 && optimistic; ++i) optimistic = c->chi[es[i]] == o.maxdim;
    std::vector<int> failed;
    int32_t* r_out = ranks_out ? ranks_out + i0 : nullptr; double* e_out = err_out ? err_out + i0 : nullptr;
    double* s_out = svals_out ? svals_out + (size_t)i0 * c->chi_cap : nullptr;
    const int32_t* fl0 = flips ? flips + i0 : nullptr;
    GateGraph* gg = (optimistic && gate_graphs_enabled(c)) ? gate_graph_find(c, nb, es + i0, fl0, o) : nullptr;
    bool done = false;
    if (gg && gg->exec && !arena_at(c, gg)) gg = nullptr;   // (not from here: run it eagerly, leave the graph alone)
    if (gg && gg->exec) {                                  // replay
      static const char* dbg = getenv("TNQS_DEBUG_QUEUE");
      if (dbg) fprintf(stderr, "[tnqs] gate batch of %d: graph replay (%lld kernels)\n", nb, gg->nlaunch);
      memcpy(gg->gates_staged, g + goff, gg->gates_bytes);
      CU(cudaGraphLaunch(gg->exec, c->stream));
      c->launches += gg->nlaunch;
      collect_optimistic_results(c, nb, gg->cap, es + i0, gg->d_ok, gg->d_err, gg->d_sv, r_out, e_out, s_out, failed);
      done = true;
    } else if (gg && gg->seen_ok >= 1 && !gg->no_graph) {  // ran eagerly with every gate committed: capture it now
      const long long l0 = c->launches;
      if (!gg->stage) {
        // descriptor bytes of a batch: a few hundred per site job (measured 2.4 KB per gate at C2), rounded up generously
        gg->stage_size = ((size_t)nb * 6144 + ((size_t)1 << 20) + 4095) / 4096 * 4096;
        gg->stage = gate_graph_stage(c, gg->stage_size);
        if (!gg->stage) gg->no_graph = true;
      }
      if (gg->stage) {
        cudaGraph_t graph = nullptr;
        GateBatch gb; gb.c = c; gb.optimistic = true; gb.defer_collect = true;
        c->cap_stage = gg->stage; c->cap_stage_size = gg->stage_size; c->cap_stage_off = 0; c->cap_last = nullptr;
        gg->arena_chunk0 = c->arena_chunk; gg->arena_off0 = c->arena_off; gg->arena_base0 = c->arena.empty() ? nullptr : c->arena[0].base;
        bool ok_cap = gg->arena_base0 && cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok_cap) {
          c->capturing = true;
          try { gb.stage1(nb, es + i0, fl0, (const double*)(g + goff), o, nullptr); gb.stage2(nullptr, nullptr, nullptr, nullptr); }
          catch (...) { ok_cap = false; }
          c->capturing = false;
          if (cudaStreamEndCapture(c->stream, &graph) != cudaSuccess || !graph) ok_cap = false;
        }
        c->cap_stage = nullptr;
        if (ok_cap && cudaGraphInstantiate(&gg->exec, graph, 0) != cudaSuccess) { gg->exec = nullptr; ok_cap = false; }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (ok_cap && gb.gates_staged) {
          static const char* dbg = getenv("TNQS_DEBUG_QUEUE");
          if (dbg) fprintf(stderr, "[tnqs] gate batch of %d captured: %zu of %zu staged bytes\n", nb, c->cap_stage_off, gg->stage_size);
          gg->nlaunch = c->launches - l0; c->launches = l0;          // the capture pass launched nothing
          gg->d_ok = gb.d_ok; gg->d_err = gb.d_err; gg->d_sv = gb.d_sv; gg->cap = gb.cap; gg->gates_staged = gb.gates_staged; gg->gates_bytes = gcount * sizeof(cplx);
          CU(cudaGraphLaunch(gg->exec, c->stream));
          c->launches += gg->nlaunch;
          collect_optimistic_results(c, nb, gg->cap, es + i0, gg->d_ok, gg->d_err, gg->d_sv, r_out, e_out, s_out, failed);
          done = true;
        } else {
          c->launches = l0;
          if (gg->exec) { cudaGraphExecDestroy(gg->exec); gg->exec = nullptr; }
          gg->no_graph = true;                                 // e.g. the arena would have had to grow: stay eager for this batch
          c->release_temps();
        }
      }
    }
    if (!done) {
      GateBatch gb; gb.c = c; gb.optimistic = optimistic;
      gb.stage1(nb, es + i0, fl0, (const double*)(g + goff), o, nullptr);
      gb.stage2(nullptr, r_out, e_out, s_out);
      failed.swap(gb.failed);
      if (optimistic && gate_graphs_enabled(c) && failed.empty()) {
        if (!gg) gg = gate_graph_find(c, nb, es + i0, fl0, o);
        if (!gg) {
          if (c->gate_graphs.size() >= 16) gate_graphs_drop(c);       // far more distinct batches than a layer has classes: start over
          gg = new GateGraph(); gg->es.assign(es + i0, es + i1); gg->fl.resize(nb);
          for (int k = 0; k < nb; ++k) gg->fl[k] = fl0 ? fl0[k] : 0;
          gg->o = o; gg->n = nb; gg->version = c->state_version;
          c->gate_graphs.push_back(gg);
        }
        ++gg->seen_ok;
      }
    }
    if (!failed.empty()) {                     // rank below the old bond dimension or an ill-conditioned QR: those gates again, conventionally
      c->release_temps();
      const int nf = (int)failed.size();
      std::vector<int> fe(nf), ff(nf), fr(nf); std::vector<double> ferr(nf), fsv((size_t)nf * c->chi_cap); std::vector<cplx> fg;
      size_t off = goff;
      std::vector<size_t> goffs(nb);
      for (int i = 0; i < nb; ++i) { goffs[i] = off; int d = c->d[c->esrc[es[i0 + i]]]; off += (size_t)d * d * d * d; }
      for (int k = 0; k < nf; ++k) {
        const int i = failed[k], d = c->d[c->esrc[es[i0 + i]]];
        fe[k] = es[i0 + i]; ff[k] = flips ? flips[i0 + i] : 0;
        fg.insert(fg.end(), g + goffs[i], g + goffs[i] + (size_t)d * d * d * d);
      }
      GateBatch gb; gb.c = c;
      gb.stage1(nf, fe.data(), ff.data(), (const double*)fg.data(), o, nullptr);
      gb.stage2(nullptr, fr.data(), ferr.data(), fsv.data());
      for (int k = 0; k < nf; ++k) {
        const int i = i0 + failed[k];
        if (ranks_out) ranks_out[i] = fr[k];
        if (err_out) err_out[i] = ferr[k];
        if (svals_out) memcpy(svals_out + (size_t)i * c->chi_cap, fsv.data() + (size_t)k * c->chi_cap, c->chi_cap * sizeof(double));
      }
    }
    if (i1 < n) c->release_temps();
    goff += gcount; i0 = i1;
  }
}

extern "C" int32_t tnqs_apply_2site(tnqs_ctx* c, int32_t n, const int32_t* e, const int32_t* flip, const double* gates,
                                    const tnqs_apply_opts* opts, int32_t* ranks, double* truncerr, double* svals) {
  API_BEGIN(c)
  REQUIRE(n >= 0 && opts && (n == 0 || (e && gates)), TNQS_ERR_INVALID_ARG, "tnqs_apply_2site: null argument");
  apply_2site_impl(c, n, e, flip, gates, *opts, ranks, truncerr, svals);
  c->release_temps();
  API_END(c)
}

static void drop_pending_gate(tnqs_ctx* c) {
  if (c->pending_gate) { delete (GateBatch*)c->pending_gate; c->pending_gate = nullptr; }
}

extern "C" int32_t tnqs_apply_2site_begin(tnqs_ctx* c, int32_t n, const int32_t* e, const int32_t* flip, const double* gates,
                                          const tnqs_apply_opts* opts, const int32_t* remote, void** send_buf, int64_t* n_cut,
                                          int64_t* slot_complex) {
  API_BEGIN(c)
  REQUIRE(n >= 1 && opts && e && gates, TNQS_ERR_INVALID_ARG, "tnqs_apply_2site_begin: null argument");
  REQUIRE(!c->pending_gate, TNQS_ERR_INVALID_ARG, "tnqs_apply_2site_begin: a batch is already pending");
  GateBatch* gb = new GateBatch(); gb->c = c;
  c->pending_gate = gb;
  try { gb->stage1(n, e, flip, gates, *opts, remote); c->sync(); }
  catch (...) { delete gb; c->pending_gate = nullptr; throw; }
  if (send_buf) *send_buf = gb->xsend;
  if (n_cut) *n_cut = (int64_t)gb->cut.size();
  if (slot_complex) *slot_complex = gb->xslot;
  API_END(c)
}

extern "C" int32_t tnqs_apply_2site_end(tnqs_ctx* c, const void* recv_buf, int32_t* ranks, double* truncerr, double* svals) {
  API_BEGIN(c)
  REQUIRE(c->pending_gate, TNQS_ERR_INVALID_ARG, "tnqs_apply_2site_end: no pending batch");
  GateBatch* gb = (GateBatch*)c->pending_gate;
  c->pending_gate = nullptr;
  try { gb->stage2((const cplx*)recv_buf, ranks, truncerr, svals); }
  catch (...) { delete gb; throw; }
  delete gb;
  c->release_temps();
  API_END(c)
}

static void apply_circuit_impl(tnqs_ctx* c, int32_t ngates, const int32_t* gate_v, const int64_t* gate_off,
                               const double* gate_data, const tnqs_apply_opts* aopts, const int32_t* edge_seq,
                               int32_t nseq, const tnqs_bp_opts* bopts, int32_t update_cache, double* truncerr,
                               int32_t* refreshes, int32_t fetch_count, const int32_t* fetch_v, double* const* fetch_data, int64_t* fetch_n) {
  REQUIRE(ngates >= 0 && aopts && bopts && (ngates == 0 || (gate_v && gate_off && gate_data)), TNQS_ERR_INVALID_ARG,
          "tnqs_apply_circuit: null argument");
  const cplx* gd = (const cplx*)gate_data;
  std::vector<char> affected(c->nv, 0), inbatch(c->nv, 0);
  std::vector<int> batch;          // gate indices of the pending batch (all 1-site or all 2-site, vertex-disjoint)
  // Consecutive one-site gates on a vertex (the Rx and Rz of a Trotter layer) are multiplied on the host and applied as one gate:
  // one pass over the tensor and one normalisation instead of one per gate.  With normalize_tensors the intermediate norms are
  // scalars that the last normalisation removes, so only rounding changes.  folded[j] = product for batch[j].
  std::vector<std::vector<cplx>> folded;
  std::vector<int> pos1(c->nv, -1);
  int batch_arity = 0, nrefresh = 0;
  auto flush = [&]() {
    if (batch.empty()) return;
    if (batch_arity == 1) {
      std::vector<int> vs; std::vector<cplx> g;
      for (size_t j = 0; j < batch.size(); ++j) {
        const int i = batch[j], v = gate_v[2 * i], d = c->d[v];
        vs.push_back(v); pos1[v] = -1;
        if (j < folded.size() && !folded[j].empty()) g.insert(g.end(), folded[j].begin(), folded[j].end());
        else g.insert(g.end(), gd + gate_off[i], gd + gate_off[i] + d * d);
      }
      folded.clear();
      apply_1site_impl(c, (int)vs.size(), vs.data(), (const double*)g.data(), aopts->normalize_tensors);
    } else {
      std::vector<int> es, fl; std::vector<cplx> g; std::vector<double> err(batch.size());
      for (int i : batch) {
        int v1 = gate_v[2 * i], v2 = gate_v[2 * i + 1], e = -1;
        for (int ee : c->inc[v1]) if (c->other(ee, v1) == v2) e = ee;
        REQUIRE(e >= 0, TNQS_ERR_INVALID_ARG, "two-site gate on vertices that do not share an edge");
        int d = c->d[v1];
        es.push_back(e); fl.push_back(c->esrc[e] == v1 ? 0 : 1);
        g.insert(g.end(), gd + gate_off[i], gd + gate_off[i] + d * d * d * d);
      }
      apply_2site_impl(c, (int)es.size(), es.data(), fl.data(), (const double*)g.data(), *aopts, nullptr, err.data(), nullptr);
      if (truncerr) for (size_t j = 0; j < batch.size(); ++j) truncerr[batch[j]] = err[j];
    }
    c->release_temps();
    for (int i : batch) { inbatch[gate_v[2 * i]] = 0; if (gate_v[2 * i + 1] >= 0) inbatch[gate_v[2 * i + 1]] = 0; }
    batch.clear();
  };
  auto refresh = [&]() {
    int it; double md;
    bp_update_impl(c, edge_seq, nseq, *bopts, &it, &md);
    c->release_temps();
    ++nrefresh;
  };
  for (int i = 0; i < ngates; ++i) {
    int v1 = gate_v[2 * i], v2 = gate_v[2 * i + 1];
    REQUIRE(v1 >= 0 && v1 < c->nv && v2 < c->nv && v1 != v2, TNQS_ERR_INVALID_ARG, "gate vertex out of range");
    int arity = v2 >= 0 ? 2 : 1;
    if (truncerr) truncerr[i] = 0.0;
    bool need = update_cache && arity == 2 && (affected[v1] || affected[v2]);        // apply_gates.jl:64
    if (need) { flush(); refresh(); std::fill(affected.begin(), affected.end(), 0); }
    if (arity == 1 && !batch.empty() && batch_arity == 1 && pos1[v1] >= 0) {
      // o[s', s] <- g_i[s', t] * o[t, s]
      const int j = pos1[v1], d = c->d[v1];
      folded.resize(batch.size());
      if (folded[j].empty()) folded[j].assign(gd + gate_off[batch[j]], gd + gate_off[batch[j]] + d * d);
      const cplx* gi = gd + gate_off[i];
      std::vector<cplx> prod((size_t)d * d, cmake(0, 0));
      for (int s2 = 0; s2 < d; ++s2) for (int t = 0; t < d; ++t) for (int s1 = 0; s1 < d; ++s1) {
        const cplx a = gi[s1 + d * t], b = folded[j][t + d * s2];
        prod[s1 + d * s2].x += a.x * b.x - a.y * b.y; prod[s1 + d * s2].y += a.x * b.y + a.y * b.x;
      }
      folded[j].swap(prod);
      continue;
    }
    bool conflict = !batch.empty() && (arity != batch_arity || inbatch[v1] || (v2 >= 0 && inbatch[v2]));
    if (conflict) flush();
    batch_arity = arity; batch.push_back(i);
    if (arity == 1) pos1[v1] = (int)batch.size() - 1;
    inbatch[v1] = 1; affected[v1] = 1;                                                // apply_gates.jl:84
    if (v2 >= 0) { inbatch[v2] = 1; affected[v2] = 1; }
  }
  flush();
  // Fetch: the site tensors are final once the last batch is done; their device -> host copies run on a second stream while
  // the closing BP refresh (which only reads the sites) runs on the main one.
  if (fetch_count > 0) {
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, ev, 0));
    cudaEventDestroy(ev);
    for (int i = 0; i < fetch_count; ++i) {
      const int v = fetch_v[i];
      REQUIRE(v >= 0 && v < c->nv && fetch_data[i], TNQS_ERR_INVALID_ARG, "tnqs_apply_circuit_fetch: bad vertex or null buffer");
      REQUIRE(fetch_n[i] >= c->site_n[v], TNQS_ERR_SHAPE, "tnqs_apply_circuit_fetch: host buffer smaller than the site tensor");
      fetch_n[i] = c->site_n[v];
      CU(cudaMemcpyAsync(fetch_data[i], c->site[v], (size_t)c->site_n[v] * sizeof(cplx), cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  try {
    if (update_cache) refresh();
    c->sync();
  } catch (...) { if (fetch_count > 0) cudaStreamSynchronize(c->copy_stream); throw; }     // no copy may outlive the call
  if (refreshes) *refreshes = nrefresh;
  if (fetch_count > 0) CU(cudaStreamSynchronize(c->copy_stream));
}

extern "C" int32_t tnqs_apply_circuit(tnqs_ctx* c, int32_t ngates, const int32_t* gate_v, const int64_t* gate_off,
                                      const double* gate_data, const tnqs_apply_opts* aopts, const int32_t* edge_seq,
                                      int32_t nseq, const tnqs_bp_opts* bopts, int32_t update_cache, double* truncerr,
                                      int32_t* refreshes) {
  API_BEGIN(c)
  apply_circuit_impl(c, ngates, gate_v, gate_off, gate_data, aopts, edge_seq, nseq, bopts, update_cache, truncerr, refreshes, 0, nullptr, nullptr, nullptr);
  API_END(c)
}
extern "C" int32_t tnqs_apply_circuit_fetch(tnqs_ctx* c, int32_t ngates, const int32_t* gate_v, const int64_t* gate_off,
                                            const double* gate_data, const tnqs_apply_opts* aopts, const int32_t* edge_seq,
                                            int32_t nseq, const tnqs_bp_opts* bopts, int32_t update_cache, double* truncerr,
                                            int32_t* refreshes, int32_t fetch_count, const int32_t* fetch_v, double* const* fetch_data, int64_t* fetch_n) {
  API_BEGIN(c)
  REQUIRE(fetch_count >= 0 && (fetch_count == 0 || (fetch_v && fetch_data && fetch_n)), TNQS_ERR_INVALID_ARG, "tnqs_apply_circuit_fetch: null argument");
  apply_circuit_impl(c, ngates, gate_v, gate_off, gate_data, aopts, edge_seq, nseq, bopts, update_cache, truncerr, refreshes, fetch_count, fetch_v, fetch_data, fetch_n);
  API_END(c)
}

// ---------------------------------------------------------------------------------------------------------------
// measurements
// ---------------------------------------------------------------------------------------------------------------
static Chain full_env_chain(tnqs_ctx* c, int v, int skip_vertex) {
  Chain ch; ch.v = v;
  for (size_t j = 0; j < c->inc[v].size(); ++j) {
    int e = c->inc[v][j], u = c->other(e, v);
    if (u == skip_vertex) continue;
    Absorb ab; ab.leg = (int)j; ab.mat = c->msg(c->dedge(u, e)); ab.ldm = c->chi[e]; ab.flags = 0;
    ch.absorb.push_back(ab);
  }
  return ch;
}

static void rdm_impl(tnqs_ctx* c, int n, const int32_t* vs, cplx* d_out /* n * dmax^2 */) {
  std::vector<std::vector<Chain>> levels(1);
  const int dd = c->dmax * c->dmax;
  cplx* tmp = c->talloc<cplx>((size_t)n * dd);
  for (int i = 0; i < n; ++i) {
    REQUIRE(vs[i] >= 0 && vs[i] < c->nv, TNQS_ERR_INVALID_ARG, "vertex out of range");
    Chain ch = full_env_chain(c, vs[i], -1);
    ch.close_leg = -1; ch.dst = tmp + (size_t)i * dd;
    levels[0].push_back(ch);
  }
  Plan p = build_plan(c, levels, scratch_budget(c));
  run_plan(c, p);
  CU(cudaMemcpyAsync(d_out, tmp, (size_t)n * dd * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream));
}

extern "C" int32_t tnqs_rdm_1site(tnqs_ctx* c, int32_t n, const int32_t* v, double* rho) {
  API_BEGIN(c)
  REQUIRE(n >= 0 && (n == 0 || (v && rho)), TNQS_ERR_INVALID_ARG, "tnqs_rdm_1site: null argument");
  if (n > 0) {
    const int dd = c->dmax * c->dmax;
    cplx* d_out = c->talloc<cplx>((size_t)n * dd);
    CU(cudaMemsetAsync(d_out, 0, (size_t)n * dd * sizeof(cplx), c->stream));
    rdm_impl(c, n, v, d_out);
    // compact d x d blocks were written with ld = d; expand to the caller's dmax x dmax layout on the host
    std::vector<cplx> h((size_t)n * dd);
    CU(cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    cplx* out = (cplx*)rho;
    for (int i = 0; i < n; ++i) {
      int d = c->d[v[i]];
      for (int b = 0; b < c->dmax; ++b) for (int a = 0; a < c->dmax; ++a)
        out[(size_t)i * dd + a + b * c->dmax] = (a < d && b < d) ? h[(size_t)i * dd + a + b * d] : cmake(0, 0);
    }
  }
  c->release_temps();
  API_END(c)
}

extern "C" int32_t tnqs_expect_2site_nn(tnqs_ctx* c, int32_t n, const int32_t* es, const double* ops, double* out) {
  API_BEGIN(c)
  REQUIRE(n >= 0 && (n == 0 || (es && ops && out)), TNQS_ERR_INVALID_ARG, "tnqs_expect_2site_nn: null argument");
  if (n > 0) {
    // level 0: environments of the first vertex with and without O1, closed on the shared bond
    // level 1: second vertex with / without O2, absorbing that environment, closed on the site index
    std::vector<std::vector<Chain>> levels(2);
    size_t opoff = 0; std::vector<cplx> hops;
    std::vector<size_t> o1(n), o2(n);
    for (int i = 0; i < n; ++i) {
      REQUIRE(es[i] >= 0 && es[i] < c->ne, TNQS_ERR_INVALID_ARG, "edge out of range");
      int a = c->esrc[es[i]], b = c->edst[es[i]], da = c->d[a], db = c->d[b];
      const cplx* src = (const cplx*)ops + opoff;
      o1[i] = hops.size(); hops.insert(hops.end(), src, src + da * da);
      o2[i] = hops.size(); hops.insert(hops.end(), src + da * da, src + da * da + db * db);
      opoff += da * da + db * db;
    }
    cplx* d_ops = c->upload(hops);
    const int dd = c->dmax * c->dmax;
    cplx* envs = c->talloc<cplx>((size_t)2 * n * c->slot);
    cplx* res = c->talloc<cplx>((size_t)2 * n * dd);
    for (int i = 0; i < n; ++i) {
      int e = es[i], a = c->esrc[e], b = c->edst[e];
      for (int w = 0; w < 2; ++w) {                       // w = 0: numerator (with operators), 1: denominator
        Chain ca = full_env_chain(c, a, b);
        ca.op = w == 0 ? d_ops + o1[i] : nullptr; ca.close_leg = c->leg(a, e); ca.dst = envs + (size_t)(2 * i + w) * c->slot;
        levels[0].push_back(ca);
        Chain cb = full_env_chain(c, b, a);
        cb.op = w == 0 ? d_ops + o2[i] : nullptr;
        Absorb ab; ab.leg = c->leg(b, e); ab.mat = ca.dst; ab.ldm = c->chi[e]; ab.flags = 0;
        cb.absorb.push_back(ab);
        cb.close_leg = -1; cb.dst = res + (size_t)(2 * i + w) * dd;
        levels[1].push_back(cb);
      }
    }
    Plan p = build_plan(c, levels, scratch_budget(c));
    run_plan(c, p);
    std::vector<cplx> h((size_t)2 * n * dd);
    CU(cudaMemcpyAsync(h.data(), res, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    for (int i = 0; i < n; ++i)
      for (int w = 0; w < 2; ++w) {
        int d = c->d[c->edst[es[i]]];
        double re = 0, im = 0;
        for (int s = 0; s < d; ++s) { cplx x = h[(size_t)(2 * i + w) * dd + s + s * d]; re += x.x; im += x.y; }
        out[4 * i + 2 * w] = re; out[4 * i + 2 * w + 1] = im;
      }
  }
  c->release_temps();
  API_END(c)
}

// Observables with one operator at each end of a path v_0 .. v_L of the network (the Steiner tree of two vertices,
// src/expect.jl:66): vertex i absorbs the messages from every neighbour off the path and the environment handed over by
// v_{i-1}, and closes on the bond towards v_{i+1} (the last vertex closes on the site index).  One level per path position.
extern "C" int32_t tnqs_expect_path(tnqs_ctx* c, int32_t n, const int32_t* path_off, const int32_t* path_v, const double* ops,
                                    double* out) {
  API_BEGIN(c)
  REQUIRE(n >= 0 && (n == 0 || (path_off && path_v && ops && out)), TNQS_ERR_INVALID_ARG, "tnqs_expect_path: null argument");
  if (n > 0) {
    size_t maxlen = 0;
    for (int i = 0; i < n; ++i) {
      int len = path_off[i + 1] - path_off[i];
      REQUIRE(len >= 2, TNQS_ERR_INVALID_ARG, "a path needs at least two vertices");
      maxlen = std::max(maxlen, (size_t)len);
    }
    std::vector<std::vector<Chain>> levels(maxlen);
    size_t opoff = 0; std::vector<cplx> hops; std::vector<size_t> o1(n), o2(n);
    for (int i = 0; i < n; ++i) {
      const int* pv = path_v + path_off[i]; int len = path_off[i + 1] - path_off[i];
      for (int j = 0; j < len; ++j) REQUIRE(pv[j] >= 0 && pv[j] < c->nv, TNQS_ERR_INVALID_ARG, "path vertex out of range");
      int da = c->d[pv[0]], db = c->d[pv[len - 1]];
      const cplx* src = (const cplx*)ops + opoff;
      o1[i] = hops.size(); hops.insert(hops.end(), src, src + da * da);
      o2[i] = hops.size(); hops.insert(hops.end(), src + da * da, src + da * da + db * db);
      opoff += da * da + db * db;
    }
    cplx* d_ops = c->upload(hops);
    const int dd = c->dmax * c->dmax;
    cplx* res = c->talloc<cplx>((size_t)2 * n * dd);
    auto edge_between = [&](int a, int b) { for (int e : c->inc[a]) if (c->other(e, a) == b) return e; return -1; };
    for (int i = 0; i < n; ++i) {
      const int* pv = path_v + path_off[i]; int len = path_off[i + 1] - path_off[i];
      for (int w = 0; w < 2; ++w) {                         // w = 0: numerator (with operators), 1: denominator
        cplx* env_prev = nullptr;
        for (int j = 0; j < len; ++j) {
          int v = pv[j], prev = j > 0 ? pv[j - 1] : -1, next = j + 1 < len ? pv[j + 1] : -1;
          Chain ch; ch.v = v;
          for (size_t l = 0; l < c->inc[v].size(); ++l) {
            int e = c->inc[v][l], u = c->other(e, v);
            if (u == next) continue;
            Absorb ab; ab.leg = (int)l; ab.ldm = c->chi[e]; ab.flags = 0;
            if (u == prev) ab.mat = env_prev; else ab.mat = c->msg(c->dedge(u, e));
            ch.absorb.push_back(ab);
          }
          if (w == 0 && j == 0) ch.op = d_ops + o1[i];
          if (w == 0 && j == len - 1) ch.op = d_ops + o2[i];
          if (next >= 0) {
            int e = edge_between(v, next);
            REQUIRE(e >= 0, TNQS_ERR_INVALID_ARG, "consecutive path vertices do not share an edge");
            ch.close_leg = c->leg(v, e); ch.dst = c->talloc<cplx>((size_t)c->slot);
            env_prev = ch.dst;
          } else { ch.close_leg = -1; ch.dst = res + (size_t)(2 * i + w) * dd; }
          levels[j].push_back(ch);
        }
      }
    }
    Plan p = build_plan(c, levels, scratch_budget(c));
    run_plan(c, p);
    std::vector<cplx> h((size_t)2 * n * dd);
    CU(cudaMemcpyAsync(h.data(), res, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    for (int i = 0; i < n; ++i)
      for (int w = 0; w < 2; ++w) {
        int d = c->d[path_v[path_off[i + 1] - 1]];
        double re = 0, im = 0;
        for (int s = 0; s < d; ++s) { cplx x = h[(size_t)(2 * i + w) * dd + s + s * d]; re += x.x; im += x.y; }
        out[4 * i + 2 * w] = re; out[4 * i + 2 * w + 1] = im;
      }
  }
  c->release_temps();
  API_END(c)
}

#include "loop_corrections.cuh"

extern "C" int32_t tnqs_loop_weights(tnqs_ctx* c, int32_t nsub, const int32_t* sub_off, const int32_t* sub_edges, double* weights) {
  API_BEGIN(c)
  REQUIRE(nsub >= 0 && (nsub == 0 || (sub_off && sub_edges && weights)), TNQS_ERR_INVALID_ARG, "tnqs_loop_weights: null argument");
  for (int i = 0; i < nsub; ++i) REQUIRE(sub_off[i + 1] >= sub_off[i], TNQS_ERR_INVALID_ARG, "tnqs_loop_weights: offsets must not decrease");
  if (nsub > 0) loop_weights_impl(c, nsub, sub_off, sub_edges, weights);
  c->release_temps();
  API_END(c)
}

extern "C" int32_t tnqs_loop_transfer(tnqs_ctx* c, int32_t len, const int32_t* walk_v, double* out) {
  API_BEGIN(c)
  REQUIRE(walk_v && out, TNQS_ERR_INVALID_ARG, "tnqs_loop_transfer: null argument");
  loop_transfer_impl(c, len, walk_v, out);
  c->release_temps();
  API_END(c)
}

static void vertex_scalars_impl(tnqs_ctx* c, cplx* d_vs /* [nv] device */) {
  std::vector<int> vs(c->nv);
  for (int v = 0; v < c->nv; ++v) vs[v] = v;
  const int dd = c->dmax * c->dmax;
  cplx* rho = c->talloc<cplx>((size_t)c->nv * dd);
  rdm_impl(c, c->nv, vs.data(), rho);
  std::vector<cplx> h((size_t)c->nv * dd), t(c->nv);
  CU(cudaMemcpyAsync(h.data(), rho, h.size() * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  for (int v = 0; v < c->nv; ++v) {
    int d = c->d[v]; double re = 0, im = 0;
    for (int s = 0; s < d; ++s) { re += h[(size_t)v * dd + s + s * d].x; im += h[(size_t)v * dd + s + s * d].y; }
    t[v] = cmake(re, im);
  }
  c->h2d(d_vs, t.data(), c->nv * sizeof(cplx));
}

extern "C" int32_t tnqs_vertex_scalars(tnqs_ctx* c, double* out) {
  API_BEGIN(c)
  REQUIRE(out, TNQS_ERR_INVALID_ARG, "null output");
  cplx* d_vs = c->talloc<cplx>(c->nv);
  vertex_scalars_impl(c, d_vs);
  CU(cudaMemcpyAsync(out, d_vs, c->nv * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  c->release_temps();
  API_END(c)
}

extern "C" int32_t tnqs_edge_scalars(tnqs_ctx* c, double* out) {
  API_BEGIN(c)
  REQUIRE(out || c->ne == 0, TNQS_ERR_INVALID_ARG, "null output");
  if (c->ne) {
    cplx* d_es = c->talloc<cplx>(c->ne);
    LAUNCH(c, TNQS_KC_MISC, k_edge_scalar, c->ne, 256, 0, c->msg_cur, c->slot, c->d_chi, d_es);
    CU(cudaMemcpyAsync(out, d_es, c->ne * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
  }
  c->release_temps();
  API_END(c)
}

// With a bra attached (bilinear form) only the ket layer can be scaled: it is divided by the vertex scalar itself, the rule
// rescale_vertices! applies to single-layer networks (beliefpropagationcache.jl:100-101; for a BilinearForm the reference has
// no rule and raises, :102-103), so that every vertex scalar becomes 1 as the loop series needs.
extern "C" int32_t tnqs_rescale(tnqs_ctx* c) {
  API_BEGIN(c)
  if (c->ne) LAUNCH(c, TNQS_KC_MISC, k_rescale_messages, c->ne, 256, 0, c->msg_cur, c->slot, c->d_chi);
  cplx* d_vs = c->talloc<cplx>(c->nv);
  vertex_scalars_impl(c, d_vs);
  std::vector<ScaleDesc> sd(c->nv);
  long long maxn = 1;
  for (int v = 0; v < c->nv; ++v) { sd[v].x = c->site[v]; sd[v].n = c->site_n[v]; sd[v].scalar = (const double*)(d_vs + v); sd[v].mode = c->bra ? 2 : 1; maxn = std::max(maxn, c->site_n[v]); }
  ScaleDesc* dsd = c->upload(sd);
  LAUNCH(c, TNQS_KC_MISC, k_scale, dim3((unsigned)std::min<long long>((maxn + 255) / 256, 1024), c->nv), 256, 0, dsd);
  c->sync();
  c->release_temps();
  API_END(c)
}

// ---------------------------------------------------------------------------------------------------------------
// symmetric gauge, truncation
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t tnqs_symmetric_gauge(tnqs_ctx* c, double regularization) {
  API_BEGIN(c)
  const int ne = c->ne;
  if (ne > 0) {
    // per edge: roots of both messages, C = rootX^T rootY, SVD, leg matrices N_a = invX^T (U sqrtS), N_c = invY^T (conj(V) sqrtS)
    std::vector<EighDesc> eg(2 * ne); std::vector<SmallMatmulDesc> mm(ne), mm2(2 * ne); std::vector<SvdDesc> sv(ne);
    std::vector<cplx*> Nmat(2 * ne);
    int* d_rank = c->talloc<int>(ne); double* d_err = c->talloc<double>(ne); double* d_sv = c->talloc<double>((size_t)ne * c->chi_cap);
    int maxchi = 1; size_t wvmax = 0;
    for (int e = 0; e < ne; ++e) {
      int x = c->chi[e]; maxchi = std::max(maxchi, x);
      for (int s = 0; s < 2; ++s) {
        EighDesc& d = eg[2 * e + s];
        d.m = c->msg(2 * e + s); d.n = x; d.ldm = x; d.root = c->talloc<cplx>((size_t)x * x); d.invroot = c->talloc<cplx>((size_t)x * x);
        d.cutoff = 0; d.shift = regularization; d.mode = 1;
      }
      SmallMatmulDesc& m = mm[e];
      m.a = eg[2 * e].root; m.b = eg[2 * e + 1].root; m.c = c->talloc<cplx>((size_t)x * x); m.m = x; m.k = x; m.n = x; m.lda = x; m.ldb = x; m.ldc = x; m.transA = 1;
      SvdDesc& s = sv[e];
      s.theta = m.c; s.m = x; s.n = x; s.work = nullptr; s.r1out = c->talloc<cplx>((size_t)x * x); s.r2out = c->talloc<cplx>((size_t)x * x);
      s.svals = d_sv + (size_t)e * c->chi_cap; s.rank = d_rank + e; s.truncerr = d_err + e; s.msg_a = nullptr; s.msg_b = nullptr;
      s.maxdim = 0; s.cutoff = 0; s.has_cutoff = 0; s.normalize = 0; s.cap = c->chi_cap; s.mindim = 1;
      wvmax = std::max(wvmax, (size_t)x * x * sizeof(cplx));
      for (int side = 0; side < 2; ++side) {
        SmallMatmulDesc& t = mm2[2 * e + side];
        Nmat[2 * e + side] = c->talloc<cplx>((size_t)x * x);
        t.a = eg[2 * e + side].invroot; t.b = side == 0 ? s.r1out : s.r2out; t.c = Nmat[2 * e + side];
        t.m = x; t.k = x; t.n = x; t.lda = x; t.ldb = x; t.ldc = x; t.transA = 1;
      }
    }
    EighDesc* deg = c->upload(eg);
    int threads = 32 * std::max(1, std::min(32, (maxchi + 1) / 2));
    LAUNCH(c, TNQS_KC_EIGH, k_eigh_sqrt, 2 * ne, threads, (size_t)2 * maxchi * maxchi * sizeof(cplx) + 3 * maxchi * sizeof(double), deg, c->d_status);
    SmallMatmulDesc* dmm = c->upload(mm);
    LAUNCH(c, TNQS_KC_MISC, k_small_matmul, ne, 256, 0, dmm);
    SvdDesc* dsv = c->upload(sv);
    LAUNCH(c, TNQS_KC_SVD, k_svd_trunc, ne, std::min(threads, 512), wvmax + (size_t)maxchi * 12 + 64, dsv, c->d_status, wvmax);
    SmallMatmulDesc* dmm2 = c->upload(mm2);
    LAUNCH(c, TNQS_KC_MISC, k_small_matmul, 2 * ne, 256, 0, dmm2);
    int st = c->read_status();
    if (st & ST_NEG_EIG) throw TnqsError(TNQS_ERR_NEGATIVE_EIGENVALUE, "negative message eigenvalue in symmetric_gauge");
    if (st & ST_NO_CONV) throw TnqsError(TNQS_ERR_SVD_NO_CONVERGENCE, "Jacobi iteration did not converge");
    // apply the leg matrices: one leg per step, all vertices that have such a leg
    size_t maxdeg = 0;
    for (int v = 0; v < c->nv; ++v) maxdeg = std::max(maxdeg, c->inc[v].size());
    std::vector<cplx*> other(c->nv);
    for (int v = 0; v < c->nv; ++v) other[v] = (cplx*)c->dalloc(c->site_n[v] * sizeof(cplx));
    for (size_t j = 0; j < maxdeg; ++j) {
      std::vector<ModeApplyDesc> ma; long long maxtiles = 1;
      std::vector<int> touched;
      for (int v = 0; v < c->nv; ++v) {
        if (j >= c->inc[v].size()) continue;
        int e = c->inc[v][j], x = c->chi[e];
        long long Lp = c->d[v];
        for (size_t i = 0; i < j; ++i) Lp *= c->chi[c->inc[v][i]];
        long long R = c->site_n[v] / (Lp * x);
        ModeApplyDesc d; d.in = c->site[v]; d.out = other[v]; d.m = Nmat[2 * e + (c->esrc[e] == v ? 0 : 1)];
        d.L = (int)Lp; d.n = x; d.np = x; d.R = (int)R; d.in_sb = Lp; d.in_sr = Lp * x; d.out_sb = Lp; d.out_sr = Lp * x; d.ldm = x; d.flags = 0;
        ma.push_back(d); touched.push_back(v);
        maxtiles = std::max(maxtiles, (Lp * R + MA_BM - 1) / MA_BM);
      }
      (void)maxtiles;
      launch_mode_apply(c, TNQS_KC_MODE_APPLY, ma);
      for (int v : touched) std::swap(c->site[v], other[v]);
      ++c->state_version;
    }
    for (int v = 0; v < c->nv; ++v) c->dfree(other[v]);
    std::vector<int> ks(ne);
    for (int e = 0; e < ne; ++e) ks[e] = e;
    int* dks = c->upload(ks);
    LAUNCH(c, TNQS_KC_MISC, k_msg_diag, ne, 128, 0, c->msg_cur, c->slot, dks, d_sv, d_rank, c->chi_cap);
    c->sync();
  }
  c->release_temps();
  API_END(c)
}

extern "C" int32_t tnqs_truncate(tnqs_ctx* c, int32_t ngroups, const int32_t* group_off, const int32_t* group_edges,
                                 const tnqs_apply_opts* aopts, const int32_t* edge_seq, int32_t nseq,
                                 const tnqs_bp_opts* bopts) {
  API_BEGIN(c)
  REQUIRE(ngroups >= 0 && aopts && bopts && (ngroups == 0 || (group_off && group_edges)), TNQS_ERR_INVALID_ARG,
          "tnqs_truncate: null argument");
  for (int g = 0; g < ngroups; ++g) {
    std::vector<int> es; std::vector<cplx> gates;
    for (int i = group_off[g]; i < group_off[g + 1]; ++i) {
      int e = group_edges[i];
      REQUIRE(e >= 0 && e < c->ne, TNQS_ERR_INVALID_ARG, "edge out of range");
      if (c->chi[e] == 1) continue;                                   // truncatable_edge (src/truncate.jl:5-10)
      int d = c->d[c->esrc[e]];
      es.push_back(e);
      for (int j = 0; j < d * d * d * d; ++j) {                      // identity o[s1',s2',s1,s2]
        int s1p = j % d, s2p = (j / d) % d, s1 = (j / (d * d)) % d, s2 = j / (d * d * d);
        gates.push_back(cmake((s1p == s1 && s2p == s2) ? 1.0 : 0.0, 0.0));
      }
    }
    if (!es.empty()) { apply_2site_impl(c, (int)es.size(), es.data(), nullptr, (const double*)gates.data(), *aopts, nullptr, nullptr, nullptr); c->release_temps(); }
    int it; double md;
    bp_update_impl(c, edge_seq, nseq, *bopts, &it, &md);
    c->release_temps();
  }
  c->sync();
  API_END(c)
}

// ---------------------------------------------------------------------------------------------------------------
// multi-GPU halo and instrumentation
// ---------------------------------------------------------------------------------------------------------------
extern "C" int32_t tnqs_message_slab(tnqs_ctx* c, void** dev_ptr, int64_t* n_complex, int64_t* slot_stride) {
  if (!c) return TNQS_ERR_INVALID_ARG;
  if (dev_ptr) *dev_ptr = c->msg_cur;
  if (n_complex) *n_complex = (int64_t)std::max(1, 2 * c->ne) * c->slot;
  if (slot_stride) *slot_stride = c->slot;
  return TNQS_OK;
}
static int32_t halo(tnqs_ctx* c, int32_t n, const int32_t* k, void* buf, int pack) {
  API_BEGIN(c)
  REQUIRE(n >= 0 && (n == 0 || (k && buf)), TNQS_ERR_INVALID_ARG, "halo: null argument");
  if (n > 0) {
    std::vector<int> ks(k, k + n);
    for (int x : ks) REQUIRE(x >= 0 && x < 2 * c->ne, TNQS_ERR_INVALID_ARG, "halo: directed edge out of range");
    int* dks = c->upload(ks);
    LAUNCH(c, TNQS_KC_MISC, k_halo_pack, n, 256, 0, c->msg_cur, c->slot, dks, (cplx*)buf, pack);
    c->sync();
  }
  c->release_temps();
  API_END(c)
}
extern "C" int32_t tnqs_halo_pack(tnqs_ctx* c, int32_t n, const int32_t* k, void* dev_buf) { return halo(c, n, k, dev_buf, 1); }
extern "C" int32_t tnqs_halo_unpack(tnqs_ctx* c, int32_t n, const int32_t* k, const void* dev_buf) { return halo(c, n, k, const_cast<void*>(dev_buf), 0); }

// ---- peer-memory halo -------------------------------------------------------------------------------------------
static void halo_release(tnqs_ctx* c) {
  for (auto& p : c->halo_peers) if (p.block && p.ipc) cudaIpcCloseMemHandle(p.block);
  c->halo_peers.clear(); c->halo_route.clear();
  if (c->halo_block) cudaFree(c->halo_block);
  c->halo_block = nullptr; c->halo_in = nullptr; c->halo_nrecv = 0; c->halo_world = 0;
}
extern "C" int32_t tnqs_halo_alloc(tnqs_ctx* c, int32_t world, int32_t nrecv, const int32_t* recv_k, void* ipc_handle, void** dev_ptr) {
  API_BEGIN(c)
  REQUIRE(c->partitioned, TNQS_ERR_INVALID_ARG, "tnqs_halo_alloc: set the partition first (tnqs_set_partition)");
  REQUIRE(world >= 1 && world <= HALO_MAX_WORLD && c->rank < world && nrecv >= 0 && (nrecv == 0 || recv_k), TNQS_ERR_INVALID_ARG, "tnqs_halo_alloc: bad argument");
  for (int i = 0; i < nrecv; ++i) REQUIRE(recv_k[i] >= 0 && recv_k[i] < 2 * c->ne, TNQS_ERR_INVALID_ARG, "tnqs_halo_alloc: directed edge out of range");
  c->sync();
  bp_cache_drop(c);
  halo_release(c);
  c->halo_world = world; c->halo_nrecv = nrecv; c->halo_recv_k.assign(recv_k, recv_k + nrecv);
  const size_t sig_bytes = HALO_SIG_DOUBLES * sizeof(double);
  const size_t bytes = sig_bytes + std::max<size_t>(1, (size_t)2 * nrecv * c->slot) * sizeof(cplx);
  CU(cudaMalloc((void**)&c->halo_block, bytes));                      // not from the stream-ordered pool: IPC needs a plain allocation
  CU(cudaMemsetAsync(c->halo_block, 0, bytes, c->stream));
  c->halo_in = (cplx*)(c->halo_block + sig_bytes);
  if (!c->d_halo_parity) { c->d_halo_parity = (int*)c->dalloc(2 * sizeof(int)); c->d_halo_out = (double*)c->dalloc(2 * sizeof(double)); }
  const int init[2] = {1, 0};                                         // the first sweep is epoch 1: it writes half 1
  c->h2d(c->d_halo_parity, init, sizeof(init));
  CU(cudaMemsetAsync(c->d_halo_out, 0, 2 * sizeof(double), c->stream));
  if (c->d_halo_recv_k) { c->dfree(c->d_halo_recv_k); c->d_halo_recv_k = nullptr; }
  c->d_halo_recv_k = (int*)c->dalloc(std::max(1, nrecv) * sizeof(int));
  c->h2d(c->d_halo_recv_k, c->halo_recv_k.data(), nrecv * sizeof(int));
  c->halo_peers.assign(world, tnqs_ctx::HaloPeer());
  if (ipc_handle) { cudaIpcMemHandle_t h; CU(cudaIpcGetMemHandle(&h, c->halo_block)); memcpy(ipc_handle, &h, sizeof(h)); }
  if (dev_ptr) *dev_ptr = c->halo_block;
  c->sync();
  API_END(c)
}
extern "C" int32_t tnqs_halo_peer(tnqs_ctx* c, int32_t peer, const void* ipc_handle, void* raw_dev_ptr, int32_t peer_nrecv) {
  API_BEGIN(c)
  REQUIRE(c->halo_block && peer >= 0 && peer < c->halo_world && peer != c->rank && peer_nrecv >= 0 && ((ipc_handle != nullptr) != (raw_dev_ptr != nullptr)),
          TNQS_ERR_INVALID_ARG, "tnqs_halo_peer: after tnqs_halo_alloc; give either an IPC handle (another process) or a device pointer (same process)");
  bp_cache_drop(c);
  tnqs_ctx::HaloPeer& p = c->halo_peers[peer];
  if (p.block && p.ipc) cudaIpcCloseMemHandle(p.block);
  p = tnqs_ctx::HaloPeer();
  if (ipc_handle) {
    cudaIpcMemHandle_t h; memcpy(&h, ipc_handle, sizeof(h));
    void* q = nullptr;
    CU(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
    p.block = (char*)q; p.ipc = true;
  } else {
    p.block = (char*)raw_dev_ptr; p.ipc = false;
  }
  p.base = (cplx*)(p.block + HALO_SIG_DOUBLES * sizeof(double));
  p.nrecv = peer_nrecv;
  ++c->state_version;
  API_END(c)
}
extern "C" int32_t tnqs_halo_route(tnqs_ctx* c, int32_t n, const int32_t* k, const int32_t* peer, const int32_t* peer_slot) {
  API_BEGIN(c)
  REQUIRE(c->halo_block && n >= 0 && (n == 0 || (k && peer && peer_slot)), TNQS_ERR_INVALID_ARG, "tnqs_halo_route: after tnqs_halo_alloc; null argument");
  bp_cache_drop(c);
  c->halo_route.assign(2 * c->ne, tnqs_ctx::HaloRoute());
  for (int i = 0; i < n; ++i) {
    REQUIRE(k[i] >= 0 && k[i] < 2 * c->ne, TNQS_ERR_INVALID_ARG, "tnqs_halo_route: directed edge out of range");
    REQUIRE(peer[i] >= 0 && peer[i] < (int)c->halo_peers.size() && c->halo_peers[peer[i]].block, TNQS_ERR_INVALID_ARG, "tnqs_halo_route: peer block not opened (tnqs_halo_peer)");
    REQUIRE(peer_slot[i] >= 0 && peer_slot[i] < c->halo_peers[peer[i]].nrecv, TNQS_ERR_INVALID_ARG, "tnqs_halo_route: slot outside the peer's buffer");
    c->halo_route[k[i]].peer = peer[i]; c->halo_route[k[i]].slot = peer_slot[i];
  }
  ++c->state_version;
  API_END(c)
}
extern "C" int32_t tnqs_halo_close(tnqs_ctx* c) {
  API_BEGIN(c)
  c->sync();
  bp_cache_drop(c);
  halo_release(c);
  ++c->state_version;
  API_END(c)
}
extern "C" int32_t tnqs_set_bra(tnqs_ctx* c, tnqs_ctx* bra) {
  API_BEGIN(c)
  if (bra) {
    REQUIRE(bra != c && bra->device == c->device && bra->nv == c->nv && bra->ne == c->ne && bra->esrc == c->esrc && bra->edst == c->edst,
            TNQS_ERR_INVALID_ARG, "tnqs_set_bra: the bra must be another handle on the same device with the same graph");
    REQUIRE(bra->d == c->d && bra->chi == c->chi, TNQS_ERR_SHAPE, "tnqs_set_bra: site and bond dimensions of ket and bra must agree");
    CU(cudaStreamSynchronize(bra->stream));
  }
  c->sync();
  bp_cache_drop(c);
  if (c->bra) { auto& us = c->bra->bra_users; us.erase(std::remove(us.begin(), us.end(), c), us.end()); }
  c->bra = bra;
  if (bra) bra->bra_users.push_back(c);
  ++c->state_version;
  API_END(c)
}
extern "C" int32_t tnqs_set_partition(tnqs_ctx* c, const int32_t* owner, int32_t rank) {
  if (!c) return TNQS_ERR_INVALID_ARG;
  if (!owner) { c->partitioned = false; c->owner.clear(); return TNQS_OK; }
  c->owner.assign(owner, owner + c->nv); c->rank = rank; c->partitioned = true;
  return TNQS_OK;
}
extern "C" int64_t tnqs_kernel_launches(const tnqs_ctx* c) { return c ? c->launches : 0; }
extern "C" int64_t tnqs_bp_sweeps_total(const tnqs_ctx* c) { return c ? c->bp_sweeps : 0; }
extern "C" int64_t tnqs_bp_message_updates_total(const tnqs_ctx* c) { return c ? c->bp_msg_updates : 0; }
extern "C" int32_t tnqs_profile_enable(tnqs_ctx* c, int32_t on) {
  if (!c) return TNQS_ERR_INVALID_ARG;
  c->collect_profile(); c->profiling = on != 0; return TNQS_OK;
}
extern "C" int32_t tnqs_profile_read(tnqs_ctx* c, double* ms, int64_t* launches, int32_t reset) {
  if (!c) return TNQS_ERR_INVALID_ARG;
  cudaSetDevice(c->device);
  c->collect_profile();
  for (int i = 0; i < TNQS_KC_COUNT; ++i) { if (ms) ms[i] = c->prof_ms[i]; if (launches) launches[i] = c->prof_launches[i]; }
  if (reset) for (int i = 0; i < TNQS_KC_COUNT; ++i) { c->prof_ms[i] = 0; c->prof_launches[i] = 0; }
  return TNQS_OK;
}
extern "C" void* tnqs_stream(tnqs_ctx* c) { return c ? (void*)c->stream : nullptr; }
