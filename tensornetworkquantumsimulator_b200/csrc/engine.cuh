// Host-side engine: context, device memory, contraction plans (levels of chains), launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <stdexcept>
#include <chrono>

#include "../../include/tnqs_cuda.h"
#include "common.cuh"
#include "kernels_gemm.cuh"
#include "kernels_small.cuh"
#include "kernels_misc.cuh"
#include "kernels_dmma.cuh"
#include "kernels_deep.cuh"

struct TnqsError {
  int code; std::string msg;
  TnqsError(int c, const std::string& m) : code(c), msg(m) {}
};

#define CU(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) throw TnqsError(_e == cudaErrorMemoryAllocation ? TNQS_ERR_OOM : TNQS_ERR_CUDA, \
  std::string(#x) + ": " + cudaGetErrorString(_e)); } while (0)
#define REQUIRE(cond, code, msg) do { if (!(cond)) throw TnqsError(code, msg); } while (0)

struct ProfEvent { cudaEvent_t a, b; int kc; };

struct tnqs_ctx {
  int device = 0;
  int num_sms = 148;
  size_t smem_optin = 227 * 1024;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;              // device -> host copies that overlap the closing BP refresh (tnqs_apply_circuit_fetch)
  int nv = 0, ne = 0, chi_cap = 1, dmax = 2;
  std::vector<int> esrc, edst, d, chi;
  std::vector<std::vector<int>> inc;               // v -> incident edge ids ascending
  std::vector<cplx*> site;                         // device pointers
  std::vector<long long> site_n;                   // element counts
  cplx* msg_cur = nullptr; cplx* msg_stage = nullptr;
  long long slot = 1;
  int* d_chi = nullptr;                            // device copy of chi[ne]
  int* d_status = nullptr;
  std::vector<int> owner; int rank = 0; bool partitioned = false;
  // peer-memory halo (tnqs_halo_*): one cudaMalloc'ed block per handle (exportable over CUDA IPC) = [signal area of the
  // device-side barrier | receive buffer: 2 parity halves of halo_nrecv message slots]; the opened blocks of the other ranks;
  // where each owned -> ghost message goes
  char* halo_block = nullptr; cplx* halo_in = nullptr; int halo_nrecv = 0; std::vector<int> halo_recv_k; int* d_halo_recv_k = nullptr;
  struct HaloPeer { char* block = nullptr; cplx* base = nullptr; int nrecv = 0; bool ipc = false; };
  std::vector<HaloPeer> halo_peers;
  struct HaloRoute { int peer = -1; int slot = 0; };
  std::vector<HaloRoute> halo_route;               // per directed edge (peer < 0: not a boundary message)
  int* d_halo_parity = nullptr;                    // device words: [0] half the next sweep writes to, [1] sweeps completed (epoch)
  double* d_halo_out = nullptr;                    // device: [0] global diff sum, [1] global message count of the last sweep
  int halo_world = 0;
  bool halo_active() const { return partitioned && halo_world > 1 && halo_block && d_halo_parity; }
  std::string err;
  long long launches = 0;
  long long bp_sweeps = 0, bp_msg_updates = 0;
  bool profiling = false;
  std::vector<ProfEvent> prof_events;
  double prof_ms[TNQS_KC_COUNT] = {0};
  long long prof_launches[TNQS_KC_COUNT] = {0};
  std::vector<void*> temp;                         // stream-ordered temporaries of the current call
  void* pending_gate = nullptr;
  size_t budget_bytes = 0; int budget_calls = 0; long long budget_version = -1;      // scratch_budget() cache
  char* gg_pool = nullptr; size_t gg_pool_size = 0, gg_pool_off = 0;               // pinned pool of the captured batches' upload stages
  std::vector<void*> gate_graphs;                  // GateGraph*: captured optimistic gate batches (tnqs_cuda.cu)
  void* bp_cache = nullptr;                        // BpPlanCache: last BP plan, reused while sites / bonds are unchanged
  long long state_version = 0;                     // bumped whenever a site pointer or a bond dimension changes                    // GateBatch between tnqs_apply_2site_begin and _end

  // Bilinear form <phi|psi> (src/Forms/bilinearform.jl:28-54): when `bra` is attached, every BP contraction of this handle closes
  // its ket tensors against the conjugated site tensors of that handle instead of its own (messages are then not Hermitian).
  tnqs_ctx* bra = nullptr;
  std::vector<tnqs_ctx*> bra_users;                // handles whose `bra` points here: detached when this handle is destroyed
  const cplx* bra_site(int v) const { return bra ? bra->site[v] : site[v]; }

  // ---- graph helpers ----
  int other(int e, int v) const { return esrc[e] == v ? edst[e] : esrc[e]; }
  int dsrc(int k) const { return (k & 1) ? edst[k >> 1] : esrc[k >> 1]; }
  int ddst(int k) const { return (k & 1) ? esrc[k >> 1] : edst[k >> 1]; }
  int dedge(int u, int e) const { return 2 * e + (esrc[e] == u ? 0 : 1); }       // u -> other(e, u)
  int leg(int v, int e) const {
    for (size_t i = 0; i < inc[v].size(); ++i) if (inc[v][i] == e) return (int)i;
    return -1;
  }
  long long site_elems(int v) const { long long n = d[v]; for (int e : inc[v]) n *= chi[e]; return n; }
  cplx* msg(int k) const { return msg_cur + (size_t)k * slot; }
  cplx* msg_st(int k) const { return msg_stage + (size_t)k * slot; }

  // ---- memory ----
  void* dalloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TNQS_DEBUG_ALLOC"); dbg = (e && e[0] == '1') ? 1 : 0; }
    auto t0 = std::chrono::steady_clock::now();
    CU(cudaMallocAsync(&p, bytes, stream));
    if (dbg) {
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > 0.5) fprintf(stderr, "[tnqs alloc] %.2f ms for %.1f MB\n", ms, bytes / 1048576.0);
    }
    return p;
  }
  void dfree(void* p) { if (p) cudaFreeAsync(p, stream); }
  // Per-call temporaries come from a bump arena (chunks are kept for the life of the handle): a batch of ~1000 gates
  // makes ~15000 small allocations, far too many for one cudaMallocAsync each.  `arena_bypass` = allocations that
  // outlive the call (the cached BP plan) go to the pool directly and are tracked in `temp`.
  struct ArenaChunk { char* base; size_t size; };
  std::vector<ArenaChunk> arena; size_t arena_chunk = 0, arena_off = 0; bool arena_bypass = false;
  void* arena_alloc(size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    while (arena_chunk < arena.size()) {
      if (arena_off + bytes <= arena[arena_chunk].size) { void* p = arena[arena_chunk].base + arena_off; arena_off += bytes; return p; }
      ++arena_chunk; arena_off = 0;
    }
    if (capturing) throw TnqsError(TNQS_ERR_OOM, "arena would have to grow during a graph capture");
    size_t sz = std::max(bytes, (size_t)256 << 20);
    ArenaChunk ch; ch.base = (char*)dalloc(sz); ch.size = sz;
    arena.push_back(ch); arena_chunk = arena.size() - 1; arena_off = bytes;
    return ch.base;
  }
  template <typename T> T* talloc(size_t count) {
    if (arena_bypass) { T* p = (T*)dalloc(count * sizeof(T)); temp.push_back(p); return p; }
    return (T*)arena_alloc(std::max<size_t>(count, 1) * sizeof(T));
  }
  void release_temps() { for (void* p : temp) dfree(p); temp.clear(); arena_chunk = 0; arena_off = 0; }
  void arena_destroy() { for (auto& ch : arena) dfree(ch.base); arena.clear(); }
  // TNQS_DEBUG_STALL=<ms>: report host-side API calls (uploads, launches) that block longer than the threshold
  static double stall_ms() {
    static double t = -2;
    if (t < -1) { const char* e = getenv("TNQS_DEBUG_STALL"); t = e ? atof(e) : -1; }
    return t;
  }
  struct StallProbe {
    const char* what; std::chrono::steady_clock::time_point t0; bool on;
    explicit StallProbe(const char* w) : what(w), on(stall_ms() >= 0) { if (on) t0 = std::chrono::steady_clock::now(); }
    ~StallProbe() {
      if (!on) return;
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > stall_ms()) fprintf(stderr, "[tnqs stall] %s blocked the host for %.2f ms\n", what, ms);
    }
  };
  // Descriptor uploads go through a pinned staging buffer owned by the handle: a copy from pageable memory makes the
  // driver stage (or page-lock) the source on every call, which is slow and erratic on a host that holds gigabytes of
  // pinned memory (measured: +35..70 ms per layer).  The buffer is a bump allocator that rewinds at every stream
  // synchronisation; a request that does not fit falls back to the pageable copy.
  char* hstage = nullptr; size_t hstage_size = 0, hstage_off = 0, hstage_want = 0; bool hstage_disabled = false;
  void hstage_resize(size_t bytes) {
    if (hstage) { cudaFreeHost(hstage); hstage = nullptr; }
    hstage_size = 0; hstage_off = 0;
    if (cudaHostAlloc((void**)&hstage, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); hstage = nullptr; hstage_disabled = true; return; }
    hstage_size = bytes;
  }
  // While a gate batch is being captured into a CUDA graph (GateGraph in tnqs_cuda.cu) every upload is staged in a pinned buffer
  // that belongs to that graph: its memcpy nodes read the same bytes at every replay (the shared staging buffer is recycled at
  // each synchronisation).  cap_last = where the latest upload was staged (the gate matrices are refreshed there before a replay).
  char* cap_stage = nullptr; size_t cap_stage_size = 0, cap_stage_off = 0; char* cap_last = nullptr; bool capturing = false;
  void h2d(void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return;
    if (capturing) {
      const size_t need = (bytes + 255) / 256 * 256;
      if (!cap_stage || cap_stage_off + need > cap_stage_size) throw TnqsError(TNQS_ERR_OOM, "capture staging buffer exhausted");
      char* s = cap_stage + cap_stage_off; cap_stage_off += need; cap_last = s;
      memcpy(s, src, bytes);
      CU(cudaMemcpyAsync(dst, s, bytes, cudaMemcpyHostToDevice, stream));
      return;
    }
    StallProbe sp("upload");
    if (!hstage && !hstage_disabled) {
      const char* e = getenv("TNQS_NO_HSTAGE");
      if (e && e[0] == '1') hstage_disabled = true; else hstage_resize((size_t)1 << 20);        // grows with the largest call (up to 64 MB)
    }
    const size_t need = (bytes + 255) / 256 * 256;
    if (hstage && hstage_off + need <= hstage_size) {
      char* s = hstage + hstage_off; hstage_off += need;
      memcpy(s, src, bytes);
      CU(cudaMemcpyAsync(dst, s, bytes, cudaMemcpyHostToDevice, stream));
    } else {
      hstage_want = std::min<size_t>((size_t)64 << 20, std::max(hstage_want, 2 * (hstage_off + need)));   // pageable this time, a bigger buffer after the next sync
      CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    }
  }
  void hstage_destroy() { if (hstage) cudaFreeHost(hstage); hstage = nullptr; hstage_size = 0; hstage_off = 0; if (hret) cudaFreeHost(hret); hret = nullptr;
                          if (hside) cudaFreeHost(hside); hside = nullptr; hside_size = 0;
                          if (gg_pool) cudaFreeHost(gg_pool); gg_pool = nullptr; gg_pool_size = 0; gg_pool_off = 0; }
  // Second pinned return page, grown on demand: results that come back on the copy stream (gate ranks) or in one batch at the end
  // of a call (truncation errors, singular values).  Valid until the next request.
  char* hside = nullptr; size_t hside_size = 0;
  void* hret_side(size_t bytes) {
    if (bytes > hside_size) {
      if (hside) cudaFreeHost(hside);
      hside = nullptr; hside_size = 0;
      const size_t want = std::max<size_t>(bytes, (size_t)64 << 10);
      if (cudaHostAlloc((void**)&hside, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); hside = nullptr; throw TnqsError(TNQS_ERR_OOM, "pinned host page for device results"); }
      hside_size = want;
    }
    return hside;
  }
  // Small device -> host reads (loop state, status word, counters) land in a pinned page and the call then waits on the
  // stream: a copy into pageable memory blocks inside the driver, which serialises every other thread of the process - fatal
  // when those threads drive other handles whose kernels this handle's kernels wait for (device-side halo barrier).
  char* hret = nullptr;
  void d2h_small(void* dst, const void* dev_src, size_t bytes) {
    constexpr size_t HRET = 256 << 10;
    if (!hret && cudaHostAlloc((void**)&hret, HRET, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); hret = nullptr; }
    if (hret && bytes <= HRET) {
      CU(cudaMemcpyAsync(hret, dev_src, bytes, cudaMemcpyDeviceToHost, stream));
      sync();
      memcpy(dst, hret, bytes);
    } else {
      CU(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, stream));
      sync();
    }
  }
  template <typename T> T* upload(const std::vector<T>& v) {
    T* p = talloc<T>(std::max<size_t>(v.size(), 1));
    h2d(p, v.data(), v.size() * sizeof(T));
    return p;
  }
  void sync() { CU(cudaStreamSynchronize(stream)); hstage_off = 0; if (hstage_want > hstage_size && !hstage_disabled) hstage_resize(hstage_want); }

  // ---- launch accounting ----
  void pre_launch(int kc) {
    ++launches; ++prof_launches[kc];
    if (profiling) {
      ProfEvent ev; ev.kc = kc;
      cudaEventCreate(&ev.a); cudaEventCreate(&ev.b);
      cudaEventRecord(ev.a, stream);
      prof_events.push_back(ev);
    }
  }
  void post_launch() {
    if (profiling) cudaEventRecord(prof_events.back().b, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw TnqsError(TNQS_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
  }
  void collect_profile() {
    if (prof_events.empty()) return;
    cudaStreamSynchronize(stream);
    for (auto& ev : prof_events) {
      float ms = 0; cudaEventElapsedTime(&ms, ev.a, ev.b);
      prof_ms[ev.kc] += ms;
      cudaEventDestroy(ev.a); cudaEventDestroy(ev.b);
    }
    prof_events.clear();
  }
  int read_status() {
    int st = 0;
    d2h_small(&st, d_status, sizeof(int));
    if (st) { int z = 0; h2d(d_status, &z, sizeof(int)); }
    return st;
  }
};

#define LAUNCH(ctx, kc, kernel, grid, block, smem, ...) do { tnqs_ctx::StallProbe _sp("launch " #kernel); (ctx)->pre_launch(kc); \
  kernel<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__); (ctx)->post_launch(); } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Contraction plans.  A chain takes a site tensor, optionally applies a one-site operator, absorbs small matrices
// on a list of legs, and closes against the conjugated site on one leg (or on the site index) into a small
// matrix.  A level is a set of independent chains; levels run in order; BP levels commit staged messages.
// ---------------------------------------------------------------------------------------------------------------
struct Absorb { int leg; const cplx* mat; int ldm; int flags; };     // leg: 0-based bond leg of the site
struct Chain {
  int v = -1;
  const cplx* op = nullptr;          // device d x d operator o[s', s] applied first, or null
  std::vector<Absorb> absorb;
  int close_leg = -1;                // bond leg to close on; -1 = close on the site index (RDM)
  // output
  cplx* dst = nullptr;               // n x n result (BP: staged message)
  cplx* dst2 = nullptr; long long dst2_pstride = 0; const int* parity = nullptr;     // peer-memory halo copy (CloseReduceDesc)
  const cplx* prev = nullptr; double* diff = nullptr; int bp_mode = 0, normalize = 0, hermitize = 0;
  int commit_k = -1;                 // BP: directed edge to commit at the end of the level
  int split_cap = 16;                // upper bound of the split-K factor of the closing (a property of the request, never of the batch)
  int split_rows = 0;                // smallest K extent of one split (0: the default, split_rows())
};

struct LaunchRec { size_t off; int count; dim3 grid; size_t smem; bool dmma = false; int tpc = 1; };

static bool dmma_disabled() {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TNQS_NO_DMMA"); disabled = (e && e[0] == '1') ? 1 : 0; }
  return disabled == 1;
}

// Kernel choice and grid of one batched mode_apply launch: the tensor-core kernel when every item has n, np <= 32.
struct ApplyLaunch { bool dmma; dim3 grid; size_t smem; int tpc; };      // tpc: row tiles per CTA of k_leg_dmma
static ApplyLaunch plan_apply_launch(const ModeApplyDesc* d, size_t count) {
  bool ok = !dmma_disabled(); long long maxrows = 1, maxtiles = 1; int maxnp = 1; size_t smem = 0;
  for (size_t i = 0; i < count; ++i) {
    ok = ok && d[i].n <= 64 && d[i].np <= 64;
    const long long rows = (long long)d[i].L * d[i].R;
    maxrows = std::max(maxrows, rows); maxnp = std::max(maxnp, d[i].np);
    if (ok) { const int rpc = leg_rows_per_cta(d[i].np); maxtiles = std::max(maxtiles, (rows + rpc - 1) / rpc); smem = std::max(smem, leg_smem_bytes(d[i].n, d[i].np)); }
  }
  if (ok) {
    // several row tiles per CTA once the launch is many waves deep (2 CTAs x 148 SMs per wave): the matrix fill is amortised
    long long tpc = std::max<long long>(1, std::min<long long>(8, maxtiles * (long long)count / (8 * 296)));
    tpc = std::min(tpc, maxtiles);
    return {true, dim3((unsigned)((maxtiles + tpc - 1) / tpc), (unsigned)count, 1), smem, (int)tpc};
  }
  return {false, dim3((unsigned)((maxrows + MA_BM - 1) / MA_BM), (unsigned)count, (unsigned)((maxnp + MA_BN - 1) / MA_BN)), 0, 1};
}
struct SubLevel {
  LaunchRec slab{0, 0, dim3(), 0}, dclose{0, 0, dim3(), 0};   // DMMA fast path (kernels_dmma.cuh)
  LaunchRec fused{0, 0, dim3(), 0};                            // the same chains as one persistent launch (k_bp_level): queue items
  LaunchRec deep{0, 0, dim3(), 0};                             // degree-5 / degree-6 chains at N = 8 (kernels_deep.cuh): k_deep_close, one item per message
  LaunchRec deepA{0, 0, dim3(), 0};                            // k_deep_absorb: one item per message, or two per degree-6 vertex whose six updates share them
  int fastN = 0, close_cap = 16;                               // close_cap: split cap of the fast chains (selects the closing CTA size)
  LaunchRec gate1{0, 0, dim3(), 0};
  std::vector<LaunchRec> steps;
  LaunchRec close{0, 0, dim3(), 0}, reduce{0, 0, dim3(), 0};
};
struct LevelPlan { std::vector<SubLevel> subs; size_t commit_off = 0; int commit_count = 0; };

struct Plan {
  std::vector<LevelPlan> levels;
  Gate1Desc* d_gate1 = nullptr; ModeApplyDesc* d_apply = nullptr; ModeCloseDesc* d_close = nullptr;
  CloseReduceDesc* d_reduce = nullptr; int* d_commit = nullptr;
  SlabDesc* d_slab = nullptr; CloseDmmaDesc* d_dclose = nullptr; int* d_ticket = nullptr;
  DeepDesc* d_deep = nullptr; DeepDesc* d_deepA = nullptr;
  int2* d_items = nullptr; int* d_qctr = nullptr; int* d_slab_done = nullptr;       // k_bp_level / k_bp_sweep queue, its counters, per-chain slab counters
  // whole-sweep persistent launch (k_bp_sweep): every chain of every level in one data-flow queue
  bool sweep = false; int sweepN = 0, sweep_items = 0; dim3 sweep_grid; size_t sweep_smem = 0;
  SweepChain* d_chains = nullptr; int* d_chain_done = nullptr; size_t nchains = 0;
  long long* d_prof = nullptr;       // TNQS_SWEEP_TRACE=1: 32 cycle counters of the sweep kernel (SweepTrace)
  int sweep_commit = 0;              // staged (JACOBI) sweeps end with a commit pass over this many messages
  // storage: a rebuilt plan of the same shape reuses the buffers of its predecessor, so device addresses (and a CUDA graph
  // captured over them) stay valid when only the site pointers inside the descriptors changed
  cplx* scratch = nullptr; size_t scratch_elems = 0; cplx* partial = nullptr; size_t partial_elems = 0;
  size_t cnt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // descriptor counts: gate1, apply, close, reduce, commit, slab, dclose, queue items, deep, deepA
  int nlaunch = 0;                   // kernel launches per run
  double flops = 0;                  // executed flops per run of the plan (8 per complex MAC)
};

// smallest K extent (rows of the matricised tensor) of one split of a closing contraction
static long long split_rows() { const char* e = getenv("TNQS_SPLIT_ROWS"); return e ? std::max(64, atoi(e)) : 512; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// DMMA fast path eligibility: BP-type chain (no operator, closes on a bond leg), site dimension 2, every leg of the
// vertex of the same dimension N in {8, 16}, 2 or 3 absorbed legs with plain N x N matrices.  Returns N or 0.
static int fast_path_N(const tnqs_ctx* c, const Chain& ch) {
  if (dmma_disabled() || ch.op || ch.close_leg < 0) return 0;
  const int v = ch.v, z = (int)c->inc[v].size(), F = (int)ch.absorb.size();
  if (c->d[v] != 2 || F != z - 1 || F < 2 || F > 3) return 0;
  const int N = c->chi[c->inc[v][0]];
  if (N != 8 && N != 16) return 0;
  for (int e : c->inc[v]) if (c->chi[e] != N) return 0;
  int expect_leg = 0;
  for (const Absorb& ab : ch.absorb) {
    if (expect_leg == ch.close_leg) ++expect_leg;
    if (ab.leg != expect_leg || ab.flags != 0 || ab.ldm != N) return 0;
    ++expect_leg;
  }
  return N;
}

// Deep path eligibility (kernels_deep.cuh): BP-type chain on a vertex of degree 5 or 6 with site dimension 2 and every leg of
// dimension 8, plain 8 x 8 matrices on all other legs in leg order.  Returns the extent of the fifth absorbed leg (8, or 1 when there
// are four) or 0.  TNQS_DEEP=0 sends these chains down the generic leg-by-leg path.
static bool deep_disabled() {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TNQS_DEEP"); disabled = (e && e[0] == '0') ? 1 : 0; }
  return disabled == 1;
}
static int deep_path_E4(const tnqs_ctx* c, const Chain& ch) {
  if (dmma_disabled() || deep_disabled() || ch.op || ch.close_leg < 0) return 0;
  const int v = ch.v, z = (int)c->inc[v].size(), F = (int)ch.absorb.size();
  if (c->d[v] != 2 || F != z - 1 || F < 4 || F > 5) return 0;
  for (int e : c->inc[v]) if (c->chi[e] != 8) return 0;
  int expect_leg = 0;
  for (const Absorb& ab : ch.absorb) {
    if (expect_leg == ch.close_leg) ++expect_leg;
    if (ab.leg != expect_leg || ab.flags != 0 || ab.ldm != 8) return 0;
    ++expect_leg;
  }
  return F == 5 ? 8 : 1;
}

// Shared absorptions (generic leg-by-leg chains): the z message updates that leave one vertex in the same level absorb the same
// messages into the same site tensor, each leaving out a different leg.  Degree 4: T x m0 x m1 and T x m2 x m3 are formed once and
// serve two updates each - 8 leg passes instead of 12; degree 3: T x m0 and T x m1 serve three - 5 instead of 6 (the association
// order changes, results agree to rounding; oracle twin: jacobi_sweep(shared=True)).  Eligible: a BP-type chain without operator whose
// absorb list is every other leg in ascending order with plain matrices.  TNQS_BP_SHARE=0 switches it off.
static bool share_disabled() {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TNQS_BP_SHARE"); disabled = (e && e[0] == '0') ? 1 : 0; }
  return disabled == 1;
}
static bool share_candidate(const tnqs_ctx* c, const Chain& ch) {
  if (share_disabled() || ch.op || ch.close_leg < 0) return false;
  const int v = ch.v, z = (int)c->inc[v].size(), F = (int)ch.absorb.size();
  if (F != z - 1 || (z != 3 && z != 4)) return false;
  int expect_leg = 0;
  for (const Absorb& ab : ch.absorb) {
    if (expect_leg == ch.close_leg) ++expect_leg;
    if (ab.leg != expect_leg || ab.flags != 0 || ab.ldm != c->chi[c->inc[v][ab.leg]]) return false;
    ++expect_leg;
  }
  return true;
}

// Build a plan for `levels` (vector of vector of chains).  All device buffers are temporaries of the current call.
template <typename T> static T* upload_into(tnqs_ctx* c, T* reuse, const std::vector<T>& v) {
  if (!reuse) return c->upload(v);
  c->h2d(reuse, v.data(), v.size() * sizeof(T));
  return reuse;
}

// The persistent one-launch-per-level kernel (k_bp_level) is opt-in (TNQS_FUSE=1): measured on C2 it does not beat the
// two-launch level (its closing items wait on slabs too often); TNQS_FUSE_LAG sets the queue lag in messages.
static bool fuse_disabled() {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("TNQS_FUSE"); disabled = (e && e[0] == '1') ? 0 : 1; }
  return disabled == 1;
}
static int fuse_lag() {
  static int lag = 0;
  if (!lag) { const char* e = getenv("TNQS_FUSE_LAG"); lag = e ? std::max(1, atoi(e)) : 12; }
  return lag;
}

// strict_levels: no chain of a level reads a message that another chain of the same level writes (the caller built the
// levels that way), so epilogues may write their message in place even while other chains of the level are in flight.
static Plan build_plan(tnqs_ctx* c, const std::vector<std::vector<Chain>>& levels, size_t scratch_budget_bytes, const Plan* reuse = nullptr,
                       bool strict_levels = false) {
  Plan plan;
  std::vector<int2> items;
  int sm_count = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); }
  std::vector<Gate1Desc> g1; std::vector<ModeApplyDesc> ma; std::vector<ModeCloseDesc> mc;
  std::vector<CloseReduceDesc> rd; std::vector<int> commit;
  std::vector<SlabDesc> sd; std::vector<CloseDmmaDesc> dc; std::vector<DeepDesc> dd, da;
  std::vector<CloseReduceDesc> frd;            // epilogues of the fast-path chains, in the order of `dc`
  // pass 1: scratch requirement per chain, sub-level split, global scratch sizing
  struct ChainMem { size_t buf_elems; int nbuf; int nsplit; size_t partial_elems; };
  size_t max_scratch = 0, max_partial = 0;
  std::vector<std::vector<std::vector<int>>> subsets(levels.size());   // level -> sub-level -> chain indices
  std::vector<std::vector<ChainMem>> mem(levels.size());
  for (size_t L = 0; L < levels.size(); ++L) {
    const auto& chains = levels[L];
    mem[L].resize(chains.size());
    std::vector<int> cur; size_t cur_bytes = 0;
    // a sub-level launches one slab / closing kernel pair, i.e. one fast-path bond dimension: chains are taken grouped by it
    // (generic chains first; they ride along with any group), and a change of N opens a new sub-level.  Inhomogeneous states
    // (a region still at 8 next to one at 16 under a cutoff-driven truncation) therefore cost one more launch pair, not an error.
    std::vector<int> order(chains.size()), fN(chains.size());
    for (size_t i = 0; i < chains.size(); ++i) { order[i] = (int)i; fN[i] = fast_path_N(c, chains[i]); }
    // chains are taken vertex by vertex: the updates leaving one vertex land in one sub-level (shared absorptions of the generic
    // chains) and next to each other in the batched launches (they read the same site tensor: L2).  TNQS_BP_VSORT=0: generic only.
    static int vsort = -1;
    if (vsort < 0) { const char* e = getenv("TNQS_BP_VSORT"); vsort = (e && e[0] == '0') ? 0 : 1; }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (fN[a] != fN[b]) return fN[a] < fN[b];
      return (fN[a] == 0 || vsort) && chains[a].v < chains[b].v;
    });
    int curN = 0;
    for (int i : order) {
      const Chain& ch = chains[i];
      int nsteps = (ch.op ? 1 : 0) + (int)ch.absorb.size();
      ChainMem cm; cm.buf_elems = (size_t)c->site_n[ch.v]; cm.nbuf = std::min(nsteps, 2); cm.nsplit = 1; cm.partial_elems = 0;
      if (fN[i] || deep_path_E4(c, ch)) cm.nbuf = 1;
      mem[L][i] = cm;
      size_t bytes = align_up(cm.buf_elems * sizeof(cplx), 256) * cm.nbuf;
      const bool otherN = fN[i] && curN && fN[i] != curN;
      // a sub-level is also cut at 65535 chains: the item index of every batched launch is blockIdx.y
      if (!cur.empty() && (cur_bytes + bytes > scratch_budget_bytes || otherN || cur.size() >= 65535)) { subsets[L].push_back(cur); cur.clear(); cur_bytes = 0; curN = 0; }
      cur.push_back(i); cur_bytes += bytes;
      if (fN[i]) curN = fN[i];
      max_scratch = std::max(max_scratch, cur_bytes);
    }
    if (!cur.empty()) subsets[L].push_back(cur);
    for (auto& sub : subsets[L]) {
      // split-K factor depends on the tensor and the request only (never on the batch), so a message is computed bit-identically whatever
      // else shares its launch: partitioned (multi-GPU) JACOBI sweeps reproduce the single-GPU result exactly.
      size_t part = 0;
      for (int i : sub) {
        const Chain& ch = chains[i];
        int n = ch.close_leg < 0 ? c->d[ch.v] : c->chi[c->inc[ch.v][ch.close_leg]];
        long long rows = c->site_n[ch.v] / n;
        int ns = (int)std::min<long long>(ch.split_cap, std::max<long long>(1, rows / (ch.split_rows ? ch.split_rows : split_rows())));
        mem[L][i].nsplit = std::max(1, ns);
        if (deep_path_E4(c, ch)) mem[L][i].nsplit = 2 * 8 * 8;          // one partial per CTA of k_deep_close
        mem[L][i].partial_elems = (size_t)mem[L][i].nsplit * n * n;
        part += mem[L][i].partial_elems;
      }
      max_partial = std::max(max_partial, part);
    }
  }
  const size_t need_scratch = max_scratch / sizeof(cplx) + 16, need_partial = max_partial + 16;
  const bool reuse_mem = reuse && reuse->scratch && reuse->scratch_elems >= need_scratch && reuse->partial_elems >= need_partial;
  cplx* scratch = reuse_mem ? reuse->scratch : c->talloc<cplx>(need_scratch);
  cplx* partial = reuse_mem ? reuse->partial : c->talloc<cplx>(need_partial);
  plan.scratch = scratch; plan.scratch_elems = reuse_mem ? reuse->scratch_elems : need_scratch;
  plan.partial = partial; plan.partial_elems = reuse_mem ? reuse->partial_elems : need_partial;
  // pass 2: descriptors
  for (size_t L = 0; L < levels.size(); ++L) {
    const auto& chains = levels[L];
    LevelPlan lp;
    lp.commit_off = commit.size();
    // one persistent launch per level for the fast-path chains; without the strict-level guarantee its epilogues stage
    // their messages and the level ends with a commit pass
    const bool fuse = !fuse_disabled() && subsets[L].size() == 1;
    bool level_staged = subsets[L].size() > 1;
    if (fuse && !strict_levels) for (const Chain& ch : chains) if (fast_path_N(c, ch) && ch.commit_k >= 0) level_staged = true;
    if (!strict_levels) {
      // the epilogues of the fast-path chains run before the generic chains of the level read their messages: without the
      // strict-level guarantee (JACOBI) a level that mixes both kinds must stage
      bool has_fast = false, has_generic = false;
      for (const Chain& ch : chains) if (ch.commit_k >= 0) { if (fast_path_N(c, ch)) has_fast = true; else has_generic = true; }
      if (has_fast && has_generic) level_staged = true;
    }
    for (auto& sub : subsets[L]) {
      SubLevel sl;
      size_t soff = 0, poff = 0;
      int maxsteps = 0;
      std::vector<cplx*> buf0(sub.size()), buf1(sub.size());
      for (size_t j = 0; j < sub.size(); ++j) {
        const ChainMem& cm = mem[L][sub[j]];
        size_t stride = align_up(cm.buf_elems * sizeof(cplx), 256) / sizeof(cplx);
        buf0[j] = scratch + soff; if (cm.nbuf >= 1) soff += stride;
        buf1[j] = scratch + soff; if (cm.nbuf >= 2) soff += stride;
        maxsteps = std::max(maxsteps, (int)chains[sub[j]].absorb.size());
      }
      // current tensor of every chain and which buffer it would write next
      std::vector<const cplx*> curp(sub.size()); std::vector<int> next(sub.size(), 0);
      for (size_t j = 0; j < sub.size(); ++j) curp[j] = c->site[chains[sub[j]].v];
      // DMMA fast path: fused absorption into the slab-major intermediate, then the tensor-core closing
      std::vector<char> fast(sub.size(), 0);
      std::vector<size_t> fast_poff(sub.size(), 0);
      sl.slab.off = sd.size(); sl.dclose.off = dc.size();
      {
        size_t pscan = 0; int maxsplit_f = 1;
        for (size_t j = 0; j < sub.size(); ++j) {
          const Chain& ch = chains[sub[j]];
          const ChainMem& cm = mem[L][sub[j]];
          int N = fast_path_N(c, ch);
          fast_poff[j] = pscan; pscan += cm.partial_elems;
          if (!N) continue;
          REQUIRE(sl.fastN == 0 || sl.fastN == N, TNQS_ERR_UNSUPPORTED, "mixed fast-path bond dimensions in one level");
          sl.fastN = N; fast[j] = 1; sl.close_cap = ch.split_cap;
          const int v = ch.v, F = (int)ch.absorb.size();
          SlabDesc a; a.t = c->site[v]; a.x = buf0[j]; a.F = F;
          CloseDmmaDesc b; b.x = buf0[j]; b.t = c->bra_site(v); b.partial = partial + fast_poff[j]; b.F = F; b.nsplit = cm.nsplit;
          b.epilogue = nullptr; b.ticket = nullptr;
          long long stride = 2;
          int ai = 0;
          for (int l = 0; l < (int)c->inc[v].size(); ++l) {
            if (l == ch.close_leg) { a.gw = stride; b.gw = stride; }
            else { a.g[ai] = stride; b.g[ai] = stride; a.m[ai] = ch.absorb[ai].mat; ++ai; }
            stride *= N;
          }
          for (; ai < 3; ++ai) { a.g[ai] = 0; b.g[ai] = 0; a.m[ai] = nullptr; }
          a.gs = 1; a.nslabs = 2 * N; a.flags = 0;
          sd.push_back(a); dc.push_back(b);
          maxsplit_f = std::max(maxsplit_f, cm.nsplit);
          plan.flops += 8.0 * (double)c->site_n[v] * N * (F + 1);
        }
        sl.slab.count = (int)(sd.size() - sl.slab.off);
        sl.dclose.count = sl.slab.count;
        if (sl.slab.count) {
          int N = sl.fastN;
          sl.slab.grid = dim3(2 * N, sl.slab.count, 1);
          sl.slab.smem = N == 16 ? SlabCfg<16>::SMEM : SlabCfg<8>::SMEM;
          sl.dclose.grid = dim3(maxsplit_f, sl.dclose.count, 1);
          sl.dclose.smem = (size_t)8 * 2 * N * N * sizeof(double);
          if (fuse) {
            // queue: absorb items of chain m, then the closing items of chain m - D (their slabs were drawn >= D chains ago)
            const int nf = sl.slab.count, D = std::min(nf, fuse_lag());
            sl.fused.off = items.size();
            auto push_close = [&](int m) { for (int j = 0; j < dc[sl.dclose.off + m].nsplit; ++j) items.push_back(make_int2(~(int)(sl.dclose.off + m), j)); };
            for (int m = 0; m < nf; ++m) {
              for (int x = 0; x < sd[sl.slab.off + m].nslabs; ++x) items.push_back(make_int2((int)(sl.slab.off + m), x));
              if (m >= D) push_close(m - D);
            }
            for (int m = std::max(0, nf - D); m < nf; ++m) push_close(m);
            sl.fused.count = (int)(items.size() - sl.fused.off);
            sl.fused.grid = dim3((unsigned)std::min(sl.fused.count, QUEUE_CTAS * sm_count), 1, 1);
            sl.fused.smem = std::max(sl.slab.smem, sl.dclose.smem);
          }
        }
      }
      // deep path: degree-5 / degree-6 chains at N = 8.  The six updates leaving a degree-6 vertex in one sub-level share their
      // three-leg products: T x m0 x m1 x m2 serves the updates on legs 3, 4, 5 and T x m3 x m4 x m5 those on legs 0, 1, 2 - two
      // k_deep_absorb items per vertex instead of six; the tile keeps the three open legs in ascending order and every update reads
      // it with its own strides (DeepDesc::xs).
      sl.deep.off = dd.size(); sl.deepA.off = da.size();
      {
        auto leg_stride = [](int l) { long long st = 2; for (int i = 0; i < l; ++i) st *= 8; return st; };
        size_t j0 = 0;
        while (j0 < sub.size()) {
          const int v = chains[sub[j0]].v;
          size_t j1 = j0;
          while (j1 < sub.size() && chains[sub[j1]].v == v) ++j1;
          const int z = (int)c->inc[v].size();
          // whole group eligible for sharing?
          bool group = !share_disabled() && (z == 6 || z == 5) && (int)(j1 - j0) == z;
          int by_leg[6] = {-1, -1, -1, -1, -1, -1}; const cplx* mat[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
          for (size_t j = j0; group && j < j1; ++j) {
            const Chain& ch = chains[sub[j]];
            group = deep_path_E4(c, ch) == (z == 6 ? 8 : 1) && by_leg[ch.close_leg] < 0;
            if (!group) break;
            by_leg[ch.close_leg] = (int)j;
            for (const Absorb& ab : ch.absorb) { if (mat[ab.leg] && mat[ab.leg] != ab.mat) group = false; mat[ab.leg] = ab.mat; }
          }
          // the per-message item pair of one chain (its own three-leg product)
          auto single = [&](size_t j) {
            const Chain& ch = chains[sub[j]];
            const int E4 = deep_path_E4(c, ch);
            if (!E4) return;
            fast[j] = 2;
            const int F = (int)ch.absorb.size();
            DeepDesc a; a.t = c->site[v]; a.bt = c->bra_site(v); a.x = buf0[j]; a.partial = partial + fast_poff[j]; a.E4 = E4;
            long long stride = 2; int ai = 0;
            for (int l = 0; l < z; ++l) {
              if (l == ch.close_leg) a.gw = stride;
              else { a.g[ai] = stride; a.m[ai] = ch.absorb[ai].mat; ++ai; }
              stride *= 8;
            }
            for (; ai < 5; ++ai) { a.g[ai] = 0; a.m[ai] = nullptr; }
            a.xs[0] = 8; a.xs[1] = 64; a.xs[2] = 64 * E4;
            dd.push_back(a); da.push_back(a);
            plan.flops += 8.0 * (double)c->site_n[v] * 8 * (F + 1);
          };
          if (group && z == 5) {
            // degree 5: T x m0 x m1 x m2 serves the updates on legs 3 and 4, T x m2 x m3 x m4 those on legs 0 and 1 (tile
            // [open leg 2][open leg 1][p0']); the update on leg 2 keeps its own product - three k_deep_absorb items instead of five
            for (int half = 0; half < 2; ++half) {
              const int s0 = half ? 2 : 0, u0 = half ? 0 : 3;
              DeepDesc a; a.t = c->site[v]; a.bt = c->bra_site(v); a.x = buf0[by_leg[u0]]; a.partial = nullptr; a.E4 = 1;
              for (int i = 0; i < 3; ++i) { a.g[i] = leg_stride(s0 + i); a.m[i] = mat[s0 + i]; }
              a.g[3] = leg_stride(u0); a.g[4] = 0; a.gw = leg_stride(u0 + 1);
              a.m[3] = a.m[4] = nullptr; a.xs[0] = 8; a.xs[1] = 64; a.xs[2] = 64;
              da.push_back(a);
              for (int o = 0; o < 2; ++o) {                                 // the update that leaves on open leg u0 + o absorbs the other one
                const int j = by_leg[u0 + o], other = u0 + 1 - o;
                DeepDesc b = a; b.partial = partial + fast_poff[j];
                b.g[3] = leg_stride(other); b.gw = leg_stride(u0 + o); b.m[3] = mat[other];
                b.xs[0] = o == 0 ? 64 : 8; b.xs[1] = 64; b.xs[2] = o == 0 ? 8 : 64;
                dd.push_back(b); fast[j] = 2;
                plan.flops += 8.0 * (double)c->site_n[v] * 8 * 5;
              }
            }
            single((size_t)by_leg[2]);
          } else if (group) {
            for (int half = 0; half < 2; ++half) {
              const int s0 = half ? 3 : 0, u0 = half ? 0 : 3;            // slab legs s0 .. s0 + 2, open legs u0 .. u0 + 2
              DeepDesc a; a.t = c->site[v]; a.bt = c->bra_site(v); a.x = buf0[by_leg[u0]]; a.partial = nullptr; a.E4 = 8;
              for (int i = 0; i < 3; ++i) { a.g[i] = leg_stride(s0 + i); a.m[i] = mat[s0 + i]; }
              a.g[3] = leg_stride(u0); a.g[4] = leg_stride(u0 + 1); a.gw = leg_stride(u0 + 2);
              a.m[3] = a.m[4] = nullptr; a.xs[0] = 8; a.xs[1] = 64; a.xs[2] = 512;
              da.push_back(a);
              for (int o = 0; o < 3; ++o) {                                 // the update that leaves on open leg u0 + o
                const int j = by_leg[u0 + o];
                const int q3 = o == 0 ? 1 : 0, q4 = o == 2 ? 1 : 2;         // the other two open legs, ascending
                DeepDesc b = a; b.partial = partial + fast_poff[j];
                b.g[3] = leg_stride(u0 + q3); b.g[4] = leg_stride(u0 + q4); b.gw = leg_stride(u0 + o);
                b.m[3] = mat[u0 + q3]; b.m[4] = mat[u0 + q4];
                const int tile_stride[3] = {8, 64, 512};                    // open legs in ascending order behind p0'
                b.xs[0] = tile_stride[q3]; b.xs[1] = tile_stride[q4]; b.xs[2] = tile_stride[o];
                dd.push_back(b); fast[j] = 2;
                plan.flops += 8.0 * (double)c->site_n[v] * 8 * 6;
              }
            }
            // (the items need not follow the chain order of the sub-level: every one carries the partial buffer of its chain)
          } else {
            for (size_t j = j0; j < j1; ++j) single(j);
          }
          j0 = j1;
        }
      }
      sl.deep.count = (int)(dd.size() - sl.deep.off);
      sl.deep.grid = dim3(2 * 8 * 8, std::max(1, sl.deep.count), 1);
      sl.deepA.count = (int)(da.size() - sl.deepA.off);
      sl.deepA.grid = dim3(2 * 8 * 8, std::max(1, sl.deepA.count), 1);
      // one-site operators
      sl.gate1.off = g1.size();
      long long maxrest = 0;
      for (size_t j = 0; j < sub.size(); ++j) {
        const Chain& ch = chains[sub[j]];
        if (!ch.op || fast[j]) continue;
        Gate1Desc g; g.in = curp[j]; g.out = buf0[j]; g.gate = ch.op; g.d = c->d[ch.v];
        g.rest = c->site_n[ch.v] / g.d; g.norm2_partial = nullptr;
        g1.push_back(g); curp[j] = buf0[j]; next[j] = 1; maxrest = std::max(maxrest, g.rest);
        plan.flops += 8.0 * g.d * g.d * g.rest;
      }
      sl.gate1.count = (int)(g1.size() - sl.gate1.off);
      sl.gate1.grid = dim3((unsigned)std::min<long long>(std::max<long long>(1, (maxrest + 255) / 256), 1024), sl.gate1.count, 1);
      // shared absorptions: groups of generic chains that are all the updates leaving one vertex
      std::vector<char> shared(sub.size(), 0);
      {
        auto make_apply = [&](int v, int leg, const cplx* in, cplx* out, const cplx* mat) {
          ModeApplyDesc d;
          long long Lp = c->d[v];
          for (int i = 0; i < leg; ++i) Lp *= c->chi[c->inc[v][i]];
          const int n = c->chi[c->inc[v][leg]];
          const long long R = c->site_n[v] / (Lp * n);
          d.in = in; d.out = out; d.m = mat;
          d.L = (int)Lp; d.n = n; d.np = n; d.R = (int)R;
          d.in_sb = Lp; d.in_sr = Lp * n; d.out_sb = Lp; d.out_sr = Lp * n; d.ldm = n; d.flags = 0;
          plan.flops += 8.0 * (double)c->site_n[v] * n;
          return d;
        };
        std::vector<ModeApplyDesc> round[3];
        size_t j0 = 0;
        while (j0 < sub.size()) {
          const int v = chains[sub[j0]].v;
          size_t j1 = j0;
          while (j1 < sub.size() && chains[sub[j1]].v == v) ++j1;
          const int z = (int)c->inc[v].size();
          bool ok = (int)(j1 - j0) == z && (z == 3 || z == 4);
          std::vector<int> by_leg(z > 0 ? z : 1, -1);           // chain (index into sub) that closes on leg l
          std::vector<const cplx*> mat(z > 0 ? z : 1, nullptr);
          for (size_t j = j0; ok && j < j1; ++j) {
            const Chain& ch = chains[sub[j]];
            ok = !fast[j] && share_candidate(c, ch) && mem[L][sub[j]].nbuf == 2 && by_leg[ch.close_leg] < 0;
            if (!ok) break;
            by_leg[ch.close_leg] = (int)j;
            for (const Absorb& ab : ch.absorb) {
              if (mat[ab.leg] && mat[ab.leg] != ab.mat) ok = false;
              mat[ab.leg] = ab.mat;
            }
          }
          if (ok) {
            const cplx* T = c->site[v];
            auto X = [&](int l) { return buf0[by_leg[l]]; };
            if (z == 4) {
              cplx* A01 = buf1[by_leg[0]]; cplx* A23 = buf1[by_leg[1]];
              round[0].push_back(make_apply(v, 0, T, A01, mat[0])); round[0].push_back(make_apply(v, 2, T, A23, mat[2]));
              round[1].push_back(make_apply(v, 1, A01, A01, mat[1])); round[1].push_back(make_apply(v, 3, A23, A23, mat[3]));
              round[2].push_back(make_apply(v, 2, A01, X(3), mat[2])); round[2].push_back(make_apply(v, 3, A01, X(2), mat[3]));
              round[2].push_back(make_apply(v, 0, A23, X(1), mat[0])); round[2].push_back(make_apply(v, 1, A23, X(0), mat[1]));
            } else {
              cplx* P = buf1[by_leg[0]]; cplx* Q = buf1[by_leg[1]];
              round[0].push_back(make_apply(v, 0, T, P, mat[0])); round[0].push_back(make_apply(v, 1, T, Q, mat[1]));
              round[1].push_back(make_apply(v, 1, P, X(2), mat[1])); round[1].push_back(make_apply(v, 2, P, X(1), mat[2]));
              round[1].push_back(make_apply(v, 2, Q, X(0), mat[2]));
            }
            for (size_t j = j0; j < j1; ++j) { shared[j] = 1; curp[j] = buf0[j]; }
          }
          j0 = j1;
        }
        for (int r = 0; r < 3; ++r) {
          if (round[r].empty()) continue;
          LaunchRec lr; lr.off = ma.size(); lr.smem = 0;
          ma.insert(ma.end(), round[r].begin(), round[r].end());
          lr.count = (int)round[r].size();
          ApplyLaunch al = plan_apply_launch(ma.data() + lr.off, lr.count);
          lr.grid = al.grid; lr.dmma = al.dmma; lr.smem = al.smem; lr.tpc = al.tpc;
          sl.steps.push_back(lr);
        }
      }
      // absorption steps
      for (int s = 0; s < maxsteps; ++s) {
        LaunchRec lr; lr.off = ma.size(); lr.smem = 0;
        long long maxtiles = 1; int maxcol = 1;
        for (size_t j = 0; j < sub.size(); ++j) {
          const Chain& ch = chains[sub[j]];
          if (fast[j] || shared[j] || s >= (int)ch.absorb.size()) continue;
          const Absorb& ab = ch.absorb[s];
          int v = ch.v;
          ModeApplyDesc d;
          long long Lp = c->d[v];
          for (int i = 0; i < ab.leg; ++i) Lp *= c->chi[c->inc[v][i]];
          int n = c->chi[c->inc[v][ab.leg]];
          long long R = c->site_n[v] / (Lp * n);
          d.in = curp[j]; d.out = next[j] == 0 ? buf0[j] : buf1[j]; d.m = ab.mat;
          d.L = (int)Lp; d.n = n; d.np = n; d.R = (int)R;
          d.in_sb = Lp; d.in_sr = Lp * n; d.out_sb = Lp; d.out_sr = Lp * n; d.ldm = ab.ldm; d.flags = ab.flags;
          ma.push_back(d);
          curp[j] = d.out; next[j] ^= 1;
          maxtiles = std::max(maxtiles, (Lp * R + MA_BM - 1) / MA_BM);
          maxcol = std::max(maxcol, (n + MA_BN - 1) / MA_BN);
          plan.flops += 8.0 * (double)c->site_n[v] * n;
        }
        lr.count = (int)(ma.size() - lr.off);
        if (lr.count) {
          ApplyLaunch al = plan_apply_launch(ma.data() + lr.off, lr.count);
          lr.grid = al.grid; lr.dmma = al.dmma; lr.smem = al.smem; lr.tpc = al.tpc;
          sl.steps.push_back(lr);
        }
        (void)maxtiles; (void)maxcol;
      }
      // closing + reduction
      sl.close.off = mc.size(); sl.reduce.off = rd.size();
      int maxsplit = 1; size_t maxnn = 1;
      for (size_t j = 0; j < sub.size(); ++j) {
        const Chain& ch = chains[sub[j]];
        const ChainMem& cm = mem[L][sub[j]];
        int v = ch.v;
        long long Lp; int n;
        if (ch.close_leg < 0) { Lp = 1; n = c->d[v]; }
        else { Lp = c->d[v]; for (int i = 0; i < ch.close_leg; ++i) Lp *= c->chi[c->inc[v][i]]; n = c->chi[c->inc[v][ch.close_leg]]; }
        long long R = c->site_n[v] / (Lp * n);
        cplx* part = partial + fast_poff[j];
        if (!fast[j]) {
          ModeCloseDesc d;
          d.a = curp[j]; d.b = c->bra_site(v); d.partial = part;
          d.L = (int)Lp; d.R = (int)R; d.n = n; d.np = n;
          d.a_sb = Lp; d.a_sr = Lp * n; d.b_sb = Lp; d.b_sr = Lp * n;
          d.nsplit = cm.nsplit; d.rows_per_split = (Lp * R + cm.nsplit - 1) / cm.nsplit;
          mc.push_back(d);
          maxsplit = std::max(maxsplit, cm.nsplit);
          plan.flops += 8.0 * (double)c->site_n[v] * n;
        }
        CloseReduceDesc r;
        r.partial = part; r.nsplit = cm.nsplit; r.n = n; r.np = n; r.dst = ch.dst; r.prev = ch.prev; r.diff = ch.diff;
        // an un-chunked level has finished all its reads when its epilogue runs: write the message in place, no commit pass
        if (ch.commit_k >= 0 && !level_staged) r.dst = c->msg(ch.commit_k);
        r.mode = ch.bp_mode; r.normalize = ch.normalize; r.hermitize = ch.hermitize;
        r.dst2 = ch.dst2; r.dst2_pstride = ch.dst2_pstride; r.parity = ch.parity;
        if (fast[j] == 1) frd.push_back(r); else rd.push_back(r);
        maxnn = std::max(maxnn, (size_t)n * n);
        REQUIRE(n <= 64, TNQS_ERR_UNSUPPORTED, "closing extent > 64 is not supported");
      }
      sl.close.count = (int)(mc.size() - sl.close.off); sl.close.grid = dim3(maxsplit, std::max(1, sl.close.count), 1);
      sl.close.dmma = !dmma_disabled();
      for (size_t i = sl.close.off; i < mc.size(); ++i) sl.close.dmma = sl.close.dmma && mc[i].n <= 64 && mc[i].np <= 64;
      sl.reduce.count = (int)(rd.size() - sl.reduce.off); sl.reduce.grid = dim3(std::max(1, sl.reduce.count), 1, 1); sl.reduce.smem = maxnn * sizeof(cplx);
      lp.subs.push_back(sl);
    }
    if (level_staged) for (const Chain& ch : chains) if (ch.commit_k >= 0) commit.push_back(ch.commit_k);
    lp.commit_count = (int)(commit.size() - lp.commit_off);
    plan.levels.push_back(lp);
  }
  // fast-path epilogues live behind the generic reduce descriptors in one array; tickets: one int per fast chain
  const size_t nrd_generic = rd.size();
  rd.insert(rd.end(), frd.begin(), frd.end());
  const size_t cnt[10] = {g1.size(), ma.size(), mc.size(), rd.size(), commit.size(), sd.size(), dc.size(), items.size(), dd.size(), da.size()};
  bool same = reuse_mem;
  for (int i = 0; i < 10 && same; ++i) same = cnt[i] == reuse->cnt[i];
  for (int i = 0; i < 10; ++i) plan.cnt[i] = cnt[i];
  plan.d_deep = upload_into(c, same ? reuse->d_deep : nullptr, dd);
  plan.d_deepA = upload_into(c, same ? reuse->d_deepA : nullptr, da);
  plan.d_gate1 = upload_into(c, same ? reuse->d_gate1 : nullptr, g1); plan.d_apply = upload_into(c, same ? reuse->d_apply : nullptr, ma);
  plan.d_close = upload_into(c, same ? reuse->d_close : nullptr, mc); plan.d_reduce = upload_into(c, same ? reuse->d_reduce : nullptr, rd);
  plan.d_commit = upload_into(c, same ? reuse->d_commit : nullptr, commit);
  plan.d_ticket = (same && reuse->d_ticket) ? reuse->d_ticket : c->talloc<int>(dc.size() + 1);
  if (!(same && reuse->d_ticket)) CU(cudaMemsetAsync(plan.d_ticket, 0, (dc.size() + 1) * sizeof(int), c->stream));
  for (size_t i = 0; i < dc.size(); ++i) { dc[i].epilogue = plan.d_reduce + nrd_generic + i; dc[i].ticket = plan.d_ticket + i; }
  plan.d_slab = upload_into(c, same ? reuse->d_slab : nullptr, sd); plan.d_dclose = upload_into(c, same ? reuse->d_dclose : nullptr, dc);
  plan.d_items = upload_into(c, same ? reuse->d_items : nullptr, items);
  const bool keep_ctr = same && reuse->d_qctr && reuse->d_slab_done;
  plan.d_qctr = keep_ctr ? reuse->d_qctr : c->talloc<int>(4);
  plan.d_slab_done = keep_ctr ? reuse->d_slab_done : c->talloc<int>(sd.size() + 1);
  if (!keep_ctr) {
    CU(cudaMemsetAsync(plan.d_qctr, 0, 4 * sizeof(int), c->stream));
    CU(cudaMemsetAsync(plan.d_slab_done, 0, (sd.size() + 1) * sizeof(int), c->stream));
  }
  for (const LevelPlan& lp : plan.levels) {
    for (const SubLevel& sl : lp.subs) plan.nlaunch += (sl.slab.count ? (sl.fused.count ? 1 : 2) : 0) + (sl.deep.count ? 2 : 0) + (sl.gate1.count ? 1 : 0) + (int)sl.steps.size() + (sl.close.count ? 1 : 0) + (sl.reduce.count ? 1 : 0);
    plan.nlaunch += lp.commit_count ? 1 : 0;
  }
  return plan;
}

// ---------------------------------------------------------------------------------------------------------------
// Whole-sweep plan (k_bp_sweep).  Eligible: every chain is a BP message update on the DMMA fast path (uniform bond
// dimension N in {8, 16}, degree 3 or 4) or a degree-2 vertex with the same N.  TNQS_SWEEP=0 falls back to level launches.
// ---------------------------------------------------------------------------------------------------------------
static bool sweep_disabled() { const char* e = getenv("TNQS_SWEEP"); return e && e[0] == '0'; }
static int sweep_lag() { const char* e = getenv("TNQS_SWEEP_LAG"); return e ? std::max(1, atoi(e)) : 8; }
static int sweep_ring() { const char* e = getenv("TNQS_SWEEP_RING"); return e ? std::max(2, atoi(e)) : 48; }

static int small_chain_N(const tnqs_ctx* c, const Chain& ch) {
  if (dmma_disabled() || ch.op || ch.close_leg < 0) return 0;
  const int v = ch.v;
  if (c->d[v] != 2 || c->inc[v].size() != 2 || ch.absorb.size() != 1) return 0;
  const int N = c->chi[c->inc[v][0]];
  if ((N != 8 && N != 16) || c->chi[c->inc[v][1]] != N) return 0;
  const Absorb& ab = ch.absorb[0];
  if (ab.leg != 1 - ch.close_leg || ab.flags != 0 || ab.ldm != N) return 0;
  return N;
}

// Returns false (plan untouched) when the request is not eligible.  in_place: REFERENCE_ORDER (messages are updated in
// place, readers and writers ordered by per-chain flags); otherwise every message goes to the staged slab and the sweep ends
// with a commit pass (JACOBI).
static bool build_sweep_plan(tnqs_ctx* c, const std::vector<std::vector<Chain>>& levels, bool in_place, const Plan* reuse, Plan& out) {
  if (sweep_disabled() || dmma_disabled()) return false;
  // JACOBI sweeps are not faster through the queue than as two whole-sweep launches (measured on C2: 11.4 vs 9.9 ms), so the
  // staged variant is opt-in (TNQS_SWEEP_JACOBI=1)
  if (!in_place) { const char* e = getenv("TNQS_SWEEP_JACOBI"); if (!(e && e[0] == '1')) return false; }
  // queue order: level-major; inside a level, chains whose producers sit early in the previous levels come first, so the
  // distance (in chains) between a producer and its consumers stays about one level and closings are rarely pulled forward
  std::vector<const Chain*> chains;
  {
    std::vector<int> writer_pos(2 * c->ne, -1);
    for (const auto& lv : levels) {
      std::vector<std::pair<int, const Chain*>> keyed;
      for (const Chain& ch : lv) {
        int key = -1;
        if (in_place) for (const Absorb& ab : ch.absorb) {
          const long long kin = (ab.mat - c->msg_cur) / c->slot;
          if (kin >= 0 && kin < 2 * c->ne) key = std::max(key, writer_pos[kin]);
        }
        keyed.push_back({key, &ch});
      }
      std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<int, const Chain*>& a, const std::pair<int, const Chain*>& b) { return a.first < b.first; });
      for (auto& kc : keyed) { if (kc.second->commit_k >= 0 && kc.second->commit_k < 2 * c->ne) writer_pos[kc.second->commit_k] = (int)chains.size(); chains.push_back(kc.second); }
    }
  }
  const int nch = (int)chains.size();
  if (nch < 2) return false;
  int N = 0; size_t max_elems = 0; int max_split = 1;
  std::vector<char> small(nch, 0);
  for (int i = 0; i < nch; ++i) {
    const Chain& ch = *chains[i];
    if (ch.bp_mode != 1 || ch.commit_k < 0) return false;
    int n = fast_path_N(c, ch);
    if (!n) { n = small_chain_N(c, ch); small[i] = 1; }
    if (!n || (N && n != N)) return false;
    N = n;
    max_elems = std::max(max_elems, (size_t)c->site_n[ch.v]);
  }
  const int R = std::max(2, std::min(nch, sweep_ring())), D = sweep_lag();
  // dependencies between chains, in queue (= level-major sequence) order
  std::vector<SweepChain> sc(nch);
  {
    std::vector<int> last_writer(2 * c->ne, -1);
    std::vector<std::vector<int>> readers(2 * c->ne);
    auto add = [](int* list, int x) -> bool {
      for (int i = 0; i < 4; ++i) { if (list[i] == x) return true; if (list[i] < 0) { list[i] = x; return true; } }
      return false;
    };
    for (int i = 0; i < nch; ++i) {
      const Chain& ch = *chains[i];
      for (int j = 0; j < 4; ++j) { sc[i].dep[j] = -1; sc[i].war[j] = -1; }
      if (i - R >= 0 && !add(sc[i].dep, i - R)) return false;             // previous user of the ring slot (X and partials)
      if (in_place) {
        for (const Absorb& ab : ch.absorb) {
          const int kin = (int)((ab.mat - c->msg_cur) / c->slot);
          if (kin < 0 || kin >= 2 * c->ne || c->msg(kin) != ab.mat) return false;
          if (last_writer[kin] >= 0 && !add(sc[i].dep, last_writer[kin])) return false;
        }
        const int k = ch.commit_k;
        for (int r : readers[k]) if (!add(sc[i].war, r)) return false;
        if (last_writer[k] >= 0 && !add(sc[i].war, last_writer[k])) return false;
        readers[k].clear(); last_writer[k] = i;
        for (const Absorb& ab : ch.absorb) readers[(int)((ab.mat - c->msg_cur) / c->slot)].push_back(i);
      }
    }
  }
  Plan plan; plan.sweep = true; plan.sweepN = N; plan.nchains = (size_t)nch;
  std::vector<SlabDesc> sd(nch); std::vector<CloseDmmaDesc> dc(nch); std::vector<CloseReduceDesc> rd(nch);
  std::vector<int> commit;
  std::vector<int> nsplit(nch, 1);
  for (int i = 0; i < nch; ++i) {
    const Chain& ch = *chains[i];
    if (!small[i]) {
      const long long rows = c->site_n[ch.v] / N;
      nsplit[i] = (int)std::max<long long>(1, std::min<long long>(ch.split_cap, rows / (ch.split_rows ? ch.split_rows : split_rows())));
    }
    max_split = std::max(max_split, nsplit[i]);
  }
  const size_t xslot = align_up(max_elems * sizeof(cplx), 256) / sizeof(cplx), pslot = (size_t)max_split * N * N;
  const size_t need_scratch = xslot * R + 16, need_partial = pslot * R + 16;
  const bool reuse_mem = reuse && reuse->scratch && reuse->scratch_elems >= need_scratch && reuse->partial_elems >= need_partial;
  plan.scratch = reuse_mem ? reuse->scratch : c->talloc<cplx>(need_scratch);
  plan.partial = reuse_mem ? reuse->partial : c->talloc<cplx>(need_partial);
  plan.scratch_elems = reuse_mem ? reuse->scratch_elems : need_scratch;
  plan.partial_elems = reuse_mem ? reuse->partial_elems : need_partial;
  for (int i = 0; i < nch; ++i) {
    const Chain& ch = *chains[i];
    const int v = ch.v, F = (int)ch.absorb.size();
    SlabDesc& a = sd[i]; CloseDmmaDesc& b = dc[i]; CloseReduceDesc& r = rd[i];
    a.t = c->site[v]; a.x = plan.scratch + (size_t)(i % R) * xslot; a.F = F;
    b.x = a.x; b.t = c->bra_site(v); b.partial = plan.partial + (size_t)(i % R) * pslot; b.F = F; b.nsplit = nsplit[i];
    b.epilogue = nullptr; b.ticket = nullptr;
    long long stride = 2; int ai = 0;
    for (int l = 0; l < (int)c->inc[v].size(); ++l) {
      if (l == ch.close_leg) { a.gw = stride; b.gw = stride; }
      else { a.g[ai] = stride; b.g[ai] = stride; a.m[ai] = ch.absorb[ai].mat; ++ai; }
      stride *= N;
    }
    for (; ai < 3; ++ai) { a.g[ai] = 0; b.g[ai] = 0; a.m[ai] = nullptr; }
    a.gs = 1; a.nslabs = small[i] ? 0 : 2 * N; a.flags = 0;
    r.partial = b.partial; r.nsplit = nsplit[i]; r.n = N; r.np = N;
    r.dst = in_place ? c->msg(ch.commit_k) : ch.dst; r.prev = ch.prev; r.diff = ch.diff;
    r.mode = ch.bp_mode; r.normalize = ch.normalize; r.hermitize = ch.hermitize;
    r.dst2 = ch.dst2; r.dst2_pstride = ch.dst2_pstride; r.parity = ch.parity;
    if (!in_place) commit.push_back(ch.commit_k);
    plan.flops += 8.0 * (double)c->site_n[v] * N * (F + 1);
  }
  // queue: absorb items of chain i, closing items of chain i - D; a chain's closing is pulled forward when a later chain
  // depends on it, so no item ever waits on a later item
  std::vector<int2> items;
  {
    std::vector<int> pending; size_t head = 0;
    std::vector<char> closed(nch, 0);
    auto emit_close = [&](int m) { for (int j = 0; j < nsplit[m]; ++j) items.push_back(make_int2(~m, j)); closed[m] = 1; };
    for (int i = 0; i < nch; ++i) {
      int upto = -1;
      for (int j = 0; j < 4; ++j) { const int p = sc[i].dep[j]; if (p >= 0 && !closed[p]) upto = std::max(upto, p); }
      while (upto >= 0 && head < pending.size() && pending[head] <= upto) emit_close(pending[head++]);
      if (small[i]) { items.push_back(make_int2(i, -1)); closed[i] = 1; }
      else {
        for (int y = 0; y < 2 * N; ++y) items.push_back(make_int2(i, y));
        pending.push_back(i);
        while (pending.size() - head > (size_t)D) emit_close(pending[head++]);
      }
    }
    while (head < pending.size()) emit_close(pending[head++]);
  }
  plan.sweep_items = (int)items.size();
  plan.sweep_grid = dim3((unsigned)std::min<long long>((long long)items.size(), (long long)QUEUE_CTAS * c->num_sms), 1, 1);
  plan.sweep_smem = N == 16 ? SlabCfg<16>::SMEM : SlabCfg<8>::SMEM;
  if (const char* e = getenv("TNQS_SWEEP_PAD_SMEM")) plan.sweep_smem += (size_t)atoi(e);      // experiment: shrink the L1 share of the SM
  plan.sweep_commit = (int)commit.size();
  if (getenv("TNQS_DEBUG_QUEUE")) {
    int occ = 0;
    if (N == 16) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bp_sweep<16>, QUEUE_THREADS, plan.sweep_smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bp_sweep<8>, QUEUE_THREADS, plan.sweep_smem);
    fprintf(stderr, "[tnqs] sweep plan: %d chains, %d items, grid %u, %zu B shared memory, %d CTAs per SM\n", nch, plan.sweep_items,
            plan.sweep_grid.x, plan.sweep_smem, occ);
  }
  const size_t cnt[8] = {0, 0, 0, rd.size(), commit.size(), sd.size(), dc.size(), items.size()};
  bool same = reuse_mem && reuse->sweep && reuse->nchains == (size_t)nch && reuse->d_chains && reuse->d_chain_done;
  for (int i = 0; i < 8 && same; ++i) same = cnt[i] == reuse->cnt[i];
  for (int i = 0; i < 8; ++i) plan.cnt[i] = cnt[i];
  plan.d_reduce = upload_into(c, same ? reuse->d_reduce : nullptr, rd);
  plan.d_commit = upload_into(c, same ? reuse->d_commit : nullptr, commit);
  plan.d_ticket = same ? reuse->d_ticket : c->talloc<int>(nch + 1);
  if (!same) CU(cudaMemsetAsync(plan.d_ticket, 0, (nch + 1) * sizeof(int), c->stream));
  for (int i = 0; i < nch; ++i) { dc[i].epilogue = plan.d_reduce + i; dc[i].ticket = plan.d_ticket + i; }
  plan.d_slab = upload_into(c, same ? reuse->d_slab : nullptr, sd);
  plan.d_dclose = upload_into(c, same ? reuse->d_dclose : nullptr, dc);
  plan.d_items = upload_into(c, same ? reuse->d_items : nullptr, items);
  plan.d_chains = upload_into(c, same ? reuse->d_chains : nullptr, sc);
  plan.d_qctr = same ? reuse->d_qctr : c->talloc<int>(8);
  plan.d_slab_done = same ? reuse->d_slab_done : c->talloc<int>(nch + 1);
  plan.d_chain_done = same ? reuse->d_chain_done : c->talloc<int>(nch + 1);
  if (!same) {
    CU(cudaMemsetAsync(plan.d_qctr, 0, 8 * sizeof(int), c->stream));
    CU(cudaMemsetAsync(plan.d_slab_done, 0, (nch + 1) * sizeof(int), c->stream));
    CU(cudaMemsetAsync(plan.d_chain_done, 0, (nch + 1) * sizeof(int), c->stream));
  }
  if (const char* e = getenv("TNQS_SWEEP_TRACE")) if (e[0] == '1') {
    plan.d_prof = (same && reuse->d_prof) ? reuse->d_prof : c->talloc<long long>(32);
    if (!(same && reuse->d_prof)) CU(cudaMemsetAsync(plan.d_prof, 0, 32 * sizeof(long long), c->stream));
  }
  plan.nlaunch = 1 + (plan.sweep_commit ? 1 : 0);
  out = plan;
  return true;
}

static bool same_rec(const LaunchRec& a, const LaunchRec& b) {
  return a.off == b.off && a.count == b.count && a.grid.x == b.grid.x && a.grid.y == b.grid.y && a.grid.z == b.grid.z && a.smem == b.smem && a.dmma == b.dmma && a.tpc == b.tpc;
}
// True when two plans issue the same launches over the same device buffers (a captured CUDA graph of one replays the other).
static bool same_launches(const Plan& a, const Plan& b) {
  if (a.sweep != b.sweep) return false;
  if (a.sweep) return a.sweepN == b.sweepN && a.sweep_items == b.sweep_items && a.sweep_grid.x == b.sweep_grid.x && a.sweep_smem == b.sweep_smem &&
                      a.sweep_commit == b.sweep_commit && a.d_items == b.d_items && a.d_qctr == b.d_qctr && a.d_slab == b.d_slab && a.d_dclose == b.d_dclose &&
                      a.d_chains == b.d_chains && a.d_slab_done == b.d_slab_done && a.d_chain_done == b.d_chain_done && a.d_commit == b.d_commit;
  if (a.levels.size() != b.levels.size() || a.d_slab != b.d_slab || a.d_items != b.d_items || a.d_qctr != b.d_qctr || a.d_slab_done != b.d_slab_done || a.d_apply != b.d_apply || a.d_close != b.d_close ||
      a.d_reduce != b.d_reduce || a.d_dclose != b.d_dclose || a.d_gate1 != b.d_gate1 || a.d_commit != b.d_commit || a.d_deep != b.d_deep || a.d_deepA != b.d_deepA) return false;
  for (size_t L = 0; L < a.levels.size(); ++L) {
    const LevelPlan &x = a.levels[L], &y = b.levels[L];
    if (x.subs.size() != y.subs.size() || x.commit_off != y.commit_off || x.commit_count != y.commit_count) return false;
    for (size_t i = 0; i < x.subs.size(); ++i) {
      const SubLevel &p = x.subs[i], &q = y.subs[i];
      if (p.fastN != q.fastN || p.close_cap != q.close_cap || p.steps.size() != q.steps.size() || !same_rec(p.slab, q.slab) || !same_rec(p.dclose, q.dclose) || !same_rec(p.fused, q.fused) || !same_rec(p.deep, q.deep) || !same_rec(p.deepA, q.deepA) ||
          !same_rec(p.gate1, q.gate1) || !same_rec(p.close, q.close) || !same_rec(p.reduce, q.reduce)) return false;
      for (size_t j = 0; j < p.steps.size(); ++j) if (!same_rec(p.steps[j], q.steps[j])) return false;
    }
  }
  return true;
}

// Upload and launch one batched mode_apply (tensor-core kernel when eligible).
static void launch_mode_apply(tnqs_ctx* c, int kc, const std::vector<ModeApplyDesc>& ma) {
  if (ma.empty()) return;
  ModeApplyDesc* dma = c->upload(ma);
  ApplyLaunch al = plan_apply_launch(ma.data(), ma.size());
  if (al.dmma) LAUNCH(c, kc, k_leg_dmma, al.grid, 256, al.smem, dma, al.tpc);
  else LAUNCH(c, kc, k_mode_apply, al.grid, 256, 0, dma);
}

// CTA size of the closing launch.  Level launches (requests with a split cap above 4: REFERENCE_ORDER) use 4-warp CTAs, so
// that 4 fit an SM and a 32-message level is one wave; whole-sweep launches (JACOBI) use 8 warps.  Like the split factor
// it is a property of the request, never of the batch: a partitioned sweep sums in the same order as the full one.
static int close_threads(int split_cap) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("TNQS_CLOSE_THREADS"); forced = e ? atoi(e) : 0; }
  if (forced) return forced;
  (void)split_cap; return 256;      // measured on C2: 4-warp CTAs (one wave per level) are not faster than 8-warp ones
}
static void run_plan(tnqs_ctx* c, const Plan& p) {
  if (p.sweep) {
    SweepQueue Q; Q.items = p.d_items; Q.n = p.sweep_items; Q.counter = p.d_qctr; Q.slabs = p.d_slab; Q.closes = p.d_dclose;
    Q.chains = p.d_chains; Q.slab_done = p.d_slab_done; Q.chain_done = p.d_chain_done;
    Q.prof = p.d_prof;
    { const char* e = getenv("TNQS_SWEEP_ABLATE"); Q.ablate = e ? atoi(e) : 0; }
    Q.nchains = (int)p.nchains; Q.ahead = std::max(1, (int)((long long)p.sweep_grid.x * p.nchains / std::max(1, p.sweep_items)));
    if (p.sweepN == 16 && (Q.ablate || Q.prof)) LAUNCH(c, TNQS_KC_MODE_APPLY, (k_bp_sweep<16, true>), p.sweep_grid, QUEUE_THREADS, p.sweep_smem, Q);
    else if (p.sweepN == 16) LAUNCH(c, TNQS_KC_MODE_APPLY, k_bp_sweep<16>, p.sweep_grid, QUEUE_THREADS, p.sweep_smem, Q);
    else LAUNCH(c, TNQS_KC_MODE_APPLY, k_bp_sweep<8>, p.sweep_grid, QUEUE_THREADS, p.sweep_smem, Q);
    if (p.sweep_commit) LAUNCH(c, TNQS_KC_BP_FINALIZE, k_msg_commit, p.sweep_commit, 128, 0, c->msg_cur, c->msg_stage, c->slot, p.d_commit, c->d_chi);
    return;
  }
  for (const LevelPlan& lp : p.levels) {
    for (const SubLevel& sl : lp.subs) {
      if (sl.fused.count) {
        LevelQueue Q; Q.items = p.d_items + sl.fused.off; Q.n = sl.fused.count; Q.counter = p.d_qctr;
        Q.slabs = p.d_slab; Q.closes = p.d_dclose; Q.slab_done = p.d_slab_done;
        if (sl.fastN == 16) LAUNCH(c, TNQS_KC_MODE_APPLY, k_bp_level<16>, sl.fused.grid, QUEUE_THREADS, sl.fused.smem, Q);
        else LAUNCH(c, TNQS_KC_MODE_APPLY, k_bp_level<8>, sl.fused.grid, QUEUE_THREADS, sl.fused.smem, Q);
      } else if (sl.slab.count) {
        if (sl.fastN == 16) {
          LAUNCH(c, TNQS_KC_MODE_APPLY, k_absorb_slab<16>, sl.slab.grid, SLAB_THREADS, sl.slab.smem, p.d_slab + sl.slab.off);
          LAUNCH(c, TNQS_KC_MODE_CLOSE, k_close_dmma<16>, sl.dclose.grid, close_threads(sl.close_cap), sl.dclose.smem, p.d_dclose + sl.dclose.off);
        } else {
          LAUNCH(c, TNQS_KC_MODE_APPLY, k_absorb_slab<8>, sl.slab.grid, SLAB_THREADS, sl.slab.smem, p.d_slab + sl.slab.off);
          LAUNCH(c, TNQS_KC_MODE_CLOSE, k_close_dmma<8>, sl.dclose.grid, close_threads(sl.close_cap), sl.dclose.smem, p.d_dclose + sl.dclose.off);
        }
      }
      if (sl.deep.count) {
        LAUNCH(c, TNQS_KC_MODE_APPLY, k_deep_absorb<8>, sl.deepA.grid, DeepCfg<8>::THREADS, DeepCfg<8>::SMEM_A, p.d_deepA + sl.deepA.off);
        LAUNCH(c, TNQS_KC_MODE_CLOSE, k_deep_close<8>, sl.deep.grid, DeepCfg<8>::THREADS, DeepCfg<8>::SMEM_B, p.d_deep + sl.deep.off);
      }
      if (sl.gate1.count) LAUNCH(c, TNQS_KC_MISC, k_gate1, sl.gate1.grid, 256, 0, p.d_gate1 + sl.gate1.off);
      for (const LaunchRec& lr : sl.steps) {
        if (lr.dmma) LAUNCH(c, TNQS_KC_MODE_APPLY, k_leg_dmma, lr.grid, 256, lr.smem, p.d_apply + lr.off, lr.tpc);
        else LAUNCH(c, TNQS_KC_MODE_APPLY, k_mode_apply, lr.grid, 256, 0, p.d_apply + lr.off);
      }
      if (sl.close.count) {
        if (sl.close.dmma) LAUNCH(c, TNQS_KC_MODE_CLOSE, k_close_leg_dmma, sl.close.grid, 256, 8 * 1024 * sizeof(double), p.d_close + sl.close.off);
        else LAUNCH(c, TNQS_KC_MODE_CLOSE, k_mode_close, sl.close.grid, 256, 0, p.d_close + sl.close.off);
      }
      if (sl.reduce.count) LAUNCH(c, TNQS_KC_BP_FINALIZE, k_close_reduce, sl.reduce.grid, 256, sl.reduce.smem, p.d_reduce + sl.reduce.off);
    }
    if (lp.commit_count)
      LAUNCH(c, TNQS_KC_BP_FINALIZE, k_msg_commit, lp.commit_count, 128, 0, c->msg_cur, c->msg_stage, c->slot,
             p.d_commit + lp.commit_off, c->d_chi);
  }
}
