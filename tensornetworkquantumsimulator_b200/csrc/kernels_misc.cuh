// HBM-bound helpers: layout permutation, one-site gates, norms / scaling, message bookkeeping, RNG fill.
// These move each element once with 16-byte accesses; their roofline is HBM (DESIGN.md "kernels").
#pragma once
#include "common.cuh"

// ---- canonical [s, b_1..b_z]  <->  matricised [outer legs (ascending), s, b_gate] ----------------------------
struct PermDesc {
  const cplx* src; cplx* dst;
  int d, chib, nouter;
  long long bstride;               // canonical stride of the gate-bond leg
  int odim[TNQS_MAX_LEGS];
  long long ostride[TNQS_MAX_LEGS];   // canonical strides of the outer legs
  long long Po;                    // product of outer dims
  int to_matrix;                   // 1: canonical -> matrix; 0: matrix -> canonical
  const double* norm2;             // matrix -> canonical only: multiply by 1/sqrt(*norm2) if non-null
  const int* commit;               // optimistic gate batches: the result is written only if *commit != 0 (null: always)
};

// grid: (ceil(max Po / 256), items); a thread owns one outer multi-index rho: it decomposes rho once, then walks the bond
// index, moving the d contiguous site-index values of the canonical side (a full 32-byte sector at d = 2) to / from d
// matrix columns (coalesced across the warp on both sides when the first outer leg is the first canonical leg).
__global__ void __launch_bounds__(256) k_permute(const PermDesc* __restrict__ descs) {
  const PermDesc p = descs[blockIdx.y];
  const long long rho = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (rho >= p.Po) return;
  if (p.commit && *p.commit == 0) return;
  double scale = 1.0;
  if (!p.to_matrix && p.norm2) { double n2 = *p.norm2; scale = n2 > 0.0 ? 1.0 / sqrt(n2) : 1.0; }
  long long off = 0, r = rho;
  for (int i = 0; i < p.nouter; ++i) { long long q = r / p.odim[i]; off += (r - q * p.odim[i]) * p.ostride[i]; r = q; }
  const int d = p.d;
  if (p.to_matrix) {
    for (int b = 0; b < p.chib; ++b)
      for (int s = 0; s < d; ++s) p.dst[rho + p.Po * (s + (long long)d * b)] = p.src[off + (long long)b * p.bstride + s];
  } else {
    for (int b = 0; b < p.chib; ++b)
      for (int s = 0; s < d; ++s) p.dst[off + (long long)b * p.bstride + s] = cscale(p.src[rho + p.Po * (s + (long long)d * b)], scale);
  }
}

// ---- one-site gates: out[s', rest] = sum_s o[s', s] in[s, rest]  (simple_update.jl:21-23) -----------------------
struct Gate1Desc { const cplx* in; cplx* out; const cplx* gate; int d; long long rest; double* norm2_partial; };

// grid: (blocks, items); block 256; each block writes one partial of sum |out|^2 (fixed order).
__global__ void __launch_bounds__(256) k_gate1(const Gate1Desc* __restrict__ descs) {
  const Gate1Desc g = descs[blockIdx.y];
  __shared__ cplx o[16];
  __shared__ double red[8];
  if (threadIdx.x < g.d * g.d) o[threadIdx.x] = g.gate[threadIdx.x];
  __syncthreads();
  double acc = 0;
  const int d = g.d;
  if (d == 2) {                                  // qubits: everything in registers, two fibres in flight per thread
    const cplx o00 = o[0], o10 = o[1], o01 = o[2], o11 = o[3];
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; r + stride < g.rest; r += 2 * stride) {
      const cplx a0 = g.in[2 * r], a1 = g.in[2 * r + 1], b0 = g.in[2 * (r + stride)], b1 = g.in[2 * (r + stride) + 1];
      cplx y0 = cmake(0, 0), y1 = cmake(0, 0), z0 = cmake(0, 0), z1 = cmake(0, 0);
      cfma(y0, o00, a0); cfma(y0, o01, a1); cfma(y1, o10, a0); cfma(y1, o11, a1);
      cfma(z0, o00, b0); cfma(z0, o01, b1); cfma(z1, o10, b0); cfma(z1, o11, b1);
      acc += cabs2(y0); acc += cabs2(y1); acc += cabs2(z0); acc += cabs2(z1);
      g.out[2 * r] = y0; g.out[2 * r + 1] = y1; g.out[2 * (r + stride)] = z0; g.out[2 * (r + stride) + 1] = z1;
    }
    for (; r < g.rest; r += stride) {
      const cplx a0 = g.in[2 * r], a1 = g.in[2 * r + 1];
      cplx y0 = cmake(0, 0), y1 = cmake(0, 0);
      cfma(y0, o00, a0); cfma(y0, o01, a1); cfma(y1, o10, a0); cfma(y1, o11, a1);
      acc += cabs2(y0); acc += cabs2(y1);
      g.out[2 * r] = y0; g.out[2 * r + 1] = y1;
    }
  } else
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < g.rest; r += (long long)gridDim.x * blockDim.x) {
    cplx x[4], y[4];
    for (int s = 0; s < d; ++s) x[s] = g.in[r * d + s];
    for (int sp = 0; sp < d; ++sp) {
      cplx a = cmake(0, 0);
      for (int s = 0; s < d; ++s) cfma(a, o[sp + d * s], x[s]);
      y[sp] = a; acc += cabs2(a);
    }
    for (int sp = 0; sp < d; ++sp) g.out[r * d + sp] = y[sp];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && g.norm2_partial) g.norm2_partial[blockIdx.x] = acc;
}

// ---- sum |x|^2 per item (one CTA per item, fixed order) ---------------------------------------------------------
struct Norm2Desc { const cplx* x; long long n; double* out; const double* partial; int npartial; };
__global__ void __launch_bounds__(1024) k_norm2(const Norm2Desc* __restrict__ descs) {
  const Norm2Desc d = descs[blockIdx.x];
  __shared__ double red[32];
  double acc = 0;
  if (d.partial) { for (int i = threadIdx.x; i < d.npartial; i += blockDim.x) acc += d.partial[i]; }
  else { for (long long i = threadIdx.x; i < d.n; i += blockDim.x) acc += cabs2(d.x[i]); }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *d.out = acc;
}

// ---- x *= factor (complex); factor = f(*scalar) ----------------------------------------------------------------
// mode 0: 1/sqrt(norm2[0]) (Frobenius normalisation, simple_update.jl:63-65)
// mode 1: rescale_vertices!: s / sqrt(vn), vn = scalar[0] + i scalar[1], s = sign(vn) if real else 1 (bpcache.jl:87-106)
// mode 0: x / sqrt(scalar[0]) (real norm^2); 1: x * sign / sqrt(complex scalar) (rescale_vertices! of a state); 2: x / complex scalar
struct ScaleDesc { cplx* x; long long n; const double* scalar; int mode; const int* commit = nullptr; };     // commit: see PermDesc
__global__ void __launch_bounds__(256) k_scale(const ScaleDesc* __restrict__ descs) {
  const ScaleDesc d = descs[blockIdx.y];
  if (d.commit && *d.commit == 0) return;
  cplx f;
  if (d.mode == 0) { double n2 = d.scalar[0]; f = cmake(n2 > 0.0 ? 1.0 / sqrt(n2) : 1.0, 0.0); }
  else {
    double re = d.scalar[0], im = d.scalar[1];
    double mod = sqrt(re * re + im * im);
    // principal complex sqrt
    double sr = sqrt(0.5 * (mod + re)), si = copysign(sqrt(0.5 * (mod - re)), im);
    if (re < 0 && im == 0.0) { sr = 0.0; si = sqrt(-re); }
    double den = sr * sr + si * si;
    cplx inv = cmake(sr / den, -si / den);
    cplx sg = cmake(1, 0);
    if (im == 0.0) sg = cmake(re > 0 ? 1.0 : (re < 0 ? -1.0 : 0.0), 0.0);
    f = cmul(sg, inv);
    if (d.mode == 2) { const double m2 = re * re + im * im; f = cmake(re / m2, -im / m2); }     // 1 / scalar
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (long long)gridDim.x * blockDim.x)
    d.x[i] = cmul(d.x[i], f);
}

// ---- messages ---------------------------------------------------------------------------------------------------
// identity default for every directed edge (tensornetworkstate.jl:64-67); grid = 2*ne
__global__ void k_msg_identity(cplx* slab, long long slot, const int* chi) {
  int k = blockIdx.x, n = chi[k >> 1];
  cplx* m = slab + (size_t)k * slot;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) m[idx] = cmake((idx % n) == (idx / n) ? 1.0 : 0.0, 0.0);
}
// commit the staged messages of a level: cur[k] <- stage[k] for k in list
__global__ void k_msg_commit(cplx* cur, const cplx* stage, long long slot, const int* list, const int* chi) {
  int k = list[blockIdx.x], n = chi[k >> 1];
  const cplx* s = stage + (size_t)k * slot; cplx* c = cur + (size_t)k * slot;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) c[idx] = s[idx];
}
// commit the messages the peers wrote into this handle's halo buffer: cur[list[j]] <- halo[half][j], half = the parity the
// sweep that just ended wrote to (= *next_parity ^ 1: the barrier has already advanced the word)
__global__ void k_halo_commit(cplx* cur, const cplx* halo, long long slot, int nrecv, const int* list, const int* chi, const int* next_parity) {
  const int k = list[blockIdx.x], n = chi[k >> 1];
  const int half = (*next_parity) ^ 1;
  const cplx* s = halo + ((size_t)half * nrecv + blockIdx.x) * slot; cplx* c = cur + (size_t)k * slot;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) c[idx] = s[idx];
}

// Device-side barrier + all-reduce of one BP sweep of a vertex-partitioned run (one CTA, thread t talks to rank t).
// Every rank owns a signal area of world x 2 slots of 4 doubles {epoch, diff sum, message count, -}; rank r writes its
// slot [r][epoch & 1] in EVERY rank's area with P2P stores over NVLink (payload, system fence, epoch), then waits until
// all slots of its own area carry the epoch: at that point every peer has finished the sweep, so the boundary messages
// its kernels stored in this rank's halo buffer (earlier in its stream) have landed.  The global sum is taken in rank
// order - the same bits on every rank, so all ranks take the same "stop / continue" decision in their WHILE graph.
// Slots alternate with the epoch parity because a peer may already be one sweep ahead.  The spin has a time-out.
constexpr int HALO_MAX_WORLD = 64;
constexpr int HALO_SIG_DOUBLES = HALO_MAX_WORLD * 2 * 4;
struct HaloBarrierArgs {
  double* sig[HALO_MAX_WORLD];     // signal area of every rank (own area at [rank])
  int world, rank;
  int* epoch;                      // device word: sweeps completed by this handle since the halo was opened
  int* next_parity;                // device word: half of the halo buffers the NEXT sweep writes to
  const double* local_sum;         // may be null (no convergence test): contributes 0
  double local_cnt;
  double* out_sum; double* out_cnt;
  int* status; int timeout_bit;
};
__global__ void __launch_bounds__(HALO_MAX_WORLD) k_halo_barrier(const HaloBarrierArgs a) {
  const int t = threadIdx.x;
  const int e = *a.epoch + 1, p = e & 1;
  __shared__ int s_bad;
  if (t == 0) s_bad = 0;
  __syncthreads();
  if (t < a.world) {
    double* dst = a.sig[t] + (size_t)(a.rank * 2 + p) * 4;
    dst[1] = a.local_sum ? *a.local_sum : 0.0;
    dst[2] = a.local_cnt;
    __threadfence_system();
    *(volatile double*)dst = (double)e;
    __threadfence_system();
    const volatile double* mine = a.sig[a.rank] + (size_t)(t * 2 + p) * 4;
    long long spins = 0;
    while (*mine != (double)e) {
      __nanosleep(200);
      if (++spins > (1ll << 23)) { s_bad = 1; break; }              // a few seconds: a peer never arrived
    }
    __threadfence_system();
  }
  __syncthreads();
  if (t == 0) {
    double s = 0, c = 0;
    const double* mine = a.sig[a.rank];
    for (int r = 0; r < a.world; ++r) { s += __ldcg(mine + (size_t)(r * 2 + p) * 4 + 1); c += __ldcg(mine + (size_t)(r * 2 + p) * 4 + 2); }
    *a.out_sum = s; *a.out_cnt = c;
    *a.epoch = e; *a.next_parity = (e + 1) & 1;
    if (s_bad) atomicOr(a.status, a.timeout_bit);
  }
}
__global__ void k_halo_pack(const cplx* slab, long long slot, const int* list, cplx* buf, int pack) {
  int k = list[blockIdx.x];
  cplx* s = const_cast<cplx*>(slab) + (size_t)k * slot; cplx* b = buf + (size_t)blockIdx.x * slot;
  for (int idx = threadIdx.x; idx < slot; idx += blockDim.x) { if (pack) b[idx] = s[idx]; else s[idx] = b[idx]; }
}
// fixed-order sum of n doubles -> out[0]
__global__ void __launch_bounds__(1024) k_sum(const double* x, int n, double* out) {
  __shared__ double red[32];
  double acc = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *out = acc;
}
// edge scalars: out[e] = sum_{b,b'} m_{2e}[b,b'] * m_{2e+1}[b,b']  (beliefpropagationcache.jl:38-40); grid = ne
__global__ void __launch_bounds__(256) k_edge_scalar(const cplx* slab, long long slot, const int* chi, cplx* out) {
  int e = blockIdx.x, n = chi[e];
  const cplx* a = slab + (size_t)(2 * e) * slot; const cplx* b = a + slot;
  __shared__ double red[8];
  double re = 0, im = 0;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) { cplx p = cmul(a[idx], b[idx]); re += p.x; im += p.y; }
  re = block_sum(re, red); im = block_sum(im, red);
  if (threadIdx.x == 0) out[e] = cmake(re, im);
}
// rescale_messages! (beliefpropagationcache.jl:129-142); grid = ne
__global__ void __launch_bounds__(256) k_rescale_messages(cplx* slab, long long slot, const int* chi) {
  int e = blockIdx.x, n = chi[e];
  cplx* a = slab + (size_t)(2 * e) * slot; cplx* b = a + slot;
  __shared__ double red[8];
  double na = 0, nb = 0;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) { na += cabs2(a[idx]); nb += cabs2(b[idx]); }
  na = sqrt(block_sum(na, red)); nb = sqrt(block_sum(nb, red));
  double re = 0, im = 0;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    cplx p = cmul(cscale(a[idx], 1.0 / na), cscale(b[idx], 1.0 / nb)); re += p.x; im += p.y;
  }
  re = block_sum(re, red); im = block_sum(im, red);
  double sg = 1.0;
  if (im == 0.0) { sg = re > 0 ? 1.0 : (re < 0 ? -1.0 : 0.0); re *= sg; }
  double mod = sqrt(re * re + im * im);
  double sr = sqrt(0.5 * (mod + re)), si = copysign(sqrt(0.5 * (mod - re)), im);
  double den = sr * sr + si * si;
  cplx inv = cmake(sr / den, -si / den);          // 1/sqrt(n)
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    a[idx] = cmul(cscale(a[idx], sg / na), inv);
    b[idx] = cmul(cscale(b[idx], 1.0 / nb), inv);
  }
}

// ---- counter-based N(0,1)+iN(0,1) fill (same formula as tensornetworkquantumsimulator_b200/rng.py) --------------
__device__ __forceinline__ unsigned long long sm64_mix(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) k_fill_random(cplx* x, long long n, unsigned long long seed, unsigned long long stream) {
  const unsigned long long G = 0x9E3779B97F4A7C15ull;
  const unsigned long long base = sm64_mix(seed * G + stream + 1ull);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z1 = sm64_mix(base + (2ull * i + 1ull) * G), z2 = sm64_mix(base + (2ull * i + 2ull) * G);
    double u1 = ((double)(z1 >> 11) + 1.0) * (1.0 / 9007199254740992.0);
    double u2 = (double)(z2 >> 11) * (1.0 / 9007199254740992.0);
    double r = sqrt(-2.0 * log(u1)), s, c;
    sincos(2.0 * 3.14159265358979323846 * u2, &s, &c);
    x[i] = cmake(r * c, r * s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Loop corrections (loop_corrections.cuh).
// k_antiproject: X <- X - mf * (mb | X) with (a | b) = sum_{i,j} a[i,j] b[i,j]: the antiprojector 1 - |m_e)(m_rev| of
// src/MessagePassing/loopcorrection.jl:49-60 applied to an environment that crosses the edge (mf = the BP message travelling the
// same way as X, mb = the one travelling against it).  One CTA (128 threads) per item; dense n x n blocks, nn = n * n.
// ---------------------------------------------------------------------------------------------------------------
struct AntiProjDesc { cplx* x; const cplx* mf; const cplx* mb; int nn; };
__global__ void __launch_bounds__(128) k_antiproject(const AntiProjDesc* __restrict__ descs) {
  const AntiProjDesc d = descs[blockIdx.x];
  __shared__ double red[8];
  double re = 0, im = 0;
  for (int i = threadIdx.x; i < d.nn; i += blockDim.x) { const cplx p = cmul(d.mb[i], d.x[i]); re += p.x; im += p.y; }
  re = block_sum(re, red); im = block_sum(im, red);
  const cplx t = cmake(re, im);
  for (int i = threadIdx.x; i < d.nn; i += blockDim.x) { const cplx p = cmul(t, d.mf[i]); d.x[i] = cmake(d.x[i].x - p.x, d.x[i].y - p.y); }
}
// the n x n identity into a zeroed buffer (rows of it are the chi x chi matrix units, n = chi^2)
__global__ void k_unit_diagonal(cplx* u, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) u[i + i * n] = cmake(1, 0);
}
