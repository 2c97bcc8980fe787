// The two GEMM-shaped primitives every tensor operation of the path reduces to (DESIGN.md "kernels"):
//   mode_apply : out[l, c, r] = sum_b in[l, b, r] * M[b, c]         (leg absorption, gate legs, Q*W, Q*R')
//   mode_close : G[b, b']     = sum_{l, r} A[l, b, r] * conj(B[l, b', r])   (message closing, RDMs, scalars)
// Batched over per-item descriptors; a tensor is viewed as [L, n, R] with explicit b / r strides so the same
// kernels serve canonical site layouts, matricised layouts and row blocks of a tall matrix.
#pragma once
#include "common.cuh"

struct ModeApplyDesc {
  const cplx* in; cplx* out; const cplx* m;
  int L, n, np, R;                 // rows = L*R, contracted extent n, produced extent np
  long long in_sb, in_sr, out_sb, out_sr;   // strides (complex elements) of the b and r indices; l stride = 1
  int ldm;                         // leading dimension of m
  int flags;                       // bit0: use conj(m); bit1: use m^T (element (b,c) = m[c + b*ldm])
  // k_leg_dmma<true> only (a product over TWO legs of a tensor in its canonical layout, e.g. (site index, gate bond)): the row
  // index l has stride sl instead of 1, and the contracted / produced index k = k0 + kd * k1 sits at k0 + k1 * sb
  long long in_sl = 1, out_sl = 1; int kd = 1;
  const int* commit = nullptr;     // optimistic gate batches: nothing is written unless *commit != 0
  double* norm2_partial = nullptr; // k_leg_dmma<true> only: CTA x of the item leaves the sum of |out|^2 over its row tiles in [x] (fixed order)
};

#define MA_BM 64
#define MA_BN 64
#define MA_BK 16

// grid: (row tiles, items, col tiles); block 256.  Thread (tx = tid%16, ty = tid/16) owns rows tx+16i, cols ty+16j.
__global__ void __launch_bounds__(256) k_mode_apply(const ModeApplyDesc* __restrict__ descs) {
  const ModeApplyDesc d = descs[blockIdx.y];
  const long long rows = (long long)d.L * d.R;
  const long long row0 = (long long)blockIdx.x * MA_BM;
  const int col0 = blockIdx.z * MA_BN;
  if (row0 >= rows || col0 >= d.np) return;

  __shared__ cplx As[MA_BK][MA_BM];
  __shared__ cplx Ms[MA_BK][MA_BN];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // loader mapping: this thread always loads row (tid % 64) of the tile, k = tid/64 + 4t
  const int lr = tid & 63, lk0 = tid >> 6;
  const long long lrow = row0 + lr;
  long long in_off = -1;
  if (lrow < rows) { long long r = lrow / d.L; in_off = (lrow - r * d.L) + r * d.in_sr; }
  const bool conj_m = d.flags & 1, trans_m = d.flags & 2;

  cplx acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = cmake(0, 0);

  for (int k0 = 0; k0 < d.n; k0 += MA_BK) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int kk = lk0 + 4 * t, k = k0 + kk;
      cplx v = cmake(0, 0);
      if (in_off >= 0 && k < d.n) v = d.in[in_off + (long long)k * d.in_sb];
      As[kk][lr] = v;
      int c = col0 + lr;
      cplx w = cmake(0, 0);
      if (k < d.n && c < d.np) {
        w = trans_m ? d.m[c + (long long)k * d.ldm] : d.m[k + (long long)c * d.ldm];
        if (conj_m) w.y = -w.y;
      }
      Ms[kk][lr] = w;
    }
    __syncthreads();
    const int kmax = min(MA_BK, d.n - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      cplx a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ms[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long row = row0 + tx + 16 * i;
    if (row >= rows) continue;
    long long r = row / d.L;
    long long off = (row - r * d.L) + r * d.out_sr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = col0 + ty + 16 * j;
      if (c < d.np) d.out[off + (long long)c * d.out_sb] = acc[i][j];
    }
  }
}

struct ModeCloseDesc {
  const cplx* a; const cplx* b; cplx* partial;     // partial[split][n*np], column-major n x np
  int L, R, n, np;
  long long a_sb, a_sr, b_sb, b_sr;
  int nsplit; long long rows_per_split;
  const int* need = nullptr;        // k_gram_dmma only: the item is skipped when *need == 0 (adaptive second pass of CholeskyQR2)
};

#define MC_SMEM 2048      // complex elements of shared staging (32 KB)

// grid: (max nsplit, items); block 256.  Threads form G groups of nb*nbp threads; every group accumulates the
// whole n x np tile (4x4 per thread) over its share of the rows; groups are then tree-reduced in a fixed order.
__global__ void __launch_bounds__(256) k_mode_close(const ModeCloseDesc* __restrict__ descs) {
  const ModeCloseDesc d = descs[blockIdx.y];
  if ((int)blockIdx.x >= d.nsplit) return;
  const long long rows = (long long)d.L * d.R;
  const long long r_begin = (long long)blockIdx.x * d.rows_per_split;
  const long long r_end = min(rows, r_begin + d.rows_per_split);

  __shared__ cplx sm[MC_SMEM];
  const int n = d.n, np = d.np;
  const int nb = (n + 3) >> 2, nbp = (np + 3) >> 2;
  const int gsz = nb * nbp;                 // threads per group (<= 256 for n, np <= 64)
  const int G = 256 / gsz;                  // groups
  const int tid = threadIdx.x;
  const int g = tid / gsz, t = tid - g * gsz;
  const bool active = g < G;
  const int bx = t % nb, by = t / nb;
  int RC = MC_SMEM / (n + np);              // rows per staged chunk
  if (RC > 256) RC = 256;
  cplx* As = sm;
  cplx* Bs = sm + (size_t)RC * n;

  cplx acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = cmake(0, 0);

  for (long long c0 = r_begin; c0 < r_end; c0 += RC) {
    const int rc = (int)min((long long)RC, r_end - c0);
    for (int idx = tid; idx < rc * n; idx += 256) {
      int rr = idx % rc, b = idx / rc;
      long long row = c0 + rr, r = row / d.L;
      As[rr * n + b] = d.a[(row - r * d.L) + r * d.a_sr + (long long)b * d.a_sb];
    }
    for (int idx = tid; idx < rc * np; idx += 256) {
      int rr = idx % rc, b = idx / rc;
      long long row = c0 + rr, r = row / d.L;
      Bs[rr * np + b] = d.b[(row - r * d.L) + r * d.b_sr + (long long)b * d.b_sb];
    }
    __syncthreads();
    if (active) {
      for (int rr = g; rr < rc; rr += G) {
        cplx a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { int bi = bx + nb * i; a[i] = bi < n ? As[rr * n + bi] : cmake(0, 0); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { int bj = by + nbp * j; b[j] = bj < np ? Bs[rr * np + bj] : cmake(0, 0); }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) cfma_cb(acc[i][j], a[i], b[j]);
      }
    }
    __syncthreads();
  }
  // fixed-order tree reduction over the groups through the staging buffer: in the round with stride s the
  // groups [s, 2s) hand their accumulators to the groups [0, s).  Writers never exceed G/2, so
  // writers * 16 * gsz <= 2048 complex = MC_SMEM.
  const int slot = 16 * gsz;
  int P = 1;
  while (P < G) P <<= 1;
  for (int s = P >> 1; s >= 1; s >>= 1) {
    if (active && g >= s && g < 2 * s) {
      cplx* dst = sm + (size_t)(g - s) * slot + t * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (active && g < s && g + s < G) {
      const cplx* src = sm + (size_t)g * slot + t * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cadd(acc[i][j], src[i * 4 + j]);
    }
    __syncthreads();
  }
  if (active && g == 0) {
    cplx* out = d.partial + (size_t)blockIdx.x * n * np;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int bi = bx + nb * i;
      if (bi >= n) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int bj = by + nbp * j;
        if (bj < np) out[bi + (size_t)bj * n] = acc[i][j];
      }
    }
  }
}

// ---- reduction of the split partials + epilogues ---------------------------------------------------------------
struct CloseReduceDesc {
  const cplx* partial; int nsplit; int n, np;
  cplx* dst;                 // STORE: n x np column-major (ld = n); BP: staged message slot
  const cplx* prev;          // BP: current message (for the diff), may be null
  double* diff;              // BP: where 1 - fidelity goes (may be null)
  int mode;                  // 0 = store, 1 = BP message epilogue
  int normalize, hermitize;
  // peer-memory halo (vertex-partitioned BP): a boundary message is also written straight into the halo buffer of the
  // GPU that owns its destination vertex (a P2P store over NVLink); the buffer is double-buffered by sweep parity
  cplx* dst2 = nullptr;            // slot in the peer's halo buffer for parity 0, or null
  long long dst2_pstride = 0;      // elements between the two parity halves of the peer's buffer
  const int* parity = nullptr;     // device word holding the parity of the running sweep
};

// Sum of the split partials + epilogue of one item, executed by a whole CTA (any block size); fixed summation order
// everywhere (deterministic).  gsm: n*np complex of shared memory; red: >= 32 doubles of shared memory.
__device__ __forceinline__ void close_reduce_body(const CloseReduceDesc& d, cplx* gsm, double* red) {
  const int nn = d.n * d.np;
  for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
    cplx s = cmake(0, 0);
    for (int p = 0; p < d.nsplit; ++p) s = cadd(s, __ldcg(d.partial + (size_t)p * nn + idx));      // partials and the previous message come from other CTAs: read through L2
    gsm[idx] = s;
  }
  __syncthreads();
  if (d.mode == 0) {
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) d.dst[idx] = gsm[idx];
    return;
  }
  // BP epilogue (abstractbeliefpropagationcache.jl:156-165, beliefpropagationcache.jl:15-19)
  const int n = d.n;
  double sre = 0, sim = 0;
  for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
    cplx v = gsm[idx];
    if (d.hermitize) {
      int i = idx % n, j = idx / n;
      cplx w = gsm[j + i * n];
      v = cmake(0.5 * (v.x + w.x), 0.5 * (v.y - w.y));
    }
    sre += v.x; sim += v.y;
  }
  sre = block_sum(sre, red);
  sim = block_sum(sim, red);
  cplx inv = cmake(1, 0);
  if (d.normalize && !(sre == 0.0 && sim == 0.0)) {
    double den = sre * sre + sim * sim;
    inv = cmake(sre / den, -sim / den);
  }
  double dre = 0, dim = 0, na = 0, nb = 0;
  cplx* const dst2 = d.dst2 ? d.dst2 + (size_t)(*d.parity) * d.dst2_pstride : nullptr;
  for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
    cplx v = gsm[idx];
    if (d.hermitize) {
      int i = idx % n, j = idx / n;
      cplx w = gsm[j + i * n];
      v = cmake(0.5 * (v.x + w.x), 0.5 * (v.y - w.y));
    }
    v = cmul(v, inv);
    cplx p = d.diff ? __ldcg(d.prev + idx) : cmake(0, 0);      // read before the store: dst may alias prev (in-place message update)
    d.dst[idx] = v;
    if (dst2) dst2[idx] = v;
    if (d.diff) {
      dre += v.x * p.x + v.y * p.y;      // conj(v) * p
      dim += v.x * p.y - v.y * p.x;
      na += cabs2(v); nb += cabs2(p);
    }
  }
  if (d.diff) {
    dre = block_sum(dre, red); dim = block_sum(dim, red);
    na = block_sum(na, red); nb = block_sum(nb, red);
    if (threadIdx.x == 0) {
      double sa = sqrt(na), sb = sqrt(nb);
      double fr = dre / (sa * sb), fi = dim / (sa * sb);
      *d.diff = 1.0 - (fr * fr + fi * fi);
    }
  }
}

// one CTA (256 threads) per item
__global__ void __launch_bounds__(256) k_close_reduce(const CloseReduceDesc* __restrict__ descs) {
  const CloseReduceDesc d = descs[blockIdx.x];
  __shared__ double red[32];
  extern __shared__ cplx gsm[];                 // n*np complex
  close_reduce_body(d, gsm, red);
}
