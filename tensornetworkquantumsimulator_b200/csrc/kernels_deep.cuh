// FP64 tensor-core (DMMA.8x8x4) path of the BP message update on HIGH-COORDINATION vertices: uniform bond dimension N = 8,
// site dimension 2, 4 or 5 absorbed legs (degree 5 and 6: the faces and the interior of the 8x8x8 cubic lattice, BASELINE
// config C4, examples/3dIsing_dynamics.jl).  There the site tensor (8 MiB at degree 6) is 64x larger than anything a CTA can
// hold and one leg absorption has an arithmetic intensity of 2 flop/B, so a chain of 5 separate leg passes + a closing pass
// (96 MiB of traffic per message) is HBM-bound at a fifth of the FP64 peak.  Two kernels bring that down to 32 MiB:
//
//   k_deep_absorb : one CTA per (s, p4, w) block of 8^4 elements: the first three absorbed legs p0, p1, p2 go through shared
//                   memory back to back (the slab machinery of k_absorb_slab; warp p3 owns sub-slab p3 through all three phases,
//                   so there is no CTA-wide barrier between them); the result is written tile-major,
//                   XA[(s, p1', p2')][w][p4][p3][p0'], in full 128-byte runs.
//   k_deep_close  : one CTA per (s, p1', p2') tile (contiguous 64 KB of XA): absorbs p3 on the way into shared memory and p4 in
//                   place, then closes G[w, w'] += sum_{p0', p3', p4'} X[p0', p3', p4', w] conj(B[s, p0', p1', p2', p3', p4', w'])
//                   with the X operand in shared memory and the bra operand streamed from global memory as DMMA fragments:
//                   the fully absorbed tensor never exists in memory.  Each CTA leaves one N x N partial; k_close_reduce sums
//                   the 2 N^2 partials of a message in a fixed order and runs the message epilogue.
//                   In a JACOBI sweep the six updates that leave a degree-6 vertex share two k_deep_absorb items (legs 0-2 for the
//                   updates on legs 3-5 and vice versa): the tile then keeps its three open legs in ascending order and every
//                   update reads it with its own strides (DeepDesc::xs; engine.cuh build_plan).
//   k_slab_batch  : the same warp-owned sub-slab with free strides, for sqrt(M) absorptions of the gate path (at the end of this file).
// (updated_message, src/MessagePassing/abstractbeliefpropagationcache.jl:144-166; same contraction, another association order.)
// Complex products are 3M (three real DMMA products), with the fragment-column permutation of absorb_phase (kernels_dmma.cuh).
#pragma once
#include "kernels_dmma.cuh"

struct DeepDesc {
  const cplx* t;          // site tensor (ket), canonical layout T[s, l0 .. l_{z-1}]
  const cplx* bt;         // bra site tensor of the closing (the ket itself for a norm network), same layout
  cplx* x;                // tile-major intermediate, 2 N^(F+1) complex
  cplx* partial;          // 2 N^2 partial results of N x N complex (column-major, M[w + N w'])
  const cplx* m[5];       // messages of the absorbed legs p0 .. p4, in absorption order (N x N, ld = N)
  long long g[5];         // canonical strides (complex elements) of p0 .. p4 (g[4] = 0 with four absorbed legs)
  long long gw;           // canonical stride of the outgoing leg
  int E4;                 // extent of p4: N (five absorbed legs) or 1 (four)
  int xs[3];              // k_deep_close: strides of p3, p4 and w inside a tile (N, N^2, N^2 E4 for a tile of its own; a tile shared by the
                          // three updates of a degree-6 vertex keeps its open legs in ascending order and each update permutes them)
};

#ifndef TNQS_DEEP_THREADS
#define TNQS_DEEP_THREADS 256
#endif
#ifndef TNQS_DEEP_CTAS
#define TNQS_DEEP_CTAS 3
#endif
#ifndef TNQS_DEEP_DEPTH
#define TNQS_DEEP_DEPTH 2      // row tiles in flight per warp in the phases that read global memory
#endif
#ifndef TNQS_DEEP_DEPTH_SMEM
#define TNQS_DEEP_DEPTH_SMEM 1 // likewise in the phases that run inside shared memory
#endif
#ifndef TNQS_DEEP_ABLATE
#define TNQS_DEEP_ABLATE 0     // experiments only (wrong results): 1 absorb: no global stores, 2 absorb: no in-place phases, 4 close: no closing loop,
#endif                         // 8 close: no absorption phases, 16 absorb: no global loads
constexpr int DEEP_THREADS = TNQS_DEEP_THREADS, DEEP_CTAS = TNQS_DEEP_CTAS;
template <int N> struct DeepCfg {
  static constexpr int SLAB = N * N * N * N;                                           // one (s, p4, w) block / one tile at E4 = N
  static constexpr size_t SMEM_A = (size_t)(SLAB + 3 * N * N) * sizeof(cplx);
  static constexpr int SB = N * N * N + 1;                                             // sub-slab stride of k_deep_close (see there)
  static constexpr size_t SMEM_B = (size_t)(N * SB + 2 * N * N) * sizeof(cplx);       // the cross-warp reduction reuses the slab
  static constexpr int THREADS = DEEP_THREADS;
};

// One absorption  out[row, c] = sum_b in[row, b] m[b, c]  over `ntiles` tiles of 8 rows, 3M form.  src(row, b) / dst(row, c) give the
// addresses; a lane (r, q) loads the elements b = 4 kk + q of row 8 ti + r and stores the columns c = 4 nt + q of the same row, so a
// phase may run in place when a row's inputs and outputs are the same set of addresses.  The next tile's elements are requested while
// the current tile multiplies.
template <int N, int D, class Src, class Dst>
__device__ __forceinline__ void deep_phase(const cplx* __restrict__ msm, int first, int count, Src src, Dst dst, int lane) {
  constexpr int KK = N / 4, NT3 = N / 8;
  const int r = lane >> 2, q = lane & 3;
  double br[KK][NT3], bi[KK][NT3], bs[KK][NT3];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk)
#pragma unroll
    for (int nt = 0; nt < NT3; ++nt) {
      const int b = 4 * kk + q, c = 4 * (2 * nt + (r & 1)) + (r >> 1);
      const cplx v = msm[b + N * c];
      br[kk][nt] = v.x; bi[kk][nt] = v.y; bs[kk][nt] = v.x + v.y;
    }
  // tiles first .. first + count - 1, D of them in flight: a[j] holds tile base + j and is refilled with tile base + j + D as soon as
  // it is consumed
  const int last = first + count;
  cplx a[D][KK];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    if (first + j < last) {
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) a[j][kk] = *src(8 * (first + j) + r, 4 * kk + q);
    }
  }
  for (int base = first; base < last; base += D) {
#pragma unroll
    for (int j = 0; j < D; ++j) {
      const int ti = base + j;
      if (ti >= last) break;                         // warp-uniform
      const int tn = ti + D;
      const bool more = tn < last;
      double t1[NT3][2], t2[NT3][2], t3[NT3][2];
#pragma unroll
      for (int nt = 0; nt < NT3; ++nt) { t1[nt][0] = t1[nt][1] = t2[nt][0] = t2[nt][1] = t3[nt][0] = t3[nt][1] = 0.0; }
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        const double ax = a[j][kk].x, ay = a[j][kk].y, as = ax + ay;
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) dmma884(t1[nt][0], t1[nt][1], ax, br[kk][nt]);
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) dmma884(t2[nt][0], t2[nt][1], ay, bi[kk][nt]);
#pragma unroll
        for (int nt = 0; nt < NT3; ++nt) dmma884(t3[nt][0], t3[nt][1], as, bs[kk][nt]);
        if (more) a[j][kk] = *src(8 * tn + r, 4 * kk + q);
      }
#pragma unroll
      for (int nt = 0; nt < NT3; ++nt)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          *dst(8 * ti + r, 4 * (2 * nt + jj) + q) = cmake(t1[nt][jj] - t2[nt][jj], t3[nt][jj] - t1[nt][jj] - t2[nt][jj]);
    }
  }
}

// grid: (2 N N, items); block 32 N threads; dynamic shared memory DeepCfg<N>::SMEM_A.  CTA x = s + 2 (p4 + E4 w).
// Shared-memory slab: element (i0, i1, i2, i3) at (i0 ^ f(i1) ^ f(i2)) + N i1 + N^2 i2 + N^3 i3 (the swizzle of k_absorb_slab inside every
// sub-slab i3 = p3).  No phase mixes different p3, so WARP p3 OWNS SUB-SLAB p3 through all three phases: after the messages are staged
// there is no CTA-wide barrier, and the warps of an SM drift apart - the global loads of one overlap the tensor-core phases of another.
template <int N>
__global__ void __launch_bounds__(DEEP_THREADS, DEEP_CTAS) k_deep_absorb(const DeepDesc* __restrict__ descs) {
  static_assert(DEEP_THREADS == 32 * N, "one warp per sub-slab");
  const DeepDesc& d = descs[blockIdx.y];
  const int E4 = d.E4;
  if ((int)blockIdx.x >= 2 * E4 * N) return;
  extern __shared__ cplx dsm[];
  cplx* const sm = dsm;
  cplx* const msm = dsm + DeepCfg<N>::SLAB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x & 1, rest = blockIdx.x >> 1, p4 = rest % E4, w = rest / E4;
  {
    const cplx *m0 = d.m[0], *m1 = d.m[1], *m2 = d.m[2];
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) { msm[idx] = __ldg(m0 + idx); msm[N * N + idx] = __ldg(m1 + idx); msm[2 * N * N + idx] = __ldg(m2 + idx); }
  }
  const cplx* const gsrc = d.t + s + (long long)p4 * d.g[4] + (long long)w * d.gw + (long long)warp * d.g[3];
  const long long g0 = d.g[0], g1 = d.g[1], g2 = d.g[2];
  const long long TS = (long long)N * N * N * E4;
  cplx* const gdst = d.x + (long long)N * N * s * TS + (long long)N * N * p4 + (long long)N * N * E4 * w + N * warp;
  cplx* const sub = sm + N * N * N * warp;
  auto at = [sub](int i0, int i1, int i2) -> cplx* { return sub + ((i0 ^ slab_f(i1) ^ slab_f(i2)) + N * i1 + N * N * i2); };
  constexpr int NTILES = N * N / 8;                  // per warp: rows (o1, o2) of its sub-slab
  __syncthreads();
  // p0: rows (p1, p2) straight from global memory -> (i0 = p0', i1 = p1, i2 = p2)
  deep_phase<N, TNQS_DEEP_DEPTH>(msm, 0, NTILES,
                [=](int rho, int b) -> const cplx* { if (TNQS_DEEP_ABLATE & 16) return msm + (rho + b) % (3 * N * N); return gsrc + (rho % N) * g1 + (rho / N) * g2 + b * g0; },
                [=](int rho, int c) -> cplx* { return at(c, rho % N, rho / N); }, lane);
  __syncwarp();
  // p1 in place: rows (i0, i2)
  if (!(TNQS_DEEP_ABLATE & 2)) deep_phase<N, TNQS_DEEP_DEPTH_SMEM>(msm + N * N, 0, NTILES,
                [=](int rho, int b) -> const cplx* { return at(rho % N, b, rho / N); },
                [=](int rho, int c) -> cplx* { return at(rho % N, c, rho / N); }, lane);
  __syncwarp();
  // p2: rows (i0 = p0', i1 = p1') -> XA[(s, p1', p2')][w][p4][p3][p0']
  deep_phase<N, TNQS_DEEP_DEPTH_SMEM>(msm + 2 * N * N, 0, NTILES,
                [=](int rho, int b) -> const cplx* { return at(rho % N, rho / N, b); },
                [=](int rho, int c) -> cplx* { if (TNQS_DEEP_ABLATE & 1) return at(rho % N, rho / N, c); return gdst + (rho / N + N * c) * TS + rho % N; }, lane);
}

// grid: (2 N N, items); block DEEP_THREADS; dynamic shared memory DeepCfg<N>::SMEM_B.  CTA y = p1' + N p2' + N^2 s.
// Shared-memory slab: element (i0 = p3', i1 = p0', i2 = p4, i3 = w) at (i0 ^ f(i1) ^ f(i2)) + N i1 + N^2 i2 + SB i3 with SB = N^3 + 1: the
// closing reads 16-byte X elements for 2 values of w x 4 consecutive p0' per quarter-warp - the swizzle spreads the four p0' over the
// even 16-byte bank groups and the odd sub-slab stride puts the second w on the odd ones.
template <int N>
__global__ void __launch_bounds__(DEEP_THREADS, DEEP_CTAS) k_deep_close(const DeepDesc* __restrict__ descs) {
  static_assert(DEEP_THREADS == 32 * N, "one warp per sub-slab");
  const DeepDesc& d = descs[blockIdx.y];
  const int E4 = d.E4;
  constexpr int SB = DeepCfg<N>::SB;
  extern __shared__ cplx dsm[];
  cplx* const sm = dsm;
  cplx* const msm = dsm + N * SB;
  cplx* const red = dsm;                             // after the closing loop: nwarps x N^2 complex on top of the slab
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int y = blockIdx.x, p1 = y % N, p2 = (y / N) % N, s = y / (N * N);
  {
    const cplx *m3 = d.m[3], *m4 = d.m[4];
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) { msm[idx] = __ldg(m3 + idx); if (E4 == N) msm[N * N + idx] = __ldg(m4 + idx); }
  }
  const long long TS = (long long)N * N * N * E4;
  const int xs3 = d.xs[0], xs4 = d.xs[1];
  const cplx* const xt = d.x + (long long)y * TS + (long long)d.xs[2] * warp;          // warp w owns sub-slab i3 = w through both absorptions
  // closing operands: descriptor scalars and the first bra elements are requested before the absorption phases
  const cplx* const bt = d.bt + s + (long long)p1 * d.g[1] + (long long)p2 * d.g[2] + (long long)r * d.gw;
  const long long g0 = d.g[0], g3 = d.g[3], g4 = d.g[4];
  constexpr int H = N / 4, U = 4;
  const int steps = H * N * E4;
  auto bra = [=](int st) -> const cplx* { const int i1 = 4 * (st % H) + q, i0 = (st / H) % N, i2 = st / (H * N); return bt + i1 * g0 + i0 * g3 + i2 * g4; };
  cplx bcur[U], bnext[U];
#pragma unroll
  for (int u = 0; u < U; ++u) bcur[u] = warp + u * nwarps < steps ? __ldg(bra(warp + u * nwarps)) : cmake(0, 0);
  auto at = [sm](int i0, int i1, int i2, int i3) -> cplx* { return sm + ((i0 ^ slab_f(i1) ^ slab_f(i2)) + N * i1 + N * N * i2 + SB * i3); };
  __syncthreads();
  // p3: rows (p0', p4) straight from the tile [w][p4][p3][p0']
  if (!(TNQS_DEEP_ABLATE & 8)) deep_phase<N, TNQS_DEEP_DEPTH>(msm, 0, E4,
                [=](int rho, int b) -> const cplx* { return xt + rho % N + xs3 * b + xs4 * (rho / N); },
                [=](int rho, int c) -> cplx* { return at(c, rho % N, rho / N, warp); }, lane);
  __syncwarp();
  if (E4 == N && !(TNQS_DEEP_ABLATE & 8)) {
    // p4 in place: rows (i0, i1)
    deep_phase<N, TNQS_DEEP_DEPTH_SMEM>(msm + N * N, 0, N * N / 8,
                  [=](int rho, int b) -> const cplx* { return at(rho % N, rho / N, b, warp); },
                  [=](int rho, int c) -> cplx* { return at(rho % N, rho / N, c, warp); }, lane);
  }
  __syncthreads();
  // closing: G[w, w'] = sum_K X[K, w] conj(B[K, w']), K = (p0', p3', p4').  A k-step = four consecutive p0' at one (p3', p4'); lane (r, q)
  // feeds X[.., w = r] as the A fragment and B[.., w' = r] as the B fragment, and owns G[r][2q], G[r][2q + 1].
  // X conj(B) in 3M form: T1 = X_r B_r, T2 = X_i B_i, T3 = (X_r + X_i)(B_r - B_i); Re = T1 + T2, Im = T3 - T1 + T2.
  double t1[2] = {0, 0}, t2[2] = {0, 0}, t3[2] = {0, 0};
  auto ket = [=](int st) -> const cplx* { const int i1 = 4 * (st % H) + q, i0 = (st / H) % N, i2 = st / (H * N); return at(i0, i1, i2, r); };
  int st = warp;
  for (; st < ((TNQS_DEEP_ABLATE & 4) ? 0 : steps); st += U * nwarps) {
    const int sn = st + U * nwarps;
#pragma unroll
    for (int u = 0; u < U; ++u) bnext[u] = sn + u * nwarps < steps ? __ldg(bra(sn + u * nwarps)) : cmake(0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (st + u * nwarps < steps) {                 // warp-uniform
        const cplx xv = *ket(st + u * nwarps), bv = bcur[u];
        dmma884(t1[0], t1[1], xv.x, bv.x);
        dmma884(t2[0], t2[1], xv.y, bv.y);
        dmma884(t3[0], t3[1], xv.x + xv.y, bv.x - bv.y);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) bcur[u] = bnext[u];
  }
  // fixed-order sum over the warps, one partial per CTA
  __syncthreads();                                   // every warp has read its last X element: the slab becomes the reduction buffer
#pragma unroll
  for (int j = 0; j < 2; ++j) red[warp * N * N + r + N * (2 * q + j)] = cmake(t1[j] + t2[j], t3[j] - t1[j] + t2[j]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    cplx acc = red[idx];
    for (int wv = 1; wv < nwarps; ++wv) acc = cadd(acc, red[wv * N * N + idx]);
    d.partial[(size_t)y * N * N + idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_slab_batch: the warp-owned sub-slab machinery of k_deep_absorb with free input and output strides, for the gate path on
// high-coordination vertices (simple_update.jl:35-36: psi~ = psi x sqrt(M_k) over 4 or 5 outer legs of dimension 8).  A warp takes a
// sub-slab (i0, i1, i2) of N^3 elements from `in`, multiplies the legs that carry a matrix (leg 0 on the way into shared memory, leg 1
// in place, leg 2 on the way out; a leg without a matrix is copied) and stores it to `out`; warp = a fourth index i3, CTA = up to three
// more.  Two launches absorb five legs with 2 x (read + write) of the tensor instead of 5 x, and the first one reads the canonical
// site tensor and writes the matricised layout, so the separate permute pass goes as well.  in == out is allowed when both stride
// sets are equal (a warp reads its whole sub-slab before it writes).
// ---------------------------------------------------------------------------------------------------------------
struct BatchSlabDesc {
  const cplx* in; cplx* out;
  const cplx* m[3];            // N x N (ld = N) per slab leg, or null
  long long gi[7], go[7];      // strides (complex elements) of i0, i1, i2, i3 (warp), and of the three digits of the CTA index
  int n3;                      // extent of i3 (<= N)
  int nc[3];                   // extents of the CTA digits
  int flags;                   // bit0: conj(m), bit1: m^T (as k_mode_apply)
};

// grid: (max CTAs per item, items); block 32 N threads; dynamic shared memory DeepCfg<N>::SMEM_A
template <int N>
__global__ void __launch_bounds__(DEEP_THREADS, DEEP_CTAS) k_slab_batch(const BatchSlabDesc* __restrict__ descs) {
  static_assert(DEEP_THREADS == 32 * N, "one warp per sub-slab");
  const BatchSlabDesc& d = descs[blockIdx.y];
  const int nc0 = d.nc[0], nc1 = d.nc[1], nc2 = d.nc[2];
  if ((int)blockIdx.x >= nc0 * nc1 * nc2) return;
  extern __shared__ cplx dsm[];
  cplx* const msm = dsm + DeepCfg<N>::SLAB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const cplx *m0 = d.m[0], *m1 = d.m[1], *m2 = d.m[2];
  {
    const bool conj_m = d.flags & 1, trans_m = d.flags & 2;
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      const int from = trans_m ? idx / N + N * (idx % N) : idx;
      cplx v;
      if (m0) { v = __ldg(m0 + from); if (conj_m) v.y = -v.y; msm[idx] = v; }
      if (m1) { v = __ldg(m1 + from); if (conj_m) v.y = -v.y; msm[N * N + idx] = v; }
      if (m2) { v = __ldg(m2 + from); if (conj_m) v.y = -v.y; msm[2 * N * N + idx] = v; }
    }
  }
  const int x = blockIdx.x, c0 = x % nc0, c1 = (x / nc0) % nc1, c2 = x / (nc0 * nc1);
  const cplx* const src = d.in + warp * d.gi[3] + c0 * d.gi[4] + c1 * d.gi[5] + c2 * d.gi[6];
  cplx* const dst = d.out + warp * d.go[3] + c0 * d.go[4] + c1 * d.go[5] + c2 * d.go[6];
  const long long gi0 = d.gi[0], gi1 = d.gi[1], gi2 = d.gi[2], go0 = d.go[0], go1 = d.go[1], go2 = d.go[2];
  const int n3 = d.n3;
  cplx* const sub = dsm + N * N * N * warp;
  auto at = [sub](int i0, int i1, int i2) -> cplx* { return sub + ((i0 ^ slab_f(i1) ^ slab_f(i2)) + N * i1 + N * N * i2); };
  constexpr int NTILES = N * N / 8;
  __syncthreads();
  if (warp >= n3) return;
  if (m0) {
    deep_phase<N, TNQS_DEEP_DEPTH>(msm, 0, NTILES,
                  [=](int rho, int b) -> const cplx* { return src + (rho % N) * gi1 + (rho / N) * gi2 + b * gi0; },
                  [=](int rho, int c) -> cplx* { return at(c, rho % N, rho / N); }, lane);
  } else {
    // plain copy: lane = (i0, two bits of i1); N^3 / 32 elements per lane, all requested before the first one is stored
    constexpr int PER = N * N * N / 32;
    cplx tmp[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int e = lane + 32 * k; tmp[k] = src[(e % N) * gi0 + ((e / N) % N) * gi1 + (e / (N * N)) * gi2]; }
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int e = lane + 32 * k; *at(e % N, (e / N) % N, e / (N * N)) = tmp[k]; }
  }
  __syncwarp();
  if (m1) {
    deep_phase<N, TNQS_DEEP_DEPTH_SMEM>(msm + N * N, 0, NTILES,
                  [=](int rho, int b) -> const cplx* { return at(rho % N, b, rho / N); },
                  [=](int rho, int c) -> cplx* { return at(rho % N, c, rho / N); }, lane);
    __syncwarp();
  }
  if (m2) {
    deep_phase<N, TNQS_DEEP_DEPTH_SMEM>(msm + 2 * N * N, 0, NTILES,
                  [=](int rho, int b) -> const cplx* { return at(rho % N, rho / N, b); },
                  [=](int rho, int c) -> cplx* { return dst + (rho % N) * go0 + (rho / N) * go1 + c * go2; }, lane);
  } else {
    constexpr int PER = N * N * N / 32;
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int e = lane + 32 * k; dst[(e % N) * go0 + ((e / N) % N) * go1 + (e / (N * N)) * go2] = *at(e % N, (e / N) % N, e / (N * N)); }
  }
}
