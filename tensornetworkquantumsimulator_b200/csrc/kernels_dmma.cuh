// FP64 tensor-core (DMMA.8x8x4, `mma.sync.m8n8k4.f64`) fast path of the BP message update for uniform bond
// dimension N (8 or 16), site dimension 2 and 2..3 absorbed legs (degree 3..4 vertices)  — DESIGN.md "kernels".
//
//   k_absorb_slab : one CTA per (message, s, w) slab.  The slab T[s, :, :, :, w] (N^F complex) goes through the
//                   F leg absorptions back to back: leg 1 straight from global memory into a padded shared-memory
//                   slab, middle legs in place in shared memory, the last leg from shared memory straight to the
//                   slab-major intermediate X[w][s][...] in global memory.  A complex product is 4 real MMAs on the
//                   interleaved (re, im) data: an absorption is the real GEMM  rows x 2N  times  2N x 2N  with
//                   [[Re m, Im m], [-Im m, Re m]] interleaved, so operand fragments are plain 8-byte loads.
//   k_close_dmma  : M[w, w'] = sum_K X[K, w] conj(T[K, w'])  as the real GEMM  P = [X_re; X_im] [T_re, T_im]  (32 x 32 x N^F
//                   per s), split over K across CTAs, fixed-order reduction over warps, partials in the layout
//                   k_close_reduce consumes.
// There is no f64 kind of tcgen05, so DMMA is the FP64 tensor path on sm_100a (measured peak 37.15 TFLOP/s).
#pragma once
#include "common.cuh"

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct SlabDesc {
  const cplx* t;          // site tensor, canonical layout
  cplx* x;                // intermediate, slab-major: x[(w*2 + s) * N^F + i1 + N*i2 + N^2*i3]
  const cplx* m[3];       // messages of the absorbed legs, in absorption order (N x N, ld = N)
  int F;                  // number of absorbed legs (2 or 3)
  long long g[3];         // canonical strides (complex elements) of the absorbed legs
  long long gw;           // canonical stride of the outgoing leg
};

template <int N> struct SlabCfg {
  static constexpr int KS = N / 2;        // k-steps of 4 real = 2 complex
  static constexpr int NT = N / 4;        // n-tiles of 8 real = 4 complex
  static constexpr int S1 = N + 2;        // padded shared stride of the 2nd slab leg (== 2 mod 8: conflict-free fragments)
  static constexpr int SLAB3 = S1 * N * N;
};

// grid: (2*N, items); block 256 (8 warps); dynamic shared memory: S1*N^(F-1) complex (max SLAB3)
template <int N>
__global__ void __launch_bounds__(256, 2) k_absorb_slab(const SlabDesc* __restrict__ descs) {
  using Cfg = SlabCfg<N>;
  constexpr int KS = Cfg::KS, NT = Cfg::NT;
  const SlabDesc d = descs[blockIdx.y];
  extern __shared__ cplx slab[];
  const int s = blockIdx.x & 1, w = blockIdx.x >> 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int r = lane >> 2, q = lane & 3, kpair = q >> 1, part = q & 1;
  const int F = d.F;
  long long NF = N; for (int i = 1; i < F; ++i) NF *= N;
  const double* gsrc = reinterpret_cast<const double*>(d.t + s + (long long)w * d.gw);
  double* gdst = reinterpret_cast<double*>(d.x + ((long long)w * 2 + s) * NF);
  double* sm = reinterpret_cast<double*>(slab);
  const long long sst[3] = {1, Cfg::S1, (long long)Cfg::S1 * N};   // shared strides
  const long long xst[3] = {1, N, (long long)N * N};               // intermediate strides
  const int nrows = (int)(NF / N);                                 // rows of every absorption

  for (int p = 0; p < F; ++p) {
    // B fragments of [[Re m, Im m], [-Im m, Re m]] (interleaved), kept in registers for the whole phase
    double bf[KS][NT];
    {
      const cplx* m = d.m[p];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          int b = 2 * ks + kpair, c = 4 * nt + (r >> 1), pp = r & 1;
          cplx v = m[b + N * c];
          bf[ks][nt] = (part == pp) ? v.x : (part ? -v.y : v.y);
        }
    }
    const bool from_global = p == 0, to_global = p == F - 1;
    const int o1 = p == 0 ? 1 : 0, o2 = p == 2 ? 1 : 2;             // the other slab legs, fastest first
    const long long src_m = from_global ? d.g[p] : sst[p];
    const long long src_o1 = from_global ? d.g[o1] : sst[o1];
    const long long src_o2 = (F == 3) ? (from_global ? d.g[o2] : sst[o2]) : 0;
    const long long dst_m = to_global ? xst[p] : sst[p];
    const long long dst_o1 = to_global ? xst[o1] : sst[o1];
    const long long dst_o2 = (F == 3) ? (to_global ? xst[o2] : sst[o2]) : 0;
    const double* src = from_global ? gsrc : sm;
    double* dst = to_global ? gdst : sm;
    for (int ti = warp; ti < nrows / 8; ti += nwarps) {
      const int rho = 8 * ti + r;
      const int i1 = rho % N, i2 = rho / N;
      const long long soff = i1 * src_o1 + i2 * src_o2;
      const long long doff = i1 * dst_o1 + i2 * dst_o2;
      double a[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) a[ks] = src[2 * (soff + (2 * ks + kpair) * src_m) + part];
      double acc[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[nt][0], acc[nt][1], a[ks], bf[ks][nt]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        double2 v; v.x = acc[nt][0]; v.y = acc[nt][1];
        *reinterpret_cast<double2*>(dst + 2 * (doff + (q + 4 * nt) * dst_m)) = v;
      }
    }
    __syncthreads();
  }
}

struct CloseDmmaDesc {
  const cplx* x;          // intermediate from k_absorb_slab (slab-major)
  const cplx* t;          // site tensor, canonical layout
  cplx* partial;          // partial[split][N*N] complex, column-major (w, w')
  int F;
  long long g[3];         // canonical strides of the absorbed legs
  long long gw;           // canonical stride of the outgoing leg
  int nsplit;
};

// grid: (max nsplit, items); block 256 (8 warps); dynamic shared memory: 8 * (2N)^2 doubles
template <int N>
__global__ void __launch_bounds__(256, 2) k_close_dmma(const CloseDmmaDesc* __restrict__ descs) {
  constexpr int MT = N / 4;               // tiles of 8 real rows = 4 w x (re, im)
  const CloseDmmaDesc d = descs[blockIdx.y];
  if ((int)blockIdx.x >= d.nsplit) return;
  extern __shared__ double red[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int s = q & 1, di = q >> 1, wl = r >> 1, part = r & 1;
  long long NF = N; for (int i = 1; i < d.F; ++i) NF *= N;
  const long long ksteps = NF / 2;
  const long long per = (ksteps + d.nsplit - 1) / d.nsplit;
  const long long k_begin = (long long)blockIdx.x * per, k_end = min(ksteps, k_begin + per);
  const double* X = reinterpret_cast<const double*>(d.x);
  const double* T = reinterpret_cast<const double*>(d.t);
  double acc[MT][MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < MT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  for (long long kap = k_begin + warp; kap < k_end; kap += 8) {
    const int j = (int)(kap % (N / 2));
    const long long rest = kap / (N / 2);
    const int i2 = (int)(rest % N), i3 = (int)(rest / N);
    const int i1 = 2 * j + di;
    const long long sigma = i1 + (long long)N * i2 + (long long)N * N * i3;
    const long long toff = s + i1 * d.g[0] + i2 * d.g[1] + (d.F == 3 ? i3 * d.g[2] : 0);
    double a[MT], b[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int w = 4 * mt + wl;
      a[mt] = X[2 * (((long long)w * 2 + s) * NF + sigma) + part];
      b[mt] = T[2 * (toff + (long long)w * d.gw) + part];
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < MT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
  }
  // fixed-order reduction over the 8 warps
  constexpr int PD = 2 * N;                        // P is PD x PD real
  double* mine = red + (size_t)warp * PD * PD;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < MT; ++nt) {
      int row = 8 * mt + r, col = 8 * nt + 2 * q;
      mine[row + PD * col] = acc[mt][nt][0];
      mine[row + PD * (col + 1)] = acc[mt][nt][1];
    }
  __syncthreads();
  cplx* out = d.partial + (size_t)blockIdx.x * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int w = idx % N, wp = idx / N;
    double rr = 0, ii = 0, ir = 0, ri = 0;
    for (int wv = 0; wv < 8; ++wv) {
      const double* P = red + (size_t)wv * PD * PD;
      rr += P[(2 * w) + PD * (2 * wp)];
      ii += P[(2 * w + 1) + PD * (2 * wp + 1)];
      ir += P[(2 * w + 1) + PD * (2 * wp)];
      ri += P[(2 * w) + PD * (2 * wp + 1)];
    }
    out[idx] = cmake(rr + ii, ir - ri);              // X conj(T): re = rr + ii, im = ir - ri
  }
}
