// Small dense factorisations, one CTA per matrix (DESIGN.md "kernels"):
//   k_eigh_sqrt  : Hermitian eigendecomposition by one-sided Jacobi -> sqrt(M), M^(-1/2)   (src/utils.jl:18-26)
//   k_house_qr   : Householder QR of a row block held in shared memory, explicit thin Q     (TSQR building block)
//   k_theta      : theta = R1 * R2 over the gate bond, then the two-site gate               (simple_update.jl:43)
//   k_svd_trunc  : one-sided Jacobi SVD + ITensors-style truncation + factor outputs         (simple_update.jl:46-61)
// All reductions have a fixed order (warp butterflies, sequential tail sums): results do not depend on scheduling.
#pragma once
#include "common.cuh"

#define TNQS_EPS 2.220446049250313e-16
#define TNQS_JACOBI_MAX_SWEEPS 60

enum { ST_NEG_EIG = 1, ST_NO_CONV = 2, ST_RANK_CAP = 4, ST_HALO_TIMEOUT = 8 };

// Round-robin (circle method) pairing: N even players, round s in [0, N-1), pair i in [0, N/2).
__device__ __forceinline__ void rr_pair(int N, int s, int i, int& p, int& q) {
  if (i == 0) { p = s % (N - 1); q = N - 1; }
  else { p = (s + i) % (N - 1); q = (s - i + 2 * (N - 1)) % (N - 1); }
  if (p > q) { int t = p; p = q; q = t; }
}

// One-sided Jacobi on the columns of W (m x n, leading dim ldw); the same rotations are accumulated into the
// rows of V (nv x n, leading dim ldv) unless V is null.  On exit the columns of W are mutually orthogonal.
// Every thread of the CTA must call.  Returns the number of sweeps used, or -1 if not converged.
// A column pair belongs to a HALF-warp (16 lanes x RPH rows per lane, m, nv <= 16 * RPH): one warp instruction serves two
// pairs, so the butterfly reductions (4 rounds instead of 5) and the rotation parameters cost half as many shuffle and
// FP64 issue slots per pair, and a CTA of 16 warps covers the 32 pairs of a 64-column step in one pass.
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int RPH>
__device__ int jacobi_columns_t(cplx* W, int ldw, int m, cplx* V, int ldv, int nv, int n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int hw = lane >> 4, sl = lane & 15;
  const int N = n + (n & 1);
  if (n < 2) return 0;
  // rotate whenever |<a_p,a_q>| > sqrt(m) eps |a_p||a_q| (LAPACK zgesvj's threshold), but call the iteration
  // converged once a whole sweep saw nothing above 8x that: rotations at the rounding-noise level never end.
  const double tol2 = (double)m * TNQS_EPS * TNQS_EPS;
  const double conv2 = 64.0 * tol2;
  // Columns whose norm has fallen below eps * sqrt(m) * ||W||_F are pure rounding noise (the null space of a
  // rank-deficient matrix, e.g. theta of a low-Schmidt-rank gate): rotating them never settles, so they are left alone.
  __shared__ double jred[32];
  double fro2 = 0;
  for (int idx = threadIdx.x; idx < m * n; idx += blockDim.x) fro2 += cabs2(W[(idx % m) + (size_t)(idx / m) * ldw]);
  fro2 = block_sum(fro2, jred);
  const double tiny2 = tol2 * fro2;
  for (int sweep = 0; sweep < TNQS_JACOBI_MAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int s = 0; s < N - 1; ++s) {
      for (int i0 = 2 * warp; i0 < N / 2; i0 += 2 * nwarps) {
        const int i = i0 + hw;
        int p = 0, q = 0;
        bool act = i < N / 2;
        if (act) { rr_pair(N, s, i, p, q); act = q < n; }
        cplx* wp = W + (size_t)p * ldw;
        cplx* wq = W + (size_t)q * ldw;
        constexpr bool KEEP = RPH <= 4;                      // columns stay in registers between the inner products and the rotation
        cplx a[KEEP ? RPH : 1], b[KEEP ? RPH : 1];
        double al = 0, be = 0, gr = 0, gi = 0;
#pragma unroll
        for (int t = 0; t < RPH; ++t) {
          const int r = sl + 16 * t;
          const bool ok = act && r < m;
          const cplx x = ok ? wp[r] : cmake(0, 0), y = ok ? wq[r] : cmake(0, 0);
          if (KEEP) { a[t] = x; b[t] = y; }
          al += cabs2(x); be += cabs2(y);
          gr += x.x * y.x + x.y * y.y;         // conj(a) * b
          gi += x.x * y.y - x.y * y.x;
        }
        al = half_sum(al); be = half_sum(be); gr = half_sum(gr); gi = half_sum(gi);
        const double g2 = gr * gr + gi * gi;
        if (!act || !(g2 > tol2 * al * be) || g2 == 0.0 || al <= tiny2 || be <= tiny2) continue;    // uniform per half-warp; no shuffles below
        if (g2 > conv2 * al * be) rotated = 1;
        // rotation parameters with reciprocal-square-root / reciprocal intrinsics: the chain of FP64 sqrt and divide
        // subroutines (about 1000 dependent cycles) was the critical path of a Jacobi step
        const double rg = rsqrt(g2);                              // 1 / |gamma|
        const cplx ph = cmake(gr * rg, -gi * rg);                 // e^{-i phi}: makes <a_p, ph*a_q> real positive
        const double zeta = 0.5 * (be - al) * rg;
        const double z2 = fma(zeta, zeta, 1.0);
        const double t = copysign(__drcp_rn(fabs(zeta) + z2 * rsqrt(z2)), zeta);
        const double c = rsqrt(fma(t, t, 1.0)), sn = c * t;
#pragma unroll
        for (int u = 0; u < RPH; ++u) {
          const int r = sl + 16 * u;
          if (r < m) {
            cplx x = KEEP ? a[u] : wp[r], y = cmul(ph, KEEP ? b[u] : wq[r]);
            wp[r] = cmake(c * x.x - sn * y.x, c * x.y - sn * y.y);
            wq[r] = cmake(sn * x.x + c * y.x, sn * x.y + c * y.y);
          }
        }
        if (V) {
          cplx* vp = V + (size_t)p * ldv;
          cplx* vq = V + (size_t)q * ldv;
#pragma unroll
          for (int u = 0; u < RPH; ++u) {
            const int r = sl + 16 * u;
            if (r < nv) {
              cplx x = vp[r], y = cmul(ph, vq[r]);
              vp[r] = cmake(c * x.x - sn * y.x, c * x.y - sn * y.y);
              vq[r] = cmake(sn * x.x + c * y.x, sn * x.y + c * y.y);
            }
          }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) return sweep + 1;
  }
  return -1;
}

__device__ int jacobi_columns(cplx* W, int ldw, int m, cplx* V, int ldv, int nv, int n) {
  const int mx = m > nv ? m : nv;
  if (mx <= 16) return jacobi_columns_t<1>(W, ldw, m, V, ldv, nv, n);
  if (mx <= 32) return jacobi_columns_t<2>(W, ldw, m, V, ldv, nv, n);
  if (mx <= 64) return jacobi_columns_t<4>(W, ldw, m, V, ldv, nv, n);
  if (mx <= 128) return jacobi_columns_t<8>(W, ldw, m, V, ldv, nv, n);
  // theta' of the gate path is (q1 d) x (q2 d) with q <= d chi <= 64: up to 256 rows on d = 4 (Pauli) sites, 192 on d = 3;
  // the host refuses anything longer (GateBatch::stage2), so no row is ever left out of the inner products
  return jacobi_columns_t<16>(W, ldw, m, V, ldv, nv, n);
}
constexpr int TNQS_JACOBI_MAX_ROWS = 256;

// ---------------------------------------------------------------------------------------------------------------
struct EighDesc {
  const cplx* m; int n; int ldm;
  cplx* root; cplx* invroot;      // n x n each (ld = n); either may be null
  double cutoff;                  // mode 0: f(lambda) = 0 if lambda == 0 or |lambda| < cutoff
  double shift;                   // mode 1: lambda <- lambda + shift, no cutoff (symmetric_gauge.jl:14-15)
  int mode;
  int* cutflag = nullptr;         // mode 0: set to 1 when an eigenvalue fell under the cutoff (then sqrt(M) M^(-1/2) is a projector, not 1)
};

// dynamic shared memory: W[n*n] V[n*n] complex, lam[n] fr[n] fi[n] doubles
__global__ void __launch_bounds__(1024) k_eigh_sqrt(const EighDesc* __restrict__ descs, int* status) {
  const EighDesc d = descs[blockIdx.x];
  extern __shared__ cplx esm[];
  const int n = d.n;
  cplx* W = esm;
  cplx* V = esm + (size_t)n * n;
  double* lam = (double*)(V + (size_t)n * n);
  double* fr = lam + n;
  double* fi = fr + n;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    int i = idx % n, j = idx / n;
    // Hermitian part: LAPACK's heev reads one triangle only; (M + M^H)/2 is the symmetric equivalent
    cplx a = d.m[i + (size_t)j * d.ldm], b = d.m[j + (size_t)i * d.ldm];
    W[idx] = cmake(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
    V[idx] = cmake(i == j ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  int sweeps = jacobi_columns(W, n, n, V, n, n, n);
  if (sweeps < 0 && threadIdx.x == 0) atomicOr(status, ST_NO_CONV);
  // W = M V  => lambda_j = Re(v_j^H w_j)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < n; j += nwarps) {
    double s = 0;
    for (int r = lane; r < n; r += 32) { cplx v = V[r + (size_t)j * n], w = W[r + (size_t)j * n]; s += v.x * w.x + v.y * w.y; }
    s = warp_sum(s);
    if (lane == 0) {
      double f = 0, g = 0;
      if (d.mode == 0) {
        if (!(s == 0.0 || fabs(s) < d.cutoff)) {
          if (s < 0) atomicOr(status, ST_NEG_EIG); else { f = sqrt(s); g = 1.0 / f; }
        } else if (d.cutflag) atomicOr(d.cutflag, 1);
      } else {
        double l = s + d.shift;
        if (l < 0) atomicOr(status, ST_NEG_EIG); else { f = sqrt(l); g = 1.0 / f; }
      }
      lam[j] = s; fr[j] = f; fi[j] = g;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    int b = idx % n, c = idx / n;
    cplx sr = cmake(0, 0), si = cmake(0, 0);
    for (int j = 0; j < n; ++j) {
      cplx vb = V[b + (size_t)j * n], vc = V[c + (size_t)j * n];
      cplx pr = cmake(vb.x * vc.x + vb.y * vc.y, vb.y * vc.x - vb.x * vc.y);   // vb * conj(vc)
      sr.x = fma(fr[j], pr.x, sr.x); sr.y = fma(fr[j], pr.y, sr.y);
      si.x = fma(fi[j], pr.x, si.x); si.y = fma(fi[j], pr.y, si.y);
    }
    if (d.root) d.root[idx] = sr;
    if (d.invroot) d.invroot[idx] = si;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Householder QR of an h x nc block (h >= nc) held in shared memory, overwritten by its explicit thin Q (zgeqr2 + zung2r
// arithmetic, R is not kept).  Every thread of the CTA must call; tau: nc complex, scal: 1 complex of shared memory.
__device__ void house_q_inplace(cplx* A, int h, int nc, cplx* tau, cplx* scal) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int j = 0; j < nc; ++j) {
    cplx* col = A + (size_t)j * h;
    if (warp == 0) {                                      // zlarfg
      double xn2 = 0;
      for (int r = j + 1 + lane; r < h; r += 32) xn2 += cabs2(col[r]);
      xn2 = warp_sum(xn2);
      if (lane == 0) {
        cplx alpha = col[j];
        if (xn2 == 0.0 && alpha.y == 0.0) { tau[j] = cmake(0, 0); scal[0] = cmake(0, 0); }
        else {
          double beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
          tau[j] = cmake((beta - alpha.x) / beta, -alpha.y / beta);
          double dr = alpha.x - beta, di = alpha.y, den = dr * dr + di * di;
          scal[0] = cmake(dr / den, -di / den);
          col[j] = cmake(beta, 0);
        }
      }
    }
    __syncthreads();
    {
      cplx sc = scal[0];
      for (int r = j + 1 + tid; r < h; r += blockDim.x) col[r] = cmul(col[r], sc);
    }
    __syncthreads();
    cplx tj = tau[j];
    if (tj.x != 0.0 || tj.y != 0.0) {                     // H(j)^H = I - conj(tau) v v^H on the trailing columns
      for (int c = j + 1 + warp; c < nc; c += nwarps) {
        cplx* cc = A + (size_t)c * h;
        cplx w = cmake(0, 0);
        for (int r = j + 1 + lane; r < h; r += 32) cfma_ca(w, col[r], cc[r]);
        w.x = warp_sum(w.x); w.y = warp_sum(w.y);
        w = cadd(w, cc[j]);                                // v_j = 1
        cplx f = cmul(cconj(tj), w);
        if (lane == 0) cc[j] = csub(cc[j], f);
        for (int r = j + 1 + lane; r < h; r += 32) cc[r] = csub(cc[r], cmul(f, col[r]));
      }
    }
    __syncthreads();
  }
  for (int i = nc - 1; i >= 0; --i) {                      // explicit Q (zung2r)
    cplx* col = A + (size_t)i * h;
    cplx ti = tau[i];
    if (ti.x != 0.0 || ti.y != 0.0) {
      for (int c = i + 1 + warp; c < nc; c += nwarps) {
        cplx* cc = A + (size_t)c * h;
        cplx w = cmake(0, 0);
        for (int r = i + 1 + lane; r < h; r += 32) cfma_ca(w, col[r], cc[r]);
        w.x = warp_sum(w.x); w.y = warp_sum(w.y);
        w = cadd(w, cc[i]);
        cplx f = cmul(ti, w);
        if (lane == 0) cc[i] = csub(cc[i], f);
        for (int r = i + 1 + lane; r < h; r += 32) cc[r] = csub(cc[r], cmul(f, col[r]));
      }
    }
    __syncthreads();
    for (int r = tid; r < h; r += blockDim.x) {
      cplx v;
      if (r < i) v = cmake(0, 0);
      else if (r == i) v = cmake(1.0 - ti.x, -ti.y);
      else v = cmul(cmake(-ti.x, -ti.y), col[r]);
      col[r] = v;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct QrDesc {
  cplx* a; long long lda; int h, nc;     // block: h rows x nc columns, column-major
  cplx* rout; int ldr;                   // R: min(h,nc) x nc, zeros below the diagonal
};

// Householder QR (LAPACK zgeqr2 + zung2r arithmetic) of one block in shared memory; block 256 threads.
// dynamic shared memory: A[h*nc] complex + tau[nc] complex + 2 complex scratch
__global__ void __launch_bounds__(256) k_house_qr(const QrDesc* __restrict__ descs) {
  const QrDesc d = descs[blockIdx.x];
  extern __shared__ cplx qsm[];
  const int h = d.h, nc = d.nc, q = min(h, nc);
  cplx* A = qsm;
  cplx* tau = qsm + (size_t)h * nc;
  cplx* scal = tau + nc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int idx = tid; idx < h * nc; idx += blockDim.x) {
    int r = idx % h, c = idx / h;
    A[idx] = d.a[r + (size_t)c * d.lda];
  }
  __syncthreads();
  for (int j = 0; j < q; ++j) {
    cplx* col = A + (size_t)j * h;
    if (warp == 0) {                                      // zlarfg
      double xn2 = 0;
      for (int r = j + 1 + lane; r < h; r += 32) xn2 += cabs2(col[r]);
      xn2 = warp_sum(xn2);
      if (lane == 0) {
        cplx alpha = col[j];
        if (xn2 == 0.0 && alpha.y == 0.0) { tau[j] = cmake(0, 0); scal[0] = cmake(0, 0); }
        else {
          double beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2), alpha.x);
          tau[j] = cmake((beta - alpha.x) / beta, -alpha.y / beta);
          double dr = alpha.x - beta, di = alpha.y, den = dr * dr + di * di;
          scal[0] = cmake(dr / den, -di / den);
          col[j] = cmake(beta, 0);
        }
      }
    }
    __syncthreads();
    {
      cplx sc = scal[0];
      for (int r = j + 1 + tid; r < h; r += blockDim.x) col[r] = cmul(col[r], sc);
    }
    __syncthreads();
    cplx tj = tau[j];
    if (tj.x != 0.0 || tj.y != 0.0) {                     // apply H(j)^H = I - conj(tau) v v^H to the trailing columns
      for (int c = j + 1 + warp; c < nc; c += nwarps) {
        cplx* cc = A + (size_t)c * h;
        cplx w = cmake(0, 0);
        for (int r = j + 1 + lane; r < h; r += 32) cfma_ca(w, col[r], cc[r]);
        w.x = warp_sum(w.x); w.y = warp_sum(w.y);
        w = cadd(w, cc[j]);                                // v_j = 1
        cplx f = cmul(cconj(tj), w);
        if (lane == 0) cc[j] = csub(cc[j], f);
        for (int r = j + 1 + lane; r < h; r += 32) cc[r] = csub(cc[r], cmul(f, col[r]));
      }
    }
    __syncthreads();
  }
  // R out
  for (int idx = tid; idx < q * nc; idx += blockDim.x) {
    int i = idx % q, c = idx / q;
    d.rout[i + (size_t)c * d.ldr] = (i <= c) ? A[i + (size_t)c * h] : cmake(0, 0);
  }
  __syncthreads();
  // explicit Q (h x q) in place (zung2r)
  for (int i = q - 1; i >= 0; --i) {
    cplx* col = A + (size_t)i * h;
    cplx ti = tau[i];
    if (ti.x != 0.0 || ti.y != 0.0) {
      for (int c = i + 1 + warp; c < q; c += nwarps) {
        cplx* cc = A + (size_t)c * h;
        cplx w = cmake(0, 0);
        for (int r = i + 1 + lane; r < h; r += 32) cfma_ca(w, col[r], cc[r]);
        w.x = warp_sum(w.x); w.y = warp_sum(w.y);
        w = cadd(w, cc[i]);
        cplx f = cmul(ti, w);
        if (lane == 0) cc[i] = csub(cc[i], f);
        for (int r = i + 1 + lane; r < h; r += 32) cc[r] = csub(cc[r], cmul(f, col[r]));
      }
    }
    __syncthreads();
    for (int r = tid; r < h; r += blockDim.x) {
      cplx v;
      if (r < i) v = cmake(0, 0);
      else if (r == i) v = cmake(1.0 - ti.x, -ti.y);
      else v = cmul(cmake(-ti.x, -ti.y), col[r]);
      col[r] = v;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < h * q; idx += blockDim.x) {
    int r = idx % h, c = idx / h;
    d.a[r + (size_t)c * d.lda] = A[idx];
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct ThetaDesc {
  const cplx* r1; int q1, ld1;           // R1[q1, s1 + d*b]
  const cplx* r2; int q2, ld2;           // R2[q2, s2 + d*b]
  const cplx* gate;                      // o[s1', s2', s1, s2], column-major d^4
  int d, chi;
  cplx* theta;                           // theta'[(q1 + q1n*s1'), (q2 + q2n*s2')], ld = q1*d
};

// one CTA per gate; thread per (q1, q2) computes the d x d block over the bond, then mixes it with the gate.
__global__ void __launch_bounds__(256) k_theta(const ThetaDesc* __restrict__ descs) {
  const ThetaDesc t = descs[blockIdx.x];
  const int d = t.d;
  const int m = t.q1 * d;
  for (int idx = threadIdx.x; idx < t.q1 * t.q2; idx += blockDim.x) {
    int a = idx % t.q1, c = idx / t.q1;
    cplx th[16];                                          // d <= 4
    for (int i = 0; i < d * d; ++i) th[i] = cmake(0, 0);
    for (int b = 0; b < t.chi; ++b)
      for (int s2 = 0; s2 < d; ++s2) {
        cplx y = t.r2[c + (size_t)(s2 + d * b) * t.ld2];
        for (int s1 = 0; s1 < d; ++s1) cfma(th[s1 + d * s2], t.r1[a + (size_t)(s1 + d * b) * t.ld1], y);
      }
    for (int s2p = 0; s2p < d; ++s2p)
      for (int s1p = 0; s1p < d; ++s1p) {
        cplx acc = cmake(0, 0);
        for (int s2 = 0; s2 < d; ++s2)
          for (int s1 = 0; s1 < d; ++s1)
            cfma(acc, t.gate[s1p + d * (s2p + d * (s1 + d * s2))], th[s1 + d * s2]);
        t.theta[(a + t.q1 * s1p) + (size_t)(c + t.q2 * s2p) * m] = acc;
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct SvdDesc {
  const cplx* theta; int m, n;           // m x n column-major, ld = m
  cplx* work;                            // global scratch of (max(m,n)*min(m,n) + min(m,n)^2) complex, or null -> shared memory
  cplx* r1out;                           // U sqrt(S): m x rank, ld = m
  cplx* r2out;                           // [i, u] = sqrt(S_u) * conj(V[i, u]): n x rank, ld = n
  double* svals;                         // [cap]
  int* rank; double* truncerr;
  cplx* msg_a; cplx* msg_b;              // message slots receiving diag(S) (rank x rank), may be null
  int maxdim; double cutoff; int has_cutoff; int normalize; int cap; int mindim;
  // Row-space reduction (optional, gate path): theta' = R1 (gate) R2 has rank <= r_o * chi with r_o the operator-Schmidt rank of
  // the gate over the (site 1 | site 2) cut - 2 for Rzz, Rxx, CZ, CNOT.  When that is at most n / 2 the kernel iterates
  // W = theta' conj(Q) (m x r_o chi) instead of theta' (m x n), Q = an orthonormal basis of the span of the vectors
  // sum_s2 B_k[s2', s2] R2[q2, s2, b] that contains every row of theta' (B_k = a basis of the row space of the gate as a
  // d^2 x d^2 matrix, found on the device so that a replayed graph may carry other gate angles).  Same singular values and left
  // vectors as theta'; the right factor comes from theta'^H W_j as before.  rs_gate == null: off.
  const cplx* rs_gate = nullptr; const cplx* rs_r2 = nullptr; int rs_q2 = 0, rs_ld2 = 0, rs_d = 0, rs_chi = 0;
};

// dynamic shared memory: W (mm x nn) if work == null, then sigma[nn] doubles, perm[nn] ints.
// Only W = theta' * V_jacobi is iterated (half the shared-memory traffic and footprint, two CTAs per SM); the right
// factor of the KEPT singular triplets is recovered afterwards from theta'^H W_j = sigma_j^2 v_j, whose relative error
// eps * sigma_max / sigma_j stays below 1e-11 for anything a relative cutoff of 1e-10 on sigma^2 can keep.
__global__ void __launch_bounds__(512, 2) k_svd_trunc(const SvdDesc* __restrict__ descs, int* status, size_t smem_wv_bytes) {
  const SvdDesc d = descs[blockIdx.x];
  extern __shared__ cplx ssm[];
  const int m = d.m, n = d.n;
  const bool swapped = m < n;
  const int mm = swapped ? n : m;
  int nn = swapped ? m : n;
  cplx* W = d.work ? d.work : ssm;
  double* sigma = (double*)((char*)ssm + smem_wv_bytes);
  __shared__ int sh_rank;
  __shared__ double sh_norm, sh_fro2;
  // ---- row-space reduction (see SvdDesc) ----
  __shared__ cplx rs_rows[256], rs_basis[256];
  __shared__ int rs_ro;
  bool reduced = false;
  if (d.rs_gate && !swapped) {
    const int dd = d.rs_d, d2 = dd * dd;
    // rows e1 = (s1', s1), columns e2 = (s2', s2) of the gate o[s1', s2', s1, s2]
    for (int idx = threadIdx.x; idx < d2 * d2; idx += blockDim.x) {
      const int e1 = idx / d2, e2 = idx - e1 * d2;
      const int s1p = e1 % dd, s1 = e1 / dd, s2p = e2 % dd, s2 = e2 / dd;
      rs_rows[idx] = d.rs_gate[s1p + dd * (s2p + dd * (s1 + dd * s2))];
    }
    __syncthreads();
    if (threadIdx.x == 0) {                     // pivoted modified Gram-Schmidt over at most 16 rows of 16 entries
      double nrm[16], maxn2 = 0;
      for (int e = 0; e < d2; ++e) { double s = 0; for (int k = 0; k < d2; ++k) s += cabs2(rs_rows[e * d2 + k]); nrm[e] = s; maxn2 = fmax(maxn2, s); }
      int ro = 0;
      for (int it = 0; it < d2; ++it) {
        int best = 0; double bn = -1;
        for (int e = 0; e < d2; ++e) if (nrm[e] > bn) { bn = nrm[e]; best = e; }
        if (!(bn > 1e-26 * maxn2) || bn <= 0.0) break;
        const double inv = 1.0 / sqrt(bn);
        for (int k = 0; k < d2; ++k) rs_basis[ro * d2 + k] = cscale(rs_rows[best * d2 + k], inv);
        for (int e = 0; e < d2; ++e) {
          if (e == best) { nrm[e] = -1; continue; }
          if (nrm[e] < 0) continue;
          cplx p = cmake(0, 0);                 // <basis, row_e>
          for (int k = 0; k < d2; ++k) cfma_ca(p, rs_basis[ro * d2 + k], rs_rows[e * d2 + k]);
          double s = 0;
          for (int k = 0; k < d2; ++k) { rs_rows[e * d2 + k] = csub(rs_rows[e * d2 + k], cmul(p, rs_basis[ro * d2 + k])); s += cabs2(rs_rows[e * d2 + k]); }
          nrm[e] = s;
        }
        ++ro;
      }
      rs_ro = ro;
    }
    __syncthreads();
    const int ro = rs_ro, rho = ro * d.rs_chi;
    reduced = ro > 0 && 2 * rho <= n && (d.has_cutoff || (d.maxdim > 0 && d.maxdim <= rho)) && d.mindim <= rho;      // uniform over the CTA
    // theta' too large for shared memory (d.work != null): the reduced matrix may still fit - Q is built in shared memory, W goes to
    // the global work buffer first and moves into shared memory once Q is dead (the host sized the buffer for the rank of the gates
    // it saw; another rank at a graph replay falls back to the full iteration in global memory)
    if (d.work && (size_t)m * rho * sizeof(cplx) > smem_wv_bytes) reduced = false;
    if (reduced) {
      cplx* Q = ssm;                             // n x rho, then W = m x rho behind it: (n + m) rho <= m n
      W = d.work ? d.work : ssm + (size_t)n * rho;
      cplx* tau = (cplx*)sigma;                  // rho complex + 1: inside the sigma / perm area, free until the iteration is over
      for (int idx = threadIdx.x; idx < n * rho; idx += blockDim.x) {
        const int row = idx % n, col = idx / n;
        const int cq = row % d.rs_q2, s2p = row / d.rs_q2, k = col % ro, b = col / ro;
        cplx v = cmake(0, 0);
        for (int s2 = 0; s2 < dd; ++s2) cfma(v, rs_basis[k * d2 + s2p + dd * s2], d.rs_r2[cq + (size_t)(s2 + dd * b) * d.rs_ld2]);
        Q[idx] = v;
      }
      __syncthreads();
      house_q_inplace(Q, n, rho, tau, tau + rho);
      for (int idx = threadIdx.x; idx < m * rho; idx += blockDim.x) {
        const int i = idx % m, j = idx / m;
        cplx v = cmake(0, 0);
        for (int cc = 0; cc < n; ++cc) cfma_cb(v, d.theta[i + (size_t)cc * m], Q[cc + (size_t)j * n]);      // theta' conj(Q)
        W[idx] = v;
      }
      if (d.work) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < m * rho; idx += blockDim.x) ssm[idx] = d.work[idx];
        W = ssm;
      }
      nn = rho;
    }
  }
  int* perm = (int*)(sigma + nn);
  if (!reduced) {
    for (int idx = threadIdx.x; idx < mm * nn; idx += blockDim.x) {
      int r = idx % mm, c = idx / mm;
      cplx v;
      if (!swapped) v = d.theta[r + (size_t)c * m];
      else v = cconj(d.theta[c + (size_t)r * m]);        // W = theta'^H
      W[idx] = v;
    }
  }
  __syncthreads();
  int sweeps = jacobi_columns(W, mm, mm, nullptr, 0, 0, nn);
  if (sweeps < 0 && threadIdx.x == 0) atomicOr(status, ST_NO_CONV);
#ifdef TNQS_TRACE_SVD
  if (threadIdx.x == 0 && blockIdx.x < 4) printf("[svd] block %d m=%d n=%d sweeps=%d\n", blockIdx.x, m, n, sweeps);
#endif
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < nn; j += nwarps) {
    double s = 0;
    for (int r = lane; r < mm; r += 32) s += cabs2(W[r + (size_t)j * mm]);
    s = warp_sum(s);
    if (lane == 0) sigma[j] = sqrt(s);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nn; j += blockDim.x) {   // descending order, ties by index
    double sj = sigma[j];
    int rk = 0;
    for (int i = 0; i < nn; ++i) { double si = sigma[i]; rk += (si > sj) || (si == sj && i < j); }
    perm[rk] = j;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // NDTensors truncate!: P = S^2 descending, sequential tail sums
    int k = nn;
    double t = 0.0, err = 0.0, fro2 = 0.0;
    for (int i = 0; i < nn; ++i) { double s = sigma[perm[i]]; fro2 += s * s; }
    if (k > 1) {
      int maxdim = d.maxdim > 0 ? min(d.maxdim, k) : k;
      while (k > maxdim) { double s = sigma[perm[k - 1]]; t += s * s; --k; }
      double scale = fro2;
      if (scale == 0.0) scale = 1.0;
      if (d.has_cutoff) {
        while (k > d.mindim) {
          double s = sigma[perm[k - 1]];
          if (!(t + s * s <= d.cutoff * scale)) break;
          t += s * s; --k;
        }
      }
      err = t / scale;
    }
    if (k > d.cap) { atomicOr(status, ST_RANK_CAP); k = d.cap; }
    double nrm = 0.0;
    for (int i = 0; i < k; ++i) { double s = sigma[perm[i]]; nrm += s * s; }
    nrm = sqrt(nrm);
    sh_rank = k; sh_norm = (d.normalize && nrm > 0.0) ? 1.0 / nrm : 1.0; sh_fro2 = fro2;
    if (d.rank) *d.rank = k;
    if (d.truncerr) *d.truncerr = err;
  }
  __syncthreads();
  const int rank = sh_rank;
  const double sn = sh_norm;
  const double tiny = sqrt((double)mm) * TNQS_EPS * sqrt(sh_fro2);      // columns below this are rounding noise
  for (int u = threadIdx.x; u < d.cap; u += blockDim.x)
    if (d.svals) d.svals[u] = u < rank ? sigma[perm[u]] * sn : 0.0;
  for (int idx = threadIdx.x; idx < rank * rank; idx += blockDim.x) {
    int i = idx % rank, j = idx / rank;
    cplx v = cmake(i == j ? sigma[perm[i]] * sn : 0.0, 0.0);
    if (d.msg_a) d.msg_a[idx] = v;
    if (d.msg_b) d.msg_b[idx] = v;
  }
  // factors.  Left-hand factor of the iterated matrix: W_j / sqrt(sigma_j); the other one from theta' and W_j.
  //   not swapped (W = theta' V):  r1 = U sqrt(S) = W_j / sqrt(s);  r2[i,u] = sqrt(s) conj(V[i,j]) = (sum_k theta'[k,i] conj(W[k,j])) / s^1.5
  //   swapped (W = theta'^H U):    r2[i,u] = conj(W[i,j]) / sqrt(s); r1[i,u] = sqrt(s) U[i,j]   = (sum_k theta'[i,k] W[k,j]) / s^1.5
  for (int idx = threadIdx.x; idx < m * rank; idx += blockDim.x) {
    int i = idx % m, u = idx / m, j = perm[u];
    double s = sigma[j];
    cplx v = cmake(0, 0);
    if (!swapped) { double f = s > 0.0 ? 1.0 / sqrt(s) : 0.0; v = cscale(W[i + (size_t)j * mm], f); }
    else if (s > tiny) {
      for (int k = 0; k < n; ++k) cfma(v, d.theta[i + (size_t)k * m], W[k + (size_t)j * mm]);
      v = cscale(v, 1.0 / (s * sqrt(s)));
    }
    d.r1out[idx] = v;
  }
  for (int idx = threadIdx.x; idx < n * rank; idx += blockDim.x) {
    int i = idx % n, u = idx / n, j = perm[u];
    double s = sigma[j];
    cplx v = cmake(0, 0);
    if (swapped) { double f = s > 0.0 ? 1.0 / sqrt(s) : 0.0; v = cscale(cconj(W[i + (size_t)j * mm]), f); }
    else if (s > tiny) {
      for (int k = 0; k < m; ++k) cfma_cb(v, d.theta[k + (size_t)i * m], W[k + (size_t)j * mm]);
      v = cscale(v, 1.0 / (s * sqrt(s)));
    }
    d.r2out[idx] = v;
  }
}

// C = op(A) * B for small matrices (n <= 64): C[i,j] = sum_k A'[i,k] B[k,j], A' = A^T if transA.  One CTA per item.
struct SmallMatmulDesc { const cplx* a; const cplx* b; cplx* c; int m, k, n; int lda, ldb, ldc; int transA; };
__global__ void __launch_bounds__(256) k_small_matmul(const SmallMatmulDesc* __restrict__ descs) {
  const SmallMatmulDesc d = descs[blockIdx.x];
  for (int idx = threadIdx.x; idx < d.m * d.n; idx += blockDim.x) {
    int i = idx % d.m, j = idx / d.m;
    cplx acc = cmake(0, 0);
    for (int k = 0; k < d.k; ++k) {
      cplx a = d.transA ? d.a[k + (size_t)i * d.lda] : d.a[i + (size_t)k * d.lda];
      cfma(acc, a, d.b[k + (size_t)j * d.ldb]);
    }
    d.c[i + (size_t)j * d.ldc] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_tri_solve: W = R^-1 B for an upper-triangular R (n x n, ld = ldr) and n x m right-hand sides (ld = ldb), by back substitution in
// shared memory; one CTA per item, a thread per column of B.  Used by the optimistic gate path: T' = psi (R^-1 R') replaces
// Q' R' when the incoming messages are of full rank (tnqs_cuda.cu, GateBatch::stage2).  dynamic shared memory: n * n + n * m complex.
// ---------------------------------------------------------------------------------------------------------------
struct TriSolveDesc { const cplx* r; int ldr; const cplx* b; int ldb; cplx* w; int ldw; int n, m; };
__global__ void __launch_bounds__(256) k_tri_solve(const TriSolveDesc* __restrict__ descs) {
  const TriSolveDesc d = descs[blockIdx.x];
  extern __shared__ cplx tsm[];
  cplx* R = tsm;                       // n x n
  cplx* X = tsm + (size_t)d.n * d.n;   // n x m
  const int n = d.n, m = d.m;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) R[idx] = d.r[(idx % n) + (size_t)(idx / n) * d.ldr];
  for (int idx = threadIdx.x; idx < n * m; idx += blockDim.x) X[idx] = d.b[(idx % n) + (size_t)(idx / n) * d.ldb];
  __syncthreads();
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    cplx* x = X + (size_t)c * n;
    for (int i = n - 1; i >= 0; --i) {
      cplx s = x[i];
      for (int k = i + 1; k < n; ++k) s = csub(s, cmul(R[i + (size_t)k * n], x[k]));
      const cplx rii = R[i + (size_t)i * n];
      const double den = rii.x * rii.x + rii.y * rii.y;
      x[i] = cmake((s.x * rii.x + s.y * rii.y) / den, (s.y * rii.x - s.x * rii.y) / den);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n * m; idx += blockDim.x) d.w[(idx % n) + (size_t)(idx / n) * d.ldw] = X[idx];
}
