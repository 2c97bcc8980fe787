/* examples/c_abi_demo.c - libtnqs_cuda.so driven from plain C99 through include/tnqs_cuda.h (no C++, no Python): a 2x2 grid of
 * spin-up qubits, one Rzz layer as a circuit, a BP refresh, the one-site density matrix of vertex 0.  Without a CUDA device
 * tnqs_create fails with TNQS_ERR_CUDA (there is no CPU fallback) and the program says so and exits 3.
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c_abi_demo.c -Ltensornetworkquantumsimulator_b200 -ltnqs_cuda -lm -o /tmp/c_abi_demo
 *   LD_LIBRARY_PATH=tensornetworkquantumsimulator_b200 /tmp/c_abi_demo */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "tnqs_cuda.h"

#define CHECK(call)                                                                                  \
  do {                                                                                               \
    int32_t st_ = (call);                                                                            \
    if (st_ != TNQS_OK) {                                                                            \
      fprintf(stderr, "%s -> %s: %s\n", #call, tnqs_status_string(st_), tnqs_last_error(ctx));       \
      if (ctx) tnqs_destroy(ctx);                                                                    \
      return 1;                                                                                      \
    }                                                                                                \
  } while (0)

int main(void) {
  /* 2x2 grid: vertices 0..3, edges (0,1) (2,3) (0,2) (1,3) */
  const int32_t esrc[4] = {0, 2, 0, 1}, edst[4] = {1, 3, 2, 3};
  const int32_t d[4] = {2, 2, 2, 2}, chi[4] = {1, 1, 1, 1};
  tnqs_ctx* ctx = NULL;
  int32_t st;
  printf("%s\n", tnqs_version());
  st = tnqs_create(&ctx, 4, 4, esrc, edst, d, chi, /*chi_cap=*/2, /*device=*/0);
  if (st == TNQS_ERR_CUDA) {
    printf("no CUDA device: %s (%s)\n", tnqs_status_string(st), tnqs_last_error(NULL));
    return 3;
  }
  if (st != TNQS_OK) { fprintf(stderr, "tnqs_create: %s\n", tnqs_last_error(NULL)); return 1; }

  /* |+> on every site: T[s, 1, 1] = 1/sqrt(2) */
  {
    const double r = 1.0 / sqrt(2.0);
    const double plus[4] = {r, 0.0, r, 0.0};             /* two ComplexF64, interleaved */
    int32_t v;
    for (v = 0; v < 4; ++v) CHECK(tnqs_set_site(ctx, v, plus, 2));
  }

  /* one layer of Rzz(0.8) = exp(-i 0.4 ZZ) on every edge, as an ordered circuit; gate array o[s1', s2', s1, s2] column-major */
  {
    double gate[4 * 16 * 2];
    int32_t gate_v[8];
    int64_t gate_off[4];
    const int32_t seq[8] = {0, 2, 4, 6, 7, 5, 3, 1};     /* any order of the 8 directed edges is a valid BP sequence */
    const tnqs_apply_opts aopts = {2, 1e-12, 1};
    const tnqs_bp_opts bopts = {25, 1e-10, 1, 0, TNQS_BP_REFERENCE_ORDER};
    double truncerr[4];
    int32_t refreshes = 0, e, i;
    memset(gate, 0, sizeof gate);
    for (e = 0; e < 4; ++e) {
      for (i = 0; i < 4; ++i) {                            /* diagonal entry (i, i), i = s1 + 2*s2 */
        const double sign = (i == 0 || i == 3) ? 1.0 : -1.0;
        gate[2 * (16 * e + i + 4 * i)] = cos(0.4);
        gate[2 * (16 * e + i + 4 * i) + 1] = -sign * sin(0.4);
      }
      gate_v[2 * e] = esrc[e];
      gate_v[2 * e + 1] = edst[e];
      gate_off[e] = 16 * e;
    }
    CHECK(tnqs_apply_circuit(ctx, 4, gate_v, gate_off, gate, &aopts, seq, 8, &bopts, 1, truncerr, &refreshes));
    printf("BP refreshes: %d, truncation errors: %.2e %.2e %.2e %.2e\n", (int)refreshes, truncerr[0], truncerr[1], truncerr[2], truncerr[3]);
  }

  {
    int32_t bond[4];
    const int32_t v0 = 0;
    double rho[8];
    CHECK(tnqs_get_bond_dims(ctx, bond));
    CHECK(tnqs_rdm_1site(ctx, 1, &v0, rho));
    printf("bond dimensions: %d %d %d %d\n", (int)bond[0], (int)bond[1], (int)bond[2], (int)bond[3]);
    /* rho[s, s'] column-major: <Z> = (rho00 - rho11) / (rho00 + rho11), <X> = 2 Re rho01 / tr */
    printf("<Z_0> = %.12f  <X_0> = %.12f  kernel launches: %lld\n", (rho[0] - rho[6]) / (rho[0] + rho[6]),
           2.0 * rho[2] / (rho[0] + rho[6]), (long long)tnqs_kernel_launches(ctx));
  }
  tnqs_destroy(ctx);
  return 0;
}
