/*
 * ilqr_synth.h — the synthetic problem-instance generator shared by the CUDA
 * library, the CPU oracle and the reference harness, so every arm of every
 * test and benchmark sees bit-identical (x0, u0).
 *
 * Header-only, plain C99.  It restates the 64-bit Mersenne Twister
 * (Matsumoto & Nishimura, MT19937-64) and the way libstdc++ turns its output
 * into uniform_real_distribution<double>(-1,1): one 64-bit draw, converted to
 * double, divided by 2^64, clamped below 1, then mapped to 2*r-1.  The harness
 * checks this restatement against the real std:: classes
 * (tests/test_oracle_ref.py::test_synth_matches_std_mt19937_64).
 *
 * Draw order (SURVEY.md §8d "Concrete synthetic inputs"): trajectory-major;
 * for each trajectory b: x0[b][0..n) = x_scale*U, then u0[b][t][0..m) =
 * u_scale*U for t < T.  If `canonical_first` is set, trajectory 0 is then
 * overwritten with x0 = 0, u0 = 0 (the reference CLI's acrobot instance,
 * src/run_ilqr.cpp:39-54) so the golden vectors apply to it.
 */
#ifndef ILQR_SYNTH_H_
#define ILQR_SYNTH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t mt[312];
  int idx;
} ilqr_mt64;

static inline void ilqr_mt64_seed(ilqr_mt64 *g, uint64_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 312; ++i)
    g->mt[i] = 6364136223846793005ULL * (g->mt[i - 1] ^ (g->mt[i - 1] >> 62)) + (uint64_t)i;
  g->idx = 312;
}

static inline uint64_t ilqr_mt64_next(ilqr_mt64 *g) {
  if (g->idx >= 312) {
    for (int i = 0; i < 312; ++i) {
      uint64_t y = (g->mt[i] & 0xFFFFFFFF80000000ULL) | (g->mt[(i + 1) % 312] & 0x7FFFFFFFULL);
      uint64_t v = g->mt[(i + 156) % 312] ^ (y >> 1);
      if (y & 1ULL) v ^= 0xB5026F5AA96619E9ULL;
      g->mt[i] = v;
    }
    g->idx = 0;
  }
  uint64_t x = g->mt[g->idx++];
  x ^= (x >> 29) & 0x5555555555555555ULL;
  x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
  x ^= (x << 37) & 0xFFF7EEE000000000ULL;
  x ^= (x >> 43);
  return x;
}

/* uniform_real_distribution<double>(-1, 1) as libstdc++ evaluates it. */
static inline double ilqr_mt64_uniform_pm1(ilqr_mt64 *g) {
  double r = (double)ilqr_mt64_next(g) / 18446744073709551616.0;
  if (r >= 1.0) r = 0.99999999999999988897769753748; /* nextafter(1,0) */
  return r * 2.0 + -1.0;
}

/* Fill x0[B][n] and u0[B][T][m] (row-major doubles). */
static inline void ilqr_synth_fill(uint64_t seed, size_t B, int T, int n, int m, double x_scale,
                                   double u_scale, int canonical_first, double *x0, double *u0) {
  ilqr_mt64 g;
  ilqr_mt64_seed(&g, seed);
  for (size_t b = 0; b < B; ++b) {
    for (int i = 0; i < n; ++i) x0[b * (size_t)n + i] = x_scale * ilqr_mt64_uniform_pm1(&g);
    double *u = u0 + b * (size_t)T * m;
    for (int t = 0; t < T * m; ++t) u[t] = u_scale * ilqr_mt64_uniform_pm1(&g);
  }
  if (canonical_first && B > 0) {
    for (int i = 0; i < n; ++i) x0[i] = 0.0;
    for (int t = 0; t < T * m; ++t) u0[t] = 0.0;
  }
}

#ifdef __cplusplus
}
#endif
#endif /* ILQR_SYNTH_H_ */
