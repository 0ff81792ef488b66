/*
 * trig.cuh — sin and cos of one argument, evaluated together, with IDENTICAL results on the GPU
 * and on the host.
 *
 * Why not ::sin / ::cos.  libdevice's and glibc's double-precision sin/cos each stay within 1–2 ulp
 * of the true value but not of each other, so any check of the kernels against a CPU run of the
 * same source could only ever be approximate; and they are the most expensive thing in an acrobot
 * dynamics evaluation (four calls, ~65 SASS instructions each: 41 % of the warp instructions of
 * the first version of the solve kernel, profiles/r1a).  This header restates the classic
 * fdlibm/musl algorithm — Cody–Waite reduction by pi/2 in two 33-bit steps, then the degree-13 /
 * degree-14 minimax kernels with the reduction tail folded in — using only IEEE-754 operations
 * whose result is defined bit for bit (add, multiply, explicitly fused multiply-add, rint), so
 * g++ (tests/emu, -ffp-contract=off) and nvcc (-fmad=false) produce the same doubles.  Error
 * < 1 ulp, the same class as the reference's libm (checked against it in tests/test_trig.py).
 *
 * |x| >= 2^19 * pi/2 (a blown-up rollout) falls back to the platform's sincos: still correct, no
 * longer bit-reproducible between host and device.
 */
#ifndef ILQR_TRIG_CUH_
#define ILQR_TRIG_CUH_

#include <math.h>

#if defined(__CUDACC__)
#define ILQR_HD __host__ __device__ __forceinline__
#else
#define ILQR_HD inline __attribute__((always_inline))
#endif

namespace ilqr {

/* fdlibm e_rem_pio2.c / k_sin.c / k_cos.c constants.  On the device they sit in constant memory so
 * the DFMAs take them as c[bank][offset] operands; as literals the compiler materialised each one
 * with two UMOVs before use (a third of the instructions of this function, profiles/r1c). */
#define ILQR_TRIG_TABLE                                                                                              \
  {6.36619772367581382433e-01 /* invpio2 0x3FE45F306DC9C883 */, 1.57079632673412561417e+00 /* pio2_1 0x3FF921FB54400000 */, \
   6.07710050630396597660e-11 /* pio2_2 0x3DD0B4611A600000 */, 2.02226624879595063154e-21 /* pio2_2t 0x3BA3198A2E037073 */, \
   -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,   \
   -2.50507602534068634195e-08, 1.58969099521155010221e-10, /* S1..S6 */                                              \
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,  \
   2.08757232129817482790e-09, -1.13596475577881948265e-11 /* C1..C6 */}
#if defined(__CUDACC__)
static __constant__ double kTrigDev[16] = ILQR_TRIG_TABLE; /* internal linkage: one copy per translation unit */
#endif
static const double kTrigHost[16] = ILQR_TRIG_TABLE;

#if defined(__CUDACC__) && defined(ILQR_NOINLINE_TRIG)
#define ILQR_HD_TRIG __host__ __device__ __noinline__
#else
#define ILQR_HD_TRIG ILQR_HD
#endif

/* |x| small enough for the two-step Cody-Waite reduction (fdlibm's "medium" range is 2^19 * pi/2) */
ILQR_HD bool sincos_in_range(double x) { return ::fabs(x) < 8.0e5; } /* false for NaN */

/* v with its sign flipped when `bit` (0 or 2) is 2: -v is exactly a sign flip, and as an integer operation on the high
 * word it is one logic instruction instead of a negation in the fp64 pipe and two selects */
ILQR_HD double flip_if2(double v, int bit) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(v) ^ (bit << 30), __double2loint(v));
#else
  unsigned long long b;
  __builtin_memcpy(&b, &v, 8);
  b ^= (unsigned long long)(unsigned)(bit & 2) << 62;
  __builtin_memcpy(&v, &b, 8);
  return v;
#endif
}

ILQR_HD int lo32(double v) {
#if defined(__CUDA_ARCH__)
  return __double2loint(v);
#else
  long long b;
  __builtin_memcpy(&b, &v, 8);
  return (int)(unsigned)(b & 0xffffffffLL);
#endif
}

/* The branch-free core: exact for sincos_in_range(x), meaningless (but harmless) otherwise.  Callers that
 * need several sincos of independent arguments call this back to back and test the ranges once afterwards,
 * so the evaluations sit in one basic block and the instruction scheduler interleaves their dependency
 * chains (a branch per call would serialise them). */
ILQR_HD void sincos_core(double x, double *sn, double *cs) {
#if defined(__CUDA_ARCH__)
  const double *tab = kTrigDev;
#else
  const double *tab = kTrigHost;
#endif
  const double invpio2 = tab[0], pio2_1 = tab[1], pio2_2 = tab[2], pio2_2t = tab[3];
  const double S1 = tab[4], S2 = tab[5], S3 = tab[6], S4 = tab[7], S5 = tab[8], S6 = tab[9];
  const double C1 = tab[10], C2 = tab[11], C3 = tab[12], C4 = tab[13], C5 = tab[14], C6 = tab[15];
  /* x = n * pi/2 + (y0 + y1), |y0| <= pi/4 (+ rounding), 118-bit pi/2 */
  const double fn = ::rint(x * invpio2);
  const int n = (int)fn;
  const double t = ::fma(-fn, pio2_1, x);
  double w = fn * pio2_2;
  const double r = t - w;
  w = ::fma(fn, pio2_2t, -((t - r) - w));
  const double y0 = r - w;
  const double y1 = (r - y0) - w;
  /* kernels on [-pi/4, pi/4] with tail y1 */
  const double z = y0 * y0;
  const double v = z * y0;
  const double ps = ::fma(z, ::fma(z, ::fma(z, ::fma(z, S6, S5), S4), S3), S2);
  const double ks = y0 - ((::fma(z, ::fma(-v, ps, 0.5 * y1), -y1)) - v * S1);
  const double pc = z * ::fma(z, ::fma(z, ::fma(z, ::fma(z, ::fma(z, C6, C5), C4), C3), C2), C1);
  const double hz = 0.5 * z;
  const double wc = 1.0 - hz;
  const double kc = wc + (((1.0 - wc) - hz) + ::fma(z, pc, -(y0 * y1)));
  /* quadrant */
  const double a = (n & 1) ? kc : ks;
  const double b = (n & 1) ? ks : kc;
  *sn = flip_if2(a, n & 2);
  *cs = flip_if2(b, (n + 1) & 2);
}

/* The platform's sincos for arguments outside the range of sincos_core.  Out of line on the device: it is never
 * reached by a sane trajectory, and inlined (several hundred instructions of Payne-Hanek reduction at every call
 * site) it sat between the hot instructions and cost instruction-cache misses (ncu, bulk regime: 12-13 % of the stall
 * samples were "no instruction").  Argument and results travel BY VALUE: a noinline callee that is handed the
 * addresses of the callers' sn[] / cs[] forces those arrays into local memory on the hot path as well (measured: -12 %). */
struct SinCos {
  double s, c;
};
#if defined(__CUDACC__) && !defined(ILQR_SINCOS_SLOW_INLINE)
static __host__ __device__ __noinline__ /* internal linkage: the header is included by several translation units */
#else
ILQR_HD
#endif
SinCos sincos_slow(double x) {
  SinCos r;
  ::sincos(x, &r.s, &r.c);
  return r;
}

ILQR_HD_TRIG void sincos_det(double x, double *sn, double *cs) {
#if defined(ILQR_TRIG_LIBM) && !defined(__CUDACC__)
  /* tests/emu only: the platform libm, to compare the kernel source bit for bit with the oracle */
  *sn = ::sin(x);
  *cs = ::cos(x);
#if defined(ILQR_TRIG_NOISE)
  /* tools/exp_attribution.py only: "another correct libm" — a deterministic function of the argument that moves about
   * one result in sixteen by one ulp, the kind of disagreement two < 1 ulp implementations have (trig.cuh and glibc
   * disagree on 3 % of arguments, tests/test_trig.py) */
  {
    unsigned long long b;
    __builtin_memcpy(&b, &x, 8);
    b ^= b >> 29;
    b *= 0x9E3779B97F4A7C15ULL;
    b ^= b >> 32;
    const unsigned h = (unsigned)(b & 63u);
    if (h == 0) *sn = ::nextafter(*sn, 2.0);
    if (h == 1) *sn = ::nextafter(*sn, -2.0);
    if (h == 2) *cs = ::nextafter(*cs, 2.0);
    if (h == 3) *cs = ::nextafter(*cs, -2.0);
  }
#endif
  return;
#endif
  if (__builtin_expect(!sincos_in_range(x), 0)) {
    const SinCos r = sincos_slow(x);
    *sn = r.s;
    *cs = r.c;
    return;
  }
  sincos_core(x, sn, cs);
}

/* K independent arguments at once: sincos_core written "one operation, K arguments" at a time, so that the
 * K dependency chains (about 25 operations of 8 cycles each) advance together in the instruction stream.  A
 * single warp issues in order: K back-to-back sincos_core calls run one after the other (measured: 661 cycles
 * for three), because the compiler does not interleave basic-block-sized chains on its own.  Bit-identical
 * to K sincos_core calls — the same operations on the same operands. */
template <int K>
ILQR_HD void sincos_coreN(const double *x, double *sn, double *cs) {
#if defined(__CUDA_ARCH__)
  const double *tab = kTrigDev;
#else
  const double *tab = kTrigHost;
#endif
  /* the four reduction constants as immediates: they head the dependency chain, and a constant-memory load there
   * is exposed latency (ncu: 3.9 % of the stall samples sat on the first multiply) */
  const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00, pio2_2 = 6.07710050630396597660e-11,
               pio2_2t = 2.02226624879595063154e-21;
  const double S1 = tab[4], S2 = tab[5], S3 = tab[6], S4 = tab[7], S5 = tab[8], S6 = tab[9];
  const double C1 = tab[10], C2 = tab[11], C3 = tab[12], C4 = tab[13], C5 = tab[14], C6 = tab[15];
  double fn[K], t[K], w[K], r[K], y0[K], y1[K], z[K], v[K], ps[K], pc[K], ks[K], kc[K], hz[K], wc[K];
  int n[K];
#define ILQR_EACH for (int i = 0; i < K; i++)
#pragma unroll
  ILQR_EACH {
    const double m = x[i] * invpio2 + 6755399441055744.0; /* 1.5 * 2^52: the sum's low mantissa bits are rint() */
    fn[i] = m - 6755399441055744.0;
    n[i] = lo32(m);
  }
#pragma unroll
  ILQR_EACH t[i] = ::fma(-fn[i], pio2_1, x[i]);
#pragma unroll
  ILQR_EACH w[i] = fn[i] * pio2_2;
#pragma unroll
  ILQR_EACH r[i] = t[i] - w[i];
#pragma unroll
  ILQR_EACH w[i] = ::fma(fn[i], pio2_2t, -((t[i] - r[i]) - w[i]));
#pragma unroll
  ILQR_EACH y0[i] = r[i] - w[i];
#pragma unroll
  ILQR_EACH y1[i] = (r[i] - y0[i]) - w[i];
#pragma unroll
  ILQR_EACH z[i] = y0[i] * y0[i];
#pragma unroll
  ILQR_EACH v[i] = z[i] * y0[i];
  /* the two polynomials, one Horner step of each per pass */
#pragma unroll
  ILQR_EACH ps[i] = ::fma(z[i], S6, S5);
#pragma unroll
  ILQR_EACH pc[i] = ::fma(z[i], C6, C5);
#pragma unroll
  ILQR_EACH ps[i] = ::fma(z[i], ps[i], S4);
#pragma unroll
  ILQR_EACH pc[i] = ::fma(z[i], pc[i], C4);
#pragma unroll
  ILQR_EACH ps[i] = ::fma(z[i], ps[i], S3);
#pragma unroll
  ILQR_EACH pc[i] = ::fma(z[i], pc[i], C3);
#pragma unroll
  ILQR_EACH ps[i] = ::fma(z[i], ps[i], S2);
#pragma unroll
  ILQR_EACH pc[i] = ::fma(z[i], pc[i], C2);
#pragma unroll
  ILQR_EACH pc[i] = z[i] * ::fma(z[i], pc[i], C1);
#pragma unroll
  ILQR_EACH ks[i] = y0[i] - ((::fma(z[i], ::fma(-v[i], ps[i], 0.5 * y1[i]), -y1[i])) - v[i] * S1);
#pragma unroll
  ILQR_EACH hz[i] = 0.5 * z[i];
#pragma unroll
  ILQR_EACH wc[i] = 1.0 - hz[i];
#pragma unroll
  ILQR_EACH kc[i] = wc[i] + (((1.0 - wc[i]) - hz[i]) + ::fma(z[i], pc[i], -(y0[i] * y1[i])));
#pragma unroll
  ILQR_EACH {
    const double a = (n[i] & 1) ? kc[i] : ks[i];
    const double b = (n[i] & 1) ? ks[i] : kc[i];
    sn[i] = flip_if2(a, n[i] & 2);
    cs[i] = flip_if2(b, (n[i] + 1) & 2);
  }
#undef ILQR_EACH
}

/* K at once with the range checks; same results as K sincos_det calls */
template <int K>
ILQR_HD void sincos_detN(const double *x, double *sn, double *cs) {
#if defined(ILQR_TRIG_LIBM) && !defined(__CUDACC__)
  for (int i = 0; i < K; i++) sincos_det(x[i], sn + i, cs + i);
  return;
#endif
  sincos_coreN<K>(x, sn, cs);
  bool ok = true;
#pragma unroll
  for (int i = 0; i < K; i++) ok = ok && sincos_in_range(x[i]);
  if (__builtin_expect(!ok, 0)) { /* the in-range ones already hold what sincos_det would give */
#pragma unroll
    for (int i = 0; i < K; i++)
      if (!sincos_in_range(x[i])) {
        const SinCos r = sincos_slow(x[i]);
        sn[i] = r.s;
        cs[i] = r.c;
      }
  }
}

/* three at once (the acrobot's arguments) */
ILQR_HD void sincos_det3(double x0, double x1, double x2, double *sn, double *cs) {
  const double x[3] = {x0, x1, x2};
  sincos_detN<3>(x, sn, cs);
}

ILQR_HD void sincos_det(float x, float *sn, float *cs) { ::sincosf(x, sn, cs); }
ILQR_HD void sincos_det3(float x0, float x1, float x2, float *sn, float *cs) {
  ::sincosf(x0, sn + 0, cs + 0);
  ::sincosf(x1, sn + 1, cs + 1);
  ::sincosf(x2, sn + 2, cs + 2);
}
template <int K>
ILQR_HD void sincos_detN(const float *x, float *sn, float *cs) {
#pragma unroll
  for (int i = 0; i < K; i++) ::sincosf(x[i], sn + i, cs + i);
}

}  // namespace ilqr
#endif
