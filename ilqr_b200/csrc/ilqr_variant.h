/*
 * ilqr_variant.h — first include of the per-model translation units.  Each of them is compiled twice (Makefile):
 *   default            -fmad=false: the reference's arithmetic, operation for operation (no fused multiply-add on its
 *                      x86-64 build), bit-identical to the oracle up to sin/cos;
 *   -DILQR_FMA_BUILD   -fmad=true: the compiler contracts a * b + c into one fused operation.  Dot products lose half
 *                      their dependent chain and a quarter of their instructions; results move by rounding (1e-13
 *                      per operation), within BASELINE's 1e-6 on K, k and cost after a handful of trips but no longer
 *                      bit-identical to anything on the CPU.  Opt-in: ilqr_desc.flags & ILQR_FLAG_FAST_FMA.
 * The second build lives in its own namespace (the macro below renames `ilqr`), so the two sets of kernels and host
 * stubs are distinct symbols in one library; ILQR_ENTRY names the extern entry point of the variant.
 */
#ifndef ILQR_VARIANT_H_
#define ILQR_VARIANT_H_
#if defined(ILQR_FMA_BUILD)
#define ilqr ilqr_fma
#define ILQR_ENTRY(name) name##_fma
#else
#define ILQR_ENTRY(name) name
#endif
#endif
