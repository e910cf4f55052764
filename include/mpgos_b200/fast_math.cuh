// fast_math.cuh -- opt-in replacements for the double-precision transcendentals that dominate transcendental-heavy right-hand
// sides on sm_100a (Keller-Miksis: pow + sincospi + sin + cos are ~80 % of a step's issue slots; 17 % of the executed
// instructions are UMOV pairs that rebuild the CUDA math library's FP64 literals, because FP64 instructions take constants only
// from the 63 uniform registers).
//
//   mpgos_b200::fm::sincos / sin / cos (double)   <= 1.5 ulp for |x| <= 1e6, the CUDA function beyond that and for NaN / Inf
//   mpgos_b200::fm::sincospi (double)             <= 1.7 ulp for |x| <= 2^30, the CUDA function beyond
//   mpgos_b200::fm::rcp (double)                  <= 1 ulp for normal finite x, the division otherwise (built-in "_fm" systems only: user code writes 1.0 / x)
//   mpgos_b200::fm::pow (double, double)          <= 3 ulp for normal positive finite x and |y log2 x| <= 1000, the CUDA function otherwise
//                                                 (x <= 0, subnormal, Inf, NaN, over/underflowing results: all special cases are CUDA's)
//   (errors measured against 200-bit references over the ranges of tests/test_fast_math_gpu.py; CUDA's own: sin/cos/sincospi/pow <= 2 ulp)
//
// What makes them cheap here: coefficient tables in __constant__ memory (one LDCU per one or two constants instead of two UMOVs per
// literal), range reduction by adding and subtracting 1.5 * 2^52 (the integer part falls out of the low mantissa word: no FRND,
// F2I or I2F on the FP64 pipe), no Payne-Hanek path and no double-double logarithm: pow forms y * log2(x) as a compensated
// product of two doubles, which is enough for 3 ulp at |y log2 x| <= 1000, where CUDA's accurate pow spends ~80 FP64 instructions.
// Instruction counts (sm_100a, per call): pow ~50 FP64 + ~15 other against ~95 + ~130; sincos ~24 + ~10 against ~50 + ~100 for
// sin + cos of the same argument; sincospi ~23 + ~10 against ~25 + ~50.
//
// OPT-IN.  Nothing in the library uses these unless asked to: the built-in systems exist in both forms ("kellermiksis" /
// "kellermiksis_fm", ...), and a user translation unit redirects the calls of its UNMODIFIED *_SystemDefinition.cuh by defining
// MPGOS_B200_FAST_MATH before including this header ahead of the system definition (or with nvcc -include):
//     #define MPGOS_B200_FAST_MATH
//     #include "mpgos_b200/fast_math.cuh"
//     #include "KellerMiksis_SystemDefinition.cuh"
//     #include "SingleSystem_PerThread_Interface.cuh"
// or, for a whole program as it is:  nvcc -DMPGOS_B200_FAST_MATH -include mpgos_b200/fast_math.cuh Program.cu ...  The redirect applies to
// device code only: host code of the same translation unit (a tutorial's main() computes its parameters with pow) keeps the C library.
// The redirect is a set of function-like macros (pow, sin, cos, sincos, sincospi): library headers that use these names themselves
// (the C++ standard headers below are taken care of; thrust, cub, ... are not) must be included BEFORE this header.
// Single-precision calls are passed on to the CUDA functions unchanged.
#ifndef MPGOS_B200_FAST_MATH_CUH
#define MPGOS_B200_FAST_MATH_CUH

#include <cuda_runtime.h>
#include <math.h>
#ifdef MPGOS_B200_FAST_MATH
// everything the interface headers and the tutorials' programs include that mentions these names (std::pow inside <random>, ...),
// before the names below become macros
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <ctime>
#include <limits>
#include <numeric>
#include <random>
#include <valarray>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>
#endif

namespace mpgos_b200 {
namespace fm {

// The coefficient table, in the order of use and in PAIRS: every load below is a volatile 16-byte constant load (LDCU.128: two
// doubles straight into uniform registers, where FP64 instructions take them from).  Volatile, because the compiler would otherwise
// hoist all ~45 constants out of the caller's step loop, run out of the 63 uniform registers, park the rest in vector registers
// (spilling some) and move them back with two R2UR per constant and use -- measured on the Keller-Miksis kernel: 220 R2UR, 127 LDC
// and 49 local-memory loads per step.
enum FmPair
{
	FM_S6C6, FM_S5C5, FM_S4C4, FM_S3C3, FM_S2C2, FM_S1C1,     // sin(r) = r + r^3 (S1 + r^2 (S2 + ...)), cos(r) = 1 - r^2/2 + r^4 (C1 + r^2 (C2 + ...)), |r| <= pi/4
	FM_2OPI_PIO2A, FM_PIO2B_PIO2C,                            // 2/pi and pi/2 in three parts (Cody-Waite with FMA)
	FM_PIHI_PILO,
	FM_LG6_LG7, FM_LG4_LG5, FM_LG2_LG3, FM_LG1_L2EH,          // log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)), s = f / (2 + f); log2(e) hi
	FM_L2EL_LN2H, FM_LN2L_E13,                                // log2(e) lo, ln 2 hi / lo
	FM_E12_E11, FM_E10_E9, FM_E8_E7, FM_E6_E5, FM_E4_E3, FM_E2_PAD,      // e^t = 1 + t + t^2/2! + ... + t^13/13!, |t| <= ln2/2
	FM_PAIRS
};

static __constant__ __align__(16) double kFm[2 * FM_PAIRS] = {
	1.58969099521155010221e-10, -1.13596475577881948265e-11,   -2.50507602534068634195e-08, 2.08757232129817482790e-09,
	2.75573137070700676789e-06, -2.75573143513906633035e-07,   -1.98412698298579493134e-04, 2.48015872894767294178e-05,
	8.33333333332248946124e-03, -1.38888888888741095749e-03,   -1.66666666666666324348e-01, 4.16666666666666019037e-02,
	0.6366197723675814, 1.5707963267948966,                    6.123233995736766e-17, -1.4973849048591698e-33,
	3.141592653589793, 1.2246467991473532e-16,
	1.531383769920937332e-01, 1.479819860511658591e-01,        2.222219843214978396e-01, 1.818357216161805012e-01,
	3.999999999940941908e-01, 2.857142874366239149e-01,        6.666666666666735130e-01, 1.4426950408889634,
	2.0355273740931033e-17, 0.6931471805599453,                2.3190468138462996e-17, 1.0 / 6227020800.0,
	1.0 / 479001600.0, 1.0 / 39916800.0,                       1.0 / 3628800.0, 1.0 / 362880.0,
	1.0 / 40320.0, 1.0 / 5040.0,                               1.0 / 720.0, 1.0 / 120.0,
	1.0 / 24.0, 1.0 / 6.0,                                     1.0 / 2.0, 0.0
};

__device__ __forceinline__ void ldpair(int pair, double& a, double& b)
{
	asm volatile("ld.const.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(__cvta_generic_to_constant(&kFm[2 * pair])));
}

// the CUDA functions, out of line: taken only outside the ranges stated above, and kept out of the callers' loop bodies
static __device__ __noinline__ void   cuda_sincos(double x, double* s, double* c)   { ::sincos(x, s, c); }      // [cold>
static __device__ __noinline__ void   cuda_sincospi(double x, double* s, double* c) { ::sincospi(x, s, c); }
static __device__ __noinline__ double cuda_pow(double x, double y)                  { return ::pow(x, y); }
static __device__ __noinline__ double cuda_rcp(double x)                            { return 1.0 / x; }
struct Rcp2 { double a, b; };
static __device__ __noinline__ Rcp2   cuda_rcp2(double a, double b)                 { Rcp2 o;  o.a = 1.0 / a;  o.b = 1.0 / b;  return o; }      // <cold]

// The three functions are real subroutines by default (MPGOS_B200_FM_INLINE 0).  Inlined into a step loop, their ~45 constants are
// loop invariants: the compiler hoists the loads, runs out of the 63 uniform registers, parks the rest in vector registers and moves
// them back with two R2UR per constant and use (measured on the Keller-Miksis kernel: 240 R2UR + 63 LDC + 47 local-memory loads per
// step, 188 bytes of spill).  A subroutine has no loop to hoist out of: every pair of constants is one LDCU.128 where it is used, and
// the six right-hand-side evaluations of a step share one copy of the code.
#ifndef MPGOS_B200_FM_INLINE
#define MPGOS_B200_FM_INLINE 0
#endif
#if MPGOS_B200_FM_INLINE
#define MPGOS_FM_FN static __device__ __forceinline__
#else
#define MPGOS_FM_FN static __device__ __noinline__
#endif

// 1.5 * 2^52: x + kMagic - kMagic rounds x to the nearest integer (|x| < 2^51), and the low mantissa word of x + kMagic holds that
// integer in two's complement
#define MPGOS_FM_MAGIC 6755399441055744.0

// sin and cos of r, |r| <= pi/4 (the polynomials of fdlibm's __kernel_sin / __kernel_cos), then the quadrant q (0..3):
// (sin, cos)(r + q pi/2); sign changes on the words that hold the sign
struct SinCos { double s, c; };      // returned by value: in registers across the subroutine boundary

__device__ __forceinline__ SinCos kernel_sincos(double r, int q)
{
	const double z = r * r;
	double ps, pc, a, b;
	ldpair(FM_S6C6, ps, pc);
	ldpair(FM_S5C5, a, b);  ps = fma(ps, z, a);  pc = fma(pc, z, b);
	ldpair(FM_S4C4, a, b);  ps = fma(ps, z, a);  pc = fma(pc, z, b);
	ldpair(FM_S3C3, a, b);  ps = fma(ps, z, a);  pc = fma(pc, z, b);
	ldpair(FM_S2C2, a, b);  ps = fma(ps, z, a);  pc = fma(pc, z, b);
	ldpair(FM_S1C1, a, b);  ps = fma(ps, z, a);  pc = fma(pc, z, b);
	const double s = fma(r * z, ps, r);
	const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
	const bool odd = (q & 1) != 0;
	const double u = odd ? c : s, v = odd ? s : c;
	const unsigned flip_s = (unsigned)(q & 2) << 30, flip_c = (unsigned)((q + 1) & 2) << 30;
	SinCos out;
	out.s = __hiloint2double((int)((unsigned)__double2hiint(u) ^ flip_s), __double2loint(u));
	out.c = __hiloint2double((int)((unsigned)__double2hiint(v) ^ flip_c), __double2loint(v));
	return out;
}

// (const: no side effects, the result depends on the argument alone -- sin(x) and cos(x) of one argument in an unmodified definition
// then share one call)
MPGOS_FM_FN __attribute__((const)) SinCos sincos_core(double x)
{
	double two_over_pi, p1, p2, p3;
	ldpair(FM_2OPI_PIO2A, two_over_pi, p1);
	ldpair(FM_PIO2B_PIO2C, p2, p3);
	const double t = fma(x, two_over_pi, MPGOS_FM_MAGIC);
	const double n = t - MPGOS_FM_MAGIC;
	double r = fma(-n, p1, x);
	r = fma(-n, p2, r);
	r = fma(-n, p3, r);
	SinCos out = kernel_sincos(r, __double2loint(t) & 3);
	if (!(fabs(x) <= 1.0e6)) cuda_sincos(x, &out.s, &out.c);           // large arguments, NaN, Inf: the CUDA function
	return out;
}
__device__ __forceinline__ void sincos(double x, double* sn, double* cs) { const SinCos r = sincos_core(x); *sn = r.s; *cs = r.c; }
__device__ __forceinline__ double sin(double x) { double s, c; mpgos_b200::fm::sincos(x, &s, &c); return s; }
__device__ __forceinline__ double cos(double x) { double s, c; mpgos_b200::fm::sincos(x, &s, &c); return c; }

MPGOS_FM_FN __attribute__((const)) SinCos sincospi_core(double x)
{
	double pi_hi, pi_lo;
	ldpair(FM_PIHI_PILO, pi_hi, pi_lo);
	const double t = fma(x, 2.0, MPGOS_FM_MAGIC);
	const double n = t - MPGOS_FM_MAGIC;
	const double r = fma(n, -0.5, x);                     // exact: |r| <= 1/4
	SinCos out = kernel_sincos(fma(r, pi_lo, r * pi_hi), __double2loint(t) & 3);
	if (!(fabs(x) <= 1073741824.0)) cuda_sincospi(x, &out.s, &out.c);
	return out;
}
__device__ __forceinline__ void sincospi(double x, double* sn, double* cs) { const SinCos r = sincospi_core(x); *sn = r.s; *cs = r.c; }

// x^y for normal positive finite x and moderate |y log2 x|; everything else (including every special case of the C standard) is
// the CUDA function's.  Straight-line: the fast path is computed unconditionally (on operands outside its range it produces
// garbage, never a fault) and replaced at the end.
MPGOS_FM_FN __attribute__((const)) double pow(double x, double y)
{
	const int hx = __double2hiint(x);
	// x = 2^e m, m in [sqrt(1/2), sqrt(2))
	const int h2 = hx + (0x3ff00000 - 0x3fe6a09e);
	const int e = (h2 >> 20) - 0x3ff;
	const double m = __hiloint2double((h2 & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
	const double ke = __hiloint2double(0x43300000, 0x80000000 ^ e) - 4503601774854144.0;      // (double)e without I2F: 2^52 + 2^31 subtracted
	// log(m): s = f / (2 + f), the reciprocal by two Newton steps on the hardware's ~20-bit one
	const double f = m - 1.0;
	const double den = 2.0 + f;
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));      // MUFU.RCP64H
	r = fma(r, fma(-den, r, 1.0), r);
	r = fma(r, fma(-den, r, 1.0), r);
	const double s = f * r;
	const double z = s * s, w = z * z;
	double lg6, lg7, lg4, lg5, lg2, lg3, lg1, l2e_hi, l2e_lo, ln2_hi, ln2_lo, c13, a, b;
	ldpair(FM_LG6_LG7, lg6, lg7);
	ldpair(FM_LG4_LG5, lg4, lg5);
	ldpair(FM_LG2_LG3, lg2, lg3);
	ldpair(FM_LG1_L2EH, lg1, l2e_hi);
	const double t1 = w * fma(w, fma(w, lg6, lg4), lg2);
	const double t2 = z * fma(w, fma(w, fma(w, lg7, lg5), lg3), lg1);
	const double hfsq = 0.5 * f * f;
	const double logm = f - (hfsq - s * (hfsq + (t2 + t1)));
	// log2(x) = e + logm * log2(e) as hi + lo, z = y * log2(x) as zh + zl
	ldpair(FM_L2EL_LN2H, l2e_lo, ln2_hi);
	const double hi = fma(logm, l2e_hi, ke);
	double lo = fma(logm, l2e_hi, ke - hi);
	lo = fma(logm, l2e_lo, lo);
	const double zh = y * hi;
	const double zl = fma(y, lo, fma(y, hi, -zh));
	// 2^zh = 2^n e^t
	const double tn = zh + MPGOS_FM_MAGIC;
	const double n = tn - MPGOS_FM_MAGIC;
	const double g = zh - n;
	ldpair(FM_LN2L_E13, ln2_lo, c13);
	const double t = fma(g, ln2_lo, g * ln2_hi);
	// e^t = 1 + t + t^2 (E(t^2) + t O(t^2)): two Horner chains in t^2 side by side (the step loops that call this are latency-bound:
	// one 13-deep chain of dependent FMAs costs more than its issue slots)
	const double tt = t * t;
	double pe, po;
	ldpair(FM_E12_E11, pe, po);      po = fma(c13, tt, po);
	ldpair(FM_E10_E9, a, b);         pe = fma(pe, tt, a);  po = fma(po, tt, b);
	ldpair(FM_E8_E7, a, b);          pe = fma(pe, tt, a);  po = fma(po, tt, b);
	ldpair(FM_E6_E5, a, b);          pe = fma(pe, tt, a);  po = fma(po, tt, b);
	ldpair(FM_E4_E3, a, b);          pe = fma(pe, tt, a);  po = fma(po, tt, b);
	pe = fma(pe, tt, 0.5);
	double p = fma(po, t, pe);       // (t^2 / 2! + t^4 / 4! + ...) / t^2 + t (t^2 / 3! + ...) / t^2
	p = fma(p, tt, t);
	p = p + 1.0;
	p = fma(p, zl * ln2_hi, p);                                                   // e^(zl ln 2) = 1 + zl ln 2
	p = p * __hiloint2double((__double2loint(tn) + 1023) << 20, 0);
	// zero, subnormal, negative, Inf, NaN base; overflow, underflow, y = NaN / Inf
	if (((unsigned)(hx - 0x00100000) >= 0x7fe00000u) | !(fabs(zh) <= 1000.0)) p = cuda_pow(x, y);      // bitwise: one branch
	return p;
}

// pow, sincospi and sincos of three unrelated arguments in ONE subroutine: x^y, (sin, cos)(pi a), (sin, cos)(b).  The three
// evaluations are independent chains of dependent FP64 instructions (latency 8 cycles each); in one basic block the compiler
// interleaves them, so a warp issues three instructions where it waited for one -- what a right-hand side of the Keller-Miksis type
// needs when few warps are resident (the block-per-system kernel: 4 per scheduler) -- and two of the three calls with their
// argument and result moves go away.  Outside the fast ranges of ANY of the three, all five results are the CUDA functions'.
struct PowSinCos2 { double p, s1, c1, s2, c2; };

static __device__ __noinline__ PowSinCos2 cuda_pow_sincospi_sincos(double x, double y, double a, double b)      // [cold>
{
	PowSinCos2 o;
	o.p = ::pow(x, y);
	::sincospi(a, &o.s1, &o.c1);
	::sincos(b, &o.s2, &o.c2);
	return o;
}                                                                                                               // <cold]

MPGOS_FM_FN __attribute__((const)) PowSinCos2 pow_sincospi_sincos(double x, double y, double a, double b)
{
	PowSinCos2 o;
	// --- reductions of the two trigonometric arguments
	double two_over_pi, p1, p2, p3, pi_hi, pi_lo;
	ldpair(FM_2OPI_PIO2A, two_over_pi, p1);
	ldpair(FM_PIO2B_PIO2C, p2, p3);
	ldpair(FM_PIHI_PILO, pi_hi, pi_lo);
	const double tb = fma(b, two_over_pi, MPGOS_FM_MAGIC);
	const double ta = fma(a, 2.0, MPGOS_FM_MAGIC);
	const double nb = tb - MPGOS_FM_MAGIC;
	const double na = ta - MPGOS_FM_MAGIC;
	double rb = fma(-nb, p1, b);
	const double ra0 = fma(na, -0.5, a);
	rb = fma(-nb, p2, rb);
	rb = fma(-nb, p3, rb);
	const double ra = fma(ra0, pi_lo, ra0 * pi_hi);
	// --- log of the base (as in pow above)
	const int hx = __double2hiint(x);
	const int h2 = hx + (0x3ff00000 - 0x3fe6a09e);
	const int e = (h2 >> 20) - 0x3ff;
	const double m = __hiloint2double((h2 & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
	const double ke = __hiloint2double(0x43300000, 0x80000000 ^ e) - 4503601774854144.0;
	const double f = m - 1.0;
	const double den = 2.0 + f;
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
	r = fma(r, fma(-den, r, 1.0), r);
	r = fma(r, fma(-den, r, 1.0), r);
	const double s = f * r;
	const double z = s * s, w = z * z;
	double lg6, lg7, lg4, lg5, lg2, lg3, lg1, l2e_hi, l2e_lo, ln2_hi, ln2_lo, c13, ca, cb;
	// --- the two sine / cosine kernels side by side with the logarithm's polynomial
	const double za = ra * ra, zb = rb * rb;
	double psa, pca, psb, pcb;
	ldpair(FM_S6C6, psa, pca);  psb = psa;  pcb = pca;
	ldpair(FM_LG6_LG7, lg6, lg7);
	ldpair(FM_S5C5, ca, cb);  psa = fma(psa, za, ca);  pca = fma(pca, za, cb);  psb = fma(psb, zb, ca);  pcb = fma(pcb, zb, cb);
	ldpair(FM_LG4_LG5, lg4, lg5);
	ldpair(FM_S4C4, ca, cb);  psa = fma(psa, za, ca);  pca = fma(pca, za, cb);  psb = fma(psb, zb, ca);  pcb = fma(pcb, zb, cb);
	ldpair(FM_LG2_LG3, lg2, lg3);
	ldpair(FM_S3C3, ca, cb);  psa = fma(psa, za, ca);  pca = fma(pca, za, cb);  psb = fma(psb, zb, ca);  pcb = fma(pcb, zb, cb);
	ldpair(FM_LG1_L2EH, lg1, l2e_hi);
	ldpair(FM_S2C2, ca, cb);  psa = fma(psa, za, ca);  pca = fma(pca, za, cb);  psb = fma(psb, zb, ca);  pcb = fma(pcb, zb, cb);
	const double t1 = w * fma(w, fma(w, lg6, lg4), lg2);
	const double t2 = z * fma(w, fma(w, fma(w, lg7, lg5), lg3), lg1);
	ldpair(FM_S1C1, ca, cb);  psa = fma(psa, za, ca);  pca = fma(pca, za, cb);  psb = fma(psb, zb, ca);  pcb = fma(pcb, zb, cb);
	const double hfsq = 0.5 * f * f;
	const double logm = f - (hfsq - s * (hfsq + (t2 + t1)));
	const double sa = fma(ra * za, psa, ra), sb = fma(rb * zb, psb, rb);
	const double cca = fma(za * za, pca, fma(za, -0.5, 1.0)), ccb = fma(zb * zb, pcb, fma(zb, -0.5, 1.0));
	{
		const int qa = __double2loint(ta), qb = __double2loint(tb);
		const bool odda = (qa & 1) != 0, oddb = (qb & 1) != 0;
		const double ua = odda ? cca : sa, va = odda ? sa : cca, ub = oddb ? ccb : sb, vb = oddb ? sb : ccb;
		o.s1 = __hiloint2double((int)((unsigned)__double2hiint(ua) ^ ((unsigned)(qa & 2) << 30)), __double2loint(ua));
		o.c1 = __hiloint2double((int)((unsigned)__double2hiint(va) ^ ((unsigned)((qa + 1) & 2) << 30)), __double2loint(va));
		o.s2 = __hiloint2double((int)((unsigned)__double2hiint(ub) ^ ((unsigned)(qb & 2) << 30)), __double2loint(ub));
		o.c2 = __hiloint2double((int)((unsigned)__double2hiint(vb) ^ ((unsigned)((qb + 1) & 2) << 30)), __double2loint(vb));
	}
	// --- y log2(x) and 2^(.) (as in pow above)
	ldpair(FM_L2EL_LN2H, l2e_lo, ln2_hi);
	const double hi = fma(logm, l2e_hi, ke);
	double lo = fma(logm, l2e_hi, ke - hi);
	lo = fma(logm, l2e_lo, lo);
	const double zh = y * hi;
	const double zl = fma(y, lo, fma(y, hi, -zh));
	const double tn = zh + MPGOS_FM_MAGIC;
	const double n = tn - MPGOS_FM_MAGIC;
	const double g = zh - n;
	ldpair(FM_LN2L_E13, ln2_lo, c13);
	const double t = fma(g, ln2_lo, g * ln2_hi);
	const double tt = t * t;
	double pe, po;
	ldpair(FM_E12_E11, pe, po);      po = fma(c13, tt, po);
	ldpair(FM_E10_E9, ca, cb);       pe = fma(pe, tt, ca);  po = fma(po, tt, cb);
	ldpair(FM_E8_E7, ca, cb);        pe = fma(pe, tt, ca);  po = fma(po, tt, cb);
	ldpair(FM_E6_E5, ca, cb);        pe = fma(pe, tt, ca);  po = fma(po, tt, cb);
	ldpair(FM_E4_E3, ca, cb);        pe = fma(pe, tt, ca);  po = fma(po, tt, cb);
	pe = fma(pe, tt, 0.5);
	double p = fma(po, t, pe);
	p = fma(p, tt, t);
	p = p + 1.0;
	p = fma(p, zl * ln2_hi, p);
	o.p = p * __hiloint2double((__double2loint(tn) + 1023) << 20, 0);
	const bool bad = ((unsigned)(hx - 0x00100000) >= 0x7fe00000u) | !(fabs(zh) <= 1000.0) | !(fabs(a) <= 1073741824.0) | !(fabs(b) <= 1.0e6);      // bitwise: one branch
	if (bad) o = cuda_pow_sincospi_sincos(x, y, a, b);
	return o;
}
__device__ __forceinline__ void pow_sincospi_sincos(double x, double y, double a, double b, double* p, double* s1, double* c1, double* s2, double* c2)
{
	const PowSinCos2 o = pow_sincospi_sincos(x, y, a, b);
	*p = o.p;  *s1 = o.s1;  *c1 = o.c1;  *s2 = o.s2;  *c2 = o.c2;
}
__device__ __forceinline__ void pow_sincospi_sincos(float x, float y, float a, float b, float* p, float* s1, float* c1, float* s2, float* c2)
{
	*p = ::powf(x, y);  ::sincospif(a, s1, c1);  ::sincosf(b, s2, c2);
}

// 1 / x for normal finite x of either sign (|x| in [2^-1020, 2^1020]); <= 1 ulp; the division otherwise.  The compiler's own
// sequence for 1.0 / x is the same Newton iteration plus a range check, a slow-path call and a correction step for correct rounding
// (~24 issue cycles against ~11).
__device__ __forceinline__ double rcp(double x)
{
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	r = fma(r, fma(-x, r, 1.0), r);
	r = fma(r, fma(-x, r, 1.0), r);
	if ((unsigned)((__double2hiint(x) & 0x7fffffff) - 0x00300000) >= 0x7fa00000u) r = cuda_rcp(x);
	return r;
}
__device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
// two reciprocals side by side: one basic block (the two Newton chains interleave), one range check for both
__device__ __forceinline__ void rcp2(double a, double b, double* ra, double* rb)
{
	double r, q;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(b));
	r = fma(r, fma(-a, r, 1.0), r);
	q = fma(q, fma(-b, q, 1.0), q);
	r = fma(r, fma(-a, r, 1.0), r);
	q = fma(q, fma(-b, q, 1.0), q);
	// (one unsigned comparison for both: the larger of the two biased exponent words; bitwise, not short-circuit: one branch)
	const unsigned ka = (unsigned)((__double2hiint(a) & 0x7fffffff) - 0x00300000), kb = (unsigned)((__double2hiint(b) & 0x7fffffff) - 0x00300000);
	if (max(ka, kb) >= 0x7fa00000u)
	{
		const Rcp2 o = cuda_rcp2(a, b);
		r = o.a;  q = o.b;
	}
	*ra = r;  *rb = q;
}
__device__ __forceinline__ void rcp2(float a, float b, float* ra, float* rb) { *ra = 1.0f / a;  *rb = 1.0f / b; }

// single precision: the CUDA functions
__device__ __forceinline__ void  sincos(float x, float* s, float* c)   { ::sincosf(x, s, c); }
__device__ __forceinline__ float sin(float x)                          { return ::sinf(x); }
__device__ __forceinline__ float cos(float x)                          { return ::cosf(x); }
__device__ __forceinline__ void  sincospi(float x, float* s, float* c) { ::sincospif(x, s, c); }
__device__ __forceinline__ float pow(float x, float y)                 { return ::powf(x, y); }

// The CUDA-library forms of the built-in systems take sin and cos of one argument from ONE ::sincos call: the CUDA library's sincos
// returns exactly the bits of its sin and its cos (14 million random arguments up to 1e300 and the special values, on the device:
// tests/test_fast_math_gpu.py::test_cuda_sincos_is_cuda_sin_and_cos pins it), so the results are those of the reference's two calls
// at one range reduction instead of two: Keller-Miksis 36.9 -> 32.8 ms, coupled bubbles 18.9 -> 17.1 ms, same checksums
// (profiles/r2_variant_ab_runs.txt).  MPGOS_B200_STD_SINCOS 0 restores the two calls.
#ifndef MPGOS_B200_STD_SINCOS
#define MPGOS_B200_STD_SINCOS 1
#endif

__device__ __forceinline__ void cuda_sincos_inline(double x, double* s, double* c) { ::sincos(x, s, c); }
__device__ __forceinline__ void cuda_sincos_inline(float x, float* s, float* c)    { *s = ::sinf(x);  *c = ::cosf(x); }      // single precision: unchanged (not checked)

// compile-time choice between these functions and CUDA's, for code that exists in both forms (the built-in systems)
template <bool FAST> struct Math
{
	template <class P> static __device__ __forceinline__ P    pow_(P x, P y)               { return mpgos_b200::fm::pow(x, y); }
	template <class P> static __device__ __forceinline__ P    sin_(P x)                    { return mpgos_b200::fm::sin(x); }
	template <class P> static __device__ __forceinline__ P    cos_(P x)                    { return mpgos_b200::fm::cos(x); }
	template <class P> static __device__ __forceinline__ void sincospi_(P x, P* s, P* c)   { mpgos_b200::fm::sincospi(x, s, c); }
	template <class P> static __device__ __forceinline__ void sincos_(P x, P* s, P* c)     { mpgos_b200::fm::sincos(x, s, c); }      // one evaluation for both
	template <class P> static __device__ __forceinline__ P    rcp_(P x)                    { return mpgos_b200::fm::rcp(x); }
	template <class P> static __device__ __forceinline__ void pow_sincospi_sincos_(P x, P y, P a, P b, P* p, P* s1, P* c1, P* s2, P* c2)
	{ mpgos_b200::fm::pow_sincospi_sincos(x, y, a, b, p, s1, c1, s2, c2); }
	template <class P> static __device__ __forceinline__ void rcp2_(P a, P b, P* ra, P* rb)      { mpgos_b200::fm::rcp2(a, b, ra, rb); }
};
template <> struct Math<false>
{
	template <class P> static __device__ __forceinline__ P    pow_(P x, P y)               { return ::pow(x, y); }
	template <class P> static __device__ __forceinline__ P    sin_(P x)                    { return ::sin(x); }
	template <class P> static __device__ __forceinline__ P    cos_(P x)                    { return ::cos(x); }
	template <class P> static __device__ __forceinline__ void sincospi_(P x, P* s, P* c)   { ::sincospi(x, s, c); }
#if MPGOS_B200_STD_SINCOS
	template <class P> static __device__ __forceinline__ void sincos_(P x, P* s, P* c)     { mpgos_b200::fm::cuda_sincos_inline(x, s, c); }
#else
	template <class P> static __device__ __forceinline__ void sincos_(P x, P* s, P* c)     { *s = ::sin(x);  *c = ::cos(x); }         // the reference's two calls
#endif
	template <class P> static __device__ __forceinline__ P    rcp_(P x)                    { return P(1.0) / x; }
	template <class P> static __device__ __forceinline__ void pow_sincospi_sincos_(P x, P y, P a, P b, P* p, P* s1, P* c1, P* s2, P* c2)
	{ *p = pow_(x, y);  sincospi_(a, s1, c1);  sincos_(b, s2, c2); }
	template <class P> static __device__ __forceinline__ void rcp2_(P a, P b, P* ra, P* rb)      { *ra = P(1.0) / a;  *rb = P(1.0) / b; }
};

} // namespace fm
} // namespace mpgos_b200

// Redirection for an unmodified system definition (see the header comment).  Function-like macros: only calls are rewritten.  The
// targets are host + device: in device code the functions above, in host code the C library (a tutorial's main() computes its
// parameters with pow and is part of the same translation unit).
namespace mpgos_b200 {
namespace fm {
namespace redirect {
#ifdef __CUDA_ARCH__
#define MPGOS_FM_PICK(fast, standard) fast
#else
#define MPGOS_FM_PICK(fast, standard) standard
#endif
// argument types as in C: float stays float, everything else (double, integers) is double
template <class A> struct Real { using type = double; };
template <> struct Real<float> { using type = float; };
template <class A, class B> struct Real2 { using type = double; };
template <> struct Real2<float, float> { using type = float; };
template <class A, class B> __host__ __device__ __forceinline__ typename Real2<A, B>::type pow(A x, B y)
{
	using R = typename Real2<A, B>::type;
	return MPGOS_FM_PICK(mpgos_b200::fm::pow((R)x, (R)y), ::pow((R)x, (R)y));
}
template <class A> __host__ __device__ __forceinline__ typename Real<A>::type sin(A x) { using R = typename Real<A>::type;  return MPGOS_FM_PICK(mpgos_b200::fm::sin((R)x), ::sin((R)x)); }
template <class A> __host__ __device__ __forceinline__ typename Real<A>::type cos(A x) { using R = typename Real<A>::type;  return MPGOS_FM_PICK(mpgos_b200::fm::cos((R)x), ::cos((R)x)); }
template <class A, class P> __host__ __device__ __forceinline__ void sincos(A x, P* s, P* c)   { MPGOS_FM_PICK(mpgos_b200::fm::sincos((P)x, s, c), ::sincos((P)x, s, c)); }
template <class A, class P> __host__ __device__ __forceinline__ void sincospi(A x, P* s, P* c) { MPGOS_FM_PICK(mpgos_b200::fm::sincospi((P)x, s, c), ::sincospi((P)x, s, c)); }
} // namespace redirect
} // namespace fm
} // namespace mpgos_b200
#ifdef MPGOS_B200_FAST_MATH
#define pow(x, y)         mpgos_b200::fm::redirect::pow(x, y)
#define sin(x)            mpgos_b200::fm::redirect::sin(x)
#define cos(x)            mpgos_b200::fm::redirect::cos(x)
#define sincos(x, s, c)   mpgos_b200::fm::redirect::sincos(x, s, c)
#define sincospi(x, s, c) mpgos_b200::fm::redirect::sincospi(x, s, c)
#endif

#endif
