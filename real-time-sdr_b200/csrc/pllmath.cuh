// Correctly-rounded float <- double-libm emulation for the PLL (host + device).
//
// The reference PLL (/root/reference/src/pll.cpp:39,49,50,52) calls glibc's DOUBLE atan2/cos/sin on
// float arguments and stores the result in a float.  Its NCO phase lives in a float that grows without
// bound, so one different last bit anywhere in the loop changes the audio for good.
// What has to be reproduced is therefore the function  float -> RN_float(RN_double(f(x))).
//
// Every function here works in two tiers:
//   fast : plain double arithmetic with a proven error of a few double ulps (own range reduction, so
//          the unbounded NCO phase never hits a slow library path); the result is accepted unless it
//          lies within kAmbigUlps double-ulps of a float rounding boundary (probability ~1e-6);
//   slow : double-double arithmetic (~2^-100), which decides those cases.
// Both tiers are pure IEEE +,-,*,/,fma, so the host build of this header (tests/, no GPU needed) and
// the sm_100a build produce the same values.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#define SDRB_UNLIKELY(x) __builtin_expect(!!(x), 0)
// Code that runs about once per 10^5 samples: out of line on the device, so that it is neither duplicated at every call
// site nor interleaved with the lines the careful path does execute.  (The careful repeat of k_pll runs cold: what it
// costs is instruction-cache misses, ~20 000 cycles per repeat measured, not arithmetic.)
#if defined(__CUDA_ARCH__)
#define SDRB_COLD __device__ __noinline__
#elif defined(__CUDACC__)
#define SDRB_COLD __host__ __device__ inline
#else
#define SDRB_COLD inline
#endif
#if defined(__CUDACC__)
#define SDRB_HD __host__ __device__ __forceinline__
#else
#define SDRB_HD inline
#endif

namespace sdrb {
namespace cr {

// ---- exact float ops: never contracted into FMA (the reference build has no FMA) ----
SDRB_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;  // host builds also use -ffp-contract=off; volatile is belt and braces
    return r;
#endif
}
SDRB_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
SDRB_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
SDRB_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
SDRB_HD double dfma(double a, double b, double c) { return fma(a, b, c); }

SDRB_HD uint64_t dbits(double v) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
#endif
}

// ---- double-double ----
struct dd {
    double hi, lo;
};
SDRB_HD dd two_sum(double a, double b) {
    double s = dadd(a, b);
    double bb = dadd(s, -a);
    double e = dadd(dadd(a, -dadd(s, -bb)), dadd(b, -bb));
    return dd{s, e};
}
SDRB_HD dd fast_two_sum(double a, double b) {  // |a| >= |b|
    double s = dadd(a, b);
    double e = dadd(b, -dadd(s, -a));
    return dd{s, e};
}
SDRB_HD dd two_prod(double a, double b) {
    double p = dmul(a, b);
    return dd{p, dfma(a, b, -p)};
}
SDRB_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = dadd(s.lo, t.hi);
    s = fast_two_sum(s.hi, s.lo);
    s.lo = dadd(s.lo, t.lo);
    return fast_two_sum(s.hi, s.lo);
}
SDRB_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = dadd(s.lo, a.lo);
    return fast_two_sum(s.hi, s.lo);
}
SDRB_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
SDRB_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = dadd(p.lo, dadd(dmul(a.hi, b.lo), dmul(a.lo, b.hi)));
    return fast_two_sum(p.hi, p.lo);
}
SDRB_HD dd dd_div_d(dd a, double d) {
    double q1 = a.hi / d;
    dd p = two_prod(q1, d);
    double rem = dadd(dadd(dadd(a.hi, -p.hi), -p.lo), a.lo);
    double q2 = rem / d;
    return fast_two_sum(q1, q2);
}
SDRB_HD dd dd_div(dd a, dd b) {
    double q1 = a.hi / b.hi;
    dd r = dd_add(a, dd_neg(dd_mul(b, dd{q1, 0.0})));
    double q2 = r.hi / b.hi;
    r = dd_add(r, dd_neg(dd_mul(b, dd{q2, 0.0})));
    double q3 = r.hi / b.hi;
    dd q = fast_two_sum(q1, q2);
    return dd_add_d(q, q3);
}

// ---- full-precision constants of the hot path ----
// A 64-bit floating-point immediate does not exist in SASS: written as literals, these constants are rebuilt with two
// register moves each, in every step of the PLL recurrence (a sixth of its instruction stream).  On the device they
// live in constant memory instead, where DFMA/DMUL/DADD read them as operands; the host build uses the same table.
#define SDRB_PLL_CONSTS                                                                                         \
    {8.33333333332248946124e-03, -1.66666666666666324348e-01, 2.75573137070700676789e-06, -1.98412698298579493134e-04, \
     1.58969099521155010221e-10, -2.50507602534068634195e-08, -1.38888888888741095749e-03, 4.16666666666666019037e-02, \
     -2.75573143513906633035e-07, 2.48015872894767294178e-05, -1.13596475577881948265e-11, 2.08757232129817482790e-09, \
     0x1.45f306dc9c883p-1, 0x1.1a62633145c07p-54, 0x1.921fb54442d18p+0}
enum PllConst { kS2, kS1, kS4, kS3, kS6, kS5, kC2, kC1, kC4, kC3, kC6, kC5, kK2OverPi, kKPio2M, kKPio2H };
#if defined(__CUDACC__)
__device__ __constant__ double c_pll_consts[15] = SDRB_PLL_CONSTS;
#endif
static const double h_pll_consts[15] = SDRB_PLL_CONSTS;
SDRB_HD double K(PllConst i) {
#if defined(__CUDA_ARCH__)
    return c_pll_consts[i];
#else
    return h_pll_consts[i];
#endif
}

// Leaner kernels for the speculative steps of the batched loop: sin r = r + r z (L1 + .. + L5 z^4), cos r = 1 - z/2 +
// z^2 (M1 + .. + M5 z^4) on |r| <= pi/4, one coefficient less than fdlibm's each (Remez fits, tests/gen_pllmath_consts.py;
// relative error below 2^-46.3 and 2^-50 including the rounding of the evaluation, checked against mpmath in tests/test_pllmath.py).  The FP64 pipe issues one warp
// instruction per 2.3 cycles and the two kernels are its busiest stretch of the recurrence: every operation less is that
// much off the chain.  The errors stay inside what the acceptance tests of the speculative step allow (2^-45 for the
// float roundings of sa / cr with the 512-ulp tie window used there, 2^-45 absolute for e).
#define SDRB_PLL_LEAN_CONSTS                                                                                     \
    {0x1.1111110cdbeb5p-7, -0x1.5555555552e41p-3, 0x1.71d752f9f8bdfp-19, -0x1.a019f946a7019p-13, -0x1.a950938183dcbp-26, \
     0.0, -0x1.6c16c169ae93cp-10, 0x1.5555555554a28p-5, -0x1.27e1089e1a501p-22, 0x1.a019fcd9727a0p-16, 0x1.1c065229821d1p-29, \
     0.0, 0x1.45f306dc9c883p-1, 0x1.1a62633145c07p-54, 0x1.921fb54442d18p+0}
// same layout as PllConst: kS2 = L2, kS1 = L1, kS4 = L4, kS3 = L3, kS6 = L5 (!), kS5 unused, kC2 = M2, kC1 = M1, kC4 = M4, kC3 = M3,
// kC6 = M5 (!), kC5 unused
enum PllLeanConst { kL5 = kS6, kM5 = kC6 };
#if defined(__CUDACC__)
__device__ double g_pll_lean_consts[15] = SDRB_PLL_LEAN_CONSTS;
#endif
static const double h_pll_lean_consts[15] = SDRB_PLL_LEAN_CONSTS;

// The same constants pinned in registers for the batched PLL kernel: left to itself the compiler re-reads the
// constant bank inside the loop (28 LDC per 8 samples, each an issue slot of an in-order warp that has none to spare).
// They are read once through a volatile global load, which cannot be rematerialised.
struct PllK {
    double v[15];
};
#if defined(__CUDACC__)
__device__ double g_pll_consts[15] = SDRB_PLL_CONSTS;
#endif
SDRB_HD void pll_k_load(PllK& kk) {
#pragma unroll
    for (int i = 0; i < 15; i++) {
#if defined(__CUDA_ARCH__)
        asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(kk.v[i]) : "l"(&g_pll_consts[i]));
#else
        kk.v[i] = h_pll_consts[i];
#endif
    }
}
SDRB_HD void pll_k_load_lean(PllK& kk) {  // the lean kernels' coefficients (sincos_poly2_lean) and the reduction constants
#pragma unroll
    for (int i = 0; i < 15; i++) {
        if (i == kS5 || i == kC5) {
            kk.v[i] = 0.0;
            continue;
        }
#if defined(__CUDA_ARCH__)
        asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(kk.v[i]) : "l"(&g_pll_lean_consts[i]));
#else
        kk.v[i] = h_pll_lean_consts[i];
#endif
    }
}

// ---- float rounding boundary test ----
// A double v rounds to float by dropping its low 29 mantissa bits; the boundary (tie) pattern of those
// bits is 0x10000000.  `true` means: v is so close to a boundary (or so small that the float is
// subnormal and the pattern does not apply) that a few-ulp error in v could change the float.
constexpr uint32_t kAmbigUlps = 256;
SDRB_HD bool near_float_boundary(double v) {
    uint64_t b = dbits(v);
    uint32_t e = (uint32_t)(b >> 52) & 0x7FFu;
    if (e == 0 && (b << 1) == 0) return false;      // exact zero
    if (e < 1023u - 126u + 1u) return true;         // float-subnormal range: let the slow tier decide
    uint32_t low = (uint32_t)b & 0x1FFFFFFFu;
    uint32_t dist = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    return dist <= kAmbigUlps;
}

// ---- pi/2 in pieces (generated with mpmath, tests/gen_pllmath_consts.py) ----
// P1..P5 carry 22 significant bits each, so k*Pi is exact in double for |k| < 2^31.
constexpr double kTwoOverPi = 0x1.45f306dc9c883p-1;
constexpr double kP1 = 0x1.921fb00000000p+0;
constexpr double kP2 = 0x1.5110b00000000p-22;
constexpr double kP3 = 0x1.1846980000000p-44;
constexpr double kP4 = 0x1.3198a00000000p-69;
constexpr double kP5 = 0x1.701b800000000p-92;
constexpr double kPTailHi = 0x1.cd129024e088ap-115;
constexpr double kPTailLo = 0x1.9f31d0082efaap-169;
constexpr double kP4Rest = 0x1.3198a2e037073p-69;  // pi/2 - P1 - P2 - P3 rounded to 53 bits
constexpr double kPiH = 0x1.921fb54442d18p+1, kPiM = 0x1.1a62633145c07p-53, kPiL = -0x1.f1976b7ed8fbcp-109;
constexpr double kPio2H = 0x1.921fb54442d18p+0, kPio2M = 0x1.1a62633145c07p-54, kPio2L = -0x1.f1976b7ed8fbcp-110;
// |x| below this uses the reductions of this file.  2^45 rad is 1.8 years of the 114 kHz loop's NCO phase (2.98 rad per
// sample at 240 kS/s); round 1 stopped at 3e9 (70 minutes), after which every step went through the library.
constexpr double kReduceLimit = 0x1p45;
// Smallest |r| for which the double-precision reductions below are trusted, given the quadrant count k: their absolute
// error is ~|k| 2^-106 (the rounding of an intermediate of size |k| (pi/2 - PIO2H) and the part of pi/2 beyond 107 bits),
// which has to stay below 2^-48 |r| for the acceptance tests to mean what they say.  Up to |k| = 2^28 that is the 2^-30 no
// float comes closer than to a multiple of pi/2 (checked exhaustively below 3e9, tests/test_pllmath.py); beyond, arguments
// this close go to the double-double tier (probability ~|k| 2^-58 per argument).
SDRB_HD double reduce_rmin(double kd) { return fmax(0x1p-30, dmul(fabs(kd), 0x1p-58)); }

// ---- sin / cos ----
// fast tier: x - k*pi/2 with four fma steps.  Steps 1 and 2 are exact (x is a float, k*P1 and k*P2
// are exact and the differences fit in 53 bits); step 3 is exact whenever |r| < 2^-12 and otherwise
// rounds with relative error 2^-53; step 4 adds k*2^-122 absolute.  |r| stays >= 2^-30 for every
// float below kReduceLimit (checked exhaustively, tests/test_pllmath.py), so r is good to ~2^-53
// relative; the callers still send |r| < 2^-30 to the slow tier.
// sin r and cos r for |r| <= pi/4: the degree-13 / degree-12 minimax kernels of fdlibm (k_sin.c, k_cos.c: S1..S6,
// C1..C6, approximation error below 2^-58), evaluated Estrin-style so the dependent depth is z, z^2, z^4 and two fmas.
// Far inside the 2^-45 the callers allow before they consult the double-double tier.
SDRB_HD void sincos_poly2(double r, double rs, double& sr, double& cr_) {  // sr = sin(rs), rs = r or |r|
    // Dependent depth 4 (z; z^2, the coefficient pairs, r z, 1 - z/2; three partial sums; the result): one level less
    // than the plain Estrin form, and every level is 8 cycles of the PLL recurrence.
    const double z = dmul(r, r);
    const double z2 = dmul(z, z);
    // sin r = r + r z (S1 + S2 z) + (r z z^2) ((S3 + S4 z) + z^2 (S5 + S6 z))
    const double s12 = dfma(K(kS2), z, K(kS1));
    const double s34 = dfma(K(kS4), z, K(kS3));
    const double s56 = dfma(K(kS6), z, K(kS5));
    const double rz = dmul(rs, z);
    const double sA = dfma(rz, s12, rs);
    const double sB = dmul(rz, z2);
    const double sC = dfma(z2, s56, s34);
    sr = dfma(sB, sC, sA);
    // cos r = (1 - z/2) + z^2 (C1 + C2 z) + z^4 ((C3 + C4 z) + z^2 (C5 + C6 z))
    const double c12 = dfma(K(kC2), z, K(kC1));
    const double c34 = dfma(K(kC4), z, K(kC3));
    const double c56 = dfma(K(kC6), z, K(kC5));
    const double ch = dfma(-0.5, z, 1.0);
    const double cA = dfma(z2, c12, ch);
    const double z4 = dmul(z2, z2);
    const double cC = dfma(z2, c56, c34);
    cr_ = dfma(z4, cC, cA);
}
SDRB_HD void sincos_poly(double r, double& sr, double& cr_) { sincos_poly2(r, r, sr, cr_); }
// sincos_poly2 with the coefficients from registers (PllK), same operations in the same order
SDRB_HD void sincos_poly2k(double r, double rs, double& sr, double& cr_, const PllK& kk) {
    const double z = dmul(r, r);
    const double z2 = dmul(z, z);
    const double s12 = dfma(kk.v[kS2], z, kk.v[kS1]);
    const double s34 = dfma(kk.v[kS4], z, kk.v[kS3]);
    const double s56 = dfma(kk.v[kS6], z, kk.v[kS5]);
    const double rz = dmul(rs, z);
    const double sA = dfma(rz, s12, rs);
    const double sB = dmul(rz, z2);
    const double sC = dfma(z2, s56, s34);
    sr = dfma(sB, sC, sA);
    const double c12 = dfma(kk.v[kC2], z, kk.v[kC1]);
    const double c34 = dfma(kk.v[kC4], z, kk.v[kC3]);
    const double c56 = dfma(kk.v[kC6], z, kk.v[kC5]);
    const double ch = dfma(-0.5, z, 1.0);
    const double cA = dfma(z2, c12, ch);
    const double z4 = dmul(z2, z2);
    const double cC = dfma(z2, c56, c34);
    cr_ = dfma(z4, cC, cA);
}
// the lean kernels (coefficients from pll_k_load_lean): 16 operations instead of 18, same dependent depth
SDRB_HD void sincos_poly2_lean(double r, double rs, double& sr, double& cr_, const PllK& kk) {
    const double z = dmul(r, r);
    const double z2 = dmul(z, z);
    const double rz = dmul(rs, z);
    const double s12 = dfma(kk.v[kS2], z, kk.v[kS1]);
    const double s34 = dfma(kk.v[kS4], z, kk.v[kS3]);
    const double sA = dfma(rz, s12, rs);
    const double sB = dmul(rz, z2);
    const double sC = dfma(z2, kk.v[kL5], s34);
    sr = dfma(sB, sC, sA);
    const double c12 = dfma(kk.v[kC2], z, kk.v[kC1]);
    const double c34 = dfma(kk.v[kC4], z, kk.v[kC3]);
    const double ch = dfma(-0.5, z, 1.0);
    const double cA = dfma(z2, c12, ch);
    const double z4 = dmul(z2, z2);
    const double cC = dfma(z2, kk.v[kM5], c34);
    cr_ = dfma(z4, cC, cA);
}
SDRB_HD double flip_sign_if(double v, unsigned flip) {  // exact negation by a sign-bit XOR (one integer op on the chain)
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)(flip << 31), __double2loint(v));
#else
    uint64_t b = dbits(v) ^ ((uint64_t)flip << 63);
    double o;
    memcpy(&o, &b, 8);
    return o;
#endif
}
SDRB_HD void sincos_quadrant(int q, double sr, double cr_, double& s, double& c) {
    double ss = (q & 1) ? cr_ : sr;
    double cs = (q & 1) ? sr : cr_;
    s = flip_sign_if(ss, ((unsigned)q >> 1) & 1u);
    c = flip_sign_if(cs, ((unsigned)(q + 1) >> 1) & 1u);
}
SDRB_HD void sincos_fast(double x, double& s, double& c, bool& tiny) {
    double kd = rint(dmul(x, kTwoOverPi));
    double r = dfma(-kd, kP1, x);
    r = dfma(-kd, kP2, r);
    r = dfma(-kd, kP3, r);
    r = dfma(-kd, kP4Rest, r);
    tiny = (kd != 0.0) && fabs(r) < reduce_rmin(kd);  // kd == 0: r = x exactly
    double sr, cr_;
    sincos_poly(r, sr, cr_);
    sincos_quadrant((int)((long long)kd & 3), sr, cr_, s, c);
}

// slow tier: the same k, the reduction carried in double-double with pi/2 to ~170 bits, Taylor series
// in double-double.  Returns RN_double(sin x), RN_double(cos x) up to ~2^-100.
SDRB_COLD void sincos_slow(double x, double& s, double& c) {
    double kd = rint(dmul(x, kTwoOverPi));
    double r2 = dfma(-kd, kP2, dfma(-kd, kP1, x));  // exact: x - k P1 has at most log2|k| + 1 significant bits, then k P2 lines up
    dd r = dd_add(dd{r2, 0.0}, dd_neg(two_prod(kd, kP3)));  // exact products (53 bits only hold them for |k| < 2^31)
    r = dd_add(r, dd_neg(two_prod(kd, kP4)));
    r = dd_add(r, dd_neg(two_prod(kd, kP5)));
    r = dd_add(r, dd_neg(two_prod(kd, kPTailHi)));
    r = dd_add_d(r, -dmul(kd, kPTailLo));
    dd z = dd_mul(r, r);
    dd ts = r, ssum = r;            // sin terms r^(2j+1)/(2j+1)!
    dd tc = dd{1.0, 0.0}, csum = tc;  // cos terms r^(2j)/(2j)!
    for (int j = 1; j <= 15; j++) {
        tc = dd_div_d(dd_mul(tc, z), (double)((2 * j - 1) * (2 * j)));
        ts = dd_div_d(dd_mul(ts, z), (double)((2 * j) * (2 * j + 1)));
        if (j & 1) {
            csum = dd_add(csum, dd_neg(tc));
            ssum = dd_add(ssum, dd_neg(ts));
        } else {
            csum = dd_add(csum, tc);
            ssum = dd_add(ssum, ts);
        }
    }
    int q = (int)((long long)kd & 3);
    double ss = (q & 1) ? csum.hi : ssum.hi;
    double cs = (q & 1) ? ssum.hi : csum.hi;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cs : cs;
}

// (float)sin((double)t), (float)cos((double)t) as glibc + the C++ float conversion produce them.
SDRB_HD void sincos_f(float t, float& s, float& c) {
    double x = (double)t;
    if (!(fabs(x) < kReduceLimit)) {  // also NaN/Inf: no own reduction, defer to the library
        s = (float)sin(x);
        c = (float)cos(x);
        return;
    }
    if (t == 0.0f) {  // keeps the sign of zero: sin(-0) = -0
        s = t;
        c = 1.0f;
        return;
    }
    double ds, dc;
    bool tiny;
    sincos_fast(x, ds, dc, tiny);
    if (tiny || near_float_boundary(ds) || near_float_boundary(dc)) sincos_slow(x, ds, dc);
    s = (float)ds;
    c = (float)dc;
}
SDRB_HD float cos_f(float t) {
    double x = (double)t;
    if (!(fabs(x) < kReduceLimit)) return (float)cos(x);
    double ds, dc;
    bool tiny;
    sincos_fast(x, ds, dc, tiny);
    if (tiny || near_float_boundary(dc)) sincos_slow(x, ds, dc);
    return (float)dc;
}

// ---- atan2 ----
// atan(i/16), i = 0..16, as double-double (mpmath).
struct AtanTab {
    double hi[17], lo[17];
};
#define SDRB_ATAN_TAB_INIT                                                                                          \
    {{0x0.0p+0, 0x1.ff55bb72cfdeap-5, 0x1.fd5ba9aac2f6ep-4, 0x1.7b97b4bce5b02p-3, 0x1.f5b75f92c80ddp-3,              \
      0x1.362773707ebccp-2, 0x1.6f61941e4def1p-2, 0x1.a64eec3cc23fdp-2, 0x1.dac670561bb4fp-2, 0x1.0657e94db30d0p-1, \
      0x1.1e00babdefeb4p-1, 0x1.345f01cce37bbp-1, 0x1.4978fa3269ee1p-1, 0x1.5d58987169b18p-1, 0x1.700a7c5784634p-1, \
      0x1.819d0b7158a4dp-1, 0x1.921fb54442d18p-1},                                                                   \
     {0x0.0p+0, -0x1.c934d86d23f1dp-60, -0x1.cd37686760c17p-59, 0x1.347b0b4f881cap-58, 0x1.8ab6e3cf7afbdp-57,        \
      -0x1.963a544b672d8p-57, -0x1.c63aae6f6e918p-56, -0x1.24dec1b50b7ffp-56, 0x1.a2b7f222f65e2p-56,                 \
      -0x1.d5b495f6349e6p-56, -0x1.928df287a668fp-58, 0x1.1021137c71102p-55, 0x1.2419a87f2a458p-56,                  \
      0x1.0028e4bc5e7cap-57, -0x1.8c34d25aadef6p-56, -0x1.bf76229d3b917p-56, 0x1.1a62633145c07p-55}}

// Reduced problem shared by both tiers: atan(num0/den0), 0 <= num0 <= den0, via
//   atan(t) = atan(c) + atan((num0 - c den0)/(den0 + c num0)),  c = i/16 nearest to t.
// num0 and den0 are floats held in doubles, c has 5 bits, so both linear combinations are exact.
struct AtanRed {
    int i;
    double nn, dn;
};
SDRB_HD AtanRed atan_reduce(double num0, double den0) {
    float est = (float)num0 / (float)den0;  // any estimate within ~1e-3 of t works
    int i = (int)(est * 16.0f + 0.5f);
    i = i < 0 ? 0 : (i > 16 ? 16 : i);
    double c = (double)i * 0.0625;
    return AtanRed{i, dfma(-c, den0, num0), dfma(c, num0, den0)};
}

SDRB_HD double atan2_fast(double ax, double ay, bool xneg, const AtanTab& tab) {
    bool swap = ay > ax;
    AtanRed rd = atan_reduce(swap ? ax : ay, swap ? ay : ax);
    double z = rd.nn / rd.dn;
    double z2 = dmul(z, z), z4 = dmul(z2, z2);
    // atan z = z + z z2 (A1 + A2 z2 + ... + A5 z2^4), Aj = (-1)^j/(2j+1); |z| <= 1/32 + eps
    double pa = dfma(0x1.999999999999ap-3, z2, -0x1.5555555555555p-2);   //  1/5, -1/3
    double pb = dfma(0x1.c71c71c71c71cp-4, z2, -0x1.2492492492492p-3);   //  1/9, -1/7
    double poly = dfma(z4, dfma(-0x1.745d1745d1746p-4, z4, pb), pa);  // pa + z4*(pb + z4*A5)
    double at = dfma(dmul(z, z2), poly, z);
    double res = dadd(tab.hi[rd.i], dadd(tab.lo[rd.i], at));
    if (swap) res = dadd(dadd(kPio2H, -res), kPio2M);
    if (xneg) res = dadd(dadd(kPiH, -res), kPiM);
    return res;
}

SDRB_COLD double atan2_slow(double ax, double ay, bool xneg, const AtanTab& tab) {
    bool swap = ay > ax;
    AtanRed rd = atan_reduce(swap ? ax : ay, swap ? ay : ax);
    dd z = dd_div(dd{rd.nn, 0.0}, dd{rd.dn, 0.0});
    dd z2 = dd_mul(z, z);
    dd term = z, sum = z;
    for (int j = 1; j <= 13; j++) {  // |z|^27/27 < 2^-130
        term = dd_mul(term, z2);
        dd t = dd_div_d(term, (double)(2 * j + 1));
        sum = (j & 1) ? dd_add(sum, dd_neg(t)) : dd_add(sum, t);
    }
    dd res = dd_add(dd{tab.hi[rd.i], tab.lo[rd.i]}, sum);
    if (swap) res = dd_add_d(dd_add(dd{kPio2H, kPio2M}, dd_neg(res)), kPio2L);
    if (xneg) res = dd_add_d(dd_add(dd{kPiH, kPiM}, dd_neg(res)), kPiL);
    return res.hi;
}

// (float)atan2((double)y, (double)x) as glibc + the float conversion produce it.
SDRB_HD float atan2_f(float yf, float xf, const AtanTab& tab) {
    const float kPiF = 3.14159274101257324f, kPio2F = 1.57079637050628662f;  // (float)pi, (float)(pi/2)
    if (yf != yf || xf != xf || fabsf(yf) == INFINITY || fabsf(xf) == INFINITY)
        return (float)atan2((double)yf, (double)xf);
    if (yf == 0.0f) {  // atan2(+-0, x): +-0 for x > 0 or +0, +-pi for x < 0 or -0
        bool xn = signbit(xf);
        return xn ? copysignf(kPiF, yf) : yf;
    }
    if (xf == 0.0f) return copysignf(kPio2F, yf);
    double ax = fabs((double)xf), ay = fabs((double)yf);
    bool xneg = xf < 0.0f;
    double v = atan2_fast(ax, ay, xneg, tab);
    if (near_float_boundary(v)) v = atan2_slow(ax, ay, xneg, tab);
    float r = (float)v;
    return yf < 0.0f ? -r : r;
}

// ---- one PLL sample, /root/reference/src/pll.cpp:34-53 ----
struct PllCoef {
    float Kp, Ki;       // :8-9
    double w;           // (2*PI) * (double)(freq/Fs), the loop-invariant head of :47
    float ncoScale, phaseAdjust;
};
struct PllState {
    float feedbackI, feedbackQ, integrator, phaseEst;
    double trigOffset;
};
SDRB_HD PllCoef pll_coef(float freq, float Fs, float ncoScale, float phaseAdjust, float normBandwidth) {
    const float Cp = 2.666f, Ci = 3.555f;
    PllCoef c;
    c.Kp = fmul(normBandwidth, Cp);
    c.Ki = fmul(fmul(normBandwidth, normBandwidth), Ci);
    c.w = dmul(2 * 3.14159265358979323846, (double)(freq / Fs));
    c.ncoScale = ncoScale;
    c.phaseAdjust = phaseAdjust;
    return c;
}
// The recurrence proper (:36-50): consumes one input sample, updates the loop state and returns trigArg.
// The NCO output of :52 is a pure function of trigArg (nco_out below), so callers that batch many
// streams keep it out of the sequential loop.
SDRB_HD float pll_step_trig(float in, PllState& st, const PllCoef& k, const AtanTab& tab) {
    float errorI = fmul(in, st.feedbackI);
    float errorQ = fmul(in, -st.feedbackQ);
    float errorD = atan2_f(errorQ, errorI, tab);
    st.integrator = fadd(st.integrator, fmul(k.Ki, errorD));
    st.phaseEst = fadd(fadd(st.phaseEst, fmul(k.Kp, errorD)), st.integrator);
    st.trigOffset = dadd(st.trigOffset, 1.0);
    float trigArg = (float)dadd(dmul(k.w, st.trigOffset), (double)st.phaseEst);
    sincos_f(trigArg, st.feedbackQ, st.feedbackI);
    return trigArg;
}
SDRB_HD float nco_out(float trigArg, const PllCoef& k) { return cos_f(fadd(fmul(trigArg, k.ncoScale), k.phaseAdjust)); }
SDRB_HD float pll_step(float in, PllState& st, const PllCoef& k, const AtanTab& tab) {
    return nco_out(pll_step_trig(in, st, k, tab), k);
}

// ---- the same recurrence with a short dependency chain (what the batched kernel runs) ----
// (cos_lean_f, used by the parallel NCO-output kernel, is defined after sincos_reduced below)
//
// The phase detector's inputs are errorI = RN(in*feedbackI), errorQ = RN(in*-feedbackQ) with
// feedbackI/Q = RN_float(cos/sin(theta)) of the PREVIOUS step, so atan2(errorQ, errorI) equals
// -theta (+pi when in < 0), wrapped to (-pi, pi], plus a perturbation delta of at most ~2^-23 rad that
// comes from the four float roundings.  Rotating (errorI, errorQ) by +theta with the double-precision
// cos/sin the previous step already produced gives  delta = atan(u/v),  u = errorI*sin + errorQ*cos,
// v = errorI*cos - errorQ*sin = in*(1 + O(2^-23));  for |delta| < 2^-22, atan(u/v) = u/in up to 2^-45.
// That replaces a table-driven atan2 with a double division by two multiplies and three adds.  The result
// is accepted only if it is farther than kAtanAbsTol from a float rounding boundary, otherwise (and for
// zero / subnormal / inconsistent inputs) the general correctly-rounded atan2_f above decides.
//
// State between steps: theta = r + kq*pi/2 (mod 2 pi) with |r| <= pi/4, and sa = |sin r|, cr = cos r in double.
// sin/cos(theta) are +-sa / +-cr in an order and with signs that depend on kq and the sign of r only.
struct PllFast {
    float fbI, fbQ, integ, phase;  // pllblock_args fields (/root/reference/include/pll.h:10-17)
    double trigOffset;
    double sa, cr;  // |sin r| and cos r of the current NCO phase, double
    double r;       // that phase reduced
    int kq;
    bool generic_next;  // sa/cr/r/kq are not valid (fbI/fbQ are): use the general atan2 at the next step
    double magic;       // 1.5 * 2^(E+29), E = binade of the current NCO phase: td + magic - magic rounds td to float precision
    uint32_t rmin_hi;   // high word of the smallest |r| the speculative step accepts in that binade (reduce_rmin, rounded up)
    uint32_t texp;      // exponent field (high word, bits 20-30) of that binade: what the unrounded phase td must show
};
SDRB_HD uint32_t dhi(double v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(v);
#else
    return (uint32_t)(dbits(v) >> 32);
#endif
}
SDRB_HD uint32_t dlo(double v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(v);
#else
    return (uint32_t)dbits(v);
#endif
}
SDRB_HD double mkd(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t b = ((uint64_t)hi << 32) | lo;
    double o;
    memcpy(&o, &b, 8);
    return o;
#endif
}
SDRB_HD uint32_t fbits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(v);
#else
    uint32_t b;
    memcpy(&b, &v, 4);
    return b;
#endif
}
SDRB_HD float bitsf(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float v;
    memcpy(&v, &b, 4);
    return v;
#endif
}
// 1.5 * 2^(E+29) for the binade E of v: adding it to a number of that binade leaves exactly 24 significant bits (RNE).
SDRB_HD double float_round_magic(double v) { return mkd(((dhi(v) & 0x7FF00000u) + (29u << 20)) | (1u << 19), 0u); }
// reduce_rmin for every argument of the binade of v, as the high word of a power of two: |k| < 2^(E+1) 2/pi, so 2^(E-57)
SDRB_HD uint32_t reduce_rmin_hi(double v) {
    const uint32_t t = (dhi(v) & 0x7FF00000u) - (57u << 20);
    return ((dhi(v) & 0x7FF00000u) > (57u << 20) && t > 0x3E100000u) ? t : 0x3E100000u;
}
constexpr double kMagicRint = 6755399441055744.0;  // 1.5 * 2^52: x + magic - magic = rint(x), integer in the low word
constexpr int kAtanTolLog2 = -43;                  // absolute error bound of the rotated phase detector, see above

// Conversions without the conversion unit.  The recurrence below crosses float <-> double four times per sample and an
// F2F costs ~19 cycles each way (and 8 cycles of a pipe that two conversions in a row have to share); for operands
// whose sign and range are known the same result takes two integer operations.
//   d2f_known: RN_float(v) for a double that is NOT within the rejection distance of a float rounding tie (the callers
//   test that separately), float-normal, and whose top nibble N (sign, three high exponent bits) is known:
//   K = (((N & 1) ^ (N >> 3)) << 31) + 0x40000000.  (v >> 28 keeps one guard bit; +1, >>1 rounds half up.)
SDRB_HD uint32_t d2f_known(double v, uint32_t K) {
    const uint32_t v28 = (dhi(v) << 4) | (dlo(v) >> 28);
    return K + ((v28 + 1u) >> 1);
}
SDRB_HD uint32_t d2f_K(uint32_t hi) {
    uint32_t K = (((hi << 3) ^ hi) & 0x80000000u) | 0x40000000u;
#if defined(__CUDA_ARCH__)
    asm("" : "+r"(K));  // one register, made beside the chain (the compiler would otherwise add the constant last, on it)
#endif
    return K;
}
//   f2d_pos: (double)p for a positive normal float.
SDRB_HD double f2d_pos(float p) {
    const uint32_t b = fbits(p);
    return mkd(0x38000000u + (b >> 3), b << 29);
}

// ---- (float)((double)num / den) for the FM discriminator (/root/reference/src/demod.cpp:11,17) ----
// The reference divides in double and stores a float; an exact double division costs the FP64 pipe a reciprocal seed and
// a dozen dependent operations.  Here: q0 = num * rc with rc ~ 1/den good to 2^-21 (a single-precision reciprocal of den's
// leading bits), one residual and one correction in double (q1 = q0 + (num - q0 den) rc, off by less than 2^-42 q1), and
// the float rounding of q1 is the reference's unless q1 lies within 2048 double-ulps of a float rounding tie (8e-6 of all
// quotients) or anything is out of the float-normal range: those return false and the caller divides exactly.
//   den_approx_f: den (positive, 2^-120 < den < 2^120, checked here) as a float, truncated: one funnel shift and one XOR.
SDRB_HD bool fm_den_in_range(double den) { return (((dhi(den) >> 20) & 0xFFFu) - (1023u - 120u)) < 240u; }  // sign bit included: negative / NaN / inf fail
SDRB_HD float fm_den_approx_f(double den) {
    const uint32_t hi = dhi(den), lo = dlo(den);
    return bitsf((((hi << 3) | (lo >> 29)) & 0x7FFFFFFFu) ^ 0x40000000u);
}
SDRB_HD bool fm_quotient_fast(float num, double den, float rc, float& out) {
    const double rd = f2d_pos(rc);
    const double nd = (double)num;
    const double q0 = dmul(nd, rd);
    const double e = dfma(-q0, den, nd);
    const double q1 = dfma(e, rd, q0);
    const uint32_t E = (dhi(q1) >> 20) & 0x7FFu;                                        // float-normal result: 2^-125 <= |q1| < 2^126
    const bool range_ok = (E - (1023u - 125u)) < 251u;
    const bool tie = (((dlo(q1) & 0x1FFFFFFFu) + 2048u) & 0x1FFFF000u) == 0x10000000u;  // low 29 bits in [tie - 2048, tie + 2047]
    out = (float)q1;
    return range_ok && !tie;
}

// v is within 2^kAtanTolLog2 of a float rounding boundary (or too small for the bound to mean anything)
SDRB_HD bool near_float_boundary_abs(double v) {
    uint64_t b = dbits(v);
    int e = (int)((b >> 52) & 0x7FFu) - 1023;
    if (e < -17 || e > 1) return true;
    uint32_t low = (uint32_t)b & 0x1FFFFFFFu;  // units of 2^(e-52)
    uint32_t dist = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    return dist <= (1u << (52 + kAtanTolLog2 - e));
}

// Quarter-turn reduction with two fused steps: r = (x - k*PIO2H) - k*PIO2M.  Each fma rounds once, relative to a
// result of the size of r, and PIO2H + PIO2M is pi/2 to 2^-107, so r is good to ~2^-52 relative for every k < 2^31
// as long as |r| >= 2^-30 (which holds for every float argument below kReduceLimit, see sincos_fast).
// Returns |sin r| and cos r; false if the result must not be trusted (tiny r).
SDRB_HD bool sincos_reduce2(double x, double& sa, double& cr_, double& r_out, int& q_out) {
    const double tm = dfma(x, kTwoOverPi, kMagicRint);
    const double kd = dadd(tm, -kMagicRint);
    const int q = (int)dlo(tm) & 3;
    const double r = dfma(-kd, kPio2M, dfma(-kd, kPio2H, x));
    const double ra = fabs(r);
    sincos_poly2(r, ra, sa, cr_);
    r_out = r;
    q_out = q;
    return !(ra < reduce_rmin(kd));
}
// (sin theta, cos theta) from the reduced pieces
SDRB_HD void pll_cs(double sa, double cr_, double r, int kq, double& s0, double& c0) {
    sincos_quadrant(kq, r < 0.0 ? -sa : sa, cr_, s0, c0);
}

// cos_f out of line on the device: the callers below reach it about once per million arguments, and inlined at every
// call site its double-double tier made k_mix five times larger than its hot path (instruction-cache misses)
#if defined(__CUDA_ARCH__)
__device__ __noinline__ float cos_f_rare(float t) { return cos_f(t); }
#else
inline float cos_f_rare(float t) { return cos_f(t); }
#endif
// cos_f for the NCO-output kernel (k_mix), which is bound by the FP64 pipe: only the kernel the quadrant asks for is
// evaluated, as ONE Horner chain whose coefficients are selected by the quadrant's parity:
//   cos(r + q pi/2) = +-cos r = +-(1 + z (-1/2 + M1 z + .. + M5 z^5))          q even
//                   = -+sin r = -+(r + (r z) (L1 + L2 z + .. + L5 z^4))        q odd
// with the lean coefficients of the speculative PLL step (relative error below 2^-46.3 / 2^-50): 8 FP64 operations after
// the reduction instead of the 18 of both kernels.  The result is accepted when it is farther than 512 double-ulps (2^-44
// relative) from a float rounding tie; anything doubtful (tie, tiny reduced argument, huge or non-finite phase) goes to
// cos_f itself.  Checked against glibc on 5M arguments (tests/test_pllmath.py::test_lean_cosine_equals_glibc).
// coefficient table of the unified chain, [parity][c5, c4, c3, c2, c1, c0, -, -]: row 0 (q even) the cosine's, row 1 the sine's
#define SDRB_COS_LEAN_TAB                                                                                                        \
    {0x1.1c065229821d1p-29, -0x1.27e1089e1a501p-22, 0x1.a019fcd9727a0p-16, -0x1.6c16c169ae93cp-10, 0x1.5555555554a28p-5, -0.5, 0.0, \
     0.0, 0.0, -0x1.a950938183dcbp-26, 0x1.71d752f9f8bdfp-19, -0x1.a019f946a7019p-13, 0x1.1111110cdbeb5p-7, -0x1.5555555552e41p-3,   \
     0.0, 0.0}
static const double h_cos_lean_tab[16] = SDRB_COS_LEAN_TAB;
#if defined(__CUDACC__)
__device__ __constant__ double c_cos_lean_tab[16] = SDRB_COS_LEAN_TAB;
#endif
// `tab`: the table above in memory that every lane can index at its own row cheaply (k_mix keeps a copy in shared memory;
// nullptr: the host table / the constant bank)
SDRB_HD float cos_lean_f(float t, const double* tab = nullptr) {
#if defined(__CUDA_ARCH__)
    if (!tab) tab = c_cos_lean_tab;
#else
    if (!tab) tab = h_cos_lean_tab;
#endif
    const double x = (double)t;
    const double tm = dfma(x, kTwoOverPi, kMagicRint);
    const double kd = dadd(tm, -kMagicRint);
    const unsigned q = dlo(tm) & 3u;
    const double r = dfma(-kd, kPio2M, dfma(-kd, kPio2H, x));
    const double z = dmul(r, r);
    const bool odd = (q & 1u) != 0u;
    const double* c = tab + 8 * (q & 1u);
    const double A = odd ? r : 1.0;
    const double B = odd ? dmul(r, z) : z;
    double p = dfma(c[0], z, c[1]);
    p = dfma(p, z, c[2]);
    p = dfma(p, z, c[3]);
    p = dfma(p, z, c[4]);
    p = dfma(p, z, c[5]);
    const double v = dfma(B, p, A);  // cos r or sin r
    // sign: q = 0 -> +cos, 1 -> -sin, 2 -> -cos, 3 -> +sin
    const double dc = flip_sign_if(v, ((q + 1u) >> 1) & 1u);
    // the tests, on integer words: phase below 2^45 (also rejects inf / NaN); reduced argument not tiny for its quadrant
    // count (reduce_rmin; a tiny r with k = 0 is harmless but rare enough to go the same way); result not within 512
    // double-ulps of a float rounding tie
    const bool big = (fbits(t) & 0x7FFFFFFFu) >= 0x56000000u;
    const bool tiny = (dhi(r) & 0x7FFFFFFFu) < reduce_rmin_hi(x);
    const bool tie = (((dlo(dc) & 0x1FFFFFFFu) + 2u * kAmbigUlps) & 0x1FFFFC00u) == 0x10000000u;
    if (big || tiny || tie) return cos_f_rare(t);
    return (float)dc;
}

SDRB_HD void pll_fast_sincos(float trigArg, PllFast& f) {
    double x = (double)trigArg;
    f.magic = float_round_magic(x);
    f.rmin_hi = reduce_rmin_hi(x);
    f.texp = dhi(x) & 0x7FF00000u;
    bool ok = fabs(x) < kReduceLimit && fabs(x) > 0x1p-100;
    if (ok) {
        double sa, cr_, r;
        int q;
        ok = sincos_reduce2(x, sa, cr_, r, q) && !near_float_boundary(sa) && !near_float_boundary(cr_);
        if (ok) {
            double ds, dc;
            pll_cs(sa, cr_, r, q, ds, dc);
            f.sa = sa;
            f.cr = cr_;
            f.r = r;
            f.kq = q;
            f.fbQ = (float)ds;
            f.fbI = (float)dc;
            f.generic_next = false;
        }
    }
    if (!ok) {
        sincos_f(trigArg, f.fbQ, f.fbI);
        f.generic_next = true;
    }
}
// fbI/fbQ from the reduced pieces (the speculative steps do not keep them up to date)
SDRB_HD void pll_fast_sync_fb(PllFast& f) {
    if (f.generic_next) return;
    double ds, dc;
    pll_cs(f.sa, f.cr, f.r, f.kq, ds, dc);
    f.fbQ = (float)ds;
    f.fbI = (float)dc;
}

SDRB_HD void pll_fast_load(PllFast& f, const PllState& st, const PllCoef& k) {
    f.integ = st.integrator;
    f.phase = st.phaseEst;
    f.trigOffset = st.trigOffset;
    f.sa = 0.0;
    f.cr = 1.0;
    f.r = 0.0;
    f.kq = 0;
    // the phase the carried feedbackI/Q were computed from (/root/reference/src/pll.cpp:47 with the carried values)
    float trigArg = (float)dadd(dmul(k.w, st.trigOffset), (double)st.phaseEst);
    pll_fast_sincos(trigArg, f);
    // a state that did not come out of this recurrence (hand-made feedback values): keep it, go general once
    if (f.fbI != st.feedbackI || f.fbQ != st.feedbackQ) f.generic_next = true;
    f.fbI = st.feedbackI;
    f.fbQ = st.feedbackQ;
}
SDRB_HD void pll_fast_store(PllFast& f, PllState& st) {
    pll_fast_sync_fb(f);
    st.feedbackI = f.fbI;
    st.feedbackQ = f.fbQ;
    st.integrator = f.integ;
    st.phaseEst = f.phase;
    st.trigOffset = f.trigOffset;
}

// 1/|in| for the rotated phase detector, or +inf when `in` is outside the range the speculative step handles
// (zero, subnormal, tiny, huge, inf, NaN): an infinite reciprocal makes the step's result NaN, which its range test
// rejects, and the careful path takes over.  `approx` is any approximation of 1/|in| good to 2^-22 relative.
SDRB_HD double pll_guard_recip(float in, double approx) {
    const uint32_t ex = (fbits(in) >> 23) & 0xFFu;  // accepted: 2^-90 <= |in| < 2^90
    return (ex - (127u - 90u)) < 180u ? approx : (double)INFINITY;
}

// One sample, careful form.  rin = 1/|in| from pll_guard_recip (only its finite values are used).  Returns trigArg.
SDRB_HD float pll_step_fast(float in, double rin, PllFast& f, const PllCoef& k, const AtanTab& tab) {
    pll_fast_sync_fb(f);  // the speculative steps leave fbI/fbQ behind
    const float x = fmul(in, f.fbI);
    const float y = fmul(in, -f.fbQ);
    float errorD;
    bool ok = !f.generic_next && x != 0.0f && y != 0.0f && rin < 1e300;
    if (ok) {
        double s0, c0;
        pll_cs(f.sa, f.cr, f.r, f.kq, s0, c0);
        double u = dfma((double)y, c0, dmul((double)x, s0));
        double w = dmul(u, in < 0.0f ? -rin : rin);
        int m = (f.kq + (in < 0.0f ? 2 : 0)) & 3;
        // -theta (+pi) = -r - m*pi/2, brought into (-pi, pi]:  m: 0 -> 0, 1 -> -pi/2, 3 -> +pi/2, 2 -> -+pi
        double mm = (m == 0) ? 0.0 : (m == 1) ? -1.0 : (m == 3) ? 1.0 : (f.r > 0.0 ? 2.0 : -2.0);
        double e = dadd(dadd(dmul(mm, kPio2H), -f.r), dfma(mm, kPio2M, w));
        ok = fabs(w) < 0x1p-22 && fabs(e) < 3.14159 && !near_float_boundary_abs(e);
        errorD = (float)e;
    }
    if (!ok) errorD = atan2_f(y, x, tab);
    f.integ = fadd(f.integ, fmul(k.Ki, errorD));
    f.phase = fadd(fadd(f.phase, fmul(k.Kp, errorD)), f.integ);
    f.trigOffset = dadd(f.trigOffset, 1.0);
    float trigArg = (float)dadd(dmul(k.w, f.trigOffset), (double)f.phase);
    pll_fast_sincos(trigArg, f);
    return trigArg;
}

// ---- speculative form: the same step without a single branch and without a conversion instruction on the chain ----
// Every acceptance test is a side computation OR-ed (bitwise, no short circuit) into `bad`; nothing on the recurrence
// waits for it.  The caller runs a few steps, looks at `bad` once, and in the (rare, ~1e-5 per step) case that any test
// failed restores the state it saved and repeats those steps with pll_step_fast.
//
// The chain per sample: sa,cr -> float (2 integer ops) -> |in|*fa, |in|*fc (FMUL) -> double (1 integer op) -> two
// DFMA -> errorD (2 integer ops) -> FMUL, FADD, FADD -> double (F2F) -> DADD, 2 DADD (float rounding) -> 2 (k) ->
// 2 DFMA (r) -> 5 deep Estrin -> sa,cr.  Signs never travel on it:
//   u/in = errorI*sin + errorQ*cos  over  in  =  (-1)^[r<0] * ( RN(|in| fc) * sa/|in|  -  RN(|in| fa) * cr/|in| )
// for every quadrant (fa, fc = RN_float(sa), RN_float(cr); the quadrant's swap of sine and cosine and their signs
// cancel between the products and the multipliers), so the float products are formed from magnitudes and the one
// remaining sign is folded into the reciprocal beforehand.
SDRB_HD unsigned ambig_rel_lo(double v) {  // within kAmbigUlps double-ulps of a float rounding tie (v float-normal)
    const uint32_t off = (dlo(v) & 0x1FFFFFFFu) - (0x10000000u - kAmbigUlps);
    return (unsigned)(off <= 2u * kAmbigUlps);
}
SDRB_HD unsigned ambig_abs(double v) {  // near_float_boundary_abs, branch-free
    const uint32_t lo = dlo(v), hi = dhi(v);
    const uint32_t E = (hi >> 20) & 0x7FFu;                      // biased exponent; accepted range e in [-17, 1]
    const uint32_t thr = 1u << ((1023u + 52u + (uint32_t)kAtanTolLog2 - E) & 31u);  // tolerance in units of 2^(e-52)
    const uint32_t off = (lo & 0x1FFFFFFFu) - 0x10000000u + thr;
    return (unsigned)((E - (1023u - 17u)) > 18u) | (unsigned)(off <= 2u * thr);
}

// `bad` is a plain flag in product builds; with -DSDRB_PLL_DIAG (tools/pll_diag.sh) each test owns a bit so that the
// careful-path counter can say which test sent a chunk there.
#if defined(SDRB_PLL_DIAG)
#define SDRB_BAD(cond, id) ((cond) ? (1u << (id)) : 0u)
#else
#if defined(SDRB_PLL_NO_TESTS)
#define SDRB_BAD(cond, id) 0u /* a bound, not a product: every acceptance test compiled out */
#else
#define SDRB_BAD(cond, id) ((unsigned)(cond))
#endif
#endif
SDRB_HD float pll_step_spec(float in, double rin, PllFast& f, const PllCoef& k, const PllK& kk, unsigned& bad) {
    // -- beside the chain: needs only `in` and the previous step's reduction --
    // rin is the raw reciprocal approximation: `in` outside 2^-90 <= |in| < 2^90 (zero, subnormal, inf, NaN too) rejects
    bad |= SDRB_BAD(((fbits(in) & 0x7FFFFFFFu) - 0x12800000u) >= (0x6C800000u - 0x12800000u), 0);
    const uint32_t rhi = dhi(f.r);
    const uint32_t rs = rhi & 0x80000000u;  // r < 0
    const unsigned m = ((unsigned)f.kq + ((fbits(in) >> 31) << 1)) & 3u;
    // -theta (+pi) = -r - m*pi/2 in (-pi, pi]:  mm = 0, -1, +-2 (the sign of r: pi - r for r > 0, -pi - r for r < 0), +1
    // for m = 0, 1, 2, 3.  (Round 1 carried the opposite sign for m = 2: |e| came out above pi, the wrap test below
    // rejected it and the careful path produced the right value - bit-exact, but every sample whose input sign disagrees
    // with the NCO's, a quarter of all samples once the loop is out of lock or the float phase grid is coarser than pi
    // (45 s into a stream for the 114 kHz loop), cost a careful repeat and k_pll ran 6x slower from then on.)
    const uint32_t mmhi = (m & 1u) ? ((m & 2u) ? 0x3FF00000u : 0xBFF00000u) : ((m & 2u) ? (0x40000000u ^ rs) : 0u);
    const double mm = mkd(mmhi, 0u);
    const double base = dfma(mm, kk.v[kKPio2M], dfma(mm, kk.v[kKPio2H], -f.r));
    const uint32_t bh = dhi(base);
    const uint32_t Ke = d2f_K(bh);
    const double ars = mkd(dhi(rin) ^ rs, dlo(rin));  // (-1)^[r<0] / |in|
    const double ma = dmul(f.sa, ars), mc = dmul(f.cr, -ars);
    // -- the chain --
    const float fa = bitsf(d2f_known(f.sa, 0xC0000000u)), fc = bitsf(d2f_known(f.cr, 0xC0000000u));
    const float pa = fmul(fabsf(in), fa), pc = fmul(fabsf(in), fc);
    const double e = dfma(f2d_pos(pc), ma, dfma(f2d_pos(pa), mc, base));
    const float errorD = bitsf(d2f_known(e, Ke));
    f.integ = fadd(f.integ, fmul(k.Ki, errorD));
    f.phase = fadd(fadd(f.phase, fmul(k.Kp, errorD)), f.integ);
    f.trigOffset = dadd(f.trigOffset, 1.0);
#if defined(SDRB_PLL_PHASE_INTCONV)
    const uint32_t pb = fbits(f.phase);  // (double)phase by hand: sign | (exponent + 896, mantissa << 29); zero/subnormal/inf/NaN -> bad
    const double phd = mkd((((pb & 0x7FFFFFFFu) >> 3) + 0x38000000u) | (pb & 0x80000000u), pb << 29);
    bad |= (unsigned)((((pb >> 23) & 0xFFu) - 1u) >= 254u);
#else
    const double phd = (double)f.phase;
#endif
    const double td = dadd(dmul(k.w, f.trigOffset), phd);
    // (double)(float)td without the round trip: td + M - M with M = 1.5 * 2^(E+29) rounds td to 24 significant bits,
    // to nearest even, exactly like the conversion, provided td lies in the binade E that M was built for.  M comes from
    // the previous step's phase (the binade changes once per doubling of the phase); a mismatch counts as bad, and so
    // the range tests of the phase (below kReduceLimit, not tiny) only have to be made where M is made (and so is the
    // smallest reduced argument that binade's quadrant counts allow, rmin_hi).
    const double xd = dadd(dadd(td, f.magic), -f.magic);
    // quarter-turn reduction and polynomials (sincos_reduce2, inlined so that its test joins `bad`).  The quadrant
    // count must come from xd, not td: once the phase passes 2^22 the float grid is coarser than pi/4, td and xd can
    // be a radian apart, and a count taken from td would leave |r| far outside the range of the kernels below.
    const double tm = dfma(xd, kk.v[kK2OverPi], kMagicRint);
    const double kd = dadd(tm, -kMagicRint);
    const int q = (int)dlo(tm) & 3;
    const double r = dfma(-kd, kk.v[kKPio2M], dfma(-kd, kk.v[kKPio2H], xd));
    const uint32_t rah = dhi(r) & 0x7FFFFFFFu;
    double sa, cr_;
    sincos_poly2k(r, mkd(rah, dlo(r)), sa, cr_, kk);  // |r| joins only at the last fma of the sine
    // -- the tests --
    // e: the wrap (|e| < pi), the float rounding of e, and base so close to +-2 that e and base could differ in the
    // nibble d2f_known was told.  (The linearisation atan(u/v) = u/in needs |u/in| < 2^-22, which holds by construction:
    // |u/in| <= sa * cr * 4 * 2^-24 from the four float roundings, none of which can underflow for 2^-90 <= |in| < 2^90
    // and sa >= 2^-31, both enforced; a NaN/inf e cannot arise from an accepted `in` either.)
    bad |= SDRB_BAD((dhi(e) & 0x7FFFFFFFu) >= 0x400921F9u, 1) | SDRB_BAD(ambig_abs(e), 2) | SDRB_BAD(((bh & 0x7FFFFFFFu) - 0x3FFFFFFFu) <= 1u, 3);
    // td left the binade of M; r tiny; sa / cr near a float rounding tie
    bad |= SDRB_BAD((((dhi(td) & 0x7FF00000u) + (29u << 20)) | (1u << 19)) != dhi(f.magic), 4) | SDRB_BAD(rah < f.rmin_hi, 5) |
           SDRB_BAD(ambig_rel_lo(sa), 6) | SDRB_BAD(ambig_rel_lo(cr_), 7);
    f.sa = sa;
    f.cr = cr_;
    f.r = r;
    f.kq = q;
    return (float)td;
}

// ---- the speculative step cut in two, for the rotated loop of the batched kernel ----
// head: the phase detector of one sample (everything up to e, the double whose float rounding is errorD); a pure function
//       of the loop state and the input, so it can be evaluated for the NEXT chunk's first sample before the current
//       chunk's acceptance flag is looked at: the branch on that flag then resolves in the shadow of the head instead of
//       stopping the recurrence (the flag's last inputs, the float-tie tests of sa / cr, are ready ~45 cycles before it).
// tail: loop filter, NCO phase, reduction and the sine / cosine kernels (e -> new sa, cr, r, kq).
// The tests are leaner than pll_step_spec's (same guarantees, fewer instructions beside the chain):
//   e   : RN_f(e - 2^-43) == RN_f(e + 2^-43), i.e. no float rounding step within the error bound of e, whatever its
//         binade; replaces the exponent-dependent window of ambig_abs (and rejects |e| < ~2^-19 by itself);
//   sa,cr : low 29 bits within 512 (sa: the lean sine is good to 2^-46.3) / 64 (cr: 2^-50) of the tie pattern, as one add and one
//           masked compare;
//   td  : exponent field against the one stored with the magic constant (one masked compare);
//   in  : the range test is made once per chunk on min / max of the four inputs (pll_chunk4r).
struct PllHead {
    double e;
    uint32_t Ke, bh;  // bh: high word of base (for the +-2 test)
};
constexpr double kAtanTol = 0x1p-43;  // 2^kAtanTolLog2
// low 29 bits in [tie - W, tie + W - 1], W a power of two: within W double-ulps of a float rounding tie
template <uint32_t W>
SDRB_HD unsigned ambig_tie29(double v) {
    return (unsigned)((((dlo(v) + W) & (0x1FFFFFFFu & ~(2u * W - 1u))) ^ 0x10000000u) == 0u);
}
SDRB_HD PllHead pll_spec_head(float in, double rin, const PllFast& f, const PllK& kk) {
    const uint32_t rhi = dhi(f.r);
    const uint32_t rs = rhi & 0x80000000u;  // r < 0
    const unsigned m = ((unsigned)f.kq + ((fbits(in) >> 31) << 1)) & 3u;
    const uint32_t mmhi = (m & 1u) ? ((m & 2u) ? 0x3FF00000u : 0xBFF00000u) : ((m & 2u) ? (0x40000000u ^ rs) : 0u);  // see pll_step_spec
    const double mm = mkd(mmhi, 0u);
    const double ars = mkd(dhi(rin) ^ rs, dlo(rin));  // (-1)^[r<0] / |in|
    // -- the chain: sa, cr -> float -> products -> double -> e --
    const float fa = bitsf(d2f_known(f.sa, 0xC0000000u));
    const double mc = dmul(f.cr, -ars);
    const float fc = bitsf(d2f_known(f.cr, 0xC0000000u));
    const double ma = dmul(f.sa, ars);
    const float pa = fmul(fabsf(in), fa), pc = fmul(fabsf(in), fc);
    const double base = dfma(mm, kk.v[kKPio2M], dfma(mm, kk.v[kKPio2H], -f.r));
    const uint32_t bh = dhi(base);
    const uint32_t Ke = d2f_K(bh);
    const double e = dfma(f2d_pos(pc), ma, dfma(f2d_pos(pa), mc, base));
    return PllHead{e, Ke, bh};
}
// The tests of a head: sa / cr are the values it was computed from (their float roundings fa, fc must not be near a tie);
// wrap (|e| < pi; NaN / inf from an unusable input land here too), float rounding of e, base at +-2.
SDRB_HD void pll_spec_head_tests(const PllHead& h, double sa, double cr_, unsigned& bad) {
    // Error bound of e.  Every term of it scales with sa cr = |sin 2r| / 2: the linearisation and the reciprocal (2^-22.5 +
    // 2^-23 relative to a perturbation of at most 2^-22 sa cr), the kernels' 2^-46.3, the double roundings of products of
    // size sa cr: together below 2^-43.3 sa cr; the only absolute terms are the four double roundings of base and e at a
    // magnitude of up to pi (<= 2^-50 together; zero in the quadrant where e = -r + ...).  Tolerance: 2^-42 sa + 2^-48 (one
    // fma; 2.4x / 4x those bounds; tests/test_pllmath.py measures the error against long-double atan2: it stays below
    // a quarter of the tolerance on 10M random states).  With the flat 2^-43 of round 1
    // every binade of small |e| rejected as often as the top one (the float grid of e shrinks with |e|, the tolerance did
    // not): 2.5e-5 rejections per sample, four fifths of them below |e| = 1/4 and avoidable, each a careful repeat that
    // stalls the whole warp for thousands of cycles.
#if defined(SDRB_PLL_FLAT_TOL)
    const double tol = kAtanTol;
#else
    const double tol = dfma(sa, 0x1p-42, 0x1p-48);
#endif
    const uint32_t elo = d2f_known(dadd(h.e, -tol), h.Ke), ehi = d2f_known(dadd(h.e, tol), h.Ke);
    bad |= SDRB_BAD((dhi(h.e) & 0x7FFFFFFFu) >= 0x400921F9u, 1) | SDRB_BAD(elo != ehi, 2) |
           SDRB_BAD(((h.bh & 0x7FFFFFFFu) - 0x3FFFFFFFu) <= 1u, 3) | SDRB_BAD(ambig_tie29<512>(sa), 6) | SDRB_BAD(ambig_tie29<64>(cr_), 7);
}
SDRB_HD float pll_spec_tail(const PllHead& h, PllFast& f, const PllCoef& k, const PllK& kk, unsigned& bad) {
    const float errorD = bitsf(d2f_known(h.e, h.Ke));
    f.integ = fadd(f.integ, fmul(k.Ki, errorD));
    f.phase = fadd(fadd(f.phase, fmul(k.Kp, errorD)), f.integ);
    f.trigOffset = dadd(f.trigOffset, 1.0);
    const double phd = (double)f.phase;  // F2F, ~20 cycles; integer forms need three dependent operations for a signed phase (measured slower)
    const double td = dadd(dmul(k.w, f.trigOffset), phd);
    const double xd = dadd(dadd(td, f.magic), -f.magic);  // (double)(float)td, see pll_step_spec
    const double tm = dfma(xd, kk.v[kK2OverPi], kMagicRint);
    const double kd = dadd(tm, -kMagicRint);
    const int q = (int)dlo(tm) & 3;
    const double r = dfma(-kd, kk.v[kKPio2M], dfma(-kd, kk.v[kKPio2H], xd));
    const uint32_t rah = dhi(r) & 0x7FFFFFFFu;
    double sa, cr_;
    sincos_poly2_lean(r, mkd(rah, dlo(r)), sa, cr_, kk);  // kk from pll_k_load_lean
    // td left the binade of the magic constant; r tiny.  (sa / cr near a float tie: tested by the head that uses them.)
    bad |= SDRB_BAD(((dhi(td) & 0x7FF00000u) ^ f.texp) != 0u, 4) | SDRB_BAD(rah < f.rmin_hi, 5);
    f.sa = sa;
    f.cr = cr_;
    f.r = r;
    f.kq = q;
    return (float)td;
}
// The loop state as the careful path needs it: the speculative steps leave sa / cr without a tie test and fbI / fbQ behind,
// so everything is derived again from the float NCO phase (/root/reference/src/pll.cpp:47 with the current values).
SDRB_HD void pll_fast_resync(PllFast& f, const PllCoef& k) {
    pll_fast_sincos((float)dadd(dmul(k.w, f.trigOffset), (double)f.phase), f);
}

#if defined(__CUDA_ARCH__)
#define SDRB_RARE __device__ __noinline__
#else
#define SDRB_RARE inline
#endif
// the careful repeat of four steps (rare: kept out of line on the device so the hot loop stays small)
// one careful step, out of line: the four steps of a repeat run through the same instruction-cache lines
SDRB_RARE float pll_step_fast_cold(float in, double rin, PllFast& f, const PllCoef& k, const AtanTab& tab) {
    return pll_step_fast(in, rin, f, k, tab);
}
// The repeat of four steps after a raised flag, on the careful path: a loop over one out-of-line step, so that the four
// steps run through the same instruction-cache lines.  (Trying each step on the speculative path first and sending only the
// one that fails to the careful step was built too: ptxas then no longer proves the warp converged at the loop's vote and
// guards it with a BRA.DIV, which cost 27 cycles per sample.)
SDRB_RARE void pll_redo4(float i0, float i1, float i2, float i3, double r0, double r1, double r2, double r3, PllFast& f,
                         const PllCoef& k, const AtanTab& tab, float& t0, float& t1, float& t2, float& t3) {
    if (!f.generic_next) pll_fast_resync(f, k);  // sa / cr may be untested (rotated loop): derive the state again from the float phase
    const float in[4] = {i0, i1, i2, i3};
    const double rr[4] = {r0, r1, r2, r3};  // raw reciprocals: guarded here
    float t[4];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int j = 0; j < 4; j++) t[j] = pll_step_fast_cold(in[j], pll_guard_recip(in[j], rr[j]), f, k, tab);
    t0 = t[0]; t1 = t[1]; t2 = t[2]; t3 = t[3];
}

// Four consecutive samples: speculative run, verified once; the careful path only on failure.
// r0..r3 = raw reciprocal approximations 1/|in| of the samples (any value for an `in` the step rejects).
// redo_ctr (device only, may be null): incremented once per lane and chunk that had to take the careful path.
SDRB_HD void pll_chunk4(float i0, float i1, float i2, float i3, double r0, double r1, double r2, double r3, PllFast& f,
                        const PllCoef& k, const PllK& kk, const AtanTab& tab, float& t0, float& t1, float& t2, float& t3,
                        unsigned long long* redo_ctr = nullptr) {
    const PllFast saved = f;
    unsigned bad = SDRB_BAD(f.generic_next, 8);
    t0 = pll_step_spec(i0, r0, f, k, kk, bad);
    t1 = pll_step_spec(i1, r1, f, k, kk, bad);
    t2 = pll_step_spec(i2, r2, f, k, kk, bad);
    t3 = pll_step_spec(i3, r3, f, k, kk, bad);
    if (bad) {  // only here does the state have to live in addressable memory (the out-of-line call)
#if defined(__CUDA_ARCH__)
        if (redo_ctr) atomicAdd(redo_ctr, 1ull);
#if defined(SDRB_PLL_DIAG)
        for (int b = 0; b < 9; b++)
            if ((bad >> b) & 1u) atomicAdd(redo_ctr + 2 * (1 + b), 1ull);  // per-test counters behind the two totals
#endif
#else
        (void)redo_ctr;
#endif
        PllFast again = saved;
        float a0, a1, a2, a3;
        pll_redo4(i0, i1, i2, i3, r0, r1, r2, r3, again, k, tab, a0, a1, a2, a3);
        f = again;
        t0 = a0; t1 = a1; t2 = a2; t3 = a3;
    }
}


// The rotated chunk: `h` is the head of this chunk's first sample, computed by the previous call (pll_chunk4r_prime before
// the first) and tested here; n0 / rn0 are the next chunk's first sample and its raw reciprocal (anything if there is none:
// the head is a pure function and its result is then never used).  `first` (0/1): the state did not come out of a
// speculative step (start of a block, after a careful repeat) and says generic_next.
SDRB_HD void pll_chunk4r_prime(float i0, double r0, const PllFast& f, const PllK& kk, PllHead& h) { h = pll_spec_head(i0, r0, f, kk); }
// The speculative part of a chunk: four tails and heads; `h` is the head of the first sample and sa0 / cr0 the values it was
// computed from, `hn` receives the head of the sample after the chunk (n0 / rn0), which the next call tests.
SDRB_HD void pll_chunk4r_core(float i0, float i1, float i2, float i3, double r1, double r2, double r3, float n0, double rn0, PllFast& f,
                              double sa0, double cr0, const PllHead& h, PllHead& hn, const PllCoef& k, const PllK& kk, float& t0,
                              float& t1, float& t2, float& t3, unsigned& bad) {
    // every |in| of the chunk within [2^-90, 2^90): min / max instead of four tests (a NaN input is dropped by fmin / fmax
    // and shows up as a NaN e, which the wrap test rejects; zero and subnormal inputs have an infinite reciprocal: same)
    const float mn = fminf(fminf(fabsf(i0), fabsf(i1)), fminf(fabsf(i2), fabsf(i3)));
    const float mx = fmaxf(fmaxf(fabsf(i0), fabsf(i1)), fmaxf(fabsf(i2), fabsf(i3)));
    bad |= SDRB_BAD(!(mn >= 0x1p-90f), 0) | SDRB_BAD(!(mx < 0x1p90f), 0);
    pll_spec_head_tests(h, sa0, cr0, bad);
    t0 = pll_spec_tail(h, f, k, kk, bad);
    PllHead g = pll_spec_head(i1, r1, f, kk);
    pll_spec_head_tests(g, f.sa, f.cr, bad);
    t1 = pll_spec_tail(g, f, k, kk, bad);
    g = pll_spec_head(i2, r2, f, kk);
    pll_spec_head_tests(g, f.sa, f.cr, bad);
    t2 = pll_spec_tail(g, f, k, kk, bad);
    g = pll_spec_head(i3, r3, f, kk);
    pll_spec_head_tests(g, f.sa, f.cr, bad);
    t3 = pll_spec_tail(g, f, k, kk, bad);
    hn = pll_spec_head(n0, rn0, f, kk);
}
// One chunk: returns the acceptance flag; `saved` receives the state before the chunk (what the careful repeat starts from).
SDRB_HD unsigned pll_chunk4r_spec(float i0, float i1, float i2, float i3, double r1, double r2, double r3, float n0, double rn0, PllFast& f,
                                  PllFast& saved, const PllHead& h, PllHead& hn, const PllCoef& k, const PllK& kk, float& t0, float& t1,
                                  float& t2, float& t3) {
    saved = f;
    unsigned bad = SDRB_BAD(f.generic_next, 8);
    pll_chunk4r_core(i0, i1, i2, i3, r1, r2, r3, n0, rn0, f, saved.sa, saved.cr, h, hn, k, kk, t0, t1, t2, t3, bad);
    return bad;
}
// The careful repeat of a chunk whose flag was raised (`bad` only selects the diagnostic counters).
SDRB_HD void pll_chunk4r_redo(unsigned bad, float i0, float i1, float i2, float i3, double r0, double r1, double r2, double r3, float n0,
                              double rn0, PllFast& f, const PllFast& saved, PllHead& hn, const PllCoef& k, const PllK& kk,
                              const AtanTab& tab, float& t0, float& t1, float& t2, float& t3, unsigned long long* redo_ctr) {
#if defined(__CUDA_ARCH__)
    if (redo_ctr) atomicAdd(redo_ctr, 1ull);
#if defined(SDRB_PLL_DIAG)
    for (int b = 0; b < 9; b++)
        if ((bad >> b) & 1u) atomicAdd(redo_ctr + 2 * (1 + b), 1ull);
#endif
#else
    if (redo_ctr) ++*redo_ctr;
#endif
    (void)bad;
    PllFast again = saved;
    float a0, a1, a2, a3;
    pll_redo4(i0, i1, i2, i3, r0, r1, r2, r3, again, k, tab, a0, a1, a2, a3);
    f = again;
    t0 = a0; t1 = a1; t2 = a2; t3 = a3;
    hn = pll_spec_head(n0, rn0, f, kk);
}
SDRB_HD void pll_chunk4r(float i0, float i1, float i2, float i3, double r0, double r1, double r2, double r3, float n0, double rn0,
                         PllFast& f, PllHead& h, const PllCoef& k, const PllK& kk, const AtanTab& tab, float& t0, float& t1,
                         float& t2, float& t3, unsigned long long* redo_ctr = nullptr) {
    PllFast saved;
    PllHead hn;
    const unsigned bad = pll_chunk4r_spec(i0, i1, i2, i3, r1, r2, r3, n0, rn0, f, saved, h, hn, k, kk, t0, t1, t2, t3);
    if (bad) pll_chunk4r_redo(bad, i0, i1, i2, i3, r0, r1, r2, r3, n0, rn0, f, saved, hn, k, kk, tab, t0, t1, t2, t3, redo_ctr);
    h = hn;
}

}  // namespace cr
}  // namespace sdrb
