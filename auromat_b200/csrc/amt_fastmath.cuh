// FP64 building blocks for the georeference kernels: reciprocal / division / (r)sqrt from the
// MUFU seed + Newton/Goldschmidt steps, and a table-driven atan2.
//
// Why not libm: ncu on the first version of k_georef_points (profiles/r01_*) showed 1280
// issued instructions per point of which only ~400 were FP64 math -- CUDA's atan2/atan/acos
// spend most of their instructions on special-case handling and on materialising polynomial
// coefficients with UMOV/IMAD.MOV.  The functions below keep their constants in __constant__
// memory (used as direct c[][] operands), have no slow paths (inputs are finite, non-denormal
// coordinates in km / ratios) and are accurate to <= 2 ulp, i.e. ~1e-14 degrees after the
// whole chain against a 1e-9 degree parity budget.
#pragma once
#include <cuda_runtime.h>

namespace amt {

__device__ __forceinline__ double mufu_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}
__device__ __forceinline__ double mufu_rsqrt(double a) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
}

// 1/a to ~2^-40: ~20-bit seed + one Newton step (2 DFMA).
__device__ __forceinline__ double rcp_nr1(double a) {
    const double r = mufu_rcp(a);
    return fma(r, fma(-a, r, 1.0), r);
}

// n/a to <= 1 ulp: q = n*r with r ~ 1/a to 2^-40, then one residual correction
// q += r*(n - a*q) (error 2^-40 * 2^-40 before rounding).  MUFU + 1 DMUL + 4 DFMA.
__device__ __forceinline__ double div_fast(double n, double a) {
    const double r = rcp_nr1(a);
    const double q = n * r;
    return fma(fma(-a, q, n), r, q);
}

// Goldschmidt from the ~20-bit MUFU seed: g -> sqrt(a) (<= 1 ulp after the residual step),
// h -> 0.5/sqrt(a) to ~2^-40 (enough wherever it only scales a small correction term).
__device__ __forceinline__ void sqrt_rsqrt(double a, double& sq, double& half_rsq) {
    const double y = mufu_rsqrt(a);
    double g = a * y;
    double h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    sq = fma(fma(-g, g, a), h, g);      // residual step: error 2^-40 * 2^-40 before rounding
    half_rsq = h;
}
// 1/sqrt(a) to ~2^-40 (one Goldschmidt step)
__device__ __forceinline__ double rsqrt_40(double a) {
    const double y = mufu_rsqrt(a);
    const double g = a * y;
    double h = 0.5 * y;
    h = fma(h, fma(-g, h, 0.5), h);
    return h + h;
}
__device__ __forceinline__ double sqrt_fast(double a) {
    double s, h;
    sqrt_rsqrt(a, s, h);
    return s;
}
// 1/sqrt(a) to ~1 ulp (no residual step on g needed).
__device__ __forceinline__ double rsqrt_fast(double a) {
    const double y = mufu_rsqrt(a);
    double g = a * y;
    double h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    h = fma(h, r, h);
    return h + h;
}

// atan(i/32), i = 0..32
__constant__ double c_atan_tab[33] = {
    0, 0.031239833430268277, 0.06241880999595735,
    0.09347678115858947, 0.12435499454676144, 0.15499674192394097,
    0.18534794999569476, 0.21535769969773805, 0.24497866312686414,
    0.27416745111965879, 0.30288486837497142, 0.3310960767041321,
    0.35877067027057225, 0.38588266939807375, 0.41241044159738732,
    0.43833655985795783, 0.46364760900080609, 0.48833395105640554,
    0.51238946031073773, 0.5358112379604637, 0.55859931534356244,
    0.58075635356767041, 0.60228734613496415, 0.6231993299340659,
    0.64350110879328437, 0.66320299270609329, 0.68231655487474807,
    0.70085440788445019, 0.71882999962162453, 0.7362574289814281,
    0.75315128096219441, 0.7695264804056583, 0.78539816339744828,
};

// Constants whose low mantissa word is non-zero cannot be instruction immediates; kept in
// constant memory they become direct c[bank][offset] operands of DFMA/DADD/DMUL instead of two
// UMOV / IMAD.MOV each (the polynomial coefficients alone were 14 of the 54 instructions of the
// first version of atan2_fast).
__constant__ double c_fm[8] = {
    1.0 / 9.0, -1.0 / 7.0, 1.0 / 5.0, -1.0 / 3.0,      // atan Taylor coefficients
    1.5707963267948966, 3.141592653589793,             // pi/2, pi
    57.29577951308232, 0.017453292519943295,           // 180/pi, pi/180
};
#define AMT_C_HALF_PI c_fm[4]
#define AMT_C_PI c_fm[5]

__device__ __forceinline__ double fabs_bits(double x) {        // |x| on the integer pipe, not the FP64 pipe
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}

// atan(mn/mx) for 0 <= mn <= mx, mx > 0: pick c = i/32 nearest to mn/mx from the MUFU
// reciprocal seed, then atan(mn/mx) = atan(c) + atan(t), t = (mn - c*mx)/(mx + c*mn),
// |t| <= 1/64 + 2^-19, where the degree-9 odd Taylor polynomial is exact to 1e-21.
__device__ __forceinline__ double atan_ratio(double mn, double mx) {
    const double q = mn * mufu_rcp(mx);
    // q + 1.5*2^47 has an ulp of exactly 1/32: the sum IS q rounded to a multiple of 1/32,
    // and its low mantissa word is the integer 32*c
    const double magic = 211106232532992.0;      // 1.5 * 2^47
    const double qi = q + magic;
    int i = __double2loint(qi);
    i = min(max(i, 0), 32);                      // also keeps NaN inputs inside the table
    const double c = qi - magic;                 // == i/32 exactly (q is in [0, 1])
    const double num = fma(-c, mx, mn);
    const double den = fma(c, mn, mx);
    const double t = div_fast(num, den);
    const double s = t * t;
    double p = c_fm[0];
    p = fma(p, s, c_fm[1]);
    p = fma(p, s, c_fm[2]);
    p = fma(p, s, c_fm[3]);
    const double ts = t * s;
    return c_atan_tab[i] + fma(ts, p, t);
}

constexpr double kPi = 3.141592653589793;
constexpr double kHalfPi = 1.5707963267948966;

// atan2(y, x), any quadrant, finite inputs, not both zero.
__device__ __forceinline__ double atan2_fast(double y, double x) {
    const double a = fabs_bits(y), b = fabs_bits(x);
    const bool swap = a > b;
    double r = atan_ratio(swap ? b : a, swap ? a : b);
    if (swap) r = AMT_C_HALF_PI - r;
    if (x < 0.0) r = AMT_C_PI - r;
    return copysign(r, y);
}

// atan2(y, x) for x >= 0 (result in [-pi/2, pi/2]); also serves atan(y/x).
__device__ __forceinline__ double atan2_posx(double y, double x) {
    const double a = fabs_bits(y);
    const bool swap = a > x;
    double r = atan_ratio(swap ? x : a, swap ? a : x);
    if (swap) r = AMT_C_HALF_PI - r;
    return copysign(r, y);
}

// acos(d) for d in [-1, 1] via atan2(sqrt((1-d)(1+d)), d); (1-d) is exact for d >= 0.5.
__device__ __forceinline__ double acos_fast(double d) {
    const double w = (1.0 - d) * (1.0 + d);
    const double s = w > 0.0 ? sqrt_fast(w) : 0.0;
    const double a = fabs_bits(d);
    const bool swap = s > a;
    if (s == 0.0) return d < 0.0 ? kPi : 0.0;
    double r = atan_ratio(swap ? a : s, swap ? s : a);
    if (swap) r = AMT_C_HALF_PI - r;
    if (d < 0.0) r = AMT_C_PI - r;
    return r;
}

}  // namespace amt
