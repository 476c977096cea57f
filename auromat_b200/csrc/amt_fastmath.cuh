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

// atan(i/32) in DEGREES, i = 0..32.  Every consumer of the arctangents wants degrees
// (lat/lon/MLat, MLT = smlon/15 + 12, elevation), so the conversion factor 180/pi is folded into
// the table and into the polynomial coefficients instead of costing a multiply per result.
__constant__ double c_atan_tab[33] = {
    0.0, 1.7899106082460694, 3.576334374997351,
    5.35582504285519, 7.125016348901798, 8.880659150520245,
    10.619655276155134, 12.339087278326195, 14.036243467926479,
    15.708637829015744, 17.35402463626132, 18.970407808486545,
    20.556045219583467, 22.109448343751673, 23.629377730656817,
    25.114834886144564, 26.56505117707799, 27.979474388480146,
    29.357753542791276, 30.699722550814414, 32.005383208083494,
    33.274887984834926, 34.5085229876684, 35.706691400602885,
    36.86989764584402, 37.99873244250467, 39.0938588862295,
    40.155999624919325, 41.18592516570965, 42.18444331578877,
    43.15238973400541, 44.09061955080086, 45.0,
};

// Constants whose low mantissa word is non-zero cannot be instruction immediates; kept in
// constant memory they become direct c[bank][offset] operands of DFMA/DADD/DMUL instead of two
// UMOV / IMAD.MOV each (the polynomial coefficients alone were 14 of the 54 instructions of the
// first version of atan2_fast).
__constant__ double c_fm[8] = {
    6.366197723675814, -8.18511135901176, 11.459155902616464, -19.098593171027442, 57.29577951308232,      // (180/pi) * {1/9, -1/7, 1/5, -1/3, 1}: atan Taylor series in degrees
    0.0, 0.0, 0.0,
};

__device__ __forceinline__ double fabs_bits(double x) {        // |x| on the integer pipe, not the FP64 pipe
    return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x));
}

// atan(mn/mx) in degrees for 0 <= mn <= mx, mx > 0: pick c = i/32 nearest to mn/mx from the MUFU
// reciprocal seed, then atan(mn/mx) = atan(c) + atan(t), t = (mn - c*mx)/(mx + c*mn),
// |t| <= 1/64 + 2^-19, where the degree-9 odd Taylor polynomial is exact to 1e-21:
// result = tab_deg[i] + t * K(1 - s/3 + s^2/5 - s^3/7 + s^4/9), s = t^2  (1 DMUL + 5 DFMA).
__device__ __forceinline__ double atan_ratio_deg(double mn, double mx) {
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
    p = fma(p, s, c_fm[4]);
    return fma(t, p, c_atan_tab[i]);
}

// atan2(y, x) in degrees, any quadrant, finite inputs, not both zero.
__device__ __forceinline__ double atan2_deg(double y, double x) {
    const double a = fabs_bits(y), b = fabs_bits(x);
    const bool swap = a > b;
    double r = atan_ratio_deg(swap ? b : a, swap ? a : b);
    if (swap) r = 90.0 - r;
    if (x < 0.0) r = 180.0 - r;
    return copysign(r, y);
}

// atan2(y, x) in degrees for x >= 0 (result in [-90, 90]); also serves atan(y/x) and asin.
__device__ __forceinline__ double atan2_posx_deg(double y, double x) {
    const double a = fabs_bits(y);
    const bool swap = a > x;
    double r = atan_ratio_deg(swap ? x : a, swap ? a : x);
    if (swap) r = 90.0 - r;
    return copysign(r, y);
}

// 90 - acos(d) = asin(d) in degrees for d in [-1, 1], via atan2(d, sqrt((1-d)(1+d)));
// (1-d) is exact for d >= 0.5.
__device__ __forceinline__ double asin_deg(double d) {
    const double w = (1.0 - d) * (1.0 + d);
    const double s = w > 0.0 ? sqrt_fast(w) : 0.0;
    if (s == 0.0) return copysign(90.0, d);
    return atan2_posx_deg(d, s);
}

}  // namespace amt
