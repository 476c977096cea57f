// FP64 building blocks for the georeference kernels: reciprocal / division / (r)sqrt from the
// MUFU seed + Newton/Goldschmidt steps, and a table-driven atan2 without octant selects.
//
// Why not libm: ncu on the first version of k_georef_points (profiles/r01_*) showed 1280
// issued instructions per point of which only ~400 were FP64 math -- CUDA's atan2/atan/acos
// spend most of their instructions on special-case handling and on materialising polynomial
// coefficients with UMOV/IMAD.MOV.  The functions below keep their constants in __constant__
// memory (used as direct c[][] operands), have no slow paths (inputs are finite, non-denormal
// coordinates in km / ratios).
//
// Accuracy contract.  The parity budget of the path is 1e-9 degrees (1.7e-11 rad).  The only
// ill-conditioned block is the ray / ellipsoid intersection, which keeps <= 1 ulp primitives
// (sqrt_fast, div_fast).  Everything after it is well conditioned; there each primitive is held
// to what its operand needs: main terms (p, s, cos e) <= 1 ulp, arctangent remainders and
// e^2-sized correction terms 2^-38 relative.  Measured end to end: <= 2e-11 degrees
// (tests/test_gpu_parity.py, scripts/host_math_check.cu runs the same code on the host).
//
// The header also compiles for the HOST (scripts/host_math_check.cu): the two MUFU seeds are
// then emulated by a 20-bit truncation of the exact value, which is what the hardware returns
// (rcp/rsqrt.approx.ftz.f64 read and write the upper 32 bits of the operand only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <cmath>

#define AMT_HD __host__ __device__ __forceinline__

namespace amt {

AMT_HD double bits_to_double(unsigned hi, unsigned lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
AMT_HD unsigned hi_word(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (unsigned)(b >> 32);
#endif
}
AMT_HD unsigned lo_word(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2loint(x);
#else
    uint64_t b;
    memcpy(&b, &x, 8);
    return (unsigned)b;
#endif
}

AMT_HD double mufu_rcp(double a) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
#else
    const double r = 1.0 / bits_to_double(hi_word(a), 0u);
    return bits_to_double(hi_word(r), 0u);
#endif
}
AMT_HD double mufu_rsqrt(double a) {
#ifdef __CUDA_ARCH__
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    return r;
#else
    const double r = 1.0 / sqrt(bits_to_double(hi_word(a), 0u));
    return bits_to_double(hi_word(r), 0u);
#endif
}

// 0.5 * y for a normal, non-tiny y: exponent decrement on the integer pipe instead of a DMUL
AMT_HD double half_of(double y) { return bits_to_double(hi_word(y) - 0x00100000u, lo_word(y)); }

// 1/a to ~2^-39: ~20-bit seed + one Newton step (2 DFMA).
AMT_HD double rcp_nr1(double a) {
    const double r = mufu_rcp(a);
    return fma(r, fma(-a, r, 1.0), r);
}

// n/a to <= 1 ulp: q = n*r with r ~ 1/a to 2^-39, then one residual correction
// q += r*(n - a*q) (error 2^-39 * 2^-39 before rounding).  MUFU + 1 DMUL + 4 DFMA.
AMT_HD double div_fast(double n, double a) {
    const double r = rcp_nr1(a);
    const double q = n * r;
    return fma(fma(-a, q, n), r, q);
}

// n/a to ~2^-38 relative straight from the seed: q0 = n*r0, q = q0 + r0*(n - a*q0)
// (the seed error enters squared).  MUFU + 1 DMUL + 2 DFMA.
AMT_HD double div_38(double n, double a) {
    const double r = mufu_rcp(a);
    const double q = n * r;
    return fma(fma(-a, q, n), r, q);
}

// Goldschmidt from the ~20-bit MUFU seed: g -> sqrt(a) (<= 1 ulp after the residual step),
// h -> 0.5/sqrt(a) to ~2^-39 (enough wherever it only scales a small correction term).
AMT_HD void sqrt_rsqrt(double a, double& sq, double& half_rsq) {
    const double y = mufu_rsqrt(a);
    double g = a * y;
    double h = half_of(y);
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    sq = fma(fma(-g, g, a), h, g);      // residual step: error 2^-39 * 2^-39 before rounding
    half_rsq = h;
}
// 1/sqrt(a) to ~2^-39 (one Goldschmidt step)
AMT_HD double rsqrt_40(double a) {
    const double y = mufu_rsqrt(a);
    const double g = a * y;
    const double r = fma(-g, half_of(y), 0.5);       // (1 - a y^2) / 2
    return fma(y, r, y);                             // y (1 + r): 1 DMUL + 2 DFMA
}
AMT_HD double sqrt_fast(double a) {
    double s, h;
    sqrt_rsqrt(a, s, h);
    return s;
}
// 1/sqrt(a) to ~1 ulp (no residual step on g needed).
AMT_HD double rsqrt_fast(double a) {
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

// ---------------------------------------------------------------------------------------
// atan2 in DEGREES without octant reduction.  For a = |y|, b = |x| the "diamond" coordinate
// w = a/(a+b) in [0,1] is monotonic in the angle; with s = round(64 w)/64 the table direction
// (1-s, s) lies within 1/64 rad of (b, a) and
//     atan2(a, b) = theta(s) + atan(t),   t = (a(1-s) - b s) / (b(1-s) + a s)
//                                           = (a - s(a+b)) / (b + s(a-b)),      |t| <= 1/64 + 2^-19.
// No swap of the arguments, hence no compare / select pairs; the quadrant of x is an offset
// into the table (theta or 180 - theta) plus a sign flip of t, the sign of y a final copysign --
// both integer operations on the high word.  atan(t) = t - t^3/3 + t^5/5 (next term t^7/7 <=
// 3.2e-14 rad = 1.9e-12 deg); the table and the series carry the factor 180/pi, so every
// arctangent comes out in degrees without a final multiply (all consumers want degrees:
// lat/lon/MLat, MLT = smlon/15 + 12, elevation).
// FP64 instructions: 14 (+2 MUFU); the version with octant selects had 19 + 6 FSEL + 2 ISETP.
// ---------------------------------------------------------------------------------------
#define AMT_ATAN_TABLE \
    0.0, 0.9093804491991414, 1.8476102659945957, \
    2.815556684211228, 3.8140748342903543, 4.844000375080679, \
    5.9061411137704996, 7.001267557495338, 8.130102354155978, \
    9.293308599397115, 10.491477012331599, 11.725112015165077, \
    12.994616791916505, 14.300277449185588, 15.642246457208728, \
    17.020525611519854, 18.43494882292201, 19.88516511385544, \
    21.370622269343183, 22.890551656248327, 24.443954780416536, \
    26.029592191513455, 27.64597536373868, 29.291362170984254, \
    30.96375653207352, 32.6609127216738, 34.380344723844864, \
    36.119340849479755, 37.874983651098205, 39.64417495714481, \
    41.42366562500265, 43.21008939175393, 45.0, \
    46.78991060824607, 48.57633437499735, 50.35582504285519, \
    52.125016348901795, 53.880659150520245, 55.619655276155136, \
    57.3390872783262, 59.036243467926475, 60.70863782901574, \
    62.35402463626132, 63.97040780848654, 65.55604521958347, \
    67.10944834375168, 68.62937773065681, 70.11483488614456, \
    71.56505117707799, 72.97947438848014, 74.35775354279127, \
    75.69972255081441, 77.0053832080835, 78.27488798483492, \
    79.5085229876684, 80.70669140060288, 81.86989764584402, \
    82.99873244250466, 84.0938588862295, 85.15599962491932, \
    86.18592516570965, 87.18444331578877, 88.15238973400541, \
    89.09061955080085, 90.0, \
    /* 65..127: padding, the x < 0 half starts at index 128 (one shift of the sign bit) */ \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    0.0, 0.0, 0.0, \
    180.0, 179.09061955080085, 178.1523897340054, \
    177.18444331578877, 176.18592516570965, 175.15599962491933, \
    174.0938588862295, 172.99873244250466, 171.86989764584402, \
    170.7066914006029, 169.5085229876684, 168.27488798483492, \
    167.0053832080835, 165.6997225508144, 164.35775354279127, \
    162.97947438848016, 161.56505117707798, 160.11483488614456, \
    158.6293777306568, 157.10944834375167, 155.55604521958347, \
    153.97040780848656, 152.35402463626133, 150.70863782901574, \
    149.03624346792648, 147.3390872783262, 145.61965527615513, \
    143.88065915052024, 142.1250163489018, 140.3558250428552, \
    138.57633437499734, 136.78991060824606, 135.0, \
    133.21008939175394, 131.42366562500266, 129.6441749571448, \
    127.8749836510982, 126.11934084947976, 124.38034472384487, \
    122.66091272167381, 120.96375653207352, 119.29136217098426, \
    117.64597536373867, 116.02959219151346, 114.44395478041653, \
    112.89055165624832, 111.37062226934319, 109.88516511385544, \
    108.43494882292201, 107.02052561151986, 105.64224645720873, \
    104.30027744918559, 102.9946167919165, 101.72511201516508, \
    100.4914770123316, 99.29330859939712, 98.13010235415598, \
    97.00126755749534, 95.9061411137705, 94.84400037508068, \
    93.81407483429035, 92.81555668421123, 91.84761026599459, \
    90.90938044919915, 90.0, \

#ifdef __CUDACC__
__constant__ double c_atan_tab[193] = {
AMT_ATAN_TABLE
};
// Constants whose low mantissa word is non-zero cannot be instruction immediates; kept in
// constant memory they become direct c[bank][offset] operands of DFMA/DADD/DMUL instead of two
// UMOV / IMAD.MOV each.
__constant__ double c_fm[4] = {
    11.459155902616464, -19.098593171027442, 57.29577951308232,           // (180/pi) * {1/5, -1/3, 1}
    24.0 / 360.0,                                                         // hours per degree (smLonToMLT)
};
#endif
static const double h_atan_tab[193] = {
AMT_ATAN_TABLE
};
static const double h_fm[4] = {11.459155902616464, -19.098593171027442, 57.29577951308232, 24.0 / 360.0};

// entry at BYTE offset `off` (index * 8 + 1024 for the x < 0 half): the callers form the offset with one
// multiply-add from the table index and the shifted sign bit
AMT_HD double atan_tab_at(unsigned off) {
#ifdef __CUDA_ARCH__
    return *(const double*)((const char*)c_atan_tab + off);
#else
    return h_atan_tab[off >> 3];
#endif
}
AMT_HD double fm_const(int i) {
#ifdef __CUDA_ARCH__
    return c_fm[i];
#else
    return h_fm[i];
#endif
}

AMT_HD double fabs_bits(double x) {        // |x| on the integer pipe, not the FP64 pipe
    return bits_to_double(hi_word(x) & 0x7fffffffu, lo_word(x));
}

// Angle of (b, a), a >= 0, b >= 0 (not both zero).  `xs` = sign bit of the caller's x (0 or
// 0x80000000): it selects the table half (theta or 180 - theta, 1024 bytes further) and flips the
// sign of the remainder -- the caller's quadrant logic in two integer instructions.
AMT_HD double atan_diamond_deg(double a, double b, unsigned xs) {
    const double sum = a + b;
    const double dif = a - b;
    const double w = a * mufu_rcp(sum);
    // w + 1.5*2^46 has an ulp of exactly 1/64: the sum IS w rounded to a multiple of 1/64,
    // and its low mantissa word is the integer 64*s
    const double magic = 105553116266496.0;      // 1.5 * 2^46
    const double wi = w + magic;
    // w is in [0, 1 + 2^-19]; the unsigned clamp also keeps NaN inputs (any bit pattern) inside the table
    const unsigned off = min(lo_word(wi), 64u) * 8u + (xs >> 21);
    const double s = wi - magic;                 // == i/64 exactly (w is in [0, 1])
    const double num = fma(-s, sum, a);
    const double den = fma(s, dif, b);
    double t = div_38(num, den);
    t = bits_to_double(hi_word(t) ^ xs, lo_word(t));
    const double z = t * t;
    double p = fm_const(0);
    p = fma(p, z, fm_const(1));
    p = fma(p, z, fm_const(2));
    return fma(t, p, atan_tab_at(off));
}

// The diamond angle is >= 0 (theta >= 0 and the remainder cannot take it below: s = 0 means t >= 0, and
// the x < 0 half starts at 180 - ...), so the sign of y is ORed in: one LOP3, no mask of the old sign.
AMT_HD double with_sign_of(double r, unsigned ysign) { return bits_to_double(hi_word(r) | ysign, lo_word(r)); }

// atan2(y, x) in degrees, any quadrant, finite inputs, not both zero.
AMT_HD double atan2_deg(double y, double x) {
    const unsigned ys = hi_word(y) & 0x80000000u, xs = hi_word(x) & 0x80000000u;
    const double r = atan_diamond_deg(bits_to_double(hi_word(y) ^ ys, lo_word(y)),
                                      bits_to_double(hi_word(x) ^ xs, lo_word(x)), xs);
    return with_sign_of(r, ys);
}

// atan2(y, x) in degrees for x >= 0 (result in [-90, 90]); also serves atan(y/x) and asin.
AMT_HD double atan2_posx_deg(double y, double x) {
    const unsigned ys = hi_word(y) & 0x80000000u;
    const double r = atan_diamond_deg(bits_to_double(hi_word(y) ^ ys, lo_word(y)), x, 0u);
    return with_sign_of(r, ys);
}

}  // namespace amt
