// Device-side FP64 math for the georeference + regrid path (sm_100a).
//
// Every function cites the reference lines whose arithmetic it reproduces (paths relative
// to the reference tree esa/auromat v1.0.8).  This translation unit is compiled with
// -fmad=false: a*b+c in the source stays a rounded multiply followed by a rounded add, which
// is what the reference's numpy passes do; fused operations appear only where written
// explicitly as fma()/__fma_rn().
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/auromat_b200.h"
#include "amt_fastmath.cuh"

namespace amt {

constexpr double kRad2Deg = 57.29577951308232;   // numpy rad2deg multiplier: 180/pi
constexpr double kDeg2Rad = 0.017453292519943295; // numpy deg2rad multiplier: pi/180

AMT_HD double qnan() { return bits_to_double(0x7ff80000u, 0u); }

// Device copy of the per-frame constants with everything pre-digested for the kernels.
struct FrameC {
    int W, H;
    int fast_center, origin_inside;
    double crpix0, crpix1;
    double cd[4];
    double rot[9];
    double cam[3];          // lineOrigin
    double rad[3];          // (1/a, 1/a, 1/b)          intersection.py:66
    double otr[3];          // (-cam) * rad             intersection.py:63,68
    double oDO;             // otr . otr                intersection.py:74
    double m_geo[9];
    double m_sm[9];
    double a, b, e2a, d;    // Bowring constants        transform.py:254-255,290
    double b_over_a;        // 2 b/a (see bowring)
    int sip_oa, sip_ob;
    int model;
    double as_xc, as_yc, as_k, as_rot;   // all-sky fisheye model (mapping/miracle.py:314-347)
    // Pure TAN headers: wcs.py:93-144 collapses to an AFFINE map from the pixel index to the
    // (un-normalised) celestial direction, dir = R (-y, x, 180/pi), (x, y) = CD (u, v):
    //     corner (ix, iy):  dir_k = aff_a + aff_b * ix + aff_c * iy      (pixel (ix - 0.5, iy - 0.5))
    //     centre (ix, iy):  dir_c = dir_k + aff_h                          (aff_h = (aff_b + aff_c) / 2)
    // coefficients formed on the host in extended precision (fill_frame).
    double aff_a[3], aff_b[3], aff_c[3], aff_h[3];
    // SIP frames: rigorous bounds of the polynomial displacement over the pixel array, in pixels
    // (sum |A_pq| U^p V^q with U, V the largest |u|, |v| of the frame): pixel (x, y) looks where the pure TAN
    // model looks at (x + fx, y + fy), |fx| <= sip_dx, |fy| <= sip_dy (k_limb_bits_sip)
    double sip_dx, sip_dy;
};

// ---------------------------------------------------------------------------------------
// Stage 1: pixel -> native TAN plane -> unit vector -> celestial (ICRS) direction.
// coordinates/wcs.py:93-144.  The reference goes (x,y) -> (phi,theta) via atan2/atan and
// back to Cartesian via sin/cos; the composition is the algebraic identity
//     (cos t cos p, cos t sin p, sin t) = (-y, x, 180/pi) / sqrt(x^2 + y^2 + (180/pi)^2)
// with p = atan2(x,-y), t = atan((180/pi)/r), which we evaluate directly: no transcendental
// call, <= 2 ulp from the reference's value.
// ---------------------------------------------------------------------------------------
// f(u,v) = sum_{p+q<=order} C_pq u^p v^q, Horner in v inside Horner in u, highest power first
// (the evaluation order of oracle/_sip_poly; each Horner step is ONE fused multiply-add here -- half the FP64
// instructions of multiply + add and one rounding instead of two: both sides are held to <= 2 ulp of the
// exact rational value of the polynomial, tests/test_gpu_parity.py::test_sip_device_polynomial_...).  Coefficients: packed triangular, index(p,q) = p*(order+1) - p*(p-1)/2 + q,
// staged in shared memory by the kernel (every lane reads the same word: broadcast).
template <int ORDER>
AMT_HD double sip_poly_fixed(const double* __restrict__ c, double u, double v) {
    double acc = 0.0;
#pragma unroll
    for (int p = ORDER; p >= 0; --p) {
        const int base = p * (ORDER + 1) - (p * (p - 1)) / 2;
        double inner = 0.0;
#pragma unroll
        for (int q = ORDER - p; q >= 0; --q) inner = fma(inner, v, c[base + q]);
        acc = fma(acc, u, inner);
    }
    return acc;
}

AMT_HD double sip_poly(const double* __restrict__ c, int order, double u, double v) {
    switch (order) {
        case 2: return sip_poly_fixed<2>(c, u, v);
        case 3: return sip_poly_fixed<3>(c, u, v);
        case 4: return sip_poly_fixed<4>(c, u, v);
        case 5: return sip_poly_fixed<5>(c, u, v);
        default: break;
    }
    double acc = 0.0;
    for (int p = order; p >= 0; --p) {
        const int base = p * (order + 1) - (p * (p - 1)) / 2;
        double inner = 0.0;
        for (int q = order - p; q >= 0; --q) inner = fma(inner, v, c[base + q]);
        acc = fma(acc, u, inner);
    }
    return acc;
}

AMT_HD void sip_distort(const double* __restrict__ ca, int oa,
                                            const double* __restrict__ cb, int ob,
                                            double& u, double& v) {
    const double fu = sip_poly(ca, oa, u, v);
    const double fv = sip_poly(cb, ob, u, v);
    u = u + fu;
    v = v + fv;
}

template <bool NORMALISE, bool SIP = true>
AMT_HD void pix2dir(const FrameC& f, const double* __restrict__ sip_a,
                                        const double* __restrict__ sip_b,
                                        double px, double py, double dir[3]) {
    // wcs.py:93-99: (px - CRPIX1) + 1, 0-based pixel coordinates
    double u = (px - f.crpix0) + 1.0;
    double v = (py - f.crpix1) + 1.0;
    if (SIP && (f.sip_oa | f.sip_ob)) sip_distort(sip_a, f.sip_oa, sip_b, f.sip_ob, u, v);
    // wcs.py:102  xy = CD . pxy
    const double x = fma(f.cd[0], u, f.cd[1] * v);
    const double y = fma(f.cd[2], u, f.cd[3] * v);
    // wcs.py:111-144 collapsed (see header comment).  The intersection and everything after
    // it are invariant to the length of the direction ("not required to be unit vectors",
    // intersection.py:152); only the fast-centre path, which averages four *unit* corner
    // directions (astrometry.py:61-62), needs the normalisation.
    const double K = 180.0 / 3.141592653589793;
    double l = -y, m = x, n = K;
    if (NORMALISE) {
        const double inv = rsqrt_fast(fma(x, x, fma(y, y, K * K)));
        l *= inv; m *= inv; n *= inv;
    }
    // wcs.py:142  lmnrot = R . lmn
    dir[0] = fma(f.rot[2], n, fma(f.rot[1], m, f.rot[0] * l));
    dir[1] = fma(f.rot[5], n, fma(f.rot[4], m, f.rot[3] * l));
    dir[2] = fma(f.rot[8], n, fma(f.rot[7], m, f.rot[6] * l));
}

// Corner and centre direction of pixel index (ix, iy) for the point kernels (directions are not
// normalised there).  Pure TAN frames use the affine form -- 9 FP64 instructions for both rays
// instead of 2 x 19 -- everything else the polynomial path.  Every kernel that evaluates a ray of
// a frame (hit test, limb solver, outline statistics, georeference, fused binning) goes through
// this one function, so their hit decisions and coordinates agree bit for bit.
template <bool SIP>
AMT_HD void dirs_kc(const FrameC& f, const double* __restrict__ sip_a, const double* __restrict__ sip_b,
                    int ix, int iy, double dk[3], double dc[3]) {
    const double fx = (double)ix, fy = (double)iy;
    if (SIP && (f.sip_oa | f.sip_ob)) {
        pix2dir<false, true>(f, sip_a, sip_b, fx - 0.5, fy - 0.5, dk);
        pix2dir<false, true>(f, sip_a, sip_b, fx, fy, dc);
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dk[k] = fma(fx, f.aff_b[k], fma(fy, f.aff_c[k], f.aff_a[k]));
        dc[k] = dk[k] + f.aff_h[k];
    }
}

// Host: coefficients of the affine ray model from (crpix, cd, rot), in extended precision.
//   u = ix - 0.5 - crpix0 + 1,  v = iy - 0.5 - crpix1 + 1   (corner of pixel index (ix, iy))
//   dir = rot[:,1] x - rot[:,0] y + rot[:,2] K,  x = cd0 u + cd1 v,  y = cd2 u + cd3 v
inline void fill_sip_bounds(FrameC& f, const double* sip_a, const double* sip_b) {
    const double U = fmax(fabs(0.5 - f.crpix0), fabs((double)f.W + 1.0 - f.crpix0)) + 1.0;
    const double V = fmax(fabs(0.5 - f.crpix1), fabs((double)f.H + 1.0 - f.crpix1)) + 1.0;
    const double* coef[2] = {sip_a, sip_b};
    const int order[2] = {f.sip_oa, f.sip_ob};
    double bound[2] = {0.0, 0.0};
    for (int w = 0; w < 2; ++w) {
        // packed triangular coefficients of sip_poly: index(p, q) = p (order + 1) - p (p - 1) / 2 + q
        for (int pp = 0; pp <= order[w]; ++pp)
            for (int q = 0; pp + q <= order[w]; ++q)
                bound[w] += fabs(coef[w][pp * (order[w] + 1) - (pp * (pp - 1)) / 2 + q]) * pow(U, pp) * pow(V, q);
    }
    f.sip_dx = bound[0];
    f.sip_dy = bound[1];
}

inline void fill_affine(FrameC& f) {
    const long double K = 180.0L / 3.14159265358979323846264338327950288L;
    const long double u0 = 0.5L - (long double)f.crpix0, v0 = 0.5L - (long double)f.crpix1;
    for (int k = 0; k < 3; ++k) {
        const long double r0 = f.rot[3 * k + 0], r1 = f.rot[3 * k + 1], r2 = f.rot[3 * k + 2];
        const long double b = r1 * (long double)f.cd[0] - r0 * (long double)f.cd[2];
        const long double c = r1 * (long double)f.cd[1] - r0 * (long double)f.cd[3];
        f.aff_b[k] = (double)b;
        f.aff_c[k] = (double)c;
        f.aff_a[k] = (double)(r2 * K + b * u0 + c * v0);
        f.aff_h[k] = (double)((b + c) * 0.5L);
    }
}

// numpy floor_divide / remainder for doubles (npy_divmod), used by the astropy-style wrap
__device__ __forceinline__ double np_floor_divide(double a, double b) {
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0 && ((b < 0.0) != (mod < 0.0))) div -= 1.0;
    if (div != 0.0) {
        double fl = floor(div);
        if (div - fl > 0.5) fl += 1.0;
        return fl;
    }
    return copysign(0.0, a / b);
}

// All-sky fisheye camera (mapping/miracle.py:314-347 calculateAzEl, :240-258 direction):
// (row, col) -> azimuth / elevation -> local Cartesian -> ECEF via `rot`.  `px, py` follow
// this file's convention (pixel centres at integers); the reference counts corners at
// integers, hence the +0.5.  Cold path (small frames): plain libm in the reference's order.
__device__ __forceinline__ double pix2dir_allsky(const FrameC& f, double px, double py, double dir[3]) {
    const double v0 = (py + 0.5) - f.as_xc;           // "X is vertical" (cal.txt)
    const double v1 = (px + 0.5) - f.as_yc;
    double az = atan2(v1, -v0);                       // signedAngleBetween(vecs, [-1, 0])
    az = az - f.as_rot;
    // Angle(az rad).wrap_at(360 deg): az - floor(az / 2pi) * 2pi, then to degrees
    const double two_pi = 6.283185307179586;
    const double wraps = np_floor_divide(az, two_pi);
    if (wraps == wraps && wraps != 0.0 && !isinf(wraps)) {
        az = az - wraps * two_pi;
        if (az >= two_pi) az -= two_pi;
        if (az < 0.0) az += two_pi;
    }
    const double az_deg = az * kRad2Deg;
    const double dist = sqrt(v0 * v0 + v1 * v1);
    const double el_deg = 90.0 - (dist / f.as_k) * kRad2Deg;
    const double el = el_deg * kDeg2Rad;
    const double azp = (-(az_deg - 180.0)) * kDeg2Rad;
    double se, ce, sa, ca;
    sincos(el, &se, &ce);
    sincos(azp, &sa, &ca);
    const double l = ce * ca, m = ce * sa, n = se;    // spherical_to_cartesian(1, el, az)
    dir[0] = (f.rot[0] * l + f.rot[1] * m) + f.rot[2] * n;
    dir[1] = (f.rot[3] * l + f.rot[4] * m) + f.rot[5] * n;
    dir[2] = (f.rot[6] * l + f.rot[7] * m) + f.rot[8] * n;
    return el_deg;
}

// ---------------------------------------------------------------------------------------
// Stage 2a: directed ray / inflated-ellipsoid intersection in the J2000 frame.
// coordinates/intersection.py:58-104:  D = dir/axes, O = -cam/axes,
//   rootTerm = (D.O)^2 - (O.O)(D.D) + (D.D);  t = (D.O -+ sqrt(rootTerm)) / (D.D);  P = dir t + cam
// This is the one ill-conditioned block of the path (grazing rays: rootTerm cancels).  The dot
// products and rootTerm are accumulated with FMA -- each partial result is rounded once
// instead of twice, i.e. at least as accurate as the reference's separate multiply/add passes --
// and sqrt / divide are the <= 1 ulp Goldschmidt / Newton forms.  Returns false when the ray
// misses (reference: NaN row).  `graze` is set when the normalised discriminant rootTerm/(D.D)
// is below kGrazeThreshold: there 1 ulp of input noise moves the footprint by > 1e-9 deg, in
// the reference as much as here.
// ---------------------------------------------------------------------------------------
constexpr double kGrazeThreshold = 1e-10;

AMT_HD bool intersect(const FrameC& f, const double dir[3], double P[3], bool& graze) {
    const double D0 = dir[0] * f.rad[0], D1 = dir[1] * f.rad[1], D2 = dir[2] * f.rad[2];
    const double dDO = fma(D2, f.otr[2], fma(D1, f.otr[1], D0 * f.otr[0]));
    const double dDD = fma(D2, D2, fma(D1, D1, D0 * D0));
    const double rt = fma(dDO, dDO, fma(-f.oDO, dDD, dDD));
    graze = rt >= 0.0 && rt < kGrazeThreshold * dDD;
    if (!(rt >= 0.0)) return false;                  // sqrt of a negative -> NaN row
    const double root = rt > 0.0 ? sqrt_fast(rt) : 0.0;
    double t = f.origin_inside ? dDO + root : dDO - root;
    if (!(t >= 0.0)) return false;                   // intersection.py:50-56 (behind the camera)
    t = div_fast(t, dDD);
    P[0] = fma(dir[0], t, f.cam[0]);                 // res = direction*dMin - (-lineOrigin)
    P[1] = fma(dir[1], t, f.cam[1]);
    P[2] = fma(dir[2], t, f.cam[2]);
    return true;
}

// The same point for a ray that is KNOWN to hit (the fused kernel reads the frame's final validity
// bitmaps, which were produced with this very discriminant): no miss / behind-the-camera tests, no
// grazing-ray bookkeeping, no branch -- the arithmetic of `intersect`, instruction for instruction, so
// the coordinates are bit-identical.  For a ray that does not hit, the result is garbage or NaN; the
// caller overwrites it.  `sgn` = +1 if the origin is inside the ellipsoid, else -1 (t = dDO + sgn root).
AMT_HD void intersect_valid(const FrameC& f, const double dir[3], double P[3]) {
    const double D0 = dir[0] * f.rad[0], D1 = dir[1] * f.rad[1], D2 = dir[2] * f.rad[2];
    const double dDO = fma(D2, f.otr[2], fma(D1, f.otr[1], D0 * f.otr[0]));
    const double dDD = fma(D2, D2, fma(D1, D1, D0 * D0));
    const double rt = fma(dDO, dDO, fma(-f.oDO, dDD, dDD));
    const double root = rt > 0.0 ? sqrt_fast(rt) : 0.0;
    const double sgn = f.origin_inside ? 1.0 : -1.0;            // uniform
    const double t = div_fast(fma(sgn, root, dDO), dDD);        // dDO +- root, exactly
    P[0] = fma(dir[0], t, f.cam[0]);
    P[1] = fma(dir[1], t, f.cam[1]);
    P[2] = fma(dir[2], t, f.cam[2]);
}

// Hit test alone (validity bitmaps without any coordinate plane): same discriminant as
// `intersect`, no square root and no division.  With the origin outside the ellipsoid
// (oDO > 1) root^2 = dDO^2 - dDD (oDO - 1) < dDO^2, so sign(dDO - root) = sign(dDO); with the
// origin inside, dDO + root >= 0 always.
AMT_HD bool intersect_hit(const FrameC& f, const double dir[3], bool& graze) {
    const double D0 = dir[0] * f.rad[0], D1 = dir[1] * f.rad[1], D2 = dir[2] * f.rad[2];
    const double dDO = fma(D2, f.otr[2], fma(D1, f.otr[1], D0 * f.otr[0]));
    const double dDD = fma(D2, D2, fma(D1, D1, D0 * D0));
    const double rt = fma(dDO, dDO, fma(-f.oDO, dDD, dDD));
    graze = rt >= 0.0 && rt < kGrazeThreshold * dDD;
    if (!(rt >= 0.0)) return false;
    if (f.origin_inside) return true;
    // exactly the t >= 0 test of `intersect` (t = dDO - root), evaluated only in the rare case
    // where the sign is not obvious
    if (dDO * dDO > 4.0 * rt) return dDO >= 0.0;
    return (dDO - (rt > 0.0 ? sqrt_fast(rt) : 0.0)) >= 0.0;
}

AMT_HD void mat3(const double* __restrict__ M, const double v[3], double o[3]) {
    o[0] = fma(M[2], v[2], fma(M[1], v[1], M[0] * v[0]));
    o[1] = fma(M[5], v[2], fma(M[4], v[1], M[3] * v[0]));
    o[2] = fma(M[8], v[2], fma(M[7], v[1], M[6] * v[0]));
}

// ---------------------------------------------------------------------------------------
// Stage 2b: ECEF -> geodetic, single-iteration Bowring 1985 (coordinates/transform.py:252-297).
//   p = sqrt(x^2+y^2), r = sqrt(p^2+z^2), tu = b z (1 + d/r)/(a p), cu3 = (1+tu^2)^(-3/2),
//   lat = atan((z + d cu3 tu^3)/(p - e^2 a cu3)), lon = atan2(y, x)
// p is a main term of the latitude (<= 1 ulp: Goldschmidt pair with residual step); 1/p, 1/r and
// cu = (1+tu^2)^(-1/2) only scale the e^2-sized corrections (2^-39 each is 1e-14 rad on the
// result).  No division besides the two inside the arctangents:
//   tu = (b/a) (z / p) (1 + d / r),  d su3 = d (tu cu)^3.  Output in DEGREES; r2 = |(x,y,z)|^2
// is handed back for the elevation (|P|^2 is invariant under the J2000 -> GEO rotation).
// ---------------------------------------------------------------------------------------
AMT_HD void bowring(double b_over_a, double e2a, double d, double x, double y, double z,
                    double& lat, double& lon, double& r2) {
    const double p2 = fma(x, x, y * y);
    double p, hp;
    sqrt_rsqrt(p2, p, hp);                       // hp = 0.5/p
    r2 = fma(z, z, p2);
    const double ir = rsqrt_40(r2);
    const double tu = ((b_over_a * z) * hp) * fma(d, ir, 1.0);        // b_over_a = 2 b/a, hp = 0.5/p: bit-identical to (b/a z)(2 hp)
    const double cu = rsqrt_40(fma(tu, tu, 1.0));
    const double tc = tu * cu;
    const double cu3 = (cu * cu) * cu;
    const double su3 = (tc * tc) * tc;
    lat = atan2_posx_deg(fma(d, su3, z), fma(-e2a, cu3, p));
    lon = atan2_deg(y, x);
}

// Reference-order Bowring with libm, used by the (cold) rotatePole path where the result
// feeds a bit-exact binning comparison against numpy.
__device__ __forceinline__ void bowring_ref(double a, double b, double e2a, double d,
                                            double x, double y, double z, double& lat, double& lon) {
    const double p2 = x * x + y * y;
    const double p = sqrt(p2);
    const double r = sqrt(p2 + z * z);
    double tu = d / r;
    tu = tu + 1.0;
    tu = tu * b;
    tu = tu * z;
    tu = tu / (a * p);
    const double tu2 = tu * tu;
    double c = 1.0 / sqrt(1.0 + tu2);
    const double cu3 = (c * c) * c;
    double su3 = tu * cu3;
    su3 = su3 * tu2;
    double tp = d * su3 + z;
    const double pm = p - cu3 * e2a;
    tp = tp / pm;
    lat = atan(tp);
    lon = atan2(y, x);
}

// geodetic -> ECEF, coordinates/transform.py:156-178 (radians in).
__device__ __forceinline__ void geodetic2ecef(double a, double e2, double lat, double lon, double h,
                                              double& x, double& y, double& z) {
    double sl, cl, so, co;
    sincos(lat, &sl, &cl);
    sincos(lon, &so, &co);
    const double n = a / sqrt(1.0 - e2 * (sl * sl));
    const double nh = n + h;
    x = (nh * cl) * co;
    y = (nh * cl) * so;
    z = (n * (1.0 - e2) + h) * sl;
}

// SM Cartesian -> (MLat deg, MLT h): transform.py:104-127 + :419-430 + :373-386.
AMT_HD void sm_to_mlat_mlt(const double S[3], double& mlat, double& mlt) {
    const double s = sqrt_fast(fma(S[0], S[0], S[1] * S[1]));
    const double smlon = atan2_deg(S[1], S[0]);
    mlat = atan2_posx_deg(S[2], s);
    mlt = fma(smlon, fm_const(3), 12.0);
}

// elevation, mapping/astrometry.py:200-212 + utils.py:28-46: 90 - acos(clip(-dir . P/|P|)) with
// `dir` a unit vector, i.e. the angle between -dir and the plane normal to P:
//     e = atan2(c, sqrt(dd |P|^2 - c^2)),   c = -dir . P,  dd = |dir|^2
// -- no normalisation, no clip (the root is clamped at 0 instead), and better conditioned near
// the nadir than acos.  UNIT_DIR: dd := 1, the reference's own behaviour for fast centres, whose
// direction is the un-normalised mean of four unit corner directions (astrometry.py:61-62).
template <bool UNIT_DIR>
AMT_HD double elevation_deg(const double dir[3], const double P[3], double r2) {
    const double c = -fma(dir[2], P[2], fma(dir[1], P[1], dir[0] * P[0]));
    double n = r2;
    if (!UNIT_DIR) n *= fma(dir[2], dir[2], fma(dir[1], dir[1], dir[0] * dir[0]));
    const double w = fma(-c, c, n);
    if (!(w > 0.0)) return c == c ? copysign(90.0, c) : c;
    return atan2_posx_deg(c, sqrt_fast(w));
}

// One J2000 intersection point -> lat/lon [deg] (+ MLat/MLT).  transform.py:324-343,403-430.
AMT_HD void point_to_geo(const FrameC& f, const double P[3], double& lat, double& lon, double& r2) {
    double G[3];
    mat3(f.m_geo, P, G);
    bowring(f.b_over_a, f.e2a, f.d, G[0], G[1], G[2], lat, lon, r2);
}
AMT_HD void point_to_geo(const FrameC& f, const double P[3], double& lat, double& lon) {
    double r2;
    point_to_geo(f, P, lat, lon, r2);
}

AMT_HD void point_to_mag(const FrameC& f, const double P[3], double& mlat, double& mlt) {
    double S[3];
    mat3(f.m_sm, P, S);
    sm_to_mlat_mlt(S, mlat, mlt);
}

// ---------------------------------------------------------------------------------------
// Stage 3 helpers: numpy semantics reproduced bit for bit.
// ---------------------------------------------------------------------------------------
// numpy.linspace(lo, hi, n+1)[k]  (util/histogram.py:185-186): fl(fl(k*step) + lo), last == hi
__device__ __forceinline__ double edge_at(double lo, double hi, double step, int n, int k) {
    return k == n ? hi : __dadd_rn(__dmul_rn((double)k, step), lo);
}

__device__ __forceinline__ double next_up(double x) {      // toward +inf, finite x
    if (x == 0.0) return __longlong_as_double(1LL);
    const long long b = __double_as_longlong(x);
    return __longlong_as_double(x > 0.0 ? b + 1 : b - 1);
}
__device__ __forceinline__ double next_down(double x) { return -next_up(-x); }

// searchsorted(edges, x, 'right') - 1 with the right-edge fix of util/histogram.py:205-224
// and outliers mapped to -1.  With NEAR, `near` is set when x lies within 1 ulp of one of
// the bin edges that decided its cell (the samples the parity statement allows to differ
// when lat/lon themselves differ in the last bit).
template <bool NEAR>
__device__ __forceinline__ int bin_index(double x, double lo, double hi, double step, double inv_step,
                                         int n, double round_scale, double eps, bool& near) {
    near = false;
    if (!NEAR) {
        // Shortcut: q = (x - lo)/step carries an error of a few ulp of q (<= n * 4e-16), and the
        // numpy edges fl(fl(k*step) + lo) differ from lo + k*step by <= 2 ulp of max(|lo|,|hi|),
        // i.e. by far less than `eps` cells (fill_grid sizes eps from exactly these bounds and
        // sets it to 1 when the grid is too fine for the argument).  A sample whose fractional
        // position is at least eps away from both ends of its cell is therefore on the same
        // side of the true edges as of the ideal ones: the floor IS searchsorted(..)-1.
        // Evaluated without floor / convert / range compares: qh = q - 1/2 rounded to the nearest
        // integer by adding 1.5 * 2^52 is floor(q) (ties only for fr = 0, rejected below); the integer
        // sits in the low mantissa word, and the high word equals that of the magic constant exactly
        // when 0 <= floor(q) < 2^32 -- so two integer compares replace x >= lo, x < hi and k < n, and a
        // NaN or an out-of-range sample falls through to the exact path.
        const double magic = 6755399441055744.0;         // 1.5 * 2^52
        const double qh = fma(__dsub_rn(x, lo), inv_step, -0.5);
        const double tq = __dadd_rn(qh, magic);
        const double d = __dsub_rn(qh, __dsub_rn(tq, magic));      // q - (floor(q) + 1/2)
        const unsigned kq = lo_word(tq);
        if (fabs(d) < 0.5 - eps && hi_word(tq) == 0x43380000u && kq < (unsigned)n) return (int)kq;
    }
    if (x >= lo && x < hi) {
        // common case, branch-free: floor guess, then one correction step against the true
        // numpy edges e(k) = fl(fl(k*step) + lo), e(n) = hi
        // lo <= x < hi, so the guess is in 0..n (n only through rounding); no clamp of the double
        // needed: guess n gives e0 = fl(n*step + lo) ~ hi, e1 = hi, and ends at n-1 below
        const double kd = floor(__dmul_rn(__dsub_rn(x, lo), inv_step));
        int k = (int)kd;
        const double e0 = __dadd_rn(__dmul_rn(kd, step), lo);
        const double e1 = k + 1 >= n ? hi : __dadd_rn(__dmul_rn(kd + 1.0, step), lo);
        k += (x >= e1) ? 1 : 0;
        k -= (x < e0) ? 1 : 0;
        k = min(max(k, 0), n - 1);
        // One correction step is exact: (x - lo)*inv_step carries a relative error of a few ulp,
        // i.e. an absolute error <= n * 4e-16 << 1 for any grid that fits in memory, so the
        // floor guess is the true cell or its direct neighbour; numpy's edges differ from
        // lo + k*step by ulps only.
        if (NEAR) {
            const double f0 = edge_at(lo, hi, step, n, k), f1 = edge_at(lo, hi, step, n, k + 1);
            near = (next_down(x) <= f0) || (next_up(x) >= f1);
        }
        return k;
    }
    if (!(x == x)) return -1;                        // NaN sorts last -> outlier
    if (x < lo) {
        if (NEAR) near = next_up(x) >= lo;
        return -1;
    }
    // x >= hi -- util/histogram.py:215-224: np.around(x, decimal) == np.around(hi, decimal)
    const double rx = rint(__dmul_rn(x, round_scale)) / round_scale;
    const double rh = rint(__dmul_rn(hi, round_scale)) / round_scale;
    if (NEAR) near = next_down(x) <= hi;
    return rx == rh ? n - 1 : -1;
}

// Angle(x deg).wrap_at(180 deg).degree (astropy `_wrap_at`): x - floor((x+180)/360)*360
// with the two rounding fix-ups.  Oracle: wrap_at_180().
__device__ __forceinline__ double wrap_at_180(double x) {
    const double wraps = np_floor_divide(x - (-180.0), 360.0);
    if (wraps == wraps && wraps != 0.0 && !isinf(wraps)) {
        x = x - wraps * 360.0;
        if (x >= 180.0) x -= 360.0;
        if (x < -180.0) x += 360.0;
    }
    return x;
}

}  // namespace amt
