// Host-side scalar arithmetic of the path, in C: the target grid of resample()
// (reference resample.py:36-61 plateCarreeResolution, :281-299 fixedGrid, :220-241,330-335 + the
// bin-range bookkeeping of util/histogram.py:185-186,215-219) and the pole-visibility test of a
// WCS frame.  These are line-by-line ports of auromat_b200/resample.py and
// auromat_b200/coordinates/geodesic.py (which follow the reference's own arithmetic): only
// + - * / sqrt floor rint and libm's sin/cos/tan/atan/atan2/log10/pow are used, in the same order, and
// the translation unit is compiled with -ffp-contract=off, so the results are BIT-IDENTICAL to the
// Python functions (tests/test_host_logic.py compares them on thousands of random boxes).  They
// let the sequence engine derive a frame's grid without returning to Python.  No CUDA calls.
#pragma once
#include <cmath>

// Python's float % (floatobject.c float_rem)
static double py_fmod(double v, double w) {
    double mod = fmod(v, w);
    if (mod != 0.0) {
        if ((w < 0) != (mod < 0)) mod += w;
    } else {
        mod = copysign(0.0, w);
    }
    return mod;
}

static const double kPyPi = 3.141592653589793;
static const double kDegToRad = kPyPi / 180.0;      // math.radians(x) = x * (pi / 180)
static const double kRadToDeg = 180.0 / kPyPi;      // math.degrees(x) = x * (180 / pi)

// auromat_b200/coordinates/geodesic.py::_inverse, sigma only (geographiclib's a12 in radians)
static double vincenty_sigma(double lat1, double lon1, double lat2, double lon2) {
    if (lat1 == lat2 && lon1 == lon2) return 0.0;
    const double f = 1 / 298.257223563;
    const double U1 = atan((1 - f) * tan(lat1 * kDegToRad));
    const double U2 = atan((1 - f) * tan(lat2 * kDegToRad));
    double L = (lon2 - lon1) * kDegToRad;
    L = py_fmod(L + kPyPi, 2 * kPyPi) - kPyPi;
    const double sU1 = sin(U1), cU1 = cos(U1), sU2 = sin(U2), cU2 = cos(U2);
    double lam = L, sigma = 0.0;
    for (int it = 0; it < 200; ++it) {
        const double sl = sin(lam), cl = cos(lam);
        const double t1 = cU2 * sl, t2 = cU1 * sU2 - sU1 * cU2 * cl;
        const double ss = sqrt(t1 * t1 + t2 * t2);
        if (ss == 0) return 0.0;
        const double cs = sU1 * sU2 + cU1 * cU2 * cl;
        sigma = atan2(ss, cs);
        const double sa = cU1 * cU2 * sl / ss;
        const double c2a = 1 - sa * sa;
        const double c2sm = c2a != 0 ? cs - 2 * sU1 * sU2 / c2a : 0.0;
        const double Cc = f / 16 * c2a * (4 + f * (4 - 3 * c2a));
        const double lam_new = L + (1 - Cc) * f * sa * (sigma + Cc * ss * (c2sm + Cc * cs * (-1 + 2 * c2sm * c2sm)));
        const bool done = fabs(lam_new - lam) < 1e-15;
        lam = lam_new;
        if (done) break;
    }
    return sigma;
}

// resample.py:36-61
extern "C" int amt_plate_carree_resolution(double lat_south, double lon_west, double lat_north, double lon_east,
                                           double arcsec_per_px, double* lat_px_per_deg, double* lon_px_per_deg) {
    CHECK_ARG(lat_px_per_deg && lon_px_per_deg, "amt_plate_carree_resolution: NULL argument");
    const double degPerPx = arcsec_per_px * (1.0 / 3600.0);
    *lat_px_per_deg = 1 / degPerPx;
    const double latMiddle = (lat_north + lat_south) / 2;
    const double dist = vincenty_sigma(latMiddle, lon_west, latMiddle, lon_east) * kRadToDeg;
    const double px = dist / degPerPx;
    const double lons = lon_west > lon_east ? lon_east + 360 - lon_west : lon_east - lon_west;
    *lon_px_per_deg = px / lons;
    return AMT_OK;
}

// np.linspace(start, stop, num)[i] = fl(fl(i*step) + start), last element == stop
static double h_linspace_at(double start, double stop, long long num, long long i) {
    if (i == num - 1) return stop;
    const double step = (stop - start) / (double)(num - 1);
    return (double)i * step + start;
}

// index of the first node that is > x (strict) or >= x; num if there is none
static long long h_first_node(double lo, double hi, long long num, double x, bool strict) {
    const double step = (hi - lo) / (double)(num - 1);
    const double guess = (x - lo) / step;
    long long i = 0;
    if (guess == guess) {
        double fl = floor(guess);
        if (fl < 0) fl = 0;
        if (fl > (double)(num - 1)) fl = (double)(num - 1);
        i = (long long)fl;
    }
    auto ok = [&](double v) { return strict ? v > x : v >= x; };
    while (i > 0 && ok(h_linspace_at(lo, hi, num, i - 1))) --i;
    while (i < num && !ok(h_linspace_at(lo, hi, num, i))) ++i;
    return i;
}
static double h_snap_down(double lo, double hi, long long num, double x) {
    long long i = h_first_node(lo, hi, num, x, true);
    if (i == num) i = 0;
    return h_linspace_at(lo, hi, num, ((i - 1) % num + num) % num);
}
static double h_snap_up(double lo, double hi, long long num, double x) {
    long long i = h_first_node(lo, hi, num, x, false);
    if (i == num) i = 0;
    return h_linspace_at(lo, hi, num, i);
}

// resample.py:281-299 + :220-241,330-335 (auromat_b200/resample.py fixedGrid + targetGrid).
// *fallback = 1 when the bin-edge bookkeeping would need the materialised edges (step within 1e-9 of
// a power of ten): the caller then derives the grid in Python.
extern "C" int amt_target_grid(double lat_px_per_deg, double lon_px_per_deg, double lat_min, double lat_max,
                               double lon_min, double lon_max, amt_grid* g, amt_grid_info* info, int32_t* fallback) {
    CHECK_ARG(g && info && fallback, "amt_target_grid: NULL argument");
    CHECK_ARG(lat_px_per_deg > 0 && lon_px_per_deg > 0, "amt_target_grid: pxPerDeg must be positive");
    *fallback = 0;
    const long long nLatAll = (long long)nearbyint(lat_px_per_deg * 180 + 1);
    const long long nLonAll = (long long)nearbyint(lon_px_per_deg * 360 + 1);
    const double latMinG = h_snap_down(-90.0, 90.0, nLatAll, lat_min);
    const double latMaxG = h_snap_up(-90.0, 90.0, nLatAll, lat_max);
    const double lonMinG = h_snap_down(-180.0, 180.0, nLonAll, lon_min);
    const double lonMaxG = h_snap_up(-180.0, 180.0, nLonAll, lon_max);
    const long long nLat = (long long)nearbyint(lat_px_per_deg * (latMaxG - latMinG) + 1);
    const long long nLon = (long long)nearbyint(lon_px_per_deg * (lonMaxG - lonMinG) + 1);
    if (nLat < 3 || nLon < 3)
        return set_err(AMT_ERR_INVALID_ARGUMENT, "the resampling grid has no interior nodes (nLat=%lld, nLon=%lld)", nLat, nLon);
    CHECK_ARG(nLat < (1LL << 31) && nLon < (1LL << 31), "amt_target_grid: grid too large");
    const double latStep = (latMinG - latMaxG) / (double)(nLat - 1);
    const double lonStep = (lonMaxG - lonMinG) / (double)(nLon - 1);
    const double latC0 = h_linspace_at(latMaxG, latMinG, nLat, 1), latCL = h_linspace_at(latMaxG, latMinG, nLat, nLat - 2);
    const double lonC0 = h_linspace_at(lonMinG, lonMaxG, nLon, 1), lonCL = h_linspace_at(lonMinG, lonMaxG, nLon, nLon - 2);
    memset(g, 0, sizeof *g);
    g->prerotate = AMT_PRE_NONE;
    g->nx = (int32_t)(nLon - 2);
    g->ny = (int32_t)(nLat - 2);
    g->lo_x = lonC0 - lonStep / 2; g->hi_x = lonCL + lonStep / 2;
    g->lo_y = latCL + latStep / 2; g->hi_y = latC0 - latStep / 2;
    g->step_x = (g->hi_x - g->lo_x) / (double)g->nx;
    g->step_y = (g->hi_y - g->lo_y) / (double)g->ny;
    auto round_scale = [&](double step, double* out) {
        const double d = -log10(step);
        if (fabs(d - nearbyint(d)) > 1e-9) {
            *out = pow(10.0, (double)((long long)d + 6));
            return true;
        }
        return false;
    };
    if (!round_scale(g->step_x, &g->round_x) || !round_scale(g->step_y, &g->round_y)) *fallback = 1;
    g->wgs_a = 6378137.0 / 1000;
    g->wgs_b = g->wgs_a * (1 - 1 / 298.257223563);
    // rotation_matrix(deg2rad(90), [1, 0, 0]): unused with AMT_PRE_NONE, filled like the Python path does
    g->rot[0] = 1.0; g->rot[4] = 6.123233995736766e-17; g->rot[5] = -1.0; g->rot[7] = 1.0; g->rot[8] = 6.123233995736766e-17;
    info->n_lat = (int32_t)nLat; info->n_lon = (int32_t)nLon;
    info->lat_min_in_grid = latMinG; info->lat_max_in_grid = latMaxG;
    info->lon_min_in_grid = lonMinG; info->lon_max_in_grid = lonMaxG;
    info->lat_step = latStep; info->lon_step = lonStep;
    info->lat_px_per_deg = lat_px_per_deg; info->lon_px_per_deg = lon_px_per_deg;
    return AMT_OK;
}

// auromat_b200/resample.py::sideScale
extern "C" double amt_side_scale(uint64_t n_samples) {
    if (n_samples < 2) n_samples = 2;
    int c = 0;
    while ((1ULL << c) < n_samples && c < 63) ++c;      // ceil(log2(n))
    int k = 62 - 7 - c;
    if (k < 1) k = 1;
    if (k > 40) k = 40;
    return ldexp(1.0, k);
}

extern "C" int amt_sip_displacement_bound(const amt_frame* fr, double* dx, double* dy) {
    CHECK_ARG(fr && dx && dy, "amt_sip_displacement_bound: NULL argument");
    GeorefParams p;
    memset(&p, 0, sizeof p);
    int rc = fill_frame(fr, p);
    if (rc) return rc;
    *dx = p.f.sip_dx;
    *dy = p.f.sip_dy;
    return AMT_OK;
}

// Where the geographic poles (on the inflated ellipsoid) appear in a WCS frame: the pole point is
// projected through the inverse WCS (TAN, SIP undone by fixed-point iteration).  in_frame[i] != 0 when
// pole i (0: north, 1: south) faces the camera, lies in front of the tangent plane and projects into
// the pixel array; (ix, iy)[i] is then the pixel that contains it.  Replaces the azimuth-sum test on a
// 50-point convex outline (reference mapping/mapping.py:705-718, geodesic.py:183-202): the pole is
// enclosed iff that pixel is defined.
static double h_sip_poly(const double* c, int order, double u, double v) {
    double acc = 0.0;
    for (int p = order; p >= 0; --p) {
        const int base = p * (order + 1) - (p * (p - 1)) / 2;
        double inner = 0.0;
        for (int q = order - p; q >= 0; --q) inner = inner * v + c[base + q];
        acc = acc * u + inner;
    }
    return acc;
}

extern "C" int amt_pole_pixels(const amt_frame* fr, int32_t ix[2], int32_t iy[2], int32_t in_frame[2]) {
    CHECK_ARG(fr && ix && iy && in_frame, "amt_pole_pixels: NULL argument");
    CHECK_ARG(fr->model == AMT_MODEL_WCS, "amt_pole_pixels: WCS frames only");
    const double a = 1.0 / fr->inv_axes[0], b = 1.0 / fr->inv_axes[2];
    const double K = 180.0 / kPyPi;
    const double det = fr->cd[0] * fr->cd[3] - fr->cd[1] * fr->cd[2];
    CHECK_ARG(det != 0.0, "amt_pole_pixels: singular CD matrix");
    for (int i = 0; i < 2; ++i) {
        in_frame[i] = 0;
        ix[i] = iy[i] = -1;
        const double z = (i == 0 ? 1.0 : -1.0) * b;
        // P = m_geo^T (0, 0, z): the pole in J2000
        const double P[3] = {fr->m_geo[6] * z, fr->m_geo[7] * z, fr->m_geo[8] * z};
        const double d[3] = {P[0] - fr->cam[0], P[1] - fr->cam[1], P[2] - fr->cam[2]};
        const double nrm[3] = {P[0] / (a * a), P[1] / (a * a), P[2] / (b * b)};
        const double facing = d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2];
        if ((facing >= 0) != (fr->origin_inside != 0)) continue;         // far side of the ellipsoid
        const double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double u3[3] = {d[0] / len, d[1] / len, d[2] / len};
        // lmn = rot^T . unit(d)
        const double l = fr->rot[0] * u3[0] + fr->rot[3] * u3[1] + fr->rot[6] * u3[2];
        const double m = fr->rot[1] * u3[0] + fr->rot[4] * u3[1] + fr->rot[7] * u3[2];
        const double n = fr->rot[2] * u3[0] + fr->rot[5] * u3[1] + fr->rot[8] * u3[2];
        if (!(n > 0)) continue;
        const double x = K * m / n, y = -K * l / n;
        double u = (fr->cd[3] * x - fr->cd[1] * y) / det, v = (fr->cd[0] * y - fr->cd[2] * x) / det;
        if (fr->sip_order_a || fr->sip_order_b) {
            const double tu = u, tv = v;
            for (int it = 0; it < 30; ++it) {
                u = tu - h_sip_poly(fr->sip_a, fr->sip_order_a, u, v);
                v = tv - h_sip_poly(fr->sip_b, fr->sip_order_b, u, v);
            }
        }
        const double px = u + fr->crpix[0] - 1, py = v + fr->crpix[1] - 1;
        const int w = fr->width, h = fr->height;
        if (px >= -0.5 && px <= w - 0.5 && py >= -0.5 && py <= h - 0.5) {
            int xi = (int)floor(px + 0.5), yi = (int)floor(py + 0.5);
            xi = xi < 0 ? 0 : (xi > w - 1 ? w - 1 : xi);
            yi = yi < 0 ? 0 : (yi > h - 1 ? h - 1 : yi);
            ix[i] = xi; iy[i] = yi;
            in_frame[i] = 1;
        }
    }
    return AMT_OK;
}
