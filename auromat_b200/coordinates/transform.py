"""Per-frame reference-frame rotations (J2000/GEI/GEO/GSE/GSM/SM, after NASA cxform) as 3x3
matrices, computed once per frame on the host; the per-pixel rotations, the Bowring
geodetic conversion and the MLat/MLT conversion run in the CUDA kernels.

Mirrors the scalar part of `auromat/coordinates/transform.py` of the reference
(:489-696) and keeps its function names.  All formulas are evaluated in the same order as
the reference so that the matrices agree bit for bit (tests/test_host_constants.py).
"""
from __future__ import annotations

import math

import numpy as np

from .geodesic import wgs84A, wgs84B
from .igrf import calcG01, calcG11, calcH11

# axis sign convention that matches cxform's hapgood_matrix (reference transform.py:489-494)
X = (-1.0, 0.0, 0.0)
Y = (0.0, 1.0, 0.0)
Z = (0.0, 0.0, -1.0)


def rotation_matrix(angle, direction):
    """3x3 rotation about `direction` by `angle` rad (Rodrigues form; reference
    coordinates/transformations.py:295-336, upper-left block).  Scalar arithmetic in the
    reference's order: R = diag(cos) + outer(d,d)*(1-cos) + skew(d*sin).  (Filled element by
    element: building a 3x3 ndarray from nested lists costs more than the arithmetic.)"""
    sina, cosa = math.sin(angle), math.cos(angle)
    if direction is X or direction is Y or direction is Z:
        dx, dy, dz = direction                       # unit axes: the normalisation is the identity
    else:
        dx, dy, dz = (float(v) for v in direction)
        n = math.sqrt(dx * dx + dy * dy + dz * dz)
        dx, dy, dz = dx / n, dy / n, dz / n
    omc = 1.0 - cosa
    sx, sy, sz = dx * sina, dy * sina, dz * sina
    m = np.empty((3, 3))
    m[0, 0] = cosa + (dx * dx) * omc + 0.0
    m[0, 1] = 0.0 + (dx * dy) * omc + -sz
    m[0, 2] = 0.0 + (dx * dz) * omc + sy
    m[1, 0] = 0.0 + (dy * dx) * omc + sz
    m[1, 1] = cosa + (dy * dy) * omc + 0.0
    m[1, 2] = 0.0 + (dy * dz) * omc + -sx
    m[2, 0] = 0.0 + (dz * dx) * omc + -sy
    m[2, 1] = 0.0 + (dz * dy) * omc + sx
    m[2, 2] = cosa + (dz * dz) * omc + 0.0
    return m


def euler_matrix_rzxz(ai, aj, ak):
    """Rotating-frame z-x-z Euler matrix (reference transformations.py:1042-1102 with
    axes='rzxz'), used for the native->celestial rotation of the TAN projection."""
    ai, ak = ak, ai
    si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
    ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    i, j, k = 2, 0, 1
    M = np.identity(3)
    M[i, i] = cj
    M[i, j] = sj * si
    M[i, k] = sj * ci
    M[j, i] = sj * sk
    M[j, j] = -cj * ss + cc
    M[j, k] = -cj * cs - sc
    M[k, i] = -sj * ck
    M[k, j] = cj * sc + cs
    M[k, k] = cj * cc - ss
    return M


def julianDate(date):
    """UTC datetime -> Julian date as one float64: JD(0h) + day fraction (one rounding).
    Stands in for astropy `Time(date, scale='utc').jd` (reference transform.py:529)."""
    y, m = date.year, date.month
    a = (14 - m) // 12
    yy, mm = y + 4800 - a, m + 12 * a - 3
    jdn = date.day + (153 * mm + 2) // 5 + 365 * yy + yy // 4 - yy // 100 + yy // 400 - 32045
    sec = date.hour * 3600 + date.minute * 60 + date.second + date.microsecond / 1e6
    return (jdn - 0.5) + sec / 86400.0


def date2es(date):
    """UTC -> ephemeris seconds since J2000 (reference transform.py:525-530)."""
    return (julianDate(date) - 2451545) * 86400


def T0(et):
    return (et / 86400.0) / 36525.0


def H(et):
    jd = (et / 86400.0) - 0.5
    hh = (jd - int(jd)) * 24.0
    if hh < 0.0:
        hh += 24.0
    return hh


def lambda0(et):
    M = 357.528 + 35999.050 * T0(et)
    lambd = 280.460 + 36000.772 * T0(et)
    return lambd + (1.915 - 0.0048 * T0(et)) * math.sin(np.deg2rad(M)) + 0.020 * math.sin(np.deg2rad(2 * M))


def epsilon(et):
    return 23.439 - 0.013 * T0(et)


def _fracYear(et):
    idx = (et + 3155803200.0) / 157788000.0
    return idx, math.fmod(idx, 1.0)


def mag_lon(et):
    idx, fy = _fracYear(et)
    return math.atan2(calcH11(idx, fy), calcG11(idx, fy)) + math.pi


def mag_lat(et):
    idx, fy = _fracYear(et)
    g01, g11, h11 = calcG01(idx, fy), calcG11(idx, fy), calcH11(idx, fy)
    l0 = mag_lon(et)
    return math.pi / 2 - math.atan((g11 * math.cos(l0) + h11 * math.sin(l0)) / g01)


def mat_P(et):
    """J2000 -> GEI (precession)."""
    t0 = T0(et)
    mat = rotation_matrix(np.deg2rad(-1.0 * (0.64062 * t0 + 0.00030 * t0 * t0)), Z)
    mat = np.dot(mat, rotation_matrix(np.deg2rad(0.55675 * t0 - 0.00012 * t0 * t0), Y))
    return np.dot(mat, rotation_matrix(np.deg2rad(-1.0 * (0.64062 * t0 + 0.00008 * t0 * t0)), Z))


def mat_T1(et):
    """GEI -> GEO (Greenwich sidereal rotation)."""
    theta = 100.461 + 36000.770 * T0(et) + 360.0 * (H(et) / 24.0)
    return rotation_matrix(np.deg2rad(theta), Z)


def mat_T2(et):
    """GEI -> GSE."""
    return np.dot(rotation_matrix(np.deg2rad(lambda0(et)), Z), rotation_matrix(np.deg2rad(epsilon(et)), X))


def vec_Qe(et):
    lat, lon = mag_lat(et), mag_lon(et)
    Qg = [math.cos(lat) * math.cos(lon), math.cos(lat) * math.sin(lon), math.sin(lat)]
    return np.dot(np.dot(mat_T2(et), mat_T1(et).T), Qg)


def mat_T3(et):
    """GSE -> GSM."""
    Qe = vec_Qe(et)
    return rotation_matrix(-math.atan2(np.deg2rad(Qe[1]), np.deg2rad(Qe[2])), X)


def mat_T4(et):
    """GSM -> SM."""
    Qe = vec_Qe(et)
    mu = math.atan2(np.deg2rad(Qe[0]), np.deg2rad(math.sqrt(Qe[1] * Qe[1] + Qe[2] * Qe[2])))
    return rotation_matrix(-mu, Y)


def mat_j2000_to_geo(et):
    return np.dot(mat_T1(et), mat_P(et))


def mat_j2000_to_sm(et):
    return mat_T4(et).dot(mat_T3(et)).dot(mat_T2(et)).dot(mat_P(et))


def mat_geo_to_sm(et):
    return mat_T4(et).dot(mat_T3(et)).dot(mat_T2(et)).dot(mat_T1(et).T)


_D2R = math.pi / 180.0        # np.deg2rad(x) == x * (pi/180)


def frameMatrices(et):
    """(J2000->GEO, J2000->SM, GEO->SM) for one ephemeris second, every primitive rotation
    evaluated once.  Bit-identical to mat_j2000_to_geo / mat_j2000_to_sm / mat_geo_to_sm."""
    t0 = T0(et)
    P = rotation_matrix((-1.0 * (0.64062 * t0 + 0.00030 * t0 * t0)) * _D2R, Z)
    P = np.dot(P, rotation_matrix((0.55675 * t0 - 0.00012 * t0 * t0) * _D2R, Y))
    P = np.dot(P, rotation_matrix((-1.0 * (0.64062 * t0 + 0.00008 * t0 * t0)) * _D2R, Z))
    T1 = rotation_matrix((100.461 + 36000.770 * t0 + 360.0 * (H(et) / 24.0)) * _D2R, Z)
    M = 357.528 + 35999.050 * t0
    lam = 280.460 + 36000.772 * t0
    lam = lam + (1.915 - 0.0048 * t0) * math.sin(M * _D2R) + 0.020 * math.sin((2 * M) * _D2R)
    T2 = np.dot(rotation_matrix(lam * _D2R, Z), rotation_matrix((23.439 - 0.013 * t0) * _D2R, X))
    idx, fy = _fracYear(et)
    g01, g11, h11 = calcG01(idx, fy), calcG11(idx, fy), calcH11(idx, fy)
    lon = math.atan2(h11, g11) + math.pi
    lat = math.pi / 2 - math.atan((g11 * math.cos(lon) + h11 * math.sin(lon)) / g01)
    Qg = [math.cos(lat) * math.cos(lon), math.cos(lat) * math.sin(lon), math.sin(lat)]
    Qe = np.dot(np.dot(T2, T1.T), Qg)
    T3 = rotation_matrix(-math.atan2(np.deg2rad(Qe[1]), np.deg2rad(Qe[2])), X)
    mu = math.atan2(np.deg2rad(Qe[0]), np.deg2rad(math.sqrt(Qe[1] * Qe[1] + Qe[2] * Qe[2])))
    T4 = rotation_matrix(-mu, Y)
    T432 = T4.dot(T3).dot(T2)
    return np.dot(T1, P), T432.dot(P), T432.dot(T1.T)


# ---- scalar helpers used by the mapping objects (single points; not the per-pixel path) ----
def ecef2GeodeticScalar(x, y, z, a=wgs84A, b=wgs84B):
    """Single-point Bowring (reference transform.py:199-230) -> (lat, lon) radians."""
    e2 = (a * a - b * b) / (a * a)
    d = (a * a - b * b) / b
    p2 = x * x + y * y
    p = math.sqrt(p2)
    r = math.sqrt(p2 + z * z)
    tu = b * z * (1 + d / r) / (a * p)
    tu2 = tu * tu
    cu3 = (1 / math.sqrt(1 + tu2)) ** 3
    su3 = cu3 * tu2 * tu
    tp = (z + d * su3) / (p - e2 * a * cu3)
    return math.atan(tp), math.atan2(y, x)


def smLonToMLT(smlons):
    """SM longitude [deg, -180..180] -> magnetic local time [h] (reference :373-386)."""
    return smlons * (24 / 360) + 12


def mltToSmLon(mlt):
    """Magnetic local time [h] -> SM longitude [deg] (reference :388-401)."""
    return (mlt - 12) / (24 / 360)
