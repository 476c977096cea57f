"""WGS84 constants and the two geodesic scalars the regridding path needs.

Mirrors `auromat/coordinates/geodesic.py` of the reference (names `wgs84A`, `wgs84B`,
`Location`, `angularDistance`).  The reference delegates to geographiclib, which is a
third-party dependency; here the auxiliary-sphere arc length is obtained with Vincenty's
inverse iteration (same quantity as geographiclib's `a12`).
"""
from __future__ import annotations

import math
from collections import namedtuple

# reference: coordinates/geodesic.py:20-21 (geographiclib Constants.WGS84_a [m], WGS84_f)
WGS84_A_M = 6378137.0
WGS84_F = 1 / 298.257223563
wgs84A = WGS84_A_M / 1000
wgs84B = wgs84A * (1 - WGS84_F)

Location = namedtuple('Location', ['lat', 'lon'])  # degrees


def _inverse(lat1, lon1, lat2, lon2):
    """Vincenty's inverse iteration on WGS84: (sigma [rad], s12 [m], azi1 [deg]) with sigma the
    arc length on the auxiliary sphere (geographiclib's `a12`), s12 the geodesic distance and
    azi1 the forward azimuth at point 1 in (-180, 180]."""
    if lat1 == lat2 and lon1 == lon2:
        return 0.0, 0.0, 0.0
    f = WGS84_F
    U1 = math.atan((1 - f) * math.tan(math.radians(lat1)))
    U2 = math.atan((1 - f) * math.tan(math.radians(lat2)))
    L = math.radians(lon2 - lon1)
    L = (L + math.pi) % (2 * math.pi) - math.pi
    sU1, cU1, sU2, cU2 = math.sin(U1), math.cos(U1), math.sin(U2), math.cos(U2)
    lam = L
    sigma = ss = cs = c2a = c2sm = 0.0
    for _ in range(200):
        sl, cl = math.sin(lam), math.cos(lam)
        # plain IEEE operations only (+ - * / sqrt) besides libm's sin/cos/tan/atan/atan2: the C port in
        # csrc/amt_host.cuh (amt_plate_carree_resolution) reproduces this function bit for bit
        t1, t2 = cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl
        ss = math.sqrt(t1 * t1 + t2 * t2)
        if ss == 0:
            return 0.0, 0.0, 0.0
        cs = sU1 * sU2 + cU1 * cU2 * cl
        sigma = math.atan2(ss, cs)
        sa = cU1 * cU2 * sl / ss
        c2a = 1 - sa * sa
        c2sm = cs - 2 * sU1 * sU2 / c2a if c2a != 0 else 0.0
        Cc = f / 16 * c2a * (4 + f * (4 - 3 * c2a))
        lam_new = L + (1 - Cc) * f * sa * (sigma + Cc * ss * (c2sm + Cc * cs * (-1 + 2 * c2sm * c2sm)))
        done = abs(lam_new - lam) < 1e-15
        lam = lam_new
        if done:
            break
    b = WGS84_A_M * (1 - f)
    u2 = c2a * (WGS84_A_M * WGS84_A_M - b * b) / (b * b)
    A = 1 + u2 / 16384 * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
    B = u2 / 1024 * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
    dsig = B * ss * (c2sm + B / 4 * (cs * (-1 + 2 * c2sm * c2sm)
                                     - B / 6 * c2sm * (-3 + 4 * ss * ss) * (-3 + 4 * c2sm * c2sm)))
    s12 = b * A * (sigma - dsig)
    azi1 = math.degrees(math.atan2(cU2 * math.sin(lam), cU1 * sU2 - sU1 * cU2 * math.cos(lam)))
    return sigma, s12, azi1


def angularDistance(location1, location2):
    """Shortest angular distance in degrees on the auxiliary sphere between two locations
    (reference: coordinates/geodesic.py:35-44)."""
    return math.degrees(_inverse(location1.lat, location1.lon, location2.lat, location2.lon)[0])


def distance(location1, location2):
    """Shortest distance in metres between two locations (reference geodesic.py:25-33)."""
    return _inverse(location1.lat, location1.lon, location2.lat, location2.lon)[1]


def course(location1, location2):
    """Azimuth in degrees when leaving `location1` towards `location2` (reference
    geodesic.py:114-122)."""
    return _inverse(location1.lat, location1.lon, location2.lat, location2.lon)[2]


def _courseDelta(a1, a2):
    """Left-turn amount from course a1 to course a2, in (-180, 180) (reference :124-140)."""
    if a2 < a1:
        a2 += 360
    turn = a2 - a1
    if turn == 180:
        return 0
    return turn - 360 if turn > 180 else turn


def _courseDeltaSum(points):
    """Sum of the course changes along a closed, non-intersecting polygon of (lat, lon) points
    (first point not repeated): one of -360, -180, 0, 180, 360 (reference :142-185)."""
    import numpy as np
    pts = np.asarray(points, dtype=float)
    assert pts.ndim == 2 and pts.shape[1] == 2
    n = len(pts)
    courses = []
    for i in range(n):
        p1 = Location(*pts[i])
        p2 = Location(*pts[(i + 1) % n])
        courses.append(course(p1, p2))
        courses.append(course(p2, p1) + 180)
    total = _courseDelta(courses[-1], courses[0])
    for i in range(1, 2 * n):
        total += _courseDelta(courses[i - 1], courses[i])
    total = float(np.around(total, decimals=1))
    assert total in (-360, -180, 0, 180, 360), total
    return total


def containsOrCrossesPole(points):
    """Whether the polygon of ordered (lat, lon) points contains (sum 0) or crosses (|sum| 180)
    one of the poles (reference geodesic.py:187-202)."""
    return abs(_courseDeltaSum(points)) != 360


def destination(location, azimuth, distance):
    """Location reached from `location` after `distance` metres along the geodesic leaving at
    `azimuth` degrees (reference geodesic.py:83-94; Vincenty's direct formula here)."""
    f, a = WGS84_F, WGS84_A_M
    b = a * (1 - f)
    alpha1 = math.radians(azimuth)
    sa1, ca1 = math.sin(alpha1), math.cos(alpha1)
    tanU1 = (1 - f) * math.tan(math.radians(location.lat))
    cU1 = 1 / math.sqrt(1 + tanU1 * tanU1)
    sU1 = tanU1 * cU1
    sigma1 = math.atan2(tanU1, ca1)
    sa = cU1 * sa1
    c2a = 1 - sa * sa
    u2 = c2a * (a * a - b * b) / (b * b)
    A = 1 + u2 / 16384 * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
    B = u2 / 1024 * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
    sigma = distance / (b * A)
    for _ in range(200):
        c2sm = math.cos(2 * sigma1 + sigma)
        ss, cs = math.sin(sigma), math.cos(sigma)
        dsig = B * ss * (c2sm + B / 4 * (cs * (-1 + 2 * c2sm ** 2)
                                         - B / 6 * c2sm * (-3 + 4 * ss ** 2) * (-3 + 4 * c2sm ** 2)))
        new = distance / (b * A) + dsig
        done = abs(new - sigma) < 1e-15
        sigma = new
        if done:
            break
    c2sm = math.cos(2 * sigma1 + sigma)
    ss, cs = math.sin(sigma), math.cos(sigma)
    t = sU1 * ss - cU1 * cs * ca1
    lat2 = math.atan2(sU1 * cs + cU1 * ss * ca1, (1 - f) * math.sqrt(sa * sa + t * t))
    lam = math.atan2(ss * sa1, cU1 * cs - sU1 * ss * ca1)
    Cc = f / 16 * c2a * (4 + f * (4 - 3 * c2a))
    L = lam - (1 - Cc) * f * sa * (sigma + Cc * ss * (c2sm + Cc * cs * (-1 + 2 * c2sm ** 2)))
    lon2 = (math.radians(location.lon) + L + 3 * math.pi) % (2 * math.pi) - math.pi
    return Location(math.degrees(lat2), math.degrees(lon2))


def intermediate(location1, location2, f=0.5):
    """Location after travelling the fraction `f` of the geodesic from `location1` to
    `location2` (reference geodesic.py:96-112)."""
    _, d, az = _inverse(location1.lat, location1.lon, location2.lat, location2.lon)
    return destination(location1, az, d * f)
