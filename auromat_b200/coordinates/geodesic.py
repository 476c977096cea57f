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


def angularDistance(location1, location2):
    """Shortest angular distance in degrees on the auxiliary sphere between two locations
    (reference: coordinates/geodesic.py:35-44)."""
    lat1, lon1, lat2, lon2 = location1.lat, location1.lon, location2.lat, location2.lon
    if lat1 == lat2 and lon1 == lon2:
        return 0.0
    f = WGS84_F
    U1 = math.atan((1 - f) * math.tan(math.radians(lat1)))
    U2 = math.atan((1 - f) * math.tan(math.radians(lat2)))
    L = math.radians(lon2 - lon1)
    L = (L + math.pi) % (2 * math.pi) - math.pi
    sU1, cU1, sU2, cU2 = math.sin(U1), math.cos(U1), math.sin(U2), math.cos(U2)
    lam = L
    sigma = 0.0
    for _ in range(200):
        sl, cl = math.sin(lam), math.cos(lam)
        ss = math.hypot(cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl)
        if ss == 0:
            return 0.0
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
    return math.degrees(sigma)
