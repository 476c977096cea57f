"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the geodesic inverse problem on WGS84 after
C. F. F. Karney, "Algorithms for geodesics", J. Geodesy 87 (2013) 43-55 -- the algorithm behind
geographiclib==1.34 `Geodesic.WGS84.Inverse(...)['a12']`, which the reference calls in
coordinates/geodesic.py:35-44 (`angularDistance`) and which is absent from /root/reference and from
this image (third-party dependency, PARITY UNPINNED against the library itself).

Restated from the paper, not from the library: the auxiliary-sphere relations (eqs. 5-12), the
longitude integral I3 with the order-6 series A3 / C3l (eqs. 8, 23-25), the distance integral I1 with
A1 / C1l (eqs. 7, 15-18).  The library solves lambda12(alpha1) = lon12 by Newton's method from a
clever starting guess; lambda12 is monotonic in alpha1 on [0, pi] for the canonical configuration
(paper, section 5), so this oracle simply BISECTS -- slower, free of the starting-guess logic, and
convergent for nearly antipodal points too.

It pins the product's `auromat_b200.coordinates.geodesic.angularDistance` (Vincenty's iteration, an
independent derivation of the same quantity): tests/test_oracle_golden.py compares the two over random
point pairs and checks this file against the published GeodSolve / geographiclib documentation examples.
"""
from __future__ import annotations

import math

WGS84_A = 6378137.0
WGS84_F = 1 / 298.257223563


def _ang_normalize(x):
    x = math.fmod(x, 360.0)
    return x - 360.0 if x > 180.0 else (x + 360.0 if x <= -180.0 else x)


def _a3(n, eps):
    """A3(eps), paper eq. 24 (order 5 in eps, exact polynomial coefficients in n)."""
    c = [1.0,
         -(1 / 2 - n / 2),
         -(1 / 4 + n / 8 - 3 * n * n / 8),
         -(1 / 16 + 3 * n / 16 + n * n / 16),
         -(3 / 64 + n / 32),
         -3 / 128]
    return sum(ck * eps ** k for k, ck in enumerate(c))


def _c3(n, eps):
    """C3l(eps), l = 1..5, paper eq. 25."""
    n2 = n * n
    e = [eps ** k for k in range(6)]
    return [0.0,
            (1 / 4 - n / 4) * e[1] + (1 / 8 - n2 / 8) * e[2] + (3 / 64 + 3 * n / 64 - n2 / 64) * e[3]
            + (5 / 128 + n / 64) * e[4] + 3 / 128 * e[5],
            (1 / 16 - 3 * n / 32 + n2 / 32) * e[2] + (3 / 64 - n / 32 - 3 * n2 / 64) * e[3]
            + (3 / 128 + n / 128) * e[4] + 5 / 256 * e[5],
            (5 / 192 - 3 * n / 64 + 5 * n2 / 192) * e[3] + (3 / 128 - 5 * n / 192) * e[4] + 7 / 512 * e[5],
            (7 / 512 - 7 * n / 256) * e[4] + 7 / 512 * e[5],
            21 / 2560 * e[5]]


def _a1(eps):
    """A1(eps), paper eq. 17."""
    e2 = eps * eps
    return (1 + e2 / 4 + e2 * e2 / 64 + e2 * e2 * e2 / 256) / (1 - eps)


def _c1(eps):
    """C1l(eps), l = 1..6, paper eq. 18."""
    e = [eps ** k for k in range(7)]
    return [0.0,
            -e[1] / 2 + 3 * e[3] / 16 - e[5] / 32,
            -e[2] / 16 + e[4] / 32 - 9 * e[6] / 2048,
            -e[3] / 48 + 3 * e[5] / 256,
            -5 * e[4] / 512 + 3 * e[6] / 512,
            -7 * e[5] / 1280,
            -7 * e[6] / 2048]


def _sin_series(c, sigma):
    return sum(cl * math.sin(2 * l * sigma) for l, cl in enumerate(c) if l)


def _lambda12(sbet1, cbet1, sbet2, cbet2, alp1, f, n, ep2):
    """Longitude difference of the geodesic that leaves point 1 with azimuth alp1 and reaches the
    parallel of point 2 (first crossing beyond the vertex side given by the canonical form), plus
    everything the caller needs: (lam12, sig12, eps, sig1, sig2, salp2, calp2)."""
    salp1, calp1 = math.sin(alp1), math.cos(alp1)
    salp0 = salp1 * cbet1                                  # eq. 5 (Clairaut)
    calp0 = math.hypot(calp1, salp1 * sbet1)
    sig1 = math.atan2(sbet1, calp1 * cbet1)                # eq. 11
    omg1 = math.atan2(salp0 * sbet1, calp1 * cbet1)        # eq. 12
    salp2 = salp0 / cbet2 if cbet2 != cbet1 else salp1
    if cbet2 != cbet1 or abs(sbet2) != -sbet1:
        t = (cbet2 - cbet1) * (cbet1 + cbet2) if cbet1 < -sbet1 else (sbet1 - sbet2) * (sbet1 + sbet2)
        calp2 = math.sqrt(max(0.0, (calp1 * cbet1) ** 2 + t)) / cbet2
    else:
        calp2 = abs(calp1)
    sig2 = math.atan2(sbet2, calp2 * cbet2)
    omg2 = math.atan2(salp0 * sbet2, calp2 * cbet2)
    sig12 = sig2 - sig1
    omg12 = omg2 - omg1
    if sig12 < 0:                                          # cannot happen in canonical form except by rounding
        sig12 += 2 * math.pi
    if omg12 < 0:
        omg12 += 2 * math.pi
    k2 = ep2 * calp0 * calp0
    eps = k2 / (2 * (1 + math.sqrt(1 + k2)) + k2)          # eq. 16
    c3 = _c3(n, eps)
    i3 = _a3(n, eps) * (sig12 + _sin_series(c3, sig2) - _sin_series(c3, sig1))      # eq. 23
    lam12 = omg12 - f * salp0 * i3                          # eq. 8
    return lam12, sig12, eps, sig1, sig2, salp2, calp2


def inverse(lat1, lon1, lat2, lon2, a=WGS84_A, f=WGS84_F):
    """{'a12': arc length on the auxiliary sphere [deg], 's12': distance [m], 'azi1', 'azi2': forward
    azimuths at the two points [deg]} of the shortest geodesic."""
    n = f / (2 - f)
    e2 = f * (2 - f)
    ep2 = e2 / (1 - e2)
    b = a * (1 - f)
    lon12 = _ang_normalize(_ang_normalize(lon2) - _ang_normalize(lon1))
    lonsign = 1 if lon12 >= 0 else -1
    lon12 *= lonsign
    swapp = 1 if abs(lat1) >= abs(lat2) else -1
    if swapp < 0:
        lonsign *= -1
        lat1, lat2 = lat2, lat1
    latsign = 1 if lat1 < 0 else -1
    lat1 *= latsign
    lat2 *= latsign                                         # now lat1 <= 0, |lat1| >= |lat2|, lon12 in [0, 180]

    def reduced(lat):
        if abs(lat) == 90:
            return math.copysign(1.0, lat), 0.0
        t = (1 - f) * math.tan(math.radians(lat))
        c = 1 / math.hypot(1.0, t)
        return t * c, c
    sbet1, cbet1 = reduced(lat1)
    sbet2, cbet2 = reduced(lat2)
    cbet1, cbet2 = max(cbet1, 1e-300), max(cbet2, 1e-300)
    lam12 = math.radians(lon12)

    if lat1 == 0 and lat2 == 0 and lam12 <= (1 - f) * math.pi:
        # equatorial geodesic (paper, section 5): runs along the equator
        sig12 = lam12 / (1 - f)
        s12 = a * lam12
        salp1 = salp2 = 1.0
        calp1 = calp2 = 0.0
    elif lam12 == 0 or lon12 == 180 or cbet1 < 1e-290:
        # meridional geodesic: alpha1 = 0 (or pi through the pole)
        alp1 = 0.0 if lam12 == 0 else math.pi
        _, sig12, eps, sig1, sig2, salp2, calp2 = _lambda12(sbet1, cbet1, sbet2, cbet2, alp1, f, n, ep2)
        salp1, calp1 = math.sin(alp1), math.cos(alp1)
        if alp1 == math.pi:
            salp1 = 0.0                                     # over the pole; sin(pi) is not exactly 0
        c1 = _c1(eps)
        s12 = b * _a1(eps) * (sig12 + _sin_series(c1, sig2) - _sin_series(c1, sig1))
    else:
        lo, hi = 0.0, math.pi
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if mid == lo or mid == hi:
                break
            if _lambda12(sbet1, cbet1, sbet2, cbet2, mid, f, n, ep2)[0] < lam12:
                lo = mid
            else:
                hi = mid
        alp1 = 0.5 * (lo + hi)
        _, sig12, eps, sig1, sig2, salp2, calp2 = _lambda12(sbet1, cbet1, sbet2, cbet2, alp1, f, n, ep2)
        salp1, calp1 = math.sin(alp1), math.cos(alp1)
        c1 = _c1(eps)
        s12 = b * _a1(eps) * (sig12 + _sin_series(c1, sig2) - _sin_series(c1, sig1))     # eq. 7, 15

    if swapp < 0:
        salp1, salp2 = salp2, salp1
        calp1, calp2 = calp2, calp1
    salp1 *= swapp * lonsign
    calp1 *= swapp * latsign
    salp2 *= swapp * lonsign
    calp2 *= swapp * latsign
    return {'a12': math.degrees(sig12), 's12': s12,
            'azi1': math.degrees(math.atan2(salp1, calp1)), 'azi2': math.degrees(math.atan2(salp2, calp2))}


def angularDistance(lat1, lon1, lat2, lon2):
    return inverse(lat1, lon1, lat2, lon2)['a12']
