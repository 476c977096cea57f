"""Synthetic inputs of the BASELINE.json configurations.

Config 1/2: one ISS Nikon D3S frame, 4256x2832, pure-TAN WCS modelled on the reference's
fixture `auromat/test/resources/ISS030-E-102170_dc.wcs` (header values quoted in SURVEY.md),
camera position POS*SHIF, time DATE-OBS - 13 s, uint8 RGB from `default_rng(0)`.
Config 3: 6000x4000 frame, same geometry scaled, SIP order 4.
Config 4: 512-frame sequence derived from config 1.
Used by bench.py, __graft_entry__.smoke() and the tests; no reference files are read.
"""
from __future__ import annotations

import math
from datetime import datetime, timedelta

import numpy as np

D3S_W, D3S_H = 4256, 2832

_BASE = {
    'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'LATPOLE': 0.0, 'LONPOLE': 180.0,
    'CRVAL1': 16.0531567459, 'CRVAL2': 23.1148929108, 'CRPIX1': 2129, 'CRPIX2': 1417,
    'CD1_1': -0.00912247310646, 'CD1_2': -0.00250608809647,
    'CD2_1': 0.00250608809647, 'CD2_2': -0.00912247310646,
    'IMAGEW': D3S_W, 'IMAGEH': D3S_H,
    'DATE-OBS': '2012-01-25T09:27:08.060000',
    'POSXSHIF': -4809.524217485676, 'POSYSHIF': 524.8117887762777, 'POSZSHIF': 4729.265809729493,
    'DATESHIF': -13.0,
}


def issHeader(width=D3S_W, height=D3S_H, sipOrder=0, seed=1):
    """Header dict of the synthetic ISS frame, geometrically scaled to width x height (the
    field of view is kept, the pixel scale changes)."""
    h = dict(_BASE)
    sx, sy = D3S_W / width, D3S_H / height
    h['IMAGEW'], h['IMAGEH'] = int(width), int(height)
    h['CRPIX1'] = _BASE['CRPIX1'] / sx
    h['CRPIX2'] = _BASE['CRPIX2'] / sy
    h['CD1_1'], h['CD1_2'] = _BASE['CD1_1'] * sx, _BASE['CD1_2'] * sy
    h['CD2_1'], h['CD2_2'] = _BASE['CD2_1'] * sx, _BASE['CD2_2'] * sy
    if sipOrder:
        rng = np.random.default_rng(seed)
        h['CTYPE1'], h['CTYPE2'] = 'RA---TAN-SIP', 'DEC--TAN-SIP'
        h['A_ORDER'] = h['B_ORDER'] = int(sipOrder)
        half = max(width, height) / 2.0
        for name in 'AB':
            for p in range(sipOrder + 1):
                for q in range(sipOrder + 1 - p):
                    if p + q < 2:
                        continue
                    # |distortion| ~ 20 px at the frame corners
                    h['%s_%d_%d' % (name, p, q)] = float(rng.normal()) * 20.0 / half ** (p + q) / sipOrder
    return h


def issImage(width=D3S_W, height=D3S_H, seed=0, dtype=np.uint8):
    rng = np.random.default_rng(seed)
    hi = 256 if dtype == np.uint8 else 65536
    return rng.integers(0, hi, (height, width, 3), dtype=dtype)


def sequenceHeaders(n, width=D3S_W, height=D3S_H):
    """Config 4: CRVAL1 advanced 0.05 deg/frame, time +1 s/frame, camera advanced along a
    circular orbit through the config-1 position (7.66 km/s)."""
    base = issHeader(width, height)
    p0 = np.array([base['POSXSHIF'], base['POSYSHIF'], base['POSZSHIF']])
    r = np.linalg.norm(p0)
    # orbit plane: spanned by p0 and a fixed prograde direction
    t = np.cross([0.0, 0.0, 1.0], p0)
    t /= np.linalg.norm(t)
    omega = 7.66 / r
    t0 = datetime.strptime(base['DATE-OBS'], '%Y-%m-%dT%H:%M:%S.%f')
    out = []
    for i in range(n):
        h = dict(base)
        h['CRVAL1'] = base['CRVAL1'] + 0.05 * i
        a = omega * i
        p = p0 * math.cos(a) + t * r * math.sin(a)
        h['POSXSHIF'], h['POSYSHIF'], h['POSZSHIF'] = (float(v) for v in p)
        h['DATE-OBS'] = (t0 + timedelta(seconds=i)).strftime('%Y-%m-%dT%H:%M:%S.%f')
        out.append(h)
    return out


def headerTimeAndCamera(header):
    """(photoTime, cameraPosGCRS) the way `getMapping` derives them from the header."""
    t = datetime.strptime(header['DATE-OBS'], '%Y-%m-%dT%H:%M:%S.%f') + timedelta(seconds=header['DATESHIF'])
    return t, np.array([header['POSXSHIF'], header['POSYSHIF'], header['POSZSHIF']])


def issHeaderLookingAt(camLat, camLon, targetLat, targetLon, width=D3S_W, height=D3S_H, camAltitude=400.0,
                       targetAltitude=110.0):
    """Synthetic ISS header whose camera sits above (camLat, camLon) at `camAltitude` km and whose
    boresight (CRVAL) points at the ground point (targetLat, targetLon) at `targetAltitude` km:
    used to place the pole or the date line inside the footprint."""
    from .coordinates import transform
    h = issHeader(width, height)
    t, _ = headerTimeAndCamera(h)
    mgeo = transform.mat_j2000_to_geo(transform.date2es(t))

    def ecef(lat, lon, r):
        la, lo = math.radians(lat), math.radians(lon)
        return np.array([r * math.cos(la) * math.cos(lo), r * math.cos(la) * math.sin(lo), r * math.sin(la)])

    cam = mgeo.T.dot(ecef(camLat, camLon, 6371.0 + camAltitude))
    tgt = mgeo.T.dot(ecef(targetLat, targetLon, 6371.0 + targetAltitude))
    d = tgt - cam
    d /= np.linalg.norm(d)
    h['CRVAL1'] = math.degrees(math.atan2(d[1], d[0])) % 360.0
    h['CRVAL2'] = math.degrees(math.asin(d[2]))
    h['POSXSHIF'], h['POSYSHIF'], h['POSZSHIF'] = (float(v) for v in cam)
    return h
