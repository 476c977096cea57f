"""Digest a FITS-WCS header (TAN, optionally TAN-SIP) plus camera position and time into the
per-frame constant block `amt_frame` consumed by the CUDA georeference kernel.

Mirrors the header handling of `auromat/coordinates/wcs.py` (:50-52 dispatch, :80-90 header
keys, :135-139 native->celestial Euler rotation).  The reference evaluates `-SIP` headers
through astropy/wcslib (:54-56); here the FITS-SIP forward polynomial is evaluated in the
same kernel as the TAN deprojection.
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from . import transform
from .geodesic import wgs84A, wgs84B


def isTanHeader(header):
    c1, c2 = header['CTYPE1'], header['CTYPE2']
    return (c1, c2) in (('RA---TAN', 'DEC--TAN'), ('RA---TAN-SIP', 'DEC--TAN-SIP'))


def nativeRotation(header):
    """reference wcs.py:135-139: euler 'rzxz' with (CRVAL1+90, 90-CRVAL2, -(LONPOLE-90))."""
    return transform.euler_matrix_rzxz(np.deg2rad(header['CRVAL1'] + 90),
                                       np.deg2rad(90 - header['CRVAL2']),
                                       np.deg2rad(-(header['LONPOLE'] - 90)))


def _sipPacked(header, prefix):
    order = int(header[prefix + '_ORDER'])
    if not (0 <= order <= _lib.AMT_SIP_MAX_ORDER):
        raise NotImplementedError('SIP order %d not supported (max %d)' % (order, _lib.AMT_SIP_MAX_ORDER))
    packed = np.zeros(_lib.AMT_SIP_MAX_COEF)
    for p in range(order + 1):
        base = p * (order + 1) - (p * (p - 1)) // 2
        for q in range(order + 1 - p):
            packed[base + q] = float(header.get('%s_%d_%d' % (prefix, p, q), 0.0))
    return order, packed


R_EARTH_SPHERE = 6378.136      # astropy.constants.R_earth [km], the 'sphere' model of mapping.py:1502-1505


def frameConstants(header, cameraPosGCRS, photoTime, altitude, fastCenterCalculation=False, earthModel='wgs84'):
    """Build the `amt_frame` for one image.

    :param header: dict-like with CTYPE1/2, LATPOLE, LONPOLE, CRVAL1/2, CRPIX1/2, CD*, IMAGEW/H
    :param cameraPosGCRS: (3,) km
    :param datetime photoTime: UTC
    :param altitude: emission altitude in km (inflation of the WGS84 ellipsoid)
    :param earthModel: 'wgs84' (ellipsoid a, b + altitude) or 'sphere' (radius R_earth + altitude): the two
        models of the reference's `inflatedEarthIntersection` (mapping/mapping.py:1474-1510); the kernels
        are generic in the three semi-axes
    """
    if not isTanHeader(header) or header['LATPOLE'] != 0.0:
        # the reference falls back to astropy.wcs for anything else (wcs.py:53-62)
        raise NotImplementedError('only TAN / TAN-SIP headers with LATPOLE == 0 are supported, got %r/%r'
                                  % (header['CTYPE1'], header['CTYPE2']))
    fr = _lib.AmtFrame()
    fr.width, fr.height = int(header['IMAGEW']), int(header['IMAGEH'])
    fr.fast_center = 1 if fastCenterCalculation else 0
    fr.crpix[:] = [float(header['CRPIX1']), float(header['CRPIX2'])]
    fr.cd[:] = [float(header['CD1_1']), float(header['CD1_2']), float(header['CD2_1']), float(header['CD2_2'])]
    fr.rot[:] = nativeRotation(header).ravel().tolist()
    cam = np.asarray(cameraPosGCRS, dtype=np.float64)
    assert cam.shape == (3,)
    fr.cam[:] = cam.tolist()
    # reference mapping/mapping.py:1497-1500 and intersection.py:66
    if earthModel == 'wgs84':
        a, b = wgs84A + altitude, wgs84B + altitude
    elif earthModel == 'sphere':
        a = b = R_EARTH_SPHERE + altitude
    else:
        raise ValueError('unsupported earth model: ' + str(earthModel))
    fr.inv_axes[:] = [1 / a, 1 / a, 1 / b]
    x, y, z = cam
    fr.origin_inside = 1 if (x / a) ** 2 + (y / a) ** 2 + (z / b) ** 2 < 1 else 0   # intersection.py:239-241
    et = transform.date2es(photoTime)
    m_geo, m_sm, _ = transform.frameMatrices(et)
    fr.m_geo[:] = m_geo.ravel().tolist()
    fr.m_sm[:] = m_sm.ravel().tolist()
    fr.wgs_a, fr.wgs_b = wgs84A, wgs84B
    if str(header['CTYPE1']).endswith('-SIP'):
        fr.sip_order_a, pa = _sipPacked(header, 'A')
        fr.sip_order_b, pb = _sipPacked(header, 'B')
        fr.sip_a[:] = pa.tolist()
        fr.sip_b[:] = pb.tolist()
    return fr
