"""Ground-based all-sky imagers (fisheye lens looking at the zenith), georeferenced by the same
CUDA kernel as the spacecraft frames with the camera model switched to AMT_MODEL_ALLSKY.

API mirror of the intersection-based part of `auromat/mapping/miracle.py` (FMI MIRACLE
network): `CalibrationData` (:37-41), `MIRACLEMapping` (:117-366) with its az/el model
(`calculateAzEl` :314-347) and direction / intersection chain (:196-258).  Everything is in
the Earth-fixed frame: station position and directions are ECEF, there is no J2000 step, and
MLat/MLT take the generic geodetic route of `BaseMapping._mLatMlt` (mapping.py:540-550).
The reference's "simple" constant-grid mode and its JPEG/cal.txt file provider are host-side
conveniences outside the hot path (the cal.txt row parser is kept).
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import numpy.ma as ma

from .. import _lib
from ..coordinates import transform
from ..coordinates.geodesic import wgs84A, wgs84B
from .mapping import BaseMapping, CORNER_PLANES, MappingCollection

CalibrationData = namedtuple('CalibrationData', ['station', 'validFrom', 'validTo', 'lat', 'lon', 'xc', 'yc',
                                                 'k', 'rotation', 'boundingBoxSimple'])

REFERENCE_WIDTH = 512     # cal.txt numbers refer to 512x512 images (miracle.py:320-324)


def parseCalibrationRow(line):
    """One data row of a MIRACLE `cal.txt` -> CalibrationData (columns: Sta Glat Glon Active to
    Xc Yc k rotation lat+ lat- lon- lon+ i1 i2 i3)."""
    p = line.split()
    return CalibrationData(p[0], float(p[3]), float(p[4]), float(p[1]), float(p[2]), float(p[5]), float(p[6]),
                           float(p[7]), float(p[8]), None)


def stationEcef(lat, lon):
    """Station position on the WGS84 ellipsoid, height 0 (reference transform.py:180-197)."""
    la, lo = np.deg2rad(lat), np.deg2rad(lon)
    a, b = wgs84A, wgs84B
    e2 = (a * a - b * b) / (a * a)
    n = a / np.sqrt(1 - e2 * np.sin(la) ** 2)
    latn = n * np.cos(la)
    return np.array([latn * np.cos(lo), latn * np.sin(lo), n * (1 - e2) * np.sin(la)])


def stationMatrix(lat, lon):
    """local (east-north-up style) -> ECEF: latitude rotation first, then longitude
    (reference miracle.py:249-252)."""
    matLat = transform.rotation_matrix(np.deg2rad(90 - lat), transform.Y)
    matLon = transform.rotation_matrix(np.deg2rad(-lon), transform.Z)
    return np.dot(matLon, matLat)


class AllSkyMapping(BaseMapping):
    """Mapping of one all-sky image defined by station calibration data.

    :param CalibrationData calData: xc, yc, k for 512-px images, rotation in radians
    :param img: square (w,w[,n]) uint8/uint16 array or device tensor
    """

    _finiteElevation = True      # the camera elevation angle, finite wherever the centre is defined

    def __init__(self, calData, img, photoTime, alti=110, identifier=None, metadata=None, device=None,
                 sanitize=True):
        if identifier is None:
            identifier = calData.station + '.' + photoTime.strftime('%Y.%m.%d.%H.%M.%S')
        self.cameraPosGEO = stationEcef(calData.lat, calData.lon)
        et = transform.date2es(photoTime)
        cameraPosGCRS = transform.mat_j2000_to_geo(et).T.dot(self.cameraPosGEO)   # latLonToJ2000(lat, lon, 0, t)
        BaseMapping.__init__(self, 110 if alti is None else alti, cameraPosGCRS, photoTime, identifier, metadata,
                             device)
        self._calData = calData
        if hasattr(img, 'data_ptr'):
            self._imgDevice = img if img.dim() == 3 else img[..., None]
            self._imgData = None
            shape = tuple(img.shape)
        else:
            if ma.isMaskedArray(img):
                img = img.data
            assert img.dtype in (np.uint8, np.uint16)
            self._imgData = img
            shape = img.shape
        assert shape[0] == shape[1], 'all-sky images are square'
        self._shape = (shape[0], shape[1])
        self._sanitize = sanitize
        self._frame = None

    shape = property(lambda self: self._shape)
    calibration = property(lambda self: self._calData)

    @property
    def img_unmasked(self):
        if self._imgData is None:
            self._imgData = self.context.to_numpy(self._imgDevice)
        return self._imgData

    @property
    def frameConstants(self):
        if self._frame is None:
            c = self._calData
            w = self._shape[0]
            s = w / REFERENCE_WIDTH
            fr = _lib.AmtFrame()
            fr.width = fr.height = w
            fr.model = _lib.AMT_MODEL_ALLSKY
            fr.allsky_xc, fr.allsky_yc, fr.allsky_k = c.xc * s, c.yc * s, c.k * s
            fr.allsky_rotation = c.rotation
            fr.rot[:] = stationMatrix(c.lat, c.lon).ravel().tolist()
            fr.cam[:] = self.cameraPosGEO.tolist()
            a, b = wgs84A + self.altitude, wgs84B + self.altitude
            fr.inv_axes[:] = [1 / a, 1 / a, 1 / b]
            x, y, z = self.cameraPosGEO
            fr.origin_inside = 1 if (x / a) ** 2 + (y / a) ** 2 + (z / b) ** 2 < 1 else 0
            fr.m_geo[:] = np.identity(3).ravel().tolist()
            fr.m_sm[:] = np.identity(3).ravel().tolist()
            fr.wgs_a, fr.wgs_b = wgs84A, wgs84B
            self._frame = fr
        return self._frame

    def _computePlanes(self, ctx, names):
        import torch
        h, w = self._shape
        if 'lat_k' not in self._planes:
            nk, nc = (h + 1) * (w + 1), h * w
            fresh = {n: ctx.empty(nk if n in CORNER_PLANES else nc, torch.float64)
                     for n in ('lat_k', 'lon_k', 'lat_c', 'lon_c', 'elev_c')}
            fresh['valid_k'], fresh['valid_c'] = ctx.new_bitmaps(w, h)
            ctx.georef(self.frameConstants, fresh)
            self._planes.update(fresh)
            if self._sanitize:
                ctx.sanitize(w, h, self._planes)
        if any(n.startswith('ml') for n in names) and 'mlat_k' not in self._planes:
            m = transform.mat_geo_to_sm(transform.date2es(self.photoTime))
            for suffix in ('k', 'c'):
                mlat, mlt = ctx.latlon_to_mlatmlt(self._planes['lat_' + suffix], self._planes['lon_' + suffix],
                                                  self.altitude, wgs84A, wgs84B, m)
                self._planes['mlat_' + suffix], self._planes['mlt_' + suffix] = mlat, mlt


def getMapping(img, calData, photoTime, altitude=110, device=None):
    """All-sky counterpart of `auromat.mapping.miracle.getMapping` for in-memory images."""
    return AllSkyMapping(calData, img, photoTime, altitude, device=device)


def getMappingCollection(images, calDatas, photoTime, altitude=110, minElevation=None, device=None,
                         identifier=None):
    """Mappings of several stations at one time as a `MappingCollection`
    (reference miracle.py:88-103), optionally masked by camera elevation."""
    mappings = []
    for img, cal in zip(images, calDatas):
        m = AllSkyMapping(cal, img, photoTime, altitude, device=device)
        if minElevation is not None:
            m = m.maskedByElevation(minElevation)
        mappings.append(m)
    if identifier is None:
        identifier = 'ALLSKY.' + photoTime.strftime('%Y.%m.%d.%H.%M.%S')
    return MappingCollection(mappings, identifier=identifier, mayOverlap=True)
