"""THEMIS all-sky imager mappings on the device planes.

API mirror of the array side of the reference's `auromat/mapping/themis.py`:
`reproject` (:224-253), `ThemisMapping` (:116-205), `bytscl` (:207-222) and the array part of
`mappingSingleASI` (:403-447, here `mappingFromCalibration`) / `getMappings` (:449-470, here
`mappingCollection`).  Reading the L1/L2 CDF files (spacepy.pycdf, downloads) is outside the hot
path (out of scope): callers hand in the arrays those files contain.

The calibration provides corner coordinates for a few reference emission heights; any other
mapping altitude goes through `amt_reproject` (one thread per corner: geodetic -> ECEF ->
direction from the station -> inflated-ellipsoid intersection -> Bowring), centres are the
four-corner means (`amt_corner_means`), and the result is an ordinary sanitised mapping that
`resample()` / `parallel.mosaic()` consume like any other.
"""
from __future__ import annotations

import numpy as np
import numpy.ma as ma

from ..coordinates import transform
from ..coordinates.geodesic import wgs84A, wgs84B
from ..runtime import get_context
from .allsky import stationEcef
from .mapping import GenericMapping, MappingCollection

# reference themis.py:24-27
stations = ['atha', 'chbg', 'ekat', 'fsim', 'fsmi', 'fykn',
            'gako', 'gbay', 'gill', 'inuv', 'kapu', 'kian',
            'kuuj', 'mcgr', 'nrsq', 'pgeo', 'pina', 'rank',
            'snap', 'snkq', 'talo', 'tpas', 'whit', 'yknf']


def reproject(latLonASI, latsRef, lonsRef, heightRef, heightNew, device=None):
    """Reproject corner coordinates given for the emission height `heightRef` [km] to
    `heightNew` [km] as seen from the imager at `latLonASI` (reference themis.py:224-253).
    numpy arrays in -> numpy arrays out; device tensors in -> device tensors out."""
    ctx = get_context(device)
    latASI, lonASI = latLonASI
    ecef = stationEcef(float(latASI), float(lonASI))
    onDevice = hasattr(latsRef, 'data_ptr')
    if onDevice:
        dlat, dlon = latsRef.contiguous(), lonsRef.contiguous()
    else:
        dlat = ctx.to_device(np.ascontiguousarray(latsRef, dtype=np.float64))
        dlon = ctx.to_device(np.ascontiguousarray(lonsRef, dtype=np.float64))
    olat, olon = ctx.reproject(dlat, dlon, ecef, heightRef, heightNew, wgs84A, wgs84B)
    if onDevice:
        return olat, olon
    return ctx.to_numpy(olat), ctx.to_numpy(olon)


def bytscl(array, max_=None, min_=None, top=255):
    """IDL BYTSCL, float formula (reference themis.py:207-222)."""
    if max_ is None:
        max_ = np.nanmax(array)
    if min_ is None:
        min_ = np.nanmin(array)
    return np.maximum(np.minimum(((top + 0.9999) * (array - min_) / (max_ - min_)).astype(np.int16), top), 0)


class ThemisMapping(GenericMapping):
    """A single-station THEMIS mapping: grey-scale image (h,w), calibrated corner coordinates
    (reference themis.py:116-205)."""

    def __init__(self, lats, lons, latsCenter, lonsCenter, elev, alti, img, cameraPosGCRS, photoTime,
                 station, minBrightness=None, maxBrightness=None, device=None):
        assert img.ndim == 2
        identifier = station + '.' + photoTime.strftime('%Y.%m.%d.%H.%M.%S')
        GenericMapping.__init__(self, lats, lons, latsCenter, lonsCenter, elev, alti, img[:, :, None],
                                cameraPosGCRS, photoTime, identifier, device=device)
        self.station = station
        self.minBrightness = minBrightness
        self.maxBrightness = maxBrightness

    def brightness_scaled(self, img):
        # brightness scaling of thm_asi_create_mosaic.pro (reference themis.py:190-198)
        if self.minBrightness is not None or self.maxBrightness is not None:
            return bytscl(img, min_=self.minBrightness, max_=self.maxBrightness, top=255)
        valid = ma.getdata(self.img)[~ma.getmaskarray(self.img)]
        med = np.median(valid[valid > 1])
        return np.minimum(img / med * 64, 255)

    @property
    def rgb(self):
        return np.require(np.repeat(self.brightness_scaled(self.img), 3, 2), dtype=np.uint8)

    @property
    def rgb_unmasked(self):
        return np.require(np.repeat(self.brightness_scaled(self.img_unmasked), 3, 2), dtype=np.uint8)

    def createResampled(self, lats, lons, latsCenter, lonsCenter, elevation, img):
        if img.ndim == 3:
            img = img[:, :, 0]
        return ThemisMapping(lats, lons, latsCenter, lonsCenter, elevation, self.altitude, img, self.cameraPosGCRS,
                             self.photoTime, self.station, self.minBrightness, self.maxBrightness, device=self._device)


def mappingFromCalibration(station, latLonASI, el, latsRef, lonsRef, heightsRef, img, imgDate, altitude=110,
                           minBrightness=None, maxBrightness=None, device=None):
    """The array part of `mappingSingleASI` (reference themis.py:403-447).

    latLonASI: (lat, lon) of the station; el: (w,w) elevation of the pixel centres [deg];
    latsRef, lonsRef: (n,w+1,w+1) corner coordinates for the reference heights `heightsRef`
    [km]; img: (w,w) raw uint16 counts of the L1 file; imgDate: datetime of the exposure."""
    ctx = get_context(device)
    heightsRef = np.asarray(heightsRef, dtype=np.float64)
    img = np.array(img, copy=True)
    assert img.ndim == 2 and img.dtype in (np.uint8, np.uint16)
    h, w = img.shape
    hit = np.flatnonzero(heightsRef == altitude)
    if len(hit):
        lats = ctx.to_device(np.ascontiguousarray(latsRef[hit[0]], dtype=np.float64))
        lons = ctx.to_device(np.ascontiguousarray(lonsRef[hit[0]], dtype=np.float64))
    else:
        lats, lons = reproject(latLonASI, ctx.to_device(np.ascontiguousarray(latsRef[0], dtype=np.float64)),
                               ctx.to_device(np.ascontiguousarray(lonsRef[0], dtype=np.float64)),
                               heightsRef[0], altitude, device=device)
    # THEMIS mappings do not span the date line: plain four-corner means (:425-426)
    latsCenter, lonsCenter = ctx.corner_means(w, h, lats, lons)
    # 2500 is the `_offset` of every THEMIS L2 file (:433-437); unsigned arithmetic wraps as in numpy
    img -= np.asarray(2500).astype(img.dtype)
    latASI, lonASI = latLonASI
    mgeo = transform.mat_j2000_to_geo(transform.date2es(imgDate))
    cameraPosGCRS = mgeo.T.dot(stationEcef(float(latASI), float(lonASI)))        # latLonToJ2000(lat, lon, 0, t)
    m = ThemisMapping(lats.reshape(h + 1, w + 1), lons.reshape(h + 1, w + 1), latsCenter.reshape(h, w),
                      lonsCenter.reshape(h, w), el, altitude, img, cameraPosGCRS, imgDate, station,
                      minBrightness, maxBrightness, device=device)
    # the calibration is unreliable at very low elevation angles (:443-446)
    return m.maskedByElevation(1)


def mappingCollection(mappings, photoTime):
    """`getMappings` without the file access (reference themis.py:449-470)."""
    return MappingCollection([m for m in mappings if m is not None],
                             'THEMIS.' + photoTime.strftime('%Y.%m.%d.%H.%M.%S'), mayOverlap=True)
