"""Export hand-off: the variables the reference's CDF / netCDF writers store, in their order,
shapes and types (reference export/cdf.py:60-290), as plain numpy arrays.

Writing the files themselves (spacepy.pycdf, netCDF4) is outside the hot path; this module is the
boundary between the device-resident mapping and such a writer: every array is pulled from the
device planes once, with a leading record (Epoch) axis where the reference adds one.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import numpy.ma as ma


def _var(data, fieldnam, units=None, depend=None, **attrs):
    a = OrderedDict(FIELDNAM=fieldnam)
    if units is not None:
        a['UNITS'] = units
    for k, d in enumerate(depend or ()):
        a['DEPEND_%d' % k] = d
    a.update(attrs)
    return dict(data=data, attrs=a)


def cdfVariables(mapping, includeBounds=True, includeMagCoords=True, includeGeoCoords=True):
    """OrderedDict name -> {'data': ndarray, 'attrs': {...}} with the variables of
    `auromat.export.cdf.write` (cdf.py:83-290), same names, order, shapes and dtypes:
    coordinates as (1, h, w) / (1, h+1, w+1) float64 (masked values filled with NaN), image bands
    widened to the next signed type with the type minimum as fill value when the image is masked
    (cdf.py:218-232), `zenith_angle` = 90 - elevation as float32."""
    pixel = ('Epoch', 'y_pixel', 'x_pixel')
    corner = ('Epoch', 'y_corner', 'x_corner')
    fill = lambda a: ma.filled(a.astype(np.float64), np.nan)[np.newaxis, :]
    out = OrderedDict()
    out['Epoch'] = _var([mapping.photoTime], 'Epoch', VAR_TYPE='support_data')
    if includeGeoCoords:
        out['lat'] = _var(fill(mapping.latsCenter), 'Latitude of pixel center', 'degrees', pixel,
                          VALIDMIN=-90.0, VALIDMAX=90.0, VAR_NOTES='Geodetic latitude')
        out['lon'] = _var(fill(mapping.lonsCenter), 'Longitude of pixel center', 'degrees', pixel,
                          VALIDMIN=-180.0, VALIDMAX=180.0, VAR_NOTES='Geodetic longitude')
        if includeBounds:
            out['lat_bounds'] = _var(fill(mapping.lats), 'Latitude of pixel corner', 'degrees', corner,
                                     VALIDMIN=-90.0, VALIDMAX=90.0, VAR_NOTES='Geodetic latitude')
            out['lon_bounds'] = _var(fill(mapping.lons), 'Longitude of pixel corner', 'degrees', corner,
                                     VALIDMIN=-180.0, VALIDMAX=180.0, VAR_NOTES='Geodetic longitude')
    out['altitude'] = _var(mapping.altitude * 1000, 'Height above reference ellipsoid', 'meters')
    if includeMagCoords:
        mlats, mlts = mapping.mLatMltCenter
        out['mlat'] = _var(fill(mlats), 'Geomagnetic latitude of pixel center', 'degrees', pixel,
                           VALIDMIN=-90.0, VALIDMAX=90.0)
        out['mlt'] = _var(fill(mlts), 'Magnetic local time of pixel center', 'hours', pixel, VALIDMIN=0.0, VALIDMAX=24.0)
        if includeBounds:
            mlats, mlts = mapping.mLatMlt
            out['mlat_bounds'] = _var(fill(mlats), 'Geomagnetic latitude of pixel corner', 'degrees', corner,
                                      VALIDMIN=-90.0, VALIDMAX=90.0)
            out['mlt_bounds'] = _var(fill(mlts), 'Magnetic local time of pixel corner', 'hours', corner,
                                     VALIDMIN=0.0, VALIDMAX=24.0)
    img = mapping.img
    if img.ndim == 2:
        img = img[:, :, None]
    if np.any(ma.getmaskarray(img)):
        widen = {np.dtype(np.uint8): np.int16, np.dtype(np.uint16): np.int32, np.dtype(np.uint32): np.int64}
        if img.dtype not in widen:
            raise NotImplementedError('Image data type not supported: ' + str(img.dtype))
        dt = widen[img.dtype]
        fillval = dt(np.iinfo(dt).min)
        data = img.astype(dt).filled(fillval)
    else:
        fillval, data = None, ma.getdata(img)
    if data.shape[2] == 1:
        bands = ['img']
    elif data.shape[2] == 3:
        bands = ['img_red', 'img_green', 'img_blue']
    else:
        raise NotImplementedError
    for i, band in enumerate(bands):
        extra = dict(VALIDMIN=np.iinfo(img.dtype).min, VALIDMAX=np.iinfo(img.dtype).max)
        if fillval:
            extra['FILLVAL'] = fillval
        out[band] = _var(data[np.newaxis, :, :, i], '', 'unitless', pixel, **extra)
    zen = 90 - ma.filled(mapping.elevation.astype(np.float32), np.nan)[np.newaxis, :]
    out['zenith_angle'] = _var(zen, 'Absolute sensor zenith angle of pixel center', 'degrees', pixel,
                               VALIDMIN=0.0, VALIDMAX=90.0)
    out['camera_pos'] = _var(np.array([mapping.cameraPosGCRS]), 'Camera position in cartesian GCRS coordinates',
                             'kilometers', ('Epoch',))
    return out
