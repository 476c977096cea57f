"""Plate-carree regridding of mappings by mean binning, on the GPU.

API mirror of `auromat/resample.py` (resample :73-157, resampleMLatMLT :63-71,
plateCarreeResolution :36-61, fixedGrid :281-299, ResampleProvider :370-394).  The host side
only derives the target grid (a handful of scalars, computed with the reference's own
arithmetic so that the grid is bit-identical); binning, accumulation and normalisation are
CUDA kernels (`amt_bin_accumulate`, `amt_normalise`) working on the device planes of the
mapping.  Only `method='mean'` is on the B200 path.
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import numpy.ma as ma

from . import _lib
from .coordinates import geodesic
from .coordinates.geodesic import Location, wgs84A, wgs84B
from .coordinates.transform import rotation_matrix
from .mapping.mapping import BaseMapping, MappingCollection


def plateCarreeResolution(boundingBox, arcsecPerPx):
    """(latPxPerDeg, lonPxPerDeg) approximating a spherical resolution at the centre of the
    bounding box (reference resample.py:36-61)."""
    degPerPx = arcsecPerPx * (1.0 / 3600.0)
    latPxPerDeg = 1 / degPerPx
    latMiddle = (boundingBox.latNorth + boundingBox.latSouth) / 2
    lonMiddleDistance = geodesic.angularDistance(Location(latMiddle, boundingBox.lonWest),
                                                 Location(latMiddle, boundingBox.lonEast))
    px = lonMiddleDistance / degPerPx
    lonEast = boundingBox.lonEast
    if boundingBox.lonWest > lonEast:
        lons = lonEast + 360 - boundingBox.lonWest
    else:
        lons = lonEast - boundingBox.lonWest
    return latPxPerDeg, px / lons


def _linspaceAt(start, stop, num, i):
    """np.linspace(start, stop, num)[i] without allocating the array: fl(fl(i*step) + start)
    with step = (stop - start)/(num - 1), last element == stop (plain IEEE double arithmetic,
    identical to numpy's)."""
    if i == num - 1:
        return float(stop)
    step = (float(stop) - float(start)) / float(num - 1)
    return float(i) * step + float(start)


def _firstNode(lo, hi, num, x, strict):
    """Index of the first linspace node that is > x (strict) or >= x; `num` if there is none."""
    step = (hi - lo) / (num - 1)
    ok = (lambda v: v > x) if strict else (lambda v: v >= x)
    guess = (x - lo) / step
    i = int(min(max(math.floor(guess), 0), num - 1)) if guess == guess else 0
    while i > 0 and ok(_linspaceAt(lo, hi, num, i - 1)):
        i -= 1
    while i < num and not ok(_linspaceAt(lo, hi, num, i)):
        i += 1
    return i


def _snapDown(lo, hi, num, x):
    """linspace[argmax(linspace > x) - 1] (reference resample.py:293,295): the last node that
    is <= x; numpy's argmax-of-all-False == 0 and index -1 wrap-around are kept."""
    i = _firstNode(lo, hi, num, x, True)
    if i == num:
        i = 0
    return _linspaceAt(lo, hi, num, (i - 1) % num)


def _snapUp(lo, hi, num, x):
    """linspace[argmax(linspace >= x)] (reference resample.py:294,296)."""
    i = _firstNode(lo, hi, num, x, False)
    if i == num:
        i = 0
    return _linspaceAt(lo, hi, num, i)


def fixedGrid(pxPerDeg, latMin, latMax, lonMin, lonMax):
    """Align a bounding box to the global plate-carree grid defined by `pxPerDeg`
    (reference resample.py:281-299).  Equivalent to indexing the global
    `np.linspace(-90, 90, ...)` / `np.linspace(-180, 180, ...)` node arrays, evaluated
    node-by-node instead of materialising up to millions of nodes."""
    latPxPerDeg, lonPxPerDeg = pxPerDeg
    nLatAll = int(round(latPxPerDeg * 180 + 1))
    nLonAll = int(round(lonPxPerDeg * 360 + 1))
    latMinInGrid = _snapDown(-90.0, 90.0, nLatAll, latMin)
    latMaxInGrid = _snapUp(-90.0, 90.0, nLatAll, latMax)
    lonMinInGrid = _snapDown(-180.0, 180.0, nLonAll, lonMin)
    lonMaxInGrid = _snapUp(-180.0, 180.0, nLonAll, lonMax)
    nLat = int(round(latPxPerDeg * (latMaxInGrid - latMinInGrid) + 1))
    nLon = int(round(lonPxPerDeg * (lonMaxInGrid - lonMinInGrid) + 1))
    return nLat, nLon, latMinInGrid, latMaxInGrid, lonMinInGrid, lonMaxInGrid


def _preRotation(mode, altitude, angle=90):
    """The part of `amt_grid` that describes the coordinate pre-rotation."""
    g = _lib.AmtGrid()
    g.prerotate = mode
    g.altitude = float(altitude)
    g.wgs_a, g.wgs_b = wgs84A, wgs84B
    # reference resample.py:186-189: rotation_matrix(deg2rad(90), [1,0,0])
    g.rot[:] = rotation_matrix(np.deg2rad(angle), [1, 0, 0]).ravel().tolist()
    return g


def targetGrid(pxPerDeg, latMin, latMax, lonMin, lonMax, prerotate=_lib.AMT_PRE_NONE, altitude=0.0):
    """Derive the binning grid exactly as the reference does (resample.py:220-241 and
    :330-335, util/histogram.py:185-186,215-219).  Returns (amt_grid, info dict)."""
    latPxPerDeg, lonPxPerDeg = pxPerDeg
    assert latPxPerDeg > 0 and lonPxPerDeg > 0
    nLat, nLon, latMinG, latMaxG, lonMinG, lonMaxG = fixedGrid(pxPerDeg, latMin, latMax, lonMin, lonMax)
    assert nLat > 1, 'nlat={}, latMax={}, latMin={}, pxperdeg={}'.format(nLat, latMaxG, latMinG, pxPerDeg)
    assert nLon > 1, 'nlon={}, lonMax={}, lonMin={}, pxperdeg={}'.format(nLon, lonMaxG, lonMinG, pxPerDeg)
    if nLat < 3 or nLon < 3:
        raise ValueError('the resampling grid has no interior nodes (nLat=%d, nLon=%d)' % (nLat, nLon))
    # np.linspace(..., retstep=True) steps (plain IEEE double arithmetic == numpy float64)
    latMinG, latMaxG, lonMinG, lonMaxG = float(latMinG), float(latMaxG), float(lonMinG), float(lonMaxG)
    latStep = (latMinG - latMaxG) / float(nLat - 1)
    lonStep = (lonMaxG - lonMinG) / float(nLon - 1)
    # first/last interior node (latSpaceCenter[1:-1], lonSpaceCenter[1:-1])
    latC0 = _linspaceAt(latMaxG, latMinG, nLat, 1)
    latCL = _linspaceAt(latMaxG, latMinG, nLat, nLat - 2)
    lonC0 = _linspaceAt(lonMinG, lonMaxG, nLon, 1)
    lonCL = _linspaceAt(lonMinG, lonMaxG, nLon, nLon - 2)
    g = _preRotation(prerotate, altitude)
    g.nx, g.ny = nLon - 2, nLat - 2
    # range_ of the histogram2d call (reference :333-334)
    g.lo_x, g.hi_x = lonC0 - lonStep / 2, lonCL + lonStep / 2
    g.lo_y, g.hi_y = latCL + latStep / 2, latC0 - latStep / 2
    g.step_x = (g.hi_x - g.lo_x) / float(g.nx)
    g.step_y = (g.hi_y - g.lo_y) / float(g.ny)

    def roundScale(lo, hi, n, step):
        # decimal = int(-log10(dedges.min())) + 6 ; dedges = diff(linspace(lo, hi, n+1)).
        # Every edge difference is `step` up to a few ulps, so the integer part of -log10 is that
        # of `step` unless step sits within 1e-9 (relative) of a power of ten; only then the
        # edges are materialised as numpy does.
        d = -math.log10(step)
        if abs(d - round(d)) > 1e-9:
            return 10.0 ** (int(d) + 6)
        k = np.arange(n + 1, dtype=np.float64)
        e = k * step + lo
        e[-1] = hi
        return float(10.0 ** (int(-np.log10(np.diff(e).min())) + 6))

    g.round_x = roundScale(g.lo_x, g.hi_x, g.nx, g.step_x)
    g.round_y = roundScale(g.lo_y, g.hi_y, g.ny, g.step_y)
    info = dict(nLat=nLat, nLon=nLon, latMinInGrid=latMinG, latMaxInGrid=latMaxG, lonMinInGrid=lonMinG,
                lonMaxInGrid=lonMaxG, latStep=latStep, lonStep=lonStep)
    return g, info


def sideScale(nSamples, maxAbs=128.0):
    """`amt_grid.side_scale` for the fixed-point elevation sums: the largest power of two (<= 2**40)
    such that nSamples values of magnitude < maxAbs, scaled and rounded, add up below 2**62 -- whatever
    the distribution of the samples over the cells.  The sums are then exact integers: independent of
    the order of the atomics, identical from run to run and across ranks."""
    k = 62 - int(math.ceil(math.log2(maxAbs))) - int(math.ceil(math.log2(max(2, int(nSamples)))))
    return float(2.0 ** max(1, min(40, k)))


def resampleMLatMLT(mapping, **kw):
    """Resample such that MLat/MLT become regular grids (reference resample.py:63-71)."""
    from .mapping.mapping import convertMappingToSM, convertSMMappingToGeo
    return convertSMMappingToGeo(resample(convertMappingToSM(mapping), **kw))


def binMappingInto(mapping, grid, count, sums, fsum, nearEdge=None):
    """Accumulate one mapping into existing sum/count grids (mosaic building block)."""
    ctx = mapping.context
    p = mapping.devicePlanes()
    img = mapping.deviceImage()
    h, w = mapping.shape
    # Only the rows that hold defined pixels are handed to the kernel (the outline statistics
    # carry the pixel box of the valid centres): an ISS limb frame is ~40 % empty sky.
    st = mapping._deviceStats()
    r0, r1 = max(0, st.row_min_c), min(h - 1, st.row_max_c)
    if r1 < r0:
        return
    lo, hi = r0 * w, (r1 + 1) * w
    flat = lambda t: t.reshape(-1)[lo:hi]
    ctx.bin_accumulate(flat(p['lat_c']), flat(p['lon_c']), flat(p['elev_c']) if fsum is not None else None,
                       img[r0:r1 + 1], grid, count, sums, fsum, nearEdge)


def deriveGrid(mapping, pxPerDeg=25, arcsecPerPx=None, containsPole=None):
    """Target grid of `resample(mapping, ...)`: (amt_grid, info).  Needs the outline statistics of
    the mapping only (bounding box, pole / date-line flags); with a pre-rotation one more outline
    reduction in the rotated frame (reference resample.py:95-117,176-227)."""
    ctx = mapping.context
    if containsPole is None:
        containsPole = mapping.containsPole
    bbox = mapping.boundingBox
    if arcsecPerPx:
        pxPerDeg = plateCarreeResolution(bbox, arcsecPerPx)
    else:
        try:
            _, _ = pxPerDeg
        except TypeError:
            assert pxPerDeg is not None
            pxPerDeg = (pxPerDeg, pxPerDeg)
    latMin, latMax, lonMin, lonMax = bbox.latSouth, bbox.latNorth, bbox.lonWest, bbox.lonEast
    mode = _lib.AMT_PRE_NONE
    if containsPole:
        mode = _lib.AMT_PRE_POLE
    elif mapping.containsDiscontinuity:
        mode = _lib.AMT_PRE_WRAP180
    planeFree = getattr(mapping, '_planeFree', False) and 'lat_c' not in mapping._planes
    if mode != _lib.AMT_PRE_NONE:
        # min/max of the rotated outline (reference resample.py:176-216)
        h, w = mapping.shape
        st = ctx.new_stats()
        if planeFree:
            vk, vc = mapping._ensureHitBitmaps()
            ctx.bbox_stats_frame(mapping.frameConstants, vk, vc, st, pre=_preRotation(mode, mapping.altitude))
        else:
            ctx.bbox_stats(w, h, mapping.devicePlanes(), st, pre=_preRotation(mode, mapping.altitude))
        s = ctx.read_stats(st)
        lonMin, lonMax = s.lon_min, s.lon_max
        if mode == _lib.AMT_PRE_POLE:
            latMin, latMax = s.lat_min, s.lat_max
    grid, info = targetGrid(pxPerDeg, latMin, latMax, lonMin, lonMax, mode, mapping.altitude)
    info['pxPerDeg'] = pxPerDeg
    info['mode'] = mode
    if getattr(mapping, '_finiteElevation', False):
        # elevations computed on the device are finite and in [-90, 90]: exact integer sums
        grid.side_scale = sideScale(mapping.shape[0] * mapping.shape[1])
    return grid, info


def resampleToDevice(mapping, pxPerDeg=25, arcsecPerPx=None, containsPole=None):
    """The device-resident part of `resample`: returns (grid, info, img, mask, elevation)
    with the outputs still in HBM."""
    import torch
    ctx = mapping.context
    grid, info = deriveGrid(mapping, pxPerDeg, arcsecPerPx, containsPole)
    planeFree = getattr(mapping, '_planeFree', False) and 'lat_c' not in mapping._planes
    img = mapping.deviceImage()
    channels = img.shape[2]
    cells = grid.nx * grid.ny
    acc = ctx.zeros((2 + channels) * cells, torch.int64)       # count | sums[channels] | fsum (f64 bits)
    count = acc[:cells]
    sums = acc[cells:(1 + channels) * cells]
    fsum = acc[(1 + channels) * cells:].view(torch.float64)
    if planeFree:
        # fused centre chain + binning: no coordinate plane is materialised
        ctx.georef_bin_fused(mapping.frameConstants, mapping._ensureHitBitmaps()[1], img, grid, count, sums, fsum)
    else:
        binMappingInto(mapping, grid, count, sums, fsum)
    outImg, outMask, outElev = ctx.normalise(grid, img.dtype, channels, count, sums, fsum)
    info['count'] = count
    return grid, info, outImg, outMask, outElev


def resample(mappingOrCollection, pxPerDeg=25, arcsecPerPx=None, containsPole=None, method='mean'):
    """Return a new mapping (or collection) whose colours and elevation are resampled onto a
    regular latitude/longitude grid (plate-carree), y=latitude, x=longitude.

    Same signature and semantics as the reference's `resample` (resample.py:73-157) for
    `method='mean'`; the interpolating methods ('nearest', 'linear', 'cubic' --
    scipy.interpolate.griddata in the reference) are not part of the B200 path.
    """
    if method != 'mean':
        raise NotImplementedError("only method='mean' runs on the B200 path (got %r)" % (method,))

    def doResample(mapping):
        ctx = mapping.context
        grid, info, outImg, outMask, outElev = resampleToDevice(mapping, pxPerDeg, arcsecPerPx, containsPole)
        lat_k, lon_k, lat_c, lon_c = ctx.plate_carree_coords(grid.nx, grid.ny, info['latMaxInGrid'],
                                                             info['latMinInGrid'], info['lonMinInGrid'],
                                                             info['lonMaxInGrid'])
        if info['mode'] == _lib.AMT_PRE_POLE:
            back = _preRotation(_lib.AMT_PRE_POLE, mapping.altitude, angle=-90)
            ctx.rotate_coords(lat_k, lon_k, back)
            ctx.rotate_coords(lat_c, lon_c, back)
        elif info['mode'] == _lib.AMT_PRE_WRAP180:
            back = _preRotation(_lib.AMT_PRE_WRAP180, mapping.altitude)
            ctx.rotate_coords(lat_k, lon_k, back)
            ctx.rotate_coords(lat_c, lon_c, back)
        img = ctx.to_numpy(outImg)
        mask = ctx.to_numpy(outMask).astype(bool)
        if mapping.img_unmasked.ndim == 2:
            img = img.reshape(img.shape[0], img.shape[1])
            imgMask = mask
        else:
            imgMask = np.repeat(mask[:, :, None], img.shape[2], 2)
        img = ma.masked_array(img, mask=imgMask)
        return mapping.createResampled(lat_k, lon_k, lat_c, lon_c, outElev, img)

    if isinstance(mappingOrCollection, BaseMapping):
        return doResample(mappingOrCollection)
    if isinstance(mappingOrCollection, MappingCollection):
        c = mappingOrCollection
        return MappingCollection([doResample(m) for m in c.mappings], c.identifier, mayOverlap=c.mayOverlap)
    raise ValueError('First argument must be a mapping or a mapping collection, but is: {}'.format(
        type(mappingOrCollection)))


def ResampleProvider(provider, **kw):
    """Wrap a mapping provider so that every returned mapping is resampled
    (reference resample.py:370-394)."""
    from .mapping.mapping import _wrapProvider
    return _wrapProvider(provider, partial(resample, **kw), 'ResamplingProvider')
