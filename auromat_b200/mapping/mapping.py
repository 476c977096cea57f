"""Georeferenced-image objects whose coordinate arrays live in B200 HBM.

API mirror of `auromat/mapping/mapping.py` of the reference (BoundingBox :44-287,
BaseMapping :293-929, GenericMapping :1233-1310, MappingCollection :1315-1373,
checkPlateCarree :931-975, convertMappingToSM :1519-1559): same names, same array shapes and
the same masked-array contract -- `lats, lons` (h+1,w+1), `latsCenter, lonsCenter,
elevation` (h,w), `img` (h,w,n), `mLatMlt`, `mLatMltCenter`, all `numpy.ma` float64 with
NaN <=> masked -- but the arrays are *device planes* produced by the CUDA kernels and are
copied to the host only when a numpy-facing property is read.  `resample()` and the masking
methods consume the device planes directly.
"""
from __future__ import annotations

import copy
import warnings
from collections import namedtuple

import numpy as np
import numpy.ma as ma

from .. import _lib
from .. import utils as polyutils
from ..coordinates import geodesic, transform
from ..coordinates.geodesic import Location, wgs84A, wgs84B
from ..runtime import get_context

Size = namedtuple('Size', ['width', 'height'])
MappingProperties = namedtuple('MappingProperties',
                               'altitude cameraPosGCRS boundingBox photoTime '
                               'centroid cameraFootpoint identifier')

PixelScales = namedtuple('PixelScales', ['width', 'height', 'diagonal'])
PixelScale = namedtuple('PixelScale', ['mean', 'median', 'min', 'max'])

CORNER_PLANES = ('lat_k', 'lon_k', 'mlat_k', 'mlt_k')
CENTER_PLANES = ('lat_c', 'lon_c', 'mlat_c', 'mlt_c', 'elev_c')


def wrapAt180(x):
    """`Angle(x deg).wrap_at(180 deg).degree`: values in [-180, 180) are returned unchanged,
    everything else is shifted by whole turns (astropy's `_wrap_at`)."""
    a = np.array(x, dtype=np.float64, copy=True, ndmin=1)
    with np.errstate(invalid='ignore'):
        wraps = (a + 180.0) // 360.0
    valid = np.isfinite(wraps) & (wraps != 0)
    if np.any(valid):
        a -= wraps * 360.0
        a[a >= 180.0] -= 360.0
        a[a < -180.0] += 360.0
    return a if np.ndim(x) else float(a[0])


class BoundingBox(object):
    """A geographical bounding box that may span the 180-degree discontinuity
    (reference mapping.py:44-287)."""

    def __init__(self, latSouth, lonWest, latNorth, lonEast):
        assert -180 <= lonWest <= 180, 'Longitude: ' + str(lonWest)
        assert -180 <= lonEast <= 180, 'Longitude: ' + str(lonEast)
        assert -90 <= latSouth <= 90, 'Latitude: ' + str(latSouth)
        assert -90 <= latNorth <= 90, 'Latitude: ' + str(latNorth)
        self._box = (latSouth, lonWest, latNorth, lonEast)

    latSouth = property(lambda self: self._box[0])
    lonWest = property(lambda self: self._box[1])
    latNorth = property(lambda self: self._box[2])
    lonEast = property(lambda self: self._box[3])
    topLeft = property(lambda self: Location(self.latNorth, self.lonWest))
    bottomLeft = property(lambda self: Location(self.latSouth, self.lonWest))
    topRight = property(lambda self: Location(self.latNorth, self.lonEast))
    bottomRight = property(lambda self: Location(self.latSouth, self.lonEast))

    @property
    def _minSphericalRectangle(self):
        """(center, size in km) of the smallest spherical rectangle around the box (reference
        mapping.py:119-172; the size is only meaningful below 180 degrees of longitude)."""
        if getattr(self, '_rect', None) is None:
            if self.containsPole:
                north = self.latNorth == 90
                center = Location(90 if north else -90, 0)
                width = geodesic.distance(center, Location(self.latSouth if north else self.latNorth, 0)) * 2
                height = width
            else:
                lonWest, lonEast = self.lonWest, self.lonEast
                if lonWest > lonEast:
                    lonEast += 360
                lonc = float(wrapAt180((lonWest + lonEast) / 2))
                if lonEast - lonWest > 180:
                    warnings.warn('The bounding box spans more than 180deg in longitude. '
                                  'The returned size of the minimum spherical rectangle will be incorrect.')
                width = geodesic.distance(self.bottomLeft, self.bottomRight)
                width2 = geodesic.distance(self.topLeft, self.topRight)
                if width2 > width:      # southern hemisphere
                    width = width2
                    wideCenter = geodesic.intermediate(self.bottomLeft, self.bottomRight, 0.5)
                    dataCenter = Location(self.latNorth, lonc)
                else:                   # northern hemisphere
                    wideCenter = geodesic.intermediate(self.topLeft, self.topRight, 0.5)
                    dataCenter = Location(self.latSouth, lonc)
                height = geodesic.distance(dataCenter, wideCenter)
                center = geodesic.intermediate(dataCenter, wideCenter, 0.5)
            self._rect = center, Size(width / 1000, height / 1000)
        return self._rect

    center = property(lambda self: self._minSphericalRectangle[0])
    size = property(lambda self: self._minSphericalRectangle[1])

    @property
    def containsDiscontinuity(self):
        return self.lonWest > self.lonEast or self.containsPole

    @property
    def containsPole(self):
        return self.lonWest == -180 and self.lonEast == 180 and (self.latNorth == 90 or self.latSouth == -90)

    @staticmethod
    def minimumBoundingBox(latLons):
        return BoundingBox.mergedBoundingBoxes([BoundingBox(lat, lon, lat, lon) for lat, lon in latLons])

    @staticmethod
    def mergedBoundingBoxes(boundingBoxes):
        """Smallest box containing all given boxes: latitude by min/max, longitude by
        removing the largest uncovered gap on the circle (reference mapping.py:236-282)."""
        boxes = list(boundingBoxes)
        latSouth = min(bb.latSouth for bb in boxes)
        latNorth = max(bb.latNorth for bb in boxes)
        spans = []
        for bb in boxes:
            w, e = bb.lonWest, bb.lonEast
            spans.append((w, e if e >= w else e + 360))
        xs = np.sort(np.array([[bb.lonWest, bb.lonEast] for bb in boxes]).ravel())
        xs = np.concatenate((xs, [xs[0] + 360]))
        best, bestIdx = -1.0, 0
        for i in range(1, len(xs)):
            lo, hi = xs[i - 1], xs[i]
            covered = any((w <= lo and e >= hi) or (w <= lo + 360 and e >= hi + 360) or
                          (w <= lo - 360 and e >= hi - 360) for w, e in spans)
            if not covered and hi - lo > best:
                best, bestIdx = hi - lo, i - 1
        lonWest = wrapAt180(xs[bestIdx + 1])
        lonEast = wrapAt180(xs[bestIdx])
        return BoundingBox(latSouth, lonWest, latNorth, lonEast)

    def __eq__(self, obj):
        return isinstance(obj, BoundingBox) and self._box == obj._box

    def __ne__(self, obj):
        return not self == obj

    def __hash__(self):
        return hash(self._box)

    def __repr__(self):
        return 'BoundingBox(latSouth={0}, lonWest={1}, latNorth={2}, lonEast={3})'.format(*self._box)


class BaseMapping(object):
    """Base class of all mappings: a georeferenced image at a given altitude.

    Guarantees (reference mapping.py:299-316) hold on the device planes after
    `_ensurePlanes()`: corner lat/lon are NaN together, centre lat/lon/elevation are NaN
    together, a centre is defined only if its 4 corners are, a corner only if one of its
    neighbouring centres is, and `img` is masked exactly where `latsCenter` is.

    Subclasses provide `_computePlanes(ctx, names)` (fill `self._planes` with device
    tensors) and the image (`img_unmasked`).
    """

    def __init__(self, altitude, cameraPosGCRS, photoTime, identifier, metadata=None, device=None):
        assert altitude >= 0
        cameraPosGCRS = np.asarray(cameraPosGCRS)
        assert cameraPosGCRS.shape == (3,)
        self._altitude = altitude
        self._cameraPosGCRS = cameraPosGCRS
        self._photoTime = photoTime
        self._identifier = identifier
        self._metadata = metadata
        self._device = device
        self._planes = {}          # name -> device tensor (float64, NaN == masked)
        self._host = {}            # name -> numpy masked array cache
        self._stats = None
        self._statsPending = None
        self._statsDevice = None   # device amt_stats block (georeference kernels count grazing rays into it)
        self._grazingCounted = False
        self._boundingBox = None
        self._imgDevice = None

    # ------------------------------------------------------------------ device side
    @property
    def context(self):
        return get_context(self._device)

    @property
    def shape(self):
        """(h, w) of the image."""
        raise NotImplementedError

    def _computePlanes(self, ctx, names):
        raise NotImplementedError

    # Pole containment: by default the per-pixel longitude winding test on the device;
    # mappings with a camera model override `_poleFlags` with an exact geometric test.
    _poleTestOnDevice = True

    def _poleFlags(self):
        return 0

    def _ensurePlanes(self, names):
        missing = [n for n in names if n not in self._planes]
        if missing:
            self._computePlanes(self.context, missing)
        return self._planes

    def devicePlanes(self, magnetic=False):
        """Device tensors of the coordinate planes (computing them if necessary)."""
        names = ['lat_k', 'lon_k', 'lat_c', 'lon_c', 'elev_c']
        if magnetic:
            names += ['mlat_k', 'mlt_k', 'mlat_c', 'mlt_c']
        return self._ensurePlanes(names)

    def prefetch(self, magnetic=True):
        """Run the georeferencing kernels now (one fused launch) instead of on first access."""
        self.devicePlanes(magnetic=magnetic)
        return self

    def deviceImage(self):
        """The raw (unmasked) image as a device tensor (h, w, n)."""
        if self._imgDevice is None:
            img = self.img_unmasked
            if img.ndim == 2:
                img = img[..., None]
            self._imgDevice = self.context.to_device(img)
        return self._imgDevice

    def _download(self, name):
        if name not in self._host:
            h, w = self.shape
            t = self._ensurePlanes([name])[name]
            shape = (h + 1, w + 1) if name.endswith('_k') else (h, w)
            arr = self.context.to_numpy(t).reshape(shape)
            self._host[name] = ma.masked_invalid(arr, copy=False)
        return self._host[name]

    def _startStats(self):
        """Enqueue the outline / bounding-box reductions and their read-back without waiting
        (the sequence pipeline overlaps the wait with the next frame's host work)."""
        if self._stats is None and self._statsPending is None:
            ctx = self.context
            p = self.devicePlanes()
            h, w = self.shape
            st = self._statsDevice if self._statsDevice is not None else ctx.new_stats()
            ctx.bbox_stats(w, h, p, st, pole_test=self._poleTestOnDevice)
            self._statsPending = ctx.start_stats_readback(st)

    def _deviceStats(self):
        if self._stats is None:
            self._startStats()
            s = self.context.finish_stats(self._statsPending)
            self._statsPending = None
            if not self._poleTestOnDevice:
                s.pole_flags = self._poleFlags()
            self._stats = s
        return self._stats

    # ------------------------------------------------------------------ plain attributes
    altitude = property(lambda self: self._altitude)
    cameraPosGCRS = property(lambda self: self._cameraPosGCRS)
    photoTime = property(lambda self: self._photoTime)
    identifier = property(lambda self: self._identifier)

    @property
    def metadata(self):
        return {} if self._metadata is None else self._metadata

    @property
    def properties(self):
        return MappingProperties(identifier=self.identifier, altitude=self.altitude,
                                 cameraPosGCRS=self.cameraPosGCRS, boundingBox=self.boundingBox,
                                 photoTime=self.photoTime, centroid=self.centroid,
                                 cameraFootpoint=self.cameraFootpoint)

    @property
    def cameraFootpoint(self):
        """Camera footpoint in geodetic coordinates (reference mapping.py:443-452)."""
        et = transform.date2es(self.photoTime)
        g = transform.mat_j2000_to_geo(et).dot(np.asarray(self.cameraPosGCRS, dtype=np.float64))
        lat, lon = transform.ecef2GeodeticScalar(g[0], g[1], g[2])
        return Location(np.rad2deg(lat), np.rad2deg(lon))

    # ------------------------------------------------------------------ numpy-facing arrays
    lats = property(lambda self: self._download('lat_k'))
    lons = property(lambda self: self._download('lon_k'))
    latsCenter = property(lambda self: self._download('lat_c'))
    lonsCenter = property(lambda self: self._download('lon_c'))
    elevation = property(lambda self: self._download('elev_c'))

    @property
    def mLatMlt(self):
        return self._download('mlat_k'), self._download('mlt_k')

    @property
    def mLatMltCenter(self):
        return self._download('mlat_c'), self._download('mlt_c')

    @property
    def img_unmasked(self):
        raise NotImplementedError

    @property
    def img(self):
        """Masked (h,w,n) image; masked exactly where `latsCenter` is."""
        if 'img' not in self._host:
            data = self.img_unmasked
            mask = ma.getmaskarray(self.latsCenter)
            if data.ndim == 3:
                mask = np.repeat(mask[:, :, None], data.shape[2], 2)
            self._host['img'] = ma.masked_array(data, mask=mask)
        return self._host['img']

    @property
    def rgb_unmasked(self):
        img = self.img_unmasked
        if img.dtype == np.uint16:
            img = (img * (255 / 65535)).astype(np.uint8)
        elif img.dtype != np.uint8:
            raise NotImplementedError
        if img.ndim == 2:
            img = img[..., None]
        if img.shape[2] == 3:
            return img
        if img.shape[2] == 1:
            return np.repeat(img, 3, 2)
        raise NotImplementedError('Unknown img format')

    @property
    def rgb(self):
        mask = np.repeat(ma.getmaskarray(self.latsCenter)[:, :, None], 3, 2)
        return ma.masked_array(self.rgb_unmasked, mask=mask)

    # ------------------------------------------------------------------ derived geometry
    @property
    def boundingBox(self):
        """Min/max of the outline (the boundary of the valid-corner mask), widened to the full
        longitude range when a pole is enclosed (reference mapping.py:694-743).  The
        reductions and the pole test run on the device (`amt_bbox_stats`)."""
        if self._boundingBox is None:
            s = self._deviceStats()
            if s.n_boundary_corners == 0:
                raise ValueError('the mapping has no defined coordinates')
            latMin, latMax, lonMin, lonMax = s.lat_min, s.lat_max, s.lon_min, s.lon_max
            if s.pole_flags:
                if latMax < 0:
                    box = (-90, -180, latMax, 180)
                else:
                    box = (latMin, -180, 90, 180)
            elif lonMax - lonMin > 180:
                box = (latMin, s.lon_min_pos, latMax, s.lon_max_neg)
            else:
                box = (latMin, lonMin, latMax, lonMax)
            self._boundingBox = BoundingBox(*box)
        return self._boundingBox

    containsDiscontinuity = property(lambda self: self.boundingBox.containsDiscontinuity)
    containsPole = property(lambda self: self.boundingBox.containsPole)

    @property
    def outline(self):
        """The complete outline of this mapping as (n,2) [lat, lon]: the ordered (clockwise in
        image coordinates) walk around the valid-corner mask; can be concave (reference
        mapping.py:655-662,672-680)."""
        return self._fullAndConvexOutlines[0]

    @property
    def outlineConvexHull(self):
        """The convex hull (in corner-index space) of the outline, as (m,2) [lat, lon]
        (reference mapping.py:664-670,682-688)."""
        return self._fullAndConvexOutlines[1]

    def _validCornerMask(self):
        """(h+1, w+1) bool array of the defined corners.  Read from the 1-bit validity bitmap on
        the device when the corner planes are not on the host yet (1.5 MB instead of 2 x 96 MB
        for a 12-Mpixel frame)."""
        h, w = self.shape
        if 'lat_k' in self._host:
            return ~ma.getmaskarray(self._host['lat_k'])
        p = self.devicePlanes()
        if 'valid_k' not in p:
            self.context.valid_bits(w, h, p)
        words = self.context.to_numpy(p['valid_k']).view(np.uint32).reshape(h + 1, -1)
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder='little')
        return bits[:, :w + 1].astype(bool)

    def _gatherCorners(self, xy):
        """[lat, lon] of the corner nodes xy (n,2) in x,y order."""
        h, w = self.shape
        if 'lat_k' in self._host and 'lon_k' in self._host:
            return np.transpose([self._host['lat_k'].data[xy[:, 1], xy[:, 0]],
                                 self._host['lon_k'].data[xy[:, 1], xy[:, 0]]])
        p = self.devicePlanes()
        idx = self.context.to_device(np.ascontiguousarray(xy[:, 1] * (w + 1) + xy[:, 0], dtype=np.int64))
        lat = self.context.to_numpy(p['lat_k'].reshape(-1).index_select(0, idx))
        lon = self.context.to_numpy(p['lon_k'].reshape(-1).index_select(0, idx))
        return np.transpose([lat, lon])

    @property
    def _fullAndConvexOutlines(self):
        if getattr(self, '_outlines', None) is None:
            outl = polyutils.outline(self._validCornerMask())
            hull = polyutils.convexHull(outl)
            both = self._gatherCorners(np.concatenate((outl, hull)))
            self._outlines = both[:len(outl)], both[len(outl):]
        return self._outlines

    @property
    def centroid(self):
        """The centroid of the outline polygon in the plate-carree projection (reference
        mapping.py:759-783)."""
        if getattr(self, '_centroid', None) is None:
            if self.containsPole:
                raise NotImplementedError
            outl = self.outline
            if self.containsDiscontinuity:
                outl = np.transpose([outl[:, 0], wrapAt180(outl[:, 1] + 180)])
                lat, lon = polyutils.polygonCentroid(outl)
                self._centroid = Location(lat, float(wrapAt180(lon + 180)))
            else:
                self._centroid = Location(*polyutils.polygonCentroid(outl))
        return self._centroid

    @property
    def arcSecPerPx(self):
        """Min, max, median and mean angular sizes (arcsec) of the width, height and diagonal of
        (up to) 1000 evenly sampled pixel polygons (reference mapping.py:785-843)."""
        if getattr(self, '_pixelScales', None) is None:
            valid = self._validCornerMask()
            ok = valid[:-1, :-1] & valid[:-1, 1:] & valid[1:, 1:] & valid[1:, :-1]
            ys, xs = np.nonzero(ok)
            polyCount = len(ys)
            sampleCount = min(polyCount, 1000)
            sel = np.round(np.linspace(0, polyCount - 1, sampleCount)).astype(int)
            ys, xs = ys[sel], xs[sel]
            # polygon vertices 0,1,2 = (y,x), (y,x+1), (y+1,x+1)
            nodes = np.concatenate((np.transpose([xs, ys]), np.transpose([xs + 1, ys]), np.transpose([xs + 1, ys + 1])))
            ll = self._gatherCorners(nodes).reshape(3, sampleCount, 2)
            v0, v1, v2 = ll[0], ll[1], ll[2]
            scales = []
            for a, b in ((v0, v1), (v1, v2), (v0, v2)):
                deg = np.array([geodesic.angularDistance(Location(*p), Location(*q)) for p, q in zip(a, b)])
                scales.append(PixelScale(mean=float(np.mean(deg)) * 3600.0, median=float(np.median(deg)) * 3600.0,
                                         min=float(deg.min()) * 3600.0, max=float(deg.max()) * 3600.0))
            self._pixelScales = PixelScales(width=scales[0], height=scales[1], diagonal=scales[2])
        return self._pixelScales

    @property
    def isPlateCarree(self):
        return isPlateCarree(self.lats, self.lons)

    def checkPlateCarree(self):
        return checkPlateCarree(self.lats, self.lons)

    # ------------------------------------------------------------------ masking
    def createMasked(self, centerMask):
        """Copy of this mapping with `centerMask` (True == masked) applied to the centre
        planes; corners left without any defined neighbour are masked as well."""
        centerMask = np.ascontiguousarray(centerMask, dtype=np.uint8)
        assert centerMask.shape == tuple(self.shape)
        return self._maskedCopy(mask=centerMask)

    def maskedByElevation(self, minElevation=10):
        """New mapping with data below `minElevation` degrees masked (reference
        mapping.py:845-864); runs entirely on the device planes."""
        m = self._maskedCopy(minElevation=float(minElevation))
        if m._deviceStats().n_valid_centers == 0:
            raise ValueError('minElevation=' + str(minElevation) + ' would mask all pixels!')
        return m

    def maskedByPolygon(self, polygon):
        """Copy of this mapping where only those pixels are retained whose four corners all lie
        inside `polygon` (ordered points of an unclosed polygon, [lat, lon]); a previously
        applied mask is ignored by the reference and kept here (the planes carry it).  Mappings
        or polygons across the date line / a pole are rotated first, best effort, in the
        reference's order (mapping.py:866-917).  The inside test runs on the device
        (`amt_polygon_center_mask`)."""
        from ..resample import _preRotation
        ctx = self.context
        polygon = np.array(polygon, dtype=np.float64)
        assert polygon.ndim == 2 and polygon.shape[1] == 2 and len(polygon) >= 3
        polyBoundingBox = BoundingBox.minimumBoundingBox(polygon)
        polyContainsPole = geodesic.containsOrCrossesPole(polygon)
        pre = None
        if self.containsDiscontinuity or polyBoundingBox.containsDiscontinuity:
            pre = _preRotation(_lib.AMT_PRE_WRAP180, self.altitude)
        elif self.containsPole or polyContainsPole:
            pre = _preRotation(_lib.AMT_PRE_POLE, self.altitude)
        dpoly = ctx.to_device(polygon)
        if pre is not None:
            # the polygon goes through the same device routine as the corner coordinates
            plat, plon = dpoly[:, 0].contiguous(), dpoly[:, 1].contiguous()
            ctx.rotate_coords(plat, plon, pre)
            import torch
            dpoly = torch.stack((plat, plon), dim=1).contiguous()
        p = self.devicePlanes()
        h, w = self.shape
        mask, nInside = ctx.polygon_center_mask(w, h, p['lat_k'], p['lon_k'], dpoly, pre)
        if int(nInside.item()) == 0:
            raise ValueError('The given mask would mask all pixels!')
        return self._maskedCopy(deviceMask=mask)

    def _maskedCopy(self, mask=None, deviceMask=None, minElevation=float('nan')):
        ctx = self.context
        # all planes are materialised first: a later georeference launch would not know the mask
        src = self.devicePlanes(magnetic=True)
        m = copy.copy(self)
        m._planes = {k: v.clone() for k, v in src.items()}
        m._host = {}
        m._stats = None
        m._statsPending = None
        m._statsDevice = None
        m._boundingBox = None
        m._outlines = m._centroid = m._pixelScales = None
        m.__dict__.pop('_ringSlot', None)          # the copy owns its planes
        m._planeFree = False
        h, w = self.shape
        dmask = deviceMask if deviceMask is not None else (ctx.to_device(mask.ravel()) if mask is not None else None)
        ctx.apply_center_mask(w, h, m._planes, dmask, minElevation)
        return m

    def _detachRing(self):
        """The sequence pipeline recycles the ring slot whose planes this mapping was showing: drop
        them (and the ring copy of the image); any later access recomputes into fresh buffers."""
        if self.__dict__.pop('_ringSlot', None) is None:
            return
        ringImg = self._imgDevice is not None and getattr(self, '_imgDevice_in', None) is not self._imgDevice
        self._planes = {}
        self.__dict__.pop('_planeBuffers', None)
        self._planeFree = False
        if ringImg:
            self._imgDevice = None

    def setDirty(self):
        self._boundingBox = None
        self._outlines = self._centroid = self._pixelScales = None
        self._stats = None
        self._statsPending = None

    # ------------------------------------------------------------------ invariants
    def checkGuarantees(self):
        """Test helper, same assertions as reference mapping.py:362-428."""
        lats, lons = self.lats, self.lons
        latsCenter, lonsCenter = self.latsCenter, self.lonsCenter
        mlat, mlt = self.mLatMlt
        mlatCenter, mltCenter = self.mLatMltCenter
        img, elevation = self.img, self.elevation
        for a in (lats, latsCenter, mlat, elevation):
            assert not np.any(np.isnan(a))
        mk, mc = ma.getmaskarray(lats), ma.getmaskarray(latsCenter)
        assert np.array_equal(mk, ma.getmaskarray(lons))
        assert np.array_equal(mc, ma.getmaskarray(lonsCenter))
        pad = np.zeros((mc.shape[0] + 2, mc.shape[1] + 2), bool)
        pad[1:-1, 1:-1] = ~mc
        assert np.all(mk | pad[1:, 1:] | pad[1:, :-1] | pad[:-1, :-1] | pad[:-1, 1:])
        ok = ~mk
        assert np.all(mc | (ok[:-1, :-1] & ok[1:, :-1] & ok[1:, 1:] & ok[:-1, 1:]))
        imgMask = ma.getmaskarray(img)
        imgMask = imgMask.reshape(imgMask.shape[0], imgMask.shape[1], -1)
        for d in range(imgMask.shape[2]):
            assert np.array_equal(imgMask[:, :, d], mc)
        assert np.array_equal(ma.getmaskarray(elevation), mc)
        assert np.array_equal(ma.getmaskarray(mlatCenter), mc)
        assert np.array_equal(ma.getmaskarray(mltCenter), mc)
        assert np.array_equal(ma.getmaskarray(mlat), mk)
        assert np.array_equal(ma.getmaskarray(mlt), mk)

    # ------------------------------------------------------------------ to be overridden
    def createResampled(self, lats, lons, latsCenter, lonsCenter, elevation, img):
        return GenericMapping(lats, lons, latsCenter, lonsCenter, elevation, self.altitude, img,
                              self.cameraPosGCRS, self.photoTime, self.identifier, metadata=self.metadata,
                              device=self._device)


def checkPlateCarree(lats, lons):
    """Raise ValueError unless the 2-D coordinate arrays describe a plate-carree grid:
    latitudes evenly spaced and decreasing, longitudes evenly spaced and increasing
    (reference mapping.py:931-960)."""
    if ma.isMaskedArray(lats):
        lats, lons = lats.data, lons.data
    if np.any(np.isnan(lats)):
        raise ValueError('coordinates contains NaNs')
    lons = np.unwrap(np.deg2rad(lons))
    if lons[0, -1] - lons[0, 0] <= 0:
        raise ValueError('longitudes are not monotonically increasing')
    if lats[0, 0] - lats[-1, 0] <= 0:
        raise ValueError('latitudes are not monotonically decreasing')
    eps = 1e-4
    dLon = np.diff(lons[0])
    if not np.max(dLon) - np.min(dLon) < eps:
        raise ValueError('longitudes are not evenly spaced; max delta: {}'.format(np.max(dLon) - np.min(dLon)))
    dLat = -np.diff(lats[:, 0])
    if not np.max(dLat) - np.min(dLat) < eps:
        raise ValueError('latitudes are not evenly spaced; max delta: {}'.format(np.max(dLat) - np.min(dLat)))


def isPlateCarree(lats, lons):
    try:
        checkPlateCarree(lats, lons)
    except Exception:
        return False
    return True


class GenericMapping(BaseMapping):
    """A mapping built from precalculated coordinate arrays (numpy, masked numpy or device
    tensors), e.g. the result of `resample()` (reference mapping.py:1233-1310).  The arrays
    are uploaded once and sanitised on the device."""

    def __init__(self, lats, lons, latsCenter, lonsCenter, elev, alti, img, cameraPosGCRS, photoTime,
                 identifier, metadata=None, device=None, sanitize=True):
        h, w = img.shape[0], img.shape[1]
        assert tuple(lats.shape) == tuple(lons.shape) == (h + 1, w + 1)
        assert tuple(latsCenter.shape) == tuple(lonsCenter.shape) == (h, w)
        assert elev is None or tuple(elev.shape) == (h, w)
        assert img.dtype in (np.uint8, np.uint16)
        BaseMapping.__init__(self, alti, cameraPosGCRS, photoTime, identifier, metadata, device)
        self._shape = (h, w)
        if ma.isMaskedArray(img):
            imgMask = ma.getmaskarray(img).reshape(h, w, -1)[:, :, 0]
            img = img.data
        else:
            imgMask = None
        self._imgData = img
        ctx = self.context

        # The coordinate values handed in stay available as the `.data` of the masked arrays
        # (the reference's sanitisation only touches masks); the device planes carry NaN
        # wherever the sanitised mask is set.
        self._raw = {}

        def up(name, a):
            if hasattr(a, 'data_ptr'):          # already a device tensor
                self._raw[name] = a.reshape(-1)
                return a.reshape(-1).clone()
            if ma.isMaskedArray(a):
                a = a.astype(np.float64).filled(np.nan)
            a = np.asarray(a, dtype=np.float64)
            self._raw[name] = a
            return ctx.to_device(a).reshape(-1)

        self._planes = dict(lat_k=up('lat_k', lats), lon_k=up('lon_k', lons), lat_c=up('lat_c', latsCenter),
                            lon_c=up('lon_c', lonsCenter))
        if elev is not None:
            self._planes['elev_c'] = up('elev_c', elev)
        else:
            import torch
            self._planes['elev_c'] = ctx.zeros(h * w, torch.float64)
        ctx.valid_bits(w, h, self._planes)
        if imgMask is not None and imgMask.any():
            # reference mapping.py:1077-1081: the image mask is applied to the centre coordinates
            ctx.apply_center_mask(w, h, self._planes, ctx.to_device(imgMask.astype(np.uint8).ravel()))
        if sanitize:
            ctx.sanitize(w, h, self._planes)

    def _download(self, name):
        if name not in self._host:
            masked = BaseMapping._download(self, name)
            raw = self._raw.get(name)
            if raw is not None:
                data = self.context.to_numpy(raw) if hasattr(raw, 'data_ptr') else raw
                data = np.asarray(data, dtype=np.float64).reshape(masked.shape)
                self._host[name] = ma.masked_array(data, mask=ma.getmaskarray(masked) | np.isnan(data))
        return self._host[name]

    shape = property(lambda self: self._shape)
    img_unmasked = property(lambda self: self._imgData)

    def _computePlanes(self, ctx, names):
        # generic geodetic -> MLat/MLT route (reference mapping.py:540-550)
        if any(n.startswith('ml') for n in names):
            m = transform.mat_geo_to_sm(transform.date2es(self.photoTime))
            for suffix in ('k', 'c'):
                mlat, mlt = ctx.latlon_to_mlatmlt(self._planes['lat_' + suffix], self._planes['lon_' + suffix],
                                                  self.altitude, wgs84A, wgs84B, m)
                self._planes['mlat_' + suffix], self._planes['mlt_' + suffix] = mlat, mlt
        unknown = [n for n in names if n not in self._planes]
        if unknown:
            raise KeyError(unknown)

    @staticmethod
    def fromMapping(mapping):
        p = mapping.devicePlanes()
        h, w = mapping.shape
        return GenericMapping(p['lat_k'].clone().reshape(h + 1, w + 1), p['lon_k'].clone().reshape(h + 1, w + 1),
                              p['lat_c'].clone().reshape(h, w), p['lon_c'].clone().reshape(h, w),
                              p['elev_c'].clone().reshape(h, w), mapping.altitude, mapping.img_unmasked,
                              mapping.cameraPosGCRS, mapping.photoTime, mapping.identifier, mapping.metadata,
                              device=mapping._device, sanitize=False)


class MappingCollection(object):
    """Mappings of (almost) the same photo time, e.g. a network of all-sky cameras
    (reference mapping.py:1315-1373)."""

    def __init__(self, mappings, identifier, mayOverlap=True):
        self._mappings = mappings
        self._identifier = identifier
        self._mayOverlap = mayOverlap

    identifier = property(lambda self: self._identifier)
    mappings = property(lambda self: self._mappings)
    mayOverlap = property(lambda self: self._mayOverlap)
    empty = property(lambda self: len(self._mappings) == 0)

    def maskedByElevation(self, minElevation=10):
        return MappingCollection([m.maskedByElevation(minElevation) for m in self.mappings],
                                 self.identifier, self.mayOverlap)

    @property
    def boundingBox(self):
        return BoundingBox.mergedBoundingBoxes([m.boundingBox for m in self.mappings])

    @property
    def photoTime(self):
        times = sorted(m.photoTime for m in self.mappings)
        return times[len(times) // 2]

    def __len__(self):
        return len(self._mappings)


class _SMMapping(GenericMapping):
    @property
    def cameraFootpoint(self):
        et = transform.date2es(self.photoTime)
        s = transform.mat_j2000_to_sm(et).dot(np.asarray(self.cameraPosGCRS, dtype=np.float64))
        mlat = np.rad2deg(np.arctan2(s[2], np.sqrt(s[0] * s[0] + s[1] * s[1])))
        mlt = transform.smLonToMLT(np.rad2deg(np.arctan2(s[1], s[0])))
        return Location(mlat, transform.mltToSmLon(mlt))


class BaseMappingProvider(object):
    """Base class of mapping providers (reference mapping.py:1376-1445): `range`, `contains`,
    `get`, `getById`, `getSequence` are provided by subclasses."""

    def __init__(self, maxTimeOffset):
        self.maxTimeOffset = maxTimeOffset   # seconds

    @property
    def range(self):
        raise NotImplementedError

    def contains(self, date):
        raise NotImplementedError

    def containsAny(self, dates):
        return any(self.contains(date) for date in dates)

    def get(self, date):
        raise NotImplementedError

    def getById(self, identifier):
        raise NotImplementedError

    def getSequence(self, dateBegin=None, dateEnd=None):
        raise NotImplementedError


def _wrapProvider(provider, fn, name):
    """Copy of `provider` whose get / getById / getSequence results pass through `fn`."""
    base = type(provider)

    class Wrapped(base):
        def get(self, *a, **k):
            return fn(super(Wrapped, self).get(*a, **k))

        def getById(self, *a, **k):
            return fn(super(Wrapped, self).getById(*a, **k))

        def getSequence(self, *a, **k):
            return map(fn, super(Wrapped, self).getSequence(*a, **k))

    Wrapped.__name__ = '%s_extended_with_%s' % (base.__name__, name)
    wrapped = copy.copy(provider)
    wrapped.__class__ = Wrapped
    return wrapped


def MaskByElevationProvider(provider, *args, **kw):
    """Wrap a mapping provider so that every returned mapping is masked by elevation
    (reference mapping.py:1447-1472); parameters as in `BaseMapping.maskedByElevation`."""
    return _wrapProvider(provider, lambda m: m.maskedByElevation(*args, **kw), 'MaskingProvider')


def convertMappingToSM(mapping):
    """Mapping whose "lat/lon" are solar-magnetic latitude / longitude (reference
    mapping.py:1519-1547); used by `resampleMLatMLT`."""
    p = mapping.devicePlanes(magnetic=True)
    h, w = mapping.shape

    def smlon(mlt):
        return transform.mltToSmLon(mlt.clone())

    return _SMMapping(p['mlat_k'].clone().reshape(h + 1, w + 1), smlon(p['mlt_k']).reshape(h + 1, w + 1),
                      p['mlat_c'].clone().reshape(h, w), smlon(p['mlt_c']).reshape(h, w),
                      p['elev_c'].clone().reshape(h, w), mapping.altitude, mapping.img_unmasked,
                      mapping.cameraPosGCRS, mapping.photoTime, mapping.identifier, device=mapping._device,
                      sanitize=False)


def convertSMMappingToGeo(mapping):
    """Inverse of `convertMappingToSM` (reference mapping.py:1549-1559): the SM "lat/lon" planes
    of a (resampled) SM mapping back to geodetic coordinates, on the device (`amt_sm_to_latlon`)."""
    ctx = mapping.context
    p = mapping.devicePlanes()
    h, w = mapping.shape
    m = transform.mat_geo_to_sm(transform.date2es(mapping.photoTime))
    # the reference converts the `.data` of the SM grids (defined everywhere on a resampled grid)
    raw = getattr(mapping, '_raw', {})
    out = {}
    for suffix in ('k', 'c'):
        la = raw.get('lat_' + suffix)
        lo = raw.get('lon_' + suffix)
        la = ctx.to_device(np.asarray(la, dtype=np.float64).ravel()) if la is not None and not hasattr(la, 'data_ptr') \
            else (la.clone() if la is not None else p['lat_' + suffix].clone())
        lo = ctx.to_device(np.asarray(lo, dtype=np.float64).ravel()) if lo is not None and not hasattr(lo, 'data_ptr') \
            else (lo.clone() if lo is not None else p['lon_' + suffix].clone())
        ctx.sm_to_latlon(la, lo, m, wgs84A, wgs84B)
        out['lat_' + suffix], out['lon_' + suffix] = la, lo
    return GenericMapping(out['lat_k'].reshape(h + 1, w + 1), out['lon_k'].reshape(h + 1, w + 1),
                          out['lat_c'].reshape(h, w), out['lon_c'].reshape(h, w),
                          mapping.elevation, mapping.altitude, mapping.img, mapping.cameraPosGCRS,
                          mapping.photoTime, mapping.identifier, device=mapping._device)
