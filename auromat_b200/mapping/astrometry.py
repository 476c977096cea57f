"""WCS-driven mapping: pixel -> celestial direction -> intersection with the inflated
ellipsoid -> geodetic / magnetic coordinates and elevation, evaluated by ONE fused CUDA
kernel per frame (`amt_georef`).

API mirror of `auromat/mapping/astrometry.py` (BaseAstrometryMapping :18-218): the reference
computes each of `lats/lons`, `latsCenter/lonsCenter`, `mLatMlt`, `mLatMltCenter`,
`elevation` as a separate chain of full-frame numpy passes; here they are planes written by
the same kernel launch.
"""
from __future__ import annotations

import numpy as np

from ..coordinates.wcs import frameConstants
from .mapping import BaseMapping, CORNER_PLANES


class BaseAstrometryMapping(BaseMapping):
    """A mapping that derives its coordinates from the camera position and a WCS solution.

    :param fastCenterCalculation: centre coordinates come from the mean of the four corner
        intersection points (reference astrometry.py:24-40,154-160) instead of their own rays.
    """

    earthModel = 'wgs84'         # or 'sphere' (reference mapping/mapping.py:1474-1510); set before first use

    def __init__(self, wcsHeader, alti, cameraPosGCRS, photoTime, identifier, metadata=None,
                 fastCenterCalculation=False, device=None, sanitize=True):
        BaseMapping.__init__(self, alti, cameraPosGCRS, photoTime, identifier, metadata, device)
        self._wcsHeader = wcsHeader
        self.fastCenterCalculation = fastCenterCalculation
        self._sanitize = sanitize and not fastCenterCalculation
        self.isSanitized = bool(fastCenterCalculation)
        self._frame = None

    wcsHeader = property(lambda self: self._wcsHeader)
    _finiteElevation = True      # computed by the georeference kernels: finite wherever the centre is defined

    @property
    def shape(self):
        return int(self._wcsHeader['IMAGEH']), int(self._wcsHeader['IMAGEW'])

    @property
    def frameConstants(self):
        """The host-computed `amt_frame` block of this image (built once)."""
        if self._frame is None:
            self._frame = frameConstants(self._wcsHeader, self.cameraPosGCRS, self.photoTime, self.altitude,
                                         self.fastCenterCalculation, self.earthModel)
        return self._frame

    def _computePlanes(self, ctx, names):
        import torch
        h, w = self.shape
        nk, nc = (h + 1) * (w + 1), h * w
        # One launch produces every plane that is asked for; lat planes are always needed
        # for the sanitisation stencil.
        want = set(names) | {'lat_k', 'lat_c'}
        if any(n.startswith('ml') for n in names):
            want |= {'mlat_k', 'mlt_k', 'mlat_c', 'mlt_c'}
        else:
            want |= {'lat_k', 'lon_k', 'lat_c', 'lon_c', 'elev_c'}
        pool = getattr(self, '_planeBuffers', None) or {}      # caller-provided ring buffers (pipeline)
        fresh = {n: pool[n] if n in pool else ctx.empty(nk if n in CORNER_PLANES else nc, torch.float64)
                 for n in want if n not in self._planes}
        # the hit bitmaps are (re)written by every launch; sanitisation then starts from them
        if 'valid_k' in pool:
            fresh['valid_k'], fresh['valid_c'] = pool['valid_k'], pool['valid_c']
        else:
            fresh['valid_k'], fresh['valid_c'] = ctx.new_bitmaps(w, h)
        # the first launch of a frame counts the grazing rays into the statistics block
        if self._statsDevice is None:
            self._statsDevice = ctx.new_stats()
        stats = None if self._grazingCounted else self._statsDevice
        self._grazingCounted = True
        # a ring slot comes with the prebuilt amt_georef_out of all its planes
        out = pool.get('_out') if len(fresh) == pool.get('_nplanes', -1) and not self._planes else None
        ctx.georef(self.frameConstants, fresh, stats, out=out)
        self._planes.update(fresh)
        hook = self.__dict__.pop('_afterGeoref', None)     # the sequence pipeline switches streams here
        if hook is not None:
            hook()
        if self._sanitize:
            ctx.sanitize(w, h, self._planes, out=out)
        self.isSanitized = True

    # ---- plane-free (fused) resampling: hit bitmaps + outline statistics only ----
    def _ensureHitBitmaps(self):
        """Validity bitmaps of this frame without any coordinate plane: ray hit ballots
        (direction + discriminant per ray) followed by the sanitisation stencils."""
        if 'valid_k' not in self._planes:
            ctx = self.context
            h, w = self.shape
            bits = {}
            bits['valid_k'], bits['valid_c'] = ctx.new_bitmaps(w, h)
            if self._statsDevice is None:
                self._statsDevice = ctx.new_stats()
            ctx.georef(self.frameConstants, bits, None if self._grazingCounted else self._statsDevice)
            self._grazingCounted = True
            hook = self.__dict__.pop('_afterGeoref', None)     # the sequence pipeline switches streams here
            if hook is not None:
                hook()
            if self._sanitize:
                ctx.sanitize(w, h, bits)
            self._planes.update(bits)
        return self._planes['valid_k'], self._planes['valid_c']

    def _startStats(self):
        if 'lat_k' in self._planes or self.fastCenterCalculation or not getattr(self, '_planeFree', False):
            return BaseMapping._startStats(self)
        if self._stats is None and self._statsPending is None:
            ctx = self.context
            vk, vc = self._ensureHitBitmaps()
            if self._statsDevice is None:
                self._statsDevice = ctx.new_stats()
            ctx.bbox_stats_frame(self.frameConstants, vk, vc, self._statsDevice)
            self._statsPending = ctx.start_stats_readback(self._statsDevice)

    def setPlaneFree(self, flag=True):
        """Let `boundingBox` / `resample` work from the hit bitmaps and the frame model alone
        (`amt_bbox_stats_frame`, `amt_georef_bin_fused`) as long as no coordinate array has
        been asked for.  Not available with fastCenterCalculation."""
        self._planeFree = bool(flag) and not self.fastCenterCalculation
        return self

    # ---- pole containment: exact geometric test instead of the reference's outline walk ----
    _poleTestOnDevice = False

    def _poleFlags(self):
        """bit0 / bit1: the geographic north / south pole (at the mapping altitude) is seen by
        a valid pixel.  The pole point is projected through the inverse WCS (`amt_pole_pixels`, host
        arithmetic in C shared with the sequence engine); replaces the azimuth-sum test on a 50-point
        convex outline (reference mapping.py:705-718)."""
        import ctypes
        from .. import _lib
        h, w = self.shape
        ix, iy, inFrame = (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)()
        _lib.check(_lib.load().amt_pole_pixels(ctypes.byref(self.frameConstants), ix, iy, inFrame))
        flags = 0
        for i, bit in ((0, 1), (1, 2)):
            if not inFrame[i]:
                continue
            if 'lat_c' in self._planes or not getattr(self, '_planeFree', False):
                lat = float(self.devicePlanes()['lat_c'][iy[i] * w + ix[i]].item())
                valid = lat == lat
            else:
                word = int(self._ensureHitBitmaps()[1][iy[i] * ((w + 31) // 32) + ix[i] // 32].item()) & 0xffffffff
                valid = bool((word >> (ix[i] % 32)) & 1)
            if valid:
                flags |= bit
        return flags

    def _invertSip(self, target):
        """Solve (u,v) + (f,g)(u,v) = target by fixed-point iteration (|distortion| << 1)."""
        fr = self.frameConstants

        def poly(packed, order, u, v):
            acc = 0.0
            for p in range(order, -1, -1):
                base = p * (order + 1) - (p * (p - 1)) // 2
                inner = 0.0
                for q in range(order - p, -1, -1):
                    inner = inner * v + packed[base + q]
                acc = acc * u + inner
            return acc
        u, v = target
        for _ in range(30):
            u = target[0] - poly(fr.sip_a, fr.sip_order_a, u, v)
            v = target[1] - poly(fr.sip_b, fr.sip_order_b, u, v)
        return np.array([u, v])

    @property
    def illConditionedCount(self):
        """Number of rays whose intersection discriminant is so small (grazing the inflated
        ellipsoid) that 1 ulp of input noise moves the footprint by more than 1e-9 deg."""
        return int(self._deviceStats().n_ill_conditioned)


def pixelDirection(fitsWcsHeader, corner=True):
    """The reference exposes the (h[+1], w[+1], 3) direction array (astrometry.py:245-269).
    The fused kernel never materialises it (24 B/px that would only be read back once)."""
    raise NotImplementedError('directions are consumed inside the fused georeference kernel')
