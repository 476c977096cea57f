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
from .mapping import BaseMapping, CENTER_PLANES, CORNER_PLANES


class BaseAstrometryMapping(BaseMapping):
    """A mapping that derives its coordinates from the camera position and a WCS solution.

    :param fastCenterCalculation: centre coordinates come from the mean of the four corner
        intersection points (reference astrometry.py:24-40,154-160) instead of their own rays.
    """

    def __init__(self, wcsHeader, alti, cameraPosGCRS, photoTime, identifier, metadata=None,
                 fastCenterCalculation=False, device=None, sanitize=True):
        BaseMapping.__init__(self, alti, cameraPosGCRS, photoTime, identifier, metadata, device)
        self._wcsHeader = wcsHeader
        self.fastCenterCalculation = fastCenterCalculation
        self._sanitize = sanitize and not fastCenterCalculation
        self.isSanitized = bool(fastCenterCalculation)
        self._illConditioned = None
        self._frame = None

    wcsHeader = property(lambda self: self._wcsHeader)

    @property
    def shape(self):
        return int(self._wcsHeader['IMAGEH']), int(self._wcsHeader['IMAGEW'])

    @property
    def frameConstants(self):
        """The host-computed `amt_frame` block of this image (built once)."""
        if self._frame is None:
            self._frame = frameConstants(self._wcsHeader, self.cameraPosGCRS, self.photoTime, self.altitude,
                                         self.fastCenterCalculation)
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
        fresh = {n: ctx.empty(nk if n in CORNER_PLANES else nc, torch.float64)
                 for n in want if n not in self._planes}
        if self._illConditioned is None:
            self._illConditioned = ctx.new_stats()
            stats = self._illConditioned
        else:
            stats = None
        ctx.georef(self.frameConstants, fresh, stats)
        self._planes.update(fresh)
        if self._sanitize:
            ctx.sanitize(w, h, self._planes)
        self.isSanitized = True

    @property
    def illConditionedCount(self):
        """Number of rays whose intersection discriminant is so small (grazing the inflated
        ellipsoid) that 1 ulp of input noise moves the footprint by more than 1e-9 deg."""
        return int(self._deviceStats().n_ill_conditioned)


def pixelDirection(fitsWcsHeader, corner=True):
    """The reference exposes the (h[+1], w[+1], 3) direction array (astrometry.py:245-269).
    The fused kernel never materialises it (24 B/px that would only be read back once)."""
    raise NotImplementedError('directions are consumed inside the fused georeference kernel')
