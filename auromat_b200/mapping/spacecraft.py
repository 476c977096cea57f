"""`getMapping` / `getMappingSequence`: the API entry of the georeferencing path.

API mirror of `auromat/mapping/spacecraft.py` (getMapping :380-426, _prepareMappingParams
:428-485, getMappingSequence :308-332, BaseSpacecraftMapping :487-555,
ArraySpacecraftMapping :583-595).  Camera positions must come from the header
(POSX/Y/Z or POS*SHIF + DATESHIF); the TLE route (pyephem + space-track download,
:454-483) is outside the hot path.
"""
from __future__ import annotations

import gc
import os
import warnings
from datetime import timedelta

import numpy as np
import numpy.ma as ma

from .. import fits
from .astrometry import BaseAstrometryMapping


def _prepareMappingParams(wcsPathOrHeader, timeshift=None, noradId=None, tleFolder=None, spacetrack=None):
    if isinstance(wcsPathOrHeader, str):
        header = fits.readHeader(wcsPathOrHeader)
    else:
        header = wcsPathOrHeader
    originalPhotoTime = fits.getPhotoTime(header)
    if originalPhotoTime is None:
        raise ValueError('DATE-OBS missing in FITS header')
    if timeshift is not None:
        photoTime = originalPhotoTime + timeshift
        cameraPosGCRS = None
    else:
        cameraPosGCRS, shifted, _ = fits.getShiftedSpacecraftPosition(header)
        if cameraPosGCRS is not None:
            photoTime = shifted
        else:
            photoTime = originalPhotoTime
            cameraPosGCRS, _ = fits.getSpacecraftPosition(header)
            if cameraPosGCRS is None:
                warnings.warn('Spacecraft position is missing in FITS header, will recalculate')
    if cameraPosGCRS is None:
        if tleFolder is None:
            raise ValueError('You need to specify tleFolder to calculate spacecraft positions')
        raise NotImplementedError('spacecraft positions from two-line elements (pyephem) are outside the '
                                  'B200 hot path; store POSX/POSY/POSZ in the header')
    return header, photoTime, originalPhotoTime, cameraPosGCRS


class BaseSpacecraftMapping(BaseAstrometryMapping):
    """Camera on a spacecraft seeing both Earth and stars; the star field gave the WCS."""

    def __init__(self, wcsHeader, alti, cameraPosGCRS, photoTime, identifier, metadata=None,
                 originalPhotoTime=None, fastCenterCalculation=False, device=None, sanitize=True):
        BaseAstrometryMapping.__init__(self, wcsHeader, alti, cameraPosGCRS, photoTime, identifier, metadata,
                                       fastCenterCalculation=fastCenterCalculation, device=device,
                                       sanitize=sanitize)
        self._originalPhotoTime = photoTime if originalPhotoTime is None else originalPhotoTime

    originalPhotoTime = property(lambda self: self._originalPhotoTime)

    @property
    def intersectsEarth(self):
        """Boolean (h,w) array: does the ray of a pixel centre hit the (un-inflated) WGS84 Earth?
        Reference spacecraft.py:508-521 -> intersection.py:165-199,229-237; here the hit ballots
        of the georeference kernel for altitude 0 (no coordinate plane is written)."""
        if 'intersectsEarth' not in self._host:
            from ..coordinates.wcs import frameConstants
            ctx = self.context
            h, w = self.shape
            fr = frameConstants(self.wcsHeader, self.cameraPosGCRS, self.photoTime, 0.0, False)
            _, bits = ctx.new_bitmaps(w, h)
            ctx.georef(fr, {'valid_c': bits})
            words = ctx.to_numpy(bits).view(np.uint32).reshape(h, (w + 31) // 32)
            unpacked = (words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1
            self._host['intersectsEarth'] = unpacked.reshape(h, -1)[:, :w].astype(bool)
        return self._host['intersectsEarth']

    def isConsistent(self, starPxCoords=None):
        """Plausibility of timestamp + astrometric solution (reference spacecraft.py:523-555)."""
        hit = self.intersectsEarth
        if np.all(hit) or not np.any(hit):
            return False
        if starPxCoords is not None:
            starPxCoords = np.asarray(starPxCoords)
            if np.any(hit[starPxCoords[:, 1], starPxCoords[:, 0]]):
                return False
        return True


class ArraySpacecraftMapping(BaseSpacecraftMapping):
    """Spacecraft mapping of an in-memory uint8/uint16 image array (h,w,n)."""

    def __init__(self, wcsHeader, alti, img, cameraPosGCRS, photoTime, identifier, metadata=None,
                 originalPhotoTime=None, fastCenterCalculation=False, device=None, sanitize=True):
        if hasattr(img, 'data_ptr'):                      # device tensor supplied by the caller
            self._imgDevice_in = img
            shape, dtype = tuple(img.shape), None
        else:
            if ma.isMaskedArray(img):
                img = img.data
            assert img.ndim == 3
            assert img.dtype in (np.uint8, np.uint16)
            self._imgDevice_in = None
            shape = img.shape
        assert (shape[0], shape[1]) == (int(wcsHeader['IMAGEH']), int(wcsHeader['IMAGEW'])), \
            'image shape does not match IMAGEW/IMAGEH'
        self._imgData = None if self._imgDevice_in is not None else img
        BaseSpacecraftMapping.__init__(self, wcsHeader, alti, cameraPosGCRS, photoTime, identifier, metadata,
                                       originalPhotoTime=originalPhotoTime,
                                       fastCenterCalculation=fastCenterCalculation, device=device,
                                       sanitize=sanitize)
        if self._imgDevice_in is not None:
            self._imgDevice = self._imgDevice_in

    @property
    def img_unmasked(self):
        if self._imgData is None:
            self._imgData = self.context.to_numpy(self._imgDevice)
        return self._imgData


class FileSpacecraftMapping(ArraySpacecraftMapping):
    """Spacecraft mapping of an image file (decoded on the host with PIL/OpenCV)."""

    def __init__(self, wcsHeader, alti, imagePath, cameraPosGCRS, photoTime, identifier, metadata=None,
                 originalPhotoTime=None, fastCenterCalculation=False, device=None, sanitize=True):
        self._imagePath = imagePath
        ArraySpacecraftMapping.__init__(self, wcsHeader, alti, fits.loadImage(imagePath), cameraPosGCRS,
                                        photoTime, identifier, metadata, originalPhotoTime,
                                        fastCenterCalculation, device, sanitize)

    imagePath = property(lambda self: self._imagePath)


def getMapping(imagePathOrArray, wcsPathOrHeader, timeshift=None, noradId=None, tleFolder=None, spacetrack=None,
               altitude=110, fastCenterCalculation=False, metadata=None, nosanitize=False, identifier=None,
               device=None):
    """Create the mapping of one image from its WCS solution, camera position and time
    (same signature as the reference's `getMapping`, plus `device`).  Nothing numeric runs
    until a coordinate property is read or `resample()` / `prefetch()` is called.

    :param imagePathOrArray: path, (h,w,n) uint8/uint16 array, or device tensor
    :param wcsPathOrHeader: path of a FITS `.wcs` file or any dict-like header
    :param datetime.timedelta timeshift: overrides the shifted timestamp of the header
    :param altitude: emission altitude in km
    :rtype: BaseSpacecraftMapping
    """
    header, photoTime, originalPhotoTime, cameraPosGCRS = \
        _prepareMappingParams(wcsPathOrHeader, timeshift, noradId, tleFolder, spacetrack)
    isImageArray = not isinstance(imagePathOrArray, str)
    isWcsHeader = not isinstance(wcsPathOrHeader, str)
    if identifier is None:
        if not isImageArray:
            identifier = os.path.splitext(os.path.basename(imagePathOrArray))[0]
        elif not isWcsHeader:
            identifier = os.path.splitext(os.path.basename(wcsPathOrHeader))[0]
    cls = ArraySpacecraftMapping if isImageArray else FileSpacecraftMapping
    return cls(header, altitude, imagePathOrArray, cameraPosGCRS, photoTime, identifier, metadata,
               originalPhotoTime=originalPhotoTime, fastCenterCalculation=fastCenterCalculation,
               device=device, sanitize=not nosanitize)


def getMappingSequence(imagePathsOrArrays, wcsPaths, metadatas=None, timeshift=None, noradId=None,
                       tleFolder=None, spacetrack=None, altitude=110, parallel=False,
                       fastCenterCalculation=False, device=None):
    """Generator of mappings for an image sequence (reference :308-332).  Frames are
    independent; `auromat_b200.parallel.shardSequence` splits them across GPUs."""
    if not metadatas:
        metadatas = [{}] * len(wcsPaths)

    def make(args):
        image, wcs, metadata = args
        m = getMapping(image, wcs, timeshift=timeshift, noradId=noradId, tleFolder=tleFolder,
                       spacetrack=spacetrack, altitude=altitude, fastCenterCalculation=fastCenterCalculation,
                       metadata=metadata, device=device)
        gc.collect()
        return m
    return map(make, zip(imagePathsOrArrays, wcsPaths, metadatas))
