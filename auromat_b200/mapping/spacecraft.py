"""`getMapping` / `getMappingSequence`: the API entry of the georeferencing path.

API mirror of `auromat/mapping/spacecraft.py` (getMapping :380-426, _prepareMappingParams
:428-485, getMappingSequence :308-332, BaseSpacecraftMapping :487-555,
ArraySpacecraftMapping :583-595, SpacecraftMappingProvider :40-218,
SpacecraftMappingPathProvider :220-292).  Camera positions must come from the header
(POSX/Y/Z or POS*SHIF + DATESHIF); the TLE route (pyephem + space-track download,
:454-483) is outside the hot path.
"""
from __future__ import annotations

import gc
import json
import os
import warnings
from datetime import datetime

import numpy as np
import numpy.ma as ma

from .. import fits
from .astrometry import BaseAstrometryMapping
from .mapping import BaseMappingProvider


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


def _parseDates(dic):
    """JSON object hook: ISO date strings under 'date' keys become datetimes (reference :599-607)."""
    for k in ('date', 'date_obs'):
        v = dic.get(k)
        if isinstance(v, str):
            for fmt in ('%Y-%m-%dT%H:%M:%S.%f', '%Y-%m-%dT%H:%M:%S', '%Y-%m-%d %H:%M:%S.%f', '%Y-%m-%d %H:%M:%S'):
                try:
                    dic[k] = datetime.strptime(v, fmt)
                    break
                except ValueError:
                    pass
    return dic


def _loadMetadata(path):
    if path and os.path.exists(path):
        with open(path, 'r') as fp:
            return json.load(fp, object_hook=_parseDates)
    return None


def _frameMetadata(metadata, identifier):
    if not metadata:
        return None
    return dict(list(metadata['sequence_metadata'].items()) + list(metadata['image_metadata'][identifier].items()))


class SpacecraftMappingProvider(BaseMappingProvider):
    """Mappings of a folder (or of parallel lists) of images and `.wcs` solutions, ordered by
    the (shifted) photo time of the headers (reference spacecraft.py:40-218).  `getSequence()`
    feeds `pipeline.resampleSequence` / `ResampleProvider` directly."""

    def __init__(self, imageSequenceFolder, wcsFolder=None, imageFileExtension=None, timeshift=None,
                 noradId=None, tleFolder=None, spacetrack=None, altitude=110, maxTimeOffset=3,
                 sequenceInParallel=False, fastCenterCalculation=False, device=None):
        BaseMappingProvider.__init__(self, maxTimeOffset=maxTimeOffset)
        if wcsFolder is None:
            assert not isinstance(imageSequenceFolder, list), \
                'The wcsFolder parameter is required if imageSequenceFolder is a list'
            wcsFolder = imageSequenceFolder
        self._lists = isinstance(imageSequenceFolder, list)
        if self._lists != isinstance(wcsFolder, list):
            raise ValueError('imageSequenceFolder and wcsFolder must be both path lists or folder paths')
        self._imageFileExtension = imageFileExtension
        if self._lists:
            self.imagePaths, self.wcsPaths = list(imageSequenceFolder), list(wcsFolder)
            self._imageFileExtension = os.path.splitext(self.imagePaths[0])[1][1:]
            self._index()
        else:
            self.imageSequenceFolder, self.wcsFolder = imageSequenceFolder, wcsFolder
            self.reload()
        self.timeshift, self.noradId, self.tleFolder, self.spacetrack = timeshift, noradId, tleFolder, spacetrack
        self.altitude, self.fastCenterCalculation, self.device = altitude, fastCenterCalculation, device
        self._sequenceInParallel = sequenceInParallel
        first = self.imagePaths[0] if self.imagePaths else os.path.join(str(imageSequenceFolder), 'x')
        self.metadata = _loadMetadata(os.path.join(os.path.dirname(first), 'metadata.json'))

    def __len__(self):
        return len(self.wcsPaths)

    def reload(self):
        """Refresh to the current disk state (folder mode only)."""
        self.wcsPaths = sorted(os.path.join(self.wcsFolder, f) for f in os.listdir(self.wcsFolder)
                               if f.endswith('.wcs'))
        try:
            ext = '.' + self.imageFileExtension
            self.imagePaths = sorted(os.path.join(self.imageSequenceFolder, f)
                                     for f in os.listdir(self.imageSequenceFolder) if f.endswith(ext))
        except ValueError:
            self.imagePaths, self.wcsPaths = [], []
        self._index()

    def _index(self):
        """Every solution needs its image; order everything by the header time."""
        ext = self._imageFileExtension
        have = {os.path.basename(p) for p in self.imagePaths}
        ids = [os.path.splitext(os.path.basename(p))[0] for p in self.wcsPaths]
        missing = [i for i in ids if '%s.%s' % (i, ext) not in have]
        assert not missing, 'wcs files without image: ' + str(missing)
        dated = sorted((fits.getShiftedPhotoTime(fits.readHeader(p)), p, i) for p, i in zip(self.wcsPaths, ids))
        self.dates = [d for d, _, _ in dated]
        self.wcsPaths = [p for _, p, _ in dated]
        self.ids = [i for _, _, i in dated]
        self._imageOf = {os.path.splitext(os.path.basename(p))[0]: p for p in self.imagePaths}

    @property
    def imageFileExtension(self):
        """e.g. 'jpg'; derived from the first solved image if not given."""
        if self._imageFileExtension is None:
            names = os.listdir(self.imageSequenceFolder)
            for w in sorted(f for f in os.listdir(self.wcsFolder) if f.endswith('.wcs')):
                base = os.path.splitext(w)[0]
                matches = [f for f in names if os.path.splitext(f)[0] == base and not f.endswith('.wcs')]
                if len(matches) == 1:
                    self._imageFileExtension = os.path.splitext(matches[0])[1][1:]
                    break
                if len(matches) > 1:
                    raise ValueError('Image file extension not given but multiple candidates exist: ' + str(matches))
            if self._imageFileExtension is None:
                raise ValueError('Image file extension could not be determined. Make sure that there exists at '
                                 'least one .wcs file and a corresponding image with the same filename base.')
        return self._imageFileExtension

    @property
    def range(self):
        return self.dates[0], self.dates[-1]

    @property
    def unsolvedIds(self):
        return sorted(i for i in self._imageOf if i not in self.ids)

    def _nearest(self, date):
        from ..utils import findNearest
        idx = findNearest(self.dates, date)
        return idx, abs(self.dates[idx] - date).total_seconds()

    def contains(self, date):
        return self._nearest(date)[1] <= self.maxTimeOffset

    def _kw(self):
        return dict(timeshift=self.timeshift, noradId=self.noradId, tleFolder=self.tleFolder,
                    spacetrack=self.spacetrack, altitude=self.altitude,
                    fastCenterCalculation=self.fastCenterCalculation, device=self.device)

    def get(self, date):
        idx, offset = self._nearest(date)
        if offset > self.maxTimeOffset:
            raise ValueError('No image found')
        identifier = self.ids[idx]
        return getMapping(self._imageOf[identifier], self.wcsPaths[idx],
                          metadata=_frameMetadata(self.metadata, identifier), **self._kw())

    def getById(self, identifier):
        matched = [i for i in self.ids if identifier in i]
        assert len(matched) == 1, 'Ambiguous identifier: ' + str(matched)
        return self.get(self.dates[self.ids.index(matched[0])])

    def getSequence(self, dateBegin=None, dateEnd=None):
        assert dateBegin is None and dateEnd is None, 'Date ranges not supported'
        metadatas = [_frameMetadata(self.metadata, i) for i in self.ids] if self.metadata else None
        return getMappingSequence([self._imageOf[i] for i in self.ids], self.wcsPaths, metadatas=metadatas,
                                  parallel=self._sequenceInParallel, **self._kw())


class SpacecraftMappingPathProvider(BaseMappingProvider):
    """Sequence-only provider over explicit image / wcs path lists (reference :220-292)."""

    def __init__(self, imagePaths, wcsPaths, metadataPath=None, timeshift=None, noradId=None, tleFolder=None,
                 spacetrack=None, altitude=110, maxTimeOffset=3, sequenceInParallel=False,
                 fastCenterCalculation=False, device=None):
        BaseMappingProvider.__init__(self, maxTimeOffset=maxTimeOffset)
        assert len(imagePaths) == len(wcsPaths)
        pairs = sorted(zip(wcsPaths, imagePaths), key=lambda wp: fits.getPhotoTime(fits.readHeader(wp[0])))
        self.wcsPaths = [w for w, _ in pairs]
        self.imagePaths = [i for _, i in pairs]
        self.timeshift, self.noradId, self.tleFolder, self.spacetrack = timeshift, noradId, tleFolder, spacetrack
        self.altitude, self.fastCenterCalculation, self.device = altitude, fastCenterCalculation, device
        self.sequenceInParallel = sequenceInParallel
        self.metadata = _loadMetadata(metadataPath)

    def __len__(self):
        return len(self.wcsPaths)

    imageFileExtension = property(lambda self: os.path.splitext(self.imagePaths[0])[1][1:])

    @property
    def range(self):
        times = [fits.getShiftedPhotoTime(fits.readHeader(p)) for p in (self.wcsPaths[0], self.wcsPaths[-1])]
        return times[0], times[1]

    def getSequence(self, dateBegin=None, dateEnd=None):
        assert dateBegin is None and dateEnd is None, 'Date ranges not supported'
        metadatas = None
        if self.metadata:
            keys = [os.path.splitext(os.path.basename(p))[0] for p in self.imagePaths]
            metadatas = [_frameMetadata(self.metadata, k) for k in keys]
        return getMappingSequence(self.imagePaths, self.wcsPaths, metadatas=metadatas, timeshift=self.timeshift,
                                  noradId=self.noradId, tleFolder=self.tleFolder, spacetrack=self.spacetrack,
                                  altitude=self.altitude, parallel=self.sequenceInParallel,
                                  fastCenterCalculation=self.fastCenterCalculation, device=self.device)
