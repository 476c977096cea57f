"""Minimal FITS header access for `.wcs` files (astrometry.net output) without astropy.

Mirrors the accessor names of `auromat/fits.py` (readHeader :29-31, getPhotoTime :365-379,
getSpacecraftPosition :393-405, getShiftedSpacecraftPosition :427-442, getNoradId :347-356).
A FITS header is a sequence of 80-byte ASCII cards `KEYWORD = value / comment` in 2880-byte
blocks, terminated by `END`.
"""
from __future__ import annotations

import functools
from datetime import datetime, timedelta

import numpy as np


def _parseValue(text):
    text = text.strip()
    if text.startswith("'"):
        end = 1
        out = []
        while end < len(text):                 # '' is an escaped quote
            if text[end] == "'":
                if end + 1 < len(text) and text[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(text[end])
            end += 1
        return ''.join(out).rstrip()
    value = text.split('/', 1)[0].strip()
    if value in ('T', 'F'):
        return value == 'T'
    try:
        return int(value)
    except ValueError:
        pass
    try:
        return float(value.replace('D', 'E'))
    except ValueError:
        return value


def readHeader(path):
    """Read the primary header of a FITS file into a plain dict (COMMENT/HISTORY dropped)."""
    header = {}
    with open(path, 'rb') as fh:
        while True:
            block = fh.read(2880)
            if len(block) < 2880:
                break
            done = False
            for i in range(0, 2880, 80):
                card = block[i:i + 80].decode('ascii', 'replace')
                key = card[:8].strip()
                if key == 'END':
                    done = True
                    break
                if card[8:10] != '= ' or key in ('COMMENT', 'HISTORY', ''):
                    continue
                header[key] = _parseValue(card[10:])
            if done:
                break
    return header


def getNoradId(header):
    noradId = header.get('NORADID')
    return None if noradId is None else int(noradId)


@functools.lru_cache(maxsize=4096)
def _parseDate(dateobs):
    try:
        return datetime.strptime(dateobs, '%Y-%m-%dT%H:%M:%S.%f')
    except ValueError:
        return datetime.strptime(dateobs, '%Y-%m-%dT%H:%M:%S')


def getPhotoTime(header):
    """DATE-OBS as datetime, or None."""
    dateobs = header.get('DATE-OBS')
    if dateobs is None:
        return None
    return _parseDate(dateobs)


def getShiftedPhotoTime(header):
    """DATE-OBS + DATESHIF if a time shift is stored, else DATE-OBS (reference fits.py:381-391)."""
    date = getPhotoTime(header)
    shift = header.get('DATESHIF')
    if date is None or shift is None:
        return date
    return date + timedelta(seconds=shift)


def getSpacecraftPosition(header):
    """([x,y,z] km in GCRS at DATE-OBS, date) or (None, None)."""
    date = getPhotoTime(header)
    x = header.get('POSX')
    if x is None or date is None:
        return None, None
    return np.array([x, header['POSY'], header['POSZ']], dtype=np.float64), date


def getShiftedSpacecraftPosition(header):
    """([x,y,z] km at the corrected photo time, corrected datetime, timedelta) or Nones."""
    date = getPhotoTime(header)
    shift = header.get('DATESHIF')
    x = header.get('POSXSHIF')
    if x is None or date is None or shift is None:
        return None, None, None
    delta = timedelta(seconds=shift)
    return np.array([x, header['POSYSHIF'], header['POSZSHIF']], dtype=np.float64), date + delta, delta


def loadImage(path):
    """Decode an 8/16-bit image file into an (h,w,3) array (grayscale is replicated)."""
    try:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise IOError('cannot read ' + path)
        if img.ndim == 3:
            img = img[:, :, ::-1]
    except ImportError:
        from PIL import Image
        img = np.asarray(Image.open(path))
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, 2)
    return np.ascontiguousarray(img[:, :, :3])
