"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference implementation (esa/auromat, mounted read-only at
/root/reference) inside this container, where several of its third-party dependencies are
absent (astropy, geographiclib, scikit-image, matplotlib, numexpr, distutils) and where
numpy>=2 removed `numpy.core.umath_tests`.  We register small stand-in modules in
`sys.modules` *before* importing `auromat.*` so that the reference's own numeric code
(`coordinates/wcs.py`, `coordinates/intersection.py`, `coordinates/transform.py`,
`mapping/mapping.py`, `mapping/astrometry.py`, `resample.py`, `util/histogram.py`) executes
as shipped.  numexpr stays absent, hence the reference's `_np` branches run.

The only source-level patch is one token in `auromat/util/histogram.py:262`
(`hist[core]` -> `hist[tuple(core)]`; list-of-slices indexing was removed from numpy) and
`np.int` -> `int` (`mapping/mapping.py:712,812`).  Both patches are applied to an in-memory
copy of the source; nothing is written to /root/reference and no reference source is copied
into this repository.

This module is used by `oracle/gen_golden.py` (to generate `tests/golden/*.npz`) and by
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is not mounted, e.g. on
the GPU box).
"""
from __future__ import annotations

import importlib
import importlib.util
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("AUROMAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "auromat"))


# --------------------------------------------------------------------------- stand-ins
class _Unit:
    """Minimal astropy.units stand-in: value * unit -> quantity, .to(unit).value."""

    __array_ufunc__ = None      # make ndarray * unit defer to __rmul__

    def __init__(self, name, in_deg=None, in_km=None):
        self.name, self.in_deg, self.in_km = name, in_deg, in_km

    def __rmul__(self, value):
        return _Quantity(value, self)

    def __mul__(self, value):
        return _Quantity(value, self)


class _Quantity:
    def __init__(self, value, unit):
        self.value, self.unit = value, unit

    def to(self, unit):
        if self.unit.in_deg is not None and unit.in_deg is not None:
            # astropy converts arcsec->deg by multiplying with the scale 1/3600
            # (a single float multiply); we keep one rounding as well.
            if self.unit is unit:
                return _Quantity(self.value, unit)
            return _Quantity(self.value * (self.unit.in_deg / unit.in_deg), unit)
        if self.unit.in_km is not None and unit.in_km is not None:
            return _Quantity(self.value * (self.unit.in_km / unit.in_km), unit)
        raise ValueError("unsupported unit conversion")


_deg = _Unit("deg", in_deg=1.0)
_arcsec = _Unit("arcsec", in_deg=1.0 / 3600.0)
_rad = _Unit("rad", in_deg=180.0 / math.pi)
_km = _Unit("km", in_km=1.0)
_m = _Unit("m", in_km=1e-3)


class _Angle:
    """astropy.coordinates.Angle stand-in: Angle(q).wrap_at(w).degree.  Like astropy, the
    value is kept in the unit it was given in and wrapped in that unit (`Angle._wrap_at`:
    wraps = (angle - (wrap - 360deg)) // 360deg; angle -= wraps*360deg; two rounding
    fix-ups); `.degree` converts at the end.  astropy==0.4, pinned by the reference, is
    absent here; this is the algorithm of current astropy."""

    def __init__(self, q):
        self._v = np.asarray(q.value, dtype=np.float64)
        self._unit = q.unit

    def wrap_at(self, wrap):
        to_native = wrap.unit.in_deg / self._unit.in_deg
        w = wrap.value * to_native
        a360 = 360.0 * (1.0 / self._unit.in_deg) if self._unit is not _deg else 360.0
        if self._unit is _rad:
            a360 = 360.0 * (math.pi / 180.0)
        a = np.array(self._v, dtype=np.float64, copy=True, ndmin=1)
        floor_ = w - a360
        with np.errstate(invalid='ignore'):
            wraps = (a - floor_) // a360
        valid = np.isfinite(wraps) & (wraps != 0)
        if np.any(valid):
            a -= wraps * a360
            a[a >= w] -= a360
            a[a < floor_] += a360
        out = _Angle.__new__(_Angle)
        out._v = a.reshape(np.shape(self._v))
        out._unit = self._unit
        return out

    @property
    def degree(self):
        d = self._v if self._unit is _deg else self._v * self._unit.in_deg
        return d if np.ndim(d) else float(d)


class _Time:
    """astropy.time.Time(datetime, scale='utc').jd stand-in (two-part JD, summed once)."""

    def __init__(self, date, scale="utc"):
        self._date = date

    @property
    def jd(self):
        return datetime_to_jd(self._date)


def datetime_to_jd(d) -> float:
    """JD = (JD of 0h, an exact half-integer) + day fraction; one final rounding.

    astropy keeps (jd1, jd2) and returns jd1 + jd2; for a UTC datetime jd1 is the
    integer-ish part and jd2 the fraction, so the result carries a single rounding of the
    sum.  Leap-second days are avoided in all synthetic inputs.
    """
    y, m = d.year, d.month
    a = (14 - m) // 12
    yy = y + 4800 - a
    mm = m + 12 * a - 3
    jdn = d.day + (153 * mm + 2) // 5 + 365 * yy + yy // 4 - yy // 100 + yy // 400 - 32045
    sec = d.hour * 3600 + d.minute * 60 + d.second + d.microsecond / 1e6
    return (jdn - 0.5) + sec / 86400.0


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_shims():
    # numpy.core.umath_tests (removed in numpy 2)
    def matrix_multiply(a, b):
        return np.matmul(a, b)

    def inner1d(a, b):
        return np.einsum("...i,...i", a, b)

    _mod("numpy.core.umath_tests", matrix_multiply=matrix_multiply, inner1d=inner1d)

    # distutils.version.LooseVersion (removed in py3.12)
    if "distutils" not in sys.modules:
        try:
            import distutils.version  # noqa: F401
        except Exception:
            class LooseVersion:
                def __init__(self, v):
                    self.v = tuple(int(x) if x.isdigit() else x for x in str(v).replace("-", ".").split("."))

                def _cmp(self, other):
                    o = other.v if isinstance(other, LooseVersion) else LooseVersion(other).v
                    return (self.v > o) - (self.v < o)

                def __lt__(self, o): return self._cmp(o) < 0
                def __le__(self, o): return self._cmp(o) <= 0
                def __gt__(self, o): return self._cmp(o) > 0
                def __ge__(self, o): return self._cmp(o) >= 0
                def __eq__(self, o): return self._cmp(o) == 0

            d = _mod("distutils")
            d.version = _mod("distutils.version", LooseVersion=LooseVersion)

    # astropy
    ap = _mod("astropy", __version__="1.0")
    ap.units = _mod("astropy.units", deg=_deg, degree=_deg, arcsec=_arcsec, rad=_rad, km=_km, m=_m)
    ap.time = _mod("astropy.time", Time=_Time)
    ang = _mod("astropy.coordinates.angles", Angle=_Angle)
    ap.coordinates = _mod("astropy.coordinates", Angle=_Angle, angles=ang)

    class _REarth:
        def to(self, unit):
            return _Quantity(6378.1366, unit)

    ap.constants = _mod("astropy.constants", R_earth=_REarth())

    class _WCS:
        def __init__(self, *a, **k):
            raise NotImplementedError("astropy.wcs is not available in this container")

    wcsmod = _mod("astropy.wcs.wcs", WCS=_WCS)
    ap.wcs = _mod("astropy.wcs", wcs=wcsmod, WCS=_WCS)
    ap.io = _mod("astropy.io")
    ap.io.fits = _mod("astropy.io.fits")

    # geographiclib: constants only; Geodesic stubbed
    class Constants:
        WGS84_a = 6378137.0
        WGS84_f = 1 / 298.257223563

    class _Geod:
        EMPTY = 0
        DISTANCE = 1
        AZIMUTH = 2
        LATITUDE = 4
        LONGITUDE = 8

        class WGS84:
            @staticmethod
            def Inverse(*a, **k):
                raise NotImplementedError("geographiclib is not available in this container")

    _mod("geographiclib")
    _mod("geographiclib.constants", Constants=Constants)
    _mod("geographiclib.geodesic", Geodesic=_Geod)

    # skimage / matplotlib / exifread stubs
    # spacepy.pycdf (CDF reader of the THEMIS provider; only the array functions are exercised)
    sp = _mod("spacepy")
    sp.pycdf = _mod("spacepy.pycdf")
    sk = _mod("skimage")
    sk.measure = _mod("skimage.measure")
    sk.io = _mod("skimage.io")
    sk.color = _mod("skimage.color")
    sk.transform = _mod("skimage.transform")
    mpl = _mod("matplotlib", __version__="3.0.0", use=lambda *a, **k: None)
    mpl.path = _mod("matplotlib.path")
    _mod("exifread")


_loaded = None


def load_reference():
    """Return a namespace with the reference's own modules, imported under the shims."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # auromat.util.histogram needs the one-token patch -> load from patched source text.
    import auromat  # noqa: F401  (package __init__ only sets up matplotlib)
    import auromat.util  # noqa: F401

    def _load_patched(modname, relpath, replacements):
        path = os.path.join(REFERENCE_ROOT, relpath)
        with open(path, "r", encoding="utf-8-sig") as fh:
            src = fh.read()
        for old, new in replacements:
            assert old in src, (relpath, old)
            src = src.replace(old, new)
        spec = importlib.util.spec_from_loader(modname, loader=None, origin=path)
        mod = importlib.util.module_from_spec(spec)
        mod.__file__ = path
        sys.modules[modname] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        return mod

    hist = _load_patched("auromat.util.histogram", "auromat/util/histogram.py",
                         [("hist = hist[core]", "hist = hist[tuple(core)]")])
    auromat.util.histogram = hist

    import auromat.coordinates.igrf as igrf
    import auromat.coordinates.transformations as transformations
    import auromat.coordinates.geodesic as geodesic
    import auromat.coordinates.transform as transform
    import auromat.coordinates.intersection as intersection
    import auromat.coordinates.wcs as wcs
    import auromat.mapping  # noqa: F401
    mapping = _load_patched("auromat.mapping.mapping", "auromat/mapping/mapping.py",
                            [(".astype(np.int)", ".astype(int)")])
    auromat.mapping.mapping = mapping
    import auromat.mapping.astrometry as astrometry
    import auromat.resample as resample
    try:
        import auromat.mapping.themis as themis
    except Exception as exc:             # the array functions are optional for the other pins
        themis = None
        print("reference themis module not importable:", exc)

    ns = types.SimpleNamespace(
        igrf=igrf, transformations=transformations, geodesic=geodesic, transform=transform,
        intersection=intersection, wcs=wcs, mapping=mapping, astrometry=astrometry,
        resample=resample, histogram=hist, themis=themis,
    )
    _loaded = ns
    return ns


if __name__ == "__main__":
    ref = load_reference()
    print("reference modules loaded:", sorted(vars(ref)))
