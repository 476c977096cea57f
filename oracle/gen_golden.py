"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (esa/auromat, imported from
/root/reference through oracle/ref_shim.py) on seeded synthetic inputs.  Run inside the build
container only (the reference tree does not exist on the GPU box):

    python oracle/gen_golden.py

The arrays stored are the reference's own outputs; tests compare the oracle restatement
(bit for bit) and the CUDA path (to the stated tolerances) against them.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from auromat_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def reference_frame(ref, hdr, cam, t, fast):
    """lat/lon/mlat/mlt/elevation through the reference's own functions
    (mapping/astrometry.py:49-212 call order)."""
    H, W = hdr['IMAGEH'], hdr['IMAGEW']
    import auromat.utils as U
    A = ref.astrometry.BaseAstrometryMapping
    with quiet():
        dk = ref.astrometry.pixelDirection(hdr, True)
        pk = ref.mapping.inflatedEarthIntersection(dk.reshape(-1, 3), cam, 110).reshape(dk.shape)
        if fast:
            dc, pc = A._calcCenters(dk), A._calcCenters(pk)
        else:
            dc = ref.astrometry.pixelDirection(hdr, False)
            pc = ref.mapping.inflatedEarthIntersection(dc.reshape(-1, 3), cam, 110).reshape(dc.shape)
        out = {}
        la, lo = ref.transform.j2000ToLatLon(pk.reshape(-1, 3), t)
        out['lats'], out['lons'] = la.reshape(H + 1, W + 1), lo.reshape(H + 1, W + 1)
        la, lo = ref.transform.j2000ToLatLon(pc.reshape(-1, 3), t)
        out['latsCenter'], out['lonsCenter'] = la.reshape(H, W), lo.reshape(H, W)
        a, b = ref.transform.j2000ToMLatMLT(pk.reshape(-1, 3), t)
        out['mlat'], out['mlt'] = a.reshape(H + 1, W + 1), b.reshape(H + 1, W + 1)
        a, b = ref.transform.j2000ToMLatMLT(pc.reshape(-1, 3), t)
        out['mlatCenter'], out['mltCenter'] = a.reshape(H, W), b.reshape(H, W)
        with np.errstate(invalid='ignore'):
            iu = U.unitVectors(pc.reshape(-1, 3))
            al = U.angleBetween((-dc).reshape(-1, 3), iu).reshape(H, W)
        np.rad2deg(al, al)
        np.subtract(90, al, al)
        out['elevation'] = al
    return out


def load_miracle():
    """Import the reference's mapping/miracle.py with ONE in-memory patch:
    `np.indices(...)` -> `.astype(float)`, because `ind += 0.5` on the integer index array
    (miracle.py:328-332) raises on numpy >= 1.10 (and silently truncated before that)."""
    import importlib.util
    path = os.path.join(ref_shim.REFERENCE_ROOT, 'auromat/mapping/miracle.py')
    src = open(path).read()
    old = 'ind = np.indices((w_, w_))'
    assert old in src
    src = src.replace(old, old + '.astype(float)')
    spec = importlib.util.spec_from_loader('auromat.mapping.miracle', loader=None, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    sys.modules['auromat.mapping.miracle'] = mod
    exec(compile(src, path, 'exec'), mod.__dict__)
    return mod


def reference_allsky(ref, mir, cal, width, t, altitude=110):
    """lat/lon (corners, centres) and camera elevation through the reference's MIRACLEMapping
    methods (miracle.py:196-258,314-347), bypassing its file-based constructor."""
    class _M(mir.MIRACLEMapping.__mro__[1]):
        def createMasked(self, m):
            pass
    m = object.__new__(_M)
    m._calData, m._simple, m._img, m._altitude, m._photoTime = cal, False, None, altitude, t
    m._img_unmasked = np.zeros((width, width, 3), np.uint8)
    m.cameraPosGEO = list(ref.transform.geodetic2EcefZero(np.deg2rad(cal.lat), np.deg2rad(cal.lon)))
    with quiet():
        lats, lons = m._calculateLatsLons(center=False)
        latc, lonc = m._calculateLatsLons(center=True)
        _, el = m.calculateAzEl(center=True)
    return dict(lats=lats, lons=lons, latsCenter=latc, lonsCenter=lonc, elevation=np.asarray(el))


def allsky_golden(ref):
    import datetime
    mir = load_miracle()
    # the SOD row of the reference's test/resources/cal.txt
    cal = mir.CalibrationData('SOD', 2011.5, 2012.5, 67.42, 26.39, 219.3, 244.2, 155.81, 0.14373, None)
    g = reference_allsky(ref, mir, cal, 96, datetime.datetime(2012, 3, 4, 17, 19, 0))
    np.savez_compressed(os.path.join(OUT, "allsky_SOD_96.npz"),
                        cal=np.array([cal.lat, cal.lon, cal.xc, cal.yc, cal.k, cal.rotation]), **g)


def themis_golden(ref):
    """mapping/themis.py:224-253 `reproject` (run unmodified) on the SOD golden's corner
    coordinates: 110 km -> 90 km and 150 km, as a THEMIS L2 calibration would hold them."""
    g = np.load(os.path.join(OUT, "allsky_SOD_96.npz"))
    asi = (float(g['cal'][0]), float(g['cal'][1]))
    lats, lons = g['lats'][::3, ::3], g['lons'][::3, ::3]
    out = dict(asi=np.array(asi), lats110=lats, lons110=lons)
    for hNew in (90.0, 150.0):
        with quiet():
            la, lo = ref.themis.reproject(asi, lats, lons, 110.0, hNew)
        out['lats%d' % hNew], out['lons%d' % hNew] = la, lo
    np.savez_compressed(os.path.join(OUT, "themis_reproject.npz"), **out)


def main():
    ref = ref_shim.load_reference()
    os.makedirs(OUT, exist_ok=True)
    W, H = 133, 89
    hdr = synthetic.issHeader(W, H)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    img = synthetic.issImage(W, H)
    for fast in (False, True):
        g = reference_frame(ref, hdr, cam, t, fast)
        # reference resampling of the reference coordinates (resample.py:159-279)
        valid = ~np.isnan(g['lats'])
        latMin, latMax = np.nanmin(g['lats']), np.nanmax(g['lats'])
        lonMin, lonMax = np.nanmin(g['lons']), np.nanmax(g['lons'])
        BB = ref.mapping.BoundingBox(latMin, lonMin, latMax, lonMax)
        cm = np.isnan(g['latsCenter'])
        imgf = img.astype(np.float64)
        imgf[cm] = np.nan
        merged = np.dstack((imgf, g['elevation']))
        with quiet():
            r = ref.resample._resample(g['latsCenter'], g['lonsCenter'], 110, merged, None, BB, (9.0, 5.0),
                                       False, False, 'mean')
        np.savez_compressed(os.path.join(OUT, "iss_frame_%dx%d_fast%d.npz" % (W, H, int(fast))),
                            bbox=np.array([latMin, lonMin, latMax, lonMax]), px_per_deg=np.array([9.0, 5.0]),
                            rs_lats=r[0], rs_lons=r[1], rs_latsCenter=r[2], rs_lonsCenter=r[3], rs_data=r[4], **g)
    # frame matrices for a few dates (transform.py:683-696), given the ephemeris second
    ets = np.array([380755615.06, 3.0e8, -1.2e8, 5.5e8])
    mats = {}
    for i, et in enumerate(ets):
        mats['geo%d' % i] = ref.transform.mat_j2000_to_geo(et)
        mats['sm%d' % i] = ref.transform.mat_j2000_to_sm(et)
        mats['geosm%d' % i] = ref.transform.mat_geo_to_sm(et)
    np.savez_compressed(os.path.join(OUT, "frame_matrices.npz"), ets=ets, **mats)
    # rotatePole + discontinuity/pole resampling on the reference's own test coordinates
    # (test/resample_test.py:24-68)
    rng = np.random.default_rng(7)
    n = 4000
    lat = rng.uniform(60, 89.9, n)
    lon = rng.uniform(-180, 180, n)
    with quiet():
        rl, ro = ref.transform.rotatePole(np.deg2rad(lat), np.deg2rad(lon), 110, angle=90, axis=[1, 0, 0])
    np.savez_compressed(os.path.join(OUT, "rotate_pole.npz"), lat=lat, lon=lon, rlat=np.rad2deg(rl), rlon=np.rad2deg(ro))
    # histogram2d with weights (util/histogram.py) incl. samples exactly on edges
    x = rng.uniform(-3, 13, 20000)
    y = rng.uniform(40, 61, 20000)
    ex = np.linspace(0.3, 10.1, 50)
    x[:49] = ex[:49]
    x[49:60] = 10.1
    y[60:70] = 60.0
    w1 = rng.integers(0, 256, 20000).astype(np.float64)
    hs, _, _ = ref.histogram.histogram2d(x, y, bins=(49, 80), range=[[0.3, 10.1], [41.0, 60.0]], weights=[None, w1])
    np.savez_compressed(os.path.join(OUT, "histogram2d.npz"), x=x, y=y, w=w1, count=hs[0], wsum=hs[1],
                        bins=np.array([49, 80]), range=np.array([[0.3, 10.1], [41.0, 60.0]]))
    allsky_golden(ref)
    themis_golden(ref)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
