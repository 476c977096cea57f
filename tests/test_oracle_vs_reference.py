"""Pins the oracle to the reference itself: runs esa/auromat's own functions (imported from
/root/reference through oracle/ref_shim.py) and the oracle restatement on the same inputs
and requires BIT-FOR-BIT equality.  Skipped where the reference tree is not mounted."""
import contextlib
import datetime
import io

import numpy as np
import pytest

import oracle.auromat_oracle as O
from auromat_b200 import synthetic
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    with contextlib.redirect_stdout(io.StringIO()):
        return ref_shim.load_reference()


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def bit_equal(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("W,H", [(266, 177), (97, 61)])
def test_stage_functions_bit_for_bit(ref, W, H):
    hdr = synthetic.issHeader(W, H)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    with quiet():
        rd = ref.astrometry.pixelDirection(hdr, corner=True)
        rdc = ref.astrometry.pixelDirection(hdr, corner=False)
        rp = ref.mapping.inflatedEarthIntersection(rd.reshape(-1, 3), cam, 110)
        rl = ref.transform.j2000ToLatLon(rp, t)
        rm = ref.transform.j2000ToMLatMLT(rp, t)
    od, odc = O.pix2dir(hdr, True), O.pix2dir(hdr, False)
    assert bit_equal(rd, od) and bit_equal(rdc, odc)
    op = O.inflated_earth_intersection(od.reshape(-1, 3), cam, 110)
    assert bit_equal(rp, op)
    ol, om = O.j2000_to_latlon(op, t), O.j2000_to_mlat_mlt(op, t)
    assert bit_equal(rl[0], ol[0]) and bit_equal(rl[1], ol[1])
    assert bit_equal(rm[0], om[0]) and bit_equal(rm[1], om[1])
    assert 0.3 < np.mean(~np.isnan(op[:, 0])) < 0.9


def test_matrices_bit_for_bit(ref):
    for d in (datetime.datetime(2012, 1, 25, 9, 26, 55, 60000), datetime.datetime(2003, 7, 1, 23, 59, 1),
              datetime.datetime(2019, 12, 31, 0, 0, 0)):
        et = ref.transform.date2es(d)
        assert et == O.date2es(d)
        for n in ('mat_P', 'mat_T1', 'mat_T2', 'mat_T3', 'mat_T4', 'mat_j2000_to_geo', 'mat_j2000_to_sm',
                  'mat_geo_to_sm'):
            assert bit_equal(getattr(ref.transform, n)(et), getattr(O, n)(et)), n


@pytest.mark.parametrize("mode", ["plain", "discontinuity", "pole"])
def test_resample_bit_for_bit(ref, mode):
    rng = np.random.default_rng(3)
    h, w = 60, 80
    yy, xx = np.mgrid[0:h + 1, 0:w + 1].astype(float)
    if mode == "plain":
        lats, lons = 70 - yy * 0.11 - xx * 0.01, -100 + xx * 0.2 + yy * 0.02
    elif mode == "discontinuity":
        lats, lons = 70 - yy * 0.11, 170 + xx * 0.25
        lons = O.wrap_at_180(lons)
    else:
        r = 1 + np.hypot(yy - h / 2, xx - w / 2) * 0.12
        ang = np.arctan2(yy - h / 2, xx - w / 2)
        lats, lons = 90 - r, np.rad2deg(ang)
    latsC = (lats[:-1, :-1] + lats[1:, 1:]) / 2
    lonsC = lons[:-1, :-1] if mode != "plain" else (lons[:-1, :-1] + lons[1:, 1:]) / 2
    latsC[:5, :7] = np.nan
    lonsC = np.where(np.isnan(latsC), np.nan, lonsC)
    data = np.dstack((rng.integers(0, 256, (h, w, 3)).astype(float), rng.uniform(0, 90, (h, w))))
    data[np.isnan(latsC)] = np.nan
    outline = np.transpose([lats.ravel(), lons.ravel()])
    if mode == "pole":
        bbox = (float(np.min(lats)), -180.0, 90.0, 180.0)
    elif mode == "discontinuity":
        bbox = (float(np.min(lats)), float(np.min(lons[lons > 0])), float(np.max(lats)), float(np.max(lons[lons <= 0])))
    else:
        bbox = (float(np.min(lats)), float(np.min(lons)), float(np.max(lats)), float(np.max(lons)))
    BB = ref.mapping.BoundingBox(*bbox)
    with quiet():
        r = ref.resample._resample(latsC.copy(), lonsC.copy(), 110, data.copy(), lambda: outline.copy(), BB,
                                   (6.0, 3.0), mode == "discontinuity", mode == "pole", 'mean')
    o = O.resample_grid(latsC.copy(), lonsC.copy(), 110, data.copy(), bbox, (6.0, 3.0),
                        contains_discontinuity=mode == "discontinuity", contains_pole=mode == "pole",
                        outline_latlon=outline.copy())
    for a, b in zip(r, o):
        assert bit_equal(a, b)
    assert np.isfinite(o[4][:, :, 0]).sum() > 100


def test_sanitize_masks_match_reference(ref):
    import numpy.ma as ma
    rng = np.random.default_rng(5)
    h, w = 40, 50
    lat_k = rng.uniform(0, 1, (h + 1, w + 1))
    lat_k[rng.uniform(size=lat_k.shape) < 0.2] = np.nan
    lat_c = rng.uniform(0, 1, (h, w))
    lat_c[rng.uniform(size=lat_c.shape) < 0.3] = np.nan
    lats, lons = ma.masked_invalid(lat_k.copy()), ma.masked_invalid(lat_k.copy())
    latsC, lonsC = ma.masked_invalid(lat_c.copy()), ma.masked_invalid(lat_c.copy())
    img = ma.masked_array(np.zeros((h, w, 3), np.uint8), mask=np.zeros((h, w, 3), bool))
    elev = ma.masked_invalid(lat_c.copy())
    with quiet():
        ref.mapping._doSanitize(lats, lons, latsC, lonsC, img, elev)
    mk, mc = O.sanitize_masks(np.isnan(lat_k), np.isnan(lat_c))
    assert np.array_equal(ma.getmaskarray(lats), mk)
    assert np.array_equal(ma.getmaskarray(latsC), mc)
    assert np.array_equal(ma.getmaskarray(img)[:, :, 0], mc)


def test_allsky_model_bit_for_bit(ref):
    import datetime
    import oracle.gen_golden as G
    mir = G.load_miracle()
    cal = mir.CalibrationData('KEV', 2011.5, 2012.5, 69.76, 27.01, 249.5, 273.8, 154.59, 0.07049, None)
    w = 64
    g = G.reference_allsky(ref, mir, cal, w, datetime.datetime(2012, 3, 4, 17, 19, 0))
    s = w / 512
    o = O.allsky_georeference(w, cal.xc * s, cal.yc * s, cal.k * s, cal.rotation, cal.lat, cal.lon, 110)
    for name in ('lats', 'lons', 'latsCenter', 'lonsCenter', 'elevation'):
        assert bit_equal(g[name], o[name]), name


def test_polygon_helpers_vs_reference(ref):
    """utils.py polygonArea / polygonCentroid / withoutConsecutiveDuplicates / convexHull run
    unmodified (convexHull through scipy's Delaunay, which is installed here)."""
    import sys
    from auromat_b200 import utils as U
    R = sys.modules['auromat.utils']
    rng = np.random.default_rng(5)
    for n in (3, 7, 40):
        ang = (np.arange(n) + 0.8 * rng.random(n)) * (2 * np.pi / n)     # origin inside => counter-clockwise
        poly = np.transpose([np.cos(ang), np.sin(ang)]) * rng.uniform(1, 5, (n, 1)) + [55.0, -99.0]   # CCW, star-shaped
        assert R.polygonArea(poly) == O.polygon_area(poly)
        assert R.polygonCentroid(poly) == O.polygon_centroid(poly)
        np.testing.assert_allclose(U.polygonCentroid(poly), R.polygonCentroid(poly), rtol=0, atol=1e-9)
        np.testing.assert_allclose(U.polygonArea(poly), R.polygonArea(poly), rtol=1e-12)
        # the reference divides signed moments by the unsigned area: clockwise input is mirrored
        # through the origin (restated as is in the oracle; the product is orientation independent)
        assert R.polygonCentroid(poly[::-1]) == O.polygon_centroid(poly[::-1])
        np.testing.assert_allclose(R.polygonCentroid(poly[::-1]), -np.array(R.polygonCentroid(poly)), rtol=1e-9)
        np.testing.assert_allclose(U.polygonCentroid(poly[::-1]), R.polygonCentroid(poly), rtol=0, atol=1e-9)
    pts = rng.integers(0, 60, (300, 2))
    with quiet():
        rh = R.convexHull(pts)
    # scipy's Delaunay hull may keep points lying exactly on a hull edge; the corner set and its order agree
    mine = U.convexHull(pts)
    assert {tuple(p) for p in mine.tolist()} <= {tuple(p) for p in rh.tolist()}
    assert U.polygonArea(mine) == pytest.approx(R.polygonArea(rh))
    strict = [p for p in rh.tolist() if tuple(p) in {tuple(q) for q in mine.tolist()}]
    assert strict == mine.tolist()
    arr = np.array([[1, 2], [1, 2], [3, 4], [3, 4], [3, 4], [1, 2]])
    assert np.array_equal(R.withoutConsecutiveDuplicates(arr), U.withoutConsecutiveDuplicates(arr))


def test_pole_test_on_the_reference_regression_outlines():
    """test/geodesic_test.py:57-2171, 2173-2669: real ISS outlines (full, reduced, hulls) that
    once fooled the pole test; none of them contains a pole.  The arrays are parsed out of the
    reference's test file."""
    import re
    from auromat_b200 import utils as U
    from auromat_b200.coordinates.geodesic import containsOrCrossesPole
    src = open('/root/reference/auromat/test/geodesic_test.py').read()
    arrays = re.findall(r"\n\s+(outline\w+) = (?:np\.array\()?(\[\[.*?\]\])", src, re.S)
    assert len(arrays) >= 5
    checked = 0
    for name, body in arrays:
        pts = np.array(eval(body), dtype=float)
        assert pts.ndim == 2 and len(pts) >= 3
        if name == 'outlineReduced100':
            # a sub-sampled outline self-intersects (the bug); the reference tests its hulls (:2161-2170)
            assert not containsOrCrossesPole(U.convexHull(pts[::2])), name + '[::2] hull'
        else:
            assert not containsOrCrossesPole(pts), name
        assert not containsOrCrossesPole(U.convexHull(pts)), name + ' hull'
        checked += 1
    assert checked >= 5


def test_themis_reproject_bit_for_bit(ref):
    if ref.themis is None:
        pytest.skip("reference themis module not importable")
    rng = np.random.default_rng(0)
    asi = (62.4, -114.5)
    lat = asi[0] + rng.uniform(-4, 4, (40, 41))
    lon = asi[1] + rng.uniform(-8, 8, (40, 41))
    lat[0, 0] = lon[0, 0] = np.nan
    for hRef, hNew in ((90.0, 110.0), (150.0, 110.0), (110.0, 0.0)):
        with quiet():
            r = ref.themis.reproject(asi, lat, lon, hRef, hNew)
        o = O.themis_reproject(asi, lat, lon, hRef, hNew)
        assert bit_equal(r[0], o[0]) and bit_equal(r[1], o[1])
    with quiet():
        cam = ref.transform.latLonToJ2000(asi[0], asi[1], 0, datetime.datetime(2012, 2, 4, 7, 56, 26))
    from auromat_b200.coordinates import transform as T
    from auromat_b200.mapping.allsky import stationEcef
    mine = T.mat_j2000_to_geo(T.date2es(datetime.datetime(2012, 2, 4, 7, 56, 26))).T.dot(stationEcef(*asi))
    np.testing.assert_allclose(mine, np.ravel(cam), rtol=0, atol=1e-9)
