"""Oracle (numpy restatement) against the golden vectors produced by the reference itself
(oracle/gen_golden.py) and against the reference's own known-answer tests."""
import datetime
import os

import numpy as np
import pytest
from numpy.testing import assert_array_almost_equal, assert_array_equal, assert_equal

import oracle.auromat_oracle as O
from auromat_b200 import synthetic

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
# numpy's sin/cos/atan kernels may differ by an ulp between CPUs; the goldens were produced
# on the build container.  1e-11 deg is 100x below the 1e-9 deg parity requirement.
TOL_DEG = 1e-11


@pytest.mark.parametrize("fast", [0, 1])
def test_georeference_matches_reference_golden(fast):
    g = np.load(os.path.join(GOLDEN, "iss_frame_133x89_fast%d.npz" % fast))
    hdr = synthetic.issHeader(133, 89)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    o = O.georeference(hdr, cam, t, 110, fast_center=bool(fast))
    for name in ('lats', 'lons', 'latsCenter', 'lonsCenter', 'mlat', 'mlt', 'mlatCenter', 'mltCenter', 'elevation'):
        assert_array_equal(np.isnan(o[name]), np.isnan(g[name]), err_msg=name)
        assert np.nanmax(np.abs(o[name] - g[name])) <= TOL_DEG, name


@pytest.mark.parametrize("fast", [0, 1])
def test_resample_matches_reference_golden(fast):
    """Binning of the *golden* coordinates is pure integer/compare work -> bit exact."""
    g = np.load(os.path.join(GOLDEN, "iss_frame_133x89_fast%d.npz" % fast))
    img = synthetic.issImage(133, 89)
    cm = np.isnan(g['latsCenter'])
    imgf = img.astype(np.float64)
    imgf[cm] = np.nan
    merged = np.dstack((imgf, g['elevation']))
    r = O.resample_grid(g['latsCenter'], g['lonsCenter'], 110, merged, tuple(g['bbox']), tuple(g['px_per_deg']))
    for a, name in zip(r, ('rs_lats', 'rs_lons', 'rs_latsCenter', 'rs_lonsCenter', 'rs_data')):
        assert_array_equal(a, g[name], err_msg=name)


def test_frame_matrices_match_reference_golden():
    g = np.load(os.path.join(GOLDEN, "frame_matrices.npz"))
    for i, et in enumerate(g['ets']):
        assert np.max(np.abs(O.mat_j2000_to_geo(et) - g['geo%d' % i])) < 1e-15
        assert np.max(np.abs(O.mat_j2000_to_sm(et) - g['sm%d' % i])) < 1e-15
        assert np.max(np.abs(O.mat_geo_to_sm(et) - g['geosm%d' % i])) < 1e-15


def test_rotate_pole_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "rotate_pole.npz"))
    la, lo = O.rotate_pole(np.deg2rad(g['lat']), np.deg2rad(g['lon']), 110, angle=90, axis=[1, 0, 0])
    assert np.max(np.abs(np.rad2deg(la) - g['rlat'])) <= TOL_DEG
    d = np.abs(np.rad2deg(lo) - g['rlon'])
    assert np.max(np.minimum(d, 360 - d)) <= TOL_DEG


def test_histogram2d_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "histogram2d.npz"))
    h = O.histogram2d_weighted(g['x'], g['y'], tuple(g['bins']), g['range'].tolist(), [None, g['w']])
    assert_array_equal(h[0], g['count'])
    assert_array_equal(h[1], g['wsum'])
    ix, iy = O.cell_indices(g['x'], g['y'], tuple(g['bins']), g['range'].tolist())
    ok = (ix >= 0) & (iy >= 0)
    cnt = np.zeros(tuple(g['bins']))
    np.add.at(cnt, (ix[ok], iy[ok]), 1)
    assert_array_equal(cnt, g['count'])


# ---- the reference's own known-answer tests, re-expressed (test/intersection_test.py:26-137) ----
def test_kat_sphere_line_intersection():
    assert_equal(O.sphere_line_intersection(2, [0, 3, 0], [0, -1, 0]), [0, 2, 0])
    d = np.array([[0, -1, 0], [-1, -1, 0]], dtype=float)
    d /= np.sqrt((d * d).sum(axis=1))[:, None]
    assert_equal(O.sphere_line_intersection(2, [0, 3, 0], d), [[0, 2, 0], [np.nan] * 3])


def test_kat_ellipsoid_line_intersection():
    p1 = np.array(O.geodetic2ecef(np.deg2rad(30), np.deg2rad(60), 0))
    p2 = np.array(O.geodetic2ecef(np.deg2rad(-30), np.deg2rad(-60), 0))
    i1 = O.ellipsoid_line_intersection(O.WGS84_A, O.WGS84_B, p1, [p1 - p2], directed=False)
    assert_array_almost_equal(i1, [p1])
    pts = O.ellipsoid_line_intersection(2, 2, [0, 3, 0], [[0, -1, 0], [0, -1, 0], [-1, -1, 0]])
    assert_equal(pts, [[0, 2, 0], [0, 2, 0], [np.nan] * 3])
    assert_equal(O.ellipsoid_line_intersects(2, 2, [0, 3, 0], [[0, -1, 0], [0, -1, 0], [-1, -1, 0]]), [True, True, False])


def test_kat_directed_intersection():
    r, origin, direction = 1, [2, 0, 0], [[1, 0, 0]]
    hit, miss = [[1, 0, 0]], [[np.nan] * 3]
    assert_array_equal(O.sphere_line_intersection(r, origin, direction, directed=False), hit)
    assert_array_equal(O.sphere_line_intersection(r, origin, direction, directed=True), miss)
    assert_array_equal(O.ellipsoid_line_intersection(r, r, origin, direction, directed=False), hit)
    assert_equal(O.ellipsoid_line_intersects(r, r, origin, direction, directed=False), [True])
    assert_array_equal(O.ellipsoid_line_intersection(r, r, origin, direction, directed=True), miss)
    assert_equal(O.ellipsoid_line_intersects(r, r, origin, direction, directed=True), [False])
    assert_array_equal(O.sphere_line_intersection(r, [-2, 0, 0], direction, directed=True), [[-1, 0, 0]])
    assert_array_equal(O.sphere_line_intersection(r, [-2, 0, 0], [[-1, 0, 0]], directed=True), miss)


def test_kat_directed_intersection_from_inside():
    r, origin = 2, [1, 0, 0]
    assert_array_equal(O.ellipsoid_line_intersection(r, r, origin, [[1, 0, 0]], directed=False), [[2, 0, 0]])
    assert_array_equal(O.ellipsoid_line_intersection(r, r, origin, [[1, 0, 0]], directed=True), [[2, 0, 0]])
    assert_equal(O.ellipsoid_line_intersects(r, r, origin, [[1, 0, 0]], directed=True), [True])
    assert_array_equal(O.sphere_line_intersection(r, origin, [[-1, 0, 0]], directed=True), [[-2, 0, 0]])


# ---- test/transform_test.py:70-129 ----
def test_kat_geodetic_round_trip_11_decimals():
    lat, lon = np.mgrid[-89:89:5, -179:179:5].astype(float)
    x, y, z = O.geodetic2ecef_zero(np.deg2rad(lat), np.deg2rad(lon))
    r = O.ecef2geodetic(x.ravel(), y.ravel(), z.ravel())
    assert_array_almost_equal(np.rad2deg(r[0]).reshape(lat.shape), lat, 11)
    assert_array_almost_equal(np.rad2deg(r[1]).reshape(lon.shape), lon, 11)
    for la in np.linspace(-89.9, 89.9, 7):
        for lo in np.linspace(-179.9, 179.9, 7):
            x, y, z = O.geodetic2ecef_zero(np.deg2rad(la), np.deg2rad(lo))
            assert_array_almost_equal(np.rad2deg(O.ecef2geodetic_scalar(x, y, z)), [la, lo], 11)


def test_kat_sscweb_frames_2_decimals():
    date = datetime.datetime(2012, 1, 25, 9, 26, 55)
    et = O.date2es(date)
    geo, j2000 = np.array([[-0.11, -0.63, 0.77]]), np.array([[-0.62, 0.16, 0.77]])
    gse, gsm, sm = np.array([[-0.72, -0.26, 0.64]]), np.array([[-0.72, -0.30, 0.62]]), np.array([[-0.43, -0.30, 0.85]])
    assert_array_almost_equal(O._matvec(O.mat_T1(et), j2000), geo, 2)
    assert_array_almost_equal(O._matvec(O.mat_T2(et), j2000), gse, 2)
    assert_array_almost_equal(O._matvec(O.mat_T3(et), gse), gsm, 2)
    assert_array_almost_equal(O._matvec(O.mat_T4(et), gsm), sm, 2)
    assert_array_almost_equal(O._matvec(O.mat_T1(et).T, geo), j2000, 2)
    assert_array_almost_equal(O._matvec(O.mat_j2000_to_geo(et), j2000), geo, 2)
    assert_array_almost_equal(O._matvec(O.mat_j2000_to_sm(et), j2000), sm, 2)
    assert_array_almost_equal(O._matvec(O.mat_geo_to_sm(et), geo), sm, 2)


def test_igrf_range():
    with pytest.raises(ValueError):
        O.mat_j2000_to_sm(O.date2es(datetime.datetime(2021, 1, 1)))


def test_vincenty_a12_sanity():
    # equator: arc on the auxiliary sphere equals the longitude difference scaled by (1-f)^-1... simply monotone
    assert abs(O.vincenty_a12(0, 0, 0, 10) - 10 / (1 - 1 / 298.257223563) * (1 - 1 / 298.257223563)) < 0.05
    a = O.vincenty_a12(55, -111, 55, -92)
    assert 10.8 < a < 10.95       # 19 deg of longitude at 55N ~ 10.87 deg of arc


def test_karney_restatement_against_published_geodesics():
    """oracle/karney_geodesic.py (Karney 2013 restated from the paper; geographiclib itself is absent)
    against the two worked examples of the GeographicLib documentation: JFK -> LHR of the GeodSolve
    manual (s12 = 5551759.400 m) and Wellington -> Salamanca of the Python package's introduction
    (a12, s12 and both azimuths as printed there)."""
    from oracle import karney_geodesic as K
    r = K.inverse(40.6, -73.8, 51.6, -0.5)
    assert abs(r['s12'] - 5551759.400) < 1e-3
    r = K.inverse(-41.32, 174.81, 40.96, -5.50)
    assert abs(r['a12'] - 179.6197069334283) < 1e-12
    assert abs(r['s12'] - 19959679.26735382) < 1e-7
    assert abs(r['azi1'] - 161.06766998615873) < 1e-9
    assert abs(r['azi2'] - 18.825195123248484) < 1e-9
    # closed forms: meridian arc pole to pole = 2 quarter meridians; equator
    assert abs(K.inverse(-90, 0, 90, 0)['a12'] - 180) < 1e-12
    assert abs(K.inverse(-90, 0, 90, 0)['s12'] - 2 * 10001965.729) < 2e-3
    assert abs(K.inverse(0, 0, 0, 90)['s12'] - 6378137.0 * np.pi / 2) < 1e-8
    assert K.inverse(10, 20, 10, 20)['a12'] == 0


def test_angular_distance_vincenty_against_karney():
    """The product's `angularDistance` (Vincenty's iteration, coordinates/geodesic.py) and the oracle's own
    Vincenty restatement against the independent Karney restatement: the auxiliary-sphere arc `a12` that
    reference geodesic.py:35-44 takes from geographiclib.  Random pairs over the globe (nearly antipodal pairs,
    where Vincenty's iteration does not converge and which no bounding box of this path produces, excluded)
    and bounding-box-sized pairs as plateCarreeResolution forms them (resample.py:281-299)."""
    from oracle import karney_geodesic as K
    from auromat_b200.coordinates import geodesic as G
    rng = np.random.default_rng(11)
    worst = 0.0
    for _ in range(3000):
        la1, la2 = rng.uniform(-89, 89, 2)
        lo1, lo2 = rng.uniform(-180, 180, 2)
        k = K.angularDistance(la1, lo1, la2, lo2)
        if k > 175:
            continue
        worst = max(worst, abs(G.angularDistance(G.Location(la1, lo1), G.Location(la2, lo2)) - k),
                    abs(O.vincenty_a12(la1, lo1, la2, lo2) - k))
    assert worst < 2e-9, worst
    worst = 0.0
    for _ in range(3000):
        la1, lo1 = rng.uniform(-85, 85), rng.uniform(-180, 180)
        la2, lo2 = float(np.clip(la1 + rng.uniform(-20, 20), -89, 89)), lo1 + rng.uniform(-40, 40)
        k = K.angularDistance(la1, lo1, la2, lo2)
        v = G.angularDistance(G.Location(la1, lo1), G.Location(la2, lo2))
        worst = max(worst, abs(v - k) / max(k, 1e-9))
    assert worst < 1e-10, worst


def test_allsky_matches_reference_golden():
    """mapping/miracle.py model (reference outputs with the documented np.indices patch)."""
    g = np.load(os.path.join(GOLDEN, "allsky_SOD_96.npz"))
    lat, lon, xc, yc, k, rot = g['cal']
    s = 96 / 512
    o = O.allsky_georeference(96, xc * s, yc * s, k * s, rot, lat, lon, 110)
    for name in ('lats', 'lons', 'latsCenter', 'lonsCenter', 'elevation'):
        assert np.array_equal(np.isnan(o[name]), np.isnan(g[name]))
        assert np.nanmax(np.abs(o[name] - g[name])) <= TOL_DEG, name


def test_themis_reproject_matches_reference_golden():
    """mapping/themis.py:224-253 run by the reference (oracle/gen_golden.py::themis_golden)."""
    g = np.load(os.path.join(GOLDEN, "themis_reproject.npz"))
    asi = tuple(g['asi'])
    for h in (90, 150):
        la, lo = O.themis_reproject(asi, g['lats110'], g['lons110'], 110.0, float(h))
        assert_array_equal(np.isnan(la), np.isnan(g['lats%d' % h]))
        assert np.nanmax(np.abs(la - g['lats%d' % h])) <= TOL_DEG
        assert np.nanmax(np.abs(lo - g['lons%d' % h])) <= TOL_DEG
    # reprojecting to the calibration's own height is the identity up to the error of the
    # single-iteration Bowring inverse at 110 km height (a few 1e-6 deg, inherent to the reference)
    la, lo = O.themis_reproject(asi, g['lats110'], g['lons110'], 110.0, 110.0)
    assert np.nanmax(np.abs(la - g['lats110'])) < 2e-5 and np.nanmax(np.abs(lo - g['lons110'])) < 2e-5


def _sip_exact(header, prefix, u, v):
    """FITS-SIP forward polynomial sum C_p_q u^p v^q in exact rational arithmetic."""
    from fractions import Fraction
    order = int(header[prefix + '_ORDER'])
    acc = Fraction(0)
    fu, fv = Fraction(u), Fraction(v)
    for p in range(order + 1):
        for q in range(order + 1 - p):
            c = header.get('%s_%d_%d' % (prefix, p, q))
            if c is not None:
                acc += Fraction(float(c)) * fu ** p * fv ** q
    return acc


def sip_kat_points(n=120, seed=9):
    rng = np.random.default_rng(seed)
    u = rng.uniform(-3000, 3000, n)
    v = rng.uniform(-2000, 2000, n)
    u[:4] = [0.0, 3000.0, -3000.0, 1.0]
    v[:4] = [0.0, -2000.0, 2000.0, -1.0]
    return u, v


def test_sip_polynomial_against_exact_rational_arithmetic():
    """Independent pin of the SIP restatement: the FITS-SIP convention u' = u + sum A_p_q u^p v^q evaluated
    with `fractions.Fraction` (no rounding at all) at 120 points of the configs[2] header; the oracle's
    fixed-order Horner evaluation must agree to a few ulp of the largest term."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    hdr = synthetic.issHeader(6000, 4000, sipOrder=4)
    oa, A, ob, B = O.sip_coefficients(hdr)
    u, v = sip_kat_points()
    fa, fb = O._sip_poly(A, u, v), O._sip_poly(B, u, v)
    for i in range(len(u)):
        ea, eb = _sip_exact(hdr, 'A', u[i], v[i]), _sip_exact(hdr, 'B', u[i], v[i])
        # |distortion| <= ~20 px: 1e-13 px absolute is a handful of ulp of the result
        assert abs(float(ea) - fa[i]) <= 1e-13 + 4e-16 * abs(float(ea)), (i, float(ea), fa[i])
        assert abs(float(eb) - fb[i]) <= 1e-13 + 4e-16 * abs(float(eb)), (i, float(eb), fb[i])
    assert np.abs(fa).max() > 1.0          # the distortion is not trivially small
