"""Outline / polygon helpers behind BaseMapping.outline, .centroid, .maskedByPolygon and the
pole test: host implementations (auromat_b200/utils.py, coordinates/geodesic.py) against the
oracle restatements and against the known answers of the reference's own tests
(test/outline_test.py:107-145, test/geodesic_test.py:15-30)."""
import numpy as np
import pytest

import oracle.auromat_oracle as O
from auromat_b200 import utils as U
from auromat_b200.coordinates import geodesic

# test/outline_test.py:108-127 -- the outline polygon of the reference's 10x10 test disc
REF_DISC_OUTLINE = [[4, 8], [3, 7], [2, 7], [1, 6], [1, 5], [1, 4], [1, 3], [1, 2], [2, 1], [3, 1], [4, 0], [5, 1],
                    [6, 1], [7, 2], [7, 3], [8, 4], [7, 5], [7, 6], [6, 7], [5, 7]]


def ref_test_image(n=10):
    """test/outline_test.py:22-35 `_testIm` (disc of radius 0.4 n in the top-left corner, one
    node removed to break the symmetry)."""
    coord = np.ones((n, n))
    r = n * 0.4
    y, x = np.ogrid[-r: r + 1, -r: r + 1]
    disc = x ** 2 + y ** 2 <= r ** 2
    im = np.zeros((n, n), bool)
    im[:disc.shape[0], :disc.shape[1]] = disc
    im[4, 0] = False
    return im


def same_cycle(a, b):
    a = [tuple(p) for p in np.asarray(a).tolist()]
    b = [tuple(p) for p in np.asarray(b).tolist()]
    return len(a) == len(b) and any(a[s:] + a[:s] == b for s in range(len(a)))


def test_outline_reproduces_the_reference_test_polygon():
    im = ref_test_image(10)
    assert same_cycle(U.outline(im), REF_DISC_OUTLINE)            # same nodes, same orientation
    assert same_cycle(O.outline_marching_squares(im), REF_DISC_OUTLINE)
    assert U.polygonArea(REF_DISC_OUTLINE) == 37.0                # outline_test.py:130
    assert O.polygon_area(REF_DISC_OUTLINE) == 37.0


def test_outline_walk_equals_marching_squares_on_random_masks():
    """Crack following (host) vs cell-wise marching squares (oracle): noisy masks with holes,
    one-node bridges, diagonal contacts and several regions."""
    rng = np.random.default_rng(7)
    checked = 0
    for _ in range(400):
        h, w = rng.integers(1, 28, 2)
        im = rng.random((h, w)) < rng.uniform(0.25, 0.95)
        if not im.any():
            continue
        a, b = U.outline(im), O.outline_marching_squares(im)
        assert same_cycle(a, b), (im.astype(int), a, b)
        assert im[a[:, 1], a[:, 0]].all()
        checked += 1
    assert checked > 300
    big = ref_test_image(800)
    assert same_cycle(U.outline(big), O.outline_marching_squares(big))
    # two regions: the larger one is returned (utils.py:120-132)
    two = np.hstack((ref_test_image(40), np.zeros((40, 2), bool), ref_test_image(40)[:, :20]))
    assert same_cycle(U.outline(two), U.outline(ref_test_image(40)))


def test_polygon_centroid_known_answer():
    poly = [(30, 50), (200, 10), (250, 50), (350, 100), (200, 180), (100, 140), (10, 200)]   # outline_test.py:134-145
    np.testing.assert_almost_equal(U.polygonCentroid(poly), (159.2903828197946, 98.88888888888))
    np.testing.assert_almost_equal(O.polygon_centroid(poly), (159.2903828197946, 98.88888888888))
    # orientation independent (the product divides by the signed area)
    np.testing.assert_almost_equal(U.polygonCentroid(poly[::-1]), (159.2903828197946, 98.88888888888))
    # a far-away small polygon keeps its precision
    sq = np.array([[0, 0], [1e-3, 0], [1e-3, 1e-3], [0, 1e-3]]) + [55.0, -99.0]
    np.testing.assert_allclose(U.polygonCentroid(sq), (55.0005, -98.9995), rtol=0, atol=1e-12)


def test_convex_hull_and_inside_test_vs_oracle():
    rng = np.random.default_rng(3)
    for n in (3, 10, 200):
        pts = rng.integers(0, 40, (n, 2))
        assert np.array_equal(U.convexHull(pts), O.convex_hull(pts))
    hull = U.convexHull(U.outline(ref_test_image(60)))
    assert U.polygonArea(hull) >= U.polygonArea(U.outline(ref_test_image(60)))
    for _ in range(20):
        poly = rng.random((rng.integers(3, 12), 2)) * 10
        q = rng.random((300, 2)) * 12 - 1
        assert np.array_equal(U.pointsInsidePolygon(q, poly), O.points_inside_polygon(q, poly))
    square = [[0, 0], [0, 4], [4, 4], [4, 0]]
    q = np.array([[2, 2], [5, 2], [-1, 2], [2, 5], [3.999, 0.001], [np.nan, 1.0]])
    assert U.pointsInsidePolygon(q, square).tolist() == [True, False, False, False, True, False]


def test_contains_or_crosses_pole_known_answers():
    """test/geodesic_test.py:15-30."""
    f = geodesic.containsOrCrossesPole
    assert not f([[1, 0], [1, 4], [5, 6], [5, 2]])
    assert not f([[1, 179], [1, -177], [5, -175], [5, -179]])          # across the date line
    assert f([[85, -135], [85, -45], [85, 45], [85, 135]])              # around the north pole
    assert f([[85, -90], [85, 0], [85, 90]])                            # crossing the north pole
    assert f([[-80, 0], [-80, -90], [-80, 180], [-80, 90]])             # south pole
    assert abs(geodesic.course(geodesic.Location(0, 0), geodesic.Location(0, 10)) - 90) < 1e-9
    assert abs(geodesic.course(geodesic.Location(10, 20), geodesic.Location(50, 20))) < 1e-9
    # geodesic length of one degree of the equator / of a meridian arc from the equator
    assert abs(geodesic.distance(geodesic.Location(0, 0), geodesic.Location(0, 1)) - 111319.4908) < 1e-3
    assert abs(geodesic.distance(geodesic.Location(0, 0), geodesic.Location(1, 0)) - 110574.3886) < 1e-3
    assert abs(geodesic.angularDistance(geodesic.Location(50, 10), geodesic.Location(51, 12))
               - O.vincenty_a12(50, 10, 51, 12)) < 1e-12


def test_bounding_box_center_and_size_known_answers():
    """test/boundingbox_test.py:12-47 (geographiclib results, 6 decimals)."""
    import warnings
    from numpy.testing import assert_array_almost_equal
    from auromat_b200.mapping.mapping import BoundingBox
    bb = BoundingBox(latSouth=-60, lonWest=80, latNorth=-30, lonEast=85)
    assert_array_almost_equal(bb.center, [-45.03119418083877, 82.5])
    assert_array_almost_equal(bb.size, [482.39311013217343, 3336.5953086140203])
    bb = BoundingBox(latSouth=-60.646114098, lonWest=82.7852215499, latNorth=-38.7515567117, lonEast=-178.546517062)
    assert_array_almost_equal(bb.center, [-54.33647117488648, 132.11935224395])
    assert_array_almost_equal(bb.size, [8084.704893634039, 3464.8889697347718])
    bb = BoundingBox(latSouth=60, lonWest=-180, latNorth=90, lonEast=180)
    assert_array_almost_equal(bb.center, [90, 0])
    assert_array_almost_equal(bb.size, [6695.78581964, 6695.78581964])
    bb = BoundingBox(latSouth=-90, lonWest=-180, latNorth=-60, lonEast=180)
    assert_array_almost_equal(bb.center, [-90, 0])
    assert_array_almost_equal(bb.size, [6695.78581964, 6695.78581964])
    bb = BoundingBox(latSouth=50, lonWest=80, latNorth=50, lonEast=80)
    assert_array_almost_equal(bb.center, [50, 80])
    assert_array_almost_equal(bb.size, 0)
    bb1 = BoundingBox(latSouth=-55, lonWest=95, latNorth=-45, lonEast=109)
    bb2 = BoundingBox(latSouth=44, lonWest=-164, latNorth=74, lonEast=-35)
    bb = BoundingBox.mergedBoundingBoxes([bb1, bb2])
    assert [bb.latSouth, bb.latNorth, bb.lonWest, bb.lonEast] == [bb1.latSouth, bb2.latNorth, bb1.lonWest, bb2.lonEast]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert_array_almost_equal(bb.center, [21.136113246, -150])


def test_provider_wrappers():
    """MaskByElevationProvider / ResampleProvider (mapping.py:1447-1472, resample.py:370-394)
    route get / getById / getSequence through the wrapped operation."""
    from auromat_b200.mapping.mapping import BaseMappingProvider, MaskByElevationProvider
    from auromat_b200 import resample as R

    class FakeMapping(object):
        def __init__(self, i):
            self.i, self.masked = i, None

        def maskedByElevation(self, minElevation=10):
            m = FakeMapping(self.i)
            m.masked = minElevation
            return m

    class Provider(BaseMappingProvider):
        def get(self, date):
            return FakeMapping(date)

        def getById(self, identifier):
            return FakeMapping(identifier)

        def getSequence(self, dateBegin=None, dateEnd=None):
            return (FakeMapping(i) for i in range(3))

        def contains(self, date):
            return date == 1

    p = Provider(maxTimeOffset=5)
    assert p.containsAny([0, 1]) and not p.containsAny([0, 2])
    w = MaskByElevationProvider(p, 7)
    assert w.get(1).masked == 7 and w.getById('x').masked == 7 and [m.masked for m in w.getSequence()] == [7, 7, 7]
    assert w.maxTimeOffset == 5 and p.get(1).masked is None and isinstance(w, Provider)
    calls = []
    orig = R.resample
    R.resample = lambda m, **kw: calls.append((m.i, kw)) or m
    try:
        rp = R.ResampleProvider(p, arcsecPerPx=100)
        rp.get(4)
        list(rp.getSequence())
    finally:
        R.resample = orig
    assert calls == [(4, {'arcsecPerPx': 100}), (0, {'arcsecPerPx': 100}), (1, {'arcsecPerPx': 100}), (2, {'arcsecPerPx': 100})]
