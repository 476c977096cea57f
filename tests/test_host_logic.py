"""CPU tests of the product's host side: per-frame constants, grid derivation, header parsing,
bounding boxes, and that the C-ABI library loads and exports every declared symbol.
No kernel is launched here."""
import ctypes
import datetime
import os
import re

import numpy as np
import pytest

import oracle.auromat_oracle as O
from auromat_b200 import _lib, fits, synthetic
from auromat_b200.coordinates import geodesic, igrf, transform, wcs
from auromat_b200.mapping.mapping import BoundingBox, wrapAt180
from auromat_b200 import resample as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "auromat_b200.h")).read()
    declared = set(re.findall(r"\b(amt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    loaded = _lib.load()
    assert loaded.amt_abi_version() == _lib.ABI_VERSION == 3


def test_struct_layouts_match_header():
    # sizes implied by include/auromat_b200.h (natural alignment, no packing)
    assert ctypes.sizeof(_lib.AmtFrame) == 4 * 4 + 8 * (2 + 4 + 9 + 3 + 3 + 9 + 9 + 2) + 2 * 4 + 8 * 2 * 55 + 2 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.AmtGeorefOut) == 11 * 8
    assert ctypes.sizeof(_lib.AmtStats) == 6 * 8 + 5 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.AmtGrid) == 4 * 4 + 8 * (6 + 2 + 3 + 9 + 1)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from auromat_b200.mapping.spacecraft import getMapping
    m = getMapping(synthetic.issImage(32, 24), synthetic.issHeader(32, 24), identifier='x')
    with pytest.raises(RuntimeError):
        m.lats                              # no CPU fallback


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "auromat_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("date", [datetime.datetime(2012, 1, 25, 9, 26, 55, 60000),
                                  datetime.datetime(2003, 7, 1, 23, 59, 1),
                                  datetime.datetime(1999, 12, 31, 12, 0, 0)])
def test_frame_matrices_bit_equal_to_oracle(date):
    et = transform.date2es(date)
    assert et == O.date2es(date)
    for name in ('mat_P', 'mat_T1', 'mat_T2', 'mat_T3', 'mat_T4', 'mat_j2000_to_geo', 'mat_j2000_to_sm',
                 'mat_geo_to_sm'):
        assert np.array_equal(getattr(transform, name)(et), getattr(O, name)(et)), name
    geo, sm, geosm = transform.frameMatrices(et)
    assert np.array_equal(geo, O.mat_j2000_to_geo(et))
    assert np.array_equal(sm, O.mat_j2000_to_sm(et))
    assert np.array_equal(geosm, O.mat_geo_to_sm(et))
    for ang, axis in ((0.3, [1, 0, 0]), (-2.1, [0, 0, -1]), (1e-3, [0, 1, 0]), (2.5, [-1, 0, 0])):
        assert np.array_equal(transform.rotation_matrix(ang, axis), O.rotation_matrix3(ang, axis))


def test_igrf_and_constants():
    assert geodesic.wgs84A == O.WGS84_A and geodesic.wgs84B == O.WGS84_B
    assert igrf.IGRF_DEFINED_UNTIL_YEAR == 2020
    with pytest.raises(ValueError):
        transform.mat_j2000_to_sm(transform.date2es(datetime.datetime(2020, 6, 1)))
    et = transform.date2es(datetime.datetime(2012, 1, 25))
    assert transform.mag_lat(et) == O.mag_lat(et) and transform.mag_lon(et) == O.mag_lon(et)


def test_frame_constants():
    hdr = synthetic.issHeader(133, 89, sipOrder=3)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    fr = wcs.frameConstants(hdr, cam, t, 110, fastCenterCalculation=True)
    assert (fr.width, fr.height, fr.fast_center, fr.origin_inside) == (133, 89, 1, 0)
    assert np.array_equal(np.array(fr.rot[:]).reshape(3, 3), O.tan_native_rotation(hdr))
    et = O.date2es(t)
    assert np.array_equal(np.array(fr.m_geo[:]).reshape(3, 3), O.mat_j2000_to_geo(et))
    assert np.array_equal(np.array(fr.m_sm[:]).reshape(3, 3), O.mat_j2000_to_sm(et))
    a, b = O.WGS84_A + 110, O.WGS84_B + 110
    assert list(fr.inv_axes[:]) == [1 / a, 1 / a, 1 / b]
    oa, A, ob, B = O.sip_coefficients(hdr)
    assert (fr.sip_order_a, fr.sip_order_b) == (3, 3)
    for (p, q), v in A.items():
        assert fr.sip_a[p * 4 - (p * (p - 1)) // 2 + q] == v
    with pytest.raises(NotImplementedError):
        bad = dict(hdr)
        bad['CTYPE1'] = 'RA---SIN'
        wcs.frameConstants(bad, cam, t, 110)


def test_fixed_grid_equals_reference_arithmetic():
    rng = np.random.default_rng(0)
    for k in range(400):
        ppd = (rng.uniform(1, 400), rng.uniform(1, 400)) if k % 3 else (36.0, float(rng.integers(1, 50)))
        la = np.sort(rng.uniform(-89, 89, 2))
        lo = np.sort(rng.uniform(-179, 179, 2))
        a = R.fixedGrid(ppd, la[0], la[1], lo[0], lo[1])
        b = O.fixed_grid(ppd, la[0], la[1], lo[0], lo[1])
        assert tuple(a) == tuple(b)
    # bounds that sit exactly on grid nodes, and the argmax wrap-around corner cases
    assert tuple(R.fixedGrid((4, 4), 10.0, 20.25, -30.0, -10.5)) == tuple(O.fixed_grid((4, 4), 10.0, 20.25, -30.0, -10.5))
    assert tuple(R.fixedGrid((2, 2), -90.0, 90.0, -180.0, 180.0)) == tuple(O.fixed_grid((2, 2), -90.0, 90.0, -180.0, 180.0))


def test_target_grid_matches_histogram_range():
    """The amt_grid handed to the kernels equals the bins/range the reference passes to
    histogram2d (resample.py:330-337) and its rounding decimals (histogram.py:215-219)."""
    ppd = (36.0, 20.7)
    latMin, latMax, lonMin, lonMax = 47.9, 61.3, -111.6, -91.9
    g, info = R.targetGrid(ppd, latMin, latMax, lonMin, lonMax)
    nLat, nLon, a, b, c, d = O.fixed_grid(ppd, latMin, latMax, lonMin, lonMax)
    latSC, latStep = np.linspace(b, a, num=nLat, retstep=True)
    lonSC, lonStep = np.linspace(c, d, num=nLon, retstep=True)
    latSC, lonSC = latSC[1:-1], lonSC[1:-1]
    assert (g.nx, g.ny) == (len(lonSC), len(latSC))
    assert (g.lo_x, g.hi_x) == (lonSC[0] - lonStep / 2, lonSC[-1] + lonStep / 2)
    assert (g.lo_y, g.hi_y) == (latSC[-1] + latStep / 2, latSC[0] - latStep / 2)
    ex = np.linspace(g.lo_x, g.hi_x, g.nx + 1)
    assert g.step_x == (ex[-1] - ex[0]) / g.nx
    assert g.round_x == 10.0 ** (int(-np.log10(np.diff(ex).min())) + 6)
    ey = np.linspace(g.lo_y, g.hi_y, g.ny + 1)
    assert g.round_y == 10.0 ** (int(-np.log10(np.diff(ey).min())) + 6)
    with pytest.raises(AssertionError):
        R.targetGrid((0, 1), 0, 1, 0, 1)


def test_plate_carree_resolution():
    bb = BoundingBox(47.9, -111.6, 61.3, -91.9)
    lat, lon = R.plateCarreeResolution(bb, 100)
    olat, olon = O.plate_carree_resolution((47.9, -111.6, 61.3, -91.9), 100)
    assert (lat, lon) == (olat, olon)
    assert lat == 1 / (100 * (1.0 / 3600.0))
    assert 19 < lon < 22           # cos(54.6 deg) * 36 px/deg
    # spans the date line
    lat2, lon2 = R.plateCarreeResolution(BoundingBox(60, 170, 70, -170), 100)
    assert 12 < lon2 < 19


def test_wrap_at_180():
    x = np.array([-540.0, -180.0, -179.9, 0.0, 179.9, 180.0, 190.0, 359.0, 360.0, 725.0, np.nan])
    w = wrapAt180(x)
    assert np.array_equal(w[:-1], [-180.0, -180.0, -179.9, 0.0, 179.9, -180.0, -170.0, -1.0, 0.0, 5.0])
    assert np.isnan(w[-1])
    assert np.array_equal(w, O.wrap_at_180(x), equal_nan=True)
    assert wrapAt180(190.0) == -170.0


def test_bounding_box():
    bb = BoundingBox(10, 170, 20, -170)
    assert bb.containsDiscontinuity and not bb.containsPole
    assert BoundingBox(60, -180, 90, 180).containsPole
    with pytest.raises(AssertionError):
        BoundingBox(0, -190, 1, 0)
    m = BoundingBox.mergedBoundingBoxes([BoundingBox(10, 160, 20, 175), BoundingBox(5, -178, 15, -170)])
    assert (m.latSouth, m.lonWest, m.latNorth, m.lonEast) == (5, 160, 20, -170)
    m = BoundingBox.mergedBoundingBoxes([BoundingBox(0, -10, 1, 10), BoundingBox(0, 5, 2, 30)])
    assert (m.lonWest, m.lonEast) == (-10, 30)
    m = BoundingBox.minimumBoundingBox([(1, 2), (3, 4), (-1, 3)])
    assert (m.latSouth, m.lonWest, m.latNorth, m.lonEast) == (-1, 2, 3, 4)
    assert BoundingBox(1, 2, 3, 4) == BoundingBox(1, 2, 3, 4)


def _card(key, value, comment=''):
    if isinstance(value, str):
        v = "'%-8s'" % value
        s = "%-8s= %-20s" % (key, v)
    elif isinstance(value, bool):
        s = "%-8s= %20s" % (key, 'T' if value else 'F')
    else:
        s = "%-8s= %20s" % (key, repr(value))
    if comment:
        s += " / " + comment
    return s[:80].ljust(80)


def test_fits_header_reader(tmp_path):
    hdr = synthetic.issHeader()
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0)]
    cards += [_card(k, v, 'x') for k, v in hdr.items()]
    cards.append("COMMENT this is ignored".ljust(80))
    cards.append(_card('NORADID', '25544'))
    cards.append("END".ljust(80))
    raw = "".join(cards)
    raw = raw.ljust((len(raw) + 2879) // 2880 * 2880)
    p = tmp_path / "frame.wcs"
    p.write_bytes(raw.encode('ascii'))
    h = fits.readHeader(str(p))
    for k, v in hdr.items():
        assert h[k] == v, k
    assert fits.getNoradId(h) == 25544
    assert fits.getPhotoTime(h) == datetime.datetime(2012, 1, 25, 9, 27, 8, 60000)
    pos, date, delta = fits.getShiftedSpacecraftPosition(h)
    assert date == datetime.datetime(2012, 1, 25, 9, 26, 55, 60000) and delta.total_seconds() == -13.0
    assert np.array_equal(pos, [hdr['POSXSHIF'], hdr['POSYSHIF'], hdr['POSZSHIF']])
    assert fits.getSpacecraftPosition(h) == (None, None)
    from auromat_b200.mapping.spacecraft import _prepareMappingParams
    header, photoTime, original, cam = _prepareMappingParams(str(p))
    assert photoTime == date and original == fits.getPhotoTime(h) and np.array_equal(cam, pos)
    # timeshift without TLE data cannot be honoured (reference spacecraft.py:441-462)
    with pytest.raises(ValueError):
        _prepareMappingParams(h, timeshift=datetime.timedelta(seconds=1))


def test_synthetic_sequence():
    hs = synthetic.sequenceHeaders(5, 133, 89)
    r = [np.linalg.norm([h['POSXSHIF'], h['POSYSHIF'], h['POSZSHIF']]) for h in hs]
    assert np.allclose(r, r[0])
    assert hs[1]['CRVAL1'] - hs[0]['CRVAL1'] == pytest.approx(0.05)
    t0, _ = synthetic.headerTimeAndCamera(hs[0])
    t4, _ = synthetic.headerTimeAndCamera(hs[4])
    assert (t4 - t0).total_seconds() == 4


def test_target_grid_round_scale_randomised():
    """The analytic shortcut for `decimal` (histogram.py:215-219) equals the numpy evaluation."""
    rng = np.random.default_rng(4)
    for _ in range(200):
        ppd = (float(rng.uniform(0.5, 400)), float(rng.uniform(0.5, 400)))
        la = np.sort(rng.uniform(-80, 80, 2))
        lo = np.sort(rng.uniform(-170, 170, 2))
        if la[1] - la[0] < 3 / ppd[0] or lo[1] - lo[0] < 3 / ppd[1]:
            continue
        g, info = R.targetGrid(ppd, la[0], la[1], lo[0], lo[1])
        ex = np.linspace(g.lo_x, g.hi_x, g.nx + 1)
        ey = np.linspace(g.lo_y, g.hi_y, g.ny + 1)
        assert g.round_x == 10.0 ** (int(-np.log10(np.diff(ex).min())) + 6)
        assert g.round_y == 10.0 ** (int(-np.log10(np.diff(ey).min())) + 6)
        assert g.step_x == (ex[-1] - ex[0]) / g.nx and g.step_y == (ey[-1] - ey[0]) / g.ny
    # power-of-ten steps take the exact path
    g, _ = R.targetGrid((10, 10), 10.0, 20.0, 30.0, 50.0)
    ex = np.linspace(g.lo_x, g.hi_x, g.nx + 1)
    assert g.round_x == 10.0 ** (int(-np.log10(np.diff(ex).min())) + 6)


def _write_wcs(path, hdr):
    cards = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0)]
    cards += [_card(k, v, 'x') for k, v in hdr.items()] + ["END".ljust(80)]
    raw = "".join(cards)
    path.write_bytes(raw.ljust((len(raw) + 2879) // 2880 * 2880).encode('ascii'))


def test_spacecraft_mapping_providers(tmp_path):
    """SpacecraftMappingProvider / SpacecraftMappingPathProvider (spacecraft.py:40-292): folder
    scan, solved/unsolved ids, ordering by header time, access by date, metadata merge.  The
    mappings are lazy -- nothing here touches a GPU."""
    import json
    from PIL import Image
    from auromat_b200.mapping.spacecraft import SpacecraftMappingPathProvider, SpacecraftMappingProvider
    from auromat_b200.utils import findNearest
    W, H = 32, 24
    hdrs = synthetic.sequenceHeaders(3, W, H)
    names = ['ISS030-E-0003', 'ISS030-E-0001', 'ISS030-E-0002']            # file order != time order
    for name, hdr in zip(names, hdrs):
        _write_wcs(tmp_path / (name + '.wcs'), hdr)
        Image.fromarray(synthetic.issImage(W, H)).save(str(tmp_path / (name + '.png')))
    Image.fromarray(synthetic.issImage(W, H)).save(str(tmp_path / 'ISS030-E-0009.png'))      # not solved
    (tmp_path / 'metadata.json').write_text(json.dumps({
        'sequence_metadata': {'lens': '24mm'},
        'image_metadata': {n: {'iso': 1000 + i} for i, n in enumerate(names)}}))
    p = SpacecraftMappingProvider(str(tmp_path), altitude=120, fastCenterCalculation=True)
    assert len(p) == 3 and p.imageFileExtension == 'png'
    assert p.ids == names and p.unsolvedIds == ['ISS030-E-0009']         # headers are 1 s apart, in `names` order
    t0, _ = synthetic.headerTimeAndCamera(hdrs[0])
    assert p.range == (t0, t0 + datetime.timedelta(seconds=2))
    assert p.contains(t0 + datetime.timedelta(seconds=1.4)) and not p.contains(t0 + datetime.timedelta(seconds=9))
    m = p.get(t0 + datetime.timedelta(seconds=1.2))
    assert m.identifier == names[1] and m.altitude == 120 and m.fastCenterCalculation
    assert m.photoTime == t0 + datetime.timedelta(seconds=1)
    assert m.metadata == {'lens': '24mm', 'iso': 1001}
    assert p.getById('E-0002').identifier == 'ISS030-E-0002'
    with pytest.raises(ValueError):
        p.get(t0 + datetime.timedelta(seconds=30))
    seq = list(p.getSequence())
    assert [s.identifier for s in seq] == names and seq[2].metadata['iso'] == 1002
    assert seq[0].img_unmasked.shape == (H, W, 3)
    q = SpacecraftMappingPathProvider([str(tmp_path / (n + '.png')) for n in reversed(names)],
                                      [str(tmp_path / (n + '.wcs')) for n in reversed(names)])
    assert [os.path.basename(w) for w in q.wcsPaths] == [n + '.wcs' for n in names] and q.imageFileExtension == 'png'
    assert [s.identifier for s in q.getSequence()] == names
    # lists instead of folders
    r = SpacecraftMappingProvider([str(tmp_path / (n + '.png')) for n in names], [str(tmp_path / (n + '.wcs')) for n in names])
    assert r.ids == names
    assert findNearest([1, 4, 9], 6) == 1 and findNearest([1, 4, 9], 7) == 2 and findNearest([1, 4, 9], 100) == 2
    assert findNearest([1, 5], 3) == 0 and findNearest([1, 4, 9], -3) == 0


def test_c_grid_derivation_is_bit_identical_to_python():
    """csrc/amt_host.cuh ports plateCarreeResolution / fixedGrid / targetGrid / sideScale to C for the
    sequence engine: on thousands of random bounding boxes and resolutions every field equals the
    Python derivation (which follows the reference's arithmetic) bit for bit."""
    import ctypes as C
    from auromat_b200 import resample as R
    from auromat_b200.mapping.mapping import BoundingBox
    lib = _lib.load()
    rng = np.random.default_rng(12)
    n_ok = 0
    for i in range(4000):
        latS = float(rng.uniform(-85, 80))
        latN = float(min(89.0, latS + rng.uniform(0.3, 30)))
        lonW = float(rng.uniform(-179, 150))
        lonE = float(min(179.5, lonW + rng.uniform(0.3, 60)))
        arcsec = float(rng.choice([10, 25, 100, 400, 977.3, 36.0, 3600.0]))
        bb = BoundingBox(latS, lonW, latN, lonE)
        ppd = R.plateCarreeResolution(bb, arcsec)
        a, b = C.c_double(), C.c_double()
        assert lib.amt_plate_carree_resolution(latS, lonW, latN, lonE, arcsec, C.byref(a), C.byref(b)) == 0
        assert (a.value, b.value) == ppd, (i, (a.value, b.value), ppd)
        if i % 3 == 0:
            ppd = (float(rng.uniform(1, 50)), float(rng.uniform(1, 50)))
        try:
            g, info = R.targetGrid(ppd, latS, latN, lonW, lonE)
        except ValueError:
            continue
        cg, ci, fb = _lib.AmtGrid(), _lib.AmtGridInfo(), C.c_int32()
        assert lib.amt_target_grid(ppd[0], ppd[1], latS, latN, lonW, lonE, C.byref(cg), C.byref(ci), C.byref(fb)) == 0
        if fb.value:
            continue
        for f in ('nx', 'ny', 'prerotate', 'lo_x', 'hi_x', 'step_x', 'lo_y', 'hi_y', 'step_y', 'round_x', 'round_y',
                  'wgs_a', 'wgs_b'):
            assert getattr(cg, f) == getattr(g, f), (i, f, getattr(cg, f), getattr(g, f))
        assert np.allclose(cg.rot[:], g.rot[:], atol=1e-15)
        assert (ci.n_lat, ci.n_lon, ci.lat_min_in_grid, ci.lat_max_in_grid, ci.lon_min_in_grid, ci.lon_max_in_grid,
                ci.lat_step, ci.lon_step) == (info['nLat'], info['nLon'], info['latMinInGrid'], info['latMaxInGrid'],
                                              info['lonMinInGrid'], info['lonMaxInGrid'], info['latStep'], info['lonStep'])
        n_ok += 1
    assert n_ok > 3000
    lib.amt_side_scale.restype = C.c_double
    for n in (1, 2, 3, 1000, 65536, 65537, 12052992, 24000000, 2 ** 31, 2 ** 40):
        assert lib.amt_side_scale(n) == R.sideScale(n), n


def test_c_sip_displacement_bound_encloses_the_polynomial():
    """amt_sip_displacement_bound (the box inflation of the TAN-SIP hit-bitmap solver): never below the
    largest displacement the oracle's SIP polynomial produces anywhere on the corner / centre lattice of the
    frame (a bound that is too small would let the solver fill a word it should have evaluated), and not
    absurdly above it; zero for a pure TAN header."""
    import ctypes as C
    from auromat_b200.coordinates.wcs import frameConstants
    lib = _lib.load()
    for W, H, order, seed in [(97, 61, 2, 3), (532, 354, 4, 4), (1064, 708, 3, 5), (700, 500, 5, 6), (6000, 4000, 4, 7)]:
        hdr = synthetic.issHeader(W, H, sipOrder=order, seed=seed)
        t, cam = synthetic.headerTimeAndCamera(hdr)
        fr = frameConstants(hdr, cam, t, 110)
        dx, dy = C.c_double(), C.c_double()
        assert lib.amt_sip_displacement_bound(C.byref(fr), C.byref(dx), C.byref(dy)) == 0
        oa, A, ob, B = O.sip_coefficients(hdr)
        # every ray the kernels evaluate: corners at x - 0.5 and centres at x, x = 0 .. W (wcs.py:93-99: + 1 - CRPIX)
        step = max(1, W // 400)
        xs = np.unique(np.concatenate([np.arange(0, W + 1, step), [W]])).astype(float)
        ys = np.unique(np.concatenate([np.arange(0, H + 1, step), [H]])).astype(float)
        worst = [0.0, 0.0]
        for off in (-0.5, 0.0):
            u, v = np.meshgrid(xs + off - hdr['CRPIX1'] + 1, ys + off - hdr['CRPIX2'] + 1)
            worst[0] = max(worst[0], float(np.abs(O._sip_poly(A, u, v)).max()))
            worst[1] = max(worst[1], float(np.abs(O._sip_poly(B, u, v)).max()))
        assert worst[0] <= dx.value <= 12 * worst[0] + 1e-9, (W, H, order, worst, dx.value)
        assert worst[1] <= dy.value <= 12 * worst[1] + 1e-9, (W, H, order, worst, dy.value)
    hdr = synthetic.issHeader(640, 426)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    fr = frameConstants(hdr, cam, t, 110)
    dx, dy = C.c_double(1.0), C.c_double(1.0)
    assert lib.amt_sip_displacement_bound(C.byref(fr), C.byref(dx), C.byref(dy)) == 0
    assert dx.value == 0.0 and dy.value == 0.0


def test_c_pole_pixels_matches_numpy_projection():
    """amt_pole_pixels (inverse WCS projection of the pole point, C) against an independent numpy
    evaluation, for cameras around the pole, elsewhere, and with SIP."""
    import ctypes as C
    from auromat_b200 import synthetic
    from auromat_b200.coordinates.geodesic import wgs84A, wgs84B
    from auromat_b200.coordinates.wcs import frameConstants
    lib = _lib.load()

    def numpy_pixels(fr, w, h, altitude):
        a, b = wgs84A + altitude, wgs84B + altitude
        cam = np.array(fr.cam[:])
        mgeo = np.array(fr.m_geo[:]).reshape(3, 3)
        rot = np.array(fr.rot[:]).reshape(3, 3)
        cd = np.array(fr.cd[:]).reshape(2, 2)
        out = []
        for sign in (1.0, -1.0):
            P = mgeo.T.dot(np.array([0.0, 0.0, sign * b]))
            d = P - cam
            normal = P / np.array([a * a, a * a, b * b])
            if (np.dot(d, normal) >= 0) != bool(fr.origin_inside):
                out.append(None)
                continue
            lmn = rot.T.dot(d / np.linalg.norm(d))
            if lmn[2] <= 0:
                out.append(None)
                continue
            K = 180.0 / np.pi
            uv = np.linalg.solve(cd, np.array([K * lmn[1] / lmn[2], -K * lmn[0] / lmn[2]]))
            px, py = uv[0] + fr.crpix[0] - 1, uv[1] + fr.crpix[1] - 1
            if -0.5 <= px <= w - 0.5 and -0.5 <= py <= h - 0.5:
                out.append((min(max(int(np.floor(px + 0.5)), 0), w - 1), min(max(int(np.floor(py + 0.5)), 0), h - 1)))
            else:
                out.append(None)
        return out

    W, H = 640, 426
    seen = 0
    for hdr in [synthetic.issHeader(W, H), synthetic.issHeaderLookingAt(80.0, 10.0, 89.5, 40.0, W, H),
                synthetic.issHeaderLookingAt(84.0, -120.0, 88.0, 100.0, W, H),
                synthetic.issHeaderLookingAt(-80.0, 10.0, -89.0, 40.0, W, H),
                synthetic.issHeaderLookingAt(50.0, 170.0, 55.0, -178.0, W, H)]:
        t, cam = synthetic.headerTimeAndCamera(hdr)
        fr = frameConstants(hdr, cam, t, 110)
        ix, iy, inf = (C.c_int32 * 2)(), (C.c_int32 * 2)(), (C.c_int32 * 2)()
        assert lib.amt_pole_pixels(C.byref(fr), ix, iy, inf) == 0
        ref = numpy_pixels(fr, W, H, 110)
        for i in range(2):
            if ref[i] is None:
                assert inf[i] == 0
            else:
                assert inf[i] == 1 and (ix[i], iy[i]) == ref[i]
                seen += 1
    assert seen >= 3


def test_seq_output_layout():
    import ctypes as C
    lib = _lib.load()
    om, os_, tot = C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.amt_seq_output_layout(484, 412, 3, _lib.AMT_U8, C.byref(om), C.byref(os_), C.byref(tot)) == 0
    cells = 484 * 412
    assert om.value == (cells * 3 + 63) // 64 * 64 and os_.value == (om.value + cells + 63) // 64 * 64
    assert tot.value == os_.value + cells * 8
    assert lib.amt_seq_output_layout(0, 4, 3, _lib.AMT_U8, None, None, None) == _lib.AMT_ERR_INVALID_ARGUMENT
