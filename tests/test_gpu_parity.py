"""GPU parity tests: the CUDA path (called through the C ABI) against the oracle on the same
seeded inputs, against the committed golden vectors of the reference, and -- at the full
BASELINE frame size -- through size-independent properties.

Tolerances (BASELINE.json north_star):
  * lat / lon / MLat / MLT: |delta| <= 1e-9 degrees (MLT: the same bound in degrees of SM
    longitude, i.e. 1e-9/15 hours), excluding a *reported* count of grazing rays
    (normalised discriminant < 1e-10) where 1 ulp of input noise exceeds the bound;
  * masks (NaN pattern), grid-cell indices, per-cell counts and integer channel sums:
    bit-exact when both sides bin the same coordinates;
  * resampled means: <= 1e-6 relative (integer channels are in fact exact).
"""
import contextlib
import io
import os

import numpy as np
import numpy.ma as ma
import pytest

pytestmark = pytest.mark.gpu

TOL_DEG = 1e-9
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; the CUDA path has no CPU fallback")
    from auromat_b200.runtime import get_context
    return get_context(0)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def oracle_frame(hdr, fast):
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    t, cam = synthetic.headerTimeAndCamera(hdr)
    g = O.georeference(hdr, cam, t, 110, fast_center=fast)
    if not fast:
        mk, mc = O.sanitize_masks(np.isnan(g['lats']), np.isnan(g['latsCenter']))
        for n in ('lats', 'lons', 'mlat', 'mlt'):
            g[n][mk] = np.nan
        for n in ('latsCenter', 'lonsCenter', 'mlatCenter', 'mltCenter', 'elevation'):
            g[n][mc] = np.nan
    return g


def gpu_arrays(m):
    return dict(lats=m.lats, lons=m.lons, latsCenter=m.latsCenter, lonsCenter=m.lonsCenter,
                elevation=m.elevation, mlat=m.mLatMlt[0], mlt=m.mLatMlt[1],
                mlatCenter=m.mLatMltCenter[0], mltCenter=m.mLatMltCenter[1])


def assert_coords_close(gpu, ref, allow=0):
    worst = {}
    for name, arr in gpu.items():
        a, b = arr.filled(np.nan), ref[name]
        assert a.shape == b.shape, name
        assert np.array_equal(np.isnan(a), np.isnan(b)), "mask mismatch in " + name
        tol = TOL_DEG / 15 if name.startswith('mlt') else TOL_DEG
        if name == 'elevation':
            # acos(dot) is ill-conditioned towards nadir (reference mapping/astrometry.py:205-209):
            # d(elev) = d(dot)/sqrt(1-dot^2); allow 8 ulp of dot on top of the base tolerance
            s = np.sin(np.deg2rad(np.clip(90 - b, 1e-9, None)))
            tol = TOL_DEG + np.rad2deg(8 * 2.2e-16 / s)
        bad = np.abs(a - b) > tol
        worst[name] = float(np.nanmax(np.abs(a - b)))
        assert np.nansum(bad) <= allow, "%s: %d values beyond tolerance, max %.3e" % (name, np.nansum(bad), worst[name])
    return worst


@pytest.mark.parametrize("fast", [False, True])
def test_georeference_vs_oracle(env, fast):
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    hdr = synthetic.issHeader(532, 354)
    m = getMapping(synthetic.issImage(532, 354), hdr, fastCenterCalculation=fast, identifier='t')
    with quiet():
        g = oracle_frame(hdr, fast)
    worst = assert_coords_close(gpu_arrays(m), g, allow=m.illConditionedCount)
    assert m.illConditionedCount < 5
    assert np.sum(~np.isnan(g['latsCenter'])) > 100000
    m.checkGuarantees()
    print("worst |delta| (deg):", worst)


@pytest.mark.parametrize("fast", [0, 1])
def test_georeference_vs_reference_golden(env, fast):
    """Against outputs of the reference itself (tests/golden, oracle/gen_golden.py)."""
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    g = dict(np.load(os.path.join(GOLDEN, "iss_frame_133x89_fast%d.npz" % fast)))
    hdr = synthetic.issHeader(133, 89)
    m = getMapping(synthetic.issImage(133, 89), hdr, fastCenterCalculation=bool(fast), identifier='t', nosanitize=True)
    assert_coords_close(gpu_arrays(m), g, allow=m.illConditionedCount)


def test_georeference_sip_vs_oracle(env):
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    hdr = synthetic.issHeader(600, 400, sipOrder=4)
    m = getMapping(synthetic.issImage(600, 400), hdr, identifier='t')
    with quiet():
        g = oracle_frame(hdr, False)
        g0 = oracle_frame(synthetic.issHeader(600, 400), False)
    assert_coords_close(gpu_arrays(m), g, allow=m.illConditionedCount)
    # the distortion really does something (~20 px at the corners)
    assert np.nanmax(np.abs(g['latsCenter'] - g0['latsCenter'])) > 1e-3


def test_camera_inside_ellipsoid_and_no_hits(env):
    """Directed intersection with the origin inside the inflated ellipsoid
    (intersection.py:85-88) and a frame that sees no Earth at all (everything masked)."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    hdr = synthetic.issHeader(200, 120)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    low = dict(hdr)
    scale = (O.WGS84_A + 50) / np.linalg.norm(cam)         # 50 km altitude: inside the 110 km shell
    low['POSXSHIF'], low['POSYSHIF'], low['POSZSHIF'] = (float(v) for v in cam * scale)
    m = getMapping(synthetic.issImage(200, 120), low, identifier='t', nosanitize=True)
    with quiet():
        g = O.georeference(low, cam * scale, t, 110)
    assert np.all(~np.isnan(g['lats']))                    # from inside, every ray hits
    assert_coords_close(gpu_arrays(m), g, allow=m.illConditionedCount)
    away = dict(hdr)
    away['CRVAL1'], away['CRVAL2'] = hdr['CRVAL1'] + 180.0, -hdr['CRVAL2']   # look away from Earth
    m2 = getMapping(synthetic.issImage(200, 120), away, identifier='t')
    assert ma.getmaskarray(m2.lats).all() and ma.getmaskarray(m2.latsCenter).all()
    with pytest.raises(ValueError):
        m2.boundingBox


def _grid_for(lats_c, lons_c, ppd, pad=0.0):
    from auromat_b200.resample import targetGrid
    return targetGrid(ppd, np.nanmin(lats_c) - pad, np.nanmax(lats_c) + pad, np.nanmin(lons_c) - pad,
                      np.nanmax(lons_c) + pad)


def test_cell_indices_bit_exact_incl_edges(env):
    """searchsorted(linspace, x, 'right') semantics incl. samples exactly on / 1 ulp beside
    bin edges, the right-most-edge pull-in and outliers (util/histogram.py:205-224)."""
    import oracle.auromat_oracle as O
    from auromat_b200.resample import targetGrid
    ctx = env
    rng = np.random.default_rng(11)
    grid, info = targetGrid((36.0, 20.7), 48.0, 61.0, -111.0, -92.0)
    ex = np.linspace(grid.lo_x, grid.hi_x, grid.nx + 1)
    ey = np.linspace(grid.lo_y, grid.hi_y, grid.ny + 1)
    n = 200000
    lon = rng.uniform(grid.lo_x - 0.5, grid.hi_x + 0.5, n)
    lat = rng.uniform(grid.lo_y - 0.5, grid.hi_y + 0.5, n)
    k = 0
    for e, arr in ((ex, lon), (ey, lat)):
        for shift in (0, 1, -1):
            v = e.copy()
            for _ in range(abs(shift)):
                v = np.nextafter(v, np.inf if shift > 0 else -np.inf)
            arr[k:k + len(v)] = v
            k += len(v)
    lon[k:k + 50] = grid.hi_x + rng.uniform(0, 2e-7, 50)       # around(x, decimal) == around(hi, decimal)
    lat[k + 50:k + 100] = grid.hi_y + rng.uniform(0, 2e-7, 50)
    lat[k + 100:k + 120] = np.nan
    ix, iy = ctx.cell_indices(ctx.to_device(lat), ctx.to_device(lon), grid)
    ok = ~np.isnan(lat)
    oix, oiy = O.cell_indices(lon[ok], lat[ok], (grid.nx, grid.ny), [[grid.lo_x, grid.hi_x], [grid.lo_y, grid.hi_y]])
    gx, gy = ix.cpu().numpy(), iy.cpu().numpy()
    assert np.array_equal(gx[ok], oix)
    assert np.array_equal(gy[ok], oiy)
    assert np.all(gx[~ok] == -1) and np.all(gy[~ok] == -1)
    assert (oix >= 0).sum() > n // 2


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("fast", [False, True])
def test_resample_bit_exact_on_same_coordinates(env, dtype, fast):
    """Feed the GPU's own lat/lon to the oracle's histogram: counts, integer sums, rounded
    means and masks must be identical; elevation means within 1e-6 relative."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample, resampleToDevice
    hdr = synthetic.issHeader(532, 354)
    img = synthetic.issImage(532, 354, dtype=dtype)
    m = getMapping(img, hdr, fastCenterCalculation=fast, identifier='t')
    geo = {k: v.filled(np.nan) for k, v in gpu_arrays(m).items()}
    for ppd in [(36.0, 20.7), 5, (11.3, 97.0)]:
        r = resample(m, pxPerDeg=ppd)
        o = O.resample_frame(geo, img, 110, px_per_deg=ppd, return_count=True)
        grid, info, _, _, _ = resampleToDevice(m, pxPerDeg=ppd)
        cnt = info['count'].cpu().numpy().reshape(grid.ny, grid.nx)
        assert np.array_equal(cnt, o['count'])
        assert r.img.dtype == dtype
        # reference GenericMapping semantics: img masked where count == 0
        assert np.array_equal(ma.getmaskarray(r.img), o['img_mask'])
        assert np.array_equal(r.img.filled(0), np.where(o['img_mask'], 0, o['img']))
        e, oe = r.elevation.filled(np.nan), o['elevation']
        assert np.array_equal(np.isnan(e), np.isnan(oe))
        assert np.nanmax(np.abs(e - oe) / np.abs(oe)) < 1e-6
        # grid coordinates: bit-identical to numpy linspace/meshgrid where defined
        for name in ('lats', 'lons', 'latsCenter', 'lonsCenter'):
            a = getattr(r, name)
            assert np.array_equal(a.data[~ma.getmaskarray(a)], o[name][~ma.getmaskarray(a)]), name
        r.checkPlateCarree()
        r.checkGuarantees()


def test_resample_vs_pure_oracle_chain(env):
    """GPU chain vs oracle chain end to end: coordinates differ in the last bits, so cell
    membership may differ only for samples within 1 ulp-ish of a bin edge (reported)."""
    import oracle.auromat_oracle as O
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resampleToDevice, targetGrid
    hdr = synthetic.issHeader(532, 354)
    img = synthetic.issImage(532, 354)
    m = getMapping(img, hdr, identifier='t')
    with quiet():
        g = oracle_frame(hdr, False)
    ppd = (36.0, 20.7)
    o = O.resample_frame(g, img, 110, px_per_deg=ppd, return_count=True)
    grid, info, outImg, outMask, outElev = resampleToDevice(m, pxPerDeg=ppd)
    cnt = info['count'].cpu().numpy().reshape(grid.ny, grid.nx)
    assert cnt.shape == o['count'].shape
    moved = int(np.abs(cnt - o['count']).sum())
    # how many GPU samples sit within 1 ulp of an edge
    ctx = env
    p = m.devicePlanes()
    near = torch.zeros(1, dtype=torch.int64, device=ctx.torch_device)
    c2, s2 = torch.zeros_like(info['count']), ctx.zeros(3 * grid.nx * grid.ny, torch.int64)
    ctx.bin_accumulate(p['lat_c'], p['lon_c'], None, m.deviceImage(), grid, c2, s2, None, near)
    n_near = int(near.item())
    print("cells with different counts: %d samples moved, %d samples within 1 ulp of an edge" % (moved, n_near))
    assert moved <= 2 * max(n_near, 4)
    same = cnt == o['count']
    gi = outImg.cpu().numpy()
    assert np.array_equal(gi[same & (cnt > 0)], o['img'][same & (cnt > 0)])
    rel = np.abs(gi.astype(float) - o['img']) / np.maximum(o['img'], 1)
    assert np.all(rel[(cnt > 0) & (o['count'] > 0)] <= 1.0)     # moved samples change a mean by < 1 count unit


def _synthetic_generic(mode, h=60, w=80):
    import oracle.auromat_oracle as O
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:h + 1, 0:w + 1].astype(float)
    if mode == "plain":
        lats, lons = 70 - yy * 0.11 - xx * 0.01, -100 + xx * 0.2 + yy * 0.02
    elif mode == "discontinuity":
        lats, lons = 70 - yy * 0.11, O.wrap_at_180(170 + xx * 0.25)
    else:
        r = 1 + np.hypot(yy - h / 2 + 0.25, xx - w / 2 + 0.25) * 0.12
        lats, lons = 90 - r, np.rad2deg(np.arctan2(yy - h / 2 + 0.25, xx - w / 2 + 0.25))
    yc, xc = np.mgrid[0:h, 0:w].astype(float) + 0.5
    if mode == "plain":
        latsC, lonsC = 70 - yc * 0.11 - xc * 0.01, -100 + xc * 0.2 + yc * 0.02
    elif mode == "discontinuity":
        latsC, lonsC = 70 - yc * 0.11, O.wrap_at_180(170 + xc * 0.25)
    else:
        r = 1 + np.hypot(yc - h / 2 + 0.25, xc - w / 2 + 0.25) * 0.12
        latsC, lonsC = 90 - r, np.rad2deg(np.arctan2(yc - h / 2 + 0.25, xc - w / 2 + 0.25))
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elev = rng.uniform(0, 90, (h, w))
    return lats, lons, latsC, lonsC, elev, img


@pytest.mark.parametrize("mode", ["plain", "discontinuity", "pole"])
def test_generic_mapping_resample_branches(env, mode):
    """GenericMapping (uploaded arrays) through the pole-rotation and date-line branches of
    resample.py:176-218,262-277 against the oracle (which is pinned bit-for-bit to the
    reference for these branches)."""
    import datetime
    import oracle.auromat_oracle as O
    from auromat_b200.mapping.mapping import GenericMapping
    from auromat_b200.resample import resample, resampleToDevice
    lats, lons, latsC, lonsC, elev, img = _synthetic_generic(mode)
    m = GenericMapping(lats, lons, latsC, lonsC, elev, 110, img, np.zeros(3), datetime.datetime(2012, 1, 1), 'g')
    assert m.containsPole == (mode == "pole")
    assert m.containsDiscontinuity == (mode != "plain")
    bb = m.boundingBox
    outline = np.transpose([lats.ravel(), lons.ravel()])
    b = O.boundary_corner_mask(np.ones_like(lats, bool))
    outline = np.transpose([lats[b], lons[b]])
    ob = O.bounding_box(lats, lons, contains_pole=(mode == "pole"))
    assert tuple(ob) == (bb.latSouth, bb.lonWest, bb.latNorth, bb.lonEast)
    data = np.dstack((img.astype(float), elev))
    ppd = (6.0, 3.0)
    o = O.resample_grid(latsC, lonsC, 110, data, ob, ppd, contains_discontinuity=(mode != "plain"),
                        contains_pole=(mode == "pole"), outline_latlon=outline, return_count=True)
    grid, info, outImg, outMask, outElev = resampleToDevice(m, pxPerDeg=ppd)
    cnt = info['count'].cpu().numpy().reshape(grid.ny, grid.nx)
    if mode == "pole":
        # rotatePole runs through sin/cos/atan on both sides -> last-bit differences may move
        # a sample across an edge; everything else is exact
        assert np.abs(cnt - o[5]).sum() <= 4
    else:
        assert np.array_equal(cnt, o[5])
        oi = np.where(np.isnan(o[4][:, :, :3]), 0, np.round(o[4][:, :, :3])).astype(np.uint8)
        assert np.array_equal(outImg.cpu().numpy(), oi)
    r = resample(m, pxPerDeg=ppd)
    tol = 1e-9
    for a, b_ in ((r.lats, o[0]), (r.lons, o[1]), (r.latsCenter, o[2]), (r.lonsCenter, o[3])):
        ok = ~ma.getmaskarray(a)
        d = np.abs(a.data[ok] - b_[ok])
        d = np.minimum(d, 360 - d)
        assert d.max() <= tol


@pytest.mark.parametrize("kernel,h,w", [("tile", 130, 170), ("registers", 130, 170), ("tile", 61, 1500), ("tile", 3, 5)])
def test_sanitize_matches_oracle(env, kernel, h, w, monkeypatch):
    """_doSanitize on the bitmaps against the oracle (itself bit-equal to the reference): the shared-memory
    tile kernel (several row tiles, partial last tile, a frame narrower than a word) and the register-only
    kernel that very wide frames fall back to."""
    import torch
    import oracle.auromat_oracle as O
    ctx = env
    if kernel == "registers":
        monkeypatch.setenv("AMT_SANITIZE_REGISTERS", "1")
    else:
        monkeypatch.delenv("AMT_SANITIZE_REGISTERS", raising=False)
    rng = np.random.default_rng(5)
    lat_k = rng.uniform(0, 1, (h + 1, w + 1))
    lat_k[rng.uniform(size=lat_k.shape) < 0.2] = np.nan
    lat_c = rng.uniform(0, 1, (h, w))
    lat_c[rng.uniform(size=lat_c.shape) < 0.3] = np.nan
    planes = dict(lat_k=ctx.to_device(lat_k.ravel()), lon_k=ctx.to_device(lat_k.ravel().copy()),
                  lat_c=ctx.to_device(lat_c.ravel()), elev_c=ctx.to_device(lat_c.ravel().copy()))
    ctx.valid_bits(w, h, planes)
    ctx.sanitize(w, h, planes)
    mk, mc = O.sanitize_masks(np.isnan(lat_k), np.isnan(lat_c))
    for n in ('lat_k', 'lon_k'):
        assert np.array_equal(np.isnan(planes[n].cpu().numpy().reshape(h + 1, w + 1)), mk)
    for n in ('lat_c', 'elev_c'):
        assert np.array_equal(np.isnan(planes[n].cpu().numpy().reshape(h, w)), mc)
    # the bitmaps agree with the planes
    wk, wc = (w + 1 + 31) // 32, (w + 31) // 32
    bk = planes['valid_k'].cpu().numpy().view(np.uint32).reshape(h + 1, wk)
    bc = planes['valid_c'].cpu().numpy().view(np.uint32).reshape(h, wc)
    unpack = lambda b, n: ((b[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(b.shape[0], -1)[:, :n]
    assert np.array_equal(unpack(bk, w + 1).astype(bool), ~mk)
    assert np.array_equal(unpack(bc, w).astype(bool), ~mc)


def test_masked_by_elevation_and_guarantees(env):
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample
    hdr = synthetic.issHeader(532, 354)
    m = getMapping(synthetic.issImage(532, 354), hdr, fastCenterCalculation=True, identifier='t')
    m.checkGuarantees()
    m10 = m.maskedByElevation(10)
    m10.checkGuarantees()
    e = m.elevation
    expect = (e < 10).filled(True)
    assert np.array_equal(ma.getmaskarray(m10.latsCenter), expect)
    assert np.array_equal(ma.getmaskarray(m10.img)[:, :, 0], expect)
    assert 0 <= e.min() and e.max() <= 90                 # reference test/elevation_test.py:12-21
    r = resample(m10, arcsecPerPx=100, method='mean')
    r.checkGuarantees()
    r.checkPlateCarree()
    with pytest.raises(ValueError):
        m.maskedByElevation(91)


def test_latlon_to_mlatmlt_generic_route(env):
    """BaseMapping._mLatMlt route for non-astrometry mappings (mapping.py:540-550)."""
    import datetime
    import oracle.auromat_oracle as O
    from auromat_b200.mapping.mapping import GenericMapping
    lats, lons, latsC, lonsC, elev, img = _synthetic_generic("plain")
    t = datetime.datetime(2012, 1, 25, 9, 26, 55)
    m = GenericMapping(lats, lons, latsC, lonsC, elev, 110, img, np.zeros(3), t, 'g')
    mlat, mlt = m.mLatMlt
    omlat, omlt = O.latlon_to_mlat_mlt(lats, lons, 110, t)
    assert np.max(np.abs(mlat.filled(np.nan) - omlat)) <= TOL_DEG
    assert np.max(np.abs(mlt.filled(np.nan) - omlt)) <= TOL_DEG / 15
    mlatc, mltc = m.mLatMltCenter
    omlat, omlt = O.latlon_to_mlat_mlt(latsC, lonsC, 110, t)
    assert np.max(np.abs(mlatc.filled(np.nan) - omlat)) <= TOL_DEG


def test_error_conventions(env):
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample
    hdr = synthetic.issHeader(64, 48)
    m = getMapping(synthetic.issImage(64, 48), hdr, identifier='t')
    with pytest.raises(NotImplementedError):
        resample(m, method='median')
    with pytest.raises(NotImplementedError):
        resample(m, method='nearest')
    with pytest.raises(ValueError):
        resample("not a mapping")
    bad = dict(hdr)
    del bad['DATE-OBS']
    with pytest.raises(ValueError):
        getMapping(synthetic.issImage(64, 48), bad)
    car = dict(hdr)
    car['CTYPE1'], car['CTYPE2'] = 'RA---CAR', 'DEC--CAR'
    with pytest.raises(NotImplementedError):
        getMapping(synthetic.issImage(64, 48), car, identifier='t').lats
    late = dict(hdr)
    late['DATE-OBS'] = '2021-03-01T00:00:00.000000'      # IGRF table ends 2020 (igrf.py:55-58)
    with pytest.raises(ValueError):
        getMapping(synthetic.issImage(64, 48), late, identifier='t').lats


# --------------------------------------------------------------- full BASELINE size
@pytest.fixture(scope="module")
def full_frame(env):
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    hdr = synthetic.issHeader()
    img = synthetic.issImage()
    m = getMapping(img, hdr, identifier='full')
    m.prefetch(magnetic=True)
    return hdr, img, m


def test_full_size_properties(env, full_frame):
    """4256x2832 (configs[1]): determinism, conservation (sum of counts == valid centres in
    range == checksum of checksums), linearity of the channel sums, mosaic additivity."""
    import torch
    from auromat_b200.resample import resampleToDevice
    ctx = env
    hdr, img, m = full_frame
    s = m._deviceStats()
    assert 6_900_000 < s.n_valid_centers < 7_100_000         # SURVEY: 7.03 M of 12.05 M centres hit
    bb = m.boundingBox
    assert 47.8 < bb.latSouth < 48.0 and 61.2 < bb.latNorth < 61.4
    assert -111.8 < bb.lonWest < -111.5 and -92.0 < bb.lonEast < -91.8
    grid, info, outImg, outMask, outElev = resampleToDevice(m, arcsecPerPx=100)
    assert (grid.ny, grid.nx) == (482, 410) or abs(grid.ny - 482) <= 2
    count = info['count'].clone()
    p = m.devicePlanes()
    dimg = m.deviceImage()
    cells = grid.nx * grid.ny
    # determinism: integer accumulators are order independent -> identical bits on a re-run
    c2, s2 = ctx.zeros(cells, torch.int64), ctx.zeros(3 * cells, torch.int64)
    ctx.bin_accumulate(p['lat_c'], p['lon_c'], None, dimg, grid, c2, s2, None)
    c3, s3 = ctx.zeros(cells, torch.int64), ctx.zeros(3 * cells, torch.int64)
    ctx.bin_accumulate(p['lat_c'], p['lon_c'], None, dimg, grid, c3, s3, None)
    assert torch.equal(c2, count) and torch.equal(c2, c3) and torch.equal(s2, s3)
    # conservation: every valid centre inside the grid range is counted exactly once
    ix, iy = ctx.cell_indices(p['lat_c'], p['lon_c'], grid)
    inside = (ix >= 0) & (iy >= 0)
    assert int(c2.sum().item()) == int(inside.sum().item())
    assert int(inside.sum().item()) <= s.n_valid_centers
    # checksum of checksums: per-channel totals equal the masked image totals
    flat = dimg.reshape(-1, 3).to(torch.int64)
    for c in range(3):
        assert int(s2[c * cells:(c + 1) * cells].sum().item()) == int(flat[inside, c].sum().item())
    # linearity: binning 255 - img gives 255*count - sums
    inv = (255 - dimg)
    c4, s4 = ctx.zeros(cells, torch.int64), ctx.zeros(3 * cells, torch.int64)
    ctx.bin_accumulate(p['lat_c'], p['lon_c'], None, inv, grid, c4, s4, None)
    assert torch.equal(s4, 255 * c2.repeat(3) - s2)
    # additivity (mosaic building block): accumulating twice doubles everything
    ctx.bin_accumulate(p['lat_c'], p['lon_c'], None, dimg, grid, c2, s2, None)
    assert torch.equal(c2, 2 * c3) and torch.equal(s2, 2 * s3)
    # means of a doubled accumulation are unchanged
    o1, m1, _ = ctx.normalise(grid, dimg.dtype, 3, c3, s3, None)
    o2, m2, _ = ctx.normalise(grid, dimg.dtype, 3, c2, s2, None)
    assert torch.equal(o1, o2) and torch.equal(m1, m2) and torch.equal(o1, outImg)


def test_full_size_sampled_rows_vs_oracle(env, full_frame):
    """Oracle on a band of rows of the full-size frame (the oracle needs ~20 s and several GB
    for the whole frame): rows 1400..1447 of corners/centres, unsanitised coordinates."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    hdr, img, m = full_frame
    t, cam = synthetic.headerTimeAndCamera(hdr)
    y0, nrow = 1400, 48
    W = hdr['IMAGEW']
    # oracle on a sub-rectangle: shift CRPIX2 so that row 0 of the band is row y0 of the frame
    band = dict(hdr)
    band['IMAGEH'] = nrow
    band['CRPIX2'] = hdr['CRPIX2'] - y0
    with quiet():
        g = O.georeference(band, cam, t, 110)
    lat = m.lats.filled(np.nan)[y0:y0 + nrow + 1]
    latc = m.latsCenter.filled(np.nan)[y0:y0 + nrow]
    mlt = m.mLatMltCenter[1].filled(np.nan)[y0:y0 + nrow]
    # interior rows of the band are not touched by sanitisation differences
    for a, b, tol in ((lat[1:-1], g['lats'][1:-1], TOL_DEG), (latc[1:-1], g['latsCenter'][1:-1], TOL_DEG),
                      (mlt[1:-1], g['mltCenter'][1:-1], TOL_DEG / 15)):
        both = ~np.isnan(a) & ~np.isnan(b)
        assert both.sum() > 0.5 * a.size
        assert np.sum(np.isnan(a) != np.isnan(b)) <= 2 * (nrow + 1)    # sanitised limb pixels only
        assert np.max(np.abs(a[both] - b[both])) <= tol
    assert W == 4256


# --------------------------------------------------------------- all-sky stations + mosaic
def _stations(n, seed=2):
    from auromat_b200.mapping.allsky import CalibrationData
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        out.append(CalibrationData('S%02d' % i, 2011.5, 2012.5, float(rng.uniform(55, 70)),
                                   float(rng.uniform(-160, -60)), 256.0 + float(rng.uniform(-20, 20)),
                                   256.0 + float(rng.uniform(-20, 20)), 155.81, float(rng.uniform(-0.2, 0.2)), None))
    return out


def test_allsky_vs_oracle_and_reference_golden(env):
    import datetime
    import oracle.auromat_oracle as O
    from auromat_b200.mapping.allsky import AllSkyMapping, CalibrationData
    t = datetime.datetime(2012, 3, 4, 17, 19, 0)
    g = np.load(os.path.join(GOLDEN, "allsky_SOD_96.npz"))
    lat, lon, xc, yc, k, rot = g['cal']
    cal = CalibrationData('SOD', 2011.5, 2012.5, lat, lon, xc, yc, k, rot, None)
    img = np.zeros((96, 96, 1), np.uint16)
    m = AllSkyMapping(cal, img, t, 110, sanitize=False)
    for name, arr in (('lats', m.lats), ('lons', m.lons), ('latsCenter', m.latsCenter),
                      ('lonsCenter', m.lonsCenter), ('elevation', m.elevation)):
        a = arr.filled(np.nan)
        assert np.array_equal(np.isnan(a), np.isnan(g[name])), name
        d = np.abs(a - g[name])
        if name.startswith('lon'):
            d = np.minimum(d, 360 - d)
        assert np.nanmax(d) <= TOL_DEG, (name, np.nanmax(d))
    # larger frame against the oracle, incl. the generic MLat/MLT route
    cal2 = _stations(1)[0]
    w = 256
    m2 = AllSkyMapping(cal2, np.zeros((w, w, 1), np.uint16), t, 110, sanitize=False)
    s = w / 512
    o = O.allsky_georeference(w, cal2.xc * s, cal2.yc * s, cal2.k * s, cal2.rotation, cal2.lat, cal2.lon, 110)
    for name, arr in (('lats', m2.lats), ('latsCenter', m2.latsCenter), ('lonsCenter', m2.lonsCenter),
                      ('elevation', m2.elevation)):
        d = np.abs(arr.filled(np.nan) - o[name])
        assert np.nanmax(np.minimum(d, 360 - d)) <= TOL_DEG, name
    omlat, omlt = O.latlon_to_mlat_mlt(o['latsCenter'], o['lonsCenter'], 110, t)
    mlat, mlt = m2.mLatMltCenter
    assert np.nanmax(np.abs(mlat.filled(np.nan) - omlat)) <= TOL_DEG
    d = np.abs(mlt.filled(np.nan) - omlt)
    assert np.nanmax(np.minimum(d, 24 - d)) <= TOL_DEG / 15


def test_mosaic_single_process_vs_oracle(env):
    """Config-5 style: several all-sky stations binned into ONE common grid; the serial oracle
    is the sum of per-station histogram2d outputs, then one division."""
    import datetime
    import oracle.auromat_oracle as O
    from auromat_b200 import parallel
    from auromat_b200.mapping.allsky import getMappingCollection
    t = datetime.datetime(2012, 3, 4, 17, 19, 0)
    cals = _stations(6)
    rng = np.random.default_rng(9)
    w = 128
    imgs = [rng.integers(0, 65536, (w, w, 1), dtype=np.uint16) for _ in cals]
    coll = getMappingCollection(imgs, cals, t, 110, minElevation=1)
    assert len(coll) == 6
    mosaic, acc = parallel.mosaic(coll.mappings, pxPerDeg=(20, 20))
    grid = acc.grid
    cnt = acc.count.cpu().numpy().reshape(grid.ny, grid.nx)
    sums = acc.sums.cpu().numpy().reshape(grid.ny, grid.nx)
    bins, rng_ = (grid.nx, grid.ny), [[grid.lo_x, grid.hi_x], [grid.lo_y, grid.hi_y]]
    tc = np.zeros((grid.ny, grid.nx))
    ts = np.zeros((grid.ny, grid.nx))
    te = np.zeros((grid.ny, grid.nx))
    for m, im in zip(coll.mappings, imgs):
        la, lo = m.latsCenter.filled(np.nan).ravel(), m.lonsCenter.filled(np.nan).ravel()
        ok = ~np.isnan(la)
        e = m.elevation.filled(np.nan).ravel()
        c, s, ee = O.histogram2d_weighted(lo[ok], la[ok], bins, rng_, [None, im.reshape(-1)[ok].astype(float), e[ok]])
        tc += c.T[::-1]
        ts += s.T[::-1]
        te += ee.T[::-1]
    assert np.array_equal(cnt, tc) and np.array_equal(sums, ts)
    assert tc.sum() > 30000 and (tc > 0).mean() > 0.02
    with np.errstate(invalid='ignore', divide='ignore'):
        mean = np.round(ts / tc)
    got = mosaic.img
    assert np.array_equal(ma.getmaskarray(got)[:, :, 0], tc == 0)
    assert np.array_equal(got.filled(0)[:, :, 0][tc > 0], mean[tc > 0].astype(np.uint16))
    ge = mosaic.elevation.filled(np.nan)
    with np.errstate(invalid='ignore', divide='ignore'):
        assert np.nanmax(np.abs(ge - te / tc) / (te / tc)) < 1e-6
    mosaic.checkPlateCarree()
    # maskedByElevation(1) of every station really removed the below-horizon fisheye corners
    for m in coll.mappings:
        assert m.elevation.min() >= 1


def test_pipelined_sequence_equals_frame_by_frame(env):
    """auromat_b200.pipeline.resampleSequence == resample(getMapping(...)) for every frame,
    with host (pinned) inputs, device inputs, and uint16 images."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence
    from auromat_b200.resample import resample
    W, H, n = 300, 200, 5
    hdrs = synthetic.sequenceHeaders(n, W, H)
    for dtype in (np.uint8, np.uint16):
        imgs = [synthetic.issImage(W, H, seed=40 + i, dtype=dtype) for i in range(n)]
        expect = [resample(getMapping(im, h, identifier='x'), arcsecPerPx=400) for im, h in zip(imgs, hdrs)]
        for source in ('host', 'device', 'host-ring', 'device-ring', 'host-full', 'host-ring-full'):
            src = imgs if source.startswith('host') else [env.to_device(im) for im in imgs]
            # ring mode = fixed plane ring + two-stream overlap; few frames, so all stay valid
            tr = {}
            got = list(resampleSequence(src, hdrs, arcsecPerPx=400, magnetic=True, ringBuffers='ring' in source,
                                        sparseUpload='full' not in source, transferStats=tr))
            assert len(got) == n
            full = sum(im.nbytes for im in imgs)
            if source.startswith('host'):
                # rows without a georeferenced pixel (above the limb, ~40 %) are not uploaded
                rows = sum(int(f.mapping._deviceStats().row_max_c - f.mapping._deviceStats().row_min_c + 1) for f in got)
                assert tr['h2d_bytes'] == (full if 'full' in source else rows * imgs[0][0].nbytes)
                assert 'full' in source or 0.4 * full < tr['h2d_bytes'] < 0.8 * full
                assert np.array_equal(got[0].mapping.img_unmasked, imgs[0])     # the mapping keeps the complete host image
            else:
                assert tr['h2d_bytes'] == 0
            for f, e in zip(got, expect):
                assert np.array_equal(ma.getmaskarray(f.img), ma.getmaskarray(e.img))
                assert np.array_equal(f.img.filled(0), e.img.filled(0))
                fe, ee = f.elevation.filled(np.nan), e.elevation.filled(np.nan)
                assert np.array_equal(np.isnan(fe), np.isnan(ee))
                assert np.nanmax(np.abs(fe - ee)) < 1e-9
                assert 'mlat_k' in f.mapping._planes
            m = got[2].toMapping()
            assert np.array_equal(m.lats.data, expect[2].lats.data)
            assert np.array_equal(ma.getmaskarray(m.latsCenter), ma.getmaskarray(expect[2].latsCenter))
    torch.cuda.synchronize()


def test_intersects_earth_and_consistency(env):
    """BaseSpacecraftMapping.intersectsEarth / isConsistent (reference spacecraft.py:508-555,
    intersection.py:165-199) from the hit ballots of the georeference kernel."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    hdr = synthetic.issHeader(532, 354)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    m = getMapping(synthetic.issImage(532, 354), hdr, identifier='t')
    d = O.pix2dir(hdr, corner=False)
    expect = O.ellipsoid_line_intersects(O.WGS84_A, O.WGS84_B, cam, d.reshape(-1, 3)).reshape(354, 532)
    got = m.intersectsEarth
    assert got.shape == expect.shape and got.dtype == bool
    assert np.sum(got != expect) <= 2            # knife-edge rays only
    assert 0.3 < got.mean() < 0.8
    # the un-inflated Earth is hit by fewer rays than the 110 km shell
    assert got.sum() < (~ma.getmaskarray(getMapping(synthetic.issImage(532, 354), hdr, identifier='t',
                                                    nosanitize=True).latsCenter)).sum()
    assert m.isConsistent()
    stars_on_earth = np.argwhere(got)[:3][:, ::-1]
    assert not m.isConsistent(stars_on_earth)
    assert m.isConsistent(np.argwhere(~got)[:3][:, ::-1])


def test_sm_roundtrip_and_resample_mlat_mlt(env):
    """convertMappingToSM / resample / convertSMMappingToGeo (reference mapping.py:1519-1559,
    resample.py:63-71, transform.py:461-485)."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.mapping import convertMappingToSM, convertSMMappingToGeo
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resampleMLatMLT
    hdr = synthetic.issHeader(300, 200)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    m = getMapping(synthetic.issImage(300, 200), hdr, fastCenterCalculation=True, identifier='t')
    sm = convertMappingToSM(m)
    mlat, mlt = m.mLatMlt
    assert np.array_equal(sm.lats.filled(np.nan), mlat.filled(np.nan), equal_nan=True)
    smlon = O.mlt_to_smlon(mlt.filled(np.nan))
    assert np.nanmax(np.abs(sm.lons.filled(np.nan) - smlon)) < 1e-12
    # smToLatLon on the device against the oracle
    geo = convertSMMappingToGeo(sm)
    ok = ~np.isnan(mlat.filled(np.nan))
    olat, olon = O.sm_to_latlon(mlat.filled(np.nan)[ok], smlon[ok], t)
    assert np.max(np.abs(geo.lats.data[ok] - olat)) <= TOL_DEG
    d = np.abs(geo.lons.data[ok] - olon)
    assert np.max(np.minimum(d, 360 - d)) <= TOL_DEG
    r = resampleMLatMLT(m, arcsecPerPx=400, method='mean')
    assert not r.isPlateCarree                       # regular in SM, not in geodetic coordinates
    assert r.img.shape[2] == 3 and (~ma.getmaskarray(r.img)).sum() > 1000
    r.checkGuarantees()


@pytest.mark.parametrize("W,H", [(1, 1), (2, 3), (31, 5), (32, 8), (33, 9), (64, 16), (65, 17), (257, 3)])
@pytest.mark.parametrize("fast", [False, True])
def test_ragged_and_tiny_frames(env, W, H, fast):
    """Frames smaller than a warp / tile, and sizes on and beside the 32-pixel word and 32x8
    tile boundaries: coordinates, masks and bitmaps against the oracle."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    hdr = synthetic.issHeader(W, H)
    # shift the reference pixel so that the tiny frame straddles the limb (mixed hit / miss)
    hdr['CRPIX2'] = hdr['CRPIX2'] - 0.18 * H
    img = synthetic.issImage(W, H)
    m = getMapping(img, hdr, fastCenterCalculation=fast, identifier='t')
    with quiet():
        g = oracle_frame(hdr, fast)
    assert_coords_close(gpu_arrays(m), g, allow=m.illConditionedCount)
    p = m.devicePlanes()
    wk, wc = (W + 1 + 31) // 32, (W + 31) // 32
    bk = p['valid_k'].cpu().numpy().view(np.uint32).reshape(H + 1, wk)
    bc = p['valid_c'].cpu().numpy().view(np.uint32).reshape(H, wc)
    unpack = lambda b, n: ((b[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(b.shape[0], -1)
    assert np.array_equal(unpack(bk, W + 1)[:, :W + 1].astype(bool), ~np.isnan(g['lats']))
    assert np.array_equal(unpack(bc, W)[:, :W].astype(bool), ~np.isnan(g['latsCenter']))
    assert not unpack(bk, W + 1)[:, W + 1:].any() and not unpack(bc, W)[:, W:].any()    # padding bits stay 0
    if (~np.isnan(g['latsCenter'])).any():
        m.checkGuarantees()


@pytest.mark.parametrize("channels", [1, 2, 4])
def test_channel_counts_and_gray_images(env, channels):
    """1-, 2- and 4-channel uint16 images through binning and normalisation."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample
    W, H = 200, 130
    hdr = synthetic.issHeader(W, H)
    rng = np.random.default_rng(channels)
    img = rng.integers(0, 65536, (H, W, channels), dtype=np.uint16)
    m = getMapping(img, hdr, fastCenterCalculation=True, identifier='t')
    r = resample(m, pxPerDeg=(9.0, 5.0))
    geo = {k: v.filled(np.nan) for k, v in gpu_arrays(m).items()}
    o = O.resample_frame(geo, img, 110, px_per_deg=(9.0, 5.0))
    assert r.img.shape == o['img'].shape and r.img.dtype == np.uint16
    assert np.array_equal(ma.getmaskarray(r.img), o['img_mask'])
    assert np.array_equal(r.img.filled(0), np.where(o['img_mask'], 0, o['img']))


@pytest.mark.parametrize("sip", [0, 3])
def test_plane_free_fused_path_equals_materialised(env, sip):
    """amt_bbox_stats_frame + amt_georef_bin_fused (no coordinate planes) give the same
    bounding box, grid, counts, integer sums and means as georeference -> planes -> binning."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence
    from auromat_b200.resample import resampleToDevice
    W, H = 532, 354
    hdr = synthetic.issHeader(W, H, sipOrder=sip)
    img = synthetic.issImage(W, H)
    a = getMapping(img, hdr, identifier='a')
    b = getMapping(img, hdr, identifier='b').setPlaneFree(True)
    ga, ia, imgA, maskA, elevA = resampleToDevice(a, arcsecPerPx=100)
    gb, ib, imgB, maskB, elevB = resampleToDevice(b, arcsecPerPx=100)
    assert 'lat_c' not in b._planes and 'lat_k' not in b._planes          # nothing was materialised
    assert a.boundingBox == b.boundingBox
    assert (ga.nx, ga.ny, ga.lo_x, ga.hi_x, ga.lo_y, ga.hi_y) == (gb.nx, gb.ny, gb.lo_x, gb.hi_x, gb.lo_y, gb.hi_y)
    assert torch.equal(ia['count'], ib['count'])
    assert torch.equal(imgA, imgB) and torch.equal(maskA, maskB)
    ea, eb = elevA.cpu().numpy(), elevB.cpu().numpy()
    assert np.array_equal(np.isnan(ea), np.isnan(eb))
    assert np.nanmax(np.abs(ea - eb)) < 1e-9
    s = b._deviceStats()
    assert (s.n_valid_corners, s.n_valid_centers) == (a._deviceStats().n_valid_corners, a._deviceStats().n_valid_centers)
    # coordinates are still available lazily
    assert np.array_equal(b.latsCenter.filled(np.nan), a.latsCenter.filled(np.nan), equal_nan=True)
    # the sequence API in plane-free mode
    frames = list(resampleSequence([img] * 3, [hdr] * 3, arcsecPerPx=100, coordinates=False))
    for f in frames:
        assert np.array_equal(f.img.filled(0), imgA.cpu().numpy() * ~maskA.cpu().numpy().astype(bool)[:, :, None])
        assert 'lat_c' not in f.mapping._planes


@pytest.mark.parametrize("case", ["pole", "dateline"])
def test_wcs_frames_over_the_pole_and_the_date_line(env, case):
    """ISS-style frames whose footprint contains the north pole / straddles the date line:
    detection (geometric pole test, longitude span), the rotatePole / wrap pre-rotations of the
    binning kernel and the outline statistics, against the oracle's resample (which is pinned
    bit-for-bit to the reference for these branches)."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample, resampleToDevice
    W, H = 300, 200
    if case == "pole":
        hdr = synthetic.issHeaderLookingAt(78.0, 20.0, 89.5, 120.0, W, H)
    else:
        hdr = synthetic.issHeaderLookingAt(52.0, 172.0, 58.0, -179.0, W, H)
    img = synthetic.issImage(W, H)
    m = getMapping(img, hdr, fastCenterCalculation=True, identifier=case)
    assert (~ma.getmaskarray(m.latsCenter)).sum() > 0.3 * W * H
    assert m.containsPole == (case == "pole")
    assert m.containsDiscontinuity
    bb = m.boundingBox
    lats, lons = m.lats.filled(np.nan), m.lons.filled(np.nan)
    ob = O.bounding_box(lats, lons, contains_pole=(case == "pole"))
    assert (bb.latSouth, bb.lonWest, bb.latNorth, bb.lonEast) == tuple(float(v) for v in ob)
    if case == "pole":
        assert bb.latNorth == 90 and (bb.lonWest, bb.lonEast) == (-180, 180)
    else:
        assert bb.lonWest > 0 > bb.lonEast
    # resample on the GPU vs the oracle fed with the GPU's own coordinates
    geo = dict(lats=lats, lons=lons, latsCenter=m.latsCenter.filled(np.nan), lonsCenter=m.lonsCenter.filled(np.nan),
               elevation=m.elevation.filled(np.nan))
    ppd = (4.0, 2.0) if case == "pole" else (6.0, 3.0)
    o = O.resample_frame(geo, img, 110, px_per_deg=ppd, contains_pole=(case == "pole"), return_count=True)
    grid, info, outImg, outMask, outElev = resampleToDevice(m, pxPerDeg=ppd)
    cnt = info['count'].cpu().numpy().reshape(grid.ny, grid.nx)
    assert cnt.shape == o['count'].shape
    if case == "pole":
        # rotatePole runs through sin/cos/atan on both sides: last-bit differences may move single samples
        assert np.abs(cnt - o['count']).sum() <= 6
    else:
        assert np.array_equal(cnt, o['count'])
        assert np.array_equal(outImg.cpu().numpy(), np.where(o['img_mask'], 0, o['img']))
    r = resample(m, pxPerDeg=ppd)
    for a, b in ((r.lats, o['lats']), (r.lons, o['lons']), (r.latsCenter, o['latsCenter']), (r.lonsCenter, o['lonsCenter'])):
        ok = ~ma.getmaskarray(a)
        d = np.abs(a.data[ok] - b[ok])
        assert np.minimum(d, 360 - d).max() <= 1e-9
    r.checkGuarantees()


def test_outline_and_centroid_golden_of_the_reference(env):
    """The reference's end-to-end known answer (test/outline_test.py:147-158): the centroid of
    `getMapping(ISS030-E-102170_dc.jpg, ISS030-E-102170_dc.wcs, fastCenterCalculation=True)` is
    (55.00295889563608, -99.21825084682715) to 6 decimals.  `synthetic.issHeader()` carries exactly the
    cards of that .wcs file (CRVAL, CRPIX, CD, POS*SHIF, DATE-OBS, DATESHIF; the image content
    does not enter)."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    m = getMapping(synthetic.issImage(), synthetic.issHeader(), fastCenterCalculation=True, identifier='kat')
    c = m.centroid
    np.testing.assert_almost_equal([c.lat, c.lon], [55.00295889563608, -99.21825084682715], decimal=6)
    outl = m.outline
    assert 10_000 < len(outl) < 20_000 and not np.isnan(outl).any()
    # the walk equals the oracle's marching squares on the downloaded corner mask, and the
    # reference's own centroid formula (unsigned area: orientation matters) gives the golden too
    valid = ~ma.getmaskarray(m.lats)
    xy = O.outline_marching_squares(valid)
    ref = np.transpose([m.lats.data[xy[:, 1], xy[:, 0]], m.lons.data[xy[:, 1], xy[:, 0]]])
    k = int(np.flatnonzero((ref == outl[0]).all(1))[0])
    assert np.array_equal(np.roll(ref, -k, axis=0), outl)
    np.testing.assert_almost_equal(O.polygon_centroid(ref), [55.00295889563608, -99.21825084682715], decimal=6)
    hull = m.outlineConvexHull
    assert 3 < len(hull) < len(outl)
    bb = m.boundingBox
    assert bb.latSouth == outl[:, 0].min() and bb.latNorth == outl[:, 0].max()
    assert bb.lonWest == outl[:, 1].min() and bb.lonEast == outl[:, 1].max()
    p = m.properties
    assert p.centroid == c and p.boundingBox == bb
    # pixel scales: ~34 arcsec/px WCS scale seen from 400 km -> sub-degree footprints per pixel
    s = m.arcSecPerPx
    assert 0 < s.width.min <= s.width.median <= s.width.max and s.diagonal.mean > s.width.mean


@pytest.mark.parametrize("case", ["plain", "dateline"])
def test_masked_by_polygon_vs_oracle(env, case):
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    W, H = 133, 89
    if case == "plain":
        hdr = synthetic.issHeader(W, H)
        polygon = [[50.0, -108.0], [58.5, -109.0], [60.0, -100.0], [57.0, -99.0], [58.0, -95.0], [49.0, -96.5]]
    else:
        hdr = synthetic.issHeaderLookingAt(52.0, 172.0, 58.0, -179.0, W, H)
        polygon = [[50.0, 170.0], [64.0, 168.0], [66.0, -172.0], [52.0, -170.0]]
    m = getMapping(synthetic.issImage(W, H), hdr, identifier=case)
    assert m.containsDiscontinuity == (case == "dateline")
    r = m.maskedByPolygon(polygon)
    omask, _ = O.polygon_center_mask(m.lats.filled(np.nan), m.lons.filled(np.nan), polygon, wrap180=(case == "dateline"))
    omask |= ma.getmaskarray(m.latsCenter)
    got = ma.getmaskarray(r.latsCenter)
    assert np.array_equal(got, omask)
    assert 0.05 < (~got).mean() < (~ma.getmaskarray(m.latsCenter)).mean()
    r.checkGuarantees()
    assert np.array_equal(r.latsCenter.compressed(), m.latsCenter[~omask].compressed())
    # every retained pixel has its four corners inside the polygon
    from auromat_b200.utils import pointsInsidePolygon
    lats, lons = m.lats.filled(np.nan), m.lons.filled(np.nan)
    if case == "dateline":
        lons = O.wrap_at_180(lons + 180)
        polygon = [[a, float(O.wrap_at_180(b + 180))] for a, b in polygon]
    ins = pointsInsidePolygon(np.transpose([lats.ravel(), lons.ravel()]), polygon).reshape(lats.shape)
    ys, xs = np.nonzero(~got)
    assert ins[ys, xs].all() and ins[ys + 1, xs + 1].all() and ins[ys, xs + 1].all() and ins[ys + 1, xs].all()
    with pytest.raises(ValueError):
        m.maskedByPolygon([[10.0, 10.0], [11.0, 10.0], [11.0, 11.0]])
    # a polygon with more vertices than one shared-memory chunk: the outline itself, shrunk towards the centroid
    if case == "plain":
        outl, c = m.outline, m.centroid
        big = np.repeat(outl, 6, axis=0)
        big = big + (np.array([c.lat, c.lon]) - big) * np.linspace(0.05, 0.06, len(big))[:, None]
        assert len(big) > 1024
        r2 = m.maskedByPolygon(big)
        o2, _ = O.polygon_center_mask(lats, lons, big)
        assert np.array_equal(ma.getmaskarray(r2.latsCenter), o2 | ma.getmaskarray(m.latsCenter))


def test_themis_reproject_and_mapping(env):
    """Ground-based THEMIS route (mapping/themis.py): altitude reprojection kernel against the
    reference's golden and the oracle; `mappingFromCalibration` (reproject -> corner means ->
    -2500 offset -> sanitise -> elevation mask) against the oracle chain; mosaic-ready."""
    import datetime
    import oracle.auromat_oracle as O
    from auromat_b200.mapping import themis
    from auromat_b200.resample import resample
    g = np.load(os.path.join(GOLDEN, "themis_reproject.npz"))
    asi = tuple(float(v) for v in g['asi'])
    for h in (90, 150):
        la, lo = themis.reproject(asi, g['lats110'], g['lons110'], 110.0, float(h))
        assert np.array_equal(np.isnan(la), np.isnan(g['lats%d' % h]))
        assert np.nanmax(np.abs(la - g['lats%d' % h])) <= TOL_DEG
        assert np.nanmax(np.abs(lo - g['lons%d' % h])) <= TOL_DEG
    # a synthetic L2 calibration: fisheye station model evaluated at three reference heights
    w = 64
    cal = dict(xc=w / 2.0, yc=w / 2.0, k=w * 155.81 / 512 * 2, rotation=0.1)
    heights = np.array([90.0, 110.0, 150.0])
    ref = [O.allsky_georeference(w, cal['xc'], cal['yc'], cal['k'], cal['rotation'], asi[0], asi[1], altitude=h)
           for h in heights]
    latsRef = np.array([r['lats'] for r in ref])
    lonsRef = np.array([r['lons'] for r in ref])
    el = ref[0]['elevation'].copy()
    el[np.isnan(ref[0]['latsCenter'])] = np.nan
    img = np.random.default_rng(4).integers(2500, 40000, (w, w), dtype=np.uint16)
    t = datetime.datetime(2012, 2, 4, 7, 56, 26)
    for alt in (110, 120):
        m = themis.mappingFromCalibration('atha', asi, el, latsRef, lonsRef, heights, img, t, altitude=alt)
        assert m.identifier == 'atha.2012.02.04.07.56.26' and m.altitude == alt
        if alt == 110:
            olat, olon = latsRef[1], lonsRef[1]
        else:
            olat, olon = O.themis_reproject(asi, latsRef[0], lonsRef[0], 90.0, 120.0)
            # the same rays georeferenced directly at 120 km: equal up to the error of the reference's
            # single-iteration Bowring inverse at these heights (~1e-6 deg in latitude)
            direct = O.allsky_georeference(w, cal['xc'], cal['yc'], cal['k'], cal['rotation'], asi[0], asi[1], altitude=120)
            assert np.nanmax(np.abs(olat - direct['lats'])) < 2e-5 and np.nanmax(np.abs(olon - direct['lons'])) < 2e-5
        oc_lat, oc_lon = O.themis_corner_means(olat, olon)
        mk, mc = O.sanitize_masks(np.isnan(olat), np.isnan(oc_lat))
        with np.errstate(invalid='ignore'):
            mc2 = mc | ~(el >= 1)
        mk2, mc2 = O.sanitize_masks(mk, mc2, after_masking=True)
        assert np.array_equal(ma.getmaskarray(m.latsCenter), mc2)
        assert np.array_equal(ma.getmaskarray(m.lats), mk2)
        ok = ~mk2
        assert np.abs(m.lats.data[ok] - olat[ok]).max() <= TOL_DEG and np.abs(m.lons.data[ok] - olon[ok]).max() <= TOL_DEG
        okc = ~mc2
        assert np.abs(m.latsCenter.data[okc] - oc_lat[okc]).max() <= TOL_DEG
        assert np.abs(m.lonsCenter.data[okc] - oc_lon[okc]).max() <= TOL_DEG
        assert np.array_equal(m.img_unmasked[:, :, 0], img - np.uint16(2500))
        m.checkGuarantees()
        assert m.rgb.shape == (w, w, 3) and m.rgb.dtype == np.uint8
        r = resample(m, pxPerDeg=8)
        assert isinstance(r, themis.ThemisMapping) and r.station == 'atha'
        r.checkGuarantees()
        r.checkPlateCarree()
    coll = themis.mappingCollection([m, None], t)
    assert len(coll.mappings) == 1 and coll.identifier == 'THEMIS.2012.02.04.07.56.26'


def test_export_handoff_variables(env):
    """auromat_b200.export.cdfVariables: names, order, shapes and dtypes of the variables the
    reference's CDF writer stores (export/cdf.py:83-290), for a frame and its resampling."""
    from auromat_b200 import synthetic
    from auromat_b200.export import cdfVariables
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample
    W, H = 133, 89
    m = getMapping(synthetic.issImage(W, H), synthetic.issHeader(W, H), identifier='x')
    v = cdfVariables(m)
    assert list(v) == ['Epoch', 'lat', 'lon', 'lat_bounds', 'lon_bounds', 'altitude', 'mlat', 'mlt', 'mlat_bounds',
                       'mlt_bounds', 'img_red', 'img_green', 'img_blue', 'zenith_angle', 'camera_pos']
    assert v['lat']['data'].shape == (1, H, W) and v['lat_bounds']['data'].shape == (1, H + 1, W + 1)
    assert v['altitude']['data'] == 110000 and v['camera_pos']['data'].shape == (1, 3)
    assert v['img_red']['data'].dtype == np.int16 and v['img_red']['attrs']['FILLVAL'] == -32768
    masked = ma.getmaskarray(m.latsCenter)
    assert np.array_equal(v['img_green']['data'][0] == -32768, masked)
    assert np.array_equal(np.isnan(v['mlt']['data'][0]), masked)
    assert v['zenith_angle']['data'].dtype == np.float32
    np.testing.assert_allclose(v['zenith_angle']['data'][0][~masked], 90 - m.elevation.compressed(), atol=1e-4)
    assert v['mlt']['attrs']['UNITS'] == 'hours' and v['lat']['attrs']['DEPEND_1'] == 'y_pixel'
    r = resample(m, arcsecPerPx=400)
    vr = cdfVariables(r, includeBounds=False, includeMagCoords=False)
    assert list(vr) == ['Epoch', 'lat', 'lon', 'altitude', 'img_red', 'img_green', 'img_blue', 'zenith_angle', 'camera_pos']
    assert vr['lat']['data'].shape[1:] == r.latsCenter.shape


def test_mosaic_over_the_pole(env):
    """A mosaic whose members enclose the pole: common +90 deg pole rotation for all members
    (resample.py:176-201 applied collection-wide).  One member: identical to resample(); two
    members: counts and sums are the sums of the members binned alone into the common grid."""
    import torch
    from auromat_b200 import parallel, synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import binMappingInto, resampleToDevice
    W, H = 300, 200
    m1 = getMapping(synthetic.issImage(W, H, seed=1), synthetic.issHeaderLookingAt(78.0, 20.0, 89.5, 120.0, W, H),
                    identifier='polar')
    m2 = getMapping(synthetic.issImage(W, H, seed=2), synthetic.issHeaderLookingAt(70.0, -60.0, 80.0, -80.0, W, H),
                    identifier='subpolar')
    assert m1.containsPole and not m2.containsPole
    ppd = (4.0, 2.0)
    mosaic, acc = parallel.mosaic([m1], ppd)
    grid, info, outImg, outMask, outElev = resampleToDevice(m1, pxPerDeg=ppd)
    assert (acc.grid.nx, acc.grid.ny) == (grid.nx, grid.ny) and acc.grid.prerotate == 2
    assert torch.equal(acc.count, info['count'])
    assert np.array_equal(mosaic.img.filled(0), np.where(outMask.cpu().numpy()[:, :, None] != 0, 0, outImg.cpu().numpy()))
    # rotated back: the mosaic's corner coordinates surround the pole
    assert mosaic.lats.max() > 89.0 and mosaic.lons.min() < -170 and mosaic.lons.max() > 170
    both, acc2 = parallel.mosaic([m1, m2], ppd)
    cells = acc2.grid.nx * acc2.grid.ny
    tot_c, tot_s = torch.zeros(cells, dtype=torch.int64, device='cuda'), torch.zeros(3 * cells, dtype=torch.int64, device='cuda')
    for m in (m1, m2):
        c = torch.zeros(cells, dtype=torch.int64, device='cuda')
        s = torch.zeros(3 * cells, dtype=torch.int64, device='cuda')
        f = torch.zeros(cells, dtype=torch.float64, device='cuda')
        binMappingInto(m, acc2.grid, c, s, f)
        # (the outermost half cells of the snapped range are not part of the grid, resample.py:232-237)
        assert 0.99 * m._deviceStats().n_valid_centers < int(c.sum().item()) <= m._deviceStats().n_valid_centers
        tot_c += c
        tot_s += s
    assert torch.equal(acc2.count, tot_c) and torch.equal(acc2.sums, tot_s)
    assert (~ma.getmaskarray(both.latsCenter)).sum() > 0


def test_pipeline_gray_images_plane_free_and_empty_frames(env):
    """Sequence pipeline corner cases: single-channel (grey-scale) images through the sparse
    row-range upload, the plane-free mode with host images, a frame that sees no Earth at all
    (the reference raises on its empty outline; so does the generator, leaving the context usable)."""
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence
    from auromat_b200.resample import resample
    W, H, n = 200, 130, 4
    hdrs = synthetic.sequenceHeaders(n, W, H)
    rng = np.random.default_rng(3)
    for shape, dtype in (((H, W, 1), np.uint8), ((H, W, 1), np.uint16)):
        imgs = [rng.integers(0, 200, shape).astype(dtype) for _ in range(n)]
        expect = [resample(getMapping(im, h, identifier='x'), arcsecPerPx=400) for im, h in zip(imgs, hdrs)]
        for ring in (False, True):
            got = list(resampleSequence(imgs, hdrs, arcsecPerPx=400, ringBuffers=ring))
            for f, e in zip(got, expect):
                assert f.img.shape == e.img.shape and f.img.dtype == e.img.dtype
                assert np.array_equal(ma.getmaskarray(f.img), ma.getmaskarray(e.img))
                assert np.array_equal(f.img.filled(0), e.img.filled(0))
    imgs = [synthetic.issImage(W, H, seed=i) for i in range(n)]
    expect = [resample(getMapping(im, h, identifier='x'), arcsecPerPx=400) for im, h in zip(imgs, hdrs)]
    tr = {}
    got = list(resampleSequence(imgs, hdrs, arcsecPerPx=400, coordinates=False, ringBuffers=True, transferStats=tr))
    assert 0 < tr['h2d_bytes'] < sum(im.nbytes for im in imgs)
    for f, e in zip(got, expect):
        assert 'lat_c' not in f.mapping._planes            # really plane-free
        assert np.array_equal(f.img.filled(0), e.img.filled(0))
    # a camera that looks away from the Earth: no valid pixel
    sky = dict(hdrs[0])
    sky['CRVAL2'] = -hdrs[0]['CRVAL2']
    sky['CRVAL1'] = (hdrs[0]['CRVAL1'] + 180.0) % 360.0
    with pytest.raises(ValueError):
        list(resampleSequence([imgs[0], imgs[1]], [hdrs[0], sky], arcsecPerPx=400, ringBuffers=True))
    again = list(resampleSequence(imgs[:2], hdrs[:2], arcsecPerPx=400, ringBuffers=True))
    assert np.array_equal(again[1].img.filled(0), expect[1].img.filled(0))


def test_config3_sip_frame_bands_vs_oracle(env):
    """BASELINE configs[2] at full size: 6000x4000 frame with SIP order-4 distortion.  Two bands of
    rows (one through the limb, one near the bottom edge where the distortion is largest) against
    the oracle, all nine planes, unsanitised so that band edges do not matter; then the fine
    10 arcsec/px resampling keeps every valid pixel it should (conservation)."""
    import torch
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resampleToDevice
    W, H = 6000, 4000
    hdr = synthetic.issHeader(W, H, sipOrder=4)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    m = getMapping(env.to_device(synthetic.issImage(W, H)), hdr, nosanitize=True, identifier='c3')
    m.prefetch(magnetic=True)
    names = dict(lats='lat_k', lons='lon_k', mlat='mlat_k', mlt='mlt_k', latsCenter='lat_c', lonsCenter='lon_c',
                 mlatCenter='mlat_c', mltCenter='mlt_c', elevation='elev_c')
    p = m.devicePlanes(magnetic=True)
    # find the first row with a valid centre: the limb band starts a little above it
    rows = torch.nonzero(~torch.isnan(p['lat_c'].reshape(H, W)).all(dim=1)).ravel()
    limb = max(0, int(rows[0].item()) - 8)
    for y0, nrow in ((limb, 24), (H - 24, 24)):
        band = dict(hdr)
        band['IMAGEH'] = nrow
        band['CRPIX2'] = hdr['CRPIX2'] - y0
        with quiet():
            g = O.georeference(band, cam, t, 110)
        n_graze = 0
        for oname, pname in names.items():
            corner = pname.endswith('_k')
            w1 = W + 1 if corner else W
            a = env.to_numpy(p[pname].reshape(-1, w1)[y0:y0 + nrow + (1 if corner else 0)])
            b = g[oname]
            assert a.shape == b.shape, oname
            nan_diff = np.isnan(a) != np.isnan(b)
            n_graze = max(n_graze, int(nan_diff.sum()))
            both = ~np.isnan(a) & ~np.isnan(b)
            d = np.abs(a[both] - b[both])
            if oname.startswith('lon'):
                d = np.minimum(d, 360 - d)
            tol = TOL_DEG / 15 if oname.startswith('mlt') else TOL_DEG
            if oname == 'elevation':
                s = np.sin(np.deg2rad(np.clip(90 - b[both], 1e-9, None)))
                tol = TOL_DEG + np.rad2deg(8 * 2.2e-16 / s)
            assert np.all(d <= tol), (oname, y0, float(d.max()))
        assert n_graze <= m.illConditionedCount + 2          # hit/miss may differ only for grazing rays
        if y0 == limb:
            assert 0 < np.isnan(g['latsCenter']).mean() < 1  # the band really crosses the limb
    del m
    ms = getMapping(env.to_device(synthetic.issImage(W, H)), hdr, identifier='c3s')
    grid, info, outImg, outMask, outElev = resampleToDevice(ms, arcsecPerPx=10)
    assert grid.nx * grid.ny > 15_000_000                    # ~20 M cells, sparsely filled
    ps = ms.devicePlanes()
    ix, iy = env.cell_indices(ps['lat_c'], ps['lon_c'], grid)
    inside = int(((ix >= 0) & (iy >= 0)).sum().item())
    assert int(info['count'].sum().item()) == inside
    assert 0.99 * ms._deviceStats().n_valid_centers < inside <= ms._deviceStats().n_valid_centers


def test_config4_full_size_sequence_equals_single_frames(env):
    """BASELINE configs[3] at full size: frames of the synthetic 4256x2832 sequence through the
    pipelined path (host images, row-range upload, plane rings) are bit-identical to
    `resample(getMapping(...))` of the same frames."""
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence
    from auromat_b200.resample import resample
    n = 5
    hdrs = synthetic.sequenceHeaders(n)
    imgs = [synthetic.issImage(seed=1000 + i) for i in range(n)]
    got = list(resampleSequence(imgs, hdrs, arcsecPerPx=100, magnetic=True, ringBuffers=True))
    for i in (0, 3, 4):
        e = resample(getMapping(imgs[i], hdrs[i], identifier='x'), arcsecPerPx=100)
        f = got[i]
        assert f.img.shape == e.img.shape
        assert np.array_equal(ma.getmaskarray(f.img), ma.getmaskarray(e.img))
        assert np.array_equal(f.img.filled(0), e.img.filled(0))
        fe, ee = f.elevation.filled(np.nan), e.elevation.filled(np.nan)
        assert np.array_equal(np.isnan(fe), np.isnan(ee)) and np.nanmax(np.abs(fe - ee)) < 1e-9
        del e


def test_config5_mosaic_of_64_stations_vs_oracle(env):
    """BASELINE configs[4]: 64 synthetic all-sky stations (256x256, uint16, elevation >= 1 deg) binned
    into one common 20 px/deg grid: counts and integer sums bit-exact against the serial oracle (sum
    of per-station histogram2d outputs)."""
    import datetime
    import oracle.auromat_oracle as O
    from auromat_b200 import parallel
    from auromat_b200.mapping.allsky import getMappingCollection
    t = datetime.datetime(2012, 3, 4, 17, 19, 0)
    cals = _stations(64)
    rng = np.random.default_rng(11)
    w = 256
    imgs = [rng.integers(0, 65536, (w, w, 1), dtype=np.uint16) for _ in cals]
    coll = getMappingCollection(imgs, cals, t, 110, minElevation=1)
    mosaic, acc = parallel.mosaic(coll.mappings, pxPerDeg=(20, 20))
    grid = acc.grid
    bins, rng_ = (grid.nx, grid.ny), [[grid.lo_x, grid.hi_x], [grid.lo_y, grid.hi_y]]
    tc = np.zeros((grid.ny, grid.nx))
    ts = np.zeros((grid.ny, grid.nx))
    for m, im in zip(coll.mappings, imgs):
        la, lo = m.latsCenter.filled(np.nan).ravel(), m.lonsCenter.filled(np.nan).ravel()
        if acc.info['mode'] == 1:
            lo = O.wrap_at_180(lo + 180)
        ok = ~np.isnan(la)
        c, s = O.histogram2d_weighted(lo[ok], la[ok], bins, rng_, [None, im.reshape(-1)[ok].astype(float)])
        tc += c.T[::-1]
        ts += s.T[::-1]
    assert np.array_equal(acc.count.cpu().numpy().reshape(grid.ny, grid.nx), tc)
    assert np.array_equal(acc.sums.cpu().numpy().reshape(grid.ny, grid.nx), ts)
    assert tc.sum() > 2_000_000 and tc.max() > 4          # overlapping stations really accumulate
    with np.errstate(invalid='ignore', divide='ignore'):
        mean = np.round(ts / tc)
    got = mosaic.img
    assert np.array_equal(ma.getmaskarray(got)[:, :, 0], tc == 0)
    assert np.array_equal(got.filled(0)[:, :, 0][tc > 0], mean[tc > 0].astype(np.uint16))


def test_collections_providers_and_footpoints(env, tmp_path):
    """The glue around the hot path with real device mappings: `resample(MappingCollection)`,
    collection members, `cameraFootpoint`, and a folder provider wrapped by
    MaskByElevationProvider / ResampleProvider (mapping.py:1315-1472, resample.py:143-157,370-394)."""
    import datetime
    from PIL import Image
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.allsky import getMappingCollection
    from auromat_b200.mapping.mapping import BoundingBox, MappingCollection, MaskByElevationProvider
    from auromat_b200.mapping.spacecraft import SpacecraftMappingProvider, getMapping
    from auromat_b200.resample import ResampleProvider, resample
    t = datetime.datetime(2012, 3, 4, 17, 19, 0)
    cals = _stations(3)
    w = 96
    imgs = [np.random.default_rng(i).integers(0, 65536, (w, w, 1), dtype=np.uint16) for i in range(3)]
    coll = getMappingCollection(imgs, cals, t, 110, minElevation=5)
    assert isinstance(coll, MappingCollection) and len(coll) == 3 and coll.photoTime == t
    assert coll.boundingBox == BoundingBox.mergedBoundingBoxes([m.boundingBox for m in coll.mappings])
    higher = coll.maskedByElevation(30)
    for a, b in zip(coll.mappings, higher.mappings):
        assert 0 < (~ma.getmaskarray(b.latsCenter)).sum() < (~ma.getmaskarray(a.latsCenter)).sum()
        assert b.elevation.min() >= 30
    rc = resample(coll, pxPerDeg=10)
    assert isinstance(rc, MappingCollection) and len(rc) == 3 and rc.identifier == coll.identifier
    for r, m in zip(rc.mappings, coll.mappings):
        r.checkPlateCarree()
        r.checkGuarantees()
        single = resample(m, pxPerDeg=10)
        assert np.array_equal(r.img.filled(0), single.img.filled(0))
    # camera footpoint of a spacecraft mapping: Bowring of the GEO camera position
    W, H = 96, 64
    hdr = synthetic.issHeader(W, H)
    tt, cam = synthetic.headerTimeAndCamera(hdr)
    m = getMapping(synthetic.issImage(W, H), hdr, identifier='fp')
    g = O.mat_j2000_to_geo(O.date2es(tt)).dot(cam)
    la, lo = O.ecef2geodetic_scalar(g[0], g[1], g[2])
    fp = m.cameraFootpoint
    assert abs(fp.lat - np.rad2deg(la)) < 1e-12 and abs(fp.lon - np.rad2deg(lo)) < 1e-12
    assert m.rgb.shape == (H, W, 3) and m.rgb_unmasked.dtype == np.uint8
    # folder provider -> elevation mask -> resampling, all through the wrappers
    hdrs = synthetic.sequenceHeaders(2, W, H)
    for i, h in enumerate(hdrs):
        cards = ["%-8s= %20s" % ('SIMPLE', 'T')] + \
                ["%-8s= %s" % (k, ("'%s'" % v) if isinstance(v, str) else repr(v)) for k, v in h.items()] + ["END"]
        raw = "".join(c.ljust(80) for c in cards)
        (tmp_path / ('f%d.wcs' % i)).write_bytes(raw.ljust((len(raw) + 2879) // 2880 * 2880).encode('ascii'))
        Image.fromarray(synthetic.issImage(W, H, seed=i)).save(str(tmp_path / ('f%d.png' % i)))
    prov = ResampleProvider(MaskByElevationProvider(SpacecraftMappingProvider(str(tmp_path)), 20), arcsecPerPx=600)
    seq = list(prov.getSequence())
    assert len(seq) == 2
    direct = resample(getMapping(synthetic.issImage(W, H, seed=1), hdrs[1], identifier='d').maskedByElevation(20),
                      arcsecPerPx=600)
    assert np.array_equal(seq[1].img.filled(0), direct.img.filled(0))
    assert np.array_equal(ma.getmaskarray(seq[1].latsCenter), ma.getmaskarray(direct.latsCenter))
    assert seq[1].elevation.min() >= 20 - 1e-9


def test_pipeline_box_upload_for_rolled_cameras(env):
    """A camera rolled by 90 degrees puts the limb (roughly) vertical: the valid pixels form a
    column band and the pipeline uploads that pixel box with one 2-D copy.  Results equal the
    full upload and the frame-by-frame path."""
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence
    from auromat_b200.resample import resample
    W, H, n = 320, 200, 3
    hdrs = []
    for h in synthetic.sequenceHeaders(n, W, H):
        h = dict(h)
        cd = np.array([[h['CD1_1'], h['CD1_2']], [h['CD2_1'], h['CD2_2']]]).dot(np.array([[0.0, -1.0], [1.0, 0.0]]))
        (h['CD1_1'], h['CD1_2']), (h['CD2_1'], h['CD2_2']) = cd[0], cd[1]
        hdrs.append(h)
    imgs = [synthetic.issImage(W, H, seed=70 + i) for i in range(n)]
    m = getMapping(imgs[0], hdrs[0], identifier='roll')
    st = m._deviceStats()
    assert st.n_valid_centers > 0.2 * W * H
    assert (st.col_max_c - st.col_min_c + 1) < 0.9 * W and (st.row_max_c - st.row_min_c + 1) == H   # a column band
    expect = [resample(getMapping(im, h, identifier='x'), arcsecPerPx=400) for im, h in zip(imgs, hdrs)]
    for ring in (False, True):
        tr = {}
        got = list(resampleSequence(imgs, hdrs, arcsecPerPx=400, ringBuffers=ring, transferStats=tr))
        box = sum((f.mapping._deviceStats().col_max_c - f.mapping._deviceStats().col_min_c + 1) *
                  (f.mapping._deviceStats().row_max_c - f.mapping._deviceStats().row_min_c + 1) * 3 for f in got)
        assert tr['h2d_bytes'] == box < 0.9 * sum(im.nbytes for im in imgs)
        for f, e in zip(got, expect):
            assert np.array_equal(ma.getmaskarray(f.img), ma.getmaskarray(e.img))
            assert np.array_equal(f.img.filled(0), e.img.filled(0))


def test_outline_queue_overflow_noise_mask(env):
    """A hand-made mapping whose validity is salt-and-pepper noise: more outline corners (1.4 M) than the node
    queue of k_outline_collect holds (2^20), so the excess is evaluated in place by the collecting kernel --
    bounding box, pixel box and counts must still equal numpy's over the sanitised masks."""
    import datetime
    import torch
    from auromat_b200.mapping.mapping import GenericMapping
    h, w = 1900, 2000
    rng = np.random.default_rng(12)
    lats = np.linspace(60, 40, h + 1)[:, None] + rng.uniform(-1e-3, 1e-3, (h + 1, w + 1))
    lons = np.linspace(-20, 25, w + 1)[None, :] + rng.uniform(-1e-3, 1e-3, (h + 1, w + 1))
    latsC = 0.25 * (lats[:-1, :-1] + lats[1:, :-1] + lats[:-1, 1:] + lats[1:, 1:])
    lonsC = 0.25 * (lons[:-1, :-1] + lons[1:, :-1] + lons[:-1, 1:] + lons[1:, 1:])
    # 2 x 2 blocks of pixels survive with probability 1/2: after sanitisation the valid region is a foam
    # in which nearly every valid corner touches an invalid neighbour
    keep = np.kron(rng.random((h // 2, w // 2)) < 0.5, np.ones((2, 2), bool))
    latsC = np.where(keep, latsC, np.nan)
    lonsC = np.where(keep, lonsC, np.nan)
    img = rng.integers(0, 255, (h, w, 1), dtype=np.uint8)
    m = GenericMapping(lats, lons, latsC, lonsC, np.full((h, w), 30.0), 110, img, np.zeros(3),
                       datetime.datetime(2012, 1, 1), 'noise')
    st = m._deviceStats()
    p = m.devicePlanes()
    vk = ~torch.isnan(p['lat_k']).reshape(h + 1, w + 1).cpu().numpy()
    vc = ~torch.isnan(p['lat_c']).reshape(h, w).cpu().numpy()
    pad = np.pad(vk, 1)
    interior = pad[1:-1, 1:-1] & pad[:-2, 1:-1] & pad[2:, 1:-1] & pad[1:-1, :-2] & pad[1:-1, 2:]
    outline = vk & ~interior
    assert outline.sum() > 2 ** 20                      # the queue does overflow
    assert st.n_boundary_corners == outline.sum() and st.n_valid_corners == vk.sum() and st.n_valid_centers == vc.sum()
    la, lo = p['lat_k'].reshape(h + 1, w + 1).cpu().numpy(), p['lon_k'].reshape(h + 1, w + 1).cpu().numpy()
    assert (st.lat_min, st.lat_max, st.lon_min, st.lon_max) == (la[outline].min(), la[outline].max(),
                                                                lo[outline].min(), lo[outline].max())
    rows, cols = np.nonzero(vc)
    assert (st.row_min_c, st.row_max_c, st.col_min_c, st.col_max_c) == (rows.min(), rows.max(), cols.min(), cols.max())


# ======================================================================= round 2: fused path
def _geometries(W, H):
    """Camera geometries for the limb solver: the ISS frame, random rolls / scales / pointings
    (limb anywhere, Earth filling the frame, Earth absent), pole and date-line views, nadir."""
    import math
    from auromat_b200 import synthetic
    yield 'iss', synthetic.issHeader(W, H)
    rng = np.random.default_rng(5)
    base = synthetic.issHeader(W, H)
    for i in range(20):
        h = dict(base)
        th = rng.uniform(0, 2 * math.pi)
        s = math.hypot(base['CD1_1'], base['CD1_2']) * rng.uniform(0.5, 2.5)
        h['CD1_1'], h['CD1_2'] = -s * math.cos(th), -s * math.sin(th)
        h['CD2_1'], h['CD2_2'] = s * math.sin(th), -s * math.cos(th)
        h['CRVAL1'] = base['CRVAL1'] + rng.uniform(-40, 40)
        h['CRVAL2'] = base['CRVAL2'] + rng.uniform(-40, 40)
        yield 'rand%d' % i, h
    yield 'pole', synthetic.issHeaderLookingAt(80.0, 10.0, 89.5, 40.0, W, H)
    yield 'dateline', synthetic.issHeaderLookingAt(50.0, 170.0, 55.0, -178.0, W, H)
    yield 'nadir', synthetic.issHeaderLookingAt(50.0, 10.0, 50.0, 10.01, W, H)


def _hit_bitmaps(ctx, frame, W, H, limb):
    os.environ.pop('AMT_NO_LIMB_SOLVER', None)
    if not limb:
        os.environ['AMT_NO_LIMB_SOLVER'] = '1'
    try:
        bits = {}
        bits['valid_k'], bits['valid_c'] = ctx.new_bitmaps(W, H)
        st = ctx.new_stats()
        ctx.georef(frame, bits, st)
        return bits, int(ctx.read_stats(st).n_ill_conditioned)
    finally:
        os.environ.pop('AMT_NO_LIMB_SOLVER', None)


@pytest.mark.parametrize("W,H", [(97, 61), (532, 354), (1064, 708)])
def test_limb_solver_equals_per_pixel_hit_test(env, W, H):
    """The O(H) limb solver (row-wise roots of the discriminant + exact evaluation near them) gives the
    same corner and centre hit bitmaps, and the same grazing-ray count, as the per-pixel hit test."""
    import torch
    from auromat_b200.mapping.spacecraft import getMapping
    seen = set()
    for name, hdr in _geometries(W, H):
        fr = getMapping(np.zeros((H, W, 1), np.uint8), hdr, identifier=name).frameConstants
        a, ga = _hit_bitmaps(env, fr, W, H, True)
        b, gb = _hit_bitmaps(env, fr, W, H, False)
        assert torch.equal(a['valid_k'], b['valid_k']), name
        assert torch.equal(a['valid_c'], b['valid_c']), name
        assert ga == gb, name
        frac = float((b['valid_c'] != 0).float().mean())
        seen.add('none' if frac == 0 else 'all' if frac == 1 else 'limb')
    assert 'limb' in seen


@pytest.mark.parametrize("W,H,order", [(97, 61, 2), (532, 354, 4), (1064, 708, 3), (700, 500, 5)])
def test_sip_limb_solver_equals_per_pixel_hit_test(env, W, H, order):
    """TAN-SIP frames: words whose displacement-inflated box is provably on one side of the limb (quadratic-form
    bound, k_limb_bits_sip) take the exact predicate of their first pixel, all others are evaluated per pixel
    -- same bitmaps and the same grazing-ray count as the per-pixel hit test, over random camera geometries
    and random SIP polynomials (displacements of a few to tens of pixels)."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    seen, evaluated = set(), []
    for k, (name, hdr) in enumerate(_geometries(W, H)):
        sip = synthetic.issHeader(W, H, sipOrder=order, seed=10 + k)
        h = dict(hdr)
        for key, v in sip.items():
            if key.startswith(('A_', 'B_', 'CTYPE')):
                h[key] = v
        fr = getMapping(np.zeros((H, W, 1), np.uint8), h, identifier=name).frameConstants
        a, ga = _hit_bitmaps(env, fr, W, H, True)
        b, gb = _hit_bitmaps(env, fr, W, H, False)
        assert torch.equal(a['valid_k'], b['valid_k']), name
        assert torch.equal(a['valid_c'], b['valid_c']), name
        assert ga == gb, name
        frac = float((b['valid_c'] != 0).float().mean())
        seen.add('none' if frac == 0 else 'all' if frac == 1 else 'limb')
    assert 'limb' in seen


@pytest.mark.parametrize("sip", [0, 4])
@pytest.mark.parametrize("dtype,channels", [(np.uint8, 3), (np.uint16, 1)])
def test_fused_kernel_equals_unfused_chain(env, sip, dtype, channels):
    """amt_georef_fused (planes + binning from the final bitmaps) == amt_georef + amt_sanitize +
    amt_bin_accumulate: all nine planes bit for bit incl. the NaN pattern, counts, integer sums and the
    fixed-point elevation sums identical; plane-free mode identical too."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import deriveGrid
    W, H = 532, 354
    hdr = synthetic.issHeader(W, H, sipOrder=sip)
    img = np.random.default_rng(1).integers(0, np.iinfo(dtype).max, (H, W, channels)).astype(dtype)
    m = getMapping(img, hdr, identifier='f')
    fr = m.frameConstants
    old = m.devicePlanes(magnetic=True)                    # amt_georef + amt_sanitize
    bits, _ = _hit_bitmaps(env, fr, W, H, True)
    env.sanitize(W, H, bits)
    assert torch.equal(bits['valid_k'], old['valid_k']) and torch.equal(bits['valid_c'], old['valid_c'])
    names = ('lat_k', 'lon_k', 'mlat_k', 'mlt_k', 'lat_c', 'lon_c', 'mlat_c', 'mlt_c', 'elev_c')
    dimg = m.deviceImage()
    for ppd in (None, (360.0, 200.0)):
        grid, info = deriveGrid(m, pxPerDeg=ppd, arcsecPerPx=None if ppd else 100)
        assert grid.side_scale > 0
        cells = grid.nx * grid.ny

        def parts(acc):
            return acc[:cells], acc[cells:(1 + channels) * cells], acc[(1 + channels) * cells:].view(torch.float64)
        accA, accB, accC = (env.zeros((2 + channels) * cells, torch.int64) for _ in range(3))
        env.bin_accumulate(old['lat_c'], old['lon_c'], old['elev_c'], dimg, grid, *parts(accA))
        new = {n: torch.full_like(old[n], 7.0) for n in names}
        env.georef_fused(fr, bits['valid_k'], bits['valid_c'], planes=new, img=dimg, grid=grid,
                         count=parts(accB)[0], sums=parts(accB)[1], fsum=parts(accB)[2])
        env.georef_fused(fr, None, bits['valid_c'], img=dimg, grid=grid, count=parts(accC)[0],
                         sums=parts(accC)[1], fsum=parts(accC)[2])
        for n in names:
            assert torch.equal(torch.isnan(old[n]), torch.isnan(new[n])), n
            assert torch.equal(torch.nan_to_num(old[n]), torch.nan_to_num(new[n])), n
        assert int(accA[:cells].sum()) > 0
        assert torch.equal(accA, accB) and torch.equal(accA, accC)
        # the fixed-point elevation mean agrees with the f64-atomic mean far inside the 1e-6 budget
        scale = grid.side_scale
        grid.side_scale = 0.0
        accF = env.zeros((2 + channels) * cells, torch.int64)
        env.bin_accumulate(old['lat_c'], old['lon_c'], old['elev_c'], dimg, grid, *parts(accF))
        assert torch.equal(accF[:(1 + channels) * cells], accA[:(1 + channels) * cells])
        fx = accA[(1 + channels) * cells:].double() / scale
        ff = parts(accF)[2]
        hit = accA[:cells] > 0
        assert float(((fx - ff)[hit].abs() / accA[:cells][hit]).max()) < 1e-9


def test_elevation_sums_are_deterministic(env):
    """Fixed-point side channel: two runs of the same binning give identical bits (the f64-atomic
    sums of round 1 differed in their last bits from run to run)."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resampleToDevice
    hdr = synthetic.issHeader(1064, 708)
    m = getMapping(synthetic.issImage(1064, 708), hdr, identifier='d')
    runs = [resampleToDevice(m, pxPerDeg=7)[4].clone() for _ in range(4)]
    for r in runs[1:]:
        assert torch.equal(torch.nan_to_num(r), torch.nan_to_num(runs[0]))


@pytest.mark.parametrize("depth", [1, 3, 5])
def test_ring_planes_lifetime(env, depth):
    """Ring mode: the planes of `frame.mapping` are the frame's own while KEEP_FRAMES + 1 further frames
    are taken; afterwards the mapping is detached from the ring and recomputes -- it never shows another
    frame's coordinates."""
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence, KEEP_FRAMES
    W, H, n = 200, 130, 14
    hdrs = synthetic.sequenceHeaders(n, W, H)
    imgs = [synthetic.issImage(W, H, seed=i) for i in range(n)]
    expect = [getMapping(im, h, identifier='x').latsCenter.filled(np.nan) for im, h in zip(imgs, hdrs)]
    assert not np.array_equal(expect[0], expect[1], equal_nan=True)
    held = []
    for i, f in enumerate(resampleSequence(imgs, hdrs, arcsecPerPx=400, magnetic=True, ringBuffers=True, depth=depth)):
        held.append(f)
        # frames taken earlier: still within the promise -> ring planes, later -> detached and recomputed
        for j in (i, i - KEEP_FRAMES - 1, i - KEEP_FRAMES - 3):
            if j >= 0:
                got = held[j].mapping.latsCenter.filled(np.nan)
                assert np.array_equal(got, expect[j], equal_nan=True), (depth, i, j)
        if i >= 1:
            held[i - 1].mapping._host.pop('lat_c', None)      # force the next access to read the device planes again
    for j, f in enumerate(held):
        f.mapping._host.pop('lat_c', None)
        assert np.array_equal(f.mapping.latsCenter.filled(np.nan), expect[j], equal_nan=True), (depth, j)
        assert np.array_equal(f.mapping.img_unmasked, imgs[j])


def test_config4_long_sequence_ring_wraparound(env):
    """BASELINE configs[3], 64 full-size frames through the ring pipeline: the ring wraps several times,
    the pinned-buffer pool does not grow after the first frames, and frames from every part of the
    sequence equal `resample(getMapping(...))`."""
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.pipeline import resampleSequence
    from auromat_b200.resample import resample
    n = 64
    hdrs = synthetic.sequenceHeaders(n)
    imgs = [synthetic.issImage(seed=1000 + i) for i in range(3)]
    check = {0, 9, 31, 62, 63}
    kept = {}
    tr = {}
    # warm the pools with a short sequence, then the long one must not page-lock anything
    list(resampleSequence([imgs[0]] * 12, hdrs[:12], arcsecPerPx=100, magnetic=True, ringBuffers=True))
    grown0 = env.__dict__['_pinned_frames'].get('grown', 0)
    for i, f in enumerate(resampleSequence([imgs[i % 3] for i in range(n)], hdrs, arcsecPerPx=100, magnetic=True,
                                           ringBuffers=True, transferStats=tr)):
        if i in check:
            kept[i] = (f.img, f.elevation.filled(np.nan))
    assert env.__dict__['_pinned_frames'].get('grown', 0) == grown0
    for i in sorted(check):
        e = resample(getMapping(imgs[i % 3], hdrs[i], identifier='x'), arcsecPerPx=100)
        gi, ge = kept[i]
        assert np.array_equal(ma.getmaskarray(gi), ma.getmaskarray(e.img)), i
        assert np.array_equal(gi.filled(0), e.img.filled(0)), i
        ee = e.elevation.filled(np.nan)
        assert np.array_equal(np.isnan(ge), np.isnan(ee)) and np.nanmax(np.abs(ge - ee)) < 1e-9, i
        del e


def test_config3_fine_grid_counts_vs_oracle(env):
    """BASELINE configs[2] numerics on a crop: a 420x100 window just below the limb of the 6000x4000 SIP
    order-4 frame resampled at 10 arcsec/px (cells far smaller than the oblique pixel footprints: one
    sample per touched cell; the limb and with it the sanitisation are inside the window) -- counts,
    rounded means and masks against the oracle's histogram of the SAME coordinates, and against the
    oracle's own chain up to the reported near-edge samples."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.resample import resample, resampleToDevice
    W, H = 6000, 4000
    full = synthetic.issHeader(W, H, sipOrder=4)
    x0, y0, w, h = 2900, 1600, 420, 100
    hdr = dict(full)
    hdr['IMAGEW'], hdr['IMAGEH'] = w, h
    hdr['CRPIX1'], hdr['CRPIX2'] = full['CRPIX1'] - x0, full['CRPIX2'] - y0
    img = synthetic.issImage(w, h, 5)
    m = getMapping(img, hdr, identifier='crop')
    geo = {k: v.filled(np.nan) for k, v in gpu_arrays(m).items()}
    bb = m.boundingBox
    from auromat_b200.resample import plateCarreeResolution
    ppd = plateCarreeResolution(bb, 10)
    r = resample(m, pxPerDeg=ppd)
    with quiet():
        o = O.resample_frame(geo, img, 110, px_per_deg=ppd, return_count=True)
    grid, info, _, _, _ = resampleToDevice(m, pxPerDeg=ppd)
    cnt = info['count'].cpu().numpy().reshape(grid.ny, grid.nx)
    assert cnt.shape == o['count'].shape and cnt.size > 20 * w * h        # far more cells than pixels
    assert np.array_equal(cnt, o['count'])
    assert cnt.max() <= 2 and int(cnt.sum()) == int((~np.isnan(geo['latsCenter'])).sum())
    assert np.array_equal(ma.getmaskarray(r.img), o['img_mask'])
    assert np.array_equal(r.img.filled(0), np.where(o['img_mask'], 0, o['img']))
    # the sequence engine (fused kernel, SIP variant) gives the same grid
    from auromat_b200.pipeline import resampleSequence
    f = next(iter(resampleSequence([img], [hdr], pxPerDeg=ppd)))
    assert np.array_equal(f.img.filled(0), r.img.filled(0)) and np.array_equal(ma.getmaskarray(f.img), ma.getmaskarray(r.img))


def test_sip_device_polynomial_against_exact_rational_arithmetic(env):
    """The device SIP polynomial (amt_sip_distort: the function the georeference kernels call) against
    exact rational arithmetic at the same 120 points as the oracle's pin: <= 2 ulp of the distorted
    coordinate.  Oracle and kernel are thereby pinned independently of each other."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.coordinates.wcs import frameConstants
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'amt_test_oracle_golden', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'test_oracle_golden.py'))
    kat = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(kat)
    _sip_exact, sip_kat_points = kat._sip_exact, kat.sip_kat_points
    hdr = synthetic.issHeader(6000, 4000, sipOrder=4)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    fr = frameConstants(hdr, cam, t, 110)
    u, v = sip_kat_points()
    uo, vo = env.sip_distort(fr, env.to_device(u), env.to_device(v))
    uo, vo = uo.cpu().numpy(), vo.cpu().numpy()
    for i in range(len(u)):
        eu = float(_sip_exact(hdr, 'A', u[i], v[i]) + __import__('fractions').Fraction(u[i]))
        ev = float(_sip_exact(hdr, 'B', u[i], v[i]) + __import__('fractions').Fraction(v[i]))
        assert abs(uo[i] - eu) <= 2 * np.spacing(abs(eu)) + 1e-13, (i, uo[i], eu)
        assert abs(vo[i] - ev) <= 2 * np.spacing(abs(ev)) + 1e-13, (i, vo[i], ev)


def test_sphere_earth_model(env):
    """`inflatedEarthIntersection(..., earthModel='sphere')` (reference mapping/mapping.py:1502-1505,
    coordinates/intersection.py:12-56): the kernels are generic in the semi-axes; against the oracle's
    restatement of sphereLineIntersection."""
    import oracle.auromat_oracle as O
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    W, H = 266, 177
    hdr = synthetic.issHeader(W, H)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    m = getMapping(synthetic.issImage(W, H), hdr, nosanitize=True, identifier='s')
    m.earthModel = 'sphere'
    with quiet():
        g = O.georeference(hdr, cam, t, 110, earth_model='sphere')
    w = assert_coords_close(gpu_arrays(m), g)
    e = getMapping(synthetic.issImage(W, H), hdr, nosanitize=True, identifier='e')
    assert np.nanmax(np.abs(e.latsCenter.filled(np.nan) - m.latsCenter.filled(np.nan))) > 1e-3   # the models differ


def test_single_consumer_multi_gpu_sequence(env):
    """parallel.resampleSequenceMultiGPU hands the frames of all devices to one consumer in sequence
    order (the shape of getMappingSequence); on every visible device count the results equal the
    single-device sequence."""
    import torch
    from auromat_b200 import synthetic
    from auromat_b200.parallel import resampleSequenceMultiGPU
    from auromat_b200.pipeline import resampleSequence
    W, H, n = 300, 200, 9
    hdrs = synthetic.sequenceHeaders(n, W, H)
    imgs = [synthetic.issImage(W, H, seed=70 + i) for i in range(n)]
    expect = [(f.img, f.mapping.identifier) for f in resampleSequence(imgs, hdrs, arcsecPerPx=400)]
    for devices in ([0], list(range(torch.cuda.device_count()))):
        got = list(resampleSequenceMultiGPU(imgs, hdrs, devices=devices, arcsecPerPx=400))
        assert len(got) == n
        for f, (e, _) in zip(got, expect):
            assert np.array_equal(ma.getmaskarray(f.img), ma.getmaskarray(e))
            assert np.array_equal(f.img.filled(0), e.filled(0))
    with pytest.raises(ValueError):           # an error in a worker reaches the consumer
        sky = dict(hdrs[0])
        sky['CRVAL2'] = -hdrs[0]['CRVAL2']
        sky['CRVAL1'] = (hdrs[0]['CRVAL1'] + 180.0) % 360.0
        list(resampleSequenceMultiGPU(imgs[:3], [hdrs[0], sky, hdrs[2]], devices=[0], arcsecPerPx=400))
