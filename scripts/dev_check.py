"""Developer check (GPU): compare the CUDA path with the oracle on a reduced frame."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, numpy.ma as ma
import torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
from auromat_b200.resample import resample, resampleToDevice
import oracle.auromat_oracle as O

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (532, 354)
for fast in (False, True):
    hdr = synthetic.issHeader(W, H)
    img = synthetic.issImage(W, H)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    m = getMapping(img, hdr, fastCenterCalculation=fast, identifier='x')
    t0 = time.time(); m.prefetch(True); torch.cuda.synchronize(); print('prefetch', time.time() - t0)
    g = O.georeference(hdr, cam, t, 110, fast_center=fast)
    if not fast:
        mk, mc = O.sanitize_masks(np.isnan(g['lats']), np.isnan(g['latsCenter']))
        for n in ('lats', 'lons', 'mlat', 'mlt'): g[n][mk] = np.nan
        for n in ('latsCenter', 'lonsCenter', 'mlatCenter', 'mltCenter', 'elevation'): g[n][mc] = np.nan
    pairs = dict(lats=m.lats, lons=m.lons, latsCenter=m.latsCenter, lonsCenter=m.lonsCenter, elevation=m.elevation,
                 mlat=m.mLatMlt[0], mlt=m.mLatMlt[1], mlatCenter=m.mLatMltCenter[0], mltCenter=m.mLatMltCenter[1])
    for n, a in pairs.items():
        a = a.filled(np.nan); b = g[n]
        mm = np.sum(np.isnan(a) != np.isnan(b))
        d = np.nanmax(np.abs(a - b))
        print('fast=%s %-11s maxabs %.3e  maskmismatch %d  valid %d' % (fast, n, d, mm, np.sum(~np.isnan(b))))
    print('ill', m.illConditionedCount, 'bbox', m.boundingBox)
    print('oracle bbox', O.bounding_box(g['lats'], g['lons']))
    # resample with GPU lat/lon fed to the oracle -> bit exact expected
    geo = {k: v.filled(np.nan) for k, v in pairs.items()}
    for ppd in [(36.0, 20.7), 5]:
        r = resample(m, pxPerDeg=ppd)
        o = O.resample_frame(geo, img, 110, px_per_deg=ppd, return_count=True)
        grid, info, oi, om, oe = resampleToDevice(m, pxPerDeg=ppd)
        cnt = info['count'].cpu().numpy().reshape(grid.ny, grid.nx)
        print('ppd', ppd, 'grid', r.img.shape, o['img'].shape, 'count equal', np.array_equal(cnt, o['count']),
              'img equal', np.array_equal(r.img.filled(0), np.where(o['img_mask'], 0, o['img'])),
              'mask equal', np.array_equal(ma.getmaskarray(r.img), o['img_mask']),
              'elev maxrel', np.nanmax(np.abs(r.elevation.filled(np.nan) - o['elevation']) / np.abs(o['elevation'])),
              'lat eq', np.array_equal(r.lats.filled(np.nan), o['lats']), np.array_equal(r.lons.filled(np.nan), o['lons']),
              np.array_equal(r.latsCenter.data, o['latsCenter']), np.array_equal(r.lonsCenter.data, o['lonsCenter']))
    r = resample(m, arcsecPerPx=100)
    print('arcsec100 grid', r.img.shape)
print('launches', m.context.launch_count)
