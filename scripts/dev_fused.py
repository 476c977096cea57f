"""GPU development check of the round-2 kernels (run on the B200 box):
  * limb solver bitmaps == per-pixel hit-test bitmaps over many camera geometries,
  * fused georeference kernel planes == amt_georef + amt_sanitize planes (bit for bit),
  * fused binning == amt_bin_accumulate on those planes (counts, integer sums, fixed-point sums),
  * timings of every kernel involved (CUDA events, 20 repetitions).
"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auromat_b200 import synthetic                                     # noqa: E402
from auromat_b200.mapping.spacecraft import getMapping                        # noqa: E402
from auromat_b200.resample import targetGrid, sideScale, plateCarreeResolution  # noqa: E402
from auromat_b200.runtime import get_context                                  # noqa: E402

ctx = get_context(0)
dev = ctx.torch_device


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def bitmaps(frame, W, H, limb):
    if limb:
        os.environ.pop('AMT_NO_LIMB_SOLVER', None)
    else:
        os.environ['AMT_NO_LIMB_SOLVER'] = '1'
    bits = {}
    bits['valid_k'], bits['valid_c'] = ctx.new_bitmaps(W, H)
    st = ctx.new_stats()
    ctx.georef(frame, bits, st)
    s = ctx.read_stats(st)
    os.environ.pop('AMT_NO_LIMB_SOLVER', None)
    return bits, int(s.n_ill_conditioned)


def headers(W, H):
    yield 'iss', synthetic.issHeader(W, H)
    rng = np.random.default_rng(5)
    base = synthetic.issHeader(W, H)
    for i in range(24):
        h = dict(base)
        th = rng.uniform(0, 2 * math.pi)
        s = math.hypot(base['CD1_1'], base['CD1_2']) * rng.uniform(0.5, 2.5)
        h['CD1_1'], h['CD1_2'], h['CD2_1'], h['CD2_2'] = -s * math.cos(th), -s * math.sin(th), s * math.sin(th), -s * math.cos(th)
        h['CRVAL1'] = base['CRVAL1'] + rng.uniform(-40, 40)
        h['CRVAL2'] = base['CRVAL2'] + rng.uniform(-40, 40)
        yield 'rand%d' % i, h
    yield 'pole', synthetic.issHeaderLookingAt(80.0, 10.0, 89.5, 40.0, W, H)
    yield 'dateline', synthetic.issHeaderLookingAt(50.0, 170.0, 55.0, -178.0, W, H)
    yield 'nadir', synthetic.issHeaderLookingAt(50.0, 10.0, 50.0, 10.01, W, H)


def main():
    ok = True
    # ---- 1. limb solver vs per-pixel hit test
    for (W, H) in ((532, 354), (1064, 708), (4256, 2832)):
        for name, hdr in headers(W, H):
            t, cam = synthetic.headerTimeAndCamera(hdr)
            m = getMapping(np.zeros((H, W, 3), np.uint8), hdr, identifier=name)
            fr = m.frameConstants
            a, ga = bitmaps(fr, W, H, True)
            b, gb = bitmaps(fr, W, H, False)
            same = torch.equal(a['valid_k'], b['valid_k']) and torch.equal(a['valid_c'], b['valid_c'])
            nk = int(sum(bin(int(v) & 0xffffffff).count('1') for v in b['valid_k'].cpu().numpy()[::97]))
            frac = float((b['valid_c'] != 0).float().mean())
            if not same or ga != gb:
                ok = False
            if not same or ga != gb or (W == 532):
                print('limb %-9s %dx%d same=%s graze %d/%d nonzero-words %.2f' % (name, W, H, same, ga, gb, frac))
    print('LIMB', 'OK' if ok else 'MISMATCH')

    # ---- 2/3. fused kernel vs georef + sanitize (+ bin), full size
    W, H = 4256, 2832
    for name, hdr in (('iss', synthetic.issHeader(W, H)), ('sip', synthetic.issHeader(W, H, sipOrder=4))):
        img = synthetic.issImage(W, H)
        m = getMapping(img, hdr, identifier=name)
        fr = m.frameConstants
        dimg = ctx.to_device(img)
        nk, nc = (W + 1) * (H + 1), W * H
        names = ('lat_k', 'lon_k', 'mlat_k', 'mlt_k', 'lat_c', 'lon_c', 'mlat_c', 'mlt_c', 'elev_c')
        old = {n: ctx.empty(nk if n.endswith('_k') else nc, torch.float64) for n in names}
        old['valid_k'], old['valid_c'] = ctx.new_bitmaps(W, H)
        ctx.georef(fr, old)
        ctx.sanitize(W, H, old)
        bits, _ = bitmaps(fr, W, H, True)
        ctx.sanitize(W, H, bits)
        print(name, 'sanitised bitmaps equal:', torch.equal(bits['valid_k'], old['valid_k']),
              torch.equal(bits['valid_c'], old['valid_c']))
        new = {n: torch.full_like(old[n], 7.0) for n in names}
        # grid from the bounding box of the old path
        st = ctx.new_stats()
        ctx.bbox_stats(W, H, old, st)
        s = ctx.read_stats(st)
        st2 = ctx.new_stats()
        ctx.bbox_stats_frame(fr, bits['valid_k'], bits['valid_c'], st2)
        s2 = ctx.read_stats(st2)
        print(name, 'stats equal:', [getattr(s, f) == getattr(s2, f) for f in ('lat_min', 'lat_max', 'lon_min', 'lon_max', 'n_valid_centers', 'n_boundary_corners')])
        from auromat_b200.mapping.mapping import BoundingBox
        bb = BoundingBox(s.lat_min, s.lon_min, s.lat_max, s.lon_max)
        for arcsec in (100, 10):
            ppd = plateCarreeResolution(bb, arcsec)
            grid, info = targetGrid(ppd, s.lat_min, s.lat_max, s.lon_min, s.lon_max)
            grid.side_scale = sideScale(W * H)
            cells = grid.nx * grid.ny
            accA = ctx.zeros(5 * cells, torch.int64)
            accB = ctx.zeros(5 * cells, torch.int64)

            def parts(acc):
                return acc[:cells], acc[cells:4 * cells], acc[4 * cells:].view(torch.float64)
            ctx.bin_accumulate(old['lat_c'], old['lon_c'], old['elev_c'], dimg, grid, *parts(accA))
            ctx.georef_fused(fr, bits['valid_k'], bits['valid_c'], planes=new, img=dimg, grid=grid,
                             count=parts(accB)[0], sums=parts(accB)[1], fsum=parts(accB)[2])
            torch.cuda.synchronize()
            if arcsec == 100:
                for n in names:
                    a, b = old[n], new[n]
                    eq = torch.equal(torch.isnan(a), torch.isnan(b))
                    d = torch.nan_to_num(a - b).abs().max().item()
                    print('  plane %-7s nan-pattern equal %s  max|diff| %.3g' % (n, eq, d))
            print('  %s %d"/px grid %dx%d: accumulators equal: %s (count sum %d)' % (
                name, arcsec, grid.nx, grid.ny, torch.equal(accA, accB), int(accA[:cells].sum())))
            # plane-free
            accC = ctx.zeros(5 * cells, torch.int64)
            ctx.georef_fused(fr, None, bits['valid_c'], img=dimg, grid=grid, count=parts(accC)[0], sums=parts(accC)[1],
                             fsum=parts(accC)[2])
            print('  plane-free equal:', torch.equal(accA, accC))
            # timings
            t = {}
            t['bin (planes)'] = timeit(lambda: ctx.bin_accumulate(old['lat_c'], old['lon_c'], old['elev_c'], dimg, grid, *parts(accA)))
            t['fused planes+mag+bin'] = timeit(lambda: ctx.georef_fused(fr, bits['valid_k'], bits['valid_c'], planes=new, img=dimg, grid=grid, count=parts(accB)[0], sums=parts(accB)[1], fsum=parts(accB)[2]))
            t['fused bin only'] = timeit(lambda: ctx.georef_fused(fr, None, bits['valid_c'], img=dimg, grid=grid, count=parts(accC)[0], sums=parts(accC)[1], fsum=parts(accC)[2]))
            t['zero acc'] = timeit(lambda: accA.zero_())
            print('  timings %d"/px [us]:' % arcsec, {k: round(v, 1) for k, v in t.items()})
        t = {}
        t['georef old (9 planes)'] = timeit(lambda: ctx.georef(fr, old))
        t['fused planes+mag'] = timeit(lambda: ctx.georef_fused(fr, bits['valid_k'], bits['valid_c'], planes=new))
        os.environ['AMT_NO_ROW_PERMUTATION'] = '1'
        t['georef old, rows in image order'] = timeit(lambda: ctx.georef(fr, old))
        t['fused planes+mag, rows in image order'] = timeit(lambda: ctx.georef_fused(fr, bits['valid_k'], bits['valid_c'], planes=new))
        os.environ.pop('AMT_NO_ROW_PERMUTATION')
        nomag = {k: v for k, v in new.items() if not k.startswith('ml')}
        t['fused planes no mag'] = timeit(lambda: ctx.georef_fused(fr, bits['valid_k'], bits['valid_c'], planes=nomag))
        raw = {}
        raw['valid_k'], raw['valid_c'] = ctx.new_bitmaps(W, H)
        t['hit bits (limb solver if TAN)'] = timeit(lambda: ctx.georef(fr, raw))
        os.environ['AMT_NO_LIMB_SOLVER'] = '1'
        t['hit bits per pixel'] = timeit(lambda: ctx.georef(fr, raw))
        os.environ.pop('AMT_NO_LIMB_SOLVER')
        t['sanitize bitmaps'] = timeit(lambda: ctx.sanitize(W, H, raw))
        t['sanitize planes'] = timeit(lambda: ctx.sanitize(W, H, old))
        t['stats frame'] = timeit(lambda: ctx.bbox_stats_frame(fr, bits['valid_k'], bits['valid_c'], st2))
        t['stats planes'] = timeit(lambda: ctx.bbox_stats(W, H, old, st))
        print(name, 'timings [us]:', {k: round(v, 1) for k, v in t.items()})


if __name__ == '__main__':
    main()
