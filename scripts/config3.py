"""BASELINE configs[2]: 24-Mpix frame (6000x4000) with SIP order-4 distortion resampled to a fine
10 arcsec/px grid (scatter-contention / footprint stress).  Prints per-stage CUDA-event times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
from auromat_b200.resample import resampleToDevice

W, H = 6000, 4000
hdr = synthetic.issHeader(W, H, sipOrder=4)
img = torch.from_numpy(synthetic.issImage(W, H)).cuda()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(3):
    t0 = ev()
    m = getMapping(img, hdr, identifier='c3'); m.prefetch(True)
    t1 = ev()
    bb = m.boundingBox
    t2 = ev()
    grid, info, oi, om, oe = resampleToDevice(m, arcsecPerPx=10)
    t3 = ev(); torch.cuda.synchronize()
    cnt = info['count']
    print('iter %d: georef+sanitize %.3f ms, stats %.3f ms, zero+bin+normalise %.3f ms | grid %dx%d = %.1f Mcells, '
          'valid px %d, filled cells %d, max count %d' % (it, t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3),
          grid.ny, grid.nx, grid.nx * grid.ny / 1e6, int(cnt.sum().item()), int((cnt > 0).sum().item()), int(cnt.max().item())))
npx = W * H
print('total %.1f Mpix/s' % (npx / (t0.elapsed_time(t3) * 1e-3) / 1e6))
