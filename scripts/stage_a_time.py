"""Developer tool (GPU): serial time of the stage-A launches of one frame (georeference, sanitise,
outline statistics) and of stage B (zero, bin, normalise), each timed with CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
from auromat_b200.resample import resampleToDevice
from auromat_b200.runtime import get_context
ctx = get_context(0)
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
W, H = synthetic.D3S_W, synthetic.D3S_H


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


acc = {}
for rep in range(8):
    m = getMapping(img, hdr, identifier='p')
    planes_needed = ['lat_k', 'lon_k', 'lat_c', 'lon_c', 'elev_c', 'mlat_k', 'mlt_k', 'mlat_c', 'mlt_c']
    torch.cuda.synchronize()
    e0 = ev()
    m._sanitize_saved = m._sanitize
    m._sanitize = False
    m._computePlanes(ctx, planes_needed)
    e1 = ev()
    ctx.sanitize(W, H, m._planes)
    e2 = ev()
    m._startStats()
    e3 = ev()
    st = m._deviceStats()
    torch.cuda.synchronize()
    e4 = ev()
    resampleToDevice(m, arcsecPerPx=100)
    e5 = ev()
    torch.cuda.synchronize()
    if rep >= 3:
        for k, a, b in (('georef', e0, e1), ('sanitise', e1, e2), ('stats', e2, e3), ('zero+bin+normalise', e4, e5)):
            acc.setdefault(k, []).append(a.elapsed_time(b) * 1e3)
for k, v in acc.items():
    print('%-20s %.1f us' % (k, min(v)))
