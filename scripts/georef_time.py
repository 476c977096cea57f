"""Developer tool (GPU): time the georeference kernel alone (all 9 planes) for the library at argv[1]."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auromat_b200._lib as L
if len(sys.argv) > 1: L.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from auromat_b200 import synthetic
from auromat_b200.coordinates.wcs import frameConstants
from auromat_b200.runtime import get_context
fast = len(sys.argv) > 2 and sys.argv[2] == 'fast'
ctx = get_context(0); W, H = 4256, 2832
hdr = synthetic.issHeader(); t, cam = synthetic.headerTimeAndCamera(hdr)
fr = frameConstants(hdr, cam, t, 110, fast)
nk, nc = (W + 1) * (H + 1), W * H
planes = {n: ctx.empty(nk if n.endswith('_k') else nc, torch.float64) for n in ('lat_k','lon_k','mlat_k','mlt_k','lat_c','lon_c','mlat_c','mlt_c','elev_c')}
planes['valid_k'], planes['valid_c'] = ctx.new_bitmaps(W, H)
for _ in range(3): ctx.georef(fr, planes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ctx.georef(fr, planes)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print('%s georef %.1f us  -> %.0f GB/s algorithmic, %.2f Gpix/s' % (sys.argv[1] if len(sys.argv) > 1 else 'default', ms * 1e3, 72.0 * nc / ms / 1e6, nc / ms / 1e6))
