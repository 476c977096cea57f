"""Developer tool (GPU): time the binning kernel alone on the configs[1] frame for the library at argv[1]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auromat_b200._lib as L
if len(sys.argv) > 1: L.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
from auromat_b200.resample import resampleToDevice
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
m = getMapping(img, hdr, identifier='p'); m.prefetch(True)
grid, info, oi, om, oe = resampleToDevice(m, arcsecPerPx=100)
ctx = m.context; p = m.devicePlanes(); cells = grid.nx * grid.ny
acc = ctx.zeros(5 * cells, torch.int64)
cnt, sums, fsum = acc[:cells], acc[cells:4 * cells], acc[4 * cells:].view(torch.float64)
for _ in range(3): ctx.bin_accumulate(p['lat_c'], p['lon_c'], p['elev_c'], img, grid, cnt, sums, fsum)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ctx.bin_accumulate(p['lat_c'], p['lon_c'], p['elev_c'], img, grid, cnt, sums, fsum)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
n = 4256 * 2832
print('%s bin %.1f us -> %.0f GB/s algorithmic (27 B/px), grid %dx%d' % (sys.argv[1] if len(sys.argv) > 1 else 'default', ms * 1e3, 27.0 * n / ms / 1e6, grid.ny, grid.nx))
ok = torch.equal(cnt, 23 * info['count'])
print('counts consistent with resampleToDevice:', ok)
