"""Developer tool (GPU): CUDA-event times of the plane-free pipeline stages on the configs[1] frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
from auromat_b200.resample import resampleToDevice
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
W, H = 4256, 2832
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
m = getMapping(img, hdr, identifier='v').setPlaneFree(True)
ctx = m.context; fr = m.frameConstants
vk, vc = ctx.new_bitmaps(W, H)
print('hits only   %.1f us' % timeit(lambda: ctx.georef(fr, {'valid_k': vk, 'valid_c': vc})))
print('sanitize    %.1f us' % timeit(lambda: ctx.sanitize(W, H, {'valid_k': vk, 'valid_c': vc})))
st = ctx.new_stats()
print('stats_frame %.1f us' % timeit(lambda: ctx.bbox_stats_frame(fr, vk, vc, st)))
grid, info, oi, om, oe = resampleToDevice(m, arcsecPerPx=100)
cells = grid.nx * grid.ny
acc = ctx.zeros(5 * cells, torch.int64)
cnt, sums, fsum = acc[:cells], acc[cells:4 * cells], acc[4 * cells:].view(torch.float64)
print('fused bin   %.1f us' % timeit(lambda: ctx.georef_bin_fused(fr, vc, img, grid, cnt, sums, fsum)))
print('fused (no elevation) %.1f us' % timeit(lambda: ctx.georef_bin_fused(fr, vc, img, grid, cnt, sums, None)))
