"""Developer tool (GPU): time sanitise and outline statistics alone (warm caches) for the library at argv[1]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import auromat_b200._lib as L
if len(sys.argv) > 1: L.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
m = getMapping(img, hdr, identifier='p'); m.prefetch(True)
ctx = m.context; p = m.devicePlanes(); W, H = synthetic.D3S_W, synthetic.D3S_H
st = ctx.new_stats()


def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print('%s: stats %.1f us, sanitise %.1f us' % (sys.argv[1] if len(sys.argv) > 1 else 'default',
      timed(lambda: ctx.bbox_stats(W, H, p, st)), timed(lambda: ctx.sanitize(W, H, p))))
