"""Developer tool (GPU): per-frame host cost of the pipelined sequence = time per frame on a tiny frame."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
W, H = 128, 96
hdr = synthetic.issHeader(W, H); img = torch.from_numpy(synthetic.issImage(W, H)).cuda()
def run(n, **kw):
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=400, toHost=False, ringBuffers=True, **kw): pass
    torch.cuda.synchronize()
himg = torch.from_numpy(synthetic.issImage(W, H)).pin_memory().numpy()
def run_e2e(n):
    for f in resampleSequence([himg] * n, [hdr] * n, arcsecPerPx=400, toHost=True, ringBuffers=True, magnetic=True): pass
    torch.cuda.synchronize()
run_e2e(20)
t0 = time.perf_counter(); run_e2e(400); print('e2e host floor %.1f us/frame' % ((time.perf_counter() - t0) / 400 * 1e6))
for kw in (dict(magnetic=True), dict(magnetic=False, coordinates=False)):
    run(20, **kw)
    t0 = time.perf_counter(); run(400, **kw); dt = (time.perf_counter() - t0) / 400
    print(kw, 'host floor %.1f us/frame' % (dt * 1e6))
if '--profile' in sys.argv:
    pr = cProfile.Profile(); pr.enable(); run_e2e(300); pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(70)
