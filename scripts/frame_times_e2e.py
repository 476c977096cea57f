import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).pin_memory().numpy()
def run(n, tag, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = []
    last = None
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=True, **kw):
        ts.append(time.perf_counter() - t0); last = f
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(tag, 'total %.2f ms' % (ts[-1] * 1e3), ' '.join('%.2f' % (t * 1e3) for t in ts[:24]))
for ring in (True, False):
    run(3, 'warm3 ring=%s' % ring, ringBuffers=ring); run(20, 'run20', ringBuffers=ring); run(20, 'run20b', ringBuffers=ring)
