import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
def run(n, tag):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ts = []
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=False, ringBuffers=True):
        ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(tag, 'total %.2f ms' % (ts[-1] * 1e3), ' '.join('%.2f' % (t * 1e3) for t in ts[:24]), 'mem GB', torch.cuda.memory_reserved() / 1e9)
run(3, 'warm3'); run(20, 'run20'); run(20, 'run20b'); run(3, 'w3'); run(20, 'run20c')
