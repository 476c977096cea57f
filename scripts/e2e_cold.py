"""Developer tool (GPU): reproduce bench.py's e2e leg in a fresh process (device leg first, then
3 warm-up + 20 timed frames) and print the wall-clock time at which every frame is yielded."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
from auromat_b200.runtime import get_context
ctx = get_context(0)
hdr = synthetic.issHeader()
host = torch.from_numpy(synthetic.issImage()).pin_memory()
dev = host.to(ctx.torch_device)


def run(img, n, toHost, log=None):
    t0 = time.perf_counter()
    last = None
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=toHost, ringBuffers=True):
        last = f
        if log is not None:
            log.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


run(dev, 3, False); print('device leg: %.4f ms/frame' % run(dev, 20, False))
run(host.numpy(), 3, True)
log = []
print('e2e leg   : %.4f ms/frame' % run(host.numpy(), 20, True, log))
print('yield times (ms):', ' '.join('%.2f' % t for t in log))
log = []
print('e2e again : %.4f ms/frame' % run(host.numpy(), 20, True, log))
print('yield times (ms):', ' '.join('%.2f' % t for t in log))
