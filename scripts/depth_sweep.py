"""Developer tool (GPU): e2e ms/frame of a 20-frame and a 200-frame sequence for pipeline depths 2..4."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
hdr = synthetic.issHeader()
host = torch.from_numpy(synthetic.issImage()).pin_memory().numpy()


def run(n, depth):
    t0 = time.perf_counter()
    for f in resampleSequence([host] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=True, ringBuffers=True,
                              depth=depth):
        pass
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


for depth in (2, 3, 4):
    run(8, depth)
    a = sorted(run(20, depth) for _ in range(5))
    b = sorted(run(200, depth) for _ in range(3))
    print('depth %d: 20 frames %.4f (median of 5, min %.4f)   200 frames %.4f ms/frame' % (depth, a[2], a[0], b[1]))
