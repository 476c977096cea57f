"""Device timeline of the sequence engine (AMT_SEQ_TRACE=1, amt_seq_trace): per frame the start / end of
stage A, of the image upload, of the fused kernel and the completion of the results, for the device-input
and the host-input (e2e) variants of the bench step.  Run on the B200 box:
    AMT_SEQ_TRACE=1 python scripts/seq_trace.py [frames]
"""
import os
import sys
import time

os.environ.setdefault('AMT_SEQ_TRACE', '1')
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auromat_b200 import synthetic                      # noqa: E402
from auromat_b200.pipeline import resampleSequence      # noqa: E402
from auromat_b200.runtime import get_context            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
W, H = 4256, 2832
ctx = get_context(0)
hdr = synthetic.issHeader(W, H)
img_np = synthetic.issImage(W, H, 1)
img_host = torch.from_numpy(img_np).pin_memory()
img_dev = img_host.to(ctx.torch_device)


def run(src, toHost, **kw):
    tr = {}
    for _ in range(2):
        tr.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f in resampleSequence([src] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=toHost, device=0,
                                  ringBuffers=True, transferStats=tr, **kw):
            pass
        f._finish()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
    return tr['trace'], wall


for name, src, toHost in (('device input', img_dev, False), ('host input (e2e)', img_host.numpy(), True)):
    trace, wall = run(src, toHost)
    base = trace[0]['a0']
    print('== %s: wall %.3f ms/frame' % (name, wall))
    print('frame   A:start   A:end |  up:start  up:end (dur) |   k:start   k:end (dur) |   out   | period')
    prev = None
    for i, t in enumerate(trace):
        r = {k: v - base for k, v in t.items()}
        print('%5d %9.3f %7.3f | %9.3f %7.3f (%.3f) | %9.3f %7.3f (%.3f) | %7.3f | %s' % (
            i, r['a0'], r['a1'], r['up0'], r['up1'], r['up1'] - r['up0'], r['k0'], r['k1'], r['k1'] - r['k0'], r['out'],
            '%.3f' % (r['k1'] - prev) if prev is not None else ''))
        prev = r['k1']
    k = np.array([t['k1'] - t['k0'] for t in trace[5:]])
    u = np.array([t['up1'] - t['up0'] for t in trace[5:]])
    per = np.diff(np.array([t['k1'] for t in trace[5:]]))
    print('median: kernel %.3f ms, upload window %.3f ms, period %.3f ms' % (np.median(k), np.median(u), np.median(per)))
