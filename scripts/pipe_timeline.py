"""Device-side timeline of the pipelined sequence (CUDA events per phase), e2e mode:
A = georeference+sanitise+stats (main stream), H = image H2D (copy stream), B = zero/bin/normalise,
D = result D2H (second stream).  Prints start/end in ms relative to the first event."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence

sparse = '--full' not in sys.argv
n = int(os.environ.get('N', 12))
hdr = synthetic.issHeader()
img = torch.from_numpy(synthetic.issImage()).pin_memory().numpy()


def run(tr):
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=True, ringBuffers=True,
                              sparseUpload=sparse, transferStats=tr):
        pass
    torch.cuda.synchronize()


run({}); run({})
tr = {'trace': []}
run(tr)
ev0 = min((e for _, _, e in tr['trace']), key=lambda e: -e.elapsed_time(tr['trace'][0][2]))
rows = {}
for tag, i, e in tr['trace']:
    rows.setdefault(i, {})[tag] = ev0.elapsed_time(e)
import time
t0 = time.perf_counter(); run({}); print("wall %.4f ms/frame over %d frames" % ((time.perf_counter() - t0) * 1e3 / n, n))
print("frame   A0     A1  |  H0     H1  |  B0     B1     D1")
for i in sorted(rows):
    r = rows[i]
    print("%3d  %6.3f %6.3f | %6.3f %6.3f | %6.3f %6.3f %6.3f" % (i, r.get('A0', 0), r.get('A1', 0), r.get('H0', 0),
                                                               r.get('H1', 0), r.get('B0', 0), r.get('B1', 0), r.get('D1', 0)))
