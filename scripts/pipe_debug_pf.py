import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import auromat_b200.pipeline as P
from auromat_b200 import synthetic
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
# monkeypatch timing into getMapping / resampleToDevice
T = []
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T.append((name, (time.perf_counter() - t0) * 1e3)); return r
    setattr(mod, name, g)
wrap(P, 'getMapping'); wrap(P, 'resampleToDevice')
import auromat_b200.mapping.mapping as MM
for n in ('prefetch', '_startStats'):
    f = getattr(MM.BaseMapping, n)
    def mk(f, n):
        def g(self, *a, **k):
            t0 = time.perf_counter(); r = f(self, *a, **k); T.append((n, (time.perf_counter() - t0) * 1e3)); return r
        return g
    setattr(MM.BaseMapping, n, mk(f, n))
f0 = P.ResampledFrame._startDownload
def sd(self, *a, **k):
    t0 = time.perf_counter(); r = f0(self, *a, **k); T.append(('startDownload', (time.perf_counter() - t0) * 1e3)); return r
P.ResampledFrame._startDownload = sd
f1 = P.ResampledFrame._finish
def fi(self, *a, **k):
    t0 = time.perf_counter(); r = f1(self, *a, **k); T.append(('finish', (time.perf_counter() - t0) * 1e3)); return r
P.ResampledFrame._finish = fi
def run(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for f in P.resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, toHost=False, ringBuffers=True, coordinates=False, magnetic=False):
        T.append(('YIELD', (time.perf_counter() - t0) * 1e3))
    torch.cuda.synchronize()
run(3); run(8); del T[:]; run(8)
print(' '.join('%s=%.2f' % t for t in T))
