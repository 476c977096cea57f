"""Developer tool (GPU): cProfile of the host side of the e2e sequence loop."""
import sys, os, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
hdr = synthetic.issHeader()
host = torch.from_numpy(synthetic.issImage()).pin_memory().numpy()
n = 300


def run():
    for f in resampleSequence([host] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=True, ringBuffers=True):
        pass
    torch.cuda.synchronize()


run()
pr = cProfile.Profile()
pr.enable(); run(); pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s).sort_stats('tottime')
st.print_stats(38)
txt = s.getvalue()
print('\n'.join(l[:150] for l in txt.splitlines()[:60]))
print('per frame: total %.3f ms' % (st.total_tt / n * 1e3))
