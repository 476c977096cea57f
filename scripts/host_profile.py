"""Developer tool (GPU): host-side cost of the pipelined sequence (device-resident inputs)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
def run(n):
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=False): pass
    torch.cuda.synchronize()
run(5)
t0 = time.perf_counter(); run(100); print('ms/frame', (time.perf_counter() - t0) / 100 * 1e3)
pr = cProfile.Profile(); pr.enable(); run(100); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(32)
