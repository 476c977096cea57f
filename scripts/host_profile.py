"""Developer tool (GPU): where does the host time of one device-resident step go?"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.mapping.spacecraft import getMapping
from auromat_b200.resample import resampleToDevice
hdr = synthetic.issHeader(); img = torch.from_numpy(synthetic.issImage()).cuda()
def step():
    m = getMapping(img, hdr, identifier='p'); m.prefetch(True)
    return resampleToDevice(m, arcsecPerPx=100)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize(); print('ms/step', (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
