"""Per-frame cost of the N>1 bench workload (cyclic 32-frame window of the configs[3] sequence)
against the N=1 workload (the configs[1] frame repeated), on one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
from auromat_b200.runtime import get_context

ctx = get_context(0)
W, H = synthetic.D3S_W, synthetic.D3S_H
img = torch.from_numpy(synthetic.issImage(W, H)).to(ctx.torch_device)
seq = synthetic.sequenceHeaders(32, W, H)
base = synthetic.issHeader(W, H)


def run(hdrs):
    last = None
    for f in resampleSequence([img] * len(hdrs), hdrs, arcsecPerPx=100, magnetic=True, toHost=False, device=0,
                              ringBuffers=True):
        last = f
    return last


def timed(hdrs, reps=3):
    run(hdrs[:8])
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(hdrs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / len(hdrs))
    return best


print("base frame x40        : %.4f ms/frame" % timed([base] * 40))
for world in (2, 4, 8):
    hd = [seq[(s * world) % 32] for s in range(40)]
    print("window, world=%d rank0 : %.4f ms/frame" % (world, timed(hd)))
for i in (0, 8, 16, 24, 31):
    m = run([seq[i]] * 4)
    st = m.mapping._deviceStats()
    print("seq[%2d]: valid centres %.3f  %.4f ms/frame" % (i, st.n_valid_centers / (W * H), timed([seq[i]] * 30)))
