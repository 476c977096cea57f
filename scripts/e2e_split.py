"""Where does the e2e overhead over the device-resident pipeline come from?  ms/frame for the four
combinations of host/device input and host/device output."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from auromat_b200 import synthetic
from auromat_b200.pipeline import resampleSequence
from auromat_b200.runtime import get_context

ctx = get_context(0)
n = 60
hdr = synthetic.issHeader()
host = torch.from_numpy(synthetic.issImage()).pin_memory().numpy()
dev = torch.from_numpy(host).to(ctx.torch_device)


def run(img, toHost, sparse=True):
    for f in resampleSequence([img] * n, [hdr] * n, arcsecPerPx=100, magnetic=True, toHost=toHost, ringBuffers=True,
                              sparseUpload=sparse):
        pass
    torch.cuda.synchronize()


for name, img, toHost, sparse in (("device in, device out", dev, False, True), ("device in, host out", dev, True, True),
                                  ("host in (sparse), device out", host, False, True),
                                  ("host in (sparse), host out", host, True, True),
                                  ("host in (full), host out", host, True, False)):
    run(img, toHost, sparse)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        run(img, toHost, sparse)
        best = min(best, (time.perf_counter() - t0) * 1e3 / n)
    print("%-32s %.4f ms/frame" % (name, best))
