"""Launches k_bin (the unfused binning kernel) and k_georef_fused once per scatter regime, for ncu
(VERDICT r1 item 8): (i) configs[1] frame at 100 arcsec/px, (ii) the same frame at pxPerDeg=5 (heavy
same-cell contention), (iii) configs[2] (24 Mpix, SIP) at 10 arcsec/px (~20 M cells, one sample per
cell).  Run:

    ncu --metrics <atomics/L2 metrics> -k regex:'k_bin|k_georef_fused' --csv --log-file gpurun_out/scatter.csv \
        python scripts/bin_scatter.py
Without ncu it prints CUDA-event times of the same launches.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auromat_b200 import synthetic                                            # noqa: E402
from auromat_b200.mapping.spacecraft import getMapping                        # noqa: E402
from auromat_b200.resample import deriveGrid                                  # noqa: E402
from auromat_b200.runtime import get_context                                  # noqa: E402


def main():
    ctx = get_context(0)
    reps = int(os.environ.get("REPS", "1"))
    for name, (W, H, sip), kw in (("100arcsec", (4256, 2832, 0), dict(arcsecPerPx=100)),
                                  ("pxPerDeg5", (4256, 2832, 0), dict(pxPerDeg=5)),
                                  ("10arcsec_sip", (6000, 4000, 4), dict(arcsecPerPx=10))):
        hdr = synthetic.issHeader(W, H, sipOrder=sip)
        img = ctx.to_device(synthetic.issImage(W, H, 1))
        m = getMapping(img, hdr, identifier=name)
        p = m.devicePlanes(magnetic=True)
        grid, info = deriveGrid(m, **kw)
        cells = grid.nx * grid.ny
        acc = ctx.zeros(5 * cells, torch.int64)
        parts = (acc[:cells], acc[cells:4 * cells], acc[4 * cells:].view(torch.float64))
        fr = m.frameConstants
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(reps):
            ctx.bin_accumulate(p['lat_c'], p['lon_c'], p['elev_c'], img, grid, *parts)
        e[1].record()
        for _ in range(reps):
            ctx.georef_fused(fr, p['valid_k'], p['valid_c'], planes=p, img=img, grid=grid, count=parts[0],
                             sums=parts[1], fsum=parts[2])
        e[2].record()
        torch.cuda.synchronize()
        print("%-14s grid %dx%d = %d cells, %d valid px: k_bin %.1f us, k_georef_fused %.1f us" % (
            name, grid.nx, grid.ny, cells, int(m._deviceStats().n_valid_centers),
            e[0].elapsed_time(e[1]) / reps * 1e3, e[1].elapsed_time(e[2]) / reps * 1e3))
        del m, p, acc, img


if __name__ == "__main__":
    main()
