"""Multi-GPU check (run under torchrun, one rank per GPU): the NCCL all-reduced mosaic of
stations sharded over the ranks equals, bit for bit in counts / integer sums / rounded
means, the mosaic one rank computes from all stations alone.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/mosaic_nccl_check.py
"""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auromat_b200 import parallel  # noqa: E402
from auromat_b200.mapping.allsky import AllSkyMapping, CalibrationData  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    n, w = 64, 256
    rng = np.random.default_rng(2)
    cals = [CalibrationData('S%02d' % i, 0, 0, float(rng.uniform(55, 70)), float(rng.uniform(-160, -60)),
                            256.0, 256.0, 155.81, 0.0, None) for i in range(n)]
    imgs = [np.random.default_rng(50 + i).integers(0, 65536, (w, w, 1), dtype=np.uint16) for i in range(n)]
    t = datetime.datetime(2012, 3, 4, 17, 19, 0)

    def build(idx):
        return [AllSkyMapping(cals[i], imgs[i], t, 110, device=local).maskedByElevation(1) for i in idx]

    mine = build(parallel.shardIndices(n))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    mos, acc = parallel.mosaic(mine, pxPerDeg=(20, 20))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        everything = build(range(n))
        grid, info = acc.grid, acc.info          # the common grid every rank derived from the gathered boxes
        ref = parallel.MosaicAccumulator(grid, info, 1, torch.uint16, everything[0].context)
        for m in everything:
            ref.add(m)
        assert torch.equal(ref.acc, acc.acc), "count / integer sums differ"
        rel = ((ref.fsum - acc.fsum).abs() / ref.fsum.abs().clamp_min(1e-300)).max().item()
        assert rel < 1e-12, rel
        img_ref = ref.finalise(everything[0]).img
        assert np.array_equal(img_ref.filled(0), mos.img.filled(0))
        print("mosaic ok: world=%d stations=%d grid=%dx%d cells, %d samples, mosaic(bin+allreduce+normalise) %.3f ms, "
              "allreduce message %.1f MB" % (world, n, grid.ny, grid.nx, int(acc.count.sum().item()), ms.item(),
                                             (acc.acc.numel() + acc.fsum.numel()) * 8 / 1e6))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
