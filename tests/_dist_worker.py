"""Worker of tests/test_distributed_cpu.py: world_size-2 gloo run of the host-side multi-GPU
logic (frame sharding, bounding-box gather, sum/count all-reduce) with CPU tensors.  The
per-station grids come from the oracle's histogram so the expected mosaic is known exactly."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle.auromat_oracle as O  # noqa: E402
from auromat_b200 import parallel  # noqa: E402
from auromat_b200.mapping.mapping import BoundingBox  # noqa: E402
from auromat_b200.resample import targetGrid  # noqa: E402


def station(i):
    rng = np.random.default_rng(100 + i)
    lat0, lon0 = 60 + 2 * i, -120 + 7 * i
    lat = lat0 + rng.uniform(-3, 3, 5000)
    lon = lon0 + rng.uniform(-6, 6, 5000)
    img = rng.integers(0, 65536, (5000, 1)).astype(np.float64)
    elev = rng.uniform(1, 90, 5000)
    return lat, lon, img, elev


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert parallel.worldInfo() == (rank, world)
    n_st = 5
    mine = parallel.shardIndices(n_st)
    assert mine == list(range(rank, n_st, world))
    assert parallel.shardSequence("abcde") == [c for i, c in enumerate("abcde") if i % world == rank]
    data = {i: station(i) for i in range(n_st)}
    boxes = [BoundingBox(float(data[i][0].min()), float(data[i][1].min()), float(data[i][0].max()),
                         float(data[i][1].max())) for i in mine]
    allb = parallel.gatherBoundingBoxes(boxes)
    assert len(allb) == n_st
    bb = BoundingBox.mergedBoundingBoxes(allb)
    grid, info = targetGrid((4.0, 2.0), bb.latSouth, bb.latNorth, bb.lonWest, bb.lonEast)
    bins = (grid.nx, grid.ny)
    rng_ = [[grid.lo_x, grid.hi_x], [grid.lo_y, grid.hi_y]]

    def hist(i):
        lat, lon, img, elev = data[i]
        return O.histogram2d_weighted(lon, lat, bins, rng_, [None, img[:, 0], elev])

    cells = grid.nx * grid.ny
    acc = torch.zeros(2 * cells, dtype=torch.int64)
    fsum = torch.zeros(cells, dtype=torch.float64)
    for i in mine:
        c, s, e = hist(i)
        acc[:cells] += torch.from_numpy(c.T[::-1].copy().astype(np.int64).ravel())
        acc[cells:] += torch.from_numpy(s.T[::-1].copy().astype(np.int64).ravel())
        fsum += torch.from_numpy(e.T[::-1].copy().ravel())
    parallel.allreduceGrids([acc, fsum])
    tc = sum(hist(i)[0] for i in range(n_st)).T[::-1]
    ts = sum(hist(i)[1] for i in range(n_st)).T[::-1]
    te = sum(hist(i)[2] for i in range(n_st)).T[::-1]
    assert np.array_equal(acc[:cells].numpy().reshape(grid.ny, grid.nx), tc)          # exact
    assert np.array_equal(acc[cells:].numpy().reshape(grid.ny, grid.nx), ts)          # exact
    assert np.allclose(fsum.numpy().reshape(grid.ny, grid.nx), te, rtol=1e-12, atol=0)
    assert tc.sum() > 20000
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
