"""Multi-GPU use of the path: one process per GPU (torchrun), `torch.distributed` for plumbing.

* Image sequences shard by frame -- frames are independent (reference
  `mapping/spacecraft.py:308-332`), so rank r takes frames r, r+N, r+2N, ... and there is NO
  data-path collective.
* Multi-camera mosaics (a `MappingCollection` of all-sky stations on one common plate-carree
  grid; "all resamplings are aligned to the same global grid", reference `resample.py:220-221`)
  have exactly one exchange step: every rank bins its stations into private sum/count grids
  and the grids are summed with ONE all-reduce (NCCL over NVLink on GPUs; gloo in the CPU
  tests).  Counts and integer channel sums are int64, so the reduction is exact and
  independent of the reduction order; only the float side channel (elevation) is summed in
  floating point.
"""
from __future__ import annotations

import numpy as np
import numpy.ma as ma

from . import _lib
from .mapping.mapping import BoundingBox, GenericMapping


def worldInfo():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def bindToLocalCpus(device_index):
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA
    node / PCIe root) so that pinned host buffers are allocated next to the GPU and the launch
    thread does not migrate.  Matters once several ranks stream images over PCIe at the same
    time.  Returns the affinity set, or None if NVML is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shardIndices(n, rank=None, world=None):
    """Indices of the frames of an n-frame sequence owned by `rank`: i with i mod world == rank."""
    if rank is None or world is None:
        rank, world = worldInfo()
    return list(range(rank, n, world))


def shardSequence(items, rank=None, world=None):
    items = list(items)
    return [items[i] for i in shardIndices(len(items), rank, world)]


def gatherBoundingBoxes(localBoxes, group=None):
    """All ranks' bounding boxes (list of BoundingBox), via all_gather_object."""
    import torch.distributed as dist
    rank, world = worldInfo()
    local = [(b.latSouth, b.lonWest, b.latNorth, b.lonEast) for b in localBoxes]
    if world == 1:
        gathered = [local]
    else:
        gathered = [None] * world
        dist.all_gather_object(gathered, local, group=group)
    return [BoundingBox(*t) for part in gathered for t in part]


def allreduceGrids(acc, group=None):
    """Sum the accumulator tensor(s) over all ranks, in place.  `acc` is the int64 tensor
    holding count | channel sums, or a list of tensors (the float64 side sums separately)."""
    import torch.distributed as dist
    _, world = worldInfo()
    if world == 1:
        return acc
    for t in (acc if isinstance(acc, (list, tuple)) else [acc]):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return acc


class MosaicAccumulator(object):
    """Sum / count grids of a common plate-carree target grid on one device.

    grid: `amt_grid` from `resample.targetGrid`; channels: image channels; the layout of the
    int64 accumulator is count | sums[channels] (planes of ny*nx cells, row 0 = north), the
    float64 accumulator holds the elevation sums."""

    def __init__(self, grid, info, channels, imgDtype, context):
        import torch
        self.grid, self.info, self.channels, self.imgDtype, self.ctx = grid, info, channels, imgDtype, context
        self.cells = grid.nx * grid.ny
        # one allocation: count | sums[channels] | elevation sums (int64 fixed point when
        # grid.side_scale > 0, else float64 bits)
        self.all = context.zeros((2 + channels) * self.cells, torch.int64)
        self.acc = self.all[:(1 + channels) * self.cells]
        self.fsum = self.all[(1 + channels) * self.cells:].view(torch.float64)

    count = property(lambda self: self.acc[:self.cells])
    sums = property(lambda self: self.acc[self.cells:])
    exactSide = property(lambda self: self.grid.side_scale > 0)

    def add(self, mapping):
        """Bin one mapping into the grids (`amt_bin_accumulate`)."""
        from .resample import binMappingInto
        assert mapping.deviceImage().shape[2] == self.channels
        binMappingInto(mapping, self.grid, self.count, self.sums, self.fsum)

    def allreduce(self, group=None):
        """ONE message when the elevation sums are integers too (exact, order independent); else the
        float plane is reduced on its own."""
        if self.exactSide:
            allreduceGrids(self.all, group)
        else:
            allreduceGrids([self.acc, self.fsum], group)

    def finalise(self, template, deviceDone=None):
        """Normalise and wrap the mosaic as a GenericMapping (metadata from `template`).
        `deviceDone`: optional CUDA event recorded once the device part (normalise, grid coordinates,
        back-rotation) is enqueued; what follows is the host-side wrapping (download + masked array)."""
        ctx, g, info = self.ctx, self.grid, self.info
        outImg, outMask, outElev = ctx.normalise(g, self.imgDtype, self.channels, self.count, self.sums, self.fsum)
        lat_k, lon_k, lat_c, lon_c = ctx.plate_carree_coords(g.nx, g.ny, info['latMaxInGrid'], info['latMinInGrid'],
                                                             info['lonMinInGrid'], info['lonMaxInGrid'])
        mode = info.get('mode', _lib.AMT_PRE_NONE)
        if mode != _lib.AMT_PRE_NONE:
            from .resample import _preRotation
            back = _preRotation(mode, template.altitude, angle=-90)     # reference resample.py:262-277
            ctx.rotate_coords(lat_k, lon_k, back)
            ctx.rotate_coords(lat_c, lon_c, back)
        if deviceDone is not None:
            deviceDone.record()
        img = ctx.to_numpy(outImg)
        mask = ctx.to_numpy(outMask).astype(bool)
        img = ma.masked_array(img, mask=np.repeat(mask[:, :, None], img.shape[2], 2))
        return GenericMapping(lat_k, lon_k, lat_c, lon_c, outElev, template.altitude, img, template.cameraPosGCRS,
                              template.photoTime, 'mosaic', device=ctx.device)


def mosaic(mappings, pxPerDeg, group=None, timings=None):
    """Compose the mappings held by ALL ranks (each rank passes its own subset) into one
    mean-binned mosaic on a common grid; every rank returns the full mosaic mapping.

    Steps: all-gather the per-mapping bounding boxes -> common `fixedGrid` on every rank ->
    local binning -> one all-reduce of the sum/count grids -> normalise.

    :param timings: optional dict that receives the CUDA-event times of this rank in ms:
        `bin_ms`, `allreduce_ms`, `normalise_ms` (normalise + target-grid coordinates on the device),
        `wrap_ms` (download of image and mask, numpy masked array, GenericMapping) and their sum
        `total_ms` (synchronises)"""
    from .resample import targetGrid
    mappings = list(mappings)
    assert mappings, 'every rank needs at least one mapping'
    try:
        _, _ = pxPerDeg
    except TypeError:
        pxPerDeg = (pxPerDeg, pxPerDeg)
    boxes = gatherBoundingBoxes([m.boundingBox for m in mappings], group)
    bb = BoundingBox.mergedBoundingBoxes(boxes)
    # pixels of ALL members on all ranks (bound of the fixed-point elevation sums) and whether every
    # member's elevation is device-computed (finite)
    nSamples = sum(m.shape[0] * m.shape[1] for m in mappings)
    exact = all(getattr(m, '_finiteElevation', False) for m in mappings)
    if worldInfo()[1] > 1:
        import torch.distributed as dist
        parts = [None] * worldInfo()[1]
        dist.all_gather_object(parts, (nSamples, exact), group=group)
        nSamples, exact = sum(p[0] for p in parts), all(p[1] for p in parts)
    mode = _lib.AMT_PRE_NONE
    latMin, latMax, lonMin, lonMax = bb.latSouth, bb.latNorth, bb.lonWest, bb.lonEast
    if bb.containsPole or any(b.containsPole for b in boxes):
        # Same trick as resample.py:176-201, applied to every member: all coordinates are rotated
        # by 90 deg about the x axis, which moves the pole onto the equator of the rotated frame;
        # the common grid is laid over the min/max of the members' ROTATED outlines (reduced on
        # the device, exchanged as four floats per mapping), and rotated back in finalise().
        from .resample import _preRotation
        mode = _lib.AMT_PRE_POLE
        local = []
        for m in mappings:
            ctx, (h, w) = m.context, m.shape
            st = ctx.new_stats()
            ctx.bbox_stats(w, h, m.devicePlanes(), st, pre=_preRotation(mode, mappings[0].altitude))
            s = ctx.read_stats(st)
            local.append((s.lat_min, s.lon_min, s.lat_max, s.lon_max))
        _, world = worldInfo()
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, local, group=group)
            local = [t for part in gathered for t in part]
        latMin, lonMin = min(t[0] for t in local), min(t[1] for t in local)
        latMax, lonMax = max(t[2] for t in local), max(t[3] for t in local)
        if lonMax - lonMin > 180:
            raise NotImplementedError('the rotated mosaic still spans more than 180 degrees of longitude')
    elif bb.containsDiscontinuity:
        # same trick as resample.py:203-218, applied to every member: rotate the longitudes by
        # 180 deg so that the mosaic does not straddle the date line; rotated back in finalise()
        from .mapping.mapping import wrapAt180
        mode = _lib.AMT_PRE_WRAP180
        lonMin, lonMax = wrapAt180(bb.lonWest + 180), wrapAt180(bb.lonEast + 180)
    grid, info = targetGrid(pxPerDeg, latMin, latMax, lonMin, lonMax, mode, mappings[0].altitude)
    info['mode'] = mode
    if exact:
        from .resample import sideScale
        grid.side_scale = sideScale(nSamples)
    m0 = mappings[0]
    img0 = m0.deviceImage()
    acc = MosaicAccumulator(grid, info, img0.shape[2], img0.dtype, m0.context)
    ev = None
    if timings is not None:
        import torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
    for m in mappings:
        acc.add(m)
    if ev:
        ev[1].record()
    acc.allreduce(group)
    if ev:
        ev[2].record()
    result = acc.finalise(m0, ev[3] if ev else None)
    if ev:
        ev[4].record()
        ev[4].synchronize()
        timings.update(bin_ms=ev[0].elapsed_time(ev[1]), allreduce_ms=ev[1].elapsed_time(ev[2]),
                       normalise_ms=ev[2].elapsed_time(ev[3]), wrap_ms=ev[3].elapsed_time(ev[4]),
                       total_ms=ev[0].elapsed_time(ev[4]))
    return result, acc


def resampleSequenceMultiGPU(imagesOrArrays, wcsHeaders, devices=None, queueDepth=8, **kw):
    """`pipeline.resampleSequence` over several GPUs of one process, frames handed to ONE consumer in
    sequence order -- the shape of the reference's `getMappingSequence` (mapping/spacecraft.py:308-332),
    which yields one mapping after the other to its caller.

    One worker thread per device runs the pipelined sequence of its shard (frame i -> device
    i mod N; one `amt_ctx` and one sequence engine per device, no state shared between them) and
    pushes finished frames into a bounded queue; the generator pops the queues round-robin, so the
    order is that of the input.  The library calls release the GIL; the per-frame Python work
    (header -> frame constants, result objects) does not, which caps a single process at a few
    thousand frames per second -- beyond that use one process per GPU (torchrun, `shardSequence`).

    :param devices: CUDA device indices (default: all visible)
    :param queueDepth: finished frames a worker may hold before it blocks
    :param kw: as for `pipeline.resampleSequence` (`device` is set per worker)
    """
    import queue
    import threading
    import torch
    from .pipeline import resampleSequence
    images, headers = list(imagesOrArrays), list(wcsHeaders)
    metadatas = kw.pop('metadatas', None)
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    n = len(devices)
    assert n >= 1
    queues = [queue.Queue(maxsize=queueDepth) for _ in devices]
    stop = threading.Event()
    _END = object()

    def worker(k):
        q = queues[k]
        try:
            idx = shardIndices(len(headers), k, n)
            with torch.cuda.device(devices[k]):
                opts = dict(kw, device=devices[k])
                if metadatas:
                    opts['metadatas'] = [metadatas[i] for i in idx]
                for f in resampleSequence([images[i] for i in idx], [headers[i] for i in idx], **opts):
                    while not stop.is_set():
                        try:
                            q.put(f, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
            q.put(_END)
        except BaseException as e:        # handed to the consumer, in order
            q.put(e)

    threads = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(n)]
    for t in threads:
        t.start()
    try:
        for i in range(len(headers)):
            item = queues[i % n].get()
            if isinstance(item, BaseException):
                raise item
            assert item is not _END
            yield item
    finally:
        stop.set()
        for t in threads:
            t.join(timeout=5)
