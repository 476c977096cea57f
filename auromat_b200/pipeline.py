"""Software-pipelined getMapping + resample over an image sequence on one GPU.

Two implementations share the public generator `resampleSequence`:

* WCS frames with their own centre rays (the default of `getMapping`) run on the C sequence engine
  (`amt_seq_*`, csrc/amt_seq.cuh): per frame TWO library calls -- stage A (hit bitmaps by the limb
  solver, sanitisation and outline statistics on the bitmaps) and, once the 104-byte statistics
  block is on the host and the grid is derived, stage B (upload of the defined pixel box, ONE fused
  kernel that writes the coordinate planes and bins the centres, normalise, download).
* fastCenterCalculation frames keep the Python-orchestrated multi-stream pipeline described below.


`getMappingSequence` of the reference (mapping/spacecraft.py:308-332) yields one mapping at a
time and `ResampleProvider` (resample.py:370-394) maps `resample` over it; every frame is
independent.  `resampleSequence` is the same composition with the host-visible phases of a
frame interleaved across frames so that the GPU never waits for the host:

  A(i)   host: per-frame constants; enqueue the georeference kernel (caller's stream, back to back
         from frame to frame) and, on an auxiliary high-priority stream, sanitise + outline
         statistics + the asynchronous read-back of the 104-byte statistics block
  B(i-d) host: bounding box -> target grid (needs the statistics of frame i-d, long finished);
         enqueue the upload of the image rows that hold valid pixels (copy stream), zero / bin /
         normalise (second high-priority stream) and the D2H of the results (their own stream)
  C      host: hand the finished frame to the caller

The only device->host dependency of the path (grid size depends on the footprint) is thereby
hidden behind the georeferencing of the next `depth` frames.  Results are identical to calling
`resample(getMapping(...))` frame by frame.
"""
from __future__ import annotations

import collections
import ctypes
import os
import weakref

import numpy as np
import numpy.ma as ma

from . import _lib
from .mapping.spacecraft import getMapping
from .resample import resampleToDevice


class _Waiter(object):
    def __init__(self, fn):
        self.synchronize = fn


class ResampledFrame(object):
    """Result of one frame: the georeferenced mapping (device planes) and its resampling.
    With toHost=True the resampled image / mask / elevation have already been copied into pinned
    host buffers when the frame is yielded; `img` and `elevation` return arrays the caller owns.
    `toMapping()` builds the same GenericMapping `resample()` returns."""

    def __init__(self, mapping, grid, info, dImg, dMask, dElev):
        self.mapping, self.grid, self.info = mapping, grid, info
        self.deviceImg, self.deviceMask, self.deviceElevation = dImg, dMask, dElev
        self._hostFlat = None
        self._event = None

    def _startDownload(self, ctx, pool, depthHint=2, stream=None):
        """Async D2H of image | mask | elevation into ONE pooled pinned byte buffer.  Buffers are
        pooled by capacity class (next power of two), not by shape: the grid size changes from
        frame to frame and page-locking (cudaHostAlloc) is a millisecond-scale call."""
        import torch
        parts = (self.deviceImg, self.deviceMask, self.deviceElevation)
        flatDev = getattr(self.deviceImg, '_amt_flat', None)      # the three parts are views of one buffer
        if flatDev is not None:
            base = flatDev.data_ptr()
            offs = [p.data_ptr() - base for p in parts]
            sizes = [p.numel() * p.element_size() for p in parts]
            total = flatDev.numel()
        else:
            sizes = [p.numel() * p.element_size() for p in parts]
            offs = [0]
            for n in sizes[:-1]:
                offs.append((offs[-1] + n + 63) // 64 * 64)
            total = offs[-1] + sizes[-1]
        cap = 1 << max(16, (total - 1).bit_length())
        bucket = pool.setdefault(cap, [])
        if not bucket and cap not in pool.setdefault('_seen', set()):
            pool['_seen'].add(cap)
            bucket.extend(torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(2 * depthHint + 6))
        if not bucket:
            pool['grown'] = pool.get('grown', 0) + 1         # page-locking inside a sequence: visible in diagnostics
        flat = bucket.pop() if bucket else torch.empty(cap, dtype=torch.uint8).pin_memory()
        if flatDev is not None:
            ctx.copy_d2h(flat.data_ptr(), flatDev.data_ptr(), total, ctx.stream())      # one cudaMemcpyAsync
        else:
            with torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(ctx.torch_device)):
                for p, o, n in zip(parts, offs, sizes):
                    flat[o:o + n].view(p.dtype).view(p.shape).copy_(p, non_blocking=True)
        # host views are built on first access
        self._hostFlat = (flat, offs, sizes, [p.dtype for p in parts], [tuple(p.shape) for p in parts])
        weakref.finalize(self, bucket.append, flat)    # the buffer returns to the pool with the frame
        self._event = torch.cuda.Event()
        if stream is not None:
            for p in (flatDev,) if flatDev is not None else parts:
                p.record_stream(stream)             # allocated on another stream, read by this one
            self._event.record(stream)
        else:
            self._event.record()

    @property
    def _host(self):
        hf = self.__dict__.get('_hostFlat')
        if hf is None:
            return None
        if '_hostViews' not in self.__dict__:
            flat, offs, sizes, dtypes, shapes = hf
            self.__dict__['_hostViews'] = tuple(flat[o:o + n].view(dt).view(sh)
                                                for o, n, dt, sh in zip(offs, sizes, dtypes, shapes))
        return self.__dict__['_hostViews']

    def _attachHost(self, flat, offs, sizes, pool_bucket, waiter):
        """Results arrive in `flat` (pinned) through the sequence engine; `waiter()` blocks until
        the copy is complete."""
        parts = (self.deviceImg, self.deviceMask, self.deviceElevation)
        self._hostFlat = (flat, offs, sizes, [p.dtype for p in parts], [tuple(p.shape) for p in parts])
        weakref.finalize(self, pool_bucket.append, flat)     # the buffer returns to the pool with the frame
        self._event = _Waiter(waiter)

    def _finish(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None

    @property
    def img(self):
        """Masked (ny, nx, n) resampled image on the host."""
        self._finish()
        if self._host is None:
            ctx = self.mapping.context
            img, mask = ctx.to_numpy(self.deviceImg), ctx.to_numpy(self.deviceMask)
        else:
            img, mask = self._host[0].numpy().copy(), self._host[1].numpy()
            if img.dtype == np.int16:
                img = img.view(np.uint16)
        mask = mask.astype(bool)
        return ma.masked_array(img, mask=np.repeat(mask[:, :, None], img.shape[2], 2))

    @property
    def elevation(self):
        self._finish()
        e = self._host[2].numpy().copy() if self._host is not None else \
            self.mapping.context.to_numpy(self.deviceElevation)
        return ma.masked_invalid(e)

    def toMapping(self):
        """The GenericMapping on the plate-carree grid that `resample()` would return."""
        self._finish()
        m, g, info = self.mapping, self.grid, self.info
        ctx = m.context
        lat_k, lon_k, lat_c, lon_c = ctx.plate_carree_coords(g.nx, g.ny, info['latMaxInGrid'], info['latMinInGrid'],
                                                             info['lonMinInGrid'], info['lonMaxInGrid'])
        if info['mode'] != _lib.AMT_PRE_NONE:
            from .resample import _preRotation
            back = _preRotation(info['mode'], m.altitude, angle=-90)
            ctx.rotate_coords(lat_k, lon_k, back)
            ctx.rotate_coords(lat_c, lon_c, back)
        return m.createResampled(lat_k, lon_k, lat_c, lon_c, self.deviceElevation, self.img)


def resampleSequence(imagesOrArrays, wcsHeaders, pxPerDeg=25, arcsecPerPx=None, altitude=110,
                     fastCenterCalculation=False, magnetic=False, metadatas=None, depth=3, toHost=True,
                     device=None, ringBuffers=False, coordinates=True, sparseUpload=True, transferStats=None):
    """Generator of `ResampledFrame` for an image sequence (frames in order).

    :param imagesOrArrays: iterable of (h,w,n) uint8/uint16 arrays (ideally pinned), device
                           tensors or image paths
    :param wcsHeaders: iterable of header dicts or `.wcs` paths, same length
    :param magnetic: also produce the MLat/MLT planes of every frame
    :param depth: frames whose georeferencing is enqueued ahead of the host-side grid derivation
        (>= 1; 1 disables the overlap, 3 keeps the GPU fed through host jitter)
    :param toHost: copy the resampled image / mask / elevation to pinned host buffers
    :param coordinates: False = plane-free mode: only the resampling is wanted, the per-pixel
        coordinate planes are never written (hit bitmaps -> outline statistics -> fused
        georeference+binning kernel); `frame.mapping` computes them lazily if asked.  Ignored
        (treated as True) with fastCenterCalculation or magnetic=True.
    :param sparseUpload: host images are copied to the device only after the frame has been
        georeferenced, and only the row range that holds georeferenced pixels (for an ISS limb
        frame ~60 % of the image): pixels that see no Earth never influence the result.
        `frame.mapping.img` still is the complete host image.
    :param transferStats: optional dict that receives `h2d_bytes` (image bytes actually copied)
    :param ringBuffers: keep the coordinate planes of the frames in a fixed ring of
        depth + AHEAD_B + KEEP_FRAMES + 1 plane sets (0.9 GB each for a 12-Mpix frame) instead of
        allocating per frame: constant memory footprint for arbitrarily long sequences.  The planes
        of `frame.mapping` are valid while KEEP_FRAMES + 1 = 3 further frames are taken from the
        generator; when its ring slot is recycled the mapping is detached from the ring and
        recomputes its planes on access (the resampled outputs of a frame are never recycled)
    """
    import torch
    from .runtime import get_context
    ctx = get_context(device)
    if not fastCenterCalculation and os.environ.get('AMT_PIPELINE', 'engine') != 'python':
        yield from _engineSequence(ctx, imagesOrArrays, wcsHeaders, pxPerDeg, arcsecPerPx, altitude, magnetic,
                                   metadatas, depth, toHost, ringBuffers, coordinates, sparseUpload, transferStats)
        return
    main = torch.cuda.current_stream(ctx.torch_device)
    # one copy stream, one image ring and one pinned-buffer pool per context: they survive
    # across sequences (torch caches device blocks per stream)
    copy = ctx.__dict__.get('_copy_stream')
    if copy is None:
        copy = ctx.__dict__['_copy_stream'] = torch.cuda.Stream(ctx.torch_device)
    stageA, stageB = collections.deque(), collections.deque()
    pool = ctx.__dict__.setdefault('_pinned_frames', {})
    imgRing = ctx.__dict__.setdefault('_image_ring', [])
    metadatas = metadatas if metadatas else None
    stats = transferStats if transferStats is not None else {}
    stats.setdefault('h2d_bytes', 0)
    trace = stats.get('trace')          # optional list: (tag, frame, timing event) per pipeline phase

    def mark(tag, i, stream):
        if trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            trace.append((tag, i, ev))

    def upload(i, img, rows=None, cols=None):
        """Host array -> device on the copy stream; returns (tensor, event).  `rows` / `cols` =
        (first, last) restrict the copy to the pixel box that holds georeferenced pixels: the
        rest of the image is never read by the binning kernel."""
        src = np.ascontiguousarray(img)
        tdtype = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.int16}[src.dtype]
        if ringBuffers:
            if len(imgRing) != ringLen or tuple(imgRing[0].shape) != src.shape or imgRing[0].dtype != tdtype:
                del imgRing[:]
                with torch.cuda.stream(copy):
                    imgRing.extend(torch.empty(src.shape, dtype=tdtype, device=ctx.torch_device)
                                   for _ in range(ringLen))
            d = imgRing[i % len(imgRing)]
            prev = slotDone.get(i % len(imgRing))
            if prev is not None:
                copy.wait_event(prev)           # the frame that used this ring slot has been binned
            else:
                copy.wait_stream(main)
        else:
            with torch.cuda.stream(copy):
                d = torch.empty(src.shape, dtype=tdtype, device=ctx.torch_device)
            d.record_stream(main)
            if second is not None:
                d.record_stream(second)
        mark('H0', i, copy)
        rowBytes = src.strides[0]
        r0, r1 = (0, src.shape[0] - 1) if rows is None else rows
        nrows = max(0, r1 - r0 + 1)
        c0, c1 = (0, src.shape[1] - 1) if cols is None else cols
        pxBytes = src.strides[1]
        if nrows and c1 >= c0 and (c1 - c0 + 1) * 10 < src.shape[1] * 9:
            # the valid pixels sit in a narrow column band (limb roughly vertical): copy the box
            widthBytes = (c1 - c0 + 1) * pxBytes
            off = r0 * rowBytes + c0 * pxBytes
            ctx.copy_h2d_2d(d.data_ptr() + off, src.ctypes.data + off, rowBytes, widthBytes, nrows, hCopy)
            nbytes = widthBytes * nrows
        else:
            nbytes = nrows * rowBytes
            if nbytes:
                # plain cudaMemcpyAsync on the copy stream (pinned source => asynchronous)
                ctx.copy_h2d(d.data_ptr() + r0 * rowBytes, src.ctypes.data + r0 * rowBytes, nbytes, hCopy)
        ev = torch.cuda.Event()
        ev.record(copy)
        mark('H1', i, copy)
        stats['h2d_bytes'] += nbytes
        keepAlive.append((ev, src))             # the host array must outlive the asynchronous copy
        while len(keepAlive) > ringLen + 2:
            keepAlive.popleft()
        if src.dtype == np.uint16:
            d = d.view(torch.uint16)
        return d, ev

    import ctypes
    hMain = ctypes.c_void_p(main.cuda_stream)
    hCopy = ctypes.c_void_p(copy.cuda_stream)
    keepAlive = collections.deque()
    ring = []
    freeStats = []
    slotDone = {}            # ring slot -> event recorded after the binning of its last user
    slotOwner = {}           # ring slot -> weakref of the mapping that shows its planes
    # frame i is handed out in iteration i + 2*depth - 1 and its planes stay valid while KEEP_FRAMES + 1
    # further frames are taken: the slot must not be reused before iteration i + 2*depth + KEEP_FRAMES + 1
    ringLen = 2 * depth + KEEP_FRAMES + 2
    # With ring buffers the phases of a frame run on separate streams (see the module docstring):
    # the georeference kernel on the caller's stream, sanitise / statistics on `aux`, zero / bin /
    # normalise on `second`, uploads on `copy`, result downloads on `dout`, so that the small
    # kernels and launch gaps hide under the long kernel of the following frames.  (Without ring
    # buffers the planes are per-frame allocations of the caller's stream and everything stays
    # on it.)
    second = None
    if ringBuffers:
        second = ctx.__dict__.get('_second_stream')
        if second is None:
            # high priority: the short stage-B kernels get SM slots as soon as blocks of the long
            # georeference kernel retire, instead of queueing behind all of its waves
            second = ctx.__dict__['_second_stream'] = torch.cuda.Stream(ctx.torch_device, priority=-1)
    hSecond = ctypes.c_void_p(second.cuda_stream) if second is not None else None
    aux = hAux = None
    if ringBuffers:
        aux = ctx.__dict__.get('_aux_stream')
        if aux is None:
            aux = ctx.__dict__['_aux_stream'] = torch.cuda.Stream(ctx.torch_device, priority=-1)
        hAux = ctypes.c_void_p(aux.cuda_stream)
    dout = ctx.__dict__.get('_dout_stream')
    if dout is None:
        dout = ctx.__dict__['_dout_stream'] = torch.cuda.Stream(ctx.torch_device)
    hDout = ctypes.c_void_p(dout.cuda_stream)

    def ringSet(i, m):
        h, w = m.shape
        if not ring or ring[0]['lat_k'].numel() != (h + 1) * (w + 1):
            del ring[:]
            names = ['lat_k', 'lon_k', 'lat_c', 'lon_c', 'elev_c'] + \
                    (['mlat_k', 'mlt_k', 'mlat_c', 'mlt_c'] if magnetic else [])
            for _ in range(ringLen):
                s = {n: ctx.empty((h + 1) * (w + 1) if n.endswith('_k') else h * w, torch.float64) for n in names}
                s['valid_k'], s['valid_c'] = ctx.new_bitmaps(w, h)
                s['_out'], s['_nplanes'] = ctx.out_struct(s), len(s)
                s['_stats'] = ctx.new_stats()
                ring.append(s)
        return ring[i % len(ring)]

    def runA(i, img, hdr):
        # Georeferencing needs the header only.  A host image is uploaded in stage B, once the
        # statistics of the frame tell which rows hold georeferenced pixels (`sparseUpload`), or
        # right away on the copy stream.
        ev = None
        late = sparseUpload and isinstance(img, np.ndarray)
        if isinstance(img, np.ndarray) and not late:
            dimg, ev = upload(i, img)
        else:
            dimg = img
        meta = metadatas[i] if metadatas else None
        m = getMapping(dimg, hdr, altitude=altitude, fastCenterCalculation=fastCenterCalculation, metadata=meta,
                       identifier=None if isinstance(hdr, str) else 'frame%06d' % i, device=ctx.device)
        if late:
            m._lateImage = img
        elif sparseUpload and isinstance(img, str):
            m._lateImage = m.img_unmasked        # image file: decoded on the host, uploaded like an array
        planeFree = not coordinates and not magnetic and not fastCenterCalculation
        if planeFree:
            m.setPlaneFree(True)
            if aux is None:
                m._startStats()
                return m, ev, i
            if not freeStats:                   # ring-owned statistics blocks (zeroed once, here)
                freeStats.extend(ctx.new_stats() for _ in range(ringLen))
            m._statsDevice = freeStats[i % ringLen]
        if ringBuffers and not planeFree:
            prev = slotDone.get(i % ringLen)
            if prev is not None:
                main.wait_event(prev)           # ring slot free: its previous frame has been binned
            old = slotOwner.get(i % ringLen)
            old = old() if old is not None else None
            if old is not None:
                old._detachRing()                  # it recomputes into fresh buffers if asked again
            m._planeBuffers = ringSet(i, m)
            m._ringSlot = (None, i % ringLen)
            slotOwner[i % ringLen] = weakref.ref(m)
            m._statsDevice = m._planeBuffers['_stats']      # ring-owned statistics block (no per-frame alloc)
        mark('A0', i, main)
        if aux is None:
            m.prefetch(magnetic=magnetic)
            m._startStats()
            mark('A1', i, main)
            return m, ev, i
        # The long georeference kernel stays alone on the caller's stream (back to back from frame
        # to frame); the short sanitise / statistics launches and the statistics read-back move to
        # an auxiliary high-priority stream and overlap the next frame's georeferencing.
        def toAux():
            evG = torch.cuda.Event()
            evG.record(main)
            mark('A1', i, main)
            aux.wait_event(evG)
            ctx.use_stream(hAux)
        m._afterGeoref = toAux
        with torch.cuda.stream(aux):
            ctx.use_stream(hMain)              # the georeference launch itself goes to the main stream
            if not planeFree:
                m.prefetch(magnetic=magnetic)
            m._startStats()                    # plane-free: hit-test launch, then sanitise + statistics
            evS = torch.cuda.Event()
            evS.record(aux)
        ctx.use_stream(hMain)
        m._statsReady = evS
        return m, ev, i

    def lateUpload(m, i):
        """Stage B of a frame whose host image has not been uploaded yet: copy the rows that hold
        at least one valid pixel (known from the outline statistics, already on the host)."""
        img = m.__dict__.pop('_lateImage', None)
        if img is None:
            return None
        st = m._deviceStats()
        dimg, ev = upload(i, img, rows=(st.row_min_c, st.row_max_c), cols=(st.col_min_c, st.col_max_c))
        m._imgDevice = dimg if dimg.dim() == 3 else dimg[..., None]
        return ev

    def runB(m, ev, i=0):
        late = lateUpload(m, i)
        ev = late if late is not None else ev
        if second is None:
            if ev is not None:
                main.wait_event(ev)
            grid, info, dImg, dMask, dElev = resampleToDevice(m, pxPerDeg=pxPerDeg, arcsecPerPx=arcsecPerPx)
            f = ResampledFrame(m, grid, info, dImg, dMask, dElev)
            if toHost:
                f._startDownload(ctx, pool, depth)
            return f
        evA = m.__dict__.pop('_statsReady', None)
        if evA is None:
            evA = torch.cuda.Event()
            evA.record(main)                    # everything of stage A of this frame
        with torch.cuda.stream(second):
            ctx.use_stream(hSecond)
            second.wait_event(evA)
            if ev is not None:
                second.wait_event(ev)
            mark('B0', i, second)
            grid, info, dImg, dMask, dElev = resampleToDevice(m, pxPerDeg=pxPerDeg, arcsecPerPx=arcsecPerPx)
            f = ResampledFrame(m, grid, info, dImg, dMask, dElev)
            # allocated under `second`, handed to the caller on `main`: the caching allocator must not
            # recycle these blocks while kernels of the caller's stream still read them
            for t in (getattr(dImg, '_amt_flat', None), info.get('count')):
                if t is not None:
                    t.record_stream(main)
            mark('B1', i, second)
            done = torch.cuda.Event()
            done.record(second)
            if toHost:
                # results leave on their own stream: the copy must not delay the next frame's binning
                dout.wait_event(done)
                ctx.use_stream(hDout)
                f._startDownload(ctx, pool, depth, stream=dout)
            mark('D1', i, dout if toHost else second)
        ctx.use_stream(hMain)
        slotDone[i % ringLen] = done
        f._done = done
        return f

    def finish(f):
        f._finish()
        done = getattr(f, '_done', None)
        if done is not None:
            main.wait_event(done)               # the caller's stream sees the finished outputs
        return f

    ctx.use_stream(hMain)
    tracing = bool(os.environ.get('AMT_SEQ_TRACE', '').strip('0'))

    def handOut(f):
        f._finish()
        if tracing:
            # device timeline of the frame (amt_seq_trace): ms since the engine was created
            t = (ctypes.c_double * 7)()
            _lib.check(lib.amt_seq_trace(eng.handle, f._slot, t))
            stats.setdefault('trace', []).append(dict(zip(
                ('a0', 'a1', 'up0', 'up1', 'k0', 'k1', 'out'), [float(v) for v in t])))
        return f

    try:
        launched = 0
        for i, (img, hdr) in enumerate(zip(imagesOrArrays, wcsHeaders)):
            stageA.append(runA(i, img, hdr))
            # the look-ahead builds up from one frame to `depth`: the first frame's upload and long kernel are
            # enqueued right after its own stage A instead of after `depth` host iterations (0.1 ms each)
            if len(stageA) >= min(depth, launched + 1):
                stageB.append(runB(*stageA.popleft()))
                launched += 1
            while len(stageB) > depth:
                ctx.use_stream(None)
                yield finish(stageB.popleft())
                ctx.use_stream(hMain)
        while stageA:
            stageB.append(runB(*stageA.popleft()))
        ctx.use_stream(None)
        while stageB:
            yield finish(stageB.popleft())
    finally:
        ctx.use_stream(None)


# --------------------------------------------------------------------------- engine path
KEEP_FRAMES = 2      # ring mode: the planes of a yielded frame stay valid while KEEP_FRAMES + 1 further frames
                     # are taken from the generator; then the slot is recycled and the mapping detached
AHEAD_B = 2          # frames whose stage B is enqueued before the oldest one is handed out


def _pinnedBuffer(pool, nbytes, hint):
    """A pinned byte buffer of at least nbytes from the per-context pool (capacity classes: the
    grid size changes from frame to frame and page-locking is a millisecond-scale call)."""
    import torch
    cap = 1 << max(16, (nbytes - 1).bit_length())
    bucket = pool.setdefault(cap, [])
    if not bucket and cap not in pool.setdefault('_seen', set()):
        pool['_seen'].add(cap)
        bucket.extend(torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(hint))
    if not bucket:
        pool['grown'] = pool.get('grown', 0) + 1         # page-locking inside a sequence: visible in diagnostics
    return (bucket.pop() if bucket else torch.empty(cap, dtype=torch.uint8).pin_memory()), bucket


class _Engine(object):
    """One `amt_seq` with its streams and ring buffers; cached per context and configuration (the
    rings survive across sequences)."""

    def __init__(self, ctx, w, h, channels, npdtype, nslots, magnetic, planes, ring, main):
        import torch
        self.ctx, self.w, self.h, self.channels, self.nslots = ctx, w, h, channels, nslots
        self.magnetic, self.planes, self.ring = magnetic, planes, ring
        self.tdtype = {np.dtype(np.uint8): torch.uint8, np.dtype(np.uint16): torch.uint16}[np.dtype(npdtype)]
        self.amtDtype = _lib.AMT_U8 if np.dtype(npdtype) == np.uint8 else _lib.AMT_U16
        dev = ctx.torch_device

        def stream(name, prio=0):
            st = ctx.__dict__.get(name)
            if st is None:
                st = ctx.__dict__[name] = torch.cuda.Stream(dev, priority=prio)
            return st
        self.main = main
        # stage A and the accumulator memset run at HIGH priority: at ordinary priority they would wait until the
        # long kernel has dispatched its last CTA (0.328 instead of 0.262 ms per frame,
        # profiles/r02_stage_a_interference.txt; AMT_SEQ_AUX_PRIORITY=0 reproduces that measurement)
        auxPriority = int(os.environ.get('AMT_SEQ_AUX_PRIORITY', '-1'))
        self.aux, self.copy, self.dout = stream('_aux_stream', auxPriority), stream('_copy_stream'), stream('_dout_stream')
        self.handle = ctypes.c_void_p()
        _lib.check(ctx.lib.amt_seq_create(ctx.handle, w, h, channels, self.amtDtype, nslots,
                                          ctypes.c_void_p(main.cuda_stream), ctypes.c_void_p(self.aux.cuda_stream),
                                          ctypes.c_void_p(self.copy.cuda_stream),
                                          ctypes.c_void_p(self.dout.cuda_stream), ctypes.byref(self.handle)))
        self.slots = [None] * nslots          # per slot: dict of tensors
        self.hasImage = [False] * nslots      # the slot's amt_seq_slot carries its image-ring pointer
        self.owner = [None] * nslots          # weakref to the mapping that currently shows the slot's planes
        self.hostStats = torch.zeros((nslots, ctypes.sizeof(_lib.AmtStats)), dtype=torch.uint8).pin_memory()
        self.devStats = torch.zeros((nslots, ctypes.sizeof(_lib.AmtStats)), dtype=torch.uint8, device=dev)
        self.imgRing = None

    def planeNames(self):
        if not self.planes:
            return []
        return ['lat_k', 'lon_k', 'lat_c', 'lon_c', 'elev_c'] + \
               (['mlat_k', 'mlt_k', 'mlat_c', 'mlt_c'] if self.magnetic else [])

    def newBuffers(self):
        import torch
        ctx, w, h = self.ctx, self.w, self.h
        b = {n: ctx.empty((h + 1) * (w + 1) if n.endswith('_k') else h * w, torch.float64) for n in self.planeNames()}
        b['valid_k'], b['valid_c'] = ctx.new_bitmaps(w, h)
        return b

    def setSlot(self, slot, buffers, needImage):
        import torch
        if needImage and self.imgRing is None:
            with torch.cuda.stream(self.copy):
                self.imgRing = torch.empty((self.nslots, self.h, self.w, self.channels), dtype=self.tdtype,
                                           device=self.ctx.torch_device)
        sl = _lib.AmtSeqSlot()
        sl.planes = self.ctx.out_struct(buffers)
        sl.d_stats = self.devStats[slot].data_ptr()
        sl.h_stats = self.hostStats[slot].data_ptr()
        sl.d_img = self.imgRing[slot].data_ptr() if self.imgRing is not None else None
        _lib.check(self.ctx.lib.amt_seq_set_slot(self.handle, slot, ctypes.byref(sl)))
        self.slots[slot] = buffers
        self.hasImage[slot] = self.imgRing is not None

    def __del__(self):
        try:
            if self.handle:
                self.ctx.lib.amt_seq_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _engineSequence(ctx, imagesOrArrays, wcsHeaders, pxPerDeg, arcsecPerPx, altitude, magnetic, metadatas, depth,
                    toHost, ringBuffers, coordinates, sparseUpload, transferStats):
    import torch
    from .resample import deriveGrid
    lib = ctx.lib
    main = torch.cuda.current_stream(ctx.torch_device)
    planes = bool(coordinates or magnetic)
    depth = max(1, int(depth))
    nslots = depth + AHEAD_B + KEEP_FRAMES + 1
    pool = ctx.__dict__.setdefault('_pinned_frames', {})
    engines = ctx.__dict__.setdefault('_engines', {})
    metadatas = metadatas if metadatas else None
    stats = transferStats if transferStats is not None else {}
    stats.setdefault('h2d_bytes', 0)
    stageA, stageB = collections.deque(), collections.deque()
    eng = None
    h2d0 = ctypes.c_uint64(0)

    def engineFor(w, h, channels, npdtype):
        key = (w, h, channels, np.dtype(npdtype).str, nslots, bool(magnetic), planes, bool(ringBuffers),
               main.cuda_stream)
        e = engines.get(key)
        if e is None:
            if len(engines) >= 4:                      # rings are large: keep a handful of configurations
                engines.pop(next(iter(engines)))
            e = engines[key] = _Engine(ctx, w, h, channels, npdtype, nslots, magnetic, planes, ringBuffers, main)
        return e

    def runA(i, img, hdr):
        nonlocal eng
        meta = metadatas[i] if metadatas else None
        m = getMapping(img, hdr, altitude=altitude, metadata=meta,
                       identifier=None if isinstance(hdr, str) else 'frame%06d' % i, device=ctx.device)
        hostImg = None
        if isinstance(img, str):
            hostImg = m.img_unmasked                    # image file: decoded on the host
        elif isinstance(img, np.ndarray):
            hostImg = img
        if hostImg is not None:
            hostImg = np.ascontiguousarray(hostImg if hostImg.ndim == 3 else hostImg[..., None])
            h, w, channels = hostImg.shape
            npdtype = hostImg.dtype
        else:
            dimg = img if img.dim() == 3 else img[..., None]
            h, w, channels = dimg.shape
            npdtype = np.uint8 if dimg.dtype == torch.uint8 else np.uint16
        if eng is None:
            eng = engineFor(w, h, channels, npdtype)
            lib.amt_seq_h2d_bytes(eng.handle, ctypes.byref(h2d0))
        elif (eng.w, eng.h, eng.channels) != (w, h, channels):
            raise ValueError('all frames of a sequence must have the same shape')
        slot = i % nslots
        # the frame that showed this slot's ring planes loses them (it recomputes on access)
        prev = eng.owner[slot]() if eng.owner[slot] is not None else None
        if prev is not None:
            prev._detachRing()
        if ringBuffers:
            if eng.slots[slot] is None:
                # the whole ring at once (8 plane sets, 7 GB for a 12-Mpix frame): a sequence reaches its steady
                # state with its first frame instead of paying a device allocation per slot over its first frames
                for s_ in range(nslots):
                    if eng.slots[s_] is None:
                        eng.setSlot(s_, eng.newBuffers(), hostImg is not None)
            elif hostImg is not None and not eng.hasImage[slot]:
                # an engine first used with device images gets its image ring with the first host image
                eng.setSlot(slot, eng.slots[slot], True)
        else:
            b = eng.newBuffers()                        # per-frame planes: valid as long as the frame lives
            for t in (b['valid_k'], b['valid_c']):
                t.record_stream(eng.aux)
            eng.setSlot(slot, b, hostImg is not None)
        _lib.check(lib.amt_seq_stage_a(eng.handle, slot, ctypes.byref(m.frameConstants)))
        return m, slot, hostImg

    try:
        pxLat, pxLon = pxPerDeg
    except TypeError:
        pxLat = pxLon = pxPerDeg
    arcsec = float(arcsecPerPx) if arcsecPerPx else 0.0

    def runB(m, slot, hostImg):
        # statistics -> pole flags -> bounding box -> target grid in C (amt_seq_plan); frames around a
        # pole / across the date line take the Python derivation (one more outline reduction)
        st, grid, ginfo, outcome = _lib.AmtStats(), _lib.AmtGrid(), _lib.AmtGridInfo(), ctypes.c_int32()
        _lib.check(lib.amt_seq_plan(eng.handle, slot, arcsec, float(pxLat or 0), float(pxLon or 0), ctypes.byref(st),
                                    ctypes.byref(grid), ctypes.byref(ginfo), ctypes.byref(outcome)))
        buffers = eng.slots[slot]
        # the mapping sees its final bitmaps; planes follow once the fused kernel is enqueued
        m._planes.update(valid_k=buffers['valid_k'], valid_c=buffers['valid_c'])
        m._grazingCounted = True
        m.isSanitized = True
        m._planeFree = True
        m._stats = st
        if outcome.value == _lib.AMT_PLAN_EMPTY:
            m.boundingBox       # nothing of the Earth in this frame: the error `resample(getMapping(...))` raises
        if outcome.value == _lib.AMT_PLAN_OK:
            grid.altitude = float(m.altitude)
            info = dict(nLat=ginfo.n_lat, nLon=ginfo.n_lon, latMinInGrid=ginfo.lat_min_in_grid,
                        latMaxInGrid=ginfo.lat_max_in_grid, lonMinInGrid=ginfo.lon_min_in_grid,
                        lonMaxInGrid=ginfo.lon_max_in_grid, latStep=ginfo.lat_step, lonStep=ginfo.lon_step,
                        pxPerDeg=(ginfo.lat_px_per_deg, ginfo.lon_px_per_deg), mode=_lib.AMT_PRE_NONE)
        else:
            grid, info = deriveGrid(m, pxPerDeg, arcsecPerPx)
        cells, C = grid.nx * grid.ny, eng.channels
        offMask, offSide, total = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(lib.amt_seq_output_layout(grid.nx, grid.ny, C, eng.amtDtype, ctypes.byref(offMask),
                                             ctypes.byref(offSide), ctypes.byref(total)))
        accBytes = (2 + C) * cells * 8
        # one device allocation per frame: accumulators | image | mask | elevation (caller's stream)
        flat = torch.empty(accBytes + total.value, dtype=torch.uint8, device=ctx.torch_device)
        flat.record_stream(eng.aux)                      # zeroed there, normalised / downloaded on `dout`
        flat.record_stream(eng.dout)
        out = flat[accBytes:]
        item = 1 if eng.amtDtype == _lib.AMT_U8 else 2
        dImg = out[:cells * C * item].view(eng.tdtype).view(grid.ny, grid.nx, C)
        dMask = out[offMask.value:offMask.value + cells].view(grid.ny, grid.nx)
        dElev = out[offSide.value:offSide.value + cells * 8].view(torch.float64).view(grid.ny, grid.nx)
        info['count'] = flat[:cells * 8].view(torch.int64)
        job = _lib.AmtSeqJob()
        job.grid = ctypes.pointer(grid)
        if hostImg is not None:
            job.h_img = hostImg.ctypes.data
            if sparseUpload:
                job.row0, job.row1, job.col0, job.col1 = st.row_min_c, st.row_max_c, st.col_min_c, st.col_max_c
            else:
                job.row0, job.row1, job.col0, job.col1 = 0, eng.h - 1, 0, eng.w - 1
            dimg = eng.imgRing[slot]
        else:
            dimg = m._imgDevice if m._imgDevice.dim() == 3 else m._imgDevice[..., None]
            job.d_img = dimg.data_ptr()
        job.d_acc = flat.data_ptr()
        job.d_out = out.data_ptr()
        job.out_bytes = total.value
        hostFlat = bucket = None
        if toHost:
            hostFlat, bucket = _pinnedBuffer(pool, total.value, 2 * depth + 6)
            job.h_out = hostFlat.data_ptr()
        _lib.check(lib.amt_seq_stage_b(eng.handle, slot, ctypes.byref(job)))
        if planes:
            m._planes.update({n: buffers[n] for n in eng.planeNames()})
            m._planeFree = False
        if hostImg is not None:
            m._imgDevice = dimg
        if ringBuffers:
            m._ringSlot = (eng, slot)
            eng.owner[slot] = weakref.ref(m)
        f = ResampledFrame(m, grid, info, dImg, dMask, dElev)
        f._src = hostImg                                 # the host image must outlive the asynchronous upload
        f._slot = slot
        handle, lib_ = eng.handle, lib
        waiter = lambda: _lib.check(lib_.amt_seq_wait_result(handle, slot))      # noqa: E731
        if toHost:
            f._attachHost(hostFlat, [0, offMask.value, offSide.value],
                          [cells * C * item, cells, cells * 8], bucket, waiter)
        else:
            f._event = _Waiter(waiter)       # normalised on the output stream: complete before the frame is handed out
        return f

    tracing = bool(os.environ.get('AMT_SEQ_TRACE', '').strip('0'))

    def handOut(f):
        f._finish()
        if tracing:
            # device timeline of the frame (amt_seq_trace): ms since the engine was created
            t = (ctypes.c_double * 7)()
            _lib.check(lib.amt_seq_trace(eng.handle, f._slot, t))
            stats.setdefault('trace', []).append(dict(zip(
                ('a0', 'a1', 'up0', 'up1', 'k0', 'k1', 'out'), [float(v) for v in t])))
        return f

    try:
        launched = 0
        for i, (img, hdr) in enumerate(zip(imagesOrArrays, wcsHeaders)):
            stageA.append(runA(i, img, hdr))
            # the look-ahead builds up from one frame to `depth`: the first frame's upload and long kernel are
            # enqueued right after its own stage A instead of after `depth` host iterations (0.1 ms each)
            if len(stageA) >= min(depth, launched + 1):
                stageB.append(runB(*stageA.popleft()))
                launched += 1
            while len(stageB) > AHEAD_B:
                yield handOut(stageB.popleft())
        while stageA:
            stageB.append(runB(*stageA.popleft()))
        while stageB:
            yield handOut(stageB.popleft())
    finally:
        if stageA or stageB:
            # abandoned mid-sequence (error, consumer stopped): frames with work in flight are being
            # dropped -- their buffers must not return to any allocator before that work is done
            torch.cuda.synchronize(ctx.torch_device)
        if eng is not None:
            now = ctypes.c_uint64(0)
            lib.amt_seq_h2d_bytes(eng.handle, ctypes.byref(now))
            stats['h2d_bytes'] += now.value - h2d0.value
            stats['pinned_grown'] = pool.get('grown', 0)
