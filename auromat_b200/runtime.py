"""Per-GPU runtime: one C-ABI context per device, torch tensors as device/pinned memory.

PyTorch is plumbing only (allocator, streams, torch.distributed); every array pass of the
hot path is one of the CUDA kernels behind `include/auromat_b200.h`.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib

_contexts = {}
_lock = threading.Lock()


def _torch():
    import torch
    return torch


class Context:
    """Owns an `amt_ctx*` for one CUDA device and wraps the C entry points with torch
    tensors (device pointers) as arguments."""

    def __init__(self, device: int):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("auromat_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                               "there is no CPU fallback")
        self.lib = _lib.load()
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        handle = C.c_void_p()
        _lib.check(self.lib.amt_ctx_create(self.device, C.byref(handle)))
        self.handle = handle
        self._stats_dev = None

    # ------------------------------------------------------------------ plumbing
    def stream(self):
        """The CUDA stream all kernels of this context are enqueued on: torch's current stream
        of the device (looked up per call unless pinned with `pin_stream`)."""
        s = self.__dict__.get('_pinned_stream')
        if s is not None:
            return s
        return C.c_void_p(_torch().cuda.current_stream(self.torch_device).cuda_stream)

    def pin_stream(self, on=True):
        """Cache the current stream handle (the sequence pipeline issues ~10 calls per frame on
        the same stream; the lookup through torch costs a few microseconds each)."""
        self.__dict__['_pinned_stream'] = \
            C.c_void_p(_torch().cuda.current_stream(self.torch_device).cuda_stream) if on else None

    def use_stream(self, handle):
        """Pin an explicit stream handle (ctypes c_void_p) or None to unpin: no torch lookup."""
        self.__dict__['_pinned_stream'] = handle

    def copy_h2d(self, d_ptr, h_ptr, nbytes, stream_handle):
        _lib.check(self.lib.amt_copy_h2d(self.handle, C.c_void_p(d_ptr), C.c_void_p(h_ptr), nbytes, stream_handle))

    def copy_h2d_2d(self, d_ptr, h_ptr, pitch, width_bytes, rows, stream_handle):
        _lib.check(self.lib.amt_copy_h2d_2d(self.handle, C.c_void_p(d_ptr), pitch, C.c_void_p(h_ptr), pitch,
                                            width_bytes, rows, stream_handle))

    def copy_d2h(self, h_ptr, d_ptr, nbytes, stream_handle):
        _lib.check(self.lib.amt_copy_d2h(self.handle, C.c_void_p(h_ptr), C.c_void_p(d_ptr), nbytes, stream_handle))

    def synchronize(self):
        _torch().cuda.current_stream(self.torch_device).synchronize()

    def empty(self, shape, dtype):
        torch = _torch()
        return torch.empty(shape, dtype=dtype, device=self.torch_device)

    def zeros(self, shape, dtype):
        torch = _torch()
        return torch.zeros(shape, dtype=dtype, device=self.torch_device)

    def to_device(self, array, non_blocking=True):
        """numpy array (or pinned torch CPU tensor) -> device tensor on the current stream."""
        torch = _torch()
        if isinstance(array, np.ndarray):
            if array.dtype == np.uint16:
                t = torch.from_numpy(np.ascontiguousarray(array).view(np.int16))
                return t.to(self.torch_device, non_blocking=non_blocking).view(torch.uint16)
            array = torch.from_numpy(np.ascontiguousarray(array))
        return array.to(self.torch_device, non_blocking=non_blocking)

    @staticmethod
    def to_numpy(tensor):
        torch = _torch()
        t = tensor.detach()
        if t.dtype == torch.uint16:
            return t.view(torch.int16).cpu().numpy().view(np.uint16)
        return t.cpu().numpy()

    @staticmethod
    def ptr(tensor):
        return C.c_void_p(tensor.data_ptr()) if tensor is not None else C.c_void_p(None)

    @property
    def launch_count(self) -> int:
        n = C.c_uint64()
        _lib.check(self.lib.amt_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def measure_fp64_peak(self) -> float:
        """Warp-lane DFMA per second of this device (register-resident DFMA stream)."""
        v = C.c_double()
        _lib.check(self.lib.amt_measure_fp64_peak(self.handle, C.byref(v)))
        return v.value

    def measure_atomic_peak(self, cells: int) -> float:
        """u64 atomicAdd per second to pseudo-random words of a `cells`-word grid (L2 scatter ceiling)."""
        v = C.c_double()
        _lib.check(self.lib.amt_measure_atomic_peak(self.handle, int(cells), C.byref(v)))
        return v.value

    # ------------------------------------------------------------------ kernels
    @staticmethod
    def out_struct(planes: dict):
        """amt_georef_out of a planes dict (name -> device tensor)."""
        out = _lib.AmtGeorefOut()
        for name, t in planes.items():
            setattr(out, "d_" + name, t.data_ptr())
        return out

    def georef(self, frame: _lib.AmtFrame, planes: dict, stats=None, out=None):
        """planes: name -> device tensor for any subset of amt_georef_out members; `out`: the
        prebuilt amt_georef_out of exactly these planes (plane rings of the sequence pipeline)."""
        if out is None:
            out = self.out_struct(planes)
        _lib.check(self.lib.amt_georef(self.handle, C.byref(frame), C.byref(out), self.ptr(stats), self.stream()))

    def new_stats(self):
        torch = _torch()
        return torch.zeros(C.sizeof(_lib.AmtStats), dtype=torch.uint8, device=self.torch_device)

    def sanitize(self, width, height, planes: dict, out=None):
        if out is None:
            out = self.out_struct(planes)
        _lib.check(self.lib.amt_sanitize(self.handle, width, height, C.byref(out), self.stream()))

    @staticmethod
    def bitmap_words(width, height):
        """(#words of the corner bitmap, #words of the centre bitmap), rows padded to 32 bits."""
        return (height + 1) * ((width + 1 + 31) // 32), height * ((width + 31) // 32)

    def new_bitmaps(self, width, height):
        torch = _torch()
        nk, nc = self.bitmap_words(width, height)
        # one allocation, corner words first: amt_sanitize then snapshots both with a single copy
        buf = torch.empty(nk + nc, dtype=torch.int32, device=self.torch_device)
        return buf[:nk], buf[nk:]

    def valid_bits(self, width, height, planes: dict):
        """Build the validity bitmaps of NaN-marked planes into planes['valid_k'/'valid_c']."""
        planes['valid_k'], planes['valid_c'] = self.new_bitmaps(width, height)
        _lib.check(self.lib.amt_valid_bits(self.handle, width, height, self.ptr(planes['lat_k']),
                                           self.ptr(planes['lat_c']), self.ptr(planes['valid_k']),
                                           self.ptr(planes['valid_c']), self.stream()))

    def bbox_stats(self, width, height, planes: dict, stats, pole_test=False, pre: "_lib.AmtGrid | None" = None):
        _lib.check(self.lib.amt_bbox_stats(self.handle, width, height, self.ptr(planes['lat_k']),
                                           self.ptr(planes['lon_k']), self.ptr(planes['valid_k']),
                                           self.ptr(planes['valid_c']), 1 if pole_test else 0,
                                           C.byref(pre) if pre is not None else None,
                                           self.ptr(stats), self.stream()))

    def bbox_stats_frame(self, frame: _lib.AmtFrame, valid_k, valid_c, stats, pre: "_lib.AmtGrid | None" = None):
        _lib.check(self.lib.amt_bbox_stats_frame(self.handle, C.byref(frame), self.ptr(valid_k), self.ptr(valid_c),
                                                 C.byref(pre) if pre is not None else None, self.ptr(stats),
                                                 self.stream()))

    def georef_bin_fused(self, frame: _lib.AmtFrame, valid_c, img, grid: _lib.AmtGrid, count, sums, fsum):
        torch = _torch()
        dtype = {torch.uint8: _lib.AMT_U8, torch.uint16: _lib.AMT_U16}.get(img.dtype)
        if dtype is None:
            raise NotImplementedError("image dtype must be uint8 or uint16, got %s" % img.dtype)
        channels = img.numel() // (frame.width * frame.height)
        _lib.check(self.lib.amt_georef_bin_fused(self.handle, C.byref(frame), self.ptr(valid_c), self.ptr(img), dtype,
                                                 channels, C.byref(grid), self.ptr(count), self.ptr(sums),
                                                 self.ptr(fsum), self.stream()))

    def georef_fused(self, frame: _lib.AmtFrame, valid_k, valid_c, planes=None, out=None, img=None, grid=None,
                     count=None, sums=None, fsum=None):
        """`amt_georef_fused`: the coordinate planes (when `planes` / `out` are given) and / or the
        binning of the centres into `grid` (when given) from the frame's final validity bitmaps."""
        torch = _torch()
        if out is None and planes is not None:
            out = self.out_struct({k: v for k, v in planes.items() if not k.startswith(('valid', '_'))})
        dtype, channels = _lib.AMT_U8, 1
        if grid is not None:
            dtype = {torch.uint8: _lib.AMT_U8, torch.uint16: _lib.AMT_U16}.get(img.dtype)
            if dtype is None:
                raise NotImplementedError("image dtype must be uint8 or uint16, got %s" % img.dtype)
            channels = img.numel() // (frame.width * frame.height)
        _lib.check(self.lib.amt_georef_fused(self.handle, C.byref(frame), C.byref(out) if out is not None else None,
                                             self.ptr(valid_k), self.ptr(valid_c), self.ptr(img), dtype, channels,
                                             C.byref(grid) if grid is not None else None, self.ptr(count),
                                             self.ptr(sums), self.ptr(fsum), self.stream()))

    def sip_distort(self, frame: _lib.AmtFrame, u, v):
        """(u', v') of the frame's FITS-SIP forward polynomial for device tensors u, v."""
        torch = _torch()
        uo, vo = torch.empty_like(u), torch.empty_like(v)
        _lib.check(self.lib.amt_sip_distort(self.handle, C.byref(frame), self.ptr(u), self.ptr(v), u.numel(),
                                            self.ptr(uo), self.ptr(vo), self.stream()))
        return uo, vo

    def apply_center_mask(self, width, height, planes: dict, mask=None, min_elevation=float("nan")):
        out = self.out_struct(planes)
        _lib.check(self.lib.amt_apply_center_mask(self.handle, width, height, self.ptr(mask),
                                                  float(min_elevation), C.byref(out), self.stream()))

    def rotate_coords(self, lat, lon, pre: _lib.AmtGrid):
        _lib.check(self.lib.amt_rotate_coords(self.handle, self.ptr(lat), self.ptr(lon), lat.numel(),
                                              C.byref(pre), self.stream()))

    def reproject(self, lat_ref, lon_ref, station_ecef, height_ref, height_new, wgs_a, wgs_b):
        """`amt_reproject` on device tensors; returns new (lat, lon) tensors of the same shape."""
        torch = _torch()
        lat_out, lon_out = torch.empty_like(lat_ref), torch.empty_like(lon_ref)
        ecef = (C.c_double * 3)(*[float(v) for v in station_ecef])
        _lib.check(self.lib.amt_reproject(self.handle, self.ptr(lat_ref), self.ptr(lon_ref), lat_ref.numel(), ecef,
                                          float(height_ref), float(height_new), float(wgs_a), float(wgs_b),
                                          self.ptr(lat_out), self.ptr(lon_out), self.stream()))
        return lat_out, lon_out

    def corner_means(self, width, height, lat_k, lon_k):
        torch = _torch()
        lat_c = torch.empty(width * height, dtype=torch.float64, device=self.torch_device)
        lon_c = torch.empty_like(lat_c)
        _lib.check(self.lib.amt_corner_means(self.handle, width, height, self.ptr(lat_k), self.ptr(lon_k),
                                             self.ptr(lat_c), self.ptr(lon_c), self.stream()))
        return lat_c, lon_c

    def polygon_center_mask(self, width, height, lat_k, lon_k, polygon, pre: "_lib.AmtGrid | None" = None):
        """(centre mask u8 (h*w), #corners inside) of `amt_polygon_center_mask`; `polygon` is a
        device tensor (n,2) of (lat, lon), already rotated like `pre` rotates the corners."""
        torch = _torch()
        mask = torch.empty(width * height, dtype=torch.uint8, device=self.torch_device)
        n_inside = torch.zeros(1, dtype=torch.int64, device=self.torch_device)
        _lib.check(self.lib.amt_polygon_center_mask(self.handle, width, height, self.ptr(lat_k), self.ptr(lon_k),
                                                    self.ptr(polygon), polygon.shape[0],
                                                    C.byref(pre) if pre is not None else None, self.ptr(mask),
                                                    self.ptr(n_inside), self.stream()))
        return mask, n_inside

    def plate_carree_coords(self, nx, ny, lat_hi, lat_lo, lon_lo, lon_hi):
        torch = _torch()
        lat_k = torch.empty((ny + 1, nx + 1), dtype=torch.float64, device=self.torch_device)
        lon_k = torch.empty_like(lat_k)
        lat_c = torch.empty((ny, nx), dtype=torch.float64, device=self.torch_device)
        lon_c = torch.empty_like(lat_c)
        _lib.check(self.lib.amt_plate_carree_coords(self.handle, nx, ny, float(lat_hi), float(lat_lo),
                                                    float(lon_lo), float(lon_hi), self.ptr(lat_k), self.ptr(lon_k),
                                                    self.ptr(lat_c), self.ptr(lon_c), self.stream()))
        return lat_k, lon_k, lat_c, lon_c

    def read_stats(self, stats) -> _lib.AmtStats:
        return self.finish_stats(self.start_stats_readback(stats))

    def start_stats_readback(self, stats):
        """Enqueue the device->host copy of an amt_stats block into a pooled pinned buffer;
        returns a handle for `finish_stats` (lets the host prepare the next frame meanwhile)."""
        torch = _torch()
        pool = self.__dict__.setdefault('_pinned_stats', [])
        host = pool.pop() if pool else torch.empty(C.sizeof(_lib.AmtStats), dtype=torch.uint8).pin_memory()
        # copy and event go to the stream the producing kernels were enqueued on (`stream()`: a
        # pinned pipeline stream or torch's current one), never to whatever stream happens to be
        # current when a generator is resumed
        st = self.stream()
        self.copy_d2h(host.data_ptr(), stats.data_ptr(), C.sizeof(_lib.AmtStats), st)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.ExternalStream(st.value or 0, device=self.torch_device) if st.value else
                  torch.cuda.default_stream(self.torch_device))
        return host, ev

    def finish_stats(self, handle) -> _lib.AmtStats:
        host, ev = handle
        ev.synchronize()
        s = _lib.AmtStats.from_buffer_copy(host.numpy().tobytes())
        self.__dict__.setdefault('_pinned_stats', []).append(host)
        return s

    def latlon_to_mlatmlt(self, lat, lon, altitude, wgs_a, wgs_b, m_geo_sm):
        torch = _torch()
        mlat = torch.empty_like(lat)
        mlt = torch.empty_like(lat)
        m = (C.c_double * 9)(*np.asarray(m_geo_sm, dtype=np.float64).ravel())
        _lib.check(self.lib.amt_latlon_to_mlatmlt(self.handle, self.ptr(lat), self.ptr(lon), lat.numel(),
                                                  float(altitude), float(wgs_a), float(wgs_b), m,
                                                  self.ptr(mlat), self.ptr(mlt), self.stream()))
        return mlat, mlt

    def sm_to_latlon(self, lat, lon, m_geo_sm, wgs_a, wgs_b):
        """In place: solar-magnetic (lat, lon) degrees -> geodetic degrees."""
        m = (C.c_double * 9)(*np.asarray(m_geo_sm, dtype=np.float64).ravel())
        _lib.check(self.lib.amt_sm_to_latlon(self.handle, self.ptr(lat), self.ptr(lon), lat.numel(), m,
                                             float(wgs_a), float(wgs_b), self.stream()))

    def bin_accumulate(self, lat_c, lon_c, side, img, grid: _lib.AmtGrid, count, sums, fsum, near_edge=None):
        torch = _torch()
        dtype = {torch.uint8: _lib.AMT_U8, torch.uint16: _lib.AMT_U16}.get(img.dtype)
        if dtype is None:
            raise NotImplementedError("image dtype must be uint8 or uint16, got %s" % img.dtype)
        n = lat_c.numel()
        channels = img.numel() // n
        _lib.check(self.lib.amt_bin_accumulate(self.handle, self.ptr(lat_c), self.ptr(lon_c), self.ptr(side),
                                               self.ptr(img), dtype, channels, n, C.byref(grid), self.ptr(count),
                                               self.ptr(sums), self.ptr(fsum), self.ptr(near_edge), self.stream()))

    def cell_indices(self, lat_c, lon_c, grid: _lib.AmtGrid):
        torch = _torch()
        ix = torch.empty(lat_c.numel(), dtype=torch.int32, device=self.torch_device)
        iy = torch.empty_like(ix)
        _lib.check(self.lib.amt_cell_indices(self.handle, self.ptr(lat_c), self.ptr(lon_c), lat_c.numel(),
                                             C.byref(grid), self.ptr(ix), self.ptr(iy), self.stream()))
        return ix, iy

    def normalise(self, grid: _lib.AmtGrid, img_dtype, channels, count, sums, fsum):
        torch = _torch()
        dtype = {torch.uint8: _lib.AMT_U8, torch.uint16: _lib.AMT_U16}[img_dtype]
        # one allocation for image | mask | side channel (64-byte aligned parts): the sequence
        # pipeline brings all three to the host with a single copy
        cells = grid.ny * grid.nx
        itemsize = 1 if img_dtype == torch.uint8 else 2
        o_mask = (cells * channels * itemsize + 63) // 64 * 64
        o_side = (o_mask + cells + 63) // 64 * 64
        flat = torch.empty(o_side + (cells * 8 if fsum is not None else 0), dtype=torch.uint8, device=self.torch_device)
        out_img = flat[:cells * channels * itemsize].view(img_dtype).view(grid.ny, grid.nx, channels)
        out_mask = flat[o_mask:o_mask + cells].view(grid.ny, grid.nx)
        out_side = flat[o_side:].view(torch.float64).view(grid.ny, grid.nx) if fsum is not None else None
        out_img._amt_flat = flat
        _lib.check(self.lib.amt_normalise(self.handle, C.byref(grid), dtype, channels, self.ptr(count),
                                          self.ptr(sums), self.ptr(fsum), self.ptr(out_img), self.ptr(out_mask),
                                          self.ptr(out_side), self.stream()))
        return out_img, out_mask, out_side

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.amt_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def get_context(device=None) -> Context:
    """The process-wide context of `device` (default: torch's current CUDA device)."""
    torch = _torch()
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("auromat_b200 needs a CUDA device (B200, sm_100a); none is visible and "
                               "there is no CPU fallback")
        device = torch.cuda.current_device()
    device = int(device)
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = Context(device)
        return ctx
