"""Host-side mirror of the batch decode interface (include/b200jpg.h) for Python callers.

PyTorch is plumbing here: it owns the output tensor and the CUDA stream; every byte of decode work happens
in the hand-written kernels of libb200jpg.so.
"""
import ctypes
from dataclasses import dataclass

import numpy as np

from . import native
from .native import lib


@dataclass
class FrameInfo:
    width: int
    height: int
    ncomp: int
    precision: int
    ycbcr: bool
    hs: tuple
    vs: tuple
    subx: tuple
    suby: tuple
    mcu_cols: int
    mcu_rows: int
    blocks_w: tuple
    blocks_h: tuple
    nscans: int
    restart_interval: int
    n_intervals: int
    ecs_bytes: int
    stored_blocks: int

    @staticmethod
    def from_struct(s):
        n = s.ncomp
        return FrameInfo(s.width, s.height, n, s.precision, bool(s.ycbcr), tuple(s.hs[:n]), tuple(s.vs[:n]),
                         tuple(s.subx[:n]), tuple(s.suby[:n]), s.mcu_cols, s.mcu_rows, tuple(s.blocks_w[:n]),
                         tuple(s.blocks_h[:n]), s.nscans, s.restart_interval, s.n_intervals, s.ecs_bytes,
                         s.stored_blocks)


def _as_buffer(data):
    """bytes / bytearray / numpy uint8 array -> (address, length, keepalive)."""
    if isinstance(data, np.ndarray):
        arr = np.ascontiguousarray(data, dtype=np.uint8)
        return arr.ctypes.data, arr.size, arr
    if hasattr(data, "data_ptr"):  # torch CPU tensor (possibly pinned)
        return data.data_ptr(), data.numel(), data
    buf = (ctypes.c_uint8 * len(data)).from_buffer_copy(data) if not isinstance(data, bytearray) else (
        ctypes.c_uint8 * len(data)).from_buffer(data)
    return ctypes.addressof(buf), len(data), buf


def parse(data):
    """Marker-level parse of one codestream on the host (no GPU). Raises NativeError on malformed input."""
    addr, n, keep = _as_buffer(data)
    s = native.FrameInfoStruct()
    rc = lib.b200jpg_parse(addr, n, ctypes.byref(s))
    native.check(rc)
    del keep
    return FrameInfo.from_struct(s)


class Context:
    def __init__(self, device=-1):
        self.handle = ctypes.c_void_p()
        rc = lib.b200jpg_create(device, ctypes.byref(self.handle))
        native.check(rc)

    def trim(self):
        """Returns the pooled (idle) device / pinned buffers to the driver."""
        if self.handle:
            lib.b200jpg_trim(self.handle)

    def close(self):
        if self.handle:
            lib.b200jpg_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        self.close()


class BatchDecoder:
    """One batch of codestreams: parse + pack on construction, then upload() / decode() on a CUDA stream."""

    def __init__(self, frames, device=-1, tolerate_bad=False, ctx=None, color_transform=True, upsample=True):
        self.ctx = ctx or Context(device)
        self._keep = []
        n = len(frames)
        ptrs = (ctypes.c_void_p * n)()
        lens = (ctypes.c_size_t * n)()
        for i, f in enumerate(frames):
            addr, ln, keep = _as_buffer(f)
            ptrs[i] = addr
            lens[i] = ln
            self._keep.append(keep)
        self.handle = ctypes.c_void_p()
        flags = 0 if color_transform else 1  # B200JPG_FLAG_NO_COLOR_TRANSFORM (JPGTAG_MATRIX_LTRAFO = none)
        if not upsample:
            flags = 3  # B200JPG_FLAG_NO_UPSAMPLE (JPGTAG_DECODER_UPSAMPLE = false): planes, no colour transformation
        self.upsample = bool(upsample)
        rc = lib.b200jpg_batch_create_ex(self.ctx.handle, ptrs, lens, n, int(tolerate_bad), flags, ctypes.byref(self.handle))
        native.check(rc, self.ctx.handle)
        self._keep = []  # the batch holds its own pinned copy
        self.n = n
        self.out_bytes = lib.b200jpg_batch_out_bytes(self.handle, -1)

    # ---- geometry ----------------------------------------------------------------------------------
    def info(self, i):
        s = native.FrameInfoStruct()
        lib.b200jpg_batch_frame_info(self.handle, i, ctypes.byref(s))
        return FrameInfo.from_struct(s)

    def out_offset(self, i):
        return lib.b200jpg_batch_out_offset(self.handle, i)

    def frame_bytes(self, i):
        return lib.b200jpg_batch_out_bytes(self.handle, i)

    @property
    def ecs_bytes(self):
        return lib.b200jpg_batch_ecs_bytes(self.handle)

    @property
    def stored_blocks(self):
        return lib.b200jpg_batch_stored_blocks(self.handle)

    @property
    def h2d_bytes(self):
        return lib.b200jpg_batch_h2d_bytes(self.handle)

    # ---- tables (multi-GPU broadcast) --------------------------------------------------------------
    def export_tables(self):
        size = lib.b200jpg_batch_export_tables(self.handle, None, 0)
        if size == 0:
            raise native.NativeError(-1024, "batch does not have exactly one table set")
        buf = np.zeros(size, dtype=np.uint8)
        lib.b200jpg_batch_export_tables(self.handle, buf.ctypes.data, size)
        return buf

    def import_tables(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        native.check(lib.b200jpg_batch_import_tables(self.handle, blob.ctypes.data, blob.size), self.ctx.handle)

    # ---- device work ---------------------------------------------------------------------------------
    @staticmethod
    def _stream_ptr(stream):
        if stream is None:
            import torch
            return torch.cuda.current_stream().cuda_stream
        if hasattr(stream, "cuda_stream"):
            return stream.cuda_stream
        return int(stream)

    def upload(self, stream=None):
        native.check(lib.b200jpg_batch_upload(self.handle, self._stream_ptr(stream)), self.ctx.handle)

    def reindex(self, stream=None):
        """Re-runs the device-side restart index alone (upload() already built it); for timing."""
        native.check(lib.b200jpg_batch_reindex(self.handle, self._stream_ptr(stream)), self.ctx.handle)

    def new_output(self, device=None):
        import torch
        return torch.empty(self.out_bytes, dtype=torch.uint8, device=device or "cuda")

    def decode(self, out, stream=None):
        """Launches both stages on `stream`; `out` is a CUDA uint8 tensor of at least out_bytes elements."""
        assert out.is_cuda and out.numel() >= self.out_bytes
        native.check(lib.b200jpg_batch_decode(self.handle, out.data_ptr(), self._stream_ptr(stream)), self.ctx.handle)
        return out

    def decode_entropy(self, stream=None):
        native.check(lib.b200jpg_batch_decode_entropy(self.handle, self._stream_ptr(stream)), self.ctx.handle)

    def reconstruct(self, out, stream=None):
        native.check(lib.b200jpg_batch_reconstruct(self.handle, out.data_ptr(), self._stream_ptr(stream)), self.ctx.handle)
        return out

    def status(self, i):
        return lib.b200jpg_batch_frame_status(self.handle, i)

    def coefficients(self, i, c):
        fi = self.info(i)
        n = fi.blocks_w[c] * fi.blocks_h[c] * 64
        buf = np.empty(n, dtype=np.int16)
        native.check(lib.b200jpg_batch_read_coefficients(self.handle, i, c, buf.ctypes.data, n), self.ctx.handle)
        return buf.reshape(fi.blocks_h[c], fi.blocks_w[c], 8, 8)

    @property
    def launches(self):
        return lib.b200jpg_batch_last_launch_count(self.handle)

    def enable_timing(self, on=True):
        lib.b200jpg_batch_enable_timing(self.handle, int(on))

    def last_timing(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        native.check(lib.b200jpg_batch_last_timing(self.handle, ctypes.byref(a), ctypes.byref(b)), self.ctx.handle)
        return a.value, b.value

    def last_unstuff_ms(self):
        return lib.b200jpg_batch_last_unstuff_ms(self.handle)

    def frame_view(self, out, i):
        """Frame i of a decoded output tensor as [H, W, C]."""
        fi = self.info(i)
        off = self.out_offset(i)
        n = fi.width * fi.height * fi.ncomp
        if fi.precision > 8:  # 12-bit frames: native-endian 16-bit samples
            import torch
            return out[off:off + 2 * n].view(torch.int16).view(fi.height, fi.width, fi.ncomp)
        return out[off:off + n].view(fi.height, fi.width, fi.ncomp)

    def plane_views(self, out, i):
        """upsample=False: the components of frame i as planes [ceil(H/suby), ceil(W/subx)] at their own resolution."""
        import torch
        assert not self.upsample
        fi = self.info(i)
        off, deep, planes = self.out_offset(i), fi.precision > 8, []
        for c in range(fi.ncomp):
            h, w = (fi.height + fi.suby[c] - 1) // fi.suby[c], (fi.width + fi.subx[c] - 1) // fi.subx[c]
            n = h * w * (2 if deep else 1)
            p = out[off:off + n]
            planes.append(p.view(torch.int16).view(h, w) if deep else p.view(h, w))
            off += n
        return planes

    def close(self):
        if getattr(self, "handle", None):
            lib.b200jpg_batch_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
