"""ctypes binding of oracle/_ref/libjpgoracle.so (the plain-C restatement of the reference). TEST-ONLY."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "libjpgoracle.so")
REF_HARNESS = os.path.join(ORACLE_DIR, "_ref", "refharness")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "jpeg")


class ScanStruct(ctypes.Structure):
    _fields_ = [("ns", ctypes.c_int), ("comp", ctypes.c_int * 4), ("td", ctypes.c_int * 4), ("ta", ctypes.c_int * 4),
                ("restart_interval", ctypes.c_int), ("ecs_offset", ctypes.c_size_t), ("ecs_end", ctypes.c_size_t),
                ("mcu_cols", ctypes.c_int), ("mcu_rows", ctypes.c_int), ("ss", ctypes.c_int), ("se", ctypes.c_int),
                ("ah", ctypes.c_int), ("al", ctypes.c_int)]


class InfoStruct(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("ncomp", ctypes.c_int), ("precision", ctypes.c_int),
                ("frame_type", ctypes.c_int), ("cid", ctypes.c_int * 4), ("hs", ctypes.c_int * 4), ("vs", ctypes.c_int * 4),
                ("tq", ctypes.c_int * 4), ("hmax", ctypes.c_int), ("vmax", ctypes.c_int), ("subx", ctypes.c_int * 4),
                ("suby", ctypes.c_int * 4), ("mcu_cols", ctypes.c_int), ("mcu_rows", ctypes.c_int), ("bw", ctypes.c_int * 4),
                ("bh", ctypes.c_int * 4), ("sbw", ctypes.c_int * 4), ("sbh", ctypes.c_int * 4), ("ycbcr", ctypes.c_int),
                ("nscans", ctypes.c_int), ("scan", ScanStruct * 16), ("quant", (ctypes.c_uint16 * 64) * 4),
                ("quant_defined", ctypes.c_int * 4)]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.jpgo_read_info.restype = ctypes.c_int
        lib.jpgo_read_info.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(InfoStruct)]
        lib.jpgo_decode.restype = ctypes.c_int
        lib.jpgo_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(InfoStruct)]
        lib.jpgo_decode16.restype = ctypes.c_int
        lib.jpgo_decode16.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(InfoStruct)]
        lib.jpgo_decode_coefficients.restype = ctypes.c_int
        lib.jpgo_decode_coefficients.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(InfoStruct), ctypes.POINTER(ctypes.c_void_p)]
        lib.jpgo_idct_block.restype = None
        lib.jpgo_idct_block.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]

    def info(self, data):
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        s = InfoStruct()
        rc = self.lib.jpgo_read_info(data.ctypes.data, data.size, ctypes.byref(s))
        return rc, s

    def decode(self, data):
        """-> (rc, pixels [H,W,C] uint8 or None)"""
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        rc, s = self.info(data)
        if rc != 0:
            return rc, None
        out = np.zeros((s.height, s.width, s.ncomp), dtype=np.uint8)
        rc = self.lib.jpgo_decode(data.ctypes.data, data.size, out.ctypes.data, out.size, ctypes.byref(s))
        return rc, (out if rc == 0 else None)

    def decode16(self, data):
        """-> (rc, pixels [H,W,C] uint16 or None): 8- and 12-bit frames, what a CTYP_UWORD client bitmap receives"""
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        rc, s = self.info(data)
        if rc != 0:
            return rc, None
        out = np.zeros((s.height, s.width, s.ncomp), dtype=np.uint16)
        rc = self.lib.jpgo_decode16(data.ctypes.data, data.size, out.ctypes.data, out.size, ctypes.byref(s))
        return rc, (out if rc == 0 else None)

    def coefficients(self, data):
        """-> (rc, info, [per component int32 [bh,bw,8,8] QUANTIZED raster-order coefficients])"""
        data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        rc, s = self.info(data)
        if rc != 0:
            return rc, s, None
        planes = [np.zeros((s.bh[c], s.bw[c], 8, 8), dtype=np.int32) for c in range(s.ncomp)]
        ptrs = (ctypes.c_void_p * 4)(*[p.ctypes.data for p in planes] + [None] * (4 - s.ncomp))
        rc = self.lib.jpgo_decode_coefficients(data.ctypes.data, data.size, ctypes.byref(s), ptrs)
        return rc, s, planes

    def idct(self, block_raster_int32, delta_raster_u16, dcoffset=128):
        src = np.ascontiguousarray(block_raster_int32, dtype=np.int32)
        q = np.ascontiguousarray(delta_raster_u16, dtype=np.uint16)
        out = np.zeros(64, dtype=np.int32)
        self.lib.jpgo_idct_block(out.ctypes.data, src.ctypes.data, q.ctypes.data, dcoffset)
        return out.reshape(8, 8)


def build():
    """Compiles the C restatement (always) and, where /root/reference exists, the reference itself."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def load():
    if not os.path.exists(LIB):
        build()
    return Oracle(ctypes.CDLL(LIB))


def have_reference():
    return os.path.exists(REF_HARNESS) and os.path.exists(REF_CLI)


def reference_decode(jpg_path, tmp_raw):
    r = subprocess.run([REF_HARNESS, "decode", jpg_path, tmp_raw], capture_output=True, text=True)
    if r.returncode != 0:
        return None
    fields = r.stdout.split()
    w, h, d = map(int, fields[:3])
    wide = len(fields) > 3  # "16": native-endian 16-bit samples (precision above 8)
    return np.fromfile(tmp_raw, dtype=np.uint16 if wide else np.uint8).reshape(h, w, d)


def with_dc_quantiser(data, value):
    """The same codestream with the DC entry of every 8-bit quantisation table replaced: the coefficients stay within
    int16, but the IDCT samples leave the ranges the GPU fast paths are built for (int16 planes, 32-bit colour)."""
    b = bytearray(data.tobytes() if hasattr(data, "tobytes") else data)
    i = 2
    while i + 4 <= len(b) and b[i] == 0xFF and b[i + 1] != 0xDA:
        seglen = (b[i + 2] << 8) | b[i + 3]
        if b[i + 1] == 0xDB:
            j = i + 4
            while j < i + 2 + seglen:
                assert b[j] >> 4 == 0, "8-bit tables expected"
                b[j + 1] = value
                j += 65
        i += 2 + seglen
    return bytes(b)


def with_fill_bytes(data, every=2, count=1):
    """The same codestream with `count` fill bytes (0xFF) in front of every `every`-th restart marker / EOI of the
    entropy coded segment (legal: entropyparser.cpp:121-125, tables.cpp:1371-1373)."""
    b = bytes(data.tobytes() if hasattr(data, "tobytes") else data)
    sos = b.find(b"\xff\xda")
    start = sos + 2 + ((b[sos + 2] << 8) | b[sos + 3])
    out, n, i = bytearray(b[:start]), 0, start
    while i < len(b):
        if b[i] == 0xFF and i + 1 < len(b) and (0xD0 <= b[i + 1] <= 0xD7 or b[i + 1] == 0xD9):
            if n % every == 0:
                out += b"\xff" * count
            n += 1
            out += b[i:i + 2]
            i += 2
        else:
            out.append(b[i])
            i += 1
    return bytes(out)


def with_swapped_restart_ids(data, first=1):
    """The same codestream with the ids of restart markers number `first` and `first + 1` exchanged (out of sequence)."""
    b = bytearray(data.tobytes() if hasattr(data, "tobytes") else data)
    sos = bytes(b).find(b"\xff\xda")
    i, at = sos + 2 + ((b[sos + 2] << 8) | b[sos + 3]), []
    while i + 1 < len(b):
        if b[i] == 0xFF and 0xD0 <= b[i + 1] <= 0xD7:
            at.append(i + 1)
            i += 2
        else:
            i += 1
    b[at[first]], b[at[first + 1]] = b[at[first + 1]], b[at[first]]
    return bytes(b)
