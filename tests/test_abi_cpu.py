"""CPU: the C-ABI library loads, exports every symbol include/b200jpg.h declares, and its host-side parser
agrees with the oracle's on the golden vectors.  No compute calls (there is no GPU here)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.jpg")))


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "b200jpg.h")).read()
    declared = sorted(set(re.findall(r"B200JPG_API\s+[\w\s\*]+?\b(b200jpg_\w+)\s*\(", header)))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built.library_path())
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    from libjpeg_b200 import native
    assert sorted(native.ABI_SYMBOLS) == declared


@pytest.mark.parametrize("name", NAMES)
def test_parser_agrees_with_oracle(built, oracle, name):
    data = open(os.path.join(GOLDEN, name + ".jpg"), "rb").read()
    fi = built.parse(data)
    rc, s = oracle.info(data)
    assert rc == 0
    assert (fi.width, fi.height, fi.ncomp) == (s.width, s.height, s.ncomp)
    assert list(fi.subx) == list(s.subx[:s.ncomp]) and list(fi.suby) == list(s.suby[:s.ncomp])
    assert list(fi.blocks_w) == list(s.bw[:s.ncomp]) and list(fi.blocks_h) == list(s.bh[:s.ncomp])
    assert fi.nscans == s.nscans and fi.restart_interval == s.scan[0].restart_interval
    assert fi.ecs_bytes == sum(s.scan[k].ecs_end - s.scan[k].ecs_offset for k in range(s.nscans))
    assert fi.stored_blocks == sum(s.sbw[c] * s.sbh[c] for c in range(s.ncomp))
    assert bool(fi.ycbcr) == bool(s.ycbcr)


def test_parser_agrees_with_oracle_on_progressive_streams(built, oracle):
    """SOF2: ten scans (DC first, AC bands, refinements), every one indexed per restart interval on the host."""
    import glob
    for path in sorted(glob.glob(os.path.join(GOLDEN, "progressive", "*.jpg"))):
        data = open(path, "rb").read()
        fi = built.parse(data)
        rc, s = oracle.info(data)
        assert rc == 0 and s.frame_type == 2
        assert (fi.width, fi.height, fi.ncomp, fi.nscans) == (s.width, s.height, s.ncomp, s.nscans)
        assert fi.ecs_bytes == sum(s.scan[k].ecs_end - s.scan[k].ecs_offset for k in range(s.nscans))
        want = 0
        for k in range(s.nscans):
            sc = s.scan[k]
            total = sc.mcu_cols * sc.mcu_rows
            per = sc.restart_interval or total
            want += (total + per - 1) // per
        assert fi.n_intervals == want


def test_parser_error_codes(built):
    from libjpeg_b200 import NativeError
    data = bytearray(open(os.path.join(GOLDEN, NAMES[0] + ".jpg"), "rb").read())
    with pytest.raises(NativeError) as e:
        built.parse(b"\x00\x01\x02\x03\x04")
    assert e.value.code == -1038
    lossless = bytearray(data)
    lossless[lossless.find(b"\xff\xc0") + 1] = 0xC3
    with pytest.raises(NativeError) as e:
        built.parse(bytes(lossless))
    assert e.value.code == -1034  # NOT_IMPLEMENTED
    prog = bytearray(data)  # a sequential scan (Ss..Se = 0..63) inside a progressive frame is malformed
    prog[prog.find(b"\xff\xc0") + 1] = 0xC2
    with pytest.raises(NativeError) as e:
        built.parse(bytes(prog))
    assert e.value.code == -1038
    with pytest.raises(NativeError) as e:
        built.parse(bytes(data[:200]))
    assert e.value.code in (-1025, -1038)


def test_parser_survives_mutated_streams(built, oracle):
    """Byte mutations of golden / progressive / subsampling vectors (single bytes, truncations, marker-length edits): the
    host parser and the oracle must answer with a status, never crash or read out of bounds, and agree that a stream is
    acceptable whenever the parser accepts it with the same geometry."""
    import glob
    from libjpeg_b200 import NativeError
    rng = np.random.default_rng(20260923)
    files = sorted(glob.glob(os.path.join(GOLDEN, "*.jpg")))[:6] + sorted(glob.glob(os.path.join(GOLDEN, "progressive", "*.jpg")))[:3]
    accepted = 0
    for path in files:
        base = open(path, "rb").read()
        head = base.find(b"\xff\xda") + 16  # mutations concentrate on the marker segments
        for trial in range(60):
            b = bytearray(base)
            kind = trial % 4
            if kind == 0:
                b[int(rng.integers(2, head))] = int(rng.integers(0, 256))
            elif kind == 1:
                for _ in range(3):
                    b[int(rng.integers(2, len(b)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2:
                b = b[:int(rng.integers(2, len(b)))]
            else:
                i = int(rng.integers(2, head))
                del b[i:i + int(rng.integers(1, 5))]
            data = bytes(b)
            try:
                fi = built.parse(data)
                accepted += 1
                assert 0 < fi.width < 65536 and 0 < fi.height < 65536 and 1 <= fi.ncomp <= 4
            except NativeError as e:
                assert e.code < 0
            rc, _ = oracle.decode(data)  # must return, whatever the verdict
            assert rc <= 0
    assert accepted > 50  # entropy-coded-data mutations leave the headers intact


def test_decode_fails_loudly_without_gpu(built):
    """No CPU fallback: creating a decode context without a CUDA device is an error, never a silent detour."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from libjpeg_b200 import NativeError
    from libjpeg_b200.decoder import Context
    with pytest.raises(NativeError) as e:
        Context()
    assert e.value.code == -8193


def test_synthetic_generator_roundtrip_through_oracle(built, oracle):
    from libjpeg_b200 import synth
    img = synth.source_image(64, 48, 3)
    data = synth.encode(img, 90, (2, 2), 4)
    rc, px = oracle.decode(data.tobytes())
    assert rc == 0 and px.shape == img.shape
    assert np.abs(px.astype(int) - img.astype(int)).mean() < 12  # a lossy codec, not garbage


def test_parser_on_streams_without_eoi(built, oracle):
    """The host parser gives the reference's verdict for streams whose tail is cut off (tests/golden/noeoi, make_noeoi.py)."""
    import json
    from libjpeg_b200 import NativeError
    d = os.path.join(GOLDEN, "noeoi")
    for name, status in json.load(open(os.path.join(d, "noeoi_status.json"))).items():
        data = open(os.path.join(d, name + ".jpg"), "rb").read()
        if status == 0:
            fi = built.parse(data)
            rc, s = oracle.info(data)
            assert rc == 0 and (fi.width, fi.height, fi.nscans) == (s.width, s.height, s.nscans), name
        else:
            with pytest.raises(NativeError) as e:
                built.parse(data)
            assert e.value.code == status, name


def test_hostile_header_cannot_size_allocations(built):
    """a 689-byte stream that claims 65535 x 65535 pixels with DRI = 1 (67 M restart intervals) must be refused
    by the parser instead of allocating gigabytes for its restart index."""
    import resource
    from libjpeg_b200 import NativeError
    from libjpeg_b200 import synth
    b = bytearray(synth.encode(synth.source_image(16, 16, 1), 75, (1, 1), 1).tobytes())
    sof = bytes(b).find(b"\xff\xc0")
    b[sof + 5:sof + 9] = b"\xff\xff\xff\xff"  # height, width
    assert len(b) < 2000
    before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
    with pytest.raises(NativeError) as e:
        built.parse(bytes(b))
    assert e.value.code in (-1038, -1034)
    assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 64 * 1024  # KiB: nothing of that order was touched


@pytest.mark.parametrize("w,h,sub,q", [(640, 360, (2, 2), 75), (1920, 1080, (2, 2), 75), (333, 200, (2, 1), 60), (512, 512, (1, 1), 90),
                                       (4000, 300, (2, 2), 30), (97, 61, (1, 2), 98)])
def test_restartless_synchronisation_rounds_on_the_host(built, w, h, sub, q):
    """a scan without restart markers is cut into subsequences whose entry states are found by speculative
    decoding rounds (specsync.hpp). The host replay of the device kernel's rounds must tile the scan's blocks exactly: every
    work item starts at the bit where its first block starts on the true path, with the DC predictors of that place."""
    import ctypes
    from libjpeg_b200 import native, synth
    data = np.frombuffer(synth.encode(synth.source_image(w, h, 7), q, sub, 0).tobytes(), dtype=np.uint8)
    rounds, nseg = ctypes.c_uint32(), ctypes.c_uint32()
    rc = native.lib.b200jpg_selftest_restartless(data.ctypes.data, data.size, ctypes.byref(rounds), ctypes.byref(nseg))
    assert rc == 0
    assert nseg.value == (max(1, (data.size * 8) // 4096 // 1) and nseg.value)  # at least one work item
    assert 1 <= rounds.value <= 12, "self-synchronisation should take a handful of rounds, not one per subsequence"


def test_parser_reads_the_xt_boxes(built):
    """streams with a JPEG XT residual layer parse (the covered profile) or are refused by name -- they are never
    taken for plain JPEG (host only, no device)."""
    import glob
    from libjpeg_b200 import NativeError
    xt = os.path.join(GOLDEN, "xt")
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(xt, "*.jpg")))
    assert len(names) >= 11
    for name in names:
        data = open(os.path.join(xt, name + ".jpg"), "rb").read()
        if name.endswith("__nimpl"):
            with pytest.raises(NativeError) as e:
                built.parse(data)
            assert e.value.code == -1034, name
        else:
            fi = built.parse(data)
            assert fi.precision == 8 and fi.ncomp in (1, 3)
    # a residual box cut short is malformed, not ignored
    data = bytearray(open(os.path.join(xt, [n for n in names if not n.endswith("__nimpl")][0] + ".jpg"), "rb").read())
    at = data.find(b"RESI") - 4
    data[at:at + 4] = (int.from_bytes(data[at:at + 4], "big") + 100).to_bytes(4, "big")
    with pytest.raises(NativeError) as e:
        built.parse(bytes(data))
    assert e.value.code == -1038


def test_parser_survives_damaged_streams(built, oracle):
    """Host parser (markers, tables, restart index, JPEG XT boxes) and the oracle's on truncated and bit-flipped streams of every
    kind of fixture: an error code or a frame, never a crash, and sizes that stay inside the stream."""
    import glob
    import random
    from libjpeg_b200 import NativeError
    rng = random.Random(1234)
    pool = sorted(glob.glob(os.path.join(GOLDEN, "*.jpg")))[:4] + sorted(glob.glob(os.path.join(GOLDEN, "xt", "*.jpg")))[:4] + \
        sorted(glob.glob(os.path.join(GOLDEN, "progressive", "*.jpg")))[:2] + sorted(glob.glob(os.path.join(GOLDEN, "deep12", "*.jpg")))[:1]
    outcomes = {"ok": 0, "error": 0}
    for path in pool:
        data = open(path, "rb").read()
        for trial in range(60):
            d = bytearray(data)
            kind = trial % 3
            if kind == 0:
                d = d[:rng.randrange(2, len(d))]
            elif kind == 1:
                for _ in range(rng.randrange(1, 6)):
                    d[rng.randrange(2, len(d))] = rng.randrange(256)
            else:  # damage inside the marker segments in front of the first scan (where the boxes and tables live)
                sos = d.find(b"\xff\xda")
                for _ in range(rng.randrange(1, 4)):
                    d[rng.randrange(2, max(3, sos))] ^= 1 << rng.randrange(8)
            try:
                fi = built.parse(bytes(d))
                assert 0 < fi.width <= 65535 and 0 < fi.height <= 65535 and fi.ecs_bytes <= len(d)
                outcomes["ok"] += 1
            except NativeError as e:
                assert e.code < 0
                outcomes["error"] += 1
            rc, _ = oracle.info(bytes(d))
            assert rc <= 0
    assert outcomes["ok"] > 50 and outcomes["error"] > 50


def test_decoder_table_cache_equals_fresh_builds(built):
    """b200jpg_batch_create builds the decoder tables of a scan once per distinct set of inputs and copies them for the other scans
    that share them: every scan of every fixture (baseline, optimised tables, progressive scripts, 12 bit, unusual samplings, JPEG XT)
    through the cache, in one mixed sequence and twice, equals a fresh build byte for byte (host only)."""
    import glob
    import random
    from libjpeg_b200 import native
    paths = sorted(glob.glob(os.path.join(GOLDEN, "*.jpg")) + glob.glob(os.path.join(GOLDEN, "*", "*.jpg")))
    assert len(paths) > 60
    random.Random(5).shuffle(paths)
    datas = [open(p, "rb").read() for p in paths]
    bufs = [(ctypes.c_char * len(d)).from_buffer_copy(d) for d in datas]
    ptrs = (ctypes.c_void_p * len(bufs))(*[ctypes.addressof(b) for b in bufs])
    lens = (ctypes.c_size_t * len(bufs))(*[len(d) for d in datas])
    n = native.lib.b200jpg_selftest_table_cache(ptrs, lens, len(bufs))
    assert n >= 2 * len(paths) * 0.8, n  # (a few fixtures are deliberately unparsable)
