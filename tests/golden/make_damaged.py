"""Regenerates tests/golden/damaged/: golden codestreams damaged on purpose, with the pixels the UNMODIFIED reference
(oracle/_ref/refharness) produces for them (its resynchronisation, entropyparser.cpp:137-199, clears the intervals the
stream no longer contains). They pin the oracle's restatement of that logic and the CUDA path's handling of streams
that end early.

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_damaged.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "damaged")


def rst_positions(b):
    sos = b.find(b"\xff\xda")
    i, at = sos + 2 + ((b[sos + 2] << 8) | b[sos + 3]), []
    while i + 1 < len(b):
        if b[i] == 0xFF and 0xD0 <= b[i + 1] <= 0xD7:
            at.append(i)
            i += 2
        else:
            i += 1
    return at


def truncated_at(b, k):
    """The stream ends (EOI) where restart marker number k stood: intervals k+1.. are missing."""
    return b[:rst_positions(b)[k]] + b"\xff\xd9"


CASES = [  # name, source golden vector, damage
    ("trunc_late_c420_127x255", "c420_127x255_z7_q30", lambda b: truncated_at(b, len(rst_positions(b)) - 3)),
    ("trunc_early_c420_127x255", "c420_127x255_z7_q30", lambda b: truncated_at(b, 5)),
    ("trunc_c422_100x60", "c422_100x60_z5_q80", lambda b: truncated_at(b, 4)),
    ("trunc_g_40x24", "g_40x24_z2_q75", lambda b: truncated_at(b, 2)),
    ("trunc_c444_64x64", "c444_64x64_z16_q90", lambda b: truncated_at(b, 1)),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    pixels = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, src, damage in CASES:
            data = damage(open(os.path.join(HERE, src + ".jpg"), "rb").read())
            jpg = os.path.join(OUT, name + ".jpg")
            open(jpg, "wb").write(data)
            raw = os.path.join(tmp, "o.raw")
            r = subprocess.run([os.path.join(REF, "refharness"), "decode", jpg, raw], capture_output=True, text=True)
            if r.returncode != 0:
                sys.exit("reference failed on %s: %s" % (name, r.stderr))
            w, h, c = (int(v) for v in r.stdout.split()[:3])
            px = np.fromfile(raw, dtype=np.uint8).reshape(h, w, c)
            pixels[name] = px[..., 0] if c == 1 else px
            print(name, len(data), "bytes ->", px.shape)
    np.savez_compressed(os.path.join(OUT, "damaged_pixels.npz"), **pixels)


if __name__ == "__main__":
    main()
