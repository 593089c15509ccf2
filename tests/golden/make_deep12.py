"""Regenerates tests/golden/deep12/: 12-bit (SOF1, and one SOF2) streams written and decoded by the UNMODIFIED reference,
pixels as native-endian 16-bit samples (what a CTYP_UWORD client bitmap receives). They pin the oracle's 12-bit path
(level shift 2048, clamp 4095, same LONG IDCT: tables.cpp:1877-1891) -- groundwork for SURVEY 8f4; the CUDA path is 8-bit.

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_deep12.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "deep12")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from libjpeg_b200.synth import source_image  # noqa: E402
import oracle_binding  # noqa: E402

CASES = [  # name, width, height, sampling option, restart interval, quality, extra options
    ("d12_420_96x80_z3_q80", 96, 80, "1x1,2x2,2x2", 3, 80, []),
    ("d12_444_64x64_z4_q95", 64, 64, None, 4, 95, []),
    ("d12_422_127x99_z5_q75", 127, 99, "1x1,2x1,2x1", 5, 75, []),
    ("d12_420_50x38_q60", 50, 38, "1x1,2x2,2x2", 0, 60, []),
    ("d12_p420_96x80_z3_q80", 96, 80, "1x1,2x2,2x2", 3, 80, ["-v"]),   # progressive
]


def main():
    os.makedirs(OUT, exist_ok=True)
    pixels = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, sub, z, q, extra in CASES:
            img = source_image(w, h, w * 1000 + h).astype(np.uint16) * 16 + (np.arange(w) % 16)[None, :, None].astype(np.uint16)
            src = os.path.join(tmp, "in.ppm")
            with open(src, "wb") as f:
                f.write(b"P6\n%d %d\n4095\n" % (w, h) + img.astype(">u2").tobytes())
            jpg = os.path.join(OUT, name + ".jpg")
            cmd = [os.path.join(REF, "jpeg"), "-q", str(q)] + extra
            if sub:
                cmd += ["-s", sub]
            if z:
                cmd += ["-z", str(z)]
            r = subprocess.run(cmd + [src, jpg], capture_output=True, text=True)
            if r.returncode != 0:
                sys.exit("reference encoder failed on %s: %s" % (name, r.stderr))
            px = oracle_binding.reference_decode(jpg, os.path.join(tmp, "o.raw"))
            if px is None or px.dtype != np.uint16:
                sys.exit("reference decoder failed on %s" % name)
            pixels[name] = px
            print(name, os.path.getsize(jpg), "bytes ->", px.shape, px.dtype, int(px.max()))
    np.savez_compressed(os.path.join(OUT, "deep12_pixels.npz"), **pixels)


if __name__ == "__main__":
    main()
