"""Regenerates tests/golden/xt/: JPEG XT (ISO/IEC 18477) streams with a residual layer -- SURVEY 8f3 / BASELINE config 5 -- written
by the reference encoder (`oracle/_ref/jpeg -r -q <base> -Q <extension> -h ...`, cmd/main.cpp) and decoded by the reference
decoder (refharness: the public API, 8-row stripes): a base image, a second DCT codestream in the RESI box of the APP11
markers, and the merging specification box SPEC that tells how the two combine (colortrafo/ycbcrtrafo.cpp:747-880).

`xt_pixels.npz`: "<name>" = the reference's pixels.  "<name>__nimpl" vectors are streams outside the profile the oracle / the
CUDA path cover (lossless residual, DCT bypass, 12-bit residual, refinement scans): NOT_IMPLEMENTED is the contract there.

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_xt.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "xt")

from libjpeg_b200.synth import source_image  # noqa: E402

# name, width, height, grey, encoder arguments
CASES = [
    ("xt_444_64x48_q75_Q90", 64, 48, False, ["-r", "-q", "75", "-Q", "90", "-h"]),
    ("xt_420_96x80_z6_q75_Q90", 96, 80, False, ["-r", "-q", "75", "-Q", "90", "-h", "-s", "1x1,2x2,2x2", "-z", "6"]),
    ("xt_420_127x99_q30_Q60", 127, 99, False, ["-r", "-q", "30", "-Q", "60", "-s", "1x1,2x2,2x2"]),
    ("xt_422_100x60_q85_Q95", 100, 60, False, ["-r", "-q", "85", "-Q", "95", "-h", "-s", "1x1,2x1,2x1"]),
    ("xt_g_40x24_q75_Q90", 40, 24, True, ["-r", "-q", "75", "-Q", "90", "-h"]),
    ("xt_420r_96x80_q75_Q80", 96, 80, False, ["-r", "-q", "75", "-Q", "80", "-h", "-s", "1x1,2x2,2x2", "-sr", "1x1,2x2,2x2"]),
    ("xt_p444_64x48_q75_Q90", 64, 48, False, ["-r", "-q", "75", "-Q", "90", "-v"]),
    ("xt_444_256x192_z8_q75_Q90", 256, 192, False, ["-r", "-q", "75", "-Q", "90", "-h", "-z", "8"]),
    ("xt_c444_80x56_q75_Q90", 80, 56, False, ["-r", "-q", "75", "-Q", "90", "-c"]),  # no decorrelation: L and R transformation = identity
    ("xt_rv444_80x56_q75_Q90", 80, 56, False, ["-r", "-q", "75", "-Q", "90", "-rv"]),  # progressive residual codestream
    ("xt_420_sr422_80x56_q60_Q85", 80, 56, False, ["-r", "-q", "60", "-Q", "85", "-qt", "3", "-s", "1x1,2x2,2x2", "-sr", "1x1,2x1,2x1"]),
]
OUTSIDE = [  # accepted by the reference, outside the covered profile
    ("xt_lossless_33x17__nimpl", 33, 17, False, ["-r", "-q", "50", "-Q", "100", "-h"]),  # residual after an RCT, int-to-int DCT
    ("xt_r12_64x48__nimpl", 64, 48, False, ["-r", "-r12", "-q", "75", "-Q", "90"]),
    ("xt_rR_64x48__nimpl", 64, 48, False, ["-r", "-q", "75", "-Q", "90", "-rR", "2"]),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    px = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, grey, args in CASES + OUTSIDE:
            img = source_image(w, h, 23)
            src = os.path.join(tmp, "s.pnm")
            if grey:
                open(src, "wb").write(b"P5\n%d %d\n255\n" % (w, h) + np.ascontiguousarray(img[:, :, 1]).tobytes())
            else:
                open(src, "wb").write(b"P6\n%d %d\n255\n" % (w, h) + img.tobytes())
            jpg = os.path.join(OUT, name + ".jpg")
            r = subprocess.run([os.path.join(REF, "jpeg")] + args + [src, jpg], capture_output=True, text=True)
            assert r.returncode == 0 and os.path.exists(jpg), (name, r.stdout[-300:], r.stderr[-300:])
            raw = os.path.join(tmp, "o.raw")
            r = subprocess.run([os.path.join(REF, "refharness"), "decode", jpg, raw], capture_output=True, text=True)
            if r.returncode != 0:
                print(name, "reference harness:", r.stdout.strip(), r.stderr.strip())
                continue
            ww, hh, c = (int(v) for v in r.stdout.split()[:3])
            px[name] = np.fromfile(raw, dtype=np.uint8).reshape(hh, ww, c)
            err = np.abs(px[name].astype(int) - (img[:, :, 1:2] if grey else img).astype(int))
            print(name, px[name].shape, os.path.getsize(jpg), "bytes, max error to the source", err.max())
    np.savez_compressed(os.path.join(OUT, "xt_pixels.npz"), **px)


if __name__ == "__main__":
    main()
