"""Regenerates tests/golden/progressive/: SOF2 streams written by the UNMODIFIED reference encoder (`jpeg -v`: DC first,
AC bands, DC / AC refinement scans) with the pixels the unmodified reference decodes from them. They pin the oracle's
restatement of codestream/sequentialscan.cpp (first passes) and codestream/refinementscan.cpp -- groundwork for SURVEY
8f2; the CUDA path does not decode progressive streams yet (b200jpg_parse reports NOT_IMPLEMENTED for SOF2).

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_progressive.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "progressive")
sys.path.insert(0, ROOT)
from libjpeg_b200.synth import source_image  # noqa: E402

# name, width, height, grey, sampling option of the reference CLI, restart interval (MCUs), quality
CASES = [
    ("p420_96x80_z3_q75", 96, 80, False, "1x1,2x2,2x2", 3, 75),
    ("p420_50x38_q75", 50, 38, False, "1x1,2x2,2x2", 0, 75),          # odd sizes, no restart markers
    ("p444_64x64_z16_q90", 64, 64, False, None, 16, 90),
    ("p420_127x255_z7_q30", 127, 255, False, "1x1,2x2,2x2", 7, 30),   # long EOB runs
    ("pg_40x24_z2_q75", 40, 24, True, None, 2, 75),
    ("p422_100x60_z5_q80", 100, 60, False, "1x1,2x1,2x1", 5, 80),
    ("p420_130x70_z9_q98", 130, 70, False, "1x1,2x2,2x2", 9, 98),     # many significant coefficients to refine
]


def main():
    os.makedirs(OUT, exist_ok=True)
    pixels = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, grey, sub, z, q in CASES:
            img = source_image(w, h, w * 1000 + h)
            src = os.path.join(tmp, "in.pnm")
            with open(src, "wb") as f:
                if grey:
                    f.write(b"P5\n%d %d\n255\n" % (w, h) + np.ascontiguousarray(img[..., 1]).tobytes())
                else:
                    f.write(b"P6\n%d %d\n255\n" % (w, h) + img.tobytes())
            jpg = os.path.join(OUT, name + ".jpg")
            cmd = [os.path.join(REF, "jpeg"), "-q", str(q), "-v"]
            if sub:
                cmd += ["-s", sub]
            if z:
                cmd += ["-z", str(z)]
            r = subprocess.run(cmd + [src, jpg], capture_output=True, text=True)
            if r.returncode != 0:
                sys.exit("reference encoder failed on %s: %s" % (name, r.stderr))
            raw = os.path.join(tmp, "o.raw")
            r = subprocess.run([os.path.join(REF, "refharness"), "decode", jpg, raw], capture_output=True, text=True)
            if r.returncode != 0:
                sys.exit("reference decoder failed on %s: %s" % (name, r.stderr))
            ww, hh, cc = (int(v) for v in r.stdout.split()[:3])
            px = np.fromfile(raw, dtype=np.uint8).reshape(hh, ww, cc)
            pixels[name] = px[..., 0] if cc == 1 else px
            print(name, os.path.getsize(jpg), "bytes ->", px.shape)
    np.savez_compressed(os.path.join(OUT, "progressive_pixels.npz"), **pixels)


if __name__ == "__main__":
    main()
