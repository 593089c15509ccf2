"""Regenerates tests/golden/subsampling/: baseline streams with 3x / 4x chroma subsampling written and decoded by the
UNMODIFIED reference. They pin the oracle's restatement of VerticalFilterCore<3|4> / HorizontalFilterCore<3|4>
(upsampling/upsampler.cpp:171-268, 310-386) -- groundwork for SURVEY 8f4; the CUDA path covers factors 1 and 2.

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_subsampling.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "subsampling")
sys.path.insert(0, ROOT)
from libjpeg_b200.synth import source_image  # noqa: E402

CASES = [  # name, width, height, sampling option, restart interval, quality
    ("c411_100x70_z4_q80", 100, 70, "1x1,4x1,4x1", 4, 80),
    ("c410_97x53_z3_q75", 97, 53, "1x1,4x2,4x2", 3, 75),
    ("c3x3_100x70_q80", 100, 70, "1x1,3x3,3x3", 0, 80),
    ("c3x1_31x200_z2_q85", 31, 200, "1x1,3x1,3x1", 2, 85),
    ("c4x4_97x53_q70", 97, 53, "1x1,4x4,4x4", 0, 70),
    ("c2x4_100x70_z5_q80", 100, 70, "1x1,2x4,2x4", 5, 80),
    ("c1x3_31x200_q90", 31, 200, "1x1,1x3,1x3", 0, 90),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    pixels = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, sub, z, q in CASES:
            src = os.path.join(tmp, "in.ppm")
            with open(src, "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (w, h) + source_image(w, h, w * 1000 + h).tobytes())
            jpg = os.path.join(OUT, name + ".jpg")
            cmd = [os.path.join(REF, "jpeg"), "-q", str(q), "-bl", "-s", sub]
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
            pixels[name] = np.fromfile(raw, dtype=np.uint8).reshape(hh, ww, cc)
            print(name, os.path.getsize(jpg), "bytes ->", pixels[name].shape)
    np.savez_compressed(os.path.join(OUT, "subsampling_pixels.npz"), **pixels)


if __name__ == "__main__":
    main()
