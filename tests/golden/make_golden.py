"""Regenerates tests/golden/ with the UNMODIFIED reference built by oracle/Makefile (oracle/_ref/jpeg as the
encoder, oracle/_ref/refharness -- public API, 8-row stripes -- as the ground-truth decoder).

Run in the build container only (needs /root/reference to have been compiled: `make -C oracle ref`):
    python tests/golden/make_golden.py
Outputs: <name>.jpg codestreams and golden_pixels.npz (name -> uint8 [H,W,C] pixels the reference wrote
through its BitMapHook).  The script is committed so the vectors can be reproduced.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
from libjpeg_b200.synth import source_image  # noqa: E402

# name, width, height, grey, reference-CLI sampling option, restart interval (MCUs), quality, extra options
CASES = [
    ("c420_96x80_z6_q75", 96, 80, False, "1x1,2x2,2x2", 6, 75, []),
    ("c420_50x38_z4_q75", 50, 38, False, "1x1,2x2,2x2", 4, 75, []),           # odd sizes, partial MCUs
    ("c444_64x64_z16_q90", 64, 64, False, None, 16, 90, []),
    ("c444_17x9_q95", 17, 9, False, None, 0, 95, []),                          # no DRI
    ("c420_33x17_q50", 33, 17, False, "1x1,2x2,2x2", 0, 50, []),
    ("c422_100x60_z5_q80", 100, 60, False, "1x1,2x1,2x1", 5, 80, []),
    ("c440_100x61_z3_q80", 100, 61, False, "1x1,1x2,1x2", 3, 80, []),
    ("c420_1x1_q75", 1, 1, False, "1x1,2x2,2x2", 0, 75, []),
    ("c420_8x8_z1_q75", 8, 8, False, "1x1,2x2,2x2", 1, 75, []),
    ("c420_127x255_z7_q30", 127, 255, False, "1x1,2x2,2x2", 7, 30, []),
    ("c420_130x70_z9_q98", 130, 70, False, "1x1,2x2,2x2", 9, 98, []),          # long Huffman codes
    ("c420_160x48_z10_q75_opt", 160, 48, False, "1x1,2x2,2x2", 10, 75, ["-h"]),  # optimised (non Annex-K) tables
    ("g_40x24_z2_q75", 40, 24, True, None, 2, 75, []),                         # one component
    ("cfg1_512x512_444_z256_q90", 512, 512, False, None, 256, 90, []),         # BASELINE.json configs[0]
]


def write_pnm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P%d\n%d %d\n255\n" % (5 if img.ndim == 2 else 6, w, h))
        f.write(img.tobytes())


def main():
    pixels = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, grey, sub, z, q, extra in CASES:
            seed = 1 if name.startswith("cfg1") else (w * 1000 + h)
            img = source_image(w, h, seed)
            if grey:
                img = np.ascontiguousarray(img[..., 1])
            pnm = os.path.join(tmp, "in.pnm")
            write_pnm(pnm, img)
            jpg = os.path.join(HERE, name + ".jpg")
            cmd = [os.path.join(REF, "jpeg"), "-q", str(q), "-bl"] + extra
            if sub:
                cmd += ["-s", sub]
            if z:
                cmd += ["-z", str(z)]
            subprocess.run(cmd + [pnm, jpg], check=True, capture_output=True)
            raw = os.path.join(tmp, "out.raw")
            r = subprocess.run([os.path.join(REF, "refharness"), "decode", jpg, raw], check=True, capture_output=True, text=True)
            ww, hh, dd = map(int, r.stdout.split())
            assert (ww, hh) == (w, h)
            pixels[name] = np.fromfile(raw, dtype=np.uint8).reshape(h, w, dd)
            print(name, os.path.getsize(jpg), "bytes", pixels[name].shape)
    np.savez_compressed(os.path.join(HERE, "golden_pixels.npz"), **pixels)


if __name__ == "__main__":
    main()
