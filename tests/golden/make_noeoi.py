"""Regenerates tests/golden/noeoi/: golden codestreams whose tail was cut off (no EOI, possibly in the middle of the
entropy coded data), with what the UNMODIFIED reference (oracle/_ref/refharness) makes of them:
  * nothing missing but the EOI, or the cut lies inside the LAST restart interval: the reference warns, feeds zero bits
    behind the end of the data (io/bitstream.cpp:103-105) and delivers the image (Frame::ParseTrailer marker/frame.cpp:1089,
    Image::ParseTrailer codestream/image.cpp:1466) -> its pixels are stored;
  * the cut removes a restart marker the scan still needs: EntropyParser::ParseRestartMarker runs out of data while
    resynchronising and throws UNEXPECTED_EOF (codestream/entropyparser.cpp:141-147) -> status -1025 is stored.

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_noeoi.py
"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(HERE, "noeoi")

CASES = [  # name, source golden vector, bytes cut from the end
    ("noeoi_c420_96x80", "c420_96x80_z6_q75", 2),
    ("cut300_c420_96x80", "c420_96x80_z6_q75", 300),
    ("cut700_c420_96x80", "c420_96x80_z6_q75", 700),
    ("cut700_c444_64x64", "c444_64x64_z16_q90", 700),
    ("cut30_c422_100x60", "c422_100x60_z5_q80", 30),
    ("cut100_c422_100x60", "c422_100x60_z5_q80", 100),
    ("cut9_g_40x24", "g_40x24_z2_q75", 9),
    ("cut40_c420_33x17", "c420_33x17_q50", 40),
    ("cut1_p420_96x80", "progressive/p420_96x80_z3_q75", 1),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    pixels, status = {}, {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, src, cut in CASES:
            data = open(os.path.join(HERE, src + ".jpg"), "rb").read()[:-cut]
            jpg = os.path.join(OUT, name + ".jpg")
            open(jpg, "wb").write(data)
            raw = os.path.join(tmp, "o.raw")
            r = subprocess.run([os.path.join(REF, "refharness"), "decode", jpg, raw], capture_output=True, text=True)
            if r.returncode != 0:
                status[name] = int(r.stdout.split()[1])
                print(name, len(data), "bytes -> reference error", status[name])
                continue
            w, h, c = (int(v) for v in r.stdout.split()[:3])
            px = np.fromfile(raw, dtype=np.uint8).reshape(h, w, c)
            pixels[name] = px[..., 0] if c == 1 else px
            status[name] = 0
            print(name, len(data), "bytes ->", px.shape)
    np.savez_compressed(os.path.join(OUT, "noeoi_pixels.npz"), **pixels)
    json.dump(status, open(os.path.join(OUT, "noeoi_status.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
