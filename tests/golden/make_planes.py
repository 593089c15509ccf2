"""Regenerates tests/golden/planes.npz: what the UNMODIFIED reference command line client writes with upsampling switched off
(`oracle/_ref/jpeg -U -c in.jpg out`: JPGTAG_DECODER_UPSAMPLE = false, one component per request, stripes of 8 * suby lines,
cmd/reconstruct.cpp:268-301) -- every component as a plane at its own resolution.  8-bit frames give bytes, 12-bit frames
big-endian 16-bit samples (stored here as uint16).

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_planes.py
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = ["c420_96x80_z6_q75", "c420_50x38_z4_q75", "c422_100x60_z5_q80", "c440_100x61_z3_q80", "c444_64x64_z16_q90", "g_40x24_z2_q75",
         "c420_127x255_z7_q30", "progressive/p420_127x255_z7_q30", "deep12/d12_420_50x38_q60", "deep12/d12_422_127x99_z5_q75"]
CASES += ["subsampling/" + os.path.basename(p)[:-4] for p in sorted(__import__("glob").glob(os.path.join(HERE, "subsampling", "*.jpg")))[:4]]


def reference_planes(path):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "o")
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "jpeg"), "-U", "-c", path, out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        planes = []
        c = 0
        while os.path.exists("%s_%d.h" % (out, c)):
            f = open("%s_%d.h" % (out, c)).read().split()  # PG ML +8 96 80
            prec, w, h = int(f[2]), int(f[3]), int(f[4])
            raw = np.fromfile("%s_%d.raw" % (out, c), dtype=np.uint8 if prec <= 8 else ">u2")
            planes.append(raw.reshape(h, w).astype(np.uint16))
            c += 1
        return planes


def main():
    out = {}
    for name in CASES:
        path = os.path.join(HERE, name + ".jpg")
        if not os.path.exists(path):
            print("skip", name)
            continue
        for c, p in enumerate(reference_planes(path)):
            out["%s__%d" % (name.replace("/", "__"), c)] = p
        print(name, [out[k].shape for k in sorted(out) if k.startswith(name.replace("/", "__") + "__")])
    np.savez_compressed(os.path.join(HERE, "planes.npz"), **out)


if __name__ == "__main__":
    main()
