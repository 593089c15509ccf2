"""Regenerates tests/golden/dnl/: streams whose frame height is 0 in the SOF and follows the first scan in a DNL marker
(`jpeg -n`, cmd/main.cpp:272; EntropyParser::ParseDNLMarker codestream/entropyparser.cpp:204-249), written by the reference
encoder, with the reference decoder's pixels.

One artefact is NOT part of the fixtures' contract: on a DNL stream the reference's chroma upsampler is built before the height
is known and its LAST pixel row differs from what the reference delivers for the very same image with the height in the SOF.
`dnl_pixels.npz` therefore holds, per vector, "<name>" = the reference's pixels of the twin stream with the height in the SOF
(what the oracle and the CUDA path produce) and "<name>__ref_dnl" = the reference's pixels of the DNL stream itself; the
tests check that the two agree everywhere but in that last row.

Run in the build container only (needs `make -C oracle ref`):  python tests/golden/make_dnl.py
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
OUT = os.path.join(HERE, "dnl")

from libjpeg_b200.synth import source_image  # noqa: E402

CASES = [("dnl_c420_96x80_z6_q75", 96, 80, "1x1,2x2,2x2", ["-z", "6"], 75), ("dnl_c420_50x38_q75", 50, 38, "1x1,2x2,2x2", [], 75),
         ("dnl_c444_64x40_z8_q90", 64, 40, None, ["-z", "8"], 90)]


def main():
    os.makedirs(OUT, exist_ok=True)
    px = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, sub, extra, q in CASES:
            ppm = os.path.join(tmp, "s.ppm")
            open(ppm, "wb").write(b"P6\n%d %d\n255\n" % (w, h) + source_image(w, h, 11).tobytes())
            args = ["-q", str(q), "-bl"] + (["-s", sub] if sub else []) + extra
            for tag, more in (("", ["-n"]), ("twin", [])):
                jpg = os.path.join(OUT, name + ".jpg") if not tag else os.path.join(tmp, "twin.jpg")
                subprocess.run([os.path.join(REF, "jpeg")] + args + more + [ppm, jpg], check=True, capture_output=True)
                raw = os.path.join(tmp, "o.raw")
                r = subprocess.run([os.path.join(REF, "refharness"), "decode", jpg, raw], capture_output=True, text=True, check=True)
                ww, hh, c = (int(v) for v in r.stdout.split()[:3])
                px[name + ("" if tag else "__ref_dnl")] = np.fromfile(raw, dtype=np.uint8).reshape(hh, ww, c)
            a, b = px[name], px[name + "__ref_dnl"]
            print(name, a.shape, "rows that differ between the reference's two decodes:", np.nonzero((a != b).any(axis=(1, 2)))[0])
    np.savez_compressed(os.path.join(OUT, "dnl_pixels.npz"), **px)


if __name__ == "__main__":
    main()
