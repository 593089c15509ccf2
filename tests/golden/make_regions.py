"""Regenerates tests/golden/regions.npz: what the UNMODIFIED reference (oracle/_ref/libjpegref.a) delivers to
tests/client/region_client.cpp -- horizontal crops, planar client bitmaps, a BitMapHook that fails. The same client linked
against libb200jpg.so must produce the same bytes and the same report line (tests/test_gpu_parity.py).

Crops start at multiples of eight (or in a component that is not subsampled): with MINX inside a block of a subsampled frame
the reference replicates the chroma edge at the crop instead of at the image border, so its first partial block differs from
its own full decode -- an artefact of its windowed upsampler, not a contract (known deviations).

Run in the build container only (needs /root/reference and `make -C oracle ref`):  python tests/golden/make_regions.py
"""
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = [  # key, golden vector, client arguments
    ("crop_c420_96x80_16_61", "c420_96x80_z6_q75", ["crop", "16", "61"]),
    ("crop_c420_96x80_8_15", "c420_96x80_z6_q75", ["crop", "8", "15"]),
    ("crop_c422_100x60_0_50", "c422_100x60_z5_q80", ["crop", "0", "50"]),
    ("crop_c444_64x64_33_999", "c444_64x64_z16_q90", ["crop", "33", "999"]),
    ("crop_g_40x24_5_5", "g_40x24_z2_q75", ["crop", "5", "5"]),
    ("crop_c420_127x255_40_100", "c420_127x255_z7_q30", ["crop", "40", "100"]),
    ("planar_c420_50x38", "c420_50x38_z4_q75", ["planar"]),
    ("planar_c440_100x61", "c440_100x61_z3_q80", ["planar"]),
    ("hookerr_c420_96x80_5", "c420_96x80_z6_q75", ["hookerr", "5"]),
    ("hookerr_c444_17x9_1", "c444_17x9_q95", ["hookerr", "1"]),
]


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "region_ref")
        subprocess.run(["g++", "-O1", "-w", "-DUSE_AUTOCONF", "-I" + os.path.join(ROOT, "oracle", "ref_config"), "-I/root/reference",
                        os.path.join(ROOT, "tests", "client", "region_client.cpp"), os.path.join(ROOT, "oracle", "_ref", "libjpegref.a"),
                        "-o", exe], check=True)
        for key, name, args in CASES:
            raw = os.path.join(tmp, "o.raw")
            r = subprocess.run([exe, os.path.join(HERE, name + ".jpg"), raw] + args, capture_output=True, text=True)
            assert r.returncode == 0, (key, r.stderr)
            out[key] = np.fromfile(raw, dtype=np.uint8)
            out[key + "__report"] = np.frombuffer(r.stdout.strip().encode(), dtype=np.uint8)
            print(key, r.stdout.strip())
    np.savez_compressed(os.path.join(HERE, "regions.npz"), **out)


if __name__ == "__main__":
    main()
