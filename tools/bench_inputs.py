"""Benchmark / test INPUTS (untimed setup, never the product path): the deterministic source images S(w,h,seed) of
SURVEY.md 8d, encoded by the REFERENCE ENCODER (oracle/_ref/jpeg, the unmodified thorfdbg/libjpeg CLI built by
oracle/Makefile; cmd/main.cpp:243,554 for -bl / -z) wherever that binary exists -- it travels to the GPU box with the
snapshot -- and by the repo's own generator (libjpeg_b200/synth.py) only as a stated fallback.

    workload  geometry                         reference encoder command line
    cfg1      512x512   4:4:4 q90 DRI=256      jpeg -q 90 -bl               -z 256 in.ppm out.jpg
    cfg2      1920x1080 4:2:0 q75 DRI=120      jpeg -q 75 -bl -s 1x1,2x2,2x2 -z 120
    cfg3      3840x2160 4:2:0 q75 DRI=240      jpeg -q 75 -bl -s 1x1,2x2,2x2 -z 240
    cfg4      3840x2160 4:2:0 q75 DRI=240 SOF2 jpeg -q 75 -v  -s 1x1,2x2,2x2 -z 240      (progressive)
    cfg3n     cfg3 without restart markers     jpeg -q 75 -bl -s 1x1,2x2,2x2             (DRI-less streams)
    cfg5      8192x8192 4:2:0 q75 DRI=512 + XT  jpeg -q 75 -bl -s 1x1,2x2,2x2 -z 512 -r -Q 90  (JPEG XT: 4:4:4 q90 DCT residual in a RESI
                                                box, 8-bit integer profile -- BASELINE config 5's geometry; its HDR profile is not covered)
"""
import hashlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "jpeg")

WORKLOADS = {
    # name: (width, height, quality, subsampling (hs, vs) of luma, restart interval in MCUs, progressive, description)
    "cfg1": (512, 512, 90, (1, 1), 256, False, "cfg1: 512x512 4:4:4 q90 baseline, DRI=256 (4 MCU rows)"),
    "cfg2": (1920, 1080, 75, (2, 2), 120, False, "cfg2: 1920x1080 4:2:0 q75 baseline, DRI=120 (one restart interval per MCU row)"),
    "cfg3": (3840, 2160, 75, (2, 2), 240, False,
             "cfg3: 3840x2160 4:2:0 q75 baseline, DRI=240 (one restart interval per MCU row), Annex-K tables"),
    "cfg4": (3840, 2160, 75, (2, 2), 240, True, "cfg4: 3840x2160 4:2:0 q75 progressive (SOF2, ten scans), DRI=240"),
    "cfg3n": (3840, 2160, 75, (2, 2), 0, False, "cfg3n: 3840x2160 4:2:0 q75 baseline WITHOUT restart markers"),
    "cfg2n": (1920, 1080, 75, (2, 2), 0, False, "cfg2n: 1920x1080 4:2:0 q75 baseline WITHOUT restart markers"),
    "cfg5": (8192, 8192, 75, (2, 2), 512, False,
             "cfg5: 8192x8192 4:2:0 q75 baseline + JPEG XT residual layer (4:4:4 q90 DCT codestream in the RESI box, 8-bit integer profile), DRI=512"),
}
EXTRA_ARGS = {"cfg5": ["-r", "-Q", "90"]}  # more encoder arguments of a workload


def have_reference_encoder():
    return os.path.exists(REF_CLI) and os.access(REF_CLI, os.X_OK)


def encoder_name(workload):
    w, h, q, sub, z, prog, _ = WORKLOADS[workload]
    if have_reference_encoder():
        return "reference: oracle/_ref/jpeg " + " ".join(_ref_args(q, sub, z, prog) + EXTRA_ARGS.get(workload, []))
    if prog or workload in EXTRA_ARGS:
        return "unavailable (the progressive workload needs the reference encoder)"
    return "synth fallback: libjpeg_b200/csrc/synth_encoder.cpp (oracle/_ref/jpeg is not in this snapshot)"


def _ref_args(q, sub, z, prog):
    a = ["-q", str(q), "-v" if prog else "-bl"]
    if sub != (1, 1):
        a += ["-s", "1x1,%dx%d,%dx%d" % (sub[0], sub[1], sub[0], sub[1])]
    if z:
        a += ["-z", str(z)]
    return a


def _cache_dir():
    d = os.environ.get("B200JPG_INPUT_CACHE") or os.path.join(tempfile.gettempdir(), "b200jpg_inputs")
    os.makedirs(d, exist_ok=True)
    return d


def encode_one(args):
    """(workload, seed) -> codestream bytes. Cached on disk under the temp directory (keyed by workload, seed, encoder)."""
    workload, seed = args
    w, h, q, sub, z, prog, _ = WORKLOADS[workload]
    from libjpeg_b200 import synth
    ref = have_reference_encoder()
    if not ref and (prog or workload in EXTRA_ARGS):
        raise RuntimeError("this workload needs the reference encoder (oracle/_ref/jpeg)")
    key = hashlib.sha1(("%s|%d|%s|v2" % (workload, seed, "ref" if ref else "synth")).encode()).hexdigest()[:16]
    path = os.path.join(_cache_dir(), "%s_%d_%s.jpg" % (workload, seed, key))
    if os.path.exists(path):
        return open(path, "rb").read()
    img = synth.source_image(w, h, seed)
    if ref:
        with tempfile.TemporaryDirectory() as tmp:
            ppm, jpg = os.path.join(tmp, "s.ppm"), os.path.join(tmp, "s.jpg")
            with open(ppm, "wb") as f:
                f.write(b"P6\n%d %d\n255\n" % (w, h))
                f.write(img.tobytes())
            r = subprocess.run([REF_CLI] + _ref_args(q, sub, z, prog) + EXTRA_ARGS.get(workload, []) + [ppm, jpg], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("reference encoder failed: " + r.stderr[-300:])
            data = open(jpg, "rb").read()
    else:
        data = synth.encode(img, q, sub, z).tobytes()
    tmp_path = path + ".%d.tmp" % os.getpid()
    with open(tmp_path, "wb") as f:
        f.write(data)
    os.replace(tmp_path, path)
    return data


def make_frames(workload, distinct, workers=1, first_seed=1):
    """`distinct` codestreams of the workload (seeds first_seed ..), encoded in `workers` processes."""
    jobs = [(workload, s) for s in range(first_seed, first_seed + distinct)]
    if workers <= 1 or distinct <= 1:
        return [encode_one(j) for j in jobs]
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(max_workers=min(workers, distinct)) as ex:
        return list(ex.map(encode_one, jobs))
