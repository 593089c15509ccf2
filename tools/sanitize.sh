#!/bin/bash
# compute-sanitizer over a small heterogeneous batch (all golden vectors + a 641x479 frame + a corrupt + a truncated stream
# + a stream whose samples leave int16 + three progressive streams)
OUT=gpurun_out/${1:-sanitize}
mkdir -p $OUT
cat > /tmp/san.py <<'PY'
import sys, glob, os
sys.path.insert(0, ".")
import numpy as np, torch
import libjpeg_b200
from libjpeg_b200 import synth
frames = [open(p, "rb").read() for p in sorted(glob.glob("tests/golden/*.jpg"))]
frames.append(synth.encode(synth.source_image(641, 479, 3), 75, (2, 2), 13).tobytes())
good = synth.encode(synth.source_image(128, 64, 5), 75, (2, 2), 8).tobytes()
bad = bytearray(good); i = bad.find(b"\xff\xda") + 14
for k in range(i + 40, i + 400):
    if bad[k] != 0xFF and bad[k - 1] != 0xFF: bad[k] = 0xF7
frames.append(bytes(bad))
frames.append(good[:good.rfind(b"\xff\xd3")] + b"\xff\xd9")
sys.path.insert(0, "tests")
import oracle_binding
frames.append(oracle_binding.with_dc_quantiser(good, 255))  # samples beyond int16: exercises the exact int32 pass
frames += [open(p, "rb").read() for p in sorted(glob.glob("tests/golden/progressive/*.jpg"))[:3]]  # progressive scans
# round 2: restart-less scans (speculative synchronisation), JPEG XT residual layers, 12-bit frames, 3x / 4x subsampling,
# four components, damaged restart markers; then the whole batch once more as planes without upsampling
frames.append(synth.encode(synth.source_image(640, 360, 7), 75, (2, 2), 0).tobytes())
frames += [open(p, "rb").read() for p in sorted(glob.glob("tests/golden/xt/*.jpg")) if "__nimpl" not in p][:4]
frames += [open(p, "rb").read() for p in sorted(glob.glob("tests/golden/deep12/*.jpg"))[:3]]
frames += [open(p, "rb").read() for p in sorted(glob.glob("tests/golden/subsampling/*.jpg"))[:4]]
frames.append(oracle_binding.with_fourth_component(synth.encode(synth.source_image(70, 50, 5), 80, (2, 2), 5, 1)))
frames.append(oracle_binding.with_restart_damage(synth.encode(synth.source_image(320, 240, 5), 75, (2, 2), 5).tobytes(), oracle_binding.RESTART_DAMAGES[0]))
dec = libjpeg_b200.BatchDecoder(frames, tolerate_bad=True)
out = dec.new_output(); dec.upload(); dec.decode(out); torch.cuda.synchronize()
print("statuses", [dec.status(i) for i in range(len(frames))])
plain = [f for i, f in enumerate(frames) if dec.status(i) == 0 and b"RESI" not in f]
dec2 = libjpeg_b200.BatchDecoder(plain, upsample=False)
out2 = dec2.new_output(); dec2.upload(); dec2.decode(out2); torch.cuda.synchronize()
print("planes statuses", [dec2.status(i) for i in range(len(plain))])
PY
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > $OUT/$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|statuses|Invalid|hazard" $OUT/$tool.log | head -12
done
