#!/bin/bash
# compute-sanitizer over a small heterogeneous batch (all golden vectors + a 641x479 frame + a corrupt + a truncated stream
# + a stream whose samples leave int16 + three progressive streams)
OUT=gpurun_out/sanitize
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
dec = libjpeg_b200.BatchDecoder(frames)
out = dec.new_output(); dec.upload(); dec.decode(out); torch.cuda.synchronize()
print("statuses", [dec.status(i) for i in range(len(frames))])
PY
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > $OUT/$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|statuses|Invalid|hazard" $OUT/$tool.log | head -12
done
