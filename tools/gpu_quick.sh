#!/bin/bash
# quick GPU iteration: parity tests, stage timings at 128 / 512 frames, optional ncu of one kernel
TAG=${1:-q}
KREGEX=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
python tools/probe.py 2>&1 | tail -14 | tee $OUT/probe.txt
if [ -n "$KREGEX" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 4 -c 1 -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 1 --frames-per-gpu 128 --distinct 4 --no-e2e --no-cpu-baseline > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
fi
