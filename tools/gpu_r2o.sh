#!/bin/bash
TAG=${1:-r2o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (new tests)"; timeout 900 python -m pytest tests -m gpu -q -k "device_bitmap or restartless or xt_residual or synthetic_frames" 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -15 | tee $OUT/pytest_gpu.txt
echo "== spec variants"; for v in "" spec3 specold; do echo "variant [$v]"; if [ -n "$v" ]; then export B200JPG_LIB=$PWD/libjpeg_b200/build/libb200jpg_$v.so; else unset B200JPG_LIB; fi; python tools/probe.py --workload cfg3n 840 1 2>&1 | grep frames: ; python tools/probe.py --workload cfg2n 1024 2>&1 | grep frames: ; done | tee $OUT/spec_variants.txt; unset B200JPG_LIB
echo "== latency"; timeout 600 python tools/latency.py 2>&1 | tail -1 | tee $OUT/latency.json
echo "== sanitizer"; timeout 900 bash tools/sanitize.sh $TAG/sanitize
ls -la $OUT
