#!/bin/bash
OUT=gpurun_out/${1:-t}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "DeprecationWarning\|os.fork" > $OUT/pytest_full.txt
tail -60 $OUT/pytest_full.txt | cut -c1-220
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "identical_geometry or beyond_the_fast or truncated_streams or restart_index" 2>&1 | grep -v "DeprecationWarning\|os.fork" | head -60 | cut -c1-250 > $OUT/sanitize.txt
head -50 $OUT/sanitize.txt
