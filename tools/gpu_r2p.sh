#!/bin/bash
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (generic reconstruction)"; timeout 900 python -m pytest tests -m gpu -q -k "xt or four_component or 12bit or subsampling or planes or without_upsampling or colour_transform or cli" 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench cfg5"; timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_cfg5.json
ls -la $OUT
