#!/bin/bash
# round-2 first GPU pass: parity tests (all, no -x), bench lines of cfg3 / cfg2 / cfg4, reference arm
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Core|NUMA" > $OUT/cpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee $OUT/bench_cfg3.json
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --distinct 32 2>&1 | tail -2 | tee $OUT/bench_cfg2.json
echo "== bench cfg4"; timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --distinct 16 2>&1 | tail -2 | tee $OUT/bench_cfg4.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
