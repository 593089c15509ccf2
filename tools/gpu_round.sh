#!/bin/bash
# One gpurun call's worth of measurement: parity tests, bench line, ncu launch list + full captures.
# usage: tools/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Core" > $OUT/cpu.txt
echo "== pytest -m gpu"; python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee $OUT/bench.json
if [ "$2" != "quick" ]; then
echo "== bench reference arm"; python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list (same workload as the bench line: 840 frames per step)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --distinct 4 --no-e2e --no-cpu-baseline > $OUT/ncu_launch_bench.log 2>&1
tail -12 $OUT/launches.csv
echo "== ncu full (one step: every kernel once)"
ncu --set full --clock-control none --import-source on -k regex:"unstuff|entropy_decode|reconstruct_kernel|idct_planes|narrow_list" -s 14 -c 7 -o $OUT/prof -f \
    python bench.py --steps 2 --warmup 3 --distinct 4 --no-e2e --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
fi
