#!/bin/bash
# the measurements committed under profiles/ for the code as it stands: GPU tests, smoke, bench lines of every workload, the
# reference arm, one-frame latencies, launch list + ncu summaries
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Core|NUMA" > $OUT/cpu.txt
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg3.json
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --distinct 32 2>&1 | tail -1 | tee $OUT/bench_cfg2.json
echo "== bench cfg4"; timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --distinct 16 2>&1 | tail -1 | tee $OUT/bench_cfg4.json
echo "== bench cfg3n"; timeout 600 python bench.py --workload cfg3n --steps 3 --warmup 3 --no-cpu-baseline --distinct 16 2>&1 | tail -1 | tee $OUT/bench_cfg3n.json
echo "== bench cfg5"; timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_cfg5.json
echo "== latency"; timeout 600 python tools/latency.py 2>&1 | tail -1 | tee $OUT/latency.json
echo "== sanitizer"; timeout 1200 bash tools/sanitize.sh $TAG/sanitize
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1
echo "== ncu cfg4"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pf_ac|pf_dc" -s 4 -c 4 -o $OUT/prof_cfg4 -f python tools/probe.py --workload cfg4 1024 > $OUT/ncu_cfg4.log 2>&1
echo "== ncu cfg3n"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spec_sync|unstuff_long|entropy_decode" -s 6 -c 3 -o $OUT/prof_cfg3n -f python tools/probe.py --workload cfg3n 840 > $OUT/ncu_cfg3n.log 2>&1
ls -la $OUT
