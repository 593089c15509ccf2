#!/bin/bash
# short closing pass: the whole GPU suite, smoke, the headline line and the lines whose kernels changed since gpu_final.sh ran
TAG=${1:-final2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg3.json
echo "== bench cfg3n"; timeout 600 python bench.py --workload cfg3n --steps 3 --warmup 3 --no-cpu-baseline --distinct 16 2>&1 | tail -1 | tee $OUT/bench_cfg3n.json
echo "== latency"; timeout 600 python tools/latency.py 2>&1 | tail -1 | tee $OUT/latency.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1
echo "== ncu cfg3n"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spec_sync|unstuff_long|entropy_decode" -s 6 -c 3 -o $OUT/prof_cfg3n -f python tools/probe.py --workload cfg3n 840 > $OUT/ncu_cfg3n.log 2>&1
ls -la $OUT
