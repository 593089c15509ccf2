#!/bin/bash
TAG=${1:-r2h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== cfg4"; python tools/probe.py --workload cfg4 1024 256 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== restart-less"; python tools/probe.py --workload cfg3n 840 64 1 2>&1 | grep frames: | tee -a $OUT/variants.txt
python tools/probe.py --workload cfg2n 1024 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== ncu launch list cfg4 256 + cfg3n 840"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none \
    -k regex:"pf_|spec_sync|unstuff_long|entropy_decode" -s 8 -c 4 --csv --log-file $OUT/cfg4_launches.csv \
    python tools/probe.py --workload cfg4 256 > $OUT/ncu_cfg4.log 2>&1
grep -E "pf_|gpu__time|thread_inst" $OUT/cfg4_launches.csv | cut -d, -f5,13,15 | head -20
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none \
    -k regex:"spec_sync|unstuff_long|entropy_decode" -s 6 -c 3 --csv --log-file $OUT/cfg3n_launches.csv \
    python tools/probe.py --workload cfg3n 840 > $OUT/ncu_cfg3n.log 2>&1
cut -d, -f5,13,15 $OUT/cfg3n_launches.csv | tail -12
