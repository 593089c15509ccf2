#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -60 | tee $OUT/pytest_gpu.txt
echo "== cfg4"; python tools/probe.py --workload cfg4 1024 256 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== pf_ac profile"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pf_ac" -s 3 -c 1 -o $OUT/prof_pf -f python tools/probe.py --workload cfg4 256 > $OUT/ncu_pf.log 2>&1
ls -la $OUT
