#!/bin/bash
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -25 | tee $OUT/pytest_gpu.txt
echo "== restart-less"; python tools/probe.py --workload cfg3n 840 64 1 2>&1 | grep frames: | tee -a $OUT/variants.txt
python tools/probe.py --workload cfg2n 1024 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== cfg4"; python tools/probe.py --workload cfg4 1024 2>&1 | grep frames: | tee -a $OUT/variants.txt
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none \
    -k regex:"spec_sync|unstuff_long|entropy_decode" -s 6 -c 3 --csv --log-file $OUT/cfg3n_launches.csv \
    python tools/probe.py --workload cfg3n 840 > $OUT/ncu_cfg3n.log 2>&1
python - $OUT/cfg3n_launches.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]: d.setdefault((int(r[iid]),r[ik][:45]),{})[r[im].split('.')[0]]=r[iv]
for k,m in sorted(d.items()): print(k,m)
PY
