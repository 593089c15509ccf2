#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== cfg4 fused / scan-by-scan"
python tools/probe.py --workload cfg4 1024 256 64 2>&1 | grep frames: | tee -a $OUT/variants.txt
B200JPG_NO_PFUSE=1 python tools/probe.py --workload cfg4 256 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== restart-less"
python tools/probe.py --workload cfg3n 840 64 1 2>&1 | grep frames: | tee -a $OUT/variants.txt
python tools/probe.py --workload cfg2n 1024 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== cfg3"; python tools/probe.py 840 512 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== ncu launch list cfg4 256"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none \
    -k regex:"pf_|unstuff|reconstruct|idct_planes|progressive" -s 20 -c 14 --csv --log-file $OUT/cfg4_launches.csv \
    python tools/probe.py --workload cfg4 256 > $OUT/ncu_cfg4.log 2>&1
python - $TAG <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/%s/cfg4_launches.csv' % sys.argv[1])) if len(r)>10]
hdr=rows[0]
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid],r[ik][:40]),{})[r[im]]=r[iv]
for (i,k),m in d.items():
    print(i,k,{a.split('.')[0].replace('__','_'):b for a,b in m.items()})
PY
echo "== ncu full restart-less"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spec_sync|unstuff_long" -s 6 -c 2 -o $OUT/prof_spec -f \
    python tools/probe.py --workload cfg3n 840 > $OUT/ncu_spec.log 2>&1
tail -2 $OUT/ncu_spec.log
