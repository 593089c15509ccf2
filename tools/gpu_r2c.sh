#!/bin/bash
# round-2 third GPU pass: where the progressive path (cfg4) spends its time -- launch list + DRAM bytes per launch
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest (new tests only)"; timeout 900 python -m pytest tests -m gpu -q -k "region_client or reference_cli or resynchronisation or without_eoi" 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== ncu launch list + dram bytes, cfg4, 256 frames"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__grid_size --clock-control none \
    -k regex:"progressive|unstuff|reconstruct420" -s 30 -c 30 --csv --log-file $OUT/cfg4_launches.csv \
    python bench.py --workload cfg4 --frames-per-gpu 256 --steps 1 --warmup 3 --distinct 4 --no-e2e --no-cpu-baseline > $OUT/ncu_cfg4.log 2>&1
python - $TAG <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/%s/cfg4_launches.csv' % (sys.argv[1] if len(sys.argv)>1 else 'r2c'))) if len(r)>10]
hdr=rows[0]
ik=hdr.index('Kernel Name'); im=hdr.index('Metric Name'); iv=hdr.index('Metric Value'); iid=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[iid],r[ik][:40]),{})[r[im]]=r[iv]
for (i,k),m in d.items():
    print(i,k,{a.split('.')[0].replace('__','_'):b for a,b in m.items()})
PY
