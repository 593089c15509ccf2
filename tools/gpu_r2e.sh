#!/bin/bash
# fused kernel with the instruction-selection work, restart-less path: parity, timings, profile
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
for v in default plain; do
  if [ $v = default ]; then unset B200JPG_LIB; else export B200JPG_LIB=$PWD/libjpeg_b200/build/libb200jpg_$v.so; fi
  echo "== variant $v"; python tools/probe.py 840 128 2>&1 | grep frames: | tee -a $OUT/variants.txt
done
unset B200JPG_LIB
echo "== restart-less cfg3n / cfg2n, then the single-work-item path"
python tools/probe.py --workload cfg3n 840 64 1 2>&1 | grep frames: | tee -a $OUT/variants.txt
python tools/probe.py --workload cfg2n 1024 2>&1 | grep frames: | tee -a $OUT/variants.txt
B200JPG_NO_SPEC=1 python tools/probe.py --workload cfg3n 64 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== ncu full, fused kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reconstruct420 -s 6 -c 1 -o $OUT/prof_fused -f \
    python tools/probe.py 840 > $OUT/ncu_fused.log 2>&1
echo "== ncu full, restart-less kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spec_sync|entropy_decode|unstuff" -s 9 -c 3 -o $OUT/prof_spec -f \
    python tools/probe.py --workload cfg3n 840 > $OUT/ncu_spec.log 2>&1
tail -2 $OUT/ncu_spec.log
