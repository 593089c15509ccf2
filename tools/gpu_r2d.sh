#!/bin/bash
# fused reconstruction: variants (CTAs per SM / prefetch) at 840 and 128 frames, then ncu of the main fused kernel
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== parity (quick)"; timeout 900 python -m pytest tests -m gpu -q -x -k "golden_vector or bench_frame or beyond_the_fast or region_client or reference_cli" 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
for v in default f12 f10; do
  if [ $v = default ]; then unset B200JPG_LIB; else export B200JPG_LIB=$PWD/libjpeg_b200/build/libb200jpg_$v.so; fi
  echo "== variant $v"; python tools/probe.py 840 128 2>&1 | grep frames: | tee -a $OUT/variants.txt
done
unset B200JPG_LIB
echo "== two-kernel"; B200JPG_NO_FUSED=1 python tools/probe.py 840 128 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== ncu full, fused kernel (default build)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reconstruct420 -s 6 -c 1 -o $OUT/prof_fused -f \
    python tools/probe.py 840 > $OUT/ncu_fused.log 2>&1
tail -2 $OUT/ncu_fused.log
bash tools/gpu_r2c.sh $TAG
