#!/bin/bash
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (two-kernel path for 4:2:0)"; B200JPG_NO_FUSED=1 timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $OUT/pytest_gpu_nofused.txt
echo "== pytest -m gpu (default)"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== fused"; python tools/probe.py 840 128 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== two-kernel, optimised arithmetic"; B200JPG_NO_FUSED=1 python tools/probe.py 840 128 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== cfg2 both"; python tools/probe.py --workload cfg2 4096 2>&1 | grep frames: | tee -a $OUT/variants.txt
B200JPG_NO_FUSED=1 python tools/probe.py --workload cfg2 4096 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== ncu full, two-kernel"
B200JPG_NO_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reconstruct_kernel|idct_planes" -s 12 -c 2 -o $OUT/prof_2k -f \
    python tools/probe.py 840 > $OUT/ncu_2k.log 2>&1
tail -2 $OUT/ncu_2k.log
