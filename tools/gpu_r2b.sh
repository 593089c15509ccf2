#!/bin/bash
# round-2 second GPU pass: opcode microbench, parity tests, fused vs two-kernel reconstruction A/B, ncu of the fused kernel
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== opbench"; ./tools/opbench 2>&1 | tee $OUT/opbench.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== bench fused"; timeout 600 python bench.py --steps 5 --warmup 3 --distinct 8 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_fused.json
echo "== bench two-kernel"; B200JPG_NO_FUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --distinct 8 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_nofused.json
echo "== ncu full (one step)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reconstruct420_kernel|entropy_decode" -s 8 -c 2 -o $OUT/prof -f \
    python bench.py --steps 2 --warmup 3 --distinct 4 --no-e2e --no-cpu-baseline > $OUT/ncu_full_bench.log 2>&1
tail -3 $OUT/ncu_full_bench.log
ls -la $OUT
