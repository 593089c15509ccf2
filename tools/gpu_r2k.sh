#!/bin/bash
# round-2 measurement pass: bench lines (cfg3 default, cfg2, cfg4, cfg3n, reference arm), recon chunking experiment,
# ncu launch list of the bench command, ncu --set full of the top kernels, pcie ceiling at N=1
TAG=${1:-r2k}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Socket|Core|NUMA" > $OUT/cpu.txt
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== bench cfg3"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_cfg3.json
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --distinct 32 2>&1 | tail -1 | tee $OUT/bench_cfg2.json
echo "== bench cfg4"; timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --distinct 16 2>&1 | tail -1 | tee $OUT/bench_cfg4.json
echo "== bench cfg3n"; timeout 600 python bench.py --workload cfg3n --steps 3 --warmup 3 --no-cpu-baseline --distinct 16 2>&1 | tail -1 | tee $OUT/bench_cfg3n.json
echo "== recon chunk"; for c in 0 4 8 16 32; do echo "chunk $c"; B200JPG_RECON_CHUNK=$c python tools/probe.py 840 2>&1 | grep frames: ; done | tee $OUT/chunk.txt
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full"; 
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^reconstruct_kernel|entropy_decode_kernel|idct_planes_kernel|unstuff_kernel" -s 6 -c 6 -o $OUT/prof_cfg3 -f \
    python tools/probe.py 840 > $OUT/ncu_cfg3.log 2>&1
B200JPG_RECON_CHUNK=8 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"^reconstruct_kernel|idct_planes_kernel" -s 212 -c 8 --csv --log-file $OUT/chunk8_dram.csv \
    python tools/probe.py 840 > $OUT/ncu_chunk8.log 2>&1
echo "== pcie"; timeout 300 python tools/pcie_ceiling.py 2>&1 | tail -2 | tee $OUT/pcie_n1.json
ls -la $OUT
