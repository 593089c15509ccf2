#!/bin/bash
TAG=${1:-r2l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|os.fork\|^$\|Docs:" | tail -40 | tee $OUT/pytest_gpu.txt
echo "== cfg4"; python tools/probe.py --workload cfg4 1024 256 2>&1 | grep frames: | tee -a $OUT/variants.txt
echo "== a1 variants"; for v in "" a1old a1pk a1sh; do echo "variant [$v]"; if [ -n "$v" ]; then export B200JPG_LIB=$PWD/libjpeg_b200/build/libb200jpg_$v.so; else unset B200JPG_LIB; fi; python tools/probe.py 840 512 2>&1 | grep frames: ; done | tee $OUT/a1_variants.txt; unset B200JPG_LIB
echo "== p2d chunks"; for c in 2 3 4; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --p2d-chunks $c 2>&1 | tail -1 > $OUT/bench_p2d$c.json; python - $OUT/bench_p2d$c.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(d["value"], d["pinned_to_device_rgb"]["chunk_frames"], d["pinned_to_device_rgb"]["value"], d["e2e"]["value"])
PY
done | tee $OUT/p2d.txt
echo "== pipes"; timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"^reconstruct_kernel|entropy_decode_kernel|idct_planes_kernel|unstuff_kernel" -s 6 -c 4 --csv --log-file $OUT/cfg3_pipes.csv python tools/probe.py 840 > $OUT/ncu_pipes.log 2>&1
echo "== pf_ac source profile"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pf_ac|pf_dc" -s 4 -c 2 -o $OUT/prof_pf -f python tools/probe.py --workload cfg4 256 > $OUT/ncu_pf.log 2>&1
ls -la $OUT
