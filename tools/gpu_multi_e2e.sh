#!/bin/bash
# the end-to-end line on N GPUs with the host share and the bare-copy ceiling in it, at two chunk sizes
N=${1:-4}
TAG=${2:-e2e$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== weak, chunk 32"; timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_weak_n${N}_chunk32.json
echo "== weak, chunk 96"; timeout 600 $TR --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunk 96 --e2e-producers 3 2>&1 | tail -1 | tee $OUT/bench_weak_n${N}_chunk96.json
ls -la $OUT
