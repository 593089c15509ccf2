#!/bin/bash
# multi-GPU evidence: platform ceiling of concurrent pinned copies, the weak-scaling bench line and the strong-scaling split of
# BASELINE config 3 (global batch 4096) on N GPUs of one box:  gpurun --gpus N -- 'bash tools/gpu_multi.sh N tag'
N=${1:-4}
TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pcie"; timeout 300 $TR --master-port 29511 tools/pcie_ceiling.py 2>&1 | tail -1 | tee $OUT/pcie_n$N.json
echo "== weak"; timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_weak_n$N.json
echo "== strong"; timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --global-batch 4096 2>&1 | tail -1 | tee $OUT/bench_strong_n$N.json
ls -la $OUT
