#!/bin/bash
# A/B of kernel variants built by tools/build_variant.sh:  tools/gpu_exp.sh <tag> <workload> <frames> <variant> [<variant> ...]   ("main" = the library in the tree)
TAG=$1; WL=$2; NF=$3; shift 3
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  echo "variant [$v]"
  if [ "$v" != "main" ]; then export B200JPG_LIB=$PWD/libjpeg_b200/build/libb200jpg_$v.so; else unset B200JPG_LIB; fi
  python tools/probe.py --workload $WL $NF 2>&1 | grep "frames:\|Error\|error" | head -4
done | tee $OUT/variants_$WL.txt
