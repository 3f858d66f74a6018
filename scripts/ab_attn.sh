#!/bin/bash
# A/B of builds of the same ABI: scripts/bench_attn.py once per library in scripts/micro/ (and the in-tree one)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in base "$@"; do
  if [ "$n" = base ]; then lib=acr_wsss_b200/libacr_b200.so; else lib=scripts/micro/libacr_b200_$n.so; fi
  echo "== $n" >> gpurun_out/ab_attn.log
  ACR_B200_LIB=$PWD/$lib timeout 300 python scripts/bench_attn.py $AB_ARGS >> gpurun_out/ab_attn.log 2>&1
done
