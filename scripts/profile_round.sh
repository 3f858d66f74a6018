#!/bin/bash
# Round profiling pass (one GPU): plain bench -> ncu launch list of the same command -> one full capture of the dominant kernel.
# Numbers printed under ncu are never bench values; only the launch list / report are kept.
set -u
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cam --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_${TAG}.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-cam --no-cpu-baseline --profile-range > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 120 python scripts/bench_attn.py 16 785 12 64 > gpurun_out/bench_attn_${TAG}.json || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_bwd_${TAG} \
  python scripts/bench_attn.py 16 785 12 64 > gpurun_out/ncu_bwd_${TAG}.log 2>&1
echo "full capture rc=$?"
