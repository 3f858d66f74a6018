#!/bin/bash
# Refresh after the 8-softmax-warp forward kernel: GPU tests, plain bench, ncu launch list of the bench command, `--set full`
# captures of attn_fwd_kernel / attn_mean_kernel (plain runs first: numbers printed under ncu are never bench values).
set -u
TAG=${1:-r02k}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3) > gpurun_out/pytest_${TAG}.log
cat gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-cam --no-cpu-baseline --profile-range > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in attn_fwd_kernel attn_mean_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 2 -f -o gpurun_out/prof_${K}_${TAG} \
    python scripts/bench_attn.py 16 785 12 64 > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  echo "$K capture rc=$?"
done
