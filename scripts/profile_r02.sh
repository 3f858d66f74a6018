#!/bin/bash
# Round-2 profiling pass (one GPU).  Plain runs first (numbers printed under ncu are never bench values), then the ncu launch
# list of the bench command, then one `--set full` capture per kernel family, then the compute-sanitizer logs.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cam --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || { echo "plain bench failed"; tail -5 gpurun_out/bench_${TAG}.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --no-cam --no-cpu-baseline --profile-range > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 120 python scripts/bench_attn.py 16 785 12 64 > gpurun_out/bench_attn_${TAG}.json || exit 1
# bench_attn.py runs the backward 11 times each without G, with fp32 G and with sign codes (the training step's mode): skip into the third block
for K in attn_bwd_kernel attn_fwd_kernel attn_mean_kernel; do
  SKIP=4; [ $K = attn_bwd_kernel ] && SKIP=26
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP -c 2 -f -o gpurun_out/prof_${K}_${TAG} \
    python scripts/bench_attn.py 16 785 12 64 > gpurun_out/ncu_${K}_${TAG}.log 2>&1
  echo "$K capture rc=$?"
done
timeout 300 python scripts/bench_refine.py > gpurun_out/bench_refine_${TAG}.json
timeout 120 python scripts/bench_cam_tc.py > gpurun_out/bench_cam_tc_${TAG}.json
timeout 120 python scripts/bench_consistency.py > gpurun_out/bench_consistency_${TAG}.json
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"pamr_iter|pamr_affinity|lattice_|consistency_rows_kernel|refine_tc_kernel|crf_head|bilinear_up_bwd" --launch-skip 0 -c 40 -f -o gpurun_out/prof_refine_${TAG} \
  python scripts/_refine_once.py > gpurun_out/ncu_refine_${TAG}.log 2>&1
echo "refine capture rc=$?"
