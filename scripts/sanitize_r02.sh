#!/bin/bash
# compute-sanitizer logs: memcheck over the GPU tests of the kernels written this round, racecheck over the
# tcgen05 attention kernels (mbarrier / TMEM pipelines) and the lattice kernels at small shapes.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "affinity_refine_tc or patch_cam or crf_from_patch or bilateral or consistency" > gpurun_out/sanitizer_memcheck_${TAG}.log 2>&1
echo "memcheck rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/bench_attn.py 2 197 3 64 > gpurun_out/sanitizer_racecheck_attn_${TAG}.log 2>&1
echo "racecheck attn rc=$?"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bilateral_matches" > gpurun_out/sanitizer_racecheck_lattice_${TAG}.log 2>&1
echo "racecheck lattice rc=$?"
