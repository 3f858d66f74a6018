#!/bin/bash
# ncu --set full captures of the refinement-side kernels (PAMR, bilateral lattice, consistency loss), one launch each.
set -u
TAG=${1:-r01g}
timeout 300 python scripts/bench_refine.py > gpurun_out/bench_refine_${TAG}.json || exit 1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"pamr_iter_smem_kernel|pamr_affinity_reg_kernel|lattice_|consistency_rows_kernel" --launch-skip 0 -c 40 -f -o gpurun_out/prof_refine_${TAG} \
  python scripts/_refine_once.py > gpurun_out/ncu_refine_${TAG}.log 2>&1
echo "refine capture rc=$?"
