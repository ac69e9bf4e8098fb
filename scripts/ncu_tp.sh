#!/bin/bash
# ncu --set full of the three test-particle kernels at ntp = 1e6 (second launch of each), into gpurun_out/<tag>_<kernel>.ncu-rep
tag=${1:-r02_tp}
for k in drift_kernel whm_tp_step_kernel helio_tp_step_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -o gpurun_out/${tag}_$k python scripts/tp_bench.py 1000000 1 > gpurun_out/${tag}_$k.log 2>&1
  echo "$k ncu rc=$?"
done
