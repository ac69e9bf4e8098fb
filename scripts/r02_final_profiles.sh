#!/bin/bash
# Round-2 evidence run (one GPU): ncu --set full of the third-law kernel and of the three tp kernels, the ncu launch list
# of a short bench run, compute-sanitizer over every kernel family.  Outputs under gpurun_out/, summarised into profiles/.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kick_flat_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_kick_flat_final python scripts/kick_bench.py 100000 1 > gpurun_out/r02_ncu_flat_final.log 2>&1; echo "ncu flat rc=$?"
bash scripts/ncu_tp.sh r02_tp_final
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/r02_launches_bench.log 2>&1; echo "ncu launches rc=$?"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/r02_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitize_$tool.log | tail -1
done
