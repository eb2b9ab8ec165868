#!/bin/bash
# GPU call 42: ncu --set full of the weight-gradient kernel on a stage-4 SubM 256 -> 256 layer (second wgrad launch of the backward
# pass: the first is conv_out) and on a stage-2 64-wide layer
LOG=gpurun_out/r2_call42.log
. "$(dirname "$0")/r2_common.sh"
step "ncu --set full: wgrad launches 1 (256->256) and 14 (narrow)"
for idx in 1 14; do
  t 150 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s $idx -c 1 -o gpurun_out/wgrad_${idx}_r02 -f \
    python scripts/bench_train_lidar.py --steps 1 --warmup 0 >> $LOG 2>&1
  ncu -i gpurun_out/wgrad_${idx}_r02.ncu-rep --page raw --csv > gpurun_out/ncu_raw_wgrad_${idx}_r02.csv 2>> $LOG
done
step done
tail -n 5 $LOG
exit 0
