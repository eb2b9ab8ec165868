#!/bin/bash
# GPU call 45 (last GPU seconds of the round): ncu --set full of the shipped weight-gradient kernel on the same stage-4 256 -> 256 launch
# as call 42, to put the DRAM traffic after the grid-order change next to the 4.2 GB before it
LOG=gpurun_out/r2_call45.log
. "$(dirname "$0")/r2_common.sh"
step "ncu --set full: wgrad launch 1 (256->256), shipped kernel"
t 70 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -o gpurun_out/wgrad_1_final_r02 -f \
  python scripts/bench_train_lidar.py --steps 1 --warmup 0 >> $LOG 2>&1
ncu -i gpurun_out/wgrad_1_final_r02.ncu-rep --page raw --csv > gpurun_out/ncu_raw_wgrad_1_final_r02.csv 2>> $LOG
step done
tail -n 3 $LOG
exit 0
