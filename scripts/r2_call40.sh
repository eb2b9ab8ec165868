#!/bin/bash
# GPU call 40: weight-gradient kernel against the fp64 sum; torch.profiler table of the LiDAR training step with the own kernel
LOG=gpurun_out/r2_call40.log
. "$(dirname "$0")/r2_common.sh"
step "sparse training tests"
t 300 python -m pytest tests/test_sparse_train_gpu.py -q -s -m gpu 2>&1 | grep -v "^$" >> $LOG
step "profile of the training step"
t 120 python scripts/profile_train_lidar.py > gpurun_out/profile_train_lidar_own.txt 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 40
cut -c1-60,120-200 gpurun_out/profile_train_lidar_own.txt | head -24
exit 0
