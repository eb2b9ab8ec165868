#!/bin/bash
# GPU call 41 (and 43: the same with index prefetch + tap-fastest grid): weight-gradient kernel: tests, training step, profile
LOG=gpurun_out/r2_call${CALL:-41}.log
. "$(dirname "$0")/r2_common.sh"
step "sparse training tests"
t 300 python -m pytest tests/test_sparse_train_gpu.py -q -s -m gpu 2>&1 | grep -v "^$" >> $LOG
step "LiDAR training step, own wgrad"
t 120 python scripts/bench_train_lidar.py --steps 6 --warmup 3 >> $LOG 2>&1
step "profile of the training step"
t 120 python scripts/profile_train_lidar.py > gpurun_out/profile_train_lidar_own.txt 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 40
cut -c1-60,120-200 gpurun_out/profile_train_lidar_own.txt | head -24
exit 0
