#!/bin/bash
# GPU call 39: own weight-gradient kernel (csrc/wgrad.cu): sparse-training tests, LiDAR training step with both weight-gradient paths
LOG=gpurun_out/r2_call39.log
. "$(dirname "$0")/r2_common.sh"
step "sparse training tests"
t 300 python -m pytest tests/test_sparse_train_gpu.py -q -s -m gpu 2>&1 | grep -v "^$" >> $LOG
step "LiDAR training step, own wgrad"
ISF_WGRAD=own t 120 python scripts/bench_train_lidar.py --steps 6 --warmup 3 >> $LOG 2>&1
step "LiDAR training step, library wgrad"
ISF_WGRAD=torch t 120 python scripts/bench_train_lidar.py --steps 6 --warmup 3 >> $LOG 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 60
exit 0
