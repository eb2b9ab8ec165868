#!/bin/bash
# GPU call 46 (the round's last GPU seconds): per-offset chunking of the weight-gradient kernel: its tests + the training step
LOG=gpurun_out/r2_call46.log
. "$(dirname "$0")/r2_common.sh"
step "sparse training tests"
t 40 python -m pytest tests/test_sparse_train_gpu.py -q -s -m gpu 2>&1 | grep -v "^$" | tail -n 16 >> $LOG
step "LiDAR training step"
t 30 python scripts/bench_train_lidar.py --steps 6 --warmup 3 >> $LOG 2>&1
step done
tail -n 20 $LOG | cut -c1-330
exit 0
