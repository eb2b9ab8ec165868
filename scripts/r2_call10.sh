#!/bin/bash
# GPU call 10 (2 GPUs): a15 over NCCL, bench.py at N=2, configs[4] at N=2; chunked tile orders on the v4 kernel
LOG=gpurun_out/r2_call10.log
. "$(dirname "$0")/r2_common.sh"
nvidia-smi -L >> $LOG 2>&1
step "a15 over NCCL: train step of DynamicVFE on 2 GPUs"
t 400 python -m pytest tests/test_dist_gpu.py -q -m gpu -s >> $LOG 2>&1
t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_vfe_train.py >> $LOG 2>&1
step "bench.py N=2"
t 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r2c10_bench_n2.json 2>> $LOG
step "configs[4] N=2"
t 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_config5.py > gpurun_out/r2_config5_n2.json 2>> $LOG
step "tile order chunk sizes on the v4 kernel"
t 300 python scripts/bench_gemm_layers.py --tile-order 16,18,20 >> $LOG 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 70
tail -c 1500 gpurun_out/r2c10_bench_n2.json; echo; tail -c 2500 gpurun_out/r2_config5_n2.json
