#!/bin/bash
# GPU call 11 (8 GPUs, falls back to what is there): a15 over NCCL, configs[4] at N = 8 / 4, bench.py at N = 8
LOG=gpurun_out/r2_call11.log
. "$(dirname "$0")/r2_common.sh"
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG" >> $LOG
step "a15 over NCCL: train step of DynamicVFE on 2 GPUs"
t 400 python -m pytest tests/test_dist_gpu.py -q -m gpu >> $LOG 2>&1
t 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dist_vfe_train.py > gpurun_out/r2_dist_vfe_train.json 2>> $LOG
step "configs[4] at N=$NG"
t 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 scripts/bench_config5.py > gpurun_out/r2_config5_n$NG.json 2>> $LOG
if [ $NG -ge 8 ]; then
  step "configs[4] at N=4"
  t 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 scripts/bench_config5.py > gpurun_out/r2_config5_n4.json 2>> $LOG
fi
step "bench.py at N=$NG"
t 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --steps 40 --warmup 5 > gpurun_out/r2c11_bench_n$NG.json 2>> $LOG
step done
grep -v "^E   \|^$\|^\[bench\|OMP_NUM\|^\*\*\*" $LOG | tail -n 40
for f in gpurun_out/r2_config5_n*.json gpurun_out/r2c11_bench_n*.json gpurun_out/r2_dist_vfe_train.json; do [ -s "$f" ] && { echo $f; tail -n1 $f | cut -c1-400; }; done
