#!/bin/bash
# GPU call 9: train-mode DynamicVFE, whole suite, smoke(), the full bench line (with CPU baseline), reference arm, 4 lanes
LOG=gpurun_out/r2_call9.log
. "$(dirname "$0")/r2_common.sh"
step "train-mode DynamicVFE"
t 300 python -m pytest tests/test_voxel_gpu.py -q -m gpu -k "train_mode" >> $LOG 2>&1
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "smoke()"
t 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
step "GEMM per layer (final defaults)"
t 300 python scripts/bench_gemm_layers.py --tile-order none >> $LOG 2>&1
step "bench: full default line, 4 lanes, reference arm"
t 600 python bench.py > gpurun_out/r2c9_bench_full.json 2>> $LOG
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --inflight 4 > gpurun_out/r2c9_bench_4lanes.json 2>> $LOG
t 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c9_bench_reference.json 2>> $LOG
step done
grep -v "^E   \|^$" $LOG | tail -n 60
for f in gpurun_out/r2c9_bench_full.json gpurun_out/r2c9_bench_4lanes.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d.get('cpu_baseline'))"; }; done
tail -c 600 gpurun_out/r2c9_bench_reference.json
