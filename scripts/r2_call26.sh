#!/bin/bash
# GPU call 26: bf16 tensor-core path of the HSF ops (configs[2]) + whole suite + bench stress with the shipped GEMM configuration
LOG=gpurun_out/r2_call26.log
. "$(dirname "$0")/r2_common.sh"
step "bf16 tests"
t 300 python -m pytest tests/test_bf16_gpu.py -q -m gpu -s >> $LOG 2>&1
step "HSF-only bench: f16, bf16, 2 levels"
t 200 python scripts/bench_hsf.py > gpurun_out/hsf_only_f16_r02b.json 2>> $LOG
t 200 python scripts/bench_hsf.py --dtype bf16 > gpurun_out/hsf_only_bf16_r02.json 2>> $LOG
t 200 python scripts/bench_hsf.py --dtype bf16 --levels 2 > gpurun_out/hsf_only_bf16_2levels_r02.json 2>> $LOG
step "default GPU suite"
t 600 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench x5 (shipped configuration)"
for i in 1 2 3 4 5; do
  t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c26_bench_$i.json 2>> $LOG; echo "rc=$?" >> $LOG
done
step done
grep -v "^E   \|^$\|^\[bench\|^  File\|^    " $LOG | tail -n 40
for f in gpurun_out/hsf_only_*r02*.json; do echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['dtype'], d['value'], d['ms_per_step'], d['stage_ms_eager'])"; done
for f in gpurun_out/r2c26_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
