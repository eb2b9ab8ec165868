#!/bin/bash
# GPU call 5: fused instance decoder (a11) bring-up + whole suite + bench A/B
LOG=gpurun_out/r2_call5.log
. "$(dirname "$0")/r2_common.sh"
step "fused decoder"
t 300 python -m pytest tests/test_fusion_gpu.py -x -q -s -m gpu -k "fused_instance_decoder" >> $LOG 2>&1
step "GEMM per layer (pipelined epilogue)"
t 300 python scripts/bench_gemm_layers.py >> $LOG 2>&1
step "default GPU suite"
t 900 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "bench: default (fused decoder), module-path decoder"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench_default.json 2>> $LOG
ISF_DECODER=torch t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench_torchdec.json 2>> $LOG
step "launch list of the IGF part (decoder kernel time)"
t 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ins_context|msda|i2s|channel_attention|heatmap|conv3x3' -c 60 --csv --log-file gpurun_out/launches_igf_r02.csv \
    python bench.py --eager-only --steps 1 --warmup 1 --frames 2 >> $LOG 2>&1
step done
grep -v "^E   \|^$" $LOG | tail -n 80
for f in gpurun_out/r2c5_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
