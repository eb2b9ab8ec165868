#!/bin/bash
# GPU call 2 of round 2: new dense conv kernel, gather4 probe, GEMM experiments (256-row tiles, timing experiments),
# the never-run rows (reference fixtures, config 5, input contract, backward), then the default suite and benches.
LOG=gpurun_out/r2_call2.log
. "$(dirname "$0")/r2_common.sh"
step "TMA-halo dense conv"
t 300 python -m pytest tests/test_conv2d_own.py -x -q -m gpu >> $LOG 2>&1
step "gather4 probe"
t 300 python scripts/probe_gather4.py >> $LOG 2>&1
step "GEMM experiments per layer"
t 300 python scripts/bench_gemm_layers.py --tile-order none --opt gemm_mt=2,gemm_debug=1,gemm_debug=2,gemm_debug=4 >> $LOG 2>&1
export ISF_EXPERIMENTAL=1
step "rows: reference fixtures"
t 400 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu >> $LOG 2>&1
step "rows: head / detector with own attention"
ISF_HEAD_ATTN=own t 300 python -m pytest tests/test_reference_fixtures_gpu.py -q -m gpu -k "head or detector" >> $LOG 2>&1
step "rows: config5, input contract, backward, properties"
t 300 python -m pytest tests/test_config5_gpu.py -q -m gpu >> $LOG 2>&1
t 200 python -m pytest tests/test_input_contract.py -q -m gpu >> $LOG 2>&1
t 300 python -m pytest tests/test_backward_ops.py -q -m gpu >> $LOG 2>&1
t 300 python -m pytest tests/test_properties.py -q -m gpu >> $LOG 2>&1
unset ISF_EXPERIMENTAL
step "default GPU suite"
t 600 python -m pytest tests -q -m gpu -x >> $LOG 2>&1
step "bench: default (tile order 31, row kernel, TMA conv), cuDNN convs, gather convs"
t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_default.json 2>> $LOG
ISF_CONV=cudnn t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_cudnn.json 2>> $LOG
ISF_CONV=gather t 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench_gather.json 2>> $LOG
step "rows at full size"
t 200 python scripts/bench_input_pipeline.py > gpurun_out/r2_input_pipeline.json 2>> $LOG
t 300 python scripts/bench_config5.py > gpurun_out/r2_config5.json 2>> $LOG
t 300 python scripts/bench_full_forward.py > gpurun_out/r2_full_forward.json 2>> $LOG
t 300 python scripts/bench_full_forward.py --raw > gpurun_out/r2_full_forward_raw.json 2>> $LOG
ISF_HEAD_ATTN=own ISF_SWIN_ATTN=own t 300 python scripts/bench_full_forward.py > gpurun_out/r2_full_forward_own_attn.json 2>> $LOG
step done
tail -n 150 $LOG
for f in gpurun_out/r2c2_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
for f in gpurun_out/r2_input_pipeline.json gpurun_out/r2_config5.json gpurun_out/r2_full_forward*.json; do [ -s "$f" ] && { echo $f; tail -c 1200 $f; echo; }; done
