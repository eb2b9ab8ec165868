#!/bin/bash
# GPU call 27: instance -> scene attention on tcgen05 (i2s_tc.cu) vs the mma.sync kernel and fp32; neck tests; bench
LOG=gpurun_out/r2_call27.log
. "$(dirname "$0")/r2_common.sh"
step "i2s tcgen05 kernel"
t 200 python -m pytest tests/test_fusion_gpu.py -q -m gpu -s -k "i2s_attention_tcgen05 or instance_to_scene" >> $LOG 2>&1
step "fusion tests"
t 400 python -m pytest tests/test_fusion_gpu.py tests/test_reference_fixtures_gpu.py -q -m gpu >> $LOG 2>&1
step "bench"
t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c27_bench.json 2>> $LOG; echo "rc=$?" >> $LOG
step done
grep -v "^$\|^\[bench\|^  File\|^    " $LOG | tail -n 60
for f in gpurun_out/r2c27_bench*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
