#!/bin/bash
# GPU call 38: the final tree (sparse-encoder train mode added): whole suite, smoke(), per-layer table, bench x2
LOG=gpurun_out/r2_call38.log
. "$(dirname "$0")/r2_common.sh"
step "default GPU suite"
t 600 python -m pytest tests -q -m gpu >> $LOG 2>&1
step "smoke()"
t 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
step "GEMM per layer"
t 200 python scripts/bench_gemm_layers.py >> $LOG 2>&1
step "bench x2"
for i in 1 2; do
  t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c38_bench_$i.json 2>> $LOG; echo "rc=$?" >> $LOG
done
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 40
for f in gpurun_out/r2c38_bench_*.json; do [ -s "$f" ] && { echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"; }; done
exit 0
