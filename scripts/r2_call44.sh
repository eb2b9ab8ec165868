#!/bin/bash
# GPU call 44: the final tree (own weight-gradient kernel added): whole GPU suite, smoke(), one bench line
LOG=gpurun_out/r2_call44.log
. "$(dirname "$0")/r2_common.sh"
step "default GPU suite"
t 400 python -m pytest tests -q -m gpu 2>&1 | tail -n 15 >> $LOG
step "smoke()"
t 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
step "bench"
t 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r2c44_bench.json 2>> $LOG; echo "rc=$?" >> $LOG
step done
grep -v "^E   \|^$\|^\[bench" $LOG | tail -n 30
python -c "import json; d=json.loads(open('gpurun_out/r2c44_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])"
exit 0
