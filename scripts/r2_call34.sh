#!/bin/bash
# GPU call 34: whole suite with the tightened bounds, measured parity values printed (pytest -s), smoke
LOG=gpurun_out/r2_call34.log
. "$(dirname "$0")/r2_common.sh"
step "default GPU suite (-s: parity lines)"
t 600 python -m pytest tests -q -m gpu -s > gpurun_out/r2c34_suite.txt 2>&1
grep -a "parity\|passed\|failed\|rel l2\|rel-L2\|bf16 grid\|i2s B=" gpurun_out/r2c34_suite.txt >> $LOG
step "smoke()"
t 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
step done
cat $LOG | tail -n 70
exit 0
