# shared by the round-2 bring-up scripts: per-step limits under an overall budget, one log
set -u
mkdir -p gpurun_out
: > $LOG
step() { echo "== [$SECONDS s] $1" | tee -a $LOG; }
# every command runs under its own limit AND the script's overall budget (seconds of box time this call may use)
BUDGET=${R2_BUDGET:-1500}
t() {
  local lim=$1; shift
  local left=$((BUDGET - SECONDS))
  if [ $left -lt 20 ]; then echo "-- skipped (budget spent): $*" | tee -a $LOG; return 99; fi
  [ $lim -gt $left ] && lim=$left
  timeout $lim "$@"
}
summarise() {
  tail -n 60 $LOG
  for f in gpurun_out/r2_bench_*.json; do
    [ -s "$f" ] || continue
    echo $f
    python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['frame_latency_ms'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'])" 2>/dev/null
  done
}
