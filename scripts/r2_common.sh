# shared by the round-2 GPU scripts: per-step limits under an overall budget, one log
set -u
mkdir -p gpurun_out
: > $LOG
step() { echo "== [$SECONDS s] $1" | tee -a $LOG; }
BUDGET=${R2_BUDGET:-1500}
t() {
  local lim=$1; shift
  local left=$((BUDGET - SECONDS))
  if [ $left -lt 20 ]; then echo "-- skipped (budget spent): $*" | tee -a $LOG; return 99; fi
  [ $lim -gt $left ] && lim=$left
  timeout $lim "$@"
}
