# Round profile (B200_PROFILING.md recipe): for each hot kernel one plain run, the launch list
# (gpu__time_duration) and one `ncu --set full` capture at the bench's own per-launch batch.
set -e
run() {  # name, kernel regex, bench args
  local name=$1 regex=$2; shift 2
  local CMD="python bench.py $* --no-cpu-baseline"
  $CMD > gpurun_out/plain_$name.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$name.csv $CMD > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_full_$name.log 2>&1
  tail -1 gpurun_out/ncu_full_$name.log
}
run gru gru_decode_kernel --workload gru64 --steps 1 --warmup 3
run scq sc_quad_kernel --workload sc1024 --steps 1 --warmup 3
run conv conv_stack_kernel --workload conv64 --steps 1 --warmup 3
run convfc conv_fc_kernel --workload conv64 --steps 1 --warmup 3
run scl scl_kernel --workload scl64 --steps 1 --warmup 3
echo profiled
