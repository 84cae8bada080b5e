# one plain run, then the launch list and one full capture of the SC kernel (B200_PROFILING.md recipe)
set -e
CMD="python bench.py --workload sc1024 --steps 2 --warmup 3 --batch 32768 --no-cpu-baseline"
$CMD > gpurun_out/plain_sc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_sc.csv $CMD > gpurun_out/ncu_launches_sc.log 2>&1
$CMD > gpurun_out/plain_sc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sc_group_kernel -s 3 -c 1 -f -o gpurun_out/prof_sc $CMD > gpurun_out/ncu_full_sc.log 2>&1
tail -3 gpurun_out/ncu_full_sc.log
