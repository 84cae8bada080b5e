set -e
CMD="python bench.py --workload conv64 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_conv.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_conv.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_stack_kernel -s 2 -c 1 -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full_conv.log 2>&1
tail -1 gpurun_out/ncu_full_conv.log
