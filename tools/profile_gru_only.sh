set -e
CMD="python bench.py --workload gru64 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_gru.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_gru.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gru_decode_kernel -s 2 -c 1 -f -o gpurun_out/prof_gru $CMD > gpurun_out/ncu_full_gru.log 2>&1
tail -1 gpurun_out/ncu_full_gru.log
