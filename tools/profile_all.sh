# Round profile: launch lists (gpu__time_duration) and one full capture per hot kernel (B200_PROFILING.md).
set -e
G="python bench.py --workload gru64 --steps 2 --warmup 3 --batch 9472 --no-cpu-baseline"
S="python bench.py --workload sc1024 --steps 2 --warmup 3 --batch 32768 --no-cpu-baseline"
$G > gpurun_out/plain_gru.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_gru.csv $G > /dev/null 2>&1
$S > gpurun_out/plain_sc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_sc.csv $S > /dev/null 2>&1
$G > gpurun_out/plain_gru2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gru_decode_kernel -s 3 -c 1 -f -o gpurun_out/prof_gru $G > gpurun_out/ncu_full_gru.log 2>&1
$S > gpurun_out/plain_sc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sc_lane_kernel -s 3 -c 1 -f -o gpurun_out/prof_sc $S > gpurun_out/ncu_full_sc.log 2>&1
echo profiled
