set -e
CMD="python bench.py --workload gru64 --steps 1 --warmup 3 --batch 9472 --no-cpu-baseline"
$CMD > gpurun_out/plain_gru.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gru_decode_kernel -s 2 -c 1 -f -o gpurun_out/prof_gru $CMD > gpurun_out/ncu_full_gru.log 2>&1
tail -2 gpurun_out/ncu_full_gru.log
