set -e
CMD="python bench.py --workload enc1024 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_enc.log 2>&1
tail -1 gpurun_out/plain_enc.log
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 2 -c 1 -f -o gpurun_out/prof_enc $CMD > gpurun_out/ncu_full_enc.log 2>&1
tail -1 gpurun_out/ncu_full_enc.log
