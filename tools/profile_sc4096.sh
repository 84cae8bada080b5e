set -e
export NPD_SC_GTOP=${NPD_SC_GTOP:-2}
CMD="python bench.py --workload sc4096 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_sc4096.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sc_quad_kernel -s 2 -c 1 -f -o gpurun_out/prof_sc4096 $CMD > gpurun_out/ncu_full_sc4096.log 2>&1
tail -1 gpurun_out/ncu_full_sc4096.log
