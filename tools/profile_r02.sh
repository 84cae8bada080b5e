# Round-2 profile (B200_PROFILING.md recipe).  Every ncu pass runs only after the same command exited 0 without ncu.
#  1. launch list (gpu__time_duration.sum) of the DEFAULT bench command;
#  2. one `ncu --set full` capture per hot kernel at the bench's own per-launch batch, summarised on the box;
#  3. compute-sanitizer memcheck + racecheck over tools/sanitize_gru.py.
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_plain_default.json 2> $O/r02_plain_default.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_default_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r02_ncu_default.log 2>&1
full() {  # name, kernel regex, skip, bench args...   (COUNT = launches to capture, default 1)
  local name=$1 regex=$2 skip=$3; shift 3
  local CMD="python bench.py $* --no-cpu-baseline --no-parity --steps 1 --warmup 3"
  $CMD > $O/r02_plain_$name.log 2>&1 || { echo "plain run of $name failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c ${COUNT:-1} -f -o $O/prof_$name $CMD > $O/r02_ncu_full_$name.log 2>&1
  python tools/ncu_summary.py $O/prof_$name.ncu-rep > $O/r02_${name}_ncu_summary.txt 2>&1
  ncu -i $O/prof_$name.ncu-rep --page source --csv > $O/src_$name.csv 2>/dev/null && python tools/ncu_hot.py $O/src_$name.csv 30 > $O/r02_${name}_hot.txt 2>&1 || true
  rm -f $O/src_$name.csv
  ls -la $O/prof_$name.ncu-rep
}
full gru64 gru_decode_kernel3 2 --workload gru64
full sc1024 sc_quad_kernel 2 --workload sc1024
full mc1024 sc_quad_kernel 6 --workload mc1024
COUNT=9 full sc4096 'split_top_kernel|sc_quad_kernel|split_out_kernel' 27 --workload sc4096
full conv64_stack conv_stack_kernel 2 --workload conv64
full conv64_fc conv_fc_kernel 2 --workload conv64
full enc1024 encode_kernel 2 --workload enc1024
rm -f $O/prof_*.ncu-rep
# compute-sanitizer is closed on this pool (profiles/r02_sanitizer_closed_on_pool.log); the small-case driver still runs plain
python tools/sanitize_gru.py > $O/r02_small_cases.log 2>&1; echo "small cases exit $?" >> $O/r02_small_cases.log; tail -3 $O/r02_small_cases.log
echo profiled
