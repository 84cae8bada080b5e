# Round-2 closing run on one B200 (after the training-step work): full GPU test suite, smoke, default bench, the training
# workloads alone, launch list of one TF32 training iteration, ncu capture of the fused forward layer-step kernel.
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_gputest_final.log
python __graft_entry__.py smoke > $O/r02_smoke_final.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
for w in train64 train64tf32 train64bf16; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; echo "$w rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/r02_train64tf32_launches.csv \
    python bench.py --workload train64tf32 --steps 1 --warmup 1 --no-cpu-baseline > $O/r02_ncu_train.log 2>&1
python tools/launch_summary.py $O/r02_train64tf32_launches.csv > $O/r02_train64tf32_launches_summary.txt 2>&1
CMD="python bench.py --workload train64tf32 --no-cpu-baseline --steps 1 --warmup 1"
ncu --set full --clock-control none --import-source on -k regex:gru_fwd_tc_kernel -s 70 -c 1 -f -o $O/prof_train_tc $CMD > $O/r02_ncu_full_train_tc.log 2>&1
python tools/ncu_summary.py $O/prof_train_tc.ncu-rep > $O/r02_train_tc_ncu_summary.txt 2>&1
ncu -i $O/prof_train_tc.ncu-rep --page source --csv > $O/src_tc.csv 2>/dev/null && python tools/ncu_hot.py $O/src_tc.csv 30 > $O/r02_train_tc_hot.txt 2>&1
rm -f $O/src_tc.csv $O/prof_train_tc.ncu-rep
echo done
