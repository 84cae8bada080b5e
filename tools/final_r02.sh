# Final round-2 run on one B200: full GPU test suite, smoke, default bench, configs 1 / 2 at 1e6 frames per SNR point, and a
# fresh ncu capture of the GRU pair kernel (byte residual).
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_gputest_final.log
python __graft_entry__.py smoke > $O/r02_smoke_final.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
python tools/run_config.py --test_size 1000000 --out $O 2>&1 | tail -4
CMD="python bench.py --workload gru64 --no-cpu-baseline --no-parity --steps 1 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:gru_decode_kernel3 -s 2 -c 1 -f -o $O/prof_gru64 $CMD > $O/r02_ncu_full_gru64.log 2>&1
python tools/ncu_summary.py $O/prof_gru64.ncu-rep > $O/r02_gru64_ncu_summary.txt 2>&1
ncu -i $O/prof_gru64.ncu-rep --page source --csv > $O/src_gru64.csv 2>/dev/null && python tools/ncu_hot.py $O/src_gru64.csv 30 > $O/r02_gru64_hot.txt 2>&1
rm -f $O/src_gru64.csv $O/prof_gru64.ncu-rep
echo done
