# Closing verification of round 2 on one B200: full GPU test suite, smoke, default bench line.
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_gputest_final.log
python __graft_entry__.py smoke > $O/r02_smoke_final.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_port_gpubox.json 2> $O/r02_bench_reference_port_gpubox.err; echo "reference arm rc=$?"
echo done
