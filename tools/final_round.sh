# full GPU suite + smoke + default bench (+ reference arm) ; outputs under gpurun_out/
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 3000 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; tail -c 800 gpurun_out/bench_reference.json
for w in gru32 sc256 sc4096 scl64 enc1024; do timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-8s value %.3e e2e %s kernel %.3f ms frac %.3f' % ('$w', d['value'], d['e2e']['value'], r['kernel_ms'], r['frac']))"; done
