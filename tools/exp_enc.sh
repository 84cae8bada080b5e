timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sweeps.py -x -q -k "gen or awgn or encode or sweep or channel or curve" 2>&1 | tail -5
for i in 1 2; do timeout 200 python bench.py --workload enc1024 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('enc1024 value %.3e kernel %.3f ms frac %.3f' % (d['value'], r['kernel_ms'], r['frac']))"; done
