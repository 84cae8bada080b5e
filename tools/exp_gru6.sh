#!/bin/bash
# clusters of four (multicast weight halves) vs plain pairs: parity, then bench both
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 value %.3e  e2e %.3e  kernel %.3f ms  frac %.3f  clocks %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['frac'], d['clocks']))"; }
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_pipe.py tests/test_gpu_sweeps.py -x -q -k "gru or GRU or rnn" 2>&1 | tail -5
for q in 1 0 1 0; do
NPD_GRU_QUAD=$q timeout 300 python bench.py --workload gru64 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | line quad=$q
done
timeout 200 python tools/stress_gru.py 2>&1 | tail -3
