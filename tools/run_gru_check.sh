#!/bin/bash
# GRU kernel check: parity tests (with a hang guard), bench line, optional trace
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_pipe.py tests/test_gpu_sweeps.py -x -q -k "gru or GRU or rnn" 2>&1 | tail -8
timeout 300 python bench.py --workload gru64 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('gru64 value %.3e  e2e %.3e  kernel %.3f ms  frac %.3f  clocks %s' % (d['value'], d['e2e']['value'], r['kernel_ms'], r['frac'], d['clocks']))"
NPD_GRU_TRACE=gpurun_out/gru_trace.txt timeout 120 python bench.py --workload gru64 --steps 1 --warmup 3 --batch 9472 --no-cpu-baseline > /dev/null 2>&1
python tools/gru_trace.py gpurun_out/gru_trace.txt 10 1 | tail -45
