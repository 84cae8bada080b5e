#!/bin/bash
# host-pipe tests + e2e numbers
python -m pytest tests/test_gpu_host_pipe.py -x -q 2>&1 | tail -15
for c in 0 1024 2048 8192; do
  NPD_HOST_CHUNK=$c python bench.py --workload sc1024 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', 'value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
done
python bench.py --workload conv64 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('conv value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
python bench.py --workload gru64 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gru value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
