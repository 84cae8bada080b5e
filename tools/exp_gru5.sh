for b in 9472 37888; do for d in 0 1 2; do NPD_GRU_DBG=$d timeout 120 python bench.py --workload gru64 --steps 3 --warmup 3 --batch $b --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('batch $b dbg $d kernel %.3f ms frac %.3f clocks %s' % (r['kernel_ms'], r['frac'], d['clocks']))"; done; done
