run() { NPD_GRU_DBG=$1 NPD_GRU_STAGES=$2 python bench.py --workload gru64 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg=%s stages=%s  kern %.3f ms frac %.4f' % (sys.argv[1], sys.argv[2], d['roofline']['kernel_ms'], d['roofline']['frac']))" $1 $2; }
for st in 2 3 4 5; do run 0 $st; run 1 $st; run 3 $st; done
