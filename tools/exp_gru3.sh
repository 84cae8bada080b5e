run() { NPD_GRU_DBG=$1 python bench.py --workload gru64 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dbg=%s kern %.3f ms frac %.4f' % (sys.argv[1], d['roofline']['kernel_ms'], d['roofline']['frac']))" $1; }
for d in 0 3; do run $d; done
