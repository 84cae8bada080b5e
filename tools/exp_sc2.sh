run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f ber %.5f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['ber']))" "$ENVTAG $*"; }
for v in 1 0 1 0; do NPD_SC_VECOUT=$v ENVTAG=vec$v run --workload sc1024 --steps 10; done
for v in 1 0; do NPD_SC_VECOUT=$v ENVTAG=vec$v run --workload sc256 --steps 10; done
