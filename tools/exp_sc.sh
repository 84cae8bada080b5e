run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f ber %.5f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['ber']))" "$ENVTAG $*"; }
ENVTAG=quad run --workload sc1024 --steps 5
ENVTAG=quad run --workload sc4096 --steps 3
ENVTAG=quad run --workload sc256 --steps 5
export NPD_SC_IMPL_LANE=1
ENVTAG=lane run --workload sc1024 --steps 5
ENVTAG=lane run --workload sc256 --steps 5
