run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f ber %.5f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['ber']))" "$ENVTAG $*"; }
ENVTAG=new run --workload sc1024 --steps 5
ENVTAG=new run --workload sc4096 --steps 3
ENVTAG=new run --workload sc256 --steps 5
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sc_ or sc" 2>&1 | tail -3
