run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f ber %.5f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['ber']))" "$ENVTAG $*"; }
for g in 2 1; do
export NPD_SC_GTOP=$g
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sc_" 2>&1 | tail -1
ENVTAG=gl$g run --workload sc4096 --steps 3
ENVTAG=gl$g run --workload sc2048 --steps 3
done
NPD_SC_GTOP=0 ENVTAG=gl0 run --workload sc2048 --steps 3
