run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))" "$ENVTAG $*"; }
for wpb in 4 2 1; do NPD_SC_WARPS=12 NPD_SC_WPB=$wpb ENVTAG=12warps_wpb$wpb run --workload sc1024 --steps 5; done
