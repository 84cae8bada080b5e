run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))" "$ENVTAG $*"; }
ENVTAG=default run --workload sc1024 --steps 5
ENVTAG=default run --workload sc4096 --steps 3
ENVTAG=default run --workload sc256 --steps 5
S="python bench.py --workload sc1024 --steps 2 --warmup 3 --no-cpu-baseline"
$S > /dev/null 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:sc_lane_kernel -s 3 -c 1 $S 2>&1 | grep -E "dram__|gpu__time"
