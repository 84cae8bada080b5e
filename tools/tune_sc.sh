run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-44s %.3e cw/s kern %.3f ms frac %.4f e2e %.3e ber %.4f bler %.4f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['ber'], d['bler']))" "$ENVTAG $*"; }
run --workload sc1024 --steps 10
for S in 6 7 8 9; do export NPD_SC_SLOG=$S; ENVTAG="SLOG=$S" run --workload sc1024 --steps 5 --batch 65536; done; unset NPD_SC_SLOG
for W in 4 8 12 20; do export NPD_SC_WARPS=$W; ENVTAG="WARPS=$W" run --workload sc64 --steps 5 ; done; unset NPD_SC_WARPS
ENVTAG="" 
for wl in sc256 sc4096; do run --workload $wl --steps 5; done
for S in 6 7; do export NPD_SC_SLOG=$S; ENVTAG="SLOG=$S" run --workload sc256 --steps 5; done; unset NPD_SC_SLOG
for S in 8 9; do export NPD_SC_SLOG=$S; ENVTAG="SLOG=$S" run --workload sc4096 --steps 3; done; unset NPD_SC_SLOG
