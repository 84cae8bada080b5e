run() { python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s %.3e cw/s kern %.3f ms frac %.4f' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))" "$ENVTAG $*"; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_pipe.py tests/test_gpu_sweeps.py -x -q -k "sc or sweep" 2>&1 | tail -1
ENVTAG=default run --workload sc4096 --steps 5
ENVTAG=default run --workload sc2048 --steps 5
ENVTAG=default run --workload sc1024 --steps 10
timeout 300 python -m neural_polar_decoder_b200.mc_sweep --N 4096 --K 2048 --snr 2 --frames 3e7 2>&1 | tail -1 | cut -c1-250
