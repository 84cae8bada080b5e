set -x
B="python bench.py --no-cpu-baseline --no-parity --steps 30 --warmup 5"
for w in 4 7; do for ww in 12 14 16; do echo "== sc1024 WPB=$w WARPS=$ww"; NPD_SC_WPB=$w NPD_SC_WARPS=$ww $B --workload sc1024 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4e kern_ms %.4f' % (d['value'], d['roofline']['kernel_ms']))"; done; done
echo "== sc2048 WPB"; for w in 4 7; do NPD_SC_WPB=$w $B --workload sc2048 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4e kern_ms %.4f' % (d['value'], d['roofline']['kernel_ms']))"; done
