for d in 0 3; do NPD_GRU_DBG=$d python bench.py --workload gru64 --steps 1 --warmup 3 --batch 9472 --no-cpu-baseline 2>/dev/null | grep "GRU MMA" | tail -1; done
