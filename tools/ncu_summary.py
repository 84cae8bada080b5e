"""Key metrics of an .ncu-rep (first profiled launch) as text + traffic per launch.
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
print("report:", rep)
print("kernel:", d.get("Kernel Name", ("", "?"))[1])
for k in KEYS:
    if k in d:
        print("%-86s %-10s %s" % (k, d[k][0], d[k][1]))


def num(k):
    u, v = d.get(k, ("", "0"))
    f = float(v.replace(",", "") or 0)
    return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(u, 1)


print("dram traffic per launch (read+write) bytes: %d" % (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")))

# more than one captured launch (e.g. the N = 4096 split path = 4 x top kernel + 4 x sub-block decode + output kernel):
# per-launch duration / DRAM bytes and the totals
if len(rows) > 3:
    def val(row, key):
        i = hdr.index(key)
        u, v = units[i], row[i]
        f = float(v.replace(",", "") or 0)
        return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1e3, "us": 1.0, "ns": 1e-3,
                    "s": 1e6}.get(u, 1)
    tot_t = tot_b = 0.0
    print("all %d captured launches:" % (len(rows) - 2))
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        t = val(r, "gpu__time_duration.sum")
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        tot_t += t
        tot_b += b
        print("  %-70s %10.1f us %14d B" % (r[hdr.index("Kernel Name")][:70], t, b))
    print("total: %.1f us, dram traffic (read+write) %d B" % (tot_t, tot_b))
