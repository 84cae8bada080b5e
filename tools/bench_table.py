"""Markdown table of a bench.py JSON line (top level + roofline.other_workloads).  usage: python tools/bench_table.py bench.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
rows = [(d["config"]["workload"], d)] + list(d.get("roofline", {}).get("other_workloads", {}).items())
print("| workload | what | codewords/s (HBM-resident) | end to end (host tensors) | ms / step | roofline | parity in the bench | CPU arm |")
print("|---|---|---|---|---|---|---|---|")
for name, r in rows:
    rf = r.get("roofline", {})
    roof = "%.2f of %s peak" % (rf.get("frac", 0), rf.get("bound"))
    sm = rf.get("smem_roofline")
    if sm:
        roof += " (%.2f of the SMEM bound)" % sm["frac"]
    pc = r.get("parity_checked")
    par = "-" if not pc else ("ok" if pc.get("ok") else ("reported" if pc.get("gated") is False else "FAILED"))
    if pc and "worst_err_over_tol" in pc:
        par += " (worst err/tol %.2f)" % pc["worst_err_over_tol"]
    cb = r.get("cpu_baseline")
    cpu = "-" if not cb else "%.3g (%s, %d cores)" % (cb["value"], cb["kind"], cb["cores"])
    e2e = r["e2e"]["value"]
    print("| `%s` | %s | %.3g | %s | %.3f | %s | %s | %s |" % (name, r["config"]["desc"].split(":")[0][:90], r["value"],
                                                             "%.3g" % e2e if e2e else "-", r["ms_per_step"], roof, par, cpu))
