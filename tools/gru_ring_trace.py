"""Ring round-trip breakdown of the CTA-pair GRU kernel from an NPD_GRU_TRACE dump (globaltimer stamps, ns).
usage: python tools/gru_ring_trace.py trace.txt [step]"""
import sys
import numpy as np
t = np.loadtxt(sys.argv[1], dtype=np.int64)
s = int(sys.argv[2]) if len(sys.argv) > 2 else 10
seen, issued, p0, p1, relay = t[s, 100:148], t[s, 150:198], t[s, 200:248], t[s, 250:298], t[s, 300:348]
base = seen[0]
print("tile  lead:seen issued | prod0:copy-out prod1:copy-out | relay:peer-landed   (ns since tile 0 seen)")
for i in range(0, 30):
    print("%3d   %8d %8d | %8d %8d | %8d" % (i, seen[i] - base, issued[i] - base, p0[i] - base, p1[i] - base, relay[i] - base))
d = np.diff(seen[6:46])
print("steady state: %.0f ns per tile" % d.mean())
print("copy-out -> landed at peer (fetch latency): mean %.0f ns" % (relay[6:40] - p1[6:40]).mean())
print("peer landed -> leader sees both halves: mean %.0f ns" % (seen[6:40] - relay[6:40]).mean())
print("leader issued tile t -> producers send tile t+6: rank0 %.0f ns, rank1 %.0f ns" % ((p0[12:40] - issued[6:34]).mean(), (p1[12:40] - issued[6:34]).mean()))
