"""Timeline of one pass of the convNet stack kernel from an NPD_CONV_TRACE dump (clock64 stamps of CTA 0).
usage: python tools/conv_trace.py trace.txt"""
import sys
import numpy as np
t = np.loadtxt(sys.argv[1], dtype=np.int64)
base = t[t > 0].min()
ev = []
for i in range(t.shape[0]):
    L, g = i // 2, i % 2
    for k, nm in enumerate(("MMA  L%d g%d activations ready", "MMA  L%d g%d issued + committed", "  EPI L%d g%d accumulators seen", "  EPI L%d g%d rows written")):
        if t[i, k]:
            ev.append((int(t[i, k] - base), nm % (L + 1, g)))
for c, n in sorted(ev):
    print("%8d  %s" % (c, n))
mma = sum(int(t[i, 1] - t[i, 0]) for i in range(t.shape[0]) if t[i, 0] and t[i, 1])
epi = sum(int(t[i, 3] - t[i, 2]) for i in range(t.shape[0]) if t[i, 2] and t[i, 3])
print("pass: %d cycles; MMA issue windows %d; epilogue windows %d" % (max(c for c, _ in ev), mma, epi))
