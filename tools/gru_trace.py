"""Pretty-print a NPD_GRU_TRACE dump (clock64 stamps of CTA 0; see gru_decode.cu trace_ev).
usage: python tools/gru_trace.py trace.txt [first_step] [n_steps]"""
import sys
import numpy as np

t = np.loadtxt(sys.argv[1], dtype=np.int64)
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 2
base = t[s0, 0]
names = {}
for l in range(2):
    for j in range(4):
        names[(l * 4 + j) * 2] = "MMA  L%dj%d begin (slot free)" % (l, j)
        names[(l * 4 + j) * 2 + 1] = "MMA  L%dj%d issued+commit" % (l, j)
        names[20 + (l * 4 + j) * 2] = "  EPI L%dj%d tmem_full seen" % (l, j)
        names[21 + (l * 4 + j) * 2] = "  EPI L%dj%d math done" % (l, j)
names[16] = "MMA  h_ready0 seen"
names[18] = "MMA  waits h_ready0"
names[17] = "MMA  h_ready1 seen"
names[36] = "  EPI step end (feedback published)"
for s in range(s0, s0 + ns):
    ev = sorted((int(t[s, k]) - base, names[k]) for k in names if t[s, k])
    print("---- step %d" % s)
    for c, n in ev:
        print("%8d  %s" % (c, n))
if t[0, 37] and t[0, 38]:
    print("prologue (setup + hoisted input projection): %d cycles; first MMA job begins %d cycles after kernel entry" % (t[0, 38] - t[0, 37], t[0, 0] - t[0, 37]))
per_step = np.diff(t[:, 0])
print("cycles per step: median %d  min %d max %d" % (np.median(per_step[2:]), per_step[2:].min(), per_step[2:].max()))
if t.shape[1] > 40 and t[s0, 41] > 0:
    import os
    ref_slot = int(os.environ.get('TRACE_REF_SLOT', '10'))
    for s in range(s0, s0 + ns):
        tl = t[s, 40:88]
        d = np.diff(np.concatenate([[t[s, ref_slot]], tl]))
        print("step %d: L1j1 per-tile issue-thread cycles: %s" % (s, " ".join(str(int(x)) for x in d)))
        print("   EPI L1j0 math window: %d..%d relative to L1j1 begin" % (t[s, 20 + 8] - t[s, 10], t[s, 21 + 8] - t[s, 10]))
