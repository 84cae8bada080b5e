"""SASS mnemonic counts per kernel of the shipped libnpd.so (evidence that the tensor-core kernels are tcgen05 / TMEM / TMA
kernels, and where local-memory spills sit).   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neural_polar_decoder_b200", "libnpd.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "STL", "LDL", "SHFL"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        op = m.group(1)
        per[name]["instr"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and ".2CTA" in op and op.startswith("UTCHMMA")):
                per[name][k] += 1
                break
print("SASS mnemonic counts per kernel of the shipped neural_polar_decoder_b200/libnpd.so (cuobjdump -sass, sm_100a): tcgen05 MMAs "
      "(UTCHMMA, .2CTA = cta_group::2), TMEM loads (LDTM), bulk copies (UBLKCP) / tensor copies (UTMALDG loads, UTMASTG stores), "
      "mbarrier ops (SYNCS), MUFU, local-memory spills (STL / LDL)\n")
for name, c in per.items():
    rest = ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])
    print("%-112s instr %6d | %s" % (name[:112], c["instr"], rest))
