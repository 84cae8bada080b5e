"""Extract the N<=256 reliability order that the reference embeds in get_code (rnn_all.py:1046,
repeated at 1086/1105/1124/1141/1159/1176 and polar.py:1173) into a data file of the product package.

This is *data* (a code-construction table "computed for SNR = 0", rnn_all.py:1036), not code: the
drop-in get_code() must produce the same info sets as the reference, so the table has
to be identical.  Run in the build container only (needs /root/reference).
"""
import json
import os
import re
import sys

REF = os.environ.get("NPD_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "neural_polar_decoder_b200",
                   "data", "polar_rs256.json")


def main():
    src = open(os.path.join(REF, "rnn_all.py")).read().split("\n")
    tables = set()
    for line in src:
        m = re.search(r"rs = np\.array\(\[(256 ,.*)\]\) - 1", line)
        if m:
            tables.add(tuple(int(t) - 1 for t in m.group(1).replace(" ", "").split(",") if t))
    assert len(tables) == 1, len(tables)
    rs = list(tables.pop())
    assert sorted(rs) == list(range(256))
    with open(OUT, "w") as f:
        json.dump({"source": "reference rnn_all.py:1046 (value-1), most reliable first",
                   "rs": rs}, f)
    print("wrote", OUT, len(rs))


if __name__ == "__main__":
    sys.exit(main())
