"""BASELINE.json configs 1 and 2 as full runs through the drop-in CLI on the GPU: the reference-trained CRISP GRU checkpoints
(tests/golden/crisp_gru_*.pt) evaluated at --test_size frames per SNR point (config 2: 10^6) next to SC, with the curves the
LIVE reference printed for the same checkpoint (100k frames, in the fixture's .json) and the two-sample z-score of every
point.  Writes one JSON per checkpoint.   python tools/run_config.py [--test_size 1000000] [--out gpurun_out]"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from neural_polar_decoder_b200 import cli  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--test_size", type=int, default=1000000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    a = ap.parse_args()
    gold = os.path.join(ROOT, "tests", "golden")
    for name in sorted(f[:-5] for f in os.listdir(gold) if f.startswith("crisp_gru_") and f.endswith(".json")):
        meta = json.load(open(os.path.join(gold, name + ".json")))
        argv = list(meta["final_stage_argv"])
        for flag in ("--load_path", "--save_path"):
            if flag in argv:
                i = argv.index(flag)
                del argv[i:i + 2]
        i = argv.index("--test_size")
        argv[i + 1] = str(a.test_size)
        argv += ["--test", "--test_load_path", os.path.join(gold, name + ".pt")]
        args = cli.get_args(argv)
        torch.manual_seed(7)
        torch.cuda.synchronize()
        t0 = time.time()
        res = cli.run_test(args, out=lambda *x: None)
        torch.cuda.synchronize()
        dt = time.time() - t0
        n1, n2 = a.test_size, meta["test_size"]
        rows = []
        for i, snr in enumerate(res["snr_range"]):
            row = {"snr_db": snr}
            for dec in ("RNN", "SC"):
                p1, p2 = res["blers_" + dec][i], meta["blers_" + dec][i]
                p = (p1 * n1 + p2 * n2) / (n1 + n2)
                se = math.sqrt(max(p * (1 - p), 1e-12) * (1.0 / n1 + 1.0 / n2))
                row.update({"bler_%s" % dec: p1, "bler_%s_reference" % dec: p2, "z_bler_%s" % dec: (p1 - p2) / se,
                            "ber_%s" % dec: res["bers_" + dec][i], "ber_%s_reference" % dec: meta["bers_" + dec][i]})
            rows.append(row)
        out = {"checkpoint": name, "argv": argv, "frames_per_snr_point": a.test_size, "reference_frames_per_snr_point": n2,
               "seconds": dt, "decoded_codewords_per_s_whole_test": 2 * a.test_size * len(rows) / dt,
               "note": "z = two-sample z-score of the block-error rates (ours on Philox noise vs the live reference on "
                       "mt19937 noise); |z| < 3.1 for every point = family-wise 95 % over 28 comparisons", "points": rows}
        path = os.path.join(a.out, "r02_config_%s.json" % name)
        json.dump(out, open(path, "w"), indent=1)
        print(name, "%.1f s" % dt, "max |z| %.2f" % max(abs(r["z_bler_RNN"]) for r in rows), "->", path)


if __name__ == "__main__":
    main()
