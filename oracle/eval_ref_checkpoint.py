"""Evaluate a reference-format CRISP GRU checkpoint with the LIVE reference's own TESTING block (rnn_all.py as __main__
with --test --test_load_path) and write <checkpoint>.json with the curves (TEST INFRASTRUCTURE ONLY; build container).
Used for checkpoints trained by this repo's GPU training loop (tools/gpu_curriculum.py): the reference itself then says
what BER / BLER that checkpoint has.   python oracle/eval_ref_checkpoint.py tests/golden/X.pt [--test_size 30000]"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from train_ref_checkpoint import run_reference_main  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("checkpoint")
    ap.add_argument("--test_size", type=int, default=30000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)
    path = os.path.abspath(a.checkpoint)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    c = ck["args"]
    argv = ["--code", "Polar", "--rate_profile", c.rate_profile, "--target_K", str(c.target_K), "--N", str(c.N), "--K", str(c.K),
            "--decoding_type", c.decoding_type, "--rnn_feature_size", str(c.rnn_feature_size), "--rnn_depth", str(c.rnn_depth),
            "--onehot", "--tfr_min", "1", "--tfr_max", "1", "--dec_train_snr", "0", "--lr", "0.001", "--gpu", "-1",
            "--id", "npd_eval", "--test_snr_start", "-2", "--test_snr_end", "2", "--snr_points", "5",
            "--test_size", str(a.test_size), "--test_batch_size", "10000", "--test", "--test_load_path", path]
    os.chdir(tempfile.mkdtemp(prefix="npd_ref_eval_"))
    t0 = time.time()
    out, g = run_reference_main(argv, seed=4242)

    def grab(v):
        x = g.get(v)
        return None if x is None else [float(t) for t in x]

    meta = {"made_by": "oracle/eval_ref_checkpoint.py (live reference rnn_all.py --test on a checkpoint trained by "
                       "tools/gpu_curriculum.py through this repo's GPU training loop)",
            "N": c.N, "K": c.K, "H": c.rnn_feature_size, "rate_profile": c.rate_profile,
            "final_stage_argv": [x for x in argv if x not in ("--test", "--test_load_path", path)],
            "test_size": a.test_size, "test_batch_size": 10000, "snr_range": grab("snr_range"),
            "bers_RNN": grab("bers_RNN_test"), "blers_RNN": grab("blers_RNN_test"),
            "bers_SC": grab("bers_SC_test"), "blers_SC": grab("blers_SC_test"), "torch": torch.__version__,
            "seconds": time.time() - t0}
    with open(path[:-3] + ".json", "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: meta[k] for k in ("bers_RNN", "bers_SC", "seconds")}))


if __name__ == "__main__":
    main()
