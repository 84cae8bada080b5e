"""Mint a short CPU-trained CRISP GRU checkpoint with the LIVE reference (TEST INFRASTRUCTURE ONLY).

Runs the reference's own script /root/reference/rnn_all.py as __main__ (through oracle/ref_shim.py, which
stubs matplotlib / IPython) for a small progressive curriculum K = K0..K in the style of run_crisp.sh
(each stage warm-started with --load_path from the previous one), then lets the script's own TESTING
block (rnn_all.py:1745-1905: torch.load(model_final.pt) -> polar_RNN_full_test) print the reference's
BER/BLER curve of the GRU decoder and of SC.  Outputs (committed):

  tests/golden/crisp_gru_N{N}_K{K}_H{H}.pt    the reference-format checkpoint {'net', 'step', 'args'} exactly as
                                          rnn_all.py:1471-1479 saves it (a pickled argparse.Namespace inside)
  tests/golden/crisp_gru_N{N}_K{K}_H{H}.json  the command line, the reference's printed curves and the test size

Build container only (needs /root/reference):  python oracle/train_ref_checkpoint.py [--N 32 --K 16 --H 256]
"""
import argparse
import contextlib
import io
import json
import os
import runpy
import shutil
import sys
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402


def run_reference_main(argv, seed):
    """Execute the reference's rnn_all.py as a script with `argv`; returns (stdout, script globals)."""
    polar = ref_shim.load("polar")
    # rnn_all.py:847 calls channel() with four extra positionals that polar.py:201 does not take
    # (SURVEY.md App. C); the wrapper drops them -- the arithmetic is untouched.
    if not getattr(polar.PolarCode.channel, "_npd_wrapped", False):
        orig = polar.PolarCode.channel

        def channel(self, code, snr, *_extra):
            return orig(self, code, snr)

        channel._npd_wrapped = True
        polar.PolarCode.channel = channel
    # reference checkpoints pickle an argparse.Namespace (rnn_all.py:1474); torch >= 2.6 needs weights_only=False
    real_load = torch.load

    def load_compat(*a, **k):
        k.setdefault("weights_only", False)
        return real_load(*a, **k)

    torch.load = load_compat
    old_argv = sys.argv
    sys.argv = ["rnn_all.py"] + argv
    torch.manual_seed(seed)
    buf = io.StringIO()

    class Tee(io.TextIOBase):
        def write(self, s):
            buf.write(s)
            sys.__stdout__.write(s)
            sys.__stdout__.flush()
            return len(s)

    try:
        with contextlib.redirect_stdout(Tee()):
            g = runpy.run_path(os.path.join(ref_shim.REF_DIR, "rnn_all.py"), run_name="__main__")
    finally:
        sys.argv = old_argv
        torch.load = real_load
    return buf.getvalue(), g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=32)
    ap.add_argument("--K", type=int, default=16)
    ap.add_argument("--K0", type=int, default=6)
    ap.add_argument("--H", type=int, default=256)
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--final_steps", type=int, default=600)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--test_size", type=int, default=100000)
    ap.add_argument("--rate_profile", default="polar")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--snr_start", type=float, default=-2.0)
    ap.add_argument("--snr_end", type=float, default=4.0)
    ap.add_argument("--snr_points", type=int, default=7)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)

    work = tempfile.mkdtemp(prefix="npd_ref_train_")
    os.chdir(work)  # the reference writes ./Supervised_RNN_Polar_Results/... relative to the cwd
    prev = None
    log = []
    t0 = time.time()
    for K in range(a.K0, a.K + 1):
        last = K == a.K
        save = os.path.join(work, "stage_K%d.pt" % K)
        argv = ["--code", "Polar", "--rate_profile", a.rate_profile, "--target_K", str(a.K), "--N", str(a.N),
                "--K", str(K), "--decoding_type", "y_input", "--rnn_feature_size", str(a.H), "--rnn_depth", "2",
                "--onehot", "--num_steps", str(a.final_steps if last else a.steps), "--batch_size", str(a.batch),
                "--tfr_min", "1", "--tfr_max", "1", "--dec_train_snr", "0", "--lr", "0.001",
                "--print_freq", "1000000", "--model_save_per", "1000000", "--gpu", "-1", "--fresh",
                "--id", "npd_stage%d" % K, "--save_path", save,
                "--test_snr_start", str(a.snr_start), "--test_snr_end", str(a.snr_end),
                "--snr_points", str(a.snr_points),
                "--test_size", str(a.test_size if last else 1000),
                "--test_batch_size", str(10000 if last else 1000)]
        if prev:
            argv += ["--load_path", prev]
        out, g = run_reference_main(argv, seed=K)
        log.append((K, argv, out, g if last else None))
        prev = save
        print("== stage K=%d done, %.0f s elapsed" % (K, time.time() - t0), flush=True)

    g = log[-1][3]

    def grab(var):
        v = g.get(var)
        return None if v is None else [float(x) for x in v]

    name = "crisp_gru_N%d_K%d_H%d" % (a.N, a.K, a.H)
    gold = os.path.join(ROOT, "tests", "golden")
    shutil.copyfile(prev, os.path.join(gold, name + ".pt"))
    meta = {
        "made_by": "oracle/train_ref_checkpoint.py (live reference rnn_all.py run as __main__)",
        "N": a.N, "K": a.K, "H": a.H, "rate_profile": a.rate_profile,
        "final_stage_argv": log[-1][1], "curriculum": [e[0] for e in log],
        "test_size": a.test_size, "test_batch_size": 10000,
        "snr_range": grab("snr_range"),
        "bers_RNN": grab("bers_RNN_test"), "blers_RNN": grab("blers_RNN_test"),
        "bers_SC": grab("bers_SC_test"), "blers_SC": grab("blers_SC_test"),
        "torch": torch.__version__, "seconds": time.time() - t0,
    }
    with open(os.path.join(gold, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: meta[k] for k in ("snr_range", "bers_RNN", "bers_SC", "seconds")}))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
