"""Mint a short CPU-trained convNet checkpoint with the LIVE reference (TEST INFRASTRUCTURE ONLY).

Runs the reference's own script /root/reference/run_models.py as __main__ (through oracle/ref_shim.py) for a short
c2n curriculum K = 1..22 in the style of run_conv_c2n.sh / run_alt.sh (every stage warm-started with
--load_previous --model_iters from the previous one), then runs the script once more with --test so that the
reference's own testXformer (run_models.py:297-371, called at 1370) prints the BER/BLER curves of the convNet, of
SC and of SC-list (L = 4) for the final checkpoint.  Outputs (committed):

  tests/golden/conv_N64_K22_E128.pt    the reference-format checkpoint {'xformer', 'step', 'args'} exactly as
                                        run_models.py:980-983 saves it
  tests/golden/conv_N64_K22_E128.json  the command lines, the file-name scheme the reference used and its printed curves

Build container only (needs /root/reference):  python oracle/train_ref_conv_checkpoint.py [--steps 60 ...]
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
    """Execute the reference's run_models.py as a script with `argv`; returns (stdout, script globals)."""
    ref_shim.load("polar")  # stubs matplotlib / IPython, puts the reference on sys.path
    real_load = torch.load

    def load_compat(*a, **k):  # the checkpoint pickles an argparse.Namespace (run_models.py:980)
        k.setdefault("weights_only", False)
        return real_load(*a, **k)

    torch.load = load_compat
    old_argv = sys.argv
    sys.argv = ["run_models.py"] + argv
    torch.manual_seed(seed)
    buf = io.StringIO()

    class Tee(io.TextIOBase):
        def write(self, s):
            buf.write(s)
            sys.__stdout__.write(s)
            sys.__stdout__.flush()
            return len(s)

    g = {}
    try:
        with contextlib.redirect_stdout(Tee()):
            try:
                g = runpy.run_path(os.path.join(ref_shim.REF_DIR, "run_models.py"), run_name="__main__")
            except SystemExit:
                pass
    finally:
        sys.argv = old_argv
        torch.load = real_load
    return buf.getvalue(), g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=64)
    ap.add_argument("--K", type=int, default=22)
    ap.add_argument("--E", type=int, default=128)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--final_steps", type=int, default=300)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--test_size", type=int, default=10000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    torch.set_num_threads(a.threads)

    work = tempfile.mkdtemp(prefix="npd_ref_conv_")
    os.chdir(work)  # the reference writes ./Supervised_Xformer_decoder_Polar_Results/... relative to the cwd
    t0 = time.time()
    common = ["--model", "conv", "--N", str(a.N), "--max_len", str(a.N), "--embed_dim", str(a.E), "--lr", "1e-3",
              "--batch_size", str(a.batch), "--mult", "1", "--num_restarts", "1", "--code", "polar",
              "--rate_profile", "polar", "--curriculum", "c2n", "--id", "c2n", "--previous_id", "c2n",
              "--validation_snr", "1", "--target_K", str(a.K), "--run", "1", "--previous_N", str(a.N)]
    stages = []
    for K in range(1, a.K + 1):
        steps = a.final_steps if K == a.K else a.steps
        snr = "-6" if K <= 2 else "-5" if K <= 4 else "-4" if K <= 8 else "-3" if K <= 12 else "-2" if K <= 16 else "-1"
        argv = common + ["--K", str(K), "--dec_train_snr", snr, "--num_steps", str(steps), "--print_freq", str(10 ** 6),
                         "--model_save_per", str(10 ** 6), "--previous_K", str(max(K - 1, 1))]
        if K > 1:
            argv += ["--load_previous", "--model_iters", str(a.steps)]
        run_reference_main(argv, seed=K)
        stages.append(argv)
        print("== conv stage K=%d done, %.0f s elapsed" % (K, time.time() - t0), flush=True)

    test_argv = common + ["--K", str(a.K), "--previous_K", str(a.K - 1), "--test", "--test_size", str(a.test_size),
                          "--test_batch_size", "1000", "--test_snr_start", "-2", "--test_snr_end", "2", "--snr_points", "5"]
    out, g = run_reference_main(test_argv, seed=1000)

    def grab(var):
        v = g.get(var)
        return None if v is None else [float(x) for x in v]

    final_dir = g.get("final_save_path")
    results_dir = g.get("results_save_path")
    ckpt = os.path.join(work, final_dir, "Models", "model_final.pt")
    name = "conv_N%d_K%d_E%d" % (a.N, a.K, a.E)
    gold = os.path.join(ROOT, "tests", "golden")
    shutil.copyfile(ckpt, os.path.join(gold, name + ".pt"))
    meta = {
        "made_by": "oracle/train_ref_conv_checkpoint.py (live reference run_models.py run as __main__)",
        "N": a.N, "K": a.K, "embed_dim": a.E, "curriculum": "c2n K=1..%d, %d steps per stage, %d final" % (a.K, a.steps, a.final_steps),
        "final_stage_argv": stages[-1], "test_argv": test_argv,
        "final_save_path": final_dir, "results_save_path": results_dir,
        "test_size": a.test_size, "test_batch_size": 1000, "snr_range": grab("snr_range"),
        "bers_Xformer": grab("bers_Xformer_test"), "blers_Xformer": grab("blers_Xformer_test"),
        "bers_SC": grab("bers_SC_test"), "blers_SC": grab("blers_SC_test"),
        "bers_SCL": grab("bers_SCL_test"), "blers_SCL": grab("blers_SCL_test"),
        "torch": torch.__version__, "seconds": time.time() - t0,
    }
    with open(os.path.join(gold, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps({k: meta[k] for k in ("snr_range", "bers_Xformer", "bers_SC", "bers_SCL", "seconds")}))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
