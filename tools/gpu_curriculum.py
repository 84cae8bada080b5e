"""A run_crisp.sh-style curriculum (K = K0 .. 22, Polar(64,22), H = 512, rev_polar, teacher forcing, lr 1e-3 with the
script's StepLR) trained ON THE GPU through the drop-in CLI's training loop (train.run_train -> npd_gru_train_step), every
stage warm-started with --load_path from the previous one like the shell script, then the TESTING block on the final
checkpoint.  Evidence that the training path is stable over thousands of iterations, not only for the three parity steps.
  python tools/gpu_curriculum.py [--steps 300 --final_steps 1500 --batch 4096 --out gpurun_out/r02_gpu_curriculum.json]"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from neural_polar_decoder_b200 import cli, train  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K0", type=int, default=8)
    ap.add_argument("--K", type=int, default=22)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--final_steps", type=int, default=1500)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--test_size", type=int, default=100000)
    ap.add_argument("--train_gemm", default="fp32", choices=["fp32", "tf32", "bf16", "fp16"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_gpu_curriculum.json"))
    a = ap.parse_args()
    a.out = os.path.abspath(a.out)
    work = tempfile.mkdtemp(prefix="npd_gpu_train_")
    os.chdir(work)
    torch.manual_seed(0)
    prev, stages, t0 = None, [], time.time()
    for K in range(a.K0, a.K + 1):
        last = K == a.K
        save = os.path.join(work, "stage_K%d.pt" % K)
        argv = ("--code Polar --rate_profile rev_polar --target_K %d --N 64 --K %d --decoding_type y_input --rnn_feature_size 512 "
                "--num_steps %d --batch_size %d --rnn_depth 2 --model_save_per 1000000 --tfr_min 1 --tfr_max 1 --dec_train_snr 0 "
                "--lr 0.001 --scheduler step --lr_decay 2000 --lr_decay_gamma 0.95 --onehot --id gpu%d --print_freq 100 "
                "--save_path %s --test_snr_start -2 --test_snr_end 2 --snr_points 5 --test_size %d --test_batch_size 10000 "
                "--train_gemm %s" % (a.K, K, a.final_steps if last else a.steps, a.batch, K, save, a.test_size, a.train_gemm)).split()
        if prev:
            argv += ["--load_path", prev]
        args = cli.get_args(argv)
        ts = time.time()
        losses = train.run_train(args, out=lambda *x: None)
        torch.cuda.synchronize()
        stages.append({"K": K, "steps": args.num_steps, "seconds": time.time() - ts, "loss_first": losses[0][1],
                       "loss_last": losses[-1][1]})
        print("stage K=%d: %d steps in %.1f s, loss %.4f -> %.4f" % (K, args.num_steps, stages[-1]["seconds"], losses[0][1],
                                                                    losses[-1][1]), flush=True)
        prev = save
    res = cli.run_test(args, out=lambda *x: None)
    out = {"what": "run_crisp.sh-style curriculum trained on the GPU (%s GEMMs, npd_gru_train_step), then the TESTING block" % a.train_gemm,
           "batch": a.batch, "stages": stages, "train_seconds": sum(s["seconds"] for s in stages), "total_seconds": time.time() - t0,
           "test_size": a.test_size, "snr_range": res["snr_range"], "bers_RNN": res["bers_RNN"], "blers_RNN": res["blers_RNN"],
           "bers_SC": res["bers_SC"], "blers_SC": res["blers_SC"],
           "cpu_reference_checkpoint_bers_RNN": json.load(open(os.path.join(ROOT, "tests/golden/crisp_gru_N64_K22_H512.json")))["bers_RNN"]}
    json.dump(out, open(a.out, "w"), indent=1)
    import shutil
    shutil.copyfile(prev, a.out[:-5] + ".pt")  # the final reference-format checkpoint {'net', 'step', 'args'}
    print(json.dumps({k: out[k] for k in ("train_seconds", "bers_RNN", "bers_SC")}))


if __name__ == "__main__":
    main()
