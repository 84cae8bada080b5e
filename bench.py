#!/usr/bin/env python
"""bench.py -- decoded codewords/sec of the Monte-Carlo decode hot path on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gru64|sc1024|mc1024|...] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic AWGN frames that is already resident in HBM
(`value`), or starts in pinned host memory and ends in host memory (`e2e`).  Prints ONE JSON line (rank 0).
The default line's top level is the headline's first half (CRISP-GRU Polar(64,22)); every other workload of
BASELINE.json's configs rides along, with the same fields, under `roofline.other_workloads` (a key the driver's
record keeps) and, for readers of the raw line, under `also`.  See DESIGN.md "Measurement" for every field.

`--impl reference` never imports the product package: it times the reference's own CPU implementation through
oracle/cpu_arm.py (the unmodified reference where /root/reference exists, the port elsewhere).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "decoded codewords/sec"
UNIT = "codewords/s"
GRU_CKPT = os.path.join(ROOT, "tests", "golden", "crisp_gru_N64_K22_H512.pt")
CONV_CKPT = os.path.join(ROOT, "tests", "golden", "conv_N64_K22_E128.pt")

WORKLOADS = {
    # name: N, K, snr_db, per-GPU batch (y = B*N*4 bytes must exceed the 126 MB L2 where the path is HBM-fed)
    "gru64": dict(kind="gru", N=64, K=22, snr=0.0, batch=37888,
                  desc="CRISP GRU(2x512, y_input, onehot) Polar(64,22), AWGN 0 dB"),
    "gru64sweep": dict(kind="grusweep", N=64, K=22, snr=0.0, batch=10000, n_snr=5,
                       desc="config 1 call pattern: polar_RNN_full_test, test_batch_size 10000 x 5 SNR points (-2..2 dB) "
                            "per step, GRU + SC + counters through the drop-in loop"),
    "gru64fast": dict(kind="gru", N=64, K=22, snr=0.0, batch=37888, precision="fast",
                      desc="the same decode with rnn_all.set_gru_precision('fast'): fp16-only recurrent state (no residual)"),
    "gru32": dict(kind="gru", N=32, K=16, snr=0.0, batch=37888,
                  desc="CRISP GRU(2x512, y_input, onehot) Polar(32,16), AWGN 0 dB, synthetic weights"),
    "sc1024": dict(kind="sc", N=1024, K=512, snr=2.0, batch=131072,
                   desc="SC Polar(1024,512), polarization-weight frozen set, AWGN 2 dB"),
    "sc256": dict(kind="sc", N=256, K=128, snr=2.0, batch=524288,
                  desc="SC Polar(256,128), reference reliability table, AWGN 2 dB"),
    "sc2048": dict(kind="sc", N=2048, K=1024, snr=2.0, batch=65536,
                   desc="SC Polar(2048,1024), polarization-weight frozen set, AWGN 2 dB"),
    "sc4096": dict(kind="sc", N=4096, K=2048, snr=2.0, batch=32768,
                   desc="SC Polar(4096,2048), polarization-weight frozen set, AWGN 2 dB"),
    "sc64": dict(kind="sc", N=64, K=22, snr=0.0, batch=2097152,
                 desc="SC Polar(64,22), reference 'polar' profile, AWGN 0 dB"),
    "pac32": dict(kind="sc", N=32, K=16, snr=2.0, batch=4194304, pac_g=53,
                  desc="PAC(32,16) SC decoder (pac_sc_decode), RM profile, g = 53, AWGN 2 dB"),
    "scl64": dict(kind="sc", N=64, K=22, snr=0.0, batch=262144, L=4,
                  desc="SC-list (L=4) Polar(64,22), reference 'polar' profile, AWGN 0 dB"),
    "enc1024": dict(kind="enc", N=1024, K=512, snr=2.0, batch=131072,
                    desc="message generation + Plotkin encoder + BPSK/AWGN channel, Polar(1024,512), 2 dB (Philox noise)"),
    # config 5: the fused generate -> encode -> noise -> decode -> count sweep (npd_mc_sc_sweep); batch = frames per step
    "mc256": dict(kind="mc", N=256, K=128, snr=2.0, batch=16 << 20, chunk=1 << 19,
                  desc="fused SC Monte-Carlo sweep Polar(256,128), 2 dB: generate + encode + AWGN + SC + count on the device"),
    "mc1024": dict(kind="mc", N=1024, K=512, snr=2.0, batch=5 << 20, chunk=1 << 17,
                   desc="fused SC Monte-Carlo sweep Polar(1024,512), 2 dB: generate + encode + AWGN + SC + count on the device"),
    "mc4096": dict(kind="mc", N=4096, K=2048, snr=2.0, batch=1 << 20, chunk=1 << 16,
                   desc="fused SC Monte-Carlo sweep Polar(4096,2048), 2 dB: generate + encode + AWGN + SC + count on the device"),
    "train64": dict(kind="train", N=64, K=22, snr=0.0, batch=4096, tf32=0,
                    desc="CRISP GRU(2x512) training iteration Polar(64,22), run_crisp.sh batch 4096, teacher-forced: forward + "
                         "BPTT + clip + AdamW on the device (fp32 GEMMs)"),
    "train64tf32": dict(kind="train", N=64, K=22, snr=0.0, batch=4096, tf32=1,
                        desc="the same training iteration with TF32 tensor-core GEMMs"),
    "train64bf16": dict(kind="train", N=64, K=22, snr=0.0, batch=4096, tf32=2,
                        desc="the same training iteration with bf16 tensor-core GEMMs (fp32 data, accumulation and optimizer)"),
    "conv64": dict(kind="conv", N=64, K=22, snr=0.0, batch=131072,
                   desc="convNet(embed_dim 128) one-shot decoder Polar(64,22), AWGN 0 dB"),
}
# what the default line carries besides its top level (gru64): name -> (steps cap, with a CPU baseline)
DEFAULT_ALSO = [("sc1024", 20, True), ("conv64", 20, True), ("gru64fast", 20, False), ("mc1024", 20, False), ("mc256", 6, False),
                ("mc4096", 10, False), ("gru64sweep", 10, False), ("sc256", 10, False), ("sc4096", 10, False),
                ("enc1024", 10, False), ("gru32", 10, False), ("pac32", 10, False), ("train64", 5, False),
                ("train64tf32", 5, False)]


def gru_weights_note(w):
    if w["kind"] in ("gru", "grusweep") and w["N"] == 64 and os.path.exists(GRU_CKPT):
        return "reference-trained checkpoint tests/golden/crisp_gru_N64_K22_H512.pt (oracle/train_ref_checkpoint.py)"
    if w["kind"] == "conv" and os.path.exists(CONV_CKPT):
        return "reference-trained checkpoint tests/golden/conv_N64_K22_E128.pt (oracle/train_ref_conv_checkpoint.py)"
    if w["kind"] in ("gru", "grusweep", "conv"):
        return "synthetic (seeded random init)"
    return None


def workload_config(name, w):
    """The `config` object -- identical in the product arm and the reference arm."""
    c = {"workload": name, "desc": w["desc"], "N": w["N"], "K": w["K"], "snr_db": w["snr"], "batch_per_gpu": w["batch"]}
    note = gru_weights_note(w)
    if note:
        c["weights"] = note
    return c


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]),
                    bf16_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")  # B200_PROFILING.md


def ncu_traffic(workload, batch):
    """DRAM bytes per launch of the workload's hot kernel, from the committed ncu capture (profiles/traffic.json),
    scaled from the capture's batch to this launch's.  -> (bytes or None, note)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f).get(workload)
    except Exception:
        t = None
    if not t:
        return None, None
    per_cw = (t["dram_read_bytes"] + t["dram_write_bytes"]) / float(t["capture_batch"])
    return per_cw * batch, "ncu --set full, %s at %d codewords per launch (%s), scaled to this launch's batch" % (
        t["kernel"], t["capture_batch"], t["source"])


def make_code(w):
    from neural_polar_decoder_b200 import PolarCode, construct
    N, K = w["N"], w["K"]
    n = int(np.log2(N))
    if w.get("pac_g"):
        from neural_polar_decoder_b200 import PAC
        return PAC(None, N, K, w["pac_g"])
    if N <= 256:
        rs = construct.reference_rs256()
        return PolarCode(n, K, None, rs=rs[rs < N])
    return PolarCode(n, K, None, F=construct.pw_frozen_set(N, K))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)  # first queries are slow: pay for them here,
            try:                                                        # before the timed region starts
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            # NVML queries briefly stall the GPU they ask about (measured: 5 ms polling added 1.0-1.4 ms to every
            # 12 ms GRU launch): a few quick samples so that short timed regions are covered, then 20 per second
            time.sleep(0.005 if len(self.samples) < 4 else 0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (oracle/cpu_arm.py: the live reference where present, the port elsewhere).  No product import here.
# ---------------------------------------------------------------------------------------------------
def _cpu_arm():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_arm
    return cpu_arm


def cpu_step_fn(w, threads):
    """-> (step(B, seed) -> (rate, B, dt), kind, how, probe batch) for the workload's decoder on the host cores."""
    ca = _cpu_arm()
    kind_w = w["kind"]
    if kind_w in ("gru", "grusweep"):
        arm = ca.make_gru_arm(w["N"], w["K"], 512, 0)
        return (lambda B, seed: ca.gru_rate(w["N"], w["K"], w["snr"], B, threads, seed, arm=arm)[:3]), arm[2], arm[3], 256
    if kind_w == "train":
        arm = ca.make_train_arm(w["N"], w["K"], 512, 0)
        return (lambda B, seed: ca.train_rate(w["N"], w["K"], w["snr"], B, threads, seed, arm=arm)[:3]), arm[2], arm[3], 128
    if kind_w == "conv":
        arm = ca.make_conv_arm(w["N"], 128, 0)
        return (lambda B, seed: ca.conv_rate(w["N"], w["snr"], B, threads, seed, arm=arm)[:3]), arm[1], arm[2], 256
    arm = ca.make_sc_arm(w["N"], w["K"], w["snr"], threads, w.get("pac_g"), int(w.get("L", 0)))
    probe = 32 if arm[2] == "reference" and w["N"] >= 1024 else max(threads, 64)
    return (lambda B, seed: ca.sc_rate(w["N"], w["K"], w["snr"], B, threads, seed, w.get("pac_g"), int(w.get("L", 0)),
                                       arm=arm)[:3]), arm[2], arm[3], probe


def cpu_baseline(w, seconds=12.0):
    """One bounded sample (about `seconds` of CPU work) of the workload's decoder on all host cores."""
    threads = os.cpu_count() or 1
    step, kind, how, probe = cpu_step_fn(w, threads)
    rate, B, dt = step(probe, 0)
    if dt < 0.5 * seconds:
        B = int(max(probe, min(w["batch"], rate * seconds)))
        rate, B, dt = step(B, 1)
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d codewords of the same workload in %.1f s: %s" % (B, dt, how)}


def reference_arm(args, name, w):
    """`--impl reference`: K timed steps of the CPU implementation, each a bounded sample of the product arm's batch."""
    threads = os.cpu_count() or 1
    total_budget = 150.0
    n_calls = max(1, args.steps + args.warmup)
    step, kind, how, probe = cpu_step_fn(w, threads)
    rate, _, dt0 = step(probe, 10 ** 6)
    per_step = min(6.0, total_budget / n_calls)
    B = int(max(probe, min(w["batch"], rate * per_step)))
    if dt0 > per_step:  # one probe-sized call already exceeds the per-step budget (live torch-op SC at large N)
        B = probe
    runs = []
    t_wall = time.perf_counter()
    for i in range(n_calls):
        r, b, dt = step(B, i)
        if i >= args.warmup:
            runs.append((b, dt))
        if time.perf_counter() - t_wall > 2 * total_budget and len(runs) >= 1:
            break  # keep the whole arm within a few minutes even on a slow host
    tot_cw, tot_t = sum(b for b, _ in runs), sum(t for _, t in runs)
    value = tot_cw / tot_t
    sample = "%d timed steps x %d codewords (of the %d-codeword batch) = %.1f s of CPU work on %d threads: %s" % (
        len(runs), B, w["batch"], tot_t, threads, how)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(runs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(name, w),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ---------------------------------------------------------------------------------------------------
# in-bench parity: >= 1000 rows of a timed step's output against the oracle (the checker, never the product path)
# ---------------------------------------------------------------------------------------------------
def parity_check(w, sample):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    kind = w["kind"]
    N, K = w["N"], w["K"]
    n = int(np.log2(N))
    rows = int(sample["y"].shape[0])
    out = {"rows": rows}
    if kind == "sc":
        info = np.asarray(sample["info"], dtype=np.int32)
        if w.get("pac_g"):
            _, ref, _ = oracle.pac_sc_decode(sample["y"], w["snr"], n, info, w["pac_g"])
            out["oracle"] = "oracle.pac_sc_decode (npd_oracle.c)"
        elif w.get("L"):
            _, ref = oracle.scl_decode(sample["y"], w["snr"], n, info, w["L"])
            out["oracle"] = "oracle.scl_decode (npd_oracle.c)"
        else:
            ref = np.concatenate(oracle.run_threaded(lambda lo, hi: oracle.sc_decode(sample["y"][lo:hi], w["snr"], n, info)[2],
                                                     rows, os.cpu_count() or 1))
            out["oracle"] = "oracle.sc_decode (npd_oracle.c)"
        out["mismatching_rows"] = int((ref != sample["decoded"]).any(axis=1).sum())
        out["criterion"] = "bit-exact decisions"
        out["ok"] = out["mismatching_rows"] == 0
        return out
    if kind == "gru":
        import torch
        from neural_polar_decoder_b200.rnn_all import gru_decode
        info = np.asarray(sample["info"])
        do, lo = oracle.gru_decode(sample["sd"], sample["y"], N, info)
        tol = 1e-2 * np.abs(lo) + 2e-3
        # (1) the timed step's free-running decisions: equal unless a near-zero logit came at or before the mismatch
        risky_before = np.cumsum(np.abs(lo) <= tol, axis=1) > 0
        mism = sample["decoded"] != do
        out["decision_mismatches"] = int(mism.sum())
        out["unexplained_mismatches"] = int((mism & ~risky_before).sum())
        # (2) logits under forced (= oracle) feedback
        _, lg = gru_decode(sample["net"], sample["loss_code"], torch.from_numpy(sample["y"]).cuda(),
                           forced=torch.from_numpy(do).cuda(), want_logits=True)
        err = np.abs(lg.cpu().numpy() - lo)
        out["max_logit_err"] = float(err.max())
        out["mean_logit_err"] = float(err.mean())
        out["worst_err_over_tol"] = float((err / tol).max())
        out["frac_beyond_tol"] = float((err > tol).mean())
        out["logit_rms"] = float(np.sqrt((lo ** 2).mean()))
        out["oracle"] = "oracle.gru_decode (fp32 torch restatement of rnn_all.py:514-547)"
        out["criterion"] = "tolerance t = 1e-2 |ref| + 2e-3 on every forced-feedback logit (rows x N entries): at most 1e-4 " \
                           "of the entries beyond t and none beyond 1.5 t (the committed fixtures are held to t itself, " \
                           "tests/test_gpu_parity.py; the maximum over 65536 entries sits at the fp16-operand floor, " \
                           "DESIGN.md 4.3e); free-running decisions equal except behind a logit within t of zero"
        out["ok"] = (out["unexplained_mismatches"] == 0 and out["frac_beyond_tol"] <= 1e-4 and
                     out["worst_err_over_tol"] <= 1.5)
        return out
    if kind == "conv":
        ref = np.asarray(oracle.conv_forward(sample["sd"], sample["y"])).reshape(rows, N)
        err = np.abs(sample["logits"] - ref)
        tol = 1e-2 * np.abs(ref) + 2e-3
        out["max_logit_err"] = float(err.max())
        out["mean_logit_err"] = float(err.mean())
        out["worst_err_over_tol"] = float((err / tol).max())
        out["frac_beyond_tol"] = float((err > tol).mean())
        out["sign_flips_outside_tol"] = int(((np.sign(sample["logits"]) != np.sign(ref)) & (np.abs(ref) > 1.5 * tol)).sum())
        out["oracle"] = "oracle.conv_forward (fp32 torch restatement of models.py:742-767)"
        out["criterion"] = "tolerance t = 1e-2 |ref| + 2e-3 on every logit (rows x N entries): at most 1e-4 of the entries " \
                           "beyond t, none beyond 1.5 t, no decision flip where |ref| > 1.5 t (the committed fixtures are held " \
                           "to t itself, tests/test_gpu_parity.py)"
        out["ok"] = (out["frac_beyond_tol"] <= 1e-4 and out["worst_err_over_tol"] <= 1.5 and
                     out["sign_flips_outside_tol"] == 0)
        return out
    if kind == "mc":
        # the sweep's counters for its first `rows` frames, reproduced by the oracle from the same Philox streams
        info = np.asarray(sample["info"], dtype=np.int32)
        ref = np.concatenate(oracle.run_threaded(lambda lo, hi: oracle.sc_decode(sample["y"][lo:hi], w["snr"], n, info)[2],
                                                 rows, os.cpu_count() or 1))
        bit, blk = oracle.count_errors(sample["msg"], ref)
        out["oracle"] = "oracle.sc_decode + oracle.count_errors on the sweep's own first frames"
        out["counts"] = sample["counts"]
        out["oracle_counts"] = [bit, blk, rows]
        out["criterion"] = "identical (bit errors, block errors, frames)"
        out["ok"] = list(sample["counts"]) == [bit, blk, rows]
        return out
    return None


def curve_check(wl, r, world):
    """BER / BLER of the timed steps against the curve the LIVE reference printed for the same checkpoint at this SNR
    (tests/golden/<checkpoint>.json): two-sample z-score of the block-error rates.  Every step decodes the same resident
    frames, so the sample is batch x n_gpus frames."""
    import math
    meta_path = wl["checkpoint"][:-3] + ".json"
    try:
        with open(meta_path) as f:
            meta = json.load(f)
        i = [abs(s - wl["snr"]) < 1e-9 for s in meta["snr_range"]].index(True)
    except Exception as e:  # no recorded point at this SNR
        return {"ok": None, "note": "no reference point: %s" % e}
    n1, n2 = wl["batch"] * world, meta["test_size"]
    key = "RNN" if "blers_RNN" in meta else "Xformer"
    p1, p2 = r["bler"], meta["blers_" + key][i]
    p = (p1 * n1 + p2 * n2) / float(n1 + n2)
    z = (p1 - p2) / math.sqrt(max(p * (1 - p), 1e-12) * (1.0 / n1 + 1.0 / n2))
    return {"snr_db": wl["snr"], "ber": r["ber"], "bler": p1, "reference_ber": meta["bers_" + key][i], "reference_bler": p2,
            "frames": n1, "reference_frames": n2, "z_bler": z, "ok": abs(z) < 3.29,
            "criterion": "|z| < 3.29 (two-sided 99.9 % two-sample interval on the block-error rate)"}


# ---------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("NPD_BENCH_WORKLOAD", "default"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU codewords per step (0 = workload default)")
    ap.add_argument("--chunk", type=int, default=0, help="mc* workloads: exact chunk size (0 = the workload's, rounded to whole "
                    "decoder rounds by sweep.sc_round_chunk)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="default workload without the other workloads")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    also = []
    if args.workload == "default":
        args.workload = "gru64"
        if not args.no_also:
            also = list(DEFAULT_ALSO)
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["batch"] = args.batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            line = reference_arm(args, args.workload, w)
            if also:  # the headline's second half and config 4, one bounded sample each
                line["cpu_baseline"]["other_workloads"] = {}
                for name in ("sc1024", "conv64"):
                    line["cpu_baseline"]["other_workloads"][name] = cpu_baseline(dict(WORKLOADS[name]), 10.0)
            print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the B200 path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION; keep stdout to the one JSON line
        # (the banner is printed at every level from VERSION up): send NCCL's own log to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    failed = []

    def run_workload(name, wl, steps, with_cpu):
        a = argparse.Namespace(**vars(args))
        a.workload, a.steps = name, steps
        kind = wl["kind"]
        if kind in ("gru", "grusweep") and wl["N"] == 64 and os.path.exists(GRU_CKPT):
            wl["checkpoint"] = GRU_CKPT
        if kind == "conv" and os.path.exists(CONV_CKPT):
            wl["checkpoint"] = CONV_CKPT
        if kind in ("gru", "grusweep"):
            from neural_polar_decoder_b200 import rnn_all
            rnn_all.set_gru_precision(wl.get("precision", "exact"))
        if kind == "enc":
            r = bench_enc(a, wl, rank, world, local_rank)
        elif kind == "sc":
            r = bench_sc(a, wl, rank, world, local_rank)
        elif kind == "mc":
            r = bench_mc(a, wl, rank, world, local_rank)
        elif kind == "train":
            r = bench_train(a, wl, rank, world, local_rank)
        else:
            from neural_polar_decoder_b200 import bench_neural
            r = bench_neural.bench(a, wl, rank, world, local_rank, ClockSampler, measured_peaks)
        r["config"] = dict(workload_config(name, wl), **{k: v for k, v in r.get("config", {}).items()
                                                         if k not in ("workload", "desc", "N", "K", "snr_db", "batch_per_gpu", "weights")})
        if kind in ("gru", "conv") and wl.get("checkpoint") and rank == 0:
            r["curve_check"] = curve_check(wl, r, world)
        sample = r.pop("_sample", None)
        if rank == 0 and sample is not None and not args.no_parity:
            r["parity_checked"] = parity_check(wl, sample)
            if wl.get("precision") == "fast":  # reported, not gated: the opt-in mode trades this error for speed
                r["parity_checked"]["gated"] = False
            elif not r["parity_checked"]["ok"]:
                failed.append(name)
        if rank == 0 and with_cpu and not args.no_cpu_baseline and world == 1:
            r["cpu_baseline"] = cpu_baseline(wl)
        return r

    res = run_workload(args.workload, w, args.steps, True)
    if also:
        keep = ("value", "unit", "steps", "ms_per_step", "dtype", "config", "e2e", "gpu_launches", "roofline", "ber", "bler",
                "frames", "cpu_baseline", "parity_checked", "curve_check", "scaling", "clocks", "metric")
        others = {}
        for name, cap, with_cpu in also:
            r = run_workload(name, dict(WORKLOADS[name]), min(args.steps, cap), with_cpu)
            others[name] = {k: r[k] for k in keep if k in r}
        res["also"] = others
        res["roofline"]["other_workloads"] = others
        res["gpu_launches_all_workloads"] = res["gpu_launches"] + sum(o.get("gpu_launches", 0) for o in others.values())
    if rank == 0:
        if failed:
            res["parity_failed"] = failed
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 1 if failed else 0


def _time_steps(torch, dist, world, dev, local_rank, warm, timed_body, steps, after_loop=None):
    """Barrier + synchronize, CUDA events around exactly `steps` iterations, max over ranks -> (ms, clocks)."""
    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    warm()
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t0.record()
    for i in range(steps):
        timed_body(i)
    if after_loop is not None:
        after_loop()
    t1.record()
    sync()
    clocks = sampler.stop()
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), clocks


def bench_enc(args, w, rank, world, local_rank):
    """Subsystem (1): fused message generation + encoder + channel (npd_gen_encode_awgn), HBM-write roofline."""
    import torch
    import torch.distributed as dist
    from neural_polar_decoder_b200 import _lib, utils
    lib = _lib.load()
    N, K, B, snr = w["N"], w["K"], w["batch"], w["snr"]
    code = make_code(w)
    h = code._handle()
    dev = torch.device("cuda", local_rank)
    sigma = float(np.float32(utils.snr_db2sigma(snr)))
    msg = torch.empty(B, K, device=dev)
    y = torch.empty(B, N, device=dev)
    st = _lib.stream_ptr()

    def step(i):
        _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), B, sigma, 2026, i, rank * B, st))

    ms, clocks = _time_steps(torch, dist, world, dev, local_rank, lambda: [step(i) for i in range(args.warmup)], step, args.steps)
    kern_ms = ms / args.steps
    peaks = measured_peaks()
    alg_bytes = (4 * N + 4 * K) * B  # y and msg written once
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    # end to end: the generator has no host input; its product read back on the host = the messages + y of one step
    e2e_B = min(B, 16384)
    t0 = time.perf_counter()
    for i in range(3):
        _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), e2e_B, sigma, 2026, i, rank * B, st))
        yh = y[:e2e_B].cpu()
    e2e_dt = time.perf_counter() - t0
    return {
        "metric": "generated + encoded + noised codewords/sec", "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": kern_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"l2_policy": "outputs larger than L2 (y = %d MB per GPU)" % (B * N * 4 >> 20)},
        "clocks": clocks, "gpu_launches": args.steps,
        "e2e": {"value": world * 3 * e2e_B / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": e2e_B * N * 4,
                "batch_per_gpu": e2e_B, "steps": 3,
                "api": "npd_gen_encode_awgn + y read back to (pageable) host memory; the generator has no host-side input"},
        "roofline": {"kernel": "encode_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm"], "traffic": ncu_traffic(args.workload, B)[0],
                     "traffic_source": ncu_traffic(args.workload, B)[1], "peak_source": peaks["src"], "kernel_ms": kern_ms,
                     "alg_bytes_per_launch": alg_bytes},
    }


def bench_sc(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from neural_polar_decoder_b200 import _lib, utils
    lib = _lib.load()
    N, K, B, snr = w["N"], w["K"], w["batch"], w["snr"]
    code = make_code(w)
    h = code._handle()
    dev = torch.device("cuda", local_rank)
    sigma = float(np.float32(utils.snr_db2sigma(snr)))
    scale = utils.llr_scale(snr)
    seed = 2026

    # synthetic frames, generated on the device by the fused generator (global codeword index =
    # rank*B + row, so the union over ranks is one reproducible stream), resident in HBM
    msg = torch.empty(B, K, device=dev)
    y = torch.empty(B, N, device=dev)
    dec = torch.empty(B, K, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), B, sigma, seed, 0, rank * B,
                                       _lib.stream_ptr()))
    st = _lib.stream_ptr()
    L = int(w.get("L", 0))

    def decode_call():
        if w.get("pac_g"):
            _lib.check(lib.npd_pac_sc_decode(h.h, _lib.ptr(y), scale, None, None, _lib.ptr(dec), None, B, st))
        elif L:
            _lib.check(lib.npd_scl_decode(h.h, _lib.ptr(y), scale, L, None, _lib.ptr(dec), B, st))
        else:
            _lib.check(lib.npd_sc_decode(h.h, _lib.ptr(y), scale, None, None, _lib.ptr(dec), B, st))

    def count_call():
        _lib.check(lib.npd_count_errors(_lib.ptr(msg), _lib.ptr(dec), B, K, _lib._vp(counts.data_ptr()), st))

    def warm():
        for _ in range(args.warmup):
            decode_call()
            count_call()
        if world > 1:
            dist.all_reduce(counts)  # warm-up of the collective too (its first call sets up NCCL channels)
        torch.cuda.synchronize()
        counts.zero_()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]

    def body(i):
        ev[i][0].record()
        decode_call()
        ev[i][1].record()
        count_call()
        ev[i][2].record()

    def after():
        if world > 1:
            dist.all_reduce(counts)  # the path's only collective: [bit errors, block errors, frames]

    elapsed_ms, clocks = _time_steps(torch, dist, world, dev, local_rank, warm, body, args.steps, after)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b, _ in ev]))
    count_ms = float(np.mean([b.elapsed_time(c) for _, b, c in ev]))
    value = world * B * args.steps / (elapsed_ms * 1e-3)
    rows = min(B, 1024)
    sample = {"y": y[:rows].cpu().numpy(), "decoded": dec[:rows].cpu().numpy(),
              "info": np.asarray(code.B if w.get("pac_g") else code.info_positions)}

    # ---- end to end through the drop-in with HOST buffers (H2D + kernel + D2H every step) ----
    e2e_B = min(B, 65536)
    y_host = y[:e2e_B].cpu().pin_memory()
    msg_host = msg[:e2e_B:61].cpu()
    e2e_steps = max(3, min(args.steps, 10))

    def host_call():
        if w.get("pac_g"):
            return code.pac_sc_decode(y_host, snr)[:2]
        if L:
            return code.scl_decode(y_host, snr, L, False, return_llr=False)
        return code.sc_decode_new(y_host, snr, return_llr=False)

    d_host = None
    for _ in range(3):  # warm-up holds the previous result like the timed loop does (pinned-pool steady state)
        _, d_host = host_call()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    errs = 0
    for _ in range(e2e_steps):
        _, d_host = host_call()  # returns host tensors
        errs += int((d_host[::61] != msg_host).sum())  # the step's result is read on the host (every 61st frame checked here)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_B * e2e_steps / float(e2e_t.item())

    peaks = measured_peaks()
    alg_bytes = (4 * N + 4 * K) * B  # fp32 y in + fp32 decisions out, per launch
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    cnt = counts.tolist()
    traffic, traffic_src = ncu_traffic(args.workload, B)
    lg = int(np.log2(N))
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"l2_policy": "inputs larger than L2 (y = %d MB per GPU)" % (B * N * 4 >> 20),
                   "step": "npd_sc_decode (y in HBM -> decisions in HBM) + npd_count_errors; one NCCL all-reduce "
                           "of the 3 counters at the end of the timed region when n_gpus > 1"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_B * N * 4,
                "d2h_bytes_per_step": e2e_B * K * 4, "batch_per_gpu": e2e_B, "steps": e2e_steps,
                "api": "PolarCode.sc_decode_new(pinned host y, snr) -> host decisions (npd_sc_decode_host: chunked H2D / "
                       "decode / D2H pipeline on three streams)", "sampled_bit_errors": errs},
        "gpu_launches": 2 * args.steps,
        "roofline": {"kernel": "scl_kernel" if L else ("sc_quad_kernel" if N >= 256 else "sc_lane_kernel"), "bound": "hbm",
                     "achieved": achieved, "peak": peaks["hbm"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm"], "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peaks["src"], "kernel_ms": kern_ms, "count_kernel_ms": count_ms,
                     "alg_bytes_per_launch": alg_bytes,
                     "llr_updates_per_s": N * lg * B / (kern_ms * 1e-3),
                     "smem_roofline": {"bound": "smem", "unit": "codewords/s",
                                       "achieved": B / (kern_ms * 1e-3),
                                       "peak": 37.2e12 / (12.0 * N * lg),
                                       "frac": (B / (kern_ms * 1e-3)) / (37.2e12 / (12.0 * N * lg)),
                                       "note": "SURVEY.md 8(d) second bound: N log2 N LLR updates x 12 B of shared-memory traffic "
                                               "each at 37.2 TB/s chip-wide (148 SMs x 128 B/clk x 1.965 GHz)"},
                     "note": "the decoder is instruction-issue bound, not HBM bound: N log2 N serial-by-level LLR updates "
                             "per codeword against 4N+4K bytes (DESIGN.md 4.1)"},
        "ber": cnt[0] / float(world * B * args.steps * K), "bler": cnt[1] / float(world * B * args.steps),
        "frames": world * B * args.steps, "_sample": sample,
    }


def bench_mc(args, w, rank, world, local_rank):
    """BASELINE config 5: the fused Monte-Carlo SC sweep (generate -> encode -> AWGN -> SC decode -> count on the device,
    polar.py:1258-1291's call pattern).  A step = one npd_mc_sc_sweep call over `batch` frames per GPU in chunks; rank r
    owns the global frame range [r*steps*batch, (r+1)*steps*batch) and ONE all-reduce of the 3 counters ends the region."""
    import torch
    import torch.distributed as dist
    from neural_polar_decoder_b200 import _lib, utils, sweep
    lib = _lib.load()
    N, K, F, snr = w["N"], w["K"], w["batch"], w["snr"]
    code = make_code(w)
    h = code._handle()
    # chunks of whole decoder rounds (the persistent SC kernel: 148 SMs x resident warps x 8 codewords per round)
    chunk = args.chunk if getattr(args, "chunk", 0) > 0 else sweep.sc_round_chunk(code, w["chunk"])
    dev = torch.device("cuda", local_rank)
    sigma = float(np.float32(utils.snr_db2sigma(snr)))
    scale = utils.llr_scale(snr)
    seed = 2026
    ws_bytes = lib.npd_mc_sc_workspace_bytes(h.h, chunk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr()
    base = rank * args.steps * F

    def sweep_call(frames, offset):
        _lib.check(lib.npd_mc_sc_sweep(h.h, frames, chunk, sigma, scale, seed, 0, offset, _lib._vp(ws.data_ptr()), ws_bytes,
                                       _lib._vp(counts.data_ptr()), st))

    # parity sample first: the sweep's counters over the first 1024 frames of this rank's range
    rows = 1024
    sweep_call(rows, base)
    torch.cuda.synchronize()
    first_counts = counts.tolist()
    msg = torch.empty(rows, K, device=dev)
    y = torch.empty(rows, N, device=dev)
    _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), rows, sigma, seed, 0, base, st))
    sample = {"y": y.cpu().numpy(), "msg": msg.cpu().numpy(), "counts": first_counts, "info": np.asarray(code.info_positions)}

    def warm():
        for _ in range(args.warmup):
            sweep_call(min(F, 4 * chunk), base)
        if world > 1:
            dist.all_reduce(counts)
        torch.cuda.synchronize()
        counts.zero_()

    def after():
        if world > 1:
            dist.all_reduce(counts)

    elapsed_ms, clocks = _time_steps(torch, dist, world, dev, local_rank, warm,
                                     lambda i: sweep_call(F, base + i * F), args.steps, after)
    frames = world * F * args.steps
    value = frames / (elapsed_ms * 1e-3)
    cnt = counts.tolist()
    assert cnt[2] == frames, (cnt, frames)

    # end to end through the drop-in call a user makes: sweep.mc_sc_sweep(...) -> BER / BLER lists on the host
    e2e_frames = min(F, 8 * chunk) * world
    e2e_steps = 3
    sweep.mc_sc_sweep(code, [snr], e2e_frames, chunk=chunk, seed=seed + 1)  # warm-up (workspace allocation)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ber, bler, fr, _ = sweep.mc_sc_sweep(code, [snr], e2e_frames, chunk=chunk, seed=seed + 2 + i)
    e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    peaks = measured_peaks()
    n_chunks = -(-F // chunk)
    alg_bytes = (4 * N + 4 * K) * F
    achieved = alg_bytes * args.steps / (elapsed_ms * 1e-3) / 1e9 / 1.0
    lg = int(np.log2(N))
    per_gpu = F * args.steps / (elapsed_ms * 1e-3)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"chunk": chunk, "chunk_rounds": chunk / float(max(1, lib.npd_sc_round_codewords(h.h))),
                   "frames_per_gpu": F * args.steps,
                   "l2_policy": "two chunks in flight; a chunk's y (%d MB) makes one HBM round trip" % (chunk * N * 4 >> 20),
                   "step": "npd_mc_sc_sweep over batch_per_gpu frames (Philox counters = global frame index); one NCCL "
                           "all-reduce of the 3 counters at the end of the timed region when n_gpus > 1"},
        "clocks": clocks,
        "e2e": {"value": e2e_steps * e2e_frames / float(e2e_t.item()), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 24,
                "batch_per_gpu": e2e_frames // world, "steps": e2e_steps,
                "api": "sweep.mc_sc_sweep(code, [snr], frames) -> (ber, bler, frames) lists on the host; the frames are "
                       "generated on the device, so the only transfer is the 24-byte counter read",
                "ber": ber[0], "bler": bler[0]},
        "gpu_launches": 3 * n_chunks * args.steps,
        "roofline": {"kernel": "encode_kernel + sc_quad_kernel + count (fused sweep, per chunk)", "bound": "hbm",
                     "achieved": achieved / world, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / world / peaks["hbm"],
                     "traffic": None, "peak_source": peaks["src"], "alg_bytes_per_launch": alg_bytes,
                     "llr_updates_per_s": N * lg * per_gpu,
                     "smem_roofline": {"bound": "smem", "unit": "codewords/s", "achieved": per_gpu,
                                       "peak": 37.2e12 / (12.0 * N * lg), "frac": per_gpu / (37.2e12 / (12.0 * N * lg))},
                     "note": "algorithmic bytes = the decoder's 4N+4K per frame; the sweep is issue-bound (generator Philox + "
                             "SC level walk), DESIGN.md 4.4"},
        "ber": cnt[0] / float(frames * K), "bler": cnt[1] / float(frames), "frames": frames, "_sample": sample,
    }


def bench_train(args, w, rank, world, local_rank):
    """one iteration of the reference's training loop (rnn_all.py:1399-1437) per step, teacher-forced, on a
    fresh batch generated on the device.  The reference has no distributed training: under torchrun every rank trains an
    independent replica ("replicas only"), value = trained codewords/s over all replicas."""
    import torch
    import torch.distributed as dist
    from neural_polar_decoder_b200 import _lib, construct, synth, utils
    from neural_polar_decoder_b200.polar import PolarCode
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder
    from neural_polar_decoder_b200.train import GRUTrainer
    lib = _lib.load()
    N, K, B, snr = w["N"], w["K"], w["batch"], w["snr"]
    dev = torch.device("cuda", local_rank)
    rs = construct.reference_rs256()
    code = PolarCode(int(np.log2(N)), K, None, rs=rs[rs < N])
    info = code.info_positions
    sd = synth.gru_state_dict(11, N, 512, 2, head_gain=2.0)
    net = RNN_Model('GRU', N + 2, 512, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', N, info, onehot=True)
    loss_code = dec._loss_code(info)
    tr = GRUTrainer(net, N, B, tf32=int(w["tf32"]))
    h = code._handle()
    sigma = float(np.float32(utils.snr_db2sigma(snr)))
    msg = torch.empty(B, K, device=dev)
    y = torch.empty(B, N, device=dev)
    gt = torch.ones(B, N, device=dev)
    info_t = torch.as_tensor(np.asarray(info), device=dev)
    st = _lib.stream_ptr()
    losses = []

    def step(i, want=False):
        _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), B, sigma, 2026, i, rank * B, st))
        gt[:, info_t] = msg
        loss, _, _ = tr.step(loss_code, y, gt, True, 1e-3, 0.25, want_loss=want)
        if want:
            losses.append(loss)

    ms, clocks = _time_steps(torch, dist, world, dev, local_rank, lambda: [step(i, i == 0) for i in range(args.warmup)],
                             lambda i: step(args.warmup + i), args.steps)
    step(args.warmup + args.steps, True)
    # end to end: pinned host y / gt in, loss out, every step
    y_h, gt_h = y.cpu().pin_memory(), gt.cpu().pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(3):
        loss, _, _ = tr.step(loss_code, y_h, gt_h, True, 1e-3, 0.25, want_loss=True)
    e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    peaks = measured_peaks()
    step_ms = ms / args.steps
    flops = 3.0 * (N * (2 * 3 * 512 * 3 * 512 + 2 * 512) + 2 * N * 3 * 512) * B  # forward + two backward GEMM families
    achieved = flops / (step_ms * 1e-3) / 1e12
    return {
        "metric": "trained codewords/sec", "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": ("f32", "tf32", "bf16", "f16")[int(w["tf32"])], "data": "synthetic",
        "config": {"step": "npd_gen_encode_awgn + gt scatter + npd_gru_train_step (64 forward steps per layer with saved gates, 64 "
                           "backward steps per layer, all-steps GEMMs for the inter-layer and weight-gradient products, "
                           "clip_grad_norm_ 0.25, AdamW lr 1e-3); independent replicas when n_gpus > 1",
                   "loss_first_last": [losses[0], losses[-1]] if losses else None},
        "clocks": clocks,
        "e2e": {"value": world * 3 * B / float(e2e_t.item()), "unit": UNIT, "h2d_bytes_per_step": 2 * B * N * 4,
                "d2h_bytes_per_step": 8, "batch_per_gpu": B, "steps": 3,
                "api": "train.GRUTrainer.step(loss_code, pinned host y, pinned host gt, ...) -> host loss"},
        "gpu_launches": (2 + (4 * 64 + 4) + (4 * 64 + 14) + 2) * args.steps,  # cuBLAS split-K reductions not counted
        "roofline": {"kernel": "cuBLAS SGEMM (library GEMMs) + cell_fwd / cell_bwd / head / adamw kernels", "bound": "tensor",
                     "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "traffic": None, "peak_source": peaks["src"],
                     "alg_flops_per_launch": flops,
                     "note": "3 x the decode's 302 MFLOP per codeword; fp32 (parity) GEMMs run on the CUDA cores, so the 16-bit "
                             "tensor peak is the contract's denominator, not this arithmetic's ceiling"},
    }


if __name__ == "__main__":
    sys.exit(main())
