"""bench.py legs for the neural decoders (CRISP GRU, convNet): device-resident throughput, end-to-end through the
drop-in API and the tensor-pipe roofline.  (The CPU legs live in oracle/cpu_arm.py: the reference arm must not
import this package.)"""
import json
import os
import time

import numpy as np
import torch

from . import _lib, construct, synth, utils
from .rnn_all import RNN_Model, RNN_decoder

METRIC = "decoded codewords/sec"
UNIT = "codewords/s"


def flops_per_codeword(N, H=512):
    """SURVEY.md App. D (hoisted form): N steps x (3 GEMV-equivalents H->3H + head) + the y projection."""
    return N * (2 * 3 * H * 3 * H + 2 * H) + 2 * N * 3 * H


def _setup(w, seed=11):
    """-> (net, info positions).  With w["checkpoint"]: the reference-trained CRISP GRU (reference checkpoint format,
    rnn_all.py:1471-1479) loaded the way the CLI loads it; otherwise seeded synthetic weights of the same architecture."""
    N, K = w["N"], w["K"]
    rs = construct.reference_rs256()
    info = np.sort(rs[rs < N][:K])
    if w.get("checkpoint"):
        from .cli import net_from_checkpoint
        net, cargs, _ = net_from_checkpoint(w["checkpoint"])
        assert (cargs.N, cargs.K, cargs.rnn_feature_size) == (N, K, 512), "checkpoint shape does not match the workload"
        return net, info
    sd = synth.gru_state_dict(seed, N, 512, 2, head_gain=8.0)
    net = RNN_Model('GRU', N + 2, 512, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return net, info


def _traffic(workload, batch):
    """roofline.traffic from the committed ncu capture (profiles/traffic.json; same helper as bench.py's)."""
    import json
    import os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
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


def bench(args, w, rank, world, local_rank, ClockSampler, measured_peaks):
    if w["kind"] == "conv":
        return bench_conv(args, w, rank, world, local_rank, ClockSampler, measured_peaks)
    if w["kind"] == "grusweep":
        return bench_grusweep(args, w, rank, world, local_rank, ClockSampler, measured_peaks)
    return bench_gru(args, w, rank, world, local_rank, ClockSampler, measured_peaks)


def bench_gru(args, w, rank, world, local_rank, ClockSampler, measured_peaks):
    import torch.distributed as dist
    from .polar import PolarCode
    N, K, B, snr = w["N"], w["K"], w["batch"], w["snr"]
    dev = torch.device("cuda", local_rank)
    net, info = _setup(w)
    rs = construct.reference_rs256()
    code = PolarCode(int(np.log2(N)), K, None, rs=rs[rs < N])
    dec = RNN_decoder('y_input', N, info, onehot=True)
    lib = _lib.load()
    h = code._handle()
    sigma = float(np.float32(utils.snr_db2sigma(snr)))
    msg = torch.empty(B, K, device=dev)
    y = torch.empty(B, N, device=dev)
    _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), B, sigma, 2026, 0, rank * B,
                                       _lib.stream_ptr()))
    gh = net.npd_handle(N)
    loss_code = dec._loss_code(info)
    decoded = torch.empty(B, N, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr()

    def step():
        _lib.check(lib.npd_gru_decode(gh.h, loss_code.h, _lib.ptr(y), None, None, None, _lib.ptr(decoded), B, None, 0, st))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tail():  # errors_ber / errors_bler of decoded[:, info_positions] against the messages, one kernel
        _lib.check(lib.npd_count_errors_info(h.h, _lib.ptr(msg), _lib.ptr(decoded), B, 0, _lib._vp(counts.data_ptr()), st))

    for _ in range(args.warmup):  # the WHOLE step: the gather / count kernels' first launches load their modules
        step()                    # (40 ms on a fresh box, which used to land in the timed region)
        tail()
    if world > 1:
        dist.all_reduce(counts)  # warm-up of the collective too (its first call sets up NCCL channels)
    counts.zero_()
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        tail()
    if world > 1:
        dist.all_reduce(counts)
    t_end.record()
    sync()
    clocks = sampler.stop()
    t = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # end to end through the drop-in (pinned host y -> H2D -> fused decode -> D2H decisions)
    e2e_B = B
    y_host = y[:e2e_B].cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 5))
    d_host = None
    for _ in range(3):  # holds the previous result like the timed loop (pinned-pool steady state)
        d_host = dec.decode(net, False, y_host)
    sync()
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(e2e_steps):
        d_host = dec.decode(net, False, y_host)
        acc += float(d_host[:, int(info[0])].sum())
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_B * e2e_steps / float(e2e_t.item())

    peaks = measured_peaks()
    fl = flops_per_codeword(N) * B
    achieved = fl / (kern_ms * 1e-3) / 1e12
    cnt = counts.tolist()
    traffic, traffic_src = _traffic(args.workload, B)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"l2_policy": "weights (4.8 MB fp16) are L2-resident by design; y = %.1f MB per GPU is read once per "
                                "launch" % (B * N * 4 / 2 ** 20),
                   "step": "npd_gru_decode (all N autoregressive steps, one launch) + npd_count_errors_info; one NCCL "
                           "all-reduce of the counters when n_gpus > 1"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_B * N * 4,
                "d2h_bytes_per_step": e2e_B * N * 4, "batch_per_gpu": e2e_B, "steps": e2e_steps,
                "api": "RNN_decoder.decode(net, False, pinned host y) -> host decisions (npd_gru_decode_host pipeline)"},
        "gpu_launches": 2 * args.steps,
        "roofline": {"kernel": "gru_decode_kernel3 (CTA-pair, cta_group::2)",
                     "bound": "tensor", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peaks["src"] + " (sustained 16-bit dense GEMM, cuBLAS bf16)", "kernel_ms": kern_ms,
                     "alg_flops_per_launch": fl, "flops_per_codeword": flops_per_codeword(N)},
        "ber": cnt[0] / float(world * B * args.steps * K), "bler": cnt[1] / float(world * B * args.steps),
        "frames": world * B * args.steps,
        "_sample": {"y": y[:1024].cpu().numpy(), "decoded": decoded[:1024].cpu().numpy(), "info": info, "net": net,
                    "loss_code": loss_code, "sd": {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}},
    }


def bench_grusweep(args, w, rank, world, local_rank, ClockSampler, measured_peaks):
    """Config 1's own call pattern (rnn_all.py:1771-1773, 1897): polar_RNN_full_test on batches of test_batch_size =
    10000 messages (host tensors, as the reference's DataLoader yields them) x n_snr SNR points.  A step = one batch:
    encode, n_snr x (channel + SC decode + count), ONE stacked GRU decode launch of n_snr x 10000 codewords + counts.
    value = GRU-decoded codewords/s (n_snr x batch per step)."""
    import torch.distributed as dist
    from . import sweep
    from .polar import PolarCode
    N, K, B, ns = w["N"], w["K"], w["batch"], w["n_snr"]
    dev = torch.device("cuda", local_rank)
    net, info = _setup(w)
    rs = construct.reference_rs256()
    code = PolarCode(int(np.log2(N)), K, None, rs=rs[rs < N])
    dec = RNN_decoder('y_input', N, info, onehot=True)
    snrs = [-2.0 + 4.0 * i / (ns - 1) for i in range(ns)]
    g = torch.Generator().manual_seed(7 + rank)
    batches = [(2 * (torch.rand(B, K, generator=g) < 0.5).float() - 1).pin_memory() for _ in range(args.steps)]

    def run(bs, seed):
        return sweep.polar_RNN_full_test(net, code, snrs, bs, False, False, False, decoder=dec, device=dev, seed=seed)

    for i in range(args.warmup):
        run(batches[:1], 100 + i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    res = run(batches, 2026)  # `steps` batches; the counters are read back once at the end of the call
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    frames = world * ns * B * args.steps
    peaks = measured_peaks()
    achieved = flops_per_codeword(N) * ns * B * args.steps / (elapsed_ms * 1e-3) / 1e12
    waves = -(-(ns * B) // (148 * 64))
    return {
        "metric": METRIC, "value": frames / (elapsed_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"snr_points": snrs,
                   "step": "sweep.polar_RNN_full_test on one 10000-message host batch: encode + %d x (channel, SC decode, "
                           "count) + one stacked GRU decode of %d codewords (%d waves of 74 CTA pairs, %.2f needed) + counts"
                           % (ns, ns * B, waves, ns * B / (148.0 * 64))},
        "clocks": clocks,
        "e2e": {"value": frames / (elapsed_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * K * 4, "d2h_bytes_per_step": 0,
                "batch_per_gpu": ns * B, "steps": args.steps,
                "api": "polar_RNN_full_test(net, code, snr_range, batches of pinned host messages) -> BER/BLER lists; the "
                       "noise is drawn on the device, the counters (%d B) are read back once per call" % (args.steps * ns * 48)},
        "gpu_launches": (1 + 3 * ns + 1 + ns) * args.steps,
        "roofline": {"kernel": "gru_decode_kernel3 inside the sweep step", "bound": "tensor", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "traffic": None, "peak_source": peaks["src"], "note": "whole sweep step, not the kernel alone"},
        "ber": res[0], "bler": res[1], "ber_sc": res[2], "bler_sc": res[3], "frames": frames,
    }


# ---------------------------------------------------------------------------------------------------
# convNet (reference models.py:691-772)
# ---------------------------------------------------------------------------------------------------
CONV_FLOPS_PER_CODEWORD = 2 * 26001408  # 26.0 M MAC (zero padding not discounted)


def _setup_conv(w, seed=21):
    import argparse
    from .models import convNet
    N = w["N"]
    if w.get("checkpoint"):  # the reference-trained convNet ({'xformer', 'step', 'args'}, run_models.py:980-983)
        from .run_models import net_from_checkpoint
        net, cargs, _ = net_from_checkpoint(w["checkpoint"])
        assert (cargs.N, cargs.embed_dim) == (N, 128), "checkpoint shape does not match the workload"
        return net
    sd = synth.conv_state_dict(seed, N, 128)
    net = convNet(argparse.Namespace(embed_dim=128, max_len=N, N=N, dont_use_bias=False, dropout=0.1))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net.eval()
    return net


def bench_conv(args, w, rank, world, local_rank, ClockSampler, measured_peaks):
    import ctypes
    import torch.distributed as dist
    from .polar import PolarCode
    N, K, B, snr = w["N"], w["K"], w["batch"], w["snr"]
    dev = torch.device("cuda", local_rank)
    net = _setup_conv(w)
    rs = construct.reference_rs256()
    code = PolarCode(int(np.log2(N)), K, None, rs=rs[rs < N])
    info = code.info_positions
    lib = _lib.load()
    h = code._handle()
    sigma = float(np.float32(utils.snr_db2sigma(snr)))
    msg = torch.empty(B, K, device=dev)
    y = torch.empty(B, N, device=dev)
    _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), B, sigma, 2026, 0, rank * B,
                                       _lib.stream_ptr()))
    ch = net.npd_handle()
    ws = ch.workspace(B, dev)
    wsp, wsn = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    logits = torch.empty(B, N, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr()
    chunks = -(-B // (wsn // (128 * 8192 * 2) * 128))

    def step():
        _lib.check(lib.npd_conv_forward(ch.h, _lib.ptr(y), _lib.ptr(logits), None, B, wsp, wsn, st))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tail():
        # decisions on the info positions = sign(logits) (run_models.py:338-339) and the error counters, one kernel
        _lib.check(lib.npd_count_errors_info(h.h, _lib.ptr(msg), _lib.ptr(logits), B, 1, _lib._vp(counts.data_ptr()), st))

    for _ in range(args.warmup):  # the whole step, so that no kernel's first launch falls into the timed region
        step()
        tail()
    if world > 1:
        dist.all_reduce(counts)  # warm-up of the collective too (its first call sets up NCCL channels)
    counts.zero_()
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        tail()
    if world > 1:
        dist.all_reduce(counts)
    t_end.record()
    sync()
    clocks = sampler.stop()
    t = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    e2e_B = B
    y_host = y[:e2e_B].cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 5))
    bits = None
    for _ in range(3):  # holds the previous result like the timed loop (pinned-pool steady state)
        bits, _ = net.decode(y_host, info, None, dev)
    sync()
    t0 = time.perf_counter()
    acc = 0.0
    for _ in range(e2e_steps):
        bits, _ = net.decode(y_host, info, None, dev)  # host bits [B,N,1]
        acc += float(bits[:, int(info[0]), 0].sum())
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_B * e2e_steps / float(e2e_t.item())

    peaks = measured_peaks()
    fl = CONV_FLOPS_PER_CODEWORD * B
    achieved = fl / (kern_ms * 1e-3) / 1e12
    cnt = counts.tolist()
    traffic, traffic_src = _traffic(args.workload, B)
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"l2_policy": "every step streams %.1f GB of fp16 activations through HBM between the two kernels "
                                "(workspace %.2f GB > L2), which evicts y and the weights' L2 lines each step"
                                % (2 * B * 16384 / 1e9, wsn / 1e9),
                   "step": "npd_conv_forward (conv_stack_kernel + conv_fc_kernel per %d-codeword chunk) + "
                           "npd_count_errors_info (sign + gather + count); one NCCL all-reduce of the counters when n_gpus > 1"
                           % (wsn // (128 * 8192 * 2) * 128)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_B * N * 4,
                "d2h_bytes_per_step": e2e_B * N * 4, "batch_per_gpu": e2e_B, "steps": e2e_steps,
                "api": "convNet.decode(pinned host y, info_positions, None, device) -> host bits (npd_conv_forward_host pipeline)"},
        "gpu_launches": (2 * chunks + 1) * args.steps,
        "roofline": {"kernel": "conv_stack_kernel + conv_fc_kernel (one npd_conv_forward call)", "bound": "tensor",
                     "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peaks["src"] + " (sustained 16-bit dense GEMM, cuBLAS bf16)", "kernel_ms": kern_ms,
                     "alg_flops_per_launch": fl, "flops_per_codeword": CONV_FLOPS_PER_CODEWORD},
        "ber": cnt[0] / float(world * B * args.steps * K), "bler": cnt[1] / float(world * B * args.steps),
        "frames": world * B * args.steps,
        "_sample": {"y": y[:1024].cpu().numpy(), "logits": logits[:1024].cpu().numpy(),
                    "sd": {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}},
    }
