"""bench.py legs for the neural decoders (CRISP GRU): device-resident throughput, end-to-end through the
drop-in API, tensor-pipe roofline, and the CPU baseline (torch eager nn.GRU stepping = what the reference's
RNN_decoder.decode does, rnn_all.py:532-547)."""
import json
import os
import time

import numpy as np
import torch

from . import _lib, construct, synth, utils
from .rnn_all import RNN_Model, RNN_decoder, gru_decode

METRIC = "decoded codewords/sec"
UNIT = "codewords/s"


def available():
    try:
        lib = _lib.load()
    except Exception:
        return False
    # the stub returns NPD_EUNSUPPORTED without touching CUDA; a real build validates its arguments first
    import ctypes
    h = ctypes.c_void_p()
    rc = lib.npd_gru_create(0, 0, None, None, None, None, None, None, None, None, None, None, ctypes.byref(h))
    return rc == _lib.NPD_EINVAL


def flops_per_codeword(N, H=512):
    """SURVEY.md App. D (hoisted form): N steps x (3 GEMV-equivalents H->3H + head) + the y projection."""
    return N * (2 * 3 * H * 3 * H + 2 * H) + 2 * N * 3 * H


def _setup(w, seed=11):
    N, K = w["N"], w["K"]
    rs = construct.reference_rs256()
    info = np.sort(rs[rs < N][:K])
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
    info_t = torch.as_tensor(info, device=dev)
    counts = torch.zeros(3, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr()
    dec_info = torch.empty(B, K, device=dev)

    def step():
        _lib.check(lib.npd_gru_decode(gh.h, loss_code.h, _lib.ptr(y), None, None, None, _lib.ptr(decoded), B, None, 0, st))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tail():
        torch.index_select(decoded, 1, info_t, out=dec_info)
        _lib.check(lib.npd_count_errors(_lib.ptr(msg), _lib.ptr(dec_info), B, K, _lib._vp(counts.data_ptr()), st))

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
        "config": {"workload": args.workload, "desc": w["desc"], "N": N, "K": K, "snr_db": snr, "batch_per_gpu": B,
                   "l2_policy": "weights (4.8 MB fp16) are L2-resident by design; y = %.1f MB per GPU is read once per "
                                "launch" % (B * N * 4 / 2 ** 20),
                   "step": "npd_gru_decode (all N autoregressive steps, one launch) + info-bit gather + "
                           "npd_count_errors; one NCCL all-reduce of the counters when n_gpus > 1",
                   "weights": "synthetic U(-1/sqrt(H), 1/sqrt(H)) (neural_polar_decoder_b200.synth, seed 11)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_B * N * 4,
                "d2h_bytes_per_step": e2e_B * N * 4, "batch_per_gpu": e2e_B, "steps": e2e_steps,
                "api": "RNN_decoder.decode(net, False, pinned host y) -> host decisions (npd_gru_decode_host pipeline)"},
        "gpu_launches": 3 * args.steps,
        "roofline": {"kernel": "gru_decode_kernel3 (CTA-pair, cta_group::2)",
                     "bound": "tensor", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peaks["src"] + " (sustained 16-bit dense GEMM, cuBLAS bf16)", "kernel_ms": kern_ms,
                     "alg_flops_per_launch": fl, "flops_per_codeword": flops_per_codeword(N)},
        "ber": cnt[0] / float(world * B * args.steps * K), "bler": cnt[1] / float(world * B * args.steps),
        "frames": world * B * args.steps,
    }


def cpu_rate(w, seconds, threads, seed=0):
    if w["kind"] == "conv":
        return cpu_rate_conv(w, seconds, threads, seed)
    return cpu_rate_gru(w, seconds, threads, seed)


def cpu_rate_gru(w, seconds, threads, seed=0):
    """CPU leg: the reference's own per-step structure -- torch eager nn.GRU(seq_len 1) + nn.Linear + sign
    feedback for N steps (rnn_all.py:532-547), fp32, all host threads."""
    torch.set_num_threads(threads)
    N, K = w["N"], w["K"]
    net, info = _setup(w)
    net = net.cpu().eval()
    info_set = set(int(i) for i in info)
    rng = np.random.RandomState(seed)

    def run(B):
        y = torch.from_numpy((rng.choice([-1.0, 1.0], size=(B, N)) + 0.9 * rng.randn(B, N)).astype(np.float32))
        t0 = time.perf_counter()
        with torch.no_grad():
            decoded = torch.ones(B, N)
            hidden = torch.zeros(2, B, net.feature_size)
            for ii in range(N):
                prev = torch.ones(B) if ii == 0 else decoded[:, ii - 1].sign()
                onehot = torch.eye(2)[(0.5 + 0.5 * prev).long()]
                out, hidden = net(torch.cat([y.unsqueeze(1), onehot.view(B, 1, 2)], 2), hidden)
                if ii in info_set:
                    decoded[:, ii] = out.squeeze().sign()
        return time.perf_counter() - t0

    dt = run(256)
    B = int(max(256, min(65536, 256 / dt * seconds)))
    dt = run(B)
    return B / dt, B, dt


def cpu_baseline(w):
    threads = os.cpu_count() or 1
    rate, B, dt = cpu_rate(w, 12.0, threads)
    how = ("torch eager Conv1d/GELU/Linear/LayerNorm stack as in models.py:742-767" if w["kind"] == "conv" else
           "torch eager nn.GRU stepping as in rnn_all.py:532-547")
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d codewords, %s, %.1f s" % (B, how, dt)}


# ---------------------------------------------------------------------------------------------------
# convNet (reference models.py:691-772)
# ---------------------------------------------------------------------------------------------------
CONV_FLOPS_PER_CODEWORD = 2 * 26001408  # 26.0 M MAC (zero padding not discounted)


def _setup_conv(w, seed=21):
    import argparse
    from .models import convNet
    N = w["N"]
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
    info_t = torch.as_tensor(info, device=dev)
    dec_info = torch.empty(B, K, device=dev)
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
        # decisions on the info positions = sign(logits) (run_models.py:338-339), then the error counters
        torch.sign(torch.index_select(logits, 1, info_t), out=dec_info)
        _lib.check(lib.npd_count_errors(_lib.ptr(msg), _lib.ptr(dec_info), B, K, _lib._vp(counts.data_ptr()), st))

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
        "config": {"workload": args.workload, "desc": w["desc"], "N": N, "K": K, "snr_db": snr, "batch_per_gpu": B,
                   "l2_policy": "every step streams %.1f GB of fp16 activations through HBM between the two kernels "
                                "(workspace %.2f GB > L2), which evicts y and the weights' L2 lines each step"
                                % (2 * B * 16384 / 1e9, wsn / 1e9),
                   "step": "npd_conv_forward (conv_stack_kernel + conv_fc_kernel per %d-codeword chunk) + sign/gather of "
                           "the info positions + npd_count_errors; one NCCL all-reduce of the counters when n_gpus > 1"
                           % (wsn // (128 * 8192 * 2) * 128),
                   "weights": "synthetic PyTorch-default-init (neural_polar_decoder_b200.synth, seed 21)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_B * N * 4,
                "d2h_bytes_per_step": e2e_B * N * 4, "batch_per_gpu": e2e_B, "steps": e2e_steps,
                "api": "convNet.decode(pinned host y, info_positions, None, device) -> host bits (npd_conv_forward_host pipeline)"},
        "gpu_launches": (2 * chunks + 4) * args.steps,
        "roofline": {"kernel": "conv_stack_kernel + conv_fc_kernel (one npd_conv_forward call)", "bound": "tensor",
                     "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peaks["src"] + " (sustained 16-bit dense GEMM, cuBLAS bf16)", "kernel_ms": kern_ms,
                     "alg_flops_per_launch": fl, "flops_per_codeword": CONV_FLOPS_PER_CODEWORD},
        "ber": cnt[0] / float(world * B * args.steps * K), "bler": cnt[1] / float(world * B * args.steps),
        "frames": world * B * args.steps,
    }


def cpu_rate_conv(w, seconds, threads, seed=0):
    """CPU leg: the reference's forward (models.py:742-767) with torch eager modules, fp32, all host threads."""
    torch.set_num_threads(threads)
    N = w["N"]
    net = _setup_conv(w).cpu()
    rng = np.random.RandomState(seed)

    def run(B):
        y = torch.from_numpy((rng.choice([-1.0, 1.0], size=(B, N)) + 0.9 * rng.randn(B, N)).astype(np.float32))
        t0 = time.perf_counter()
        with torch.no_grad():
            x2 = net.layers1(y.unsqueeze(1))
            x3 = net.layers2(x2) + x2
            x4 = net.layers3(x3) + x3
            x5 = net.layers4(x4) + x4
            x6 = net.layers5(x5)
            out = net.layer_norm(net.dropout(net.layersFin(torch.flatten(x6, start_dim=1))))
            out.sign()
        return time.perf_counter() - t0

    dt = run(256)
    B = int(max(256, min(1 << 17, 256 / dt * seconds)))
    dt = run(B)
    return B / dt, B, dt
