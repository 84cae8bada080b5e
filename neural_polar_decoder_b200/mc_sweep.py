"""Monte-Carlo SC / SC-list BER/BLER sweep from the command line (BASELINE.json config 5; the call pattern of the
reference's `python polar.py` sweep, polar.py:1258-1291, with the frames generated, decoded and counted on the GPU).

  python -m neural_polar_decoder_b200.mc_sweep --N 1024 --K 512 --snr 1 1.5 2 2.5 --frames 1e9
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m neural_polar_decoder_b200.mc_sweep ...

Under torchrun every rank decodes a contiguous range of the global frame indices (Philox counters = global index, so
the counts do not depend on the GPU count) and ONE all-reduce (NCCL) of the int64 [n_snr, 3] counters ends the run.
The frozen set is the reference's reliability table for N <= 256 and the polarization-weight construction above
(`construct.pw_frozen_set`)."""
import argparse
import json
import os
import time

import numpy as np
import torch


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--K", type=int, default=512)
    ap.add_argument("--snr", type=float, nargs="+", default=[1.0, 1.5, 2.0, 2.5, 3.0])
    ap.add_argument("--frames", type=float, default=1e8, help="frames per SNR point (whole job)")
    ap.add_argument("--list_size", type=int, default=0, help="0 = SC (fused sweep), L >= 1 = SC-list with L paths")
    ap.add_argument("--chunk", type=int, default=1 << 17)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)

    import torch.distributed as dist
    from . import PolarCode, construct, sweep
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = int(np.log2(a.N))
    if a.N <= 256:
        rs = construct.reference_rs256()
        code = PolarCode(n, a.K, None, rs=rs[rs < a.N])
    else:
        code = PolarCode(n, a.K, None, F=construct.pw_frozen_set(a.N, a.K))
    total = int(a.frames)
    torch.cuda.synchronize()
    t0 = time.time()
    if a.list_size <= 0:
        ber, bler, frames, _ = sweep.mc_sc_sweep(code, a.snr, total, chunk=a.chunk, seed=a.seed)
    else:
        ber, bler, frames, _ = _scl_sweep(code, a, total)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"N": a.N, "K": a.K, "decoder": "SC" if a.list_size <= 0 else "SCL-%d" % a.list_size,
                          "snr_db": a.snr, "ber": ber, "bler": bler, "frames": frames, "n_gpus": world,
                          "seconds": dt, "codewords_per_s": sum(frames) / dt}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _scl_sweep(code, a, total):
    """SC-list variant of the sweep: same generator and counters, scl_decode in the middle."""
    from . import sweep
    info = torch.as_tensor(np.asarray(code.info_positions), device="cuda")

    def run_point(snr):
        def decode_fn(y):
            _, dec = code.scl_decode(y, snr, a.list_size, False, return_llr=False)
            full = torch.ones(y.shape[0], code.N, device=y.device)
            full[:, info] = dec
            return full
        return sweep.mc_decoder_sweep(code, decode_fn, [snr], total, chunk=min(a.chunk, 1 << 16), seed=a.seed)

    ber, bler, frames = [], [], []
    for si, snr in enumerate(a.snr):
        b, bl, fr, _ = run_point(snr)
        ber += b; bler += bl; frames += fr
    return ber, bler, frames, None


if __name__ == "__main__":
    raise SystemExit(main())
