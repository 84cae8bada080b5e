"""CPU, world_size 2 over gloo: the multi-GPU Monte-Carlo plumbing (contiguous frame shards, counter-based
streams indexed by the GLOBAL frame number, one all-reduce of the [n_snr,3] counters).  The decode itself is
played by the oracle here; on the GPU box the same plumbing wraps the CUDA kernels (tests -m gpu)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _local_counts(lo, hi, snrs, N, K, info, seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    n = int(np.log2(N))
    counts = torch.zeros(len(snrs), 3, dtype=torch.int64)
    if hi > lo:
        msg = oracle.gen_msg(seed, lo, hi - lo, K)
        x = oracle.polar_encode(msg, n, info)
        for si, snr in enumerate(snrs):
            y = x + np.float32(10 ** (-snr / 20)) * oracle.gen_noise(seed, lo, si, hi - lo, N)
            _, _, dec = oracle.sc_decode(y, snr, n, info)
            bit, blk = oracle.count_errors(msg, dec)
            counts[si] = torch.tensor([bit, blk, hi - lo])
    return counts


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from neural_polar_decoder_b200 import construct, sweep
    N, K = 64, 22
    rs = construct.reference_rs256()
    info = np.sort(rs[rs < N][:K])
    lo, hi = sweep.shard_range(total, rank, world)
    counts = _local_counts(lo, hi, [0.0, 2.0], N, K, info, seed=9)
    sweep.reduce_counts(counts)
    if rank == 0:
        torch.save(counts, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sweep_equals_single_rank(tmp_path):
    sys.path.insert(0, ROOT)
    from neural_polar_decoder_b200 import construct, sweep
    total = 1001  # odd: ranks get 501 / 500 frames
    out = str(tmp_path / "counts.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    got = torch.load(out)
    N, K = 64, 22
    rs = construct.reference_rs256()
    info = np.sort(rs[rs < N][:K])
    want = _local_counts(0, total, [0.0, 2.0], N, K, info, seed=9)
    assert torch.equal(got, want), (got, want)
    ber, bler, frames = sweep.finalize(got, K)
    assert frames == [total, total] and 0 < bler[1] < bler[0] < 1
