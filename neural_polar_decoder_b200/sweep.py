"""BER/BLER sweep loops of the reference on the B200 path, plus the multi-GPU Monte-Carlo
driver.

  polar_RNN_full_test  reference rnn_all.py:821-966   (GRU vs SC on a polar code)
  test_full_data       reference rnn_all.py:730-776   (GRU vs SC ("Dumer") on a PAC code)
  testXformer          reference run_models.py:297-371 (one-shot model, e.g. convNet, vs SC)
  mc_sc_sweep / mc_decoder_sweep  -- generate -> encode -> channel -> decode -> count entirely on the device,
                       sharded over torch.distributed ranks with ONE all-reduce of the counters at the end.

The reference loops read module globals (`args`, `code`, `decoder`, `device`); here they are keyword
arguments.  Semantics kept: the same codewords are re-noised for every SNR point, per-batch error RATES are
averaged (`+= ber / num_test_batches`), results are Python lists indexed by SNR point.  Counters stay on the
device and are read back once per call (the reference syncs with .item() after every decoder call).
SC-list decoding (polar.scl_decode) is run where the reference runs it; decoders the hot path does not cover (ML/MAP,
RNN list, Fano) are skipped: their lists stay 0."""
import numpy as np
import torch

from . import _lib, rng, utils


def _counts_buf(n_rows, device):
    return torch.zeros(n_rows, 2, dtype=torch.int64, device=device)


def _count_into(buf_row, a, b):
    a = a.contiguous()
    b = b.contiguous()
    _lib.check(_lib.load().npd_count_errors(_lib.ptr(a), _lib.ptr(b), a.shape[0], a.shape[1],
                                            _lib._vp(buf_row.data_ptr()), _lib.stream_ptr()))


def _count_info_into(buf_row, handle, msg, full, take_sign=False):
    """msg [B,K] against full[:, info positions of `handle`] (full [B,N]) without materialising the gather."""
    msg = msg.contiguous()
    full = full.contiguous()
    _lib.check(_lib.load().npd_count_errors_info(handle.h, _lib.ptr(msg), _lib.ptr(full), msg.shape[0], int(take_sign),
                                                 _lib._vp(buf_row.data_ptr()), _lib.stream_ptr()))


# The sequential decoders run one CTA pair per 128 codewords for the whole decode, so a launch costs whole waves of
# 74 pairs (148 SMs): the reference's test_batch_size of 10000 is 1.07 waves -- two waves of time.  The sweep loops
# therefore stack the SNR points of a batch (same codewords, independent noise) into ONE decode launch; the Philox
# noise of a frame depends on (seed, SNR point, global frame index) only, so the results equal per-point launches.
_STACK_ROWS = 1 << 20


def _snr_groups(n_snr, batch):
    per = max(1, min(n_snr, _STACK_ROWS // max(int(batch), 1)))
    return [list(range(i, min(i + per, n_snr))) for i in range(0, n_snr, per)]


def _rates(counts, sizes, K, n_snr, n_dec):
    """counts [batches, n_snr, n_dec, 2] -> per decoder (ber list, bler list) as averages of per-batch rates."""
    c = counts.cpu().numpy().astype(np.float64)
    nb = len(sizes)
    out = []
    for d in range(n_dec):
        ber = [float(sum(c[k, s, d, 0] / (sizes[k] * K) for k in range(nb)) / nb) for s in range(n_snr)]
        bler = [float(sum(c[k, s, d, 1] / sizes[k] for k in range(nb)) / nb) for s in range(n_snr)]
        out.append((ber, bler))
    return out


def polar_RNN_full_test(net, polar, snr_range, Test_Data_Generator, run_ML=False, run_SCL=False, run_RNNL=False,
                        decoder=None, device=None, seed=None, list_size=4):
    """-> (bers_RNN, blers_RNN, bers_SC, blers_SC, bers_SCL, blers_SCL, bers_RNNL, blers_RNNL, bers_ML, blers_ML)."""
    assert decoder is not None, "pass the RNN_decoder (a module global in the reference)"
    if run_ML or run_RNNL:
        # run_RNNL is hard-wired False at the reference's only call site (rnn_all.py:1897); run_ML is a 2^K brute-force
        # search (892-930).  Neither is on the B200 path: fail instead of returning silent zero curves.
        raise NotImplementedError("polar_RNN_full_test: run_ML / run_RNNL are not on the accelerated path")
    _lib.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    snr_range = list(snr_range)
    nb, ns = len(Test_Data_Generator), len(snr_range)
    info = torch.as_tensor(np.asarray(polar.info_positions), device=device)
    counts = torch.zeros(nb, ns, 3, 2, dtype=torch.int64, device=device)
    sizes, frame0 = [], 0
    seed = rng.get_seed() if seed is None else seed
    # --loss_only (rnn_all.py:850-891, 1189-1192): both decoders run genie-aided outside `loss_inds` and only the
    # message columns `msg_indices` are scored
    loss_inds = getattr(polar, "loss_inds", None)
    sel = None
    if loss_inds is not None:
        sel = torch.as_tensor(np.asarray(polar.msg_indices), device=device)
    n_scored = polar.K if sel is None else int(sel.numel())

    def scored(t):
        return t if sel is None else t.index_select(1, sel)

    with torch.cuda.device(device):
        for k, msg_bits in enumerate(Test_Data_Generator):
            msg = _lib.to_device_f32(msg_bits, device)
            sizes.append(msg.shape[0])
            x = polar.encode_plotkin(msg)
            gt = None
            if loss_inds is not None:  # rnn_all.py:844-845
                gt = torch.ones(msg.shape[0], polar.N, device=device)
                gt[:, info] = msg
            msg_s = scored(msg)
            b = msg.shape[0]
            for group in _snr_groups(ns, b):
                ys = torch.empty(len(group) * b, polar.N, device=device)
                for j, si in enumerate(group):
                    snr = snr_range[si]
                    y = polar.channel(x, snr, point=(1 << 31) | si, cw_offset=frame0, seed=seed)
                    ys[j * b:(j + 1) * b] = y
                    _, dec_sc = polar.sc_decode_new(y, snr, gt, return_llr=False)
                    _count_into(counts[k, si, 1], msg_s, scored(dec_sc))  # .sign() is the identity on {-1,0,+1}
                    if run_SCL:  # rnn_all.py:857-870 (args.list_size in the reference)
                        _, dec_scl = polar.scl_decode(y, snr, list_size, False, return_llr=False)
                        _count_into(counts[k, si, 2], msg_s, scored(dec_scl))
                if loss_inds is None:  # one launch for all SNR points of the group
                    dec = decoder.decode(net, False, ys)
                else:
                    dec = decoder.decode(net, False, ys, gt.repeat(len(group), 1), loss_inds=loss_inds)
                for j, si in enumerate(group):
                    d = dec[j * b:(j + 1) * b]
                    if sel is None:
                        _count_info_into(counts[k, si, 0], polar._handle(), msg, d)
                    else:
                        _count_into(counts[k, si, 0], msg_s, scored(d.index_select(1, info)))
            frame0 += msg.shape[0]
    (ber_r, bler_r), (ber_s, bler_s), (ber_l, bler_l) = _rates(counts, sizes, n_scored, ns, 3)
    zeros = [0. for _ in snr_range]
    return (ber_r, bler_r, ber_s, bler_s, ber_l, bler_l, list(zeros), list(zeros), list(zeros), list(zeros))


def test_full_data(net, code, snr_range, Test_Data_Generator, run_fano=False, run_dumer=True, decoder=None,
                   device=None, seed=None):
    """PAC sweep -> (bers_RNN, blers_RNN, bers_Dumer, blers_Dumer, bers_ML, blers_ML, bers_fano, blers_fano)."""
    assert decoder is not None
    _lib.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    snr_range = list(snr_range)
    nb, ns = len(Test_Data_Generator), len(snr_range)
    info = torch.as_tensor(np.asarray(code.B), device=device)
    counts = torch.zeros(nb, ns, 2, 2, dtype=torch.int64, device=device)
    sizes, frame0 = [], 0
    seed = rng.get_seed() if seed is None else seed
    with torch.cuda.device(device):
        for k, msg_bits in enumerate(Test_Data_Generator):
            msg = _lib.to_device_f32(msg_bits, device)
            sizes.append(msg.shape[0])
            x = code.pac_encode(msg)
            b = msg.shape[0]
            for group in _snr_groups(ns, b):
                ys = torch.empty(len(group) * b, code.N, device=device)
                for j, si in enumerate(group):
                    snr = snr_range[si]
                    y = code.channel(x, snr, point=(1 << 31) | si, cw_offset=frame0, seed=seed)
                    ys[j * b:(j + 1) * b] = y
                    if run_dumer:
                        _, v_hat, _ = code.pac_sc_decode(y, snr)
                        _count_into(counts[k, si, 1], msg, v_hat)
                dec = decoder.decode(net, False, ys)  # one launch for all SNR points of the group
                for j, si in enumerate(group):
                    _count_info_into(counts[k, si, 0], code._handle(), msg, dec[j * b:(j + 1) * b])
            frame0 += msg.shape[0]
    (ber_r, bler_r), (ber_d, bler_d) = _rates(counts, sizes, code.K, ns, 2)
    zeros = [0. for _ in snr_range]
    return ber_r, bler_r, ber_d, bler_d, list(zeros), list(zeros), list(zeros), list(zeros)


def testXformer(net, polar, snr_range, Test_Data_Generator, device, Test_Data_Mask=None, run_ML=False,
                bitwise_snr_idx=-1, seed=None, run_SCL=True, list_size=4):
    """One-shot decoders (net.decode(y, info_positions, mask, device) -> (bits[B,N,1], mask)).
    -> the reference's 11 values; ML / bitwise-MAP entries stay 0 (out of the hot path); run_SCL=False skips the
    list decoder the reference hard-wires at L = 4."""
    _lib.require_cuda()
    device = torch.device(device)
    snr_range = list(snr_range)
    nb, ns = len(Test_Data_Generator), len(snr_range)
    info = torch.as_tensor(np.asarray(polar.info_positions), device=device)
    counts = torch.zeros(nb, ns, 3, 2, dtype=torch.int64, device=device)
    bitwise = torch.zeros((1, polar.K), device=device)
    sizes, frame0 = [], 0
    seed = rng.get_seed() if seed is None else seed
    with torch.cuda.device(device):
        for k, msg_bits in enumerate(Test_Data_Generator):
            msg = _lib.to_device_f32(msg_bits, device)
            sizes.append(msg.shape[0])
            x = polar.encode_plotkin(msg)
            for si, snr in enumerate(snr_range):
                y = polar.channel(x, snr, point=(1 << 31) | si, cw_offset=frame0, seed=seed)
                _, dec_sc = polar.sc_decode_new(y, snr, return_llr=False)
                _count_into(counts[k, si, 1], msg, dec_sc)
                if run_SCL and not run_ML:  # run_models.py:328-333: list size 4 whenever the ML decoder is off
                    _, dec_scl = polar.scl_decode(y, snr, list_size, False, return_llr=False)
                    _count_into(counts[k, si, 2], msg, dec_scl)
                bits, _ = net.decode(y, polar.info_positions, None, device)
                dec = bits.reshape(bits.shape[0], -1).index_select(1, info)
                _count_into(counts[k, si, 0], msg, dec)
                if si == bitwise_snr_idx % ns and bitwise_snr_idx != -1:
                    bitwise += utils.errors_bitwise_ber(msg, dec.sign()).reshape(1, -1) / nb
            frame0 += msg.shape[0]
    (ber_x, bler_x), (ber_s, bler_s), (ber_l, bler_l) = _rates(counts, sizes, polar.K, ns, 3)
    zeros = [0. for _ in snr_range]
    return (ber_x, bler_x, ber_s, bler_s, ber_l, bler_l, list(zeros), list(zeros), bitwise, list(zeros), list(zeros))


# ---------------------------------------------------------------------------------------------------
# multi-GPU Monte-Carlo drivers: contiguous global frame ranges per rank, Philox counters =
# global frame index, one all-reduce(sum) of an int64 [n_snr, 3] tensor at the end.
# ---------------------------------------------------------------------------------------------------
def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` frames owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counts(counts, group=None):
    """Sum the [n_snr, 3] counter tensor over all ranks (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def finalize(counts, K):
    """[n_snr, 3] (bit errors, block errors, frames) -> (ber list, bler list, frames list)."""
    c = counts.cpu().numpy().astype(np.float64)
    frames = np.maximum(c[:, 2], 1)
    return (c[:, 0] / (frames * K)).tolist(), (c[:, 1] / frames).tolist(), c[:, 2].astype(np.int64).tolist()


def sc_round_chunk(polar, chunk):
    """`chunk` rounded to a whole number of rounds of the persistent SC kernel (npd_sc_round_codewords): a chunk of 9.2
    rounds costs 10.  Unchanged for codes the dynamically scheduled kernels decode."""
    rnd = int(_lib.load().npd_sc_round_codewords(polar._handle().h))
    if rnd <= 0 or chunk < rnd:
        return int(chunk)
    return (int(chunk) + rnd // 2) // rnd * rnd


def mc_sc_sweep(polar, snr_range, total_frames, chunk=1 << 17, seed=0, rank=None, world=None, group=None):
    """SC BER/BLER of `total_frames` frames per SNR point, fused on the device (npd_mc_sc_sweep).  `chunk` is rounded to
    whole rounds of the decoder (sc_round_chunk); the counts do not depend on it (Philox counters = global frame index)."""
    import torch.distributed as dist
    _lib.require_cuda()
    collective = rank is None  # explicit (rank, world) = a caller-simulated shard: no all-reduce, the caller sums
    if rank is None:
        rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(total_frames, rank, world)
    lib = _lib.load()
    h = polar._handle()
    dev = torch.device("cuda", torch.cuda.current_device())
    snr_range = list(snr_range)
    counts = torch.zeros(len(snr_range), 3, dtype=torch.int64, device=dev)
    chunk = int(max(1, min(sc_round_chunk(polar, chunk), max(hi - lo, 1))))
    ws_bytes = lib.npd_mc_sc_workspace_bytes(h.h, chunk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    for si, snr in enumerate(snr_range):
        if hi > lo:
            _lib.check(lib.npd_mc_sc_sweep(h.h, hi - lo, chunk, float(np.float32(utils.snr_db2sigma(snr))),
                                           utils.llr_scale(snr), int(seed), si, lo, _lib._vp(ws.data_ptr()), ws_bytes,
                                           _lib._vp(counts[si].data_ptr()), _lib.stream_ptr()))
    if collective:
        reduce_counts(counts, group)
    return finalize(counts, polar.K) + (counts,)


WAVE = 148 * 64  # codewords in one wave of the sequential decoders (64 per CTA, 148 SMs = 74 CTA pairs)


def round_to_waves(chunk):
    """Chunk sizes for the sequential decoders: a whole number of waves (a 3.46-wave chunk costs 4 waves of time)."""
    return max(WAVE, (int(chunk) + WAVE // 2) // WAVE * WAVE)


def mc_gru_sweep(polar, net, decoder, snr_range, total_frames, chunk=2 * WAVE, seed=0, rank=None, world=None, group=None):
    """GRU BER/BLER of `total_frames` frames per SNR point with generate -> decode -> count fused behind ONE library
    call per SNR point (npd_mc_gru_sweep).  Same sharding and counters as mc_sc_sweep."""
    import torch.distributed as dist
    _lib.require_cuda()
    collective = rank is None  # explicit (rank, world) = a caller-simulated shard: no all-reduce, the caller sums
    if rank is None:
        rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(total_frames, rank, world)
    lib = _lib.load()
    h = polar._handle()
    gh = net.npd_handle(polar.N)
    loss = decoder._loss_code(decoder.info_inds)
    dev = torch.device("cuda", torch.cuda.current_device())
    snr_range = list(snr_range)
    counts = torch.zeros(len(snr_range), 3, dtype=torch.int64, device=dev)
    chunk = int(max(1, min(round_to_waves(chunk), max(hi - lo, 1))))
    ws_bytes = lib.npd_mc_gru_workspace_bytes(gh.h, h.h, chunk)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    for si, snr in enumerate(snr_range):
        if hi > lo:
            _lib.check(lib.npd_mc_gru_sweep(gh.h, h.h, loss.h, hi - lo, chunk, float(np.float32(utils.snr_db2sigma(snr))),
                                            int(seed), si, lo, _lib._vp(ws.data_ptr()), ws_bytes,
                                            _lib._vp(counts[si].data_ptr()), _lib.stream_ptr()))
    if collective:
        reduce_counts(counts, group)
    return finalize(counts, polar.K) + (counts,)


def mc_decoder_sweep(polar, decode_fn, snr_range, total_frames, chunk=2 * WAVE, seed=0, rank=None, world=None,
                     group=None):
    """Same for any decoder: decode_fn(y[B,N]) -> decisions [B,N] (e.g. lambda y: decoder.decode(net, False, y))."""
    import torch.distributed as dist
    _lib.require_cuda()
    collective = rank is None  # explicit (rank, world) = a caller-simulated shard: no all-reduce, the caller sums
    if rank is None:
        rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(total_frames, rank, world)
    lib = _lib.load()
    h = polar._handle()
    dev = torch.device("cuda", torch.cuda.current_device())
    snr_range = list(snr_range)
    info = torch.as_tensor(np.asarray(polar.info_positions), device=dev)
    counts = torch.zeros(len(snr_range), 3, dtype=torch.int64, device=dev)
    for si, snr in enumerate(snr_range):
        for f0 in range(lo, hi, chunk):
            b = min(chunk, hi - f0)
            msg = torch.empty(b, polar.K, device=dev)
            y = torch.empty(b, polar.N, device=dev)
            _lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), b,
                                               float(np.float32(utils.snr_db2sigma(snr))), int(seed), si, f0,
                                               _lib.stream_ptr()))
            _count_info_into(counts[si], h, msg, decode_fn(y))
            counts[si, 2] += b
    if collective:
        reduce_counts(counts, group)
    return finalize(counts, polar.K) + (counts,)
