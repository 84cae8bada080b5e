"""GPU: the sweep loops and the Monte-Carlo drivers against the oracle on the very same
device-generated noise, split invariance of the sharded sweep, and the BER/BLER curve of SC Polar(64,22)
against the reference's published-in-survey curve within binomial 95% intervals."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _polar64():
    from neural_polar_decoder_b200.rnn_all import get_code
    return get_code("Polar", "polar", 64, 22)


def test_polar_rnn_full_test_matches_oracle_counts():
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder
    from neural_polar_decoder_b200.sweep import polar_RNN_full_test
    code = _polar64()
    N, K = 64, 22
    sd = synth.gru_state_dict(11, N, 512, 2, head_gain=8.0)
    net = RNN_Model('GRU', N + 2, 512, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', N, code.info_positions, onehot=True)
    rs = np.random.RandomState(0)
    batches = [torch.from_numpy((1.0 - 2.0 * rs.randint(0, 2, size=(b, K))).astype(np.float32)) for b in (300, 300, 170)]
    snrs = [-1.0, 1.0, 3.0]
    res = polar_RNN_full_test(net, code, snrs, batches, decoder=dec, seed=77)
    assert len(res) == 10 and all(len(r) == 3 for r in res)
    # recompute the SC lists with the oracle on the same noise (same seed / point / frame offsets)
    want_ber, want_bler = [0.0] * 3, [0.0] * 3
    f0 = 0
    for msg in batches:
        x = code.encode_plotkin(msg.cuda())
        for si, snr in enumerate(snrs):
            y = code.channel(x, snr, point=(1 << 31) | si, cw_offset=f0, seed=77).cpu().numpy()
            _, _, d = oracle.sc_decode(y, snr, 6, code.info_positions)
            bit, blk = oracle.count_errors(msg.numpy(), d)
            want_ber[si] += bit / (msg.shape[0] * K) / len(batches)
            want_bler[si] += blk / msg.shape[0] / len(batches)
        f0 += msg.shape[0]
    assert res[2] == pytest.approx(want_ber, abs=1e-12) and res[3] == pytest.approx(want_bler, abs=1e-12)
    assert res[4] == [0.0] * 3  # SCL skipped
    # GRU lists: recompute through the fused decode API (the sweep must just be bookkeeping around it)
    info = torch.as_tensor(code.info_positions).cuda()
    f0, ber = 0, [0.0] * 3
    for msg in batches:
        x = code.encode_plotkin(msg.cuda())
        for si, snr in enumerate(snrs):
            y = code.channel(x, snr, point=(1 << 31) | si, cw_offset=f0, seed=77)
            d = dec.decode(net, False, y)[:, info].cpu().numpy()
            ber[si] += (d != msg.numpy()).mean() / len(batches)
        f0 += msg.shape[0]
    assert res[0] == pytest.approx(ber, abs=1e-9)


def test_pac_sweep_and_gru_on_pac_code():
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.sweep import test_full_data as pac_sweep
    pac = get_code("PAC", "RM", 32, 16, g=53)
    N, K = 32, 16
    sd = synth.gru_state_dict(12, N, 512, 2, head_gain=8.0)
    net = RNN_Model('GRU', N + 2, 512, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', N, pac.info_inds, onehot=True)
    rs = np.random.RandomState(1)
    batches = [torch.from_numpy((1.0 - 2.0 * rs.randint(0, 2, size=(256, K))).astype(np.float32)) for _ in range(2)]
    res = pac_sweep(net, pac, [0.0, 2.0], batches, decoder=dec, seed=3)
    assert len(res) == 8
    f0, want = 0, [0.0, 0.0]
    for msg in batches:
        x = pac.pac_encode(msg.cuda())
        for si, snr in enumerate([0.0, 2.0]):
            y = pac.channel(x, snr, point=(1 << 31) | si, cw_offset=f0, seed=3).cpu().numpy()
            _, v, _ = oracle.pac_sc_decode(y, snr, 5, pac.B, 53)
            want[si] += oracle.count_errors(msg.numpy(), v)[1] / 256 / 2
        f0 += 256
    assert res[3] == pytest.approx(want, abs=1e-12)
    assert 0.0 < res[3][1] < res[3][0] < 1.0


def test_mc_sc_sweep_split_invariance_and_curve():
    """Counter-based streams: 1 rank == sum over 3 simulated ranks, bit for bit; and the SC curve of
    Polar(64,22) agrees with the reference run recorded in BASELINE.md (100k frames, seed 0) within
    two-sample intervals at a family-wise 95 % level (tests/mcstats.py; block errors binomial, bit errors with the
    per-frame variance estimated from a sample)."""
    from neural_polar_decoder_b200.sweep import mc_sc_sweep
    code = _polar64()
    snrs = [-2.0, -1.0, 0.0, 1.0, 2.0]
    frames = 100000
    ber, bler, fr, counts = mc_sc_sweep(code, snrs, frames, chunk=40000, seed=1, rank=0, world=1)
    parts = [mc_sc_sweep(code, snrs, frames, chunk=7777, seed=1, rank=r, world=3)[3] for r in range(3)]
    assert torch.equal(counts, parts[0] + parts[1] + parts[2])
    assert fr == [frames] * 5
    ref_bler = [0.49265, 0.28446, 0.12171, 0.03666, 0.00732]  # BASELINE.md 2, reference sc_decode_new
    ref_ber = [0.18100, 0.09901, 0.03978, 0.01123, 0.00195]
    import mcstats
    z = mcstats.z_familywise(10)  # family-wise 95 % over the 10 intervals asserted below
    # per-frame variance of the bit-error fraction from a 20k-frame sample of the same decoder
    g = torch.Generator().manual_seed(5)
    msg = (1.0 - 2.0 * torch.randint(0, 2, (20000, code.K), generator=g).float()).cuda()
    x = code.encode_plotkin(msg)
    for i, snr in enumerate(snrs):
        hw = mcstats.bler_halfwidth(bler[i], frames, ref_bler[i], 100000, z)
        assert abs(bler[i] - ref_bler[i]) <= hw, ("bler", snr, bler[i], ref_bler[i], hw)
        _, d = code.sc_decode_new(code.channel(x, snr, point=77 + i, seed=8), snr, return_llr=False)
        hw = mcstats.ber_halfwidth(mcstats.frame_fraction_var(msg, d), frames, 100000, z)
        assert abs(ber[i] - ref_ber[i]) <= hw + 5e-6, ("ber", snr, ber[i], ref_ber[i], hw)  # ref_ber has 5 digits


def test_sc_round_chunk_and_chunk_invariance():
    """npd_sc_round_codewords: a whole number of resident groups of 8 per SM for the persistent kernel (N >= 256), 0 for the
    lane-kernel codes; sweep.sc_round_chunk rounds to it; the sweep's counters do not depend on the chunk size."""
    from neural_polar_decoder_b200 import PolarCode, _lib, construct
    from neural_polar_decoder_b200.sweep import mc_sc_sweep, sc_round_chunk
    lib = _lib.load()
    assert lib.npd_sc_round_codewords(_polar64()._handle().h) == 0
    assert sc_round_chunk(_polar64(), 40000) == 40000
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for n, K in ((8, 128), (10, 512), (12, 2048)):
        code = PolarCode(n, K, None, F=construct.pw_frozen_set(1 << n, K))
        rnd = lib.npd_sc_round_codewords(code._handle().h)
        assert rnd > 0 and rnd % (8 * sms) == 0, (n, rnd)
        assert sc_round_chunk(code, 3 * rnd + 5) == 3 * rnd and sc_round_chunk(code, rnd // 2) == rnd // 2
    code = PolarCode(10, 512, None, F=construct.pw_frozen_set(1024, 512))
    rnd = lib.npd_sc_round_codewords(code._handle().h)
    frames = 2 * rnd + 1234
    a = mc_sc_sweep(code, [2.0], frames, chunk=rnd, seed=3, rank=0, world=1)[3]
    b = mc_sc_sweep(code, [2.0], frames, chunk=5000, seed=3, rank=0, world=1)[3]
    assert torch.equal(a, b) and int(a[0, 2]) == frames


def test_mc_decoder_sweep_gru_statistics():
    """GRU sweep through the generic driver: frame counts and split invariance over simulated ranks."""
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder
    from neural_polar_decoder_b200.sweep import mc_decoder_sweep
    code = _polar64()
    sd = synth.gru_state_dict(11, 64, 512, 2, head_gain=8.0)
    net = RNN_Model('GRU', 66, 512, 1, 2, 64, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', 64, code.info_positions, onehot=True)
    fn = lambda y: dec.decode(net, False, y)
    full = mc_decoder_sweep(code, fn, [0.0], 3000, chunk=1024, seed=4, rank=0, world=1)[3]
    parts = [mc_decoder_sweep(code, fn, [0.0], 3000, chunk=999, seed=4, rank=r, world=2)[3] for r in range(2)]
    assert torch.equal(full, parts[0] + parts[1]) and int(full[0, 2]) == 3000


@pytest.mark.gpu
def test_mc_sweep_cli(capsys):
    """`python -m neural_polar_decoder_b200.mc_sweep` (BASELINE config 5 call pattern): SC and SC-list sweeps print one
    JSON line; the SC-list curve lies below the SC curve and the frame count is what was asked for."""
    import json
    from neural_polar_decoder_b200 import mc_sweep
    assert mc_sweep.main(["--N", "256", "--K", "128", "--snr", "1.5", "2.5", "--frames", "200000"]) == 0
    sc = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert sc["frames"] == [200000, 200000] and sc["bler"][0] > sc["bler"][1] > 0
    assert mc_sweep.main(["--N", "256", "--K", "128", "--snr", "1.5", "2.5", "--frames", "200000", "--list_size", "4"]) == 0
    scl = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert scl["decoder"] == "SCL-4" and all(l < s for l, s in zip(scl["bler"], sc["bler"]))


def test_count_errors_info_equals_gathered_count():
    from neural_polar_decoder_b200 import _lib
    from neural_polar_decoder_b200.sweep import _count_info_into, _count_into
    code = _polar64()
    rs = np.random.RandomState(3)
    msg = torch.from_numpy((1.0 - 2.0 * rs.randint(0, 2, size=(777, 22))).astype(np.float32)).cuda()
    full = torch.from_numpy(rs.choice([-1.0, 0.0, 1.0], size=(777, 64), p=[0.45, 0.1, 0.45]).astype(np.float32)).cuda()
    info = torch.as_tensor(code.info_positions).cuda()
    full[:, info] = torch.where(torch.from_numpy(rs.rand(777, 22) < 0.9).cuda(), msg, full[:, info])
    a = torch.zeros(2, dtype=torch.int64, device="cuda")
    b = torch.zeros(2, dtype=torch.int64, device="cuda")
    _count_info_into(a, code._handle(), msg, full)
    _count_into(b, msg, full.index_select(1, info))
    bit, blk = oracle.count_errors(msg.cpu().numpy(), full[:, info].cpu().numpy())
    assert a.tolist() == b.tolist() == [bit, blk] and 0 < blk < 777


def test_mc_gru_sweep_equals_generic_driver():
    """npd_mc_gru_sweep (generate -> decode -> count behind one call) == the generic Python driver, any chunking."""
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder
    from neural_polar_decoder_b200.sweep import mc_decoder_sweep, mc_gru_sweep
    code = _polar64()
    sd = synth.gru_state_dict(11, 64, 512, 2, head_gain=8.0)
    net = RNN_Model('GRU', 66, 512, 1, 2, 64, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', 64, code.info_positions, onehot=True)
    want = mc_decoder_sweep(code, lambda y: dec.decode(net, False, y), [0.0, 2.0], 5000, chunk=1999, seed=4, rank=0, world=1)[3]
    got = mc_gru_sweep(code, net, dec, [0.0, 2.0], 5000, seed=4, rank=0, world=1)[3]
    assert torch.equal(got, want) and got[:, 2].tolist() == [5000, 5000]
    parts = [mc_gru_sweep(code, net, dec, [0.0, 2.0], 5000, seed=4, rank=r, world=3)[3] for r in range(3)]
    assert torch.equal(parts[0] + parts[1] + parts[2], want)


def test_loss_only_sweep_scores_genie_aided_subset(golden):
    """--loss_only (rnn_all.py:850-891, 1189-1192): SC and the GRU run genie-aided outside loss_inds and only the
    message columns msg_indices are scored.  SC list rebuilt with the oracle (use_gt) on the same noise."""
    import argparse
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.sweep import polar_RNN_full_test
    code = get_code("Polar", "rev_polar", 64, 22, args=argparse.Namespace(target_K=22, loss_only=6))
    g = golden("misc")
    assert np.array_equal(code.loss_inds, g["lossonly_inds"]) and len(code.msg_indices) == 6
    sd = synth.gru_state_dict(11, 64, 512, 2, head_gain=8.0)
    net = RNN_Model('GRU', 66, 512, 1, 2, 64, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', 64, code.info_positions, onehot=True)
    rs = np.random.RandomState(2)
    batches = [torch.from_numpy((1.0 - 2.0 * rs.randint(0, 2, size=(400, 22))).astype(np.float32)) for _ in range(2)]
    snrs = [-2.0, 0.0]
    res = polar_RNN_full_test(net, code, snrs, batches, decoder=dec, seed=9)
    want, want_rnn, f0 = [0.0, 0.0], [0.0, 0.0], 0
    sel = code.msg_indices
    for msg in batches:
        x = code.encode_plotkin(msg.cuda())
        gt = np.ones((400, 64), np.float32)
        gt[:, code.info_positions] = msg.numpy()
        for si, snr in enumerate(snrs):
            y = code.channel(x, snr, point=(1 << 31) | si, cw_offset=f0, seed=9)
            _, _, d = oracle.sc_decode(y.cpu().numpy(), snr, 6, code.info_positions, use_gt=gt)
            want[si] += (d[:, sel] != msg.numpy()[:, sel]).mean() / 2
            dr = dec.decode(net, False, y, torch.from_numpy(gt).cuda(), loss_inds=code.loss_inds).cpu().numpy()
            want_rnn[si] += (dr[:, code.info_positions][:, sel] != msg.numpy()[:, sel]).mean() / 2
        f0 += 400
    assert res[2] == pytest.approx(want, abs=1e-12)
    assert res[0] == pytest.approx(want_rnn, abs=1e-12)
    # genie-aided outside loss_inds: every scored SC decision equals the genie value where loss_inds is empty -> sanity
    assert all(0.0 <= v < 0.5 for v in res[2])
