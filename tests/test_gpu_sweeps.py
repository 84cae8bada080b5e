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
