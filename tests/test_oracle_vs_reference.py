"""CPU, build container only: the oracle side by side with the LIVE reference (/root/reference through
oracle/ref_shim.py) on fresh seeded inputs -- not on the committed fixtures.  Skipped where the reference is
absent (the GPU box); there tests/test_oracle_golden.py pins the oracle against the fixtures minted here.

Bit-exact for the encoder / SC / SC-list / PAC / counters, fp32 round-off for the GRU and convNet restatements."""
import numpy as np
import pytest
import torch

import oracle
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="live reference not present")


def _msgs(rs, B, K):
    return (1.0 - 2.0 * rs.randint(0, 2, size=(B, K))).astype(np.float32)


@pytest.mark.parametrize("N,K,prof,snr", [(8, 4, "polar", 1.0), (32, 16, "polar", 0.0), (64, 22, "rev_polar", -1.0),
                                          (128, 64, "polar", 2.0)])
def test_encoder_and_sc_side_by_side(N, K, prof, snr):
    rs = np.random.RandomState(N + K)
    code = ref_shim.get_code("Polar", prof, N, K)
    info = np.asarray(code.info_positions, dtype=np.int32)
    n = int(np.log2(N))
    msg = _msgs(rs, 96, K)
    x_ref = code.encode_plotkin(torch.from_numpy(msg)).numpy()
    assert np.array_equal(oracle.polar_encode(msg, n, info), x_ref)
    sigma = 10 ** (-snr / 20)
    y = (x_ref + sigma * rs.randn(*x_ref.shape)).astype(np.float32)
    y[:8] = np.round(y[:8] * 2) / 2  # half-integer LLR inputs: exact cancellations -> sign(0) = 0 ties
    llr_ref, dec_ref = code.sc_decode_new(torch.from_numpy(y), snr)
    llr, _, dec = oracle.sc_decode(y, snr, n, info)
    assert np.array_equal(llr, llr_ref.numpy())
    assert np.array_equal(dec, dec_ref.numpy())
    # genie-aided (--loss_only path, rnn_all.py:862)
    gt = np.ones((96, N), np.float32)
    gt[:, info] = msg
    llr_ref, dec_ref = code.sc_decode_new(torch.from_numpy(y), snr, torch.from_numpy(gt))
    llr, _, dec = oracle.sc_decode(y, snr, n, info, use_gt=gt)
    assert np.array_equal(llr, llr_ref.numpy()) and np.array_equal(dec, dec_ref.numpy())


def test_scl_side_by_side():
    rs = np.random.RandomState(3)
    code = ref_shim.get_code("Polar", "polar", 32, 16)
    info = np.asarray(code.info_positions, dtype=np.int32)
    msg = _msgs(rs, 24, 16)
    x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
    y = (x + rs.randn(*x.shape)).astype(np.float32)
    for L in (1, 2, 4):
        llr_ref, dec_ref = code.scl_decode(torch.from_numpy(y), 0.0, L, False)
        llr, dec = oracle.scl_decode(y, 0.0, 5, info, L)
        assert np.array_equal(dec, dec_ref.numpy()), L
        assert np.array_equal(llr, llr_ref.numpy()), L


def test_pac_side_by_side():
    rs = np.random.RandomState(4)
    pac = ref_shim.get_code("PAC", "RM", 32, 16, g=53)
    info = np.asarray(pac.B, dtype=np.int32)
    msg = _msgs(rs, 64, 16)
    x_ref = pac.pac_encode(torch.from_numpy(msg)).numpy()
    assert np.array_equal(oracle.pac_encode(msg, 5, info, 53), x_ref)
    y = (x_ref + 0.9 * rs.randn(*x_ref.shape)).astype(np.float32)
    y[:6] = np.round(y[:6] * 2) / 2
    llr_ref, v_ref, u_ref = pac.pac_sc_decode(torch.from_numpy(y), 1.0)
    llr, v, u = oracle.pac_sc_decode(y, 1.0, 5, info, 53)
    assert np.array_equal(llr, llr_ref.numpy())
    assert np.array_equal(v, v_ref.numpy())
    assert np.array_equal(u, u_ref.numpy())


def test_error_counters_side_by_side():
    ru = ref_shim.load("utils")
    rs = np.random.RandomState(6)
    a = _msgs(rs, 200, 22)
    b = a.copy()
    b[rs.rand(200, 22) < 0.02] *= -1
    b[5, 3] = 0.0
    bits, blocks = oracle.count_errors(a, b)
    assert bits / a.size == pytest.approx(ru.errors_ber(torch.from_numpy(a), torch.from_numpy(b)).item(), abs=1e-7)
    assert blocks / 200 == pytest.approx(float(ru.errors_bler(torch.from_numpy(a), torch.from_numpy(b))), abs=1e-12)


def test_gru_decode_side_by_side():
    """RNN_decoder.decode test branch (rnn_all.py:514-561) on a seeded random-init RNN_Model vs oracle.gru_decode:
    the hoisted-projection restatement equals the literal loop within fp32 round-off (SURVEY App. B last item)."""
    ra = ref_shim.load("rnn_all")
    N, K, H = 32, 16, 64
    code = ref_shim.get_code("Polar", "polar", N, K)
    torch.manual_seed(11)
    net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0)
    dec = ra.RNN_decoder("y_input", N, code.info_inds, onehot=True)
    rs = np.random.RandomState(7)
    y = (_msgs(rs, 64, N) + 0.8 * rs.randn(64, N)).astype(np.float32)
    ref = dec.decode(net, False, torch.from_numpy(y)).numpy()
    got, logits = oracle.gru_decode(net.state_dict(), y, N, np.asarray(code.info_inds), H=H)
    got = np.asarray(got)
    logits = np.asarray(logits)
    # decisions may only differ behind a near-zero logit (sequential feedback divergence)
    bad = (got != ref).any(axis=1)
    near = (np.abs(logits) < 1e-4).any(axis=1)
    assert not (bad & ~near).any()
    assert bad.mean() < 0.05


def test_conv_forward_side_by_side():
    rm = ref_shim.load("models")
    import argparse
    torch.manual_seed(12)
    cfg = argparse.Namespace(embed_dim=128, max_len=64, N=64, dont_use_bias=False, dropout=0.0)
    net = rm.convNet(cfg)
    net.eval()
    rs = np.random.RandomState(8)
    y = (_msgs(rs, 32, 64) + 0.7 * rs.randn(32, 64)).astype(np.float32)
    with torch.no_grad():
        ref = net(torch.from_numpy(y), None, None, "cpu")[3].reshape(32, 64).numpy()
    got = np.asarray(oracle.conv_forward(net.state_dict(), y)).reshape(32, 64)
    np.testing.assert_allclose(got, ref, atol=2e-5, rtol=1e-5)
