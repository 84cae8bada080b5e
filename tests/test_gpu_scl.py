"""GPU tests of the SC-list decoder (npd_scl_decode; reference polar.py:793-876)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import oracle  # noqa: E402

pytestmark = pytest.mark.gpu


def _code(N, K):
    from neural_polar_decoder_b200 import PolarCode, construct
    n = int(np.log2(N))
    rs = construct.reference_rs256()
    return PolarCode(n, K, None, rs=rs[rs < N]) if N <= 256 else PolarCode(n, K, None, F=construct.pw_frozen_set(N, K))


def test_scl_reference_fixtures_bit_exact(golden):
    """Chosen path's decisions and leaf LLRs against the live reference, list sizes 1..32."""
    g = golden("scl")
    for nm in [str(s) for s in g["names"]]:
        N, K, L = [int(v[1:]) if v.startswith("L") else int(v) for v in nm.split("_")[1:]]
        code = _code(N, K)
        assert np.array_equal(code.info_positions, g[nm + "_info"])
        y = torch.from_numpy(g[nm + "_y"]).cuda()
        llr, dec = code.scl_decode(y, float(g[nm + "_snr"]), L)
        assert np.array_equal(dec.cpu().numpy(), g[nm + "_dec"]), nm
        assert np.array_equal(llr.cpu().numpy(), g[nm + "_llr"]), nm
        # host tensors go through the chunked pipeline and give the same result
        llr_h, dec_h = code.scl_decode(torch.from_numpy(g[nm + "_y"]), float(g[nm + "_snr"]), L)
        assert not dec_h.is_cuda and np.array_equal(dec_h.numpy(), g[nm + "_dec"]) and np.array_equal(llr_h.numpy(), g[nm + "_llr"])


@pytest.mark.parametrize("N,K,L,B,snr", [(64, 22, 4, 3000, 0.0), (128, 64, 8, 600, 1.0), (32, 16, 3, 2000, 0.0),
                                         (64, 32, 32, 300, 1.0), (512, 256, 4, 100, 2.0), (1024, 512, 2, 40, 2.0)])
def test_scl_vs_oracle_seeded(N, K, L, B, snr):
    """Seeded AWGN frames against the C restatement (itself pinned to the reference), incl. a non-power-of-two list."""
    code = _code(N, K)
    n = int(np.log2(N))
    r = np.random.RandomState(N + L)
    msg = (1.0 - 2.0 * r.randint(0, 2, size=(B, K))).astype(np.float32)
    x = oracle.polar_encode(msg, n, code.info_positions)
    y = (x + 10 ** (-snr / 20) * r.randn(B, N)).astype(np.float32)
    llr_o, dec_o = oracle.scl_decode(y, snr, n, code.info_positions, L)
    llr, dec = code.scl_decode(torch.from_numpy(y).cuda(), snr, L)
    assert np.array_equal(dec.cpu().numpy(), dec_o)
    assert np.array_equal(llr.cpu().numpy(), llr_o)
    _, dec2 = code.scl_decode(torch.from_numpy(y).cuda(), snr, L, return_llr=False)
    assert np.array_equal(dec2.cpu().numpy(), dec_o)


def test_scl_list_of_one_is_sc_and_lists_help():
    code = _code(64, 22)
    g = torch.Generator().manual_seed(7)
    B = 20000
    msg = 1.0 - 2.0 * torch.randint(0, 2, (B, 22), generator=g).float()
    x = code.encode_plotkin(msg.cuda())
    y = x + 10 ** (0.0 / 20) * torch.randn(B, 64, generator=g).cuda()
    _, sc = code.sc_decode_new(y, 0.0, return_llr=False)
    _, l1 = code.scl_decode(y, 0.0, 1, return_llr=False)
    assert torch.equal(sc, l1)
    bler = [float((code.scl_decode(y, 0.0, L, return_llr=False)[1] != msg.cuda()).any(1).float().mean()) for L in (1, 4, 16)]
    assert bler[0] > bler[1] > bler[2] > 0, bler  # the ML pick among a longer list decodes more frames


def test_scl_envelope_and_errors():
    from neural_polar_decoder_b200 import _lib
    code = _code(64, 22)
    y = torch.zeros(4, 64).cuda()
    with pytest.raises(_lib.NpdError):
        code.scl_decode(y, 0.0, 64)       # list size > 32
    with pytest.raises(NotImplementedError):
        code.scl_decode(y, 0.0, 4, use_CRC=True)
    big = _code(1024, 512)
    with pytest.raises(_lib.NpdError):
        big.scl_decode(torch.zeros(2, 1024).cuda(), 0.0, 32)  # state does not fit in shared memory
    llr, dec = code.scl_decode(torch.empty(0, 64).cuda(), 0.0, 4)
    assert llr.shape == (0, 64) and dec.shape == (0, 22)


def test_sweeps_run_the_list_decoder():
    """polar_RNN_full_test(run_SCL=True) / testXformer fill the SCL lists like the reference does."""
    from neural_polar_decoder_b200 import rnn_all, sweep, synth
    N, K = 32, 16
    code = rnn_all.get_code('Polar', 'polar', N, K)
    net = rnn_all.RNN_Model('GRU', N + 2, 128, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.gru_state_dict(5, N, 128, 2, head_gain=6.0).items()})
    dec = rnn_all.RNN_decoder('y_input', N, code.info_inds, onehot=True)
    msgs = 1.0 - 2.0 * torch.randint(0, 2, (4000, K), generator=torch.Generator().manual_seed(3)).float()
    loader = torch.utils.data.DataLoader(msgs, batch_size=2000, shuffle=False)
    out = sweep.polar_RNN_full_test(net, code, [0.0, 2.0], loader, False, True, False, decoder=dec, list_size=4, seed=1)
    bers_sc, blers_sc, bers_scl, blers_scl = out[2], out[3], out[4], out[5]
    assert all(0 < l <= s for l, s in zip(blers_scl, blers_sc)) and blers_scl[1] < blers_scl[0]
    out0 = sweep.polar_RNN_full_test(net, code, [0.0, 2.0], loader, False, False, False, decoder=dec, seed=1)
    assert out0[4] == [0.0, 0.0] and out0[2] == bers_sc


@pytest.mark.parametrize("N,K", [(16, 16), (16, 1), (8, 0), (2, 1)])
def test_scl_degenerate_codes(N, K):
    """All-information, single-information-bit, no-information and two-bit codes against the oracle."""
    from neural_polar_decoder_b200 import PolarCode
    n = int(np.log2(N))
    frozen = np.arange(N - K)  # the K last positions carry information
    code = PolarCode(n, K, None, F=frozen)
    r = np.random.RandomState(N * 7 + K)
    y = (r.choice([-1.0, 1.0], size=(300, N)) + 0.9 * r.randn(300, N)).astype(np.float32)
    for L in (1, 4):
        llr_o, dec_o = oracle.scl_decode(y, 1.0, n, code.info_positions, L)
        llr, dec = code.scl_decode(torch.from_numpy(y).cuda(), 1.0, L)
        assert dec.shape == (300, K)
        assert np.array_equal(dec.cpu().numpy(), dec_o) and np.array_equal(llr.cpu().numpy(), llr_o)
