"""GPU tests of the host-buffer entry points (npd_*_host, include/npd.h): host tensors through the chunked
copy/decode/copy pipeline must give exactly what the device entry points give on the same inputs."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _auto_chunk_after_test():
    yield
    from neural_polar_decoder_b200 import _lib
    _lib.check(_lib.load().npd_host_set_chunk(0))


def _set_chunk(_monkeypatch, rows):
    """Force small pipeline chunks so that small batches still run several chunks through all three streams."""
    from neural_polar_decoder_b200 import _lib
    _lib.check(_lib.load().npd_host_set_chunk(int(rows)))


def _frames(code, B, snr, seed):
    g = torch.Generator().manual_seed(seed)
    msg = 1.0 - 2.0 * torch.randint(0, 2, (B, code.K), generator=g).float()
    x = code.encode_plotkin(msg.cuda()).cpu()
    y = x + 10 ** (-snr / 20) * torch.randn(B, code.N, generator=g)
    return msg, y


@pytest.mark.parametrize("n,K,B,chunk", [(6, 22, 1000, 256), (10, 512, 3000, 512), (8, 128, 777, 0), (5, 16, 1, 0),
                                         (12, 2048, 700, 256)])
@pytest.mark.parametrize("pinned", [False, True])
def test_sc_host_equals_device(n, K, B, chunk, pinned, monkeypatch):
    from neural_polar_decoder_b200 import PolarCode, construct
    N = 1 << n
    rs = construct.reference_rs256()
    code = PolarCode(n, K, None, rs=rs[rs < N]) if N <= 256 else PolarCode(n, K, None, F=construct.pw_frozen_set(N, K))
    _, y = _frames(code, B, 1.0, 5)
    y[0, :4] = 0.0  # exact ties take the flagged re-decode path inside a chunk
    if chunk:
        _set_chunk(monkeypatch, chunk)
    yh = y.pin_memory() if pinned else y
    llr_d, dec_d = code.sc_decode_new(y.cuda(), 1.0)
    llr_h, dec_h = code.sc_decode_new(yh, 1.0)
    assert not dec_h.is_cuda and dec_h.is_pinned() == pinned
    assert torch.equal(dec_h, dec_d.cpu()) and torch.equal(llr_h, llr_d.cpu())
    _, dec_h2 = code.sc_decode_new(yh, 1.0, return_llr=False)
    assert torch.equal(dec_h2, dec_d.cpu())
    gt = torch.sign(torch.randn(B, N))
    _, g_d = code.sc_decode_new(y.cuda(), 1.0, use_gt=gt.cuda())
    _, g_h = code.sc_decode_new(yh, 1.0, use_gt=gt)
    assert torch.equal(g_h, g_d.cpu())


def test_sc_host_empty():
    from neural_polar_decoder_b200 import PolarCode, construct
    rs = construct.reference_rs256()
    code = PolarCode(5, 16, None, rs=rs[rs < 32])
    llr, dec = code.sc_decode_new(torch.empty(0, 32), 0.0)
    assert llr.shape == (0, 32) and dec.shape == (0, 16)


def test_pac_host_equals_device(monkeypatch):
    from neural_polar_decoder_b200.pac_code import PAC
    _set_chunk(monkeypatch, 256)
    code = PAC(None, 32, 16, 53)
    g = torch.Generator().manual_seed(3)
    msg = 1.0 - 2.0 * torch.randint(0, 2, (900, 16), generator=g).float()
    y = code.pac_encode(msg.cuda()).cpu() + 0.8 * torch.randn(900, 32, generator=g)
    d = code.pac_sc_decode(y.cuda(), 1.0)
    h = code.pac_sc_decode(y, 1.0)
    for a, b in zip(h, d):
        assert not a.is_cuda and torch.equal(a, b.cpu())


def test_gru_host_equals_device(monkeypatch):
    from neural_polar_decoder_b200 import rnn_all, synth
    _set_chunk(monkeypatch, 128)
    N, K = 32, 16
    code = rnn_all.get_code('Polar', 'polar', N, K)
    sd = synth.gru_state_dict(4, N, 512, 2, head_gain=6.0)
    net = rnn_all.RNN_Model('GRU', N + 2, 512, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = rnn_all.RNN_decoder('y_input', N, code.info_inds, onehot=True)
    _, y = _frames(code, 500, 0.0, 9)
    d_dev, l_dev = dec.decode(net, False, y.cuda(), return_logits=True)
    d_host, l_host = dec.decode(net, False, y.pin_memory(), return_logits=True)
    assert not d_host.is_cuda
    assert torch.equal(d_host, d_dev.cpu()) and torch.equal(l_host, l_dev.cpu())


def test_conv_host_equals_device(monkeypatch):
    import argparse
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.models import convNet
    _set_chunk(monkeypatch, 128)
    net = convNet(argparse.Namespace(embed_dim=128, max_len=64, N=64, dont_use_bias=False, dropout=0.1))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.conv_state_dict(2, 64, 128).items()})
    net.eval()
    y = torch.randn(700, 64, generator=torch.Generator().manual_seed(1))
    out_d = net.forward(y.cuda(), None, None, torch.device("cuda"))
    out_h = net.forward(y, None, None, torch.device("cuda"))
    assert torch.equal(out_h[3], out_d[3].cpu()) and torch.equal(out_h[4], out_d[4].cpu())
    assert torch.equal(out_h[1], out_d[1].cpu())


def test_conv_decode_is_sign_of_forward():
    """convNet.decode (npd_conv_decode: the sign is taken in the kernel's epilogue) == sign of forward()'s logits,
    device and host paths."""
    import argparse
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.models import convNet
    net = convNet(argparse.Namespace(embed_dim=128, max_len=64, N=64, dont_use_bias=False, dropout=0.1))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.conv_state_dict(2, 64, 128).items()})
    net.eval()
    y = torch.randn(333, 64, generator=torch.Generator().manual_seed(4))
    logits = net.forward(y.cuda(), None, None, torch.device("cuda"))[3]
    bits_d, _ = net.decode(y.cuda(), None, None, torch.device("cuda"))
    bits_h, _ = net.decode(y, None, None, torch.device("cuda"))
    assert bits_d.shape == (333, 64, 1) and torch.equal(bits_d, logits.sign())
    assert not bits_h.is_cuda and torch.equal(bits_h, logits.sign().cpu())
