"""GPU: parity at BASELINE.json's FULL batch sizes through size-independent properties of the domain (the oracle finishes only
small cases; tests/test_gpu_parity.py covers those bit for bit):

  * noiseless round trip: decode(encode(m)) == m for every frame of a full launch (SC at N = 256 / 1024 / 4096, PAC(32,16));
  * linearity of the encoder over GF(2): enc(m1 xor m2) = enc(m1) * enc(m2) in the BPSK domain;
  * symmetry of min-sum SC for a linear code: multiplying y by any codeword x(m') multiplies the decisions by m'
    (f and g only move signs; exact in floating point), with real noise;
  * row independence of the neural decoders: decisions and logits of a codeword do not depend on the batch it travels in
    (split / permute the full launch);
  * a checksum of checksums: the fused sweep's counters over 2^20 frames equal the sum over disjoint sub-ranges."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _code(N, K):
    from neural_polar_decoder_b200 import PolarCode, construct
    n = int(np.log2(N))
    if N <= 256:
        rs = construct.reference_rs256()
        return PolarCode(n, K, None, rs=rs[rs < N])
    return PolarCode(n, K, None, F=construct.pw_frozen_set(N, K))


def _msgs(B, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return 1.0 - 2.0 * torch.randint(0, 2, (B, K), generator=g, device="cuda").float()


@pytest.mark.parametrize("N,K,B", [(256, 128, 524288), (1024, 512, 131072), (4096, 2048, 32768)])
def test_sc_noiseless_round_trip_full_batch(N, K, B):
    code = _code(N, K)
    msg = _msgs(B, K, N)
    x = code.encode_plotkin(msg)
    assert x.abs().eq(1).all()
    _, dec = code.sc_decode_new(x, 3.0, return_llr=False)   # the bench's throughput path (quad / split kernels)
    assert torch.equal(dec, msg)
    if N <= 1024:
        llr, dec2 = code.sc_decode_new(x[:4096], 3.0)        # leaf-LLR path (lane kernel)
        assert torch.equal(dec2, msg[:4096])
        fr = torch.as_tensor(code.frozen_positions, device="cuda")
        assert (llr[:, fr] > 999.0).all()                    # the frozen prior is added, not forced (polar.py:471-472)


def test_pac_noiseless_round_trip_full_batch():
    from neural_polar_decoder_b200 import PAC
    pac = PAC(None, 32, 16, 53)
    msg = _msgs(1 << 20, 16, 5)
    x = pac.pac_encode(msg)
    _, v, u = pac.pac_sc_decode(x, 3.0)
    assert torch.equal(v, msg)
    assert torch.equal(pac.pac_encode(v), x)


@pytest.mark.parametrize("N,K,B", [(1024, 512, 131072), (64, 22, 1 << 20)])
def test_encoder_linearity_full_batch(N, K, B):
    code = _code(N, K)
    m1, m2 = _msgs(B, K, 1), _msgs(B, K, 2)
    assert torch.equal(code.encode_plotkin(m1 * m2), code.encode_plotkin(m1) * code.encode_plotkin(m2))
    assert torch.equal(code.encode_plotkin(torch.ones(8, K, device="cuda")), torch.ones(8, N, device="cuda"))


@pytest.mark.parametrize("N,K,B,snr", [(1024, 512, 131072, 2.0), (4096, 2048, 32768, 2.0), (256, 128, 262144, 1.0)])
def test_sc_codeword_symmetry_full_batch(N, K, B, snr):
    """dec(y * x(m')) == dec(y) * m' on real noise: every f / g of the min-sum recursion only moves signs."""
    code = _code(N, K)
    msg, flip = _msgs(B, K, 3), _msgs(B, K, 4)
    y = code.channel(code.encode_plotkin(msg), snr, point=7, seed=11)
    _, d1 = code.sc_decode_new(y, snr, return_llr=False)
    _, d2 = code.sc_decode_new(y * code.encode_plotkin(flip), snr, return_llr=False)
    assert torch.equal(d2, d1 * flip)
    assert 0.0 < (d1 != msg).any(dim=1).float().mean().item() < 0.9   # real block errors were in play


def test_gru_row_independence_full_launch():
    """37888 codewords (the bench's launch): decisions AND logits of a codeword are the same bits whether it is decoded in the
    full launch, in the reference's batch of 10000, or in a permuted launch."""
    import os
    from conftest import GOLDEN
    from neural_polar_decoder_b200 import cli, synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, gru_decode
    N, K, B = 64, 22, 37888
    code = _code(N, K)
    path = os.path.join(GOLDEN, "crisp_gru_N64_K22_H512.pt")
    if os.path.exists(path):
        net, _, _ = cli.net_from_checkpoint(path)
    else:
        net = RNN_Model('GRU', N + 2, 512, 1, 2, N, 0, 0)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.gru_state_dict(11, N, 512, 2, head_gain=8.0).items()})
    dec = RNN_decoder('y_input', N, code.info_positions, onehot=True)
    lc = dec._loss_code(dec.info_inds)
    y = code.channel(code.encode_plotkin(_msgs(B, K, 9)), 0.0, point=3, seed=5)
    d, lg = gru_decode(net, lc, y, want_logits=True)
    d2, lg2 = gru_decode(net, lc, y[:10000].contiguous(), want_logits=True)
    assert torch.equal(d[:10000], d2) and torch.equal(lg[:10000], lg2)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    d3, lg3 = gru_decode(net, lc, y[perm].contiguous(), want_logits=True)
    assert torch.equal(d3, d[perm]) and torch.equal(lg3, lg[perm])
    assert torch.equal(dec.decode(net, False, y), d)   # the logits output does not change the decisions


def test_conv_row_independence_full_launch():
    import argparse
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.models import convNet
    N, B = 64, 65536
    net = convNet(argparse.Namespace(embed_dim=128, max_len=N, N=N, dont_use_bias=False, dropout=0.0))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.conv_state_dict(4, N, 128).items()})
    net.eval()
    code = _code(N, 22)
    y = code.channel(code.encode_plotkin(_msgs(B, 22, 2)), 0.0, point=1, seed=2)
    lg = net.logits(y)
    assert torch.equal(net.logits(y[:1000].contiguous()), lg[:1000])
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    assert torch.equal(net.logits(y[perm].contiguous()), lg[perm])


def test_sweep_checksum_of_checksums():
    from neural_polar_decoder_b200.sweep import mc_sc_sweep
    code = _code(1024, 512)
    total = 1 << 20
    whole = mc_sc_sweep(code, [2.0, 2.5], total, chunk=1 << 17, seed=77, rank=0, world=1)[3]
    parts = [mc_sc_sweep(code, [2.0, 2.5], total, chunk=50000, seed=77, rank=r, world=5)[3] for r in range(5)]
    assert torch.equal(whole, sum(parts[1:], parts[0]))
    assert whole[:, 2].tolist() == [total, total] and (whole[0, 1] > whole[1, 1] > 0)
