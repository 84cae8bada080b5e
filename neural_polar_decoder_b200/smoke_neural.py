"""Neural-decoder half of __graft_entry__.smoke(): one small GRU decode and one small convNet forward on
cuda:0, each checked against the fp32 oracle (oracle/ is test infrastructure; only smoke() hands it in)."""
import argparse

import numpy as np
import torch

from . import construct, synth
from .models import convNet
from .rnn_all import RNN_Model, RNN_decoder, gru_decode


def run(oracle):
    rng = np.random.RandomState(0)
    # ---- CRISP GRU: the flagship shape Polar(64,22), H = 512 (CTA-pair kernel), then Polar(32,16), H = 256, ragged ----
    for N, K, H, B in ((64, 22, 512, 192), (32, 16, 256, 100)):
        _gru(oracle, rng, N, K, H, B)
    _conv(oracle, rng)


def _gru(oracle, rng, N, K, H, B):
    sd = synth.gru_state_dict(3, N, H, 2, head_gain=6.0)
    net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    rs = construct.reference_rs256()
    info = np.sort(rs[rs < N][:K])
    dec = RNN_decoder('y_input', N, info, onehot=True)
    y = (rng.choice([-1.0, 1.0], size=(B, N)) + 0.8 * rng.randn(B, N)).astype(np.float32)
    do, lo = oracle.gru_decode(sd, y, N, info)
    d, lg = gru_decode(net, dec._loss_code(info), torch.from_numpy(y).cuda(), forced=torch.from_numpy(do).cuda(),
                       want_logits=True)
    err = np.abs(lg.cpu().numpy() - lo)
    tol = 1e-2 * np.abs(lo) + 2e-3  # north_star: 1e-2 relative; SURVEY 7: + 2e-3 absolute for near-zero logits
    assert (err <= tol).all(), "GRU logits differ from the oracle: max err %.3e" % err.max()
    print("smoke: GRU(2x%d) Polar(%d,%d) logits within tolerance on %d frames (max err %.2e)" % (H, N, K, B, err.max()))


def _conv(oracle, rng):
    # ---- convNet, Polar(64,22) ----
    N, E, B = 64, 128, 50
    sd = synth.conv_state_dict(4, N, E)
    cnet = convNet(argparse.Namespace(embed_dim=E, max_len=N, N=N, dont_use_bias=False, dropout=0.0))
    cnet.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    cnet.eval()
    y = (rng.choice([-1.0, 1.0], size=(B, N)) + 0.9 * rng.randn(B, N)).astype(np.float32)
    ref = oracle.conv_forward(sd, y)
    lg = cnet.logits(torch.from_numpy(y).cuda()).cpu().numpy()
    err = np.abs(lg - ref)
    assert (err <= 1e-2 * np.abs(ref) + 2e-3).all(), "convNet logits differ from the oracle: max err %.3e" % err.max()
    print("smoke: convNet Polar(%d,22) logits within tolerance on %d frames (max err %.2e)" % (N, B, err.max()))
