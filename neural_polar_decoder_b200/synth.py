"""Deterministic synthetic parameters (no checkpoints ship with the reference and there is no network).

The arrays are drawn with numpy's RandomState -- not torch's RNG -- so the same seed gives the same
weights on every box and torch version; tests/golden/ stores reference outputs for these seeds.
Distributions follow PyTorch's default initialisers: nn.GRU / nn.Linear U(-1/sqrt(fan), 1/sqrt(fan))."""
from collections import OrderedDict

import numpy as np


def gru_state_dict(seed, N, H=512, layers=2, head_gain=1.0):
    """state_dict of RNN_Model('GRU', N+2, H, 1, layers, ...) (reference rnn_all.py:294-343, keys as in
    SURVEY.md a7) as numpy float32 arrays.  head_gain scales linear.weight so that |logit| ~ O(1)."""
    rs = np.random.RandomState(seed)
    k = 1.0 / np.sqrt(H)
    sd = OrderedDict()
    for l in range(layers):
        fan_in = N + 2 if l == 0 else H
        sd["rnn.weight_ih_l%d" % l] = rs.uniform(-k, k, (3 * H, fan_in)).astype(np.float32)
        sd["rnn.weight_hh_l%d" % l] = rs.uniform(-k, k, (3 * H, H)).astype(np.float32)
        sd["rnn.bias_ih_l%d" % l] = rs.uniform(-k, k, (3 * H,)).astype(np.float32)
        sd["rnn.bias_hh_l%d" % l] = rs.uniform(-k, k, (3 * H,)).astype(np.float32)
    sd["linear.weight"] = (head_gain * rs.uniform(-k, k, (1, H))).astype(np.float32)
    sd["linear.bias"] = rs.uniform(-k, k, (1,)).astype(np.float32)
    return sd


def gru_y_state_dict(seed, N, H, in_size, y_hidden, y_depth, y_out, head_gain=1.0):
    """state_dict of RNN_Model('GRU', in_size, H, 1, 2, N, y_hidden, y_depth, y_output_size=y_out) (reference
    rnn_all.py:1317 'y_h0': in_size = 1 + onehot, y_out = 2H; 1320 use_ynn: in_size = N + 1 + onehot, y_out = N):
    the GRU and head of gru_state_dict with a first layer of in_size columns, plus the y-MLP y_linears.{i}
    (rnn_all.py:323-329)."""
    sd = gru_state_dict(seed, in_size - 2, H, 2, head_gain)
    rs = np.random.RandomState(seed + 7919)
    dims = [(y_hidden, N)] + [(y_hidden, y_hidden)] * max(0, y_depth - 2) + [(y_out, y_hidden)]
    for i, (o, n_in) in enumerate(dims):
        k = 1.0 / np.sqrt(n_in)
        sd["y_linears.%d.weight" % i] = rs.uniform(-k, k, (o, n_in)).astype(np.float32)
        sd["y_linears.%d.bias" % i] = rs.uniform(-k, k, (o,)).astype(np.float32)
    return sd


def with_mlp_head(sd, seed, H, y_hidden, depth, head_gain=1.0):
    """Replace the Linear(H,1) head of a GRU state_dict by the nn.Sequential head of out_linear_depth = depth > 1
    (reference rnn_all.py:335-343): linear.{0,2,...}: Linear(H,Yh) SELU [Linear(Yh,Yh) SELU]* Linear(Yh,1)."""
    sd = OrderedDict((k, v) for k, v in sd.items() if not k.startswith("linear."))
    rs = np.random.RandomState(seed + 104729)
    dims = [(y_hidden, H)] + [(y_hidden, y_hidden)] * (depth - 2) + [(1, y_hidden)]
    for i, (o, n_in) in enumerate(dims):
        k = (head_gain if i == len(dims) - 1 else 2.0) / np.sqrt(n_in)
        sd["linear.%d.weight" % (2 * i)] = rs.uniform(-k, k, (o, n_in)).astype(np.float32)
        sd["linear.%d.bias" % (2 * i)] = rs.uniform(-k, k, (o,)).astype(np.float32)
    return sd


CONV_LAYERS = [  # (name, C_out, C_in) in state_dict order (reference models.py:701-730)
    ("layers1.0", 64, 1), ("layers1.2", 64, 64), ("layers2.0", 64, 64), ("layers2.2", 64, 64),
    ("layers3.0", 64, 64), ("layers3.2", 64, 64), ("layers4.0", 64, 64), ("layers4.2", 64, 64),
    ("layers5.0", 128, 64), ("layers5.2", 128, 128),
]


def conv_state_dict(seed, N=64, embed_dim=128):
    """state_dict of convNet(config) with embed_dim, max_len = N (reference models.py:692-740)."""
    rs = np.random.RandomState(seed)
    C = embed_dim // 2
    sd = OrderedDict()
    for name, co, ci in CONV_LAYERS:
        co = co * C // 64
        ci = ci if ci == 1 else ci * C // 64
        k = 1.0 / np.sqrt(ci * 7)
        sd[name + ".weight"] = rs.uniform(-k, k, (co, ci, 7)).astype(np.float32)
        sd[name + ".bias"] = rs.uniform(-k, k, (co,)).astype(np.float32)
    dims = [(4 * N, embed_dim * N), (N, 4 * N), (N, N)]
    for idx, (o, i) in zip((0, 2, 4), dims):
        k = 1.0 / np.sqrt(i)
        sd["layersFin.%d.weight" % idx] = rs.uniform(-k, k, (o, i)).astype(np.float32)
        sd["layersFin.%d.bias" % idx] = rs.uniform(-k, k, (o,)).astype(np.float32)
    sd["layer_norm.weight"] = (1.0 + 0.1 * rs.uniform(-1, 1, (N,))).astype(np.float32)
    sd["layer_norm.bias"] = (0.1 * rs.uniform(-1, 1, (N,))).astype(np.float32)
    return sd
