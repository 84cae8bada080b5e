"""Drop-in for the hot-path half of the reference's rnn_all.py / rnn.py: RNN_Model (parameter
container with the reference's state_dict keys), RNN_decoder.decode (test branch, 'y_input', onehot) and
get_code.  The N-step autoregressive decode is ONE libnpd.so launch (csrc/gru_decode.cu).

Reference: rnn_all.py:258-260 (get_onehot), 294-398 (RNN_Model), 400-561 (RNN_decoder),
1015-1196 (get_code).  Out of scope: training branches, y_h0 / y_h0_out conditioning,
LSTM / bidirectional / LayerNorm variants, the list decoder."""
import ctypes
import random

import numpy as np
import torch
import torch.nn as nn

from . import _lib, construct
from .pac_code import PAC
from .polar import PolarCode


def get_onehot(actions):
    """reference rnn_all.py:258-260."""
    inds = (0.5 + 0.5 * actions).long()
    return torch.eye(2, device=inds.device)[inds].reshape(actions.shape[0], -1)


class RNN_Model(nn.Module):
    """Same constructor signature and parameter names as the reference (rnn_all.py:294-343) so that
    reference checkpoints load with load_state_dict.  forward() keeps the reference semantics through
    torch's own nn.GRU (used for single steps / training by callers); the Monte-Carlo decode path does
    not go through forward() -- RNN_decoder.decode hands the weights to the fused kernel."""

    def __init__(self, rnn_type, input_size, feature_size, output_size, num_rnn_layers, y_size, y_hidden_size,
                 y_depth, activation='relu', dropout=0., skip=False, out_linear_depth=1, y_output_size=None,
                 bidirectional=False, use_layernorm=False):
        super().__init__()
        assert rnn_type in ['GRU', 'LSTM']
        self.rnn_type = rnn_type
        self.input_size = input_size
        self.feature_size = feature_size
        self.output_size = output_size
        self.num_rnn_layers = num_rnn_layers
        self.bidirectional = bidirectional
        self.y_size, self.y_hidden_size, self.y_depth = y_size, y_hidden_size, y_depth
        self.out_linear_depth = out_linear_depth
        self.activation, self.dropout, self.skip = activation, dropout, skip
        self.use_layernorm = use_layernorm
        self.rnn = getattr(nn, rnn_type)(input_size, feature_size, num_rnn_layers, bidirectional=bidirectional,
                                         batch_first=True)
        self.drop = nn.Dropout(dropout)
        self.layernorm = nn.LayerNorm(feature_size) if use_layernorm else nn.Identity()
        if out_linear_depth != 1 or (y_hidden_size > 0 and y_depth > 0):
            raise NotImplementedError("only the run_crisp.sh configuration (y_input: y_depth=0, "
                                      "out_linear_depth=1) is on the accelerated path")
        self.linear = nn.Linear((int(bidirectional) + 1) * feature_size, output_size)
        self._npd = None  # (key, handle)

    def forward(self, input, hidden, Fy=None):
        out, hidden = self.rnn(input, hidden)
        out = self.layernorm(self.drop(out))
        decoded = self.linear(out if Fy is None else torch.cat([Fy, out], -1))
        return decoded.view(-1, self.output_size), hidden

    # ---- libnpd handle (weights repacked once per parameter version) ----
    def _supported(self):
        return (self.rnn_type == 'GRU' and self.num_rnn_layers == 2 and not self.bidirectional and
                not self.use_layernorm and self.output_size == 1 and self.out_linear_depth == 1)

    def npd_handle(self, N):
        if not self._supported():
            raise NotImplementedError("fused decode supports GRU, 2 layers, unidirectional, no LayerNorm, 1 output")
        sd = self.state_dict()
        key = (N, torch.cuda.current_device()) + tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if self._npd is not None and self._npd[0] == key:
            return self._npd[1]
        handle = GruHandle(N, self.feature_size, sd)
        self._npd = (key, handle)
        return handle


class GruHandle:
    """Owns an npd_gru_t (fp16 weight tile streams in HBM)."""

    def __init__(self, N, H, sd):
        _lib.require_cuda()
        lib = _lib.load()

        def host(name, shape):
            t = sd[name]
            t = torch.as_tensor(t).detach().float().cpu().contiguous()
            assert tuple(t.shape) == tuple(shape), (name, tuple(t.shape), shape)
            return t

        self._keep = [
            host("rnn.weight_ih_l0", (3 * H, N + 2)), host("rnn.weight_hh_l0", (3 * H, H)),
            host("rnn.bias_ih_l0", (3 * H,)), host("rnn.bias_hh_l0", (3 * H,)),
            host("rnn.weight_ih_l1", (3 * H, H)), host("rnn.weight_hh_l1", (3 * H, H)),
            host("rnn.bias_ih_l1", (3 * H,)), host("rnn.bias_hh_l1", (3 * H,)),
            host("linear.weight", (1, H)), host("linear.bias", (1,)),
        ]
        h = ctypes.c_void_p()
        _lib.check(lib.npd_gru_create(int(N), int(H), *[ctypes.c_void_p(t.data_ptr()) for t in self._keep],
                                      ctypes.byref(h)))
        self.h, self.N, self.H = h, N, H
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "h", None) and _lib._lib is not None:
                _lib._lib.npd_gru_destroy(self.h)
                self.h = None
        except Exception:
            pass


def gru_decode(net_or_handle, code_handle, y, forced=None, want_logits=False, genie=None):
    """One fused launch: y [B,N] (device) -> (decoded [B,N], logits [B,N] or None)."""
    B, N = y.shape
    handle = net_or_handle if isinstance(net_or_handle, GruHandle) else net_or_handle.npd_handle(N)
    decoded = torch.empty(B, N, dtype=torch.float32, device=y.device)
    logits = torch.empty(B, N, dtype=torch.float32, device=y.device) if want_logits else None
    if B > 0:
        _lib.check(_lib.load().npd_gru_decode(handle.h, code_handle.h, _lib.ptr(y), _lib.ptr(forced), _lib.ptr(genie),
                                              _lib.ptr(logits), _lib.ptr(decoded), B, None, 0, _lib.stream_ptr()))
    return decoded, logits


def gru_decode_host(net_or_handle, code_handle, y, forced=None, want_logits=False, genie=None):
    """Host tensors in / out through the library's chunked copy/decode/copy pipeline (npd_gru_decode_host)."""
    _lib.require_cuda()
    yh = _lib.host_f32(y)
    B, N = yh.shape
    handle = net_or_handle if isinstance(net_or_handle, GruHandle) else net_or_handle.npd_handle(N)
    forced = None if forced is None else _lib.host_f32(forced.cpu())
    genie = None if genie is None else _lib.host_f32(genie.cpu())
    decoded = _lib.host_out((B, N), yh)
    logits = _lib.host_out((B, N), yh) if want_logits else None
    if B > 0:
        _lib.check(_lib.load().npd_gru_decode_host(handle.h, code_handle.h, _lib.hptr(yh), _lib.hptr(forced),
                                                   _lib.hptr(genie), _lib.hptr(logits), _lib.hptr(decoded), B))
    return decoded, logits


class RNN_decoder:
    """reference rnn_all.py:400-561."""

    def __init__(self, decoding_type, N, info_inds, onehot=False, reverse_order=False):
        self.decoding_type = decoding_type
        self.N = N
        self.info_inds = info_inds
        self.onehot = onehot
        self.reverse_order = reverse_order
        self._codes = {}

    def _loss_code(self, loss_inds):
        """libnpd code object whose 'info' set is the loss-position set (only its bit mask is used)."""
        inds = np.sort(np.asarray(loss_inds, dtype=np.int64))
        key = (torch.cuda.current_device(), inds.tobytes())
        h = self._codes.get(key)
        if h is None:
            h = _lib.CodeHandle(int(np.log2(self.N)), inds)
            self._codes[key] = h
        return h

    def decode(self, net, train, y, gt=None, teacher_forcing_ratio=0., loss_inds=None, return_logits=False):
        """reference rnn_all.py:408-561 for decoding_type 'y_input' with one-hot feedback; all N autoregressive
        steps are one kernel launch.

        train=False (rnn_all.py:514-561): hidden state from zero, decisions sign(logit) on loss_inds (default: the
        info positions), +1 elsewhere.  gt = genie tensor [B,N]: decoded starts as gt.clone(), so positions outside
        loss_inds keep and feed back their genie value (rnn_all.py:519-522; the `loss_inds=code.loss_inds` call of
        polar_RNN_full_test, 887).
        train=True is served for EVALUATION only (no autograd graph; the reference's test_model(tf=True) calls it
        under torch.no_grad(), rnn_all.py:982-984): teacher forcing (425-461) returns the raw outputs of all N
        steps with gt fed back; student forcing (462-512) returns raw outputs on the info positions, 1 elsewhere.
        Training itself (gradients) is out of scope of the B200 path."""
        if self.decoding_type != 'y_input' or not self.onehot or self.reverse_order:
            raise NotImplementedError("accelerated path: decoding_type='y_input', onehot=True, forward order")
        if getattr(net, "y_depth", 0) != 0:
            raise NotImplementedError("y_input with a y-MLP (y_depth > 0) is out of scope")
        on_host = torch.is_tensor(y) and not y.is_cuda
        run = gru_decode_host if on_host else gru_decode
        yd = _lib.host_f32(y) if on_host else _lib.to_device_f32(y)
        assert yd.dim() == 2 and yd.shape[1] == self.N
        dev_ctx = torch.cuda.device(torch.cuda.current_device() if on_host else yd.device)

        def like_y(t):
            return None if t is None else (_lib.host_f32(t.cpu()) if on_host else _lib.to_device_f32(t, yd.device))

        if train:
            if torch.is_grad_enabled() and any(p.requires_grad for p in net.parameters()):
                raise NotImplementedError("training (rnn_all.py:422-512 with gradients) is out of scope of the B200 "
                                          "path; call under torch.no_grad() for teacher-/student-forced evaluation")
            with dev_ctx:
                if random.random() < teacher_forcing_ratio:  # rnn_all.py:425
                    assert gt is not None and gt.shape[1] == self.N
                    _, logits = run(net, self._loss_code(self.info_inds), yd, forced=like_y(gt), want_logits=True)
                    return logits
                decoded, logits = run(net, self._loss_code(self.info_inds), yd, want_logits=True)
                mask = torch.zeros(self.N, dtype=torch.bool, device=logits.device)
                mask[torch.as_tensor(np.asarray(self.info_inds), device=logits.device)] = True
                return torch.where(mask.unsqueeze(0), logits, torch.ones_like(logits))
        if loss_inds is None:
            loss_inds = self.info_inds
        with dev_ctx:
            decoded, logits = run(net, self._loss_code(loss_inds), yd, want_logits=return_logits, genie=like_y(gt))
        return (decoded, logits) if return_logits else decoded


def get_code(code_type, rate_profile, N, K, g=None, args=None):
    """reference rnn_all.py:1015-1196 for the rate profiles that need no data files."""
    n = int(np.log2(N))
    target_K = getattr(args, "target_K", None) or K
    if code_type == 'PAC':
        code = PAC(args, N, K, g, rate_profile=rate_profile)
        code.info_inds = code.B
        code.frozen_inds = np.array(sorted(set(range(N)) - set(code.B.tolist())))
        code.encode = code.pac_encode
        return code
    rs = construct.reference_rs256()
    rs = rs[rs < N]
    if rate_profile == 'polar':
        code = PolarCode(n, K, args, rs=rs)
    elif rate_profile == 'RM':
        rmweight = np.array([construct.count_set_bits(i) for i in range(N)])
        code = PolarCode(n, K, args, F=np.sort(np.argsort(rmweight)[:-K]))
    elif rate_profile == 'sorted':
        first = np.sort(rs[:target_K].copy())
        rs[:target_K] = first[::-1]
        code = PolarCode(n, K, args, rs=rs)
    elif rate_profile == 'sorted_last':
        first = np.sort(rs[:target_K].copy())
        rs[:target_K] = first
        code = PolarCode(n, K, args, rs=rs)
    elif rate_profile == 'rev_polar':
        first = rs[:target_K].copy()
        rs[:target_K] = first[::-1]
        code = PolarCode(n, K, args, rs=rs)
    else:
        raise NotImplementedError("rate_profile %r" % (rate_profile,))
    code.info_inds = code.info_positions
    code.frozen_inds = code.frozen_positions
    code.rate_profile = rate_profile
    code.encode = code.encode_plotkin
    code.msg_indices = np.arange(K)
    return code


if __name__ == '__main__':
    # `python -m neural_polar_decoder_b200.rnn_all <run_crisp.sh flags> --test` (reference rnn_all.py:1198-1949,
    # evaluation half; see cli.py)
    import sys
    from .cli import main
    sys.exit(main())
