"""Drop-in for the hot-path half of the reference's rnn_all.py / rnn.py: RNN_Model (parameter
container with the reference's state_dict keys), RNN_decoder.decode ('y_input' and 'y_h0') and
get_code.  The N-step autoregressive decode is ONE libnpd.so launch (csrc/gru_decode.cu).

Reference: rnn_all.py:258-260 (get_onehot), 294-398 (RNN_Model), 400-561 (RNN_decoder),
1015-1196 (get_code).  MLP heads (out_linear_depth > 1), 'y_h0', use_ynn, scalar feedback and reverse order are
covered by the same kernel; the training iteration (gradients) lives in train.py / csrc/gru_train.cu.  Out of scope
(SURVEY.md 2): LSTM / bidirectional / LayerNorm variants, the GRU list decoder."""
import ctypes
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

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
        D = int(bidirectional) + 1
        self.y_output_size = D * num_rnn_layers * feature_size if y_output_size is None else y_output_size
        if y_hidden_size > 0 and y_depth > 0:  # the y-MLP of 'y_h0' / use_ynn (rnn_all.py:323-329), same parameter names
            self.y_linears = nn.ModuleList([nn.Linear(y_size, y_hidden_size, bias=True)])
            self.y_linears.extend([nn.Linear(y_hidden_size, y_hidden_size, bias=True) for _ in range(1, y_depth - 1)])
            self.y_linears.append(nn.Linear(y_hidden_size, self.y_output_size - (y_size if skip else 0), bias=True))
        if out_linear_depth == 1:
            self.linear = nn.Linear(D * feature_size, output_size)
        else:  # rnn_all.py:335-343 (MLP head: npd_gru_set_head_mlp)
            layers = [nn.Linear(D * feature_size, y_hidden_size)]
            for _ in range(1, out_linear_depth - 1):
                layers += [nn.SELU(), nn.Linear(y_hidden_size, y_hidden_size)]
            layers += [nn.SELU(), nn.Linear(y_hidden_size, output_size)]
            self.linear = nn.Sequential(*layers)
        self._npd = None  # (key, handle)

    def act(self, inputs):
        """rnn_all.py:346-360 (unknown names fall through to identity there too)."""
        fn = {'tanh': torch.tanh, 'elu': F.elu, 'relu': F.relu, 'selu': F.selu, 'sigmoid': torch.sigmoid}
        return fn.get(self.activation, lambda x: x)(inputs)

    def get_Fy(self, y):
        """rnn_all.py:377-384: the y-MLP.  The reference skips the activation only for `ii == y_depth`, which with
        y_depth layers (y_depth >= 2) never happens -- the last layer is activated too -- and with y_depth == 1 (two
        layers) hits the second one.  A small dense MLP run once per batch: torch (cuBLAS) on y's device."""
        x = y
        for ii, layer in enumerate(self.y_linears):
            x = F.linear(x, layer.weight.to(y.device), layer.bias.to(y.device))
            if ii != self.y_depth:
                x = self.act(x)
        return x

    def get_h0(self, y):
        """rnn_all.py:362-375: y-MLP output [B, L*H] read as [B, H, L] and permuted to [L, B, H]."""
        x = self.get_Fy(y)
        if self.skip:
            x = torch.cat([y, x], 1)
        x = x.reshape(-1, self.feature_size, (int(self.bidirectional) + 1) * self.num_rnn_layers).permute(2, 0, 1)
        x = x.contiguous()
        return x if self.rnn_type == 'GRU' else (x, x)

    def forward(self, input, hidden, Fy=None):
        out, hidden = self.rnn(input, hidden)
        out = self.layernorm(self.drop(out))
        decoded = self.linear(out if Fy is None else torch.cat([Fy, out], -1))
        return decoded.view(-1, self.output_size), hidden

    # ---- libnpd handle (weights repacked once per parameter version) ----
    def _supported(self):
        return (self.rnn_type == 'GRU' and self.num_rnn_layers == 2 and not self.bidirectional and
                not self.use_layernorm and self.output_size == 1 and self.out_linear_depth >= 1)

    def npd_handle(self, N, onehot=True, y_in=True):
        if not self._supported():
            raise NotImplementedError("fused decode supports GRU, 2 layers, unidirectional, no LayerNorm, 1 output")
        sd = self.state_dict()
        key = (N, onehot, y_in, _GRU_PRECISION, torch.cuda.current_device()) + tuple((k, v._version, v.data_ptr()) for k, v in sd.items())
        if self._npd is not None and self._npd[0] == key:
            return self._npd[1]
        handle = GruHandle(N, self.feature_size, sd, onehot=onehot, y_in=y_in, head_depth=self.out_linear_depth,
                           y_hidden=self.y_hidden_size)
        self._npd = (key, handle)
        return handle


_GRU_PRECISION = "exact"


def set_gru_precision(mode):
    """'exact' (default): the CTA-pair kernel keeps the fp16 rounding residual of the recurrent state (include/npd.h,
    NPD_GRU_OPT_RESIDUAL_STATE) -- logits of trained checkpoints within 1e-2 |ref| + 2e-3 of the reference;
    'fast': fp16-only state, ~14 % faster, logit error up to ~3 x larger.  Applies to handles created afterwards."""
    global _GRU_PRECISION
    assert mode in ("exact", "fast")
    _GRU_PRECISION = mode


class GruHandle:
    """Owns an npd_gru_t (fp16 weight tile streams in HBM)."""

    def __init__(self, N, H, sd, onehot=True, y_in=True, head_depth=1, y_hidden=0):
        """onehot / y_in describe the columns of rnn.weight_ih_l0: [y (N, when y_in) | feedback (2 one-hot, else 1)].
        The library takes the run_crisp.sh form [3H, N+2]; the other forms are widened to it exactly: absent y
        columns are zeros ('y_h0', rnn_all.py:526-528), a scalar +-1 feedback column w is the one-hot pair [-w, +w]."""
        _lib.require_cuda()
        lib = _lib.load()

        def host(name, shape):
            t = sd[name]
            t = torch.as_tensor(t).detach().float().cpu().contiguous()
            assert tuple(t.shape) == tuple(shape), (name, tuple(t.shape), shape)
            return t

        w_in = host("rnn.weight_ih_l0", (3 * H, (N if y_in else 0) + (2 if onehot else 1)))
        wy = w_in[:, :N] if y_in else torch.zeros(3 * H, N)
        fb = w_in[:, -2:] if onehot else torch.cat([-w_in[:, -1:], w_in[:, -1:]], 1)
        self._keep = [
            torch.cat([wy, fb], 1).contiguous(), host("rnn.weight_hh_l0", (3 * H, H)),
            host("rnn.bias_ih_l0", (3 * H,)), host("rnn.bias_hh_l0", (3 * H,)),
            host("rnn.weight_ih_l1", (3 * H, H)), host("rnn.weight_hh_l1", (3 * H, H)),
            host("rnn.bias_ih_l1", (3 * H,)), host("rnn.bias_hh_l1", (3 * H,)),
        ]
        mlp = None
        if head_depth == 1:
            self._keep += [host("linear.weight", (1, H)), host("linear.bias", (1,))]
        else:  # nn.Sequential head (rnn_all.py:335-343): Linear at indices 0, 2, 4, ... with SELU in between
            self._keep += [torch.zeros(1, H), torch.zeros(1)]
            dims = [(y_hidden, H)] + [(y_hidden, y_hidden)] * (head_depth - 2) + [(1, y_hidden)]
            mlp = torch.cat([t.reshape(-1) for i, (o, n_in) in enumerate(dims)
                             for t in (host("linear.%d.weight" % (2 * i), (o, n_in)), host("linear.%d.bias" % (2 * i), (o,)))])
        h = ctypes.c_void_p()
        _lib.check(lib.npd_gru_create(int(N), int(H), *[ctypes.c_void_p(t.data_ptr()) for t in self._keep],
                                      ctypes.byref(h)))
        self.h, self.N, self.H = h, N, H
        self._keep = None
        if mlp is not None:
            _lib.check(lib.npd_gru_set_head_mlp(h, int(head_depth), int(y_hidden), ctypes.c_void_p(mlp.data_ptr())))
        _lib.check(lib.npd_gru_set_option(h, 1, int(_GRU_PRECISION == "exact")))  # NPD_GRU_OPT_RESIDUAL_STATE

    def __del__(self):
        try:
            if getattr(self, "h", None) and _lib._lib is not None:
                _lib._lib.npd_gru_destroy(self.h)
                self.h = None
        except Exception:
            pass


def gru_decode(net_or_handle, code_handle, y, forced=None, want_logits=False, genie=None, h0=None):
    """One fused launch: y [B,N] (device) -> (decoded [B,N], logits [B,N] or None).  h0 [2,B,H]: initial state."""
    B, N = y.shape
    handle = net_or_handle if isinstance(net_or_handle, GruHandle) else net_or_handle.npd_handle(N)
    decoded = torch.empty(B, N, dtype=torch.float32, device=y.device)
    logits = torch.empty(B, N, dtype=torch.float32, device=y.device) if want_logits else None
    if h0 is not None:
        h0 = h0.float().contiguous()
        assert tuple(h0.shape) == (2, B, handle.H) and h0.device == y.device, (tuple(h0.shape), h0.device)
    if B > 0:
        lib = _lib.load()
        ws_bytes = int(lib.npd_gru_workspace_bytes(handle.h, B))  # non-zero only for MLP heads
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=y.device) if ws_bytes else None
        _lib.check(lib.npd_gru_decode_h0(handle.h, code_handle.h, _lib.ptr(y), _lib.ptr(h0), _lib.ptr(forced),
                                         _lib.ptr(genie), _lib.ptr(logits), _lib.ptr(decoded), B, _lib.ptr(ws), ws_bytes,
                                         _lib.stream_ptr()))
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
        """reference rnn_all.py:408-561 for decoding_type 'y_input' (zero initial state, step input [Fy | feedback],
        Fy = y or the y-MLP net.get_Fy(y) under use_ynn, 532-547) and 'y_h0' (initial state net.get_h0(y), step input
        = feedback only, 523-531), feedback one-hot or the scalar +-1, forward or reverse order; all N autoregressive
        steps are one kernel launch.  ('y_h0_out' cannot run in the reference either: forward() concatenates
        [Fy (L*H) | out (H)] into a head built for H inputs, rnn_all.py:334-343, 393-396.)

        train=False (rnn_all.py:514-561): decisions sign(logit) on loss_inds (default: the info positions), +1
        elsewhere.  gt = genie tensor [B,N]: decoded starts as gt.clone(), so positions outside loss_inds keep and
        feed back their genie value (rnn_all.py:519-522; the `loss_inds=code.loss_inds` call of
        polar_RNN_full_test, 887).  reverse_order: step ii handles position N-1-ii (gt flipped on entry, 417-419,
        result flipped back, 558-561).
        train=True is served for EVALUATION only (no autograd graph; the reference's test_model(tf=True) calls it
        under torch.no_grad(), rnn_all.py:982-984): teacher forcing (425-461) returns the raw outputs of all N
        steps with gt fed back; student forcing (462-512) returns raw outputs on the steps `ii in info_inds`, 1
        elsewhere.  The training iteration with gradients is train.GRUTrainer.step (csrc/gru_train.cu).
        One deviation, on a measure-zero input: a scalar (non-one-hot) feedback of exactly 0 (sign of a zero logit or
        of a zero genie value) enters the reference as 0 * w and this path as -w, the one-hot convention (258-260)."""
        if self.decoding_type not in ('y_input', 'y_h0'):
            raise NotImplementedError("decoding_type %r: the accelerated path serves 'y_input' and 'y_h0'"
                                      % (self.decoding_type,))
        N = self.N
        y_h0 = self.decoding_type == 'y_h0'
        ynn = (not y_h0) and getattr(net, "y_depth", 0) != 0
        on_host = torch.is_tensor(y) and not y.is_cuda
        pipe = on_host and not (y_h0 or ynn)   # the library's chunked host pipeline; the y-MLP variants go via the device
        run = gru_decode_host if pipe else gru_decode
        yd = _lib.host_f32(y) if pipe else _lib.to_device_f32(y)
        assert yd.dim() == 2 and yd.shape[1] == N
        dev_ctx = torch.cuda.device(torch.cuda.current_device() if pipe else yd.device)
        handle = net.npd_handle(N, onehot=bool(self.onehot), y_in=not y_h0)
        flip = (lambda t: t.flip(1)) if self.reverse_order else (lambda t: t)

        def like_y(t):
            return None if t is None else (_lib.host_f32(t.cpu()) if pipe else _lib.to_device_f32(t, yd.device))

        def back(t):
            t = flip(t)
            return t.cpu() if (on_host and not pipe) else t

        kw = {}
        if gt is not None:
            gt = flip(gt)
        with dev_ctx:
            if y_h0:
                with torch.no_grad():
                    kw["h0"] = net.get_h0(yd)
            elif ynn:
                with torch.no_grad():
                    yd = net.get_Fy(yd).contiguous()
                assert yd.shape[1] == N, "use_ynn: the y-MLP must emit N values (y_output_size = N, rnn_all.py:1320)"
            if train:
                if torch.is_grad_enabled() and any(p.requires_grad for p in net.parameters()):
                    raise NotImplementedError("decode(train=True) returns no autograd graph on the B200 path: the whole training "
                                              "iteration (forward, backward, clip, AdamW) is train.GRUTrainer.step; call "
                                              "under torch.no_grad() for teacher-/student-forced evaluation")
                if random.random() < teacher_forcing_ratio:  # rnn_all.py:425
                    assert gt is not None and gt.shape[1] == N
                    _, logits = run(handle, self._loss_code(self.info_inds), yd, forced=like_y(gt), want_logits=True, **kw)
                    return back(logits)
                _, logits = run(handle, self._loss_code(self.info_inds), yd, want_logits=True, **kw)
                mask = torch.zeros(N, dtype=torch.bool, device=logits.device)
                mask[torch.as_tensor(np.asarray(self.info_inds), device=logits.device)] = True  # `ii in info_inds`
                return back(torch.where(mask.unsqueeze(0), logits, torch.ones_like(logits)))
            if loss_inds is None:
                loss_inds = self.info_inds
            steps = np.asarray(loss_inds, dtype=np.int64)
            if self.reverse_order:
                steps = N - 1 - steps  # `jj in loss_inds` with jj = N-1-ii (rnn_all.py:416, 546)
            decoded, logits = run(handle, self._loss_code(steps), yd, want_logits=return_logits, genie=like_y(gt), **kw)
        return (back(decoded), back(logits)) if return_logits else back(decoded)


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
    elif rate_profile == 'sorted':  # rnn_all.py:1089-1091: the target_K best positions in ascending index order
        first = np.sort(rs[:target_K].copy())
        rs[:target_K] = first
        code = PolarCode(n, K, args, rs=rs)
    elif rate_profile == 'sorted_last':  # rnn_all.py:1108-1110: ... in descending index order
        first = np.sort(rs[:target_K].copy())
        rs[:target_K] = first[::-1]
        code = PolarCode(n, K, args, rs=rs)
    elif rate_profile == 'random':  # rnn_all.py:1179-1180
        first = rs[:target_K].copy()
        rs[:target_K] = np.random.RandomState(seed=getattr(args, "random_seed", 42)).permutation(first)
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
    loss_only = getattr(args, "loss_only", None)
    if loss_only is not None:  # rnn_all.py:1189-1192: score (and train on) the first `loss_only` positions of rs only
        code.loss_inds = np.sort(rs[:loss_only].copy())
        code.msg_indices = np.where(np.isin(code.info_inds, code.loss_inds))[0]
    else:
        code.loss_inds = None
        code.msg_indices = np.arange(K)
    return code


if __name__ == '__main__':
    # `python -m neural_polar_decoder_b200.rnn_all <run_crisp.sh flags> --test` (reference rnn_all.py:1198-1949,
    # evaluation half; see cli.py)
    import sys
    from .cli import main
    sys.exit(main())
