"""ctypes front-end of the CPU oracle (oracle/npd_oracle.c) plus the fp32 torch oracles for the two
floating-point decoders (GRU, CNN).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Parity status: pinned against the live reference through tests/golden/ (oracle/gen_golden.py) and,
where /root/reference is present, tests/test_oracle_vs_reference.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libnpd_oracle.so")
    src = os.path.join(_HERE, "npd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libnpd_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.npdo_sc_decode.restype = ctypes.c_int
        _LIB.npdo_pac_sc_decode.restype = ctypes.c_int
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a, t=_f32p):
    return None if a is None else a.ctypes.data_as(t)


def frozen_mask(N, info):
    m = np.ones(N, dtype=np.uint8)
    m[np.asarray(info, dtype=np.int64)] = 0
    return m


def llr_scale(snr_db: float) -> np.float32:
    """fp32(2/sigma^2) exactly as torch computes `(2/sigma**2) * tensor` (polar.py:467-469,
    utils.py:5-6): double arithmetic, then the Python scalar is rounded to fp32."""
    sigma = 10 ** (-snr_db * 1.0 / 20)
    return np.float32(2 / sigma ** 2)


def polar_encode(msg, n, info):
    msg = _f32(msg)
    B, K = msg.shape
    info = np.ascontiguousarray(info, dtype=np.int32)
    x = np.empty((B, 1 << n), dtype=np.float32)
    lib().npdo_polar_encode(_ptr(msg), ctypes.c_int64(B), n, K, _ptr(info, _i32p), _ptr(x))
    return x


def sc_decode(y, snr_db, n, info, infty=1000.0, use_gt=None, scale=None):
    """-> (leaf_llr[B,N], u_hat[B,N], decoded[B,K]) following polar.py:465-484."""
    y = _f32(y)
    B, N = y.shape
    assert N == 1 << n
    info = np.ascontiguousarray(info, dtype=np.int32)
    K = info.shape[0]
    fr = frozen_mask(N, info)
    gt = None if use_gt is None else _f32(use_gt)
    leaf = np.empty((B, N), dtype=np.float32)
    uh = np.empty((B, N), dtype=np.float32)
    dec = np.empty((B, K), dtype=np.float32)
    s = llr_scale(snr_db) if scale is None else np.float32(scale)
    rc = lib().npdo_sc_decode(_ptr(y), ctypes.c_int64(B), n, K, _ptr(info, _i32p), _ptr(fr, _u8p),
                              ctypes.c_float(s), ctypes.c_float(infty), _ptr(gt), _ptr(leaf),
                              _ptr(uh), _ptr(dec))
    assert rc == 0
    return leaf, uh, dec


def scl_decode(y, snr_db, n, info, L, infty=1000.0, scale=None):
    """-> (leaf_llr[B,N] of the chosen path, decoded[B,K]) following polar.py:793-876 (use_CRC=False)."""
    y = _f32(y)
    B, N = y.shape
    assert N == 1 << n
    info = np.ascontiguousarray(info, dtype=np.int32)
    K = info.shape[0]
    fr = frozen_mask(N, info)
    leaf = np.empty((B, N), dtype=np.float32)
    dec = np.empty((B, K), dtype=np.float32)
    s = llr_scale(snr_db) if scale is None else np.float32(scale)
    rc = lib().npdo_scl_decode(_ptr(y), ctypes.c_int64(B), n, K, _ptr(info, _i32p), _ptr(fr, _u8p),
                               ctypes.c_float(s), ctypes.c_float(infty), int(L), _ptr(leaf), _ptr(dec))
    assert rc == 0
    return leaf, dec


def pac_g_array(g: int):
    """pac_code.py:101-103: M = floor(log2 g)+1 bits, MSB first, mapped to 1-2*bit."""
    M = int(np.floor(np.log2(g))) + 1
    bits = np.array([(g >> (M - 1 - i)) & 1 for i in range(M)], dtype=np.float32)
    return (1.0 - 2.0 * bits).astype(np.float32)


def pac_encode(msg, n, info, g):
    msg = _f32(msg)
    B, K = msg.shape
    info = np.ascontiguousarray(info, dtype=np.int32)
    ga = pac_g_array(g)
    x = np.empty((B, 1 << n), dtype=np.float32)
    lib().npdo_pac_encode(_ptr(msg), ctypes.c_int64(B), n, K, _ptr(info, _i32p), _ptr(ga),
                          len(ga), _ptr(x))
    return x


def pac_sc_decode(y, snr_db, n, info, g, use_gt_codeword=None, scale=None):
    """-> (leaf_llr[B,N], v_hat[B,K], u_hat[B,N]) following pac_code.py:534-573."""
    y = _f32(y)
    B, N = y.shape
    info = np.ascontiguousarray(info, dtype=np.int32)
    K = info.shape[0]
    fr = frozen_mask(N, info)
    ga = pac_g_array(g)
    gt = None if use_gt_codeword is None else _f32(use_gt_codeword)
    leaf = np.empty((B, N), dtype=np.float32)
    vh = np.empty((B, K), dtype=np.float32)
    uh = np.empty((B, N), dtype=np.float32)
    s = llr_scale(snr_db) if scale is None else np.float32(scale)
    rc = lib().npdo_pac_sc_decode(_ptr(y), ctypes.c_int64(B), n, K, _ptr(info, _i32p),
                                  _ptr(fr, _u8p), _ptr(ga), len(ga), ctypes.c_float(s), _ptr(gt),
                                  _ptr(leaf), _ptr(vh), _ptr(uh))
    assert rc == 0
    return leaf, vh, uh


def count_errors(a, b):
    """-> (bit_errors, block_errors) = numerators of utils.py:17-25 / 37-51."""
    a = _f32(a)
    b = _f32(b)
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    out = np.zeros(2, dtype=np.uint64)
    lib().npdo_count_errors(_ptr(a), _ptr(b), ctypes.c_int64(a.shape[0]), a.shape[1],
                            _ptr(out, _u64p))
    return int(out[0]), int(out[1])


def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().npdo_philox4x32_10(_ptr(c, _u32p), _ptr(k, _u32p), _ptr(o, _u32p))
    return o


def gen_msg(seed, cw0, B, K):
    m = np.empty((B, K), dtype=np.float32)
    lib().npdo_gen_msg(ctypes.c_uint64(seed), ctypes.c_uint64(cw0), ctypes.c_int64(B), K, _ptr(m))
    return m


def gen_noise(seed, cw0, pt, B, N):
    z = np.empty((B, N), dtype=np.float32)
    lib().npdo_gen_noise(ctypes.c_uint64(seed), ctypes.c_uint64(cw0), ctypes.c_uint32(pt),
                         ctypes.c_int64(B), N, _ptr(z))
    return z


# --------------------------------------------------------------------------------------------------
# fp32 torch oracles for the floating-point decoders.
# --------------------------------------------------------------------------------------------------

def y_mlp(sd, y, y_depth, activation="relu"):
    """RNN_Model.get_Fy (rnn_all.py:377-384): Linear layers y_linears.{i}; the activation (act, 346-360) follows every
    layer whose index differs from y_depth -- i.e. all of them when there are y_depth layers (y_depth >= 2), and only
    the first of the two layers that y_depth == 1 builds (323-329)."""
    import torch
    import torch.nn.functional as F
    act = {"tanh": torch.tanh, "elu": F.elu, "relu": F.relu, "selu": F.selu, "sigmoid": torch.sigmoid}.get(
        activation, lambda t: t)
    x = torch.as_tensor(np.asarray(y)).float()
    ii = 0
    while ("y_linears.%d.weight" % ii) in sd:
        W = torch.as_tensor(np.asarray(sd["y_linears.%d.weight" % ii])).float()
        b = torch.as_tensor(np.asarray(sd["y_linears.%d.bias" % ii])).float()
        x = x @ W.t() + b
        if ii != y_depth:
            x = act(x)
        ii += 1
    return x


def gru_h0(sd, y, y_depth, H, L=2, activation="relu"):
    """RNN_Model.get_h0 (rnn_all.py:362-375, skip=False): the y-MLP output [B, L*H] is READ AS [B, H, L] and permuted
    to [L, B, H] -- layer l's state is the stride-L slice x[:, l::L]."""
    x = y_mlp(sd, y, y_depth, activation)
    return x.reshape(-1, H, L).permute(2, 0, 1).contiguous().numpy()


def gru_decode(sd, y, N, info, forced=None, H=None, round_bf16=False, genie=None, h0=None, onehot=True, y_in=True):
    """fp32 restatement of RNN_decoder.decode(net, False, y) for decoding_type 'y_input', onehot,
    GRU (rnn_all.py:514-521, 532-547) with RNN_Model.forward (387-398) written out gate by gate
    (PyTorch nn.GRU gate order r,z,n).  The y-part of the layer-0 input projection is
    hoisted out of the step loop -- bit-for-bit equal to the literal form was
    checked against the live reference in tests/test_oracle_vs_reference.py within fp32 round-off.

    sd: state_dict-like mapping with the a7 key names (torch tensors or numpy arrays).
    forced: optional [B,N] of +-1: when given, step ii feeds back forced[:, ii-1] instead of the
            decoder's own decision (used for logit parity under identical feedback).
    genie: optional [B,N]: decoded starts as this tensor instead of ones (gt.clone(), rnn_all.py:519-522), so
            positions outside `info` (the loss positions) keep and feed back their genie value.
    h0: optional [L,B,H] initial state (decoding_type 'y_h0': net.get_h0(y), rnn_all.py:523-524; default zeros, 538).
    y_in: False when the step input carries no y part ('y_h0', 526-528: input = feedback only).
    onehot: False = the feedback enters as the scalar prev (RNN_decoder onehot=False, onehot_fn = identity, 410-411).
    round_bf16: emulate the kernel's operand rounding (weights, h and y operands to bf16, fp32
            accumulate) to size tolerances; not a parity target.
    -> (decoded[B,N] in {-1,0,+1} with +1 on non-info positions, logits[B,N])
    """
    import torch

    def T(v):
        return torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).float().cpu()

    def rb(t):
        return t.to(torch.bfloat16).float() if round_bf16 else t

    y = T(y)
    B = y.shape[0]
    L = 0
    while ("rnn.weight_ih_l%d" % L) in sd:
        L += 1
    Wih = [T(sd["rnn.weight_ih_l%d" % l]) for l in range(L)]
    Whh = [T(sd["rnn.weight_hh_l%d" % l]) for l in range(L)]
    bih = [T(sd["rnn.bias_ih_l%d" % l]) for l in range(L)]
    bhh = [T(sd["rnn.bias_hh_l%d" % l]) for l in range(L)]
    if "linear.weight" in sd:
        Wo, bo, head = T(sd["linear.weight"]), T(sd["linear.bias"]), None
    else:  # out_linear_depth > 1 (rnn_all.py:335-343): nn.Sequential of Linear (even indices) with SELU in between
        head, i = [], 0
        while ("linear.%d.weight" % i) in sd:
            head.append((T(sd["linear.%d.weight" % i]), T(sd["linear.%d.bias" % i])))
            i += 2
    Hs = Whh[0].shape[1]
    info_set = set(int(i) for i in info)
    h = [torch.zeros(B, Hs) for _ in range(L)] if h0 is None else [T(h0)[l].clone() for l in range(L)]
    decoded = torch.ones(B, N) if genie is None else T(genie).clone()
    logits = torch.zeros(B, N)
    ny = N if y_in else 0
    Gy = rb(y) @ rb(Wih[0][:, :N]).t() if y_in else torch.zeros(B, 3 * Hs)  # hoisted y projection
    if onehot:
        col_m1 = Wih[0][:, ny]      # onehot(-1) = [1,0]  (rnn_all.py:258-260)
        col_p1 = Wih[0][:, ny + 1]  # onehot(+1) = [0,1]
    else:
        col_fb = Wih[0][:, ny]      # scalar feedback column
    for ii in range(N):
        if ii == 0:
            prev = torch.ones(B)
        elif forced is not None:
            prev = T(forced)[:, ii - 1]
        else:
            prev = decoded[:, ii - 1].sign()
        if onehot:
            sel = (0.5 + 0.5 * prev).long()  # 0 -> column N, 1 -> column N+1 (sign 0 maps to 0)
            gi = Gy + torch.where(sel[:, None] == 1, col_p1[None, :], col_m1[None, :]) + bih[0]
        else:
            gi = Gy + prev[:, None] * col_fb[None, :] + bih[0]
        x = None
        for l in range(L):
            if l > 0:
                gi = rb(x) @ rb(Wih[l]).t() + bih[l]
            gh = rb(h[l]) @ rb(Whh[l]).t() + bhh[l]
            r = torch.sigmoid(gi[:, :Hs] + gh[:, :Hs])
            z = torch.sigmoid(gi[:, Hs:2 * Hs] + gh[:, Hs:2 * Hs])
            n = torch.tanh(gi[:, 2 * Hs:] + r * gh[:, 2 * Hs:])
            h[l] = (1 - z) * n + z * h[l]
            x = h[l]
        if head is not None:
            out = x
            for li, (Wl, bl) in enumerate(head):
                out = out @ Wl.t() + bl
                if li + 1 < len(head):
                    out = torch.nn.functional.selu(out)
            out = out.view(-1)
        else:
            out = (rb(x) @ rb(Wo).t() + bo).view(-1) if round_bf16 else (x @ Wo.t() + bo).view(-1)
        logits[:, ii] = out
        if ii in info_set:
            decoded[:, ii] = out.sign()
    return decoded.numpy(), logits.numpy()


TRAIN_KEYS = ("rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0", "rnn.weight_ih_l1",
              "rnn.weight_hh_l1", "rnn.bias_ih_l1", "rnn.bias_hh_l1", "linear.weight", "linear.bias")


def train_shapes(N, H):
    I = N + 2
    return [(3 * H, I), (3 * H, H), (3 * H,), (3 * H,), (3 * H, H), (3 * H, H), (3 * H,), (3 * H,), (1, H), (1,)]


def gru_train_step(blob, y, gt, N, H, info, teacher, lr=1e-3, clip=0.25, state=None):
    """fp32 torch-autograd restatement of one iteration of the reference's training loop (rnn_all.py:1399-1437) for
    GRU / 'y_input' / onehot / 2 layers / Linear(H,1): teacher-forced (425-449) or student-forced (463-489) pass,
    MSELoss on the info positions (1413), backward, clip_grad_norm_ (1432), torch.optim.AdamW step (1346, 1435).
    The GRU cell is written out gate by gate (no nn.GRU) so that the restatement is independent of torch's fused RNN.

    blob : parameters as one fp32 vector in TRAIN_KEYS order;  state: None or (exp_avg, exp_avg_sq, step)
    -> (new blob, clipped grad blob, loss, total grad norm, logits [B,N], new state)"""
    import torch
    y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
    gt = torch.from_numpy(np.ascontiguousarray(gt, dtype=np.float32))
    B = y.shape[0]
    params, o = [], 0
    for shp in train_shapes(N, H):
        n = int(np.prod(shp))
        params.append(torch.from_numpy(np.array(blob[o:o + n], dtype=np.float32).reshape(shp)).requires_grad_(True))
        o += n
    Wih0, Whh0, bih0, bhh0, Wih1, Whh1, bih1, bhh1, Wo, bo = params
    info_set = set(int(i) for i in info)
    eye = torch.eye(2)

    def cell(x, h, Wi, Wh, bi, bh):
        gi, gh = x @ Wi.t() + bi, h @ Wh.t() + bh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        return (1 - z) * n + z * h

    h0 = torch.zeros(B, H)
    h1 = torch.zeros(B, H)
    decoded = torch.ones(B, N)
    outs = []
    for ii in range(N):
        if ii == 0:
            prev = torch.ones(B)
        elif teacher:
            prev = gt[:, ii - 1]
        else:
            prev = decoded[:, ii - 1].sign().detach()
        x = torch.cat([y, eye[(0.5 + 0.5 * prev).long()]], 1)
        h0 = cell(x, h0, Wih0, Whh0, bih0, bhh0)
        h1 = cell(h0, h1, Wih1, Whh1, bih1, bhh1)
        out = (h1 @ Wo.t() + bo).view(-1)
        outs.append(out)
        if teacher or ii in info_set:
            decoded = torch.cat([decoded[:, :ii], out.view(-1, 1), decoded[:, ii + 1:]], 1)
    logits = torch.stack(outs, 1)
    idx = torch.as_tensor(sorted(info_set))
    loss = torch.nn.functional.mse_loss(logits[:, idx], gt[:, idx])
    loss.backward()
    norm = torch.nn.utils.clip_grad_norm_(params, clip) if clip and clip > 0 else torch.zeros(())
    opt = torch.optim.AdamW(params, lr=lr)
    if state is not None:
        for p_, m_, v_ in zip(params, _split(state[0], N, H), _split(state[1], N, H)):
            opt.state[p_] = {"step": torch.tensor(float(state[2])), "exp_avg": torch.from_numpy(m_.copy()),
                             "exp_avg_sq": torch.from_numpy(v_.copy())}
    grad = np.concatenate([p_.grad.numpy().reshape(-1) for p_ in params]).astype(np.float32)
    opt.step()
    new = np.concatenate([p_.detach().numpy().reshape(-1) for p_ in params]).astype(np.float32)
    st = (np.concatenate([opt.state[p_]["exp_avg"].numpy().reshape(-1) for p_ in params]),
          np.concatenate([opt.state[p_]["exp_avg_sq"].numpy().reshape(-1) for p_ in params]),
          int(opt.state[params[0]]["step"]))
    return new, grad, float(loss.item()), float(norm), logits.detach().numpy(), st


def _split(blob, N, H):
    out, o = [], 0
    for shp in train_shapes(N, H):
        n = int(np.prod(shp))
        out.append(np.asarray(blob[o:o + n], dtype=np.float32).reshape(shp))
        o += n
    return out


_CONV_SPEC = [  # (sequential name, index, dilation)   models.py:701-730, padding = 3*dilation
    ("layers1", 0, 1), ("layers1", 2, 2),
    ("layers2", 0, 4), ("layers2", 2, 1),
    ("layers3", 0, 2), ("layers3", 2, 4),
    ("layers4", 0, 1), ("layers4", 2, 2),
    ("layers5", 0, 4), ("layers5", 2, 1),
]


def conv_forward(sd, y, round_bf16=False, return_in4=False):
    """fp32 restatement of convNet.forward (models.py:742-767): 10 dilated k=7 Conv1d + exact GELU with
    three residual adds, flatten, 3-layer MLP, LayerNorm(N, eps=1e-6).  -> logits[B,N]
    (+ input4 [B,C,N], forward()'s 5th return value, when return_in4)."""
    import torch
    import torch.nn.functional as F

    def T(v):
        return torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).float().cpu()

    def rb(t):
        return t.to(torch.bfloat16).float() if round_bf16 else t

    def conv(x, name, idx, dil):
        w = T(sd["%s.%d.weight" % (name, idx)])
        b = sd.get("%s.%d.bias" % (name, idx))
        b = None if b is None else T(b)
        return F.gelu(F.conv1d(rb(x), rb(w), b, padding=3 * dil, dilation=dil))

    x = T(y).unsqueeze(1)
    acts = []
    cur = x
    for gi in range(5):
        a = conv(cur, *_CONV_SPEC[2 * gi])
        a = conv(a, *_CONV_SPEC[2 * gi + 1])
        if gi in (1, 2, 3):
            a = a + cur
        cur = a
        acts.append(cur)
    v = torch.flatten(cur, start_dim=1)
    for li, idx in enumerate((0, 2, 4)):
        w = T(sd["layersFin.%d.weight" % idx])
        b = sd.get("layersFin.%d.bias" % idx)
        v = rb(v) @ rb(w).t()
        if b is not None:
            v = v + T(b)
        if li < 2:
            v = F.gelu(v)
    Nn = v.shape[1]
    v = F.layer_norm(v, (Nn,), T(sd["layer_norm.weight"]), T(sd["layer_norm.bias"]), 1e-6)
    if return_in4:
        return v.numpy(), acts[2].numpy()
    return v.numpy()


def run_threaded(fn, B, threads):
    """Split rows [0,B) over `threads` Python threads (ctypes releases the GIL inside the C oracle).
    fn(lo, hi) is called once per slice; returns the list of results in slice order."""
    from concurrent.futures import ThreadPoolExecutor
    threads = max(1, min(int(threads), B))
    bounds = [(B * t) // threads for t in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        return list(ex.map(lambda t: fn(bounds[t], bounds[t + 1]), range(threads)))
