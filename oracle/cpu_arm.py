"""CPU arm of bench.py (TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package, and it never
imports the product package: the reference arm must not map libnpd.so).

Times the reference's own CPU implementation of the hot path on the host cores:

  kind "reference"  the UNMODIFIED reference classes imported from /root/reference through oracle/ref_shim.py
                    (rnn_all.RNN_decoder.decode + rnn_all.RNN_Model, polar.PolarCode.sc_decode_new / scl_decode,
                    pac_code.PAC.pac_sc_decode, models.convNet.decode) -- available in the build container only: a
                    Python reference cannot travel to the GPU box;
  kind "port"       where /root/reference is absent (the GPU box): the restatement of the same call structure --
                    torch eager nn.GRU(seq_len 1) + nn.Linear stepping exactly as rnn_all.py:532-547, the torch eager
                    Conv1d / GELU / Linear / LayerNorm stack of models.py:742-767, and the C oracle (npd_oracle.c) for
                    the SC family.  NOTE the C oracle is ~10^3 x faster than the reference's torch-op SC loop
                    (1.8 cw/s at N = 1024), so the "port" SC figure flatters the CPU.

Every function returns (codewords per second, codewords decoded, seconds, kind, how)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def have_reference():
    return ref_shim.available()


# ---- code constructions (restated here so that this arm never imports the product package) ----------------------
def reference_rs(N):
    """The reference's N <= 256 reliability table (rnn_all.py:1046), read from the package's data file."""
    with open(os.path.join(ROOT, "neural_polar_decoder_b200", "data", "polar_rs256.json")) as f:
        rs = np.asarray(json.load(f)["rs"], dtype=np.int64)
    return rs[rs < N]


def pw_info_set(N, K):
    """Polarization-weight construction (beta = 2^(1/4)): the K largest sum_j b_j(i) beta^j carry information;
    ties cannot occur.  Must equal neural_polar_decoder_b200.construct (tests/test_cabi_and_host.py checks)."""
    n = int(np.log2(N))
    beta = 2.0 ** 0.25
    idx = np.arange(N)
    w = np.zeros(N)
    for j in range(n):
        w += ((idx >> j) & 1) * beta ** j
    order = np.lexsort((-idx, -w))
    return np.sort(order[:K]).astype(np.int32)


def info_set(N, K, pac_g=None):
    if pac_g:
        wt = np.array([bin(i).count("1") for i in range(N)])
        return np.sort(np.argsort(wt)[-K:]).astype(np.int32)  # pac_code.py:132-135 (RM profile)
    if N <= 256:
        return np.sort(reference_rs(N)[:K]).astype(np.int32)
    return pw_info_set(N, K)


def trained_gru(N, K, H=512):
    """(state_dict, meta) of the reference-trained checkpoint fixture for this shape, or (None, None)."""
    name = "crisp_gru_N%d_K%d_H%d" % (N, K, H)
    pt, js = os.path.join(GOLD, name + ".pt"), os.path.join(GOLD, name + ".json")
    if not (os.path.exists(pt) and os.path.exists(js)):
        return None, None
    import torch
    ckpt = torch.load(pt, map_location="cpu", weights_only=False)
    with open(js) as f:
        return ckpt["net"], json.load(f)


def _frames(rs, B, N, K, info, snr, n, pac_g=None):
    import oracle
    msg = (1.0 - 2.0 * rs.randint(0, 2, size=(B, K))).astype(np.float32)
    x = oracle.pac_encode(msg, n, info, pac_g) if pac_g else oracle.polar_encode(msg, n, info)
    return msg, (x + 10 ** (-snr / 20) * rs.randn(B, N)).astype(np.float32)


# ---- CRISP GRU ---------------------------------------------------------------------------------------------------
class _PortGRU:
    """nn.GRU + nn.Linear with the reference's state_dict keys (rnn_all.py:307, 333-334), stepped as 532-547."""

    def __init__(self, N, H, sd, seed):
        import torch
        from torch import nn
        torch.manual_seed(seed)

        class M(nn.Module):
            def __init__(s):
                super().__init__()
                s.rnn = nn.GRU(N + 2, H, 2, batch_first=True)
                s.linear = nn.Linear(H, 1)

            def forward(s, x, h):
                out, h = s.rnn(x, h)
                return s.linear(out).squeeze(-1), h

        self.net = M().eval()
        if sd is not None:
            self.net.load_state_dict({k: v for k, v in sd.items() if k.startswith(("rnn.", "linear."))})
        self.N, self.H = N, H

    def decode(self, y, info):
        import torch
        B, N = y.shape
        info = set(int(i) for i in info)
        eye = torch.eye(2)
        with torch.no_grad():
            decoded = torch.ones(B, N)
            hidden = torch.zeros(2, B, self.H)
            for ii in range(N):
                prev = torch.ones(B) if ii == 0 else decoded[:, ii - 1].sign()
                onehot = eye[(0.5 + 0.5 * prev).long()]
                out, hidden = self.net(torch.cat([y.unsqueeze(1), onehot.view(B, 1, 2)], 2), hidden)
                if ii in info:
                    decoded[:, ii] = out.squeeze().sign()
        return decoded


def make_gru_arm(N, K, H=512, seed=0):
    """-> (decode_fn(y_numpy[B,N]) -> decisions, info, kind, how)."""
    import torch
    sd, _ = trained_gru(N, K, H)
    info = info_set(N, K)
    if have_reference():
        ra = ref_shim.load("rnn_all")
        ra.args = ref_shim.make_args(N=N, K=K)
        torch.manual_seed(seed)
        net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0)
        if sd is not None:
            net.load_state_dict(sd)
        dec = ra.RNN_decoder("y_input", N, info, onehot=True)
        return (lambda y: dec.decode(net, False, torch.from_numpy(y))), info, "reference", \
            "unmodified rnn_all.RNN_decoder.decode(net, False, y) on rnn_all.RNN_Model (torch CPU eager)"
    port = _PortGRU(N, H, sd, seed)
    return (lambda y: port.decode(torch.from_numpy(y), info)), info, "port", \
        "torch eager nn.GRU(seq_len 1) + nn.Linear stepping as rnn_all.py:532-547"


def gru_rate(N, K, snr, B, threads, seed=0, H=512, arm=None):
    import torch
    torch.set_num_threads(threads)
    fn, info, kind, how = arm if arm is not None else make_gru_arm(N, K, H, seed)
    rs = np.random.RandomState(seed)
    _, y = _frames(rs, B, N, K, info, snr, int(np.log2(N)))
    t0 = time.perf_counter()
    fn(y)
    dt = time.perf_counter() - t0
    return B / dt, B, dt, kind, how


# ---- CRISP GRU training iteration (rnn_all.py:1399-1437) -------------------------------------------------------------
def make_train_arm(N, K, H=512, seed=0):
    """-> (step_fn(y, gt) running one forward / backward / clip / AdamW iteration, info, kind, how)."""
    import torch
    info = info_set(N, K)
    idx = torch.as_tensor(info.astype(np.int64))
    if have_reference():
        ra = ref_shim.load("rnn_all")
        ra.args = ref_shim.make_args(N=N, K=K)
        torch.manual_seed(seed)
        net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0)
        dec = ra.RNN_decoder("y_input", N, info, onehot=True)
        fwd = lambda y, gt: dec.decode(net, True, y, gt, 1.0)  # noqa: E731
        kind, how = "reference", "unmodified rnn_all.RNN_decoder.decode(net, True, y, gt, 1.0) + MSELoss + backward + clip + AdamW"
    else:
        port = _PortGRU(N, H, None, seed)
        net = port.net.train()
        eye = torch.eye(2)

        def fwd(y, gt):
            B = y.shape[0]
            hidden = torch.zeros(2, B, H)
            outs = []
            for ii in range(N):
                prev = torch.ones(B) if ii == 0 else gt[:, ii - 1]
                out, hidden = net(torch.cat([y.unsqueeze(1), eye[(0.5 + 0.5 * prev).long()].view(B, 1, 2)], 2), hidden)
                outs.append(out.view(-1))
            return torch.stack(outs, 1)
        kind, how = "port", "torch autograd over nn.GRU(seq_len 1) steps as rnn_all.py:425-449 + MSELoss + backward + clip + AdamW"
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    loss_fn = torch.nn.MSELoss()

    def step(y, gt):
        y, gt = torch.from_numpy(y), torch.from_numpy(gt)
        loss = loss_fn(fwd(y, gt)[:, idx], gt[:, idx])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 0.25)
        opt.step()
        opt.zero_grad()
        return float(loss.item())
    return step, info, kind, how


def train_rate(N, K, snr, B, threads, seed=0, H=512, arm=None):
    import torch
    torch.set_num_threads(threads)
    fn, info, kind, how = arm if arm is not None else make_train_arm(N, K, H, seed)
    rs = np.random.RandomState(seed)
    msg, y = _frames(rs, B, N, K, info, snr, int(np.log2(N)))
    gt = np.ones((B, N), np.float32)
    gt[:, info] = msg
    t0 = time.perf_counter()
    fn(y, gt)
    dt = time.perf_counter() - t0
    return B / dt, B, dt, kind, how


# ---- SC family ---------------------------------------------------------------------------------------------------
def make_sc_arm(N, K, snr, threads, pac_g=None, L=0):
    """-> (decode_fn(y), info, kind, how)."""
    import oracle
    import torch
    n = int(np.log2(N))
    info = info_set(N, K, pac_g)
    if have_reference():
        torch.set_num_threads(threads)
        if pac_g:
            code = ref_shim.get_code("PAC", "RM", N, K, g=pac_g)
            assert np.array_equal(np.sort(np.asarray(code.B)), info)
            return (lambda y: code.pac_sc_decode(torch.from_numpy(y), snr)), info, "reference", \
                "unmodified pac_code.PAC.pac_sc_decode (torch CPU eager)"
        polar = ref_shim.load("polar")
        F = np.array(sorted(set(range(N)) - set(info.tolist())))
        code = polar.PolarCode(n, K, ref_shim.make_args(N=N, K=K), F=F, use_cuda=False)
        if L:
            return (lambda y: code.scl_decode(torch.from_numpy(y), snr, L, False)), info, "reference", \
                "unmodified polar.PolarCode.scl_decode (torch CPU eager)"
        return (lambda y: code.sc_decode_new(torch.from_numpy(y), snr)), info, "reference", \
            "unmodified polar.PolarCode.sc_decode_new (torch CPU eager)"
    if pac_g:
        f = lambda lo, hi, y: oracle.pac_sc_decode(y[lo:hi], snr, n, info, pac_g)  # noqa: E731
    elif L:
        f = lambda lo, hi, y: oracle.scl_decode(y[lo:hi], snr, n, info, L)  # noqa: E731
    else:
        f = lambda lo, hi, y: oracle.sc_decode(y[lo:hi], snr, n, info)  # noqa: E731
    return (lambda y: oracle.run_threaded(lambda lo, hi: f(lo, hi, y), y.shape[0], threads)), info, "port", \
        "C oracle (npd_oracle.c, the restatement of polar.py:465-484) on %d threads" % threads


def sc_rate(N, K, snr, B, threads, seed=0, pac_g=None, L=0, arm=None):
    fn, info, kind, how = arm if arm is not None else make_sc_arm(N, K, snr, threads, pac_g, L)
    rs = np.random.RandomState(seed)
    _, y = _frames(rs, B, N, K, info, snr, int(np.log2(N)), pac_g)
    t0 = time.perf_counter()
    fn(y)
    dt = time.perf_counter() - t0
    return B / dt, B, dt, kind, how


# ---- convNet -----------------------------------------------------------------------------------------------------
def make_conv_arm(N=64, E=128, seed=0):
    import argparse
    import torch
    if have_reference():
        rm = ref_shim.load("models")
        torch.manual_seed(seed)
        net = rm.convNet(argparse.Namespace(embed_dim=E, max_len=N, N=N, dont_use_bias=False, dropout=0.1))
        ck = os.path.join(GOLD, "conv_N%d_K22_E%d.pt" % (N, E))
        if os.path.exists(ck):  # the reference-trained checkpoint, like the product arm
            net.load_state_dict(torch.load(ck, map_location="cpu", weights_only=False)["xformer"])
        net.eval()
        info = info_set(N, 22)

        def fn(y):
            with torch.no_grad():
                return net.decode(torch.from_numpy(y), info, None, "cpu")
        return fn, "reference", "unmodified models.convNet.decode (torch CPU eager)"
    from torch import nn
    torch.manual_seed(seed)

    def conv(ci, co, d):
        return nn.Conv1d(ci, co, 7, padding=3 * d, dilation=d)
    C = 64
    l1 = nn.Sequential(conv(1, C, 1), nn.GELU(), conv(C, C, 2), nn.GELU())
    l2 = nn.Sequential(conv(C, C, 4), nn.GELU(), conv(C, C, 1), nn.GELU())
    l3 = nn.Sequential(conv(C, C, 2), nn.GELU(), conv(C, C, 4), nn.GELU())
    l4 = nn.Sequential(conv(C, C, 1), nn.GELU(), conv(C, C, 2), nn.GELU())
    l5 = nn.Sequential(conv(C, E, 4), nn.GELU(), conv(E, E, 1), nn.GELU())
    fin = nn.Sequential(nn.Linear(E * N, 256), nn.GELU(), nn.Linear(256, 64), nn.GELU(), nn.Linear(64, N))
    ln = nn.LayerNorm(N, eps=1e-6)
    for m in (l1, l2, l3, l4, l5, fin, ln):
        m.eval()

    def fn(y):
        with torch.no_grad():
            x2 = l1(torch.from_numpy(y).unsqueeze(1))
            x3 = l2(x2) + x2
            x4 = l3(x3) + x3
            x5 = l4(x4) + x4
            return ln(fin(torch.flatten(l5(x5), start_dim=1))).sign()
    return fn, "port", "torch eager Conv1d/GELU/Linear/LayerNorm stack as in models.py:742-767"


def conv_rate(N, snr, B, threads, seed=0, arm=None):
    import torch
    torch.set_num_threads(threads)
    fn, kind, how = arm if arm is not None else make_conv_arm(N, 128, seed)
    rs = np.random.RandomState(seed)
    _, y = _frames(rs, B, N, 22, info_set(N, 22), snr, int(np.log2(N)))
    t0 = time.perf_counter()
    fn(y)
    dt = time.perf_counter() - t0
    return B / dt, B, dt, kind, how
