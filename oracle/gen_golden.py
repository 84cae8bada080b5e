"""Mint tests/golden/*.npz from the LIVE reference (/root/reference through oracle/ref_shim.py).

Run in the build container only:  python oracle/gen_golden.py [--big]
The fixtures pin the oracle (tests/test_oracle_golden.py, CPU) and are the committed parity targets
of the GPU tests (tests/test_gpu_parity.py).  Inputs are generated with numpy RandomState so that the
script is reproducible; outputs are whatever the reference returns on this torch build.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
from neural_polar_decoder_b200 import construct, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def bpsk_msgs(rs, B, K):
    return (1.0 - 2.0 * rs.randint(0, 2, size=(B, K))).astype(np.float32)


def noisy(rs, x, snr, ties=False):
    sigma = 10 ** (-snr / 20)
    y = (x + sigma * rs.randn(*x.shape)).astype(np.float32)
    if ties:
        y = (np.round(y * 2) / 2).astype(np.float32)  # half-integers: exact cancellations -> sign(0)
    return y


def polar_cases(big):
    polar = ref_shim.load("polar")
    rs = np.random.RandomState(1234)
    out = {}
    meta = []

    def add(name, code, y, snr, use_gt=None, msg=None, x=None):
        llr, dec = code.sc_decode_new(torch.from_numpy(y), snr,
                                      use_gt=None if use_gt is None else torch.from_numpy(use_gt))
        out[name + "_y"] = y
        out[name + "_snr"] = np.float64(snr)
        out[name + "_info"] = np.asarray(code.info_positions, dtype=np.int32)
        out[name + "_llr"] = llr.numpy()
        out[name + "_dec"] = dec.numpy()
        if use_gt is not None:
            out[name + "_gt"] = use_gt
        if msg is not None:
            out[name + "_msg"] = msg
            out[name + "_x"] = x
        meta.append(name)

    # SURVEY.md App. B KAT1-4 (inputs written out; outputs recomputed from the live reference)
    c84 = polar.PolarCode(3, 4, ref_shim.make_args(8, 4), use_cuda=False)
    msg = np.array([[-1, 1, 1, -1], [1, -1, -1, -1]], dtype=np.float32)
    x = c84.encode_plotkin(torch.from_numpy(msg)).numpy()
    y = np.array([[0.9, -1.2, 0.3, -0.4, -1.1, 0.8, -0.2, 1.5],
                  [-0.5, 0.5, 1, -1, 0.25, -0.25, 2, -2]], dtype=np.float32)
    add("kat1", c84, y, 1.0, msg=msg, x=x)
    add("kat2", c84, np.array([[1, -1, 1, 1, 1, 1, 1, 1]], dtype=np.float32), 0.0)
    c22 = polar.PolarCode(1, 2, ref_shim.make_args(2, 2), use_cuda=False)
    add("kat3", c22, np.array([[1, -1]], dtype=np.float32), 0.0)
    add("kat4", c22, np.array([[0, -1]], dtype=np.float32), 0.0)

    configs = [(2, 1, None), (4, 2, None), (8, 4, None), (16, 8, None), (16, 16, None), (16, 1, None),
               (32, 16, "polar"), (64, 22, "polar"), (64, 22, "rev_polar"), (128, 64, "polar"),
               (256, 128, "polar"), (512, 256, "pw"), (1024, 512, "pw"), (1024, 512, None)]
    if big:
        configs.append((2048, 1024, "pw"))
    for N, K, prof in configs:
        n = int(np.log2(N))
        if prof in ("polar", "rev_polar"):
            code = ref_shim.get_code("Polar", prof, N, K)
        elif prof == "pw":
            code = polar.PolarCode(n, K, ref_shim.make_args(N, K), F=construct.pw_frozen_set(N, K),
                                   use_cuda=False)
        else:
            code = polar.PolarCode(n, K, ref_shim.make_args(N, K), use_cuda=False)
        B = 37 if N <= 64 else (9 if N <= 256 else (3 if N <= 1024 else 2))
        msg = bpsk_msgs(rs, B, K)
        x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
        tag = "p%d_%d_%s" % (N, K, prof or "last")
        snrs = (0.0, 2.0) if N <= 256 else (2.0,)
        for si, snr in enumerate(snrs):
            add("%s_s%d" % (tag, si), code, noisy(rs, x, snr), snr, msg=msg, x=x)
        add(tag + "_ties", code, noisy(rs, x, 0.0, ties=True), 0.0)
        if N <= 256:
            u = np.ones((B, N), dtype=np.float32)
            u[:, code.info_positions] = msg
            add(tag + "_gt", code, noisy(rs, x, 1.0), 1.0, use_gt=u)
        print("polar", tag, "done", flush=True)
    out["names"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "polar_sc.npz"), **out)


def pac_cases():
    rs = np.random.RandomState(4321)
    out = {}
    meta = []
    for N, K, g in [(32, 16, 53), (16, 8, 13), (64, 32, 53), (8, 4, 7), (128, 64, 133), (32, 16, 3)]:
        pac = ref_shim.get_code("PAC", "RM", N, K, g=g)
        B = 33 if N <= 64 else 7
        msg = bpsk_msgs(rs, B, K)
        x = pac.pac_encode(torch.from_numpy(msg)).numpy()
        for tag, snr, ties, gt in [("a", 0.0, False, False), ("b", 2.0, False, False),
                                   ("t", 0.0, True, False), ("g", 1.0, False, True)]:
            y = noisy(rs, x, snr, ties)
            name = "pac%d_%d_%d_%s" % (N, K, g, tag)
            gtc = None
            if gt:
                # genie mode wants the true u (pre-transform) sequence: conv-encoded rate-profiled msg
                v = pac.rate_profiler(torch.from_numpy(msg), scheme="RM")
                gtc = pac.convolutional_encode(v).numpy()
            llr, vh, uh = pac.pac_sc_decode(torch.from_numpy(y), snr,
                                            use_gt_codeword=None if gtc is None else torch.from_numpy(gtc))
            out[name + "_y"] = y
            out[name + "_snr"] = np.float64(snr)
            out[name + "_g"] = np.int64(g)
            out[name + "_info"] = np.asarray(pac.B, dtype=np.int32)
            out[name + "_msg"] = msg
            out[name + "_x"] = x
            out[name + "_llr"] = llr.numpy()
            out[name + "_v"] = vh.numpy()
            out[name + "_u"] = uh.numpy()
            if gtc is not None:
                out[name + "_gt"] = gtc
            meta.append(name)
        print("pac", N, K, g, "done", flush=True)
    out["names"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "pac_sc.npz"), **out)


def _ref_gru_logits(ra, net, dec, y, N, info, forced=None):
    """Step the reference RNN_Model exactly as RNN_decoder.decode's test branch does
    (rnn_all.py:532-547), also collecting the head output of every step."""
    net.eval()
    B = y.shape[0]
    decoded = torch.ones(B, N)
    logits = torch.zeros(B, N)
    hidden = torch.zeros(net.num_rnn_layers, B, net.feature_size)
    info = set(int(i) for i in info)
    with torch.no_grad():
        for ii in range(N):
            if ii == 0:
                prev = torch.ones(B)
            elif forced is not None:
                prev = forced[:, ii - 1]
            else:
                prev = decoded[:, ii - 1].sign()
            inp = torch.cat([y.unsqueeze(1), ra.get_onehot(prev).view(-1, 1, net.input_size - N)], 2)
            out, hidden = net(inp, hidden)
            logits[:, ii] = out.squeeze()
            if ii in info:
                decoded[:, ii] = out.squeeze().sign()
    return decoded, logits


def gru_cases():
    ra = ref_shim.load("rnn_all")
    rs = np.random.RandomState(99)
    out = {}
    meta = []
    for name, N, K, H, seed, gain, B in [("gru64", 64, 22, 512, 11, 8.0, 160), ("gru32", 32, 16, 512, 12, 8.0, 70),
                                         ("gru16_h64", 16, 8, 64, 13, 4.0, 21)]:
        code = ref_shim.get_code("Polar", "polar", N, K)
        sd = synth.gru_state_dict(seed, N, H, 2, head_gain=gain)
        net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0, out_linear_depth=1)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        dec = ra.RNN_decoder("y_input", N, code.info_positions, onehot=True)
        msg = bpsk_msgs(rs, B, K)
        x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
        y = noisy(rs, x, 1.0)
        d_ref = dec.decode(net, False, torch.from_numpy(y))
        d2, lg = _ref_gru_logits(ra, net, dec, torch.from_numpy(y), N, code.info_positions)
        assert torch.equal(d_ref, d2)
        out[name + "_y"] = y
        out[name + "_info"] = np.asarray(code.info_positions, dtype=np.int32)
        out[name + "_cfg"] = np.array([N, K, H, seed], dtype=np.int64)
        out[name + "_gain"] = np.float64(gain)
        out[name + "_decoded"] = d_ref.numpy()
        out[name + "_logits"] = lg.numpy()
        if H <= 64:  # small net: store the weights themselves too (pins synth.gru_state_dict)
            for k, v in sd.items():
                out[name + "_w_" + k] = v
        meta.append(name)
        print("gru", name, "|logit| mean", float(lg.abs().mean()), flush=True)
    out["names"] = np.array(meta)
    np.savez_compressed(os.path.join(OUT, "gru.npz"), **out)


def gru_trained_cases(ckpt="crisp_gru_N64_K22_H512", out_name="gru_trained"):
    """Logits of a TRAINED CRISP GRU from the live reference on noisy codewords at -2 / 0 / 2 dB: where |logit| ~ 1, which
    synthetic weights never reach.  Checkpoints: crisp_gru_N64_K22_H512.pt (trained by the reference on CPU,
    oracle/train_ref_checkpoint.py) and crisp_gru_N64_K22_H512_gputrained.pt (same curriculum, 5700 iterations at batch 4096
    through this repo's GPU training loop, tools/gpu_curriculum.py: BER within 1.4x of SC) -- both in the reference's
    checkpoint format, both evaluated here by the reference's own RNN_Model / RNN_decoder."""
    ra = ref_shim.load("rnn_all")
    path = os.path.join(OUT, ckpt + ".pt")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    a = ck["args"]
    N, K, H = a.N, a.K, a.rnn_feature_size
    ra.args = ref_shim.make_args(N=N, K=K, target_K=a.target_K)
    code = ref_shim.get_code("Polar", a.rate_profile, N, K, target_K=a.target_K)
    net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0)
    net.load_state_dict(ck["net"])
    dec = ra.RNN_decoder("y_input", N, code.info_inds, onehot=True)
    rs = np.random.RandomState(64)
    out = {"cfg": np.array([N, K, H], dtype=np.int64), "info": np.asarray(code.info_inds, dtype=np.int32)}
    ys, ds, ls, snrs = [], [], [], []
    for snr in (-2.0, 0.0, 2.0):
        msg = bpsk_msgs(rs, 96, K)
        y = noisy(rs, code.encode_plotkin(torch.from_numpy(msg)).numpy(), snr)
        d_ref = dec.decode(net, False, torch.from_numpy(y))
        d2, lg = _ref_gru_logits(ra, net, dec, torch.from_numpy(y), N, code.info_inds)
        assert torch.equal(d_ref, d2)
        ys.append(y); ds.append(d_ref.numpy()); ls.append(lg.numpy()); snrs += [snr] * 96
        print("gru_trained snr", snr, "|logit| mean on info", float(lg[:, code.info_inds].abs().mean()),
              "BER", float((d_ref[:, code.info_inds].numpy() != msg).mean()), flush=True)
    out.update(y=np.concatenate(ys), decoded=np.concatenate(ds), logits=np.concatenate(ls), snr=np.array(snrs))
    np.savez_compressed(os.path.join(OUT, out_name + ".npz"), **out)


TRAIN_KEYS = ("rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0", "rnn.weight_ih_l1",
              "rnn.weight_hh_l1", "rnn.bias_ih_l1", "rnn.bias_hh_l1", "linear.weight", "linear.bias")


def _blob(named):
    return np.concatenate([named[k].detach().cpu().numpy().reshape(-1) for k in TRAIN_KEYS]).astype(np.float32)


def gru_train_cases():
    """Three iterations of the reference's OWN training-loop body (rnn_all.py:1399-1437) on a seeded RNN_Model:
    decoder.decode(net, True, y, gt, tfr) -> MSELoss on the info positions -> backward -> clip_grad_norm_(0.25) ->
    AdamW.step().  tfr = 1 draws teacher forcing, tfr = 0 student forcing (`random.random() < tfr`, 425)."""
    ra = ref_shim.load("rnn_all")
    N, K, H, B = 16, 8, 64, 96
    ra.args = ref_shim.make_args(N=N, K=K)
    code = ref_shim.get_code("Polar", "polar", N, K)
    torch.manual_seed(5)
    net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0)
    dec = ra.RNN_decoder("y_input", N, code.info_inds, onehot=True)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    loss_fn = torch.nn.MSELoss()
    rs = np.random.RandomState(41)
    out = {"cfg": np.array([N, K, H, B], dtype=np.int64), "info": np.asarray(code.info_inds, dtype=np.int32),
           "lr": np.float64(1e-3), "clip": np.float64(0.25), "p0": _blob(dict(net.named_parameters()))}
    for step, tfr in enumerate([1.0, 0.0, 1.0]):
        msg = torch.from_numpy(bpsk_msgs(rs, B, K))
        gt = torch.ones(B, N)
        gt[:, code.info_inds] = msg
        y = torch.from_numpy(noisy(rs, code.encode(msg).numpy(), 0.0))
        decoded = dec.decode(net, True, y, gt, tfr)
        loss = loss_fn(decoded[:, code.info_inds], msg)
        loss.backward()
        norm = torch.nn.utils.clip_grad_norm_(net.parameters(), 0.25)
        out["s%d_y" % step] = y.numpy()
        out["s%d_gt" % step] = gt.numpy()
        out["s%d_teacher" % step] = np.int64(tfr >= 1.0)
        out["s%d_loss" % step] = np.float64(loss.item())
        out["s%d_norm" % step] = np.float64(float(norm))
        out["s%d_grad" % step] = _blob({k: p.grad for k, p in net.named_parameters()})
        out["s%d_decoded" % step] = decoded.detach().numpy()
        opt.step()
        opt.zero_grad()
        out["s%d_p" % step] = _blob(dict(net.named_parameters()))
        print("gru_train step", step, "loss", loss.item(), "norm", float(norm), flush=True)
    np.savez_compressed(os.path.join(OUT, "gru_train.npz"), **out)


def scl_cases():
    """SC-list decoder (polar.py:793-876, use_CRC=False) on real-valued noise at several list sizes."""
    rs = np.random.RandomState(2718)
    out = {}
    names = []
    for N, K, L, B, snr in [(64, 22, 4, 160, 0.0), (64, 22, 8, 60, -1.0), (32, 16, 2, 120, 1.0), (32, 16, 4, 120, 0.0),
                            (16, 8, 16, 80, 0.0), (8, 4, 32, 40, -2.0), (128, 64, 4, 40, 1.0), (256, 128, 2, 12, 2.0),
                            (64, 22, 1, 50, 0.0)]:
        code = ref_shim.get_code("Polar", "polar", N, K)
        msg = bpsk_msgs(rs, B, K)
        x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
        y = noisy(rs, x, snr)
        llr, dec = code.scl_decode(torch.from_numpy(y), snr, L, use_CRC=False)
        nm = "scl_%d_%d_L%d" % (N, K, L)
        out[nm + "_y"] = y
        out[nm + "_snr"] = np.float64(snr)
        out[nm + "_info"] = np.asarray(code.info_positions, dtype=np.int32)
        out[nm + "_msg"] = msg
        out[nm + "_llr"] = llr.numpy()
        out[nm + "_dec"] = dec.numpy()
        names.append(nm)
        print(nm, "BLER", float((dec.numpy() != msg).any(1).mean()), flush=True)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "scl.npz"), **out)


def gru_mode_cases():
    """Genie-aided decode (gt / loss_inds, rnn_all.py:519-522, 887) and the teacher- / student-forced passes that
    test_model(tf=True) runs under no_grad (rnn_all.py:982-984 -> 425-461, 462-512), from the live reference."""
    ra = ref_shim.load("rnn_all")
    ra.args = ref_shim.make_args(32, 16)
    rs = np.random.RandomState(314)
    out = {}
    N, K, H, seed, gain, B = 32, 16, 512, 12, 8.0, 48
    code = ref_shim.get_code("Polar", "polar", N, K)
    sd = synth.gru_state_dict(seed, N, H, 2, head_gain=gain)
    net = ra.RNN_Model("GRU", N + 2, H, 1, 2, N, 0, 0, out_linear_depth=1)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = ra.RNN_decoder("y_input", N, code.info_positions, onehot=True)
    msg = bpsk_msgs(rs, B, K)
    x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
    y = torch.from_numpy(noisy(rs, x, 1.0))
    gt = torch.ones(B, N)
    gt[:, code.info_positions] = torch.from_numpy(msg)
    loss_inds = np.asarray(code.info_positions)[-6:]          # the curriculum's "last bits only" (loss_only)
    genie_sub = dec.decode(net, False, y, gt, loss_inds=loss_inds)
    genie_all = dec.decode(net, False, y, gt)                   # loss_inds = info set: gt only matters on frozen (= +1)
    with torch.no_grad():
        tf = dec.decode(net, True, y, gt, 1)                    # teacher forced raw outputs, all N positions
        sf = dec.decode(net, True, y, gt, 0)                    # student forced raw outputs on info positions
    out.update(cfg=np.array([N, K, H, seed], dtype=np.int64), gain=np.float64(gain), y=y.numpy(), gt=gt.numpy(),
               info=np.asarray(code.info_positions, dtype=np.int32), loss_inds=loss_inds.astype(np.int32),
               genie_sub=genie_sub.numpy(), genie_all=genie_all.numpy(), teacher=tf.numpy(), student=sf.numpy())
    np.savez_compressed(os.path.join(OUT, "gru_modes.npz"), **out)
    print("gru modes: teacher |out| mean", float(tf.abs().mean()), flush=True)


GRU_COND = [  # name, decoding_type, onehot, reverse, y_hidden, y_depth, activation, H, seed, out_linear_depth
    ("h0_onehot", "y_h0", True, False, 48, 2, "selu", 512, 31, 1),
    ("h0_scalar", "y_h0", False, False, 40, 1, "relu", 256, 32, 1),
    ("h0_reverse", "y_h0", True, True, 48, 3, "tanh", 256, 33, 1),
    ("ynn_onehot", "y_input", True, False, 48, 2, "relu", 256, 34, 1),
    ("yin_scalar_rev", "y_input", False, True, 0, 0, "relu", 256, 35, 1),
    ("yin_head3", "y_input", True, False, 128, 0, "relu", 512, 36, 3),   # rnn_all.py:1322 with --out_linear_depth 3
    ("yin_head2", "y_input", True, False, 48, 0, "relu", 256, 37, 2),
    ("h0_head4", "y_h0", True, False, 64, 2, "selu", 128, 38, 4),
]


def gru_cond_cases():
    """The decoder's other conditionings, from the live reference: 'y_h0' (initial state = y-MLP, input = feedback only,
    rnn_all.py:523-531, model of 1317), 'y_input' through the y-MLP (use_ynn, 1320), scalar instead of one-hot
    feedback (410-411), reverse order (414-419, 558-561).  Per case: free-running decode, genie decode on the last
    six info positions, teacher-forced raw outputs (train=True under no_grad)."""
    ra = ref_shim.load("rnn_all")
    ra.args = ref_shim.make_args(32, 16)
    rs = np.random.RandomState(2718)
    N, K, B, gain = 32, 16, 40, 6.0
    code = ref_shim.get_code("Polar", "polar", N, K)
    info = np.asarray(code.info_positions)
    out = {"names": np.array([c[0] for c in GRU_COND]), "info": info.astype(np.int32), "gain": np.float64(gain)}
    for name, dtype, onehot, rev, yh, yd, act, H, seed, od in GRU_COND:
        y_h0 = dtype == "y_h0"
        in_size = (0 if y_h0 else N) + 1 + int(onehot)
        if yd > 0:
            sd = synth.gru_y_state_dict(seed, N, H, in_size, yh, yd, 2 * H if y_h0 else N, head_gain=gain)
        else:
            sd = synth.gru_state_dict(seed, in_size - 2, H, 2, head_gain=gain)
        if od > 1:
            sd = synth.with_mlp_head(sd, seed, H, yh, od, head_gain=gain)
        net = ra.RNN_Model("GRU", in_size, H, 1, 2, N, yh, yd, act, out_linear_depth=od,
                           y_output_size=None if (y_h0 or yd == 0) else N)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        dec = ra.RNN_decoder(dtype, N, info, onehot=onehot, reverse_order=rev)
        msg = bpsk_msgs(rs, B, K)
        x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
        y = torch.from_numpy(noisy(rs, x, 1.0))
        gt = torch.ones(B, N)
        gt[:, info] = torch.from_numpy(msg)
        free = dec.decode(net, False, y)
        genie = dec.decode(net, False, y, gt, loss_inds=info[-6:])
        with torch.no_grad():
            teacher = dec.decode(net, True, y, gt, 1)
        out[name + "_cfg"] = np.array([H, seed, yh, yd, int(onehot), int(rev), od], dtype=np.int64)
        out[name + "_type"] = np.array(dtype)
        out[name + "_act"] = np.array(act)
        out[name + "_y"] = y.numpy()
        out[name + "_gt"] = gt.numpy()
        out[name + "_free"] = free.numpy()
        out[name + "_genie"] = genie.numpy()
        out[name + "_teacher"] = teacher.numpy()
        print("gru cond", name, "teacher |out| mean", float(teacher.abs().mean()), flush=True)
    np.savez_compressed(os.path.join(OUT, "gru_cond.npz"), **out)


def conv_cases():
    import argparse as ap
    md = ref_shim.load("models")
    rs = np.random.RandomState(7)
    out = {}
    for name, N, K, E, seed, B in [("conv64", 64, 22, 128, 21, 72)]:
        code = ref_shim.get_code("Polar", "polar", N, K)
        sd = synth.conv_state_dict(seed, N, E)
        net = md.convNet(ap.Namespace(embed_dim=E, max_len=N, N=N, dont_use_bias=False, dropout=0.1))
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        net.eval()
        msg = bpsk_msgs(rs, B, K)
        x = code.encode_plotkin(torch.from_numpy(msg)).numpy()
        y = noisy(rs, x, 1.0)
        with torch.no_grad():
            probs, bits, _, logits, in4 = net(torch.from_numpy(y), None, None, "cpu")
        out[name + "_y"] = y
        out[name + "_cfg"] = np.array([N, K, E, seed], dtype=np.int64)
        out[name + "_logits"] = logits.squeeze(-1).numpy()
        out[name + "_bits"] = bits.squeeze(-1).numpy()
        print("conv", name, "|logit| mean", float(logits.abs().mean()), flush=True)
    out["names"] = np.array(["conv64"])
    np.savez_compressed(os.path.join(OUT, "conv.npz"), **out)


def conv_trained_cases():
    """Logits of the reference-TRAINED convNet (tests/golden/conv_N64_K22_E128.pt, oracle/train_ref_conv_checkpoint.py)
    from the live reference's forward on noisy codewords at -2 / 0 / 2 dB."""
    md = ref_shim.load("models")
    ck = torch.load(os.path.join(OUT, "conv_N64_K22_E128.pt"), map_location="cpu", weights_only=False)
    a = ck["args"]
    net = md.convNet(a)
    net.load_state_dict(ck["xformer"])
    net.eval()
    code = ref_shim.get_code("Polar", "polar", a.N, a.K)
    rs = np.random.RandomState(65)
    ys, ls, bs = [], [], []
    for snr in (-2.0, 0.0, 2.0):
        msg = bpsk_msgs(rs, 96, a.K)
        y = noisy(rs, code.encode_plotkin(torch.from_numpy(msg)).numpy(), snr)
        with torch.no_grad():
            _, bits, _, logits, _ = net(torch.from_numpy(y), None, None, "cpu")
        ys.append(y); ls.append(logits.squeeze(-1).numpy()); bs.append(bits.squeeze(-1).numpy())
        print("conv_trained snr", snr, "|logit| mean", float(logits.abs().mean()),
              "BER", float((bits.squeeze(-1)[:, code.info_positions].numpy() != msg).mean()), flush=True)
    np.savez_compressed(os.path.join(OUT, "conv_trained.npz"), y=np.concatenate(ys), logits=np.concatenate(ls),
                        bits=np.concatenate(bs), info=np.asarray(code.info_positions, dtype=np.int32))


def misc_cases():
    """Info sets and the reference's error counters on small hand-made inputs."""
    ra = ref_shim.load("rnn_all")
    out = {}
    for N, K, prof in [(64, 22, "polar"), (64, 22, "rev_polar"), (32, 16, "polar"), (16, 8, "polar"),
                       (128, 64, "polar"), (256, 128, "polar"), (64, 22, "RM")]:
        code = ref_shim.get_code("Polar", prof, N, K)
        out["info_%s_%d_%d" % (prof, N, K)] = np.asarray(code.info_positions, dtype=np.int32)
    # curriculum stages K < target_K for every rate profile (rnn_all.py:1075-1181) and --loss_only (1189-1192)
    for prof in ("polar", "sorted", "sorted_last", "rev_polar", "random"):
        for N, K, tK in [(64, 8, 22), (64, 14, 22), (32, 6, 16)]:
            code = ref_shim.get_code("Polar", prof, N, K, target_K=tK, random_seed=42)
            out["cur_%s_%d_%d_%d" % (prof, N, K, tK)] = np.asarray(code.info_positions, dtype=np.int32)
    code = ref_shim.get_code("Polar", "rev_polar", 64, 22, target_K=22, loss_only=6)
    out["lossonly_inds"] = np.asarray(code.loss_inds, dtype=np.int32)
    out["lossonly_msg"] = np.asarray(code.msg_indices, dtype=np.int32)
    for N, K, g in [(32, 16, 53), (64, 32, 53), (128, 64, 133)]:
        pac = ref_shim.get_code("PAC", "RM", N, K, g=g)
        out["pacinfo_%d_%d" % (N, K)] = np.asarray(pac.B, dtype=np.int32)
        out["pacg_%d" % g] = np.asarray(pac.g_array, dtype=np.int32)
    rs = np.random.RandomState(5)
    a = (1.0 - 2.0 * rs.randint(0, 2, size=(50, 22))).astype(np.float32)
    b = a.copy()
    flip = rs.rand(50, 22) < 0.07
    b[flip] *= -1
    b[3, 4] = 0.0   # a tie counts as an error (round(0) != +-1)
    b[9, :] = a[9, :]
    out["err_a"] = a
    out["err_b"] = b
    out["err_ber"] = np.float64(ra.errors_ber(torch.from_numpy(a), torch.from_numpy(b)).item())
    out["err_bler"] = np.float64(ra.errors_bler(torch.from_numpy(a), torch.from_numpy(b)))
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--big", action="store_true", help="also run the reference at N=2048 (minutes)")
    p.add_argument("--only", default="")
    a = p.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    todo = a.only.split(",") if a.only else ["misc", "pac", "gru", "gru_modes", "gru_cond", "scl", "conv", "polar"]
    if "gru_train" in todo:
        gru_train_cases()
    if "gru_trained" in todo:
        gru_trained_cases()
    if "gru_trained_gpu" in todo:
        gru_trained_cases("crisp_gru_N64_K22_H512_gputrained", "gru_trained_gpu")
    if "gru_trained_gpu_tenth" in todo:  # tools/gpu_curriculum.py --train_gemm tf32 --steps 500 --final_steps 10000
        gru_trained_cases("crisp_gru_N64_K22_H512_gputrained_tenth", "gru_trained_gpu_tenth")
    if "conv_trained" in todo:
        conv_trained_cases()
    if "gru_cond" in todo:
        gru_cond_cases()
    if "gru_modes" in todo:
        gru_mode_cases()
    if "scl" in todo:
        scl_cases()
    if "misc" in todo:
        misc_cases()
    if "pac" in todo:
        pac_cases()
    if "gru" in todo:
        gru_cases()
    if "conv" in todo:
        conv_cases()
    if "polar" in todo:
        polar_cases(a.big)
    print("golden fixtures written to", OUT)
