"""GPU: the CRISP GRU training step (csrc/gru_train.cu) against three iterations of the LIVE reference's
training-loop body (tests/golden/gru_train.npz: loss, clipped gradients, AdamW-updated parameters; teacher-, student-,
teacher-forced) and against the oracle's autograd restatement at the flagship shape.  fp32 tolerances."""
import argparse

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _net(N, H, blob):
    from neural_polar_decoder_b200.rnn_all import RNN_Model
    from neural_polar_decoder_b200.train import PARAM_KEYS
    net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
    sd, o = net.state_dict(), 0
    for k in PARAM_KEYS:
        n = sd[k].numel()
        sd[k].copy_(torch.from_numpy(np.asarray(blob[o:o + n]).reshape(tuple(sd[k].shape))))
        o += n
    return net


def _relerr(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_train_steps_match_live_reference_fixture(golden):
    from neural_polar_decoder_b200.rnn_all import RNN_decoder
    from neural_polar_decoder_b200.train import GRUTrainer
    g = golden("gru_train")
    N, K, H, B = [int(v) for v in g["cfg"]]
    net = _net(N, H, g["p0"])
    dec = RNN_decoder('y_input', N, g["info"], onehot=True)
    tr = GRUTrainer(net, N, B)
    code = dec._loss_code(g["info"])
    prev = g["p0"].astype(np.float64)
    m = np.zeros_like(prev)
    v = np.zeros_like(prev)
    for s in range(3):
        teacher = bool(g["s%d_teacher" % s])
        loss, norm, logits = tr.step(code, torch.from_numpy(g["s%d_y" % s]).cuda(), torch.from_numpy(g["s%d_gt" % s]).cuda(),
                                     teacher, float(g["lr"]), float(g["clip"]), want_logits=True)
        assert loss == pytest.approx(float(g["s%d_loss" % s]), rel=2e-5), s
        assert norm == pytest.approx(float(g["s%d_norm" % s]), rel=2e-4), s
        if teacher:  # decoded_vhat of the teacher-forced pass = every step's logit
            np.testing.assert_allclose(logits.cpu().numpy(), g["s%d_decoded" % s], atol=5e-6)
        grad, ref = tr.get("grads"), g["s%d_grad" % s]
        assert _relerr(grad, ref) <= 2e-4, (s, _relerr(grad, ref))
        p, ref_p = tr.get("params"), g["s%d_p" % s]
        # (a) the update arithmetic: torch.optim.AdamW recomputed on the host from OUR clipped gradient
        lr, b1, b2, eps, wd = float(g["lr"]), 0.9, 0.999, 1e-8, 0.01
        m = b1 * m + (1 - b1) * grad.astype(np.float64)
        v = b2 * v + (1 - b2) * grad.astype(np.float64) ** 2
        expect = prev * (1 - lr * wd) - lr / (1 - b1 ** (s + 1)) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** (s + 1)) + eps)
        assert np.abs(p - expect).max() <= 2e-7, (s, np.abs(p - expect).max())
        # (b) end to end against the reference's parameters: Adam normalises every entry's step to ~lr, so entries whose
        # gradient is at round-off level may land lr apart; everything else must agree to fp32 round-off
        d = np.abs(p - ref_p)
        assert np.median(d) <= 1e-7 and np.mean(d > 1e-5) < 0.02, (s, float(np.median(d)), float(np.mean(d > 1e-5)))
        prev = p.astype(np.float64)
    # the nn.Module receives the trained weights and the decode kernel then runs on them
    tr.sync_to_net()
    from neural_polar_decoder_b200.train import _blob
    assert np.array_equal(_blob(net), tr.get("params"))


@pytest.mark.parametrize("teacher", [True, False])
def test_train_step_flagship_shape_vs_oracle(teacher):
    """Polar(64,22), H = 512 (run_crisp.sh): gradients of one iteration against the oracle's autograd restatement."""
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.train import GRUTrainer, _blob
    N, K, H, B = 64, 22, 512, 48
    code = get_code("Polar", "rev_polar", N, K, args=argparse.Namespace(target_K=22))
    sd = synth.gru_state_dict(5, N, H, 2, head_gain=2.0)
    net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', N, code.info_inds, onehot=True)
    rs = np.random.RandomState(3)
    msg = (1.0 - 2.0 * rs.randint(0, 2, size=(B, K))).astype(np.float32)
    gt = np.ones((B, N), np.float32)
    gt[:, code.info_inds] = msg
    y = (code.encode(torch.from_numpy(msg).cuda()).cpu().numpy() + rs.randn(B, N)).astype(np.float32)
    blob = _blob(net)
    new, grad, loss, norm, logits, _ = oracle.gru_train_step(blob, y, gt, N, H, code.info_inds, teacher, 1e-3, 0.25)
    tr = GRUTrainer(net, N, B)
    l2, n2, lg = tr.step(dec._loss_code(code.info_inds), torch.from_numpy(y).cuda(), torch.from_numpy(gt).cuda(), teacher,
                         1e-3, 0.25, want_logits=True)
    assert l2 == pytest.approx(loss, rel=5e-5) and n2 == pytest.approx(norm, rel=5e-4)
    if teacher:
        np.testing.assert_allclose(lg.cpu().numpy(), logits, atol=2e-5)
    assert _relerr(tr.get("grads"), grad) <= 5e-4
    d = np.abs(tr.get("params") - new)  # see test_train_steps_match_live_reference_fixture (b)
    assert np.median(d) <= 1e-7 and np.mean(d > 1e-5) < 0.02, (float(np.median(d)), float(np.mean(d > 1e-5)))


@pytest.mark.parametrize("teacher", [True, False])
def test_train_step_ragged_shapes_vs_oracle(teacher):
    """Batch not a multiple of the gate-gradient kernel's 8-row groups, smaller than the trainer's max_batch (the all-steps
    operands are laid out for the CALL's batch), H not a multiple of the warp / block widths; two calls on one trainer."""
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.train import GRUTrainer, _blob
    N, K, H = 16, 7, 24
    code = get_code("Polar", "polar", N, K)
    sd = synth.gru_state_dict(9, N, H, 2, head_gain=2.0)
    net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    dec = RNN_decoder('y_input', N, code.info_inds, onehot=True)
    tr = GRUTrainer(net, N, 40)
    blob = _blob(net)
    rs = np.random.RandomState(4)
    for B in (13, 37, 1):
        msg = (1.0 - 2.0 * rs.randint(0, 2, size=(B, K))).astype(np.float32)
        gt = np.ones((B, N), np.float32)
        gt[:, code.info_inds] = msg
        y = (code.encode(torch.from_numpy(msg).cuda()).cpu().numpy() + rs.randn(B, N)).astype(np.float32)
        _, grad, loss, norm, logits, _ = oracle.gru_train_step(blob, y, gt, N, H, code.info_inds, teacher, 1e-3, 0.25)
        l2, n2, lg = tr.step(dec._loss_code(code.info_inds), torch.from_numpy(y).cuda(), torch.from_numpy(gt).cuda(), teacher,
                             1e-3, 0.25, want_logits=True, apply_update=False)
        assert l2 == pytest.approx(loss, rel=5e-5) and n2 == pytest.approx(norm, rel=5e-4), B
        if teacher:
            np.testing.assert_allclose(lg.cpu().numpy(), logits, atol=2e-5)
        ours = tr.get("grads") * min(1.0, 0.25 / (n2 + 1e-6))  # apply_update=False leaves the gradient unclipped
        assert _relerr(ours, grad) <= 5e-4, (B, _relerr(ours, grad))


def test_training_loop_learns_and_writes_reference_checkpoint(tmp_path, monkeypatch):
    """`python -m neural_polar_decoder_b200.rnn_all <run_crisp.sh-style flags>` without --test: a short K = 4 stage trains
    on the GPU (loss falls), leaves {'net','step','args'} where the reference would, and the TESTING block then loads it."""
    from neural_polar_decoder_b200 import cli
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    import random
    random.seed(0)
    argv = ("--code Polar --rate_profile rev_polar --target_K 8 --N 16 --K 4 --decoding_type y_input --rnn_feature_size 128 "
            "--num_steps 150 --batch_size 512 --rnn_depth 2 --tfr_min 1 --tfr_max 1 --dec_train_snr 0 --lr 0.003 "
            "--scheduler step --lr_decay 100 --lr_decay_gamma 0.5 --onehot --id t1 --print_freq 50 --test_size 4000 "
            "--test_batch_size 2000 --test_snr_start 0 --test_snr_end 4 --snr_points 3 --fresh").split()
    args = cli.get_args(argv)
    from neural_polar_decoder_b200.train import run_train
    losses = run_train(args, out=lambda *a: None)
    assert losses[-1][1] < 0.6 * losses[0][1], losses
    results, final = cli.result_paths(args)
    ck = cli.load_checkpoint(results + "/Models/model_final.pt")
    assert set(ck) == {"net", "step", "args"} and ck["step"] == 150
    assert cli.load_checkpoint(final)["step"] == 150
    res = cli.run_test(args, out=lambda *a: None)
    assert res["bers_RNN"][-1] < 0.2 and res["step"] == 150


@pytest.mark.parametrize("mode,rel", [(1, 2e-2), (2, 1e-1), (3, 3e-2)])
def test_train_step_tensor_core_gemm_modes(mode, rel):
    """tf32 / bf16 / fp16 tensor-core GEMMs (fp32 data, accumulation and optimizer): gradients within the mode's operand
    precision of the fp32 parity mode's, same loss to 3 digits."""
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.train import GRUTrainer
    N, K, H, B = 32, 16, 256, 256
    code = get_code("Polar", "polar", N, K)
    sd = synth.gru_state_dict(5, N, H, 2, head_gain=2.0)
    rs = np.random.RandomState(3)
    msg = (1.0 - 2.0 * rs.randint(0, 2, size=(B, K))).astype(np.float32)
    gt = np.ones((B, N), np.float32)
    gt[:, code.info_inds] = msg
    y = torch.from_numpy((code.encode(torch.from_numpy(msg).cuda()).cpu().numpy() + rs.randn(B, N)).astype(np.float32)).cuda()
    res = {}
    for m in (0, mode):
        net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        dec = RNN_decoder('y_input', N, code.info_inds, onehot=True)
        tr = GRUTrainer(net, N, B, tf32=m)
        loss, norm, _ = tr.step(dec._loss_code(code.info_inds), y, torch.from_numpy(gt).cuda(), True, 1e-3, 0.25, apply_update=False)
        res[m] = (loss, norm, tr.get("grads"))
    assert res[mode][0] == pytest.approx(res[0][0], rel=2e-3)
    assert _relerr(res[mode][2], res[0][2]) <= rel, _relerr(res[mode][2], res[0][2])


@pytest.mark.parametrize("teacher", [True, False])
def test_train_fused_tf32_layer_step_matches_library_path(teacher):
    """GEMM mode 1, H % 128 == 0, batch % 128 == 0: the forward layer-steps run on the fused tcgen05 kernel
    (csrc/gru_train_tc.cuh: TMA-fed kind::tf32 MMAs, gates in the epilogue, TMA stores of the saved quantities).  The same rows
    inside a batch of 257 take the library-GEMM + gate-kernel path; logits of a row do not depend on the batch, so the two
    must agree to TF32 round-off, and both must sit within TF32 precision of the fp32 parity mode."""
    from neural_polar_decoder_b200 import synth
    from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder, get_code
    from neural_polar_decoder_b200.train import GRUTrainer
    N, K, H, B = 32, 16, 256, 256
    code = get_code("Polar", "polar", N, K)
    sd = synth.gru_state_dict(6, N, H, 2, head_gain=2.0)
    rs = np.random.RandomState(8)
    msg = (1.0 - 2.0 * rs.randint(0, 2, size=(B + 1, K))).astype(np.float32)
    gt = np.ones((B + 1, N), np.float32)
    gt[:, code.info_inds] = msg
    y = torch.from_numpy((code.encode(torch.from_numpy(msg).cuda()).cpu().numpy() + rs.randn(B + 1, N)).astype(np.float32)).cuda()
    gt = torch.from_numpy(gt).cuda()
    out = {}
    for name, mode, rows in (("fused", 1, B), ("library", 1, B + 1), ("fp32", 0, B)):
        net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        dec = RNN_decoder('y_input', N, code.info_inds, onehot=True)
        tr = GRUTrainer(net, N, B + 1, tf32=mode)
        loss, norm, logits = tr.step(dec._loss_code(code.info_inds), y[:rows].contiguous(), gt[:rows].contiguous(), teacher, 1e-3,
                                     0.25, apply_update=False, want_logits=True)
        out[name] = (loss, logits[:B].cpu().numpy(), tr.get("grads"))
    if teacher:  # student forcing feeds decisions back: a TF32-sized logit difference near zero may flip a later input
        assert np.abs(out["fused"][1] - out["library"][1]).max() <= 5e-3
        assert np.abs(out["fused"][1] - out["fp32"][1]).max() <= 5e-3
    assert out["fused"][0] == pytest.approx(out["fp32"][0], rel=2e-3)
    assert _relerr(out["fused"][2], out["fp32"][2]) <= 2e-2


def test_training_loop_on_a_pac_code(tmp_path, monkeypatch):
    """BASELINE config 3's code family: the same training loop on PAC(16,8) (RM profile, convolutional pre-coder): the loss set
    is the PAC rate profile, the genie tensor carries the message on it; loss falls and the TESTING block runs the PAC
    branch (pac_sc_decode next to the GRU)."""
    from neural_polar_decoder_b200 import cli
    from neural_polar_decoder_b200.train import run_train
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    import random
    random.seed(0)
    argv = ("--code PAC --rate_profile RM --N 16 --K 8 --decoding_type y_input --rnn_feature_size 128 --num_steps 300 "
            "--batch_size 512 --rnn_depth 2 --tfr_min 1 --tfr_max 1 --dec_train_snr 1 --lr 0.003 --onehot --id tpac --print_freq 50 "
            "--test_size 4000 --test_batch_size 2000 --test_snr_start 0 --test_snr_end 4 --snr_points 3 --fresh "
            "--train_gemm tf32").split()
    args = cli.get_args(argv)
    losses = run_train(args, out=lambda *a: None)
    assert losses[-1][1] < 0.85 * losses[0][1], losses  # (the channel's Philox stream depends on what ran before: keep a margin)
    res = cli.run_test(args, out=lambda *a: None)
    assert res["step"] == 300 and res["bers_RNN"][-1] < 0.35
