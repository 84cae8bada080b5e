"""CPU: the oracle (oracle/npd_oracle.c + oracle/oracle.py) against the fixtures minted from the live
reference (oracle/gen_golden.py).  This is what pins the oracle; bit-exact for SC / encoder / counters,
fp32 round-off for the GRU / CNN restatements."""
import numpy as np
import pytest

import oracle


def _n(N):
    return int(np.log2(N))


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32_10
    assert [hex(v) for v in oracle.philox4x32_10([0, 0, 0, 0], [0, 0])] == \
        ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(v) for v in oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(v) for v in oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                                 [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_survey_kat_values(golden):
    """SURVEY.md App. B KAT1-4 as literal numbers (independent of the fixture file)."""
    info = np.array([4, 5, 6, 7], dtype=np.int32)
    x = oracle.polar_encode(np.array([[-1, 1, 1, -1], [1, -1, -1, -1]], np.float32), 3, info)
    assert x.tolist() == [[1, -1, -1, -1, 1, -1, -1, -1], [-1, 1, 1, -1, -1, 1, 1, -1]]
    llr, _, dec = oracle.sc_decode(np.array([[1, -1, 1, 1, 1, 1, 1, 1]], np.float32), 0.0, 3, info)
    assert llr.tolist() == [[998, 1000, 1000, 1004, 0, 0, 4, 8]]
    assert dec.tolist() == [[0, 0, 1, 1]]
    llr, _, dec = oracle.sc_decode(np.array([[0, -1]], np.float32), 0.0, 1, np.array([0, 1], np.int32))
    assert llr.tolist() == [[0, -2]] and dec.tolist() == [[0, -1]]
    y0 = np.array([[0.9, -1.2, 0.3, -0.4, -1.1, 0.8, -0.2, 1.5]], np.float32)
    llr, _, dec = oracle.sc_decode(y0, 1.0, 3, info)
    np.testing.assert_allclose(llr[0], [1000.50354, 1001.51068, 1002.76965, 994.20892, 0.25178510,
                                        -1.25892544, 0.75535518, 4.53213167], rtol=1e-7)
    assert dec.tolist() == [[1, -1, 1, 1]]


def test_polar_sc_against_reference_fixtures(golden):
    g = golden("polar_sc")
    names = [str(s) for s in g["names"]]
    assert len(names) > 30
    ties = 0
    for nm in names:
        y, info, snr = g[nm + "_y"], g[nm + "_info"], float(g[nm + "_snr"])
        n = _n(y.shape[1])
        gt = g[nm + "_gt"] if nm + "_gt" in g.files else None
        llr, _, dec = oracle.sc_decode(y, snr, n, info, use_gt=gt)
        assert np.array_equal(llr, g[nm + "_llr"]), nm
        assert np.array_equal(dec, g[nm + "_dec"]), nm
        ties += int((g[nm + "_dec"] == 0).sum())
        if nm + "_msg" in g.files:
            assert np.array_equal(oracle.polar_encode(g[nm + "_msg"], n, info), g[nm + "_x"]), nm
    assert ties > 0, "fixtures must exercise sign(0) = 0"


def test_pac_against_reference_fixtures(golden):
    g = golden("pac_sc")
    for nm in [str(s) for s in g["names"]]:
        y, info, snr, gp = g[nm + "_y"], g[nm + "_info"], float(g[nm + "_snr"]), int(g[nm + "_g"])
        n = _n(y.shape[1])
        assert np.array_equal(oracle.pac_encode(g[nm + "_msg"], n, info, gp), g[nm + "_x"]), nm
        gt = g[nm + "_gt"] if nm + "_gt" in g.files else None
        llr, v, u = oracle.pac_sc_decode(y, snr, n, info, gp, use_gt_codeword=gt)
        assert np.array_equal(llr, g[nm + "_llr"]), nm
        assert np.array_equal(v, g[nm + "_v"]), nm
        assert np.array_equal(u, g[nm + "_u"]), nm


def test_error_counters(golden):
    g = golden("misc")
    a, b = g["err_a"], g["err_b"]
    bit, blk = oracle.count_errors(a, b)
    assert bit / a.size == pytest.approx(float(g["err_ber"]), abs=1e-7)
    assert blk / a.shape[0] == pytest.approx(float(g["err_bler"]), abs=1e-12)


def test_gru_oracle_against_reference_fixtures(golden):
    from neural_polar_decoder_b200 import synth
    g = golden("gru")
    for nm in [str(s) for s in g["names"]]:
        N, K, H, seed = [int(v) for v in g[nm + "_cfg"]]
        sd = synth.gru_state_dict(seed, N, H, 2, head_gain=float(g[nm + "_gain"]))
        if H <= 64:  # weights stored in the fixture: pins synth's determinism
            for k, v in sd.items():
                assert np.array_equal(v, g[nm + "_w_" + k]), k
        dec, lg = oracle.gru_decode(sd, g[nm + "_y"], N, g[nm + "_info"], forced=g[nm + "_decoded"])
        np.testing.assert_allclose(lg, g[nm + "_logits"], rtol=0, atol=2e-5)
        # decisions identical wherever the reference logit is not within round-off of zero
        safe = np.abs(g[nm + "_logits"]) > 1e-4
        assert np.array_equal(dec[safe], g[nm + "_decoded"][safe])


@pytest.mark.parametrize("ckpt,fixture", [("crisp_gru_N64_K22_H512", "gru_trained"),
                                          ("crisp_gru_N64_K22_H512_gputrained", "gru_trained_gpu"),
                                          ("crisp_gru_N64_K22_H512_gputrained_tenth", "gru_trained_gpu_tenth")])
def test_gru_oracle_on_trained_checkpoints(golden, ckpt, fixture):
    """The oracle is the checker of bench.py's in-bench parity test on the TRAINED Polar(64,22), H = 512 checkpoints: pin it on
    those weights too -- forced-feedback logits and free-running decisions against what the live reference produced
    (oracle/gen_golden.py gru_trained_cases), at fp32 round-off."""
    import os
    import torch
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, ckpt + ".pt")
    if not os.path.exists(path):
        pytest.skip("no %s checkpoint" % ckpt)
    g = golden(fixture)
    N, K, H = [int(v) for v in g["cfg"]]
    sd = torch.load(path, map_location="cpu", weights_only=False)["net"]
    rows = slice(0, 288, 3)  # 96 of the 288 frames, all three SNR points
    dec, lg = oracle.gru_decode(sd, g["y"][rows], N, g["info"], forced=g["decoded"][rows])
    np.testing.assert_allclose(lg, g["logits"][rows], rtol=0, atol=3e-5)
    free, _ = oracle.gru_decode(sd, g["y"][rows], N, g["info"])
    safe = np.cumsum(np.abs(g["logits"][rows]) <= 1e-4, axis=1) == 0  # no logit within round-off of zero so far
    assert np.array_equal(free[safe], g["decoded"][rows][safe])


def test_conv_oracle_against_reference_fixtures(golden):
    from neural_polar_decoder_b200 import synth
    g = golden("conv")
    N, K, E, seed = [int(v) for v in g["conv64_cfg"]]
    lg = oracle.conv_forward(synth.conv_state_dict(seed, N, E), g["conv64_y"])
    np.testing.assert_allclose(lg, g["conv64_logits"], rtol=0, atol=2e-5)


def test_gru_oracle_modes_against_reference_fixture(golden):
    """Genie-aided decode and the teacher- / student-forced evaluation passes (rnn_all.py:519-522, 425-512)."""
    from neural_polar_decoder_b200 import synth
    g = golden("gru_modes")
    N, K, H, seed = [int(v) for v in g["cfg"]]
    sd = synth.gru_state_dict(seed, N, H, 2, head_gain=float(g["gain"]))
    y, gt, info = g["y"], g["gt"], g["info"]
    # teacher forcing: every step is fed gt -> the raw head outputs are the logits under forced = gt
    _, lg = oracle.gru_decode(sd, y, N, info, forced=gt)
    np.testing.assert_allclose(lg, g["teacher"], rtol=0, atol=2e-5)
    # student forcing: raw outputs on the info positions, ones elsewhere, own decisions fed back
    dec, lg = oracle.gru_decode(sd, y, N, info)
    student = np.where(np.isin(np.arange(N), info)[None, :], lg, 1.0)
    safe_rows = (np.abs(lg[:, info]) > 1e-4).all(axis=1)
    np.testing.assert_allclose(student[safe_rows], g["student"][safe_rows], rtol=0, atol=2e-5)
    # genie: decoded starts as gt; only loss_inds are decided (and fed back as decided)
    for key, loss in (("genie_sub", g["loss_inds"]), ("genie_all", info)):
        dec, lg = oracle.gru_decode(sd, y, N, loss, genie=gt)
        safe_rows = (np.abs(lg[:, loss]) > 1e-4).all(axis=1)
        assert safe_rows.mean() > 0.9
        assert np.array_equal(dec[safe_rows], g[key][safe_rows]), key


def _cond_case(g, nm):
    """(state_dict, decode kwargs, y as the kernel sees it, loss-step helper) of one gru_cond.npz case."""
    from neural_polar_decoder_b200 import synth
    H, seed, yh, yd, onehot, rev, od = [int(v) for v in g[nm + "_cfg"]]
    y_h0 = str(g[nm + "_type"]) == "y_h0"
    N = g[nm + "_y"].shape[1]
    in_size = (0 if y_h0 else N) + 1 + onehot
    gain = float(g["gain"])
    if yd > 0:
        sd = synth.gru_y_state_dict(seed, N, H, in_size, yh, yd, 2 * H if y_h0 else N, head_gain=gain)
    else:
        sd = synth.gru_state_dict(seed, in_size - 2, H, 2, head_gain=gain)
    if od > 1:
        sd = synth.with_mlp_head(sd, seed, H, yh, od, head_gain=gain)
    kw = dict(onehot=bool(onehot), y_in=not y_h0)
    y = g[nm + "_y"]
    if y_h0:
        kw["h0"] = oracle.gru_h0(sd, y, yd, H, activation=str(g[nm + "_act"]))
    elif yd > 0:
        y = oracle.y_mlp(sd, y, yd, str(g[nm + "_act"])).numpy()
    return sd, kw, y, bool(rev)


def test_gru_oracle_conditionings_against_reference_fixture(golden):
    """'y_h0' (initial state from the y-MLP), use_ynn, scalar feedback, reverse order and MLP heads (rnn_all.py:335-343,
    410-419, 523-531, 1317-1322) against the live reference: teacher-forced raw outputs, free-running and genie decodes."""
    g = golden("gru_cond")
    info = g["info"]
    for nm in [str(s) for s in g["names"]]:
        sd, kw, y, rev = _cond_case(g, nm)
        N = y.shape[1]
        fl = (lambda a: np.ascontiguousarray(a[:, ::-1])) if rev else (lambda a: a)
        gt = fl(g[nm + "_gt"])                       # the reference flips gt on entry (417-419) ...
        _, lg = oracle.gru_decode(sd, y, N, info, forced=gt, **kw)
        np.testing.assert_allclose(fl(lg), g[nm + "_teacher"], rtol=0, atol=2e-5, err_msg=nm)  # ... and its result on exit
        steps = N - 1 - info if rev else info        # `jj in loss_inds`, jj = N-1-ii
        dec, lg = oracle.gru_decode(sd, y, N, steps, **kw)
        safe = (np.abs(lg[:, steps]) > 1e-4).all(axis=1)
        assert safe.mean() > 0.9 and np.array_equal(fl(dec)[safe], g[nm + "_free"][safe]), nm
        loss = info[-6:]
        steps = N - 1 - loss if rev else loss
        dec, lg = oracle.gru_decode(sd, y, N, steps, genie=gt, **kw)
        safe = (np.abs(lg[:, steps]) > 1e-4).all(axis=1)
        assert safe.mean() > 0.9 and np.array_equal(fl(dec)[safe], g[nm + "_genie"][safe]), nm


def test_scl_oracle_against_reference_fixtures(golden):
    """SC-list restatement vs the live reference's scl_decode (polar.py:793-876): chosen path's decisions and
    leaf LLRs, bit for bit, at list sizes 1..32."""
    g = golden("scl")
    for nm in [str(s) for s in g["names"]]:
        N, K, L = [int(v[1:]) if v.startswith("L") else int(v) for v in nm.split("_")[1:]]
        n = int(np.log2(N))
        llr, dec = oracle.scl_decode(g[nm + "_y"], float(g[nm + "_snr"]), n, g[nm + "_info"], L)
        assert np.array_equal(dec, g[nm + "_dec"]), nm
        assert np.array_equal(llr, g[nm + "_llr"]), nm
        if L == 1:  # a list of one is the SC decoder (without the frozen prior in the recursion)
            _, _, sc = oracle.sc_decode(g[nm + "_y"], float(g[nm + "_snr"]), n, g[nm + "_info"])
            assert np.array_equal(dec, sc)


def test_gru_train_step_against_reference_fixture(golden):
    """oracle.gru_train_step (gate-by-gate autograd restatement) against three iterations of the live reference's
    training-loop body: loss, clipped gradients, updated parameters (teacher-, student-, teacher-forced)."""
    g = golden("gru_train")
    N, K, H, B = [int(v) for v in g["cfg"]]
    blob, state = g["p0"], None
    for s in range(3):
        new, grad, loss, norm, logits, state = oracle.gru_train_step(
            blob, g["s%d_y" % s], g["s%d_gt" % s], N, H, g["info"], bool(g["s%d_teacher" % s]), float(g["lr"]),
            float(g["clip"]), state)
        assert loss == pytest.approx(float(g["s%d_loss" % s]), rel=1e-5)
        assert norm == pytest.approx(float(g["s%d_norm" % s]), rel=1e-4)
        ref_grad = g["s%d_grad" % s]
        assert np.abs(grad - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max() + 1e-9
        assert np.abs(new - g["s%d_p" % s]).max() <= 2e-6
        if bool(g["s%d_teacher" % s]):
            np.testing.assert_allclose(logits, g["s%d_decoded" % s], atol=2e-6)
        blob = new
    assert state[2] == 3
