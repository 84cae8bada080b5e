"""CLI / checkpoint compatibility with the reference's rnn_all.py: argument surface, derived
fields, the results-path scheme, checkpoint loading.  CPU tests use known answers taken from the reference
(run_crisp.sh embeds the final-net file name its own first line produces; the results path was printed by
the live reference during oracle/train_ref_checkpoint.py)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neural_polar_decoder_b200 import cli  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# run_crisp.sh line 1 (reference), verbatim flags
CRISP_LINE1 = ("--code Polar --rate_profile rev_polar --target_K 22 --N 64 --K 8 --decoding_type y_input "
               "--rnn_feature_size 512 --y_hidden_size 128 --y_depth 3 --num_steps 10000 --batch_size 4096 --rnn_depth 2 "
               "--model_save_per 10000 --tfr_min 1 --tfr_max 1 --dec_train_snr 0 --lr 0.001 --scheduler step "
               "--lr_decay 2000 --lr_decay_gamma 0.95 --onehot --id run1 --test_codes y --test_bitwise --testing_snr 0 "
               "--progressive_path Supervised_RNN_Polar_Results/progressive/N64_K22_H2E --test_snr_start -3 "
               "--test_snr_end 3").split()
# ... and the --load_path run_crisp.sh line 2 uses for the net that line 1 saved
CRISP_LINE1_FINAL = ("./Supervised_RNN_Polar_Results/final_nets/Scheme_rev_polar_22/N64_K8_y_input_onehot_GRU_depth_2_"
                     "fsize_512_y_depth_0_hsize_0_snr_0.0_bs_4096_tfr_1.0_activ_selu_init_He_optim_AdamW_lr_0.001_"
                     "decay_2000_step_loss_MSE_run1.pt")


def test_run_crisp_line_parses_and_derives_fields():
    a = cli.get_args(CRISP_LINE1)
    assert (a.N, a.K, a.target_K, a.rnn_feature_size, a.rnn_depth) == (64, 8, 22, 512, 2)
    assert a.onehot is True and a.test_codes is True and a.test_bitwise is True
    # y_input without --use_ynn zeroes the y-MLP sizes (rnn_all.py:251-254)
    assert (a.y_depth, a.y_hidden_size) == (0, 0)
    assert a.g == 91 and a.M == 6 and a.are_we_doing_ML is False and a.tfr_min == 1.0
    assert cli.snr_grid(a) == [-3.0 + i for i in range(7)]


def test_defaults_match_reference():
    a = cli.get_args([])
    assert (a.N, a.K, a.code, a.rate_profile, a.decoding_type) == (32, 12, "PAC", "RM", "y_h0")
    assert a.target_K == 16 and a.g == 53 and a.test_size == 100000 and a.test_batch_size == 10000
    assert (a.test_snr_start, a.test_snr_end, a.snr_points) == (-2.0, 4.0, 7)
    assert a.run_dumer is True and a.tfr_min == 0.0
    assert cli.get_args(["--K", "20", "--N", "32"]).target_K == 20


def test_result_paths_match_reference_known_answers():
    _, final = cli.result_paths(cli.get_args(CRISP_LINE1))
    assert final == CRISP_LINE1_FINAL
    a = cli.get_args("--code Polar --rate_profile polar --target_K 16 --N 32 --K 16 --decoding_type y_input "
                     "--rnn_feature_size 128 --rnn_depth 2 --onehot --batch_size 1024 --tfr_min 1 --tfr_max 1 "
                     "--dec_train_snr 0 --lr 0.001 --id npd_stage16".split())
    results, _ = cli.result_paths(a)
    assert results == ("./Supervised_RNN_Polar_Results/Polar_16_32/Scheme_polar/y_input_onehot/GRU_depth_2_fsize_128/"
                       "y_depth_0_hsize_0/Dec_snr_0.0_bs_1024/tfr_1.0/Activ_selu_Init_He/Optim_AdamW_LR_0.001_loss_MSE/"
                       "npd_stage16")


def _trained_names():
    return sorted(f[:-5] for f in os.listdir(GOLD) if f.startswith("crisp_gru_") and f.endswith(".json"))


def _trained(name=None):
    names = _trained_names()
    if not names:
        pytest.skip("no trained reference checkpoint fixture")
    name = names[0] if name is None else name
    with open(os.path.join(GOLD, name + ".json")) as f:
        return os.path.join(GOLD, name + ".pt"), json.load(f)


def test_reference_checkpoint_loads_with_reference_keys():
    path, meta = _trained()
    ckpt = cli.load_checkpoint(path)
    assert set(ckpt) == {"net", "step", "args"}
    H, N = meta["H"], meta["N"]
    assert tuple(ckpt["net"]["rnn.weight_ih_l0"].shape) == (3 * H, N + 2)
    net, cargs, step = cli.net_from_checkpoint(path)  # strict load_state_dict: keys and shapes as in the reference
    assert (cargs.N, cargs.K, cargs.rnn_feature_size) == (N, meta["K"], H) and step == ckpt["step"]
    assert not net.training


@pytest.mark.gpu
@pytest.mark.parametrize("name", _trained_names())
def test_cli_test_mode_reproduces_reference_curve(capsys, name):
    """`--test` through the drop-in on a reference-trained checkpoint: the GRU and SC BER/BLER curves must agree
    with the curves the live reference printed for the same checkpoint within two-sample Monte-Carlo intervals at a
    family-wise 95 % level (independent noise on both sides; tests/mcstats.py)."""
    path, meta = _trained(name)
    argv = [a for a in meta["final_stage_argv"]]
    for flag in ("--load_path", "--save_path"):
        if flag in argv:
            i = argv.index(flag)
            del argv[i:i + 2]
    argv += ["--test", "--test_load_path", path]
    args = cli.get_args(argv)
    torch.manual_seed(1)
    res = cli.run_test(args)
    printed = capsys.readouterr().out
    assert "BERs of RNN:" in printed and "BERs of SC decoding:" in printed and "Model loaded at step" in printed
    _check_curves(args, res, meta, path)


def _frame_vars(args, path, snr_range, frames=20000):
    """Per-frame bit-error-fraction variances (GRU, SC) per SNR point from a fresh sample through the drop-in."""
    from neural_polar_decoder_b200.rnn_all import RNN_decoder, get_code
    import mcstats
    code = get_code(args.code, args.rate_profile, args.N, args.K, args.g, args=args)
    decoder = RNN_decoder(args.decoding_type, args.N, code.info_inds, args.onehot, args.reverse_order)
    net, _, _ = cli.net_from_checkpoint(path)
    g = torch.Generator().manual_seed(99)
    msg = (1.0 - 2.0 * torch.randint(0, 2, (frames, args.K), generator=g).float()).cuda()
    x = code.encode_plotkin(msg)
    info = torch.as_tensor(np.asarray(code.info_positions), device="cuda")
    out = []
    for si, snr in enumerate(snr_range):
        y = code.channel(x, snr, point=(1 << 30) | si, seed=1234)
        _, dsc = code.sc_decode_new(y, snr, return_llr=False)
        drn = decoder.decode(net, False, y).index_select(1, info)
        out.append((mcstats.frame_fraction_var(msg, drn), mcstats.frame_fraction_var(msg, dsc)))
    return out


def _check_curves(args, res, meta, path):
    import mcstats
    n_ref = n_ours = meta["test_size"]
    assert np.allclose(res["snr_range"], meta["snr_range"])
    ns = len(meta["snr_range"])
    z = mcstats.z_familywise(4 * ns)  # family-wise 95 % over the 4 curves x ns points asserted here
    fv = _frame_vars(args, path, meta["snr_range"])
    print("curve test: %d comparisons, z = %.3f" % (4 * ns, z))
    for i in range(ns):
        for name, which in (("RNN", 0), ("SC", 1)):
            a, b = res["blers_" + name][i], meta["blers_" + name][i]
            hw = mcstats.bler_halfwidth(a, n_ours, b, n_ref, z)
            assert abs(a - b) <= hw, ("bler", name, i, a, b, hw)
            a, b = res["bers_" + name][i], meta["bers_" + name][i]
            hw = mcstats.ber_halfwidth(fv[i][which], n_ours, n_ref, z)
            assert abs(a - b) <= hw + 1e-6, ("ber", name, i, a, b, hw)


# ---- run_models.py (convNet, config 4) ------------------------------------------------------------------------------
RUN_ALT_LAST = ("--model conv --N 64 --max_len 64 --K 22 --dec_train_snr -1 --embed_dim 128 --num_steps 1000 --lr 1e-3 "
                "--batch_size 8192 --mult 1 --num_restarts 1 --print_freq 200 --code polar --previous_K 21 --previous_N 64 "
                "--load_previous --model_iters 1000 --rate_profile polar --curriculum c2n --id c2n --previous_id c2n "
                "--validation_snr 1 --target_K 22 --run 1").split()


def test_run_models_args_and_paths():
    from neural_polar_decoder_b200 import run_models
    a = run_models.get_args(RUN_ALT_LAST)
    assert (a.model, a.N, a.K, a.embed_dim, a.curriculum, a.g, a.are_we_doing_ML) == ("conv", 64, 22, 128, "c2n", 91, False)
    results, final, previous = run_models.result_paths(a)
    # the scheme of run_models.py:575-611 (the ".pt" path is a DIRECTORY in the reference)
    assert results == "./Supervised_Xformer_decoder_Polar_Results/Polar_22_64/Scheme_polar/conv/8_depth_6/c2n/1"
    assert final == "./Supervised_Xformer_decoder_Polar_Results/final_nets/Scheme_polar/N64_K22_conv_8_depth_6.pt/c2n/1"
    assert previous == "./Supervised_Xformer_decoder_Polar_Results/Polar_21_64/Scheme_polar/conv/8_depth_6/c2n/1"
    polar, target, info, tinfo = run_models.build_code(a)
    assert info.tolist() == tinfo.tolist() == [27, 29, 30, 31, 39, 43, 45, 46, 47] + list(range(51, 64))
    b = run_models.get_args("--model conv --N 64 --K 5 --target_K 22 --code polar --rate_profile polar --curriculum l2r".split())
    assert run_models.build_code(b)[2].tolist() == tinfo[:5].tolist()
    with pytest.raises(NotImplementedError):
        run_models.build_model(run_models.get_args(["--model", "gpt", "--code", "polar"]))
    with pytest.raises(SystemExit):
        run_models.main(["--model", "conv", "--code", "polar", "--N", "64", "--K", "22"])


def _conv_trained():
    pt, js = os.path.join(GOLD, "conv_N64_K22_E128.pt"), os.path.join(GOLD, "conv_N64_K22_E128.json")
    if not os.path.exists(pt):
        pytest.skip("no trained reference convNet checkpoint fixture")
    with open(js) as f:
        return pt, json.load(f)


def test_reference_conv_checkpoint_loads_with_reference_keys():
    from neural_polar_decoder_b200 import run_models
    pt, meta = _conv_trained()
    ck = cli.load_checkpoint(pt)
    assert set(ck) == {"xformer", "step", "args"}
    net, cargs, step = run_models.net_from_checkpoint(pt)  # strict load_state_dict
    assert (cargs.model, cargs.N, cargs.embed_dim) == ("conv", 64, 128) and not net.training


@pytest.mark.gpu
def test_run_models_test_mode_synthetic_checkpoint(tmp_path, monkeypatch, capsys):
    """`--model conv --test` end to end on a checkpoint written in the reference's format at the reference's path."""
    import argparse
    from neural_polar_decoder_b200 import run_models, synth
    monkeypatch.chdir(tmp_path)
    argv = [a for a in RUN_ALT_LAST if a not in ("--load_previous",)] + ["--test", "--test_size", "4000", "--test_batch_size",
                                                                         "2000", "--snr_points", "3", "--test_snr_end", "2"]
    args = run_models.get_args(argv)
    args.model_iters = None
    _, final, _ = run_models.result_paths(args)
    os.makedirs(final + "/Models")
    sd = {k: torch.from_numpy(v) for k, v in synth.conv_state_dict(3, 64, 128).items()}
    torch.save({"xformer": sd, "step": 7, "args": argparse.Namespace(**vars(args))}, final + "/Models/model_final.pt")
    res = run_models.run_test(args)
    out = capsys.readouterr().out
    assert "Model loaded at step 7" in out and "BERs of Xformer:" in out and "BLERs of Xformer:" in out
    assert len(res["bers_Xformer"]) == 3 and res["bers_SC"][0] > res["bers_SC"][-1] > 0
    assert all(l <= s + 1e-9 for l, s in zip(res["blers_SCL"], res["blers_SC"]))


@pytest.mark.gpu
def test_run_models_reproduces_reference_conv_curve(capsys):
    """The reference-trained convNet checkpoint through `--model conv --test`: convNet / SC / SC-list(4) curves within
    two-sample Monte-Carlo intervals (family-wise 95 %) of what the live reference's testXformer printed."""
    import mcstats
    from neural_polar_decoder_b200 import run_models
    pt, meta = _conv_trained()
    argv = [a for a in meta["test_argv"]] + ["--test_load_path", pt]
    i = argv.index("--test_size")
    argv[i + 1] = "100000"
    args = run_models.get_args(argv)
    torch.manual_seed(2)
    res = run_models.run_test(args)
    n_ref, n_ours, ns = meta["test_size"], 100000, len(meta["snr_range"])
    z = mcstats.z_familywise(3 * ns)
    for i in range(ns):
        for name in ("Xformer", "SC", "SCL"):
            a, b = res["blers_" + name][i], meta["blers_" + name][i]
            hw = mcstats.bler_halfwidth(a, n_ours, b, n_ref, z)
            assert abs(a - b) <= hw, (name, i, a, b, hw)
