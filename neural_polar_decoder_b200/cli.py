"""Command-line and checkpoint compatibility for the evaluation half of the reference's rnn_all.py
(SURVEY.md 8f-1): `python -m neural_polar_decoder_b200.rnn_all <the run_crisp.sh flags> --test` loads a
reference-format checkpoint and prints the BER/BLER sweep the reference's TESTING block prints
(rnn_all.py:1745-1905), with every decode on the B200 path.

What is mirrored
  * the argument names, defaults and derived fields of get_args (rnn_all.py:48-255): training-only flags are
    accepted (a run_crisp.sh line can be pasted unchanged) and ignored;
  * the results / final-net path scheme (rnn_all.py:1216-1248), so `--test` finds `Models/model_final.pt` where a
    reference training run left it; `--test_load_path` / `--model_iters` select a file as in 1749-1754;
  * checkpoint format {'net': state_dict, 'step': int, 'args': Namespace} (rnn_all.py:1471-1479);
  * the printed lines "Test SNRs :", "BERs of RNN:", "BERs of SC decoding:", "BERs of ML:", "Time taken".
  * without --test: the training loop (rnn_all.py:1386-1479) on the GPU for the run_crisp.sh configuration
    (train.py / csrc/gru_train.cu), writing the reference's checkpoints to the reference's paths.
Not mirrored: plotting, validation prints inside the training loop, Fano / ML / RNN-list decoders (their lists print as
zeros).
"""
import argparse
import math
import os
import time

import numpy as np
import torch


def str2bool(v):
    if isinstance(v, bool):
        return v
    s = str(v).lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


# (flag, kind, default[, choices]); kind: a type, 'flag' (store_true) or 'bool' (str2bool, nargs='?', const=True)
_ARGS = [
    ("id", str, None), ("N", int, 32), ("K", int, 12), ("target_K", int, None), ("test", "flag", False),
    ("code", str, "PAC", ["PAC", "Polar"]),
    ("rate_profile", str, "RM", ["RM", "rev_RM", "polar", "sorted", "sorted_last", "rev_polar", "custom", "random"]),
    ("random_seed", int, 42), ("info_ind", int, 63), ("rnn_type", str, "GRU", ["GRU", "LSTM"]),
    ("bidirectional", "bool", False), ("decoding_type", str, "y_h0", ["y_h0", "y_input", "y_h0_out"]),
    ("target", str, "gt", ["gt", "llr"]), ("onehot", "bool", False), ("mult", int, 1), ("print_freq", int, 100),
    ("rnn_feature_size", int, 256), ("rnn_pool_type", str, "last", ["last", "average"]), ("rnn_depth", int, 2),
    ("y_depth", int, 3), ("y_hidden_size", int, 128), ("out_linear_depth", int, 1), ("dropout", float, 0.),
    ("use_skip", "bool", False), ("use_layernorm", "bool", False), ("weight0", float, None),
    ("test_codes", "bool", False), ("test_bitwise", "bool", False), ("num_steps", int, 200000),
    ("batch_size", int, 4096), ("activation", str, "selu", ["selu", "relu", "elu", "tanh", "sigmoid"]),
    ("initialization", str, "He", ["Dontknow", "He", "Xavier"]),
    ("optimizer_type", str, "AdamW", ["Adam", "RMS", "AdamW"]), ("scheduler", str, None, ["cosine", "step"]),
    ("loss", str, "MSE", ["Huber", "MSE", "BCE"]), ("loss_on_all", "flag", False), ("loss_only", int, None),
    ("split_batch", "flag", False), ("lr", float, 0.001), ("lr_decay", int, None), ("lr_decay_gamma", float, 0.1),
    ("clip", float, 0.25), ("no_detach", "flag", False), ("tfr_min", float, None), ("tfr_max", float, 0.),
    ("tfr_decay", float, 10000), ("teacher_steps", int, -10000), ("dec_train_snr", float, -1.),
    ("validation_snr", float, None), ("testing_snr", float, None), ("do_range_training", "bool", False),
    ("model_save_per", int, 10000), ("test_snr_start", float, -2.), ("test_snr_end", float, 4.),
    ("snr_points", int, 7), ("test_batch_size", int, 10000), ("test_size", int, 100000),
    ("noise_type", str, "awgn", ["awgn", "fading", "radar", "t-dist"]), ("vv", float, 5), ("radar_prob", float, 0.05),
    ("radar_power", float, 5.0), ("model_iters", int, None), ("test_load_path", str, None), ("list_size", int, None),
    ("run_fano", "flag", False), ("random_test", "flag", False), ("save_path", str, None),
    ("progressive_path", str, None), ("load_path", str, None), ("run_dumer", "bool", True), ("run_ML", "bool", False),
    ("hard_decision", "flag", False), ("gpu", int, -2), ("anomaly", "flag", False), ("only_args", "flag", False),
    ("use_ynn", "flag", False), ("reverse_order", "flag", False), ("print_cust", "flag", False), ("fresh", "flag", False),
    # not a reference flag: arithmetic of the training step's GEMMs (npd_gru_trainer_create; fp32 = the reference's)
    ("train_gemm", str, "fp32", ["fp32", "tf32", "bf16", "fp16"]),
]

# default PAC generator polynomial per code length (rnn_all.py:217-235)
_PAC_G = {4: 7, 8: 13, 16: 21, 32: 53}


def get_args(argv=None):
    """reference rnn_all.py:48-255: same flags and defaults, same derived fields."""
    p = argparse.ArgumentParser(description="CRISP sequential decoder evaluation on the B200 path")
    for spec in _ARGS:
        name, kind, default = spec[0], spec[1], spec[2]
        kw = {"default": default}
        if kind == "flag":
            kw["action"] = "store_true"
        elif kind == "bool":
            kw.update(type=str2bool, nargs="?", const=True)
        else:
            kw["type"] = kind
            if len(spec) > 3:
                kw["choices"] = spec[3]
        p.add_argument("--" + name, **kw)
    args = p.parse_args(argv)
    if args.target_K is None:
        args.target_K = args.N // 2 if args.K <= args.N // 2 else args.K
    args.g = _PAC_G.get(args.N, 91)
    args.M = int(math.log(args.N, 2))
    args.are_we_doing_ML = bool(args.K <= 0 or args.run_ML)
    if args.tfr_min is None:
        args.tfr_min = args.tfr_max
    if args.decoding_type == "y_input" and not args.use_ynn:
        args.y_depth = 0
        if not args.out_linear_depth > 1:
            args.y_hidden_size = 0
    return args


def result_paths(args):
    """(results_save_path, final_save_path) exactly as the reference composes them (rnn_all.py:1216-1248)."""
    ID = "" if args.id is None else args.id
    lr_ = args.lr if args.scheduler is None else str(args.lr) + "_decay_{}_{}".format(args.lr_decay, args.scheduler)
    if args.tfr_min != args.tfr_max:
        tfr_ = "tfr_min_{}_max_{}_decay_{}_init_{}".format(args.tfr_max, args.tfr_min, args.tfr_decay, args.teacher_steps)
    else:
        tfr_ = "tfr_{}".format(args.tfr_min)
    g_ = "_g_{}".format(args.g) if args.code == "PAC" else ""
    lo_ = "" if args.loss_only is None else "_lo{}".format(args.loss_only)
    y_ = "y_depth_{}_hsize_{}{}{}".format(args.y_depth, args.y_hidden_size,
                                          "_out" + str(args.out_linear_depth) if args.out_linear_depth > 1 else "",
                                          "_LN" if args.use_layernorm else "")
    rnn_type = args.rnn_type if not args.bidirectional else "Bi-" + args.rnn_type
    profile = args.rate_profile
    if args.rate_profile == "random" and args.random_seed != 42:
        profile = "random{}".format(args.random_seed)
    if args.target_K != args.N // 2:
        profile = "{}_{}".format(profile, args.target_K)
    if args.use_skip:
        y_ = y_ + "_skip"
    dec = args.decoding_type if not args.onehot else args.decoding_type + "_onehot"
    bs = args.batch_size * args.mult
    root = "./Supervised_RNN_{}_Results".format(args.code)
    results = (root + "/{code}_{K}_{N}{lo}/Scheme_{prof}{g}/{dec}/{rnn}_depth_{depth}_fsize_{fs}/{y}/Dec_snr_{snr}_bs_{bs}/"
               "{tfr}/Activ_{act}_Init_{init}/Optim_{opt}_LR_{lr}_loss_{loss}/{ID}").format(
        code=args.code, K=args.K, N=args.N, lo=lo_, prof=profile, g=g_, dec=dec, rnn=rnn_type, depth=args.rnn_depth,
        fs=args.rnn_feature_size, y=y_, snr=args.dec_train_snr, bs=bs, tfr=tfr_, act=args.activation,
        init=args.initialization, opt=args.optimizer_type, lr=lr_, loss=args.loss, ID=ID)
    if args.save_path is not None:
        final = args.save_path
    else:
        final = (root + "/final_nets/Scheme_{prof}/N{N}_K{K}{lo}{g}_{dec}_{rnn}_depth_{depth}_fsize_{fs}_{y}_snr_{snr}_bs_{bs}_"
                 "{tfr}_activ_{act}_init_{init}_optim_{opt}_lr_{lr}_loss_{loss}_{ID}.pt").format(
            prof=profile, N=args.N, K=args.K, lo=lo_, g=g_, dec=dec, rnn=rnn_type, depth=args.rnn_depth,
            fs=args.rnn_feature_size, y=y_, snr=args.dec_train_snr, bs=bs, tfr=tfr_, act=args.activation,
            init=args.initialization, opt=args.optimizer_type, lr=lr_, loss=args.loss, ID=ID)
    return results, final


def load_checkpoint(path):
    """A reference checkpoint pickles an argparse.Namespace next to the state_dict (rnn_all.py:1471-1479), which
    torch >= 2.6 refuses under its weights_only default; load it the way the reference's torch 1.12 did."""
    try:
        return torch.load(path, map_location="cpu", weights_only=False)
    except TypeError:  # torch < 1.13 has no weights_only argument
        return torch.load(path, map_location="cpu")


def build_net(args):
    """The RNN_Model the reference builds for `args` (rnn_all.py:1313-1324)."""
    from .rnn_all import RNN_Model
    onehot = int(args.onehot)
    if args.decoding_type == "y_input":
        if args.use_ynn:
            return RNN_Model(args.rnn_type, args.N + 1 + onehot, args.rnn_feature_size, 1, args.rnn_depth, args.N,
                             args.y_hidden_size, args.y_depth, args.activation, args.dropout, args.use_skip,
                             y_output_size=args.N, out_linear_depth=args.out_linear_depth,
                             bidirectional=args.bidirectional, use_layernorm=args.use_layernorm)
        return RNN_Model(args.rnn_type, args.N + 1 + onehot, args.rnn_feature_size, 1, args.rnn_depth, args.N,
                         args.y_hidden_size if args.out_linear_depth > 1 else 0, 0, args.activation, args.dropout,
                         args.use_skip, out_linear_depth=args.out_linear_depth, bidirectional=args.bidirectional,
                         use_layernorm=args.use_layernorm)
    if args.decoding_type == "y_h0":  # the reference's default: initial state from the y-MLP (rnn_all.py:1316-1317)
        return RNN_Model(args.rnn_type, 1 + onehot, args.rnn_feature_size, 1, args.rnn_depth, args.N,
                         args.y_hidden_size, args.y_depth, args.activation, args.dropout, args.use_skip,
                         bidirectional=args.bidirectional, use_layernorm=args.use_layernorm)
    # 'y_h0_out' (rnn_all.py:1323-1324) cannot run in the reference: its head is built for H inputs and fed [Fy | out]
    raise NotImplementedError("decoding_type %r: the B200 path covers 'y_input' and 'y_h0'" % args.decoding_type)


def net_from_checkpoint(path_or_ckpt):
    """-> (net in eval mode, checkpoint args Namespace, step): rebuilds the model from the hyper-parameters the
    reference stored inside the checkpoint."""
    ckpt = load_checkpoint(path_or_ckpt) if isinstance(path_or_ckpt, (str, os.PathLike)) else path_or_ckpt
    cargs = ckpt["args"]
    net = build_net(cargs)
    net.load_state_dict(ckpt["net"])
    net.eval()
    return net, cargs, ckpt.get("step")


def snr_grid(args):
    """reference rnn_all.py:1765-1769."""
    if args.snr_points == 1 and args.test_snr_start == args.test_snr_end:
        return [args.test_snr_start]
    step = (args.test_snr_end - args.test_snr_start) * 1.0 / (args.snr_points - 1)
    return [step * i + args.test_snr_start for i in range(args.snr_points)]


def run_test(args, out=print):
    """The reference's TESTING block (rnn_all.py:1745-1905) without plotting.  Returns a dict of the curves."""
    from . import sweep
    from .rnn_all import RNN_decoder, get_code
    code = get_code(args.code, args.rate_profile, args.N, args.K, args.g, args=args)
    decoder = RNN_decoder(args.decoding_type, args.N, code.info_inds, args.onehot, args.reverse_order)
    net = build_net(args)
    results_path, _ = result_paths(args)
    out("TESTING :")
    if args.model_iters is not None:
        path = results_path + "/Models/model_{0}.pt".format(args.model_iters)
    elif args.test_load_path is not None:
        path = args.test_load_path
    else:
        path = results_path + "/Models/model_final.pt"
    ckpt = load_checkpoint(path)
    net.load_state_dict(ckpt["net"])
    net.eval()
    out("Model loaded at step {}".format(ckpt["step"]))

    snr_range = snr_grid(args)
    test_msg_bits = 2 * (torch.rand(args.test_size, args.K) < 0.5).float() - 1
    loader = torch.utils.data.DataLoader(test_msg_bits, batch_size=args.test_batch_size, shuffle=False)
    start = time.time()
    res = {"snr_range": snr_range, "step": ckpt["step"]}
    if args.code == "PAC":
        out("Testing on random data")
        r = sweep.test_full_data(net, code, snr_range, loader, run_fano=False, run_dumer=args.run_dumer, decoder=decoder)
        keys = ("bers_RNN", "blers_RNN", "bers_SC", "blers_SC", "bers_ML", "blers_ML", "bers_fano", "blers_fano")
        res.update(dict(zip(keys, r)))
        out("Test SNRs : ", snr_range)
        out("BERs of RNN: {0}".format(res["bers_RNN"]))
        out("BERs of SC decoding: {0}".format(res["bers_SC"]))
        out("BERs of ML: {0}".format(res["bers_ML"]))
        out("BERs of Fano: {0}".format(res["bers_fano"]))
    else:
        r = sweep.polar_RNN_full_test(net, code, snr_range, loader, args.are_we_doing_ML, args.list_size is not None,
                                      False, decoder=decoder, list_size=args.list_size or 4)
        keys = ("bers_RNN", "blers_RNN", "bers_SC", "blers_SC", "bers_SCL", "blers_SCL", "bers_RNNL", "blers_RNNL",
                "bers_ML", "blers_ML")
        res.update(dict(zip(keys, r)))
        out("Test SNRs : ", snr_range)
        out("BERs of RNN: {0}".format(res["bers_RNN"]))
        out("BERs of SC decoding: {0}".format(res["bers_SC"]))
        if args.list_size is not None:
            out("BERs of SCL decoding, L={1}: {0}".format(res["bers_SCL"], args.list_size))
            out("BERs of RNNL decoding, L={1}: {0}".format(res["bers_RNNL"], args.list_size))
        out("BERs of ML: {0}".format(res["bers_ML"]))
    torch.cuda.synchronize()
    out("Time taken = {} seconds".format(time.time() - start))
    out("BLERs of RNN: {0}".format(res["blers_RNN"]))
    out("BLERs of SC decoding: {0}".format(res["blers_SC"]))
    return res


def main(argv=None):
    args = get_args(argv)
    if args.only_args:
        print("Loaded args. Exiting")
        return 0
    if args.gpu >= 0:
        torch.cuda.set_device(args.gpu)
    if not args.test:
        # the reference trains and then falls through into its TESTING block (rnn_all.py:1386-1479, 1745-1905)
        from .train import run_train
        run_train(args)
    run_test(args)
    return 0
