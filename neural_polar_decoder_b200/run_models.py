"""Command-line and checkpoint compatibility for the evaluation half of the reference's run_models.py
(SURVEY.md 8f-1, config 4): `python -m neural_polar_decoder_b200.run_models --model conv <the run_alt.sh /
run_conv_c2n.sh flags> --test` loads a reference-format checkpoint ({'xformer', 'step', 'args'}, run_models.py:980-983)
and prints what the reference's TESTING block prints (run_models.py:1104-1380), every decode on the B200 path.

Mirrored: the argument names / defaults / derived fields of get_args (run_models.py:45-218; training-only flags are
accepted and ignored), the results / final-net path scheme (571-611; the reference uses `<...>.pt/<id>/<run>` as a
DIRECTORY holding Models/model_final.pt), the code construction and curriculum index selection (617-706), checkpoint
selection (1310-1318), the SNR grid (1327-1331), the test set and loader (1333-1338) and the printed lines (1371-1379).
Not mirrored: training (raises unless --test), plotting, --plot_progressive, the PAC / Fano test branch (1419-1505),
models other than `conv` (transformers / rnnAttn are out of the hot path: SURVEY.md 2), the ML / bitwise-MAP decoders
(are_we_doing_ML is only true for K <= 16 and N <= 32; their lists print as zeros)."""
import argparse
import random
import time

import numpy as np
import torch

from .cli import load_checkpoint, str2bool

# (flag, kind, default[, choices]); kind: a type, 'flag' (store_true) or 'bool' (str2bool, nargs='?', const=True)
_ARGS = [
    ("id", str, None), ("previous_id", str, None), ("code", str, "pac", ["pac", "polar"]),
    ("previous_code", str, None, [None, "pac", "polar"]), ("N", int, 32), ("previous_N", int, 32), ("max_len", int, 32),
    ("K", int, 8), ("previous_K", int, 8), ("test", "flag", False), ("plot_progressive", "flag", False),
    ("do_range_training", "flag", False), ("rate_profile", str, "RM", ["RM", "polar", "sorted", "last", "custom"]),
    ("previous_rate_profile", str, None, [None, "RM", "polar", "sorted", "last", "custom"]), ("embed_dim", int, 64),
    ("dropout", int, 0.1), ("n_head", int, 8), ("n_layers", int, 6), ("num_devices", int, 2), ("load_previous", "flag", False),
    ("parallel", "flag", False), ("dont_use_bias", "flag", False), ("include_previous_block_errors", "flag", False),
    ("dec_train_snr", float, -1.), ("test_snr_start", float, -2.), ("test_snr_end", float, 4.), ("model_iters", int, None),
    ("run", int, None), ("num_steps", int, 400000), ("batch_size", int, 128), ("mult", int, 1), ("lr", float, 1e-3),
    ("cosine", "flag", False), ("num_restarts", int, 1), ("print_freq", int, 1000),
    ("activation", str, "selu", ["selu", "relu", "elu", "tanh", "sigmoid"]),
    ("curriculum", str, "c2n", ["c2n", "n2c", "r2l", "l2r", "random"]), ("target_K", int, 16),
    ("model", str, "gpt", ["simple", "conv", "encoder", "decoder", "gpt", "denoiser", "bigConv", "small", "multConv",
                           "rnnAttn", "bitConv"]),
    ("initialization", str, "Xavier", ["Dontknow", "He", "Xavier"]),
    ("optimizer_type", str, "AdamW", ["Adam", "RMS", "AdamW", "SGD"]), ("loss", str, "MSE", ["Huber", "MSE", "NLL", "Block"]),
    ("loss_on_all", "flag", False), ("split_batch", "flag", False), ("lr_decay", int, None), ("T_anneal", int, None),
    ("lr_decay_gamma", float, None), ("clip", float, 0.25), ("validation_snr", float, None), ("no_detach", "flag", False),
    ("tfr_min", float, None), ("tfr_max", float, 0.), ("tfr_decay", float, 10000), ("teacher_steps", int, -10000),
    ("model_save_per", int, 5000), ("snr_points", int, 7), ("test_batch_size", int, 1000), ("test_size", int, 50000),
    ("test_load_path", str, None), ("run_fano", "flag", False), ("random_test", "flag", False), ("save_path", str, None),
    ("load_path", str, None), ("run_dumer", "bool", True), ("hard_decision", "flag", False), ("gpu", int, -1),
    ("anomaly", "flag", False), ("only_args", "flag", False),
]
_PAC_G = {4: 7, 8: 13, 16: 21, 32: 53}


def get_args(argv=None):
    """reference run_models.py:45-218: same flags and defaults, same derived fields."""
    p = argparse.ArgumentParser(description="one-shot decoder (convNet) evaluation on the B200 path")
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
    # extension: the reference hard-wires the CPU list decoder (L = 4) into testXformer whenever ML is off
    p.add_argument("--no_scl", action="store_true", help="skip the SC-list (L=4) curve of testXformer")
    args = p.parse_args(argv)
    args.g = _PAC_G.get(args.N, 91)
    args.are_we_doing_ML = bool(args.K <= 16 and args.N <= 32)
    return args


def result_paths(args):
    """(results_save_path, final_save_path, previous_save_path) as run_models.py:571-611 composes them."""
    if args.previous_code is None:
        args.previous_code = args.code
    if args.previous_rate_profile is None:
        args.previous_rate_profile = args.rate_profile
    ID = "" if args.id is None else args.id
    root = {"polar": "./Supervised_Xformer_decoder_Polar_Results", "pac": "./Supervised_Xformer_decoder_PAC_Results"}
    tail = "{}/{}_depth_{}".format(args.model, args.n_head, args.n_layers)
    results = "{}/Polar_{}_{}/Scheme_{}/{}".format(root[args.code], args.K, args.N, args.rate_profile, tail)
    if args.save_path is None:
        final = "{}/final_nets/Scheme_{}/N{}_K{}_{}_{}_depth_{}.pt".format(root[args.code], args.rate_profile, args.N, args.K,
                                                                            args.model, args.n_head, args.n_layers)
    else:
        final = args.save_path
    if ID != "":
        results, final = results + "/" + ID, final + "/" + ID
    previous = "{}/Polar_{}_{}/Scheme_{}/{}".format(root[args.previous_code], args.previous_K, args.previous_N,
                                                    args.previous_rate_profile, tail)
    if args.previous_id is not None:
        previous = previous + "/" + args.previous_id
    if args.run is not None:
        results, final, previous = (p + "/{0}".format(args.run) for p in (results, final, previous))
    return results, final, previous


def build_code(args):
    """(polar, polarTarget, info_inds, target_info_inds) -- run_models.py:617-706 for --code polar."""
    from . import PolarCode, construct
    if args.code != "polar":
        raise NotImplementedError("run_models --code pac (run_models.py:1419-1505) is not on the B200 path")
    n = int(np.log2(args.N))
    if args.rate_profile == "polar":
        rs = construct.reference_rs256()
        polar = PolarCode(n, args.K, args, rs=rs)
        target = PolarCode(n, args.target_K, args, rs=rs)
    elif args.rate_profile == "RM":
        w = np.array([construct.count_set_bits(i) for i in range(args.N)])
        polar = PolarCode(n, args.K, args, F=np.sort(np.argsort(w)[:-args.K]))
        target = PolarCode(n, args.target_K, args, F=np.sort(np.argsort(w)[:-args.target_K]))
    else:
        raise NotImplementedError("rate_profile %r is not constructed by the reference's run_models.py either" % args.rate_profile)
    cur = args.curriculum
    if cur == "c2n":
        info = polar.info_positions
    elif cur == "n2c":
        info = target.unsorted_info_positions[:args.K].copy()
    elif cur == "l2r":
        info = target.info_positions[:args.K].copy()
    elif cur == "r2l":
        info = target.info_positions[-args.K:].copy()
    else:  # 'random' (run_models.py:689-694)
        perm = target.info_positions.copy()
        random.Random(42).shuffle(perm)
        info = perm[:args.K].copy()
    return polar, target, np.sort(info), np.sort(target.info_positions)


def build_model(args):
    from .models import convNet
    if args.model != "conv":
        raise NotImplementedError("--model %s: the B200 path covers the convNet alternate model (`--model conv`)" % args.model)
    return convNet(args)


def snr_grid(args):
    """run_models.py:1327-1331."""
    if args.snr_points == 1 and args.test_snr_start == args.test_snr_end:
        return [args.test_snr_start]
    step = (args.test_snr_end - args.test_snr_start) * 1.0 / (args.snr_points - 1)
    return [step * i + args.test_snr_start for i in range(args.snr_points)]


def net_from_checkpoint(path_or_ckpt):
    """-> (convNet in eval mode, checkpoint args, step) rebuilt from the hyper-parameters stored in the checkpoint."""
    ckpt = load_checkpoint(path_or_ckpt) if isinstance(path_or_ckpt, str) else path_or_ckpt
    net = build_model(ckpt["args"])
    net.load_state_dict(ckpt["xformer"])
    net.eval()
    return net, ckpt["args"], ckpt.get("step")


def run_test(args, out=print):
    """The reference's TESTING block (run_models.py:1104, 1307-1380) without plotting.  Returns a dict of the curves."""
    from . import sweep
    polar, _, info_inds, target_info = build_code(args)
    out("Info positions : {}".format(info_inds))
    out("Target Info positions : {}".format(target_info))
    out("Code : {0} ".format(args.code))
    out("Type of training : {0}".format(args.curriculum))
    out("Rate Profile : {0}".format(args.rate_profile))
    xformer = build_model(args)
    results_path, final_path, _ = result_paths(args)
    out("TESTING :")
    if args.plot_progressive:
        raise NotImplementedError("--plot_progressive (run_models.py:1107-1305) is plotting only")
    if args.model_iters is not None:
        path = results_path + "/Models/model_{0}.pt".format(args.model_iters)
    elif args.test_load_path is not None:
        path = args.test_load_path
    else:
        path = final_path + "/Models/model_final.pt"
    ckpt = load_checkpoint(path)
    xformer.load_state_dict(ckpt["xformer"])
    xformer.eval()
    out("Model loaded at step {}".format(ckpt["step"]))
    device = torch.device("cuda", torch.cuda.current_device())
    snr_range = snr_grid(args)
    test_msg_bits = 2 * (torch.rand(args.test_size, args.K) < 0.5).float() - 1
    loader = torch.utils.data.DataLoader(test_msg_bits, batch_size=args.test_batch_size, shuffle=False)
    start = time.time()
    r = sweep.testXformer(xformer, polar, snr_range, loader, device, run_ML=False, run_SCL=not args.no_scl)
    keys = ("bers_Xformer", "blers_Xformer", "bers_SC", "blers_SC", "bers_SCL", "blers_SCL", "bers_ML", "blers_ML",
            "bers_bitwise_Xformer", "bers_bitwise_MAP", "blers_bitwise_MAP")
    res = dict(zip(keys, r))
    res.update(snr_range=snr_range, step=ckpt["step"])
    torch.cuda.synchronize()
    out("Test SNRs : ", snr_range)
    out("BERs of Xformer: {0}".format(res["bers_Xformer"]))
    out("BERs of SC decoding: {0}".format(res["bers_SC"]))
    out("BERs of ML: {0}".format(res["bers_ML"]))
    out("BLERs of ML: {0}".format(res["blers_ML"]))
    out("BERs of bitML: {0}".format(res["bers_bitwise_MAP"]))
    out("BLERs of bitML: {0}".format(res["blers_bitwise_MAP"]))
    out("BLERs of Xformer: {0}".format(res["blers_Xformer"]))
    out("Time taken = {} seconds".format(time.time() - start))
    return res


def main(argv=None):
    args = get_args(argv)
    if args.only_args:
        print("Loaded args. Exiting")
        return 0
    if not args.test:
        raise SystemExit("neural_polar_decoder_b200 accelerates the evaluation path only: pass --test "
                         "(training stays with the reference's run_models.py; its checkpoints load here unchanged)")
    if args.gpu >= 0:
        torch.cuda.set_device(args.gpu)
    run_test(args)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
