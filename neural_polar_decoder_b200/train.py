"""GPU training step of the CRISP GRU sequential decoder: the body of the reference's training loop,
rnn_all.py:1399-1437, behind libnpd.so's npd_gru_train_step (csrc/gru_train.cu), plus the loop around it
(rnn_all.py:1386-1479: message generation, encoder, channel, teacher-forcing schedule, StepLR, reference-format
checkpoints) so that a run_crisp.sh line without --test trains on the B200.

Covered configuration = what run_crisp.sh trains: rnn_type GRU, decoding_type 'y_input', --onehot, rnn_depth 2,
out_linear_depth 1, no LayerNorm / dropout, loss MSE on the info positions (target 'gt'), AdamW (torch defaults), grad-norm
clipping.  Everything else raises NotImplementedError instead of silently training something different."""
import ctypes
import math
import os
import random
import time

import numpy as np
import torch

from . import _lib

PARAM_KEYS = ("rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0", "rnn.weight_ih_l1",
              "rnn.weight_hh_l1", "rnn.bias_ih_l1", "rnn.bias_hh_l1", "linear.weight", "linear.bias")


def _blob(net):
    sd = net.state_dict()
    missing = [k for k in PARAM_KEYS if k not in sd]
    if missing or len(sd) != len(PARAM_KEYS):
        raise NotImplementedError("GPU training covers the 2-layer GRU + Linear(H,1) head of run_crisp.sh; this RNN_Model has "
                                  "parameters %s" % sorted(sd))
    return np.concatenate([sd[k].detach().cpu().numpy().astype(np.float32).reshape(-1) for k in PARAM_KEYS])


class GRUTrainer:
    """Device-side training state (parameters, gradients, Adam moments, saved activations) of one RNN_Model."""

    def __init__(self, net, N, max_batch, tf32=0):
        """tf32: GEMM arithmetic, 0 = fp32 (parity), 1 = TF32, 2 = bf16, 3 = fp16 tensor cores (include/npd.h)."""
        _lib.require_cuda()
        self.net, self.N, self.H, self.max_batch = net, int(N), int(net.feature_size), int(max_batch)
        if net.input_size != self.N + 2:
            raise NotImplementedError("GPU training covers decoding_type 'y_input' with --onehot (input width N + 2)")
        self.lib = _lib.load()
        blob = _blob(net)
        self.n_params = int(self.lib.npd_gru_trainer_param_count(self.N, self.H))
        assert blob.size == self.n_params, (blob.size, self.n_params)
        h = ctypes.c_void_p()
        _lib.check(self.lib.npd_gru_trainer_create(self.N, self.H, self.max_batch, blob.ctypes.data, int(tf32),
                                                   ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.npd_gru_trainer_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def get(self, what="params"):
        """-> fp32 numpy blob in PARAM_KEYS order: 'params', 'grads' (last step, after clipping), 'exp_avg', 'exp_avg_sq'."""
        out = np.empty(self.n_params, dtype=np.float32)
        _lib.check(self.lib.npd_gru_trainer_get(self.h, ("params", "grads", "exp_avg", "exp_avg_sq").index(what),
                                                out.ctypes.data))
        return out

    def set_params(self, net=None, reset_optimizer=False):
        blob = _blob(self.net if net is None else net)
        _lib.check(self.lib.npd_gru_trainer_set_params(self.h, blob.ctypes.data, int(reset_optimizer)))

    def sync_to_net(self, net=None):
        """Copy the trained parameters back into the nn.Module (for state_dict() / checkpoints / the decode kernel)."""
        net = self.net if net is None else net
        blob, o = self.get("params"), 0
        with torch.no_grad():
            sd = net.state_dict()
            for k in PARAM_KEYS:
                n = sd[k].numel()
                sd[k].copy_(torch.from_numpy(blob[o:o + n].reshape(tuple(sd[k].shape))))
                o += n
        if hasattr(net, "_npd"):
            net._npd = None  # repacked decode weights are stale now
        return net

    def step(self, loss_code, y, gt, teacher_forced, lr, clip=0.25, apply_update=True, want_logits=False, want_loss=True):
        """One training iteration on device tensors y, gt [B,N].  -> (loss, grad norm before clipping, logits or None)."""
        y = _lib.to_device_f32(y)
        gt = _lib.to_device_f32(gt, y.device)
        B = y.shape[0]
        assert y.shape == gt.shape == (B, self.N)
        logits = torch.empty(B, self.N, device=y.device) if want_logits else None
        res = (ctypes.c_float * 2)()
        with torch.cuda.device(y.device):
            _lib.check(self.lib.npd_gru_train_step(self.h, loss_code.h, _lib.ptr(y), _lib.ptr(gt), int(bool(teacher_forced)), B,
                                                   float(lr), float(clip), int(bool(apply_update)),
                                                   ctypes.cast(res, ctypes.c_void_p) if want_loss else None, _lib.ptr(logits),
                                                   _lib.stream_ptr()))
        return (float(res[0]), float(res[1]), logits) if want_loss else (None, None, logits)


def check_trainable(args):
    bad = []
    if args.rnn_type != "GRU" or args.bidirectional:
        bad.append("--rnn_type GRU, unidirectional")
    if args.decoding_type != "y_input" or not args.onehot or args.use_ynn:
        bad.append("--decoding_type y_input --onehot without --use_ynn")
    if args.rnn_depth != 2 or args.out_linear_depth != 1 or args.use_layernorm or args.dropout:
        bad.append("--rnn_depth 2, Linear(H,1) head, no LayerNorm / dropout")
    if args.loss != "MSE" or args.target != "gt" or args.loss_on_all or args.loss_only is not None:
        bad.append("--loss MSE --target gt on the info positions")
    if args.optimizer_type != "AdamW" or args.scheduler == "cosine" or args.mult != 1:
        bad.append("--optimizer_type AdamW, no scheduler or --scheduler step, --mult 1")
    if args.reverse_order:
        bad.append("forward decoding order")
    if bad:
        raise NotImplementedError("GPU training (csrc/gru_train.cu) covers the run_crisp.sh configuration only: " + "; ".join(bad))


def run_train(args, out=print):
    """The reference's training loop (rnn_all.py:1386-1479) with every iteration's forward / backward / update on the GPU.
    Writes the reference's checkpoints ({'net', 'step', 'args'}) to the reference's paths.  Returns the losses."""
    from . import cli
    from .rnn_all import RNN_decoder, get_code
    check_trainable(args)
    _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    code = get_code(args.code, args.rate_profile, args.N, args.K, args.g, args=args)
    decoder = RNN_decoder(args.decoding_type, args.N, code.info_inds, args.onehot, args.reverse_order)
    net = cli.build_net(args)
    if args.load_path is not None:
        net.load_state_dict(cli.load_checkpoint(args.load_path)["net"])
        out("Pretrained model loaded")
    results_path, final_path = cli.result_paths(args)
    os.makedirs(results_path + "/Models", exist_ok=True)
    if os.path.dirname(final_path):
        os.makedirs(os.path.dirname(final_path), exist_ok=True)
    bs = args.batch_size
    trainer = GRUTrainer(net, args.N, bs, tf32=("fp32", "tf32", "bf16", "fp16").index(getattr(args, "train_gemm", "fp32")))
    loss_code = decoder._loss_code(code.info_inds)
    info = torch.as_tensor(np.asarray(code.info_inds), device=dev)
    out("Training ({}, {}). Need to save for: {} \n Save path: {}".format(args.K, args.N, args.model_save_per, results_path))
    range_snr = [args.dec_train_snr, args.dec_train_snr + 1, args.dec_train_snr + 2]
    losses, start = [], time.time()
    for i_step in range(args.num_steps):
        train_snr = range_snr[i_step % 3] if args.do_range_training else args.dec_train_snr
        msg_bits = 1 - 2 * (torch.rand(bs, args.K, device=dev) < 0.5).float()          # rnn_all.py:1400
        gt = torch.ones(bs, args.N, device=dev)
        gt[:, info] = msg_bits
        y = code.channel(code.encode(msg_bits), train_snr)
        if i_step > args.teacher_steps:                                                # 1406
            tfr = args.tfr_min + (args.tfr_max - args.tfr_min) * math.exp(-1 * (i_step - args.teacher_steps) / args.tfr_decay)
        else:
            tfr = args.tfr_max
        teacher = random.random() < tfr                                                # 425
        lr = args.lr if args.scheduler is None else args.lr * args.lr_decay_gamma ** (i_step // args.lr_decay)  # StepLR
        want = i_step % args.print_freq == 0 or i_step + 1 == args.num_steps
        loss, _, _ = trainer.step(loss_code, y, gt, teacher, lr, args.clip, want_loss=want)
        if want:
            losses.append((i_step, loss))
            out("[%d/%d] At %d dB, Loss: %.7f" % (i_step, args.num_steps, train_snr, loss))
        if i_step == 10:
            torch.cuda.synchronize()
            out("Time for one step is {0:.4f} minutes".format((time.time() - start) / 11 / 60))
        if (i_step + 1) % args.model_save_per == 0 or i_step + 1 == args.num_steps:  # 1471-1479
            trainer.sync_to_net()
            ck = {"net": net.state_dict(), "step": i_step + 1, "args": args}
            torch.save(ck, results_path + "/Models/model_{0}.pt".format(i_step + 1))
            torch.save(ck, results_path + "/Models/model_final.pt")
            torch.save(ck, final_path)
    out("Complete")
    return losses
