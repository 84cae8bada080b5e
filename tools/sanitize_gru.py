"""Tiny invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_gru.py
GRU pair + single-CTA kernels, SC lane / quad / group kernels, SC-list, fused sweep (with flagged re-decodes), encoder,
counters, convNet stack + FC, and the GRU training step."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from neural_polar_decoder_b200 import PolarCode, construct, rnn_all, sweep, synth
from neural_polar_decoder_b200.models import convNet
from neural_polar_decoder_b200.train import GRUTrainer

N, K, B = 32, 16, int(os.environ.get("B", "130"))
code = rnn_all.get_code('Polar', 'polar', N, K)
y = torch.randn(B, N, generator=torch.Generator().manual_seed(1)).cuda()
for H in (256, 128):  # CTA-pair kernel, single-CTA kernel
    net = rnn_all.RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.gru_state_dict(5, N, H, 2, head_gain=6.0).items()})
    dec = rnn_all.RNN_decoder('y_input', N, code.info_inds, onehot=True)
    d, lg = dec.decode(net, False, y, return_logits=True)
    torch.cuda.synchronize()
    print("gru H=%d ok" % H, float(lg.abs().mean()))
_, s = code.scl_decode(y, 1.0, 4)
_, c = code.sc_decode_new(y, 1.0)
torch.cuda.synchronize()
print("sc lane / scl ok", float((s == c).float().mean()))
# quad kernel + exact re-decode of flagged rows (ties), fused sweep with a small frozen prior (many flagged codewords)
big = PolarCode(8, 128, None, F=construct.pw_frozen_set(256, 128))
yb = torch.randn(100, 256, generator=torch.Generator().manual_seed(2)).cuda()
yb[:3] = torch.round(yb[:3] * 2) / 2
_, dq = big.sc_decode_new(yb, 1.0, return_llr=False)
small = PolarCode(8, 128, None, F=construct.pw_frozen_set(256, 128), infty=3.0)
print("sc quad ok", float(dq.abs().mean()), "sweep", sweep.mc_sc_sweep(small, [1.0], 700, chunk=256, seed=3, rank=0, world=1)[:3])
tr = GRUTrainer(net, N, 64)
gt = torch.ones(64, N).cuda()
print("train step ok", tr.step(dec._loss_code(code.info_inds), y[:64], gt, True, 1e-3, 0.25)[:2])
print("train step (student) ok", tr.step(dec._loss_code(code.info_inds), y[:64], gt, False, 1e-3, 0.25)[:2])
cnet = convNet(argparse.Namespace(embed_dim=128, max_len=64, N=64, dont_use_bias=False, dropout=0.0))
cnet.load_state_dict({k: torch.from_numpy(v) for k, v in synth.conv_state_dict(4, 64, 128).items()})
cnet.eval()
lgc = cnet.logits(torch.randn(20, 64, generator=torch.Generator().manual_seed(4)).cuda())
torch.cuda.synchronize()
print("conv ok", float(lgc.abs().mean()))
