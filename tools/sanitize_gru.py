"""Tiny GRU / SCL / SC decodes for compute-sanitizer (memcheck / racecheck):  compute-sanitizer --tool memcheck python tools/sanitize_gru.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_polar_decoder_b200 import rnn_all, synth
N, K, H, B = 32, 16, int(os.environ.get("H", "256")), int(os.environ.get("B", "130"))
code = rnn_all.get_code('Polar', 'polar', N, K)
net = rnn_all.RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.gru_state_dict(5, N, H, 2, head_gain=6.0).items()})
dec = rnn_all.RNN_decoder('y_input', N, code.info_inds, onehot=True)
y = torch.randn(B, N, generator=torch.Generator().manual_seed(1)).cuda()
d, lg = dec.decode(net, False, y, return_logits=True)
torch.cuda.synchronize()
print("gru ok", float(lg.abs().mean()))
_, s = code.scl_decode(y, 1.0, 4)
_, c = code.sc_decode_new(y, 1.0)
torch.cuda.synchronize()
print("sc/scl ok", float((s == c).float().mean()))
