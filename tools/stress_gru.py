"""Stress the GRU kernels: many ragged batch sizes, repeated launches, concurrent streams; every result is compared
with the single-CTA kernel's (NPD_GRU_PAIR=0 in a subprocess would be cleaner; here decisions must be deterministic
and identical across repeats, and logits must match the fp32 oracle within tolerance)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import oracle
from neural_polar_decoder_b200 import rnn_all, synth, construct
N, K, H = 64, 22, 512
sd = synth.gru_state_dict(7, N, H, 2, head_gain=8.0)
net = rnn_all.RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0)
net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
rs = construct.reference_rs256(); info = np.sort(rs[rs < N][:K])
dec = rnn_all.RNN_decoder('y_input', N, info, onehot=True)
g = torch.Generator().manual_seed(0)
bad = 0
for B in [1, 63, 64, 65, 127, 128, 129, 200, 9471, 9472, 9473, 20000]:
    y = torch.randn(B, N, generator=g).cuda()
    ref = None
    for rep in range(4):
        d, lg = dec.decode(net, False, y, return_logits=True)
        torch.cuda.synchronize()
        if ref is None:
            ref = (d.clone(), lg.clone())
        elif not (torch.equal(d, ref[0]) and torch.equal(lg, ref[1])):
            bad += 1
            print("NON-DETERMINISTIC at B=%d rep %d: max logit diff %.3e" % (B, rep, float((lg - ref[1]).abs().max())))
    if B <= 200:
        do, lo = oracle.gru_decode(sd, y.cpu().numpy(), N, info)
        d2, lg2 = rnn_all.gru_decode(net, dec._loss_code(info), y, forced=torch.from_numpy(do).cuda(), want_logits=True)
        err = np.abs(lg2.cpu().numpy() - lo); tol = 1e-2 * (np.abs(lo) + np.sqrt((lo ** 2).mean()))
        if not (err <= tol).all():
            bad += 1
            print("TOLERANCE at B=%d: %.3e" % (B, err.max()))
    print("B=%d ok" % B, flush=True)
# concurrent streams
ys = [torch.randn(9472, N, generator=g).cuda() for _ in range(3)]
refs = [dec.decode(net, False, y) for y in ys]
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(3)]
for it in range(5):
    outs = []
    for s, y in zip(streams, ys):
        with torch.cuda.stream(s):
            outs.append(dec.decode(net, False, y))
    torch.cuda.synchronize()
    for o, r in zip(outs, refs):
        if not torch.equal(o, r):
            bad += 1
            print("CONCURRENT mismatch iteration", it)
print("stress done, failures:", bad)
sys.exit(1 if bad else 0)
