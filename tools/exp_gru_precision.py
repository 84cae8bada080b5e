"""Round-2 experiment: precision variants of the GRU pair kernel (libnpd_v*.so built with -DNPD_GRU_ACT / GATEFIX / LO) on the
reference-trained Polar(64,22) checkpoint: forced-feedback logit error against the live reference's fixture and the decode
time of 37888 codewords.  usage: python tools/exp_gru_precision.py v0 v1 ..."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, numpy as np, torch
sys.path.insert(0, %r)
from neural_polar_decoder_b200 import _lib
_lib.LIB_PATH = os.path.join(%r, "neural_polar_decoder_b200", "libnpd%%s.so" %% ("" if sys.argv[1] == "default" else "_" + sys.argv[1]))
CK = os.environ.get("CK", "crisp_gru_N64_K22_H512"); FX = os.environ.get("FX", "gru_trained")
from neural_polar_decoder_b200 import cli
from neural_polar_decoder_b200.rnn_all import RNN_decoder, gru_decode
g = np.load(os.path.join(%r, "tests/golden/" + FX + ".npz"))
net, cargs, _ = cli.net_from_checkpoint(os.path.join(%r, "tests/golden/" + CK + ".pt"))
if os.environ.get("FAST"):
    from neural_polar_decoder_b200 import rnn_all
    rnn_all.set_gru_precision("fast")
N = 64
dec = RNN_decoder('y_input', N, g["info"], onehot=True)
code = dec._loss_code(dec.info_inds)
y = torch.from_numpy(g["y"]).cuda(); ref = g["logits"]
_, lg = gru_decode(net, code, y, forced=torch.from_numpy(g["decoded"]).cuda(), want_logits=True)
err = np.abs(lg.cpu().numpy() - ref); tol = 1e-2 * np.abs(ref) + 2e-3
# bigger sample against the oracle (1024 fresh frames at 0 dB, forced feedback = the oracle's decisions)
sys.path.insert(0, os.path.join(%r, "oracle"))
import oracle, cpu_arm
rs = np.random.RandomState(1)
info = np.asarray(g["info"])
_, yn = cpu_arm._frames(rs, 1024, N, 22, info.astype(np.int32), 0.0, 6)
sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
do, lo = oracle.gru_decode(sd, yn, N, info)
_, lg2 = gru_decode(net, code, torch.from_numpy(yn).cuda(), forced=torch.from_numpy(do).cuda(), want_logits=True)
e2 = np.abs(lg2.cpu().numpy() - lo); t2 = 1e-2 * np.abs(lo) + 2e-3
print("%%s: 1024 frames vs oracle: max %%.2e worst err/tol %%.2f frac>tol %%.1e  |logit| rms %%.2f" %% (sys.argv[1], e2.max(), (e2 / t2).max(), (e2 > t2).mean(), np.sqrt((lo ** 2).mean())))
w = np.unravel_index((e2 / t2).argmax(), e2.shape); print("   worst entry: ref %%.4f err %%.2e step %%d" %% (lo[w], e2[w], w[1]))
B = 37888
yb = y.repeat((B + y.shape[0] - 1) // y.shape[0], 1)[:B].contiguous()
for _ in range(3): gru_decode(net, code, yb)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(10): gru_decode(net, code, yb)
torch.cuda.synchronize(); dt = (time.time() - t0) / 10
print("%%s: logit err max %%.2e mean %%.2e worst err/tol %%.2f | %%.3f ms per 37888 codewords" %% (sys.argv[1], err.max(), err.mean(), (err / tol).max(), dt * 1e3))
''' % (ROOT, ROOT, ROOT, ROOT, ROOT)
for v in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", CHILD, v])
