"""Where do the 1.4 ms per step outside the GRU kernel go? (bench-only experiment)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from neural_polar_decoder_b200 import _lib, construct, synth
from neural_polar_decoder_b200.rnn_all import RNN_Model, RNN_decoder
N, K, H, B = 64, 22, 512, 37888
lib = _lib.load()
sd = synth.gru_state_dict(11, N, H, 2)
net = RNN_Model('GRU', N + 2, H, 1, 2, N, 0, 0); net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
rs = construct.reference_rs256(); info = np.sort(rs[rs < N][:K])
dec = RNN_decoder('y_input', N, info, onehot=True)
gh = net.npd_handle(N); code = dec._loss_code(info)
y = torch.randn(B, N, device="cuda"); decoded = torch.empty(B, N, device="cuda")
msg = torch.ones(B, K, device="cuda"); dec_info = torch.empty(B, K, device="cuda")
info_t = torch.as_tensor(info, device="cuda"); counts = torch.zeros(3, dtype=torch.int64, device="cuda")
st = _lib.stream_ptr()
def k(): _lib.check(lib.npd_gru_decode(gh.h, code.h, _lib.ptr(y), None, None, None, _lib.ptr(decoded), B, None, 0, st))
def sel(): torch.index_select(decoded, 1, info_t, out=dec_info)
def cnt(): _lib.check(lib.npd_count_errors(_lib.ptr(msg), _lib.ptr(dec_info), B, K, _lib._vp(counts.data_ptr()), st))
def timeit(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    print("%-28s %.3f ms per call" % (name, a.elapsed_time(b) / n))
timeit("kernel only", k)
timeit("index_select", sel)
timeit("count", cnt)
timeit("kernel+select+count", lambda: (k(), sel(), cnt()))
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
timeit("kernel with events", lambda: (e[0].record(), k(), e[1].record()))
timeit("kernel only (again)", k)
