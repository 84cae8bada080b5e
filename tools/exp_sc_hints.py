"""Round-2 experiment: L2 cache hints in the SC quad kernel (libnpd_s0 = none, libnpd_s1 = streaming decision stores + evict-first
loads in the last pass over y), at 12 / 14 / 16 resident warps per SM (debug-knob builds: NPD_SC_WPB / NPD_SC_WARPS)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, time, numpy as np, torch
sys.path.insert(0, %r)
from neural_polar_decoder_b200 import _lib
_lib.LIB_PATH = os.path.join(%r, "neural_polar_decoder_b200", "libnpd_%%s.so" %% sys.argv[1])
from neural_polar_decoder_b200 import PolarCode, construct, utils
N, K, B = int(sys.argv[2]), int(sys.argv[2]) // 2, int(sys.argv[3])
code = PolarCode(int(np.log2(N)), K, None, F=construct.pw_frozen_set(N, K))
lib = _lib.load(); h = code._handle()
msg = torch.empty(B, K, device="cuda"); y = torch.empty(B, N, device="cuda"); dec = torch.empty(B, K, device="cuda")
_lib.check(lib.npd_gen_encode_awgn(h.h, _lib.ptr(msg), None, _lib.ptr(y), B, float(np.float32(utils.snr_db2sigma(2.0))), 1, 0, 0, _lib.stream_ptr()))
def run(): _lib.check(lib.npd_sc_decode(h.h, _lib.ptr(y), utils.llr_scale(2.0), None, None, _lib.ptr(dec), B, _lib.stream_ptr()))
for _ in range(5): run()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(30): run()
torch.cuda.synchronize(); dt = (time.time() - t0) / 30
print("%%s N=%%d WPB=%%s WARPS=%%s: %%.4f ms  %%.3e cw/s  ber %%.4f" %% (sys.argv[1], N, os.environ.get("NPD_SC_WPB"), os.environ.get("NPD_SC_WARPS"), dt * 1e3, B / dt, (dec != msg).float().mean().item()))
''' % (ROOT, ROOT)
for lib in ("s0", "s1"):
    for wpb, warps in ((4, 12), (7, 14), (4, 16), (5, 15)):
        env = dict(os.environ, NPD_SC_WPB=str(wpb), NPD_SC_WARPS=str(warps))
        subprocess.run([sys.executable, "-c", CHILD, lib, "1024", "131072"], env=env)
