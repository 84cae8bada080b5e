"""PCIe probe: pinned H2D / D2H / bidirectional bandwidth and the npd_sc_decode_host call by itself."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import numpy as np
from neural_polar_decoder_b200 import PolarCode, construct, _lib
from neural_polar_decoder_b200.utils import llr_scale

MB = 256
h = torch.empty(MB << 18, dtype=torch.float32).pin_memory()
h2 = torch.empty(MB << 18, dtype=torch.float32).pin_memory()
d = torch.empty(MB << 18, dtype=torch.float32, device="cuda")
d2 = torch.empty(MB << 18, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


dt = t(lambda: d.copy_(h, non_blocking=True)); print("H2D %.1f GB/s" % (MB / 1024 / dt))
dt = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H %.1f GB/s" % (MB / 1024 / dt))


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


dt = t(both); print("bidir %.1f GB/s each" % (MB / 1024 / dt))

N, K, B = 1024, 512, 65536
code = PolarCode(10, K, None, F=construct.pw_frozen_set(N, K))
y = torch.randn(B, N).pin_memory()
dec = torch.empty(B, K).pin_memory()
lib = _lib.load()
hh = code._handle()
for chunk in (0, 1024, 4096, 16384):
    _lib.check(lib.npd_host_set_chunk(chunk))
    dt = t(lambda: _lib.check(lib.npd_sc_decode_host(hh.h, _lib.hptr(y), llr_scale(2.0), None, None, _lib.hptr(dec), B)))
    print("npd_sc_decode_host chunk %d: %.2f ms  %.2e cw/s  H2D %.1f GB/s" % (chunk, dt * 1e3, B / dt, B * N * 4 / dt / 1e9))
dt = t(lambda: code.sc_decode_new(y, 2.0, return_llr=False))
print("sc_decode_new(host): %.2f ms %.2e cw/s" % (dt * 1e3, B / dt))
t0 = time.perf_counter(); x = torch.empty(B, K, pin_memory=True); print("pinned alloc %.2f ms" % ((time.perf_counter() - t0) * 1e3))
del x
t0 = time.perf_counter(); x = torch.empty(B, K, pin_memory=True); print("pinned alloc again %.2f ms" % ((time.perf_counter() - t0) * 1e3))
