"""ctypes binding of libnpd.so (include/npd.h).  The CUDA library is the product: if it is missing or
there is no CUDA device the package raises -- there is no CPU or eager-PyTorch fallback."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnpd.so")

NPD_OK, NPD_EINVAL, NPD_ECUDA, NPD_EUNSUPPORTED, NPD_ENOMEM = 0, -1, -2, -3, -4

_c = ctypes
_vp, _i64, _u64, _u32, _f32, _int, _sz = (_c.c_void_p, _c.c_int64, _c.c_uint64, _c.c_uint32,
                                            _c.c_float, _c.c_int, _c.c_size_t)

# name -> (restype, argtypes); every symbol declared in include/npd.h
SIGNATURES = {
    "npd_version": (_int, []),
    "npd_last_error": (_c.c_char_p, []),
    "npd_device_info": (_int, [_c.POINTER(_int)] * 3 + [_c.c_char_p, _int]),
    "npd_code_create": (_int, [_int, _int, _vp, _f32, _u32, _c.POINTER(_vp)]),
    "npd_code_destroy": (_int, [_vp]),
    "npd_polar_encode": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "npd_awgn": (_int, [_vp, _vp, _i64, _int, _f32, _u64, _u32, _u64, _vp]),
    "npd_gen_encode_awgn": (_int, [_vp, _vp, _vp, _vp, _i64, _f32, _u64, _u32, _u64, _vp]),
    "npd_sc_decode": (_int, [_vp, _vp, _f32, _vp, _vp, _vp, _i64, _vp]),
    "npd_sc_round_codewords": (_i64, [_vp]),
    "npd_scl_decode": (_int, [_vp, _vp, _f32, _int, _vp, _vp, _i64, _vp]),
    "npd_scl_decode_host": (_int, [_vp, _vp, _f32, _int, _vp, _vp, _i64]),
    "npd_pac_sc_decode": (_int, [_vp, _vp, _f32, _vp, _vp, _vp, _vp, _i64, _vp]),
    "npd_count_errors": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "npd_count_errors_info": (_int, [_vp, _vp, _vp, _i64, _int, _vp, _vp]),
    "npd_mc_gru_workspace_bytes": (_sz, [_vp, _vp, _i64]),
    "npd_mc_gru_sweep": (_int, [_vp, _vp, _vp, _i64, _i64, _f32, _u64, _u32, _u64, _vp, _sz, _vp, _vp]),
    "npd_mc_sc_workspace_bytes": (_sz, [_vp, _i64]),
    "npd_mc_sc_sweep": (_int, [_vp, _i64, _i64, _f32, _f32, _u64, _u32, _u64, _vp, _sz, _vp, _vp]),
    "npd_gru_create": (_int, [_int, _int] + [_vp] * 10 + [_c.POINTER(_vp)]),
    "npd_gru_destroy": (_int, [_vp]),
    "npd_gru_set_head_mlp": (_int, [_vp, _int, _int, _vp]),
    "npd_gru_workspace_bytes": (_sz, [_vp, _i64]),
    "npd_gru_set_option": (_int, [_vp, _int, _int]),
    "npd_gru_decode": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "npd_gru_decode_h0": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "npd_gru_trainer_param_count": (_sz, [_int, _int]),
    "npd_gru_trainer_create": (_int, [_int, _int, _i64, _vp, _int, _c.POINTER(_vp)]),
    "npd_gru_trainer_destroy": (_int, [_vp]),
    "npd_gru_trainer_get": (_int, [_vp, _int, _vp]),
    "npd_gru_trainer_set_params": (_int, [_vp, _vp, _int]),
    "npd_gru_train_step": (_int, [_vp, _vp, _vp, _vp, _int, _i64, _f32, _f32, _int, _vp, _vp, _vp]),
    "npd_conv_create": (_int, [_int, _int, _vp, _sz, _c.POINTER(_vp)]),
    "npd_conv_destroy": (_int, [_vp]),
    "npd_conv_workspace_bytes": (_sz, [_vp, _i64]),
    "npd_conv_forward": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "npd_sc_decode_host": (_int, [_vp, _vp, _f32, _vp, _vp, _vp, _i64]),
    "npd_host_set_chunk": (_int, [_i64]),
    "npd_pac_sc_decode_host": (_int, [_vp, _vp, _f32, _vp, _vp, _vp, _vp, _i64]),
    "npd_gru_decode_host": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "npd_conv_forward_host": (_int, [_vp, _vp, _vp, _vp, _i64]),
    "npd_conv_decode": (_int, [_vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "npd_conv_decode_host": (_int, [_vp, _vp, _vp, _i64]),
}

_lib = None


class NpdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libnpd error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load libnpd.so (no CUDA call is made here, so this works on a CPU-only box)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "neural_polar_decoder_b200: %s is missing -- build it with `python -c 'import "
            "__graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise NpdError(rc, load().npd_last_error().decode("utf-8", "replace"))


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("neural_polar_decoder_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")


def stream_ptr():
    return _vp(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous float32 CUDA tensor (or None)."""
    if t is None:
        return _vp(0)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return _vp(t.data_ptr())


def to_device_f32(t, device=None):
    """Contiguous float32 CUDA view/copy of `t`; host tensors are copied (pinned -> async)."""
    require_cuda()
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_cuda:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def host_f32(t):
    """Contiguous float32 HOST tensor for the npd_*_host entry points (no copy when it already is one)."""
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    assert not t.is_cuda
    return t.to(torch.float32).contiguous()


def host_out(shape, like):
    """Fresh host output tensor; pinned when the caller's input is pinned (the copies are then truly async)."""
    return torch.empty(shape, dtype=torch.float32, pin_memory=bool(like.is_pinned()))


def hptr(t):
    return _vp(0) if t is None else _vp(t.data_ptr())


def to_host(t, like):
    """Device tensor -> host tensor shaped like the caller's input: pinned in, pinned out (async copy +
    one stream sync), pageable otherwise."""
    if t is None:
        return None
    if like.is_pinned():
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return out
    return t.to(like.device)


class CodeHandle:
    """Owns an npd_code_t (device-side info / frozen tables)."""

    def __init__(self, n, info_positions, infty=1000.0, pac_g=0):
        import numpy as np
        require_cuda()
        lib = load()
        info = np.ascontiguousarray(np.asarray(info_positions, dtype=np.int32))
        self.n, self.N, self.K = int(n), 1 << int(n), int(info.shape[0])
        h = _vp()
        check(lib.npd_code_create(self.n, self.K, _vp(info.ctypes.data), float(infty), int(pac_g),
                                  ctypes.byref(h)))
        self.h = h
        self.device = torch.cuda.current_device()

    def __del__(self):
        try:
            if getattr(self, "h", None) and _lib is not None:
                _lib.npd_code_destroy(self.h)
                self.h = None
        except Exception:
            pass
