"""Drop-in for the hot-path half of the reference's pac_code.py: PAC with pac_encode, channel and
pac_sc_decode (reference pac_code.py:97-119, 121-176, 193-231, 534-573).  Out of scope: Fano decoding,
soft/diff SC variants, brute-force MAP, the rate profiles that need the absent data/ pickles."""
import numpy as np
import torch

from . import _lib, rng
from .construct import count_set_bits
from .utils import llr_scale, snr_db2sigma


def dec2bitarray(in_number, bit_width):
    """MSB-first bit array (reference pac_code.py:40-62)."""
    return np.array([(int(in_number) >> (bit_width - 1 - i)) & 1 for i in range(bit_width)], dtype=int)


class PAC:
    def __init__(self, args, N, K, g, infty=1000., rate_profile='RM'):
        self.N = int(N)
        self.n = int(np.log2(N))
        self.K = int(K)
        self.args = args
        self.g = int(g)
        M = int(np.floor(np.log2(g))) + 1
        self.g_array = 1 - 2 * dec2bitarray(g, M)
        self.rate_profile = rate_profile
        self.infty = infty
        self.B = self._profile(rate_profile)
        self.unsorted_info_positions = self.B.copy()
        self._handles = {}

    # reference pac_code.py:121-176 (schemes that need no data files)
    def _profile(self, scheme, custom_info_positions=None):
        N, K = self.N, self.K
        if custom_info_positions is not None:
            return np.sort(np.asarray(custom_info_positions).copy())
        target_K = getattr(self.args, "target_K", None) or N // 2
        rmweight = np.array([count_set_bits(i) for i in range(N)])
        if scheme == 'RM':
            B = np.argsort(rmweight)[-K:]
        elif scheme == 'rev_RM':
            B = np.argsort(rmweight)[-target_K:][:K].copy()
        elif scheme == 'sorted':
            B = np.sort(np.argsort(rmweight)[-int(target_K):])[:K].copy()
        elif scheme == 'sorted_last':
            B = np.sort(np.argsort(rmweight)[-int(target_K):])[-K:].copy()
        elif scheme == 'last':
            B = np.arange(N - 1, N - K - 1, -1)
        elif scheme == 'freeze_even':
            B = np.arange(N - 1, -1, -2)
        elif scheme == 'freeze_odd':
            B = np.arange(N - 2, -1, -2)
        else:
            raise ValueError("rate profile %r needs data files the reference does not ship "
                             "(pac_code.py:141-160)" % (scheme,))
        return np.sort(B)

    def _handle(self, info=None):
        _lib.require_cuda()
        info = self.B if info is None else info
        key = (torch.cuda.current_device(), np.asarray(info).tobytes())
        h = self._handles.get(key)
        if h is None:
            h = _lib.CodeHandle(self.n, info, self.infty, self.g)
            self._handles[key] = h
        return h

    def extract(self, v_hat, B=None):
        return v_hat[:, self.B if B is None else B]

    def pac_encode(self, msg_bits, scheme=None, custom_info_positions=None):
        """reference pac_code.py:220-224: rate profile -> convolutional pre-coder -> Plotkin transform."""
        if custom_info_positions is not None or scheme is not None:
            self.B = self._profile(scheme if scheme is not None else self.rate_profile, custom_info_positions)
        src_dev = msg_bits.device
        msg = _lib.to_device_f32(msg_bits)
        with torch.cuda.device(msg.device):
            h = self._handle()
            assert msg.dim() == 2 and msg.shape[1] == h.K
            x = torch.empty(msg.shape[0], self.N, dtype=torch.float32, device=msg.device)
            if msg.shape[0] > 0:
                _lib.check(_lib.load().npd_polar_encode(h.h, _lib.ptr(msg), _lib.ptr(x), msg.shape[0],
                                                        _lib.stream_ptr()))
        return x if src_dev.type == "cuda" else x.to(src_dev)

    def channel(self, code, snr, *_ignored, point=None, cw_offset=0, seed=None):
        """reference pac_code.py:226-231."""
        sigma = snr_db2sigma(snr)
        src_dev = code.device
        x = _lib.to_device_f32(code)
        y = torch.empty_like(x)
        if x.numel() == 0:
            return y if src_dev.type == "cuda" else y.to(src_dev)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().npd_awgn(
                _lib.ptr(x), _lib.ptr(y), x.numel() // x.shape[-1], x.shape[-1], float(np.float32(sigma)),
                rng.get_seed() if seed is None else int(seed),
                rng.next_stream() if point is None else int(point), int(cw_offset), _lib.stream_ptr()))
        return y if src_dev.type == "cuda" else y.to(src_dev)

    def pac_sc_decode(self, corrupted_codewords, snr, use_gt_codeword=None):
        """reference pac_code.py:534-573 -> (leaf LLRs [B,N], v_hat[:, B] [B,K], u_hat [B,N])."""
        src_dev = corrupted_codewords.device
        if src_dev.type != "cuda":
            # host tensors: chunked copy/decode/copy pipeline inside the library (npd_pac_sc_decode_host)
            _lib.require_cuda()
            y = _lib.host_f32(corrupted_codewords)
            assert y.dim() == 2 and y.shape[1] == self.N
            Bn = y.shape[0]
            gt = None if use_gt_codeword is None else _lib.host_f32(use_gt_codeword.cpu())
            llr, v, u = (_lib.host_out((Bn, self.N), y), _lib.host_out((Bn, self.K), y),
                         _lib.host_out((Bn, self.N), y))
            if Bn > 0:
                _lib.check(_lib.load().npd_pac_sc_decode_host(self._handle().h, _lib.hptr(y), llr_scale(snr),
                                                              _lib.hptr(gt), _lib.hptr(llr), _lib.hptr(v),
                                                              _lib.hptr(u), Bn))
            return llr, v, u
        y = _lib.to_device_f32(corrupted_codewords)
        assert y.dim() == 2 and y.shape[1] == self.N
        Bn = y.shape[0]
        gt = None if use_gt_codeword is None else _lib.to_device_f32(use_gt_codeword, y.device)
        with torch.cuda.device(y.device):
            h = self._handle()
            llr = torch.empty(Bn, self.N, dtype=torch.float32, device=y.device)
            v = torch.empty(Bn, self.K, dtype=torch.float32, device=y.device)
            u = torch.empty(Bn, self.N, dtype=torch.float32, device=y.device)
            if Bn > 0:
                _lib.check(_lib.load().npd_pac_sc_decode(h.h, _lib.ptr(y), llr_scale(snr), _lib.ptr(gt),
                                                         _lib.ptr(llr), _lib.ptr(v), _lib.ptr(u), Bn,
                                                         _lib.stream_ptr()))
        return llr, v, u
