"""Drop-in for the hot-path half of the reference's polar.py: PolarCode with encode_plotkin, channel
and sc_decode_new, each forwarding to one libnpd.so entry point (hand-written sm_100a kernels).

Signatures, attribute names and tensor contracts follow reference polar.py:66-148, 201-207, 465-484.
`scl_decode` (SC-list, polar.py:777-876) is covered too.  Out of scope: the recursive exact-LSE
`sc_decode`, soft/neural SC variants, CRC, bitwise MAP and the analysis/plotting helpers.
"""
import numpy as np
import torch

from . import _lib, rng
from .utils import errors_ber, errors_bler, llr_scale, snr_db2sigma  # noqa: F401 (re-exported like the reference)


class PolarCode:
    def __init__(self, n, K, args=None, F=None, rs=None, use_cuda=True, infty=1000.):
        # reference polar.py:66-117 (index-set construction; integer work, kept on the host)
        assert n >= 1
        self.args = args
        self.n = int(n)
        self.N = 2 ** self.n
        self.K = int(K)
        self.infty = infty
        self.device = torch.device("cuda" if use_cuda else "cpu")

        if F is not None:
            F = np.array(F, dtype=np.int64)
            assert len(F) == self.N - self.K
            self.frozen_positions = np.sort(F)
            self.unsorted_frozen_positions = self.frozen_positions
            self.info_positions = np.array(sorted(set(range(self.N)) - set(self.frozen_positions.tolist())),
                                           dtype=np.int64)
            self.unsorted_info_positions = self.info_positions
        else:
            if rs is None:
                # reference default: "increasing order of reliability" = 1023..0, i.e. the K largest
                # indices carry information (polar.py:93-96); only defined for N <= 1024
                self.reliability_seq = np.arange(1023, -1, -1)
                self.rs = self.reliability_seq[self.reliability_seq < self.N]
            else:
                self.reliability_seq = np.asarray(rs)
                self.rs = self.reliability_seq[self.reliability_seq < self.N]
                assert len(self.rs) == self.N
            head = self.rs[:self.K].copy()
            self.unsorted_info_positions = np.flip(head.copy())
            self.info_positions = np.sort(head)
            self.frozen_positions = np.sort(self.rs[self.K:].copy())
            self.unsorted_frozen_positions = self.frozen_positions
        assert len(self.info_positions) == self.K
        self._handles = {}

    # ---- libnpd code objects (one per device and info set) ----
    def _handle(self, info_positions=None):
        _lib.require_cuda()
        info = self.info_positions if info_positions is None else np.sort(np.asarray(info_positions))
        key = (torch.cuda.current_device(), info.tobytes(), float(self.infty))
        h = self._handles.get(key)
        if h is None:
            h = _lib.CodeHandle(self.n, info, self.infty, 0)
            self._handles[key] = h
        return h

    @property
    def G(self):
        # dense generator matrix (reference polar.py:73-77); not used by the hot path, built lazily
        g = np.array([1])
        for _ in range(self.n):
            g = np.kron(g, np.array([[1, 0], [1, 1]]))
        return torch.from_numpy(g).float()

    def encode_plotkin(self, message, scaling=None, custom_info_positions=None):
        """reference polar.py:128-148.  message [B,K] of +-1 -> codeword [B,N] of +-1 (same device)."""
        src_dev = message.device
        msg = _lib.to_device_f32(message)
        if custom_info_positions is not None:
            # `u[:, custom_info_positions] = message` (polar.py:133-136): message column j goes to position custom[j]
            # in the caller's order; the code handle keeps positions sorted, so permute the columns instead
            custom = np.asarray(custom_info_positions)
            order = np.argsort(custom, kind="stable")
            if not np.array_equal(order, np.arange(len(order))):
                msg = msg.index_select(1, torch.as_tensor(order, device=msg.device)).contiguous()
        with torch.cuda.device(msg.device):
            h = self._handle(custom_info_positions)
            assert msg.dim() == 2 and msg.shape[1] == h.K, (tuple(msg.shape), h.K)
            x = torch.empty(msg.shape[0], self.N, dtype=torch.float32, device=msg.device)
            if msg.shape[0] > 0:
                _lib.check(_lib.load().npd_polar_encode(h.h, _lib.ptr(msg), _lib.ptr(x), msg.shape[0],
                                                        _lib.stream_ptr()))
        if scaling is not None:  # polar.py:146-147 (unused by the eval loops)
            scaling = scaling.to(x.device)
            x = (scaling * np.sqrt(self.N) * x) / torch.norm(scaling)
        return x if src_dev.type == "cuda" else x.to(src_dev)

    def channel(self, code, snr, *_ignored, point=None, cw_offset=0, seed=None):
        """reference polar.py:201-207: code + sigma * N(0,1).  Extra positionals are swallowed (the
        reference's own callers pass four more, rnn_all.py:847).  Noise is device Philox (rng.py)."""
        sigma = snr_db2sigma(snr)
        src_dev = code.device
        x = _lib.to_device_f32(code)
        y = torch.empty_like(x)
        cols = x.shape[-1]
        if x.numel() == 0:
            return y if src_dev.type == "cuda" else y.to(src_dev)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().npd_awgn(
                _lib.ptr(x), _lib.ptr(y), x.numel() // cols, cols,
                float(np.float32(sigma)), rng.get_seed() if seed is None else int(seed),
                rng.next_stream() if point is None else int(point), int(cw_offset), _lib.stream_ptr()))
        return y if src_dev.type == "cuda" else y.to(src_dev)

    def sc_decode_new(self, corrupted_codewords, snr, use_gt=None, return_llr=True):
        """reference polar.py:465-484: min-sum SC, bit-exact in fp32.
        -> (leaf LLRs [B,N] incl. the +infty frozen prior, u_hat[:, info_positions] [B,K] in {-1,0,+1}).
        return_llr=False skips the LLR output (first element is None)."""
        src = corrupted_codewords
        if torch.is_tensor(src) and not src.is_cuda:
            return self._sc_decode_host(src, snr, use_gt, return_llr)
        y = _lib.to_device_f32(corrupted_codewords)
        assert y.dim() == 2 and y.shape[1] == self.N, tuple(y.shape)
        B = y.shape[0]
        gt = None if use_gt is None else _lib.to_device_f32(use_gt, y.device)
        with torch.cuda.device(y.device):
            h = self._handle()
            llr = torch.empty(B, self.N, dtype=torch.float32, device=y.device) if return_llr else None
            dec = torch.empty(B, self.K, dtype=torch.float32, device=y.device)
            if B > 0:
                _lib.check(_lib.load().npd_sc_decode(h.h, _lib.ptr(y), llr_scale(snr), _lib.ptr(gt),
                                                     _lib.ptr(llr), _lib.ptr(dec), B, _lib.stream_ptr()))
        return llr, dec

    def scl_decode(self, corrupted_codewords, snr, L=1, use_CRC=False, return_llr=True):
        """reference polar.py:793-876: SC-list decoding with L paths and the ML pick among the list.
        -> (leaf LLRs [B,N] of the chosen path, its u_hat[:, info_positions] [B,K]), on the caller's device.
        use_CRC=True (the reference's per-codeword Python CRC check) is not on the accelerated path."""
        if use_CRC:
            raise NotImplementedError("CRC-aided list decoding (polar.py:738-775, 848-867) is out of scope of the B200 path")
        src = corrupted_codewords
        L = int(L)
        if torch.is_tensor(src) and not src.is_cuda:
            _lib.require_cuda()
            y = _lib.host_f32(src)
            assert y.dim() == 2 and y.shape[1] == self.N, tuple(y.shape)
            B = y.shape[0]
            llr = _lib.host_out((B, self.N), y) if return_llr else None
            dec = _lib.host_out((B, self.K), y)
            if B > 0:
                _lib.check(_lib.load().npd_scl_decode_host(self._handle().h, _lib.hptr(y), llr_scale(snr), L,
                                                           _lib.hptr(llr), _lib.hptr(dec), B))
            return llr, dec
        y = _lib.to_device_f32(src)
        assert y.dim() == 2 and y.shape[1] == self.N, tuple(y.shape)
        B = y.shape[0]
        with torch.cuda.device(y.device):
            h = self._handle()
            llr = torch.empty(B, self.N, dtype=torch.float32, device=y.device) if return_llr else None
            dec = torch.empty(B, self.K, dtype=torch.float32, device=y.device)
            if B > 0:
                _lib.check(_lib.load().npd_scl_decode(h.h, _lib.ptr(y), llr_scale(snr), L, _lib.ptr(llr), _lib.ptr(dec),
                                                      B, _lib.stream_ptr()))
        return llr, dec

    def _sc_decode_host(self, y, snr, use_gt, return_llr):
        """Host tensors in, host tensors out: npd_sc_decode_host overlaps the chunked H2D copy, the
        kernels and the D2H copy on the library's own streams (synchronous)."""
        _lib.require_cuda()
        y = _lib.host_f32(y)
        assert y.dim() == 2 and y.shape[1] == self.N, tuple(y.shape)
        B = y.shape[0]
        gt = None if use_gt is None else _lib.host_f32(use_gt.cpu() if torch.is_tensor(use_gt) else use_gt)
        llr = _lib.host_out((B, self.N), y) if return_llr else None
        dec = _lib.host_out((B, self.K), y)
        if B > 0:
            _lib.check(_lib.load().npd_sc_decode_host(self._handle().h, _lib.hptr(y), llr_scale(snr), _lib.hptr(gt),
                                                      _lib.hptr(llr), _lib.hptr(dec), B))
        return llr, dec
