"""Metrics of the hot path (reference utils.py:5-6, 17-51) on top of libnpd's fused counter."""
import numpy as np
import torch

from . import _lib


def snr_db2sigma(train_snr):
    """reference utils.py:5-6 (Python double)."""
    return 10 ** (-train_snr * 1.0 / 20)


def llr_scale(snr_db):
    """fp32(2/sigma^2): torch multiplies an fp32 tensor by a Python scalar after rounding the scalar
    to fp32 (reference polar.py:467-469)."""
    sigma = snr_db2sigma(snr_db)
    return float(np.float32(2 / sigma ** 2))


def count_errors(y_true, y_pred):
    """-> (bit_errors, block_errors, numel, rows): numerators of errors_ber / errors_bler computed by
    npd_count_errors on the device (one 16-byte D2H read)."""
    a = _lib.to_device_f32(y_true)
    b = _lib.to_device_f32(y_pred, a.device)
    rows = a.shape[0]
    a = a.reshape(rows, -1)
    b = b.reshape(rows, -1)
    assert a.shape == b.shape, (a.shape, b.shape)
    counts = torch.zeros(2, dtype=torch.int64, device=a.device)
    if rows > 0 and a.shape[1] > 0:
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().npd_count_errors(_lib.ptr(a), _lib.ptr(b), rows, a.shape[1],
                                                    _lib._vp(counts.data_ptr()), _lib.stream_ptr()))
    c = counts.tolist()
    return c[0], c[1], a.numel(), rows


def errors_ber(y_true, y_pred, mask=None):
    """reference utils.py:17-25: fraction of elements whose rounded values differ.  Returns a 0-dim
    tensor on the inputs' device like the reference.  `mask` (rarely used by the eval loops) is
    applied on the device with plain tensor ops."""
    if mask is not None:
        yt = y_true.reshape(y_true.shape[0], -1)
        yp = y_pred.reshape(y_pred.shape[0], -1)
        m = mask.reshape(mask.shape[0], -1).to(yt.dtype)
        return (m * torch.ne(torch.round(yt), torch.round(yp)).to(yt.dtype)).sum() / m.sum()
    bit, _, numel, _ = count_errors(y_true, y_pred)
    dev = y_true.device if torch.is_tensor(y_true) else "cpu"
    return torch.tensor(bit / max(numel, 1), dtype=torch.float32, device=dev)


def errors_bler(y_true, y_pred, get_pos=False):
    """reference utils.py:37-51: fraction of rows with any mismatch (Python float)."""
    if get_pos:
        yt = torch.round(y_true.reshape(y_true.shape[0], -1))
        yp = torch.round(y_pred.reshape(y_pred.shape[0], -1))
        bad = (yt != yp).any(dim=1)
        return bad.float().mean().item(), list(torch.nonzero(bad).flatten().tolist())
    _, blk, _, rows = count_errors(y_true, y_pred)
    return blk * 1.0 / max(rows, 1)


def errors_bitwise_ber(y_true, y_pred, mask=None):
    """reference utils.py:27-35 (per-position error rate; diagnostics only, plain tensor ops)."""
    yt = y_true.reshape(y_true.shape[0], -1)
    yp = y_pred.reshape(y_pred.shape[0], -1)
    ne = torch.ne(torch.round(yt), torch.round(yp)).float()
    if mask is None:
        return ne.mean(dim=0).unsqueeze(-1)
    m = mask.reshape(mask.shape[0], -1).float()
    return ((m * ne).sum(0) / m.sum(0)).unsqueeze(-1)
