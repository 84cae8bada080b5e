"""Counter-based noise/message streams of the drop-in (Philox4x32-10 on the device).

The reference draws torch.randn on the CPU default generator (polar.py:204); that stream cannot be
reproduced on the device, so the drop-in's channel() is a *statistical* equivalent: key = seed,
counter = (global codeword index, sample quad, stream id).  Every channel() call takes a fresh stream
id so that successive calls are independent, and a sweep that passes explicit (point, cw_offset)
values gets noise that does not depend on how the batch is split over GPUs."""
import threading

_state = threading.local()


def manual_seed(seed: int):
    _state.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state.calls = 0


def get_seed() -> int:
    if not hasattr(_state, "seed"):
        manual_seed(0)
    return _state.seed


def next_stream() -> int:
    """Stream id for an anonymous channel() call (wraps at 2^31; ids >= 2^31 are reserved for sweeps)."""
    get_seed()
    s = _state.calls
    _state.calls = (s + 1) & 0x7FFFFFFF
    return s
