"""Code constructions used by the drop-in (integer work, host side).

* N <= 256: the reliability order the reference embeds in get_code (rnn_all.py:1046, "computed for
  SNR = 0"), shipped as data/polar_rs256.json (extracted by tools/extract_tables.py).
* N  > 256: the reference has no construction; the builder-supplied one is the
  polarization-weight order W(i) = sum_j b_j(i) * beta^j with beta = 2^(1/4), handed to both the
  oracle and the kernels through PolarCode(..., F=...).
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_RS256 = None


def reference_rs256():
    """Most-reliable-first order over 0..255 (reference rnn_all.py:1046)."""
    global _RS256
    if _RS256 is None:
        with open(os.path.join(_HERE, "data", "polar_rs256.json")) as f:
            _RS256 = np.array(json.load(f)["rs"], dtype=np.int64)
    return _RS256.copy()


def count_set_bits(i):
    return bin(int(i)).count("1")


def polarization_weight_order(N, beta=2 ** 0.25):
    """Most-reliable-first index order by polarization weight (ties broken by index, descending)."""
    n = int(np.log2(N))
    idx = np.arange(N)
    w = np.zeros(N)
    for j in range(n):
        w += ((idx >> j) & 1) * (beta ** j)
    return np.lexsort((-idx, -w))  # primary: weight descending


def pw_frozen_set(N, K):
    """Sorted frozen positions (N-K least reliable) for PolarCode(n, K, args, F=...)."""
    order = polarization_weight_order(N)
    return np.sort(order[K:])


def rm_info_set(N, K):
    """pac_code.py:115-119 / 132-135: K indices of highest Hamming weight (numpy argsort order)."""
    rmweight = np.array([count_set_bits(i) for i in range(N)])
    B = np.argsort(rmweight)[-K:]
    return np.sort(B)
