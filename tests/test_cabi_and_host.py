"""CPU: the C-ABI library loads and exports every symbol include/npd.h declares (no compute call without
a GPU), and the host-side logic of the drop-in (index sets, sharding, rate averaging, RNG streams)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "npd.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(npd_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported_and_bound():
    from neural_polar_decoder_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libnpd.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding and header disagree"
    assert _lib.load().npd_version() == 100


def test_no_cpu_fallback():
    """Without a CUDA device every hot-path entry point raises instead of computing on the host."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from neural_polar_decoder_b200 import PolarCode, PAC, errors_ber
    code = PolarCode(3, 4)
    with pytest.raises(RuntimeError):
        code.encode_plotkin(torch.ones(2, 4))
    with pytest.raises(RuntimeError):
        code.sc_decode_new(torch.ones(2, 8), 1.0)
    with pytest.raises(RuntimeError):
        PAC(None, 8, 4, 7).pac_sc_decode(torch.ones(2, 8), 1.0)
    with pytest.raises(RuntimeError):
        errors_ber(torch.ones(2, 4), torch.ones(2, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "neural_polar_decoder_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src and "ref_shim" not in src, fn


def test_info_sets_match_reference(golden):
    from neural_polar_decoder_b200 import PAC
    from neural_polar_decoder_b200.rnn_all import get_code
    g = golden("misc")
    import argparse
    for key in g.files:
        if key.startswith("info_"):
            _, prof, N, K = key.split("_", 3) if not key.startswith("info_rev") else (None, "rev_polar") + tuple(key.split("_")[3:])
            N, K = int(N), int(K)
            code = get_code("Polar", prof, N, K, args=argparse.Namespace(target_K=K))
            assert np.array_equal(code.info_positions, g[key]), key
            assert np.array_equal(np.sort(np.concatenate([code.info_positions, code.frozen_positions])), np.arange(N))
        if key.startswith("pacinfo_"):
            _, N, K = key.split("_")
            pac = PAC(None, int(N), int(K), 53 if int(N) < 128 else 133)
            assert np.array_equal(pac.B, g[key]), key
    assert np.array_equal(PAC(None, 32, 16, 53).g_array, g["pacg_53"])
    # SURVEY.md KAT5
    code = get_code("Polar", "polar", 64, 22)
    assert code.info_positions.tolist() == [27, 29, 30, 31, 39, 43, 45, 46, 47] + list(range(51, 64))


def test_curriculum_info_sets_match_reference(golden):
    """K < target_K (every run_crisp.sh stage but the last) for every rate profile, and --loss_only."""
    import argparse
    from neural_polar_decoder_b200.rnn_all import get_code
    g = golden("misc")
    seen = 0
    for key in g.files:
        if not key.startswith("cur_"):
            continue
        prof, N, K, tK = key[4:].rsplit("_", 3)
        code = get_code("Polar", prof, int(N), int(K), args=argparse.Namespace(target_K=int(tK), random_seed=42))
        assert np.array_equal(code.info_positions, g[key]), key
        seen += 1
    assert seen == 15
    code = get_code("Polar", "rev_polar", 64, 22, args=argparse.Namespace(target_K=22, loss_only=6))
    assert np.array_equal(code.loss_inds, g["lossonly_inds"])
    assert np.array_equal(code.msg_indices, g["lossonly_msg"])


def test_polarcode_constructor_variants():
    from neural_polar_decoder_b200 import PolarCode, construct
    c = PolarCode(3, 4)  # rs=None default: the K largest indices (polar.py:93-96)
    assert c.info_positions.tolist() == [4, 5, 6, 7] and c.frozen_positions.tolist() == [0, 1, 2, 3]
    F = construct.pw_frozen_set(1024, 512)
    c = PolarCode(10, 512, None, F=F)
    assert len(c.info_positions) == 512 and len(set(c.info_positions) & set(F)) == 0
    with pytest.raises(AssertionError):
        PolarCode(3, 4, None, F=np.array([0, 1, 2]))
    assert tuple(c.G.shape) == (1024, 1024) and c.G[1023].sum() == 1024  # lazy dense generator


def test_llr_scale_is_fp32_rounded_scalar():
    from neural_polar_decoder_b200 import utils
    for snr in (-2.0, 0.0, 1.0, 3.5):
        sigma = 10 ** (-snr / 20)
        y = torch.tensor([0.3, -1.7, 2.9])
        assert torch.equal((2 / sigma ** 2) * y, torch.tensor(utils.llr_scale(snr)) * y)


def test_sharding_and_rates():
    from neural_polar_decoder_b200 import sweep
    for total in (0, 1, 7, 100, 10 ** 9 + 3):
        for world in (1, 2, 3, 8):
            cuts = [sweep.shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    # per-batch rates are averaged like the reference (+= ber / num_test_batches)
    counts = torch.zeros(2, 1, 1, 2, dtype=torch.int64)
    counts[0, 0, 0] = torch.tensor([10, 2])
    counts[1, 0, 0] = torch.tensor([1, 1])
    (ber, bler), = sweep._rates(counts, [100, 10], 4, 1, 1)
    assert ber[0] == pytest.approx((10 / 400 + 1 / 40) / 2) and bler[0] == pytest.approx((2 / 100 + 1 / 10) / 2)
    b, l, f = sweep.finalize(torch.tensor([[8, 2, 4], [0, 0, 0]]), 2)
    assert b == [1.0, 0.0] and l == [0.5, 0.0] and f == [4, 0]


def test_rng_streams():
    from neural_polar_decoder_b200 import rng
    rng.manual_seed(5)
    assert rng.get_seed() == 5 and [rng.next_stream() for _ in range(3)] == [0, 1, 2]
    rng.manual_seed(5)
    assert rng.next_stream() == 0


def test_synthetic_weights_are_deterministic():
    from neural_polar_decoder_b200 import synth
    a, b = synth.gru_state_dict(3, 16, 64), synth.gru_state_dict(3, 16, 64)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert a["rnn.weight_ih_l0"].shape == (192, 18) and a["linear.weight"].shape == (1, 64)
    c = synth.conv_state_dict(1)
    assert sum(v.size for v in c.values()) == 2492096  # SURVEY.md a9 parameter count
