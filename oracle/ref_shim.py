"""Import shim for the live CRISP reference (TEST INFRASTRUCTURE ONLY).

The reference (/root/reference, read-only, pure Python/PyTorch) cannot be imported as shipped in this
image: it imports matplotlib / IPython at module top (polar.py:7-11, pac_code.py:5, rnn_all.py:15-17,
32-36, models.py:11,17-19) and rnn_all.py parses sys.argv at import-free __main__ only.  This module
stubs those imports, puts /root/reference on sys.path and hands back the reference modules so that
`oracle/gen_golden.py` can mint fixtures and `tests/test_oracle_vs_reference.py` can pin the oracle.

/root/reference does not exist on the GPU box: nothing that runs there may import this module
(`available()` returns False there and callers skip).
"""
import os
import sys
import types
import importlib

REF_DIR = os.environ.get("NPD_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "polar.py"))


class _Anything(types.ModuleType):
    """A module whose every attribute is a callable returning another _Anything-like dummy."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _Dummy:
    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __getitem__(self, k):
        return _Dummy()

    def __setitem__(self, k, v):
        pass

    def __contains__(self, k):
        return False

    def update(self, *a, **k):
        pass


def _stub_plotting():
    for name in ("matplotlib", "matplotlib.pyplot", "IPython", "IPython.display"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            m = _Anything(name)
            if name == "matplotlib":
                m.get_backend = lambda: "agg"
                m.use = lambda *a, **k: None
            if name == "matplotlib.pyplot":
                m.rcParams = _Dummy()
            sys.modules[name] = m
    if isinstance(sys.modules.get("matplotlib"), _Anything):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if isinstance(sys.modules.get("IPython"), _Anything):
        sys.modules["IPython"].display = sys.modules["IPython.display"]


_loaded = {}


def load(name: str):
    """Return reference module `name` (utils, polar, pac_code, rnn_all, models)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_DIR)
    if name in _loaded:
        return _loaded[name]
    _stub_plotting()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    # The reference's own module names (utils, polar, models ...) are generic; import them under a
    # guard so that a same-named module of the repo under test can never shadow them.
    for generic in ("utils", "polar", "pac_code", "rnn_all", "models"):
        mod = sys.modules.get(generic)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_DIR):
            del sys.modules[generic]
    argv = sys.argv
    sys.argv = [argv[0]]
    try:
        mod = importlib.import_module(name)
    finally:
        sys.argv = argv
    _loaded[name] = mod
    return mod


def make_args(N=64, K=22, **kw):
    """Namespace with the fields the hot path reads (polar.py:400 hard_decision, pac_code.py:128
    target_K, rnn_all.py:1183-1192 loss_only / K)."""
    import argparse
    ns = argparse.Namespace(hard_decision=True, target_K=K, loss_only=None, N=N, K=K,
                            soft_sign=False, no_detach=False, random_seed=0, g=None)
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns


def get_code(code_type, rate_profile, N, K, g=None, **kw):
    """Reference get_code (rnn_all.py:1015-1196) with its module-global `args` provided."""
    ra = load("rnn_all")
    ra.args = make_args(N=N, K=K, g=g, **kw)
    return ra.get_code(code_type, rate_profile, N, K, g)
