"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference model code (``/root/reference/magvit2_pytorch/
magvit2_pytorch.py`` + ``attend.py``) in THIS container so that golden vectors
can be generated from the reference itself (tests/golden/, made by
oracle/make_golden.py) and so that oracle/restated.py can be validated.

``/root/reference`` does not exist on the GPU box: nothing that runs there may
import this module.  It is only used by ``oracle/make_golden.py`` and by the
CPU tests that are skipped when the reference tree is absent.

``import magvit2_pytorch`` fails here because the package ``__init__`` pulls the
trainer and five PyPI packages that are not installed.  We
register a synthetic package whose ``__path__`` is the reference directory and
import the model module directly, after installing restated stand-ins for the
model-side third-party modules (oracle/shims/).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MV2_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "magvit2_pytorch", "magvit2_pytorch.py"))


def _install_shims():
    from oracle.shims import quantizers, taylor

    if "vector_quantize_pytorch" not in sys.modules:
        m = types.ModuleType("vector_quantize_pytorch")
        m.LFQ, m.FSQ = quantizers.LFQ, quantizers.FSQ
        m.__shim__ = True
        sys.modules["vector_quantize_pytorch"] = m
    if "taylor_series_linear_attention" not in sys.modules:
        m = types.ModuleType("taylor_series_linear_attention")
        m.TaylorSeriesLinearAttn = taylor.TaylorSeriesLinearAttn
        m.__shim__ = True
        sys.modules["taylor_series_linear_attention"] = m
    if "gateloop_transformer" not in sys.modules:
        from oracle.shims import gateloop
        m = types.ModuleType("gateloop_transformer")
        m.SimpleGateLoopLayer = gateloop.SimpleGateLoopLayer
        m.__shim__ = True
        sys.modules["gateloop_transformer"] = m
    if "kornia" not in sys.modules:
        k = types.ModuleType("kornia")
        kf = types.ModuleType("kornia.filters")

        def filter3d(*a, **kw):  # only reached with antialias=True / discriminator blur
            raise NotImplementedError("kornia.filters.filter3d is not on the hot path")

        kf.filter3d = filter3d
        k.filters = kf
        k.__shim__ = True
        sys.modules["kornia"] = k
        sys.modules["kornia.filters"] = kf


_ref_module = None


def load_reference():
    """Return the reference ``magvit2_pytorch.magvit2_pytorch`` module object."""
    global _ref_module
    if _ref_module is not None:
        return _ref_module
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_shims()
    pkg_name = "magvit2_pytorch"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "magvit2_pytorch")]
        pkg.__synthetic__ = True
        sys.modules[pkg_name] = pkg
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _ref_module = importlib.import_module("magvit2_pytorch.magvit2_pytorch")
    return _ref_module


def build_reference_tokenizer(**kwargs):
    """Construct the reference ``VideoTokenizer`` (VGG/GAN branches disabled: they
    are not on the tokenize/decode path and VGG would need a download)."""
    ref = load_reference()
    kwargs.setdefault("perceptual_loss_weight", 0.)
    kwargs.setdefault("use_gan", False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref.VideoTokenizer(**kwargs)
