"""CPU-side checks of the drop-in boundary: libgifgan.so loads without a GPU, exports every entry point that
include/gifgan.h declares, and the ctypes table in gifgan/_cabi.py mirrors the header one to one.  No compute
calls are made here (those are the `-m gpu` tests)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gifgan.h")
LIB = os.path.join(ROOT, "gif-gan_b200", "lib", "libgifgan.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_survey_abi():
    names = declared_symbols()
    for must in ("gg_conv2d_fwd", "gg_conv2d_dgrad", "gg_conv2d_wgrad", "gg_deconv2d_fwd", "gg_deconv2d_dgrad", "gg_deconv2d_wgrad",
                 "gg_conv3d_fwd", "gg_bn_fwd_train", "gg_bn_fwd_infer", "gg_bn_bwd", "gg_linear_fwd", "gg_lstm_step_fwd", "gg_lstm_step_bwd",
                 "gg_sigmoid_ce", "gg_adam", "gg_last_error", "gg_version"):
        assert must in names, must


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    lib.gg_version.restype = ctypes.c_int
    assert lib.gg_version() == 100
    lib.gg_last_error.restype = ctypes.c_char_p
    assert lib.gg_last_error() is not None


def test_ctypes_table_mirrors_the_header():
    from gifgan import _cabi
    declared = set(declared_symbols())
    bound = set(_cabi.SIGNATURES)
    assert declared - bound == set(), sorted(declared - bound)
    assert bound - declared == set(), sorted(bound - declared)


def test_no_cpu_fallback_in_the_binding():
    """The product must fail loudly off-GPU: ptr() rejects host tensors."""
    import pytest
    import torch
    from gifgan import _cabi
    with pytest.raises(Exception):
        _cabi.ptr(torch.zeros(4))
