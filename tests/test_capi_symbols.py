"""CPU-only: libqzb200.so loads and exports exactly the symbols include/qzb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qzb200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qz_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def so_path():
    from alphazero_quoridor_b200 import build
    return build.build()


def test_header_declares_something():
    syms = declared_symbols()
    assert "qz_env_step" in syms and "qz_env_legal_mask" in syms and "qz_rollout" in syms


def test_library_exports_every_declared_symbol(so_path):
    lib = ctypes.CDLL(so_path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libqzb200.so does not export %s" % name


def test_shim_signatures_cover_header(so_path):
    from alphazero_quoridor_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.qz_version() == _lib.ABI_VERSION


def test_argument_errors_need_no_gpu(so_path):
    """Argument validation happens before any CUDA call, so it is testable on a CPU box."""
    from alphazero_quoridor_b200 import _lib
    lib = _lib.load()
    assert lib.qz_env_reset(None, 0, None) == 0                 # empty batch is a no-op
    assert lib.qz_env_reset(None, 4, None) == -1                # QZ_E_NULL
    assert b"NULL" in lib.qz_last_error_string()
    assert lib.qz_env_reset(ctypes.c_void_p(12), 4, None) == -3   # QZ_E_ALIGN
    assert lib.qz_env_encode(ctypes.c_void_p(8), ctypes.c_void_p(8), 7, 0, 26, 1, None) == -2   # bad dtype
    assert lib.qz_env_legal_mask(None, None, -1, None) == -2


def test_no_cpu_fallback_without_gpu(so_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from alphazero_quoridor_b200 import _lib
    from alphazero_quoridor_b200.quoridor import BatchedQuoridor, Quoridor
    with pytest.raises(_lib.QzError):
        BatchedQuoridor(4)
    g = Quoridor()            # pure host attributes, constructing is fine
    with pytest.raises(_lib.QzError):
        g.actions()           # ... but any rules query needs the device
