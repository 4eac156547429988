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


def test_more_argument_errors_need_no_gpu(so_path):
    """Every entry point validates its arguments before touching CUDA (return < 0, message set)."""
    import ctypes as C
    from alphazero_quoridor_b200 import _lib, tree
    lib = _lib.load()
    p8 = C.c_void_p(64)
    # rollouts
    assert lib.qz_rollout(None, 0, None, 1, 0, 0, 0, None, 1000, None, None, None, None, 0, None) == 0      # empty
    assert lib.qz_rollout(None, 4, None, 1, 4, 0, 0, None, 1000, p8, None, None, p8, 0, None) == -1          # states NULL
    assert lib.qz_rollout(p8, 4, None, 1, 4, 0, 0, None, 0, p8, None, None, p8, 0, None) == -2               # limit < 1
    assert lib.qz_rollout(p8, 2, None, 1, 4, 0, 0, None, 1000, p8, None, None, p8, 0, None) == -2            # 2 states x 1 < 4
    assert lib.qz_rollout(C.c_void_p(68), 4, None, 1, 4, 0, 0, None, 1000, p8, None, None, p8, 0, None) == -3
    assert lib.qz_rollout_workspace_bytes(1000) >= 64 + 1000 * 24 + 4000
    assert lib.qz_env_random_play(None, 0, None, 3000, 5, None) == -1
    assert lib.qz_env_random_play(p8, 0, None, 70000, 5, None) == -2
    assert lib.qz_env_sample_legal(p8, None, 0, None, p8, 5, None) == -1
    # trees
    t = tree.QzTree()
    assert lib.qz_mcts_init(None, None, None, None) == -1
    assert lib.qz_mcts_select(C.byref(t), 5.0, 0, 1, 0, None, None) == -2                                      # zero dimensions
    t.n_games, t.node_cap, t.max_depth, t.leaves_per_game = 4, 100, 16, 2
    assert lib.qz_mcts_select(C.byref(t), 5.0, 0, 1, 0, None, None) == -2                                      # node_cap < one full block
    t.node_cap = 1000
    assert lib.qz_mcts_select(C.byref(t), 5.0, 0, 1, 0, None, None) == -1                                      # arrays NULL
    assert b"NULL" in lib.qz_last_error_string()
    assert lib.qz_stub_eval(p8, p8, 9, p8, p8, 4, None) == -2                                                # unknown stub kind
    assert lib.qz_device_sm_count(None) == -1
