"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os
import subprocess

from gnnepcsaft_b200 import _lib


def test_build_and_symbols():
    import __graft_entry__
    __graft_entry__.build()
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.exported_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), name
    assert _lib.load().gnn_abi_version() == 1
    assert b"unsupported" in _lib.load().gnn_error_string(-3)


def test_library_is_sm100a_and_free_of_torch():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in ldd and "libc10" not in ldd


def test_argument_errors_without_gpu():
    lib = _lib.load()
    assert lib.gnn_linear_fwd(None, 0, None, 0, None, -1, 0, 0, 0, None, 0, None) == -1
    assert lib.gnn_pna_aggregate_fwd(None, None, None, None, None, 5, 2, 3, None, None) == -1
    assert lib.gnn_build_csr_workspace_bytes(1000, 2000) > 0
