"""CPU: the C-ABI library loads and exports every symbol include/speecht5_b200.h declares; host-side plumbing."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from speecht5_b200.build import build
    build()
    from speecht5_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "speecht5_b200.h")).read()
    declared = set(re.findall(r"\b(st5_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"st5_gemm_args", "st5_attn_args"}
    assert len(declared) >= 19
    from speecht5_b200._lib import EXPORTS
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_error_text(lib):
    assert lib.st5_version() >= 100
    assert isinstance(lib.st5_last_error(), bytes)


def test_struct_layout_matches_header():
    from speecht5_b200._lib import AttnArgs, GemmArgs
    # C layout computed by hand from the header: ints first, then 8-byte aligned pointers / int64
    assert ctypes.sizeof(GemmArgs) == 5 * 4 + 6 * 4 + 4 + 8 * 16 + 4 + 4 + 16 + 8 + 8
    assert ctypes.sizeof(AttnArgs) % 8 == 0 and AttnArgs.q.offset == 32


def test_product_path_refuses_cpu_tensors():
    from speecht5_b200 import kernels as K
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.gemm(a, a, torch.zeros(8, 8), M=8, N=8, K=8)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from speecht5_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
