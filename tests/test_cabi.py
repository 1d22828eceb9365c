"""The C-ABI library loads and exports every symbol include/defer_b200.h declares (no GPU needed)."""
import ctypes
import re
from pathlib import Path

import pytest

from defer_b200 import _cabi as A

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "defer_b200.h").read_text()
    return sorted(set(re.findall(r"DEFER_API\s+[\w\s\*]+?\b(defer_\w+)\s*\(", text)))


def test_library_builds_and_loads():
    lib = A.load()
    assert lib.defer_abi_version() == A.ABI_VERSION
    assert isinstance(lib.defer_last_error(), bytes)


def test_every_declared_symbol_is_exported_and_bound():
    lib = A.load()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in A.PROTOTYPES, f"{name} has no ctypes prototype"
    assert sorted(A.PROTOTYPES) == declared


def test_struct_layouts_match_header():
    assert ctypes.sizeof(A.BufDesc) == 16
    assert ctypes.sizeof(A.OpDesc) == 17 * 4
    assert ctypes.sizeof(A.StageConfig) == 12 * 4


def test_errors_are_codes_not_crashes():
    lib = A.load()
    n = ctypes.c_int(-1)
    rc = lib.defer_device_count(ctypes.byref(n))
    if rc != A.OK:    # CPU box: the product must fail loudly, there is no CPU fallback
        assert rc == A.ERR_CUDA and b"cuda" in lib.defer_last_error().lower()
        with pytest.raises(A.DeferError):
            A.device_count()
    assert lib.defer_stage_step(None, 0) == A.ERR_INVALID
    assert b"null" in lib.defer_last_error()


def test_product_never_imports_oracle():
    for p in (ROOT / "defer_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
