"""The C-ABI shared library loads on a GPU-less box and exports every symbol the header declares.
No compute call is made here."""

import os
import re

from torchstore_b200 import _native
from torchstore_b200._build import LIB_PATH, REPO_ROOT


def header_functions():
    text = open(os.path.join(REPO_ROOT, "include", "tstore_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsb_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(LIB_PATH), "run __graft_entry__.build()"


def test_every_declared_symbol_is_exported_and_bound():
    lib = _native.lib()
    names = header_functions()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), f"{name} declared in tstore_b200.h but not exported"
    assert set(names) == set(_native.EXPORTED_SYMBOLS), set(names) ^ set(_native.EXPORTED_SYMBOLS)


def test_abi_version_and_struct_sizes():
    import ctypes as C

    assert _native.lib().tsb_abi_version() == _native.TSB_ABI_VERSION
    assert C.sizeof(_native.Region) == 64 + 8 * 4 + 4 * 2 + 8 + 8
    assert C.sizeof(_native.Rect) == 16 + 3 * 6 * 8 + 16
    assert C.sizeof(_native.PlanInfo) == 7 * 8 + 6 * 4


def test_no_gpu_means_loud_failure_not_fallback():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.TsbError):
        _native.init()
    with pytest.raises(_native.TsbError):
        _native.plan_create(0, _native.make_rect_array(1), 0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback when the CUDA extension is absent: first use raises with build instructions."""
    import pytest

    monkeypatch.setenv("TSTORE_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_native, "_lib", None)
    with pytest.raises(RuntimeError, match="libtstore_b200.so not found"):
        _native.lib()
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _native.init()


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under torchstore_b200/ may import it."""
    pkg = os.path.join(REPO_ROOT, "torchstore_b200")
    offenders = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
                    offenders.append(os.path.join(root, f))
    assert offenders == []
