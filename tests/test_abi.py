"""CPU-side checks: the C-ABI library loads and exports every symbol include/rrl_b200.h declares; argument
validation works without a GPU (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rrl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rrl_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as ge
    ge.build()
    import rrl_b200
    return rrl_b200._native


def test_every_declared_symbol_is_exported(native):
    L = native.lib()
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(native.EXPORTED) == declared


def test_version_errors_and_workspace_sizing(native):
    L = native.lib()
    assert L.rrl_version() == 100
    assert L.rrl_error_string(0) == b"ok"
    assert b"workspace" in L.rrl_error_string(-2)
    assert L.rrl_workspace_bytes(0, 1, 1, 1) == 0 and L.rrl_workspace_bytes(1, 0, 1, 1) == 0
    small, big = L.rrl_workspace_bytes(1, 1024, 1024, 20000), L.rrl_workspace_bytes(32, 1024, 1024, 15000)
    assert 0 < small < big < 4 << 30
    # the radix-sort scratch of clouds above 4096 triplets covers every cloud of up to 512 pairs (one sort per batch):
    # it grows with the batch up to that cap and the whole workspace stays linear in triplets and lines
    one, eight = L.rrl_workspace_bytes(1, 8192, 8192, 1000), L.rrl_workspace_bytes(8, 8192, 8192, 1000)
    fixed = 4 << 20                                                           # CUB's temporary storage allowance
    assert 7 * (one - fixed) < eight - fixed < 9 * (one - fixed)
    large = L.rrl_workspace_bytes(1, 500000, 500000, 100000)
    assert 100 << 20 < large < 200 << 20
    assert L.rrl_workspace_bytes(1, 500000, 300, 100000) < large              # ragged pair: sized per cloud
    assert L.rrl_sampler_workspace_bytes(1, 20000, 10) >= 200000
    # null pointers / bad windows are rejected before anything touches a device
    assert L.rrl_loss_forward(None, None, None, 1, 8, 8, 8, 1, 1, 5, 5, None, 0, None, None, None, None, None) == -1
    assert L.rrl_se3_exp(None, 1, None, None, None) == -1
    assert L.rrl_sample_lines(None, None, None, None, 1, 1, 1, 1, 10, 0, 0, None, None, None, None, 0, None) == -1


def test_product_path_refuses_cpu_tensors(native):
    import torch
    import rrl_b200
    with pytest.raises(rrl_b200.NativeError):
        rrl_b200.intersected_line_loss(torch.zeros(1, 4, 9), torch.zeros(1, 4, 9), torch.zeros(1, 4, 6))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "a-robust-registration-loss_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/ ", ""), f
