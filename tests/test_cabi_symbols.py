"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/xb200.h
declares, the Python binding declares a signature for each of them, and the product fails loudly without a GPU."""
import os
import subprocess

import pytest

import xerus_b200 as xb
from xerus_b200 import _lib


def test_library_is_built_in_tree():
    assert os.path.exists(xb.LIB_PATH), "run `python -m xerus_b200.build` (or __graft_entry__.build())"
    assert os.path.dirname(xb.LIB_PATH).endswith("xerus_b200")


def test_every_declared_symbol_is_exported():
    syms = xb.declared_symbols()
    assert len(syms) >= 55
    L = xb.lib()
    for s in syms:
        assert hasattr(L, s), "include/xb200.h declares %s but libxb200.so does not export it" % s
    nm = subprocess.run(["nm", "-D", "--defined-only", xb.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if " T " in line}
    assert set(syms) <= exported


def test_binding_covers_header():
    declared = set(xb.declared_symbols())
    bound = set(_lib._SIGS) | {"xb_last_error", "xb_version"}
    assert declared == bound, (declared - bound, bound - declared)


def test_no_torch_types_in_the_abi():
    text = open(_lib.HEADER).read()
    assert "torch" not in text and "at::" not in text and "#include <cuda" not in text


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "xerus_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "tt_oracle" not in src and "oracle/" not in src and "import oracle" not in src, f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(xb.XerusError) as e:
        xb.blasWrapper.two_norm([1.0, 2.0])
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_host_side_argument_checks():
    # REQUIRE-style failures that are decided on the host mirror (reference: ttNetwork.h:131-134, ttNetwork.cpp:648)
    with pytest.raises(xb.XerusError):
        xb.TTTensor.random([2, 2, 2], [2])
    with pytest.raises(xb.XerusError):
        xb.reshuffle([[1.0, 2.0]], [0, 0])
    assert xb.reduce_to_maximal_ranks([32] * 7, [4] * 8) == [4, 16, 32, 32, 32, 16, 4]
    assert xb.reduce_to_maximal_ranks([256] * 31, [2] * 32)[:9] == [2, 4, 8, 16, 32, 64, 128, 256, 256]
