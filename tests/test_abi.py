"""CPU-side contract tests: the C-ABI library loads and exports every symbol
the header declares; the host-side logic (Philox permutation, batch sharding)
matches the oracle.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import philox as px

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "scdbm_b200.h")).read()
    return sorted(set(re.findall(r"\b(scdbm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/scdbm_b200.h but not exported"
    assert set(pkg.PROTOTYPES) | {"scdbm_last_error"} == set(syms)


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "scdbm_b200.h")).read()
    assert len(re.findall(r"src/[a-z]+\.jl:\d+", text)) >= 20


def test_no_gpu_fails_loudly(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.ScdbmError) as e:
        pkg.Context(0, 1)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pk = os.path.join(ROOT, "scdbm.jl_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_host_permutation_matches_oracle(pkg):
    for n, seed, tick in [(1, 0, 0), (7, 5, 1), (500, 20261019, 3), (1001, 2 ** 40 + 17, 9)]:
        a = pkg.hostrng.permutation(n, seed, tick)
        b = px.Rng(seed).permutation(n, tick)
        assert np.array_equal(a, b) and sorted(a) == list(range(n))


def test_mostevenbatches(pkg):
    """src/misc.jl:157-169"""
    assert pkg.mostevenbatches(10, 3) == [4, 3, 3]
    assert pkg.mostevenbatches(1_000_000, 8) == [125_000] * 8
    assert sum(pkg.mostevenbatches(65_537, 8)) == 65_537
