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


# ------------------------------------------------ the Julia facade (cannot be executed here: no Julia toolchain)
_JL2C = {"Int32": ctypes.c_int32, "Int64": ctypes.c_int64, "UInt32": ctypes.c_uint32, "UInt64": ctypes.c_uint64,
         "Float64": ctypes.c_double, "Float32": ctypes.c_float, "UInt16": ctypes.c_uint16, "UInt8": ctypes.c_uint8}


def _julia_ccalls():
    src = open(os.path.join(ROOT, "julia", "scDBM.jl")).read()
    calls = re.findall(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(([^()]*)\)", src)
    return src, [(name, ret, [a.strip() for a in args.split(",") if a.strip()]) for name, ret, args in calls]


def _compatible(jl, ct):
    """Julia ccall argument type vs the ctypes argtype of the Python binding (both mirror include/scdbm_b200.h)."""
    if jl == "Cstring":
        return ct is ctypes.c_char_p
    m = re.fullmatch(r"(?:Ptr|Ref)\{(.+)\}", jl)
    if m:
        inner = m.group(1)
        if ct is ctypes.c_void_p:                       # opaque handles and void* take any pointer
            return True
        if inner == "Ptr{Cvoid}":                       # handle out-parameter
            return ct is ctypes.POINTER(ctypes.c_void_p)
        return inner in _JL2C and ct is ctypes.POINTER(_JL2C[inner])
    return jl in _JL2C and ct is _JL2C[jl]


def test_julia_facade_ccalls_match_the_abi(pkg):
    """Every `ccall((:sym, LIB), ret, (argtypes...))` of julia/scDBM.jl: the symbol is declared in the header and
    exported, the return type is Int32 (Cstring for scdbm_last_error), arity and every C type agree with
    `_lib.PROTOTYPES`."""
    src, calls = _julia_ccalls()
    assert len(calls) >= 55, f"only {len(calls)} ccalls parsed"
    assert src.count("ccall(") == len(calls), "a ccall in julia/scDBM.jl does not follow the parsable (:sym, LIB) form"
    syms = set(header_symbols())
    for name, ret, args in calls:
        assert name in syms, f"{name}: not declared in include/scdbm_b200.h"
        if name == "scdbm_last_error":
            assert ret == "Cstring" and args == []
            continue
        assert ret == "Int32", f"{name}: every entry point returns int32 status"
        proto = pkg.PROTOTYPES[name]
        assert len(args) == len(proto), f"{name}: {len(args)} arguments in the ccall, {len(proto)} in the ABI"
        for i, (jl, ct) in enumerate(zip(args, proto)):
            assert _compatible(jl, ct), f"{name}: argument {i} is {jl} in Julia but {ct} in the ABI"


def test_julia_facade_defines_what_it_exports():
    src, calls = _julia_ccalls()
    exports = re.search(r"^export (.*?)\n\n", src, re.S | re.M).group(1)
    names = [n.strip() for n in exports.replace("\n", " ").split(",") if n.strip()]
    assert len(names) >= 40
    for n in names:
        pat = (rf"(^|\n)\s*(function |mutable struct |struct |const |Base\.@kwdef struct |abstract type )?{re.escape(n)}(?![\w!])[ ({{=:<\n]"
               rf"|for f in \([^)]*:{re.escape(n)}[,)]")     # the allocating conditional operators are generated by @eval loops
        assert re.search(pat, src), f"`{n}` is exported but never defined in julia/scDBM.jl"
    assert "using Random" in src and "UNEXECUTED" in src
    # the entry points a drop-in user needs are all bound
    used = {c[0] for c in calls}
    for must in ("scdbm_rbm_train_epoch", "scdbm_dbm_pcd_step", "scdbm_dbm_pcd_step_dp", "scdbm_gibbs", "scdbm_gibbs_negbin_rbm",
                 "scdbm_gibbs_cond", "scdbm_sample_dropout", "scdbm_mle_theta", "scdbm_meanfield", "scdbm_ais_run",
                 "scdbm_comm_init", "scdbm_allreduce_gradient", "scdbm_mat_upload_compressed", "scdbm_mat_download_csr"):
        assert must in used, f"{must} is not bound by the Julia facade"
