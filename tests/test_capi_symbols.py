"""CPU-side checks of the C-ABI library: it loads, exports every symbol the
header declares, and its pure host helpers agree with the reference's float
expression.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from panfeed_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "panfeed_b200.h")).read()
    declared = set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", header))
    declared -= {"pf_pattern_words(S", "pf_kmer_pattern_words("}
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(capi.EXPORTS) == declared


def test_struct_sizes_match_header():
    assert capi.SEQ_DTYPE.itemsize == 48
    assert ctypes.sizeof(capi.Params) == 48
    assert ctypes.sizeof(capi.Batch) == 72


def test_ctypes_mirrors_match_the_library():
    """Every struct the binding mirrors has the size the library was compiled with."""
    lib = ctypes.CDLL(capi.LIB_PATH)
    lib.pf_struct_size.restype = ctypes.c_uint32
    lib.pf_struct_size.argtypes = [ctypes.c_int]
    mirrors = {0: ctypes.sizeof(capi.Params), 1: capi.SEQ_DTYPE.itemsize,
               2: capi.CLUSTER_DTYPE.itemsize, 3: ctypes.sizeof(capi.Batch),
               4: ctypes.sizeof(capi.BatchResult), 5: ctypes.sizeof(capi.Stats),
               6: ctypes.sizeof(capi.SynthParams), 7: ctypes.sizeof(capi.CutResult)}
    for which, size in mirrors.items():
        assert lib.pf_struct_size(which) == size, which
    assert lib.pf_struct_size(99) == 0


def test_maf_window_known_answers():
    for kat in helpers.hot_kats()["maf_windows"]:
        got = capi.maf_window(kat["maf"], kat["n"])
        assert kat["contiguous"]
        if kat["lo"] is None:
            assert got is None
        else:
            assert got == (kat["lo"], kat["hi"]), kat


def test_maf_window_matches_float_expression_exhaustively():
    rng = np.random.default_rng(0)
    for n in list(range(1, 70)) + [int(x) for x in rng.integers(70, 5000, 40)]:
        for maf in (0.0, 0.01, 0.05, 0.1, 0.123, 0.25, 0.3333, 0.49, 0.5):
            keep = []
            for c in range(n + 1):
                af = np.float64(c) / n
                if af >= 0.5:
                    af = 1 - af
                if not (af < maf):
                    keep.append(c)
            got = capi.maf_window(maf, n)
            if not keep:
                assert got is None
            else:
                assert keep == list(range(keep[0], keep[-1] + 1))
                assert got == (keep[0], keep[-1]), (n, maf)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PfError):
        capi.Context(31, 8)
