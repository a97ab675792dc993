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
               6: ctypes.sizeof(capi.SynthParams), 7: ctypes.sizeof(capi.CutResult),
               8: ctypes.sizeof(capi.CutPlanes)}
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


@pytest.mark.parametrize("k", [0, 65, 100])
def test_kmer_length_outside_1_to_64_is_refused_with_the_documented_error(k):
    """Known difference from the reference, which accepts any -k (panfeed.py:59-67 slice strings):
    this build packs a k-mer into one or two 64-bit words and refuses k outside 1..64 with
    PF_ERR_UNSUPPORTED (-4) before touching the device; the CLI exits with the same message
    (README "Known differences", `panfeed --help`)."""
    with pytest.raises(capi.PfError) as e:
        capi.Context(k, 8)
    assert e.value.code == -4 and "1..64" in str(e.value)


def test_cli_refuses_long_kmers(tmp_path, capsys):
    from panfeed_b200.__main__ import get_options, main
    _parser_of(get_options)                       # --help names the limit
    with pytest.raises(SystemExit) as e:
        main(["-g", "fixture/gffs/", "-p", "fixture/gene_presence_absence.csv", "-k", "65",
              "-o", str(tmp_path / "out")])
    assert e.value.code == 1
    assert not (tmp_path / "out").exists()


def _parser_of(get_options):
    """The argparse parser get_options builds (its --help text names the k limit)."""
    import argparse
    made = []
    orig = argparse.ArgumentParser.parse_args

    def grab(self, argv=None):
        made.append(self)
        raise SystemExit(0)
    argparse.ArgumentParser.parse_args = grab
    try:
        try:
            get_options([])
        except SystemExit:
            pass
    finally:
        argparse.ArgumentParser.parse_args = orig
    assert "1..64" in made[0].format_help()
    return made[0]
