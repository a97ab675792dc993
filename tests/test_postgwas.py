"""panfeed_b200.postgwas (get-clusters / get-kmers with the native TSV row filter) against the
stdout of the unmodified reference tools (tests/golden/make_golden_postgwas.py), and
pf_tsv_filter against a plain Python scan.  Pure host code: runs without a GPU."""
import gzip
import json
import os

import numpy as np
import pytest

from panfeed_b200 import capi, postgwas

import helpers

POST = os.path.join(helpers.EXPECTED, "postgwas")
CASES = json.load(open(os.path.join(POST, "cases.json")))


@pytest.fixture(scope="module")
def outputs(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("postgwas")
    for name in helpers.FILES:
        (tmp / name).write_text(helpers.golden(CASES["mode"], name))
    return tmp


@pytest.mark.parametrize("case", sorted(CASES["cases"]))
def test_postgwas_tools_match_reference_stdout(case, outputs, capsys):
    tool, extra = CASES["cases"][case]
    args = ["-a", os.path.join(POST, "associations.tsv"), "-p", str(outputs / "kmers_to_hashes.tsv")]
    if tool == "get_kmers":
        args += ["-k", str(outputs / "kmers.tsv")]
        postgwas.get_kmers_main(args + extra)
    else:
        postgwas.get_clusters_main(args + extra)
    got = capsys.readouterr().out
    want = gzip.open(os.path.join(POST, case + ".txt.gz"), "rt").read()
    # clusters come out of a Python set in both: any order; the header of get-kmers is printed once, first
    assert sorted(got.split("\n")) == sorted(want.split("\n"))
    assert got.split("\n")[0] == want.split("\n")[0] or tool == "get_clusters"
    assert len(got) == len(want)


def test_postgwas_gz_inputs_take_the_pandas_route(outputs, capsys, tmp_path):
    gz = tmp_path / "kmers_to_hashes.tsv.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write((outputs / "kmers_to_hashes.tsv").read_text())
    postgwas.get_clusters_main(["-a", os.path.join(POST, "associations.tsv"), "-p", str(gz), "--threshold", "0.3"])
    got = capsys.readouterr().out
    want = gzip.open(os.path.join(POST, "clusters_t0.3.txt.gz"), "rt").read()
    assert sorted(got.split("\n")) == sorted(want.split("\n"))


@pytest.mark.parametrize("threads", [1, 4])
def test_tsv_filter_matches_python_scan(threads, tmp_path):
    """Rows in file order; header excluded; CRLF and a last line without newline; keys that are
    prefixes of fields or sit in other columns must not match; multi-threaded pieces (> 4 MiB each)."""
    rng = np.random.default_rng(5)
    keys = [("h%05d" % i) for i in rng.choice(100000, 300, replace=False)]
    lines = ["cluster\tk-mer\thashed_pattern"]
    n = 600_000 if threads > 1 else 5_000
    hs = rng.integers(0, 100000, n)
    cs = rng.integers(0, 50, n)
    for i in range(n):
        lines.append("cl%d\t%s\th%05d" % (cs[i], "ACGT" if i % 7 else "", hs[i]))
    lines[10] = "h%s\tACGT\tx" % keys[0][1:]                    # a key in another column
    lines[11] = "cl1\tACGT\t%s0" % keys[1]                      # a key as a prefix
    lines[12] = "cl1\tACGT\t%s\r" % keys[2]                     # CRLF
    lines[13] = ""                                              # blank line
    lines[14] = "cl1\tACGT"                                     # short row
    path = tmp_path / "t.tsv"
    path.write_text("\n".join(lines[:-1]) + "\n" + "cl9\tAAAA\t" + keys[3])      # no trailing newline
    lines[-1] = "cl9\tAAAA\t" + keys[3]
    ks = set(keys)
    want = [ln.rstrip("\r") for ln in lines[1:] if len(ln.rstrip("\r").split("\t")) > 2 and
            ln.rstrip("\r").split("\t")[2] in ks]
    body, rows = capi.tsv_filter(str(path), 2, keys, True, threads)
    assert rows == len(want)
    assert body.decode().split("\n")[:-1] == want
    # another column, no header skipping, empty key set, missing file
    body, rows = capi.tsv_filter(str(path), 0, ["cl9", "cluster"], False, threads)
    assert body.decode().split("\n")[0] == lines[0] and rows == 1 + sum(ln.startswith("cl9\t") for ln in lines)
    assert capi.tsv_filter(str(path), 2, [], True, threads) == (b"", 0)
    with pytest.raises(capi.PfError):
        capi.tsv_filter(str(tmp_path / "nope.tsv"), 0, ["x"])
