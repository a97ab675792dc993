"""The drop-in CLI (panfeed_b200.__main__) on the toy pangenome: its three files
must equal the unmodified reference's after sorting, in every argument
combination of tests/unit_test.sh plus the extra modes."""
import gzip
import os

import pytest

import helpers

pytestmark = pytest.mark.gpu


def _read(path):
    if os.path.exists(path + ".gz"):
        return gzip.open(path + ".gz", "rt").read()
    return open(path).read()


@pytest.mark.parametrize("feeder", ["python", "native"])
@pytest.mark.parametrize("mode", sorted(helpers.modes()))
def test_cli_outputs_match_reference(mode, feeder, tmp_path):
    """feeder = native: GFF/FASTA parsing and cluster cutting by the library (the default);
    python: the loops that mirror the reference's input.py (--python-feeder)."""
    from panfeed_b200.__main__ import main
    args = [a for a in helpers.modes()[mode]] + (["--native-feeder"] if feeder == "native" else ["--python-feeder"])
    out = str(tmp_path / "out")
    cwd = os.getcwd()
    os.chdir(helpers.GOLDEN)
    try:
        main(args + ["--output", out])
    finally:
        os.chdir(cwd)
    for name in helpers.FILES:
        got = _read(os.path.join(out, name))
        want = helpers.golden(mode, name)
        assert got.split("\n")[0] == [x for x in want.split("\n") if x.startswith(
            ("cluster\t", "hashed_pattern"))][0]
        assert helpers.sorted_lines(got) == helpers.sorted_lines(want), (mode, name)
    if mode == "compress":
        assert os.path.exists(os.path.join(out, "kmers.tsv.gz"))


def test_cli_multiple_files(tmp_path):
    """--multiple-files: one directory per cluster, pattern set reset per cluster
    (panfeed.py:35-43,153-167).  Checked against the oracle port per cluster."""
    import numpy as np
    import pandas as pd
    from oracle import ref_port
    from panfeed_b200.__main__ import main
    out = str(tmp_path / "out")
    cwd = os.getcwd()
    os.chdir(helpers.GOLDEN)
    try:
        main(["--gff", "fixture/gffs/", "--presence-absence", "fixture/gene_presence_absence.csv",
              "--targets", "fixture/stroi.txt", "--multiple-files", "--output", out])
        table = pd.read_csv("fixture/gene_presence_absence.csv", sep=",", index_col=0,
                            low_memory=False).drop(columns=["Non-unique Gene name", "Annotation"])
        genomes = ref_port.load_inputs("fixture/gffs/")
        stroi = {x.rstrip("\n") for x in open("fixture/stroi.txt")}
        h2p_head, k2h_head = ref_port.headers(table.columns)
        for item in ref_port.feed_clusters(table, genomes, 0, 0, False):
            res = ref_port.kmer_stage(item, 31, stroi, True, False)
            a, b, c = ref_port.pattern_stage((res,), True, 0.01, False, set())
            d = os.path.join(out, item[1])
            assert helpers.sorted_lines(open(os.path.join(d, "kmers.tsv")).read()) == \
                helpers.sorted_lines(ref_port.KMERS_HEADER + a)
            assert helpers.sorted_lines(open(os.path.join(d, "hashes_to_patterns.tsv")).read()) == \
                helpers.sorted_lines(h2p_head + b)
            assert helpers.sorted_lines(open(os.path.join(d, "kmers_to_hashes.tsv")).read()) == \
                helpers.sorted_lines(k2h_head + c)
    finally:
        os.chdir(cwd)
