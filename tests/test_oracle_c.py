"""Pins oracle/oracle.c (fast C restatement) against the goldens of the
unmodified reference and against oracle/ref_port.py."""
import os

import numpy as np
import pandas as pd
import pytest

import helpers
import render
from oracle import oracle_c, ref_port


def _items(kw):
    table = pd.read_csv(kw["presence_absence"], sep=",", index_col=0,
                        low_memory=False).drop(
                            columns=["Non-unique Gene name", "Annotation"])
    genomes = ref_port.load_inputs(kw["gff"], kw["fasta"])
    stroi = ({x.rstrip("\n") for x in open(kw["targets"])}
             if kw["targets"] else set())
    genes = ({x.rstrip("\n") for x in open(kw["genes"])}
             if kw["genes"] else None)
    items = list(ref_port.feed_clusters(table, genomes, kw["upstream"],
                                        kw["downstream"],
                                        kw["downstream_start_codon"], genes))
    return items, stroi, len(table.columns)


@pytest.mark.parametrize("mode", sorted(helpers.modes()))
@pytest.mark.parametrize("threads", [1, 3])
def test_c_oracle_matches_reference_goldens(mode, threads):
    if threads == 3 and mode not in ("basic", "cm_nofilter_up"):
        pytest.skip("threaded run checked on two modes")
    kw = helpers.cli_kwargs(helpers.modes()[mode])
    cwd = os.getcwd()
    os.chdir(helpers.GOLDEN)
    try:
        items, stroi, S = _items(kw)
    finally:
        os.chdir(cwd)
    out = oracle_c.run(items, stroi, kw["k"], canonical=not kw["non_canonical"],
                       consider_missing=kw["consider_missing"],
                       cluster_equal_filter=kw["no_filter"], maf=kw["maf"],
                       n_threads=threads)
    res = dict(out)
    res["row_kmer"] = [x.decode() for x in out["row_kmer"]]
    res["pos_kmer"] = [x.decode() for x in out["pos_kmer"]]
    got = render.render(res, out["ids"], out["seq_meta"], out["seqs"], S,
                        kw["k"], kw["consider_missing"],
                        not kw["non_canonical"])
    for name in helpers.FILES:
        assert sorted(got[name]) == render.golden_body(
            helpers.golden(mode, name)), (mode, name)


def test_c_oracle_random_clusters_vs_port():
    rng = np.random.default_rng(5)
    comp = str.maketrans("ACGTN", "TGCAN")
    for trial in range(6):
        S = int(rng.integers(2, 70))
        k = int(rng.choice([5, 12, 31, 32]))
        names = [f"g{i:03d}" for i in range(S)]
        order = list(rng.permutation(names))
        items = []
        for c in range(3):
            L = int(rng.integers(k, 90))
            anc = "".join(rng.choice(list("ACGT"), L))
            presab = np.zeros(S, dtype=int)
            cluster, absent = {}, []
            for s in order:
                if rng.random() < 0.3:
                    absent.append(s)
                    continue
                presab[names.index(s)] = 1
                q = "".join(ch if rng.random() > 0.05 else
                            rng.choice(list("ACGTN")) for ch in anc)
                cluster[s] = [ref_port.CutSeq(q, q.translate(comp), s + "_x",
                                              "c", 11, 10 + L,
                                              int(rng.choice([1, -1])), 3)]
            for s in absent:
                cluster[s] = []
            items.append((cluster, f"cl{c}", presab))
        canon = bool(trial % 2 == 0)
        cm = bool(trial % 3 == 0)
        nf = bool(trial % 2 == 1)
        stroi = set(order[:3])
        pats = set()
        want = {n: [] for n in helpers.FILES}
        for it in items:
            r = ref_port.kmer_stage(it, k, stroi, canon, cm)
            a, b, c = ref_port.pattern_stage((r,), not nf, 0.05, cm, pats)
            want["kmers.tsv"] += a.split("\n")[:-1]
            want["hashes_to_patterns.tsv"] += b.split("\n")[:-1]
            want["kmers_to_hashes.tsv"] += c.split("\n")[:-1]
        out = oracle_c.run(items, stroi, k, canon, cm, nf, 0.05)
        res = dict(out)
        res["row_kmer"] = [x.decode() for x in out["row_kmer"]]
        res["pos_kmer"] = [x.decode() for x in out["pos_kmer"]]
        got = render.render(res, out["ids"], out["seq_meta"], out["seqs"], S,
                            k, cm, canon)
        for name in helpers.FILES:
            assert sorted(got[name]) == sorted(want[name]), (trial, name)
