"""BASELINE.json config #2 at FULL size (500 genomes x 4,000 clusters, 1.9e9 bases)
on one B200, checked through size-independent properties and an exact oracle
comparison on a sample of clusters."""
import numpy as np
import pytest

from panfeed_b200 import capi, packer
from oracle import oracle_c

pytestmark = pytest.mark.gpu

S, C, L, K = 500, 4000, 1200, 31


def _checksum(r, W):
    """Order-independent checksum of the (cluster, k-mer, count, bitset) rows."""
    pat = r["new_kmer_patterns"][r["row_pattern"].astype(np.int64)][:, :W]
    h = (r["row_cluster"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ r["row_kmer"]
    h ^= r["row_count"].astype(np.uint64) << np.uint64(40)
    w = np.arange(1, W + 1, dtype=np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
    h = h + (pat.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)
    h = (h ^ (h >> np.uint64(31))) * np.uint64(0xBF58476D1CE4E5B9)
    return int(np.bitwise_xor.reduce(h)), int(h.sum(dtype=np.uint64))


@pytest.fixture(scope="module")
def full_run():
    hb = capi.synth_batch(0, 20261018 + 2, S, C, total_clusters=C, gene_len=L)
    ctx = capi.Context(K, S, maf=0.01)
    ctx.submit(hb)
    r = ctx.collect()
    st = ctx.stats()
    ctx.close()
    return hb, r, st


def test_full_size_properties(full_run):
    hb, r, st = full_run
    W = (S + 31) // 32
    assert st["bases"] == hb.n_bases and st["bases"] > 1.8e9
    assert st["instances"] == int(np.maximum(hb.seqs["len"].astype(np.int64) - K + 1, 0).sum())
    n = len(r["row_cluster"])
    assert n == st["rows"] and n > 1e7
    pat = r["new_kmer_patterns"]
    # patterns are unique (the table is keyed on the full bitset)
    assert len(np.unique(pat, axis=0)) == len(pat) == st["kmer_patterns"]
    # every pattern is referenced, ids are dense
    assert np.array_equal(np.unique(r["row_pattern"]), np.arange(len(pat)))
    # count == popcount(bitset), inside the MAF window [5, 495] of n = 500, maf = 0.01
    pc = np.unpackbits(pat.view(np.uint8), axis=1).sum(axis=1)
    assert np.array_equal(pc[r["row_pattern"].astype(np.int64)], r["row_count"])
    lo, hi = capi.maf_window(0.01, S)
    assert (lo, hi) == (5, 495)
    assert r["row_count"].min() >= lo and r["row_count"].max() <= hi
    # a k-mer's samples are a subset of its cluster's presence
    bits = pat[r["row_pattern"].astype(np.int64)][:, :W]
    pres = hb.presence[r["row_cluster"].astype(np.int64)]
    assert not np.any(bits & ~pres)
    # (cluster, k-mer) rows are unique, k-mers canonical and < 4^k
    key = r["row_cluster"].astype(np.uint64) << np.uint64(40)
    assert len(np.unique(np.stack([r["row_cluster"].astype(np.uint64), r["row_kmer"]], axis=1), axis=0)) == n
    assert int(r["row_kmer"].max()) < 4 ** K
    # cluster rows: one pattern id per cluster, pointing at the cluster's own presence bits
    assert np.array_equal(r["new_cluster_patterns"][r["cluster_pattern"].astype(np.int64)], hb.presence)


@pytest.mark.parametrize("engine", ["records", "fullsort"])
def test_full_size_engines_agree(full_run, engine):
    """The block-aggregation engine (default), the record path in partition mode (fused,
    unstable, hashed) and the plain full-sort engine produce the same multiset of rows
    (checksum of checksums)."""
    hb, r, st = full_run
    assert st["engine"] == 2
    W = (S + 31) // 32
    ctx = capi.Context(K, S, maf=0.01, mode=1 if engine == "fullsort" else 0, debug_flags=2)
    ctx.submit(hb)
    r2 = ctx.collect()
    st2 = ctx.stats()
    ctx.close()
    assert st2["engine"] == (1 if engine == "fullsort" else 0)
    assert st2["rows"] == st["rows"] and st2["kmer_patterns"] == st["kmer_patterns"]
    if engine == "records":      # (the full-sort engine reports prefix-runs, a lower bound)
        assert st2["unique_kmers"] == st["unique_kmers"]
    assert _checksum(r, W) == _checksum(r2, W)


def test_full_size_sample_of_clusters_matches_oracle(full_run):
    hb, r, st = full_run
    W = (S + 31) // 32
    rng = np.random.default_rng(1)
    first = np.searchsorted(hb.seqs["cluster"], np.arange(C + 1))
    idx = np.arange(S)
    sh = (62 - 2 * np.arange(32)).astype(np.uint64)
    order = np.argsort(r["row_cluster"], kind="stable")
    bounds = np.searchsorted(r["row_cluster"][order], np.arange(C + 1))
    for c in sorted(rng.choice(C, 24, replace=False).tolist()):
        seqs = hb.seqs[first[c]:first[c + 1]]
        if len(seqs) == 0:
            continue
        w0 = int(seqs["base_off"][0]) // 32
        w1 = int(seqs["base_off"][-1] + (seqs["len"][-1] + 63) // 64 * 64) // 32
        codes = ((hb.packed[w0:w1, None] >> sh[None, :]) & np.uint64(3)).astype(np.uint8).ravel()
        ascii_plane = np.frombuffer(b"ACGT", np.uint8)[codes]
        o = np.zeros(len(seqs), oracle_c.SEQ_DTYPE)
        for f in ("len", "sample", "start", "end", "offset", "strand"):
            o[f] = seqs[f]
        o["off"] = seqs["base_off"] - np.uint64(w0 * 32)
        presab = ((hb.presence[c:c + 1, idx >> 5] >> (idx & 31)) & 1).astype(np.uint8)
        want = oracle_c.run_arrays(ascii_plane, o, presab, K, True, False, False, 0.01, n_threads=4)
        rows = order[bounds[c]:bounds[c + 1]]
        got = sorted(zip(packer.kmers_to_str(r["row_kmer"][rows], K).tolist(), r["row_count"][rows].tolist(),
                         [r["new_kmer_patterns"][p][:W].tobytes() for p in r["row_pattern"][rows]]))
        exp = sorted(zip(want["row_kmer"].tolist(), want["row_count"].tolist(),
                         [want["kmer_pattern_bits"][p].tobytes() for p in want["row_pattern"]]))
        assert got == exp, f"cluster {c}"


@pytest.mark.parametrize("samples,clusters,cm", [(10000, 48, False), (50000, 8, True)])
def test_sample_sliced_engine_agrees_with_records(samples, clusters, cm):
    """BASELINE configs #4 / #5 sample counts (10,000 / 50,000 genomes, the latter with the
    cluster-absent encoding): the block engine in 512-sample slices and the record engine
    produce the same multiset of (cluster, k-mer, count, bitset) rows and the same patterns."""
    hb = capi.synth_batch(0, 20261018 + 4, samples, clusters, total_clusters=clusters, gene_len=L)
    W = (samples + 31) // 32
    res = {}
    for name, flags in (("block", 0), ("records", 2)):
        ctx = capi.Context(K, samples, consider_missing=cm, maf=0.01, debug_flags=flags)
        ctx.submit(hb)
        r = ctx.collect()
        st = ctx.stats()
        ctx.close()
        res[name] = (st, _checksum(r, W), len(r["new_kmer_patterns"]))
    assert res["block"][0]["engine"] == 2 and res["records"][0]["engine"] == 0
    assert res["block"][0]["rows"] == res["records"][0]["rows"] > 10000
    assert res["block"][0]["unique_kmers"] == res["records"][0]["unique_kmers"]
    assert res["block"][1] == res["records"][1]
    assert res["block"][2] == res["records"][2]
