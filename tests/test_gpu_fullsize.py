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


def _oracle_cluster(hb, c, first, samples, cm):
    """C oracle on cluster c of a synthetic HostBatch -> sorted (k-mer, count, bitset bytes) rows."""
    W = (samples + 31) // 32
    sh = (62 - 2 * np.arange(32)).astype(np.uint64)
    idx = np.arange(samples)
    seqs = hb.seqs[first[c]:first[c + 1]]
    w0 = int(seqs["base_off"][0]) // 32
    w1 = int(seqs["base_off"][-1] + (seqs["len"][-1] + 63) // 64 * 64) // 32
    codes = ((hb.packed[w0:w1, None] >> sh[None, :]) & np.uint64(3)).astype(np.uint8).ravel()
    ascii_plane = np.frombuffer(b"ACGT", np.uint8)[codes]
    o = np.zeros(len(seqs), oracle_c.SEQ_DTYPE)
    for f in ("len", "sample", "start", "end", "offset", "strand"):
        o[f] = seqs[f]
    o["off"] = seqs["base_off"] - np.uint64(w0 * 32)
    presab = ((hb.presence[c:c + 1, idx >> 5] >> (idx & 31)) & 1).astype(np.uint8)
    want = oracle_c.run_arrays(ascii_plane, o, presab, K, True, cm, False, 0.01, n_threads=8)
    return sorted(zip(want["row_kmer"].tolist(), want["row_count"].tolist(),
                      [want["kmer_pattern_bits"][p][:W].tobytes() for p in want["row_pattern"]]))


@pytest.mark.parametrize("samples,clusters,cm,check", [(10000, 30, False, (3, 17)), (50000, 6, True, (4,))])
def test_configs_4_and_5_full_length_clusters_match_oracle(samples, clusters, cm, check):
    """BASELINE configs #4 / #5 at their real shape - 10,000 / 50,000 genomes, 1.2-kb clusters,
    the latter with the cluster-absent encoding - through the pipelined submit (several
    sub-batches): sampled clusters must equal the C oracle row by row (k-mer, sample count, full
    bitset), and the whole batch must satisfy the size-independent properties."""
    hb = capi.synth_batch(0, 20261018 + (5 if cm else 4), samples, clusters, total_clusters=clusters, gene_len=L)
    W = (samples + 31) // 32
    ctx = capi.Context(K, samples, consider_missing=cm, maf=0.01)
    ctx.submit(hb)
    r = ctx.collect()
    st = ctx.stats()
    ctx.close()
    assert st["engine"] == 2 and st["sub_batches"] >= 2
    assert st["instances"] == int(np.maximum(hb.seqs["len"].astype(np.int64) - K + 1, 0).sum())
    pat = r["new_kmer_patterns"]
    assert len(np.unique(pat, axis=0)) == len(pat) == st["kmer_patterns"]
    bits = pat[r["row_pattern"].astype(np.int64)][:, :W]
    pres = hb.presence[r["row_cluster"].astype(np.int64)]
    assert not np.any(bits & ~pres)
    pc = np.unpackbits(bits.view(np.uint8), axis=1).sum(axis=1)
    assert np.array_equal(pc, r["row_count"])
    if cm:      # the key's last word names the cluster pattern that gives the NaN plane
        cp = r["new_cluster_patterns"][pat[r["row_pattern"].astype(np.int64)][:, W].astype(np.int64)]
        assert np.array_equal(cp, pres)
    first = np.searchsorted(hb.seqs["cluster"], np.arange(clusters + 1))
    order = np.argsort(r["row_cluster"], kind="stable")
    bounds = np.searchsorted(r["row_cluster"][order], np.arange(clusters + 1))
    for c in check:
        rows = order[bounds[c]:bounds[c + 1]]
        got = sorted(zip(packer.kmers_to_str(r["row_kmer"][rows], K).tolist(), r["row_count"][rows].tolist(),
                         [r["new_kmer_patterns"][p][:W].tobytes() for p in r["row_pattern"][rows]]))
        assert got == _oracle_cluster(hb, c, first, samples, cm), f"cluster {c}"
        assert len(got) > 100


def test_config3_second_pass_positions_full_size():
    """BASELINE config #3 at full size (200 clusters x 500 genomes, every sample a --targets
    strain: 9.3e7 positional records).  The compact form (used_strand bit plane) and the 21-byte
    record form must agree on every record; the records of sampled sequences are checked field
    by field against a host computation from the packed bases (panfeed.py:64-107)."""
    n_cl = 200
    hb = capi.synth_batch(0, 20261018 + 3, S, n_cl, total_clusters=n_cl, gene_len=L, all_targets=True)
    out = {}
    for mode in (2, 1):
        ctx = capi.Context(K, S, emit_positions=mode, maf=0.01)
        ctx.submit(hb)
        out[mode] = (ctx.collect(), ctx.stats())
        ctx.close()
    (rc, stc), (rr, st) = out[2], out[1]
    nwin = np.maximum(hb.seqs["len"].astype(np.int64) - K + 1, 0)
    assert rc["n_pos"] == rr["n_pos"] == int(nwin.sum()) == st["instances"] > 9e7
    assert len(rc["pos_seq"]) == 0 and len(rc["pos_strand_bits"]) == len(hb.packed)
    assert stc["rows"] == st["rows"] and np.array_equal(np.sort(rc["row_kmer"]), np.sort(rr["row_kmer"]))
    # every record: sequence, position, coordinates, strand bit
    seq = rr["pos_seq"].astype(np.int64)
    assert np.array_equal(np.bincount(seq, minlength=len(hb.seqs)), nwin)
    q = hb.seqs
    p = rr["pos_gene_start"].astype(np.int64) + q["offset"][seq]
    assert p.min() >= 0 and np.all(p < nwin[seq])
    want_c0 = np.where(q["strand"][seq] > 0, q["start"][seq] + p, q["end"][seq] - p - K)
    assert np.array_equal(rr["pos_contig_start"], want_c0)
    i = q["base_off"][seq].astype(np.int64) + p
    bit = (rc["pos_strand_bits"][i >> 5] >> (i & 31).astype(np.uint32)) & 1
    assert np.array_equal(bit.astype(np.uint8), rr["pos_flags"] & 1)
    del seq, p, want_c0, i, bit
    # sampled sequences: canonical k-mer and strand of every window from the packed bases
    rng = np.random.default_rng(3)
    starts = np.zeros(len(hb.seqs) + 1, np.int64)
    np.cumsum(nwin, out=starts[1:])
    order = np.argsort(rr["pos_seq"], kind="stable")
    sh = (62 - 2 * np.arange(32)).astype(np.uint64)
    mask = (1 << (2 * K)) - 1
    for s in rng.choice(len(hb.seqs), 40, replace=False).tolist():
        w0 = int(q["base_off"][s]) // 32
        codes = ((hb.packed[w0:w0 + (int(q["len"][s]) + 31) // 32, None] >> sh[None, :]) & np.uint64(3)).ravel()
        codes = [int(x) for x in codes[:int(q["len"][s])]]
        rec = order[starts[s]:starts[s + 1]]
        assert np.all(rr["pos_seq"][rec] == s)
        got = {int(g) + int(q["offset"][s]): (int(km), int(fl)) for g, km, fl in
               zip(rr["pos_gene_start"][rec], rr["pos_kmer"][rec], rr["pos_flags"][rec])}
        fwd = rc_ = 0
        for j, b in enumerate(codes):
            fwd = ((fwd << 2) | b) & mask
            rc_ = (rc_ >> 2) | ((3 - b) << (2 * (K - 1)))
            if j >= K - 1:
                pos = j - K + 1
                assert got[pos] == ((rc_, 1) if rc_ < fwd else (fwd, 0)), (s, pos)
