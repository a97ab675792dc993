"""pf_format_positions (native, multi-threaded host code of the library) against a plain
Python rendering of the reference's f-strings (/root/reference/panfeed/panfeed.py:90-107).
Pure host code: runs without a GPU."""
import numpy as np
import pytest

from panfeed_b200 import capi

_COMP = bytes.maketrans(b"ACTGNYRWSKMDVHBX", b"TGACNRYWSMKHBDVX")


def _pack2(s):
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch)
    return v


def _pack4(s):
    v = 0
    for ch in s:
        v = (v << 4) | capi.AMB_ALPHABET.index(ch)
    return v >> 64, v & ((1 << 64) - 1)


def _records(rng, n, k, n_seqs, amb_rate):
    kmers, flags, wide, texts = [], [], [], []
    for _ in range(n):
        if rng.random() < amb_rate:
            t = "".join(rng.choice(list("ACGTNRYKMSWBDHVX"), k))
            hi, lo = _pack4(t)
            kmers.append(len(wide))
            wide.append((hi, lo))
            flags.append(2 | int(rng.integers(0, 2)))
        else:
            t = "".join(rng.choice(list("ACGT"), k))
            kmers.append(_pack2(t))
            flags.append(int(rng.integers(0, 2)))
        texts.append(t)
    r = {"pos_kmer": np.array(kmers, np.uint64), "pos_seq": rng.integers(0, n_seqs, n).astype(np.uint32),
         "pos_contig_start": rng.integers(-50, 5_000_000, n).astype(np.int32),
         "pos_gene_start": rng.integers(-300, 3000, n).astype(np.int32),
         "pos_flags": np.array(flags, np.uint8),
         "pos_wide_kmer": np.array(wide, np.uint64).reshape(-1, 2)}
    return r, texts


@pytest.mark.parametrize("k", [1, 15, 31, 32])
@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("threads", [1, 4])
def test_native_position_rows_match_python(k, canonical, threads):
    rng = np.random.default_rng(100 * k + canonical)
    n_seqs = 37
    leads = [f"cl{int(rng.integers(0, 9))}\tstrain_{i}\tgene{i}_x\tcontig{i % 5}\t{s}\t".encode()
             for i, s in enumerate(rng.choice([1, -1], n_seqs))]
    strand = np.array([int(x.split(b"\t")[4]) for x in leads], np.int32)
    n = 70_000 if threads > 1 else 3_000
    r, texts = _records(rng, n, k, n_seqs, 0.02)
    got = capi.format_positions(r, k, canonical, leads, strand, n_threads=threads)
    want = []
    for i in range(n):
        s = int(r["pos_seq"][i])
        c0, g0 = int(r["pos_contig_start"][i]), int(r["pos_gene_start"][i])
        head = leads[s].decode() + f"{c0}\t{c0 + k}\t{g0}\t{g0 + k}\t"
        if canonical:
            used = -1 if r["pos_flags"][i] & 1 else 1
            want.append(f"{head}{used}\t{texts[i]}\n")
        else:
            st = int(strand[s])
            rc = texts[i].encode().translate(_COMP)[::-1].decode()
            want.append(f"{head}{st}\t{texts[i]}\n{head}{-st}\t{rc}\n")
    assert got.decode() == "".join(want)


def test_native_position_rows_empty_and_bad_range():
    r = {"pos_kmer": np.zeros(0, np.uint64), "pos_seq": np.zeros(0, np.uint32),
         "pos_contig_start": np.zeros(0, np.int32), "pos_gene_start": np.zeros(0, np.int32),
         "pos_flags": np.zeros(0, np.uint8)}
    assert capi.format_positions(r, 31, True, [], []) == b""


@pytest.mark.parametrize("S", [1, 31, 32, 33, 500, 1300])
@pytest.mark.parametrize("with_nan", [False, True])
def test_native_pattern_rows_match_python(S, with_nan):
    """pf_format_patterns against the reference's join (panfeed.py:183-187,217-223)."""
    rng = np.random.default_rng(S + with_nan)
    n, W = 300, (S + 31) // 32
    bits = rng.integers(0, 2, (n, S)).astype(np.uint8)
    pres = rng.integers(0, 2, (n, S)).astype(np.uint8) if with_nan else np.ones((n, S), np.uint8)
    bits &= pres                                   # a k-mer's bits are a subset of the presence

    def words(m):
        pad = np.zeros((n, W * 32), np.uint8)
        pad[:, :S] = m
        return np.packbits(pad.reshape(n, W, 32), axis=2, bitorder="little").view(np.uint32).reshape(n, W)

    ids = ["%024d" % i for i in range(n)]
    got = capi.format_patterns(np.hstack([words(bits), np.full((n, 1), 7, np.uint32)]), S, ids,
                               words(pres) if with_nan else None, n_threads=3)
    want = "".join(pid + "\t" + "\t".join("" if not p else str(int(v)) for v, p in zip(b, pr)) + "\n"
                   for pid, b, pr in zip(ids, bits, pres))
    assert got.decode() == want


def _kmer_rows_python(r, k, tags, kmer_ids, cluster_ids):
    """kmers_to_hashes text rendered with numpy / Python (what the host mirror did before the
    native formatter): per cluster the header, the narrow rows by k-mer, then the wide rows."""
    from panfeed_b200 import packer
    out = []
    nk = packer.kmers_to_str(r["row_kmer"], k)
    wk = packer.wide_kmers_to_str(r["wide_row_kmer"], k)
    for c, tag in enumerate(tags):
        out.append(tag + b"\t\t" + cluster_ids[r["cluster_pattern"][c]] + b"\n")
        sel = np.flatnonzero(r["row_cluster"] == c)
        sel = sel[np.argsort(r["row_kmer"][sel], kind="stable")]
        for i in sel:
            out.append(tag + b"\t" + nk[i] + b"\t" + kmer_ids[r["row_pattern"][i]] + b"\n")
        sel = np.flatnonzero(r["wide_row_cluster"] == c)
        if len(sel):
            w = r["wide_row_kmer"][sel]
            sel = sel[np.lexsort((w[:, 1], w[:, 0]))]
        for i in sel:
            out.append(tag + b"\t" + wk[i] + b"\t" + kmer_ids[r["wide_row_pattern"][i]] + b"\n")
    return out


@pytest.mark.parametrize("k", [1, 17, 31, 32])
@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("layout", ["interleaved", "runs"])
def test_native_kmer_rows_match_python(k, threads, layout):
    """pf_format_kmer_rows against the reference's f-strings (panfeed.py:177, :208): header row
    per cluster (also for clusters without rows), k-mer rows sorted inside the cluster.
    layout: the rows of the clusters interleaved (record engines: counting sort inside), or one
    run per cluster with the clusters in any order (the block engine: run boundaries only)."""
    rng = np.random.default_rng(7 * k + threads)
    nc = 23
    n = 40_000 if threads > 1 else 2_000
    nw = 300
    ids = np.array([("%022d==" % i).encode() for i in range(500)], "S24")
    cids = np.array([("c%021d==" % i).encode() for i in range(40)], "S24")
    tags = [str(int(x)).encode() for x in rng.integers(0, 100000, nc)]
    row_cluster = rng.integers(0, nc, n).astype(np.uint32)
    row_cluster[row_cluster == 5] = 6                      # cluster 5 has no narrow rows
    if layout == "runs":
        rank = rng.permutation(nc)                         # clusters in shuffled order, each one run
        row_cluster = row_cluster[np.argsort(rank[row_cluster], kind="stable")]
        assert len(np.flatnonzero(np.diff(row_cluster.astype(np.int64)))) == len(np.unique(row_cluster)) - 1
        assert (np.diff(row_cluster.astype(np.int64)) < 0).any()
    # distinct k-mers per cluster are what the library returns; duplicates across clusters are fine
    row_kmer = rng.integers(0, 1 << min(62, 2 * k), n, dtype=np.uint64) if k < 32 else \
        rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    if k == 1:
        row_kmer = rng.integers(0, 4, n, dtype=np.uint64)
    wide = np.zeros((nw, 2), np.uint64)
    for i in range(nw):
        t = "".join(rng.choice(list("ACGTNRYKMSWBDHVX"), k))
        wide[i] = _pack4(t)
    r = {"cluster_pattern": rng.integers(0, 40, nc).astype(np.uint32),
         "row_cluster": row_cluster, "row_kmer": row_kmer,
         "row_pattern": rng.integers(0, 500, n).astype(np.uint32),
         "wide_row_cluster": rng.integers(0, nc, nw).astype(np.uint32), "wide_row_kmer": wide,
         "wide_row_pattern": rng.integers(0, 500, nw).astype(np.uint32)}
    got, off = capi.format_kmer_rows(r, k, tags, ids, cids, n_threads=threads)
    want = _kmer_rows_python(r, k, tags, ids, cids)
    # equal k-mers inside a cluster (possible in this random input only) may swap: compare sorted
    assert sorted(got.split(b"\n")) == sorted(b"".join(want).split(b"\n"))
    assert len(got) == len(b"".join(want)) == int(off[-1])
    # per-cluster slices: header first, then rows in k-mer order
    for c in range(nc):
        lines = got[int(off[c]):int(off[c + 1])].split(b"\n")[:-1]
        assert lines[0] == tags[c] + b"\t\t" + cids[r["cluster_pattern"][c]]
        n_c = int((row_cluster == c).sum())
        kms = [ln.split(b"\t")[1] for ln in lines[1:1 + n_c]]
        assert kms == sorted(kms)
        assert len(lines) == 1 + n_c + int((r["wide_row_cluster"] == c).sum())


def test_native_kmer_rows_reject_bad_indices():
    ids = np.array([b"x" * 24], "S24")
    r = {"cluster_pattern": np.zeros(2, np.uint32), "row_cluster": np.array([0, 2], np.uint32),
         "row_kmer": np.zeros(2, np.uint64), "row_pattern": np.zeros(2, np.uint32),
         "wide_row_cluster": np.zeros(0, np.uint32), "wide_row_kmer": np.zeros((0, 2), np.uint64),
         "wide_row_pattern": np.zeros(0, np.uint32)}
    with pytest.raises(capi.PfError):
        capi.format_kmer_rows(r, 5, [b"0", b"1"], ids, ids)          # row of cluster 2 of 2
    r["row_cluster"] = np.array([0, 1], np.uint32)
    r["row_pattern"] = np.array([0, 1], np.uint32)
    with pytest.raises(capi.PfError):
        capi.format_kmer_rows(r, 5, [b"0", b"1"], ids, ids)          # pattern 1 of 1


@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("member_bytes", [0, 1 << 16, 1 << 20])
def test_native_gzip_members_round_trip(threads, member_bytes):
    """pf_gzip_members: independently deflated members, concatenated - Python's gzip must read
    them back as the one text (what `--compress` readers do with the three outputs)."""
    import gzip
    rng = np.random.default_rng(member_bytes + threads)
    rows = [b"%d\t%s\t%s\n" % (i % 97, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 31)), b"x" * 22 + b"==")
            for i in range(40_000)]
    data = b"".join(rows)
    z = capi.gzip_members(data, 9, member_bytes, threads)
    assert gzip.decompress(z) == data
    assert len(z) < len(data) // 2
    if member_bytes == 1 << 16:
        assert z.count(b"\x1f\x8b\x08") >= len(data) // member_bytes      # one header per member
    # pieces written one after the other form one stream
    assert gzip.decompress(capi.gzip_members(data[:1000], 1) + capi.gzip_members(data[1000:5000], 9)) == data[:5000]
    assert gzip.decompress(capi.gzip_members(b"")) == b""


def test_gzip_text_writer_matches_gzip_open(tmp_path):
    """The --compress handles of the host mirror: same text back as gzip.open(..., "wt")."""
    import gzip
    from panfeed_b200.input import GzipTextWriter
    path = str(tmp_path / "kmers_to_hashes.tsv.gz")
    w = GzipTextWriter(path, flush_bytes=10_000)
    text = []
    for i in range(3000):
        row = f"{i}\t{'ACGT' * 7}\t{'h' * 22}==\n"
        w.write(row)
        text.append(row)
        if i % 500 == 0:
            w.flush()
    w.close()
    w.close()
    assert gzip.open(path, "rt").read() == "".join(text)
    empty = str(tmp_path / "empty.tsv.gz")
    GzipTextWriter(empty).close()
    assert gzip.open(empty, "rt").read() == ""


@pytest.mark.parametrize("k", [1, 15, 31, 32])
@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("threads", [1, 4])
def test_native_compact_position_rows_match_python(k, canonical, threads):
    """pf_format_positions_compact (emit_positions = 2: the device only returns the used_strand
    bit of every window) against a Python rendering of panfeed.py:64-107 on the same batch:
    k-mer text from the packed planes, coordinates from the descriptors."""
    from panfeed_b200 import packer
    rng = np.random.default_rng(7 * k + canonical)
    n_seqs = 90 if threads > 1 else 12
    seqs_txt = []
    for i in range(n_seqs):
        L = int(rng.integers(0, 140))
        alphabet = list("ACGTNRYK") if i % 7 == 3 else list("ACGT")
        seqs_txt.append("".join(rng.choice(alphabet, L)) if L else "")
    packed, base_off, is_amb, amb_plane, amb_off = capi.pack_sequences([s.encode() for s in seqs_txt])
    seqs = np.zeros(n_seqs, capi.SEQ_DTYPE)
    seqs["base_off"], seqs["amb_off"] = base_off, amb_off
    seqs["len"] = [len(s) for s in seqs_txt]
    seqs["flags"] = (rng.random(n_seqs) < 0.7) * capi.PF_SEQ_TARGET + is_amb * capi.PF_SEQ_AMBIGUOUS
    seqs["start"] = rng.integers(1, 4_000_000, n_seqs)
    seqs["offset"] = rng.integers(0, 120, n_seqs)
    seqs["strand"] = rng.choice([1, -1], n_seqs)
    # coordinates that run through a change of their number of digits, upwards and downwards
    # (the formatter steps decimal texts from row to row), and through zero
    edge = [1, 5, 9, 95, 995, 9_990, 99_995, 999_990, 1, 8, 99, 950, 9_900, 99_900]
    ne = min(len(edge), n_seqs)
    seqs["start"][:ne] = edge[:ne]
    seqs["strand"][:ne] = ([1, -1] * len(edge))[:ne]
    seqs["offset"][:ne] = ([0, 3, 10, 99, 100, 101, 7] * 2)[:ne]
    seqs["end"] = seqs["start"] + seqs["len"] - 1
    hb = capi.HostBatch(packed, seqs, np.zeros(1, capi.CLUSTER_DTYPE), np.zeros((1, 1), np.uint32), amb_plane)
    leads = [f"cl{i % 4}\tstrain_{i}\tgene{i}\tctg{i % 3}\t{int(seqs['strand'][i])}\t".encode() for i in range(n_seqs)]
    # the bit plane the device would return: 1 where the reverse complement is the canonical k-mer
    bits = np.zeros(len(packed), np.uint32)
    want = []
    for i, s in enumerate(seqs_txt):
        q = seqs[i]
        for p in range(len(s) - k + 1):
            fwd = s[p:p + k]
            rc = fwd.encode().translate(_COMP)[::-1].decode()
            c0 = int(q["start"]) + p if q["strand"] > 0 else int(q["end"]) - p - k
            g0 = p - int(q["offset"])
            use_rc = not fwd <= rc
            if use_rc:
                j = int(q["base_off"]) + p
                bits[j >> 5] |= np.uint32(1 << (j & 31))
            if not (q["flags"] & capi.PF_SEQ_TARGET):
                continue
            head = leads[i].decode() + f"{c0}\t{c0 + k}\t{g0}\t{g0 + k}\t"
            if canonical:
                want.append(f"{head}{-1 if use_rc else 1}\t{rc if use_rc else fwd}\n")
            else:
                st = int(q["strand"])
                want.append(f"{head}{st}\t{fwd}\n{head}{-st}\t{rc}\n")
    got = capi.format_positions_compact(hb, bits if canonical else None, k, canonical, leads, n_threads=threads)
    assert got.decode() == "".join(want)


def test_native_base64_ids_match_python():
    """pf_base64_ids (the text of the pattern ids, panfeed.py:175-176: b2a_base64(md5)[:24])
    against Python's base64 and the numpy rendering."""
    import base64
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, (70_000, 16)).astype(np.uint8)
    d[0], d[1] = 0, 255
    for threads in (1, 0):
        ids = capi.base64_ids(d, threads)
        assert ids.dtype == np.dtype("S24") and (ids == capi.base64_ids_numpy(d)).all()
        assert [bytes(x) for x in ids[:50]] == [base64.b64encode(x.tobytes()) for x in d[:50]]
    assert capi.base64_ids(d[:0]).shape == (0,)
