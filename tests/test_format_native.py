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
