"""The reference's hot-path callables, re-hosted on the CUDA library.

Same names, argument meaning and output files as
`/root/reference/panfeed/panfeed.py`:

    cluster_cutter(cluster_gen, klength, stroi, multiple_files, canon,
                   consider_missing_cluster, output, compress)     panfeed.py:23
    pattern_hasher(cluster_dict_iter, kmer_stroi, hash_pat, kmer_hash, genepres,
                   patfilt, maf, output, patterns, consider_missing_cluster,
                   compress)                                        panfeed.py:132
    write_headers(hash_pat, kmer_hash, genepres)                    panfeed.py:116

What changes is where the work happens.  `cluster_cutter` only packs the
cluster (host); the k-mer extraction, sorting, presence-bitset reduction, MAF /
same-as-cluster filters and the global pattern dedup all run on the GPU inside
`pattern_hasher`, many clusters per launch.  The second element of the tuple
`cluster_cutter` returns is therefore a `PackedCluster`, not a dict, and
`patterns` is a `PatternStore` (the device-resident pattern table plus the id
strings handed out so far) instead of a Python set.  There is no CPU path.
"""
import binascii
import hashlib
import logging
import os

import numpy as np

from . import capi, packer
from .input import create_hash_files, create_kmer_stroi

logger = logging.getLogger("panfeed.panfeed")

# k-mer records per GPU batch (whole clusters are never split)
BATCH_RECORDS = 192 * 1024 * 1024

_COMP = bytes.maketrans(b"ACTGNYRWSKMDVHBX", b"TGACNRYWSMKHBDVX")


def pattern_id(vec):
    """base64(md5(raw little-endian bytes of the vector))[:24] (panfeed.py:175-176,
    206-207): int64 bytes for cluster rows, float64 bytes for k-mer rows.  Host-side
    definition of the id, kept for tools and tests; the product derives ids on the
    device (`Context.pattern_ids`, kernel K5)."""
    return binascii.b2a_base64(hashlib.md5(np.ascontiguousarray(vec).view(np.uint8)).digest()
                               ).decode()[:24]


def cluster_cutter(cluster_gen, klength, stroi, multiple_files, canon,
                   consider_missing_cluster, output, compress=False):
    """Pack one cluster for the GPU.  Returns the reference's 4-tuple shape
    (idx, <payload>, clusterpresab, memchunk); the positional rows the reference
    formats here (panfeed.py:90-107) come back from the device in
    `pattern_hasher`, so memchunk is always None."""
    cluster, idx, clusterpresab = cluster_gen
    logger.debug(f"Packing sequences of {idx}")
    pc = packer.PackedCluster(cluster, idx, clusterpresab, stroi)
    pc.k, pc.canonical, pc.consider_missing = klength, bool(canon), bool(consider_missing_cluster)
    return idx, pc, clusterpresab, None


def write_headers(hash_pat, kmer_hash, genepres):
    hash_pat.write("hashed_pattern" + "".join(f"\t{s}" for s in sorted(genepres.columns)) + "\n")
    hash_pat.flush()
    kmer_hash.write("cluster\tk-mer\thashed_pattern\n")
    kmer_hash.flush()


class PatternStore:
    """What `patterns` is in this build: the GPU context (pattern tables and
    pools live in HBM) and the ids of every pattern written so far."""

    def __init__(self):
        self.ctx = None
        self.key = None
        self.kmer_ids = []        # pool index -> id string (float64 namespace)
        self.cluster_ids = []     # pool index -> id string (int64 namespace)
        self.cluster_bits = []    # pool index -> presence words (for the NaN plane)
        self.seen = set()         # the reference's `patterns` set of id strings

    def __len__(self):
        return len(self.seen)

    def context(self, k, S, canonical, consider_missing, cluster_equal_filter, maf,
                device=0, sort_bits=0):
        key = (k, S, canonical, consider_missing, cluster_equal_filter, maf, device)
        if self.ctx is None:
            self.ctx = capi.Context(k, S, canonical, consider_missing, cluster_equal_filter,
                                    emit_positions=True, maf=maf, sort_bits=sort_bits,
                                    device=device)
            self.key = key
        elif key != self.key:
            raise ValueError("pattern_hasher was called with options that differ from the ones "
                             "its PatternStore was created with")
        return self.ctx

    def reset(self):
        """`patterns = set()` of --multiple-files (panfeed.py:165)."""
        if self.ctx is not None:
            self.ctx.reset_patterns()
        self.kmer_ids, self.cluster_ids, self.cluster_bits = [], [], []
        self.seen = set()

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None


def _run_batch(store, ctx, pcs, idxs, S, k, canonical, consider_missing):
    """One GPU batch -> (kmers.tsv text, hashes_to_patterns text, per-cluster
    kmers_to_hashes texts)."""
    hb, meta, ids = packer.pack_batch(pcs, list(range(len(pcs))))
    ctx.submit(hb)
    r = ctx.collect()

    # ---- new patterns -> ids (MD5 on the device, K5) + hashes_to_patterns rows ----
    pat_text = []
    new_cp = r["new_cluster_patterns"]
    if len(new_cp):
        new_ids = [x.decode() for x in ctx.pattern_ids(True, r["cluster_pattern_base"], len(new_cp))]
        fresh = []
        for pid in new_ids:
            fresh.append(pid not in store.seen)
            store.seen.add(pid)
        store.cluster_ids += new_ids
        store.cluster_bits += [w for w in new_cp]
        sel = np.array(fresh, bool)
        pat_text.append(capi.format_patterns(new_cp[sel], S, [p for p, f in zip(new_ids, fresh) if f]).decode())
    new_kp = r["new_kmer_patterns"]
    if len(new_kp):
        W = (S + 31) // 32
        present = None
        if consider_missing:      # NaN cells = samples whose cluster is absent (the key's last word
            present = np.stack([store.cluster_bits[c] for c in new_kp[:, W]])   # names that cluster pattern)
        new_ids = [x.decode() for x in ctx.pattern_ids(False, r["kmer_pattern_base"], len(new_kp))]
        fresh = []
        for pid in new_ids:
            fresh.append(pid not in store.seen)
            store.seen.add(pid)
        store.kmer_ids += new_ids
        sel = np.array(fresh, bool)
        pat_text.append(capi.format_patterns(new_kp[sel], S, [p for p, f in zip(new_ids, fresh) if f],
                                             None if present is None else present[sel]).decode())

    # ---- kmers_to_hashes rows, cluster by cluster: formatted by the library's host threads
    #      (pf_format_kmer_rows; inside a cluster the plain k-mers in alphabetical order, then the
    #      rows holding N/IUPAC symbols - the reference's order is arbitrary too) --------------
    kmer_ids = np.array(store.kmer_ids, dtype="S24") if store.kmer_ids else np.zeros(0, "S24")
    cluster_ids = np.array(store.cluster_ids, dtype="S24") if store.cluster_ids else np.zeros(0, "S24")
    text, off = capi.format_kmer_rows(r, k, [str(idx).encode() for idx in idxs], kmer_ids, cluster_ids)
    hash_texts = [text[int(off[c]):int(off[c + 1])].decode() for c in range(len(idxs))]

    # ---- kmers.tsv rows from the positional records: formatted by the library's host
    #      threads (pf_format_positions), 1e8 rows are too many for Python ----------------
    pos_text = ""
    if len(r["pos_seq"]):
        leads = [f"{idxs[c]}\t{m[0]}\t{m[1]}\t{m[2]}\t{st}\t".encode()
                 for c, st, m in zip(hb.seqs["cluster"].tolist(), hb.seqs["strand"].tolist(), meta)]
        pos_text = capi.format_positions(r, k, canonical, leads, hb.seqs["strand"]).decode()
    return pos_text, "".join(pat_text), hash_texts


def pattern_hasher(cluster_dict_iter, kmer_stroi, hash_pat, kmer_hash, genepres, patfilt, maf,
                   output, patterns=None, consider_missing_cluster=False, compress=False,
                   device=0, sort_bits=0):
    """Consumes `cluster_cutter` results, runs K1..K4 on the GPU in batches of
    whole clusters and writes the three files exactly as the reference does
    (panfeed.py:152-233).  Returns the updated `patterns` (a PatternStore)."""
    multiple_files = hash_pat is None or kmer_hash is None
    if patterns is None or isinstance(patterns, set):
        patterns = PatternStore()
    S = len(genepres.columns)

    pending, pending_records = [], 0

    def flush():
        nonlocal pending, pending_records, hash_pat, kmer_hash
        if not pending:
            return
        pcs = [p[1] for p in pending]
        idxs = [p[0] for p in pending]
        first = pcs[0]
        ctx = patterns.context(first.k, S, first.canonical, bool(consider_missing_cluster),
                               patfilt == False, maf, device, sort_bits)  # noqa: E712
        pos_text, pat_text, hash_texts = _run_batch(patterns, ctx, pcs, idxs, S, first.k,
                                                    first.canonical, bool(consider_missing_cluster))
        if multiple_files:
            path = os.path.join(output, idxs[0])
            if not os.path.exists(path):
                os.mkdir(path)
            ks = create_kmer_stroi(path, compress)
            ks.write(pos_text)
            ks.close()
            hp, kh = create_hash_files(path, compress)
            write_headers(hp, kh, genepres)
            hp.write(pat_text)
            kh.write("".join(hash_texts))
            hp.close()
            kh.close()
        else:
            if kmer_stroi is not None:
                kmer_stroi.write(pos_text)
            hash_pat.write(pat_text)
            kmer_hash.write("".join(hash_texts))
        pending, pending_records = [], 0

    for idx, pc, clusterpresab, memchunk in cluster_dict_iter:
        if multiple_files:
            flush()
            patterns.reset()
        n = pc.n_records(pc.k, pc.canonical)
        if pending and pending_records + n > BATCH_RECORDS:
            flush()
        pending.append((idx, pc))
        pending_records += n
        if multiple_files:
            flush()
    flush()
    if not multiple_files:
        hash_pat.flush()
        kmer_hash.flush()
    return patterns
