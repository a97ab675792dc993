"""ORACLE — TEST INFRASTRUCTURE ONLY (ctypes wrapper of oracle/oracle.c).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs
may import this.  Takes the same per-cluster items the reference's
iter_gene_clusters yields (input.py:468): (dict strain -> [Seqinfo-like],
idx, clusterpresab) and returns numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


class OrSeq(C.Structure):
    _fields_ = [("off", C.c_uint64), ("len", C.c_uint32),
                ("cluster", C.c_uint32), ("sample", C.c_uint32),
                ("flags", C.c_uint32), ("start", C.c_int32),
                ("end", C.c_int32), ("offset", C.c_int32),
                ("strand", C.c_int32)]


SEQ_DTYPE = np.dtype([("off", "<u8"), ("len", "<u4"), ("cluster", "<u4"),
                      ("sample", "<u4"), ("flags", "<u4"), ("start", "<i4"),
                      ("end", "<i4"), ("offset", "<i4"), ("strand", "<i4")])


class OrParams(C.Structure):
    _fields_ = [("k", C.c_uint32), ("n_samples", C.c_uint32),
                ("canonical", C.c_uint32), ("consider_missing", C.c_uint32),
                ("cluster_equal_filter", C.c_uint32),
                ("n_threads", C.c_uint32), ("maf", C.c_double)]


class OrResult(C.Structure):
    _fields_ = [("n_rows", C.c_uint64),
                ("row_cluster", C.POINTER(C.c_uint32)),
                ("row_kmer", C.POINTER(C.c_char)),
                ("row_count", C.POINTER(C.c_uint32)),
                ("row_pattern", C.POINTER(C.c_uint32)),
                ("n_clusters", C.c_uint32),
                ("cluster_pattern", C.POINTER(C.c_uint32)),
                ("n_kmer_patterns", C.c_uint64),
                ("kmer_pattern_bits", C.POINTER(C.c_uint32)),
                ("kmer_pattern_cluster", C.POINTER(C.c_uint32)),
                ("n_cluster_patterns", C.c_uint64),
                ("cluster_pattern_bits", C.POINTER(C.c_uint32)),
                ("n_pos", C.c_uint64),
                ("pos_seq", C.POINTER(C.c_uint32)),
                ("pos_pos", C.POINTER(C.c_uint32)),
                ("pos_used_strand", C.POINTER(C.c_int32)),
                ("pos_kmer", C.POINTER(C.c_char)),
                ("n_unique", C.c_uint64), ("n_instances", C.c_uint64)]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or (
            os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "oracle.c"))):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"],
                       check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_run.restype = C.c_int
        _lib.oracle_run.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32,
                                    C.c_uint32, C.c_void_p,
                                    C.POINTER(OrParams), C.POINTER(OrResult)]
        _lib.oracle_free.argtypes = [C.POINTER(OrResult)]
    return _lib


def flatten(items, stroi):
    """Cluster items -> (ascii bases, seq array, presab matrix, ids, seq meta)."""
    chunks, seqs, meta, presabs, ids = [], [], [], [], []
    off = 0
    for ci, (cluster, idx, presab) in enumerate(items):
        rank = {s: i for i, s in enumerate(sorted(cluster.keys()))}
        ids.append(idx)
        presabs.append(np.asarray(presab, dtype=np.uint8))
        for strain in cluster.keys():
            for q in cluster[strain]:
                b = q.sequence.encode()
                chunks.append(b)
                seqs.append((off, len(b), ci, rank[strain],
                             1 if strain in stroi else 0, q.start, q.end,
                             q.offset, q.strand))
                meta.append((strain, q.id, q.chromosome))
                off += len(b)
    arr = np.array(seqs, dtype=SEQ_DTYPE) if seqs else np.zeros(0, SEQ_DTYPE)
    pres = (np.stack(presabs) if presabs else np.zeros((0, 0), np.uint8))
    return b"".join(chunks), arr, np.ascontiguousarray(pres), ids, meta


def run_arrays(bases, seqs, presab, k, canonical=True, consider_missing=False,
               cluster_equal_filter=False, maf=0.01, n_threads=1):
    """bases: bytes/uint8 array (ASCII); seqs: SEQ_DTYPE array;
    presab: uint8 [n_clusters, S]."""
    L = lib()
    n_clusters, S = presab.shape
    p = OrParams(k, S, int(canonical), int(consider_missing),
                 int(cluster_equal_filter), n_threads, maf)
    res = OrResult()
    if isinstance(bases, np.ndarray):
        bptr = bases.ctypes.data_as(C.c_char_p)
    else:
        bptr = C.c_char_p(bases)
    seqs = np.ascontiguousarray(seqs)
    presab = np.ascontiguousarray(presab, dtype=np.uint8)
    rc = L.oracle_run(bptr, seqs.ctypes.data, len(seqs), n_clusters,
                      presab.ctypes.data, C.byref(p), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"oracle_run failed: {rc}")
    W = (S + 31) // 32

    def arr(ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    out = {
        "row_cluster": arr(res.row_cluster, res.n_rows, np.uint32),
        "row_kmer": (np.frombuffer(C.string_at(res.row_kmer, res.n_rows * k),
                                   dtype=f"S{k}").copy()
                     if res.n_rows else np.zeros(0, f"S{k}")),
        "row_count": arr(res.row_count, res.n_rows, np.uint32),
        "row_pattern": arr(res.row_pattern, res.n_rows, np.uint32),
        "cluster_pattern": arr(res.cluster_pattern, res.n_clusters, np.uint32),
        "kmer_pattern_bits": arr(res.kmer_pattern_bits,
                                 res.n_kmer_patterns * W,
                                 np.uint32).reshape(-1, W),
        "kmer_pattern_cluster": arr(res.kmer_pattern_cluster,
                                    res.n_kmer_patterns, np.uint32),
        "cluster_pattern_bits": arr(res.cluster_pattern_bits,
                                    res.n_cluster_patterns * W,
                                    np.uint32).reshape(-1, W),
        "pos_seq": arr(res.pos_seq, res.n_pos, np.uint32),
        "pos_pos": arr(res.pos_pos, res.n_pos, np.uint32),
        "pos_used_strand": arr(res.pos_used_strand, res.n_pos, np.int32),
        "pos_kmer": (np.frombuffer(C.string_at(res.pos_kmer, res.n_pos * k),
                                   dtype=f"S{k}").copy()
                     if res.n_pos else np.zeros(0, f"S{k}")),
        "n_unique": int(res.n_unique), "n_instances": int(res.n_instances),
        "S": S, "k": k,
    }
    L.oracle_free(C.byref(res))
    return out


def run(items, stroi, k, canonical=True, consider_missing=False,
        cluster_equal_filter=False, maf=0.01, n_threads=1):
    items = list(items)
    bases, seqs, presab, ids, meta = flatten(items, stroi)
    if presab.size == 0:
        presab = presab.reshape(len(items), 0)
    out = run_arrays(bases, seqs, presab, k, canonical, consider_missing,
                     cluster_equal_filter, maf, n_threads)
    out["ids"], out["seq_meta"], out["seqs"] = ids, meta, seqs
    return out


# ---- rendering to the reference's text formats (for comparisons) ----------
def ternary(bits_row, mask_row, S):
    """-> float64[S] with NaN where mask says the cluster is absent."""
    idx = np.arange(S)
    v = ((bits_row[idx >> 5] >> (idx & 31)) & 1).astype(np.float64)
    if mask_row is not None:
        m = ((mask_row[idx >> 5] >> (idx & 31)) & 1).astype(bool)
        v[~m] = np.nan
    return v


def binary_int(bits_row, S):
    idx = np.arange(S)
    return ((bits_row[idx >> 5] >> (idx & 31)) & 1).astype(np.int64)
