"""Independent renderer: result arrays (oracle C or CUDA path) -> the lines of
the reference's three files, so results can be compared with the goldens after
sorting (row order and pattern numbering are free).
Formats: /root/reference/panfeed/panfeed.py:104-107,177,187,208,223."""
import binascii
import hashlib

import numpy as np

_COMP = str.maketrans("ACTGNYRWSKMDVHBX", "TGACNRYWSMKHBDVX")


def md5_id(vec):
    return binascii.b2a_base64(hashlib.md5(
        np.ascontiguousarray(vec).view(np.uint8)).digest()).decode()[:24]


def expand(bits_row, S):
    idx = np.arange(S)
    return (bits_row[idx >> 5] >> (idx & 31)) & 1


def render(res, ids, seq_meta, seqs, S, k, consider_missing, canonical):
    """res: dict with row_cluster,row_kmer(str list),row_pattern,cluster_pattern,
    kmer_pattern_bits,kmer_pattern_cluster,cluster_pattern_bits,
    pos_seq,pos_pos,pos_used_strand,pos_kmer(str list)."""
    cp_ids, cp_lines = [], []
    for row in res["cluster_pattern_bits"]:
        v = expand(row, S).astype(np.int64)
        pid = md5_id(v)
        cp_ids.append(pid)
        cp_lines.append(pid + "\t" + "\t".join(map(str, v)))
    kp_ids, kp_lines = [], []
    for row, cl in zip(res["kmer_pattern_bits"], res["kmer_pattern_cluster"]):
        v = expand(row, S).astype(np.float64)
        if consider_missing:
            m = expand(res["cluster_pattern_bits"][cl], S).astype(bool)
            v[~m] = np.nan
            cells = "\t".join("" if np.isnan(x) else str(int(x)) for x in v)
        else:
            cells = "\t".join(map(str, v.astype(np.uint8)))
        pid = md5_id(v)
        kp_ids.append(pid)
        kp_lines.append(pid + "\t" + cells)
    k2h = [f"{ids[c]}\t\t{cp_ids[p]}" for c, p in enumerate(res["cluster_pattern"])]
    for c, kmer, p in zip(res["row_cluster"], res["row_kmer"], res["row_pattern"]):
        k2h.append(f"{ids[c]}\t{kmer}\t{kp_ids[p]}")
    # hashes_to_patterns: the reference's set is shared by both namespaces
    h2p = sorted(set(cp_lines) | set(kp_lines))
    kt = []
    for si, pos, used, kmer in zip(res["pos_seq"], res["pos_pos"],
                                   res["pos_used_strand"], res["pos_kmer"]):
        q = seqs[si]
        strain, fid, contig = seq_meta[si]
        pos = int(pos)
        if q["strand"] > 0:
            c0 = int(q["start"]) + pos
            c1 = c0 + k
        else:
            c1 = int(q["end"]) - pos
            c0 = c1 - k
        g0 = pos - int(q["offset"])
        g1 = g0 + k
        lead = (f"{ids[q['cluster']]}\t{strain}\t{fid}\t{contig}\t{q['strand']}"
                f"\t{c0}\t{c1}\t{g0}\t{g1}\t")
        if canonical:
            kt.append(f"{lead}{used}\t{kmer}")
        else:
            kt.append(f"{lead}{used}\t{kmer}")
            kt.append(f"{lead}{-used}\t{kmer.translate(_COMP)[::-1]}")
    return {"kmers.tsv": kt, "kmers_to_hashes.tsv": k2h,
            "hashes_to_patterns.tsv": h2p}


def golden_body(text):
    """Golden file text -> sorted data lines (header dropped)."""
    lines = [x for x in text.split("\n") if x != ""]
    header = [x for x in lines if x.startswith(("cluster\t", "hashed_pattern"))]
    assert len(header) == 1
    lines.remove(header[0])
    return sorted(lines)
