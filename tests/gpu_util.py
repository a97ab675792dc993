"""Drive the CUDA path through the C-ABI for the parity tests and render its
arrays to the reference's line formats (independently of panfeed_b200.engine)."""
import numpy as np

from panfeed_b200 import capi, packer

_COMP = str.maketrans("ACTGNYRWSKMDVHBX", "TGACNRYWSMKHBDVX")


def run_gpu(items, stroi, S, k, canonical=True, consider_missing=False,
            cluster_equal_filter=False, maf=0.01, batch_clusters=3,
            sort_bits=0, mode=0, debug_flags=0):
    """items: list of (cluster dict, idx, presab).  Returns a dict shaped like
    oracle_c.run()'s, plus the positional arrays the kernels computed."""
    ctx = capi.Context(k, S, canonical, consider_missing, cluster_equal_filter,
                       emit_positions=bool(stroi), maf=maf, sort_bits=sort_bits,
                       mode=mode, debug_flags=debug_flags)
    res = {"row_cluster": [], "row_kmer": [], "row_count": [], "row_pattern": [],
           "cluster_pattern": [], "kp": [], "cp": [], "pos": [], "ids": [],
           "seq_meta": [], "seqs": [], "seq_cluster_idx": []}
    seq_base = 0
    try:
        # batch_clusters: clusters per batch, or a list of batch sizes used in turn
        sizes = batch_clusters if isinstance(batch_clusters, (list, tuple)) else [batch_clusters]
        starts, b0, i = [], 0, 0
        while b0 < len(items):
            starts.append((b0, sizes[i % len(sizes)]))
            b0 += sizes[i % len(sizes)]
            i += 1
        for b0, bc in starts:
            chunk = items[b0:b0 + bc]
            pcs = [packer.PackedCluster(c, idx, pa, stroi) for c, idx, pa in chunk]
            hb, meta, ids = packer.pack_batch(pcs, list(range(b0, b0 + len(chunk))))
            ctx.submit(hb)
            r = ctx.collect()
            assert r["kmer_pattern_base"] == sum(len(x) for x in res["kp"])
            assert r["cluster_pattern_base"] == sum(len(x) for x in res["cp"])
            res["row_cluster"] += [r["row_cluster"], r["wide_row_cluster"]]
            res["row_kmer"] += [packer.kmers_to_str(r["row_kmer"], k),
                                packer.wide_kmers_to_str(r["wide_row_kmer"], k)]
            res["row_count"] += [r["row_count"], r["wide_row_count"]]
            res["row_pattern"] += [r["row_pattern"], r["wide_row_pattern"]]
            res["cluster_pattern"].append(r["cluster_pattern"])
            res["kp"].append(r["new_kmer_patterns"])
            res["cp"].append(r["new_cluster_patterns"])
            kmer_s = packer.kmers_to_str(r["pos_kmer"], k)
            wide_s = packer.wide_kmers_to_str(r["pos_wide_kmer"], k)
            isw = (r["pos_flags"] & 2) != 0
            if isw.any():
                kmer_s = kmer_s.copy()
                kmer_s[isw] = wide_s[r["pos_kmer"][isw].astype(np.int64)]
            res["pos"].append((r["pos_seq"].astype(np.int64) + seq_base, kmer_s,
                               r["pos_contig_start"], r["pos_gene_start"],
                               r["pos_flags"]))
            res["ids"] += ids
            res["seq_cluster_idx"] += [ids[c] for c in hb.seqs["cluster"]]
            res["seq_meta"] += meta
            res["seqs"].append(hb.seqs)
            seq_base += len(hb.seqs)
        stats = ctx.stats()
    finally:
        ctx.close()
    W = (S + 31) // 32
    out = {
        "row_cluster": np.concatenate(res["row_cluster"]) if res["row_cluster"] else np.zeros(0, np.uint32),
        "row_kmer": np.concatenate(res["row_kmer"]) if res["row_kmer"] else np.zeros(0, f"S{k}"),
        "row_count": np.concatenate(res["row_count"]) if res["row_count"] else np.zeros(0, np.uint32),
        "row_pattern": np.concatenate(res["row_pattern"]) if res["row_pattern"] else np.zeros(0, np.uint32),
        "cluster_pattern": np.concatenate(res["cluster_pattern"]) if res["cluster_pattern"] else np.zeros(0, np.uint32),
        "ids": res["ids"], "seq_meta": res["seq_meta"],
        "seqs": np.concatenate(res["seqs"]) if res["seqs"] else np.zeros(0, capi.SEQ_DTYPE),
        "stats": stats, "seq_cluster_idx": res["seq_cluster_idx"],
    }
    kp = np.concatenate(res["kp"]) if res["kp"] else np.zeros((0, W), np.uint32)
    out["kmer_pattern_bits"] = kp[:, :W]
    out["kmer_pattern_cluster"] = (kp[:, W] if consider_missing
                                   else np.full(len(kp), 0xffffffff, np.uint32))
    out["cluster_pattern_bits"] = (np.concatenate(res["cp"]) if res["cp"]
                                   else np.zeros((0, W), np.uint32))
    out["pos"] = res["pos"]
    return out


def kmers_tsv_lines(out, k, canonical):
    """Positional records -> kmers.tsv lines (panfeed.py:104-107)."""
    lines = []
    seqs, meta, ids = out["seqs"], out["seq_meta"], out["ids"]
    cl_base = 0
    for (pseq, kmer_s, cstart, gstart, flags) in out["pos"]:
        for si, km, c0, g0, fl in zip(pseq, kmer_s, cstart, gstart, flags):
            q = seqs[si]
            strain, fid, contig = meta[si]
            km = km.decode()
            lead = (f"{out['seq_cluster_idx'][si]}\t{strain}\t{fid}\t{contig}\t"
                    f"{q['strand']}\t{c0}\t{c0 + k}\t{g0}\t{g0 + k}\t")
            if canonical:
                lines.append(f"{lead}{-1 if fl & 1 else 1}\t{km}")
            else:
                lines.append(f"{lead}{q['strand']}\t{km}")
                lines.append(f"{lead}{-q['strand']}\t{km.translate(_COMP)[::-1]}")
    return lines
