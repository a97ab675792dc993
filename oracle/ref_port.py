"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by `panfeed_b200`.

Plain-Python/numpy restatement of the reference's per-gene-cluster k-mer
streaming path, written to be read side by side with the reference:

    kmer_stage()     restates  panfeed/panfeed.py:23-113   (cluster_cutter)
    blank_vector()   restates  panfeed/panfeed.py:16-20    (init_presabs_vector)
    pattern_stage()  restates  panfeed/panfeed.py:132-235  (pattern_hasher)
    headers()        restates  panfeed/panfeed.py:116-129  (write_headers)
    feed_clusters()  restates  panfeed/input.py:335-468    (iter_gene_clusters)
    read_gff_cds()   restates  panfeed/input.py:274-332    (parse_gff)
    run()            restates  panfeed/__main__.py:226-369 single-process loop

Parity pin: tests/test_oracle_golden.py checks `run()` byte-for-byte against
the outputs of the unmodified reference committed under tests/golden/expected
(18 argument combinations incl. the 12 of tests/unit_test.sh) and the
hot-function known answers in hot_kats.json.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py --impl reference / cpu_baseline` may
import this module.

Only behaviour is restated; loops are pure Python, so use it for small cases
(seconds up to ~10^5 k-mer instances).  oracle/oracle.c is the fast twin.
"""
import binascii
import hashlib
import io
import os
from collections import namedtuple

import numpy as np
import pandas as pd

CutSeq = namedtuple("CutSeq", "sequence compsequence id chromosome start end "
                              "strand offset")
CdsFeature = namedtuple("CdsFeature", "id chromosome start end strand")

# complement table of pyfaidx (third-party, un-vendored)
_COMPLEMENT = str.maketrans("ACTGNactgnYRWSKMDVHBXyrwskmdvhbx",
                            "TGACNtgacnRYWSMKHBDVXrywsmkhbdvx")

KMERS_HEADER = ("cluster\tstrain\tfeature_id\tcontig\tfeature_strand\t"
                "contig_start\tcontig_end\tgene_start\tgene_end\tstrand\t"
                "k-mer\n")                                  # input.py:243


def pattern_id(vec):
    """panfeed.py:175-176 / 206-207: base64(md5(raw bytes))[:24]."""
    return binascii.b2a_base64(
        hashlib.md5(np.ascontiguousarray(vec).view(np.uint8)).digest()
    ).decode()[:24]


def blank_vector(n, clusterpresab, missing_nan):
    v = np.zeros(n, dtype=np.float64)
    if missing_nan:
        v[clusterpresab == 0] = np.nan
    return v


def kmer_stage(item, k, stroi, canon, consider_missing):
    """One cluster -> (idx, {kmer: float64[S]} in first-seen order,
    clusterpresab, kmers.tsv chunk)."""
    cluster, idx, clusterpresab = item
    rank = {s: i for i, s in enumerate(sorted(cluster.keys()))}
    template = blank_vector(len(rank), clusterpresab, consider_missing)
    table = {}
    rows = []

    def mark(kmer, col):
        vec = table.get(kmer)
        if vec is None:                 # reference deep-copies every time;
            vec = template.copy()       # only the first copy is ever kept
            table[kmer] = vec
        vec[col] = 1

    for strain in cluster.keys():
        col = rank[strain]
        for seq in cluster[strain]:
            fwd_all, comp_all = seq.sequence, seq.compsequence
            for pos in range(len(fwd_all) - k + 1):
                fwd = fwd_all[pos:pos + k]
                rev = comp_all[pos:pos + k][::-1]
                if canon:
                    if fwd <= rev:
                        chosen, used = fwd, 1
                    else:
                        chosen, used = rev, -1
                    mark(chosen, col)
                else:
                    used = seq.strand
                    mark(fwd, col)
                    mark(rev, col)
                if strain in stroi:
                    if seq.strand > 0:
                        c0 = seq.start + pos
                        c1 = seq.start + pos + k
                    else:
                        c1 = seq.end - pos
                        c0 = seq.end - pos - k
                    g0 = pos - seq.offset
                    g1 = pos + k - seq.offset
                    lead = (f"{idx}\t{strain}\t{seq.id}\t{seq.chromosome}\t"
                            f"{seq.strand}\t{c0}\t{c1}\t{g0}\t{g1}\t")
                    if canon:
                        rows.append(f"{lead}{used}\t{chosen}\n")
                    else:
                        rows.append(f"{lead}{used}\t{fwd}\n")
                        rows.append(f"{lead}{-used}\t{rev}\n")
    return idx, table, clusterpresab, "".join(rows)


def headers(sample_names):
    h2p = "hashed_pattern" + "".join(f"\t{s}" for s in sorted(sample_names))
    return h2p + "\n", "cluster\tk-mer\thashed_pattern\n"


def _cells(vec, consider_missing):
    if not consider_missing:
        return "\t".join(map(str, vec.astype(np.uint8)))
    return "\t".join("" if np.isnan(x) else str(int(x)) for x in vec)


def pattern_stage(results, patfilt, maf, consider_missing, patterns):
    """Iterable of kmer_stage() results -> (kmers.tsv text,
    hashes_to_patterns text, kmers_to_hashes text); updates `patterns`."""
    out_k, out_p, out_h = [], [], []
    for idx, table, clusterpresab, chunk in results:
        if chunk is not None:
            out_k.append(chunk)
        cid = pattern_id(clusterpresab)            # int64 bytes
        out_h.append(f"{idx}\t\t{cid}\n")
        if cid not in patterns:
            patterns.add(cid)
            out_p.append(f"{cid}\t{_cells(clusterpresab, consider_missing)}\n")
        for kmer, vec in table.items():
            if not consider_missing:
                af = vec.sum() / vec.shape[0]
            else:
                seen = vec[~np.isnan(vec)]
                af = seen.sum() / seen.shape[0]
            if af >= 0.5:
                af = 1 - af
            if af < maf:
                continue
            # NB the flag is inverted in the reference: the filter runs only
            # when patfilt is False, i.e. when --no-filter IS given.
            if patfilt == False and tuple(vec) == tuple(clusterpresab):  # noqa
                continue
            pid = pattern_id(vec)                  # float64 bytes
            out_h.append(f"{idx}\t{kmer}\t{pid}\n")
            if pid in patterns:
                continue
            patterns.add(pid)
            out_p.append(f"{pid}\t{_cells(vec, consider_missing)}\n")
    return "".join(out_k), "".join(out_p), "".join(out_h)


# --------------------------------------------------------------------------
# feeder
# --------------------------------------------------------------------------
def read_fasta(path, upper=True):
    recs, name, buf = {}, None, []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    recs[name] = "".join(buf)
                toks = line[1:].split()
                name, buf = (toks[0] if toks else ""), []
            elif name is not None:
                buf.append(line.strip())
    if name is not None:
        recs[name] = "".join(buf)
    if upper:
        recs = {n: s.upper() for n, s in recs.items()}
    return recs


def read_gff_fasta(path, upper=True):
    """Nucleotides after the ##FASTA marker (input.py:105-108)."""
    text = open(path).read().split("##FASTA")[1]
    tmp = io.StringIO(text)
    recs, name, buf = {}, None, []
    for line in tmp:
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if name is not None:
                recs[name] = "".join(buf)
            toks = line[1:].split()
            name, buf = (toks[0] if toks else ""), []
        elif name is not None:
            buf.append(line.strip())
    if name is not None:
        recs[name] = "".join(buf)
    if upper:
        recs = {n: s.upper() for n, s in recs.items()}
    return recs


def read_gff_cds(path):
    feats = {}
    with open(path) as fh:
        for line in fh:
            if line.lstrip().startswith("##FASTA"):
                break
            if line.lstrip().startswith("#"):
                continue
            cols = line.split("\t")
            try:
                if cols[2] != "CDS":
                    continue
                chrom, start, end = cols[0], int(cols[3]), int(cols[4])
                strand = 1 if cols[6] == "+" else -1
                fid = None
                for kv in cols[8].split(";"):
                    if kv.startswith("ID") and "=" in kv:
                        fid = kv.split("=")[1]
                if fid is None:
                    continue
                feats[fid] = CdsFeature(fid, chrom, start, end, strand)
            except Exception:
                continue
    return feats


def feed_clusters(table, genomes, up, down, down_start_codon, gene_list=None):
    """`table`: panaroo frame (clusters x strains); `genomes`:
    {strain: ({contig: upper-case str}, {id: CdsFeature})}."""
    for idx, row in table.iterrows():
        if gene_list is not None and idx not in gene_list:
            continue
        strains = row.index
        rank = {s: i for i, s in enumerate(sorted(strains))}
        have = row.dropna()
        presab = np.zeros(len(strains), dtype=int)
        for s in have.index:
            presab[rank[s]] = 1
        seqs = {}
        for strain, cell in have.items():
            strain = str(strain)
            if strain not in genomes:
                continue
            contigs, feats = genomes[strain]
            seqs[strain] = []
            for gene in cell.split(";"):
                f = feats.get(gene)
                if f is None or f.chromosome not in contigs:
                    continue
                ctg = contigs[f.chromosome]
                off_u = f.start - 1 if (f.strand > 0 and f.start - 1 - up < 0) else up
                off_d = f.start - 1 if (f.strand < 0 and f.start - 1 - down < 0) else down
                if not down_start_codon:
                    if f.strand > 0:
                        a, b = f.start - 1 - off_u, f.end + off_d
                        s0, s1 = f.start - off_u, f.end + off_d
                    else:
                        a, b = f.start - 1 - off_d, f.end + off_u
                        s0, s1 = f.start - off_d, f.end + off_u
                else:
                    if f.strand > 0:
                        a, b = f.start - 1 - off_u, f.start + off_d
                        s0, s1 = f.start - off_u, f.start + off_d
                    else:
                        a, b = f.end - 1 - off_d, f.end + off_u
                        s0, s1 = f.end - off_d, f.end + off_u
                piece = ctg[a:b]                       # python slice semantics
                if f.strand < 0:
                    piece = piece.translate(_COMPLEMENT)[::-1]
                comp = piece.translate(_COMPLEMENT)    # complement, not reversed
                seqs[strain].append(CutSeq(piece, comp, f.id, f.chromosome,
                                           s0, s1, f.strand, off_u))
        for s in strains.difference(have.index):
            seqs[s] = []
        yield seqs, idx, presab


def load_inputs(gff, fasta=None):
    """input.py:16-138 reduced to what the path needs: genome -> (contigs, CDS)."""
    def listing(arg, exts):
        if os.path.isfile(arg):
            files = [x.rstrip() for x in open(arg)]
        else:
            files = [os.path.join(arg, f) for f in os.listdir(arg)]
        out = {}
        for f in files:
            base = os.path.split(f)[-1]
            if any(f.endswith(e) for e in exts):
                out[".".join(base.split(".")[:-1])] = f
        return out
    gffs = listing(gff, (".gff",))
    fastas = listing(fasta, (".fasta", ".fna")) if fasta is not None else {}
    genomes = {}
    for g in sorted(gffs):
        if g in fastas:
            contigs = read_fasta(fastas[g])
        else:
            contigs = read_gff_fasta(gffs[g])
        genomes[g] = (contigs, read_gff_cds(gffs[g]))
    return genomes


def run(gff, presence_absence, targets=None, genes=None, fasta=None, k=31,
        maf=0.01, upstream=0, downstream=0, downstream_start_codon=False,
        non_canonical=False, no_filter=False, consider_missing=False):
    """Whole single-process CLI run -> dict of the three file contents."""
    table = pd.read_csv(presence_absence, sep=",", index_col=0,
                        low_memory=False).drop(
                            columns=["Non-unique Gene name", "Annotation"])
    stroi = ({x.rstrip("\n") for x in open(targets)}
             if targets is not None else "")
    gene_list = ({x.rstrip("\n") for x in open(genes)}
                 if genes is not None else None)
    genomes = load_inputs(gff, fasta)
    h2p, k2h = headers(table.columns)
    out = {"kmers.tsv": [KMERS_HEADER], "hashes_to_patterns.tsv": [h2p],
           "kmers_to_hashes.tsv": [k2h]}
    patterns = set()
    for item in feed_clusters(table, genomes, upstream, downstream,
                              downstream_start_codon, gene_list):
        res = kmer_stage(item, k, stroi, not non_canonical, consider_missing)
        a, b, c = pattern_stage((res,), not no_filter, maf, consider_missing,
                                patterns)
        out["kmers.tsv"].append(a)
        out["hashes_to_patterns.tsv"].append(b)
        out["kmers_to_hashes.tsv"].append(c)
    return {name: "".join(parts) for name, parts in out.items()}
