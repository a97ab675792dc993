"""The native feeder (pf_feeder_*: GFF3 / FASTA parsing and cluster cutting in the library) against
the Python feeder of panfeed_b200/input.py, which mirrors the reference's input.py:274-468 and is
what the golden files were checked with.  Both must hand the packer identical batches.
Pure host code: runs without a GPU."""
import os

import numpy as np
import pandas as pd
import pytest

from panfeed_b200 import feeder as nf
from panfeed_b200 import input as pyin
from panfeed_b200 import capi, packer, panfeed

import helpers

FIX = os.path.join(helpers.GOLDEN, "fixture")


def _table():
    return pd.read_csv(os.path.join(FIX, "gene_presence_absence.csv"), sep=",", index_col=0,
                       low_memory=False).drop(columns=["Non-unique Gene name", "Annotation"])


def _both(up, down, dsc, stroi, fastadir, gene_list=None, canon=True):
    gffdir = os.path.join(FIX, "gffs")
    filelist, fastalist = pyin.what_are_my_inputfiles(gffdir, fastadir)
    table = _table()
    data = pyin.prep_data_n_fasta(filelist, fastalist, gffdir, fastadir, None)
    py = [panfeed.cluster_cutter(x, 31, stroi, False, canon, False, None)
          for x in pyin.iter_gene_clusters(table, data, up, down, dsc, True, gene_list)]
    native, index = nf.prep_feeder(filelist, fastalist, gffdir, fastadir)
    nat = list(nf.iter_packed_clusters(table, native, index, up, down, dsc, stroi, 31, canon, False,
                                       gene_list))
    return py, nat, native


@pytest.mark.parametrize("up,down,dsc", [(0, 0, False), (100, 50, False), (40, 40, True), (5000, 7000, False),
                                          (3, 0, True)])
@pytest.mark.parametrize("fasta", [False, True])
def test_native_feeder_batches_equal_python_feeder(up, down, dsc, fasta):
    stroi = {"s00", "s05"}
    fastadir = os.path.join(FIX, "fastas") if fasta else None
    py, nat, native = _both(up, down, dsc, stroi, fastadir)
    assert [p[0] for p in py] == [n[0] for n in nat]
    for (idx, a, presab_a, _), (_, b, presab_b, _) in zip(py, nat):
        assert (presab_a == presab_b).all(), idx
        assert a.seq_bytes == b.seq_bytes, idx
        for f in ("sample", "target", "start", "end", "offset", "strand"):
            assert (np.asarray(getattr(a, f)) == np.asarray(getattr(b, f))).all(), (idx, f)
        assert a.meta == b.meta, idx
        assert a.n_records(31, True) == b.n_records(31, True)
        assert (b.k, b.canonical, b.consider_missing) == (31, True, False)
    hb_a, meta_a, ids_a = packer.pack_batch([p[1] for p in py])
    hb_b, meta_b, ids_b = packer.pack_batch([n[1] for n in nat])
    assert ids_a == ids_b
    assert list(meta_a) == list(meta_b) and len(meta_a) == len(hb_a.seqs)
    assert meta_b[0] == meta_a[0] and meta_b[len(meta_b) - 1] == meta_a[-1]
    assert (hb_a.packed == hb_b.packed).all()
    assert hb_a.seqs.tobytes() == hb_b.seqs.tobytes()
    assert (hb_a.presence == hb_b.presence).all()
    assert hb_a.clusters.tobytes() == hb_b.clusters.tobytes()
    assert (hb_a.amb is None) == (hb_b.amb is None)
    if hb_a.amb is not None:
        assert (hb_a.amb == hb_b.amb).all()
    native.close()


def test_native_feeder_gene_list_and_missing_genes(caplog):
    """--genes keeps the listed clusters; a feature id that is not in the GFF is skipped with the
    reference's warning (the fixture's group_acc1 names s07_refound_1, which no GFF has)."""
    genes = {line.rstrip("\n") for line in open(os.path.join(FIX, "genes.txt"))}
    py, nat, native = _both(10, 10, False, "", None, genes)
    assert [p[0] for p in py] == [n[0] for n in nat] and len(nat) == len(genes & set(_table().index))
    import logging
    caplog.clear()
    with caplog.at_level(logging.WARNING, logger="panfeed.input"):
        py, nat, native = _both(0, 0, False, "", None, {"group_acc1"})
    msgs = [r.getMessage() for r in caplog.records]
    assert msgs.count("Could not find gene s07_refound_1 from group_acc1 in s07") == 2     # both feeders
    assert py[0][1].seq_bytes == nat[0][1].seq_bytes
    filelist, fastalist = pyin.what_are_my_inputfiles(os.path.join(FIX, "gffs"), None)
    with pytest.raises(KeyError):
        list(nf.iter_packed_clusters(_table(), native, {g: i for i, g in enumerate(filelist)}, 0, 0, False, "",
                                     31, True, False, {"group_acc1"}, raise_missing=True))


GFF = """##gff-version 3
##sequence-region ctg1 1 60
ctg1\tprodigal\tCDS\t5\t16\t.\t+\t0\tID=g1;Name=x
ctg1\tprodigal\tgene\t5\t16\t.\t+\t0\tID=skipped_type
  # an indented comment
ctg1\tprodigal\tCDS\tabc\t16\t.\t+\t0\tID=bad_int
ctg1\tprodigal\tCDS\t20\t31\t.\t-\t0\tName=y;ID=g2=tail;IDX=overrides;z=1
ctg1\tprodigal\tCDS\t 40 \t50_0\t.\t?\t0\tlocus=1;ID=g3;
ctg2\tprodigal\tCDS\t1\t9\t.\t+\t0\tID=g4;z
short\tline\tCDS
ctgX\tprodigal\tCDS\t1\t9\t.\t+\t0\tID=no_contig;z=2
ctg1\tprodigal\tCDS\t1\t6\t.\t+\t0\tID=g1;again=1
ctg1\tprodigal\tCDS\t2\t7\t.\t+\t0\tID=trail
##FASTA
>ctg1 some description
acgtacgtacGTACGTNNRYacgtacgtac
  ACGTACGTACGTACGTACGTACGTACGTAC\r
>ctg2
ACGTAC
>ctg2 again
TTTTTTTTTTTT
"""


def test_native_feeder_parsing_quirks_match_python(tmp_path):
    """Hand-written GFF with the reference parser's quirks: non-CDS rows, malformed rows, the last
    ID-like attribute winning ("IDX=overrides"), ids cut at a second '=', int() tolerance
    (" 40 ", "50_0"), any strand but '+' read as minus, repeated ids / contig names replacing the
    earlier ones, lower case and padded sequence lines, windows running off both contig ends."""
    path = tmp_path / "q.gff"
    path.write_text(GFF)
    feats = pyin.parse_gff(str(path))
    contigs = pyin.read_fasta_text(open(path).read().split("##FASTA")[1].split("\n"))
    native = nf.NativeFeeder()
    g = native.add_genome("q", str(path))
    info = native.genome_info(g)
    assert info["features"] == len(feats) and info["contigs"] == len(contigs)
    assert info["bases"] == sum(len(v) for v in contigs.values())
    got = {}
    for i in range(info["features"]):
        ident, contig, start, end, strand = native.feature(g, i)
        got[ident] = (contig, start, end, strand)
    assert got == {k: (v.chromosome, v.start, v.end, v.strand) for k, v in feats.items()}
    # an ID that ends the line keeps its newline (the reference splits the unstripped line)
    assert set(got) == {"g1", "overrides", "g3", "g4", "no_contig", "trail\n"}
    assert got["g1"][1:3] == (1, 6)
    table = pd.DataFrame({"q": ["g1;overrides;g3;nope;g4;no_contig;"]}, index=["cl"])
    saw_amb = False
    for up, down, dsc in [(0, 0, False), (3, 4, False), (100, 100, False), (2, 2, True), (30, 0, True)]:
        py = list(pyin.iter_gene_clusters(table, {"q": (contigs, feats)}, up, down, dsc, True))
        a = packer.PackedCluster(py[0][0], "cl", py[0][2], {"q"})
        (_, b, presab, _), = nf.iter_packed_clusters(table, native, {"q": g}, up, down, dsc, {"q"}, 5, True, False)
        assert a.seq_bytes == b.seq_bytes, (up, down, dsc)
        assert a.meta == b.meta
        for f in ("sample", "target", "start", "end", "offset", "strand"):
            assert (np.asarray(getattr(a, f)) == np.asarray(getattr(b, f))).all(), (up, down, dsc, f)
        # batches mixing clusters packed at cutting time (N / IUPAC sequences included) with clusters
        # that still hold ASCII: the planes must be those of packing everything at once
        want, _, _ = packer.pack_batch([a, a, a, a])
        for mix in ([b, b, b, b], [a, b, a, b], [b, a, a, b], [a, a, b, b]):
            got, meta, _ = packer.pack_batch(mix)
            assert (got.packed == want.packed).all(), (up, down, dsc)
            assert got.seqs.tobytes() == want.seqs.tobytes(), (up, down, dsc)
            assert (got.amb is None) == (want.amb is None)
            if want.amb is not None:
                assert (got.amb == want.amb).all()
            assert list(meta) == a.meta * 4
        saw_amb |= want.amb is not None
    assert saw_amb                                               # the quirk contig holds N / R / Y
    # the same genome from memory, with the FASTA given separately
    g2 = native.add_genome_text("q2", GFF, ">ctg1\nACGT\n")
    assert native.genome_info(g2) == {"features": len(feats), "contigs": 1, "bases": 4}
    with pytest.raises(Exception):
        native.add_genome_text("q3", GFF.split("##FASTA")[0])       # no sequences at all
    native.close()


def test_cut_packed_equals_cut_then_pack():
    """pf_feeder_cut_packed (contig windows straight into the 2-bit / 4-bit planes, host threads)
    against pf_feeder_cut + pf_pack_2bit / pf_pack_4bit over its ASCII result, on random genomes:
    both strands, windows clamped at both contig ends, empty windows, N / IUPAC runs, lengths
    around the 32- and 64-base word boundaries, enough bases for several threads."""
    rng = np.random.default_rng(7)
    native = nf.NativeFeeder()
    cells, genomes = [], []
    alphabet = np.frombuffer(b"ACGT", np.uint8)
    for gi in range(6):
        n_genes = 180
        lens = rng.integers(1, 2600, n_genes)
        lens[:12] = [1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1, 2]
        contig = alphabet[rng.integers(0, 4, int(lens.sum()) + 50)].copy()
        for _ in range(25):                                    # N / IUPAC runs and lower case
            at = int(rng.integers(0, len(contig) - 8))
            contig[at:at + int(rng.integers(1, 8))] = np.frombuffer(b"NRYKMSWBDHVX", np.uint8)[rng.integers(0, 12)]
        text = contig.tobytes().decode()
        text = text[:500].lower() + text[500:]
        rows, pos, ids = [], 1, []
        for j, ln in enumerate(lens.tolist()):
            strand = "+" if rng.random() < 0.5 else "-"
            rows.append(f"c\tx\tCDS\t{pos}\t{pos + ln - 1}\t.\t{strand}\t0\tID=g{gi}_{j};x=1")
            ids.append(f"g{gi}_{j}")
            pos += ln
        rows.append(f"c\tx\tCDS\t{len(text) + 6000}\t{len(text) + 6010}\t.\t+\t0\tID=g{gi}_off;x=1")     # past the contig end
        ids.append(f"g{gi}_off")
        g = native.add_genome_text(f"s{gi}", "\n".join(rows) + "\n##FASTA\n>c\n" +
                                   "\n".join(text[i:i + 70] for i in range(0, len(text), 70)) + "\n")
        for j in range(0, len(ids), 3):
            cells.append(";".join(ids[j:j + 3]) + (";nope_%d" % j if j % 30 == 0 else ""))     # misses in every thread's range
            genomes.append(g)
    blob = "\n".join(cells).encode()
    genomes = np.array(genomes, np.uint32)
    for up, down, dsc in [(0, 0, False), (17, 40, False), (5000, 5000, False), (10, 25, True)]:
        a = native.cut(genomes, blob, up, down, dsc, prepack=False)
        want = capi.pack_blob(a["ascii"], a["seq_off"])
        b = native.cut(genomes, blob, up, down, dsc, prepack=True)
        for nt in (1, 3, 7):                                  # the look-ups and the packing on that many threads
            c = native.cut(genomes, blob, up, down, dsc, prepack=True, n_threads=nt)
            assert all((b[f] == c[f]).all() for f in ("seq_off", "cell", "feature", "start", "end", "offset",
                                                      "strand", "packed", "base_off", "is_amb", "amb_off"))
            assert b["missing"] == c["missing"] and len(b["missing"]) > 3
        assert b["ascii"] is None and b["n_seqs"] == a["n_seqs"] > 1000
        for f in ("seq_off", "cell", "feature", "start", "end", "offset", "strand"):
            assert (a[f] == b[f]).all(), f
        assert (np.diff(a["seq_off"].astype(np.int64)) == 0).any()          # empty windows are in
        assert (want[0] == b["packed"]).all()
        assert (want[1] == b["base_off"]).all()
        assert (want[2] == b["is_amb"]).all() and want[2].any() and not want[2].all()
        assert (want[3] == b["amb_plane"]).all()
        assert (want[4] == b["amb_off"]).all()
    # a symbol outside the 16 IUPAC codes is refused and named, like the packer does
    g = native.add_genome_text("bad", "c\tx\tCDS\t1\t12\t.\t-\t0\tID=z;x=1\n##FASTA\n>c\nACGTAC*TACGT\n")
    with pytest.raises(ValueError, match=r"\*"):
        native.cut(np.array([g], np.uint32), b"z", 0, 0, False, prepack=True)
    native.close()


@pytest.mark.parametrize("up,down,dsc,first_cells,target", [(0, 0, False, 16384, 96 << 20), (100, 50, False, 7, 3000),
                                                            (40, 40, True, 1, 1), (30, 10, False, 25, 40000)])
def test_packed_batches_equal_packed_clusters(up, down, dsc, first_cells, target):
    """iter_packed_batches (one library call and a few array operations per GPU batch) hands over
    exactly the pf_batch that pack_batch builds from the per-cluster objects of
    iter_packed_clusters, whatever the batch boundaries; sequence metadata on demand."""
    stroi = {"s00", "s05"}
    gffdir = os.path.join(FIX, "gffs")
    filelist, fastalist = pyin.what_are_my_inputfiles(gffdir, None)
    table = _table()
    native, index = nf.prep_feeder(filelist, fastalist, gffdir, None)
    clusters = list(nf.iter_packed_clusters(table, native, index, up, down, dsc, stroi, 31, True, False))
    batches = list(nf.iter_packed_batches(table, native, index, up, down, dsc, stroi, 31, True, False,
                                          target_bases=target, first_cells=first_cells))
    assert sum(len(b[0]) for b in batches) == len(clusters)
    if first_cells < 100:
        assert len(batches) > 1
    at = 0
    for idxs, pb, _, _ in batches:
        mine = clusters[at:at + len(idxs)]
        at += len(idxs)
        assert idxs == [c[0] for c in mine] == pb.idxs
        want, meta, _ = packer.pack_batch([c[1] for c in mine])
        got = pb.hb
        assert (got.packed == want.packed).all()
        assert got.seqs.tobytes() == want.seqs.tobytes()
        assert got.clusters.tobytes() == want.clusters.tobytes()
        assert (got.presence == want.presence).all() and got.presence.dtype == want.presence.dtype
        assert (got.amb is None) == (want.amb is None)
        if want.amb is not None:
            assert (got.amb == want.amb).all()
        assert [pb.meta(i) for i in range(len(got.seqs))] == list(meta)
        assert pb.n_records(31, True) == sum(c[1].n_records(31, True) for c in mine)
        assert (pb.k, pb.canonical, pb.consider_missing) == (31, True, False)
    native.close()


def test_prefetch_order_errors_and_early_exit():
    """feeder.prefetch: same items in the same order; a producer error surfaces at the consumer
    after the items before it; a consumer that stops early does not leave the producer stuck."""
    import threading
    import time
    assert list(nf.prefetch(iter(range(100)), depth=3)) == list(range(100))
    assert list(nf.prefetch(iter(()))) == []

    def failing():
        yield 1
        yield 2
        raise KeyError("gene")
    got = []
    with pytest.raises(KeyError, match="gene"):
        for x in nf.prefetch(failing()):
            got.append(x)
    assert got == [1, 2]

    produced = []

    def endless():
        i = 0
        while True:
            produced.append(i)
            yield i
            i += 1
    before = threading.active_count()
    gen = nf.prefetch(endless(), depth=2)
    assert next(gen) == 0
    gen.close()
    time.sleep(0.3)
    n = len(produced)
    time.sleep(0.3)
    assert len(produced) == n and threading.active_count() == before


@pytest.mark.parametrize("layout", ["regular60", "regular_exact", "one_line", "crlf", "lower", "ragged", "blank_tail",
                                    "blank_inside", "padded", "separate_fasta", "tiny"])
def test_contig_views_match_python_feeder(layout, tmp_path):
    """Regular FASTA records are used where they lie in the (mapped) file text, anything else is
    copied line by line like the reference does: both must cut the same sequences, whatever the
    line width, the length of the last line, the line ends, the case, blank or padded lines, a
    repeated contig name, genes on both strands and windows past both contig ends."""
    rng = np.random.default_rng(hash(layout) % 1000)
    n_contigs = 1 if layout == "tiny" else 4
    lens = [7] if layout == "tiny" else [240, 1200, 61, 3000]
    if layout == "regular_exact":
        lens = [240, 1200, 60, 3000]
    contigs = ["".join(rng.choice(list("ACGT"), n)) for n in lens]
    contigs[0] = contigs[0][:3] + "N" + contigs[0][4:]

    def lines_of(seq, ci):
        if layout in ("regular60", "regular_exact", "crlf", "separate_fasta", "tiny", "blank_tail"):
            out = [seq[i:i + 60] for i in range(0, len(seq), 60)]
        elif layout == "one_line":
            out = [seq]
        elif layout == "lower":
            out = [seq[i:i + 70].lower() if (i // 70) % 2 else seq[i:i + 70] for i in range(0, len(seq), 70)]
        elif layout == "ragged":
            out, i = [], 0
            while i < len(seq):
                w = int(rng.integers(1, 90))
                out.append(seq[i:i + w])
                i += w
        elif layout == "blank_inside":
            out = [seq[i:i + 50] for i in range(0, len(seq), 50)]
            out.insert(len(out) // 2, "")
        elif layout == "padded":
            out = [("  " if i % 120 == 0 else "") + seq[i:i + 60] + (" " if i % 180 == 0 else "")
                   for i in range(0, len(seq), 60)]
        if layout == "blank_tail":
            out += ["", ""]
        return out

    rows, ids = [], []
    for ci, seq in enumerate(contigs):
        pos = 1
        j = 0
        while pos + 5 < len(seq):
            ln = int(rng.integers(3, 200))
            end = min(len(seq), pos + ln - 1)
            strand = "+" if rng.random() < 0.5 else "-"
            rows.append(f"c{ci}\tx\tCDS\t{pos}\t{end}\t.\t{strand}\t0\tID=g{ci}_{j};x=1")
            ids.append(f"g{ci}_{j}")
            pos = end + int(rng.integers(1, 40))
            j += 1
    fasta = []
    for ci, seq in enumerate(contigs):
        fasta.append(f">c{ci} description {ci}")
        fasta += lines_of(seq, ci)
    if layout != "tiny":                                      # a repeated name replaces the earlier record
        fasta.append(">c1 again")
        contigs[1] = contigs[1][::-1]
        fasta += lines_of(contigs[1], 1)
    nl = "\r\n" if layout == "crlf" else "\n"
    gff_text = nl.join(["##gff-version 3"] + rows) + nl
    fasta_text = nl.join(fasta) + nl
    gff_path = tmp_path / "g.gff"
    fasta_path = None
    if layout == "separate_fasta":
        fasta_path = tmp_path / "g.fna"
        gff_path.write_bytes(gff_text.encode())
        fasta_path.write_bytes(fasta_text.encode())
        pcontigs = pyin.read_fasta_text(open(fasta_path).read().split("\n"))
    else:
        gff_path.write_bytes((gff_text + "##FASTA" + nl + fasta_text).encode())
        pcontigs = pyin.read_fasta_text(open(gff_path).read().split("##FASTA")[1].split("\n"))
    feats = pyin.parse_gff(str(gff_path))
    assert [len(pcontigs[f"c{ci}"]) for ci in range(n_contigs)] == lens
    native = nf.NativeFeeder()
    g = native.add_genome("q", str(gff_path), None if fasta_path is None else str(fasta_path))
    info = native.genome_info(g)
    assert info == {"features": len(feats), "contigs": n_contigs, "bases": sum(lens)}
    listed = [native.contig(g, ci) for ci in range(n_contigs)]
    assert [(x[0], x[1]) for x in listed] == [(f"c{ci}", lens[ci]) for ci in range(n_contigs)]
    if layout != "ragged":                      # (a short ragged record can come out regular by chance)
        assert [x[2] for x in listed] == [layout not in ("blank_inside", "padded")] * n_contigs
    table = pd.DataFrame({"q": [";".join(ids)]}, index=["cl"])
    for up, down, dsc in [(0, 0, False), (25, 10, False), (5000, 5000, False), (9, 9, True)]:
        py = list(pyin.iter_gene_clusters(table, {"q": (pcontigs, feats)}, up, down, dsc, True))
        want = [q.sequence.encode() for q in py[0][0]["q"]]
        cut = native.cut(np.zeros(1, np.uint32) + g, ";".join(ids).encode(), up, down, dsc, prepack=False)
        off = cut["seq_off"].astype(np.int64)
        got = [cut["ascii"][off[i]:off[i + 1]] for i in range(cut["n_seqs"])]
        assert got == want, (layout, up, down, dsc)
        packed = native.cut(np.zeros(1, np.uint32) + g, ";".join(ids).encode(), up, down, dsc, prepack=True)
        ref = capi.pack_blob(cut["ascii"], cut["seq_off"])
        assert (packed["packed"] == ref[0]).all() and (packed["is_amb"] == ref[2]).all()
        assert (packed["amb_plane"] is None) == (ref[3] is None)
        if ref[3] is not None:
            assert (packed["amb_plane"] == ref[3]).all() and (packed["amb_off"] == ref[4]).all()
    native.close()


TABLE = '''Gene,Non-unique Gene name,Annotation,s02,s00,s01,"s 03"
cl1,,"hypothetical protein, putative",g2_1,g0_1;g0_2,,g3_1
cl2,x,"two
lines, with a comma",,g0_3,g1_3,NA

"cl3",y,plain,g2_4,"g0_4",nan,g3_4
cl4,,,g2_5
cl5,,"a ""quoted"" word",#N/A,None,g1_6,n/a
cl6,,,-,0,1.5,g3_7
'''


@pytest.mark.parametrize("crlf", [False, True])
def test_panaroo_table_equals_pandas(crlf, tmp_path):
    """feeder.PanarooTable (pf_table_*: the mapped file, cells as offsets) against what the
    reference does - pd.read_csv(..., index_col=0, low_memory=False).drop(columns=[...]) - on a
    table with quoted fields (commas, line ends, escaped quotes in a dropped column), a blank
    line, a short row, pandas' missing-value strings, unsorted columns, both line ends."""
    path = tmp_path / "t.csv"
    path.write_bytes((TABLE.replace("\n", "\r\n") if crlf else TABLE).encode())
    want = pd.read_csv(path, sep=",", index_col=0, low_memory=False, dtype=str).drop(
        columns=["Non-unique Gene name", "Annotation"])
    got = nf.PanarooTable(str(path))
    assert got.columns == [str(c) for c in want.columns] == ["s02", "s00", "s01", "s 03"]
    assert got.index == [str(i) for i in want.index] and got.shape == want.shape
    assert (got.n_present() == want.notna().sum(axis=1).to_numpy()).all()
    vals = want.to_numpy(dtype=object)
    for order in (None, [1, 2, 0, 3], [3, 2, 1, 0]):
        for rows in ([0, 1, 2, 3, 4, 5], [4, 0], [3], []):
            pres, blob = got.cells(rows, order)
            v = vals[rows][:, order] if order is not None else vals[rows]
            assert (pres == pd.notna(v)).all() and pres.shape == (len(rows), 4)
            assert blob == "\n".join(v[pd.notna(v)]).encode()
    sub = got.take([5, 1, 2])
    assert sub.index == ["cl6", "cl2", "cl3"] and sub.shape == (3, 4)
    assert (sub.n_present() == want.iloc[[5, 1, 2]].notna().sum(axis=1).to_numpy()).all()
    assert sub.cells([1], None)[1] == b"g0_3\ng1_3"
    assert sub.take([2, 0]).index == ["cl3", "cl6"]
    (tmp_path / "bad.csv").write_text("Gene,Annotation,s0\ncl1,,g\n")
    with pytest.raises(KeyError):
        nf.PanarooTable(str(tmp_path / "bad.csv"))
    (tmp_path / "long.csv").write_text("Gene,Non-unique Gene name,Annotation,s0\ncl1,,,g,extra,more\n")
    with pytest.raises(capi.PfError, match="Expected 4 fields"):
        nf.PanarooTable(str(tmp_path / "long.csv"))


def test_batches_from_native_table_equal_batches_from_dataframe():
    stroi = {"s00", "s05"}
    gffdir = os.path.join(FIX, "gffs")
    filelist, fastalist = pyin.what_are_my_inputfiles(gffdir, None)
    native, index = nf.prep_feeder(filelist, fastalist, gffdir, None)
    table = nf.PanarooTable(os.path.join(FIX, "gene_presence_absence.csv"))
    a = list(nf.iter_packed_batches(_table(), native, index, 30, 10, False, stroi, 31, True, False, first_cells=20,
                                    target_bases=20000))
    b = list(nf.iter_packed_batches(table, native, index, 30, 10, False, stroi, 31, True, False, first_cells=20,
                                    target_bases=20000))
    assert len(a) == len(b) > 1
    for (ia, pa, _, _), (ib, pb, _, _) in zip(a, b):
        assert ia == ib
        assert (pa.hb.packed == pb.hb.packed).all() and pa.hb.seqs.tobytes() == pb.hb.seqs.tobytes()
        assert (pa.hb.presence == pb.hb.presence).all()
    genes = {"group_acc1", "group_core0"} & set(table.index) or set(table.index[:2])
    c = list(nf.iter_packed_clusters(table.take([0, 2, 3]), native, index, 0, 0, False, stroi, 31, True, False, genes))
    d = list(nf.iter_packed_clusters(_table().iloc[[0, 2, 3]], native, index, 0, 0, False, stroi, 31, True, False, genes))
    assert [x[0] for x in c] == [x[0] for x in d]
    native.close()


def test_panaroo_table_fuzz_against_pandas(tmp_path):
    """Random small tables - quoted fields with commas / line ends / escaped quotes (in the dropped
    columns), pandas' missing-value strings, padded cells, short rows, blank lines, both line
    ends, shuffled columns: PanarooTable must see the table pandas sees."""
    import random
    rnd = random.Random(11)
    atoms = ["g1", "g_2;g_3", "", "NA", "nan", "x y", " lead", "trail ", "a,b", 'q"q', "line\nbreak", "N/A", "None",
             "0", "-", "#NA", "null", "é", "\t"]

    def field(a):
        if any(ch in a for ch in ',"\n') or rnd.random() < 0.15:
            return '"' + a.replace('"', '""') + '"'
        return a
    path = str(tmp_path / "t.csv")
    compared = 0
    for _ in range(150):
        ncol, nrow = rnd.randint(1, 6), rnd.randint(0, 8)
        cols = [f"s{j}" for j in range(ncol)]
        rnd.shuffle(cols)
        lines = ["Gene,Non-unique Gene name,Annotation," + ",".join(cols)]
        for r in range(nrow):
            cells = [rnd.choice([a for a in atoms if '"' not in a]) for _ in range(ncol)]
            keep = ncol if rnd.random() < 0.8 else rnd.randint(0, ncol)
            lines.append(",".join([f"cl{r}", field(rnd.choice(["", "x"])), field(rnd.choice(atoms))] +
                                  [field(c) for c in cells[:keep]]))
            if rnd.random() < 0.1:
                lines.append("")
        nl = rnd.choice(["\n", "\r\n"])
        with open(path, "wb") as fh:
            fh.write((nl.join(lines) + (nl if rnd.random() < 0.8 else "")).encode())
        want = pd.read_csv(path, sep=",", index_col=0, low_memory=False, dtype=str).drop(
            columns=["Non-unique Gene name", "Annotation"])
        got = nf.PanarooTable(path)
        vals = want.to_numpy(dtype=object)
        pres, blob = got.cells(np.arange(got.shape[0]), None)
        assert got.columns == [str(c) for c in want.columns] and got.shape == want.shape
        assert got.index == [str(i) for i in want.index]
        assert (pres == pd.notna(vals)).all()
        assert blob == b"\n".join(x.encode() for x in vals[pd.notna(vals)])
        compared += 1
    assert compared == 150


def test_native_feeder_fuzz_against_python_feeder(tmp_path):
    """Random GFF3 + FASTA text built from the quirks both parsers tolerate (short / long rows,
    int() oddities, any strand symbol, ID attributes in every form, comments, repeated and
    unnamed contigs, ragged / padded / blank sequence lines, both line ends): features, contigs
    and the cut sequences with their Seqinfo fields must be the Python feeder's."""
    import logging
    import random
    rnd = random.Random(5)
    ints = ["1", "12", "40", "99", "150"] * 3 + [" 7", "7 ", "1_0", "+5", "-3", "0", "abc", "", "1.0", "1__0", "_1", "99999"]
    strands = ["+", "-", "+", "-", ".", "?", "", "+ "]
    attrs = ["ID=g{}", "ID=g{};Name=x", "ID=g{};x=1", "ID=g{};x=2", "Name=y;ID=g{}", "ID=g{}=tail", "locus=1;ID=g{};", "IDX=o{}", "ID", "Name=z",
             "ID=g{};ID=h{}", "ID=", "id=g{}"]
    path = str(tmp_path / "g.gff")
    cuts = 0
    logging.disable(logging.WARNING)
    try:
        for _ in range(120):
            fasta_lines = []
            for c in range(rnd.randint(1, 3)):
                name = rnd.choice([f"c{c}", f"c{c} desc", f" c{c}", "c0", ""])
                seq = "".join(rnd.choice("ACGTacgtNnRY") for _ in range(rnd.randint(0, 300)))
                fasta_lines.append(">" + name)
                w, style, i = rnd.choice([10, 60, 70, 1000]), rnd.random(), 0
                while i < len(seq):
                    ww = w if style < 0.7 else rnd.randint(1, 80)
                    piece = seq[i:i + ww]
                    if rnd.random() < 0.05:
                        piece = " " + piece
                    if rnd.random() < 0.05:
                        piece = piece + " "
                    fasta_lines.append(piece)
                    if rnd.random() < 0.03:
                        fasta_lines.append("")
                    i += ww
            gff_lines = ["##gff-version 3"]
            for g in range(rnd.randint(0, 12)):
                cols = [rnd.choice(["c0", "c1", "c2", "cX"]), "src", rnd.choice(["CDS", "CDS", "gene", "cds"]),
                        rnd.choice(ints), rnd.choice(ints), ".", rnd.choice(strands), "0",
                        rnd.choice(attrs).format(g, g)]
                line = "\t".join((cols + ["extra"])[:rnd.choice([9, 9, 9, 9, 8, 5, 3, 2, 10])])
                if rnd.random() < 0.05:
                    line = "  # comment"
                if rnd.random() < 0.05:
                    line = "#" + line
                gff_lines.append(line)
            nl = rnd.choice(["\n", "\n", "\r\n"])
            text = nl.join(gff_lines) + nl + "##FASTA" + nl + nl.join(fasta_lines) + (nl if rnd.random() < 0.9 else "")
            with open(path, "wb") as fh:
                fh.write(text.encode())
            feats = pyin.parse_gff(path)
            pcont = pyin.read_fasta_text(open(path).read().split("##FASTA")[1].split("\n"))
            native = nf.NativeFeeder()
            g = native.add_genome("q", path)
            info = native.genome_info(g)
            got = {}
            for i in range(info["features"]):
                ident, contig, start, end, strand = native.feature(g, i)
                got[ident] = (contig, start, end, strand)
            assert got == {k: (v.chromosome, v.start, v.end, v.strand) for k, v in feats.items()}, text
            listed = [native.contig(g, i) for i in range(info["contigs"])]
            assert {x[0]: x[1] for x in listed} == {k: len(v) for k, v in pcont.items()}, text
            ids = [i for i in feats if "\n" not in i and ";" not in i]      # (what a table cell can name)
            if ids:
                table = pd.DataFrame({"q": [";".join(ids)]}, index=["cl"])
                for up, down, dsc in [(0, 0, False), (rnd.randint(0, 50), rnd.randint(0, 50), rnd.random() < 0.5)]:
                    py = list(pyin.iter_gene_clusters(table, {"q": (pcont, feats)}, up, down, dsc, True))
                    cut = native.cut(np.zeros(1, np.uint32) + g, ";".join(ids).encode(), up, down, dsc)
                    off = cut["seq_off"].astype(np.int64)
                    assert [cut["ascii"][off[i]:off[i + 1]] for i in range(cut["n_seqs"])] == \
                        [q.sequence.encode() for q in py[0][0]["q"]], text
                    assert [(q.start, q.end, q.offset, q.strand) for q in py[0][0]["q"]] == \
                        list(zip(cut["start"].tolist(), cut["end"].tolist(), cut["offset"].tolist(),
                                 cut["strand"].tolist())), text
                    cuts += cut["n_seqs"]
                    # the same windows packed on the spot (contigs in place or copied, two threads)
                    pk = native.cut(np.zeros(1, np.uint32) + g, ";".join(ids).encode(), up, down, dsc, prepack=True,
                                    n_threads=2)
                    ref = capi.pack_blob(cut["ascii"], cut["seq_off"])
                    assert (pk["packed"] == ref[0]).all() and (pk["base_off"] == ref[1]).all(), text
                    assert (pk["is_amb"] == ref[2]).all() and (pk["amb_off"] == ref[4]).all(), text
                    assert (pk["amb_plane"] is None) == (ref[3] is None), text
                    if ref[3] is not None:
                        assert (pk["amb_plane"] == ref[3]).all(), text
            native.close()
    finally:
        logging.disable(logging.NOTSET)
    assert cuts > 20
