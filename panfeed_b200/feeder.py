"""Native feeder: GFF3/FASTA parsing and per-cluster sequence cutting in the library
(`pf_feeder_*`, csrc/pf_feeder.cu) instead of the Python loops of `input.py`.

`PanarooTable` stands where the `genepres` DataFrame does (reference input.py:198-201),
`prep_feeder` where `input.prep_data_n_fasta` does (input.py:68-138) and `iter_packed_batches` /
`iter_packed_clusters` where `iter_gene_clusters` + `cluster_cutter` do (input.py:335-468,
panfeed.py:23-113): they yield the 4-tuples `pattern_hasher` consumes - whole GPU batches cut and
packed by one library call (`NativePackedBatch`, the CLI's default), or one `NativePackedCluster`
per panaroo row (`--multiple-files`).  `prefetch` runs either on a background thread.  The
Python path (`--python-feeder`) stays the behavioural reference; `tests/test_feeder_native.py`
checks that both produce identical batches on the fixture genomes, on hand-written quirks and
on random GFF3 / FASTA / table text.
"""
import ctypes as C
import logging
import os
import sys

import numpy as np
import pandas as pd

from . import capi
from .input import _genome_name, _listing

logger = logging.getLogger("panfeed.input")


class NativeFeeder:
    """Genomes parsed once by the library; `cut` returns the sequences of one cluster."""

    def __init__(self):
        self.lib = capi.load()
        self.h = C.c_void_p()
        rc = self.lib.pf_feeder_create(C.byref(self.h))
        if rc != 0:
            raise capi.PfError(rc, "pf_feeder_create failed")
        self.names = []
        self._feature_names = {}

    def close(self):
        if self.h:
            self.lib.pf_feeder_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _error(self, rc, what):
        msg = self.lib.pf_feeder_last_error(self.h)
        raise capi.PfError(rc, f"{what}: {msg.decode() if msg else ''}")

    def add_genome(self, name, gff_path, fasta_path=None):
        """-> genome index.  Malformed GFF lines are skipped like the reference does
        (input.py:326-330), with one warning per file."""
        skipped = C.c_uint32()
        rc = self.lib.pf_feeder_add_genome(self.h, name.encode(), os.fsencode(gff_path),
                                           None if fasta_path is None else os.fsencode(fasta_path),
                                           C.byref(skipped))
        if rc < 0:
            self._error(rc, f"reading {gff_path}")
        if skipped.value:
            logger.warning(f"skipped {skipped.value} malformed feature lines of {gff_path}")
        self.names.append(name)
        return rc

    def add_genomes(self, names, gff_paths, fasta_paths=None, n_threads=0):
        """Several genomes, read and parsed by the library's host threads -> index of the first."""
        n = len(names)
        arr = C.c_char_p * n
        c_names = arr(*[x.encode() for x in names])
        c_gff = arr(*[os.fsencode(x) for x in gff_paths])
        c_fa = None
        if fasta_paths is not None and any(x is not None for x in fasta_paths):
            c_fa = arr(*[None if x is None else os.fsencode(x) for x in fasta_paths])
        skipped = np.zeros(max(n, 1), np.uint32)
        rc = self.lib.pf_feeder_add_genomes(self.h, n, c_names, c_gff, c_fa, skipped.ctypes.data, int(n_threads))
        if rc < 0:
            self._error(rc, "reading the genomes")
        for i in np.flatnonzero(skipped[:n]):
            logger.warning(f"skipped {skipped[i]} malformed feature lines of {gff_paths[i]}")
        self.names.extend(names)
        return rc

    def add_genome_text(self, name, gff_text, fasta_text=None):
        gff = gff_text.encode() if isinstance(gff_text, str) else gff_text
        fasta = fasta_text.encode() if isinstance(fasta_text, str) else fasta_text
        skipped = C.c_uint32()
        rc = self.lib.pf_feeder_add_genome_text(self.h, name.encode(), gff, len(gff), fasta,
                                                0 if fasta is None else len(fasta), C.byref(skipped))
        if rc < 0:
            self._error(rc, f"parsing genome {name}")
        self.names.append(name)
        return rc

    def genome_info(self, genome):
        nf, ncg, nb = C.c_uint32(), C.c_uint32(), C.c_uint64()
        rc = self.lib.pf_feeder_genome_info(self.h, genome, C.byref(nf), C.byref(ncg), C.byref(nb))
        if rc != 0:
            self._error(rc, "pf_feeder_genome_info")
        return {"features": nf.value, "contigs": ncg.value, "bases": nb.value}

    def contig(self, genome, contig):
        """-> (name, number of bases, used in place) of a parsed contig."""
        name, n, in_place = C.c_char_p(), C.c_uint64(), C.c_uint32()
        rc = self.lib.pf_feeder_contig(self.h, genome, contig, C.byref(name), C.byref(n), C.byref(in_place))
        if rc != 0:
            self._error(rc, "pf_feeder_contig")
        return name.value.decode(), n.value, bool(in_place.value)

    def feature(self, genome, feature):
        """-> (id, contig, start, end, strand) of a parsed feature."""
        ident, contig = C.c_char_p(), C.c_char_p()
        start, end, strand = C.c_int64(), C.c_int64(), C.c_int32()
        rc = self.lib.pf_feeder_feature(self.h, genome, feature, C.byref(ident), C.byref(contig),
                                        C.byref(start), C.byref(end), C.byref(strand))
        if rc != 0:
            self._error(rc, "pf_feeder_feature")
        return ident.value.decode(), contig.value.decode(), start.value, end.value, strand.value

    def feature_names(self, genome, feature):
        key = (genome, feature)
        got = self._feature_names.get(key)
        if got is None:
            got = self._feature_names[key] = self.feature(genome, feature)[:2]
        return got

    def cut(self, genomes, cells_blob, up, down, down_start_codon, prepack=False, n_threads=0):
        """genomes: uint32 array (genome index per cell); cells_blob: the cells joined with
        newlines (bytes).  -> dict of copies of the pf_cut_result arrays.  prepack: the sequences
        go from the contigs straight into the 2-bit / 4-bit planes (pf_feeder_cut_packed, host
        threads: "packed", "base_off", "is_amb", "amb_plane", "amb_off" as capi.pack_blob would
        return them for the ASCII cut) and there is no ASCII text ("ascii" is None)."""
        genomes = np.ascontiguousarray(genomes, dtype=np.uint32)
        res = capi.CutResult()
        planes = capi.CutPlanes()
        if prepack:
            rc = self.lib.pf_feeder_cut_packed(self.h, len(genomes), genomes.ctypes.data, cells_blob, len(cells_blob),
                                               int(up), int(down), int(bool(down_start_codon)), int(n_threads),
                                               C.byref(res), C.byref(planes))
            if rc == -4 and planes.bad_symbol:
                raise ValueError(f"unsupported sequence symbol {chr(planes.bad_symbol)!r}: only "
                                 f"{capi.AMB_ALPHABET} (IUPAC, upper case) are accepted")
        else:
            rc = self.lib.pf_feeder_cut(self.h, len(genomes), genomes.ctypes.data, cells_blob, len(cells_blob),
                                        int(up), int(down), int(bool(down_start_codon)), C.byref(res))
        if rc != 0:
            self._error(rc, "pf_feeder_cut")
        n = res.n_seqs

        def arr(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

        seq_off = arr(res.seq_off, n + 1, np.uint64) if n else np.zeros(1, np.uint64)
        out = {"n_seqs": n, "seq_off": seq_off,
               "cell": arr(res.cell, n, np.uint32), "feature": arr(res.feature, n, np.uint32),
               "start": arr(res.start, n, np.int32), "end": arr(res.end, n, np.int32),
               "offset": arr(res.offset, n, np.int32), "strand": arr(res.strand, n, np.int32),
               "missing": []}
        if prepack:
            out["ascii"] = None
            out["packed"] = arr(planes.packed, int(planes.n_words), np.uint64)
            out["base_off"] = arr(planes.base_off, n, np.uint64)
            out["is_amb"] = arr(planes.is_amb, n, np.uint8).astype(bool)
            out["amb_plane"] = arr(planes.amb_plane, int(planes.n_amb_words), np.uint64) if planes.n_amb_words else None
            out["amb_off"] = arr(planes.amb_off, n, np.uint64)
        else:
            out["ascii"] = C.string_at(res.ascii, int(seq_off[-1])) if n else b""
        if res.n_missing:
            moff = arr(res.missing_off, res.n_missing + 1, np.uint64)
            text = C.string_at(res.missing_text, int(moff[-1]))
            mcell = arr(res.missing_cell, res.n_missing, np.uint32)
            mkind = arr(res.missing_kind, res.n_missing, np.uint8)
            out["missing"] = [(int(mcell[i]), int(mkind[i]), text[int(moff[i]):int(moff[i + 1])].decode())
                              for i in range(res.n_missing)]
        return out


class NativePackedCluster:
    """The attributes of `packer.PackedCluster`, filled from a native cut whose sequences were
    packed on the spot: `packed_words` / `amb_words` are this cluster's slices of the 2-bit / 4-bit
    planes (every sequence starts on a 64-base boundary, so a cluster is a whole number of words),
    `base_rel` / `amb_rel` the offsets inside them.  `seq_bytes` (decoded from the planes) and
    `meta` are built only if somebody asks."""
    __slots__ = ("idx", "clusterpresab", "packed_words", "base_rel", "is_amb", "amb_words", "amb_rel",
                 "seq_len", "sample", "target", "start", "end", "offset", "strand", "k", "canonical",
                 "consider_missing", "_feeder", "_genome", "_feature", "_order", "_meta")

    def __init__(self, idx, clusterpresab, packed_words, base_rel, is_amb, amb_words, amb_rel, seq_len,
                 sample, target, start, end, offset, strand, genome, feature, order, feeder):
        self.idx = idx
        self.clusterpresab = clusterpresab
        self.packed_words, self.base_rel = packed_words, base_rel
        self.is_amb, self.amb_words, self.amb_rel = is_amb, amb_words, amb_rel
        self.seq_len = seq_len
        self.sample = sample
        self.target = target
        self.start, self.end, self.offset, self.strand = start, end, offset, strand
        self._feeder, self._genome, self._feature, self._order = feeder, genome, feature, order
        self._meta = None

    @property
    def seq_bytes(self):
        """The sequences as ASCII, decoded from the planes (tests, debugging)."""
        acgt = np.frombuffer(b"ACGT", np.uint8)
        amb = np.frombuffer(capi.AMB_ALPHABET.encode(), np.uint8)
        sh2 = (62 - 2 * np.arange(32)).astype(np.uint64)
        sh4 = (60 - 4 * np.arange(16)).astype(np.uint64)
        out = []
        for i, n in enumerate(self.seq_len.tolist()):
            if self.is_amb[i]:
                w0 = int(self.amb_rel[i]) // 16
                words = self.amb_words[w0:w0 + (n + 15) // 16]
                codes = ((words[:, None] >> sh4[None, :]) & np.uint64(15)).astype(np.uint8).ravel()[:n]
                out.append(amb[codes].tobytes())
            else:
                w0 = int(self.base_rel[i]) // 32
                words = self.packed_words[w0:w0 + (n + 31) // 32]
                codes = ((words[:, None] >> sh2[None, :]) & np.uint64(3)).astype(np.uint8).ravel()[:n]
                out.append(acgt[codes].tobytes())
        return out

    @property
    def meta(self):
        """[(strain, feature id, contig)] per sequence (for the kmers.tsv rows)."""
        if self._meta is None:
            self._meta = [(self._order[int(self.sample[i])],) +
                          self._feeder.feature_names(int(self._genome[i]), int(self._feature[i]))
                          for i in range(len(self._feature))]
        return self._meta

    def n_records(self, k, canonical):
        n = int(np.maximum(self.seq_len - k + 1, 0).sum())
        return n if canonical else 2 * n


def prep_feeder(filelist, fastalist, gffdir, fastadir, output=None):
    """-> (NativeFeeder, {genome name: genome index}).  Same file discovery as
    `input.prep_data_n_fasta` (reference input.py:68-138); the nucleotides stay in the library."""
    gff_files, gff_is_list = _listing(gffdir)
    if gff_is_list:
        gff_path = {_genome_name(f): f for f in gff_files if f.endswith(".gff")}
    else:
        gff_path = {g: os.path.join(gffdir, f"{g}.gff") for g in filelist}
    fasta_path = {}
    if fastadir is not None:
        files, is_list = _listing(fastadir)
        for f in files:
            if f.endswith(".fna") or f.endswith(".fasta"):
                fasta_path[_genome_name(f)] = f if is_list else os.path.join(fastadir, f)
    feeder = NativeFeeder()
    gffs, fastas = [], []
    for genome in filelist:
        logger.debug(f"Handling {genome}")
        fasta = None
        if genome in fastalist:
            if genome not in fasta_path:
                logger.error(f"Neither {genome}.fna not {genome}.fasta found in {fastadir}")
                sys.exit(1)
            fasta = fasta_path[genome]
        gffs.append(gff_path[genome])
        fastas.append(fasta)
    first = feeder.add_genomes(list(filelist), gffs, fastas)
    index = {genome: first + i for i, genome in enumerate(filelist)}
    return feeder, index


def _na_values():
    """The strings pandas reads as a missing value (read_csv's default na_values)."""
    try:
        from pandas._libs.parsers import STR_NA_VALUES
        return sorted(STR_NA_VALUES)
    except Exception:      # noqa: BLE001 - a pandas without that private name: its documented defaults
        return ["", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>",
                "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null"]


class PanarooTable:
    """panaroo's gene_presence_absence.csv read by the library (pf_table_*): stands where the
    reference's `genepres` DataFrame does (pd.read_csv(..., index_col=0).drop(columns=[...]),
    input.py:198-201) for the native feeder - `columns`, `index`, a row subset (`take`), the
    cells of a run of rows - without a Python object per cell (pandas needs ~0.5 us and ~60 bytes
    for each; BASELINE configs #4 / #5 hold 5e7 / 4e8 of them)."""

    DROP = ("Non-unique Gene name", "Annotation")

    def __init__(self, path=None, _parent=None, _rows=None):
        if _parent is not None:
            self._owner, self.lib, self.h = _parent._owner, _parent.lib, _parent.h
            self.columns, self._labels = _parent.columns, _parent._labels
            self._rows = np.asarray(_rows, np.uint64)
            return
        self._owner = self
        self.lib = capi.load()
        self.h = C.c_void_p()
        if self.lib.pf_table_create(C.byref(self.h)) != 0:
            raise capi.PfError(-1, "pf_table_create failed")
        drop = (C.c_char_p * len(self.DROP))(*[d.encode() for d in self.DROP])
        na = [x.encode() for x in _na_values()]
        rc = self.lib.pf_table_load(self.h, os.fsencode(path), drop, len(self.DROP), (C.c_char_p * len(na))(*na),
                                    len(na), 0)
        if rc != 0:
            msg = self.lib.pf_table_last_error(self.h).decode()
            if msg.startswith("column not found"):
                raise KeyError(msg)                  # what DataFrame.drop raises
            raise capi.PfError(rc, f"reading {path}: {msg}")
        n_rows, n_cols = C.c_uint64(), C.c_uint32()
        self.lib.pf_table_shape(self.h, C.byref(n_rows), C.byref(n_cols))
        self.columns = self._names(0, n_cols.value)
        self._labels = self._names(1, n_rows.value)
        self._rows = np.arange(n_rows.value, dtype=np.uint64)

    def _names(self, row_labels, n):
        blob, off = C.c_void_p(), C.c_void_p()
        self.lib.pf_table_names(self.h, row_labels, C.byref(blob), C.byref(off))
        if n == 0:
            return []
        o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint64)), shape=(n + 1,))
        text = C.string_at(blob, int(o[-1]))
        return [text[int(o[i]):int(o[i + 1])].decode() for i in range(n)]

    def __del__(self):
        try:
            if self._owner is self and self.h:
                self.lib.pf_table_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:      # noqa: BLE001
            pass

    @property
    def index(self):
        return [self._labels[int(r)] for r in self._rows]

    @property
    def shape(self):
        return (len(self._rows), len(self.columns))

    def take(self, positions):
        """The rows at `positions` (of this view), like DataFrame.iloc[positions]."""
        return PanarooTable(_parent=self, _rows=self._rows[np.asarray(positions, np.int64)])

    def n_present(self):
        """Number of present cells of every row (DataFrame.notna().sum(axis=1))."""
        n_all = np.zeros(len(self._labels), np.uint32)
        if len(n_all):
            self.lib.pf_table_row_counts(self.h, n_all.ctypes.data)
        return n_all[self._rows.astype(np.int64)].astype(np.int64)

    def cells(self, positions, col_order=None):
        """(present [n, n_cols] bool, the present cells row by row joined with newlines) of the
        rows at `positions`, columns in the order `col_order`."""
        rows = np.ascontiguousarray(self._rows[np.asarray(positions, np.int64)], dtype=np.uint64)
        order = None if col_order is None else np.ascontiguousarray(col_order, dtype=np.uint32)
        S = len(self.columns)
        need, n_cells = C.c_uint64(), C.c_uint64()
        args = (self.h, rows.ctypes.data, len(rows), None if order is None else order.ctypes.data)
        rc = self.lib.pf_table_cells(*args, None, None, 0, C.byref(need), C.byref(n_cells))
        if rc != 0:
            raise capi.PfError(rc, "pf_table_cells (sizing) failed")
        present = np.zeros((len(rows), S), np.uint8)
        blob = C.create_string_buffer(max(1, int(need.value)))
        rc = self.lib.pf_table_cells(*args, present.ctypes.data, blob, int(need.value), C.byref(need), C.byref(n_cells))
        if rc != 0:
            raise capi.PfError(rc, "pf_table_cells failed")
        return present.astype(bool), blob.raw[:int(need.value)]


def prefetch(iterator, depth=2):
    """Runs `iterator` on a background thread, `depth` items ahead of the consumer: the library
    cuts and packs the next GPU batch (GIL released inside the calls) while the current one is on
    the device and its rows are formatted and written.  Exceptions of the producer are re-raised
    at the consumer, in order."""
    import queue
    import threading
    q = queue.Queue(maxsize=max(1, int(depth)))
    done = object()
    stop = threading.Event()

    def put(item):
        while not stop.is_set():
            try:
                q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def work():
        try:
            for item in iterator:
                if not put((item, None)):
                    return
            put((done, None))
        except BaseException as exc:      # noqa: BLE001 - handed to the consumer
            put((done, exc))

    th = threading.Thread(target=work, name="pf-feeder", daemon=True)
    th.start()
    try:
        while True:
            item, exc = q.get()
            if exc is not None:
                raise exc
            if item is done:
                return
            yield item
    finally:
        stop.set()                        # the consumer left early (error downstream): the producer finishes
        th.join()                         # the cut it is in, sees the flag and goes; the feeder outlives it


class _CutPlan:
    """What both iterators share: the panaroo table as arrays in sample-rank order, the rows to cut
    and the library call that cuts a run of them."""

    def __init__(self, panaroo, feeder, genome_index, up, down, down_start_codon, stroi, gene_list, raise_missing):
        cols = [str(c) for c in panaroo.columns]
        missing = set(cols).difference(genome_index.keys())
        if missing:
            raise KeyError(f"{len(missing)} strains of the pangenome table have no GFF "
                           f"(e.g. {sorted(missing)[0]}); the reference would emit misaligned "
                           "presence vectors here, this build refuses")
        self.feeder, self.up, self.down, self.dsc = feeder, up, down, down_start_codon
        self.raise_missing = raise_missing
        self.order = sorted(cols)
        rank = {s: i for i, s in enumerate(self.order)}
        perm = np.argsort(np.array([rank[c] for c in cols]), kind="stable")        # table columns in rank order
        self.genome_of_rank = np.array([genome_index[s] for s in self.order], np.uint32)
        self.target_of_rank = np.array([s in stroi for s in self.order], bool)
        self.perm = perm
        self.table = panaroo if isinstance(panaroo, PanarooTable) else None
        if self.table is not None:
            self.n_rows, self.S = panaroo.shape
            all_cells = panaroo.n_present()
        else:
            self.values = panaroo.to_numpy(dtype=object)[:, perm]
            self.present_all = pd.notna(self.values)
            self.n_rows, self.S = self.values.shape
            all_cells = self.present_all.sum(axis=1)
        self.index = list(panaroo.index)
        rows = []
        for i, idx in enumerate(self.index):
            if gene_list is not None and idx not in gene_list:
                logger.debug(f"Skipping {idx} ({i + 1}/{self.n_rows})")
                continue
            rows.append(i)
        self.rows = np.array(rows, np.int64)
        self.n_cells = all_cells[self.rows] if len(self.rows) else np.zeros(0, np.int64)

    def take(self, at, cell_budget):
        """-> `to`: rows at .. to hold at least one cluster, then as many as fit the cell budget."""
        to, budget = at + 1, cell_budget - int(self.n_cells[at])
        while to < len(self.rows) and budget - int(self.n_cells[to]) >= 0:
            budget -= int(self.n_cells[to])
            to += 1
        return to

    def cut(self, at, to):
        """Cut rows at .. to -> (sel, pres, rr, cc, cut): the table rows, their presence matrix, the
        (cluster, rank) of every present cell in row-major order, the library's result."""
        sel = self.rows[at:to]
        if self.table is not None:
            pres, blob = self.table.cells(sel, self.perm)
            rr, cc = np.nonzero(pres)                # row-major: cluster by cluster, ranks ascending
        else:
            pres = self.present_all[sel]
            rr, cc = np.nonzero(pres)
            cells = self.values[sel][pres]
            blob = "\n".join(cells).encode() if len(cells) else b""
        cut = self.feeder.cut(self.genome_of_rank[cc], blob, self.up, self.down, self.dsc, prepack=True)
        for cell, kind, name in cut["missing"]:
            idx, strain = self.index[sel[rr[cell]]], self.order[cc[cell]]
            msg = (f"Could not find gene {name} from {idx} in {strain}" if kind == 0
                   else f"Could not find chromosome {name} in {strain}")
            logger.warning(msg)
            if self.raise_missing:
                raise KeyError(msg)
        return sel, pres, rr, cc, cut


def iter_packed_clusters(panaroo, feeder, genome_index, up, down, down_start_codon, stroi, klength,
                         canon, consider_missing_cluster, gene_list=None, raise_missing=False,
                         cells_per_call=16384):
    """Yields what `cluster_cutter(iter_gene_clusters(...))` yields — (idx, packed cluster,
    int presence vector, None) per panaroo row — with the cutting done by the library, several
    clusters (about `cells_per_call` table cells) per call: the Python work per cluster is a
    handful of array slices."""
    plan = _CutPlan(panaroo, feeder, genome_index, up, down, down_start_codon, stroi, gene_list, raise_missing)
    index, order, n_rows = plan.index, plan.order, plan.n_rows
    at = 0
    while at < len(plan.rows):
        to = plan.take(at, cells_per_call)
        sel, pres, rr, cc, cut = plan.cut(at, to)
        seq_cluster = rr[cut["cell"]]
        sample = cc[cut["cell"]].astype(np.uint32)
        target = plan.target_of_rank[sample]
        genome = plan.genome_of_rank[sample]
        seq_len = np.diff(cut["seq_off"]).astype(np.int64)
        first = np.searchsorted(seq_cluster, np.arange(len(sel) + 1))
        # word ranges of the clusters in the planes of this cut
        packed, base_off, is_amb = cut["packed"], cut["base_off"].astype(np.int64), cut["is_amb"]
        word_at = np.concatenate([base_off // 32, [len(packed)]]).astype(np.int64)
        amb_plane = cut["amb_plane"] if cut["amb_plane"] is not None else np.zeros(0, np.uint64)
        padded = (seq_len + 63) // 64 * 64
        amb_sym_at = np.concatenate([[0], np.cumsum(np.where(is_amb, padded, 0))]).astype(np.int64)
        amb_off = cut["amb_off"].astype(np.int64)
        for j, i in enumerate(sel):
            idx = index[i]
            logger.debug(f"Extracting sequences from {idx} ({i + 1}/{n_rows})")
            a, b = int(first[j]), int(first[j + 1])
            clusterpresab = pres[j].astype(int)
            amb_c = is_amb[a:b]
            pc = NativePackedCluster(
                idx, clusterpresab, packed[int(word_at[a]):int(word_at[b])], base_off[a:b] - (base_off[a] if b > a else 0),
                amb_c, amb_plane[int(amb_sym_at[a]) // 16:int(amb_sym_at[b]) // 16],
                np.where(amb_c, amb_off[a:b] - amb_sym_at[a], 0), seq_len[a:b], sample[a:b], target[a:b],
                cut["start"][a:b], cut["end"][a:b], cut["offset"][a:b], cut["strand"][a:b], genome[a:b],
                cut["feature"][a:b], order, feeder)
            pc.k, pc.canonical, pc.consider_missing = klength, bool(canon), bool(consider_missing_cluster)
            yield idx, pc, clusterpresab, None
        at = to


class NativePackedBatch:
    """A run of clusters cut by the library, already in the form of a pf_batch: `hb` is what
    `packer.pack_batch` makes of the same clusters (the planes of the cut ARE the planes of the
    batch), `idxs` their panaroo row names.  `pattern_hasher` submits it as it is: no Python
    object per cluster, no second copy of the planes.  (strain, feature id, contig) of a sequence
    are looked up only when a kmers.tsv row needs them (`meta`)."""
    __slots__ = ("hb", "idxs", "k", "canonical", "consider_missing", "_feeder", "_genome", "_feature", "_order")

    def __init__(self, hb, idxs, genome, feature, order, feeder):
        self.hb, self.idxs = hb, idxs
        self._feeder, self._genome, self._feature, self._order = feeder, genome, feature, order

    def meta(self, i):
        return (self._order[int(self.hb.seqs["sample"][i])],) + \
            self._feeder.feature_names(int(self._genome[i]), int(self._feature[i]))

    def n_records(self, k, canonical):
        n = int(np.maximum(self.hb.seqs["len"].astype(np.int64) - k + 1, 0).sum())
        return n if canonical else 2 * n


def iter_packed_batches(panaroo, feeder, genome_index, up, down, down_start_codon, stroi, klength,
                        canon, consider_missing_cluster, gene_list=None, raise_missing=False,
                        target_bases=96 * 1024 * 1024, first_cells=16384):
    """The batch-level form of `iter_packed_clusters`: yields (idxs, NativePackedBatch, None, None),
    one library call and a few array operations per BATCH of clusters.  The first call cuts about
    `first_cells` table cells; from its bases per cell the following calls are sized to about
    `target_bases` bases (whole clusters, at least one)."""
    plan = _CutPlan(panaroo, feeder, genome_index, up, down, down_start_codon, stroi, gene_list, raise_missing)
    W = (plan.S + 31) // 32
    at, cells_per_call = 0, int(first_cells)
    seen_cells = seen_bases = 0
    while at < len(plan.rows):
        to = plan.take(at, cells_per_call)
        sel, pres, rr, cc, cut = plan.cut(at, to)
        n = cut["n_seqs"]
        cell = cut["cell"]
        seqs = np.zeros(n, capi.SEQ_DTYPE)
        sample = cc[cell].astype(np.uint32)
        seq_len = np.diff(cut["seq_off"]).astype(np.uint32) if n else np.zeros(0, np.uint32)
        seqs["base_off"] = cut["base_off"]
        seqs["len"] = seq_len
        seqs["cluster"] = rr[cell]
        seqs["sample"] = sample
        seqs["flags"] = (plan.target_of_rank[sample].astype(np.uint32) * capi.PF_SEQ_TARGET |
                         cut["is_amb"].astype(np.uint32) * capi.PF_SEQ_AMBIGUOUS)
        for f in ("start", "end", "offset", "strand"):
            seqs[f] = cut[f]
        seqs["amb_off"] = cut["amb_off"]
        clusters = np.zeros(len(sel), capi.CLUSTER_DTYPE)
        clusters["id"] = np.arange(len(sel), dtype=np.uint32)
        bits = np.zeros((len(sel), W * 32), np.uint8)
        bits[:, :plan.S] = pres
        presence = np.packbits(bits, axis=1, bitorder="little").view(np.uint32).reshape(len(sel), W)
        hb = capi.HostBatch(cut["packed"], seqs, clusters, presence, cut["amb_plane"])
        idxs = [plan.index[i] for i in sel]
        logger.debug(f"Extracted sequences of {len(idxs)} clusters ({idxs[0]} ..; {to}/{len(plan.rows)})")
        pb = NativePackedBatch(hb, idxs, plan.genome_of_rank[sample], cut["feature"], plan.order, feeder)
        pb.k, pb.canonical, pb.consider_missing = klength, bool(canon), bool(consider_missing_cluster)
        yield idxs, pb, None, None
        seen_cells += int(pres.sum())
        seen_bases += int(seq_len.sum())
        if seen_bases:
            cells_per_call = int(min(max(first_cells, target_bases * seen_cells // seen_bases), 1 << 22))
        at = to
