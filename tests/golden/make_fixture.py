#!/usr/bin/env python
"""Deterministic toy pangenome used by the parity tests.

The reference's own fixture (`tests/test_files.tar.gz`) is listed in
`/root/reference/.MISSING_LARGE_BLOBS`, so this script writes a stand-in with
the layout `tests/unit_test.sh:18-52` expects:

    fixture/gffs/<genome>.gff            GFF3 + ##FASTA
    fixture/fastas/<genome>.fasta        same nucleotides (file-of-files mode)
    fixture/gene_presence_absence.csv    panaroo table
    fixture/stroi.txt                    --targets
    fixture/genes.txt                    --genes (second pass)
    fixture/genes_acgt.txt               --genes: the clusters without N/IUPAC symbols (k > 32 modes)
    fixture/input_gffs.txt, input_fastas.txt

It deliberately covers: both strands, paralogs (';'), a refound gene whose ID
is not in any GFF, genes whose flanks run over the contig start on either
strand and over the contig end, an N run and IUPAC codes, lower-case bases,
a gene shorter than k, a singleton cluster, CSV columns not in sorted order,
a CDS line without ID, a non-CDS feature, a malformed line.
"""
import os
import random
import sys

COMP = str.maketrans("ACGTN", "TGCAN")


def revcomp(s):
    return s.translate(COMP)[::-1]


def rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def mutate(rng, s, rate):
    out = list(s)
    for i, c in enumerate(out):
        if rng.random() < rate:
            out[i] = rng.choice([b for b in "ACGT" if b != c])
    return "".join(out)


def build(outdir, seed=20261018):
    rng = random.Random(seed)
    genomes = [f"s{i:02d}" for i in range(8)]
    # CSV column order is NOT the sorted order (input.py:368-369 sorts)
    csv_order = ["s03", "s00", "s07", "s01", "s05", "s02", "s06", "s04"]

    # cluster -> ancestral length, presence probability
    clusters = [
        ("group_core", 240, 1.0),
        ("group_acc1", 180, 0.6),
        ("group_acc2", 210, 0.4),
        ("group_para", 150, 0.9),
        ("group_short", 25, 0.7),     # shorter than k=31 without flanks
        ("group_single", 120, 0.0),   # forced singleton below
        ("group_edge", 160, 1.0),     # always first/last on a contig
    ]
    anc = {c: rand_seq(rng, L) for c, L, _ in clusters}
    founders = {c: [mutate(rng, anc[c], 0.02) for _ in range(3)] for c in anc}

    cells = {c: {} for c in anc}      # cluster -> genome -> [gene ids]
    gff_dir = os.path.join(outdir, "gffs")
    fa_dir = os.path.join(outdir, "fastas")
    os.makedirs(gff_dir)
    os.makedirs(fa_dir)

    for gi, g in enumerate(genomes):
        genes = []                    # (cluster, seq)
        for c, L, p in clusters:
            if c == "group_single":
                present = (g == "s05")
            elif c == "group_core" and g == "s06":
                present = False       # core gene missing from one genome
            else:
                present = rng.random() < p
            if not present:
                continue
            s = mutate(rng, rng.choice(founders[c]), 0.004)
            genes.append((c, s))
            if c == "group_para" and rng.random() < 0.5:
                genes.append((c, mutate(rng, rng.choice(founders[c]), 0.01)))
        # the edge cluster goes first on contig 1 and last on contig 2
        edge = [x for x in genes if x[0] == "group_edge"]
        rest = [x for x in genes if x[0] != "group_edge"]
        rng.shuffle(rest)
        half = len(rest) // 2
        contigs = [edge[:1] + rest[:half], rest[half:]]
        if gi % 2 == 1 and edge:
            contigs = [rest[:half], rest[half:] + edge[:1]]

        gff_lines = ["##gff-version 3"]
        fasta = []
        counter = 0
        for ci, members in enumerate(contigs):
            cname = f"{g}_ctg{ci + 1}"
            seq = ""
            for mi, (c, s) in enumerate(members):
                if c == "group_edge":
                    spacer = rng.randint(3, 40)     # < upstream/downstream 100
                else:
                    spacer = rng.randint(20, 160)
                if not (c == "group_edge" and mi == len(members) - 1):
                    seq += rand_seq(rng, spacer)
                elif mi > 0:
                    seq += rand_seq(rng, rng.randint(60, 160))
                strand = rng.choice("+-")
                start = len(seq) + 1
                seq += s if strand == "+" else revcomp(s)
                end = len(seq)
                gid = f"{g}_{counter:05d}"
                counter += 1
                gff_lines.append(f"{cname}\tsynth\tgene\t{start}\t{end}\t.\t"
                                 f"{strand}\t.\tID=gene_{gid}")
                gff_lines.append(f"{cname}\tsynth\tCDS\t{start}\t{end}\t.\t"
                                 f"{strand}\t0\tID={gid};product=x")
                cells[c].setdefault(g, []).append(gid)
            # trailing sequence: short after an edge gene that is last
            last_edge = members and members[-1][0] == "group_edge"
            seq += rand_seq(rng, rng.randint(5, 30) if last_edge
                            else rng.randint(40, 170))
            # decorations: N run / IUPAC / lowercase inside genome s02, s04
            if g == "s02" and ci == 0 and len(seq) > 200:
                seq = seq[:150] + "NNNNN" + seq[155:]
                seq = seq[:260] + "R" + seq[261:]
            if g == "s04" and ci == 1 and len(seq) > 120:
                seq = seq[:60] + seq[60:120].lower() + seq[120:]
            fasta.append((cname, seq))
        # odd lines the parser must survive (input.py:296-330)
        gff_lines.append(f"{g}_ctg1\tsynth\tCDS\t5\t20\t.\t+\t0\tproduct=noid")
        gff_lines.append("# a comment")
        if g == "s01":
            gff_lines.append(f"{g}_ctg1\tsynth\tCDS\tnotanint\t20\t.\t+\t0\tID=bad")
        with open(os.path.join(gff_dir, f"{g}.gff"), "w") as fh:
            fh.write("\n".join(gff_lines) + "\n##FASTA\n")
            for cname, seq in fasta:
                fh.write(f">{cname} len={len(seq)}\n")
                for i in range(0, len(seq), 60):
                    fh.write(seq[i:i + 60] + "\n")
        with open(os.path.join(fa_dir, f"{g}.fasta"), "w") as fh:
            for cname, seq in fasta:
                fh.write(f">{cname}\n")
                for i in range(0, len(seq), 70):
                    fh.write(seq[i:i + 70] + "\n")

    # a refound gene (ID absent from the GFF), input.py:395-402
    cells["group_acc1"].setdefault("s07", []).append("s07_refound_1")
    with open(os.path.join(outdir, "gene_presence_absence.csv"), "w") as fh:
        fh.write("Gene,Non-unique Gene name,Annotation," +
                 ",".join(csv_order) + "\n")
        for c, _, _ in clusters:
            row = [c, "", "hypothetical protein"]
            for g in csv_order:
                row.append(";".join(cells[c].get(g, [])))
            fh.write(",".join(row) + "\n")
    with open(os.path.join(outdir, "stroi.txt"), "w") as fh:
        fh.write("s01\ns02\ns05\n")
    with open(os.path.join(outdir, "genes.txt"), "w") as fh:
        fh.write("group_acc1\ngroup_para\ngroup_edge\n")
    # the clusters without N/IUPAC symbols (k > 32 runs: two-word k-mers take plain bases only)
    with open(os.path.join(outdir, "genes_acgt.txt"), "w") as fh:
        fh.write("group_core\ngroup_acc1\ngroup_acc2\ngroup_short\ngroup_single\n")
    with open(os.path.join(outdir, "input_gffs.txt"), "w") as fh:
        for g in genomes:
            fh.write(os.path.join("fixture", "gffs", f"{g}.gff") + "\n")
    with open(os.path.join(outdir, "input_fastas.txt"), "w") as fh:
        for g in genomes:
            fh.write(os.path.join("fixture", "fastas", f"{g}.fasta") + "\n")


if __name__ == "__main__":
    build(sys.argv[1] if len(sys.argv) > 1 else
          os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixture"))
