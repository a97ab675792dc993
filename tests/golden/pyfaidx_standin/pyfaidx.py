"""Test-only stand-in for the un-vendored third-party `pyfaidx` package.

The reference (`/root/reference/panfeed/input.py:9`) imports pyfaidx at module
top; it is not installed in this image and there is no network.  This module
exposes only the surface the reference's call sites touch:

    Fasta(path, sequence_always_upper=True, rebuild=False)   input.py:262-271
    fa[chrom] -> record, KeyError if absent                  input.py:404-411
    rec[a:b]  -> sequence (python slice, truncating)          input.py:430-446
    -seq (reverse complement), seq[::-1], str(seq), len(seq)  input.py:434-455
    fa.close()                                                input.py:462

It is used ONLY by tests/golden/make_golden.py to run the unmodified reference
in the build container and is never imported by the product.
"""
import os

_COMP = str.maketrans("ACTGNactgnYRWSKMDVHBXyrwskmdvhbx",
                      "TGACNtgacnRYWSMKHBDVXrywsmkhbdvx")


class Sequence:
    def __init__(self, seq):
        self.seq = seq

    def __getitem__(self, n):
        if isinstance(n, slice):
            return Sequence(self.seq[n])
        return Sequence(self.seq[n])

    def __neg__(self):
        return Sequence(self.seq.translate(_COMP)[::-1])

    def __str__(self):
        return self.seq

    def __len__(self):
        return len(self.seq)


class Fasta:
    def __init__(self, filename, sequence_always_upper=False, rebuild=True,
                 **_kw):
        self.filename = filename
        self._records = {}
        name, chunks = None, []
        with open(filename) as fh:
            for line in fh:
                line = line.rstrip("\n").rstrip("\r")
                if line.startswith(">"):
                    if name is not None:
                        self._records[name] = "".join(chunks)
                    name = line[1:].split()[0] if line[1:].split() else ""
                    chunks = []
                elif name is not None:
                    chunks.append(line.strip())
        if name is not None:
            self._records[name] = "".join(chunks)
        if sequence_always_upper:
            self._records = {k: v.upper() for k, v in self._records.items()}
        fai = filename + ".fai"
        if rebuild or not os.path.exists(fai):
            with open(fai, "w") as out:
                for k, v in self._records.items():
                    out.write(f"{k}\t{len(v)}\t0\t0\t0\n")

    def __getitem__(self, name):
        return Sequence(self._records[name])

    def keys(self):
        return self._records.keys()

    def close(self):
        pass
