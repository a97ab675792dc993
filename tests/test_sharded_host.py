"""Host-side pieces of the sharded (torchrun) product path that need no GPU: the per-rank piece
files and their merge (plain and --compress, where every piece is a run of gzip members), and
the incrementally grown id table of the PatternStore."""
import gzip
import os

import numpy as np
import pytest

from panfeed_b200 import input as pfin
from panfeed_b200.panfeed import _IdTable


@pytest.mark.parametrize("compress", [False, True])
def test_piece_files_merge_into_the_three_outputs(tmp_path, compress):
    out = str(tmp_path / "out")
    os.mkdir(out)
    world = 3
    want = {name: [] for name in pfin.OUTPUT_NAMES}
    for rank in range(world):
        handles = pfin.create_part_files(out, rank, compress)
        for name, h in zip(pfin.OUTPUT_NAMES, handles):
            text = "".join(f"{name}\trank{rank}\trow{i}\n" for i in range(0 if rank == 1 and name == "kmers.tsv" else 50 + rank))
            h.write(text)
            want[name].append(text)
            h.close()
    headers = ("H1\tx\n", "H2\ty\n", "H3\tz\n")
    pfin.merge_part_files(out, world, headers, compress)
    assert not os.path.exists(os.path.join(out, ".parts"))
    for name, head in zip(pfin.OUTPUT_NAMES, headers):
        path = os.path.join(out, name + (".gz" if compress else ""))
        got = gzip.open(path, "rt").read() if compress else open(path).read()
        assert got == head + "".join(want[name])
    assert sorted(os.listdir(out)) == sorted(n + (".gz" if compress else "") for n in pfin.OUTPUT_NAMES)


def test_id_table_grows_without_rebuilding():
    t = _IdTable()
    rng = np.random.default_rng(0)
    ref = []
    for n in (0, 3, 1000, 1, 5000):
        ids = np.array([("%022d==" % int(x)).encode() for x in rng.integers(0, 10**15, n)], dtype="S24")
        t.append(ids)
        ref += ids.tolist()
        v = t.view()
        assert len(v) == len(ref) and v.dtype == np.dtype("S24")
    assert t.view().tolist() == ref
