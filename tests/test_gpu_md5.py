"""K5: device-side MD5 pattern ids equal hashlib on the reference's byte images,
for every message-length residue (padding) and both namespaces."""
import binascii
import hashlib

import numpy as np
import pytest

import helpers
from panfeed_b200 import capi, packer
from oracle import ref_port

pytestmark = pytest.mark.gpu


def _id(v):
    return binascii.b2a_base64(hashlib.md5(np.ascontiguousarray(v).view(np.uint8)).digest()).decode()[:24]


@pytest.mark.parametrize("cm", [False, True])
@pytest.mark.parametrize("S", [1, 6, 7, 8, 9, 12, 31, 32, 33, 63, 64, 65, 500, 1003])
def test_pattern_ids_match_hashlib(S, cm):
    rng = np.random.default_rng(S)
    comp = str.maketrans("ACGT", "TGCA")
    names = [f"g{i:04d}" for i in range(S)]
    items = []
    for c in range(3):
        anc = rng.choice(list("ACGT"), 90)
        presab = np.zeros(S, dtype=int)
        cluster = {}
        for i, s in enumerate(names):
            if S > 1 and rng.random() < 0.3:
                cluster[s] = []
                continue
            presab[i] = 1
            q = anc.copy()
            m = rng.random(90) < 0.05
            q[m] = rng.choice(list("ACGT"), int(m.sum()))
            q = "".join(q)
            cluster[s] = [ref_port.CutSeq(q, q.translate(comp), "x", "c", 1, 90, 1, 0)]
        if presab.sum() == 0:
            presab[0] = 1
            q = "".join(anc)
            cluster[names[0]] = [ref_port.CutSeq(q, q.translate(comp), "x", "c", 1, 90, 1, 0)]
        items.append((cluster, f"c{c}", presab))
    pcs = [packer.PackedCluster(c, idx, pa, set()) for c, idx, pa in items]
    hb, _, _ = packer.pack_batch(pcs)
    ctx = capi.Context(21, S, True, cm, False, False, 0.0)
    ctx.submit(hb)
    r = ctx.collect()
    W = (S + 31) // 32
    idx = np.arange(S)
    cids = ctx.pattern_ids(True, 0, len(r["new_cluster_patterns"]))
    for row, got in zip(r["new_cluster_patterns"], cids):
        v = ((row[idx >> 5] >> (idx & 31)) & 1).astype(np.int64)
        assert got.decode() == _id(v)
    kids = ctx.pattern_ids(False, 0, len(r["new_kmer_patterns"]))
    assert len(kids) > 0
    for row, got in zip(r["new_kmer_patterns"], kids):
        v = ((row[idx >> 5] >> (idx & 31)) & 1).astype(np.float64)
        if cm:
            pres = r["new_cluster_patterns"][row[W]]
            v[((pres[idx >> 5] >> (idx & 31)) & 1) == 0] = np.nan
        assert got.decode() == _id(v)
    ctx.close()


def test_known_answer_ids():
    """The S=12 known answers produced by the reference (hot_kats.json)."""
    kats = {k["desc"]: k for k in helpers.hot_kats()["md5_ids"]}
    assert _id(np.ones(12, dtype=np.int64)) == kats["int64 all-ones S=12"]["id"]
