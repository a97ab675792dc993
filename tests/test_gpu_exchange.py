"""Device side of the global pattern dedup (pf_exchange_*), world = 2 emulated
on ONE GPU: two contexts own disjoint cluster shards; the buffers an NCCL
all-to-all would move are routed by hand.  The result must number patterns
exactly like one context that saw every cluster."""
import numpy as np
import pytest
import torch

import gpu_util  # noqa: F401  (path setup)
from panfeed_b200 import capi, dist as pfdist, packer
from test_gpu_parity import _random_items

pytestmark = pytest.mark.gpu


def _submit(ctx, items, stroi, base):
    pcs = [packer.PackedCluster(c, idx, pa, stroi) for c, idx, pa in items]
    hb, _, _ = packer.pack_batch(pcs, list(range(base, base + len(items))))
    ctx.submit(hb)
    return ctx.collect()


@pytest.mark.parametrize("cm", [False, True])
def test_exchange_two_contexts_one_gpu(cm):
    dev = torch.device("cuda", 0)
    S, k = 48, 21
    rng = np.random.default_rng(11)
    items, stroi = _random_items(rng, S, k, 6, 260)
    shards = [items[0::2], items[1::2]]
    ctxs = [capi.Context(k, S, True, cm, False, False, 0.02) for _ in range(2)]
    res = [_submit(c, sh, set(), 100 * i) for i, (c, sh) in enumerate(zip(ctxs, shards))]
    bes = [pfdist.DeviceBackend(c, dev) for c in ctxs]
    world = 2
    l2g = {}
    for ns in (pfdist.CLUSTER, pfdist.KMER):
        kw = bes[0].key_words(ns)
        sends, counts = [], []
        for r, be in enumerate(bes):
            n = be.n_local(ns)
            send = torch.empty((n, kw), dtype=torch.int32, device=dev)
            remap = l2g[(pfdist.CLUSTER, r)] if (cm and ns == pfdist.KMER) else None
            counts.append(be.pack(ns, world, remap, send))
            sends.append(send)
        # route: owner o receives, from every rank r, rows [off_r[o], off_r[o] + counts[r][o])
        recvs = []
        for o in range(world):
            parts = []
            for r in range(world):
                off = sum(counts[r][:o])
                parts.append(sends[r][off:off + counts[r][o]])
            recvs.append(torch.cat(parts).contiguous())
        uniq, n_unique = [], []
        for o, be in enumerate(bes):
            u = torch.empty(recvs[o].shape[0], dtype=torch.int32, device=dev)
            nu = torch.zeros(1, dtype=torch.int32, device=dev)
            be.dedup(ns, recvs[o], u, nu, True)    # asynchronous on the context's stream
            torch.cuda.synchronize()
            n_unique.append(int(nu.cpu().item()))
            uniq.append(u)
        owner_base = torch.tensor([0, n_unique[0]], dtype=torch.int32, device=dev)
        writers = []
        for r, be in enumerate(bes):
            back = []
            for o in range(world):
                off = sum(counts[q][o] for q in range(r))
                back.append(uniq[o][off:off + counts[r][o]])
            returned = torch.cat(back).contiguous()
            out = torch.empty(be.n_local(ns), dtype=torch.int32, device=dev)
            wr = torch.empty(be.n_local(ns), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            be.unpack(ns, returned, owner_base, out, wr)
            torch.cuda.synchronize()
            l2g[(ns, r)] = out
            writers.append(wr.cpu().numpy())
        l2g[(ns, "total")] = sum(n_unique)
        l2g[(ns, "owned")] = [be.unique_keys(ns) for be in bes]
        # exactly one writer per global pattern
        written = np.concatenate([l2g[(ns, r)].cpu().numpy()[writers[r] != 0] for r in range(world)])
        assert sorted(written.tolist()) == list(range(sum(n_unique)))

    # reference: one context over all clusters
    one = capi.Context(k, S, True, cm, False, False, 0.02)
    r_all = _submit(one, items, set(), 0)
    assert l2g[(pfdist.CLUSTER, "total")] == len(r_all["new_cluster_patterns"])
    assert l2g[(pfdist.KMER, "total")] == len(r_all["new_kmer_patterns"])
    W = (S + 31) // 32
    # global ids are consistent: same full (ternary) pattern <-> same id
    seen = {}
    for r, ctx in enumerate(ctxs):
        kp = ctx.export_patterns(False, 0, bes[r].n_local(pfdist.KMER))
        cp = ctx.export_patterns(True, 0, bes[r].n_local(pfdist.CLUSTER))
        g = l2g[(pfdist.KMER, r)].cpu().numpy()
        for row, gid in zip(kp, g):
            key = row[:W].tobytes() + (cp[row[W]].tobytes() if cm else b"")
            assert seen.setdefault(key, int(gid)) == int(gid)
    assert len(set(seen.values())) == len(seen) == l2g[(pfdist.KMER, "total")]
    want = set()
    cp_all = r_all["new_cluster_patterns"]
    for row in r_all["new_kmer_patterns"]:
        want.add(row[:W].tobytes() + (cp_all[row[W]].tobytes() if cm else b""))
    assert set(seen) == want
    for c in ctxs + [one]:
        c.close()
