"""Host-side logic of the multi-GPU path on CPU: cluster sharding and the
all-to-all choreography of the global pattern dedup, world_size 2 over gloo.
The four device primitives (pf_exchange_*) are replaced by a numpy model HERE
(test double only; the product has no CPU path) so that the collective plumbing,
id bases and remap tables can be checked without a GPU."""
import os
import subprocess
import sys

import numpy as np

from panfeed_b200 import dist as pfdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_clusters_balanced_and_complete():
    rng = np.random.default_rng(1)
    w = rng.integers(1, 1000, 200)
    for world in (1, 2, 3, 8):
        shards = pfdist.shard_clusters(w, world)
        assert sorted(c for s in shards for c in s) == list(range(200))
        loads = [int(w[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= int(w.max())
        assert all(s == sorted(s) for s in shards)


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from panfeed_b200 import dist as pfdist

class NumpyBackend:
    """Model of pf_exchange_pack/dedup/unpack on host arrays."""
    def __init__(self, pools, consider_missing):
        self.pools, self.consider_missing, self.perm, self.uniq = pools, consider_missing, {}, {}
    def key_words(self, ns): return self.pools[ns].shape[1]
    def n_local(self, ns): return self.pools[ns].shape[0]
    def _keys(self, ns, mask_remap):
        keys = self.pools[ns].copy()
        if mask_remap is not None:
            keys[:, -1] = mask_remap.numpy().astype(np.uint32)[keys[:, -1]]
        return keys
    def pack(self, ns, world, mask_remap, send):
        keys = self._keys(ns, mask_remap)
        owner = np.array([hash(k.tobytes()) % world for k in keys], dtype=np.int64) if len(keys) else np.zeros(0, np.int64)
        order = np.argsort(owner, kind="stable")
        perm = np.empty(len(keys), np.int64); perm[order] = np.arange(len(keys))
        self.perm[ns] = perm
        self.owner = getattr(self, "owner", {}); self.owner[ns] = owner
        send.copy_(torch.from_numpy(keys[order].astype(np.int32).reshape(send.shape)))
        return [int((owner == r).sum()) for r in range(world)]
    def dedup(self, ns, recv, unique_index, n_unique, keep_unique=False):
        keys = recv.numpy().astype(np.uint32)
        seen, idx = {}, []
        for k in keys:
            first = k.tobytes() not in seen
            u = seen.setdefault(k.tobytes(), len(seen))
            idx.append(u - (1 << 31) if first else u)        # bit 31 (int32): first copy = writer
        unique_index.copy_(torch.tensor(idx, dtype=torch.int32))
        self.uniq[ns] = np.array([np.frombuffer(b, np.uint32) for b in seen]).reshape(len(seen), keys.shape[1])
        n_unique.fill_(len(seen))
    def unique_keys(self, ns): return self.uniq[ns]
    def unpack(self, ns, returned, owner_base, l2g, writer):
        v = returned[torch.from_numpy(self.perm[ns])].numpy().astype(np.int64) & 0xffffffff
        base = owner_base.numpy().astype(np.int64)[self.owner[ns]]
        l2g.copy_(torch.from_numpy(((v & 0x7fffffff) + base).astype(np.int32)))
        if writer is not None:
            writer.copy_(torch.from_numpy((v >> 31).astype(np.uint8)))

os.environ["PYTHONHASHSEED"] = "0"
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cm = sys.argv[2] == "1"
rng = np.random.default_rng(7)            # same stream on both ranks
W = 3
universe_c = rng.integers(0, 2**32, (6, W), dtype=np.uint64).astype(np.uint32)
universe_k = rng.integers(0, 2**32, (40, W), dtype=np.uint64).astype(np.uint32)
pick = np.random.default_rng(100 + rank)
cl_local = universe_c[pick.choice(6, 4, replace=False)]
rows = pick.choice(40, 25, replace=False)
if cm:
    kp_local = np.concatenate([universe_k[rows], pick.integers(0, 4, (25, 1)).astype(np.uint32)], axis=1)
else:
    kp_local = universe_k[rows]
be = NumpyBackend({pfdist.CLUSTER: cl_local, pfdist.KMER: kp_local}, cm)
be.consider_missing = cm
ex = pfdist.PatternExchange(None, torch.device("cpu"), backend=be)
out = ex.run(want_unique=True, want_writer=True)
# gather everything on every rank and check global consistency
def gather(obj):
    lst = [None] * world
    dist.all_gather_object(lst, obj)
    return lst
cl_g = out["cluster"]["local_to_global"].numpy()
km_g = out["kmer"]["local_to_global"].numpy()
true_k = kp_local.copy()
if cm:
    true_k[:, -1] = cl_g[kp_local[:, -1]]
all_cl = gather((cl_local, cl_g, out["cluster"]["owned_keys"], out["cluster"]["owned_base"], out["cluster"]["writer"].numpy()))
all_km = gather((true_k, km_g, out["kmer"]["owned_keys"], out["kmer"]["owned_base"], out["kmer"]["writer"].numpy()))
for name, allx, total in (("cluster", all_cl, out["cluster"]["n_global"]), ("kmer", all_km, out["kmer"]["n_global"])):
    key_to_id = {}
    for keys, ids, _, _, _ in allx:
        for k, i in zip(keys, ids):
            assert key_to_id.setdefault(k.tobytes(), int(i)) == int(i), name + ": same key, two ids"
    assert len(set(key_to_id.values())) == len(key_to_id), name + ": two keys share an id"
    assert sorted(key_to_id.values()) == list(range(total)), name + ": ids not dense"
    # the owner's exported unique keys sit at owned_base + index
    for _, _, owned, base, _ in allx:
        for j, k in enumerate(owned):
            assert key_to_id[k.tobytes()] == base + j
    # every global pattern has exactly one writer over all ranks
    written = [int(i) for _, ids, _, _, wr in allx for i, w in zip(ids, wr) if w]
    assert sorted(written) == list(range(total)), name + ": writer flags"
print(f"rank {rank} ok cm={cm} clusters={out['cluster']['n_global']} kmers={out['kmer']['n_global']}")
dist.destroy_process_group()
'''


def _run(cm):
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(WORKER)
        path = f.name
    env = dict(os.environ, PYTHONHASHSEED="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29540 + cm), path, ROOT, str(cm)],
                       capture_output=True, text=True, timeout=300, env=env)
    os.unlink(path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok ") == 2, r.stdout


def test_exchange_world2_gloo():
    _run(0)


def test_exchange_world2_gloo_consider_missing():
    _run(1)
