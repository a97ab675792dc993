"""Multi-GPU: cluster sharding and the global pattern dedup exchange
(SURVEY.md §8(e)).

Gene clusters are independent for K1-K3 and for the rows of kmers_to_hashes /
kmers.tsv, so whole clusters are sharded over ranks with no data-path
collective.  The only global state of the reference is the `patterns` set its
single writer owns (`/root/reference/panfeed/__main__.py:70`,
`panfeed.py:210-212`); here every rank dedups locally (K4) and one exchange
makes the numbering global:

    owner(pattern) = hash(full key) % world                (x_classify kernel)
    all-to-all(v) of the full keys to their owners         (NCCL, torch.distributed)
    owner dedups on the full key, numbers its uniques       (k4_probe + scan + x_finish)
    all-gather of the unique counts -> global id = base[owner] + unique index
    reverse all-to-all of the ids, scatter to local order   (x_unpack kernel)

With --consider-missing a k-mer pattern key ends with the id of the cluster
pattern that gives its NaN plane, so cluster patterns are exchanged first and the
k-mer keys are rewritten with the GLOBAL cluster-pattern id before hashing.

torch is plumbing here (device buffers + the collective); the kernels are in
libpanfeed_b200.so behind pf_exchange_*.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

KMER, CLUSTER = 0, 1


def shard_clusters(weights, world):
    """Greedy longest-processing-time partition of clusters by weight (sum of
    sequence lengths).  -> list (per rank) of ascending cluster indices."""
    weights = np.asarray(weights, dtype=np.int64)
    order = np.argsort(-weights, kind="stable")
    load = np.zeros(world, np.int64)
    shards = [[] for _ in range(world)]
    for c in order:
        r = int(np.argmin(load))
        shards[r].append(int(c))
        load[r] += int(weights[c])
    return [sorted(s) for s in shards]


class DeviceBackend:
    """pf_exchange_* on one pf_ctx; buffers are torch tensors on its device."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device

    def key_words(self, ns):
        return self.ctx.W if ns == CLUSTER else self.ctx.Wk

    def n_local(self, ns):
        st = self.ctx.stats()
        return int(st["cluster_patterns"] if ns == CLUSTER else st["kmer_patterns"])

    def pack(self, ns, world, mask_remap, send):
        counts = (C.c_uint64 * world)()
        self.ctx._check(self.ctx.lib.pf_exchange_pack(
            self.ctx.h, ns, world, mask_remap.data_ptr() if mask_remap is not None else None,
            send.data_ptr() if send.numel() else None, send.shape[0], counts))
        return [int(c) for c in counts]

    def dedup(self, ns, recv, unique_index):
        n_unique = C.c_uint64()
        self.ctx._check(self.ctx.lib.pf_exchange_dedup(
            self.ctx.h, ns, recv.data_ptr() if recv.numel() else None, recv.shape[0],
            unique_index.data_ptr() if unique_index.numel() else None, C.byref(n_unique)))
        return int(n_unique.value)

    def unique_keys(self, ns, n_unique):
        out = np.zeros((n_unique, self.key_words(ns)), np.uint32)
        self.ctx._check(self.ctx.lib.pf_exchange_unique_export(self.ctx.h, ns, out.ctypes.data))
        return out

    def unpack(self, ns, returned, local_to_global):
        self.ctx._check(self.ctx.lib.pf_exchange_unpack(
            self.ctx.h, ns, returned.data_ptr() if returned.numel() else None,
            local_to_global.data_ptr() if local_to_global.numel() else None))


class PatternExchange:
    """Global numbering of the patterns of all ranks.  `run()` returns, per
    namespace, the local->global id table and this rank's owned unique keys."""

    def __init__(self, ctx, device, backend=None, group=None):
        self.backend = backend if backend is not None else DeviceBackend(ctx, device)
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.consider_missing = bool(getattr(ctx, "consider_missing", False)) if ctx is not None \
            else bool(getattr(backend, "consider_missing", False))

    def _exchange(self, ns, mask_remap):
        import time
        t0 = time.perf_counter()
        be, world, dev = self.backend, self.world, self.device
        kw = be.key_words(ns)
        n_local = be.n_local(ns)
        send = torch.empty((n_local, kw), dtype=torch.int32, device=dev)
        send_counts = be.pack(ns, world, mask_remap, send)
        t1 = time.perf_counter()
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = [int(x) for x in rc.tolist()]
        n_recv = sum(recv_counts)
        recv = torch.empty((n_recv, kw), dtype=torch.int32, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts,
                               input_split_sizes=send_counts, group=self.group)
        uniq_idx = torch.empty(n_recv, dtype=torch.int32, device=dev)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)      # NCCL ran on torch's stream; the library uses its own
        t2 = time.perf_counter()
        n_unique = be.dedup(ns, recv, uniq_idx)
        t3 = time.perf_counter()
        nu = torch.tensor([n_unique], dtype=torch.int64, device=dev)
        all_nu = [torch.empty_like(nu) for _ in range(world)]
        dist.all_gather(all_nu, nu, group=self.group)
        counts = [int(x) for x in torch.cat(all_nu).tolist()]      # one read-back, not one per rank
        base = sum(counts[:self.rank])
        uniq_idx += base
        returned = torch.empty(n_local, dtype=torch.int32, device=dev)
        dist.all_to_all_single(returned, uniq_idx, output_split_sizes=send_counts,
                               input_split_sizes=recv_counts, group=self.group)
        l2g = torch.empty(n_local, dtype=torch.int32, device=dev)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t4 = time.perf_counter()
        be.unpack(ns, returned, l2g)
        t5 = time.perf_counter()
        return {"local_to_global": l2g, "n_global": sum(counts), "n_owned": n_unique,
                "owned_base": base, "bytes_sent": int(send.numel() * 4),
                "ms": {"pack": (t1 - t0) * 1e3, "a2a_keys": (t2 - t1) * 1e3, "dedup": (t3 - t2) * 1e3,
                       "a2a_ids": (t4 - t3) * 1e3, "unpack": (t5 - t4) * 1e3}}

    def run(self, want_unique=False):
        cl = self._exchange(CLUSTER, None)
        remap = cl["local_to_global"] if self.consider_missing else None
        km = self._exchange(KMER, remap)
        out = {"cluster": cl, "kmer": km}
        if want_unique:
            cl["owned_keys"] = self.backend.unique_keys(CLUSTER, cl["n_owned"])
            km["owned_keys"] = self.backend.unique_keys(KMER, km["n_owned"])
        return out
