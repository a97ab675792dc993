// K3 (partition mode) — per-k-mer sample presence from PARTIALLY sorted records.
//
// After `p` radix passes the records of a cluster are ordered by the leading
// 8p bits of the mixed key only.  Instead of spending more global passes, one
// CTA takes all prefix-runs that START inside its 2048-record tile (a handful
// of buckets, typically a few thousand records and ~100 distinct k-mers),
// streams their records once through a shared-memory open-addressing table
// keyed on the FULL 64-bit key, and does the whole reduction on chip:
//
//   phase 1  find-or-insert the key (64-bit CAS in shared memory), count its
//            distinct samples.  Records of one key arrive in ascending sample
//            order (stable passes, sample-ordered packing), so across 2048-record
//            chunks only "equal to the last sample seen" can repeat; inside a
//            chunk a small (slot, sample) pair set removes duplicates exactly.
//   phase 2  keys whose count lies in the cluster's integer MAF window get a row
//            (one atomicAdd per CTA on the global row counter).
//   phase 3  the records are streamed again (L2-resident) and the surviving
//            keys' sample bits are OR-ed into shared-memory bitsets, as many rows
//            per round as fit, then written out coalesced with the row's
//            (cluster, un-mixed k-mer, count).
//
// Replaces `cluster_dict[kmer][sortstrain[strain]] = 1` and the filters of
// /root/reference/panfeed/panfeed.py:77-88,190-204.  Exact for any input: if a
// CTA meets more distinct keys than its table holds it raises a flag and the
// host re-runs the batch with 8 more sorted bits (at 64 bits a tile can hold
// at most 2048 + 1 distinct keys, which always fits).
#pragma once
#include "pf_common.cuh"
#include "k3_reduce.cuh"

namespace pf {

constexpr int kLocalThreads = 256;
constexpr int kLocalItems = 8;
constexpr int kLocalTile = kLocalThreads * kLocalItems;   // 2048 records
constexpr int kLocalSlots = 4096;                         // >= 2 * kLocalTile
constexpr uint32_t kLocalMaxUnique = 3072;
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint32_t kEmptyPair = 0xffffffffu;
constexpr uint32_t kNoRow = 0xffffffffu;
constexpr uint32_t kLocalMaxSamples = 1u << 19;           // (slot:13 | sample:19) pair word
constexpr int kLocalPoolWords = kLocalSlots + 1 + kLocalSlots;

struct LocalSmem {
  uint64_t keys[kLocalSlots + 1];      // slot kLocalSlots: home of the key equal to kEmptyKey
  uint32_t cnt[kLocalSlots + 1];       // distinct samples
  uint32_t last_prev[kLocalSlots + 1]; // 1 + largest sample before this chunk; phase 2+: row index
  uint32_t last_next[kLocalSlots + 1]; // 1 + largest sample in this chunk   } phase 3: bitset pool
  uint32_t pairs[kLocalSlots];         // chunk-local (slot, sample) set       }
  uint32_t n_unique, n_pass, row_base, special_used, overflow, ok;
};

enum { LC_ROWS = 0, LC_UNIQUE = 1, LC_TABLE_OVERFLOW = 2, LC_ROW_OVERFLOW = 3 };

__device__ __forceinline__ uint32_t local_find_or_insert(LocalSmem& sm, uint64_t key) {
  if (key == kEmptyKey) {
    if (atomicExch(&sm.special_used, 1u) == 0u) atomicAdd(&sm.n_unique, 1u);
    return kLocalSlots;
  }
  uint32_t h = (uint32_t)key & (kLocalSlots - 1);
  for (int probes = 0; probes < kLocalSlots; ++probes) {
    const uint64_t cur = sm.keys[h];
    if (cur == key) return h;
    if (cur == kEmptyKey) {
      const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long*>(&sm.keys[h]),
                                     (unsigned long long)kEmptyKey, (unsigned long long)key);
      if (old == kEmptyKey) { atomicAdd(&sm.n_unique, 1u); return h; }
      if (old == key) return h;
    }
    h = (h + 1) & (kLocalSlots - 1);
  }
  sm.overflow = 1;
  return kNoRow;
}

__device__ __forceinline__ uint32_t local_find(const LocalSmem& sm, uint64_t key) {
  if (key == kEmptyKey) return kLocalSlots;
  uint32_t h = (uint32_t)key & (kLocalSlots - 1);
  while (sm.keys[h] != key) h = (h + 1) & (kLocalSlots - 1);
  return h;
}

__global__ void __launch_bounds__(kLocalThreads, 2)
k3_local(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
         const TileDev* __restrict__ ltiles, uint32_t n_ltiles,
         const uint32_t* __restrict__ tile_first_run /* [n_ltiles + 1] */,
         const uint32_t* __restrict__ run_start, uint32_t n_records,
         const ClusterDev* __restrict__ clusters, RowOut out, uint32_t row_capacity,
         uint32_t* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char local_raw[];
  LocalSmem& sm = *reinterpret_cast<LocalSmem*>(local_raw);
  const uint32_t tid = threadIdx.x;
  const uint32_t t = blockIdx.x;
  const uint32_t r0 = tile_first_run[t], r1 = tile_first_run[t + 1];
  if (r0 == r1) return;                       // no prefix-run starts in this tile
  const uint32_t n_runs = tile_first_run[n_ltiles];
  const uint32_t a = run_start[r0];
  const uint32_t b = (r1 < n_runs) ? run_start[r1] : n_records;
  const uint32_t seg = ltiles[t].seg;
  const ClusterDev cl = clusters[seg];
  const uint32_t W = out.pattern_words;

  for (uint32_t i = tid; i <= kLocalSlots; i += kLocalThreads) {
    sm.keys[i] = kEmptyKey;
    sm.cnt[i] = 0;
    sm.last_prev[i] = 0;
    sm.last_next[i] = 0;
  }
  if (tid == 0) { sm.n_unique = 0; sm.n_pass = 0; sm.special_used = 0; sm.overflow = 0; sm.ok = 1; }

  // ---- phase 1: group by full key, count distinct samples ---------------------
  for (uint32_t c0 = a; c0 < b; c0 += kLocalTile) {
    for (uint32_t i = tid; i < kLocalSlots; i += kLocalThreads) {
      sm.pairs[i] = kEmptyPair;
      if (c0 > a) {
        const uint32_t n = sm.last_next[i];
        if (n > sm.last_prev[i]) sm.last_prev[i] = n;
      }
    }
    if (tid == 0 && c0 > a && sm.last_next[kLocalSlots] > sm.last_prev[kLocalSlots])
      sm.last_prev[kLocalSlots] = sm.last_next[kLocalSlots];
    __syncthreads();
    uint64_t key[kLocalItems];
    uint32_t val[kLocalItems];
#pragma unroll
    for (int j = 0; j < kLocalItems; ++j) {
      const uint32_t i = c0 + j * kLocalThreads + tid;
      val[j] = kInvalidSample;
      key[j] = 0;
      if (i < b) { key[j] = keys[i]; val[j] = vals[i]; }
    }
#pragma unroll
    for (int j = 0; j < kLocalItems; ++j) {
      const uint32_t v = val[j];
      if (v == kInvalidSample) continue;
      const uint32_t slot = local_find_or_insert(sm, key[j]);
      if (slot == kNoRow) continue;
      const uint32_t e = (slot << 19) | v;
      uint32_t h = (e * 0x9e3779b1u) >> 20;              // 12 bits
      bool fresh = false;
      for (;;) {
        const uint32_t old = atomicCAS(&sm.pairs[h], kEmptyPair, e);
        if (old == kEmptyPair) { fresh = true; break; }
        if (old == e) break;
        h = (h + 1) & (kLocalSlots - 1);
      }
      if (fresh) {
        if (sm.last_prev[slot] != v + 1u) atomicAdd(&sm.cnt[slot], 1u);
        atomicMax(&sm.last_next[slot], v + 1u);
      }
    }
    __syncthreads();
    if (sm.overflow || sm.n_unique > kLocalMaxUnique) {
      if (tid == 0) atomicExch(&counters[LC_TABLE_OVERFLOW], 1u);
      return;
    }
  }

  // ---- phase 2: which keys survive, row allocation -----------------------------
  for (uint32_t i = tid; i <= kLocalSlots; i += kLocalThreads) {
    const bool used = (i < kLocalSlots) ? (sm.keys[i] != kEmptyKey) : (sm.special_used != 0u);
    uint32_t row = kNoRow;
    if (used) {
      const uint32_t c = sm.cnt[i];
      if (c >= cl.lo && c <= cl.hi) row = atomicAdd(&sm.n_pass, 1u);
    }
    sm.last_prev[i] = row;
  }
  __syncthreads();
  if (tid == 0) {
    atomicAdd(&counters[LC_UNIQUE], sm.n_unique);
    if (sm.n_pass) {
      const uint32_t base = atomicAdd(&counters[LC_ROWS], sm.n_pass);
      sm.row_base = base;
      if ((uint64_t)base + sm.n_pass > row_capacity) {
        sm.ok = 0;
        atomicExch(&counters[LC_ROW_OVERFLOW], 1u);
      }
    }
  }
  __syncthreads();
  const uint32_t n_pass = sm.n_pass;
  if (n_pass == 0 || !sm.ok) return;
  const uint32_t base = sm.row_base;
  for (uint32_t i = tid; i <= kLocalSlots; i += kLocalThreads) {
    const uint32_t row = sm.last_prev[i];
    if (row != kNoRow) {
      const size_t g = (size_t)base + row;
      out.cluster[g] = cl.id;
      out.kmer[g] = unmix64(i < kLocalSlots ? sm.keys[i] : kEmptyKey);
      out.count[g] = sm.cnt[i];
      if (out.key_words > W) out.cand[g * out.key_words + W] = out.cluster_pattern[seg];
    }
  }

  // ---- phase 3: bitsets of the surviving keys, `per_round` rows at a time ----------
  uint32_t* pool = sm.last_next;                 // last_next + pairs are contiguous
  const uint32_t per_round = max(1u, (uint32_t)kLocalPoolWords / W);
  for (uint32_t lo = 0; lo < n_pass; lo += per_round) {
    const uint32_t rows = min(per_round, n_pass - lo);
    __syncthreads();
    for (uint32_t i = tid; i < rows * W; i += kLocalThreads) pool[i] = 0;
    __syncthreads();
    for (uint32_t c0 = a; c0 < b; c0 += kLocalTile) {
      uint64_t key[kLocalItems];
      uint32_t val[kLocalItems];
#pragma unroll
      for (int j = 0; j < kLocalItems; ++j) {
        const uint32_t i = c0 + j * kLocalThreads + tid;
        val[j] = kInvalidSample;
        key[j] = 0;
        if (i < b) { key[j] = keys[i]; val[j] = vals[i]; }
      }
#pragma unroll
      for (int j = 0; j < kLocalItems; ++j) {
        const uint32_t v = val[j];
        if (v == kInvalidSample) continue;
        const uint32_t row = sm.last_prev[local_find(sm, key[j])];
        if (row - lo < rows)                      // also false for kNoRow
          atomicOr(&pool[(row - lo) * W + (v >> 5)], 1u << (v & 31u));
      }
    }
    __syncthreads();
    for (uint32_t i = tid; i < rows * W; i += kLocalThreads) {
      const uint32_t r = i / W, w = i - r * W;
      out.cand[((size_t)base + lo + r) * out.key_words + w] = pool[i];
    }
  }
}

}  // namespace pf
