// Block aggregation (S <= 1024): K1 + K2 + K3 without ever materialising a record.
//
// The sequences of one gene cluster are near-identical copies of one gene, so
// the windows that START inside the same short position block [p0, p0 + B) of
// every sequence of the cluster hold only a few hundred distinct k-mers between
// them (one per haplotype and position), however many samples there are.
//
//   kA_block_aggregate  one CTA per (cluster, position block).  Every thread walks
//                       runs of 16 consecutive windows of one sequence with a ROLLING
//                       forward / reverse-complement k-mer (two funnel shifts each per
//                       window instead of a fresh extraction), looks the canonical
//                       k-mer up in a shared-memory open-addressing table and ORs the
//                       sample's bit into that k-mer's W-word bitset, also in shared
//                       memory.  Lanes of a warp work on DIFFERENT sequences at the
//                       SAME positions: their table probes hit the same slot (a
//                       broadcast) and their bits fall into different words.  The CTA
//                       ends by writing its distinct k-mers and bitsets ("partials")
//                       to its own slab in HBM: ~1/30 of the bytes the records would
//                       have taken, written once.
//   kB_merge            a k-mer can start in two blocks (indels, clamped flanks,
//                       paralogs), so partials of one cluster are merged by FULL key:
//                       the key space of the cluster is cut into G hash groups small
//                       enough for one CTA; CTA (cluster, g) scans all partial keys of
//                       the cluster (L2-resident), keeps those of its group, ORs their
//                       bitsets in shared memory, then counts, applies the integer MAF
//                       window and emits rows exactly like k3_local_direct.
//
// Exact for any input: nothing depends on the sequences being aligned — alignment
// only decides how few partials there are.  A block holding more distinct k-mers
// than the shared-memory table takes raises a flag and the host reruns the batch
// through the record path (k2_extract_scatter + k3_local_direct).
//
// Shared-memory table of both kernels: `slots` 64-bit keys (open addressing, linear
// probing, claimed with one 64-bit atomicCAS) and one W-word bitset PER SLOT, so a
// lookup is one LDS.64 and the OR goes straight to slot * stride: no dense-id
// indirection on the hot path.  The all-ones word marks an empty slot; it is a valid
// k-mer only for k = 32 in --non-canonical mode, which stays on the record path.
//
// Replaces the window loop and `cluster_dict[kmer][sortstrain[strain]] = 1` of
// /root/reference/panfeed/panfeed.py:54-88 and the filters of :190-204.
#pragma once
#include "pf_common.cuh"
#include "k1_extract.cuh"
#include "k3_local.cuh"

namespace pf {

struct SeqLite {           // 16 bytes: what kA needs of a sequence, one 128-bit load
  uint32_t word_off;       // first 64-bit word in the 2-bit plane
  uint32_t len;
  uint32_t sample_flags;   // sample rank | (ambiguous ? 1u << 31 : 0)
  uint32_t amb_word_off;   // first 32-bit word in the ambiguity bit plane
};
struct ClusterBlk {        // 16 bytes
  uint32_t seq_start, n_seqs, max_nwin, reserved;
};

constexpr int kBlkThreads = 256;
constexpr int kBlkWarps = kBlkThreads / 32;
constexpr int kBlkRun = 16;                       // windows per task
constexpr uint32_t kBlkOverflow = 0xffffffffu;    // slab count of a block that did not fit

__global__ void plan_seq_lite(const SeqDev* __restrict__ seqs, uint32_t n, SeqLite* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SeqDev d = seqs[i];
  SeqLite l;
  l.word_off = (uint32_t)(d.base_off >> 5);
  l.len = d.len;
  l.sample_flags = d.sample | ((d.flags & 2u) ? 0x80000000u : 0u);
  l.amb_word_off = (uint32_t)(d.amb_off >> 5);
  out[i] = l;
}

// one warp per cluster: its sequence range (sequences are sorted by cluster), the
// longest window count and from it the number of position blocks
__global__ void plan_cluster_blocks(const SeqDev* __restrict__ seqs, uint32_t n_seqs, uint32_t n_clusters,
                                    int k, uint32_t block_windows, ClusterBlk* __restrict__ cb,
                                    uint32_t* __restrict__ n_items) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_clusters) return;
  const uint32_t lane = lane_id();
  auto lower = [&](uint32_t v) {              // first sequence with cluster >= v
    uint32_t lo = 0, hi = n_seqs;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (seqs[mid].cluster < v) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const uint32_t s0 = lower(c), s1 = lower(c + 1);
  uint32_t mx = 0;
  for (uint32_t s = s0 + lane; s < s1; s += 32) {
    const uint32_t len = seqs[s].len;
    if (len >= (uint32_t)k) mx = max(mx, len - (uint32_t)k + 1u);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) mx = max(mx, __shfl_xor_sync(kFull, mx, m));
  if (lane == 0) {
    ClusterBlk b;
    b.seq_start = s0; b.n_seqs = s1 - s0; b.max_nwin = mx; b.reserved = 0;
    cb[c] = b;
    n_items[c] = (mx + block_windows - 1) / block_windows;
  }
}

// ---------------------------------------------------------------------------
// shared-memory table
// ---------------------------------------------------------------------------
struct BlkHead {           // 32 bytes, then keys[slots], list[slots] (u16), pool[(slots + 1) * WS]
                           // (row `slots` is a scratch row: lookups of a full table land there)
  uint32_t n_unique, overflow, n_pass, row_base, ok, work, pad[2];
};
__host__ __device__ inline uint32_t blk_smem_bytes(uint32_t slots, uint32_t W) {
  return (uint32_t)sizeof(BlkHead) + slots * 8u + slots * 2u + (slots + 1u) * (W | 1u) * 4u;
}
struct BlkView {
  BlkHead* h;
  uint64_t* keys;
  uint16_t* list;
  uint32_t* pool;
  uint32_t mask, shift;    // slots - 1, 32 - log2(slots)
};
__device__ __forceinline__ BlkView blk_view(unsigned char* raw, uint32_t slots) {
  BlkView v;
  v.h = reinterpret_cast<BlkHead*>(raw);
  v.keys = reinterpret_cast<uint64_t*>(raw + sizeof(BlkHead));
  v.list = reinterpret_cast<uint16_t*>(v.keys + slots);
  v.pool = reinterpret_cast<uint32_t*>(v.list + slots);
  v.mask = slots - 1u;
  v.shift = 32u - (uint32_t)__popc(v.mask);
  return v;
}
__device__ __forceinline__ uint32_t blk_hash(uint32_t kh, uint32_t kl, uint32_t shift) {
  return ((kl ^ (kh * 0x85ebca6bu)) * 0x9e3779b1u) >> shift;
}

// Collision chain of the lookup (the home slot holds another key): linear probing from h.
// Returns the slot (the scratch row `slots` if the chain is too long; overflow is then flagged).
__device__ __noinline__ uint32_t blk_probe_chain(const BlkView v, uint64_t key, uint32_t h) {
  const uint32_t limit = min(v.mask, 96u);
  for (uint32_t probes = 0; probes < limit; ++probes) {
    uint64_t cur = *reinterpret_cast<const volatile uint64_t*>(&v.keys[h]);
    if (cur == ~0ull)
      cur = atomicCAS(reinterpret_cast<unsigned long long*>(&v.keys[h]), ~0ull, (unsigned long long)key);
    if (cur == ~0ull || cur == key) return h;
    h = (h + 1u) & v.mask;
  }
  v.h->overflow = 1u;
  return v.mask + 1u;
}
// The home slot h held `kk` != key: claim it if it is empty (one CAS, inline: this is how
// every k-mer seen by a single sample enters the table), else walk the chain.
__device__ __forceinline__ uint32_t blk_resolve(const BlkView v, uint64_t key, uint32_t h, uint64_t kk) {
  if (kk == ~0ull) {
    kk = atomicCAS(reinterpret_cast<unsigned long long*>(&v.keys[h]), ~0ull, (unsigned long long)key);
    if (kk == ~0ull || kk == key) return h;
  }
  return blk_probe_chain(v, key, (h + 1u) & v.mask);
}
// list[pos] = h for every occupied slot; one shared-memory atomic per warp.  Returns nothing;
// *counter ends as the number of occupied slots.  Call with all threads of the CTA.
__device__ __forceinline__ void blk_list_occupied(const BlkView v, uint32_t slots, uint32_t* counter) {
  const uint32_t lane = lane_id();
  for (uint32_t h0 = 0; h0 < slots; h0 += kBlkThreads) {
    const uint32_t h = h0 + threadIdx.x;
    const bool used = h < slots && v.keys[h] != ~0ull;
    const uint32_t m = __ballot_sync(kFull, used);
    if (m == 0u) continue;
    uint32_t base = 0;
    if (lane == (uint32_t)__ffs(m) - 1u) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(kFull, base, __ffs(m) - 1);
    if (used) v.list[base + __popc(m & lanemask_lt())] = (uint16_t)h;
  }
}

// ---------------------------------------------------------------------------
// kA
// ---------------------------------------------------------------------------
struct BlkPlan {
  const uint32_t* item_base;     // [n_clusters + 1] exclusive scan of blocks per cluster
  const uint32_t* item_cluster;  // [n_items] inverse of item_base
  const ClusterBlk* cblk;
  uint32_t n_clusters;
  uint32_t block_windows;        // B, multiple of kBlkRun, kBlkRun .. kBlkRun * kBlkWarps
  uint32_t slots, max_unique;    // table size of kA (power of two), fill limit
  uint32_t W, WP;                // bitset words, slab row stride (W rounded up to 4)
};

template <bool CANON, bool KHI /* k > 16 */>
__global__ void __launch_bounds__(kBlkThreads)
kA_block_aggregate(const uint64_t* __restrict__ bases, const uint32_t* __restrict__ ambbits,
                   const SeqLite* __restrict__ seqs, BlkPlan plan, int k,
                   uint64_t* __restrict__ slab_keys, uint32_t* __restrict__ slab_rows,
                   uint32_t* __restrict__ slab_base, uint32_t* __restrict__ slab_count,
                   uint32_t partial_capacity, uint32_t* __restrict__ counters,
                   const uint32_t* __restrict__ item_list /* null: item = blockIdx.x */,
                   uint32_t* __restrict__ rescue_items /* out: items whose table overflowed */) {
  extern __shared__ __align__(16) unsigned char blk_raw[];
  const BlkView v = blk_view(blk_raw, plan.slots);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t item = item_list ? item_list[blockIdx.x] : blockIdx.x;

  const uint32_t c = plan.item_cluster[item];
  const ClusterBlk cb = plan.cblk[c];
  const uint32_t p0 = (item - plan.item_base[c]) * plan.block_windows;
  const uint32_t W = plan.W, WS = W | 1u;

  for (uint32_t i = tid; i < plan.slots; i += kBlkThreads) v.keys[i] = ~0ull;
  for (uint32_t i = tid; i < plan.slots * WS; i += kBlkThreads) v.pool[i] = 0;
  if (tid == 0) { v.h->n_unique = 0; v.h->overflow = 0; v.h->n_pass = 0; }
  __syncthreads();

  // task = (sequence, run of 16 windows).  Warp -> run of the block and phase over the
  // sequences; lane -> a contiguous chunk of the cluster's sequences (sample order), so
  // the 32 lanes sit in 32 different sample ranges at the same positions.
  const uint32_t n_sub = plan.block_windows / kBlkRun;
  const uint32_t sub = warp % n_sub, phase = warp / n_sub, n_phase = max(1u, (uint32_t)kBlkWarps / n_sub);
  const uint32_t chunk = (cb.n_seqs + 31u) >> 5;
  const uint32_t s_rel = p0 + sub * kBlkRun;                  // first window of the run
  const uint32_t sh64 = 64u - 2u * (uint32_t)k;
  const uint32_t mask_h = KHI ? (k == 32 ? 0xffffffffu : ((1u << (2 * k - 32)) - 1u)) : 0u;
  const uint32_t mask_l = KHI ? 0xffffffffu : (k == 16 ? 0xffffffffu : ((1u << (2 * k)) - 1u));
  const uint32_t hshift = v.shift;
  const uint32_t WS4 = WS * 4u;

  if (warp < n_sub * n_phase) {
    for (uint32_t j = phase; j < chunk; j += n_phase) {
      const uint32_t si = lane * chunk + j;
      if (si >= cb.n_seqs) continue;
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(seqs + cb.seq_start + si));
      const uint32_t len = raw.y;
      if (len < (uint32_t)k) continue;
      const uint32_t nwin = len - (uint32_t)k + 1u;
      if (s_rel >= nwin) continue;
      const uint32_t nv = min((uint32_t)kBlkRun, nwin - s_rel);
      const uint32_t sample = raw.z & 0x7fffffffu;
      const bool amb = (raw.z >> 31) != 0u;
      unsigned char* row_word = reinterpret_cast<unsigned char*>(v.pool + (sample >> 5));
      const uint32_t bit = 1u << (sample & 31u);

      // three words cover the run: 16 + k - 1 <= 47 bases from an offset < 32
      const uint64_t* w = bases + raw.x + (s_rel >> 5);
      const uint64_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
      const uint32_t o0 = s_rel & 31u;
      const uint64_t x = (w0 << (2u * o0)) | ((w1 >> 1) >> (63u - 2u * o0));
      const uint64_t f0 = x >> sh64;
      uint32_t fh = (uint32_t)(f0 >> 32), fl = (uint32_t)f0;
      uint64_t r0 = __brevll(~f0);                            // left-aligned reverse complement (+ junk below)
      r0 = ((r0 & 0xaaaaaaaaaaaaaaaaULL) >> 1) | ((r0 & 0x5555555555555555ULL) << 1);
      uint32_t rh = (uint32_t)(r0 >> 32), rl = (uint32_t)r0;
      // the 16 bases that follow the first window, next base in the top two bits
      const uint32_t off = o0 + (uint32_t)k;
      const uint64_t pa = off >= 32u ? w1 : w0, pb = off >= 32u ? w2 : w1;
      const uint32_t o1 = off & 31u;
      uint32_t cs = (uint32_t)(((pa << (2u * o1)) | ((pb >> 1) >> (63u - 2u * o1))) >> 32);
      // the same 16 bases complemented and in reverse order: next base in the low two bits
      uint32_t rs = __brev(~cs);
      rs = ((rs & 0xaaaaaaaau) >> 1) | ((rs & 0x55555555u) << 1);

      auto roll = [&]() {
        fh = __funnelshift_l(fl, fh, 2) & mask_h;
        fl = __funnelshift_l(cs, fl, 2);
        if (!KHI) fl &= mask_l;
        cs <<= 2;
        rl = __funnelshift_r(rl, rh, 2);
        rh = __funnelshift_r(rh, rs, 2);          // complement of the entering base on top
        rs >>= 2;
      };
      auto rc_right = [&]() -> uint64_t {
        if (KHI) return ((uint64_t)(rh >> sh64) << 32) | __funnelshift_r(rl, rh, sh64);
        return (uint64_t)(rh >> (sh64 - 32u));
      };
      auto put = [&](uint64_t key) {
        uint32_t h = blk_hash((uint32_t)(key >> 32), (uint32_t)key, hshift);
        const uint64_t kk = *reinterpret_cast<const volatile uint64_t*>(&v.keys[h]);
        if (kk != key) h = blk_resolve(v, key, h, kk);
        atomicOr(reinterpret_cast<uint32_t*>(row_word + h * WS4), bit);
      };

      if (nv == (uint32_t)kBlkRun && !amb) {
        // fast path: 4 windows at a time so that the probes of a batch overlap
#pragma unroll
        for (int q0 = 0; q0 < kBlkRun; q0 += 4) {
          uint64_t key[4], key2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint64_t f = ((uint64_t)fh << 32) | fl;
            const uint64_t r = rc_right();
            if (CANON) key[q] = r < f ? r : f;
            else { key[q] = f; key2[q] = r; }
            roll();
          }
          uint32_t h[4];
          uint64_t kk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            h[q] = blk_hash((uint32_t)(key[q] >> 32), (uint32_t)key[q], hshift);
            kk[q] = *reinterpret_cast<const volatile uint64_t*>(&v.keys[h[q]]);
          }
          if ((kk[0] != key[0]) | (kk[1] != key[1]) | (kk[2] != key[2]) | (kk[3] != key[3])) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (kk[q] != key[q]) h[q] = blk_resolve(v, key[q], h[q], kk[q]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) atomicOr(reinterpret_cast<uint32_t*>(row_word + h[q] * WS4), bit);
          if (!CANON) {
#pragma unroll
            for (int q = 0; q < 4; ++q) put(key2[q]);
          }
        }
      } else {
        const uint32_t* ab = amb ? ambbits + raw.w : nullptr;
        for (uint32_t q = 0; q < nv; ++q) {
          bool dead = false;
          if (amb) {
            const uint32_t p = s_rel + q;
            const uint32_t wi = p >> 5, bs = p & 31u;
            const uint64_t two = ((uint64_t)ab[wi] << 32) | (uint64_t)ab[wi + 1];
            dead = ((two << bs) >> (64 - k)) != 0ull;
          }
          if (!dead) {
            const uint64_t f = ((uint64_t)fh << 32) | fl;
            const uint64_t r = rc_right();
            if (CANON) put(r < f ? r : f);
            else { put(f); put(r); }
          }
          roll();
        }
      }
      if (*reinterpret_cast<const volatile uint32_t*>(&v.h->overflow)) break;
    }
  }
  __syncthreads();

  if (v.h->overflow) {
    if (tid == 0) {
      slab_count[item] = kBlkOverflow;
      slab_base[item] = 0;
      atomicExch(&counters[LC_TABLE_OVERFLOW], 1u);
      rescue_items[atomicAdd(&counters[LC_RESCUE], 1u)] = item;
    }
    return;
  }
  blk_list_occupied(v, plan.slots, &v.h->n_pass);
  __syncthreads();
  // the slab of this block: n partial rows from a bump allocator
  const uint32_t n = v.h->n_pass;
  if (tid == 0) {
    const uint32_t b = atomicAdd(&counters[LC_PARTIALS], n);
    v.h->row_base = b;
    v.h->ok = 1;
    slab_base[item] = b;
    slab_count[item] = n;
    if ((uint64_t)b + n > partial_capacity) {
      v.h->ok = 0;
      slab_count[item] = 0;
      atomicExch(&counters[LC_PARTIAL_OVERFLOW], 1u);
    }
  }
  __syncthreads();
  if (!v.h->ok) return;
  const size_t base = v.h->row_base;
  const uint32_t WP = plan.WP;
  for (uint32_t r = tid; r < n; r += kBlkThreads) {
    const uint32_t h = v.list[r];
    slab_keys[base + r] = v.keys[h];
    const uint32_t* src = v.pool + h * WS;
    uint4* dst = reinterpret_cast<uint4*>(slab_rows + (base + r) * WP);
    for (uint32_t q = 0; q < WP; q += 4) {
      uint4 x;
      x.x = src[q];
      x.y = q + 1 < W ? src[q + 1] : 0u;
      x.z = q + 2 < W ? src[q + 2] : 0u;
      x.w = q + 3 < W ? src[q + 3] : 0u;
      dst[q >> 2] = x;
    }
  }
}

// ---------------------------------------------------------------------------
// kB
// ---------------------------------------------------------------------------
// one warp per cluster: partial rows of the cluster -> groups of about `target` distinct keys
__global__ void plan_merge_groups(const uint32_t* __restrict__ item_base, uint32_t n_clusters,
                                  const uint32_t* __restrict__ slab_count, uint32_t target,
                                  uint32_t* __restrict__ n_groups) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_clusters) return;
  const uint32_t lane = lane_id();
  uint32_t p = 0;
  for (uint32_t i = item_base[c] + lane; i < item_base[c + 1]; i += 32) {
    const uint32_t n = slab_count[i];
    if (n != kBlkOverflow) p += n;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(kFull, p, m);
  if (lane == 0) n_groups[c] = (p + target - 1) / target;
}

// expand an exclusive scan into its inverse map: out[base[c] + i] = c
__global__ void plan_expand_owner(const uint32_t* __restrict__ base, uint32_t n_clusters,
                                  uint32_t* __restrict__ out) {
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= n_clusters) return;
  for (uint32_t i = base[c] + lane_id(); i < base[c + 1]; i += 32) out[i] = c;
}

__device__ __forceinline__ uint32_t merge_group_of(uint64_t key, uint32_t G) {
  return __umulhi((uint32_t)(mix64(key) >> 32), G);
}

// Counting sort of the partial rows by (cluster, hash group), one warp per slab.
// SCATTER = false: group sizes (group_cnt, zeroed by the caller).
// SCATTER = true:  group_cnt holds the scanned offsets (cursors); part_list[cursor++] = row.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
kB0_group(const uint64_t* __restrict__ slab_keys, const uint32_t* __restrict__ slab_base,
          const uint32_t* __restrict__ slab_count, const uint32_t* __restrict__ item_cluster,
          uint32_t n_items, const uint32_t* __restrict__ group_base,
          uint32_t* __restrict__ group_cnt, uint32_t* __restrict__ part_list) {
  const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= n_items) return;
  const uint32_t n = slab_count[item];
  if (n == kBlkOverflow || n == 0) return;
  const uint32_t c = item_cluster[item];
  const uint32_t gb = group_base[c], G = group_base[c + 1] - gb;
  const uint32_t base = slab_base[item];
  for (uint32_t i = lane_id(); i < n; i += 32) {
    const uint32_t g = gb + merge_group_of(slab_keys[base + i], G);
    if (SCATTER) part_list[atomicAdd(&group_cnt[g], 1u)] = base + i;
    else atomicAdd(&group_cnt[g], 1u);
  }
}

__global__ void __launch_bounds__(kBlkThreads)
kB_merge(const uint64_t* __restrict__ slab_keys, const uint32_t* __restrict__ slab_rows, BlkPlan plan,
         const uint32_t* __restrict__ group_base /* [n_clusters + 1] */,
         const uint32_t* __restrict__ group_cluster, const uint32_t* __restrict__ group_off,
         const uint32_t* __restrict__ part_list, uint32_t merge_slots, uint32_t merge_max_unique,
         const ClusterDev* __restrict__ clusters, RowOut out, uint32_t row_capacity,
         uint32_t* __restrict__ counters, uint32_t* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char blk_raw[];
  const BlkView v = blk_view(blk_raw, merge_slots);
  const uint32_t tid = threadIdx.x;
  const uint32_t W = plan.W, WS = W | 1u, WP = plan.WP;
  const uint32_t total_work = group_base[plan.n_clusters];

  for (;;) {
    __syncthreads();
    if (tid == 0) v.h->work = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t work = v.h->work;
    if (work >= total_work) return;
    const uint32_t c = group_cluster[work];
    const uint32_t p0 = group_off[work], p1 = group_off[work + 1];
    if (p0 == p1) continue;
    const ClusterDev cl = clusters[c];

    for (uint32_t i = tid; i < merge_slots; i += kBlkThreads) v.keys[i] = ~0ull;
    for (uint32_t i = tid; i < merge_slots * WS; i += kBlkThreads) v.pool[i] = 0;
    if (tid == 0) { v.h->n_unique = 0; v.h->overflow = 0; v.h->n_pass = 0; v.h->ok = 1; }
    __syncthreads();
    for (uint32_t i = p0 + tid; i < p1; i += kBlkThreads) {
      const uint32_t pi = part_list[i];
      const uint64_t key = slab_keys[pi];
      uint32_t h = blk_hash((uint32_t)(key >> 32), (uint32_t)key, v.shift);
      const uint64_t kk = *reinterpret_cast<const volatile uint64_t*>(&v.keys[h]);
      if (kk != key) h = blk_resolve(v, key, h, kk);
      if (h >= merge_slots) continue;
      const uint4* src = reinterpret_cast<const uint4*>(slab_rows + (size_t)pi * WP);
      uint32_t* dst = v.pool + h * WS;
      for (uint32_t q = 0; q < WP / 4; ++q) {
        const uint4 x = __ldg(src + q);
        if (x.x) atomicOr(dst + 4 * q, x.x);
        if (x.y) atomicOr(dst + 4 * q + 1, x.y);
        if (x.z) atomicOr(dst + 4 * q + 2, x.z);
        if (x.w) atomicOr(dst + 4 * q + 3, x.w);
      }
    }
    __syncthreads();
    if (v.h->overflow) {
      if (tid == 0) atomicExch(&counters[LC_TABLE_OVERFLOW], 2u);
      continue;
    }
    // ---- counts, filter, rows (as k3_local_direct) ------------------------------------
    for (uint32_t h0 = 0; h0 < merge_slots; h0 += kBlkThreads) {
      const uint32_t h = h0 + tid;
      const bool used = h < merge_slots && v.keys[h] != ~0ull;
      uint32_t cnt = 0;
      if (used) {
        const uint32_t* bits = v.pool + h * WS;
        for (uint32_t w = 0; w < W; ++w) cnt += __popc(bits[w]);
      }
      const bool pass = used && cnt >= cl.lo && cnt <= cl.hi;
      const uint32_t mu = __ballot_sync(kFull, used), mp = __ballot_sync(kFull, pass);
      const uint32_t lane = tid & 31u;
      if (lane == 0 && mu) atomicAdd(&v.h->n_unique, (uint32_t)__popc(mu));
      if (mp) {
        uint32_t b = 0;
        if (lane == (uint32_t)__ffs(mp) - 1u) b = atomicAdd(&v.h->n_pass, (uint32_t)__popc(mp));
        b = __shfl_sync(kFull, b, __ffs(mp) - 1);
        if (pass) v.list[b + __popc(mp & lanemask_lt())] = (uint16_t)h;
      }
    }
    __syncthreads();
    if (tid == 0) {
      atomicAdd(&counters[LC_UNIQUE], v.h->n_unique);
      if (v.h->n_pass) {
        const uint32_t b = atomicAdd(&counters[LC_ROWS], v.h->n_pass);
        v.h->row_base = b;
        if ((uint64_t)b + v.h->n_pass > row_capacity) {
          v.h->ok = 0;
          atomicExch(&counters[LC_ROW_OVERFLOW], 1u);
        }
      }
    }
    __syncthreads();
    const uint32_t n_pass = v.h->n_pass;
    if (n_pass == 0 || !v.h->ok) continue;
    const uint32_t rbase = v.h->row_base;
    for (uint32_t r = tid; r < n_pass; r += kBlkThreads) {
      const uint32_t h = v.list[r];
      const uint32_t* bits = v.pool + h * WS;
      uint32_t cnt = 0;
      for (uint32_t w = 0; w < W; ++w) cnt += __popc(bits[w]);
      const size_t gi = (size_t)rbase + r;
      out.cluster[gi] = cl.id;
      out.kmer[gi] = v.keys[h];
      out.count[gi] = cnt;
      if (out.key_words > W) out.cand[gi * out.key_words + W] = out.cluster_pattern[c];
    }
    for (uint32_t i = tid; i < n_pass * W; i += kBlkThreads) {
      const uint32_t r = i / W, wd = i - r * W;
      out.cand[((size_t)rbase + r) * out.key_words + wd] = v.pool[(uint32_t)v.list[r] * WS + wd];
    }
  }
}

}  // namespace pf
