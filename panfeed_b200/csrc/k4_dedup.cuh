// K4 — global pattern deduplication in a GPU open-addressing table keyed on
// the FULL bitset.  Replaces `if khash in patterns: continue; patterns.add()`
// (/root/reference/panfeed/panfeed.py:179-180,210-212) where the reference keys
// a Python set on base64(md5(vector bytes)): here two different patterns can
// never be merged, whatever their hashes.
//
// The table holds 32-bit entries: an index into the persistent pattern pool,
// or (top bit set) the index of a candidate row of the current batch.  One warp
// per candidate: hash W words cooperatively, linear-probe, claim an empty slot
// with one atomicCAS, and on an occupied slot compare all W words against the
// occupant (pool entry or another candidate of this batch).  After the kernel
// the winners are numbered by a scan and copied to the pool in that order.
#pragma once
#include "pf_common.cuh"

namespace pf {

constexpr uint32_t kEmptySlot = 0xffffffffu;
constexpr uint32_t kTentative = 0x80000000u;

__device__ __forceinline__ uint64_t warp_hash_words(const uint32_t* __restrict__ key, uint32_t n) {
  uint64_t h = 0;
  for (uint32_t w = lane_id(); w < n; w += 32) h += word_hash(key[w], w);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) h += __shfl_xor_sync(kFull, h, m);
  return fmix64(h);
}

__device__ __forceinline__ bool warp_equal_words(const uint32_t* __restrict__ a,
                                                 const uint32_t* __restrict__ b, uint32_t n) {
  for (uint32_t w0 = 0; w0 < n; w0 += 32) {
    const uint32_t w = w0 + lane_id();
    const bool ok = (w >= n) || (a[w] == b[w]);
    if (!__all_sync(kFull, ok)) return false;
  }
  return true;
}

// rep[row]  : pool index of an existing equal pattern, or kTentative|q where q
//             is the candidate row that claimed the slot (q == row: this row won)
// slot_of[row] (winners only): the table slot it claimed
__global__ void __launch_bounds__(256)
k4_probe(const uint32_t* __restrict__ cand, uint32_t n_rows, uint32_t key_words,
         const uint32_t* __restrict__ pool, uint32_t* __restrict__ table, uint32_t table_mask,
         uint32_t* __restrict__ rep, uint32_t* __restrict__ slot_of,
         uint32_t* __restrict__ winner_flag) {
  const uint32_t lane = lane_id();
  const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows;
       row += total_warps) {
    const uint32_t* key = cand + (size_t)row * key_words;
    uint32_t slot = (uint32_t)warp_hash_words(key, key_words) & table_mask;
    uint32_t result = kEmptySlot;
    bool won = false;
    for (;;) {
      uint32_t e = 0;
      if (lane == 0) {
        e = *reinterpret_cast<volatile uint32_t*>(table + slot);
        if (e == kEmptySlot) {
          const uint32_t old = atomicCAS(table + slot, kEmptySlot, kTentative | row);
          if (old == kEmptySlot) { won = true; e = kTentative | row; }
          else e = old;
        }
      }
      e = __shfl_sync(kFull, e, 0);
      won = __shfl_sync(kFull, (int)won, 0) != 0;
      if (won) { result = e; break; }
      const uint32_t* other = (e & kTentative)
                                  ? cand + (size_t)(e & ~kTentative) * key_words
                                  : pool + (size_t)e * key_words;
      if (warp_equal_words(key, other, key_words)) { result = e; break; }
      slot = (slot + 1u) & table_mask;
    }
    if (lane == 0) {
      rep[row] = result;
      winner_flag[row] = won ? 1u : 0u;
      if (won) slot_of[row] = slot;
    }
  }
}

// winners: copy the candidate into the pool at pool_base + rank, fix the slot.
// everyone: translate rep into a final pool index.
__global__ void __launch_bounds__(256)
k4_commit(const uint32_t* __restrict__ cand, uint32_t n_rows, uint32_t key_words,
          uint32_t* __restrict__ pool, uint32_t pool_base, uint32_t* __restrict__ table,
          const uint32_t* __restrict__ rep, const uint32_t* __restrict__ slot_of,
          const uint32_t* __restrict__ winner_rank /* exclusive scan of winner_flag, n_rows+1 */,
          uint32_t* __restrict__ row_pattern) {
  const uint32_t lane = lane_id();
  const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows;
       row += total_warps) {
    const uint32_t r = rep[row];
    uint32_t id;
    if (r & kTentative) {
      const uint32_t q = r & ~kTentative;
      id = pool_base + winner_rank[q];
      if (q == row) {
        const uint32_t* src = cand + (size_t)row * key_words;
        uint32_t* dst = pool + (size_t)id * key_words;
        for (uint32_t w = lane; w < key_words; w += 32) dst[w] = src[w];
        if (lane == 0) table[slot_of[row]] = id;
      }
    } else {
      id = r;
    }
    if (lane == 0) row_pattern[row] = id;
  }
}

// re-insert pool entries [0, n) into a fresh (larger) table; entries are distinct
__global__ void __launch_bounds__(256)
k4_rehash(const uint32_t* __restrict__ pool, uint32_t n, uint32_t key_words,
          uint32_t* __restrict__ table, uint32_t table_mask) {
  const uint32_t lane = lane_id();
  const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += total_warps) {
    uint32_t slot = (uint32_t)warp_hash_words(pool + (size_t)e * key_words, key_words) & table_mask;
    if (lane == 0) {
      while (atomicCAS(table + slot, kEmptySlot, e) != kEmptySlot) slot = (slot + 1u) & table_mask;
    }
  }
}

}  // namespace pf
