// K1 fused into the first radix pass (partition mode, direct variant).
//
// The unfused path writes every 12-byte record once (k1_extract), reads the keys
// again for the histogram and reads the records a third time in the first pass.
// A k-mer is cheaper to recompute from the 2-bit plane (0.25 B/base, L2-resident
// per cluster) than to move, so:
//
//   k1_histogram_fused   walks the sequences like k1_extract but only counts the
//                        leading digit(s) of mix64(k-mer) per cluster (per-warp
//                        shared-memory histograms, flushed with one atomicAdd per
//                        non-empty bin and cluster).  No record is written.
//   k2_extract_scatter   is k2_scatter_pass whose tile load is replaced by the
//                        extraction: record index -> (sequence, window) by a short
//                        binary search in the per-sequence record offsets, two
//                        64-bit loads of the plane, funnel shift, reverse complement,
//                        canonical choice, mix.  Records first exist in HBM already
//                        partitioned by their leading digit.
//
// Same window semantics as k1_extract (/root/reference/panfeed/panfeed.py:54-88).
#pragma once
#include "k1_extract.cuh"
#include "k2_onesweep.cuh"

namespace pf {

// forward / reverse-complement k-mer of window p of sequence d, and whether the
// window touches a non-ACGT symbol (then the narrow record is dead)
__device__ __forceinline__ void window_kmers(const uint64_t* __restrict__ bases,
                                             const uint32_t* __restrict__ ambbits,
                                             uint64_t base_off, uint64_t amb_off, uint32_t flags,
                                             uint32_t p, int k, uint64_t& fwd, uint64_t& rc,
                                             bool& is_amb) {
  const uint64_t* w = bases + (base_off >> 5) + (p >> 5);
  const uint64_t w0 = __ldg(w), w1 = __ldg(w + 1);
  const uint32_t sh = 2u * (p & 31u);
  const uint64_t x = (w0 << sh) | ((w1 >> 1) >> (63u - sh));
  fwd = x >> (64 - 2 * k);
  rc = revcomp2(fwd, k);
  is_amb = false;
  if (flags & 2u) {
    const uint32_t* ab = ambbits + (amb_off >> 5);
    const uint32_t wi = p >> 5, bs = p & 31u;
    const uint64_t two = ((uint64_t)ab[wi] << 32) | (uint64_t)ab[wi + 1];
    is_amb = ((two << bs) >> (64 - k)) != 0ull;
  }
}

constexpr int kHistSeqsPerChunk = 8;

template <bool CANON>
__global__ void __launch_bounds__(kK1Warps * 32)
k1_histogram_fused(const uint64_t* __restrict__ bases, const uint32_t* __restrict__ ambbits,
                   const SeqDev* __restrict__ seqs, uint32_t n_seqs, int k, int passes, int shift0,
                   uint32_t* __restrict__ seg_hist /* [seg][passes][256] */) {
  __shared__ __align__(16) uint64_t stage_all[kK1Warps][kK1StageWords];
  __shared__ uint32_t hist_all[kK1Warps][2 * kRadix];      // passes <= 2
  const uint32_t lane = lane_id();
  const uint32_t warp = threadIdx.x >> 5;
  uint64_t* stage = stage_all[warp];
  uint32_t* h = hist_all[warp];
  const uint32_t n_chunks = (n_seqs + kHistSeqsPerChunk - 1) / kHistSeqsPerChunk;
  const uint32_t stride = gridDim.x * kK1Warps;
  const int kshift = 64 - 2 * k;
  for (uint32_t i = lane; i < 2 * kRadix; i += 32) h[i] = 0;
  __syncwarp();

  for (uint32_t chunk = blockIdx.x * kK1Warps + warp; chunk < n_chunks; chunk += stride) {
    const uint32_t s0 = chunk * kHistSeqsPerChunk;
    const uint32_t s1 = min(n_seqs, s0 + kHistSeqsPerChunk);
    uint32_t cur_cluster = seqs[s0].cluster;
    for (uint32_t s = s0; s < s1; ++s) {
      const SeqDev d = seqs[s];
      if (d.cluster != cur_cluster) {
        __syncwarp();
        for (uint32_t i = lane; i < (uint32_t)passes * kRadix; i += 32) {
          const uint32_t c = h[i];
          if (c) { atomicAdd(&seg_hist[(size_t)cur_cluster * passes * kRadix + i], c); h[i] = 0; }
        }
        __syncwarp();
        cur_cluster = d.cluster;
      }
      if (d.len < (uint32_t)k) continue;
      const uint32_t nwin = d.len - (uint32_t)k + 1u;
      const uint64_t* w = bases + (d.base_off >> 5);
      const bool amb = (d.flags & 2u) != 0u;
      const uint32_t* ab = amb ? (ambbits + (d.amb_off >> 5)) : nullptr;
      uint4 cur = ld_stream128(w + 2 * lane);
      uint4 halo = make_uint4(0, 0, 0, 0);
      if (lane == 0) halo = ld_stream128(w + 64);
      for (uint32_t c0 = 0; c0 < nwin; c0 += 2048u) {
        __syncwarp();
        reinterpret_cast<uint4*>(stage)[lane] = cur;
        if (lane == 0) reinterpret_cast<uint4*>(stage)[32] = halo;
        __syncwarp();
        const uint32_t next = c0 + 2048u;
        if (next < nwin) {
          const uint64_t* wn = w + (next >> 5);
          cur = ld_stream128(wn + 2 * lane);
          if (lane == 0) halo = ld_stream128(wn + 64);
        }
        const uint32_t iters = min(64u, (nwin - c0 + 31u) >> 5);
        for (uint32_t it = 0; it < iters; ++it) {
          const uint32_t p = c0 + it * 32u + lane;
          if (p >= nwin) continue;
          const uint64_t w0 = stage[it], w1 = stage[it + 1];
          const uint32_t sh = 2u * lane;
          const uint64_t x = (w0 << sh) | ((w1 >> 1) >> (63u - sh));
          const uint64_t fwd = x >> kshift;
          const uint64_t rc = revcomp2(fwd, k);
          bool is_amb = false;
          if (amb) {
            const uint32_t wi = p >> 5, bs = p & 31u;
            const uint64_t two = ((uint64_t)ab[wi] << 32) | (uint64_t)ab[wi + 1];
            is_amb = ((two << bs) >> (64 - k)) != 0ull;
          }
          if (CANON) {
            const uint64_t key = is_amb ? 0ull : mix64(rc < fwd ? rc : fwd);
            for (int q = 0; q < passes; ++q) atomicAdd(&h[q * kRadix + key_digit(key, shift0 + 8 * q)], 1u);
          } else {
            const uint64_t k0 = is_amb ? 0ull : mix64(fwd), k1 = is_amb ? 0ull : mix64(rc);
            for (int q = 0; q < passes; ++q) {
              atomicAdd(&h[q * kRadix + key_digit(k0, shift0 + 8 * q)], 1u);
              atomicAdd(&h[q * kRadix + key_digit(k1, shift0 + 8 * q)], 1u);
            }
          }
        }
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < (uint32_t)passes * kRadix; i += 32) {
      const uint32_t c = h[i];
      if (c) { atomicAdd(&seg_hist[(size_t)cur_cluster * passes * kRadix + i], c); h[i] = 0; }
    }
    __syncwarp();
  }
}

template <bool CANON>
__global__ void __launch_bounds__(kSortThreads, 3)
k2_extract_scatter(const uint64_t* __restrict__ bases, const uint32_t* __restrict__ ambbits,
                   const SeqDev* __restrict__ seqs, const uint32_t* __restrict__ seq_rec_off /* [n_seqs+1] */,
                   uint32_t n_seqs, const uint32_t* __restrict__ tile_first_seq /* [n_tiles+1] */, int k,
                   uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                   const TileDev* __restrict__ tiles, uint32_t n_tiles,
                   uint32_t* __restrict__ cursors /* copy of the scanned histogram: next free slot of
                                                     every (segment, digit) bucket */,
                   int passes, int shift) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<uint64_t>& sm = *reinterpret_cast<ScatterSmem<uint64_t>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kWarpItems = kSortItems * 32;

  sm.hist[tid] = 0;
  __syncthreads();
  const uint32_t tile = blockIdx.x;
  if (tile >= n_tiles) return;
  const TileDev td = tiles[tile];
  const uint32_t seq_lo = tile_first_seq[tile];
  const uint32_t seq_hi = min(n_seqs, tile_first_seq[tile + 1] + 1u);   // exclusive bound of candidates

  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
  uint16_t rank[kSortItems];
  const uint32_t wbase = warp * kWarpItems + lane;
  // items of a thread ascend, so does their sequence: keep the current one in registers
  uint32_t s = seq_lo, s_lo = seq_rec_off[seq_lo], s_hi = seq_rec_off[seq_lo + 1];
  bool loaded = false;
  const uint64_t* s_words = nullptr;
  const uint32_t* s_amb = nullptr;
  uint32_t s_sample = 0;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    key[i] = 0;
    val[i] = 0;
    if (idx < td.count) {
      const uint32_t r = td.start + idx;
      if (r >= s_hi) {
        // sequence holding record r: first index in (s, seq_hi) whose offset exceeds r, minus 1
        uint32_t lo = s + 1, hi = seq_hi;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (seq_rec_off[mid] <= r) lo = mid + 1; else hi = mid;
        }
        s = lo - 1;
        s_lo = seq_rec_off[s];
        s_hi = seq_rec_off[s + 1];
        loaded = false;
      }
      if (!loaded) {
        const SeqDev* d = seqs + s;
        s_words = bases + (d->base_off >> 5);
        s_amb = (d->flags & 2u) ? ambbits + (d->amb_off >> 5) : nullptr;
        s_sample = d->sample;
        loaded = true;
      }
      const uint32_t q = r - s_lo;
      const uint32_t p = CANON ? q : (q >> 1);
      const uint64_t w0 = __ldg(s_words + (p >> 5)), w1 = __ldg(s_words + (p >> 5) + 1);
      const uint32_t sh = 2u * (p & 31u);
      const uint64_t x = (w0 << sh) | ((w1 >> 1) >> (63u - sh));
      const uint64_t fwd = x >> (64 - 2 * k);
      const uint64_t rc = revcomp2(fwd, k);
      bool is_amb = false;
      if (s_amb) {
        const uint32_t wi = p >> 5, bs = p & 31u;
        const uint64_t two = ((uint64_t)s_amb[wi] << 32) | (uint64_t)s_amb[wi + 1];
        is_amb = ((two << bs) >> (64 - k)) != 0ull;
      }
      uint64_t kk;
      if (CANON) kk = rc < fwd ? rc : fwd;
      else kk = (q & 1u) ? rc : fwd;
      key[i] = is_amb ? 0ull : mix64(kk);
      val[i] = is_amb ? kInvalidSample : s_sample;
    }
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    rank[i] = 0;
    if (idx < td.count) rank[i] = (uint16_t)atomicAdd(&sm.hist[key_digit(key[i], shift)], 1u);
  }
  __syncthreads();
  {
    const uint32_t total = sm.hist[tid];
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(kFull, incl, o);
      if (lane >= (uint32_t)o) incl += n;
    }
    if (lane == 31) sm.warp_sums[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) woff += (w < (int)warp) ? sm.warp_sums[w] : 0u;
    const uint32_t excl = woff + incl - total;
    sm.excl[tid] = excl;
    // The bucket sizes are exact (histogram), so space inside a bucket can be handed out
    // in any order: one global atomicAdd per (tile, digit) replaces the look-back chain.
    // The order of records inside a bucket is then arbitrary, which the direct local
    // reduce does not care about.
    const size_t ds = (size_t)td.seg * passes * kRadix + tid;      // pass 0
    const uint32_t first = total ? atomicAdd(&cursors[ds], total) : 0u;
    sm.gbase[tid] = first - excl;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    if (idx < td.count) {
      const uint32_t slot = sm.excl[key_digit(key[i], shift)] + rank[i];
      sm.keys[slot] = key[i];
      sm.vals[slot] = val[i];
    }
  }
  __syncthreads();
  for (uint32_t idx = tid; idx < td.count; idx += kSortThreads) {
    const uint64_t kq = sm.keys[idx];
    const uint32_t g = sm.gbase[key_digit(kq, shift)] + idx;   // 32-bit wrap-around is intended
    keys_out[g] = kq;
    vals_out[g] = sm.vals[idx];
  }
}

}  // namespace pf
