// K2 — segmented onesweep radix sort of (mixed k-mer key, sample rank) records.
//
// Replaces the grouping the reference gets from its per-cluster Python dict
// (/root/reference/panfeed/panfeed.py:45,77-88): after the sort, all instances
// of one k-mer of one cluster are adjacent.
//
// * Segments = gene clusters.  A tile (4096 records) never crosses a segment,
//   so the cluster id costs no key bits and no extra pass.
// * LSD passes of 8 bits over the LEADING `sort_bits` of the mixed key only
//   (K3 resolves the rest exactly), least-significant digit first; the sort is
//   stable, so records of equal key keep K1's order (sample ranks ascending).
// * One pass = one read and one write of every record ("onesweep"): per-tile
//   digit counts are chained with a decoupled look-back (aggregate / inclusive
//   flags in one 32-bit word per (tile, digit)), the per-segment digit starts
//   come from one up-front histogram kernel over all passes.
// * Tiles are claimed through an atomic ticket so a tile only ever waits for
//   tiles that are already resident.  Spins are bounded by a watchdog that
//   raises an error flag instead of hanging the GPU.
#pragma once
#include "pf_common.cuh"

namespace pf {

constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValMask = ~kFlagMask;
constexpr uint32_t kSpinLimit = 1u << 24;
constexpr int kMaxPasses = 8;
static_assert(kSortThreads == kRadix, "one thread per digit in the scan / look-back phase");

// Histogram of every pass's digit for every segment, from one read of the keys.
// CTA b owns tiles [b*per, (b+1)*per); it flushes its shared histogram to
// seg_hist[seg][pass][256] whenever the segment changes.
template <typename KeyT>
__global__ void __launch_bounds__(256)
k2_histogram(const KeyT* __restrict__ keys, const TileDev* __restrict__ tiles,
             uint32_t n_tiles, uint32_t tiles_per_cta, int passes, int shift0,
             uint32_t* __restrict__ seg_hist) {
  __shared__ uint32_t h[kMaxPasses][kRadix];
  const uint32_t t0 = blockIdx.x * tiles_per_cta;
  const uint32_t t1 = min(n_tiles, t0 + tiles_per_cta);
  if (t0 >= t1) return;
  for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  uint32_t cur_seg = tiles[t0].seg;
  for (uint32_t t = t0; t < t1; ++t) {
    const TileDev td = tiles[t];
    if (td.seg != cur_seg) {
      __syncthreads();
      for (int i = threadIdx.x; i < passes * kRadix; i += 256) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&seg_hist[(size_t)cur_seg * passes * kRadix + i], c);
        (&h[0][0])[i] = 0;
      }
      __syncthreads();
      cur_seg = td.seg;
    }
    for (uint32_t i = threadIdx.x; i < td.count; i += 256) {
      const KeyT key = keys[(size_t)td.start + i];
#pragma unroll
      for (int p = 0; p < kMaxPasses; ++p)
        if (p < passes) atomicAdd(&h[p][key_digit(key, shift0 + 8 * p)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * kRadix; i += 256) {
    const uint32_t c = (&h[0][0])[i];
    if (c) atomicAdd(&seg_hist[(size_t)cur_seg * passes * kRadix + i], c);
  }
}

// In place: counts -> absolute start of each digit's bucket
// (segment start + exclusive scan).  One warp per (segment, pass).
__global__ void k2_scan_histogram(uint32_t* __restrict__ seg_hist, const uint32_t* __restrict__ seg_start,
                                  uint32_t n_rows /* segments*passes */, int passes) {
  const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const uint32_t lane = lane_id();
  uint32_t* h = seg_hist + (size_t)row * kRadix;
  uint32_t v[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = h[lane * 8 + i]; sum += v[i]; }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(kFull, incl, o);
    if (lane >= (uint32_t)o) incl += n;
  }
  uint32_t run = seg_start[row / passes] + incl - sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) { h[lane * 8 + i] = run; run += v[i]; }
}

template <typename KeyT>
struct SortSmem {
  KeyT keys[kSortTile];
  uint32_t vals[kSortTile];
  uint32_t warp_hist[kSortThreads / 32][kRadix];
  uint32_t excl[kRadix];      // first slot of each digit inside the sorted tile
  uint32_t gbase[kRadix];     // global index of tile-sorted slot 0 of each digit, minus excl
  uint32_t warp_sums[kSortThreads / 32];
  uint32_t tile;
};

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads)
k2_onesweep_pass(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                 KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                 const TileDev* __restrict__ tiles, uint32_t n_tiles,
                 const uint32_t* __restrict__ digit_start /* [seg][passes][256] */,
                 int pass, int passes, int shift,
                 uint32_t* __restrict__ lookback /* [n_tiles][256], zeroed */,
                 uint32_t* __restrict__ ticket, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortSmem<KeyT>& sm = *reinterpret_cast<SortSmem<KeyT>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kWarpItems = kSortItems * 32;

  if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
  for (int i = tid; i < kWarps * kRadix; i += kSortThreads) (&sm.warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = sm.tile;
  if (tile >= n_tiles) return;
  const TileDev td = tiles[tile];

  KeyT key[kSortItems];
  uint32_t val[kSortItems];
  uint32_t rank[kSortItems];
  const uint32_t wbase = warp * kWarpItems + lane;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    if (idx < td.count) {
      key[i] = keys_in[(size_t)td.start + idx];
      val[i] = vals_in[(size_t)td.start + idx];
    } else {
      key[i] = KeyTraits<KeyT>::zero();
      val[i] = 0;
    }
  }

  // ---- rank each item among equal digits of its warp (stable) -------------
  uint32_t* wh = sm.warp_hist[warp];
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    const bool valid = idx < td.count;
    const uint32_t d = key_digit(key[i], shift);
    const uint32_t peers = __match_any_sync(kFull, valid ? d : (256u + lane));
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader && valid) {
      base = wh[d];
      wh[d] = base + __popc(peers);
    }
    base = __shfl_sync(kFull, base, leader);
    rank[i] = base + __popc(peers & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit: offsets of each warp, tile total, tile-exclusive scan ----
  uint32_t total = 0;
  {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = sm.warp_hist[w][tid];
      sm.warp_hist[w][tid] = total;
      total += c;
    }
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

    // ---- decoupled look-back over the earlier tiles of this segment --------
    uint32_t* my = lookback + (size_t)tile * kRadix + tid;
    uint32_t prev = 0;
    if (tile != td.first_tile) {
      st_relaxed(my, kFlagAgg | total);
      uint32_t j = tile;
      uint32_t spins = 0;
      bool failed = false;
      for (;;) {
        --j;
        const uint32_t* p = lookback + (size_t)j * kRadix + tid;
        uint32_t v = ld_relaxed(p);
        while ((v & kFlagMask) == 0u) {
          if (++spins > kSpinLimit) { failed = true; break; }
          __nanosleep(40);
          v = ld_relaxed(p);
        }
        if (failed) { atomicExch(err, 1u); break; }
        prev += v & kValMask;
        if ((v & kFlagMask) == kFlagIncl || j == td.first_tile) break;
      }
    }
    st_relaxed(my, kFlagIncl | ((prev + total) & kValMask));
    const size_t ds = ((size_t)td.seg * passes + pass) * kRadix + tid;
    sm.gbase[tid] = digit_start[ds] + prev - excl;
  }
  __syncthreads();

  // ---- reorder the tile in shared memory, then write digit runs coalesced --
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    if (idx < td.count) {
      const uint32_t d = key_digit(key[i], shift);
      const uint32_t slot = sm.excl[d] + sm.warp_hist[warp][d] + rank[i];
      sm.keys[slot] = key[i];
      sm.vals[slot] = val[i];
    }
  }
  __syncthreads();
  for (uint32_t idx = tid; idx < td.count; idx += kSortThreads) {
    const KeyT k = sm.keys[idx];
    const uint32_t d = key_digit(k, shift);
    const uint32_t g = sm.gbase[d] + idx;      // 32-bit wrap-around is intended
    keys_out[g] = k;
    vals_out[g] = sm.vals[idx];
  }
}


// ---------------------------------------------------------------------------
// Unstable variant of the pass, for the direct local reduce (S <= 1024), which
// ORs sample bits and therefore does not care about the order of records inside
// a bucket.  The per-warp match/ballot ranking (a 16-deep dependent chain) is
// replaced by one shared-memory atomicAdd per record on a per-tile histogram
// (the 16 atomics of a thread are independent, so they pipeline), and the
// decoupled look-back chain by one global atomicAdd per (tile, digit) on a cursor
// array initialised with the exact bucket starts: tiles no longer wait for each other.
template <typename KeyT>
struct ScatterSmem {
  KeyT keys[kSortTile];
  uint32_t vals[kSortTile];
  uint32_t hist[kRadix];
  uint32_t excl[kRadix];
  uint32_t gbase[kRadix];
  uint32_t warp_sums[kSortThreads / 32];
  uint32_t tile;
};

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads, 3)
k2_scatter_pass(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                const TileDev* __restrict__ tiles, uint32_t n_tiles,
                uint32_t* __restrict__ cursors /* copy of the scanned histogram */, int pass, int passes,
                int shift) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem<KeyT>& sm = *reinterpret_cast<ScatterSmem<KeyT>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kWarpItems = kSortItems * 32;

  sm.hist[tid] = 0;
  __syncthreads();
  const uint32_t tile = blockIdx.x;
  if (tile >= n_tiles) return;
  const TileDev td = tiles[tile];

  KeyT key[kSortItems];
  uint32_t val[kSortItems];
  uint16_t rank[kSortItems];
  const uint32_t wbase = warp * kWarpItems + lane;
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint32_t idx = wbase + i * 32;
    if (idx < td.count) {
      key[i] = keys_in[(size_t)td.start + idx];
      val[i] = vals_in[(size_t)td.start + idx];
    } else {
      key[i] = KeyTraits<KeyT>::zero();
      val[i] = 0;
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
    // exact bucket sizes are known: claim space with one atomicAdd per (tile, digit)
    const size_t ds = ((size_t)td.seg * passes + pass) * kRadix + tid;
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
    const KeyT k = sm.keys[idx];
    const uint32_t g = sm.gbase[key_digit(k, shift)] + idx;   // 32-bit wrap-around is intended
    keys_out[g] = k;
    vals_out[g] = sm.vals[idx];
  }
}

}  // namespace pf
