// libpanfeed_b200.so — C-ABI entry points (include/panfeed_b200.h) and the host
// side of the pipeline: batch validation and planning, buffer management,
// kernel sequencing on one CUDA stream, result staging.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/panfeed_b200.h"
#include "k1_extract.cuh"
#include "k1_fused.cuh"
#include "k2_onesweep.cuh"
#include "k3_reduce.cuh"
#include "k3_local.cuh"
#include "k3_block.cuh"
#include "k4_dedup.cuh"
#include "k5_md5.cuh"
#include "synth.cuh"

using namespace pf;

namespace pf {
// ---- planning helpers that run on the device (tile lists are pure functions of the
//      per-cluster record ranges; generating them there saves host loops and H2D) -------
__global__ void plan_expand_tiles(const ClusterDev* __restrict__ clusters, uint32_t n_clusters,
                                  const uint32_t* __restrict__ tile_base, uint32_t tile_size, int wide,
                                  TileDev* __restrict__ tiles) {
  const uint32_t c = blockIdx.x;
  if (c >= n_clusters) return;
  const uint32_t lo = wide ? clusters[c].wrec_start : clusters[c].rec_start;
  const uint32_t hi = wide ? clusters[c].wrec_end : clusters[c].rec_end;
  const uint32_t first = tile_base[c];
  const uint32_t n = (hi - lo + tile_size - 1) / tile_size;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    TileDev t;
    t.start = lo + i * tile_size;
    t.count = min(tile_size, hi - t.start);
    t.seg = c;
    t.first_tile = first;
    tiles[first + i] = t;
  }
}
__global__ void plan_seq_rec_off(const SeqDev* __restrict__ seqs, uint32_t n_seqs, uint32_t total,
                                 uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seqs) out[i] = seqs[i].rec_off;
  else if (i == n_seqs) out[i] = total;
}
// Counter read-back without a copy engine: the D2H engine may be busy with the previous
// batch's rows, and a 64-byte memcpy queued behind them would stall the pipeline's host side.
// `dst` is pinned host memory (device-accessible under UVA).
__global__ void mirror_counters(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n) {
  if (threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
  __threadfence_system();
}
// pipelined submit: positional records of a sub-batch index its own sequences / wide k-mers
__global__ void pos_rebase(uint32_t* __restrict__ pos_seq, uint64_t* __restrict__ pos_kmer,
                           const uint8_t* __restrict__ pos_flags, uint32_t n, uint32_t seq_base,
                           uint64_t wide_base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pos_seq[i] += seq_base;
  if (wide_base && (pos_flags[i] & 2u)) pos_kmer[i] += wide_base;
}
// sequence holding the first record of every tile (last s with rec_off[s] <= start, non-empty)
__global__ void plan_tile_first_seq(const TileDev* __restrict__ tiles, uint32_t n_tiles,
                                    const uint32_t* __restrict__ seq_rec_off, uint32_t n_seqs,
                                    uint32_t* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  if (t == n_tiles) { out[t] = n_seqs ? n_seqs - 1 : 0; return; }
  const uint32_t r = tiles[t].start;
  uint32_t lo = 0, hi = n_seqs;                   // first index in [0, n_seqs] with rec_off > r
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (seq_rec_off[mid] <= r) lo = mid + 1; else hi = mid;
  }
  out[t] = lo ? lo - 1 : 0;
}
}  // namespace pf

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PatternSpace {
  uint32_t key_words = 0;
  DevBuf pool;            // n x key_words
  uint64_t n = 0;         // committed patterns
  DevBuf table;           // table_size x u32
  uint32_t table_size = 0;
  // exchange state
  DevBuf x_owner, x_pos, x_perm, x_counts, x_unique, x_table, x_rep, x_slot, x_winner;
  uint64_t x_n_unique = 0;
};

struct WidthState {       // per key width (narrow u64 / wide Key128)
  DevBuf keys[2], vals[2];
  DevBuf tiles, seg_start, seg_hist, lookback, cursors;
  DevBuf ltiles, tile_first_run;     // partition mode: 2048-record tiles of the local reduce
  PinBuf h_tiles, h_seg_start, h_ltiles;       // h_tiles / h_ltiles now hold per-cluster tile bases
  DevBuf d_tile_base, d_ltile_base;
  uint32_t n_tiles = 0, n_ltiles = 0, max_seg = 0;
  uint32_t n_records = 0;
  uint32_t n_runs = 0;
  uint32_t n_rows = 0;
  int sort_bits = 0, passes = 0;
  int final_buf = 0;      // which of keys[]/vals[] holds the sorted records
};

enum Ev { EV_START, EV_EXTRACT, EV_HIST, EV_SORT, EV_MARK, EV_COUNTED, EV_REDUCE, EV_DEDUP, EV_END, EV_COUNT };

}  // namespace

// Everything that belongs to ONE batch of whole clusters: what pf_upload builds, the record /
// row buffers of that batch and its result arrays on the device.  A context holds two of these
// so that the upload of sub-batch j+1 and the D2H of sub-batch j-1 can overlap the kernels of
// sub-batch j (pf_submit on a large batch; see submit_pipelined).
struct BatchState {
  bool have_batch = false, executed = false;
  uint32_t n_seqs = 0, n_clusters = 0, n_wide_seqs = 0;
  uint64_t n_words = 0, n_amb_words = 0, n_bases = 0;
  uint32_t n_pos = 0, n_pos_wide = 0;
  PinBuf h_seqs, h_clusters, h_wide_seqs;
  DevBuf d_bases, d_amb, d_ambbits, d_seqs, d_clusters, d_wide_seqs, d_presence;
  WidthState nar, wid;
  // rows (narrow first, then wide)
  DevBuf d_row_cluster, d_row_kmer, d_wrow_kmer, d_row_count, d_row_pattern;
  DevBuf d_cl_pattern;
  DevBuf d_pos_kmer, d_pos_seq, d_pos_cstart, d_pos_gstart, d_pos_flags, d_pos_wide;
  DevBuf d_seq_rec_off, d_tile_first_seq;
  PinBuf h_seq_rec_off, h_tile_first_seq;
  std::vector<std::pair<uint32_t, uint32_t>> nar_ranges;   // narrow record range of every cluster
  uint32_t n_items = 0;          // (cluster, block) work items of the batch (block aggregation)
  DevBuf d_seq_lite, d_cblk, d_item_base, d_item_cluster, d_plan_total, d_bsum_slot, d_slice_seq, d_item_desc;
  PinBuf h_plan;
  uint64_t kp_base = 0, cp_base = 0;   // pool sizes before the batch
  bool rows_prefetched = false;
  uint64_t row_cap = 0;          // capacity of the row arrays above
  cudaEvent_t ev[EV_COUNT]{};    // stage timestamps of the batch's pf_execute
  PinBuf h_done;                 // pinned mirror of the batch's new k-mer pattern count (K4)
};

struct pf_ctx : BatchState {
  pf_params prm{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // D2H of finished row arrays while K4 still runs
  cudaStream_t up_stream = nullptr;     // H2D of the next sub-batch while the current one computes
  cudaEvent_t ev_rows = nullptr;
  std::string err;
  uint32_t W = 0, Wk = 0;
  std::unordered_map<uint32_t, std::pair<uint32_t, uint32_t>> maf_cache;
  std::mutex maf_mu;       // the upload helper thread of the pipelined submit shares the cache

  BatchState alt;          // the other batch slot (pipelined submit)
  DevBuf d_counters;       // u32[C_COUNT]: tickets, n_runs, errors, totals
  PinBuf h_counters;
  DevBuf d_bsum;
  DevBuf d_cand;
  DevBuf d_rep, d_slot_of, d_winner;
  DevBuf d_cl_rep, d_cl_slot, d_cl_winner;
  bool partition = true;   // mode 0: few radix passes + shared-memory hash grouping (k3_local)
  bool use_direct = true;  // S <= 1024: bitsets for every distinct key in shared memory
  bool runs_from_hist = false;   // one pass: prefix-runs are the digit buckets of the histogram
  uint32_t local_tile = 0;       // records per tile of the local reduce (4096 direct / 2048 general)
  bool fused = false;            // K1 fused into the histogram and the first pass (no record write in K1)
  DevBuf d_digests;
  int extra_bits = 0;      // sort bits added after a table overflow (sticky)
  double row_ratio = 1.0 / 48;   // surviving rows per record, learned from earlier batches
  uint64_t unique_last = 0;
  uint32_t rescued_last = 0;
  // block aggregation (k3_block.cuh): no records, partial (k-mer, bitset) rows per position block
  bool block_mode = false;       // S <= 1024 and not disabled: kA_block_aggregate + kB1..kB3
  uint32_t block_windows = 16;   // windows per position block (= kBlkRun)
  uint32_t blk_slots = 1024, blk_cap = 448, blk_cslots = 128;   // shared memory of kA: k-mer key slots / rows, chunk slots
  uint32_t n_slices = 1, slice_samples = 0, Ws = 0;   // sample slices of the block engine (S > 1024): slices,
                                                       // samples per slice, bitset words of a partial row
  uint32_t block_fallbacks = 0;
  double partial_ratio = 1.0 / 16;   // partial rows per window, learned from earlier batches
  uint64_t partial_cap = 0;
  uint64_t partials_last = 0;
  bool used_block = false;       // the last batch went through kA/kB
  DevBuf d_slab_base, d_slab_count, d_slab_keys, d_slab_rows,
      d_group_base /* merge-table offsets per (cluster, slice) */, d_mtable, d_pslot,
      d_table2_base, d_table2, d_next, d_pslice, d_cta_cluster, d_slab_cnt, d_rescue[2];
  PatternSpace kp, cp;     // k-mer patterns, cluster patterns
  // pinned results
  PinBuf r_row_cluster, r_row_kmer, r_wrow_kmer, r_row_count, r_row_pattern, r_cl_pattern;
  PinBuf r_new_kp, r_new_cp, r_pos_kmer, r_pos_seq, r_pos_cstart, r_pos_gstart, r_pos_flags,
      r_pos_wide;
  PinBuf r_wrow_cluster, r_wrow_count, r_wrow_pattern;   // pipelined submit: wide rows apart
  cudaEvent_t ev_h2d[2]{}, ev_d2h[2]{};
  pf_stats stats{};
  // the k-mer pattern count of the last pf_execute is folded into kp.n lazily (its K4 may still run)
  bool kp_pending = false;
  uint64_t kp_pending_base = 0;
  const uint32_t* kp_pending_count = nullptr;
  std::atomic<uint32_t> launches{0};
  // pipelined submit (submit_pipelined / collect_pipelined)
  bool prefetch_rows = false;    // start the D2H of the row arrays under K4 (PF_PREFETCH_ROWS=1): off by
                                 // default, the copy engine it occupies delays every small read-back that
                                 // follows (multi-GPU exchange), and large submits are pipelined anyway
  bool pipe_pending = false;     // results of a pipelined submit wait for pf_collect
  uint32_t pipe_subs = 1;        // sub-batches of the last submit
  bool pipe_mode = false;        // inside submit_pipelined: pf_execute leaves the D2H to it
  double pipe_ms[10] = {0};      // stage times summed over the sub-batches
  uint64_t pipe_rows = 0, pipe_wide_rows = 0, pipe_pos = 0, pipe_pos_wide = 0;
  uint32_t pipe_clusters = 0;
  uint64_t pipe_kp_base = 0, pipe_cp_base = 0, pipe_kp_copied = 0;
  uint64_t pipe_row_cap = 0, pipe_wide_cap = 0;
  cudaEvent_t ev_up[2]{}, ev_exec_end[2]{}, ev_out_done[2]{}, ev_pipe[2]{};
  uint32_t pipe_min_seqs = 200000;   // batches with fewer sequences are not split
  uint32_t pipe_target_seqs = 262144;   // sequences per sub-batch
};

namespace {

int fail(pf_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(ctx, PF_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                  __FILE__, __LINE__);                                                 \
  } while (0)

int dev_ensure(pf_ctx* ctx, DevBuf& b, size_t bytes, bool keep = false) {
  if (bytes <= b.cap) return PF_OK;
  size_t want = std::max(bytes, b.cap + b.cap / 2);
  want = (want + 255) & ~size_t(255);
  void* np = nullptr;
  CU(cudaMalloc(&np, want));
  if (keep && b.p && b.cap) {
    CU(cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if (b.p) CU(cudaFree(b.p));
  b.p = np;
  b.cap = want;
  return PF_OK;
}
int pin_ensure(pf_ctx* ctx, PinBuf& b, size_t bytes) {
  if (bytes <= b.cap) return PF_OK;
  size_t want = std::max(bytes, b.cap + b.cap / 2);
  want = (want + 4095) & ~size_t(4095);
  if (b.p) CU(cudaFreeHost(b.p));
  b.p = nullptr; b.cap = 0;
  CU(cudaMallocHost(&b.p, want));
  b.cap = want;
  return PF_OK;
}
#define TRY(x) do { int r_ = (x); if (r_ != PF_OK) return r_; } while (0)

// PF_DEBUG_SYNC=1: synchronise after every stage so a device fault names its kernel.
bool debug_sync(const char* name) {
  static const char* v = getenv("PF_DEBUG_SYNC");
  return v && (v[0] == '1' || strstr(name, v) != nullptr);
}
#define STAGE(name)                                                                      \
  do {                                                                                   \
    if (debug_sync(name)) {                                                                  \
      cudaError_t e_ = cudaStreamSynchronize(ctx->stream);                               \
      if (e_ != cudaSuccess)                                                             \
        return fail(ctx, PF_ERR_CUDA, "stage %s: %s", name, cudaGetErrorString(e_));     \
    }                                                                                    \
  } while (0)

inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
constexpr int kGridPersist = 148 * 4;
constexpr uint32_t kBlkMaxSmem = 220u * 1024u;   // largest kA table we ask for

// counters layout in d_counters
enum { C_TICKET_N = 0, C_TICKET_W = 1, C_RUNS_N = 2, C_RUNS_W = 3, C_ERR = 4, C_ROWS_N = 5,
       C_ROWS_W = 6, C_NEW_KP = 7, C_NEW_CP = 8, C_TICKET_MARK_N = 9, C_TICKET_MARK_W = 10,
       C_LOCAL = 11 /* LC_COUNT words: rows, unique, table overflow, row overflow, rescue runs,
                        partial rows, partial overflow */, C_TICKET_MERGE = 18, C_COUNT = 24 };
static_assert(C_LOCAL + LC_COUNT <= C_TICKET_MERGE, "counter layout");

bool keep_count(double maf, uint32_t c, uint32_t n) {
  double af = (double)c / (double)n;      // numpy: vec.sum() / vec.shape[0]
  if (af >= 0.5) af = 1 - af;
  return !(af < maf);
}

}  // namespace

extern "C" int pf_maf_window(double maf, uint32_t n, uint32_t* lo, uint32_t* hi) {
  if (!lo || !hi) return PF_ERR_INVALID;
  if (n == 0) { *lo = 0; *hi = 0; return 1; }   // 0/0 = NaN: neither comparison fires
  const uint32_t mid = (uint32_t)(((uint64_t)n + 1) / 2);   // first c with c/n >= 0.5
  // rising part [0, mid): kept counts form a suffix
  uint32_t a = 0, b = mid;                 // first kept in [a, b) or b
  while (a < b) { uint32_t m = a + (b - a) / 2; if (keep_count(maf, m, n)) b = m; else a = m + 1; }
  const uint32_t rise_lo = a;              // == mid if none
  // falling part [mid, n]: kept counts form a prefix
  uint32_t x = mid, y = n + 1;             // first dropped in [x, y) or y
  while (x < y) { uint32_t m = x + (y - x) / 2; if (!keep_count(maf, m, n)) y = m; else x = m + 1; }
  const uint32_t fall_end = x;             // kept: [mid, fall_end)
  const bool rise = rise_lo < mid, fall = fall_end > mid;
  if (!rise && !fall) { *lo = 1; *hi = 0; return 0; }
  *lo = rise ? rise_lo : mid;
  *hi = fall ? fall_end - 1 : mid - 1;
  return 1;
}

extern "C" uint32_t pf_pattern_words(uint32_t n_samples) { return (n_samples + 31u) / 32u; }
extern "C" int pf_abi_version(void) { return PF_ABI_VERSION; }
extern "C" uint32_t pf_struct_size(int which) {
  switch (which) {
    case 0: return (uint32_t)sizeof(pf_params);
    case 1: return (uint32_t)sizeof(pf_seq_desc);
    case 2: return (uint32_t)sizeof(pf_cluster_desc);
    case 3: return (uint32_t)sizeof(pf_batch);
    case 4: return (uint32_t)sizeof(pf_batch_result);
    case 5: return (uint32_t)sizeof(pf_stats);
    case 6: return (uint32_t)sizeof(pf_synth_params);
    default: return 0;
  }
}
extern "C" const char* pf_last_error(const pf_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}
extern "C" uint32_t pf_kmer_pattern_words(const pf_ctx* ctx) { return ctx ? ctx->Wk : 0; }
extern "C" void* pf_stream(pf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int pf_create(pf_ctx** out, int device, const pf_params* p) {
  pf_ctx* ctx = nullptr;
  if (!out || !p) return fail(nullptr, PF_ERR_INVALID, "pf_create: null argument");
  *out = nullptr;
  if (p->abi_version != PF_ABI_VERSION)
    return fail(nullptr, PF_ERR_INVALID, "pf_create: ABI version %u != %u", p->abi_version, PF_ABI_VERSION);
  if (p->k < 1 || p->k > 32)
    return fail(nullptr, PF_ERR_UNSUPPORTED, "k=%u unsupported: the 64-bit 2-bit path covers 1..32", p->k);
  if (p->n_samples < 1) return fail(nullptr, PF_ERR_INVALID, "n_samples must be >= 1");
  if (p->sort_bits != 0 && (p->sort_bits % 8 != 0 || p->sort_bits < 8 || p->sort_bits > 64))
    return fail(nullptr, PF_ERR_INVALID, "sort_bits must be 0 or a multiple of 8 in 8..64");
  if (p->mode > 1) return fail(nullptr, PF_ERR_INVALID, "mode must be 0 (partition) or 1 (full sort)");
  if (!(p->maf <= 0.5) || p->maf < 0)
    return fail(nullptr, PF_ERR_INVALID, "--maf should be in [0, 0.5]");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, PF_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(nullptr, PF_ERR_INVALID, "device %d out of range", device);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, PF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10)
    return fail(nullptr, PF_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only",
                prop.major, prop.minor);
  ctx = new pf_ctx();
  ctx->prm = *p;
  ctx->device = device;
  ctx->W = pf_pattern_words(p->n_samples);
  ctx->Wk = ctx->W + (p->consider_missing ? 1u : 0u);
  ctx->kp.key_words = ctx->Wk;
  ctx->cp.key_words = ctx->W;
  // partition mode needs (slot:13 | sample:19) pair words and a bitset row that fits the pool
  ctx->partition = (p->mode == 0) && p->n_samples < kLocalMaxSamples && ctx->W <= (uint32_t)kLocalPoolWords;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, PF_ERR_CUDA, "cudaStreamCreate failed");
  }
  cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_exec_end[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_out_done[i], cudaEventDisableTiming);
    cudaEventCreate(&ctx->ev_pipe[i]);
  }
  if (const char* e = getenv("PF_PREFETCH_ROWS")) ctx->prefetch_rows = atoi(e) != 0;
  if (const char* e = getenv("PF_PIPELINE_SEQS")) {      // 0 disables the pipelined submit
    const long v = atol(e);
    if (v <= 0) ctx->pipe_min_seqs = 0xffffffffu;
    else { ctx->pipe_target_seqs = (uint32_t)std::max<long>(1024, v); ctx->pipe_min_seqs = ctx->pipe_target_seqs + ctx->pipe_target_seqs / 2; }
  }
  cudaEventCreateWithFlags(&ctx->ev_rows, cudaEventDisableTiming);
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  for (auto& ev : ctx->alt.ev) cudaEventCreate(&ev);
  for (auto& ev : ctx->ev_h2d) cudaEventCreate(&ev);
  for (auto& ev : ctx->ev_d2h) cudaEventCreate(&ev);
  // opt in to > 48 KB dynamic shared memory for the sort passes
  cudaFuncSetAttribute(k2_onesweep_pass<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(SortSmem<uint64_t>));
  cudaFuncSetAttribute(k2_onesweep_pass<Key128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(SortSmem<Key128>));
  cudaFuncSetAttribute(k2_scatter_pass<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(ScatterSmem<uint64_t>));
  cudaFuncSetAttribute(k2_extract_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(ScatterSmem<uint64_t>));
  cudaFuncSetAttribute(k2_extract_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)sizeof(ScatterSmem<uint64_t>));
  cudaFuncSetAttribute(k3_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LocalSmem));
  cudaFuncSetAttribute(k3_local_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DirectSmem));
  ctx->use_direct = ctx->W <= kDirectMaxWords;
  {
    // block aggregation: debug_flags bit 1 disables it; PF_BLOCK_WINDOWS / PF_BLOCK_SMEM_KB tune it
    ctx->block_mode = ctx->partition && !(p->debug_flags & 2u);
    if (ctx->W > kDirectMaxWords) {            // S > 1024: partial rows per slice of 512 samples
      ctx->slice_samples = 512;
      ctx->n_slices = (p->n_samples + 511u) / 512u;
      ctx->Ws = 16;
      if (ctx->n_slices > 0xffffu) ctx->block_mode = false;
    } else {
      ctx->slice_samples = 32u * ctx->W;
      ctx->n_slices = 1;
      ctx->Ws = ctx->W;
    }
    if (const char* e = getenv("PF_BLOCK_SLICED")) if (atoi(e) == 0 && ctx->n_slices > 1) ctx->block_mode = false;
    if (const char* e = getenv("PF_BLOCK_MODE")) ctx->block_mode = ctx->block_mode && atoi(e) != 0;
    // tables of kA: one k-mer slot per expected distinct k-mer of a 16-window run at ~50 % load
    // (about one haplotype per 30 samples and position), a quarter as many chunk slots;
    // overflowing blocks are rerun with both doubled
    ctx->block_windows = (uint32_t)kBlkRun;
    // (about one new k-mer per sample and run of 16 windows, plus the haplotypes' own)
    uint32_t slots = 256;
    const uint32_t s_eff = std::min<uint32_t>(p->n_samples, ctx->slice_samples);   // samples a kA block sees
    while (slots < 2u * s_eff && slots < 4096u) slots *= 2;
    if (const char* e = getenv("PF_BLOCK_SLOTS")) {
      const int v = atoi(e);
      if (v >= 64 && v <= 8192 && (v & (v - 1)) == 0) slots = (uint32_t)v;
    }
    uint32_t cap = std::min<uint32_t>(slots * 13u / 16u, std::max<uint32_t>(96u, s_eff * 9u / 10u));
    if (const char* e = getenv("PF_BLOCK_CAP")) {
      const int v = atoi(e);
      if (v >= 32 && (uint32_t)v <= slots * 13u / 16u) cap = (uint32_t)v;
    }
    uint32_t cslots = std::max<uint32_t>(64u, slots / 8u);
    if (const char* e = getenv("PF_BLOCK_CSLOTS")) {
      const int v = atoi(e);
      if (v >= 32 && v <= 8192 && (v & (v - 1)) == 0) cslots = (uint32_t)v;
    }
    while (slots > 64u && blkA_smem_bytes(slots, cap, cslots, ctx->Ws) > kBlkMaxSmem) {
      slots /= 2; cap /= 2; cslots = std::max<uint32_t>(32u, cslots / 2);
    }
    ctx->blk_slots = slots;
    ctx->blk_cap = cap;
    ctx->blk_cslots = cslots;
    if (p->k == 32 && !p->canonical) ctx->block_mode = false;   // all-T k-mer == the empty-slot mark
    cudaFuncSetAttribute(kA_block_aggregate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkMaxSmem);
    cudaFuncSetAttribute(kA_block_aggregate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlkMaxSmem);
  }
  const int k3_smem = (int)(8 * ctx->W * sizeof(uint32_t));
  if (k3_smem > 48 * 1024) {
    cudaFuncSetAttribute(k3_runs<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3_smem);
    cudaFuncSetAttribute(k3_runs<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3_smem);
    cudaFuncSetAttribute(k3_runs<Key128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3_smem);
    cudaFuncSetAttribute(k3_runs<Key128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3_smem);
  }
  if (k3_smem > 200 * 1024) {
    pf_destroy(ctx);
    return fail(nullptr, PF_ERR_UNSUPPORTED, "n_samples=%u needs %d B of shared memory per CTA", p->n_samples, k3_smem);
  }
  *out = ctx;
  return PF_OK;
}

extern "C" void pf_destroy(pf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  auto fd = [](DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; };
  auto fp = [](PinBuf& b) { if (b.p) cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; };
  for (BatchState* bs : {static_cast<BatchState*>(ctx), &ctx->alt}) {
    for (DevBuf* b : {&bs->d_bases, &bs->d_amb, &bs->d_ambbits, &bs->d_seqs, &bs->d_clusters, &bs->d_wide_seqs,
                      &bs->d_presence, &bs->d_row_cluster, &bs->d_row_kmer, &bs->d_wrow_kmer, &bs->d_row_count,
                      &bs->d_row_pattern, &bs->d_cl_pattern, &bs->d_pos_kmer, &bs->d_pos_seq, &bs->d_pos_cstart,
                      &bs->d_pos_gstart, &bs->d_pos_flags, &bs->d_pos_wide, &bs->d_seq_rec_off,
                      &bs->d_tile_first_seq, &bs->d_seq_lite, &bs->d_cblk, &bs->d_item_base, &bs->d_item_cluster,
                      &bs->d_plan_total, &bs->d_bsum_slot, &bs->d_slice_seq, &bs->d_item_desc})
      fd(*b);
    for (WidthState* w : {&bs->nar, &bs->wid}) {
      for (DevBuf* b : {&w->keys[0], &w->keys[1], &w->vals[0], &w->vals[1], &w->tiles, &w->seg_start,
                        &w->seg_hist, &w->lookback, &w->cursors, &w->ltiles, &w->tile_first_run,
                        &w->d_tile_base, &w->d_ltile_base})
        fd(*b);
      fp(w->h_tiles); fp(w->h_seg_start); fp(w->h_ltiles);
    }
    for (PinBuf* b : {&bs->h_seqs, &bs->h_clusters, &bs->h_wide_seqs, &bs->h_seq_rec_off, &bs->h_tile_first_seq,
                      &bs->h_plan, &bs->h_done})
      fp(*b);
  }
  for (DevBuf* b : {&ctx->d_counters, &ctx->d_bsum, &ctx->d_cand, &ctx->d_rep, &ctx->d_slot_of, &ctx->d_winner,
                    &ctx->d_cl_rep, &ctx->d_cl_slot, &ctx->d_cl_winner, &ctx->d_digests, &ctx->d_slab_base,
                    &ctx->d_slab_count, &ctx->d_slab_keys, &ctx->d_slab_rows, &ctx->d_group_base, &ctx->d_mtable,
                    &ctx->d_pslot, &ctx->d_rescue[0], &ctx->d_rescue[1], &ctx->d_table2_base, &ctx->d_table2,
                    &ctx->d_next, &ctx->d_pslice, &ctx->d_cta_cluster, &ctx->d_slab_cnt})
    fd(*b);
  for (PatternSpace* s : {&ctx->kp, &ctx->cp})
    for (DevBuf* b : {&s->pool, &s->table, &s->x_owner, &s->x_pos, &s->x_perm, &s->x_counts, &s->x_unique,
                      &s->x_table, &s->x_rep, &s->x_slot, &s->x_winner})
      fd(*b);
  for (PinBuf* b : {&ctx->h_counters, &ctx->r_row_cluster, &ctx->r_row_kmer, &ctx->r_wrow_kmer, &ctx->r_row_count,
                    &ctx->r_row_pattern, &ctx->r_cl_pattern, &ctx->r_new_kp, &ctx->r_new_cp, &ctx->r_pos_kmer,
                    &ctx->r_pos_seq, &ctx->r_pos_cstart, &ctx->r_pos_gstart, &ctx->r_pos_flags, &ctx->r_pos_wide,
                    &ctx->r_wrow_cluster, &ctx->r_wrow_count, &ctx->r_wrow_pattern})
    fp(*b);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->alt.ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_h2d) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_d2h) if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
    if (ctx->ev_exec_end[i]) cudaEventDestroy(ctx->ev_exec_end[i]);
    if (ctx->ev_out_done[i]) cudaEventDestroy(ctx->ev_out_done[i]);
    if (ctx->ev_pipe[i]) cudaEventDestroy(ctx->ev_pipe[i]);
  }
  if (ctx->ev_rows) cudaEventDestroy(ctx->ev_rows);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

namespace {

int auto_sort_bits(const pf_ctx* ctx, uint32_t max_seg_records, bool narrow) {
  if (ctx->prm.sort_bits) return std::min(64, (int)ctx->prm.sort_bits + (narrow ? ctx->extra_bits : 0));
  if (narrow && ctx->partition) {
    // enough 8-bit passes that a prefix bucket of the largest cluster averages <= 4096 records
    int passes = 1;
    uint64_t buckets = 256;
    while ((uint64_t)max_seg_records / buckets > 4096 && passes < 8) { ++passes; buckets <<= 8; }
    return std::min(64, 8 * passes + ctx->extra_bits);
  }
  int lg = 0;
  while ((1ull << lg) < (uint64_t)std::max<uint32_t>(max_seg_records, 1)) ++lg;
  int bits = ((lg + 12 + 7) / 8) * 8;      // expected shared prefixes per segment <= n / 8192
  return std::min(64, std::max(16, bits));
}

// Build the tile list of one key width from the per-cluster record ranges.
// Tile list of the local reduce (partition mode): 8192-record tiles for the direct
// variant, 2048 for the general one (its exactness guarantee needs <= 2048).
int plan_local_tiles(pf_ctx* ctx, BatchState& B, cudaStream_t st) {
  WidthState& w = B.nar;
  if (!ctx->partition) { w.n_ltiles = 0; return PF_OK; }
  const uint32_t tile = ctx->use_direct ? (uint32_t)kDirectTile : (uint32_t)kLocalTile;
  ctx->local_tile = tile;
  const uint32_t nc = (uint32_t)B.nar_ranges.size();
  TRY(pin_ensure(ctx, w.h_ltiles, std::max<size_t>(1, nc) * 4));
  uint32_t* base = w.h_ltiles.as<uint32_t>();
  uint64_t nl = 0;
  for (uint32_t c = 0; c < nc; ++c) {
    base[c] = (uint32_t)nl;
    nl += cdiv(B.nar_ranges[c].second - B.nar_ranges[c].first, tile);
  }
  w.n_ltiles = (uint32_t)nl;
  if (w.n_ltiles) {
    TRY(dev_ensure(ctx, w.ltiles, (size_t)w.n_ltiles * sizeof(TileDev)));
    TRY(dev_ensure(ctx, w.tile_first_run, ((size_t)w.n_ltiles + 1) * 4));
    TRY(dev_ensure(ctx, w.lookback, std::max<size_t>((size_t)w.n_tiles * kRadix * 4, (size_t)w.n_ltiles * 8)));
    TRY(dev_ensure(ctx, w.d_ltile_base, (size_t)nc * 4));
    CU(cudaMemcpyAsync(w.d_ltile_base.p, base, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    // needs d_clusters: the caller uploads it first
    plan_expand_tiles<<<nc, 128, 0, st>>>(B.d_clusters.as<ClusterDev>(), nc, w.d_ltile_base.as<uint32_t>(),
                                                   tile, 0, w.ltiles.as<TileDev>());
    ctx->launches++;     // (the pinned `base` array must not be rewritten before this copy ran:
                         //  pf_upload ends with a sync, the re-plan in pf_execute syncs itself)
  }
  return PF_OK;
}

int plan_tiles(pf_ctx* ctx, WidthState& w, const std::vector<std::pair<uint32_t, uint32_t>>& ranges, bool narrow) {
  uint64_t n_tiles = 0;
  uint32_t max_seg = 0;
  for (auto& r : ranges) {
    n_tiles += cdiv(r.second - r.first, kSortTile);
    max_seg = std::max(max_seg, r.second - r.first);
  }
  if (max_seg >= (1u << 30))
    return fail(ctx, PF_ERR_INVALID, "a cluster has %u k-mer records; the limit per cluster is 2^30", max_seg);
  w.n_tiles = (uint32_t)n_tiles;
  w.max_seg = max_seg;
  w.sort_bits = auto_sort_bits(ctx, max_seg, narrow);
  w.passes = w.sort_bits / 8;
  TRY(pin_ensure(ctx, w.h_tiles, std::max<size_t>(1, ranges.size()) * 4));
  TRY(pin_ensure(ctx, w.h_seg_start, std::max<size_t>(1, ranges.size()) * sizeof(uint32_t)));
  uint32_t* tb = w.h_tiles.as<uint32_t>();
  uint32_t* ss = w.h_seg_start.as<uint32_t>();
  uint32_t ti = 0;
  for (uint32_t c = 0; c < ranges.size(); ++c) {
    ss[c] = ranges[c].first;
    tb[c] = ti;
    ti += cdiv(ranges[c].second - ranges[c].first, kSortTile);
  }
  return PF_OK;
}

}  // namespace

namespace {
int plan_blocks(pf_ctx* ctx, BatchState& B, cudaStream_t st, bool sync);
int plan_blocks_finish(pf_ctx* ctx, BatchState& B, cudaStream_t st);
int scan_inplace(pf_ctx* ctx, uint32_t* data, uint32_t n, uint32_t* total_dev, cudaStream_t st = nullptr, DevBuf* scratch = nullptr);
}

namespace {
// A sub-range of a caller batch: sequences [s0,s1) of clusters [c0,c1), whose bases are words
// [w0,w1) of the 2-bit plane and [a0,a1) of the 4-bit plane.
struct SubRange { uint32_t s0, s1, c0, c1; uint64_t w0, w1, a0, a1; };

// Validate + plan + H2D of the sub-range into the CURRENT batch slot, all asynchronous on `st`
// (the caller's buffers must stay valid until `st` has passed).  upload_finish completes it.
int upload_async(pf_ctx* ctx, BatchState& B, const pf_batch* full, const SubRange& r, cudaStream_t st) {
  pf_batch view = *full;
  view.seqs = full->seqs ? full->seqs + r.s0 : nullptr;
  view.n_seqs = r.s1 - r.s0;
  view.clusters = full->clusters ? full->clusters + r.c0 : nullptr;
  view.n_clusters = r.c1 - r.c0;
  view.cluster_presence = full->cluster_presence ? full->cluster_presence + (size_t)r.c0 * ctx->W : nullptr;
  view.packed_bases = full->packed_bases ? full->packed_bases + r.w0 : nullptr;
  view.n_words = r.w1 - r.w0;
  view.amb_codes = full->amb_codes ? full->amb_codes + r.a0 : nullptr;
  view.n_amb_words = r.a1 - r.a0;
  const pf_batch* b = &view;
  const uint32_t rc = r.c0;                       // rebase of cluster indices
  const uint64_t rb = r.w0 * 32ull, ra = r.a0 * 16ull;   // ... of base / symbol offsets
  B.have_batch = false;
  B.executed = false;
  const pf_params& P = ctx->prm;
  const uint32_t k = P.k, S = P.n_samples, W = ctx->W;
  if (b->n_seqs && (!b->seqs || !b->packed_bases)) return fail(ctx, PF_ERR_INVALID, "null seqs/packed_bases");
  if (b->n_clusters && (!b->clusters || !b->cluster_presence))
    return fail(ctx, PF_ERR_INVALID, "null clusters/cluster_presence");
  if (b->n_clusters == 0 && b->n_seqs) return fail(ctx, PF_ERR_INVALID, "sequences without clusters");

  TRY(pin_ensure(ctx, B.h_seqs, std::max<size_t>(1, b->n_seqs) * sizeof(SeqDev)));
  TRY(pin_ensure(ctx, B.h_clusters, std::max<size_t>(1, b->n_clusters) * sizeof(ClusterDev)));
  TRY(pin_ensure(ctx, B.h_wide_seqs, std::max<size_t>(1, b->n_seqs) * sizeof(uint32_t)));
  SeqDev* hs = B.h_seqs.as<SeqDev>();
  ClusterDev* hc = B.h_clusters.as<ClusterDev>();
  uint32_t* hw = B.h_wide_seqs.as<uint32_t>();

  const uint32_t mult = P.canonical ? 1u : 2u;
  // the packed plane is the bulk of the transfer: start it before the host-side planning
  const size_t slack_words = 80;
  TRY(dev_ensure(ctx, B.d_bases, (b->n_words + slack_words) * 8));
  CU(cudaEventRecord(ctx->ev_h2d[0], st));
  if (b->n_words) CU(cudaMemcpyAsync(B.d_bases.p, b->packed_bases, b->n_words * 8, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync((char*)B.d_bases.p + b->n_words * 8, 0, slack_words * 8, st));

  // ---- planning, in parallel over chunks of sequences --------------------------------
  // phase A: validate + per-sequence sizes, per-chunk sums; phase B: prefix over chunks;
  // phase C: offsets.  Cluster ranges come from the first sequence of every cluster.
  uint64_t rec = 0, wrec = 0, pos = 0, pwide = 0, bases = 0;
  uint32_t n_wide = 0;
  std::vector<std::pair<uint32_t, uint32_t>> nr(b->n_clusters), wr(b->n_clusters);
  {
    const uint32_t n = b->n_seqs;
    // PF_HOST_THREADS caps the planning threads (several contexts / ranks share the host's cores)
    static const uint32_t host_thr = []() { const char* e = getenv("PF_HOST_THREADS"); const int v = e ? atoi(e) : 0;
                                            return v > 0 ? (uint32_t)v : 16u; }();
    const uint32_t n_thr = std::max(1u, std::min<uint32_t>(std::min(host_thr, std::thread::hardware_concurrency()),
                                                           (n + 65535u) / 65536u));
    struct Part { uint64_t rec = 0, wrec = 0, pos = 0, pwide = 0, bases = 0; uint32_t wide = 0; std::string err; };
    std::vector<Part> parts(n_thr);
    const uint32_t per = (n + n_thr - 1) / std::max(1u, n_thr);
    auto run = [&](auto&& fn) {
      if (n_thr == 1) { fn(0u); return; }
      std::vector<std::thread> th;
      for (uint32_t t = 0; t < n_thr; ++t) th.emplace_back(fn, t);
      for (auto& x : th) x.join();
    };
    auto errf = [](Part& p, const char* fmt, uint32_t i, uint32_t a2 = 0, uint32_t a3 = 0) {
      char buf[256];
      snprintf(buf, sizeof buf, fmt, i, a2, a3);
      p.err = buf;
    };
    run([&](uint32_t t) {
      Part& p = parts[t];
      const uint32_t i0 = std::min(n, t * per), i1 = std::min(n, i0 + per);
      for (uint32_t i = i0; i < i1; ++i) {
        pf_seq_desc q = b->seqs[i];
        q.cluster -= rc; q.base_off -= rb; q.amb_off -= (q.flags & PF_SEQ_AMBIGUOUS) ? ra : 0;
        if (q.cluster >= b->n_clusters) return errf(p, "seq %u: cluster %u out of range", i, q.cluster);
        if (i && q.cluster + rc < b->seqs[i - 1].cluster) return errf(p, "seq %u: clusters must be non-decreasing", i);
        if (q.sample >= S) return errf(p, "seq %u: sample rank %u >= n_samples %u", i, q.sample, S);
        if (i && q.cluster + rc == b->seqs[i - 1].cluster && q.sample < b->seqs[i - 1].sample)
          return errf(p, "seq %u: sample ranks must be non-decreasing inside a cluster", i);
        if (!((b->cluster_presence[(size_t)q.cluster * W + (q.sample >> 5)] >> (q.sample & 31)) & 1u))
          return errf(p, "seq %u: sample %u is not marked present in cluster %u", i, q.sample, q.cluster);
        if (q.base_off & 63u) return errf(p, "seq %u: base_off must be a multiple of 64", i);
        if (q.base_off + q.len > b->n_words * 32ull) return errf(p, "seq %u: bases run past the packed plane", i);
        if (q.strand != 1 && q.strand != -1) return errf(p, "seq %u: strand must be +1/-1", i);
        const bool amb = (q.flags & PF_SEQ_AMBIGUOUS) != 0;
        if (amb) {
          if (!b->amb_codes) return errf(p, "seq %u is ambiguous but amb_codes is NULL", i);
          if (q.amb_off & 31u) return errf(p, "seq %u: amb_off must be a multiple of 32", i);
          if (q.amb_off + q.len > b->n_amb_words * 16ull) return errf(p, "seq %u: symbols run past the 4-bit plane", i);
        }
        const bool target = P.emit_positions && (q.flags & PF_SEQ_TARGET);
        const uint32_t nwin = q.len >= k ? q.len - k + 1 : 0;
        p.rec += (uint64_t)nwin * mult;
        if (target) p.pos += nwin;
        if (amb) { p.wrec += (uint64_t)nwin * mult; p.wide++; if (target) p.pwide += nwin; }
        p.bases += q.len;
      }
    });
    for (auto& p : parts) if (!p.err.empty()) return fail(ctx, PF_ERR_INVALID, "%s", p.err.c_str());
    std::vector<Part> base(n_thr);
    for (uint32_t t = 0; t < n_thr; ++t) {
      base[t].rec = rec; base[t].wrec = wrec; base[t].pos = pos; base[t].pwide = pwide; base[t].wide = n_wide;
      rec += parts[t].rec; wrec += parts[t].wrec; pos += parts[t].pos; pwide += parts[t].pwide;
      n_wide += parts[t].wide; bases += parts[t].bases;
    }
    if (rec >= (1ull << 32) - kSortTile || wrec >= (1ull << 32) - kSortTile)
      return fail(ctx, PF_ERR_INVALID, "batch holds more than 2^32 k-mer records; split it");
    if (pos >= (1ull << 32)) return fail(ctx, PF_ERR_INVALID, "more than 2^32 positional records; split the batch");
    constexpr uint32_t kUnset = 0xffffffffu;
    for (uint32_t c = 0; c < b->n_clusters; ++c) { nr[c] = {kUnset, kUnset}; wr[c] = {kUnset, kUnset}; }
    run([&](uint32_t t) {
      Part o = base[t];
      const uint32_t i0 = std::min(n, t * per), i1 = std::min(n, i0 + per);
      for (uint32_t i = i0; i < i1; ++i) {
        pf_seq_desc q = b->seqs[i];
        q.cluster -= rc; q.base_off -= rb; q.amb_off -= (q.flags & PF_SEQ_AMBIGUOUS) ? ra : 0;
        const bool amb = (q.flags & PF_SEQ_AMBIGUOUS) != 0;
        const bool target = P.emit_positions && (q.flags & PF_SEQ_TARGET);
        const uint32_t nwin = q.len >= k ? q.len - k + 1 : 0;
        if (i == 0 || q.cluster + rc != b->seqs[i - 1].cluster) {     // first sequence of its cluster
          nr[q.cluster].first = (uint32_t)o.rec;
          wr[q.cluster].first = (uint32_t)o.wrec;
        }
        SeqDev& d = hs[i];
        d.base_off = q.base_off; d.amb_off = amb ? q.amb_off : 0; d.len = q.len; d.sample = q.sample;
        d.cluster = q.cluster; d.flags = (target ? 1u : 0u) | (amb ? 2u : 0u);
        d.start = q.start; d.end = q.end; d.offset = q.offset; d.strand = q.strand;
        d.rec_off = (uint32_t)o.rec; d.pos_off = (uint32_t)o.pos; d.wrec_off = (uint32_t)o.wrec;
        d.pwide_off = (uint32_t)o.pwide;
        o.rec += (uint64_t)nwin * mult;
        if (target) o.pos += nwin;
        if (amb) { o.wrec += (uint64_t)nwin * mult; hw[o.wide++] = i; if (target) o.pwide += nwin; }
      }
    });
    // clusters without sequences are empty ranges at the start of the next non-empty one
    uint32_t next_n = (uint32_t)rec, next_w = (uint32_t)wrec;
    for (uint32_t c = b->n_clusters; c-- > 0;) {
      if (nr[c].first == kUnset) { nr[c].first = next_n; wr[c].first = next_w; }
      nr[c].second = next_n; wr[c].second = next_w;
      next_n = nr[c].first; next_w = wr[c].first;
    }
  }
  for (uint32_t c = 0; c < b->n_clusters; ++c) {
    uint32_t np = 0;
    for (uint32_t w = 0; w < W; ++w) {
      uint32_t word = b->cluster_presence[(size_t)c * W + w];
      if (w == W - 1 && (S & 31u)) {
        if (word >> (S & 31u)) return fail(ctx, PF_ERR_INVALID, "cluster %u: presence bits beyond n_samples", c);
      }
      np += (uint32_t)__builtin_popcount(word);
    }
    ClusterDev& d = hc[c];
    d.rec_start = nr[c].first; d.rec_end = nr[c].second;
    d.wrec_start = wr[c].first; d.wrec_end = wr[c].second;
    d.id = b->clusters[c].id; d.n_present = np;
    const uint32_t n = P.consider_missing ? np : S;
    uint32_t lo, hi;
    {
      std::lock_guard<std::mutex> lk(ctx->maf_mu);
      auto it = ctx->maf_cache.find(n);
      if (it == ctx->maf_cache.end()) {
        uint32_t wl, wh;
        pf_maf_window(P.maf, n, &wl, &wh);
        it = ctx->maf_cache.emplace(n, std::make_pair(wl, wh)).first;
      }
      lo = it->second.first; hi = it->second.second;
    }
    // "same as cluster" (panfeed.py:202-204): k-mer bits are a subset of the
    // cluster's, so equality <=> count == n_present; NaN entries never compare equal.
    if (P.cluster_equal_filter && (!P.consider_missing || np == S)) {
      if (np == 0) { lo = 1; hi = 0; }
      else hi = std::min(hi, np - 1);
    }
    if (lo == 0) lo = 1;                    // a k-mer row always has >= 1 sample
    d.lo = lo; d.hi = hi;
  }

  B.n_seqs = b->n_seqs; B.n_clusters = b->n_clusters; B.n_wide_seqs = n_wide;
  B.n_words = b->n_words; B.n_amb_words = n_wide ? b->n_amb_words : 0; B.n_bases = bases;
  B.n_pos = (uint32_t)pos; B.n_pos_wide = (uint32_t)pwide;
  B.nar.n_records = (uint32_t)rec; B.wid.n_records = (uint32_t)wrec;
  TRY(plan_tiles(ctx, B.nar, nr, true));
  TRY(plan_tiles(ctx, B.wid, wr, false));

  // ---- device buffers + H2D ------------------------------------------------
  TRY(dev_ensure(ctx, B.d_seqs, std::max<size_t>(1, b->n_seqs) * sizeof(SeqDev)));
  TRY(dev_ensure(ctx, B.d_clusters, std::max<size_t>(1, b->n_clusters) * sizeof(ClusterDev)));
  TRY(dev_ensure(ctx, B.d_presence, std::max<size_t>(1, (size_t)b->n_clusters * W) * 4));
  TRY(dev_ensure(ctx, ctx->d_counters, C_COUNT * 4));
  TRY(pin_ensure(ctx, ctx->h_counters, C_COUNT * 4 + 32 * 4));
  for (WidthState* w : {&B.nar, &B.wid}) {
    TRY(dev_ensure(ctx, w->tiles, std::max<size_t>(1, w->n_tiles) * sizeof(TileDev)));
    TRY(dev_ensure(ctx, w->seg_start, std::max<size_t>(1, b->n_clusters) * 4));
    TRY(dev_ensure(ctx, w->seg_hist, std::max<size_t>(1, (size_t)b->n_clusters * w->passes * kRadix) * 4));
    TRY(dev_ensure(ctx, w->lookback, std::max<size_t>(1, (size_t)w->n_tiles * kRadix) * 4));
  }
  if (b->n_seqs) CU(cudaMemcpyAsync(B.d_seqs.p, hs, b->n_seqs * sizeof(SeqDev), cudaMemcpyHostToDevice, st));
  if (b->n_clusters) {
    CU(cudaMemcpyAsync(B.d_clusters.p, hc, b->n_clusters * sizeof(ClusterDev), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(B.d_presence.p, b->cluster_presence, (size_t)b->n_clusters * W * 4, cudaMemcpyHostToDevice, st));
  }
  // tile lists are generated on the device from the per-cluster record ranges
  for (WidthState* w : {&B.nar, &B.wid}) {
    if (b->n_clusters) CU(cudaMemcpyAsync(w->seg_start.p, w->h_seg_start.p, b->n_clusters * 4, cudaMemcpyHostToDevice, st));
    if (w->n_tiles) {
      TRY(dev_ensure(ctx, w->d_tile_base, (size_t)b->n_clusters * 4));
      CU(cudaMemcpyAsync(w->d_tile_base.p, w->h_tiles.p, (size_t)b->n_clusters * 4, cudaMemcpyHostToDevice, st));
      plan_expand_tiles<<<b->n_clusters, 128, 0, st>>>(B.d_clusters.as<ClusterDev>(), b->n_clusters,
                                                       w->d_tile_base.as<uint32_t>(), kSortTile,
                                                       w == &B.wid ? 1 : 0, w->tiles.as<TileDev>());
      ctx->launches++;
    }
  }
  B.nar_ranges = nr;
  TRY(plan_local_tiles(ctx, B, st));
  // fused first pass: record index -> sequence lookup tables
  if (b->n_seqs) {
    TRY(dev_ensure(ctx, B.d_seq_rec_off, ((size_t)b->n_seqs + 1) * 4));
    TRY(dev_ensure(ctx, B.d_tile_first_seq, ((size_t)B.nar.n_tiles + 1) * 4));
    plan_seq_rec_off<<<cdiv((uint64_t)b->n_seqs + 1, 256), 256, 0, st>>>(
        B.d_seqs.as<SeqDev>(), b->n_seqs, (uint32_t)rec, B.d_seq_rec_off.as<uint32_t>());
    plan_tile_first_seq<<<cdiv((uint64_t)B.nar.n_tiles + 1, 256), 256, 0, st>>>(
        B.nar.tiles.as<TileDev>(), B.nar.n_tiles, B.d_seq_rec_off.as<uint32_t>(), b->n_seqs,
        B.d_tile_first_seq.as<uint32_t>());
    ctx->launches += 2;
  }
  if (n_wide) {
    const size_t bit_words = (b->n_amb_words + 1) / 2 + 4;
    TRY(dev_ensure(ctx, B.d_amb, (b->n_amb_words + 8) * 8));
    TRY(dev_ensure(ctx, B.d_ambbits, bit_words * 4));
    TRY(dev_ensure(ctx, B.d_wide_seqs, n_wide * 4));
    CU(cudaMemcpyAsync(B.d_amb.p, b->amb_codes, b->n_amb_words * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync((char*)B.d_amb.p + b->n_amb_words * 8, 0, 8 * 8, st));
    CU(cudaMemcpyAsync(B.d_wide_seqs.p, hw, n_wide * 4, cudaMemcpyHostToDevice, st));
    k1_amb_bits<<<cdiv(bit_words, 256), 256, 0, st>>>(B.d_amb.as<uint64_t>(), b->n_amb_words,
                                                       B.d_ambbits.as<uint32_t>(), bit_words);
    ctx->launches++;
  }
  if (ctx->block_mode && b->n_seqs) {
    if (b->n_words >= (1ull << 32)) return fail(ctx, PF_ERR_INVALID, "packed plane holds 2^32 words or more; split the batch");
    TRY(dev_ensure(ctx, B.d_seq_lite, (size_t)b->n_seqs * sizeof(SeqLite)));
    plan_seq_lite<<<cdiv(b->n_seqs, 256), 256, 0, st>>>(B.d_seqs.as<SeqDev>(), b->n_seqs, B.d_seq_lite.as<SeqLite>());
    ctx->launches++;
  }
  CU(cudaEventRecord(ctx->ev_h2d[1], st));
  if (ctx->block_mode) TRY(plan_blocks(ctx, B, st, false));
  CU(cudaGetLastError());
  return PF_OK;
}

// Second half of an upload: wait for the copies and the planning kernels, read n_items back.
int upload_finish(pf_ctx* ctx, BatchState& B, cudaStream_t up) {
  CU(cudaStreamSynchronize(up));
  if (ctx->block_mode) TRY(plan_blocks_finish(ctx, B, up));
  CU(cudaStreamSynchronize(up));
  CU(cudaGetLastError());
  B.have_batch = true;
  return PF_OK;
}
}  // namespace

extern "C" int pf_upload(pf_ctx* ctx, const pf_batch* b) {
  if (!ctx) return PF_ERR_INVALID;
  if (!b) return fail(ctx, PF_ERR_INVALID, "pf_upload: null batch");
  CU(cudaSetDevice(ctx->device));
  if (ctx->pipe_pending) return fail(ctx, PF_ERR_STATE, "pf_upload: results of the previous pf_submit were not collected");
  SubRange r{0, b->n_seqs, 0, b->n_clusters, 0, b->n_words, 0, b->n_amb_words};
  TRY(upload_async(ctx, *ctx, b, r, ctx->stream));
  // caller buffers may be pageable: upload_finish makes sure the copies have consumed them
  return upload_finish(ctx, *ctx, ctx->stream);
}

namespace {

template <typename KeyT>
int hist_width(pf_ctx* ctx, WidthState& w, bool fused = false) {
  if (w.n_records == 0) return PF_OK;
  cudaStream_t st = ctx->stream;
  const int shift0 = KeyTraits<KeyT>::kBits - w.sort_bits;
  const uint32_t n_seg = ctx->n_clusters;
  CU(cudaMemsetAsync(w.seg_hist.p, 0, (size_t)n_seg * w.passes * kRadix * 4, st));
  if (fused) {
    const uint32_t chunks = cdiv(ctx->n_seqs, kHistSeqsPerChunk);
    const uint32_t grid = std::min<uint32_t>(cdiv(chunks, kK1Warps), 148 * 8);
    if (ctx->prm.canonical)
      k1_histogram_fused<true><<<grid, kK1Warps * 32, 0, st>>>(
          ctx->d_bases.as<uint64_t>(), ctx->d_ambbits.as<uint32_t>(), ctx->d_seqs.as<SeqDev>(), ctx->n_seqs,
          (int)ctx->prm.k, w.passes, shift0, w.seg_hist.as<uint32_t>());
    else
      k1_histogram_fused<false><<<grid, kK1Warps * 32, 0, st>>>(
          ctx->d_bases.as<uint64_t>(), ctx->d_ambbits.as<uint32_t>(), ctx->d_seqs.as<SeqDev>(), ctx->n_seqs,
          (int)ctx->prm.k, w.passes, shift0, w.seg_hist.as<uint32_t>());
    const uint32_t rows = n_seg * w.passes;
    k2_scan_histogram<<<cdiv(rows, 8), 256, 0, st>>>(w.seg_hist.as<uint32_t>(), w.seg_start.as<uint32_t>(), rows, w.passes);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return PF_OK;
  }
  const uint32_t hist_ctas = std::min<uint32_t>(w.n_tiles, 148 * 8);
  const uint32_t per = cdiv(w.n_tiles, hist_ctas);
  k2_histogram<KeyT><<<cdiv(w.n_tiles, per), 256, 0, st>>>(w.keys[0].as<KeyT>(), w.tiles.as<TileDev>(),
                                                           w.n_tiles, per, w.passes, shift0,
                                                           w.seg_hist.as<uint32_t>());
  const uint32_t rows = n_seg * w.passes;
  k2_scan_histogram<<<cdiv(rows, 8), 256, 0, st>>>(w.seg_hist.as<uint32_t>(), w.seg_start.as<uint32_t>(), rows, w.passes);
  ctx->launches += 2;
  CU(cudaGetLastError());
  return PF_OK;
}

template <typename KeyT>
int passes_width(pf_ctx* ctx, WidthState& w, int ticket_idx, bool unstable = false, bool fused = false) {
  if (w.n_records == 0) { w.final_buf = 0; return PF_OK; }
  cudaStream_t st = ctx->stream;
  const int shift0 = KeyTraits<KeyT>::kBits - w.sort_bits;
  uint32_t* counters = ctx->d_counters.as<uint32_t>();
  int src = 0;
  for (int p = 0; p < w.passes; ++p) {
    const bool atomic_pass = (fused || unstable) && p == 0;
    if (atomic_pass) {
      const size_t hist_bytes = (size_t)ctx->n_clusters * w.passes * kRadix * 4;
      TRY(dev_ensure(ctx, w.cursors, std::max<size_t>(hist_bytes, 4)));
      CU(cudaMemcpyAsync(w.cursors.p, w.seg_hist.p, hist_bytes, cudaMemcpyDeviceToDevice, st));
    } else {
      CU(cudaMemsetAsync(w.lookback.p, 0, (size_t)w.n_tiles * kRadix * 4, st));
      CU(cudaMemsetAsync(counters + ticket_idx, 0, 4, st));
    }
    if (fused && p == 0) {
      if (ctx->prm.canonical)
        k2_extract_scatter<true><<<w.n_tiles, kSortThreads, sizeof(ScatterSmem<uint64_t>), st>>>(
            ctx->d_bases.as<uint64_t>(), ctx->d_ambbits.as<uint32_t>(), ctx->d_seqs.as<SeqDev>(),
            ctx->d_seq_rec_off.as<uint32_t>(), ctx->n_seqs, ctx->d_tile_first_seq.as<uint32_t>(), (int)ctx->prm.k,
            w.keys[src ^ 1].template as<uint64_t>(), w.vals[src ^ 1].template as<uint32_t>(), w.tiles.template as<TileDev>(),
            w.n_tiles, w.cursors.template as<uint32_t>(), w.passes, shift0);
      else
        k2_extract_scatter<false><<<w.n_tiles, kSortThreads, sizeof(ScatterSmem<uint64_t>), st>>>(
            ctx->d_bases.as<uint64_t>(), ctx->d_ambbits.as<uint32_t>(), ctx->d_seqs.as<SeqDev>(),
            ctx->d_seq_rec_off.as<uint32_t>(), ctx->n_seqs, ctx->d_tile_first_seq.as<uint32_t>(), (int)ctx->prm.k,
            w.keys[src ^ 1].template as<uint64_t>(), w.vals[src ^ 1].template as<uint32_t>(), w.tiles.template as<TileDev>(),
            w.n_tiles, w.cursors.template as<uint32_t>(), w.passes, shift0);
    } else if (unstable && p == 0)   // LSD: only the first pass may ignore the incoming order
      k2_scatter_pass<KeyT><<<w.n_tiles, kSortThreads, sizeof(ScatterSmem<KeyT>), st>>>(
          w.keys[src].as<KeyT>(), w.vals[src].as<uint32_t>(), w.keys[src ^ 1].as<KeyT>(),
          w.vals[src ^ 1].as<uint32_t>(), w.tiles.as<TileDev>(), w.n_tiles, w.cursors.as<uint32_t>(), p,
          w.passes, shift0 + 8 * p);
    else
      k2_onesweep_pass<KeyT><<<w.n_tiles, kSortThreads, sizeof(SortSmem<KeyT>), st>>>(
          w.keys[src].as<KeyT>(), w.vals[src].as<uint32_t>(), w.keys[src ^ 1].as<KeyT>(),
          w.vals[src ^ 1].as<uint32_t>(), w.tiles.as<TileDev>(), w.n_tiles, w.seg_hist.as<uint32_t>(), p,
          w.passes, shift0 + 8 * p, w.lookback.as<uint32_t>(), counters + ticket_idx, counters + C_ERR);
    ctx->launches++;
    src ^= 1;
  }
  w.final_buf = src;
  CU(cudaGetLastError());
  return PF_OK;
}

template <typename KeyT>
int mark_width(pf_ctx* ctx, WidthState& w, int ticket_idx, int runs_idx, bool local_tiles) {
  if (w.n_records == 0) return PF_OK;
  cudaStream_t st = ctx->stream;
  uint32_t* counters = ctx->d_counters.as<uint32_t>();
  const int other = w.final_buf ^ 1;
  // the idle ping-pong buffers hold the run lists: keys[other] = run_start | run_seg, vals[other] = nrows
  uint32_t* run_start = w.keys[other].as<uint32_t>();
  uint32_t* run_seg = run_start + ((size_t)w.n_records + 1);
  const uint32_t nt = local_tiles ? w.n_ltiles : w.n_tiles;
  CU(cudaMemsetAsync(w.lookback.p, 0, (size_t)nt * 8, st));
  CU(cudaMemsetAsync(counters + ticket_idx, 0, 4, st));
  if (local_tiles && ctx->local_tile == (uint32_t)kDirectTile)
    k3_mark_runs<KeyT, kDirectTile / kSortThreads><<<nt, kSortThreads, 0, st>>>(
        w.keys[w.final_buf].as<KeyT>(), w.ltiles.as<TileDev>(), nt, w.sort_bits, run_start, run_seg,
        w.tile_first_run.as<uint32_t>(), w.lookback.as<uint64_t>(), counters + ticket_idx,
        counters + runs_idx, counters + C_ERR);
  else if (local_tiles)
    k3_mark_runs<KeyT, kLocalItems><<<nt, kSortThreads, 0, st>>>(
        w.keys[w.final_buf].as<KeyT>(), w.ltiles.as<TileDev>(), nt, w.sort_bits, run_start, run_seg,
        w.tile_first_run.as<uint32_t>(), w.lookback.as<uint64_t>(), counters + ticket_idx,
        counters + runs_idx, counters + C_ERR);
  else
    k3_mark_runs<KeyT, kSortItems><<<nt, kSortThreads, 0, st>>>(
        w.keys[w.final_buf].as<KeyT>(), w.tiles.as<TileDev>(), nt, w.sort_bits, run_start, run_seg,
        nullptr, w.lookback.as<uint64_t>(), counters + ticket_idx, counters + runs_idx, counters + C_ERR);
  ctx->launches++;
  CU(cudaGetLastError());
  return PF_OK;
}

int scan_inplace(pf_ctx* ctx, uint32_t* data, uint32_t n, uint32_t* total_dev, cudaStream_t st, DevBuf* scratch) {
  if (!st) st = ctx->stream;
  DevBuf& bsum = scratch ? *scratch : ctx->d_bsum;
  const uint32_t nb = std::max(1u, cdiv(n, kScanBlock));
  TRY(dev_ensure(ctx, bsum, ((size_t)nb + 1) * 4));
  scan_block_sums<<<nb, 256, 0, st>>>(data, n, bsum.as<uint32_t>());
  scan_of_sums<<<1, 1024, 0, st>>>(bsum.as<uint32_t>(), nb, total_dev);
  scan_apply<<<nb, 256, 0, st>>>(data, n, bsum.as<uint32_t>(), nb);
  ctx->launches += 3;
  CU(cudaGetLastError());
  return PF_OK;
}

template <typename KeyT, bool EMIT>
int runs_width(pf_ctx* ctx, WidthState& w, RowOut out) {
  if (w.n_runs == 0) return PF_OK;
  cudaStream_t st = ctx->stream;
  const int other = w.final_buf ^ 1;
  uint32_t* run_start = w.keys[other].as<uint32_t>();
  uint32_t* run_seg = run_start + ((size_t)w.n_records + 1);
  uint32_t* nrows = w.vals[other].as<uint32_t>();
  const size_t smem = (size_t)8 * ctx->W * 4;
  const uint32_t grid = std::min<uint32_t>(cdiv(w.n_runs, 8), kGridPersist * 2);
  k3_runs<KeyT, EMIT><<<grid, 256, smem, st>>>(w.keys[w.final_buf].as<KeyT>(), w.vals[w.final_buf].as<uint32_t>(),
                                               run_start, run_seg, w.n_runs, w.n_records,
                                               ctx->d_clusters.as<ClusterDev>(), nrows, out);
  ctx->launches++;
  CU(cudaGetLastError());
  return PF_OK;
}

// Ensure the table of a pattern space can take `extra` more patterns at <= 50 % load.
int table_reserve(pf_ctx* ctx, PatternSpace& s, uint64_t extra) {
  cudaStream_t st = ctx->stream;
  const uint64_t need = (s.n + extra) * 2 + 16;
  if (need >= (1ull << 31)) return fail(ctx, PF_ERR_NOMEM, "pattern table would exceed 2^31 slots");
  TRY(dev_ensure(ctx, s.pool, std::max<size_t>(1, (s.n + extra)) * s.key_words * 4, true));
  if (need <= s.table_size) return PF_OK;
  uint32_t size = std::max<uint32_t>(1024, s.table_size);
  while (size < need) size *= 2;
  DevBuf nt;
  TRY(dev_ensure(ctx, nt, (size_t)size * 4));
  CU(cudaMemsetAsync(nt.p, 0xff, (size_t)size * 4, st));
  if (s.n) {
    k4_rehash<<<std::min<uint32_t>(cdiv(s.n, 8), kGridPersist), 256, 0, st>>>(
        s.pool.as<uint32_t>(), (uint32_t)s.n, s.key_words, nt.as<uint32_t>(), size - 1);
    ctx->launches++;
  }
  CU(cudaStreamSynchronize(st));
  if (s.table.p) CU(cudaFree(s.table.p));
  s.table = nt;
  s.table_size = size;
  return PF_OK;
}

// Dedup n candidate keys against a pattern space; ids to `ids_out`, number of
// new patterns to counters[new_idx] (device).
int dedup(pf_ctx* ctx, PatternSpace& s, const uint32_t* cand, uint32_t n, DevBuf& rep, DevBuf& slot_of,
          DevBuf& winner, uint32_t* ids_out, int new_idx) {
  cudaStream_t st = ctx->stream;
  uint32_t* counters = ctx->d_counters.as<uint32_t>();
  if (n == 0) { CU(cudaMemsetAsync(counters + new_idx, 0, 4, st)); return PF_OK; }
  TRY(table_reserve(ctx, s, n));
  TRY(dev_ensure(ctx, rep, (size_t)n * 4));
  TRY(dev_ensure(ctx, slot_of, (size_t)n * 4));
  TRY(dev_ensure(ctx, winner, ((size_t)n + 1) * 4));
  // lanes per row: at most 4 words per lane
  const int L = s.key_words <= 16 ? 4 : s.key_words <= 32 ? 8 : s.key_words <= 64 ? 16 : 32;
  const uint32_t grid = std::min<uint32_t>(cdiv(n, 256 / L), kGridPersist * 4);
#define PF_K4(LL)                                                                                      \
  do {                                                                                                 \
    k4_probe<LL><<<grid, 256, 0, st>>>(cand, n, s.key_words, s.pool.as<uint32_t>(), s.table.as<uint32_t>(), \
                                       s.table_size - 1, rep.as<uint32_t>(), slot_of.as<uint32_t>(),   \
                                       winner.as<uint32_t>());                                         \
    ctx->launches++;                                                                                   \
    TRY(scan_inplace(ctx, winner.as<uint32_t>(), n, counters + new_idx));                              \
    k4_commit<LL><<<grid, 256, 0, st>>>(cand, n, s.key_words, s.pool.as<uint32_t>(), (uint32_t)s.n,     \
                                        s.table.as<uint32_t>(), rep.as<uint32_t>(), slot_of.as<uint32_t>(), \
                                        winner.as<uint32_t>(), ids_out);                               \
    ctx->launches++;                                                                                   \
  } while (0)
  if (L == 4) PF_K4(4); else if (L == 8) PF_K4(8); else if (L == 16) PF_K4(16); else PF_K4(32);
#undef PF_K4
  CU(cudaGetLastError());
  return PF_OK;
}

// A previous pf_execute that was never collected: fold its new-pattern count in.
int finalize_pending(pf_ctx* ctx) {
  if (!ctx->kp_pending) return PF_OK;
  CU(cudaStreamSynchronize(ctx->stream));
  // (a row prefetch of results nobody collected may still be draining on the copy stream: it is
  //  not waited for — the next prefetch queues behind it on that stream, pf_collect syncs it)
  ctx->kp.n = ctx->kp_pending_base + *ctx->kp_pending_count;
  ctx->kp_pending = false;
  return PF_OK;
}

void fill_timings(pf_ctx* ctx, const BatchState* slot = nullptr) {
  const BatchState& B = slot ? *slot : *ctx;
  pf_stats& s = ctx->stats;
  auto ms = [](cudaEvent_t a, cudaEvent_t b) { float m = 0; if (cudaEventElapsedTime(&m, a, b) != cudaSuccess) { cudaGetLastError(); m = 0; } return m; };
  s.ms_h2d = ms(ctx->ev_h2d[0], ctx->ev_h2d[1]);
  s.ms_extract = ms(B.ev[EV_START], B.ev[EV_EXTRACT]);
  s.ms_hist = ms(B.ev[EV_EXTRACT], B.ev[EV_HIST]);
  s.ms_sort = ms(B.ev[EV_HIST], B.ev[EV_SORT]);
  s.ms_mark = ms(B.ev[EV_SORT], B.ev[EV_MARK]);
  s.ms_count = ms(B.ev[EV_MARK], B.ev[EV_COUNTED]);
  s.ms_reduce = ms(B.ev[EV_COUNTED], B.ev[EV_REDUCE]);
  s.ms_dedup = ms(B.ev[EV_REDUCE], B.ev[EV_DEDUP]);
  s.ms_total = ms(B.ev[EV_START], B.ev[EV_END]);
  s.sort_passes = (uint32_t)ctx->nar.passes;
  s.engine = ctx->used_block ? 2u : (ctx->partition ? 0u : 1u);
  s.block_windows = ctx->block_windows;
  s.block_slots = ctx->blk_slots;
  s.partial_rows = ctx->used_block ? ctx->partials_last : 0;
  s.sub_batches = ctx->pipe_subs;
}

int check_device_error(pf_ctx* ctx) {
  const uint32_t e = ctx->h_counters.as<uint32_t>()[C_ERR];
  if (e) return fail(ctx, PF_ERR_INTERNAL, "device watchdog: look-back chain stalled (code %u)", e);
  return PF_OK;
}

}  // namespace

namespace {

int ensure_rows(pf_ctx* ctx, uint64_t rows, uint64_t narrow_rows, bool keep) {
  TRY(dev_ensure(ctx, ctx->d_row_cluster, std::max<size_t>(1, rows) * 4, keep));
  TRY(dev_ensure(ctx, ctx->d_row_count, std::max<size_t>(1, rows) * 4, keep));
  TRY(dev_ensure(ctx, ctx->d_row_pattern, std::max<size_t>(1, rows) * 4, keep));
  TRY(dev_ensure(ctx, ctx->d_row_kmer, std::max<size_t>(1, narrow_rows) * 8, keep));
  TRY(dev_ensure(ctx, ctx->d_cand, std::max<size_t>(1, rows) * ctx->Wk * 4, keep));
  return PF_OK;
}

int launch_k1(pf_ctx* ctx) {
  cudaStream_t st = ctx->stream;
  const pf_params& P = ctx->prm;
  WidthState& N = ctx->nar;
  WidthState& Wd = ctx->wid;
  PosOut po{ctx->d_pos_kmer.as<uint64_t>(), ctx->d_pos_seq.as<uint32_t>(), ctx->d_pos_cstart.as<int32_t>(),
            ctx->d_pos_gstart.as<int32_t>(), ctx->d_pos_flags.as<uint8_t>()};
  if (ctx->n_seqs && N.n_records && !(ctx->fused && ctx->n_pos == 0)) {
    const uint32_t grid = std::min<uint32_t>(cdiv(ctx->n_seqs, kK1Warps), 148 * 8);
#define PF_K1(CANON, REC)                                                                              \
    k1_extract<CANON, REC><<<grid, kK1Warps * 32, 0, st>>>(ctx->d_bases.as<uint64_t>(), ctx->d_ambbits.as<uint32_t>(), \
                                                      ctx->d_seqs.as<SeqDev>(), ctx->n_seqs, (int)P.k,  \
                                                      N.keys[0].as<uint64_t>(), N.vals[0].as<uint32_t>(), po)
    if (ctx->fused) { if (P.canonical) PF_K1(true, false); else PF_K1(false, false); }
    else { if (P.canonical) PF_K1(true, true); else PF_K1(false, true); }
#undef PF_K1
    ctx->launches++;
  }
  if (ctx->n_wide_seqs && Wd.n_records) {
    const uint32_t grid = std::min<uint32_t>(cdiv(ctx->n_wide_seqs, 8), 148 * 8);
    if (P.canonical)
      k1_extract_wide<true><<<grid, 256, 0, st>>>(ctx->d_amb.as<uint64_t>(), ctx->d_seqs.as<SeqDev>(),
                                                 ctx->d_wide_seqs.as<uint32_t>(), ctx->n_wide_seqs, (int)P.k,
                                                 Wd.keys[0].as<Key128>(), Wd.vals[0].as<uint32_t>(),
                                                 ctx->d_pos_wide.as<uint64_t>(), ctx->d_pos_flags.as<uint8_t>());
    else
      k1_extract_wide<false><<<grid, 256, 0, st>>>(ctx->d_amb.as<uint64_t>(), ctx->d_seqs.as<SeqDev>(),
                                                  ctx->d_wide_seqs.as<uint32_t>(), ctx->n_wide_seqs, (int)P.k,
                                                  Wd.keys[0].as<Key128>(), Wd.vals[0].as<uint32_t>(),
                                                  ctx->d_pos_wide.as<uint64_t>(), ctx->d_pos_flags.as<uint8_t>());
    ctx->launches++;
  }
  CU(cudaGetLastError());
  return PF_OK;
}

int launch_local(pf_ctx* ctx, RowOut ro, uint32_t n_rescue = 0) {
  WidthState& N = ctx->nar;
  if (N.n_records == 0) return PF_OK;
  cudaStream_t st = ctx->stream;
  uint32_t* counters = ctx->d_counters.as<uint32_t>();
  const int other = N.final_buf ^ 1;
  const uint32_t* run_start = ctx->runs_from_hist ? N.seg_hist.as<uint32_t>() : N.keys[other].as<uint32_t>();
  const uint32_t* run_seg = ctx->runs_from_hist ? nullptr : N.keys[other].as<uint32_t>() + ((size_t)N.n_records + 1);
  uint32_t* rescue = N.vals[other].as<uint32_t>();       // idle in partition mode
  const uint32_t cap = (uint32_t)std::min<uint64_t>(ctx->row_cap, 0x7fffffffu);
  if (n_rescue) {
    // second launch of the direct variant: one CTA per run of the tiles that overflowed;
    // rows/unique counters keep accumulating
    k3_local_direct<<<n_rescue, kLocalThreads, sizeof(DirectSmem), st>>>(
        N.keys[N.final_buf].as<uint64_t>(), N.vals[N.final_buf].as<uint32_t>(), N.ltiles.as<TileDev>(),
        N.n_ltiles, N.tile_first_run.as<uint32_t>(), run_start, N.n_records,
        ctx->d_clusters.as<ClusterDev>(), ro, cap, counters + C_LOCAL, rescue, rescue, run_seg);
    ctx->launches++;
    CU(cudaGetLastError());
    return PF_OK;
  }
  CU(cudaMemsetAsync(counters + C_LOCAL, 0, 5 * 4, st));
  if (ctx->use_direct)
    k3_local_direct<<<N.n_ltiles, kLocalThreads, sizeof(DirectSmem), st>>>(
        N.keys[N.final_buf].as<uint64_t>(), N.vals[N.final_buf].as<uint32_t>(), N.ltiles.as<TileDev>(),
        N.n_ltiles, N.tile_first_run.as<uint32_t>(), run_start, N.n_records,
        ctx->d_clusters.as<ClusterDev>(), ro, cap, counters + C_LOCAL, rescue, nullptr, run_seg);
  else
    k3_local<<<N.n_ltiles, kLocalThreads, sizeof(LocalSmem), st>>>(
        N.keys[N.final_buf].as<uint64_t>(), N.vals[N.final_buf].as<uint32_t>(), N.ltiles.as<TileDev>(),
        N.n_ltiles, N.tile_first_run.as<uint32_t>(), run_start, N.n_records,
        ctx->d_clusters.as<ClusterDev>(), ro, cap, counters + C_LOCAL);
  ctx->launches++;
  CU(cudaGetLastError());
  return PF_OK;
}

// ---- block aggregation: work items = (cluster, position block) -----------------------
// (re)computes the per-cluster block counts for ctx->block_windows; syncs to learn n_items
int plan_blocks(pf_ctx* ctx, BatchState& B, cudaStream_t st, bool sync) {
  const uint32_t nc = B.n_clusters;
  B.n_items = 0;
  if (nc == 0 || B.n_seqs == 0) return PF_OK;
  TRY(dev_ensure(ctx, B.d_cblk, (size_t)nc * sizeof(ClusterBlk)));
  TRY(dev_ensure(ctx, B.d_item_base, ((size_t)nc + 1) * 4));
  TRY(dev_ensure(ctx, B.d_plan_total, 16));
  TRY(pin_ensure(ctx, B.h_plan, 16));
  plan_cluster_blocks<<<cdiv((uint64_t)nc * 32, 256), 256, 0, st>>>(
      B.d_seqs.as<SeqDev>(), B.n_seqs, nc, (int)ctx->prm.k, ctx->block_windows,
      B.d_cblk.as<ClusterBlk>(), B.d_item_base.as<uint32_t>());
  TRY(scan_inplace(ctx, B.d_item_base.as<uint32_t>(), nc, B.d_plan_total.as<uint32_t>(), st, &B.d_bsum_slot));
  if (ctx->n_slices > 1) {
    TRY(dev_ensure(ctx, B.d_slice_seq, (size_t)nc * (ctx->n_slices + 1) * 4));
    plan_cluster_slices<<<cdiv((uint64_t)nc * 32, 256), 256, 0, st>>>(
        B.d_seqs.as<SeqDev>(), B.d_cblk.as<ClusterBlk>(), nc, ctx->n_slices, ctx->slice_samples,
        B.d_slice_seq.as<uint32_t>());
  }
  // n_items is read back through pinned memory (no copy engine); plan_blocks_finish takes it
  mirror_counters<<<1, 32, 0, st>>>(B.h_plan.as<uint32_t>(), B.d_plan_total.as<uint32_t>(), 1);
  CU(cudaGetLastError());
  if (sync) return plan_blocks_finish(ctx, B, st);
  return PF_OK;
}
// after the stream has passed plan_blocks: the item -> cluster map
int plan_blocks_finish(pf_ctx* ctx, BatchState& B, cudaStream_t st) {
  const uint32_t nc = B.n_clusters;
  if (nc == 0 || B.n_seqs == 0) return PF_OK;
  CU(cudaStreamSynchronize(st));
  B.n_items = B.h_plan.as<uint32_t>()[0];
  TRY(dev_ensure(ctx, B.d_item_cluster, std::max<size_t>(1, B.n_items) * 4));
  plan_expand_owner<<<cdiv((uint64_t)nc * 32, 256), 256, 0, st>>>(B.d_item_base.as<uint32_t>(), nc,
                                                                  B.d_item_cluster.as<uint32_t>());
  const uint64_t n_ka = (uint64_t)B.n_items * ctx->n_slices;
  if (n_ka >= (1ull << 31)) return fail(ctx, PF_ERR_INVALID, "too many (cluster, run, slice) work items; split the batch");
  TRY(dev_ensure(ctx, B.d_item_desc, std::max<size_t>(1, n_ka) * 16));
  if (n_ka)
    plan_item_desc<<<cdiv(n_ka, 256), 256, 0, st>>>(
        B.d_item_base.as<uint32_t>(), B.d_item_cluster.as<uint32_t>(), B.d_cblk.as<ClusterBlk>(),
        ctx->n_slices > 1 ? B.d_slice_seq.as<uint32_t>() : nullptr, B.n_items, ctx->n_slices,
        (uint32_t)kBlkRun, B.d_item_desc.as<uint4>());
  CU(cudaGetLastError());
  return PF_OK;
}

BlkPlan block_plan(const pf_ctx* ctx) {
  BlkPlan bp;
  bp.item_base = ctx->d_item_base.as<uint32_t>();
  bp.item_cluster = ctx->d_item_cluster.as<uint32_t>();
  bp.cblk = ctx->d_cblk.as<ClusterBlk>();
  bp.n_clusters = ctx->n_clusters;
  bp.block_windows = ctx->block_windows;
  bp.slots = ctx->blk_slots;
  bp.cap = ctx->blk_cap;
  bp.cslots = ctx->blk_cslots;
  bp.W = ctx->Ws;
  bp.WP = (ctx->Ws + 3u) & ~3u;
  bp.n_slices = ctx->n_slices;
  bp.slice_samples = ctx->slice_samples;
  bp.slice_seq = ctx->n_slices > 1 ? ctx->d_slice_seq.as<uint32_t>() : nullptr;
  bp.item_desc = ctx->d_item_desc.as<uint4>();
  return bp;
}

// items == nullptr: all (cluster, block) items with the context's table size; else the listed
// items (a rescue launch) with `slots` slots.  Items that overflow are appended to `rescue_out`.
int launch_block_aggregate(pf_ctx* ctx, const uint32_t* items, uint32_t n, uint32_t slots, uint32_t cap,
                           uint32_t cslots, uint32_t* rescue_out) {
  if (n == 0) return PF_OK;
  cudaStream_t st = ctx->stream;
  BlkPlan bp = block_plan(ctx);
  bp.slots = slots;
  bp.cap = cap;
  bp.cslots = cslots;
  const uint32_t cap32 = (uint32_t)std::min<uint64_t>(ctx->partial_cap, 0xfffffff0u);
  uint32_t* counters = ctx->d_counters.as<uint32_t>() + C_LOCAL;
  const uint32_t smem = blkA_smem_bytes(slots, cap, cslots, ctx->Ws);
#define PF_KA(CANON)                                                                                   \
  kA_block_aggregate<CANON><<<n, kBlkThreads, smem, st>>>(                                             \
      ctx->d_bases.as<uint64_t>(), ctx->d_ambbits.as<uint32_t>(), ctx->d_seq_lite.as<SeqLite>(), bp,   \
      (int)ctx->prm.k, ctx->d_slab_keys.as<uint64_t>(), ctx->d_slab_rows.as<uint32_t>(),               \
      ctx->d_slab_base.as<uint32_t>(), ctx->d_slab_count.as<uint32_t>(), ctx->d_slab_cnt.as<uint32_t>(), \
      cap32, counters, items,                                                                          \
      rescue_out)
  if (ctx->prm.canonical) PF_KA(true); else PF_KA(false);
#undef PF_KA
  ctx->launches++;
  CU(cudaGetLastError());
  return PF_OK;
}

// kB1..kB3 (kB1, kB2, kB4, kB5 with sample slices) over the `n_partials` partial rows kA left
int launch_block_merge(pf_ctx* ctx, RowOut ro, uint32_t n_partials) {
  if (ctx->n_items == 0 || n_partials == 0) return PF_OK;
  cudaStream_t st = ctx->stream;
  const uint32_t nc = ctx->n_clusters, ns = ctx->n_slices;
  const uint32_t n_cs = nc * ns;                    // (cluster, slice) merge tables
  const uint32_t n_it = ctx->n_items * ns;          // kA work items
  uint32_t* counters = ctx->d_counters.as<uint32_t>();
  const uint32_t WP = (ctx->Ws + 3u) & ~3u;
  // per-(cluster, slice) tables: 1.5 slots per partial row (+2), offsets by a scan
  TRY(dev_ensure(ctx, ctx->d_group_base, ((size_t)n_cs + 1) * 4));
  if (ns > 1) {
    TRY(dev_ensure(ctx, ctx->d_table2_base, ((size_t)nc + 1) * 4));
    CU(cudaMemsetAsync(ctx->d_table2_base.p, 0, ((size_t)nc + 1) * 4, st));
  }
  plan_merge_tables<<<cdiv((uint64_t)n_cs * 32, 256), 256, 0, st>>>(
      ctx->d_item_base.as<uint32_t>(), nc, ns, ctx->d_slab_count.as<uint32_t>(), ctx->d_group_base.as<uint32_t>(),
      ns > 1 ? ctx->d_table2_base.as<uint32_t>() : nullptr);
  ctx->launches++;
  TRY(scan_inplace(ctx, ctx->d_group_base.as<uint32_t>(), n_cs, ctx->d_plan_total.as<uint32_t>() + 1));
  const uint64_t n_slots = (uint64_t)n_partials + n_partials / 2 + 2ull * n_cs + 16;
  if (n_slots >= (1ull << 32)) return fail(ctx, PF_ERR_INVALID, "too many partial rows for one batch; split it");
  TRY(dev_ensure(ctx, ctx->d_mtable, n_slots * sizeof(MergeEntry)));
  TRY(dev_ensure(ctx, ctx->d_pslot, (size_t)n_partials * 4));
  if (ns > 1) TRY(dev_ensure(ctx, ctx->d_pslice, (size_t)n_partials * 2));
  CU(cudaMemsetAsync(ctx->d_mtable.p, 0xff, n_slots * sizeof(MergeEntry), st));
  const uint32_t g0 = cdiv((uint64_t)n_it * 32, 256);
  kB1_insert<<<g0, 256, 0, st>>>(ctx->d_slab_keys.as<uint64_t>(), ctx->d_slab_base.as<uint32_t>(),
                                 ctx->d_slab_count.as<uint32_t>(), ctx->d_item_cluster.as<uint32_t>(), n_it, ns,
                                 ctx->d_group_base.as<uint32_t>(), ctx->d_mtable.as<MergeEntry>(),
                                 ctx->d_pslot.as<uint32_t>(), ns > 1 ? ctx->d_pslice.as<uint16_t>() : nullptr);
  CU(cudaMemsetAsync(counters + C_LOCAL + LC_RESCUE, 0, 4, st));     // kB2 counts the folded rows there
  kB2_fold<<<cdiv(n_partials, 256), 256, 0, st>>>(n_partials, ctx->d_pslot.as<uint32_t>(),
                                                  ctx->d_mtable.as<MergeEntry>(), ctx->d_slab_rows.as<uint32_t>(),
                                                  ctx->d_slab_cnt.as<uint32_t>(), WP, counters + C_LOCAL);
  const uint32_t cap = (uint32_t)std::min<uint64_t>(ctx->row_cap, 0x7fffffffu);
  if (ns == 1) {
    kB3_emit<<<g0, 256, 0, st>>>(ctx->d_slab_keys.as<uint64_t>(), ctx->d_slab_rows.as<uint32_t>(),
                                 ctx->d_slab_base.as<uint32_t>(), ctx->d_slab_count.as<uint32_t>(),
                                 ctx->d_item_cluster.as<uint32_t>(), ctx->n_items, ctx->d_slab_cnt.as<uint32_t>(),
                                 ctx->d_clusters.as<ClusterDev>(), ro, cap, counters + C_LOCAL, ctx->W, WP);
    ctx->launches += 3;
    CU(cudaGetLastError());
    return PF_OK;
  }
  // ---- sample slices: link the per-slice rows of a k-mer, filter on the summed count, assemble ----
  plan_table2_ctas<<<cdiv(nc, 256), 256, 0, st>>>(ctx->d_table2_base.as<uint32_t>(), nc);
  TRY(scan_inplace(ctx, ctx->d_table2_base.as<uint32_t>(), nc, ctx->d_plan_total.as<uint32_t>() + 2));
  const uint64_t max_ctas = ((uint64_t)n_partials + n_partials / 2) / 256 + 2ull * nc + 2;
  TRY(dev_ensure(ctx, ctx->d_table2, max_ctas * 256 * sizeof(LinkEntry)));
  TRY(dev_ensure(ctx, ctx->d_cta_cluster, max_ctas * 4));
  TRY(dev_ensure(ctx, ctx->d_next, (size_t)n_partials * 4));
  CU(cudaMemsetAsync(ctx->d_table2.p, 0xff, max_ctas * 256 * sizeof(LinkEntry), st));
  plan_expand_owner<<<cdiv((uint64_t)nc * 32, 256), 256, 0, st>>>(ctx->d_table2_base.as<uint32_t>(), nc,
                                                                  ctx->d_cta_cluster.as<uint32_t>());
  kB4_link<<<g0, 256, 0, st>>>(ctx->d_slab_keys.as<uint64_t>(), ctx->d_slab_rows.as<uint32_t>(),
                               ctx->d_slab_base.as<uint32_t>(), ctx->d_slab_count.as<uint32_t>(),
                               ctx->d_item_cluster.as<uint32_t>(), n_it, ns, ctx->d_slab_cnt.as<uint32_t>(),
                               ctx->d_table2_base.as<uint32_t>(), ctx->d_table2.as<LinkEntry>(),
                               ctx->d_next.as<uint32_t>(), WP);
  // grid of kB5 = CTAs of all cross-slice tables: one small read-back
  TRY(pin_ensure(ctx, ctx->h_plan, 16));
  mirror_counters<<<1, 32, 0, st>>>(ctx->h_plan.as<uint32_t>() + 2, ctx->d_plan_total.as<uint32_t>() + 2, 1);
  CU(cudaStreamSynchronize(st));
  const uint32_t n_ctas = ctx->h_plan.as<uint32_t>()[2];
  if (n_ctas > max_ctas) return fail(ctx, PF_ERR_INTERNAL, "cross-slice table larger than planned");
  if (n_ctas)
    kB5_emit<<<n_ctas, 256, 0, st>>>(ctx->d_table2.as<LinkEntry>(), ctx->d_cta_cluster.as<uint32_t>(),
                                     ctx->d_next.as<uint32_t>(), ctx->d_pslice.as<uint16_t>(),
                                     ctx->d_slab_rows.as<uint32_t>(), ctx->d_clusters.as<ClusterDev>(), ro, cap,
                                     counters + C_LOCAL, ctx->W, ctx->Ws, WP);
  ctx->launches += 7;
  CU(cudaGetLastError());
  return PF_OK;
}

// partition mode, one pass: no key read is needed to find the prefix-runs
int tiles_from_hist(pf_ctx* ctx) {
  WidthState& N = ctx->nar;
  if (N.n_records == 0) return PF_OK;
  k3_tiles_from_hist<<<cdiv((uint64_t)N.n_ltiles + 1, 256), 256, 0, ctx->stream>>>(
      N.ltiles.as<TileDev>(), N.n_ltiles, N.seg_hist.as<uint32_t>(), ctx->n_clusters,
      N.tile_first_run.as<uint32_t>());
  ctx->launches++;
  CU(cudaGetLastError());
  return PF_OK;
}

}  // namespace

extern "C" int pf_execute(pf_ctx* ctx) {
  if (!ctx) return PF_ERR_INVALID;
  if (!ctx->have_batch) return fail(ctx, PF_ERR_STATE, "pf_execute: no batch uploaded");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t launches0 = ctx->launches;
  WidthState& N = ctx->nar;
  WidthState& Wd = ctx->wid;
  uint32_t* counters = ctx->d_counters.as<uint32_t>();
  uint32_t* hcnt = ctx->h_counters.as<uint32_t>();
  ctx->executed = false;
  ctx->cp_base = ctx->cp.n;
  const bool part = ctx->partition;

  // record buffers (ping-pong); the idle one later holds the run lists, so give
  // it room for n+1 run starts + n run segments (8n+4 bytes <= 8n+16)
  bool blk = ctx->block_mode && part && N.n_records > 0 && ctx->n_items > 0;
  const uint32_t n_ka_items = ctx->n_items * ctx->n_slices;
  auto ensure_records = [&]() -> int {     // the record path needs the narrow ping-pong buffers
    for (int i = 0; i < 2; ++i) {
      TRY(dev_ensure(ctx, N.keys[i], ((size_t)N.n_records + 2) * 8));
      TRY(dev_ensure(ctx, N.vals[i], ((size_t)N.n_records + 2) * 4));
    }
    return PF_OK;
  };
  if (!blk) TRY(ensure_records());
  for (int i = 0; i < 2; ++i) {
    TRY(dev_ensure(ctx, Wd.keys[i], ((size_t)Wd.n_records + 2) * 16));
    TRY(dev_ensure(ctx, Wd.vals[i], ((size_t)Wd.n_records + 2) * 4));
  }
  if (ctx->n_pos) {
    TRY(dev_ensure(ctx, ctx->d_pos_kmer, (size_t)ctx->n_pos * 8));
    TRY(dev_ensure(ctx, ctx->d_pos_seq, (size_t)ctx->n_pos * 4));
    TRY(dev_ensure(ctx, ctx->d_pos_cstart, (size_t)ctx->n_pos * 4));
    TRY(dev_ensure(ctx, ctx->d_pos_gstart, (size_t)ctx->n_pos * 4));
    TRY(dev_ensure(ctx, ctx->d_pos_flags, (size_t)ctx->n_pos));
  }
  if (ctx->n_pos_wide) TRY(dev_ensure(ctx, ctx->d_pos_wide, (size_t)ctx->n_pos_wide * 16));
  TRY(dev_ensure(ctx, ctx->d_cl_pattern, std::max<size_t>(1, ctx->n_clusters) * 4));
  if (part) {
    ctx->row_cap = std::max<uint64_t>(ctx->row_cap, std::max<uint64_t>(
        65536, (uint64_t)(ctx->row_ratio * 1.3 * (double)N.n_records) + 4096));
    TRY(ensure_rows(ctx, ctx->row_cap, ctx->row_cap, false));
  }

  CU(cudaMemsetAsync(counters, 0, C_COUNT * 4, st));
  CU(cudaEventRecord(ctx->ev[EV_START], st));

  // ---- cluster rows (int64 namespace) first: their ids are the NaN-plane word
  //      of k-mer pattern keys in cluster-absent mode ------------------------------
  TRY(dedup(ctx, ctx->cp, ctx->d_presence.as<uint32_t>(), ctx->n_clusters, ctx->d_cl_rep, ctx->d_cl_slot,
            ctx->d_cl_winner, ctx->d_cl_pattern.as<uint32_t>(), C_NEW_CP));
  STAGE("k4 cluster rows");

  RowOut ro{};
  ro.key_words = ctx->Wk; ro.pattern_words = ctx->W;
  ro.cluster_pattern = ctx->d_cl_pattern.as<uint32_t>();

  const uint64_t n_windows = N.n_records / (ctx->prm.canonical ? 1u : 2u);
  ctx->used_block = false;
  for (int attempt = 0;; ++attempt) {
    if (blk) {
      // ---- block aggregation: kA (K1+K2+K3 grouping, no records) + kB (merge, filter, rows) ----
      if (attempt > 12) return fail(ctx, PF_ERR_INTERNAL, "block aggregation did not converge");
      ctx->fused = true;                    // K1 proper only emits positional records
      TRY(launch_k1(ctx));
      STAGE("k1_extract");
      CU(cudaEventRecord(ctx->ev[EV_EXTRACT], st));
      TRY(hist_width<Key128>(ctx, Wd));
      CU(cudaEventRecord(ctx->ev[EV_HIST], st));
      ctx->partial_cap = std::max<uint64_t>(ctx->partial_cap, std::max<uint64_t>(
          65536, (uint64_t)(ctx->partial_ratio * 1.3 * (double)n_windows) + 4096));
      ctx->partial_cap = std::min<uint64_t>(ctx->partial_cap, 0xfffffff0ull);
      const uint32_t WP = (ctx->Ws + 3u) & ~3u;
      TRY(dev_ensure(ctx, ctx->d_slab_base, std::max<size_t>(1, n_ka_items) * 4));
      TRY(dev_ensure(ctx, ctx->d_slab_count, std::max<size_t>(1, n_ka_items) * 4));
      TRY(dev_ensure(ctx, ctx->d_rescue[0], std::max<size_t>(1, n_ka_items) * 4));
      TRY(dev_ensure(ctx, ctx->d_rescue[1], std::max<size_t>(1, n_ka_items) * 4));
      TRY(dev_ensure(ctx, ctx->d_slab_keys, ctx->partial_cap * 8));
      TRY(dev_ensure(ctx, ctx->d_slab_cnt, ctx->partial_cap * 4));
      TRY(dev_ensure(ctx, ctx->d_slab_rows, ctx->partial_cap * WP * 4));
      CU(cudaMemsetAsync(counters + C_LOCAL, 0, LC_COUNT * 4, st));
      TRY(launch_block_aggregate(ctx, nullptr, n_ka_items, ctx->blk_slots, ctx->blk_cap, ctx->blk_cslots,
                                 ctx->d_rescue[0].as<uint32_t>()));
      STAGE("kA_block_aggregate");
      TRY(passes_width<Key128>(ctx, Wd, C_TICKET_W));
      TRY(mark_width<Key128>(ctx, Wd, C_TICKET_MARK_W, C_RUNS_W, false));
      // blocks holding more distinct k-mers than the table takes are rerun with a table twice
      // the size, then four times, ...; past the largest table the batch takes the record path
      mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
      CU(cudaStreamSynchronize(st));
      TRY(check_device_error(ctx));
      bool too_big = false;
      static const bool dbg = getenv("PF_DEBUG_BLOCK") != nullptr;
      if (dbg)
        fprintf(stderr, "[pf] kA: items %u slots %u cap %u cslots %u -> overflow %u rescue %u partials %u part_ovf %u\n",
                ctx->n_items, ctx->blk_slots, ctx->blk_cap, ctx->blk_cslots, hcnt[C_LOCAL + LC_TABLE_OVERFLOW],
                hcnt[C_LOCAL + LC_RESCUE], hcnt[C_LOCAL + LC_PARTIALS], hcnt[C_LOCAL + LC_PARTIAL_OVERFLOW]);
      {
        uint32_t slots = ctx->blk_slots, cap = ctx->blk_cap, cslots = ctx->blk_cslots;
        int cur = 0;
        const uint32_t first_rescue = hcnt[C_LOCAL + LC_RESCUE];
        while (hcnt[C_LOCAL + LC_TABLE_OVERFLOW] == 1u) {
          const uint32_t n_resc = hcnt[C_LOCAL + LC_RESCUE];
          // twice the key slots, as many rows as then fit (the chunk table stays: a full one only
          // sends runs down the direct path)
          uint32_t slots2 = slots * 2u, cap2 = 0;
          if (slots2 <= 8192u && blkA_smem_bytes(slots2, 0u, cslots, ctx->Ws) < kBlkMaxSmem) {
            const uint32_t per_row = 8u + (ctx->Ws | 1u) * 4u;
            const uint32_t fit = (kBlkMaxSmem - blkA_smem_bytes(slots2, 0u, cslots, ctx->Ws)) / per_row;
            cap2 = std::min<uint32_t>(std::min<uint32_t>(fit, slots2 * 13u / 16u),
                                      std::max<uint32_t>(cap * 2u, slots2 * 5u / 8u));
          }
          if (cap2 <= cap) {        // no more rows with more slots: try all the rows the current slots allow
            slots2 = slots;
            const uint32_t per_row = 8u + (ctx->Ws | 1u) * 4u;
            const uint32_t fit = (kBlkMaxSmem - blkA_smem_bytes(slots, 0u, cslots, ctx->Ws)) / per_row;
            cap2 = std::min<uint32_t>(fit, slots * 13u / 16u);
          }
          if (cap2 <= cap) { too_big = true; break; }
          slots = slots2;
          cap = cap2;
          CU(cudaMemsetAsync(counters + C_LOCAL + LC_TABLE_OVERFLOW, 0, 4, st));
          CU(cudaMemsetAsync(counters + C_LOCAL + LC_RESCUE, 0, 4, st));
          TRY(launch_block_aggregate(ctx, ctx->d_rescue[cur].as<uint32_t>(), n_resc, slots, cap, cslots,
                                     ctx->d_rescue[cur ^ 1].as<uint32_t>()));
          cur ^= 1;
          mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
          CU(cudaStreamSynchronize(st));
          if (dbg)
            fprintf(stderr, "[pf] kA rescue: %u items slots %u cap %u cslots %u -> overflow %u rescue %u\n", n_resc, slots,
                    cap, cslots, hcnt[C_LOCAL + LC_TABLE_OVERFLOW], hcnt[C_LOCAL + LC_RESCUE]);
        }
        // many rescued blocks: start the next batches with more rows (and key slots to match)
        if (!too_big && first_rescue > n_ka_items / 8u) {
          uint32_t ns = ctx->blk_slots, ncap = ctx->blk_cap + ctx->blk_cap / 2;
          while (ncap > ns * 13u / 16u) ns *= 2;
          if (ns <= 8192u && blkA_smem_bytes(ns, ncap, ctx->blk_cslots, ctx->Ws) <= kBlkMaxSmem) {
            ctx->blk_slots = ns;
            ctx->blk_cap = ncap;
          }
        }
      }
      CU(cudaEventRecord(ctx->ev[EV_SORT], st));
      CU(cudaEventRecord(ctx->ev[EV_MARK], st));
      if (too_big) {
        // a position block holds more distinct k-mers than the largest table: shorter blocks,
        // then the record path (for this batch; for good after the second time)
        if (ctx->block_windows > (uint32_t)kBlkRun) {
          ctx->block_windows /= 2;
          TRY(plan_blocks(ctx, *ctx, st, true));
        } else {
          blk = false;
          if (++ctx->block_fallbacks >= 2) ctx->block_mode = false;
          TRY(ensure_records());
        }
        continue;
      }
      if (hcnt[C_LOCAL + LC_PARTIAL_OVERFLOW]) {
        ctx->partial_cap = (uint64_t)hcnt[C_LOCAL + LC_PARTIALS] + 4096;
        continue;
      }
      ro.cluster = ctx->d_row_cluster.as<uint32_t>();
      ro.count = ctx->d_row_count.as<uint32_t>();
      ro.cand = ctx->d_cand.as<uint32_t>();
      ro.kmer = ctx->d_row_kmer.as<uint64_t>();
      ro.row_base = 0;
      TRY(launch_block_merge(ctx, ro, hcnt[C_LOCAL + LC_PARTIALS]));
      STAGE("kB_merge");
      mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
      CU(cudaStreamSynchronize(st));
      TRY(check_device_error(ctx));
      ctx->rescued_last = 0;
      ctx->cp.n = ctx->cp_base + hcnt[C_NEW_CP];
      N.n_runs = 0;
      Wd.n_runs = Wd.n_records ? hcnt[C_RUNS_W] : 0;
      if (hcnt[C_LOCAL + LC_PARTIAL_OVERFLOW]) {
        ctx->partial_cap = (uint64_t)hcnt[C_LOCAL + LC_PARTIALS] + 4096;
        continue;
      }
      if (hcnt[C_LOCAL + LC_ROW_OVERFLOW]) {
        ctx->row_cap = (uint64_t)hcnt[C_LOCAL + LC_ROWS] + 1024;
        TRY(ensure_rows(ctx, ctx->row_cap, ctx->row_cap, false));
        continue;
      }
      ctx->used_block = true;
      ctx->partials_last = hcnt[C_LOCAL + LC_PARTIALS];
      if (n_windows) ctx->partial_ratio = std::max(1e-4, (double)ctx->partials_last / (double)n_windows);
      break;
    }
    // ---- K1 + K2 -------------------------------------------------------------
    ctx->fused = part && ctx->use_direct && N.passes <= 2 && !(ctx->prm.debug_flags & 1u);
    TRY(launch_k1(ctx));
    STAGE("k1_extract");
    CU(cudaEventRecord(ctx->ev[EV_EXTRACT], st));
    TRY(hist_width<uint64_t>(ctx, N, ctx->fused));
    TRY(hist_width<Key128>(ctx, Wd));
    STAGE("k2_histogram");
    CU(cudaEventRecord(ctx->ev[EV_HIST], st));
    TRY(passes_width<uint64_t>(ctx, N, C_TICKET_N, part && ctx->use_direct, ctx->fused));
    TRY(passes_width<Key128>(ctx, Wd, C_TICKET_W));
    STAGE("k2_onesweep_pass");
    CU(cudaEventRecord(ctx->ev[EV_SORT], st));
    // ---- K3: runs -----------------------------------------------------------------
    ctx->runs_from_hist = part && N.passes == 1;
    if (ctx->runs_from_hist) TRY(tiles_from_hist(ctx));
    else TRY(mark_width<uint64_t>(ctx, N, C_TICKET_MARK_N, C_RUNS_N, part));
    STAGE("k3_mark_runs/k3_tiles_from_hist");
    TRY(mark_width<Key128>(ctx, Wd, C_TICKET_MARK_W, C_RUNS_W, false));
    STAGE("k3_mark_runs<wide>");
    CU(cudaEventRecord(ctx->ev[EV_MARK], st));
    // local reduce (+ rescue launch of the direct variant), then one sync to read the counters
    auto run_local = [&]() -> int {
      if (part) {
        ro.cluster = ctx->d_row_cluster.as<uint32_t>();
        ro.count = ctx->d_row_count.as<uint32_t>();
        ro.cand = ctx->d_cand.as<uint32_t>();
        ro.kmer = ctx->d_row_kmer.as<uint64_t>();
        ro.row_base = 0;
        TRY(launch_local(ctx, ro));
        STAGE("k3_local");
      }
      mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
      CU(cudaStreamSynchronize(st));
      TRY(check_device_error(ctx));
      ctx->rescued_last = 0;
      if (part && ctx->use_direct && N.n_records && hcnt[C_LOCAL + LC_RESCUE] &&
          !hcnt[C_LOCAL + LC_TABLE_OVERFLOW]) {
        ctx->rescued_last = hcnt[C_LOCAL + LC_RESCUE];
        TRY(launch_local(ctx, ro, hcnt[C_LOCAL + LC_RESCUE]));
        mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
        CU(cudaStreamSynchronize(st));
      }
      return PF_OK;
    };
    TRY(run_local());
    ctx->cp.n = ctx->cp_base + hcnt[C_NEW_CP];
    N.n_runs = N.n_records ? hcnt[C_RUNS_N] : 0;
    Wd.n_runs = Wd.n_records ? hcnt[C_RUNS_W] : 0;
    if (!part || N.n_records == 0) break;
    if (attempt > 10) return fail(ctx, PF_ERR_INTERNAL, "partition mode did not converge");
    if (hcnt[C_LOCAL + LC_TABLE_OVERFLOW]) {
      // one prefix-run alone holds more distinct k-mers than a CTA can hold: sort 8 more
      // bits (runs get 256x smaller) and redo the batch.  At 64 bits the general variant
      // always fits (<= 2048 + 1 keys per tile); the direct one falls back to it.
      if (N.sort_bits < 64) {
        ctx->extra_bits += 8;
        N.sort_bits = auto_sort_bits(ctx, N.max_seg, true);
        N.passes = N.sort_bits / 8;
        TRY(dev_ensure(ctx, N.seg_hist, std::max<size_t>(1, (size_t)ctx->n_clusters * N.passes * kRadix) * 4));
      } else if (ctx->use_direct) {
        ctx->use_direct = false;
        TRY(plan_local_tiles(ctx, *ctx, st));
        CU(cudaStreamSynchronize(st));
      } else {
        return fail(ctx, PF_ERR_INTERNAL, "local table overflow at 64 sorted bits");
      }
      continue;
    }
    if (hcnt[C_LOCAL + LC_ROW_OVERFLOW]) {
      ctx->row_cap = (uint64_t)hcnt[C_LOCAL + LC_ROWS] + 1024;
      TRY(ensure_rows(ctx, ctx->row_cap, ctx->row_cap, false));
      TRY(run_local());
      if (hcnt[C_LOCAL + LC_ROW_OVERFLOW] || hcnt[C_LOCAL + LC_TABLE_OVERFLOW])
        return fail(ctx, PF_ERR_INTERNAL, "local reduce overflowed twice");
    }
    break;
  }

  if (part) {
    N.n_rows = N.n_records ? hcnt[C_LOCAL + LC_ROWS] : 0;
    ctx->unique_last = N.n_records ? hcnt[C_LOCAL + LC_UNIQUE] : 0;
    if (ctx->used_block && ctx->n_slices == 1)   // distinct k-mers = partial rows - rows folded into an earlier one
      ctx->unique_last = (uint64_t)hcnt[C_LOCAL + LC_PARTIALS] - hcnt[C_LOCAL + LC_RESCUE];
    if (N.n_records) ctx->row_ratio = std::max(1e-4, (double)N.n_rows / (double)N.n_records);
  } else {
    TRY((runs_width<uint64_t, false>(ctx, N, ro)));
  }
  TRY((runs_width<Key128, false>(ctx, Wd, ro)));
  CU(cudaEventRecord(ctx->ev[EV_COUNTED], st));
  const bool need_scan = (!part && N.n_runs) || Wd.n_runs;
  if (need_scan) {
    if (!part && N.n_runs)
      TRY(scan_inplace(ctx, N.vals[N.final_buf ^ 1].as<uint32_t>(), N.n_runs, counters + C_ROWS_N));
    if (Wd.n_runs) TRY(scan_inplace(ctx, Wd.vals[Wd.final_buf ^ 1].as<uint32_t>(), Wd.n_runs, counters + C_ROWS_W));
    mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
    CU(cudaStreamSynchronize(st));
    if (!part) N.n_rows = N.n_runs ? hcnt[C_ROWS_N] : 0;
    Wd.n_rows = Wd.n_runs ? hcnt[C_ROWS_W] : 0;
  } else {
    if (!part) N.n_rows = 0;
    Wd.n_rows = 0;
  }
  if (!part) ctx->unique_last = N.n_runs;
  ctx->unique_last += Wd.n_runs;
  const uint64_t rows = (uint64_t)N.n_rows + Wd.n_rows;
  if (rows >= (1ull << 31)) return fail(ctx, PF_ERR_INVALID, "batch yields 2^31 rows or more; split it");

  if (!part || Wd.n_rows) TRY(ensure_rows(ctx, std::max<uint64_t>(rows, part ? ctx->row_cap : 0),
                                          std::max<uint64_t>(N.n_rows, part ? ctx->row_cap : 0), part));
  TRY(dev_ensure(ctx, ctx->d_wrow_kmer, std::max<size_t>(1, Wd.n_rows) * 16));
  ro.cluster = ctx->d_row_cluster.as<uint32_t>();
  ro.count = ctx->d_row_count.as<uint32_t>();
  ro.cand = ctx->d_cand.as<uint32_t>();
  if (!part && N.n_rows) {
    ro.kmer = ctx->d_row_kmer.as<uint64_t>();
    ro.row_base = 0;
    TRY((runs_width<uint64_t, true>(ctx, N, ro)));
  }
  if (Wd.n_rows) {
    ro.kmer = ctx->d_wrow_kmer.as<uint64_t>();
    ro.row_base = N.n_rows;
    TRY((runs_width<Key128, true>(ctx, Wd, ro)));
  }
  CU(cudaEventRecord(ctx->ev[EV_REDUCE], st));

  // the (cluster, k-mer, count) arrays of the rows are final: start their D2H on the copy
  // stream while K4 numbers the patterns
  ctx->rows_prefetched = false;
  if (rows && !ctx->pipe_mode && ctx->prefetch_rows) {
    TRY(pin_ensure(ctx, ctx->r_row_cluster, rows * 4));
    TRY(pin_ensure(ctx, ctx->r_row_count, rows * 4));
    TRY(pin_ensure(ctx, ctx->r_row_kmer, std::max<size_t>(8, (size_t)N.n_rows * 8)));
    TRY(pin_ensure(ctx, ctx->r_wrow_kmer, std::max<size_t>(8, (size_t)Wd.n_rows * 16)));
    CU(cudaEventRecord(ctx->ev_rows, st));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rows, 0));
    CU(cudaMemcpyAsync(ctx->r_row_cluster.p, ctx->d_row_cluster.p, rows * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->r_row_count.p, ctx->d_row_count.p, rows * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (N.n_rows) CU(cudaMemcpyAsync(ctx->r_row_kmer.p, ctx->d_row_kmer.p, (size_t)N.n_rows * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (Wd.n_rows) CU(cudaMemcpyAsync(ctx->r_wrow_kmer.p, ctx->d_wrow_kmer.p, (size_t)Wd.n_rows * 16, cudaMemcpyDeviceToHost, ctx->copy_stream));
    ctx->rows_prefetched = true;
  }

  // ---- K4 ---------------------------------------------------------------
  // the previous batch's K4 is long over by now (this call has synchronised the stream at least
  // once since): fold its pattern count in, then number this batch's patterns behind it
  TRY(finalize_pending(ctx));
  ctx->kp_base = ctx->kp.n;
  TRY(dedup(ctx, ctx->kp, ctx->d_cand.as<uint32_t>(), (uint32_t)rows, ctx->d_rep, ctx->d_slot_of,
            ctx->d_winner, ctx->d_row_pattern.as<uint32_t>(), C_NEW_KP));
  CU(cudaEventRecord(ctx->ev[EV_DEDUP], st));
  TRY(pin_ensure(ctx, ctx->h_done, 16));
  mirror_counters<<<1, 32, 0, st>>>(hcnt, counters, C_COUNT);
  mirror_counters<<<1, 32, 0, st>>>(ctx->h_done.as<uint32_t>(), counters + C_NEW_KP, 1);
  CU(cudaEventRecord(ctx->ev[EV_END], st));
  ctx->kp_pending = true;
  ctx->kp_pending_base = ctx->kp_base;
  ctx->kp_pending_count = ctx->h_done.as<uint32_t>();
  ctx->executed = true;
  ctx->stats.launches = ctx->launches - launches0;
  return PF_OK;
}

namespace {

// grow a pinned result array, keeping the `used` bytes already copied into it
int pin_grow(pf_ctx* ctx, PinBuf& b, size_t need, size_t used) {
  if (need <= b.cap) return PF_OK;
  CU(cudaStreamSynchronize(ctx->copy_stream));     // copies into the old array are in flight
  size_t want = std::max(need, b.cap + b.cap / 2);
  want = (want + 4095) & ~size_t(4095);
  void* np = nullptr;
  CU(cudaMallocHost(&np, want));
  if (b.p && used) memcpy(np, b.p, used);
  if (b.p) CU(cudaFreeHost(b.p));
  b.p = np;
  b.cap = want;
  return PF_OK;
}

// cut a batch into sub-batches of whole clusters with about `target` sequences each
std::vector<SubRange> split_batch(const pf_batch* b, uint32_t target) {
  std::vector<SubRange> subs;
  const uint32_t n = b->n_seqs;
  uint32_t s0 = 0, c0 = 0;
  while (s0 < n) {
    uint32_t s1 = n, c1 = b->n_clusters;
    if ((uint64_t)s0 + target + target / 2 < n) {
      // first sequence of the cluster that holds sequence s0 + target (clusters are sorted)
      const uint32_t c = b->seqs[s0 + target].cluster;
      uint32_t lo = s0, hi = s0 + target;
      while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (b->seqs[mid].cluster < c) lo = mid + 1; else hi = mid;
      }
      if (lo > s0) { s1 = lo; c1 = c; }
      else {            // one cluster longer than the target: take it whole
        lo = s0 + target; hi = n;
        while (lo < hi) {
          const uint32_t mid = lo + (hi - lo) / 2;
          if (b->seqs[mid].cluster <= c) lo = mid + 1; else hi = mid;
        }
        s1 = lo;
        c1 = s1 < n ? b->seqs[s1].cluster : b->n_clusters;
      }
    }
    SubRange r{};
    r.s0 = s0; r.s1 = s1; r.c0 = c0; r.c1 = std::max(c1, c0);
    r.w0 = b->seqs[s0].base_off >> 5;
    r.w1 = s1 < n ? (b->seqs[s1].base_off >> 5) : b->n_words;
    r.a0 = 0; r.a1 = 0;
    subs.push_back(r);
    s0 = s1;
    c0 = r.c1;
  }
  if (!subs.empty()) subs.back().c1 = b->n_clusters;     // trailing clusters without sequences
  return subs;
}

void pipe_add_timings(pf_ctx* ctx, const BatchState* slot = nullptr) {
  fill_timings(ctx, slot);
  const pf_stats& t = ctx->stats;
  const float v[10] = {t.ms_h2d, t.ms_extract, t.ms_hist, t.ms_sort, t.ms_mark, t.ms_count, t.ms_reduce,
                       t.ms_dedup, 0.f, 0.f};
  for (int i = 0; i < 10; ++i) ctx->pipe_ms[i] += v[i];
}

// D2H of the current slot's results behind everything it executed, appended to the pinned
// result arrays of the pipelined submit
int pipe_enqueue_results(pf_ctx* ctx, const SubRange& r, int slot) {
  cudaStream_t st = ctx->stream, cp = ctx->copy_stream;
  WidthState& N = ctx->nar;
  WidthState& Wd = ctx->wid;
  const uint64_t nr = N.n_rows, nw = Wd.n_rows;
  const uint64_t R = ctx->pipe_rows, RW = ctx->pipe_wide_rows;
  TRY(pin_grow(ctx, ctx->r_row_cluster, (R + nr) * 4, R * 4));
  TRY(pin_grow(ctx, ctx->r_row_count, (R + nr) * 4, R * 4));
  TRY(pin_grow(ctx, ctx->r_row_pattern, (R + nr) * 4, R * 4));
  TRY(pin_grow(ctx, ctx->r_row_kmer, (R + nr) * 8, R * 8));
  TRY(pin_grow(ctx, ctx->r_wrow_cluster, std::max<uint64_t>(8, (RW + nw) * 4), RW * 4));
  TRY(pin_grow(ctx, ctx->r_wrow_count, std::max<uint64_t>(8, (RW + nw) * 4), RW * 4));
  TRY(pin_grow(ctx, ctx->r_wrow_pattern, std::max<uint64_t>(8, (RW + nw) * 4), RW * 4));
  TRY(pin_grow(ctx, ctx->r_wrow_kmer, std::max<uint64_t>(8, (RW + nw) * 16), RW * 16));
  if (ctx->n_pos) {
    const uint64_t P0 = ctx->pipe_pos, np = ctx->n_pos;
    TRY(pin_grow(ctx, ctx->r_pos_kmer, (P0 + np) * 8, P0 * 8));
    TRY(pin_grow(ctx, ctx->r_pos_seq, (P0 + np) * 4, P0 * 4));
    TRY(pin_grow(ctx, ctx->r_pos_cstart, (P0 + np) * 4, P0 * 4));
    TRY(pin_grow(ctx, ctx->r_pos_gstart, (P0 + np) * 4, P0 * 4));
    TRY(pin_grow(ctx, ctx->r_pos_flags, (P0 + np), P0));
    if (ctx->n_pos_wide)
      TRY(pin_grow(ctx, ctx->r_pos_wide, (ctx->pipe_pos_wide + ctx->n_pos_wide) * 16, ctx->pipe_pos_wide * 16));
    pos_rebase<<<cdiv(ctx->n_pos, 256), 256, 0, st>>>(ctx->d_pos_seq.as<uint32_t>(), ctx->d_pos_kmer.as<uint64_t>(),
                                                      ctx->d_pos_flags.as<uint8_t>(), ctx->n_pos, r.s0,
                                                      ctx->pipe_pos_wide);
    ctx->launches++;
  }
  CU(cudaEventRecord(ctx->ev_rows, st));
  CU(cudaStreamWaitEvent(cp, ctx->ev_rows, 0));
  auto d2h = [&](PinBuf& dst, size_t dst_off, const void* src, size_t bytes) -> int {
    if (bytes) CU(cudaMemcpyAsync((char*)dst.p + dst_off, src, bytes, cudaMemcpyDeviceToHost, cp));
    return PF_OK;
  };
  TRY(d2h(ctx->r_row_cluster, R * 4, ctx->d_row_cluster.p, nr * 4));
  TRY(d2h(ctx->r_row_count, R * 4, ctx->d_row_count.p, nr * 4));
  TRY(d2h(ctx->r_row_pattern, R * 4, ctx->d_row_pattern.p, nr * 4));
  TRY(d2h(ctx->r_row_kmer, R * 8, ctx->d_row_kmer.p, nr * 8));
  TRY(d2h(ctx->r_wrow_cluster, RW * 4, ctx->d_row_cluster.as<uint32_t>() + nr, nw * 4));
  TRY(d2h(ctx->r_wrow_count, RW * 4, ctx->d_row_count.as<uint32_t>() + nr, nw * 4));
  TRY(d2h(ctx->r_wrow_pattern, RW * 4, ctx->d_row_pattern.as<uint32_t>() + nr, nw * 4));
  TRY(d2h(ctx->r_wrow_kmer, RW * 16, ctx->d_wrow_kmer.p, nw * 16));
  TRY(d2h(ctx->r_cl_pattern, (size_t)ctx->pipe_clusters * 4, ctx->d_cl_pattern.p, (size_t)ctx->n_clusters * 4));
  if (ctx->n_pos) {
    const uint64_t P0 = ctx->pipe_pos, np = ctx->n_pos;
    TRY(d2h(ctx->r_pos_kmer, P0 * 8, ctx->d_pos_kmer.p, np * 8));
    TRY(d2h(ctx->r_pos_seq, P0 * 4, ctx->d_pos_seq.p, np * 4));
    TRY(d2h(ctx->r_pos_cstart, P0 * 4, ctx->d_pos_cstart.p, np * 4));
    TRY(d2h(ctx->r_pos_gstart, P0 * 4, ctx->d_pos_gstart.p, np * 4));
    TRY(d2h(ctx->r_pos_flags, P0, ctx->d_pos_flags.p, np));
    if (ctx->n_pos_wide)
      TRY(d2h(ctx->r_pos_wide, ctx->pipe_pos_wide * 16, ctx->d_pos_wide.p, (size_t)ctx->n_pos_wide * 16));
  }
  CU(cudaEventRecord(ctx->ev_out_done[slot], cp));
  ctx->pipe_rows += nr;
  ctx->pipe_wide_rows += nw;
  ctx->pipe_pos += ctx->n_pos;
  ctx->pipe_pos_wide += ctx->n_pos_wide;
  ctx->pipe_clusters += ctx->n_clusters;
  pf_stats& s = ctx->stats;
  s.bases += ctx->n_bases;
  s.instances += (uint64_t)N.n_records + Wd.n_records;
  s.unique_kmers += ctx->unique_last;
  s.rows += nr + nw;
  return PF_OK;
}

// D2H (copy stream) of the k-mer patterns numbered since the last call; the compute stream
// must have passed the K4 that appended them
int pipe_copy_new_patterns(pf_ctx* ctx) {
  const uint64_t done = ctx->kp.n, from = ctx->pipe_kp_copied, base = ctx->pipe_kp_base;
  if (done <= from) return PF_OK;
  const size_t row = (size_t)ctx->Wk * 4;
  TRY(pin_grow(ctx, ctx->r_new_kp, (done - base) * row, (from - base) * row));
  CU(cudaMemcpyAsync((char*)ctx->r_new_kp.p + (from - base) * row, ctx->kp.pool.as<uint32_t>() + from * ctx->Wk,
                     (done - from) * row, cudaMemcpyDeviceToHost, ctx->copy_stream));
  ctx->pipe_kp_copied = done;
  return PF_OK;
}

// pf_submit of a large batch: sub-batches of whole clusters flow through two batch slots, so
// that the H2D of sub-batch j+1 (up_stream) and the D2H of sub-batch j-1 (copy_stream) run
// under the kernels of sub-batch j (stream).  The pattern tables are shared, K4 of the
// sub-batches is ordered by the compute stream, so pattern ids are those of one big batch.
int submit_pipelined(pf_ctx* ctx, const pf_batch* b, const std::vector<SubRange>& subs) {
  cudaStream_t st = ctx->stream, up = ctx->up_stream;
  TRY(finalize_pending(ctx));
  ctx->executed = false;
  ctx->alt.executed = false;
  ctx->pipe_rows = ctx->pipe_wide_rows = ctx->pipe_pos = ctx->pipe_pos_wide = 0;
  ctx->pipe_clusters = 0;
  ctx->pipe_kp_base = ctx->kp.n;
  ctx->pipe_kp_copied = ctx->kp.n;
  ctx->pipe_cp_base = ctx->cp.n;
  for (double& m : ctx->pipe_ms) m = 0;
  TRY(pin_ensure(ctx, ctx->r_cl_pattern, std::max<size_t>(8, (size_t)b->n_clusters * 4)));
  {
    // row arrays: learned rows-per-base ratio, grown on demand
    const uint64_t est = (uint64_t)(ctx->row_ratio * 1.3 * (double)b->n_words * 32.0) + 65536;
    TRY(pin_grow(ctx, ctx->r_row_cluster, est * 4, 0));
    TRY(pin_grow(ctx, ctx->r_row_count, est * 4, 0));
    TRY(pin_grow(ctx, ctx->r_row_pattern, est * 4, 0));
    TRY(pin_grow(ctx, ctx->r_row_kmer, est * 8, 0));
  }
  auto swap_slots = [&]() { std::swap(static_cast<BatchState&>(*ctx), ctx->alt); };
  struct Guard { pf_ctx* c; ~Guard() { c->pipe_mode = false; } } guard{ctx};
  ctx->pipe_mode = true;
  int cur = 0;
  const size_t J = subs.size();
  ctx->pipe_subs = (uint32_t)J;
  static const bool dbg = getenv("PF_DEBUG_PIPE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  double t_up = 0, t_ex = 0, t_out = 0, t_fin = 0;
  CU(cudaEventRecord(ctx->ev_pipe[0], st));
  TRY(upload_async(ctx, *ctx, b, subs[0], up));
  TRY(upload_finish(ctx, *ctx, up));
  for (size_t j = 0; j < J; ++j) {
    // a helper thread validates, plans and uploads sub-batch j+1 into the other slot while
    // this thread drives the kernels of sub-batch j (pf_execute blocks on its read-backs)
    std::thread helper;
    int up_rc = PF_OK;
    if (j + 1 < J) {
      // the kernels of the other slot's last occupant (sub-batch j-1) must be done with its
      // inputs; its result arrays are still draining, but an upload does not touch those
      CU(cudaStreamWaitEvent(up, ctx->ev_exec_end[cur ^ 1], 0));
      helper = std::thread([&, j]() {
        cudaSetDevice(ctx->device);
        const double t0 = now();
        up_rc = upload_async(ctx, ctx->alt, b, subs[j + 1], up);
        t_up += now() - t0;
      });
    }
    // this slot's previous rows must have left the device before they are overwritten
    int rc = PF_OK;
    if (cudaStreamWaitEvent(st, ctx->ev_out_done[cur], 0) != cudaSuccess) rc = fail(ctx, PF_ERR_CUDA, "cudaStreamWaitEvent failed");
    double t0 = now();
    if (rc == PF_OK) rc = pf_execute(ctx);
    t_ex += now() - t0;
    if (rc == PF_OK && cudaEventRecord(ctx->ev_exec_end[cur], st) != cudaSuccess) rc = fail(ctx, PF_ERR_CUDA, "cudaEventRecord failed");
    t0 = now();
    if (rc == PF_OK) rc = pipe_enqueue_results(ctx, subs[j], cur);
    t_out += now() - t0;
    if (helper.joinable()) helper.join();
    if (rc != PF_OK) return rc;
    if (up_rc != PF_OK) return up_rc;
    t0 = now();
    // pf_execute folded sub-batch j-1's pattern count in before its own K4: those patterns are
    // final, and so are the other slot's stage timestamps
    TRY(pipe_copy_new_patterns(ctx));
    if (j > 0) pipe_add_timings(ctx, &ctx->alt);
    if (j + 1 < J) {
      ctx->executed = false;
      swap_slots();
      cur ^= 1;
      TRY(upload_finish(ctx, *ctx, up));
      t_fin += now() - t0;
    }
  }
  if (dbg)
    fprintf(stderr, "[pf] pipeline: %zu sub-batches, host ms: total %.2f  upload_async %.2f  execute %.2f  enqueue %.2f  finish %.2f\n",
            J, now() - t_begin, t_up, t_ex, t_out, t_fin);
  CU(cudaEventRecord(ctx->ev_pipe[1], st));
  ctx->pipe_pending = true;
  return PF_OK;
}

}  // namespace

extern "C" int pf_submit(pf_ctx* ctx, const pf_batch* batch) {
  if (!ctx) return PF_ERR_INVALID;
  if (!batch) return fail(ctx, PF_ERR_INVALID, "pf_submit: null batch");
  if (ctx->pipe_pending) return fail(ctx, PF_ERR_STATE, "pf_submit: results of the previous pf_submit were not collected");
  if (batch->n_seqs >= ctx->pipe_min_seqs && batch->n_amb_words == 0 && batch->seqs && batch->n_clusters > 1) {
    CU(cudaSetDevice(ctx->device));
    const std::vector<SubRange> subs = split_batch(batch, ctx->pipe_target_seqs);
    if (subs.size() > 1) return submit_pipelined(ctx, batch, subs);
  }
  ctx->pipe_subs = 1;
  int r = pf_upload(ctx, batch);
  if (r != PF_OK) return r;
  return pf_execute(ctx);
}

namespace {
int collect_pipelined(pf_ctx* ctx, pf_batch_result* out) {
  cudaStream_t st = ctx->stream;
  CU(cudaStreamSynchronize(st));
  TRY(check_device_error(ctx));
  uint32_t* hcnt = ctx->h_counters.as<uint32_t>();
  ctx->kp.n = ctx->kp_base + hcnt[C_NEW_KP];      // last sub-batch
  ctx->kp_pending = false;
  ctx->executed = false;
  ctx->alt.executed = false;
  ctx->pipe_pending = false;
  pipe_add_timings(ctx);
  const uint64_t new_kp = ctx->kp.n - ctx->pipe_kp_base, new_cp = ctx->cp.n - ctx->pipe_cp_base;
  CU(cudaEventRecord(ctx->ev_d2h[0], st));
  TRY(pipe_copy_new_patterns(ctx));
  if (!ctx->r_new_kp.p) TRY(pin_ensure(ctx, ctx->r_new_kp, 8));
  TRY(pin_ensure(ctx, ctx->r_new_cp, std::max<size_t>(8, new_cp * ctx->W * 4)));
  if (new_cp)
    CU(cudaMemcpyAsync(ctx->r_new_cp.p, ctx->cp.pool.as<uint32_t>() + ctx->pipe_cp_base * ctx->W, new_cp * ctx->W * 4,
                       cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(ctx->ev_d2h[1], st));
  CU(cudaStreamSynchronize(st));
  CU(cudaStreamSynchronize(ctx->copy_stream));
  if (out) {
    memset(out, 0, sizeof *out);
    out->n_rows = ctx->pipe_rows;
    out->row_cluster = ctx->r_row_cluster.as<uint32_t>();
    out->row_kmer = ctx->r_row_kmer.as<uint64_t>();
    out->row_count = ctx->r_row_count.as<uint32_t>();
    out->row_pattern = ctx->r_row_pattern.as<uint32_t>();
    out->n_wide_rows = ctx->pipe_wide_rows;
    out->wide_row_cluster = ctx->r_wrow_cluster.as<uint32_t>();
    out->wide_row_kmer = ctx->r_wrow_kmer.as<uint64_t>();
    out->wide_row_count = ctx->r_wrow_count.as<uint32_t>();
    out->wide_row_pattern = ctx->r_wrow_pattern.as<uint32_t>();
    out->n_clusters = ctx->pipe_clusters;
    out->cluster_pattern = ctx->r_cl_pattern.as<uint32_t>();
    out->kmer_pattern_base = ctx->pipe_kp_base;
    out->n_new_kmer_patterns = new_kp;
    out->new_kmer_patterns = ctx->r_new_kp.as<uint32_t>();
    out->cluster_pattern_base = ctx->pipe_cp_base;
    out->n_new_cluster_patterns = new_cp;
    out->new_cluster_patterns = ctx->r_new_cp.as<uint32_t>();
    out->n_pos = ctx->pipe_pos;
    out->pos_kmer = ctx->r_pos_kmer.as<uint64_t>();
    out->pos_seq = ctx->r_pos_seq.as<uint32_t>();
    out->pos_contig_start = ctx->r_pos_cstart.as<int32_t>();
    out->pos_gene_start = ctx->r_pos_gstart.as<int32_t>();
    out->pos_flags = ctx->r_pos_flags.as<uint8_t>();
    out->pos_wide_kmer = ctx->r_pos_wide.as<uint64_t>();
    out->n_pos_wide = ctx->pipe_pos_wide;
  }
  pf_stats& s = ctx->stats;
  s.batches++;
  s.kmer_patterns = ctx->kp.n;
  s.cluster_patterns = ctx->cp.n;
  s.total_launches = ctx->launches;
  s.ms_h2d = (float)ctx->pipe_ms[0]; s.ms_extract = (float)ctx->pipe_ms[1]; s.ms_hist = (float)ctx->pipe_ms[2];
  s.ms_sort = (float)ctx->pipe_ms[3]; s.ms_mark = (float)ctx->pipe_ms[4]; s.ms_count = (float)ctx->pipe_ms[5];
  s.ms_reduce = (float)ctx->pipe_ms[6]; s.ms_dedup = (float)ctx->pipe_ms[7];
  float m = 0;
  if (cudaEventElapsedTime(&m, ctx->ev_pipe[0], ctx->ev_pipe[1]) == cudaSuccess) s.ms_total = m;
  if (cudaEventElapsedTime(&m, ctx->ev_d2h[0], ctx->ev_d2h[1]) == cudaSuccess) s.ms_d2h = m;
  return PF_OK;
}
}  // namespace

extern "C" int pf_collect(pf_ctx* ctx, pf_batch_result* out) {
  if (!ctx) return PF_ERR_INVALID;
  if (ctx->pipe_pending) { CU(cudaSetDevice(ctx->device)); return collect_pipelined(ctx, out); }
  if (!ctx->executed) return fail(ctx, PF_ERR_STATE, "pf_collect: nothing executed");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CU(cudaStreamSynchronize(st));
  uint32_t* hcnt = ctx->h_counters.as<uint32_t>();
  TRY(check_device_error(ctx));
  WidthState& N = ctx->nar;
  WidthState& Wd = ctx->wid;
  const uint64_t rows = (uint64_t)N.n_rows + Wd.n_rows;
  const uint64_t new_kp = hcnt[C_NEW_KP];
  const uint64_t new_cp = ctx->cp.n - ctx->cp_base;
  ctx->kp.n = ctx->kp_base + new_kp;
  ctx->kp_pending = false;
  ctx->executed = false;      // results are handed out once

  CU(cudaEventRecord(ctx->ev_d2h[0], st));
  auto d2h = [&](PinBuf& dst, const void* src, size_t bytes) -> int {
    TRY(pin_ensure(ctx, dst, std::max<size_t>(bytes, 8)));
    if (bytes) CU(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyDeviceToHost, st));
    return PF_OK;
  };
  if (!ctx->rows_prefetched) {
    TRY(d2h(ctx->r_row_cluster, ctx->d_row_cluster.p, rows * 4));
    TRY(d2h(ctx->r_row_count, ctx->d_row_count.p, rows * 4));
    TRY(d2h(ctx->r_row_kmer, ctx->d_row_kmer.p, (size_t)N.n_rows * 8));
    TRY(d2h(ctx->r_wrow_kmer, ctx->d_wrow_kmer.p, (size_t)Wd.n_rows * 16));
  }
  TRY(d2h(ctx->r_row_pattern, ctx->d_row_pattern.p, rows * 4));
  TRY(d2h(ctx->r_cl_pattern, ctx->d_cl_pattern.p, (size_t)ctx->n_clusters * 4));
  TRY(d2h(ctx->r_new_kp, ctx->kp.pool.as<uint32_t>() + ctx->kp_base * ctx->Wk, new_kp * ctx->Wk * 4));
  TRY(d2h(ctx->r_new_cp, ctx->cp.pool.as<uint32_t>() + ctx->cp_base * ctx->W, new_cp * ctx->W * 4));
  if (ctx->n_pos) {
    TRY(d2h(ctx->r_pos_kmer, ctx->d_pos_kmer.p, (size_t)ctx->n_pos * 8));
    TRY(d2h(ctx->r_pos_seq, ctx->d_pos_seq.p, (size_t)ctx->n_pos * 4));
    TRY(d2h(ctx->r_pos_cstart, ctx->d_pos_cstart.p, (size_t)ctx->n_pos * 4));
    TRY(d2h(ctx->r_pos_gstart, ctx->d_pos_gstart.p, (size_t)ctx->n_pos * 4));
    TRY(d2h(ctx->r_pos_flags, ctx->d_pos_flags.p, (size_t)ctx->n_pos));
  }
  if (ctx->n_pos_wide) TRY(d2h(ctx->r_pos_wide, ctx->d_pos_wide.p, (size_t)ctx->n_pos_wide * 16));
  CU(cudaEventRecord(ctx->ev_d2h[1], st));
  CU(cudaStreamSynchronize(st));
  if (ctx->rows_prefetched) CU(cudaStreamSynchronize(ctx->copy_stream));
  ctx->rows_prefetched = false;

  if (out) {
    memset(out, 0, sizeof *out);
    out->n_rows = N.n_rows;
    out->row_cluster = ctx->r_row_cluster.as<uint32_t>();
    out->row_kmer = ctx->r_row_kmer.as<uint64_t>();
    out->row_count = ctx->r_row_count.as<uint32_t>();
    out->row_pattern = ctx->r_row_pattern.as<uint32_t>();
    out->n_wide_rows = Wd.n_rows;
    out->wide_row_cluster = out->row_cluster + N.n_rows;
    out->wide_row_kmer = ctx->r_wrow_kmer.as<uint64_t>();
    out->wide_row_count = out->row_count + N.n_rows;
    out->wide_row_pattern = out->row_pattern + N.n_rows;
    out->n_clusters = ctx->n_clusters;
    out->cluster_pattern = ctx->r_cl_pattern.as<uint32_t>();
    out->kmer_pattern_base = ctx->kp_base;
    out->n_new_kmer_patterns = new_kp;
    out->new_kmer_patterns = ctx->r_new_kp.as<uint32_t>();
    out->cluster_pattern_base = ctx->cp_base;
    out->n_new_cluster_patterns = new_cp;
    out->new_cluster_patterns = ctx->r_new_cp.as<uint32_t>();
    out->n_pos = ctx->n_pos;
    out->pos_kmer = ctx->r_pos_kmer.as<uint64_t>();
    out->pos_seq = ctx->r_pos_seq.as<uint32_t>();
    out->pos_contig_start = ctx->r_pos_cstart.as<int32_t>();
    out->pos_gene_start = ctx->r_pos_gstart.as<int32_t>();
    out->pos_flags = ctx->r_pos_flags.as<uint8_t>();
    out->pos_wide_kmer = ctx->r_pos_wide.as<uint64_t>();
    out->n_pos_wide = ctx->n_pos_wide;
  }
  // ---- stats ---------------------------------------------------------------
  pf_stats& s = ctx->stats;
  s.batches++;
  s.bases += ctx->n_bases;
  s.instances += (uint64_t)N.n_records + Wd.n_records;
  s.unique_kmers += ctx->unique_last;
  s.rows += rows;
  s.kmer_patterns = ctx->kp.n;
  s.cluster_patterns = ctx->cp.n;
  s.sort_passes = (uint32_t)N.passes;
  s.total_launches = ctx->launches;
  fill_timings(ctx);
  {
    float m = 0;
    if (cudaEventElapsedTime(&m, ctx->ev_d2h[0], ctx->ev_d2h[1]) == cudaSuccess) s.ms_d2h = m;
  }
  return PF_OK;
}

extern "C" int pf_reset_patterns(pf_ctx* ctx) {
  if (!ctx) return PF_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->executed = false;
  ctx->kp_pending = false;
  for (PatternSpace* s : {&ctx->kp, &ctx->cp}) {
    s->n = 0;
    s->x_n_unique = 0;
    if (s->table.p) CU(cudaMemsetAsync(s->table.p, 0xff, (size_t)s->table_size * 4, ctx->stream));
  }
  ctx->kp_base = ctx->cp_base = 0;
  ctx->stats.kmer_patterns = ctx->stats.cluster_patterns = 0;
  CU(cudaStreamSynchronize(ctx->stream));
  return PF_OK;
}

extern "C" int pf_patterns_export(pf_ctx* ctx, int cluster_namespace, uint64_t first, uint64_t count,
                                  uint32_t* host_out) {
  if (!ctx || !host_out) return PF_ERR_INVALID;
  TRY(finalize_pending(ctx));
  PatternSpace& s = cluster_namespace ? ctx->cp : ctx->kp;
  if (first + count > s.n) return fail(ctx, PF_ERR_INVALID, "pattern range [%llu,%llu) beyond %llu",
                                       (unsigned long long)first, (unsigned long long)(first + count),
                                       (unsigned long long)s.n);
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  if (count)
    CU(cudaMemcpy(host_out, s.pool.as<uint32_t>() + first * s.key_words, count * s.key_words * 4,
                  cudaMemcpyDeviceToHost));
  return PF_OK;
}

extern "C" int pf_pattern_ids(pf_ctx* ctx, int cluster_namespace, uint64_t first, uint64_t count,
                              uint8_t* host_digests) {
  if (!ctx || (count && !host_digests)) return PF_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  TRY(finalize_pending(ctx));
  PatternSpace& s = cluster_namespace ? ctx->cp : ctx->kp;
  if (first + count > s.n) return fail(ctx, PF_ERR_INVALID, "pattern range [%llu,%llu) beyond %llu",
                                       (unsigned long long)first, (unsigned long long)(first + count),
                                       (unsigned long long)s.n);
  if (count == 0) return PF_OK;
  TRY(dev_ensure(ctx, ctx->d_digests, count * 16));
  k5_md5_ids<<<cdiv(count, 128), 128, 0, ctx->stream>>>(
      s.pool.as<uint32_t>(), (uint32_t)first, (uint32_t)count, s.key_words, ctx->W, ctx->prm.n_samples,
      cluster_namespace ? 1 : 0, ctx->cp.pool.as<uint32_t>(), ctx->d_digests.as<uint8_t>());
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(host_digests, ctx->d_digests.p, count * 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return PF_OK;
}

extern "C" int pf_stats_get(pf_ctx* ctx, pf_stats* out) {
  if (!ctx || !out) return PF_ERR_INVALID;
  if (ctx->executed) {
    TRY(finalize_pending(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    fill_timings(ctx);
  }
  ctx->stats.kmer_patterns = ctx->kp.n;
  ctx->stats.cluster_patterns = ctx->cp.n;
  *out = ctx->stats;
  out->total_launches = ctx->launches;
  return PF_OK;
}

// ---------------------------------------------------------------------------
// synthetic pangenome
// ---------------------------------------------------------------------------
namespace {
struct SynthCell { bool present; uint32_t copies; };
inline uint64_t thr64(double p) {
  if (p <= 0) return 0;
  if (p >= 1) return ~0ull;
  return (uint64_t)(p * 18446744073709551616.0);
}
inline SynthCell synth_cell(const pf_synth_params* p, uint32_t gc, uint32_t s) {
  // core / accessory is a per-cluster coin flip (core_fraction), so any shard of the
  // pangenome has the same mix
  double pc = 0.99;
  if (synth_hash(p->seed, gc, 0, 0, kTagCore) >= thr64(p->core_fraction)) {
    const uint64_t h = synth_hash(p->seed, gc, 0, 0, kTagAccessoryP);
    pc = 0.05 + 0.90 * ((double)(h >> 11) / 9007199254740992.0);
  }
  SynthCell c;
  c.present = synth_hash(p->seed, gc, s, 0, kTagPresence) < thr64(pc);
  c.copies = c.present ? (synth_hash(p->seed, gc, s, 0, kTagParalog) < thr64(p->paralog_rate) ? 2u : 1u) : 0u;
  return c;
}
}  // namespace

extern "C" int pf_synth_plan(const pf_synth_params* p, uint32_t* n_seqs, uint64_t* n_words) {
  if (!p || !n_seqs || !n_words || p->gene_len == 0 || p->n_founders == 0) return PF_ERR_INVALID;
  const uint64_t wps = ((uint64_t)p->gene_len + 63) / 64 * 2;   // words per sequence, 64-base aligned
  uint64_t n = 0;
  for (uint32_t c = 0; c < p->n_clusters; ++c)
    for (uint32_t s = 0; s < p->n_samples; ++s) n += synth_cell(p, p->first_cluster + c, s).copies;
  if (n >= (1ull << 32)) return PF_ERR_INVALID;
  *n_seqs = (uint32_t)n;
  *n_words = n * wps;
  return PF_OK;
}

extern "C" int pf_synth_fill(int device, const pf_synth_params* p, pf_seq_desc* seqs,
                             pf_cluster_desc* clusters, uint32_t* presence, uint64_t* packed_bases) {
  pf_ctx* ctx = nullptr;
  if (!p || !seqs || !clusters || !presence || !packed_bases) return fail(nullptr, PF_ERR_INVALID, "pf_synth_fill: null argument");
  const uint32_t W = pf_pattern_words(p->n_samples);
  const uint64_t wps = ((uint64_t)p->gene_len + 63) / 64 * 2;
  std::vector<SynthSeq> ss;
  memset(presence, 0, (size_t)p->n_clusters * W * 4);
  uint64_t n = 0;
  for (uint32_t c = 0; c < p->n_clusters; ++c) {
    const uint32_t gc = p->first_cluster + c;
    clusters[c].id = gc;
    clusters[c].reserved = 0;
    for (uint32_t s = 0; s < p->n_samples; ++s) {
      const SynthCell cell = synth_cell(p, gc, s);
      if (!cell.present) continue;
      presence[(size_t)c * W + (s >> 5)] |= 1u << (s & 31);
      for (uint32_t cp = 0; cp < cell.copies; ++cp) {
        pf_seq_desc& q = seqs[n];
        q.base_off = n * wps * 32;
        q.len = p->gene_len;
        q.cluster = c;
        q.sample = s;
        q.flags = p->all_targets ? PF_SEQ_TARGET : 0u;
        const uint64_t inst = ((uint64_t)s << 8) | cp;
        q.strand = (synth_hash(p->seed, gc, inst, 0, kTagStrand) & 1u) ? 1 : -1;
        q.start = 1 + (int32_t)(synth_hash(p->seed, gc, inst, 0, kTagStart) % 1000000u);
        q.end = q.start + (int32_t)p->gene_len - 1;
        q.offset = 100;
        q.amb_off = 0;
        ss.push_back(SynthSeq{n * wps, gc, s, cp, p->gene_len});
        ++n;
      }
    }
  }
  if (n == 0) return PF_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, PF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  SynthSeq* d_ss = nullptr;
  uint64_t* d_out = nullptr;
  CU(cudaMalloc(&d_ss, ss.size() * sizeof(SynthSeq)));
  CU(cudaMalloc(&d_out, n * wps * 8));
  CU(cudaMemcpy(d_ss, ss.data(), ss.size() * sizeof(SynthSeq), cudaMemcpyHostToDevice));
  const uint64_t threads = n * wps;
  synth_bases<<<cdiv(threads, 256), 256>>>(d_ss, (uint32_t)n, (uint32_t)wps, p->seed, p->n_founders,
                                           thr64(p->founder_div), thr64(p->private_div), d_out);
  CU(cudaGetLastError());
  CU(cudaMemcpy(packed_bases, d_out, n * wps * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_ss);
  cudaFree(d_out);
  return PF_OK;
}

// ---------------------------------------------------------------------------
// multi-GPU exchange
// ---------------------------------------------------------------------------
namespace pf {

// owner of every local pattern + its position inside the owner's bucket.  L lanes hash one
// key (as in K4: short keys would leave most of a warp idle), a warp handles 32 consecutive
// patterns and hands out their bucket positions with ONE atomic per distinct owner — a counter
// per rank shared by millions of patterns would serialise in L2.
template <int L>
__global__ void __launch_bounds__(256)
x_classify(const uint32_t* __restrict__ pool, uint32_t n, uint32_t key_words,
           const uint32_t* __restrict__ mask_remap, uint32_t world, uint32_t* __restrict__ owner,
           uint32_t* __restrict__ pos, uint32_t* __restrict__ counts) {
  const uint32_t lane = lane_id(), gl = lane & (L - 1), g = lane / L;
  constexpr uint32_t G = 32 / L;                   // patterns hashed at once by a warp
  const uint32_t total_warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t e0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; e0 < n; e0 += total_warps * 32u) {
    uint32_t my_owner = 0xffffffffu;
    const uint32_t cnt = min(32u, n - e0);
    for (uint32_t p0 = 0; p0 < 32u; p0 += G) {     // (uniform trip count: the shuffles need all lanes)
      const uint32_t p = p0 + g;
      uint64_t h = 0;
      if (p < cnt) {
        const uint32_t* key = pool + (size_t)(e0 + p) * key_words;
        for (uint32_t w = gl; w < key_words; w += L) {
          uint32_t v = key[w];
          if (mask_remap && w == key_words - 1) v = mask_remap[v];
          h += word_hash(v, w);
        }
      }
#pragma unroll
      for (int m = L / 2; m >= 1; m >>= 1) h += __shfl_xor_sync(kFull, h, m);
      h = fmix64(h);
      const uint32_t o = (uint32_t)((h >> 32) % world);
      // lane p of the warp keeps pattern p's owner: it sits in group p - p0, any lane of it
      const uint32_t got = __shfl_sync(kFull, o, (lane - p0) * L);
      if (lane >= p0 && lane < p0 + G && lane < cnt) my_owner = got;
    }
    const uint32_t m = __match_any_sync(kFull, my_owner);
    if (lane < cnt) {
      const int leader = __ffs(m) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(&counts[my_owner], (uint32_t)__popc(m));
      base = __shfl_sync(m, base, leader);
      owner[e0 + lane] = my_owner;
      pos[e0 + lane] = base + __popc(m & lanemask_lt());
    }
  }
}

// exclusive offsets of the `world` buckets behind the counts (counts[world .. 2 world))
__global__ void x_offsets(uint32_t* __restrict__ counts, uint32_t world) {
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t r = 0; r < world; ++r) { counts[world + r] = run; run += counts[r]; }
  }
}

template <int L>
__global__ void __launch_bounds__(256)
x_pack(const uint32_t* __restrict__ pool, uint32_t n, uint32_t key_words,
       const uint32_t* __restrict__ mask_remap, const uint32_t* __restrict__ owner,
       const uint32_t* __restrict__ pos, const uint32_t* __restrict__ offsets,
       uint32_t* __restrict__ send, uint32_t* __restrict__ perm) {
  const uint32_t gl = lane_id() & (L - 1);
  const uint32_t total = gridDim.x * (blockDim.x / L);
  for (uint32_t e = blockIdx.x * (blockDim.x / L) + threadIdx.x / L; e < n; e += total) {
    const uint32_t dst = offsets[owner[e]] + pos[e];
    const uint32_t* key = pool + (size_t)e * key_words;
    uint32_t* out = send + (size_t)dst * key_words;
    for (uint32_t w = gl; w < key_words; w += L) {
      uint32_t v = key[w];
      if (mask_remap && w == key_words - 1) v = mask_remap[v];
      out[w] = v;
    }
    if (gl == 0) perm[e] = dst;
  }
}

// unique index of every received key + compacted unique keys
template <int L>
__global__ void __launch_bounds__(256)
x_finish(const uint32_t* __restrict__ recv, uint32_t n, uint32_t key_words,
         const uint32_t* __restrict__ rep, const uint32_t* __restrict__ winner_rank,
         uint32_t* __restrict__ unique_index, uint32_t* __restrict__ unique_keys) {
  const uint32_t gl = lane_id() & (L - 1);
  const uint32_t total = gridDim.x * (blockDim.x / L);
  for (uint32_t e = blockIdx.x * (blockDim.x / L) + threadIdx.x / L; e < n; e += total) {
    const uint32_t q = rep[e] & ~kTentative;         // every rep is tentative here (empty pool)
    const uint32_t u = winner_rank[q];
    if (gl == 0) unique_index[e] = u;
    if (q == e) {
      const uint32_t* src = recv + (size_t)e * key_words;
      uint32_t* dst = unique_keys + (size_t)u * key_words;
      for (uint32_t w = gl; w < key_words; w += L) dst[w] = src[w];
    }
  }
}

__global__ void x_unpack(const uint32_t* __restrict__ returned, const uint32_t* __restrict__ perm,
                         uint32_t n, uint32_t* __restrict__ local_to_global) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) local_to_global[i] = returned[perm[i]];
}

}  // namespace pf

extern "C" int pf_exchange_pack(pf_ctx* ctx, int cluster_namespace, uint32_t world,
                                const uint32_t* mask_remap_dev, uint32_t* send_words_dev,
                                uint64_t capacity_patterns, uint64_t* counts_host) {
  if (!ctx || !counts_host || world == 0) return PF_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  TRY(finalize_pending(ctx));
  PatternSpace& s = cluster_namespace ? ctx->cp : ctx->kp;
  cudaStream_t st = ctx->stream;
  const uint32_t n = (uint32_t)s.n;
  if (n > capacity_patterns) return fail(ctx, PF_ERR_INVALID, "send buffer too small: %u patterns", n);
  if (mask_remap_dev && (cluster_namespace || !ctx->prm.consider_missing))
    return fail(ctx, PF_ERR_INVALID, "mask_remap only applies to k-mer patterns with consider_missing");
  TRY(dev_ensure(ctx, s.x_owner, std::max<size_t>(1, n) * 4));
  TRY(dev_ensure(ctx, s.x_pos, std::max<size_t>(1, n) * 4));
  TRY(dev_ensure(ctx, s.x_perm, std::max<size_t>(1, n) * 4));
  TRY(dev_ensure(ctx, s.x_counts, (size_t)world * 2 * 4));
  CU(cudaMemsetAsync(s.x_counts.p, 0, (size_t)world * 2 * 4, st));
  std::vector<uint32_t> counts(world, 0);
  if (n) {
    if (!send_words_dev) return fail(ctx, PF_ERR_INVALID, "null send buffer");
    if (world > 32) return fail(ctx, PF_ERR_UNSUPPORTED, "exchange over more than 32 ranks");
    TRY(pin_ensure(ctx, ctx->h_counters, C_COUNT * 4 + 32 * 4));
    uint32_t* hx = ctx->h_counters.as<uint32_t>() + C_COUNT;
    const int L = s.key_words <= 16 ? 4 : s.key_words <= 32 ? 8 : s.key_words <= 64 ? 16 : 32;
    const uint32_t cgrid = std::min<uint32_t>(cdiv(n, 256), kGridPersist * 2);
    const uint32_t pgrid = std::min<uint32_t>(cdiv(n, 256 / L), kGridPersist * 4);
#define PF_XC(LL)                                                                                          \
    do {                                                                                                   \
      x_classify<LL><<<cgrid, 256, 0, st>>>(s.pool.as<uint32_t>(), n, s.key_words, mask_remap_dev, world,  \
                                            s.x_owner.as<uint32_t>(), s.x_pos.as<uint32_t>(),              \
                                            s.x_counts.as<uint32_t>());                                    \
      x_offsets<<<1, 32, 0, st>>>(s.x_counts.as<uint32_t>(), world);                                       \
      x_pack<LL><<<pgrid, 256, 0, st>>>(s.pool.as<uint32_t>(), n, s.key_words, mask_remap_dev,             \
                                        s.x_owner.as<uint32_t>(), s.x_pos.as<uint32_t>(),                  \
                                        s.x_counts.as<uint32_t>() + world, send_words_dev,                 \
                                        s.x_perm.as<uint32_t>());                                          \
    } while (0)
    if (L == 4) PF_XC(4); else if (L == 8) PF_XC(8); else if (L == 16) PF_XC(16); else PF_XC(32);
#undef PF_XC
    mirror_counters<<<1, 32, 0, st>>>(hx, s.x_counts.as<uint32_t>(), world);
    ctx->launches += 3;
    CU(cudaStreamSynchronize(st));           // the only sync: bucket sizes for the caller's all-to-all
    CU(cudaGetLastError());
    for (uint32_t r = 0; r < world; ++r) counts[r] = hx[r];
  }
  for (uint32_t r = 0; r < world; ++r) counts_host[r] = counts[r];
  return PF_OK;
}

extern "C" int pf_exchange_dedup(pf_ctx* ctx, int cluster_namespace, const uint32_t* recv_words_dev,
                                 uint64_t n_recv, uint32_t* recv_unique_index_dev, uint64_t* n_unique_host) {
  if (!ctx || !n_unique_host) return PF_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  PatternSpace& s = cluster_namespace ? ctx->cp : ctx->kp;
  cudaStream_t st = ctx->stream;
  *n_unique_host = 0;
  s.x_n_unique = 0;
  if (n_recv == 0) return PF_OK;
  if (n_recv >= (1ull << 30)) return fail(ctx, PF_ERR_INVALID, "too many received patterns");
  if (!recv_words_dev || !recv_unique_index_dev) return fail(ctx, PF_ERR_INVALID, "null exchange buffer");
  const uint32_t n = (uint32_t)n_recv;
  uint32_t size = 1024;
  while (size < 2ull * n + 16) size *= 2;
  DevBuf& table = s.x_table;          // scratch kept across calls: no cudaMalloc in the steady state
  DevBuf& rep = s.x_rep;
  DevBuf& slot_of = s.x_slot;
  DevBuf& winner = s.x_winner;
  TRY(dev_ensure(ctx, table, (size_t)size * 4));
  TRY(dev_ensure(ctx, rep, (size_t)n * 4));
  TRY(dev_ensure(ctx, slot_of, (size_t)n * 4));
  TRY(dev_ensure(ctx, winner, ((size_t)n + 1) * 4));
  TRY(dev_ensure(ctx, s.x_unique, (size_t)n * s.key_words * 4));
  CU(cudaMemsetAsync(table.p, 0xff, (size_t)size * 4, st));
  uint32_t* counters = ctx->d_counters.p ? ctx->d_counters.as<uint32_t>() : nullptr;
  if (!counters) { TRY(dev_ensure(ctx, ctx->d_counters, C_COUNT * 4)); TRY(pin_ensure(ctx, ctx->h_counters, C_COUNT * 4)); counters = ctx->d_counters.as<uint32_t>(); }
  const uint32_t grid = std::min<uint32_t>(cdiv(n, 8), kGridPersist * 2);
  {
    // lanes per pattern as in K4: short keys would leave most of a warp idle
    const int L = s.key_words <= 16 ? 4 : s.key_words <= 32 ? 8 : s.key_words <= 64 ? 16 : 32;
    const uint32_t pgrid = std::min<uint32_t>(cdiv(n, 256 / L), kGridPersist * 4);
#define PF_XP(LL)                                                                                          \
    k4_probe<LL><<<pgrid, 256, 0, st>>>(recv_words_dev, n, s.key_words, nullptr, table.as<uint32_t>(), size - 1, \
                                        rep.as<uint32_t>(), slot_of.as<uint32_t>(), winner.as<uint32_t>())
    if (L == 4) PF_XP(4); else if (L == 8) PF_XP(8); else if (L == 16) PF_XP(16); else PF_XP(32);
#undef PF_XP
  }
  TRY(scan_inplace(ctx, winner.as<uint32_t>(), n, counters + C_NEW_KP));
  {
    const int L = s.key_words <= 16 ? 4 : s.key_words <= 32 ? 8 : s.key_words <= 64 ? 16 : 32;
    const uint32_t fgrid = std::min<uint32_t>(cdiv(n, 256 / L), kGridPersist * 4);
#define PF_XF(LL)                                                                                          \
    x_finish<LL><<<fgrid, 256, 0, st>>>(recv_words_dev, n, s.key_words, rep.as<uint32_t>(), winner.as<uint32_t>(), \
                                        recv_unique_index_dev, s.x_unique.as<uint32_t>())
    if (L == 4) PF_XF(4); else if (L == 8) PF_XF(8); else if (L == 16) PF_XF(16); else PF_XF(32);
#undef PF_XF
  }
  ctx->launches += 2;
  TRY(pin_ensure(ctx, ctx->h_counters, C_COUNT * 4 + 32 * 4));
  uint32_t* hx = ctx->h_counters.as<uint32_t>() + C_COUNT;
  mirror_counters<<<1, 32, 0, st>>>(hx, counters + C_NEW_KP, 1);
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  const uint32_t total = hx[0];
  s.x_n_unique = total;
  *n_unique_host = total;
  return PF_OK;
}

extern "C" int pf_exchange_unique_export(pf_ctx* ctx, int cluster_namespace, uint32_t* host_out) {
  if (!ctx) return PF_ERR_INVALID;
  PatternSpace& s = cluster_namespace ? ctx->cp : ctx->kp;
  if (s.x_n_unique == 0) return PF_OK;
  if (!host_out) return PF_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpy(host_out, s.x_unique.p, s.x_n_unique * s.key_words * 4, cudaMemcpyDeviceToHost));
  return PF_OK;
}

extern "C" int pf_exchange_unpack(pf_ctx* ctx, int cluster_namespace, const uint32_t* returned_ids_dev,
                                  uint32_t* local_to_global_dev) {
  if (!ctx) return PF_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  PatternSpace& s = cluster_namespace ? ctx->cp : ctx->kp;
  const uint32_t n = (uint32_t)s.n;
  if (n == 0) return PF_OK;
  if (!returned_ids_dev || !local_to_global_dev) return fail(ctx, PF_ERR_INVALID, "null exchange buffer");
  x_unpack<<<cdiv(n, 256), 256, 0, ctx->stream>>>(returned_ids_dev, s.x_perm.as<uint32_t>(), n, local_to_global_dev);
  ctx->launches++;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaGetLastError());
  return PF_OK;
}
