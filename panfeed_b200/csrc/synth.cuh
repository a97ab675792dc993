// Synthetic pangenome generator: deterministic, counter-based,
// so any (cluster, sample, copy, position) can be produced independently on the
// device.  Not part of the reference; it only feeds benchmarks and parity tests.
#pragma once
#include "pf_common.cuh"

namespace pf {

__host__ __device__ __forceinline__ uint64_t synth_hash(uint64_t seed, uint64_t a, uint64_t b,
                                                        uint64_t c, uint64_t tag) {
  uint64_t h = seed ^ (tag * 0x9e3779b97f4a7c15ULL);
  h = fmix64(h + a * 0xd6e8feb86659fd93ULL);
  h = fmix64(h + b * 0xca5a826395121157ULL);
  h = fmix64(h + c * 0x2545f4914f6cdd1dULL);
  return h;
}

enum : uint64_t { kTagAnc = 1, kTagFounderSnp = 2, kTagFounderPick = 3, kTagPrivate = 4,
                  kTagPresence = 5, kTagAccessoryP = 6, kTagParalog = 7, kTagStrand = 8,
                  kTagStart = 9, kTagCore = 10 };

struct SynthSeq {        // one generated sequence
  uint64_t word_off;     // first word in the packed plane
  uint32_t cluster;      // global cluster index
  uint32_t sample;
  uint32_t copy;
  uint32_t len;
};

// one thread = one 64-bit word = 32 bases
__global__ void synth_bases(const SynthSeq* __restrict__ seqs, uint32_t n_seqs,
                            uint32_t words_per_seq, uint64_t seed, uint32_t n_founders,
                            uint64_t founder_thr, uint64_t private_thr,
                            uint64_t* __restrict__ out) {
  const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t s = gid / words_per_seq;
  if (s >= n_seqs) return;
  const uint32_t wi = (uint32_t)(gid % words_per_seq);
  const SynthSeq q = seqs[s];
  const uint64_t inst = ((uint64_t)q.sample << 8) | q.copy;
  const uint32_t founder = (uint32_t)(synth_hash(seed, q.cluster, inst, 0, kTagFounderPick) % n_founders);
  uint64_t word = 0;
  for (uint32_t j = 0; j < 32; ++j) {
    const uint32_t pos = wi * 32 + j;
    uint32_t b = 0;
    if (pos < q.len) {
      b = (uint32_t)(synth_hash(seed, q.cluster, pos, 0, kTagAnc) >> 61) & 3u;
      const uint64_t hf = synth_hash(seed, q.cluster, pos, founder, kTagFounderSnp);
      if (hf < founder_thr) b = (b + 1u + (uint32_t)((hf >> 3) % 3u)) & 3u;
      const uint64_t hp = synth_hash(seed, q.cluster, pos, inst, kTagPrivate);
      if (hp < private_thr) b = (b + 1u + (uint32_t)((hp >> 3) % 3u)) & 3u;
    }
    word |= (uint64_t)b << (62 - 2 * j);
  }
  out[q.word_off + wi] = word;
}

}  // namespace pf
