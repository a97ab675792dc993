// pf_synth_api.cuh — pf_synth_plan / pf_synth_fill: the synthetic pangenome of SURVEY.md 8(d).
// Part of libpanfeed_b200.so's single translation unit: included once, in order, by pf_api.cu.

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
