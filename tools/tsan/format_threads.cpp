#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../../include/panfeed_b200.h"
int main() {
  std::mt19937_64 rng(1);
  const uint32_t nc = 64, n = 200000, k = 31, S = 500, W = 16;
  std::vector<uint32_t> cl(nc), rc(n), rp(n); std::vector<uint64_t> rk(n);
  for (uint32_t c = 0; c < nc; ++c) cl[c] = c % 10;
  for (uint32_t i = 0; i < n; ++i) { rc[i] = (i * (uint64_t)nc) / n; rk[i] = rng() >> 2; rp[i] = rng() % 5000; }
  pf_batch_result r; memset(&r, 0, sizeof r);
  r.n_clusters = nc; r.cluster_pattern = cl.data(); r.n_rows = n; r.row_cluster = rc.data(); r.row_kmer = rk.data(); r.row_pattern = rp.data();
  std::string tags; std::vector<uint64_t> toff(1, 0);
  for (uint32_t c = 0; c < nc; ++c) { tags += "group_" + std::to_string(c); toff.push_back(tags.size()); }
  std::vector<char> kid(5000 * 24, 'x'), cid(10 * 24, 'y');
  uint64_t need = 0; std::vector<uint64_t> coff(nc + 1);
  int a = pf_format_kmer_rows(&r, k, tags.data(), toff.data(), kid.data(), 5000, cid.data(), 10, nullptr, 0, &need, coff.data(), 6);
  std::vector<char> out(need);
  int b = pf_format_kmer_rows(&r, k, tags.data(), toff.data(), kid.data(), 5000, cid.data(), 10, out.data(), out.size(), &need, coff.data(), 6);
  // interleaved clusters: the counting-sort path
  for (uint32_t i = 0; i < n; ++i) rc[i] = rng() % nc;
  int b2 = pf_format_kmer_rows(&r, k, tags.data(), toff.data(), kid.data(), 5000, cid.data(), 10, out.data(), out.size(), &need, coff.data(), 6);
  printf("kmer rows %d %d %d, %lu bytes\n", a, b, b2, (unsigned long)need);
  const uint32_t np = 60000;
  std::vector<uint32_t> words((size_t)np * W), pres((size_t)np * W); for (auto& x : words) x = (uint32_t)rng(); for (auto& x : pres) x = (uint32_t)rng();
  std::vector<char> ids((size_t)np * 24, 'z');
  for (const uint32_t* pw : {(const uint32_t*)nullptr, (const uint32_t*)pres.data()}) {
    pf_format_patterns(words.data(), np, W, S, ids.data(), pw, W, nullptr, 0, &need, 6);
    std::vector<char> o2(need);
    int c = pf_format_patterns(words.data(), np, W, S, ids.data(), pw, W, o2.data(), o2.size(), &need, 6);
    printf("patterns %d, %lu bytes\n", c, (unsigned long)need);
  }
  std::vector<uint8_t> dig((size_t)np * 16); for (auto& x : dig) x = (uint8_t)rng();
  std::vector<char> b64((size_t)np * 24);
  printf("base64 %d\n", pf_base64_ids(dig.data(), np, b64.data(), 6));
  // compact positions
  const uint32_t nseq = 600, L = 700, wp = (L + 63) / 64 * 2;
  std::vector<uint64_t> packed((size_t)nseq * wp); for (auto& x : packed) x = rng();
  std::vector<pf_seq_desc> seqs(nseq); memset(seqs.data(), 0, nseq * sizeof(pf_seq_desc));
  for (uint32_t i = 0; i < nseq; ++i) { seqs[i].base_off = (uint64_t)i * wp * 32; seqs[i].len = L; seqs[i].flags = PF_SEQ_TARGET; seqs[i].start = 990 + i; seqs[i].end = seqs[i].start + L - 1; seqs[i].offset = i % 120; seqs[i].strand = i % 2 ? 1 : -1; }
  pf_batch bt; memset(&bt, 0, sizeof bt); bt.packed_bases = packed.data(); bt.n_words = packed.size(); bt.seqs = seqs.data(); bt.n_seqs = nseq;
  std::vector<uint32_t> bits(packed.size()); for (auto& x : bits) x = (uint32_t)rng();
  std::string leads; std::vector<uint64_t> loff(1, 0);
  for (uint32_t i = 0; i < nseq; ++i) { leads += "cl\tstrain\tgene\tctg\t1\t"; loff.push_back(leads.size()); }
  pf_format_positions_compact(&bt, bits.data(), k, 1, 0, nseq, leads.data(), loff.data(), nullptr, 0, &need, 6);
  std::vector<char> o3(need);
  int d = pf_format_positions_compact(&bt, bits.data(), k, 1, 0, nseq, leads.data(), loff.data(), o3.data(), o3.size(), &need, 6);
  printf("positions %d, %lu bytes\n", d, (unsigned long)need);
}
