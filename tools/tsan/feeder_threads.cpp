#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/panfeed_b200.h"
int main(int argc, char** argv) {
  const int G = 32;
  std::vector<std::string> names, paths;
  for (int i = 0; i < G; ++i) { char b[256]; snprintf(b, 256, "s%04d", i); names.push_back(b); snprintf(b, 256, "%s/s%04d.gff", argv[1], i); paths.push_back(b); }
  std::vector<const char*> cn, cp;
  for (int i = 0; i < G; ++i) { cn.push_back(names[i].c_str()); cp.push_back(paths[i].c_str()); }
  pf_feeder* f; pf_feeder_create(&f);
  int rc = pf_feeder_add_genomes(f, G, cn.data(), cp.data(), nullptr, nullptr, 6);
  std::string blob; std::vector<uint32_t> genome;
  for (int c = 0; c < 300; ++c) for (int g = 0; g < G; ++g) { char b[64]; snprintf(b, 64, "s%04d_%05d", g, c); if (!blob.empty()) blob += '\n'; blob += b; genome.push_back(g); }
  pf_cut_result r; pf_cut_planes p;
  for (int rep = 0; rep < 2; ++rep) {
    int rc2 = pf_feeder_cut_packed(f, genome.size(), genome.data(), blob.data(), blob.size(), 100, 100, 0, 5, &r, &p);
    int rc3 = pf_feeder_cut(f, genome.size(), genome.data(), blob.data(), blob.size(), 50, 20, 0, &r);
    printf("rc %d %d %d seqs %u words %lu\n", rc, rc2, rc3, r.n_seqs, (unsigned long)p.n_words);
  }
  pf_table* t; pf_table_create(&t);
  const char* drop[2] = {"Non-unique Gene name", "Annotation"};
  const char* na[2] = {"", "NA"};
  rc = pf_table_load(t, argv[2], drop, 2, na, 2, 6);
  uint64_t nr; uint32_t nc; pf_table_shape(t, &nr, &nc);
  printf("table rc %d %lu x %u (%s)\n", rc, (unsigned long)nr, nc, pf_table_last_error(t));
  pf_table_destroy(t);
  pf_feeder_destroy(f);
}
