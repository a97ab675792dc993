// Native (host threads) row filter over the TSV files the path writes, for the joins that close
// the two-pass loop: panfeed-get-clusters (/root/reference/panfeed/get_clusters.py:90-101) and
// panfeed-get-kmers (get_kmers.py:108-145) read kmers_to_hashes.tsv / kmers.tsv through pandas in
// chunks of 100,000 rows and keep the rows whose hash / cluster is in a set:
//
//     pd.concat([x[x['hashed_pattern'].isin(passing_hashes)] for x in iter_h])
//
// On first-pass outputs of 1e8-1e9 rows that scan is the whole run time.  Here the file is mapped,
// cut into pieces on line boundaries, and every piece is scanned by a thread: one pass over the
// bytes, one hash-set probe per row, the matching lines copied out in file order.
// Fields are split on tabs only (the files panfeed writes hold no quoted fields).  No device code.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/panfeed_b200.h"
#include "pf_host.h"

extern "C" void pf_free(void* p) { free(p); }

extern "C" int pf_tsv_filter(const char* path, uint32_t column, const char* keys_blob, const uint64_t* key_off,
                             uint64_t n_keys, int skip_header, char** out, uint64_t* out_len, uint64_t* n_rows,
                             uint32_t n_threads) {
  if (!path || !out || !out_len || (n_keys && (!keys_blob || !key_off))) return PF_ERR_INVALID;
  *out = nullptr;
  *out_len = 0;
  if (n_rows) *n_rows = 0;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return PF_ERR_INVALID;
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); return PF_ERR_INVALID; }
  const size_t size = (size_t)st.st_size;
  if (size == 0) { close(fd); return PF_OK; }
  void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return PF_ERR_NOMEM;
  const char* data = static_cast<const char*>(map);
  const char* end = data + size;
  std::unordered_set<std::string_view> keys;
  keys.reserve((size_t)n_keys * 2 + 16);
  for (uint64_t i = 0; i < n_keys; ++i)
    keys.emplace(keys_blob + key_off[i], (size_t)(key_off[i + 1] - key_off[i]));
  const char* body = data;
  if (skip_header) {
    const char* nl = (const char*)memchr(data, '\n', size);
    body = nl ? nl + 1 : end;
  }
  const size_t body_size = (size_t)(end - body);
  const uint32_t hw = pf_host_threads();
  uint32_t nt = n_threads ? n_threads : hw;
  nt = (uint32_t)std::max<size_t>(1, std::min<size_t>(nt, body_size >> 22));      // >= 4 MiB per thread
  // piece boundaries on line starts
  std::vector<const char*> cut(nt + 1);
  cut[0] = body;
  cut[nt] = end;
  for (uint32_t t = 1; t < nt; ++t) {
    const char* p = body + body_size * t / nt;
    if (p < cut[t - 1]) p = cut[t - 1];
    const char* nl = p < end ? (const char*)memchr(p, '\n', (size_t)(end - p)) : nullptr;
    cut[t] = nl ? nl + 1 : end;
  }
  std::vector<std::string> part(nt);
  std::vector<uint64_t> rows(nt, 0);
  auto scan = [&](uint32_t t) {
    std::string& o = part[t];
    const char* p = cut[t];
    const char* e = cut[t + 1];
    while (p < e) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
      const char* le = nl ? nl : e;                       // line without its newline
      const char* l1 = (le > p && le[-1] == '\r') ? le - 1 : le;
      // field `column`
      const char* f0 = p;
      uint32_t c = 0;
      while (c < column && f0 <= l1) {
        const char* tab = (const char*)memchr(f0, '\t', (size_t)(l1 - f0));
        if (!tab) { f0 = l1 + 1; break; }
        f0 = tab + 1;
        ++c;
      }
      if (c == column && f0 <= l1) {
        const char* tab = (const char*)memchr(f0, '\t', (size_t)(l1 - f0));
        const char* f1 = tab ? tab : l1;
        if (keys.count(std::string_view(f0, (size_t)(f1 - f0)))) {
          o.append(p, (size_t)(l1 - p));
          o.push_back('\n');
          ++rows[t];
        }
      }
      p = nl ? nl + 1 : e;
    }
  };
  if (nt == 1) scan(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nt; ++t) th.emplace_back(scan, t);
    for (auto& x : th) x.join();
  }
  munmap(map, size);
  uint64_t total = 0, nr = 0;
  for (uint32_t t = 0; t < nt; ++t) { total += part[t].size(); nr += rows[t]; }
  if (n_rows) *n_rows = nr;
  *out_len = total;
  if (total == 0) return PF_OK;
  char* buf = static_cast<char*>(malloc(total));
  if (!buf) return PF_ERR_NOMEM;
  uint64_t at = 0;
  for (uint32_t t = 0; t < nt; ++t) {
    memcpy(buf + at, part[t].data(), part[t].size());
    at += part[t].size();
  }
  *out = buf;
  return PF_OK;
}
