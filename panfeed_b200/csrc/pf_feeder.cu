// Native feeder (host only): GFF3 + FASTA -> the cut, oriented sequences of gene clusters.
//
// What the reference does per cluster in Python (/root/reference/panfeed/input.py): parse_gff
// (:274-332, CDS features with an ID= attribute), the pyfaidx contigs (:262-266, upper case),
// and iter_gene_clusters (:335-468: per present strain the ';'-separated feature ids of the
// panaroo cell, the up/downstream window of :413-446, reverse complement on the minus strand).
// At ~1 us of interpreter time per attribute access that loop cannot feed one GPU at BASELINE
// configs #4 / #5.  Here the genomes are parsed once into flat arrays and a
// cluster is cut by one call that appends ASCII sequences + descriptors to buffers the packer
// (pf_pack_2bit / pf_pack_4bit) reads directly.  Semantics follow panfeed_b200/input.py (the Python
// mirror of the reference, which the CPU tests compare this file with) line by line, including
// Python's slice clamping and the tolerance for malformed GFF lines.  No device code in this file.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/panfeed_b200.h"
#include "pf_host.h"

namespace {

struct FeatureRec {
  std::string id, contig;
  int64_t start = 0, end = 0;
  int32_t strand = 1;
};

// The text of one input file: a private mapping of the file (no copy out of the page cache and
// no first touch of as much fresh memory, which costs more than the parsing itself and does
// not scale over threads) or, where a file cannot be mapped, a buffer read with fread.
struct FileText {
  char* p = nullptr;
  size_t n = 0, mapped = 0;
  std::string owned;
  FileText() = default;
  FileText(const FileText&) = delete;
  FileText& operator=(const FileText&) = delete;
  FileText(FileText&& o) noexcept { *this = std::move(o); }
  FileText& operator=(FileText&& o) noexcept {
    if (this != &o) {
      release();
      const bool in_owned = !o.mapped && o.p;
      owned = std::move(o.owned);
      p = in_owned ? &owned[0] : o.p;
      n = o.n; mapped = o.mapped;
      o.p = nullptr; o.n = 0; o.mapped = 0;
    }
    return *this;
  }
  ~FileText() { release(); }
  void release() {
    if (mapped) munmap(p, mapped);
    p = nullptr; n = 0; mapped = 0;
    owned = std::string();
  }
  const char* data() const { return p; }
  size_t size() const { return n; }
};

// A contig is either a string of its own (upper case, the copying parser) or a VIEW of the
// genome's file text: a FASTA record whose lines all hold `width` symbols (the last one may be
// shorter) needs no copy - base i sits at lines[i + i / width] - and no second first touch of
// as much memory again; its case is folded when a window is cut.
struct Contig {
  std::string owned;
  const char* lines = nullptr;
  uint64_t width = 0, n = 0;
  uint64_t size() const { return lines ? n : (uint64_t)owned.size(); }
};

struct Genome {
  std::string name;
  std::vector<FeatureRec> features;
  std::unordered_map<std::string, uint32_t> feature_of;       // id -> index (the last one wins)
  std::vector<std::string> contig_name;
  std::vector<Contig> contigs;
  std::unordered_map<std::string, uint32_t> contig_of;        // name -> index (the last one wins)
  FileText text;                                              // the file the views point into
};

inline bool str_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }

// Python's int(str): optional whitespace, sign, decimal digits with single underscores between them
bool py_int(const char* p, const char* e, int64_t* out) {
  while (p < e && str_space((unsigned char)*p)) ++p;
  while (e > p && str_space((unsigned char)e[-1])) --e;
  if (p == e) return false;
  bool neg = false;
  if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
  if (p == e || *p < '0' || *p > '9') return false;
  int64_t v = 0;
  bool last_us = false;
  for (; p < e; ++p) {
    if (*p == '_') { if (last_us) return false; last_us = true; continue; }
    if (*p < '0' || *p > '9') return false;
    last_us = false;
    if (v > (INT64_MAX - 9) / 10) return false;
    v = v * 10 + (*p - '0');
  }
  if (last_us) return false;
  *out = neg ? -v : v;
  return true;
}

// text-mode reading like Python's open(path): "\r\n" and lone "\r" become "\n" - in place (the
// mapping is private: copy on write), and only if the text holds a '\r' at all.
bool read_text(const char* path, FileText* out) {
  out->release();
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
    if (m != MAP_FAILED) {
      // (no MADV_SEQUENTIAL: the pages are read again, at random, when windows are cut)
      out->p = (char*)m;
      out->n = out->mapped = (size_t)st.st_size;
    }
  }
  if (!out->mapped) {                                 // a pipe, an empty file, no mmap: read it
    std::string& b = out->owned;
    b.reserve(1 << 16);                               // (never the in-object small-string buffer: views point here)
    char buf[1 << 16];
    ssize_t got;
    while ((got = read(fd, buf, sizeof buf)) > 0) b.append(buf, (size_t)got);
    out->p = &b[0];
    out->n = b.size();
  }
  close(fd);
  char* b = out->p;
  const size_t n = out->n;
  const char* first = n ? (const char*)memchr(b, '\r', n) : nullptr;
  if (first) {
    size_t w = (size_t)(first - b);
    for (size_t i = w; i < n; ++i) {
      if (b[i] == '\r') {
        b[w++] = '\n';
        if (i + 1 < n && b[i + 1] == '\n') ++i;
      } else {
        b[w++] = b[i];
      }
    }
    out->n = w;
  }
  return true;
}

// input.py parse_gff (reference :274-332): lines up to "##FASTA"; cols[2] == "CDS"; int(cols[3]),
// int(cols[4]); strand '+' -> 1 else -1; every ';' field of cols[8] that starts with "ID" and holds
// '=' sets the id to the text between its first and second '='; lines that raise are skipped.
void parse_gff_text(const char* p, const char* e, Genome* g, uint32_t* skipped) {
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', e - p);
    const char* le = nl ? nl + 1 : e;                 // the line, newline included (as Python iterates)
    const char* h = p;
    while (h < le && str_space((unsigned char)*h)) ++h;
    if (le - h >= 7 && memcmp(h, "##FASTA", 7) == 0) break;
    if (h < le && *h == '#') { p = le; continue; }
    const char* col[10];
    int nc = 0;
    col[nc++] = p;
    for (const char* q = p; q < le && nc < 10; ++q)
      if (*q == '\t') col[nc++] = q + 1;
    // col[i] .. col[i+1]-1 is column i; the last parsed column runs to the end of the line unless more tabs follow
    auto col_end = [&](int i) -> const char* {
      if (i + 1 < nc) return col[i + 1] - 1;
      const char* q = col[i];
      while (q < le && *q != '\t') ++q;
      return q;
    };
    bool bad = false;
    do {
      if (nc < 3) { bad = true; break; }              // cols[2] raises IndexError
      const char* t0 = col[2];
      const char* t1 = col_end(2);
      if (!(t1 - t0 == 3 && memcmp(t0, "CDS", 3) == 0)) break;      // not wanted: skipped silently
      if (nc < 5) { bad = true; break; }
      int64_t start, end;
      if (!py_int(col[3], col_end(3), &start) || !py_int(col[4], col_end(4), &end)) { bad = true; break; }
      if (nc < 7) { bad = true; break; }
      const int32_t strand = (col_end(6) - col[6] == 1 && *col[6] == '+') ? 1 : -1;
      if (nc < 9) { bad = true; break; }
      const char* a0 = col[8];
      const char* a1 = col_end(8);
      bool have = false;
      std::string ident;
      for (const char* f0 = a0; f0 <= a1;) {
        const char* f1 = (const char*)memchr(f0, ';', a1 - f0);
        if (!f1) f1 = a1;
        if (f1 - f0 >= 2 && f0[0] == 'I' && f0[1] == 'D') {
          const char* eq = (const char*)memchr(f0, '=', f1 - f0);
          if (eq) {
            const char* eq2 = (const char*)memchr(eq + 1, '=', f1 - (eq + 1));
            ident.assign(eq + 1, eq2 ? eq2 : f1);
            have = true;
          }
        }
        f0 = f1 + 1;
      }
      if (!have) break;
      FeatureRec fr;
      fr.id = ident;
      fr.contig.assign(col[0], col_end(0));
      fr.start = start; fr.end = end; fr.strand = strand;
      auto it = g->feature_of.find(ident);
      if (it == g->feature_of.end()) {
        g->feature_of.emplace(ident, (uint32_t)g->features.size());
        g->features.push_back(std::move(fr));
      } else {
        g->features[it->second] = std::move(fr);      // dict assignment: the last line wins
      }
    } while (false);
    if (bad) ++*skipped;
    p = le;
  }
}

// A FASTA record as a view (see Contig): body .. the next line that starts with '>' (or e).  Regular
// means: every line holds the same number of symbols except a shorter last one, no line starts or
// ends with whitespace (the copying parser would strip it), blank lines only after the last one.
// Returns the end of the record, or nullptr if the record has to be copied.
const char* fasta_record_view(const char* body, const char* e, Contig* c) {
  const char* q = body;
  uint64_t width = 0, n = 0;
  bool last_short = false, ended = false;
  while (q < e && *q != '>') {
    const char* nl = (const char*)memchr(q, '\n', e - q);
    const char* le = nl ? nl : e;
    const uint64_t len = (uint64_t)(le - q);
    if (len == 0) {
      ended = true;
    } else {
      if (ended || last_short || str_space((unsigned char)q[0]) || str_space((unsigned char)le[-1])) return nullptr;
      if (width == 0) width = len;
      else if (len > width) return nullptr;
      if (len < width) last_short = true;
      n += len;
    }
    q = nl ? nl + 1 : e;
  }
  if (n == 0) return nullptr;                       // nothing to point at: an (empty) string of its own
  c->owned.clear();
  c->lines = body;
  c->width = width;
  c->n = n;
  return q;
}

// input.py read_fasta_text: '>' lines name a record (first whitespace-separated word, "" if none),
// other lines are stripped and appended; sequences are upper-cased; a repeated name replaces.
// views: [p, e) outlives the genome (it is the genome's own `text`): regular records stay where they are.
void parse_fasta_text(const char* p, const char* e, Genome* g, bool views) {
  bool have = false;
  std::string name;
  Contig cur;
  auto close = [&]() {
    if (!have) return;
    auto it = g->contig_of.find(name);
    if (it == g->contig_of.end()) {
      g->contig_of.emplace(name, (uint32_t)g->contig_name.size());
      g->contig_name.push_back(name);
      g->contigs.push_back(std::move(cur));
    } else {
      g->contigs[it->second] = std::move(cur);
    }
    cur = Contig();
  };
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', e - p);
    const char* le = nl ? nl : e;
    const char* l1 = le;
    while (l1 > p && (l1[-1] == '\r' || l1[-1] == '\n')) --l1;
    if (l1 > p && *p == '>') {
      close();
      const char* q = p + 1;
      while (q < l1 && str_space((unsigned char)*q)) ++q;
      const char* w = q;
      while (w < l1 && !str_space((unsigned char)*w)) ++w;
      name.assign(q, w);
      have = true;
      const char* body = nl ? nl + 1 : e;
      if (views) {
        const char* end = fasta_record_view(body, e, &cur);
        if (end) { p = end; continue; }
      }
      // one allocation per contig: its lines end where the next record starts (growing the string
      // line by line re-allocates and page-faults its way up, which serialises the parsing threads)
      const char* next = body < e ? (*body == '>' ? body : (const char*)memmem(body, (size_t)(e - body), "\n>", 2)) : e;
      cur.owned.reserve((size_t)((next ? next : e) - body));
    } else if (have) {
      const char* a = p;
      const char* b = l1;
      while (a < b && str_space((unsigned char)*a)) ++a;
      while (b > a && str_space((unsigned char)b[-1])) --b;
      std::string& seq = cur.owned;
      const size_t at = seq.size();
      seq.append(a, b);
      char* q = &seq[at];                               // str.upper() of the reference's contigs (ASCII)
      for (size_t i = 0, m = (size_t)(b - a); i < m; ++i) {
        const unsigned char c = (unsigned char)q[i];
        q[i] = (char)(c - (((unsigned)(c - 'a') < 26u) << 5));
      }
    }
    p = nl ? nl + 1 : e;
  }
  close();
}

// pyfaidx's complement table restricted to upper case (input.py:448-452 uses it on upper-case contigs)
struct CompLut {
  unsigned char t[256];
  CompLut() {
    for (int i = 0; i < 256; ++i) t[i] = (unsigned char)i;
    const char* a = "ACTGNactgnYRWSKMDVHBXyrwskmdvhbx";
    const char* b = "TGACNtgacnRYWSMKHBDVXrywsmkhbdvx";
    for (int i = 0; a[i]; ++i) t[(unsigned char)a[i]] = (unsigned char)b[i];
  }
};
const CompLut kComp;

// bases [lo, hi) of a contig as contiguous bytes: where they are when the contig owns its text
// (upper case), gathered line by line into `tmp` when it is a view of the file (case as in the
// file: the consumers fold it)
const unsigned char* contig_window(const Contig& c, int64_t lo, int64_t hi, std::vector<unsigned char>& tmp) {
  if (!c.lines) return reinterpret_cast<const unsigned char*>(c.owned.data()) + lo;
  tmp.resize((size_t)(hi - lo));
  unsigned char* dst = tmp.data();
  uint64_t pos = (uint64_t)lo;
  const uint64_t end = (uint64_t)hi;
  while (pos < end) {
    const uint64_t col = pos % c.width;
    const uint64_t take = std::min<uint64_t>(c.width - col, end - pos);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(c.lines) + pos + pos / c.width;
    memcpy(dst, src, (size_t)take);
    dst += take;
    pos += take;
  }
  return tmp.data();
}

}  // namespace

struct pf_feeder {
  std::vector<Genome> genomes;
  std::string err;
  // result of the last pf_feeder_cut
  std::unique_ptr<char[]> ascii;
  size_t ascii_cap = 0;
  std::vector<uint64_t> seq_off;
  std::vector<uint32_t> cell, feature;
  std::vector<int32_t> start, end, offset, strand;
  std::vector<uint32_t> miss_cell;
  std::vector<uint8_t> miss_kind;
  std::string miss_text;
  std::vector<uint64_t> miss_off;
  // planes of the last pf_feeder_cut_packed (capacity is kept from call to call)
  std::unique_ptr<uint64_t[]> packed, amb_plane;
  size_t packed_cap = 0, amb_cap = 0;
  std::vector<uint64_t> base_off, amb_off;
  std::vector<uint8_t> is_amb;
};

extern "C" int pf_feeder_create(pf_feeder** out) {
  if (!out) return PF_ERR_INVALID;
  *out = new pf_feeder();
  return PF_OK;
}

extern "C" void pf_feeder_destroy(pf_feeder* f) { delete f; }

extern "C" const char* pf_feeder_last_error(const pf_feeder* f) { return f ? f->err.c_str() : "null feeder"; }

namespace {
// views: gff (or fasta, when given) is g.text and stays alive with the genome
int add_genome_parsed(pf_feeder* f, Genome&& g_in, const char* gff, uint64_t gff_len, const char* fasta,
                      uint64_t fasta_len, uint32_t* skipped_lines, bool views) {
  Genome g = std::move(g_in);
  const char* name = g.name.c_str();
  uint32_t skipped = 0;
  parse_gff_text(gff, gff + gff_len, &g, &skipped);
  if (fasta) {
    parse_fasta_text(fasta, fasta + fasta_len, &g, views);
  } else {
    // open(gff).read().split("##FASTA")[1]: the text between the first and the second marker
    const char* a = gff_len ? (const char*)memmem(gff, gff_len, "##FASTA", 7) : nullptr;
    if (!a) {
      f->err = std::string("genome ") + name + ": no FASTA file and no ##FASTA section in the GFF";
      return PF_ERR_INVALID;
    }
    const char* end = gff + gff_len;
    const char* b = (const char*)memmem(a + 7, (size_t)(end - (a + 7)), "##FASTA", 7);
    if (!b) b = end;
    parse_fasta_text(a + 7, b, &g, views);
  }
  if (skipped_lines) *skipped_lines = skipped;
  f->genomes.push_back(std::move(g));
  return (int)f->genomes.size() - 1;
}
}  // namespace

extern "C" int pf_feeder_add_genome_text(pf_feeder* f, const char* name, const char* gff, uint64_t gff_len,
                                         const char* fasta, uint64_t fasta_len, uint32_t* skipped_lines) {
  if (!f || !name || (!gff && gff_len)) return PF_ERR_INVALID;
  Genome g;
  g.name = name;
  return add_genome_parsed(f, std::move(g), gff, gff_len, fasta, fasta_len, skipped_lines, false);   // the caller's memory
}

namespace {
int add_genome_files(pf_feeder* f, const char* name, const char* gff_path, const char* fasta_path,
                     uint32_t* skipped_lines) {
  FileText gff, fasta;
  if (!read_text(gff_path, &gff)) { f->err = std::string("cannot read ") + gff_path; return PF_ERR_INVALID; }
  if (fasta_path && !read_text(fasta_path, &fasta)) { f->err = std::string("cannot read ") + fasta_path; return PF_ERR_INVALID; }
  // the text that holds the sequences becomes the genome's: regular FASTA records are used in place
  Genome g;
  g.name = name;
  g.text = std::move(fasta_path ? fasta : gff);     // (mapping or heap buffer: the address stays)
  const char* t = g.text.data();
  const uint64_t tn = g.text.size();
  const int rc = fasta_path ? add_genome_parsed(f, std::move(g), gff.data(), gff.size(), t, tn, skipped_lines, true)
                            : add_genome_parsed(f, std::move(g), t, tn, nullptr, 0, skipped_lines, true);
  if (rc >= 0) {
    bool used = false;
    for (const Contig& c : f->genomes[rc].contigs) used |= c.lines != nullptr;
    if (!used) f->genomes[rc].text.release();       // nothing points into it
  }
  return rc;
}
}  // namespace

extern "C" int pf_feeder_add_genome(pf_feeder* f, const char* name, const char* gff_path, const char* fasta_path,
                                    uint32_t* skipped_lines) {
  if (!f || !name || !gff_path) return PF_ERR_INVALID;
  return add_genome_files(f, name, gff_path, fasta_path, skipped_lines);
}

// Many genomes at once: the files are read and parsed on host threads (a genome is a few MB of
// text; 500 of them are seconds of single-threaded parsing), then appended in the order given.
// Returns the index of the first one; the others follow consecutively.
extern "C" int pf_feeder_add_genomes(pf_feeder* f, uint32_t n, const char* const* names, const char* const* gff_paths,
                                     const char* const* fasta_paths /* NULL, or NULL entries: ##FASTA section */,
                                     uint32_t* skipped_lines /* [n] or NULL */, uint32_t n_threads) {
  if (!f || (n && (!names || !gff_paths))) return PF_ERR_INVALID;
  const int first = (int)f->genomes.size();
  if (n == 0) return first;
  std::vector<pf_feeder> local(n);                       // one scratch feeder per genome: no shared state
  std::vector<int> rc(n, PF_OK);
  uint32_t nt = n_threads ? n_threads : pf_host_threads();
  nt = std::min(nt, n);
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < nt; ++t)
    th.emplace_back([&, t]() {
      for (uint32_t i = t; i < n; i += nt) {
        uint32_t sk = 0;
        if (!names[i] || !gff_paths[i]) { rc[i] = PF_ERR_INVALID; continue; }
        rc[i] = add_genome_files(&local[i], names[i], gff_paths[i], fasta_paths ? fasta_paths[i] : nullptr, &sk);
        if (skipped_lines) skipped_lines[i] = sk;
      }
    });
  for (auto& x : th) x.join();
  for (uint32_t i = 0; i < n; ++i)
    if (rc[i] < 0) { f->err = local[i].err; return rc[i]; }
  for (uint32_t i = 0; i < n; ++i) f->genomes.push_back(std::move(local[i].genomes[0]));
  return first;
}

extern "C" int pf_feeder_genome_info(const pf_feeder* f, uint32_t genome, uint32_t* n_features, uint32_t* n_contigs,
                                     uint64_t* n_bases) {
  if (!f || genome >= f->genomes.size()) return PF_ERR_INVALID;
  const Genome& g = f->genomes[genome];
  if (n_features) *n_features = (uint32_t)g.features.size();
  if (n_contigs) *n_contigs = (uint32_t)g.contig_name.size();
  if (n_bases) { uint64_t n = 0; for (auto& c : g.contigs) n += c.size(); *n_bases = n; }
  return PF_OK;
}

extern "C" int pf_feeder_contig(const pf_feeder* f, uint32_t genome, uint32_t contig, const char** name,
                                uint64_t* n_bases, uint32_t* in_place) {
  if (!f || genome >= f->genomes.size() || contig >= f->genomes[genome].contigs.size()) return PF_ERR_INVALID;
  const Genome& g = f->genomes[genome];
  if (name) *name = g.contig_name[contig].c_str();
  if (n_bases) *n_bases = g.contigs[contig].size();
  if (in_place) *in_place = g.contigs[contig].lines ? 1u : 0u;
  return PF_OK;
}

extern "C" int pf_feeder_feature(const pf_feeder* f, uint32_t genome, uint32_t feature, const char** id,
                                 const char** contig, int64_t* start, int64_t* end, int32_t* strand) {
  if (!f || genome >= f->genomes.size() || feature >= f->genomes[genome].features.size()) return PF_ERR_INVALID;
  const FeatureRec& r = f->genomes[genome].features[feature];
  if (id) *id = r.id.c_str();
  if (contig) *contig = r.contig.c_str();
  if (start) *start = r.start;
  if (end) *end = r.end;
  if (strand) *strand = r.strand;
  return PF_OK;
}

// One cluster (input.py:335-468 / cut_window): cells_blob holds n_cells panaroo cells separated by
// '\n'; cell i belongs to genome genome[i].  For every ';'-separated feature id, in order: the
// feature and its contig are looked up (a miss is recorded, the gene skipped), the window
// [a, b) is sliced with Python's clamping and, on the minus strand, reverse-complemented.
namespace {
struct Piece { const Contig* contig; int64_t lo, hi; bool minus; };

// pass 1: look the genes up, place the windows.  The cells are independent, so ranges of them go
// to host threads (a look-up is two hash probes, ~0.5 us: on a many-core host the serial form is
// what bounds a cut); every range fills its own arrays, which are laid end to end afterwards.
struct Placed {
  std::vector<Piece> pieces;
  std::vector<uint32_t> cell, feature;
  std::vector<int32_t> start, end, offset, strand;
  std::vector<uint32_t> miss_cell;
  std::vector<uint8_t> miss_kind;
  std::string miss_text;
  std::vector<uint64_t> miss_off;          // ends of the names inside miss_text
};

void place_cells(const pf_feeder* f, uint32_t c0, uint32_t c1, const uint32_t* genome,
                 const std::vector<const char*>& cell_at, int32_t up, int32_t down, int32_t down_start_codon,
                 Placed& o) {
  std::string gene;
  for (uint32_t ci = c0; ci < c1; ++ci) {
    const Genome& g = f->genomes[genome[ci]];
    const char* p = cell_at[ci];
    const char* ce = cell_at[ci + 1] - 1;          // the separator (or the end of the blob) after the cell
    for (const char* g0 = p; g0 <= ce;) {
      const char* g1 = (const char*)memchr(g0, ';', ce - g0);
      if (!g1) g1 = ce;
      gene.assign(g0, g1);
      g0 = g1 + 1;
      auto miss = [&](uint8_t kind, const std::string& what) {
        o.miss_cell.push_back(ci);
        o.miss_kind.push_back(kind);
        o.miss_text += what;
        o.miss_off.push_back(o.miss_text.size());
      };
      auto fi = g.feature_of.find(gene);
      if (fi == g.feature_of.end()) { miss(0, gene); continue; }
      const FeatureRec& ft = g.features[fi->second];
      auto cit = g.contig_of.find(ft.contig);
      if (cit == g.contig_of.end()) { miss(1, ft.contig); continue; }
      const Contig& contig = g.contigs[cit->second];
      // cut_window (reference input.py:413-446)
      const bool over_up = ft.strand > 0 && ft.start - 1 - up < 0;
      const bool over_down = ft.strand < 0 && ft.start - 1 - down < 0;
      const int64_t offset = over_up ? ft.start - 1 : up;
      const int64_t offset_d = over_down ? ft.start - 1 : down;
      int64_t a, b, seq_start, seq_end;
      if (ft.strand > 0) {
        a = ft.start - 1 - offset;
        seq_start = ft.start - offset;
        b = seq_end = (down_start_codon ? ft.start : ft.end) + offset_d;
      } else {
        b = seq_end = ft.end + offset;
        if (down_start_codon) { a = ft.end - 1 - offset_d; seq_start = ft.end - offset_d; }
        else { a = ft.start - 1 - offset_d; seq_start = ft.start - offset_d; }
      }
      // contig[a:b] with Python's slice semantics
      const int64_t n = (int64_t)contig.size();
      const int64_t lo = a < 0 ? std::max<int64_t>(a + n, 0) : std::min(a, n);
      int64_t hi = b < 0 ? std::max<int64_t>(b + n, 0) : std::min(b, n);
      if (hi < lo) hi = lo;
      o.pieces.push_back(Piece{&contig, lo, hi, ft.strand < 0});
      o.cell.push_back(ci);
      o.feature.push_back(fi->second);
      o.start.push_back((int32_t)seq_start);
      o.end.push_back((int32_t)seq_end);
      o.offset.push_back((int32_t)offset);
      o.strand.push_back(ft.strand);
    }
  }
}

int place_windows(pf_feeder* f, uint32_t n_cells, const uint32_t* genome, const char* cells_blob,
                  uint64_t cells_len, int32_t up, int32_t down, int32_t down_start_codon,
                  std::vector<Piece>& pieces, uint32_t n_threads = 0) {
  f->seq_off.assign(1, 0);
  f->cell.clear(); f->feature.clear();
  f->start.clear(); f->end.clear(); f->offset.clear(); f->strand.clear();
  f->miss_cell.clear(); f->miss_kind.clear(); f->miss_text.clear(); f->miss_off.assign(1, 0);
  pieces.clear();
  // where the cells start: cell i is [cell_at[i], cell_at[i + 1] - 1)
  std::vector<const char*> cell_at(n_cells + 1);
  const char* p = cells_blob;
  const char* e = cells_blob + cells_len;
  for (uint32_t ci = 0; ci < n_cells; ++ci) {
    if (genome[ci] >= f->genomes.size()) { f->err = "pf_feeder_cut: genome index out of range"; return PF_ERR_INVALID; }
    cell_at[ci] = p;
    const char* ce = (const char*)memchr(p, '\n', e - p);
    if (!ce) {
      if (ci + 1 != n_cells) { f->err = "pf_feeder_cut: fewer cells in the blob than n_cells"; return PF_ERR_INVALID; }
      ce = e;
    }
    p = ce + 1;                                     // (one past the end after the last cell: never read)
  }
  cell_at[n_cells] = p;
  // all cores, >= 2,048 cells (~1 ms) per thread; a caller that names a thread count gets it
  uint32_t nt = n_threads ? std::min(n_threads, std::max(1u, n_cells / 8u))
                          : std::min(pf_host_threads(), n_cells / 2048u);
  nt = std::max(1u, nt);
  std::vector<Placed> part(nt);
  if (nt == 1) {
    place_cells(f, 0, n_cells, genome, cell_at, up, down, down_start_codon, part[0]);
  } else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        place_cells(f, (uint32_t)((uint64_t)n_cells * t / nt), (uint32_t)((uint64_t)n_cells * (t + 1) / nt), genome,
                    cell_at, up, down, down_start_codon, part[t]);
      });
    for (auto& x : th) x.join();
  }
  size_t n_seq = 0;
  for (const Placed& o : part) n_seq += o.pieces.size();
  pieces.reserve(n_seq);
  f->seq_off.reserve(n_seq + 1);
  for (const Placed& o : part) {
    pieces.insert(pieces.end(), o.pieces.begin(), o.pieces.end());
    f->cell.insert(f->cell.end(), o.cell.begin(), o.cell.end());
    f->feature.insert(f->feature.end(), o.feature.begin(), o.feature.end());
    f->start.insert(f->start.end(), o.start.begin(), o.start.end());
    f->end.insert(f->end.end(), o.end.begin(), o.end.end());
    f->offset.insert(f->offset.end(), o.offset.begin(), o.offset.end());
    f->strand.insert(f->strand.end(), o.strand.begin(), o.strand.end());
    const uint64_t text0 = f->miss_text.size();
    f->miss_cell.insert(f->miss_cell.end(), o.miss_cell.begin(), o.miss_cell.end());
    f->miss_kind.insert(f->miss_kind.end(), o.miss_kind.begin(), o.miss_kind.end());
    f->miss_text += o.miss_text;
    for (uint64_t end : o.miss_off) f->miss_off.push_back(text0 + end);
  }
  for (const Piece& pc : pieces) f->seq_off.push_back(f->seq_off.back() + (uint64_t)(pc.hi - pc.lo));
  return PF_OK;
}

void fill_cut_result(const pf_feeder* f, pf_cut_result* out) {
  memset(out, 0, sizeof *out);
  out->n_seqs = (uint32_t)f->cell.size();
  out->seq_off = f->seq_off.data();
  out->cell = f->cell.data();
  out->feature = f->feature.data();
  out->start = f->start.data();
  out->end = f->end.data();
  out->offset = f->offset.data();
  out->strand = f->strand.data();
  out->n_missing = (uint32_t)f->miss_cell.size();
  out->missing_cell = f->miss_cell.data();
  out->missing_kind = f->miss_kind.data();
  out->missing_text = f->miss_text.data();
  out->missing_off = f->miss_off.data();
}

// [0, n_seq) split into nt ranges of about equal bases; range t = [cut[t], cut[t + 1])
std::vector<size_t> split_by_bases(const std::vector<uint64_t>& seq_off, uint32_t nt) {
  const size_t n_seq = seq_off.size() - 1;
  const uint64_t total = seq_off.back();
  std::vector<size_t> cut(nt + 1, n_seq);
  cut[0] = 0;
  for (uint32_t t = 1; t < nt; ++t) {
    const uint64_t want = total * t / nt;
    size_t s = (size_t)(std::upper_bound(seq_off.begin(), seq_off.end(), want) - seq_off.begin() - 1);
    cut[t] = std::max(std::min(s, n_seq), cut[t - 1]);
  }
  return cut;
}
}  // namespace

extern "C" int pf_feeder_cut(pf_feeder* f, uint32_t n_cells, const uint32_t* genome, const char* cells_blob,
                             uint64_t cells_len, int32_t up, int32_t down, int32_t down_start_codon,
                             pf_cut_result* out) {
  if (!f || !out || (n_cells && (!genome || !cells_blob))) return PF_ERR_INVALID;
  std::vector<Piece> pieces;
  const int rc = place_windows(f, n_cells, genome, cells_blob, cells_len, up, down, down_start_codon, pieces);
  if (rc != PF_OK) return rc;
  // pass 2 (host threads): copy / reverse-complement the windows into place
  const size_t total = (size_t)f->seq_off.back();
  if (total > f->ascii_cap) {
    f->ascii.reset(new char[total + total / 4 + 64]);
    f->ascii_cap = total + total / 4 + 64;
  }
  {
    char* base = f->ascii.get();
    const size_t n_seq = pieces.size();
    auto fill = [&](size_t s0, size_t s1) {
      std::vector<unsigned char> tmp;
      for (size_t s = s0; s < s1; ++s) {
        const Piece& pc = pieces[s];
        char* dst = base + f->seq_off[s];
        const int64_t len = pc.hi - pc.lo;
        if (len <= 0) continue;
        const unsigned char* src = contig_window(*pc.contig, pc.lo, pc.hi, tmp);
        auto up = [](unsigned char c) { return (unsigned char)(c - (((unsigned)(c - 'a') < 26u) << 5)); };
        if (pc.minus) {
          for (int64_t i = 0; i < len; ++i) dst[i] = (char)kComp.t[up(src[len - 1 - i])];
        } else {
          for (int64_t i = 0; i < len; ++i) dst[i] = (char)up(src[i]);
        }
      }
    };
    uint32_t nt = std::min<uint32_t>(pf_host_threads(), 8u);
    nt = (uint32_t)std::min<size_t>(nt, std::max<size_t>(1, total >> 20));       // >= 1 MiB per thread
    if (nt <= 1 || n_seq < 2 * nt) fill(0, n_seq);
    else {
      const std::vector<size_t> cut = split_by_bases(f->seq_off, nt);             // by bytes, not by sequences
      std::vector<std::thread> th;
      for (uint32_t t = 0; t < nt; ++t) th.emplace_back(fill, cut[t], cut[t + 1]);
      for (auto& x : th) x.join();
    }
  }
  fill_cut_result(f, out);
  out->ascii = f->ascii.get();
  return PF_OK;
}

// The same cut with the sequences packed on the spot: contig windows go straight into the 2-bit
// plane of a pf_batch (and, for sequences holding N / IUPAC symbols, the 4-bit plane) on host
// threads, without the ASCII copy and the second pass of pf_pack_2bit / pf_pack_4bit over it.
// Layout and codes are exactly those of pf_pack_plan / pf_pack_2bit / pf_pack_4bit on the ASCII
// result of pf_feeder_cut (tests/test_feeder_native.py compares the two).
namespace {
struct PackLut {
  uint8_t two_fwd[256], two_rc[256], four_fwd[256], four_rc[256];
  PackLut() {
    uint8_t two[256], four[256];
    memset(two, 255, sizeof two);
    memset(four, 255, sizeof four);
    two[(unsigned char)'A'] = 0; two[(unsigned char)'C'] = 1; two[(unsigned char)'G'] = 2; two[(unsigned char)'T'] = 3;
    const char* amb = "ABCDGHKMNRSTVWXY";                       // the 4-bit alphabet of pf_pack_4bit
    for (int i = 0; i < 16; ++i) four[(unsigned char)amb[i]] = (uint8_t)i;
    for (int c = 0; c < 256; ++c) {
      // any case (contigs used in place keep the file's; the reference upper-cases them).  2-bit
      // tables: 0x80 marks a symbol outside ACGT (its low bits pack as A, like pf_pack_2bit)
      const int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
      two_fwd[c] = two[u] == 255 ? 0x80 : two[u]; four_fwd[c] = four[u];
      two_rc[c] = two[kComp.t[u]] == 255 ? 0x80 : two[kComp.t[u]]; four_rc[c] = four[kComp.t[u]];
    }
  }
};
const PackLut kPack;

// 8 symbols at a time (SWAR): x holds them as bytes, the FIRST one in the top byte.  If all are
// ACGT in either case, *code16 = their 2-bit codes (A 0, C 1, G 2, T 3), first symbol in the top
// bits, and the result is true; otherwise false (the caller takes the table loop).
inline bool pack8(uint64_t x, uint32_t* code16) {
  const uint64_t k01 = 0x0101010101010101ull;
  x &= ~(0x20 * k01);                                             // fold the case bit
  const uint64_t code = ((x >> 1) ^ (x >> 2)) & (3 * k01);        // A 0x41, C 0x43, G 0x47, T 0x54 -> 0, 1, 2, 3
  const uint64_t b0 = code & k01, b1 = (code >> 1) & k01, b01 = b0 & b1;
  // the letter each code stands for: 0x41 + {0, 2, 6, 0x13}; equal to x iff x was that letter
  const uint64_t letter = 0x41 * k01 + (b0 << 1) + (b1 << 2) + (b1 << 1) + (b01 << 3) + (b01 << 1) + b01;
  if (x != letter) return false;
  uint64_t y = code;
  y = (y | (y >> 6)) & 0x000F000F000F000Full;
  y = (y | (y >> 12)) & 0x000000FF000000FFull;
  y = (y | (y >> 24)) & 0xFFFFull;
  *code16 = (uint32_t)y;
  return true;
}
}  // namespace

extern "C" int pf_feeder_cut_packed(pf_feeder* f, uint32_t n_cells, const uint32_t* genome, const char* cells_blob,
                                    uint64_t cells_len, int32_t up, int32_t down, int32_t down_start_codon,
                                    uint32_t n_threads, pf_cut_result* out, pf_cut_planes* planes) {
  if (!f || !out || !planes || (n_cells && (!genome || !cells_blob))) return PF_ERR_INVALID;
  std::vector<Piece> pieces;
  const int rc = place_windows(f, n_cells, genome, cells_blob, cells_len, up, down, down_start_codon, pieces, n_threads);
  if (rc != PF_OK) return rc;
  const size_t n_seq = pieces.size();
  // pf_pack_plan: every sequence starts on a 64-base boundary
  f->base_off.resize(n_seq);
  f->amb_off.assign(n_seq, 0);
  f->is_amb.assign(n_seq, 0);
  uint64_t pos = 0;
  for (size_t s = 0; s < n_seq; ++s) {
    f->base_off[s] = pos;
    pos += (uint64_t)(pieces[s].hi - pieces[s].lo + 63) / 64 * 64;
  }
  const size_t n_words = (size_t)(pos / 32);
  if (n_words > f->packed_cap) {
    f->packed_cap = n_words + n_words / 4 + 64;
    f->packed.reset(new uint64_t[f->packed_cap]);
  }
  uint32_t nt = n_threads ? n_threads : pf_host_threads();
  if (!n_threads) nt = (uint32_t)std::min<size_t>(nt, std::max<size_t>(1, (size_t)f->seq_off.back() >> 18));    // >= 256 k bases per thread
  nt = (uint32_t)std::max<size_t>(1, std::min<size_t>(nt, n_seq));
  const std::vector<size_t> cut = split_by_bases(f->seq_off, nt);
  auto run = [&](auto&& fn) {
    if (nt == 1) { fn(0u); return; }
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nt; ++t) th.emplace_back(fn, t);
    for (auto& x : th) x.join();
  };
  // pf_pack_2bit: padding and non-ACGT symbols pack as A; a sequence with any of the latter is flagged
  run([&](uint32_t t) {
    std::vector<unsigned char> tmp;
    for (size_t s = cut[t]; s < cut[t + 1]; ++s) {
      const Piece& pc = pieces[s];
      const uint64_t len = (uint64_t)(pc.hi - pc.lo);
      const unsigned char* src = len ? contig_window(*pc.contig, pc.lo, pc.hi, tmp) : nullptr;
      uint64_t* dst = f->packed.get() + f->base_off[s] / 32;
      const uint64_t words = (len + 63) / 64 * 2;
      uint32_t bad = 0;
      for (uint64_t w = 0; w < words; ++w) {
        const uint64_t p0 = w * 32;
        const uint32_t m = (uint32_t)std::min<uint64_t>(32, len > p0 ? len - p0 : 0);
        uint64_t v = 0;
        if (m == 32) {                                            // a full word: four groups of 8 symbols
          bool ok = true;
          for (uint32_t gq = 0; gq < 4 && ok; ++gq) {
            uint64_t x;
            uint32_t c16;
            if (pc.minus) {                                       // the 8 symbols END at q: the last byte loaded is the first symbol
              memcpy(&x, src + len - p0 - 8 * gq - 8, 8);
              ok = pack8(x, &c16);
              c16 ^= 0xFFFFu;                                     // complement: A <-> T, C <-> G
            } else {
              memcpy(&x, src + p0 + 8 * gq, 8);
              ok = pack8(__builtin_bswap64(x), &c16);
            }
            v = (v << 16) | c16;
          }
          if (ok) { dst[w] = v; continue; }
          v = 0;
        }
        if (pc.minus) {
          const unsigned char* q = src + len - 1 - p0;            // symbol j of the word: complement of q[-j]
          for (uint32_t j = 0; j < m; ++j) { const uint32_t c = kPack.two_rc[*(q - j)]; bad |= c; v = (v << 2) | (c & 3u); }
        } else {
          const unsigned char* q = src + p0;
          for (uint32_t j = 0; j < m; ++j) { const uint32_t c = kPack.two_fwd[q[j]]; bad |= c; v = (v << 2) | (c & 3u); }
        }
        dst[w] = m ? v << (2 * (32 - m)) : 0;
      }
      f->is_amb[s] = (bad & 0x80u) ? 1 : 0;
    }
  });
  // pf_pack_4bit for the flagged sequences (rare: N / IUPAC symbols)
  uint64_t apos = 0;
  std::vector<uint32_t> amb_seqs;
  for (size_t s = 0; s < n_seq; ++s)
    if (f->is_amb[s]) {
      f->amb_off[s] = apos;
      apos += (uint64_t)(pieces[s].hi - pieces[s].lo + 63) / 64 * 64;
      amb_seqs.push_back((uint32_t)s);
    }
  const size_t n_amb_words = (size_t)(apos / 16);
  int bad_symbol = 0;
  if (n_amb_words) {
    if (n_amb_words > f->amb_cap) {
      f->amb_cap = n_amb_words + n_amb_words / 4 + 64;
      f->amb_plane.reset(new uint64_t[f->amb_cap]);
    }
    const uint64_t a_code = kPack.four_fwd[(unsigned char)'A'];
    std::vector<unsigned char> tmp;
    for (uint32_t s : amb_seqs) {
      const Piece& pc = pieces[s];
      const uint64_t len = (uint64_t)(pc.hi - pc.lo), padded = (len + 63) / 64 * 64;
      const unsigned char* src = contig_window(*pc.contig, pc.lo, pc.hi, tmp);
      uint64_t* dst = f->amb_plane.get() + f->amb_off[s] / 16;
      for (uint64_t w = 0; w < padded / 16; ++w) {
        uint64_t v = 0;
        for (uint32_t j = 0; j < 16; ++j) {
          const uint64_t p = w * 16 + j;
          uint64_t c = a_code;                                   // padding packs as A
          if (p < len) {
            const unsigned char sym = pc.minus ? src[len - 1 - p] : src[p];
            c = pc.minus ? kPack.four_rc[sym] : kPack.four_fwd[sym];
            if (c == 255u) {
              if (!bad_symbol) {                                 // as the reference would see it: upper case, complemented
                const unsigned char u = (unsigned char)(sym - (((unsigned)(sym - 'a') < 26u) << 5));
                bad_symbol = pc.minus ? kComp.t[u] : u;
              }
              c = a_code;
            }
          }
          v = (v << 4) | c;
        }
        dst[w] = v;
      }
    }
  }
  fill_cut_result(f, out);
  out->ascii = nullptr;
  memset(planes, 0, sizeof *planes);
  planes->packed = f->packed.get();
  planes->n_words = n_words;
  planes->base_off = f->base_off.data();
  planes->is_amb = f->is_amb.data();
  planes->amb_plane = n_amb_words ? f->amb_plane.get() : nullptr;
  planes->n_amb_words = n_amb_words;
  planes->amb_off = f->amb_off.data();
  planes->bad_symbol = bad_symbol;
  if (bad_symbol) {
    f->err = std::string("unsupported sequence symbol '") + (char)bad_symbol + "'";
    return PF_ERR_UNSUPPORTED;
  }
  return PF_OK;
}

// ---------------------------------------------------------------------------
// The pangenome table (panaroo's gene_presence_absence.csv): what the reference reads with
// pd.read_csv(path, sep=",", index_col=0, low_memory=False).drop(columns=[...])
// (input.py:198-201) and walks with iterrows().  pandas spends ~0.5 us per cell on a wide table
// and holds every cell as a Python object (BASELINE config #5 is 4e8 cells); here the file is
// mapped, lines and fields are located on host threads and a cell is an (offset, length) into
// the mapping.  Semantics kept: RFC-4180 quoting (delimiters and line ends inside quotes),
// blank lines skipped, short rows padded with missing cells, pandas' strings for a missing
// value (handed over by the caller) and the empty cell mean "absent".
// ---------------------------------------------------------------------------
struct pf_table {
  std::string err;
  FileText text;
  std::vector<std::string> columns, rows;            // kept column names, row labels
  std::string col_blob, row_blob;                    // the same, back to back
  std::vector<uint64_t> col_off, row_off;
  std::vector<uint64_t> cell_off;                    // [n_rows * n_cols] into text
  std::vector<uint32_t> cell_len;                    // 0: absent
};

namespace {
// end of the record that starts at p (the '\n' that ends it, or e); quotes may hold line ends
const char* csv_record_end(const char* p, const char* e) {
  const char* nl = (const char*)memchr(p, '\n', e - p);
  if (!nl) nl = e;
  if (!memchr(p, '"', nl - p)) return nl;
  bool in_q = false;
  for (const char* q = p; q < e; ++q) {
    if (*q == '"') in_q = !in_q;                     // ("" inside quotes toggles twice)
    else if (*q == '\n' && !in_q) return q;
  }
  return e;
}

// the fields of one record: fn(index, begin, end, plain) - [begin, end) is the field's text with
// the enclosing quotes removed; plain = false if it still holds "" escapes or text after the
// closing quote (the caller decides whether it can live with that)
template <typename F>
uint32_t csv_fields(const char* p, const char* e, F&& fn) {
  uint32_t n = 0;
  for (;;) {
    if (p < e && *p == '"') {
      const char* q = p + 1;
      bool plain = true;
      const char* close = nullptr;
      while (q < e) {
        const char* c = (const char*)memchr(q, '"', e - q);
        if (!c) break;
        if (c + 1 < e && c[1] == '"') { plain = false; q = c + 2; continue; }
        close = c;
        break;
      }
      const char* fe = close ? close : e;
      const char* next = close ? close + 1 : e;
      const char* comma = next < e ? (const char*)memchr(next, ',', e - next) : nullptr;
      if (close && (comma ? comma : e) != next) plain = false;            // text after the closing quote
      fn(n++, p + 1, fe, plain);
      if (!comma) return n;
      p = comma + 1;
    } else {
      const char* comma = (const char*)memchr(p, ',', e - p);
      fn(n++, p, comma ? comma : e, true);
      if (!comma) return n;
      p = comma + 1;
    }
  }
}
}  // namespace

extern "C" int pf_table_create(pf_table** out) {
  if (!out) return PF_ERR_INVALID;
  *out = new pf_table();
  return PF_OK;
}
extern "C" void pf_table_destroy(pf_table* t) { delete t; }
extern "C" const char* pf_table_last_error(const pf_table* t) { return t ? t->err.c_str() : "null table"; }

extern "C" int pf_table_load(pf_table* t, const char* path, const char* const* drop_columns, uint32_t n_drop,
                             const char* const* na_values, uint32_t n_na, uint32_t n_threads) {
  if (!t || !path || (n_drop && !drop_columns) || (n_na && !na_values)) return PF_ERR_INVALID;
  if (!read_text(path, &t->text)) { t->err = std::string("cannot read ") + path; return PF_ERR_INVALID; }
  const char* p = t->text.data();
  const char* e = p + t->text.size();
  if (p == e) { t->err = "empty table"; return PF_ERR_INVALID; }
  // header: the first field names the index, the dropped names must exist (pandas raises KeyError)
  const char* he = csv_record_end(p, e);
  std::vector<std::string> header;
  csv_fields(p, he, [&](uint32_t, const char* a, const char* b, bool) { header.emplace_back(a, b); });
  if (!header.empty() && !header.back().empty() && header.back().back() == '\r') header.back().pop_back();
  const uint32_t n_fields = (uint32_t)header.size();
  std::vector<int32_t> kept_of(n_fields, -1);          // field -> kept column, -1 dropped / index
  std::vector<bool> dropped(n_fields, false);
  for (uint32_t d = 0; d < n_drop; ++d) {
    bool found = false;
    for (uint32_t c = 1; c < n_fields; ++c)
      if (header[c] == drop_columns[d]) { dropped[c] = true; found = true; }
    if (!found) { t->err = std::string("column not found: ") + drop_columns[d]; return PF_ERR_INVALID; }
  }
  t->columns.clear();
  for (uint32_t c = 1; c < n_fields; ++c)
    if (!dropped[c]) { kept_of[c] = (int32_t)t->columns.size(); t->columns.push_back(header[c]); }
  const uint32_t S = (uint32_t)t->columns.size();
  // records (blank lines are skipped)
  std::vector<const char*> rec;                          // starts; rec_end[i] by csv_record_end again in the workers
  std::vector<const char*> rec_end;
  for (const char* q = he < e ? he + 1 : e; q < e;) {
    const char* re = csv_record_end(q, e);
    if (re > q) { rec.push_back(q); rec_end.push_back(re); }
    q = re < e ? re + 1 : e;
  }
  const uint64_t R = rec.size();
  std::unordered_map<std::string, int> na;
  size_t na_max = 0;
  for (uint32_t i = 0; i < n_na; ++i) { na.emplace(na_values[i], 1); na_max = std::max(na_max, strlen(na_values[i])); }
  t->cell_off.assign((size_t)R * S, 0);
  t->cell_len.assign((size_t)R * S, 0);
  t->rows.assign(R, std::string());
  uint32_t nt = n_threads ? n_threads : pf_host_threads();
  nt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(nt, R / 64 + 1));
  std::vector<std::string> errs(nt);
  const char* base = t->text.data();
  auto work = [&](uint32_t w) {
    const uint64_t r0 = R * w / nt, r1 = R * (w + 1) / nt;
    for (uint64_t r = r0; r < r1 && errs[w].empty(); ++r) {
      uint64_t* off = t->cell_off.data() + (size_t)r * S;
      uint32_t* len = t->cell_len.data() + (size_t)r * S;
      const uint32_t got = csv_fields(rec[r], rec_end[r], [&](uint32_t f, const char* a, const char* b, bool plain) {
        if (f == 0) {
          if (!plain) errs[w] = "escaped quotes in a row label (record " + std::to_string(r + 2) + ")";
          t->rows[r].assign(a, b);
          return;
        }
        if (f >= n_fields || kept_of[f] < 0) return;
        if (!plain) { errs[w] = "escaped quotes in a cell of record " + std::to_string(r + 2); return; }
        const size_t n = (size_t)(b - a);
        if (n == 0 || (n <= na_max && na.count(std::string(a, b)))) return;          // absent
        if (n >= (1ull << 32)) { errs[w] = "cell too long"; return; }
        off[kept_of[f]] = (uint64_t)(a - base);
        len[kept_of[f]] = (uint32_t)n;
      });
      if (got > n_fields)
        errs[w] = "Expected " + std::to_string(n_fields) + " fields in record " + std::to_string(r + 2) + ", saw " +
                  std::to_string(got);
    }
  };
  if (nt == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (uint32_t w = 0; w < nt; ++w) th.emplace_back(work, w);
    for (auto& x : th) x.join();
  }
  for (const std::string& m : errs)
    if (!m.empty()) { t->err = m; return PF_ERR_INVALID; }
  auto blob = [](const std::vector<std::string>& v, std::string* b, std::vector<uint64_t>* off) {
    b->clear();
    off->assign(1, 0);
    for (const std::string& s : v) { *b += s; off->push_back(b->size()); }
  };
  blob(t->columns, &t->col_blob, &t->col_off);
  blob(t->rows, &t->row_blob, &t->row_off);
  return PF_OK;
}

extern "C" int pf_table_shape(const pf_table* t, uint64_t* n_rows, uint32_t* n_cols) {
  if (!t) return PF_ERR_INVALID;
  if (n_rows) *n_rows = t->rows.size();
  if (n_cols) *n_cols = (uint32_t)t->columns.size();
  return PF_OK;
}

extern "C" int pf_table_names(const pf_table* t, int row_labels, const char** blob, const uint64_t** off) {
  if (!t || !blob || !off) return PF_ERR_INVALID;
  *blob = row_labels ? t->row_blob.data() : t->col_blob.data();
  *off = row_labels ? t->row_off.data() : t->col_off.data();
  return PF_OK;
}

extern "C" int pf_table_row_counts(const pf_table* t, uint32_t* n_present) {
  if (!t || !n_present) return PF_ERR_INVALID;
  const size_t S = t->columns.size();
  for (size_t r = 0; r < t->rows.size(); ++r) {
    uint32_t n = 0;
    const uint32_t* len = t->cell_len.data() + r * S;
    for (size_t c = 0; c < S; ++c) n += len[c] != 0;
    n_present[r] = n;
  }
  return PF_OK;
}

// The cells of rows[0 .. n_sel) with the columns in the order col_order (col_order[j] = table
// column shown at position j; NULL = table order): present[i * n_cols + j] = 1 where row i has a
// cell there, and those cells, row by row in that order, joined with '\n' - the cells_blob of
// pf_feeder_cut.  blob NULL: sizing (blob_len, n_cells).
extern "C" int pf_table_cells(const pf_table* t, const uint64_t* rows, uint64_t n_sel, const uint32_t* col_order,
                              uint8_t* present, char* blob, uint64_t blob_cap, uint64_t* blob_len, uint64_t* n_cells) {
  if (!t || (n_sel && !rows) || !blob_len) return PF_ERR_INVALID;
  const size_t S = t->columns.size();
  const char* base = t->text.data();
  uint64_t bytes = 0, cells = 0;
  for (uint64_t i = 0; i < n_sel; ++i) {
    if (rows[i] >= t->rows.size()) return PF_ERR_INVALID;
    const uint32_t* len = t->cell_len.data() + (size_t)rows[i] * S;
    for (size_t j = 0; j < S; ++j) {
      const uint32_t l = len[col_order ? col_order[j] : j];
      if (l) { bytes += l; ++cells; }
    }
  }
  *blob_len = bytes + (cells ? cells - 1 : 0);
  if (n_cells) *n_cells = cells;
  if (!blob && !present) return PF_OK;
  if (blob && blob_cap < *blob_len) return PF_ERR_NOMEM;
  char* p = blob;
  bool first = true;
  for (uint64_t i = 0; i < n_sel; ++i) {
    const uint64_t* off = t->cell_off.data() + (size_t)rows[i] * S;
    const uint32_t* len = t->cell_len.data() + (size_t)rows[i] * S;
    for (size_t j = 0; j < S; ++j) {
      const size_t c = col_order ? col_order[j] : j;
      if (col_order && c >= S) return PF_ERR_INVALID;
      if (present) present[i * S + j] = len[c] != 0;
      if (len[c] && blob) {
        if (!first) *p++ = '\n';
        first = false;
        memcpy(p, base + off[c], len[c]);
        p += len[c];
      }
    }
  }
  return PF_OK;
}
