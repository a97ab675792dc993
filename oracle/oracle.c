/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into libpanfeed_b200.so.
 *
 * Plain-C restatement of the reference's per-gene-cluster k-mer streaming
 * path, the fast twin of oracle/ref_port.py for shapes the Python reference
 * cannot finish.  It works on ASCII bases, not on the
 * product's 2-bit planes, so it also checks the packer.
 *
 *   kmer_stage()    follows /root/reference/panfeed/panfeed.py:45-88
 *                   (dict k-mer -> presence vector, first-seen order; string
 *                   compare for the canonical choice, :69-75)
 *   positional rows follow panfeed.py:90-107
 *   pattern_stage() follows panfeed.py:175-223 (cluster row, MAF in float64
 *                   :190-200, same-as-cluster filter :202-204, global dedup
 *                   :210-212).  The reference dedups on md5(vector bytes); here
 *                   the key is the vector itself (identical unless md5 collides).
 *   complement      pyfaidx's table (third-party, un-vendored, unpinned:
 *                   pyproject.toml:29), input.py:448-452
 *
 * Parity pin: tests/test_oracle_c.py checks it against ref_port.py, which is
 * itself pinned byte-for-byte to the unmodified reference's outputs.
 *
 * Presence vectors are {0,1,NaN}; they are held as two bit planes of
 * W = ceil(S/32) words: `bits` (1) and the cluster's presence (not NaN).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct or_seq {
  uint64_t off;      /* first base in `bases` */
  uint32_t len;
  uint32_t cluster;  /* cluster index, non-decreasing */
  uint32_t sample;   /* rank in sorted(strains) */
  uint32_t flags;    /* bit0: strain in --targets */
  int32_t start, end, offset, strand;
} or_seq;

typedef struct or_params {
  uint32_t k, n_samples, canonical, consider_missing, cluster_equal_filter;
  uint32_t n_threads;
  double maf;
} or_params;

typedef struct or_result {
  uint64_t n_rows;
  uint32_t* row_cluster;
  char* row_kmer;          /* n_rows x k bytes */
  uint32_t* row_count;
  uint32_t* row_pattern;   /* index into kmer patterns */
  uint32_t n_clusters;
  uint32_t* cluster_pattern;
  uint64_t n_kmer_patterns;
  uint32_t* kmer_pattern_bits;     /* n x W */
  uint32_t* kmer_pattern_cluster;  /* n: cluster-pattern index giving the NaN plane, or ~0 */
  uint64_t n_cluster_patterns;
  uint32_t* cluster_pattern_bits;  /* n x W */
  uint64_t n_pos;
  uint32_t* pos_seq;
  uint32_t* pos_pos;
  int32_t* pos_used_strand;  /* canonical: +1/-1 ; else feature strand (row 2 uses the negation) */
  char* pos_kmer;            /* n_pos x k: canonical k-mer, or the forward one */
  uint64_t n_unique;         /* U: dict entries over all clusters */
  uint64_t n_instances;      /* M */
} or_result;

static unsigned char COMP[256];
static pthread_once_t comp_once = PTHREAD_ONCE_INIT;
static void comp_init(void) {
  const char* a = "ACTGNactgnYRWSKMDVHBXyrwskmdvhbx";
  const char* b = "TGACNtgacnRYWSMKHBDVXrywsmkhbdvx";
  for (int i = 0; i < 256; ++i) COMP[i] = (unsigned char)i;
  for (int i = 0; a[i]; ++i) COMP[(unsigned char)a[i]] = (unsigned char)b[i];
}

static inline uint64_t hash_bytes(const void* p, size_t n) {
  const unsigned char* s = (const unsigned char*)p;
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
  h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
  return h;
}

/* ---- per-cluster dict: k-mer string -> presence bits, insertion ordered ---- */
typedef struct kdict {
  uint32_t k, W;
  uint32_t cap, n;       /* table capacity (pow2), entries */
  int32_t* slots;        /* -1 empty else entry index */
  uint32_t ecap;
  size_t* key;           /* offset of the k-mer bytes in the arena */
  uint32_t* bits;        /* n x W */
  char* arena; size_t arena_n, arena_cap;
} kdict;

static void kd_init(kdict* d, uint32_t k, uint32_t W) {
  memset(d, 0, sizeof *d);
  d->k = k; d->W = W; d->cap = 1024;
  d->slots = (int32_t*)malloc(sizeof(int32_t) * d->cap);
  memset(d->slots, 0xff, sizeof(int32_t) * d->cap);
  d->ecap = 512;
  d->key = (size_t*)malloc(sizeof(size_t) * d->ecap);
  d->bits = (uint32_t*)calloc((size_t)d->ecap * W, 4);
}
static void kd_free(kdict* d) {
  free(d->slots); free(d->key); free(d->bits); free(d->arena);
}
static void kd_grow(kdict* d) {
  uint32_t ncap = d->cap * 2;
  int32_t* ns = (int32_t*)malloc(sizeof(int32_t) * ncap);
  memset(ns, 0xff, sizeof(int32_t) * ncap);
  for (uint32_t e = 0; e < d->n; ++e) {
    uint64_t h = hash_bytes(d->arena + d->key[e], d->k) & (ncap - 1);
    while (ns[h] >= 0) h = (h + 1) & (ncap - 1);
    ns[h] = (int32_t)e;
  }
  free(d->slots); d->slots = ns; d->cap = ncap;
}
/* returns the entry index of `kmer`, inserting it (first-seen order) if new */
static uint32_t kd_get(kdict* d, const char* kmer) {
  uint64_t h = hash_bytes(kmer, d->k) & (d->cap - 1);
  while (d->slots[h] >= 0) {
    uint32_t e = (uint32_t)d->slots[h];
    if (memcmp(d->arena + d->key[e], kmer, d->k) == 0) return e;
    h = (h + 1) & (d->cap - 1);
  }
  if (d->n == d->ecap) {
    uint32_t ne = d->ecap * 2;
    d->key = (size_t*)realloc(d->key, sizeof(size_t) * ne);
    d->bits = (uint32_t*)realloc(d->bits, (size_t)ne * d->W * 4);
    memset(d->bits + (size_t)d->ecap * d->W, 0, (size_t)(ne - d->ecap) * d->W * 4);
    d->ecap = ne;
  }
  uint32_t e = d->n++;
  if (d->arena_n + d->k > d->arena_cap) {
    d->arena_cap = d->arena_cap ? d->arena_cap * 2 : 1 << 16;
    while (d->arena_n + d->k > d->arena_cap) d->arena_cap *= 2;
    d->arena = (char*)realloc(d->arena, d->arena_cap);
  }
  memcpy(d->arena + d->arena_n, kmer, d->k);
  d->key[e] = d->arena_n;
  d->arena_n += d->k;
  d->slots[h] = (int32_t)e;
  if (d->n * 2 > d->cap) kd_grow(d);
  return e;
}
static inline const char* kd_key(const kdict* d, uint32_t e) {
  return d->arena + d->key[e];
}

/* ---- per-cluster result of stage 1 ---- */
typedef struct cl_out {
  uint32_t n_rows;
  char* kmers;       /* n_rows x k */
  uint32_t* counts;
  uint32_t* bits;    /* n_rows x W */
  uint32_t n_unique;
  uint64_t n_inst;
  uint64_t n_pos;
  uint32_t* pos_seq; uint32_t* pos_pos; int32_t* pos_us; char* pos_kmer;
} cl_out;

typedef struct job {
  const char* bases; const or_seq* seqs; uint32_t n_seqs; uint32_t n_clusters;
  const uint8_t* presab; const or_params* p;
  uint32_t* first_seq;   /* n_clusters + 1 */
  cl_out* out;
  uint32_t next;         /* atomic cluster ticket */
} job;

static inline uint32_t popc(uint32_t x) { return (uint32_t)__builtin_popcount(x); }

static void do_cluster(job* j, uint32_t c) {
  const or_params* p = j->p;
  const uint32_t k = p->k, S = p->n_samples, W = (S + 31) / 32;
  cl_out* o = &j->out[c];
  memset(o, 0, sizeof *o);
  kdict d; kd_init(&d, k, W);
  char* rev = (char*)malloc(k);
  size_t pos_cap = 0;
  for (uint32_t si = j->first_seq[c]; si < j->first_seq[c + 1]; ++si) {
    const or_seq* s = &j->seqs[si];
    if (s->len < k) continue;
    const char* seq = j->bases + s->off;
    const uint32_t nk = s->len - k + 1;
    const int target = (int)(s->flags & 1u);
    if (target && o->n_pos + nk > pos_cap) {
      pos_cap = (o->n_pos + nk) * 2;
      o->pos_seq = (uint32_t*)realloc(o->pos_seq, pos_cap * 4);
      o->pos_pos = (uint32_t*)realloc(o->pos_pos, pos_cap * 4);
      o->pos_us = (int32_t*)realloc(o->pos_us, pos_cap * 4);
      o->pos_kmer = (char*)realloc(o->pos_kmer, pos_cap * k);
    }
    for (uint32_t pos = 0; pos < nk; ++pos) {
      const char* fwd = seq + pos;
      for (uint32_t i = 0; i < k; ++i)
        rev[i] = (char)COMP[(unsigned char)fwd[k - 1 - i]];
      const char* chosen = fwd; int32_t used;
      if (p->canonical) {
        if (memcmp(fwd, rev, k) <= 0) { chosen = fwd; used = 1; }
        else { chosen = rev; used = -1; }
        uint32_t e = kd_get(&d, chosen);
        d.bits[(size_t)e * W + (s->sample >> 5)] |= 1u << (s->sample & 31);
      } else {
        used = s->strand;
        uint32_t e = kd_get(&d, fwd);
        d.bits[(size_t)e * W + (s->sample >> 5)] |= 1u << (s->sample & 31);
        e = kd_get(&d, rev);
        d.bits[(size_t)e * W + (s->sample >> 5)] |= 1u << (s->sample & 31);
      }
      o->n_inst += p->canonical ? 1 : 2;
      if (target) {
        o->pos_seq[o->n_pos] = si; o->pos_pos[o->n_pos] = pos;
        o->pos_us[o->n_pos] = used;
        memcpy(o->pos_kmer + (size_t)o->n_pos * k, chosen, k);
        o->n_pos++;
      }
    }
  }
  free(rev);
  /* pattern-stage filters that depend on the row alone (panfeed.py:190-204) */
  const uint8_t* pa = j->presab + (size_t)c * S;
  uint32_t n_present = 0;
  uint32_t* cbits = (uint32_t*)calloc(W, 4);
  for (uint32_t s = 0; s < S; ++s) if (pa[s]) { n_present++; cbits[s >> 5] |= 1u << (s & 31); }
  o->n_unique = d.n;
  o->kmers = (char*)malloc((size_t)(d.n ? d.n : 1) * k);
  o->counts = (uint32_t*)malloc((size_t)(d.n ? d.n : 1) * 4);
  o->bits = (uint32_t*)malloc((size_t)(d.n ? d.n : 1) * W * 4);
  for (uint32_t e = 0; e < d.n; ++e) {
    const uint32_t* b = d.bits + (size_t)e * W;
    uint32_t cnt = 0;
    for (uint32_t w = 0; w < W; ++w) cnt += popc(b[w]);
    double denom = p->consider_missing ? (double)n_present : (double)S;
    double af = (double)cnt / denom;      /* vec.sum() / shape[0] */
    if (af >= 0.5) af = 1 - af;
    if (af < p->maf) continue;
    if (p->cluster_equal_filter) {
      int equal = 1;
      if (p->consider_missing && n_present != S) equal = 0;   /* NaN != anything */
      else for (uint32_t w = 0; w < W; ++w) if (b[w] != cbits[w]) { equal = 0; break; }
      if (equal) continue;
    }
    memcpy(o->kmers + (size_t)o->n_rows * k, kd_key(&d, e), k);
    o->counts[o->n_rows] = cnt;
    memcpy(o->bits + (size_t)o->n_rows * W, b, (size_t)W * 4);
    o->n_rows++;
  }
  free(cbits);
  kd_free(&d);
}

static void* worker(void* arg) {
  job* j = (job*)arg;
  for (;;) {
    uint32_t c = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (c >= j->n_clusters) break;
    do_cluster(j, c);
  }
  return NULL;
}

/* ---- global pattern set keyed on the full vector ---- */
typedef struct pset {
  uint32_t words;       /* key length in words */
  uint64_t cap, n, ecap;
  int64_t* slots;
  uint32_t* keys;       /* n x words */
} pset;
static void ps_init(pset* s, uint32_t words) {
  memset(s, 0, sizeof *s); s->words = words; s->cap = 1024; s->ecap = 256;
  s->slots = (int64_t*)malloc(8 * s->cap); memset(s->slots, 0xff, 8 * s->cap);
  s->keys = (uint32_t*)malloc((size_t)s->ecap * words * 4);
}
static uint64_t ps_add(pset* s, const uint32_t* key) {
  uint64_t h = hash_bytes(key, (size_t)s->words * 4) & (s->cap - 1);
  while (s->slots[h] >= 0) {
    if (memcmp(s->keys + (size_t)s->slots[h] * s->words, key, (size_t)s->words * 4) == 0)
      return (uint64_t)s->slots[h];
    h = (h + 1) & (s->cap - 1);
  }
  if (s->n == s->ecap) { s->ecap *= 2; s->keys = (uint32_t*)realloc(s->keys, (size_t)s->ecap * s->words * 4); }
  memcpy(s->keys + (size_t)s->n * s->words, key, (size_t)s->words * 4);
  s->slots[h] = (int64_t)s->n;
  uint64_t id = s->n++;
  if (s->n * 2 > s->cap) {
    uint64_t nc = s->cap * 2; int64_t* ns = (int64_t*)malloc(8 * nc); memset(ns, 0xff, 8 * nc);
    for (uint64_t e = 0; e < s->n; ++e) {
      uint64_t g = hash_bytes(s->keys + (size_t)e * s->words, (size_t)s->words * 4) & (nc - 1);
      while (ns[g] >= 0) g = (g + 1) & (nc - 1);
      ns[g] = (int64_t)e;
    }
    free(s->slots); s->slots = ns; s->cap = nc;
  }
  return id;
}

int oracle_run(const char* bases, const or_seq* seqs, uint32_t n_seqs,
               uint32_t n_clusters, const uint8_t* presab,
               const or_params* p, or_result* out) {
  pthread_once(&comp_once, comp_init);
  memset(out, 0, sizeof *out);
  const uint32_t S = p->n_samples, W = (S + 31) / 32, k = p->k;
  job j; memset(&j, 0, sizeof j);
  j.bases = bases; j.seqs = seqs; j.n_seqs = n_seqs; j.n_clusters = n_clusters;
  j.presab = presab; j.p = p;
  j.first_seq = (uint32_t*)calloc(n_clusters + 1, 4);
  for (uint32_t i = 0; i < n_seqs; ++i) {
    if (seqs[i].cluster >= n_clusters) return -1;
    if (i && seqs[i].cluster < seqs[i - 1].cluster) return -1;
    j.first_seq[seqs[i].cluster + 1]++;
  }
  for (uint32_t c = 0; c < n_clusters; ++c) j.first_seq[c + 1] += j.first_seq[c];
  j.out = (cl_out*)calloc(n_clusters ? n_clusters : 1, sizeof(cl_out));
  uint32_t nt = p->n_threads ? p->n_threads : 1;
  if (nt > 256) nt = 256;
  if (nt == 1) worker(&j);
  else {
    pthread_t th[256];
    for (uint32_t t = 0; t < nt; ++t) pthread_create(&th[t], NULL, worker, &j);
    for (uint32_t t = 0; t < nt; ++t) pthread_join(th[t], NULL);
  }
  /* serial pattern stage in cluster order = the reference's single writer */
  uint64_t n_rows = 0, n_pos = 0;
  for (uint32_t c = 0; c < n_clusters; ++c) { n_rows += j.out[c].n_rows; n_pos += j.out[c].n_pos; }
  out->n_rows = n_rows; out->n_clusters = n_clusters; out->n_pos = n_pos;
  out->row_cluster = (uint32_t*)malloc((n_rows ? n_rows : 1) * 4);
  out->row_kmer = (char*)malloc((n_rows ? n_rows : 1) * (size_t)k);
  out->row_count = (uint32_t*)malloc((n_rows ? n_rows : 1) * 4);
  out->row_pattern = (uint32_t*)malloc((n_rows ? n_rows : 1) * 4);
  out->cluster_pattern = (uint32_t*)malloc((n_clusters ? n_clusters : 1) * 4);
  out->pos_seq = (uint32_t*)malloc((n_pos ? n_pos : 1) * 4);
  out->pos_pos = (uint32_t*)malloc((n_pos ? n_pos : 1) * 4);
  out->pos_used_strand = (int32_t*)malloc((n_pos ? n_pos : 1) * 4);
  out->pos_kmer = (char*)malloc((n_pos ? n_pos : 1) * (size_t)k);
  pset pc, pk; ps_init(&pc, W); ps_init(&pk, W + 1);
  uint32_t* key = (uint32_t*)malloc((size_t)(W + 1) * 4);
  uint64_t r = 0, q = 0;
  for (uint32_t c = 0; c < n_clusters; ++c) {
    cl_out* o = &j.out[c];
    memset(key, 0, (size_t)(W + 1) * 4);
    const uint8_t* pa = presab + (size_t)c * S;
    for (uint32_t s = 0; s < S; ++s) if (pa[s]) key[s >> 5] |= 1u << (s & 31);
    uint32_t cid = (uint32_t)ps_add(&pc, key);
    out->cluster_pattern[c] = cid;
    for (uint32_t i = 0; i < o->n_rows; ++i, ++r) {
      memcpy(key, o->bits + (size_t)i * W, (size_t)W * 4);
      key[W] = p->consider_missing ? cid : 0xffffffffu;
      out->row_cluster[r] = c;
      memcpy(out->row_kmer + r * k, o->kmers + (size_t)i * k, k);
      out->row_count[r] = o->counts[i];
      out->row_pattern[r] = (uint32_t)ps_add(&pk, key);
    }
    for (uint64_t i = 0; i < o->n_pos; ++i, ++q) {
      out->pos_seq[q] = o->pos_seq[i]; out->pos_pos[q] = o->pos_pos[i];
      out->pos_used_strand[q] = o->pos_us[i];
      memcpy(out->pos_kmer + q * k, o->pos_kmer + i * k, k);
    }
    out->n_unique += o->n_unique; out->n_instances += o->n_inst;
    free(o->kmers); free(o->counts); free(o->bits);
    free(o->pos_seq); free(o->pos_pos); free(o->pos_us); free(o->pos_kmer);
  }
  free(key);
  out->n_cluster_patterns = pc.n;
  out->cluster_pattern_bits = pc.keys; free(pc.slots);
  out->n_kmer_patterns = pk.n;
  out->kmer_pattern_bits = (uint32_t*)malloc((pk.n ? pk.n : 1) * (size_t)W * 4);
  out->kmer_pattern_cluster = (uint32_t*)malloc((pk.n ? pk.n : 1) * 4);
  for (uint64_t e = 0; e < pk.n; ++e) {
    memcpy(out->kmer_pattern_bits + e * W, pk.keys + e * (W + 1), (size_t)W * 4);
    out->kmer_pattern_cluster[e] = pk.keys[e * (W + 1) + W];
  }
  free(pk.keys); free(pk.slots);
  free(j.first_seq); free(j.out);
  return 0;
}

void oracle_free(or_result* r) {
  free(r->row_cluster); free(r->row_kmer); free(r->row_count); free(r->row_pattern);
  free(r->cluster_pattern); free(r->kmer_pattern_bits); free(r->kmer_pattern_cluster);
  free(r->cluster_pattern_bits); free(r->pos_seq); free(r->pos_pos);
  free(r->pos_used_strand); free(r->pos_kmer);
  memset(r, 0, sizeof *r);
}
