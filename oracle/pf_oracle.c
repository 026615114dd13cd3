/*
 * pf_oracle.c -- CPU ORACLE (test infrastructure; see pf_oracle.h header comment).
 *
 * Plain-C restatement of the reference's `query` path and of the `build` path needed to
 * create databases.  Pinning: the FxHasher restatement below is checked against a real rustc-hash 2.x
 * build executed in this image (tests/test_hash_pin_cpu.py, via hashbrown bucket order of an
 * FxHashMap<Vec<u8>,_>); PARITY UNPINNED at the bitvec-serde / bincode (file format) boundaries.
 * Citations are file:line relative to the reference root.
 */
#define _GNU_SOURCE
#include "pf_oracle.h"

#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[512];
static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *pfo_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------------------------
 * L0: rustc-hash 2.x FxHasher on a 64-bit target (crate not vendored).
 * Call chain in the reference: BloomFilter::contains (bloom_filter.rs:312) ->
 * HashIter::from (hash_iter.rs:31-45) -> BuildHasher::hash_one -> HashSeed::build_hasher
 * (hasher.rs:15-20: FxHasher::default() then write_usize(seed)) -> <Vec<u8> as Hash>::hash
 * (write_usize(len) length prefix, then write(bytes)) -> finish().
 * ---------------------------------------------------------------------------------------- */
#define FX_K 0xf1357aea2e62a9c5ULL
#define FX_SEED1 0x243f6a8885a308d3ULL
#define FX_SEED2 0x13198a2e03707344ULL
#define FX_PREVENT 0xa4093822299f31d0ULL

static inline uint64_t le64(const uint8_t *p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v; /* host is little-endian x86-64 */
}
static inline uint32_t le32(const uint8_t *p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return v;
}
static inline uint64_t multiply_mix(uint64_t x, uint64_t y) {
    __uint128_t p = (__uint128_t)x * (__uint128_t)y;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
}
static inline uint64_t rotl64(uint64_t x, int r) {
    r &= 63;
    return r ? (x << r) | (x >> (64 - r)) : x;
}

/* rustc-hash 2.x `hash_bytes` (wyhash-inspired byte-string compressor). */
uint64_t pfo_hash_bytes(const uint8_t *b, size_t len) {
    uint64_t s0 = FX_SEED1, s1 = FX_SEED2;
    if (len <= 16) {
        if (len >= 8) {
            s0 ^= le64(b);
            s1 ^= le64(b + len - 8);
        } else if (len >= 4) {
            s0 ^= (uint64_t)le32(b);
            s1 ^= (uint64_t)le32(b + len - 4);
        } else if (len > 0) {
            uint8_t lo = b[0], mid = b[len / 2], hi = b[len - 1];
            s0 ^= (uint64_t)lo;
            s1 ^= ((uint64_t)hi << 8) | (uint64_t)mid;
        }
    } else {
        size_t off = 0;
        while (off < len - 16) {
            uint64_t x = le64(b + off), y = le64(b + off + 8);
            uint64_t t = multiply_mix(s0 ^ x, FX_PREVENT ^ y);
            s0 = s1;
            s1 = t;
            off += 16;
        }
        s0 ^= le64(b + len - 16);
        s1 ^= le64(b + len - 8);
    }
    return multiply_mix(s0, s1) ^ (uint64_t)len;
}

/* FxHasher::add_to_hash: hash = (hash + i) * K (wrapping). */
static inline uint64_t fx_add(uint64_t s, uint64_t x) { return (s + x) * FX_K; }

/* hash_one(&Vec<u8>) through HashSeed (hasher.rs:15-20). */
uint64_t pfo_fx_hash(uint64_t seed, const uint8_t *item, size_t len, int rot) {
    uint64_t s = 0;                            /* FxHasher::default() */
    s = fx_add(s, seed);                       /* hasher.write_usize(self.seed)   hasher.rs:18 */
    s = fx_add(s, (uint64_t)len);              /* <[u8] as Hash>: write_length_prefix -> write_usize(len) */
    s = fx_add(s, pfo_hash_bytes(item, len));  /* Hasher::write(bytes) -> write_u64(hash_bytes(bytes)) */
    return rotl64(s, rot);                     /* finish(): rotate_left(ROTATE) */
}

/* HashIter::next (hash_iter.rs:13-27). */
void pfo_hash_iter(uint64_t h1, uint64_t h2, uint32_t count, uint64_t *out) {
    for (uint32_t i = 0; i < count; i++) {
        if (i == 0) out[i] = h1;
        else if (i == 1) out[i] = h2;
        else out[i] = (h1 + (uint64_t)i) * h2;
    }
}

/* ------------------------------------------------------------------------------------------
 * L4: canonical k-mers on raw ASCII bytes (file_parser.rs:114-148).
 * bio::alphabets::dna::complement table: identity except the IUPAC pairs below, with
 * lower-case variants mapped likewise (case preserved).
 * ---------------------------------------------------------------------------------------- */
/* The table is filled once, at load time (constructor), before any thread can use it; comp_init() stays as a
 * guard for static linking.  Every entry is computed by a pure function and stored exactly once with its
 * final value, so even a concurrent second initialisation could only rewrite identical bytes (round 1 filled
 * the table with the identity first and patched it afterwards: a late thread doing that inside the OpenMP
 * region of query_impl made 'A'->'A' visible to threads already canonicalising). */
static uint8_t g_comp[256];
static volatile int g_comp_init = 0;
static uint8_t comp_of(int v) {
    static const char a[] = "AGCTYRWSKMDVHBN", b[] = "TCGARYWSMKHBDVN";
    for (int i = 0; a[i]; i++) {
        if (v == (uint8_t)a[i]) return (uint8_t)b[i];
        if (v == (uint8_t)a[i] + 32) return (uint8_t)(b[i] + 32);
    }
    return (uint8_t)v;
}
__attribute__((constructor)) static void comp_init(void) {
    if (g_comp_init) return;
    for (int v = 0; v < 256; v++) g_comp[v] = comp_of(v);
    __sync_synchronize();
    g_comp_init = 1;
}
uint8_t pfo_complement(uint8_t b) {
    comp_init();
    return g_comp[b];
}
void pfo_revcomp(const uint8_t *kmer, size_t k, uint8_t *out) {
    comp_init();
    for (size_t i = 0; i < k; i++) out[i] = g_comp[kmer[k - 1 - i]];
}
/* get_lex_less (file_parser.rs:114-121): Less -> forward, Greater -> revcomp, Equal -> forward */
void pfo_get_lex_less(const uint8_t *kmer, size_t k, uint8_t *out) {
    comp_init();
    /* compare forward with reverse complement bytewise without materialising it */
    int c = 0;
    for (size_t i = 0; i < k && c == 0; i++) {
        uint8_t f = kmer[i], r = g_comp[kmer[k - 1 - i]];
        c = (f < r) ? -1 : (f > r) ? 1 : 0;
    }
    if (c <= 0) memcpy(out, kmer, k);
    else pfo_revcomp(kmer, k, out);
}
size_t pfo_num_kmers(size_t seq_len, size_t k) {
    if (k > seq_len || k == 0) return 0; /* file_parser.rs:136-138 */
    return seq_len - k + 1;
}
void pfo_get_kmers(const uint8_t *seq, size_t len, size_t k, uint8_t *out) {
    size_t n = pfo_num_kmers(len, k);
    for (size_t i = 0; i < n; i++) pfo_get_lex_less(seq + i, k, out + i * k);
}

/* ------------------------------------------------------------------------------------------
 * L1: filter geometry in f32 (bloom_filter.rs:342-357).  Rust's f32::ln lowers to logf,
 * f32::round is round-half-away-from-zero (roundf), `as usize`/`as u32` saturate.
 * ---------------------------------------------------------------------------------------- */
#define LN_2_F32 0.693147180559945309417232121458176568f
uint64_t pfo_needed_bits(float fpr, uint32_t num_items) {
    volatile float ln22 = LN_2_F32 * LN_2_F32;
    volatile float inv = 1.0f / fpr;
    volatile float l = logf(inv);
    volatile float q = l / ln22;
    volatile float v = (float)num_items * q;
    float r = roundf(v);
    if (!(r > 0.0f)) return 0;
    if (r >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)r;
}
uint32_t pfo_optimal_num_hashes(uint64_t num_bits, uint32_t num_items) {
    volatile float a = (float)num_bits / (float)num_items;
    volatile float b = a * LN_2_F32;
    float r = roundf(b);
    uint32_t k;
    if (!(r > 0.0f)) k = 0;
    else if (r >= 4294967296.0f) k = UINT32_MAX;
    else k = (uint32_t)r;
    if (k < 2) k = 2;
    if (k > 200) k = 200;
    return k;
}

/* ------------------------------------------------------------------------------------------
 * L1: bloom filter (bloom_filter.rs).  bits: BitVec<usize, Lsb0>: bit idx lives in word
 * idx/64 at position idx%64 counted from the LSB.
 * ---------------------------------------------------------------------------------------- */
pfo_filter *pfo_filter_new(uint64_t m, uint32_t K, uint64_t seed1, uint64_t seed2) {
    pfo_filter *f = (pfo_filter *)calloc(1, sizeof *f);
    if (!f) return NULL;
    f->m = m;
    f->nwords = (m + 63) / 64;
    f->words = (uint64_t *)calloc(f->nwords ? f->nwords : 1, 8);
    f->K = K;
    f->seed1 = seed1;
    f->seed2 = seed2;
    if (!f->words) {
        free(f);
        return NULL;
    }
    return f;
}
void pfo_filter_free(pfo_filter *f) {
    if (!f) return;
    free(f->words);
    free(f);
}
/* insert (bloom_filter.rs:291-307) */
int pfo_filter_insert(pfo_filter *f, const uint8_t *item, size_t len, int rot) {
    uint64_t h1 = pfo_fx_hash(f->seed1, item, len, rot), h2 = pfo_fx_hash(f->seed2, item, len, rot);
    int contained = 1;
    for (uint32_t i = 0; i < f->K; i++) {
        uint64_t g = i == 0 ? h1 : i == 1 ? h2 : (h1 + (uint64_t)i) * h2;
        uint64_t idx = g % f->m;
        contained = (int)((f->words[idx >> 6] >> (idx & 63)) & 1);
        f->words[idx >> 6] |= 1ULL << (idx & 63);
    }
    return !contained;
}
/* contains (bloom_filter.rs:312-332): early exit at the first clear bit */
int pfo_filter_contains(const pfo_filter *f, const uint8_t *item, size_t len, int rot, uint64_t *probes) {
    uint64_t h1 = pfo_fx_hash(f->seed1, item, len, rot), h2 = pfo_fx_hash(f->seed2, item, len, rot);
    for (uint32_t i = 0; i < f->K; i++) {
        uint64_t g = i == 0 ? h1 : i == 1 ? h2 : (h1 + (uint64_t)i) * h2;
        uint64_t idx = g % f->m;
        if (probes) (*probes)++;
        if (!((f->words[idx >> 6] >> (idx & 63)) & 1)) return 0;
    }
    return 1;
}
/* union (bloom_filter.rs:275-278) */
void pfo_filter_union(pfo_filter *dst, const pfo_filter *src) {
    uint64_t n = dst->nwords < src->nwords ? dst->nwords : src->nwords;
    for (uint64_t i = 0; i < n; i++) dst->words[i] |= src->words[i];
}
/* distance (bloom_filter.rs:142-150): Hamming distance over the raw words */
uint64_t pfo_filter_distance(const pfo_filter *a, const pfo_filter *b) {
    uint64_t n = a->nwords < b->nwords ? a->nwords : b->nwords, d = 0;
    for (uint64_t i = 0; i < n; i++) d += (uint64_t)__builtin_popcountll(a->words[i] ^ b->words[i]);
    return d;
}

/* ---- bincode 1.3 (little-endian, fixed-width ints, u64 lengths, Option = u8 tag) ---- */
static int w_u8(FILE *fp, uint8_t v) { return fwrite(&v, 1, 1, fp) == 1 ? 0 : -1; }
static int w_u32(FILE *fp, uint32_t v) { return fwrite(&v, 4, 1, fp) == 1 ? 0 : -1; }
static int w_u64(FILE *fp, uint64_t v) { return fwrite(&v, 8, 1, fp) == 1 ? 0 : -1; }
static int w_f32(FILE *fp, float v) { return fwrite(&v, 4, 1, fp) == 1 ? 0 : -1; }
static int w_str(FILE *fp, const char *s) {
    uint64_t n = strlen(s);
    if (w_u64(fp, n)) return -1;
    return n == 0 || fwrite(s, 1, n, fp) == n ? 0 : -1;
}
static int r_u8(FILE *fp, uint8_t *v) { return fread(v, 1, 1, fp) == 1 ? 0 : -1; }
static int r_u32(FILE *fp, uint32_t *v) { return fread(v, 4, 1, fp) == 1 ? 0 : -1; }
static int r_u64(FILE *fp, uint64_t *v) { return fread(v, 8, 1, fp) == 1 ? 0 : -1; }
static int r_f32(FILE *fp, float *v) { return fread(v, 4, 1, fp) == 1 ? 0 : -1; }
static char *r_str(FILE *fp) {
    uint64_t n;
    if (r_u64(fp, &n) || n > (1u << 20)) return NULL;
    char *s = (char *)malloc(n + 1);
    if (!s) return NULL;
    if (n && fread(s, 1, n, fp) != n) {
        free(s);
        return NULL;
    }
    s[n] = 0;
    return s;
}

#define BITVEC_ORDER "bitvec::order::Lsb0"

/* BloomFilter serialisation (bloom_filter.rs:84-93 through bincode::serialize_into, :176-207):
 * bits (bitvec serde BitSeq{order, head{width,index}, bits, data}), num_hashes u32,
 * hash_builder_one{seed}, hash_builder_two{seed}, file_path Option<PathBuf>; `modified` skipped. */
int pfo_filter_save(const pfo_filter *f, const char *path, const char *recorded_path) {
    FILE *fp = fopen(path, "wb");
    if (!fp) {
        set_err("cannot create %s: %s", path, strerror(errno));
        return -1;
    }
    int e = 0;
    e |= w_str(fp, BITVEC_ORDER);
    e |= w_u8(fp, 64); /* head.width: bits of usize */
    e |= w_u8(fp, 0);  /* head.index */
    e |= w_u64(fp, f->m);
    e |= w_u64(fp, f->nwords);
    if (f->nwords && fwrite(f->words, 8, f->nwords, fp) != f->nwords) e = -1;
    e |= w_u32(fp, f->K);
    e |= w_u64(fp, f->seed1);
    e |= w_u64(fp, f->seed2);
    if (recorded_path) {
        e |= w_u8(fp, 1);
        e |= w_str(fp, recorded_path);
    } else {
        e |= w_u8(fp, 0);
    }
    if (fclose(fp)) e = -1;
    if (e) set_err("write error on %s", path);
    return e ? -1 : 0;
}
/* load_from_file (bloom_filter.rs:153-174); the stored file_path is ignored (:171). */
pfo_filter *pfo_filter_load(const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        set_err("Failed to open Bloom filter file: %s", path);
        return NULL;
    }
    pfo_filter *f = NULL;
    char *order = r_str(fp);
    uint8_t width = 0, index = 0, tag = 0;
    uint64_t m = 0, nwords = 0, s1 = 0, s2 = 0;
    uint32_t K = 0;
    if (!order || strcmp(order, BITVEC_ORDER)) goto bad;
    if (r_u8(fp, &width) || r_u8(fp, &index) || width != 64 || index != 0) goto bad;
    if (r_u64(fp, &m) || r_u64(fp, &nwords) || nwords != (m + 63) / 64) goto bad;
    f = pfo_filter_new(m, 0, 0, 0);
    if (!f) goto bad;
    if (nwords && fread(f->words, 8, nwords, fp) != nwords) goto bad;
    if (r_u32(fp, &K) || r_u64(fp, &s1) || r_u64(fp, &s2) || r_u8(fp, &tag)) goto bad;
    if (tag == 1) {
        char *p = r_str(fp);
        if (!p) goto bad;
        free(p);
    } else if (tag != 0)
        goto bad;
    f->K = K;
    f->seed1 = s1;
    f->seed2 = s2;
    free(order);
    fclose(fp);
    return f;
bad:
    set_err("Failed to deserialize Bloom filter from file: %s", path);
    free(order);
    pfo_filter_free(f);
    fclose(fp);
    return NULL;
}

/* ------------------------------------------------------------------------------------------
 * L2: tree (bloom_tree.rs).  Nodes are boxed recursively in the reference; here they are
 * heap nodes with child pointers.  Filters are keyed by their path string: identical paths
 * share one filter (the reference's cache is keyed the same way, cache.rs:56-77).
 * ---------------------------------------------------------------------------------------- */
typedef struct pfo_node {
    struct pfo_node *left, *right;
    char *bf_path;  /* bloom_filter_path: relative file name "<id>.bf" */
    char *tax_id;   /* Option<String>; NULL = None */
    uint64_t mapped_reads;
    int filter;     /* index into tree->filters */
    /* cache for the kernel-schedule restatement (invalidated whenever the tree changes) */
    int analysed, mono;
    uint64_t pop;
    uint32_t steps;  /* plan of the current pfo_query_batch_sched call */
    uint32_t stride; /* k-mer sampling stride of the plan (1 = every k-mer) */
} pfo_node;

struct pfo_tree {
    pfo_node *root;
    float fpr;
    uint32_t largest_genome;
    uint64_t kmer_size;
    uint64_t seed1, seed2;
    int rot;
    /* filter store keyed by path */
    pfo_filter **filters;
    char **filter_keys;
    int n_filters, cap_filters;
    uint64_t m;
    uint32_t K;
    /* naming of internal nodes */
    int name_mode;
    uint64_t name_state;
    uint64_t name_counter;
    uint8_t *name_used; /* 65536 flags for mode 1 */
    char *directory;
    /* caches rebuilt lazily */
    pfo_node **leaves;
    uint64_t n_leaves;
    pfo_node **pre;
    uint64_t n_pre;
    int dirty;
    /* reference-faithful filter cache (pfo_tree_load_lazy): BFLruCache (cache.rs:13-17, 55-88) -- an LRU of `lru_cap`
     * filters keyed by the node's file name, a miss re-reads and decodes "<db>/<name>.bf" (bloom_filter.rs:153-174) */
    int lru_cap, lru_n;
    pfo_filter **lru_filter;
    const char **lru_key; /* points at the owning node's bf_path */
    uint64_t *lru_stamp, lru_clock;
    uint64_t lru_loads, lru_hits, lru_bytes;
};

static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static int tree_find_filter(const pfo_tree *t, const char *key) {
    for (int i = 0; i < t->n_filters; i++)
        if (!strcmp(t->filter_keys[i], key)) return i;
    return -1;
}
static int tree_add_filter(pfo_tree *t, const char *key, pfo_filter *f) {
    if (t->n_filters == t->cap_filters) {
        int nc = t->cap_filters ? t->cap_filters * 2 : 64;
        t->filters = (pfo_filter **)realloc(t->filters, (size_t)nc * sizeof *t->filters);
        t->filter_keys = (char **)realloc(t->filter_keys, (size_t)nc * sizeof *t->filter_keys);
        t->cap_filters = nc;
    }
    t->filters[t->n_filters] = f;
    t->filter_keys[t->n_filters] = strdup(key);
    return t->n_filters++;
}

pfo_tree *pfo_tree_new(uint64_t kmer_size, float fpr, uint32_t largest_genome, uint64_t seed1,
                       uint64_t seed2, int rot, int name_mode, uint64_t name_seed) {
    pfo_tree *t = (pfo_tree *)calloc(1, sizeof *t);
    if (!t) return NULL;
    t->kmer_size = kmer_size;
    t->fpr = fpr;
    t->largest_genome = largest_genome;
    t->seed1 = seed1;
    t->seed2 = seed2;
    t->rot = rot;
    /* with_rate (bloom_filter.rs:229-240) */
    t->m = pfo_needed_bits(fpr, largest_genome);
    t->K = pfo_optimal_num_hashes(t->m, largest_genome);
    t->name_mode = name_mode;
    t->name_state = name_seed;
    if (name_mode == 1) t->name_used = (uint8_t *)calloc(65536, 1);
    t->dirty = 1;
    return t;
}
static void node_free(pfo_node *n) {
    if (!n) return;
    node_free(n->left);
    node_free(n->right);
    free(n->bf_path);
    free(n->tax_id);
    free(n);
}
void pfo_tree_free(pfo_tree *t) {
    if (!t) return;
    node_free(t->root);
    for (int i = 0; i < t->n_filters; i++) {
        pfo_filter_free(t->filters[i]);
        free(t->filter_keys[i]);
    }
    free(t->filters);
    free(t->filter_keys);
    for (int i = 0; i < t->lru_n; i++) pfo_filter_free(t->lru_filter[i]);
    free(t->lru_filter);
    free(t->lru_key);
    free(t->lru_stamp);
    free(t->name_used);
    free(t->directory);
    free(t->leaves);
    free(t->pre);
    free(t);
}

/* make_bloom_node (bloom_tree.rs:279-299): new empty filter "<nodeid>.bf" + node */
static pfo_node *make_bloom_node(pfo_tree *t, const char *nodeid) {
    pfo_node *n = (pfo_node *)calloc(1, sizeof *n);
    size_t L = strlen(nodeid);
    n->bf_path = (char *)malloc(L + 4);
    memcpy(n->bf_path, nodeid, L);
    memcpy(n->bf_path + L, ".bf", 4);
    n->tax_id = strdup(nodeid);
    pfo_filter *f = pfo_filter_new(t->m, t->K, t->seed1, t->seed2);
    int existing = tree_find_filter(t, n->bf_path);
    if (existing >= 0) {
        /* cache.add_filter replaces the entry under the same key (cache.rs:83-87) */
        pfo_filter_free(t->filters[existing]);
        t->filters[existing] = f;
        n->filter = existing;
    } else {
        n->filter = tree_add_filter(t, n->bf_path, f);
    }
    return n;
}
static int is_leaf(const pfo_node *n) { return !n->left && !n->right; } /* bloom_tree.rs:416-418 */

/* init_internal_node (bloom_tree.rs:226-246) */
static pfo_node *init_internal_node(pfo_tree *t, pfo_node *current, pfo_node *new_node) {
    char name[64];
    if (t->name_mode == 1 && t->name_counter < 65536) {
        uint16_t n2;
        do {
            n2 = (uint16_t)splitmix64(&t->name_state);
        } while (t->name_used[n2]);
        t->name_used[n2] = 1;
        snprintf(name, sizeof name, "Internal_Node_%u", (unsigned)n2);
    } else if (t->name_mode == 1) {
        snprintf(name, sizeof name, "Internal_Node_%llu", (unsigned long long)t->name_counter);
    } else {
        snprintf(name, sizeof name, "Internal_Node_%llu", (unsigned long long)t->name_counter);
    }
    t->name_counter++;
    pfo_node *in = make_bloom_node(t, name);
    pfo_filter_union(t->filters[in->filter], t->filters[new_node->filter]); /* :237 */
    pfo_filter_union(t->filters[in->filter], t->filters[current->filter]);  /* :238 */
    in->left = current;   /* :242 existing on the left */
    in->right = new_node; /* :243 new on the right */
    return in;
}
/* add_to_tree (bloom_tree.rs:187-214) */
static pfo_node *add_to_tree(pfo_tree *t, pfo_node *current, pfo_node *node) {
    if (current->left && current->right) {
        pfo_filter_union(t->filters[current->filter], t->filters[node->filter]); /* :194 */
        uint64_t rd = pfo_filter_distance(t->filters[current->right->filter], t->filters[node->filter]);
        uint64_t ld = pfo_filter_distance(t->filters[current->left->filter], t->filters[node->filter]);
        if (rd < ld) current->right = add_to_tree(t, current->right, node); /* :200-202 */
        else current->left = add_to_tree(t, current->left, node);          /* :203-206 ties go left */
    } else if (!current->left && !current->right) {
        current = init_internal_node(t, current, node); /* :207-208 */
    } else {
        set_err("Node with only one child encountered - should not happen.");
    }
    return current;
}
/* insert (bloom_tree.rs:128-145) + init_leaf_node (:154-170) */
int pfo_tree_insert(pfo_tree *t, const char *id, const uint8_t *seq, size_t len) {
    pfo_node *leaf = make_bloom_node(t, id);
    pfo_filter *f = t->filters[leaf->filter];
    size_t k = (size_t)t->kmer_size, n = pfo_num_kmers(len, k);
    uint8_t *km = (uint8_t *)malloc(k ? k : 1);
    for (size_t i = 0; i < n; i++) {
        pfo_get_lex_less(seq + i, k, km);
        pfo_filter_insert(f, km, k, t->rot);
    }
    free(km);
    if (!t->root) t->root = leaf;
    else t->root = add_to_tree(t, t->root, leaf);
    t->dirty = 1;
    return 0;
}

/* tree.bin: BloomTree through bincode (bloom_tree.rs:28-61, 339-355).  Pre-order recursion. */
static int node_write(FILE *fp, const pfo_node *n) {
    int e = 0;
    if (n->left) {
        e |= w_u8(fp, 1);
        e |= node_write(fp, n->left);
    } else
        e |= w_u8(fp, 0);
    if (n->right) {
        e |= w_u8(fp, 1);
        e |= node_write(fp, n->right);
    } else
        e |= w_u8(fp, 0);
    e |= w_str(fp, n->bf_path);
    if (n->tax_id) {
        e |= w_u8(fp, 1);
        e |= w_str(fp, n->tax_id);
    } else
        e |= w_u8(fp, 0);
    e |= w_u64(fp, n->mapped_reads);
    return e;
}
static void join_path(char *out, size_t cap, const char *dir, const char *name) {
    size_t L = strlen(dir);
    snprintf(out, cap, "%s%s%s", dir, (L && dir[L - 1] == '/') ? "" : "/", name);
}
int pfo_tree_save(const pfo_tree *t, const char *dir) {
    mkdir(dir, 0777);
    char p[4096];
    join_path(p, sizeof p, dir, "tree.bin");
    FILE *fp = fopen(p, "wb");
    if (!fp) {
        set_err("cannot create %s: %s", p, strerror(errno));
        return -1;
    }
    int e = 0;
    if (t->root) {
        e |= w_u8(fp, 1);
        e |= node_write(fp, t->root);
    } else
        e |= w_u8(fp, 0);
    e |= w_f32(fp, t->fpr);
    e |= w_u32(fp, t->largest_genome);
    e |= w_u64(fp, t->kmer_size);
    e |= w_u64(fp, t->seed1);
    e |= w_u64(fp, t->seed2);
    if (fclose(fp)) e = -1;
    if (e) {
        set_err("write error on %s", p);
        return -1;
    }
    /* every filter is written by Drop with file_path = directory.join(name) (bloom_filter.rs:105-117) */
    for (int i = 0; i < t->n_filters; i++) {
        join_path(p, sizeof p, dir, t->filter_keys[i]);
        if (pfo_filter_save(t->filters[i], p, p)) return -1;
    }
    return 0;
}
static pfo_node *node_read(FILE *fp, int depth, int *err) {
    if (depth > 1000000) {
        *err = 1;
        return NULL;
    }
    pfo_node *n = (pfo_node *)calloc(1, sizeof *n);
    uint8_t tag;
    if (r_u8(fp, &tag)) goto bad;
    if (tag == 1) {
        n->left = node_read(fp, depth + 1, err);
        if (*err) goto bad;
    } else if (tag != 0)
        goto bad;
    if (r_u8(fp, &tag)) goto bad;
    if (tag == 1) {
        n->right = node_read(fp, depth + 1, err);
        if (*err) goto bad;
    } else if (tag != 0)
        goto bad;
    n->bf_path = r_str(fp);
    if (!n->bf_path) goto bad;
    if (r_u8(fp, &tag)) goto bad;
    if (tag == 1) {
        n->tax_id = r_str(fp);
        if (!n->tax_id) goto bad;
    } else if (tag != 0)
        goto bad;
    if (r_u64(fp, &n->mapped_reads)) goto bad;
    return n;
bad:
    *err = 1;
    node_free(n);
    return NULL;
}
static int load_filters(pfo_tree *t, pfo_node *n, const char *dir) {
    if (!n) return 0;
    int idx = tree_find_filter(t, n->bf_path);
    if (idx < 0) {
        char p[4096];
        join_path(p, sizeof p, dir, n->bf_path);
        pfo_filter *f = pfo_filter_load(p);
        if (!f) return -1;
        idx = tree_add_filter(t, n->bf_path, f);
    }
    n->filter = idx;
    if (load_filters(t, n->left, dir)) return -1;
    return load_filters(t, n->right, dir);
}
/* get_filter (cache.rs:56-77): hit -> most recently used; miss -> load_from_file, insert, evict the least recently
 * used.  Called from the (serial) recursion only, like the reference's write-locked cache. */
static const pfo_filter *lru_get(pfo_tree *t, const pfo_node *n) {
    t->lru_clock++;
    for (int i = 0; i < t->lru_n; i++)
        if (!strcmp(t->lru_key[i], n->bf_path)) {
            t->lru_stamp[i] = t->lru_clock;
            t->lru_hits++;
            return t->lru_filter[i];
        }
    char p[4096];
    join_path(p, sizeof p, t->directory, n->bf_path);
    pfo_filter *f = pfo_filter_load(p);
    if (!f) return NULL;
    t->lru_loads++;
    t->lru_bytes += (f->m + 63) / 64 * 8;
    int slot = t->lru_n;
    if (t->lru_n == t->lru_cap) {
        slot = 0;
        for (int i = 1; i < t->lru_n; i++)
            if (t->lru_stamp[i] < t->lru_stamp[slot]) slot = i;
        pfo_filter_free(t->lru_filter[slot]);
    } else {
        t->lru_n++;
    }
    t->lru_filter[slot] = f;
    t->lru_key[slot] = n->bf_path;
    t->lru_stamp[slot] = t->lru_clock;
    return f;
}
static const pfo_filter *node_filter(pfo_tree *t, const pfo_node *n) {
    return t->lru_cap ? lru_get(t, n) : t->filters[n->filter];
}

/* BloomTree::load (bloom_tree.rs:364-386).  lru_cap == 0: all filters are loaded eagerly and kept (same bits as the
 * reference's lazy cache, no disk traffic while querying).  lru_cap > 0: reference-faithful -- only tree.bin is read
 * here and filters come through an LRU of that capacity (`--cache-size`, main.rs:119-122). */
static pfo_tree *tree_load_impl(const char *dir, int rot, int lru_cap) {
    char p[4096];
    join_path(p, sizeof p, dir, "tree.bin");
    FILE *fp = fopen(p, "rb");
    if (!fp) {
        set_err("Must provide a directory in where a tree has been stored (%s)", p);
        return NULL;
    }
    pfo_tree *t = (pfo_tree *)calloc(1, sizeof *t);
    t->rot = rot;
    t->dirty = 1;
    uint8_t tag;
    int err = 0;
    if (r_u8(fp, &tag)) err = 1;
    if (!err && tag == 1) t->root = node_read(fp, 0, &err);
    else if (!err && tag != 0) err = 1;
    if (!err && (r_f32(fp, &t->fpr) || r_u32(fp, &t->largest_genome) || r_u64(fp, &t->kmer_size) ||
                 r_u64(fp, &t->seed1) || r_u64(fp, &t->seed2)))
        err = 1;
    fclose(fp);
    if (err) {
        set_err("Failed to deserialize %s", p);
        pfo_tree_free(t);
        return NULL;
    }
    t->directory = strdup(dir);
    if (lru_cap > 0) {
        t->lru_cap = lru_cap;
        t->lru_filter = (pfo_filter **)calloc((size_t)lru_cap, sizeof *t->lru_filter);
        t->lru_key = (const char **)calloc((size_t)lru_cap, sizeof *t->lru_key);
        t->lru_stamp = (uint64_t *)calloc((size_t)lru_cap, sizeof *t->lru_stamp);
        if (t->root) {
            const pfo_filter *f = lru_get(t, t->root);
            if (!f) {
                pfo_tree_free(t);
                return NULL;
            }
            t->m = f->m;
            t->K = f->K;
            t->lru_loads = t->lru_hits = t->lru_bytes = 0;
        }
        return t;
    }
    if (load_filters(t, t->root, dir)) {
        pfo_tree_free(t);
        return NULL;
    }
    if (t->n_filters) {
        t->m = t->filters[0]->m;
        t->K = t->filters[0]->K;
    }
    return t;
}
pfo_tree *pfo_tree_load(const char *dir, int rot) { return tree_load_impl(dir, rot, 0); }
pfo_tree *pfo_tree_load_lazy(const char *dir, int rot, int cache_size) {
    return tree_load_impl(dir, rot, cache_size < 1 ? 1 : cache_size);
}
void pfo_tree_cache_stats(const pfo_tree *t, uint64_t *loads, uint64_t *hits, uint64_t *bytes) {
    *loads = t->lru_loads;
    *hits = t->lru_hits;
    *bytes = t->lru_bytes;
}
/* prune_tree (bloom_tree.rs:302-330): nodes at depth >= search_depth lose their children */
static void prune_rec(pfo_node *n, uint64_t depth, uint64_t search_depth) {
    if (!n) return;
    if (depth < search_depth) {
        prune_rec(n->left, depth + 1, search_depth);
        prune_rec(n->right, depth + 1, search_depth);
    } else {
        node_free(n->left);
        node_free(n->right);
        n->left = n->right = NULL;
    }
}
void pfo_tree_prune(pfo_tree *t, uint64_t search_depth) {
    prune_rec(t->root, 0, search_depth);
    t->dirty = 1;
}

static void collect(pfo_tree *t, pfo_node *n) {
    if (!n) return;
    t->pre[t->n_pre++] = n;
    if (is_leaf(n)) {
        t->leaves[t->n_leaves++] = n;
        return;
    }
    collect(t, n->left);  /* left-first DFS (query.rs:208-216) */
    collect(t, n->right);
}
static uint64_t count_nodes(const pfo_node *n) { return n ? 1 + count_nodes(n->left) + count_nodes(n->right) : 0; }
static void refresh(pfo_tree *t) {
    if (!t->dirty) return;
    uint64_t n = count_nodes(t->root);
    free(t->leaves);
    free(t->pre);
    t->leaves = (pfo_node **)malloc((n ? n : 1) * sizeof *t->leaves);
    t->pre = (pfo_node **)malloc((n ? n : 1) * sizeof *t->pre);
    t->n_leaves = t->n_pre = 0;
    collect(t, t->root);
    for (uint64_t i = 0; i < t->n_pre; i++) t->pre[i]->analysed = 0;
    t->dirty = 0;
}
uint64_t pfo_tree_num_nodes(const pfo_tree *t) {
    refresh((pfo_tree *)t);
    return t->n_pre;
}
uint64_t pfo_tree_num_leaves(const pfo_tree *t) {
    refresh((pfo_tree *)t);
    return t->n_leaves;
}
uint64_t pfo_tree_kmer_size(const pfo_tree *t) { return t->kmer_size; }
uint64_t pfo_tree_num_bits(const pfo_tree *t) { return t->m; }
uint32_t pfo_tree_num_hashes(const pfo_tree *t) { return t->K; }
void pfo_tree_seeds(const pfo_tree *t, uint64_t *s1, uint64_t *s2) {
    *s1 = t->seed1;
    *s2 = t->seed2;
}
const char *pfo_tree_leaf_id(const pfo_tree *t, uint64_t i) {
    refresh((pfo_tree *)t);
    return i < t->n_leaves ? (t->leaves[i]->tax_id ? t->leaves[i]->tax_id : "") : NULL;
}
uint64_t pfo_tree_leaf_count(const pfo_tree *t, uint64_t i) {
    refresh((pfo_tree *)t);
    return i < t->n_leaves ? t->leaves[i]->mapped_reads : 0;
}
void pfo_tree_reset_counts(pfo_tree *t) {
    refresh(t);
    for (uint64_t i = 0; i < t->n_pre; i++) t->pre[i]->mapped_reads = 0;
}
static void depth_rec(const pfo_node *n, uint32_t d, uint8_t *leaf, uint32_t *depth, uint64_t cap, uint64_t *i) {
    if (!n) return;
    if (*i < cap) {
        leaf[*i] = (uint8_t)is_leaf(n);
        depth[*i] = d;
    }
    (*i)++;
    depth_rec(n->left, d + 1, leaf, depth, cap, i);
    depth_rec(n->right, d + 1, leaf, depth, cap, i);
}
uint64_t pfo_tree_preorder(const pfo_tree *t, uint8_t *leaf, uint32_t *depth, uint64_t cap) {
    uint64_t i = 0;
    depth_rec(t->root, 0, leaf, depth, cap, &i);
    return i;
}
const char *pfo_tree_node_name(const pfo_tree *t, uint64_t i) {
    refresh((pfo_tree *)t);
    return i < t->n_pre ? (t->pre[i]->tax_id ? t->pre[i]->tax_id : "") : NULL;
}

/* ------------------------------------------------------------------------------------------
 * L3: query (query.rs:38-158)
 * ---------------------------------------------------------------------------------------- */
/* (threshold * read.kmers.len() as f32).ceil() as usize   (query.rs:48); `as` saturates, NaN -> 0 */
uint64_t pfo_need(float threshold, uint64_t n_kmers) {
    volatile float prod = threshold * (float)n_kmers;
    float c = ceilf(prod);
    if (!(c > 0.0f)) return 0;
    if (c >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)c;
}

typedef struct {
    const uint8_t *kmers; /* n_k * k canonical k-mer bytes (DNASequence.kmers, file_parser.rs:155) */
    uint64_t n_k;
    uint64_t need;
} oread;

/* Probe count of one (read,node) pair under the GPU kernel's deterministic schedule (pf_kernels.cuh,
 * probe_group): need == 0 -> pass, no probes; need > n_k -> fail, no probes.  k-mers are taken in groups of
 * 32*G consecutive positions.  Inside a group the phases are: step 0 of the first 32 k-mers; step 0 of the
 * rest; then step i = 1..n_steps-1 of every k-mer still alive (none alive -> group done).  After every phase
 * the pair fails as soon as total misses > n_k - need; after every group it passes as soon as hits >= need.
 * With n_steps < K a surviving k-mer counts as a hit (sound pre-test of interior nodes).  The outcome at
 * exact nodes (n_steps == K) is identical to query_passes; only the amount of work differs. */
static uint64_t pair_sched(const pfo_filter *f, uint32_t n_steps, uint32_t stride, uint32_t G, const oread *r, size_t k,
                           int rot, int *pass_out) {
    uint64_t probes = 0;
    if (n_steps == 0 || r->need == 0) { /* skipped interior node, or nothing to reach */
        *pass_out = 1;
        return 0;
    }
    if (r->need > r->n_k) {
        *pass_out = 0;
        return 0;
    }
    /* sampled pre-test (stride > 1, verified interior nodes only): slots q = 0..n_s-1 stand for the k-mers
     * off + q*stride, centred in the read; the k-mers that are not probed count as not proven absent */
    uint64_t n_s = r->n_k, off = 0;
    if (stride > 1) {
        n_s = r->n_k / stride ? r->n_k / stride : 1;
        off = (r->n_k - 1 - (n_s - 1) * stride) / 2;
    }
    uint64_t allowed = r->n_k - r->need, misses = 0, hits = r->n_k - n_s;
    const uint64_t gsz = 32ull * G;
    uint64_t *h1 = (uint64_t *)malloc(gsz * 8), *h2 = (uint64_t *)malloc(gsz * 8);
    uint8_t *alive = (uint8_t *)malloc(gsz);
    int decided = 0, result = 0;
    for (uint64_t base = 0; base < n_s && !decided; base += gsz) {
        uint64_t cnt = n_s - base < gsz ? n_s - base : gsz;
        for (uint64_t j = 0; j < cnt; j++) {
            const uint8_t *km = r->kmers + (off + (base + j) * (stride > 1 ? stride : 1)) * k;
            h1[j] = pfo_fx_hash(f->seed1, km, k, rot);
            h2[j] = pfo_fx_hash(f->seed2, km, k, rot);
            alive[j] = 1;
        }
        uint64_t dead = 0;
        /* phase list: (step, lo, hi) */
        for (uint32_t ph = 0; !decided; ph++) {
            uint32_t step;
            uint64_t lo, hi;
            if (ph == 0) {
                step = 0, lo = 0, hi = cnt < 32 ? cnt : 32;
            } else if (ph == 1) {
                if (cnt <= 32) continue;
                step = 0, lo = 32, hi = cnt;
            } else {
                step = ph - 1;
                if (step >= n_steps) break;
                if (cnt - dead == 0) break;
                lo = 0, hi = cnt;
            }
            for (uint64_t j = lo; j < hi; j++) {
                if (!alive[j]) continue;
                uint64_t g = step == 0 ? h1[j] : step == 1 ? h2[j] : (h1[j] + (uint64_t)step) * h2[j];
                uint64_t idx = g % f->m;
                probes++;
                if (!((f->words[idx >> 6] >> (idx & 63)) & 1)) {
                    alive[j] = 0;
                    dead++;
                }
            }
            if (misses + dead > allowed) {
                decided = 1;
                result = 0;
            }
        }
        if (decided) break;
        misses += dead;
        hits += cnt - dead;
        if (hits >= r->need) {
            decided = 1;
            result = 1;
        }
    }
    if (!decided) result = hits >= r->need;
    free(h1);
    free(h2);
    free(alive);
    *pass_out = result;
    return probes;
}

/* query_passes (query.rs:38-49) */
static int query_passes(const pfo_filter *f, const oread *r, size_t k, int rot, uint64_t *probes_ref) {
    uint64_t matches = 0, pr = 0;
    for (uint64_t i = 0; i < r->n_k; i++) matches += (uint64_t)pfo_filter_contains(f, r->kmers + i * k, k, rot, &pr);
    *probes_ref = pr;
    return matches >= r->need;
}


typedef struct {
    pfo_tree *t;
    const oread *reads;
    int want_hits;
    pfo_query_result *out;
    uint64_t hit_cap;
    uint64_t leaf_cursor; /* DFS leaf index of the next leaf reached by the recursion */
} qctx;

static void push_hit(qctx *c, uint32_t read, uint32_t leaf) {
    pfo_query_result *o = c->out;
    if (o->n_hits == c->hit_cap) {
        c->hit_cap = c->hit_cap ? c->hit_cap * 2 : 1024;
        o->hit_read = (uint32_t *)realloc(o->hit_read, c->hit_cap * 4);
        o->hit_leaf = (uint32_t *)realloc(o->hit_leaf, c->hit_cap * 4);
    }
    o->hit_read[o->n_hits] = read;
    o->hit_leaf[o->n_hits] = leaf;
    o->n_hits++;
}
static uint64_t subtree_leaves(const pfo_node *n) {
    if (!n) return 0;
    if (is_leaf(n)) return 1;
    return subtree_leaves(n->left) + subtree_leaves(n->right);
}

/* _query_batch (query.rs:99-158): pre-order, left then right, on the surviving subset */
static void query_rec(qctx *c, pfo_node *node, const uint32_t *set, uint64_t n_set) {
    const pfo_filter *f = node_filter(c->t, node); /* bf_cache.get_filter(&node.bloom_filter_path), query.rs:107-110 */
    if (!f) return;                                /* pfo_filter_load has set the error message */
    size_t k = (size_t)c->t->kmer_size;
    int rot = c->t->rot;
    uint8_t *flag = (uint8_t *)malloc(n_set ? n_set : 1);
    uint64_t pr_ref = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : pr_ref)
    for (uint64_t i = 0; i < n_set; i++) {
        uint64_t p = 0;
        flag[i] = (uint8_t)query_passes(f, &c->reads[set[i]], k, rot, &p); /* :113-117 */
        pr_ref += p;
    }
    c->out->probes_ref += pr_ref;
    c->out->pairs += n_set;
    uint32_t *pass = (uint32_t *)malloc((n_set ? n_set : 1) * 4);
    uint64_t n_pass = 0;
    for (uint64_t i = 0; i < n_set; i++)
        if (flag[i]) pass[n_pass++] = set[i];
    free(flag);
    if (!is_leaf(node)) {
        if (n_pass) { /* :122 */
            if (node->left) query_rec(c, node->left, pass, n_pass);
            if (node->right) query_rec(c, node->right, pass, n_pass);
        } else {
            c->leaf_cursor += subtree_leaves(node);
        }
    } else {
        node->mapped_reads += n_pass; /* :143 */
        if (c->want_hits)
            for (uint64_t i = 0; i < n_pass; i++) push_hit(c, pass[i], (uint32_t)c->leaf_cursor);
        c->leaf_cursor++;
    }
    free(pass);
}

/* ---- restatement of the GPU kernel's schedule (not of the reference): step-limited pre-test at
 * verified-monotone interior nodes, exact evaluation at leaves and everywhere else.  Used to check that the
 * schedule decides exactly like the reference and to predict the kernel's pair and probe counts. ---- */
static uint64_t filter_popcount(const pfo_filter *f) {
    uint64_t c = 0;
    for (uint64_t i = 0; i < f->nwords; i++) c += (uint64_t)__builtin_popcountll(f->words[i]);
    return c;
}
static int filter_contains_filter(const pfo_filter *parent, const pfo_filter *child) {
    uint64_t n = parent->nwords < child->nwords ? parent->nwords : child->nwords;
    for (uint64_t i = 0; i < n; i++)
        if (child->words[i] & ~parent->words[i]) return 0;
    return 1;
}
/* ---- expected-cost step plan -------------------------------------------------------------------------
 * The same text (modulo the node accessors) lives in oracle/pf_oracle.c and phagefilter_b200/csrc/pf_query.cu;
 * only +,-,*,/ and sqrt on doubles and a fixed table, so both produce the same table bit for bit.
 * Model: a read unrelated to the subtree has n absent k-mers; at a node with fill f probed for s steps each
 * k-mer is proven absent with probability p = 1 - f^s after (1-f^s)/(1-f) expected probes; the read is pruned
 * when more than `allowed` k-mers are proven absent: P = Phi((n p - allowed - 0.5) / sqrt(n p (1-p))).
 * Bottom-up, a verified interior node picks s in {0 (skip), 1..K} and a k-mer sampling stride t in {1,2,4,8} minimising
 *     probes(f,s) * n_s/n + PAIR_OVERHEAD + (1 - P) * (cost(left) + cost(right))    [s = 0: just the children]
 * with n_s = floor(n/t) sampled k-mers and P the prune probability of n_s trials.
 * Leaves and unverified nodes are exact (s = K). */
#define PF_PLAN_PAIR_OVERHEAD 0.25
static double plan_phi(double z) { /* standard normal CDF: 33-point table on [-4,4], linear interpolation */
    static const double T[33] = {3.167124183312e-05, 8.841728520081e-05, 2.326290790355e-04, 5.770250423908e-04, 1.349898031630e-03, 2.979763235055e-03, 6.209665325776e-03, 1.222447265504e-02, 2.275013194818e-02, 4.005915686382e-02, 6.680720126886e-02, 1.056497736669e-01, 1.586552539315e-01, 2.266273523769e-01, 3.085375387260e-01, 4.012936743171e-01, 5.000000000000e-01, 5.987063256829e-01, 6.914624612740e-01, 7.733726476231e-01, 8.413447460685e-01, 8.943502263331e-01, 9.331927987311e-01, 9.599408431362e-01, 9.772498680518e-01, 9.877755273450e-01, 9.937903346742e-01, 9.970202367649e-01, 9.986501019684e-01, 9.994229749576e-01, 9.997673709210e-01, 9.999115827148e-01, 9.999683287582e-01};
    if (z <= -4.0) return 0.0;
    if (z >= 4.0) return 1.0;
    const double x = (z + 4.0) * 4.0;
    const int i = (int)x;
    return T[i] + (T[i + 1] - T[i]) * (x - (double)i);
}
static double plan_probe_cost(double f, uint32_t s) { /* expected probes per absent k-mer over s steps */
    double c = 0.0, p = 1.0;
    for (uint32_t i = 0; i < s; ++i) {
        c += p;
        p *= f;
    }
    return c;
}
static double plan_prune_prob(double n, double allowed, double f, uint32_t s) {
    double surv = 1.0;
    for (uint32_t i = 0; i < s; ++i) surv *= f;
    const double p = 1.0 - surv, mean = n * p, var = n * p * (1.0 - p);
    if (var < 1e-9) return mean > allowed ? 1.0 : 0.0;
    return plan_phi((mean - allowed - 0.5) / sqrt(var));
}
/* best (steps, stride) for one verified interior node; *cost_out = its expected cost.  Stride t > 1 probes only
 * floor(n / t) k-mers, every t-th one centred in the read: an unprobed k-mer is not proven absent, so the test stays
 * sound; its probes shrink by n_s / n and its pruning power is that of n_s trials.  t is a power of two <= 8 (a
 * substitution spoils k >= 17 consecutive k-mers, so reads with errors are still caught) and never leaves fewer
 * than 16 k-mers (a gather costs per sector, not per instruction, so half a warp round is still efficient).  A sample is only ever probed for ONE step: with few k-mers per pair, further
 * steps are dependent round trips with ever fewer probes in flight (measured: 66 G probes/s instead of 250). */
static uint32_t plan_choose(double f, uint32_t K, uint64_t n_nominal, double allowed, double below, uint32_t *stride_out,
                            double *cost_out) {
    uint32_t best_s = 0, best_t = 1;
    double best = below;
    const double n = (double)n_nominal;
    for (uint32_t t = 1; t <= 8; t *= 2) {
        const uint64_t n_s = n_nominal / t;
        if (t > 1 && n_s < 16) break;
        const double ns = (double)(n_s ? n_s : 1), frac = ns / n;
        for (uint32_t s = 1; s <= (t > 1 ? 1u : K); ++s) { /* a sample is probed for one step only, see above */
            const double c = plan_probe_cost(f, s) * frac + PF_PLAN_PAIR_OVERHEAD + (1.0 - plan_prune_prob(ns, allowed, f, s)) * below;
            if (c < best) {
                best = c;
                best_s = s;
                best_t = t;
            }
        }
    }
    *stride_out = best_s ? best_t : 1;
    *cost_out = best;
    return best_s;
}
static double plan_exact_cost(double f, uint32_t K, double n, double allowed, double below) {
    return plan_probe_cost(f, K) + PF_PLAN_PAIR_OVERHEAD + (1.0 - plan_prune_prob(n, allowed, f, K)) * below;
}

static double plan_rec(const pfo_tree *t, pfo_node *n, uint64_t n_nominal, double allowed, int lazy) {
    const double nn = (double)n_nominal;
    n->stride = 1;
    const pfo_filter *f = t->filters[n->filter];
    const uint32_t K = f->K;
    if (!n->analysed) {
        n->mono = (!n->left || filter_contains_filter(f, t->filters[n->left->filter])) &&
                  (!n->right || filter_contains_filter(f, t->filters[n->right->filter]));
        n->pop = filter_popcount(f);
        n->analysed = 1;
    }
    const double fill = (double)n->pop / (double)f->m;
    if (is_leaf(n)) {
        n->steps = K;
        return plan_probe_cost(fill, K) + PF_PLAN_PAIR_OVERHEAD;
    }
    const double cl = n->left ? plan_rec(t, n->left, n_nominal, allowed, lazy) : 0.0;
    const double cr = n->right ? plan_rec(t, n->right, n_nominal, allowed, lazy) : 0.0;
    const double below = cl + cr;
    if (!lazy || !n->mono) {
        n->steps = K;
        return plan_exact_cost(fill, K, nn, allowed, below);
    }
    double c = 0.0;
    n->steps = plan_choose(fill, K, n_nominal, allowed, below, &n->stride, &c);
    return c;
}
static void plan_steps(pfo_tree *t, float threshold, uint64_t n_nominal, int lazy) {
    const uint64_t need = pfo_need(threshold, n_nominal);
    const double allowed = need > n_nominal ? 0.0 : (double)(n_nominal - need);
    if (t->root) plan_rec(t, t->root, n_nominal, allowed, lazy);
}
static uint32_t pfo_node_steps(const pfo_tree *t, pfo_node *n, float threshold, int lazy) {
    (void)t;
    (void)threshold;
    (void)lazy;
    return n->steps; /* filled by plan_steps for this query */
}

typedef struct {
    pfo_tree *t;
    const oread *reads;
    int want_hits, lazy;
    uint32_t group_rounds;
    float threshold;
    pfo_query_result *out;
    uint64_t hit_cap;
    uint64_t leaf_cursor;
} sctx;

/* top != 0: every ancestor was skipped by the plan.  The kernel never materialises that region: it starts
 * the frontier at the first tested nodes (entry nodes), so skipped top nodes count no pairs. */
static void sched_rec(sctx *c, pfo_node *node, const uint32_t *set, uint64_t n_set, int top) {
    if (top && node->steps == 0 && !is_leaf(node)) {
        if (node->left) sched_rec(c, node->left, set, n_set, 1);
        if (node->right) sched_rec(c, node->right, set, n_set, 1);
        return;
    }
    const pfo_filter *f = c->t->filters[node->filter];
    size_t k = (size_t)c->t->kmer_size;
    int rot = c->t->rot;
    uint32_t n_steps = pfo_node_steps(c->t, node, c->threshold, c->lazy);
    uint8_t *flag = (uint8_t *)malloc(n_set ? n_set : 1);
    uint64_t pr = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : pr)
    for (uint64_t i = 0; i < n_set; i++) {
        int pass = 0;
        pr += pair_sched(f, n_steps, node->stride ? node->stride : 1, c->group_rounds, &c->reads[set[i]], k, rot, &pass);
        flag[i] = (uint8_t)pass;
    }
    c->out->probes_sched += pr;
    c->out->pairs += n_set;
    uint32_t *pass = (uint32_t *)malloc((n_set ? n_set : 1) * 4);
    uint64_t n_pass = 0;
    for (uint64_t i = 0; i < n_set; i++)
        if (flag[i]) pass[n_pass++] = set[i];
    free(flag);
    if (!is_leaf(node)) {
        if (n_pass) {
            if (node->left) sched_rec(c, node->left, pass, n_pass, 0);
            if (node->right) sched_rec(c, node->right, pass, n_pass, 0);
        } else {
            c->leaf_cursor += subtree_leaves(node);
        }
    } else {
        if (c->want_hits) {
            qctx q = {c->t, c->reads, 1, c->out, c->hit_cap, 0};
            for (uint64_t i = 0; i < n_pass; i++) push_hit(&q, pass[i], (uint32_t)c->leaf_cursor);
            c->hit_cap = q.hit_cap;
        }
        c->leaf_cursor++;
    }
    free(pass);
}

static int query_impl(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, float threshold,
                      int threads, int want_hits, int sched, int lazy, uint32_t group_rounds, pfo_query_result *out);

int pfo_query_batch(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, float threshold,
                    int threads, int want_hits, pfo_query_result *out) {
    return query_impl(t, seqs, offs, n_reads, threshold, threads, want_hits, 0, 0, 1, out);
}
/* Leaf counters are NOT touched; hits/pairs/probes_sched describe the kernel schedule. */
int pfo_query_batch_sched(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, float threshold,
                          int threads, int want_hits, int lazy, uint32_t group_rounds, pfo_query_result *out) {
    return query_impl(t, seqs, offs, n_reads, threshold, threads, want_hits, 1, lazy, group_rounds ? group_rounds : 1,
                      out);
}

static int query_impl(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, float threshold,
                      int threads, int want_hits, int sched, int lazy, uint32_t group_rounds, pfo_query_result *out) {
    memset(out, 0, sizeof *out);
    g_err[0] = 0;
    if (!t->root) return 0; /* root.take().map(...) on None (query.rs:72-80) */
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    else omp_set_num_threads(omp_get_num_procs());
#else
    (void)threads;
#endif
    size_t k = (size_t)t->kmer_size;
    oread *reads = (oread *)calloc(n_reads ? n_reads : 1, sizeof *reads);
    /* DNASequence.kmers: one canonical byte vector per k-mer (file_parser.rs:135-148) */
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_reads; i++) total += pfo_num_kmers((size_t)(offs[i + 1] - offs[i]), k) * k;
    uint8_t *kbuf = (uint8_t *)malloc(total ? total : 1);
    uint64_t *koff = (uint64_t *)malloc(((uint64_t)n_reads + 1) * 8);
    koff[0] = 0;
    for (uint32_t i = 0; i < n_reads; i++)
        koff[i + 1] = koff[i] + pfo_num_kmers((size_t)(offs[i + 1] - offs[i]), k) * k;
#pragma omp parallel for schedule(dynamic, 256)
    for (uint32_t i = 0; i < n_reads; i++) {
        size_t len = (size_t)(offs[i + 1] - offs[i]);
        reads[i].n_k = pfo_num_kmers(len, k);
        reads[i].kmers = kbuf + koff[i];
        pfo_get_kmers(seqs + offs[i], len, k, kbuf + koff[i]);
        reads[i].need = pfo_need(threshold, reads[i].n_k);
    }
    uint32_t *all = (uint32_t *)malloc((n_reads ? n_reads : 1) * 4);
    for (uint32_t i = 0; i < n_reads; i++) all[i] = i;
    if (sched && t->lru_cap) {
        set_err("the kernel-schedule restatement needs every filter resident (pfo_tree_load, not pfo_tree_load_lazy)");
        free(all);
        free(koff);
        free(kbuf);
        free(reads);
        return -1;
    }
    if (sched) {
        refresh(t);
        {   /* nominal read: mean length of the block, as the kernel's planner */
            uint64_t mean_len = n_reads ? (offs[n_reads] - offs[0]) / n_reads : 0;
            uint64_t nk = pfo_num_kmers((size_t)mean_len, k);
            plan_steps(t, threshold, nk ? nk : 1, lazy);
        }
        sctx c = {t, reads, want_hits, lazy, group_rounds, threshold, out, 0, 0};
        sched_rec(&c, t->root, all, n_reads, 1);
    } else {
        qctx c = {t, reads, want_hits, out, 0, 0};
        query_rec(&c, t->root, all, n_reads);
    }
    free(all);
    free(koff);
    free(kbuf);
    free(reads);
    return g_err[0] ? -1 : 0;
}
/* The reference's query driver loop (main.rs:334-368): the read set is cut into blocks of `block_size` reads
 * (`--block-size-reads`, default 100, main.rs:115-118) and query_batch runs once per block, serially; hit read indices
 * are relative to the whole set.  With a tree from pfo_tree_load_lazy every block re-walks the tree through the LRU. */
int pfo_query_blocks(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, float threshold,
                     int threads, int want_hits, uint32_t block_size, pfo_query_result *out) {
    memset(out, 0, sizeof *out);
    if (block_size == 0) block_size = 1;
    uint64_t cap = 0;
    for (uint32_t b0 = 0; b0 < n_reads; b0 += block_size) {
        const uint32_t nb = n_reads - b0 < block_size ? n_reads - b0 : block_size;
        pfo_query_result r;
        if (query_impl(t, seqs, offs + b0, nb, threshold, threads, want_hits, 0, 0, 1, &r)) {
            pfo_query_result_free(&r);
            return -1;
        }
        if (r.n_hits) {
            if (out->n_hits + r.n_hits > cap) {
                cap = (out->n_hits + r.n_hits) * 2;
                out->hit_read = (uint32_t *)realloc(out->hit_read, cap * 4);
                out->hit_leaf = (uint32_t *)realloc(out->hit_leaf, cap * 4);
            }
            for (uint64_t i = 0; i < r.n_hits; i++) {
                out->hit_read[out->n_hits + i] = r.hit_read[i] + b0;
                out->hit_leaf[out->n_hits + i] = r.hit_leaf[i];
            }
            out->n_hits += r.n_hits;
        }
        out->pairs += r.pairs;
        out->probes_ref += r.probes_ref;
        pfo_query_result_free(&r);
    }
    return 0;
}
void pfo_query_result_free(pfo_query_result *r) {
    free(r->hit_read);
    free(r->hit_leaf);
    memset(r, 0, sizeof *r);
}

/* save_leaf_counts (query.rs:173-183): "{id},{count}\n" for count > 0, DFS leaf order */
uint64_t pfo_classification_csv(const pfo_tree *t, char *buf, uint64_t cap) {
    refresh((pfo_tree *)t);
    uint64_t w = 0;
    for (uint64_t i = 0; i < t->n_leaves; i++) {
        const pfo_node *n = t->leaves[i];
        if (n->mapped_reads == 0) continue;
        char line[1024];
        int L = snprintf(line, sizeof line, "%s,%llu\n", n->tax_id ? n->tax_id : "", (unsigned long long)n->mapped_reads);
        if (w + (uint64_t)L <= cap && buf) memcpy(buf + w, line, (size_t)L);
        w += (uint64_t)L;
    }
    return w;
}
