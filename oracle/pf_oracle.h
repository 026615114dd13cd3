/*
 * pf_oracle.h -- CPU ORACLE for the PhageFilter `query` hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may load it.  The product
 * (phagefilter_b200/, libpfgpu.so) never links, imports or executes anything here.
 *
 * It is a plain-C restatement of the reference's algorithm (Dreycey/PhageFilter, Rust).
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  The reference cannot be compiled in this environment (no Rust
 * toolchain), and three pieces of arithmetic live in un-vendored crates
 * (rustc-hash ^2.1, bitvec 1.0.1 + serde, bincode 1.3.3, bio 2.2.0); their published
 * algorithms are restated here.
 *
 * PARITY UNPINNED at the bitvec-serde / bincode (file format) boundaries: the reference's
 * own tests hold no hash value, bit index or `.bf` byte image.
 * The rustc-hash arithmetic IS pinned, against a real rustc-hash 2.x build executed in this
 * image (tests/test_hash_pin_cpu.py: hashbrown bucket order of an FxHashMap<Vec<u8>,_> inside
 * the outlines_core extension; 100 % agreement with finish() = rotate_left(26), chance level
 * with 20).  Pinned against the reference's own tests: canonicalisation (file_parser.rs:396-407),
 * get_kmers windows (:380-393), the HashIter derivation (hash_iter.rs:75-90), filter
 * geometry (bloom_filter.rs:342-357), tree topology (bloom_tree.rs:458-734), query
 * semantics on the toy trees (query.rs:249-380) and get_ext_id strings (result_map.rs:78-103).
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rustc-hash 2.1.1 finish() rotate; 2.0.0/2.1.0 are believed to use 20.  Runtime parameter. */
#define PFO_DEFAULT_ROT 26

/* ---- L0 hashing (bloom_filter/hasher.rs:12-21, hash_iter.rs:13-45, rustc-hash FxHasher) ---- */
uint64_t pfo_hash_bytes(const uint8_t *bytes, size_t len);
uint64_t pfo_fx_hash(uint64_t seed, const uint8_t *item, size_t len, int rot);
/* g_i of HashIter for i in [0,count): out[i]  (hash_iter.rs:13-27) */
void pfo_hash_iter(uint64_t h1, uint64_t h2, uint32_t count, uint64_t *out);

/* ---- L4 canonical k-mers (file_parser.rs:114-148; bio::alphabets::dna::revcomp) ---- */
uint8_t pfo_complement(uint8_t b);
void pfo_revcomp(const uint8_t *kmer, size_t k, uint8_t *out);
void pfo_get_lex_less(const uint8_t *kmer, size_t k, uint8_t *out);
/* number of k-mers of a sequence: 0 if k==0 or k>len (file_parser.rs:136-139) */
size_t pfo_num_kmers(size_t seq_len, size_t k);
/* writes n_k * k bytes of canonical k-mers */
void pfo_get_kmers(const uint8_t *seq, size_t len, size_t k, uint8_t *out);

/* ---- L1 filter geometry (bloom_filter.rs:342-357) ---- */
uint64_t pfo_needed_bits(float fpr, uint32_t num_items);
uint32_t pfo_optimal_num_hashes(uint64_t num_bits, uint32_t num_items);

/* ---- L1 bloom filter (bloom_filter.rs:84-93, 142-151, 275-332) ---- */
typedef struct pfo_filter {
    uint64_t m;        /* bits.len() */
    uint64_t nwords;   /* ceil(m/64) */
    uint64_t *words;   /* BitVec<usize, Lsb0> raw storage */
    uint32_t K;        /* num_hashes */
    uint64_t seed1, seed2;
} pfo_filter;

pfo_filter *pfo_filter_new(uint64_t m, uint32_t K, uint64_t seed1, uint64_t seed2);
void pfo_filter_free(pfo_filter *f);
/* returns 1 if the item was NOT present before (bloom_filter.rs:291-307) */
int pfo_filter_insert(pfo_filter *f, const uint8_t *item, size_t len, int rot);
/* bloom_filter.rs:312-332; *probes (optional) += number of bits examined */
int pfo_filter_contains(const pfo_filter *f, const uint8_t *item, size_t len, int rot, uint64_t *probes);
void pfo_filter_union(pfo_filter *dst, const pfo_filter *src);
uint64_t pfo_filter_distance(const pfo_filter *a, const pfo_filter *b);
/* .bf (bincode + bitvec serde) -- SURVEY App. B.  0 on success. */
int pfo_filter_save(const pfo_filter *f, const char *path, const char *recorded_path);
pfo_filter *pfo_filter_load(const char *path);

/* ---- L2 tree (bloom_tree.rs:28-61, 128-299, 302-330, 339-386) ---- */
typedef struct pfo_tree pfo_tree;

/* name_mode: 0 = "Internal_Node_<counter>" (unique, deterministic),
 *            1 = "Internal_Node_<u16>" drawn from splitmix64(name_seed) -- reference-style
 *                random u16 names (bloom_tree.rs:232-234) but still made unique by redraw. */
pfo_tree *pfo_tree_new(uint64_t kmer_size, float fpr, uint32_t largest_genome,
                       uint64_t seed1, uint64_t seed2, int rot, int name_mode, uint64_t name_seed);
void pfo_tree_free(pfo_tree *t);
/* BloomTree::insert (bloom_tree.rs:128-145) for one genome record */
int pfo_tree_insert(pfo_tree *t, const char *id, const uint8_t *seq, size_t len);
int pfo_tree_save(const pfo_tree *t, const char *dir);
pfo_tree *pfo_tree_load(const char *dir, int rot);
void pfo_tree_prune(pfo_tree *t, uint64_t search_depth);
const char *pfo_last_error(void);

/* introspection */
uint64_t pfo_tree_num_nodes(const pfo_tree *t);
uint64_t pfo_tree_num_leaves(const pfo_tree *t);   /* after pruning */
uint64_t pfo_tree_kmer_size(const pfo_tree *t);
uint64_t pfo_tree_num_bits(const pfo_tree *t);
uint32_t pfo_tree_num_hashes(const pfo_tree *t);
void pfo_tree_seeds(const pfo_tree *t, uint64_t *s1, uint64_t *s2);
/* leaf ids in left-first DFS order (query.rs:197-218); returns pointer owned by the tree */
const char *pfo_tree_leaf_id(const pfo_tree *t, uint64_t dfs_leaf_index);
uint64_t pfo_tree_leaf_count(const pfo_tree *t, uint64_t dfs_leaf_index); /* mapped_reads */
void pfo_tree_reset_counts(pfo_tree *t);
/* pre-order topology dump for tests: for each node (pre-order): is_leaf, depth; name via pfo_tree_node_name */
uint64_t pfo_tree_preorder(const pfo_tree *t, uint8_t *is_leaf, uint32_t *depth, uint64_t cap);
const char *pfo_tree_node_name(const pfo_tree *t, uint64_t preorder_index);

/* ---- L3 query (query.rs:38-158) ---- */
typedef struct pfo_query_result {
    uint64_t n_hits;          /* (read, leaf) pairs: a leaf passed for that read */
    uint32_t *hit_read;       /* index into the block */
    uint32_t *hit_leaf;       /* DFS leaf index */
    uint64_t pairs;           /* (read,node) pairs evaluated */
    uint64_t probes_ref;      /* bit probes under the reference's semantics (k-mer early exit only) */
    uint64_t probes_sched;    /* bit probes under the GPU kernel's schedule (pfo_query_batch_sched only) */
} pfo_query_result;

/* query_batch on one block of reads given as ASCII bytes: seqs concatenated, offs[n+1].
 * Leaf counters accumulate inside the tree across calls (query.rs:142, main.rs:337-342).
 * threads<=0: all cores. */
int pfo_query_batch(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads,
                    float threshold, int threads, int want_hits, pfo_query_result *out);
void pfo_query_result_free(pfo_query_result *r);
/* The reference's driver loop around query_batch (main.rs:334-368): blocks of `block_size` reads, serially. */
int pfo_query_blocks(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, float threshold,
                     int threads, int want_hits, uint32_t block_size, pfo_query_result *out);
/* Reference-faithful filter store: only tree.bin is read; filters come through an LRU of `cache_size` entries and a
 * miss re-reads "<db>/<name>.bf" from disk (BFLruCache, cache.rs:55-88; --cache-size, main.rs:119-122).  Query only. */
pfo_tree *pfo_tree_load_lazy(const char *dir, int rot, int cache_size);
void pfo_tree_cache_stats(const pfo_tree *t, uint64_t *loads, uint64_t *hits, uint64_t *bytes);
/* Restatement of the GPU kernel's schedule (NOT of the reference): same decisions, different amount of
 * work.  lazy=1: verified-monotone interior nodes get the step-limited pre-test (see pfo_node_steps);
 * lazy=0: every node exact with read-level early exit.  Fills hits, pairs and probes_sched; leaf counters
 * of the tree are not touched. */
int pfo_query_batch_sched(pfo_tree *t, const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads,
                          float threshold, int threads, int want_hits, int lazy, uint32_t group_rounds,
                          pfo_query_result *out);
/* need = (threshold * n_k as f32).ceil() as usize (query.rs:48) -- saturating cast */
uint64_t pfo_need(float threshold, uint64_t n_kmers);
/* save_leaf_counts (query.rs:173-183): returns bytes written to buf (cap), or needed size */
uint64_t pfo_classification_csv(const pfo_tree *t, char *buf, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
