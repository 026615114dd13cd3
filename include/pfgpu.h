/*
 * pfgpu.h -- C ABI of libpfgpu: the B200 (sm_100a) implementation of PhageFilter's `query`
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no C++ / torch types.
 *
 * The reference (Dreycey/PhageFilter, one Rust binary crate) has no FFI of its own; the seam
 * is the one `src/main.rs` already uses around the query loop.  Each entry point below cites
 * the reference interface it replaces (file:line relative to the reference root).  The Rust
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns PF_OK (0) or a non-zero pf_status; pf_last_error() gives a
 *     thread-local message.  Nothing unwinds across the boundary.
 *   - one pf_db per GPU; calls on one handle are serialised by the caller (the reference
 *     calls query_batch serially from the main thread, main.rs:337-342).
 *   - the caller owns its input buffers; the library owns device memory and every array it
 *     returns (valid until the next call on the same handle).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef PFGPU_H
#define PFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum pf_status {
    PF_OK = 0,
    PF_ERR_ARG = 1,     /* bad argument */
    PF_ERR_IO = 2,      /* file missing / unreadable (reference: panic in BloomTree::load, bloom_tree.rs:366-375) */
    PF_ERR_FORMAT = 3,  /* tree.bin / .bf does not decode (reference: panic, bloom_filter.rs:163-168) */
    PF_ERR_CUDA = 4,    /* CUDA runtime error or no device */
    PF_ERR_NOMEM = 5,   /* host or device allocation failed (frontier too large: split the block) */
    PF_ERR_NCCL = 6,
    PF_ERR_STATE = 7
} pf_status;

const char *pf_last_error(void);
/* library version string, e.g. "pfgpu 0.1 sm_100a" */
const char *pf_version(void);

/* ------------------------------------------------------------------------------------------
 * Database handle: flattened, level-ordered, device-resident gSBT.
 * Replaces BloomTree::load (bloom_tree.rs:364-386), BloomTree::prune_tree (:302-330) and the
 * lazily loading BFLruCache (cache.rs:38-77): every distinct `.bf` is decoded once and kept
 * in HBM; nodes that name the same file share one filter.
 * ---------------------------------------------------------------------------------------- */
typedef struct pf_db pf_db;

typedef struct pf_db_info_t {
    uint64_t kmer_size;        /* BloomTree.kmer_size */
    uint64_t num_bits;         /* m = bits.len() of every filter */
    uint64_t words_per_filter; /* u64 words per filter in HBM (padded to 128 B) */
    uint64_t n_nodes;          /* after pruning */
    uint64_t n_leaves;         /* after pruning */
    uint64_t n_filters;        /* distinct .bf files resident */
    uint64_t n_levels;         /* tree depth + 1 */
    uint64_t filter_bytes;     /* HBM bytes held by filters */
    uint64_t seed1, seed2;     /* hash_states (bloom_tree.rs:46-47) */
    uint32_t num_hashes;       /* K */
    uint32_t largest_genome;
    float false_pos_rate;
    int32_t device;
    int32_t hash_rot;          /* FxHasher::finish rotate in use (26 = rustc-hash 2.1.1) */
    int32_t fast_path;         /* 1 if the 2-bit register path covers this k (17..32) */
    uint64_t n_internal;       /* interior nodes */
    uint64_t n_monotone;       /* interior nodes whose filter was verified to contain both children's */
} pf_db_info_t;

/* search_depth < 0: no pruning (main.rs:293-299 passes Some(depth)). */
int pf_db_open(const char *db_path, int device, int64_t search_depth, pf_db **out);
int pf_db_info(const pf_db *db, pf_db_info_t *out);
/* tax_id of the leaf with left-first DFS index `dfs_leaf` (query.rs:197-218 order). */
const char *pf_db_leaf_id(const pf_db *db, uint64_t dfs_leaf);
/* rustc-hash's finish() rotate differs between crate versions; default 26. */
int pf_db_set_hash_rot(pf_db *db, int rot);
/* Self-certifying known-answer test: rebuild leaf `dfs_leaf` from its genome with rot in {26,20}
 * and compare with the stored bits.  *rot_out = matching rotate, or -1 if neither matches. */
int pf_db_detect_hash_rot(pf_db *db, uint64_t dfs_leaf, const uint8_t *genome, uint64_t len, int *rot_out);
void pf_db_close(pf_db *db);

/* ------------------------------------------------------------------------------------------
 * Read batches.  Replaces DNASequence.kmers (file_parser.rs:135-155): k-mers are never
 * materialised; reads travel as 2-bit codes (A=0,C=1,G=2,T=3; base j of a read sits at bits
 * [2j,2j+2) of its little-endian bit stream, 16 bases per uint32 word) and the canonical
 * k-mer bytes the reference hashes are re-created in registers.  A read holding any byte other
 * than upper-case A/C/G/T is an *exception read*: its raw bytes travel verbatim and take the
 * byte-exact slow path (the reference hashes raw bytes, file_parser.rs:135-148).
 * ---------------------------------------------------------------------------------------- */
typedef struct pf_read_batch {
    uint32_t n_reads;
    uint32_t n_exc;            /* number of exception reads */
    const uint32_t *lengths;   /* [n_reads] bases per read */
    const uint64_t *word_off;  /* [n_reads] first word of each read in `packed`; always even */
    const uint32_t *packed;    /* [n_words] 2-bit codes; must be followed by >= 4 readable pad words */
    uint64_t n_words;          /* including the trailing pad */
    const uint32_t *exc_index; /* [n_reads] index into exc_off for exception reads, 0xFFFFFFFF otherwise; NULL if n_exc==0 */
    const uint64_t *exc_off;   /* [n_exc+1] byte offsets into exc_bytes */
    const uint8_t *exc_bytes;  /* raw bytes of the exception reads */
    uint32_t max_length;       /* longest read in bases; 0 = unknown (the library then scans `lengths`) */
    uint64_t total_bases;      /* sum of lengths; 0 = unknown */
} pf_read_batch;

/* Host-side packer from raw ASCII reads (seqs concatenated, offs[n_reads+1]) into pinned memory, on several
 * host threads (PF_PACK_THREADS overrides the count).  *out must be NULL (a new pf_packed is allocated) or a
 * pf_packed from an earlier call, whose buffers are then recycled -- page-locking memory is slower than
 * packing it. */
typedef struct pf_packed pf_packed;
int pf_pack_reads(const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, pf_packed **out);
/* Same, for reads scattered in memory (e.g. slices of a parsed FASTQ buffer): one pointer and length per read. */
int pf_pack_reads_ptrs(const uint8_t *const *seq_ptrs, const uint32_t *lengths, uint32_t n_reads, pf_packed **out);
const pf_read_batch *pf_packed_batch(const pf_packed *p);
/* Gives *inout (NULL: a new pf_packed) page-locked buffers at least as large as `model` has reached, without packing
 * anything: a host that rotates several batches (an ingest thread ahead of the query thread) locks all of them once,
 * up front -- page-locking holds the driver's lock and would otherwise stall the first queries. */
int pf_packed_reserve_like(pf_packed **inout, const pf_packed *model);
void pf_packed_free(pf_packed *p);
void *pf_alloc_pinned(size_t bytes);
void pf_free_pinned(void *p);
/* Host threads other than the one that opened the handle (an ingest thread that packs the next batch while the
 * query thread runs, main.rs:322-342 split in two) call this once so that their pinned allocations belong to the
 * handle's GPU.  A no-op without a CUDA device. */
int pf_thread_set_device(int device);

/* Per-read matched leaves of one block: CSR over the block's reads; leaves are DFS leaf
 * indices, ascending within a read.  Replaces ResultMap (result_map.rs:9-46). */
typedef struct pf_hits {
    uint64_t n_hits;
    const uint64_t *read_off; /* [n_reads+1] */
    const uint32_t *leaf;     /* [n_hits] */
} pf_hits;

/* query::query_batch (query.rs:66-82) for one block: H2D copy of the batch, level-synchronous
 * descent on the GPU, leaf counters accumulated in the handle (mapped_reads, query.rs:143),
 * and -- if want_hits -- the (read -> leaves) map copied back.  threshold is f32 and the pass
 * bound is ceilf(threshold * (float)n_kmers) exactly as query.rs:48. */
int pf_query_block(pf_db *db, const pf_read_batch *in, float threshold, int want_hits, pf_hits *out);

/* Same query on a batch already resident in HBM (used to time the device path alone). */
typedef struct pf_dev_batch pf_dev_batch;
int pf_batch_upload(pf_db *db, const pf_read_batch *in, pf_dev_batch **out);
int pf_query_device(pf_db *db, pf_dev_batch *batch, float threshold, int want_hits, pf_hits *out);
/* Pipelined ingest: starts the H2D copies of `in` on the handle's copy stream and returns at once, so the
 * upload of block i+1 overlaps pf_query_device on block i (which waits for its own batch's copies only).
 * *inout == NULL allocates a device batch; otherwise that batch is reused (it must not be in use by a
 * running query).  The caller keeps the host arrays alive until pf_query_device on the batch has returned. */
int pf_batch_upload_async(pf_db *db, const pf_read_batch *in, pf_dev_batch **inout);
void pf_batch_free(pf_db *db, pf_dev_batch *batch);

/* Leaf counters: BloomNode.mapped_reads in DFS leaf order (query.rs:197-218). */
int pf_leaf_counts(pf_db *db, uint64_t *counts /* [n_leaves] */);
int pf_reset_counts(pf_db *db);
/* query::save_leaf_counts (query.rs:173-183): "{id},{count}\n" for count > 0. */
int pf_save_leaf_counts(pf_db *db, const char *csv_path);

/* Work / timing counters since the last pf_reset_stats. */
typedef struct pf_stats_t {
    uint64_t blocks;          /* pf_query_* calls */
    uint64_t reads;
    uint64_t pairs;           /* (read,node) pairs evaluated */
    uint64_t probes_issued;   /* bloom bit probes issued by the kernel */
    uint64_t levels;          /* tree levels traversed */
    uint64_t probe_launches;  /* launches of the probe kernel */
    uint64_t other_launches;  /* init / scan / scatter / pack kernels */
    uint64_t h2d_bytes, d2h_bytes;
    double probe_kernel_ms;   /* CUDA-event time summed over probe launches */
    double device_ms;         /* CUDA-event time of the query calls, first launch to last */
    uint64_t group_rounds;    /* rounds of 32 k-mers each lane owned at once in the last call (1..8) */
    uint64_t memo_hits;       /* k-mers answered by the k-mer memo instead of K - 1 probes (pf_db_set_memo) */
    uint64_t memo_lookups;    /* k-mers looked up in the memo (one 8-byte read each) */
    uint64_t sliced_blocks;   /* query calls evaluated on bit-sliced tiles (pf_db_set_mode); `pairs` are then (read, tile)
                                 pairs (see sliced_pairs, sector_loads) */
    uint64_t sliced_tiles;    /* tiles of the current tiling */
    uint64_t sliced_table_bytes; /* HBM held by their tables */
    uint64_t chunk_splits;    /* times a chunk of reads was cut in half because its frontier outgrew the pair index */
    uint64_t sector_loads;    /* row (32-byte sector) loads of the sliced kernel; `probes_issued` = bit probes of the node-at-a-time kernel */
    uint64_t sliced_pairs;    /* (read, tile) pairs; `pairs` counts those and the (read, node) pairs */
    double sliced_kernel_ms;  /* CUDA-event time of the sliced kernel's launches (`probe_kernel_ms`: the node-at-a-time kernel's) */
    uint64_t sliced_launches; /* timed launches of the sliced kernels (one per tile-tree depth; the entry depth may be two kernels timed as one) */
    uint64_t line_loads;      /* 128-byte line loads of the entry line kernel (one line = one row of up to four entry tiles); not in sector_loads */
    double entry_kernel_ms;   /* the share of sliced_kernel_ms spent at the entry depth (every read against every entry tile) */
} pf_stats_t;
int pf_get_stats(pf_db *db, pf_stats_t *out);
/* The CUDA stream (cudaStream_t) every kernel and copy of this handle is issued on, so a caller can
 * bracket calls with its own events on the launching stream. */
void *pf_db_stream(pf_db *db);
int pf_reset_stats(pf_db *db);
/* 0 (default): read-level early exit on.  1: reference-faithful probing -- every k-mer of every
 * pair is probed until its first clear bit (probes_issued then equals the reference's count). */
int pf_db_set_exhaustive(pf_db *db, int on);
/* 1 (default): interior nodes whose filter was verified at load to be a bitwise superset of their
 * children's filters ("child passes => parent passes") get a step-limited, sound cannot-pass test
 * instead of the exact one; leaves and unverified nodes stay exact, so results are identical.
 * 0: every node is evaluated exactly (the frontier then equals the reference's, query.rs:113-141). */
int pf_db_set_lazy(pf_db *db, int on);
/* How a block is evaluated.  1: node at a time -- a warp per (read, node) pair probing that node's filter (one bit per
 * 32-byte sector touched; filters stay L2-resident because the frontier is node-major).  2: bit-sliced tiles -- the
 * filters of up to 256 nodes are stored transposed, so one sector load answers a bloom probe for every node of the tile
 * (exact per-node k-mer counts, then the reference's descent rule on the pass bits).  0 (default): per (threshold, read
 * length) whichever the cost model expects to be faster; the tiles are built on first use from the resident filters
 * (about as much HBM again).  Results are identical in every mode.  The environment variable PF_MODE
 * (auto|pair|sliced) sets the initial value. */
int pf_db_set_mode(pf_db *db, int mode);
/* Sliced path, below the cut through the tree that every read is tested against.  0: tiles all the way down (about as much
 * HBM again as the filters).  1: tiles for the cut only (a few hundred MB); the (read, node) pairs they cannot rule out are
 * handed to the node-at-a-time descent.  -1 (default): 0 if the tables fit the free HBM, else 1.  Results are identical.
 * Environment: PF_SLICED_HANDOVER=0|1. */
int pf_db_set_handover(pf_db *db, int handover);
/* Sliced path: most columns (tree nodes) a tile may hold -- 32, 64, 128 or 256 (default).  Narrower tiles only make the
 * tiling finer (more entry tiles, more tile-tree depths); the tests use it to drive small trees through the multi-tile
 * code paths (entry tiles sharing 128-byte lines, several groups of them).  Results are identical. */
int pf_db_set_tile_cols(pf_db *db, int cols);
/* 1 (default): k-mer memo at exact nodes.  BloomFilter::contains depends on a k-mer only through its 64-bit
 * hash_bytes value, so once a k-mer has passed all K probes at a node, every later occurrence of the same value at that
 * node within the block (sequencing depth: 30x in BASELINE config 2) is a hit after ONE table look-up instead of K - 1
 * further probes.  Exact (a table entry is the full 64-bit value, tables are per node and zeroed per level and block),
 * so results do not change; the number of probes issued then depends on timing, which is why the work-count parity
 * tests switch it off.  budget_bytes = 0 keeps the current budget (default 256 MiB). */
int pf_db_set_memo(pf_db *db, int on, uint64_t budget_bytes);
/* hash_bytes of every k-mer is computed once per batch and cached in HBM (8 B per k-mer); a batch whose
 * cache would exceed `bytes` (default 16 GiB) is processed in several chunks of reads. */
int pf_db_set_hash_cache_bytes(pf_db *db, uint64_t bytes);
/* (read, node) pairs are indexed with 32 bits.  A block whose frontier would outgrow that is not refused: the library
 * works through it in chunks of reads and halves a chunk whose frontier grows past `pairs` (default and maximum
 * 0xFF000000).  Results do not depend on the value; tests set it low to exercise the splitting. */
int pf_db_set_frontier_cap(pf_db *db, uint64_t pairs);
/* Probe steps per node (level order, n_nodes entries) a query with `threshold` uses for a batch whose reads
 * of mean length have `nominal_kmers` k-mers (K = exact, 0 = skipped). */
int pf_db_node_steps(pf_db *db, float threshold, uint64_t nominal_kmers, uint32_t *steps);
/* Same, plus the k-mer sampling stride per node (1 = every k-mer; t > 1 only at verified interior nodes: every t-th
 * k-mer, centred in the read, is probed -- a sound cannot-pass test on a sample).  Either array may be NULL. */
int pf_db_node_plan(pf_db *db, float threshold, uint64_t nominal_kmers, uint32_t *steps, uint32_t *strides);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: reads are sharded by rank, every rank holds a replica of the tree, and the
 * per-leaf counters are combined with ONE ncclAllReduce(sum, u64[n_leaves]).
 * The 128-byte id is created on rank 0 and handed to the other ranks by the host.
 * ---------------------------------------------------------------------------------------- */
int pf_nccl_unique_id(void *id128);
int pf_comm_init(pf_db *db, int nranks, int rank, const void *id128);
int pf_allreduce_counts(pf_db *db);

/* ------------------------------------------------------------------------------------------
 * Subtree shards: trees larger than one GPU's HBM (BASELINE config 5 at the default geometry).
 * The tree is cut at one level: the levels above it are replicated, every subtree rooted at the cut level
 * is owned by one rank, and a rank keeps only the top's and its own subtrees' filters resident.  Queries
 * are collective: the ranks' reads are gathered, every rank descends the top with its own reads, the
 * surviving (read, node) pairs cross NVLink to the subtrees' owners in one all-to-all, and the (read, leaf)
 * hits return to the reads' owners in a second one.  Leaf counters stay partial per rank until
 * pf_allreduce_counts.  Results equal those of the replicated tree (same nodes see the same reads,
 * query.rs:99-158).
 * ---------------------------------------------------------------------------------------- */
typedef struct pf_shard_info_t {
    int32_t sharded, rank, nranks;
    uint32_t cut_level;        /* levels < cut_level are replicated */
    uint64_t top_nodes;        /* replicated nodes */
    uint64_t owned_nodes;      /* nodes of this rank's subtrees */
    uint64_t resident_filters; /* distinct .bf held in this rank's HBM */
    uint64_t resident_bytes;
} pf_shard_info_t;
typedef struct pf_shard_stats_t {
    uint64_t queries, collectives;
    uint64_t reads_gathered;   /* reads of all ranks seen by this rank */
    uint64_t pairs_top;        /* (read,node) pairs evaluated above the cut (own reads) */
    uint64_t pairs_subtrees;   /* pairs evaluated in this rank's subtrees (reads of every rank) */
    uint64_t pairs_sent, pairs_received; /* frontier exchange, excluding the rank's own slice */
    uint64_t hits_sent;
    uint64_t bytes_sent, bytes_received;
} pf_shard_stats_t;
/* Collective over `nranks` processes (one GPU each); id128 comes from pf_nccl_unique_id on rank 0.
 * cut_level < 0: chosen to minimise the filters one rank holds. */
int pf_db_open_sharded(const char *db_path, int device, int64_t search_depth, int nranks, int rank, const void *id128,
                       int64_t cut_level, pf_db **out);
int pf_shard_info(const pf_db *db, pf_shard_info_t *out);
int pf_shard_stats(pf_db *db, pf_shard_stats_t *out);
/* The partition alone, from tree.bin (host only, no GPU needed): owner_out[u] = -1 for replicated nodes, else
 * the owning rank, nodes in level order. */
int pf_shard_plan(const char *db_path, int64_t search_depth, int nranks, int64_t cut_level, uint32_t *cut_level_out,
                  int32_t *owner_out, uint64_t owner_cap, uint64_t *n_nodes_out);
/* query_batch (query.rs:66-82) on a sharded handle; every rank calls it with its own block of reads (possibly
 * empty) and gets the hit lists of its own reads.  pf_query_block / pf_query_device refuse sharded handles. */
int pf_query_sharded(pf_db *db, const pf_read_batch *in, float threshold, int want_hits, pf_hits *out);
int pf_query_sharded_device(pf_db *db, pf_dev_batch *batch, float threshold, int want_hits, pf_hits *out);

/* ------------------------------------------------------------------------------------------
 * Database builder on the GPU (the `build`/`add` side the query consumes):
 * BloomTree::new (bloom_tree.rs:100-119), ::insert (:128-145), ::save (:339-355).
 * Writes the reference's on-disk format (tree.bin + one .bf per node).
 * name_mode 0: "Internal_Node_<counter>"; 1: "Internal_Node_<u16>" from splitmix64(name_seed),
 * redrawn until unused (bloom_tree.rs:232-234 draws a random u16).
 * ---------------------------------------------------------------------------------------- */
typedef struct pf_builder pf_builder;
int pf_builder_create(uint64_t kmer_size, float false_pos_rate, uint32_t largest_genome, uint64_t seed1,
                      uint64_t seed2, int device, int name_mode, uint64_t name_seed, pf_builder **out);
/* BloomTree::load for `add` (main.rs:220-221): continue inserting into an existing database. */
int pf_builder_open(const char *db_path, int device, pf_builder **out);
int pf_builder_set_hash_rot(pf_builder *b, int rot);
int pf_builder_insert(pf_builder *b, const char *id, const uint8_t *seq, uint64_t len);
int pf_builder_save(pf_builder *b, const char *db_path);
void pf_builder_free(pf_builder *b);
/* filter geometry in f32 (bloom_filter.rs:342-357) */
uint64_t pf_needed_bits(float false_pos_rate, uint32_t num_items);
uint32_t pf_optimal_num_hashes(uint64_t num_bits, uint32_t num_items);

/* ------------------------------------------------------------------------------------------
 * Host-only view of the tiling of the sliced path (no device needed; used by the CPU tests).  The tree is given as
 * level-ordered arrays (children 0xFFFFFFFF = none, leaf[u] = DFS leaf index or -1, pop[u] = set bits of the node's
 * filter, mono[u] = filter verified a superset of both children's).  Per node: skipped or (tile, column); per tile: parent
 * tile (-1 = entry tile) and width in columns.  handover 1: tiles for the cut only.
 * ---------------------------------------------------------------------------------------- */
int pf_plan_tiles(uint64_t n_nodes, const uint32_t *left, const uint32_t *right, const int32_t *leaf, const uint64_t *pop,
                  const uint8_t *mono, uint64_t num_bits, uint32_t num_hashes, float threshold, uint64_t nominal_kmers,
                  int handover, uint8_t *skip_out, int32_t *node_tile_out, uint32_t *node_col_out, int32_t *tile_parent_out,
                  uint32_t *tile_width_out, uint64_t tile_cap, uint64_t *n_tiles_out, uint64_t *n_entry_out);
/* The same plan -- with tiles of at most tile_cols columns (pf_db_set_tile_cols) -- and where its tables lie: per tile the
 * first u32 word of its table, the words from one row to the next (32 when the tile shares 128-byte lines with up to three
 * other entry tiles, else the words of a row), the words of a row, the line group (-1: a table of its own), the pre-test
 * depth and whether the pre-test is all the tile does; the entry tiles in the order the kernels walk them (the first
 * n_line_tiles share lines, four per group); the u32 words all tables take and the rows of one table. */
int pf_plan_tile_layout(uint64_t n_nodes, const uint32_t *left, const uint32_t *right, const int32_t *leaf, const uint64_t *pop,
                        const uint8_t *mono, uint64_t num_bits, uint32_t num_hashes, float threshold, uint64_t nominal_kmers,
                        int handover, int tile_cols, uint64_t tile_cap, uint64_t *table_off_out, uint32_t *row_stride_out,
                        uint32_t *row_words_out, int32_t *tile_group_out, uint32_t *pre_steps_out, uint8_t *filter_only_out,
                        uint32_t *entry_order_out, uint64_t *n_tiles_out, uint64_t *n_entry_out, uint64_t *n_line_tiles_out,
                        uint64_t *table_words_out, uint64_t *rows_out);

/* ------------------------------------------------------------------------------------------
 * Roofline micro-benchmark: all SMs issue independent random 32-byte-sector loads over a
 * working set of `bytes` (1.8 MB -> L2-resident filter; >= L2 size -> HBM).  Reports sectors/s.
 * ---------------------------------------------------------------------------------------- */
int pf_microbench_sectors(int device, uint64_t bytes, int iters, double *sectors_per_s);

#ifdef __cplusplus
}
#endif
#endif /* PFGPU_H */
