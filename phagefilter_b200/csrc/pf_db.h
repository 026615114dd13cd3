// pf_db.h -- internal state of libpfgpu shared by pf_query.cu (replicated tree) and pf_shard.cu (subtree shards):
// grow-only device / pinned arrays, the device copy of a read batch, the NCCL entry points (dlopen'ed so the
// library loads without NCCL) and the pf_db handle itself.
#pragma once
#include <nccl.h>

#include <algorithm>
#include <string>
#include <vector>

#include "pf_common.h"
#include "pf_format.h"
#include "pf_kernels.cuh"

namespace pf {

template <class T>
struct DevBuf {  // grow-only device array
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return PF_OK;
        size_t want = std::max(n, cap + cap / 2);
        T *q = nullptr;
        cudaError_t e = cudaMalloc(&q, want * sizeof(T));
        if (e != cudaSuccess) {
            cudaGetLastError();
            // retry with the exact size before giving up
            want = n;
            e = cudaMalloc(&q, want * sizeof(T));
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_error("device allocation of %zu bytes failed (frontier too large: use smaller read blocks)",
                          want * sizeof(T));
                return PF_ERR_NOMEM;
            }
        }
        if (p) cudaFree(p);
        p = q;
        cap = want;
        return PF_OK;
    }
    // grow while preserving the first `keep` elements (hit lists accumulate across levels)
    int grow_keep(size_t n, size_t keep, cudaStream_t s) {
        if (n <= cap) return PF_OK;
        T *old = p;
        p = nullptr;
        size_t old_cap = cap;
        cap = 0;
        int rc = ensure(std::max(n, old_cap + old_cap / 2));
        if (rc != PF_OK) {
            p = old;
            cap = old_cap;
            return rc;
        }
        if (old && keep) cudaMemcpyAsync(p, old, keep * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (old) {
            cudaStreamSynchronize(s);
            cudaFree(old);
        }
        return PF_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

template <class T>
struct PinnedBuf {  // grow-only pinned host array
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return PF_OK;
        const size_t want = std::max(n, cap + cap / 2);
        T *q = nullptr;
        if (cudaMallocHost(&q, want * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            set_error("pinned host allocation of %zu bytes failed", want * sizeof(T));
            return PF_ERR_NOMEM;
        }
        if (p) cudaFreeHost(p);
        p = q;
        cap = want;
        return PF_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace pf

namespace pf {
struct ShardState;
struct SlicedState;
}
using namespace pf;  // internal header: the handle types below live at global scope (C ABI)

constexpr int PF_SPLIT_CHUNK = -100;  // internal: the chunk's frontier outgrew frontier_cap, retry with fewer reads

struct pf_dev_batch {
    uint32_t n_reads = 0, n_exc = 0;
    uint64_t n_words = 0, exc_nbytes = 0;
    DevBuf<uint32_t> lengths, packed, exc_index;
    DevBuf<uint64_t> word_off, exc_off, kmer_off;
    DevBuf<uint32_t> kcnt;                 // scratch of the device-side k-mer prefix sum
    DevBuf<unsigned long long> kbsum;
    DevBuf<uint8_t> exc_bytes;
    std::vector<uint64_t> h_kmer_off;  // [n_reads + 1] prefix sum of k-mer counts (host copy for chunking)
    cudaEvent_t ready = nullptr;       // recorded after the H2D copies of an asynchronous upload
    uint64_t kmer_size = 0, max_kmers = 0;
    uint64_t total_bases_bound = 0;  // upper bound on the batch's k-mers when the prefix sum is device-only
    uint64_t nominal_kmers = 1;      // k-mers of a read of mean length: what the step plan is made for
    uint64_t bytes = 0;
    uint32_t max_length = 0;         // longest read in bases (header of the subtree-sharded exchange)
    uint64_t total_bases = 0;        // sum of lengths
    void release() {
        lengths.release();
        packed.release();
        exc_index.release();
        word_off.release();
        kmer_off.release();
        kcnt.release();
        kbsum.release();
        exc_off.release();
        exc_bytes.release();
        if (ready) cudaEventDestroy(ready);
        ready = nullptr;
    }
};

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
namespace pf {
extern NcclApi g_nccl;
int load_nccl();  // PF_OK or PF_ERR_NCCL (message set)
}
#define PF_NCCL_OK(expr)                                                                               \
    do {                                                                                               \
        ncclResult_t _r = (expr);                                                                      \
        if (_r != ncclSuccess) {                                                                       \
            pf::set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__,                               \
                          pf::g_nccl.GetErrorString ? pf::g_nccl.GetErrorString(_r) : "error");        \
            return PF_ERR_NCCL;                                                                        \
        }                                                                                              \
    } while (0)
struct pf_db {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the next batch overlaps the query of the current one
    int sm_count = 148;
    // host copy of the flattened tree (level order)
    HostTree tree;
    std::vector<uint32_t> h_left, h_right, h_slot;
    std::vector<int32_t> h_leaf;
    std::vector<int32_t> h_pre;         // level-order id -> index into tree.nodes
    std::vector<uint32_t> level_start;  // n_levels + 1
    std::vector<std::string> leaf_ids;  // DFS leaf order
    uint64_t n_nodes = 0, n_leaves = 0, n_slots = 0, wpf = 0;
    BfHeader geom;
    HashParams hp{};
    int exhaustive = 0;
    int lazy = 1;                      // step-limited pre-test at verified-monotone interior nodes
    std::vector<uint64_t> h_pop;       // set bits of each node's filter
    std::vector<uint8_t> h_mono;       // interior node whose filter contains both children's filters
    std::vector<uint32_t> h_steps;     // probe steps per node for the current (threshold, mode)
    std::vector<uint32_t> h_stride;    // k-mer sampling stride per node (1 = every k-mer)
    std::vector<uint32_t> level_min_stride;  // per level: smallest stride if every tested node is a 1-step sample, else 0
    uint32_t *d_steps = nullptr;
    std::vector<uint32_t> h_entry;        // entry nodes of the current plan, ordered by level
    std::vector<uint32_t> entry_start;    // [n_levels + 1] offsets into h_entry per level
    uint32_t *d_entry = nullptr;
    float steps_theta = -1.f;
    uint64_t steps_n = 0;  // nominal k-mers per read the plan was made for
    int steps_mode = -1;
    uint64_t n_internal = 0, n_monotone = 0;
    // k-mer memo of exact nodes (see ProbeArgs::memo): regions are handed out per level, zeroed before the level runs
    int memo = 1;
    uint64_t memo_budget_bytes = 256ULL << 20;
    std::vector<uint32_t> h_node_memo;          // per node: region index inside its level, or NONE32
    std::vector<uint32_t> level_memo_regions;   // per level: regions in use
    std::vector<uint32_t> level_memo_entries;   // per level: entries of all its regions / 4096, 0 = memo off
    std::vector<uint64_t> level_memo_kmers;     // per level: k-mers held by the filters of its memo nodes (set bits / K)
    uint32_t *d_node_memo = nullptr;
    DevBuf<unsigned long long> memo_table;
    // device tree
    uint32_t *d_left = nullptr, *d_right = nullptr, *d_slot = nullptr;
    int32_t *d_leaf = nullptr;
    uint64_t *d_filters = nullptr;
    // accumulators and per-block scratch
    unsigned long long *d_counts = nullptr, *d_blk_counts = nullptr, *d_blk_snapshot = nullptr;
    // largest frontier (pairs) one chunk of reads may produce; 32-bit pair indices and ticket counters that every
    // persistent warp bumps once more after the last pair (~2e5 x 32) leave this much room
    uint64_t frontier_cap = 0xFF000000ULL;
    uint32_t *d_node_pass = nullptr, *d_cursor = nullptr;  // contiguous [2 * n_nodes], then:
    uint32_t *d_node_pass_copies = nullptr;                // [NODE_PASS_COPIES * n_nodes] counters the probe kernel adds to
    unsigned long long *d_next_base = nullptr, *d_hit_base = nullptr;
    unsigned int *d_work = nullptr;  // one counter per level
    unsigned long long *d_probes = nullptr;
    LevelTotals *d_totals = nullptr, *h_totals = nullptr;
    DevBuf<uint32_t> fr_read[2], fr_node[2], hit_read, hit_leaf;
    DevBuf<uint8_t> pass;
    DevBuf<uint64_t> hb;                       // cached hash_bytes per k-mer of the current chunk
    DevBuf<uint32_t> idx0;                     // cached step-0 bit index per k-mer (m < 2^31)
    DevBuf<uint8_t> read_flag;                 // hybrid hand-over: reads that go on node by node
    uint64_t hash_cache_bytes = 16ULL << 30;   // chunk reads so the cache stays below this
    pf_dev_batch own_batch;  // device copy used by pf_query_block
    // outputs: per-read hit lists as CSR, built on the device, returned through pinned host arrays
    DevBuf<uint32_t> read_hits, csr_leaf;
    DevBuf<unsigned long long> csr_off, csr_bsum;
    PinnedBuf<uint64_t> pin_off;
    PinnedBuf<uint32_t> pin_leaf;
    std::vector<uint64_t> out_off;
    // timing
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_probe;  // 2 per level
    pf_stats_t stats{};
    ncclComm_t comm = nullptr;
    // ---- subtree sharding (pf_shard.cu): levels < cut_level are replicated ("top"), every node of level
    // cut_level roots a subtree owned by exactly one rank; only top + owned filters are resident.
    int sharded = 0, rank = 0, nranks = 1;
    uint32_t cut_level = 0;
    int64_t cut_level_req = -1;         // requested cut level (-1: chosen to minimise resident filters per rank)
    std::vector<int32_t> h_owner;      // per node: -1 = top (replicated), else owning rank
    std::vector<uint32_t> cut_lo;      // [nranks + 1] node-id boundaries of the owners' ranges inside level cut_level
    pf::ShardState *shard = nullptr;   // exchange buffers (pf_shard.cu)
    // ---- bit-sliced tiles (pf_sliced.cu): 0 = choose per (threshold, read length) by cost model, 1 = node-at-a-time
    // descent only, 2 = sliced tiles only
    int mode = 0;
    uint32_t tile_cols = 256;  // sliced path: most columns per tile (pf_db_set_tile_cols)
    int handover = -1;  // sliced path below the cut: 0 tiles all the way down, 1 hand-over to the node-at-a-time descent, -1 tiles if they fit
    // L2 residency (access-policy window on the stream): bytes of L2 set aside for persisting lines, largest window, and
    // per level the range of filter slots its nodes use
    uint64_t l2_persist_bytes = 0, l2_window_max = 0;
    int l2_persist_policy = 0;
    bool l2_window_set = false;
    std::vector<uint32_t> level_slot_lo, level_slot_hi;
    double related_share = -1.0;       // running estimate of the share of reads with at least one hit (-1: none yet)
    double plan_cost = 0.0;            // expected bit probes PER K-MER of a read unrelated to the database under the current step plan
    pf::SlicedState *sliced = nullptr;
    // hand-over from the tiles to the node-at-a-time descent: (read, node) pairs, node-major, cut per level
    const uint32_t *inj_read = nullptr, *inj_node = nullptr;
    std::vector<uint64_t> inj_level_off;  // [n_levels + 1]; empty or all-equal: nothing to inject
    std::vector<uint8_t> nccl_id;      // ncclUniqueId handed to pf_db_open_sharded (the analysis at open is collective)
};

namespace pf {
// One descent over levels [l_begin, l_end) of the flattened tree: entry nodes of each level are injected for the
// reads [inj_r0, inj_r0 + inj_n), every level is probed, scanned and scattered (query.rs:99-158).
struct Descent {
    uint64_t n = 0;  // pairs in the current frontier
    int cur = 0;     // ping-pong buffer holding it
    uint64_t hits_total = 0, hits_before = 0, probes = 0, pairs = 0, levels = 0, probe_launches = 0, other_launches = 0;
    uint64_t memo_hits = 0, memo_lookups = 0;
    uint64_t sectors = 0, sliced_pairs = 0;  // row loads and (read, tile) pairs of the sliced kernel
    uint64_t lines = 0;                      // 128-byte line loads of the entry line kernel (not in `sectors`)
    size_t n_ev = 0;
    std::vector<uint8_t> ev_sliced;          // per event pair: recorded around a sliced-kernel launch
};
int run_levels(pf_db *db, const pf_dev_batch *bt, float threshold, int want_hits, uint32_t G, uint64_t kmer_base,
               size_t l_begin, size_t l_end, uint32_t inj_r0, uint32_t inj_n, Descent &st);
int update_steps(pf_db *db, float threshold, uint64_t n_nominal);
int batch_upload_impl(pf_db *db, const pf_read_batch *in, pf_dev_batch *b, cudaStream_t s);
void launch_hash(const HashArgs &a, int grid, cudaStream_t s);
uint32_t group_rounds_for(uint64_t max_kmers, bool small_m);
uint32_t hash_grab(uint64_t max_kmers);
int db_open_impl(pf_db *db, const char *db_path, int64_t search_depth);
void db_free(pf_db *db);
void flatten(pf_db *db, int64_t search_depth);     // prune_tree + level-order numbering (host only)
int comm_init_impl(pf_db *db, int nranks, int rank, const void *id128);
struct ShardPlan {
    uint32_t cut_level = 0;
    std::vector<int32_t> owner;    // per node: -1 = replicated top, else owning rank
    std::vector<uint32_t> cut_lo;  // [nranks + 1] owners' node-id ranges inside the cut level
    uint64_t top_nodes = 0, max_owned = 0;
};
void plan_shards(const std::vector<uint32_t> &level_start, const std::vector<uint32_t> &left,
                 const std::vector<uint32_t> &right, int nranks, int64_t cut_req, ShardPlan &out);
void shard_free(pf_db *db);                       // pf_shard.cu
int shard_plan(pf_db *db, int64_t cut_level_req);  // pf_shard.cu: cut level, owners, resident slots (after flatten)
int finish_csr(pf_db *db, uint32_t n_reads, uint64_t hits_total, int want_hits, uint32_t out_r0, uint32_t out_n,
               pf_hits *out, uint64_t *other_launches, uint64_t *d2h);
void account_stats(pf_db *db, const Descent &st, uint64_t n_reads, uint64_t d2h);
// pf_sliced.cu
void sliced_free(pf_db *db);
int sliced_prepare(pf_db *db, float threshold, uint64_t n_nominal, bool *use_sliced);
int sliced_begin_block(pf_db *db);
int sliced_set_hit_cursor(pf_db *db, uint64_t hits);
uint64_t sliced_entry_tiles(const pf_db *db);
bool sliced_hybrid(const pf_db *db);
// hf: null, or the chunk's hash-kernel arguments when the entry line kernel hashes on the fly (sliced_fused) -- the hash
// values of the reads that survive the entry depth are then made here, between the entry depth and the next one
int run_sliced(pf_db *db, const pf_dev_batch *bt, float threshold, int want_hits, uint64_t kmer_base, uint32_t r0,
               uint32_t n_chunk, Descent &st, const HashArgs *hf);
bool sliced_fused(const pf_db *db, const pf_dev_batch *bt);
}  // namespace pf
