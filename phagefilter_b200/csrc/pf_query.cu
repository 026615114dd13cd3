// pf_query.cu -- pf_db: flattened device-resident gSBT and the level-synchronous query.
//
// Replaces, behind the C ABI of include/pfgpu.h:
//   BloomTree::load / prune_tree        bloom_tree.rs:364-386, 302-330
//   BFLruCache::get_filter              cache.rs:56-77      (all filters resident in HBM instead)
//   query::query_batch / _query_batch   query.rs:66-158
//   query::save_leaf_counts             query.rs:173-218
#include <dlfcn.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>

#include "pf_db.h"

namespace pf {

static thread_local std::string g_error;
void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
}

// pf_build.cu
int build_leaf_filter(const uint8_t *h_seq, uint64_t len, const HashParams &hp, uint64_t *d_filter, uint64_t wpf,
                      cudaStream_t s);
int filters_equal(const uint64_t *a, const uint64_t *b, uint64_t n_words, cudaStream_t s, bool *equal);

}  // namespace pf

using namespace pf;

namespace pf {
NcclApi g_nccl;
}
int pf::load_nccl() {
    if (g_nccl.h) return PF_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("NCCL not found: %s", dlerror());
        return PF_ERR_NCCL;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.Broadcast = (decltype(g_nccl.Broadcast))dlsym(h, "ncclBroadcast");
    g_nccl.Send = (decltype(g_nccl.Send))dlsym(h, "ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(h, "ncclRecv");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.AllGather ||
        !g_nccl.Broadcast || !g_nccl.Send || !g_nccl.Recv || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        set_error("NCCL symbols missing");
        return PF_ERR_NCCL;
    }
    g_nccl.h = h;
    return PF_OK;
}

void pf::db_free(pf_db *db) {
    if (!db) return;
    cudaSetDevice(db->device);
    shard_free(db);
    sliced_free(db);
    if (db->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(db->comm);
    cudaFree(db->d_left);
    cudaFree(db->d_right);
    cudaFree(db->d_slot);
    cudaFree(db->d_leaf);
    cudaFree(db->d_filters);
    cudaFree(db->d_steps);
    cudaFree(db->d_entry);
    cudaFree(db->d_counts);
    cudaFree(db->d_blk_counts);
    cudaFree(db->d_blk_snapshot);
    cudaFree(db->d_node_pass);
    cudaFree(db->d_next_base);
    cudaFree(db->d_hit_base);
    cudaFree(db->d_work);
    cudaFree(db->d_probes);
    cudaFree(db->d_node_memo);
    db->memo_table.release();
    cudaFree(db->d_totals);
    if (db->h_totals) cudaFreeHost(db->h_totals);
    for (int i = 0; i < 2; i++) {
        db->fr_read[i].release();
        db->fr_node[i].release();
    }
    db->hit_read.release();
    db->hit_leaf.release();
    db->pass.release();
    db->hb.release();
    db->idx0.release();
    db->read_flag.release();
    db->read_hits.release();
    db->csr_leaf.release();
    db->csr_off.release();
    db->csr_bsum.release();
    db->pin_off.release();
    db->pin_leaf.release();
    db->own_batch.release();
    if (db->ev_begin) cudaEventDestroy(db->ev_begin);
    if (db->ev_end) cudaEventDestroy(db->ev_end);
    for (auto e : db->ev_probe) cudaEventDestroy(e);
    if (db->stream) cudaStreamDestroy(db->stream);
    if (db->copy_stream) cudaStreamDestroy(db->copy_stream);
    delete db;
}

// prune_tree (bloom_tree.rs:302-330) + level-order flattening + DFS leaf numbering.
void pf::flatten(pf_db *db, int64_t search_depth) {
    const HostTree &t = db->tree;
    db->h_left.clear();
    db->h_right.clear();
    db->h_slot.clear();
    db->h_leaf.clear();
    db->level_start.clear();
    db->leaf_ids.clear();
    if (t.root < 0) {
        db->n_nodes = db->n_leaves = 0;
        db->level_start.push_back(0);
        return;
    }
    // level-order ids; `keep_children[pre]` is false for nodes cut by the search depth
    std::vector<int32_t> order;               // level-order -> pre-order index
    std::vector<uint32_t> depth_of;           // by level-order id
    std::vector<int32_t> bfs_of(t.nodes.size(), -1);
    order.push_back(t.root);
    depth_of.push_back(0);
    bfs_of[t.root] = 0;
    db->level_start.push_back(0);
    for (size_t q = 0; q < order.size(); ++q) {
        const HostNode &n = t.nodes[order[q]];
        uint32_t d = depth_of[q];
        if (q > 0 && d != depth_of[q - 1]) db->level_start.push_back((uint32_t)q);
        bool cut = search_depth >= 0 && (int64_t)d >= search_depth;
        if (cut) continue;
        for (int32_t c : {n.left, n.right}) {
            if (c < 0) continue;
            bfs_of[c] = (int32_t)order.size();
            order.push_back(c);
            depth_of.push_back(d + 1);
        }
    }
    db->level_start.push_back((uint32_t)order.size());
    db->n_nodes = order.size();
    db->h_left.assign(db->n_nodes, NONE32);
    db->h_right.assign(db->n_nodes, NONE32);
    db->h_leaf.assign(db->n_nodes, -1);
    for (size_t q = 0; q < order.size(); ++q) {
        const HostNode &n = t.nodes[order[q]];
        bool cut = search_depth >= 0 && (int64_t)depth_of[q] >= search_depth;
        if (!cut) {
            if (n.left >= 0) db->h_left[q] = (uint32_t)bfs_of[n.left];
            if (n.right >= 0) db->h_right[q] = (uint32_t)bfs_of[n.right];
        }
    }
    // left-first DFS leaf order (query.rs:197-218)
    std::vector<uint32_t> st{0};
    while (!st.empty()) {
        uint32_t u = st.back();
        st.pop_back();
        if (db->h_left[u] == NONE32 && db->h_right[u] == NONE32) {  // is_leafnode, bloom_tree.rs:416-418
            db->h_leaf[u] = (int32_t)db->leaf_ids.size();
            db->leaf_ids.push_back(t.nodes[order[u]].tax_id);
            continue;
        }
        if (db->h_right[u] != NONE32) st.push_back(db->h_right[u]);
        if (db->h_left[u] != NONE32) st.push_back(db->h_left[u]);
    }
    db->n_leaves = db->leaf_ids.size();
    db->h_pre = order;
    // filter slots: one per distinct path string (cache.rs keys filters by path)
    std::map<std::string, uint32_t> slot_of;
    db->h_slot.resize(db->n_nodes);
    for (size_t q = 0; q < order.size(); ++q) {
        const std::string &p = t.nodes[order[q]].bf_path;
        auto it = slot_of.find(p);
        if (it == slot_of.end()) it = slot_of.emplace(p, (uint32_t)slot_of.size()).first;
        db->h_slot[q] = it->second;
    }
    db->n_slots = slot_of.size();
}


// slots are numbered in level order, so the filters of one level are (nearly) one contiguous range of the buffer
static void level_slot_ranges(pf_db *db) {
    const size_t n_levels = db->level_start.size() - 1;
    db->level_slot_lo.assign(n_levels, NONE32);
    db->level_slot_hi.assign(n_levels, 0);
    for (size_t l = 0; l < n_levels; ++l)
        for (uint32_t u = db->level_start[l]; u < db->level_start[l + 1]; ++u) {
            if (db->h_slot[u] == NONE32) continue;  // subtree shards: filter held by another rank
            db->level_slot_lo[l] = std::min(db->level_slot_lo[l], db->h_slot[u]);
            db->level_slot_hi[l] = std::max(db->level_slot_hi[l], db->h_slot[u]);
        }
}

template <class T>
static int upload(T **dst, const std::vector<T> &v, cudaStream_t s) {
    PF_CUDA_OK(cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) PF_CUDA_OK(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return PF_OK;
}

static int analyse_tree(pf_db *db);

namespace {
struct OpenTiming {  // PF_TIMING=1: where pf_db_open spends its time (stderr)
    bool on = getenv("PF_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), t = t0;
    std::string line;
    void lap(const char *name) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof b, " %s=%.1f", name, std::chrono::duration<double, std::milli>(n - t).count());
        line += b;
        t = n;
    }
    void report(const pf_db *db) {
        if (!on) return;
        fprintf(stderr, "[pf_db_open] total=%.1f ms%s (%llu filters, %.1f MB)\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), line.c_str(),
                (unsigned long long)db->n_slots, db->n_slots * db->wpf * 8 / 1e6);
    }
};
}  // namespace

int pf::db_open_impl(pf_db *db, const char *db_path, int64_t search_depth) {
    OpenTiming timing;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libpfgpu has no CPU fallback");
        return PF_ERR_CUDA;
    }
    if (db->device < 0 || db->device >= ndev) {
        set_error("device %d out of range (%d devices)", db->device, ndev);
        return PF_ERR_ARG;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    PF_CUDA_OK(cudaStreamCreateWithFlags(&db->stream, cudaStreamNonBlocking));
    PF_CUDA_OK(cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking));
    PF_CUDA_OK(cudaDeviceGetAttribute(&db->sm_count, cudaDevAttrMultiProcessorCount, db->device));
    {   // L2 set-aside for persisting accesses: the filters a level is probing are pinned there with an access-policy
        // window on the stream (run_levels), everything streamed is loaded evict-first (pf_kernels.cuh)
        int persist = 0, window = 0;
        // PF_L2_PERSIST: 0 (default) no window; 1: levels whose filters fit the set-aside; 2: every level (hitRatio < 1).
        // Measured on BASELINE config 2 (profiles/r2h_*): the filters being probed are L2 hits already (the frontier is
        // node-major; what misses is the compulsory streaming of frontier / index / hash arrays, loaded evict-first), so the
        // window changes neither hit rate nor DRAM bytes, while the set-aside takes L2 from the leaf levels whose filters
        // together exceed it: 4.40 ms per step without, 5.71 (policy 1) and 5.62 (policy 2) with.
        const char *off = getenv("PF_L2_PERSIST");
        db->l2_persist_policy = off ? atoi(off) : 0;
        if (db->l2_persist_policy != 0 && cudaDeviceGetAttribute(&persist, cudaDevAttrMaxPersistingL2CacheSize, db->device) == cudaSuccess &&
            cudaDeviceGetAttribute(&window, cudaDevAttrMaxAccessPolicyWindowSize, db->device) == cudaSuccess && persist > 0 &&
            window > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist) == cudaSuccess) {
            db->l2_persist_bytes = (uint64_t)persist;
            db->l2_window_max = (uint64_t)window;
        }
        cudaGetLastError();
    }
    timing.lap("cuda_init");
    if (db->sharded) {  // the load-time analysis of a sharded tree ends in an all-reduce
        int crc = comm_init_impl(db, db->nranks, db->rank, db->nccl_id.data());
        if (crc != PF_OK) return crc;
    }
    std::string err;
    std::string dir(db_path);
    if (!read_tree_bin(join_path(dir, "tree.bin"), db->tree, err)) {
        set_error("%s", err.c_str());
        return err.rfind("cannot open", 0) == 0 ? PF_ERR_IO : PF_ERR_FORMAT;
    }
    flatten(db, search_depth);
    if (db->n_nodes == 0) {
        set_error("database has no root node (reference panics in save_leaf_counts, main.rs:374)");
        return PF_ERR_FORMAT;
    }
    db->h_owner.assign(db->n_nodes, -1);
    if (db->sharded) {  // subtree shards: choose the cut, the owners, and keep only top + owned filters resident
        int prc = shard_plan(db, db->cut_level_req);
        if (prc != PF_OK) return prc;
    }
    level_slot_ranges(db);
    // decode every distinct filter once; geometry must be uniform
    std::vector<std::string> slot_path(db->n_slots);
    for (size_t q = 0; q < db->n_nodes; ++q)
        if (db->h_slot[q] != NONE32) slot_path[db->h_slot[q]] = db->tree.nodes[db->h_pre[q]].bf_path;
    BfHeader first{};
    if (!read_bf_header(join_path(dir, slot_path[0]), first, err)) {
        set_error("%s", err.c_str());
        return err.rfind("Failed to open", 0) == 0 ? PF_ERR_IO : PF_ERR_FORMAT;
    }
    if (first.num_bits == 0) {
        set_error("filter with zero bits (the reference would divide by zero in contains)");
        return PF_ERR_FORMAT;
    }
    db->geom = first;
    db->wpf = (first.n_words + 15) / 16 * 16;  // 128-byte aligned slots
    size_t total_words = (size_t)db->n_slots * db->wpf;
    cudaError_t ce = cudaMalloc(&db->d_filters, total_words * 8);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        set_error("cannot hold %zu filters (%.1f GB) in device memory", (size_t)db->n_slots, total_words * 8 / 1e9);
        return PF_ERR_NOMEM;
    }
    timing.lap("tree+alloc");
    // Filters are decoded by several host threads into two half-rings of pinned staging buffers: while the copies
    // of one wave run, the next wave is read into the other half.
    const unsigned hc = std::thread::hardware_concurrency();
    const uint64_t wave = std::min<uint64_t>(std::max<uint64_t>((32u << 20) / (db->wpf * 8), 1), std::min<uint64_t>(hc ? hc : 4, 16));
    uint64_t *stage = nullptr;
    cudaEvent_t staged[2] = {nullptr, nullptr};
    PF_CUDA_OK(cudaMallocHost(&stage, 2 * wave * db->wpf * 8));
    for (int i = 0; i < 2; i++) PF_CUDA_OK(cudaEventCreateWithFlags(&staged[i], cudaEventDisableTiming));
    int rc = PF_OK;
    std::vector<std::string> errs(wave);
    std::vector<BfHeader> hdrs(wave);
    for (uint64_t s0 = 0, w = 0; s0 < db->n_slots && rc == PF_OK; s0 += wave, ++w) {
        const uint64_t cnt = std::min<uint64_t>(wave, db->n_slots - s0);
        uint64_t *half = stage + (w & 1) * wave * db->wpf;
        cudaEventSynchronize(staged[w & 1]);
        auto load = [&](uint64_t i) {
            uint64_t *dst = half + i * db->wpf;
            errs[i].clear();
            if (read_bf(join_path(dir, slot_path[s0 + i]), hdrs[i], dst, db->wpf, errs[i]))
                memset(dst + hdrs[i].n_words, 0, (db->wpf - hdrs[i].n_words) * 8);  // pad words of the 128-byte slot
        };
        std::vector<std::thread> th;
        for (uint64_t i = 1; i < cnt; ++i) th.emplace_back(load, i);
        load(0);
        for (auto &t : th) t.join();
        for (uint64_t i = 0; i < cnt && rc == PF_OK; ++i) {
            const BfHeader &h = hdrs[i];
            if (!errs[i].empty()) {
                set_error("%s", errs[i].c_str());
                rc = errs[i].rfind("Failed to open", 0) == 0 ? PF_ERR_IO : PF_ERR_FORMAT;
            } else if (h.num_bits != first.num_bits || h.num_hashes != first.num_hashes || h.seed1 != first.seed1 || h.seed2 != first.seed2) {
                set_error("filter %s differs in geometry/seeds from the rest of the database", slot_path[s0 + i].c_str());
                rc = PF_ERR_FORMAT;
            }
        }
        if (rc != PF_OK) break;
        cudaMemcpyAsync(db->d_filters + s0 * db->wpf, half, cnt * db->wpf * 8, cudaMemcpyHostToDevice, db->stream);
        cudaEventRecord(staged[w & 1], db->stream);
    }
    cudaStreamSynchronize(db->stream);
    cudaFreeHost(stage);
    for (int i = 0; i < 2; i++) cudaEventDestroy(staged[i]);
    timing.lap("filters");
    if (rc != PF_OK) return rc;
    PF_CUDA_OK(cudaGetLastError());

    db->hp = make_hash_params(first.seed1, first.seed2, db->tree.kmer_size, first.num_bits, first.num_hashes, 26);
    if ((rc = upload(&db->d_left, db->h_left, db->stream))) return rc;
    if ((rc = upload(&db->d_right, db->h_right, db->stream))) return rc;
    if ((rc = upload(&db->d_slot, db->h_slot, db->stream))) return rc;
    if ((rc = upload(&db->d_leaf, db->h_leaf, db->stream))) return rc;
    size_t nn = db->n_nodes, nl = std::max<uint64_t>(db->n_leaves, 1), nlev = db->level_start.size();
    PF_CUDA_OK(cudaMalloc(&db->d_counts, nl * 8));
    PF_CUDA_OK(cudaMalloc(&db->d_blk_counts, nl * 8));
    PF_CUDA_OK(cudaMalloc(&db->d_blk_snapshot, nl * 8));
    // [node totals | cursors | NODE_PASS_COPIES x per-node counters], zeroed together per chunk of reads
    PF_CUDA_OK(cudaMalloc(&db->d_node_pass, (2 + NODE_PASS_COPIES) * nn * 4));
    db->d_cursor = db->d_node_pass + nn;
    db->d_node_pass_copies = db->d_node_pass + 2 * nn;
    PF_CUDA_OK(cudaMalloc(&db->d_next_base, nn * 8));
    PF_CUDA_OK(cudaMalloc(&db->d_hit_base, nn * 8));
    PF_CUDA_OK(cudaMalloc(&db->d_work, nlev * 4));
    PF_CUDA_OK(cudaMalloc(&db->d_probes, 24));  // [probes issued, k-mers answered by the memo, memo look-ups]
    PF_CUDA_OK(cudaMalloc(&db->d_node_memo, nn * 4));
    PF_CUDA_OK(cudaMalloc(&db->d_totals, sizeof(LevelTotals)));
    PF_CUDA_OK(cudaMallocHost(&db->h_totals, sizeof(LevelTotals)));
    PF_CUDA_OK(cudaMemsetAsync(db->d_counts, 0, nl * 8, db->stream));
    PF_CUDA_OK(cudaMalloc(&db->d_steps, nn * 4));
    PF_CUDA_OK(cudaMalloc(&db->d_entry, nn * 4));
    timing.lap("tables");
    if ((rc = analyse_tree(db))) return rc;
    PF_CUDA_OK(cudaEventCreate(&db->ev_begin));
    PF_CUDA_OK(cudaEventCreate(&db->ev_end));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    timing.lap("analyse");
    timing.report(db);
    return PF_OK;
}

// Load-time analysis: fill of every filter and, per interior node, whether its filter is a bitwise
// superset of its children's.  Where that holds, "child passes => parent passes" for every read and
// threshold (a k-mer contained in the child has all K bits set in the parent too), so the parent only
// needs a sound cannot-pass test; the reference's u16 name collisions (bloom_tree.rs:232-234) can break
// the superset property, and such nodes stay exact.
static int analyse_tree(pf_db *db) {
    const size_t nn = db->n_nodes;
    // Who measures what.  Replicated tree: this rank measures everything.  Subtree shards: a node's fill is
    // measured by its owner (rank 0 for the replicated top), a (node, child) superset check by the child's owner
    // (which also holds the parent); one all-reduce(sum) then gives every rank the same global tables, so every
    // rank derives the same step plan.
    auto mine = [&](uint32_t u) {
        const int32_t o = db->h_owner[u];
        return o == db->rank || (o < 0 && db->rank == 0);
    };
    std::vector<uint32_t> chk_left(nn, NONE32), chk_right(nn, NONE32);
    for (size_t u = 0; u < nn; ++u) {
        if (db->h_left[u] != NONE32 && mine(db->h_left[u])) chk_left[u] = db->h_left[u];
        if (db->h_right[u] != NONE32 && mine(db->h_right[u])) chk_right[u] = db->h_right[u];
    }
    unsigned long long *d_pop = nullptr, *d_nodeinfo = nullptr;  // d_nodeinfo: [0,nn) violations, [nn,2nn) fill
    uint32_t *d_chk = nullptr;
    PF_CUDA_OK(cudaMalloc(&d_pop, std::max<size_t>(db->n_slots, 1) * 8));
    PF_CUDA_OK(cudaMalloc(&d_nodeinfo, 2 * nn * 8));
    PF_CUDA_OK(cudaMalloc(&d_chk, 2 * nn * 4));
    PF_CUDA_OK(cudaMemsetAsync(d_pop, 0, std::max<size_t>(db->n_slots, 1) * 8, db->stream));
    PF_CUDA_OK(cudaMemsetAsync(d_nodeinfo, 0, 2 * nn * 8, db->stream));
    PF_CUDA_OK(cudaMemcpyAsync(d_chk, chk_left.data(), nn * 4, cudaMemcpyHostToDevice, db->stream));
    PF_CUDA_OK(cudaMemcpyAsync(d_chk + nn, chk_right.data(), nn * 4, cudaMemcpyHostToDevice, db->stream));
    const uint32_t bx = (uint32_t)std::min<uint64_t>((db->wpf + 1023) / 1024, 64);
    for (uint64_t s0 = 0; s0 < db->n_slots; s0 += 65535) {
        const uint32_t ny = (uint32_t)std::min<uint64_t>(65535, db->n_slots - s0);
        fill_kernel<<<dim3(bx, ny), 256, 0, db->stream>>>(db->d_filters, db->wpf, (uint32_t)s0, d_pop);
    }
    for (uint64_t u0 = 0; u0 < nn; u0 += 65535) {
        const uint32_t ny = (uint32_t)std::min<uint64_t>(65535, nn - u0);
        subset_kernel<<<dim3(bx, ny), 256, 0, db->stream>>>(db->d_filters, db->wpf, db->d_slot, d_chk, d_chk + nn,
                                                           (uint32_t)u0, d_nodeinfo);
    }
    std::vector<unsigned long long> pop(std::max<size_t>(db->n_slots, 1)), info(2 * nn, 0);
    cudaMemcpyAsync(pop.data(), d_pop, db->n_slots * 8, cudaMemcpyDeviceToHost, db->stream);
    cudaError_t e = cudaStreamSynchronize(db->stream);
    if (e == cudaSuccess && db->sharded) {
        for (size_t u = 0; u < nn; ++u) info[nn + u] = mine((uint32_t)u) ? pop[db->h_slot[u]] : 0;
        cudaMemcpyAsync(d_nodeinfo + nn, info.data() + nn, nn * 8, cudaMemcpyHostToDevice, db->stream);
        ncclResult_t r = g_nccl.AllReduce(d_nodeinfo, d_nodeinfo, 2 * nn, ncclUint64, ncclSum, db->comm, db->stream);
        if (r != ncclSuccess) {
            cudaFree(d_pop);
            cudaFree(d_nodeinfo);
            cudaFree(d_chk);
            set_error("ncclAllReduce (tree analysis): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
            return PF_ERR_NCCL;
        }
    }
    if (e == cudaSuccess) {
        cudaMemcpyAsync(info.data(), d_nodeinfo, 2 * nn * 8, cudaMemcpyDeviceToHost, db->stream);
        e = cudaStreamSynchronize(db->stream);
    }
    cudaFree(d_pop);
    cudaFree(d_nodeinfo);
    cudaFree(d_chk);
    PF_CUDA_OK(e);
    PF_CUDA_OK(cudaGetLastError());
    db->h_pop.resize(nn);
    db->h_mono.assign(nn, 0);
    db->n_internal = db->n_monotone = 0;
    for (size_t u = 0; u < nn; ++u) {
        db->h_pop[u] = db->sharded ? info[nn + u] : pop[db->h_slot[u]];
        if (db->h_leaf[u] >= 0) continue;
        db->n_internal++;
        db->h_mono[u] = info[u] == 0;
        db->n_monotone += db->h_mono[u];
    }
    return PF_OK;
}

// Probe steps per node.  Leaves, unverified nodes and the reference-faithful modes use all K steps (exact).
// A verified-monotone interior node never has to CONFIRM a pass (its children imply it); it is only worth
// probing if that prunes reads that do not belong below it.
/* ---- expected-cost step plan -------------------------------------------------------------------------
 * Only +,-,*,/ and sqrt on doubles and a fixed table are used, so the CPU checker's independent restatement of
 * this planner reproduces the table bit for bit (the tests compare the kernel's work counts with it).
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

// (threshold * n as f32).ceil() as usize  (query.rs:48), on the host
static uint64_t host_need(float threshold, uint64_t n) {
    volatile float prod = threshold * (float)n;
    const float c = ceilf(prod);
    if (!(c > 0.0f)) return 0;
    if (c >= 18446744073709551616.0f) return ~0ULL;
    return (uint64_t)c;
}
static void plan_steps(pf_db *db, float threshold, uint64_t n_nominal, std::vector<uint32_t> &steps,
                       std::vector<uint32_t> &strides) {
    const uint32_t K = db->geom.num_hashes;
    const size_t nn = db->n_nodes;
    const uint64_t need = host_need(threshold, n_nominal);
    const double n = (double)n_nominal, allowed = need > n_nominal ? 0.0 : (double)(n_nominal - need);
    std::vector<double> cost(nn, 0.0);
    steps.assign(nn, K);
    strides.assign(nn, 1);
    for (size_t u = nn; u-- > 0;) {  // children have larger level-order ids than their parent
        const double f = (double)db->h_pop[u] / (double)db->geom.num_bits;
        const uint32_t l = db->h_left[u], r = db->h_right[u];
        if (db->h_leaf[u] >= 0) {
            cost[u] = plan_probe_cost(f, K) + PF_PLAN_PAIR_OVERHEAD;
            continue;
        }
        const double below = (l != NONE32 ? cost[l] : 0.0) + (r != NONE32 ? cost[r] : 0.0);
        if (!db->h_mono[u]) {
            steps[u] = K;
            cost[u] = plan_exact_cost(f, K, n, allowed, below);
        } else {
            steps[u] = plan_choose(f, K, n_nominal, allowed, below, &strides[u], &cost[u]);
        }
    }
    db->plan_cost = nn ? cost[0] : 0.0;
}
int pf::update_steps(pf_db *db, float threshold, uint64_t n_nominal) {
    const int mode = db->exhaustive ? 2 : (db->lazy ? 1 : 0);
    if (mode == db->steps_mode && (mode != 1 || (threshold == db->steps_theta && n_nominal == db->steps_n))) return PF_OK;
    const uint32_t K = db->geom.num_hashes;
    db->h_steps.assign(db->n_nodes, K);
    db->h_stride.assign(db->n_nodes, 1);
    if (mode == 1) plan_steps(db, threshold, n_nominal, db->h_steps, db->h_stride);
    // entry nodes: descend from the root through skipped (0-step) interior nodes; level-order ids are already
    // sorted by level, so a sorted list is grouped by level
    db->h_entry.clear();
    {
        std::vector<uint32_t> st{0};
        while (!st.empty()) {
            const uint32_t u = st.back();
            st.pop_back();
            if (db->h_steps[u] == 0 && db->h_leaf[u] < 0) {
                if (db->h_left[u] != NONE32) st.push_back(db->h_left[u]);
                if (db->h_right[u] != NONE32) st.push_back(db->h_right[u]);
            } else {
                db->h_entry.push_back(u);
            }
        }
        // subtree shards: below the cut only the owner of a subtree injects at its entry nodes (for the reads of
        // every rank); above the cut each rank injects its own reads
        if (db->sharded) {
            size_t w = 0;
            for (uint32_t u : db->h_entry)
                if (db->h_owner[u] < 0 || db->h_owner[u] == db->rank) db->h_entry[w++] = u;
            db->h_entry.resize(w);
        }
        std::sort(db->h_entry.begin(), db->h_entry.end());
    }
    const size_t n_levels = db->level_start.size() - 1;
    db->entry_start.assign(n_levels + 1, 0);
    for (size_t l = 0, e = 0; l < n_levels; ++l) {
        while (e < db->h_entry.size() && db->h_entry[e] < db->level_start[l + 1]) ++e;
        db->entry_start[l + 1] = (uint32_t)e;
    }
    // device plan: steps | stride << 8 (stride 1 = every k-mer)
    // per level: the smallest sampling stride over its tested nodes if ALL of them are single-step samples, else 0
    db->level_min_stride.assign(n_levels, 0);
    for (size_t l = 0; l < n_levels && mode == 1; ++l) {
        uint32_t mn = 0xFFFFFFFFu;
        for (uint32_t u = db->level_start[l]; u < db->level_start[l + 1]; ++u) {
            if (db->h_steps[u] == 0 && db->h_leaf[u] < 0) continue;  // skipped: settled without probes
            if (db->h_steps[u] != 1 || db->h_stride[u] < 2) {
                mn = 0;
                break;
            }
            mn = std::min(mn, db->h_stride[u]);
        }
        db->level_min_stride[l] = mn == 0xFFFFFFFFu ? 0 : mn;
    }
    // memo regions: every exact multi-step node of a level gets a power-of-two region of about four times the k-mers its
    // filter holds (set bits / K), 2^12 .. 2^18 entries of 8 B; a level whose regions exceed the budget runs without
    {
        db->h_node_memo.assign(db->n_nodes, NONE32);
        db->level_memo_regions.assign(n_levels, 0);
        db->level_memo_entries.assign(n_levels, 0);
        db->level_memo_kmers.assign(n_levels, 0);
        for (size_t l = 0; l < n_levels && db->memo && mode != 2 && K > 1; ++l) {
            uint64_t total = 0, kmers = 0;
            uint32_t cnt = 0;
            auto log2_entries = [&](uint32_t u) {
                const uint64_t want = 4 * (db->h_pop[u] / K);
                uint32_t lg = 12;
                while (lg < 18 && (1ULL << lg) < want) ++lg;
                return lg;
            };
            for (uint32_t u = db->level_start[l]; u < db->level_start[l + 1]; ++u)
                if (db->h_steps[u] == K && db->h_slot[u] != NONE32) {
                    total += 1ULL << log2_entries(u);
                    kmers += db->h_pop[u] / K;
                    ++cnt;
                }
            if (!cnt || total * 8 > db->memo_budget_bytes || (total >> 12) >= (1ULL << 27)) continue;
            uint64_t off = 0;
            for (uint32_t u = db->level_start[l]; u < db->level_start[l + 1]; ++u)
                if (db->h_steps[u] == K && db->h_slot[u] != NONE32) {
                    const uint32_t lg = log2_entries(u);
                    db->h_node_memo[u] = (uint32_t)((off >> 12) << 5) | lg;
                    off += 1ULL << lg;
                }
            db->level_memo_regions[l] = cnt;
            db->level_memo_kmers[l] = kmers;
            db->level_memo_entries[l] = (uint32_t)(total >> 12);  // in units of 4096 entries
        }
        PF_CUDA_OK(cudaMemcpyAsync(db->d_node_memo, db->h_node_memo.data(), db->n_nodes * 4, cudaMemcpyHostToDevice, db->stream));
    }
    std::vector<uint32_t> enc(db->n_nodes);
    for (size_t u = 0; u < db->n_nodes; ++u) enc[u] = db->h_steps[u] | (db->h_stride[u] << 8);
    PF_CUDA_OK(cudaMemcpyAsync(db->d_steps, enc.data(), db->n_nodes * 4, cudaMemcpyHostToDevice, db->stream));
    PF_CUDA_OK(cudaMemcpyAsync(db->d_entry, db->h_entry.data(), db->h_entry.size() * 4, cudaMemcpyHostToDevice, db->stream));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    db->steps_mode = mode;
    db->steps_theta = threshold;
    db->steps_n = n_nominal;
    return PF_OK;
}

// ---- kernel dispatch ---------------------------------------------------------------------------
static bool fast_path_ok(const HashParams &hp) { return hp.k >= 17 && hp.k <= 32; }
template <int KM>
static void launch_hash_k(const HashArgs &a, int grid, cudaStream_t s) {
    hash_kernel<KM><<<grid, HASH_THREADS, 0, s>>>(a);
}
void pf::launch_hash(const HashArgs &a, int grid, cudaStream_t s) {
    switch (a.k) {
#define PF_CASE(K) \
    case K:        \
        launch_hash_k<K>(a, grid, s); \
        break;
        PF_CASE(17) PF_CASE(18) PF_CASE(19) PF_CASE(20) PF_CASE(21) PF_CASE(22) PF_CASE(23) PF_CASE(24)
        PF_CASE(25) PF_CASE(26) PF_CASE(27) PF_CASE(28) PF_CASE(29) PF_CASE(30) PF_CASE(31) PF_CASE(32)
#undef PF_CASE
        default:
            launch_hash_k<0>(a, grid, s);  // byte path for every other k
    }
}
// Rounds of 32 k-mers a lane owns at once: enough to cover the longest read of the batch, at most 8.
// reads per ticket of the hash kernel's work counter: many short reads per ticket, long reads one by one
uint32_t pf::hash_grab(uint64_t max_kmers) { return max_kmers <= 256 ? 16u : (max_kmers <= 2048 ? 4u : 1u); }
uint32_t pf::group_rounds_for(uint64_t max_kmers, bool small_m) {
    if (!small_m) return 1;  // 64-bit remainder path (m >= 2^31): kept simple
    return (uint32_t)std::min<uint64_t>(8, std::max<uint64_t>(1, (max_kmers + 31) / 32));
}
static void launch_probe(const ProbeArgs &a, uint32_t G, int sm_count, cudaStream_t s) {
    const int g3 = sm_count * 3, g4 = sm_count * 4, g6 = sm_count * 6;
    if (!a.hp.small_m) {
        probe_kernel<1, false><<<g6, PROBE_THREADS, 0, s>>>(a);
        return;
    }
    switch (G) {
        case 1: probe_kernel<1, true><<<g6, PROBE_THREADS, 0, s>>>(a); break;
        case 2: probe_kernel<2, true><<<g6, PROBE_THREADS, 0, s>>>(a); break;
        case 3: probe_kernel<3, true><<<g4, PROBE_THREADS, 0, s>>>(a); break;
        case 4: probe_kernel<4, true><<<g4, PROBE_THREADS, 0, s>>>(a); break;
        case 5: probe_kernel<5, true><<<g4, PROBE_THREADS, 0, s>>>(a); break;
        case 6: probe_kernel<6, true><<<g3, PROBE_THREADS, 0, s>>>(a); break;
        case 7: probe_kernel<7, true><<<g3, PROBE_THREADS, 0, s>>>(a); break;
        default: probe_kernel<8, true><<<g3, PROBE_THREADS, 0, s>>>(a); break;
    }
}

static int ensure_events(pf_db *db, size_t n) {
    while (db->ev_probe.size() < n) {
        cudaEvent_t e;
        PF_CUDA_OK(cudaEventCreate(&e));
        db->ev_probe.push_back(e);
    }
    return PF_OK;
}

// Event times and work counters of one finished query call (after the final stream synchronize).
void pf::account_stats(pf_db *db, const Descent &st, uint64_t n_reads, uint64_t d2h) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, db->ev_begin, db->ev_end);
    db->stats.device_ms += ms;
    for (size_t e = 0; e + 1 < st.n_ev; e += 2) {
        float pm = 0.f;
        cudaEventElapsedTime(&pm, db->ev_probe[e], db->ev_probe[e + 1]);
        if (e / 2 < st.ev_sliced.size() && st.ev_sliced[e / 2]) {
            db->stats.sliced_kernel_ms += pm;
            db->stats.sliced_launches++;
            if (st.ev_sliced[e / 2] == 2) db->stats.entry_kernel_ms += pm;
        }
        else {
            db->stats.probe_kernel_ms += pm;
        }
    }
    db->stats.sector_loads += st.sectors;
    db->stats.line_loads += st.lines;
    db->stats.sliced_pairs += st.sliced_pairs;
    db->stats.blocks++;
    db->stats.reads += n_reads;
    db->stats.pairs += st.pairs;
    db->stats.probes_issued += st.probes;
    db->stats.memo_hits += st.memo_hits;
    db->stats.memo_lookups += st.memo_lookups;
    db->stats.levels += st.levels;
    db->stats.probe_launches += st.probe_launches;
    db->stats.other_launches += st.other_launches;
    db->stats.d2h_bytes += d2h;
}

// Levels [l_begin, l_end) of the level-synchronous descent over the frontier held in `st` (query.rs:99-158):
// per level  [inject entry nodes] -> probe -> scan -> 32-byte D2H of the totals -> scatter.
int pf::run_levels(pf_db *db, const pf_dev_batch *bt, float threshold, int want_hits, uint32_t G, uint64_t kmer_base,
                   size_t l_begin, size_t l_end, uint32_t inj_r0, uint32_t inj_n, Descent &st) {
    cudaStream_t s = db->stream;
    int rc;
    // pairs handed over by the tiles (pf_sliced.cu) replace the plan's own entry nodes
    const bool handed = !db->inj_level_off.empty() && db->inj_level_off.back() > db->inj_level_off.front();
    const size_t last_entry_level = [&] {
        size_t l = l_begin;
        for (size_t i = l_begin; i < l_end; ++i)
            if (handed ? db->inj_level_off[i + 1] > db->inj_level_off[i] : db->entry_start[i + 1] > db->entry_start[i]) l = i;
        return l;
    }();
    for (size_t l = l_begin; l < l_end && (st.n > 0 || l <= last_entry_level); ++l) {
        const int cur = st.cur;
        uint64_t n = st.n;
        // entry nodes of this level: append (read, node) pairs for every read of the injected range
        if (handed && db->inj_level_off[l + 1] > db->inj_level_off[l]) {
            const uint64_t o = db->inj_level_off[l], add = db->inj_level_off[l + 1] - o;
            if (n + add > db->frontier_cap) return PF_SPLIT_CHUNK;
            if ((rc = db->fr_read[cur].grow_keep(n + add, n, s)) || (rc = db->fr_node[cur].grow_keep(n + add, n, s))) return rc;
            PF_CUDA_OK(cudaMemcpyAsync(db->fr_read[cur].p + n, db->inj_read + o, add * 4, cudaMemcpyDeviceToDevice, s));
            PF_CUDA_OK(cudaMemcpyAsync(db->fr_node[cur].p + n, db->inj_node + o, add * 4, cudaMemcpyDeviceToDevice, s));
            n += add;
            st.n = n;
        }
        const uint32_t e0 = db->entry_start[l], n_entry = db->entry_start[l + 1] - e0;
        if (n_entry && inj_n && !handed) {
            const uint64_t add = (uint64_t)inj_n * n_entry;
            if (n + add > db->frontier_cap) return PF_SPLIT_CHUNK;  // query_impl retries with fewer reads
            if ((rc = db->fr_read[cur].grow_keep(n + add, n, s)) || (rc = db->fr_node[cur].grow_keep(n + add, n, s)))
                return rc;
            inject_frontier_kernel<<<(uint32_t)std::min<uint64_t>((add + 255) / 256, 8192), 256, 0, s>>>(
                db->fr_read[cur].p + n, db->fr_node[cur].p + n, inj_r0, inj_n, db->d_entry + e0, n_entry);
            st.other_launches++;
            n += add;
            st.n = n;
        }
        if (n == 0) continue;
        if ((rc = db->pass.ensure(n))) return rc;
        ProbeArgs a{};
        a.fr_read = db->fr_read[cur].p;
        a.fr_node = db->fr_node[cur].p;
        a.n_pairs = (uint32_t)n;
        a.lengths = bt->lengths.p;
        a.kmer_off = bt->kmer_off.p;
        a.hb = db->hb.p;
        a.idx0 = db->idx0.p;
        a.kmer_base = kmer_base;
        a.node_slot = db->d_slot;
        a.node_steps = db->d_steps;
        a.filters = db->d_filters;
        a.words_per_filter = db->wpf;
        a.pass = db->pass.p;
        a.node_pass = db->d_node_pass_copies;
        a.n_nodes = (uint32_t)db->n_nodes;
        a.work_ctr = db->d_work + l;
        a.probes = db->d_probes;
        a.hp = db->hp;
        a.threshold = threshold;
        a.exhaustive = db->exhaustive;
        // 32 pairs per ticket when pairs are cheap and plentiful; 8 when a pair is long (per-pair path, many k-mer groups)
        // or the level has few tickets per warp, where coarse tickets leave warps idle at the end of the launch
        // (cheap pairs = every tested node of the level is a single-step sample that fits one round: the ticket atomic
        // and the metadata round trips are then a large share of a pair's cost)
        const bool cheap = db->level_min_stride[l] > 1 && bt->max_kmers / db->level_min_stride[l] <= 32;
        a.grab = (cheap && n >= (uint64_t)db->sm_count * 32u * PROBE_GRAB * 8u) ? PROBE_GRAB : PROBE_CHUNK;
        // the memo pays when the level's pairs bring each k-mer of its exact nodes several times (sequencing depth):
        // instances = pairs x k-mers per read against the k-mers those filters hold
        if (db->level_memo_entries[l] && (double)n * (double)bt->nominal_kmers >= 4.0 * (double)db->level_memo_kmers[l]) {
            const size_t words = (size_t)db->level_memo_entries[l] << 12;
            if ((rc = db->memo_table.ensure(words))) return rc;
            zero_kernel<<<db->sm_count * 8, 256, 0, s>>>(reinterpret_cast<uint4 *>(db->memo_table.p), words / 2);
            st.other_launches++;
            a.memo = db->memo_table.p;
            a.node_memo = db->d_node_memo;
            // chunk order: up to 8 nodes of the level are worked on at the same time
            const uint32_t n_chunks = (uint32_t)((n + a.grab - 1) / a.grab);
            a.order_streams = std::max(1u, std::min(8u, db->level_memo_regions[l]));
            a.order_span = (n_chunks + a.order_streams - 1) / a.order_streams;
        }
        if ((rc = ensure_events(db, st.n_ev + 2))) return rc;
        const uint64_t level_filter_bytes =
            db->level_slot_lo[l] == NONE32 ? 0 : (uint64_t)(db->level_slot_hi[l] - db->level_slot_lo[l] + 1) * db->wpf * 8;
        const bool window = db->l2_persist_bytes && level_filter_bytes &&
                            (db->l2_persist_policy >= 2 || level_filter_bytes <= std::min(db->l2_persist_bytes, db->l2_window_max));
        if (db->l2_window_set && !window) {  // the previous level's persisting lines would only take L2 away from this one
            cudaStreamAttrValue v{};
            v.accessPolicyWindow.num_bytes = 0;
            PF_CUDA_OK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
            db->l2_window_set = false;
        }
        if (window) {
            db->l2_window_set = true;
            // the level's filters as persisting lines: with more bytes than the set-aside, hitRatio makes that share of
            // the window's lines persisting and the rest ordinary, instead of thrashing the set-aside
            cudaStreamAttrValue v{};
            const uint64_t bytes = std::min<uint64_t>((uint64_t)(db->level_slot_hi[l] - db->level_slot_lo[l] + 1) * db->wpf * 8,
                                                      db->l2_window_max);
            v.accessPolicyWindow.base_ptr = (void *)(db->d_filters + (uint64_t)db->level_slot_lo[l] * db->wpf);
            v.accessPolicyWindow.num_bytes = bytes;
            v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)db->l2_persist_bytes / (double)bytes);
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            PF_CUDA_OK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
        }
        PF_CUDA_OK(cudaEventRecord(db->ev_probe[st.n_ev], s));
        launch_probe(a, G, db->sm_count, s);
        PF_CUDA_OK(cudaEventRecord(db->ev_probe[st.n_ev + 1], s));
        st.n_ev += 2;
        st.ev_sliced.push_back(0);
        st.probe_launches++;
        st.pairs += n;
        st.levels++;
        level_scan_kernel<<<1, 1024, 0, s>>>(db->level_start[l], db->level_start[l + 1], db->d_node_pass_copies,
                                             (uint32_t)db->n_nodes, db->d_node_pass, db->d_left,
                                             db->d_right, db->d_leaf, db->d_next_base, db->d_hit_base,
                                             db->d_blk_counts, db->d_totals, db->d_probes);
        st.other_launches++;
        PF_CUDA_OK(cudaMemcpyAsync(db->h_totals, db->d_totals, sizeof(LevelTotals), cudaMemcpyDeviceToHost, s));
        PF_CUDA_OK(cudaStreamSynchronize(s));
        const uint64_t next_n = db->h_totals->next_pairs;
        st.hits_total = db->h_totals->hits_total;
        st.probes = db->h_totals->probes;
        st.memo_hits = db->h_totals->memo_hits;
        st.memo_lookups = db->h_totals->memo_lookups;
        if (next_n > db->frontier_cap) return PF_SPLIT_CHUNK;
        const int nxt = cur ^ 1;
        if (next_n && ((rc = db->fr_read[nxt].ensure(next_n)) || (rc = db->fr_node[nxt].ensure(next_n)))) return rc;
        // hits of earlier levels and chunks live in the same arrays: grow with copy
        if (want_hits && st.hits_total &&
            ((rc = db->hit_read.grow_keep(st.hits_total, st.hits_before, s)) ||
             (rc = db->hit_leaf.grow_keep(st.hits_total, st.hits_before, s))))
            return rc;
        st.hits_before = st.hits_total;
        scatter_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(
            db->fr_read[cur].p, db->fr_node[cur].p, db->pass.p, (uint32_t)n, db->d_node_pass, db->d_cursor,
            db->d_left, db->d_right, db->d_leaf, db->d_next_base, db->d_hit_base, db->fr_read[nxt].p,
            db->fr_node[nxt].p, db->hit_read.p, db->hit_leaf.p, db->read_hits.p, want_hits);
        st.other_launches++;
        st.n = next_n;
        st.cur = nxt;
    }
    if (db->l2_window_set) {  // later kernels on this stream (CSR, the next block's hashing) use L2 normally
        cudaStreamAttrValue v{};
        v.accessPolicyWindow.num_bytes = 0;
        PF_CUDA_OK(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &v));
        db->l2_window_set = false;
    }
    return PF_OK;
}

// Per-read hit lists (ResultMap, result_map.rs:9-46) as CSR over reads [0, n_reads), leaves ascending within a
// read, built on the device from (hit_read, hit_leaf)[0, hits_total) and the per-read counts in read_hits; the
// offsets of reads [out_r0, out_r0 + out_n] and the leaf array go back through pinned host memory.
int pf::finish_csr(pf_db *db, uint32_t n_reads, uint64_t hits_total, int want_hits, uint32_t out_r0, uint32_t out_n,
                   pf_hits *out, uint64_t *other_launches, uint64_t *d2h) {
    cudaStream_t s = db->stream;
    int rc;
    if (!want_hits) return PF_OK;
    const uint32_t nb = (n_reads + 1023u) / 1024u;
    if ((rc = db->csr_off.ensure((size_t)n_reads + 1)) || (rc = db->csr_bsum.ensure(nb)) ||
        (rc = db->csr_leaf.ensure(std::max<uint64_t>(hits_total, 1))))
        return rc;
    csr_block_sums_kernel<<<nb, 1024, 0, s>>>(db->read_hits.p, n_reads, db->csr_bsum.p);
    csr_scan_sums_kernel<<<1, 1024, 0, s>>>(db->csr_bsum.p, nb);
    csr_offsets_kernel<<<nb, 1024, 0, s>>>(db->read_hits.p, n_reads, db->csr_bsum.p, db->csr_off.p);
    *other_launches += 3;
    if (hits_total) {
        csr_fill_kernel<<<(uint32_t)((hits_total + 255) / 256), 256, 0, s>>>(db->hit_read.p, db->hit_leaf.p, hits_total,
                                                                            db->csr_off.p, db->read_hits.p, db->csr_leaf.p);
        csr_sort_kernel<<<(n_reads + 255) / 256, 256, 0, s>>>(db->csr_off.p, n_reads, db->csr_leaf.p);
        *other_launches += 2;
    }
    if ((rc = db->pin_off.ensure((size_t)out_n + 1)) || (rc = db->pin_leaf.ensure(std::max<uint64_t>(hits_total, 1))))
        return rc;
    PF_CUDA_OK(cudaMemcpyAsync(db->pin_off.p, db->csr_off.p + out_r0, ((size_t)out_n + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (hits_total) PF_CUDA_OK(cudaMemcpyAsync(db->pin_leaf.p, db->csr_leaf.p, hits_total * 4, cudaMemcpyDeviceToHost, s));
    *d2h += ((uint64_t)out_n + 1) * 8 + hits_total * 4;
    (void)out;
    return PF_OK;
}

// Between the tiles and the node-at-a-time descent (hybrid evaluation): flag the reads that own a handed-over pair, give
// them step-0 indices (hash kernel restricted to flagged reads), and make the running hit total the level scan continues
// from equal to what the tiles have already emitted.
static int hand_over(pf_db *db, const pf_dev_batch *bt, HashArgs h, uint64_t chunk_kmers, bool fused, Descent &st) {
    cudaStream_t s = db->stream;
    int rc;
    const uint64_t n_inj = db->inj_level_off.back();
    if (db->hp.small_m || fused) {  // fused: the entry depth hashed on the fly, the survivors' values are cached only now
        if (db->hp.small_m && (rc = db->idx0.ensure(std::max<uint64_t>(chunk_kmers, 1)))) return rc;
        if ((rc = db->read_flag.ensure(bt->n_reads))) return rc;
        PF_CUDA_OK(cudaMemsetAsync(db->read_flag.p, 0, bt->n_reads, s));
        flag_reads_kernel<<<(uint32_t)std::min<uint64_t>((n_inj + 255) / 256, 65535), 256, 0, s>>>(db->inj_read, n_inj, db->read_flag.p);
        PF_CUDA_OK(cudaMemsetAsync(h.work_ctr, 0, 4, s));
        h.idx0 = db->hp.small_m ? db->idx0.p : nullptr;
        h.flags = db->read_flag.p;
        launch_hash(h, db->sm_count * 8, s);
        st.other_launches += 2;
    }
    unsigned long long h64 = st.hits_total;
    PF_CUDA_OK(cudaMemcpyAsync(&db->d_totals->hits_total, &h64, 8, cudaMemcpyHostToDevice, s));
    PF_CUDA_OK(cudaStreamSynchronize(s));  // h64 lives on this stack frame
    return PF_OK;
}

// The level-synchronous descent for one resident batch.  Reads are processed in chunks whose cached k-mer
// hashes (8 B per k-mer) fit the hash-cache budget; hits and counters accumulate over the chunks.
static int query_impl(pf_db *db, const pf_dev_batch *bt, float threshold, int want_hits, pf_hits *out) {
    PF_CUDA_OK(cudaSetDevice(db->device));
    cudaStream_t s = db->stream;
    const uint32_t n_reads = bt->n_reads;
    const size_t n_levels = db->level_start.size() - 1;
    int rc;
    if (db->sharded) {
        set_error("this handle holds a subtree shard: use pf_query_sharded / pf_query_sharded_device");
        return PF_ERR_STATE;
    }
    if (out) *out = pf_hits{};
    if (n_reads == 0) {
        db->out_off.assign(1, 0);
        if (out) out->read_off = db->out_off.data();
        return PF_OK;
    }
    if (bt->kmer_size != db->tree.kmer_size) {
        set_error("batch was uploaded for k=%llu but the database has k=%llu", (unsigned long long)bt->kmer_size,
                  (unsigned long long)db->tree.kmer_size);
        return PF_ERR_STATE;
    }
    if ((rc = update_steps(db, threshold, bt->nominal_kmers))) return rc;
    bool sliced = false;
    if ((rc = sliced_prepare(db, threshold, bt->nominal_kmers, &sliced))) return rc;
    PF_CUDA_OK(cudaEventRecord(db->ev_begin, s));
    if (sliced && (rc = sliced_begin_block(db))) return rc;
    PF_CUDA_OK(cudaMemsetAsync(db->d_blk_counts, 0, std::max<uint64_t>(db->n_leaves, 1) * 8, s));
    PF_CUDA_OK(cudaMemsetAsync(db->d_probes, 0, 24, s));
    PF_CUDA_OK(cudaMemsetAsync(db->d_totals, 0, sizeof(LevelTotals), s));
    if (want_hits) {
        if ((rc = db->read_hits.ensure(n_reads))) return rc;
        PF_CUDA_OK(cudaMemsetAsync(db->read_hits.p, 0, (size_t)n_reads * 4, s));
    }
    Descent st;
    const uint32_t G = group_rounds_for(bt->max_kmers, db->hp.small_m != 0);
    db->stats.group_rounds = G;
    const uint64_t budget_kmers = std::max<uint64_t>(db->hash_cache_bytes / 12, 1);  // 8 B hash + 4 B index
    const std::vector<uint64_t> &ko = bt->h_kmer_off;
    const bool one_chunk = ko.empty();  // prefix sum lives on the device only; the whole batch fits the cache
    uint32_t max_chunk_reads = 0xFFFFFFFFu;
    for (uint32_t r0 = 0; r0 < n_reads;) {
        // chunk [r0, r1): as many reads as the hash cache holds (always at least one)
        uint32_t r1 = n_reads;
        if (!one_chunk) {
            r1 = (uint32_t)(std::upper_bound(ko.begin() + r0 + 1, ko.begin() + n_reads + 1, ko[r0] + budget_kmers) -
                            ko.begin()) - 1;
            if (r1 <= r0) r1 = r0 + 1;
        }
        if (sliced) {  // (read, entry tile) pairs of a chunk are indexed with 32 bits
            // ... and every pair owns 32 B of column bits and a slot in the list of pairs with an output: 64 M pairs
            // per chunk keep that at ~2.3 GB whatever the block size
            const uint64_t cap = std::max<uint64_t>(1, std::min<uint64_t>(db->frontier_cap, 1ULL << 26) /
                                                           std::max<uint64_t>(sliced_entry_tiles(db), 1));
            if (r1 - r0 > cap) r1 = r0 + (uint32_t)cap;
        }
        if (r1 - r0 > max_chunk_reads) r1 = r0 + max_chunk_reads;  // an earlier attempt overflowed the frontier
        const uint32_t n_chunk = r1 - r0;
        // with the prefix sum on the device only, a cut chunk still uses the whole batch's k-mer index space
        const uint64_t chunk_kmers = one_chunk ? bt->total_bases_bound : ko[r1] - ko[r0];
        const uint64_t kmer_base = one_chunk ? 0 : ko[r0];
        // A frontier that outgrows the 32-bit pair index makes the chunk start over with half the reads; the chunk's
        // side effects so far (leaf histogram, hit list, per-read hit counts) are undone first.
        const uint64_t hits_at_chunk = st.hits_total;
        PF_CUDA_OK(cudaMemcpyAsync(db->d_blk_snapshot, db->d_blk_counts, std::max<uint64_t>(db->n_leaves, 1) * 8,
                                   cudaMemcpyDeviceToDevice, s));
        if ((rc = db->hb.ensure(std::max<uint64_t>(chunk_kmers, 1)))) return rc;
        if (db->hp.small_m && !sliced && (rc = db->idx0.ensure(std::max<uint64_t>(chunk_kmers, 1)))) return rc;
        PF_CUDA_OK(cudaMemsetAsync(db->d_node_pass, 0, (2 + NODE_PASS_COPIES) * db->n_nodes * 4, s));
        PF_CUDA_OK(cudaMemsetAsync(db->d_work, 0, db->level_start.size() * 4, s));
        HashArgs h{};
        h.lengths = bt->lengths.p;
        h.word_off = bt->word_off.p;
        h.packed = bt->packed.p;
        h.exc_index = bt->n_exc ? bt->exc_index.p : nullptr;
        h.exc_off = bt->exc_off.p;
        h.exc_bytes = bt->exc_bytes.p;
        h.kmer_off = bt->kmer_off.p;
        h.hb = db->hb.p;
        h.idx0 = (db->hp.small_m && !sliced) ? db->idx0.p : nullptr;
        h.hp = db->hp;
        h.kmer_base = kmer_base;
        h.read0 = r0;
        h.n_reads = n_chunk;
        h.k = db->hp.k;
        h.work_ctr = db->d_work + n_levels;
        h.grab = hash_grab(bt->max_kmers);
        // Sliced path with every entry tile on shared lines: the entry kernel hashes on the fly and only the reads that
        // survive it get their values cached (run_sliced); reads with bytes other than ACGT are hashed here, byte-exact.
        const bool fused = sliced && sliced_fused(db, bt);
        if (fused) {
            if (bt->n_exc && chunk_kmers) {
                if ((rc = db->read_flag.ensure(bt->n_reads))) return rc;
                flag_exc_kernel<<<(uint32_t)std::min<uint32_t>((n_chunk + 255) / 256, 4096), 256, 0, s>>>(bt->exc_index.p, r0, n_chunk,
                                                                                                           db->read_flag.p);
                HashArgs he = h;
                he.flags = db->read_flag.p;
                launch_hash(he, db->sm_count * 8, s);
                st.other_launches += 2;
            }
        } else if (chunk_kmers) {
            launch_hash(h, db->sm_count * 8, s);
            st.other_launches++;
        }
        st.n = 0;
        st.cur = 0;
        db->inj_level_off.clear();
        if (sliced) {
            // the tiles append hits through their own cursor: it continues where the block's hit list stands (earlier
            // chunks may have added hits through the node-at-a-time descent after a hand-over)
            if (r0 != 0 && (rc = sliced_set_hit_cursor(db, st.hits_total))) return rc;
            rc = run_sliced(db, bt, threshold, want_hits, kmer_base, r0, n_chunk, st, fused ? &h : nullptr);
            if (rc == PF_OK && sliced_hybrid(db) && !db->inj_level_off.empty() && db->inj_level_off.back() > 0) {
                // the reads that survived the tiles go down the tree node by node.  That descent uses the step-0 bit indices
                // next to the hash values; they are made now, for the surviving reads only.
                rc = hand_over(db, bt, h, chunk_kmers, fused, st);
                if (rc == PF_OK) rc = run_levels(db, bt, threshold, want_hits, G, kmer_base, 0, n_levels, r0, 0, st);
                db->inj_level_off.clear();
            }
        } else {
            rc = run_levels(db, bt, threshold, want_hits, G, kmer_base, 0, n_levels, r0, n_chunk, st);
        }
        if (rc == PF_SPLIT_CHUNK) {
            if (n_chunk <= 1) {
                set_error("one read alone produces a frontier beyond the 32-bit pair index");
                return PF_ERR_NOMEM;
            }
            max_chunk_reads = n_chunk / 2;
            PF_CUDA_OK(cudaMemcpyAsync(db->d_blk_counts, db->d_blk_snapshot, std::max<uint64_t>(db->n_leaves, 1) * 8,
                                       cudaMemcpyDeviceToDevice, s));
            if (want_hits) PF_CUDA_OK(cudaMemsetAsync(db->read_hits.p + r0, 0, (size_t)n_chunk * 4, s));
            st.hits_total = st.hits_before = hits_at_chunk;
            unsigned long long h64 = hits_at_chunk;
            PF_CUDA_OK(cudaMemcpyAsync(&db->d_totals->hits_total, &h64, 8, cudaMemcpyHostToDevice, s));
            if (sliced && (rc = sliced_set_hit_cursor(db, hits_at_chunk))) return rc;
            PF_CUDA_OK(cudaStreamSynchronize(s));
            db->stats.chunk_splits++;
            continue;
        }
        if (rc) return rc;
        r0 = r1;
    }
    add_counts_kernel<<<(uint32_t)((db->n_leaves + 255) / 256), 256, 0, s>>>(db->d_counts, db->d_blk_counts,
                                                                            (uint32_t)db->n_leaves);
    st.other_launches++;
    uint64_t d2h = st.levels * sizeof(LevelTotals);
    if ((rc = finish_csr(db, n_reads, st.hits_total, want_hits, 0, n_reads, out, &st.other_launches, &d2h))) return rc;
    PF_CUDA_OK(cudaEventRecord(db->ev_end, s));
    PF_CUDA_OK(cudaStreamSynchronize(s));
    PF_CUDA_OK(cudaGetLastError());
    if (out) {
        if (want_hits) {
            out->n_hits = st.hits_total;
            out->read_off = db->pin_off.p;
            out->leaf = db->pin_leaf.p;
        } else {
            db->out_off.assign((size_t)n_reads + 1, 0);
            out->read_off = db->out_off.data();
        }
    }
    account_stats(db, st, n_reads, d2h);
    if (sliced) db->stats.sliced_blocks++;
    {   // feeds the choice between the two evaluation paths (sliced_prepare)
        const double share = std::min(1.0, (double)st.hits_total / (double)n_reads);
        db->related_share = db->related_share < 0 ? share : 0.7 * db->related_share + 0.3 * share;
    }
    return PF_OK;
}


int pf::batch_upload_impl(pf_db *db, const pf_read_batch *in, pf_dev_batch *b, cudaStream_t s) {
    if (!in || (in->n_reads && (!in->lengths || !in->word_off || !in->packed))) {
        set_error("pf_read_batch: null array");
        return PF_ERR_ARG;
    }
    if (in->n_exc && (!in->exc_index || !in->exc_off || !in->exc_bytes)) {
        set_error("pf_read_batch: exception arrays missing");
        return PF_ERR_ARG;
    }
    if (in->n_reads > 0xFF000000u) {
        set_error("pf_read_batch: at most 0xFF000000 reads per batch (32-bit work tickets)");
        return PF_ERR_ARG;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    int rc;
    b->n_reads = in->n_reads;
    b->n_exc = in->n_exc;
    b->n_words = in->n_words;
    b->bytes = 0;
    if (in->n_reads == 0) return PF_OK;
    if ((rc = b->lengths.ensure(in->n_reads)) || (rc = b->word_off.ensure(in->n_reads)) ||
        (rc = b->packed.ensure(in->n_words + 4)))
        return rc;
    PF_CUDA_OK(cudaMemcpyAsync(b->lengths.p, in->lengths, (size_t)in->n_reads * 4, cudaMemcpyHostToDevice, s));
    PF_CUDA_OK(cudaMemcpyAsync(b->word_off.p, in->word_off, (size_t)in->n_reads * 8, cudaMemcpyHostToDevice, s));
    PF_CUDA_OK(cudaMemcpyAsync(b->packed.p, in->packed, (size_t)in->n_words * 4, cudaMemcpyHostToDevice, s));
    PF_CUDA_OK(cudaMemsetAsync(b->packed.p + in->n_words, 0, 16, s));
    b->bytes = (uint64_t)in->n_reads * 12 + in->n_words * 4;
    // k-mer offsets: bookkeeping derived from `lengths`, not part of the host batch.  When the packer supplied
    // max_length / total_bases and the batch fits the hash cache in one chunk, the prefix sum is done on the
    // device; otherwise the host builds it (it is then also needed to cut the batch into chunks).
    const uint32_t k = (uint32_t)db->tree.kmer_size;
    b->kmer_size = db->tree.kmer_size;
    if ((rc = b->kmer_off.ensure((size_t)in->n_reads + 1))) return rc;
    const bool one_chunk = in->max_length != 0 && in->total_bases != 0 && in->total_bases <= db->hash_cache_bytes / 12;
    b->total_bases_bound = in->total_bases;
    {
        uint64_t tb = in->total_bases;
        uint32_t ml = in->max_length;
        if (tb == 0 || ml == 0) {
            tb = 0;
            for (uint32_t r = 0; r < in->n_reads; ++r) {
                tb += in->lengths[r];
                ml = std::max(ml, in->lengths[r]);
            }
        }
        b->total_bases = tb;
        b->max_length = ml;
        const uint64_t mean_len = tb / in->n_reads;
        b->nominal_kmers = std::max<uint64_t>(1, kmers_of((uint32_t)std::min<uint64_t>(mean_len, 0xFFFFFFFFu), k));
    }
    if (one_chunk) {
        b->h_kmer_off.clear();
        b->max_kmers = kmers_of(in->max_length, k);
        const uint32_t nb = (in->n_reads + 1023u) / 1024u;
        if ((rc = b->kcnt.ensure(in->n_reads)) || (rc = b->kbsum.ensure(nb))) return rc;
        kmer_counts_kernel<<<(in->n_reads + 255) / 256, 256, 0, s>>>(b->lengths.p, in->n_reads, k, b->kcnt.p);
        csr_block_sums_kernel<<<nb, 1024, 0, s>>>(b->kcnt.p, in->n_reads, b->kbsum.p);
        csr_scan_sums_kernel<<<1, 1024, 0, s>>>(b->kbsum.p, nb);
        csr_offsets_kernel<<<nb, 1024, 0, s>>>(b->kcnt.p, in->n_reads, b->kbsum.p,
                                               reinterpret_cast<unsigned long long *>(b->kmer_off.p));
        db->stats.other_launches += 4;
    } else {
        b->h_kmer_off.assign((size_t)in->n_reads + 1, 0);
        b->max_kmers = 0;
        for (uint32_t r = 0; r < in->n_reads; ++r) {
            const uint64_t nk = kmers_of(in->lengths[r], k);
            b->h_kmer_off[r + 1] = b->h_kmer_off[r] + nk;
            b->max_kmers = std::max(b->max_kmers, nk);
        }
        PF_CUDA_OK(cudaMemcpyAsync(b->kmer_off.p, b->h_kmer_off.data(), ((size_t)in->n_reads + 1) * 8,
                                   cudaMemcpyHostToDevice, s));
    }
    if (in->n_exc) {
        b->exc_nbytes = in->exc_off[in->n_exc];
        if ((rc = b->exc_index.ensure(in->n_reads)) || (rc = b->exc_off.ensure((size_t)in->n_exc + 1)) ||
            (rc = b->exc_bytes.ensure(std::max<uint64_t>(b->exc_nbytes, 1))))
            return rc;
        PF_CUDA_OK(cudaMemcpyAsync(b->exc_index.p, in->exc_index, (size_t)in->n_reads * 4, cudaMemcpyHostToDevice, s));
        PF_CUDA_OK(cudaMemcpyAsync(b->exc_off.p, in->exc_off, ((size_t)in->n_exc + 1) * 8, cudaMemcpyHostToDevice, s));
        if (b->exc_nbytes)
            PF_CUDA_OK(cudaMemcpyAsync(b->exc_bytes.p, in->exc_bytes, b->exc_nbytes, cudaMemcpyHostToDevice, s));
        b->bytes += (uint64_t)in->n_reads * 4 + ((uint64_t)in->n_exc + 1) * 8 + b->exc_nbytes;
    }
    db->stats.h2d_bytes += b->bytes;
    if (!b->ready) PF_CUDA_OK(cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming));
    PF_CUDA_OK(cudaEventRecord(b->ready, s));
    return PF_OK;
}

// =================================== C ABI ======================================================
extern "C" {

const char *pf_last_error(void) { return g_error.c_str(); }
const char *pf_version(void) { return "pfgpu 0.2 sm_100a"; }

int pf_db_open(const char *db_path, int device, int64_t search_depth, pf_db **out) {
    if (!db_path || !out) {
        set_error("pf_db_open: null argument");
        return PF_ERR_ARG;
    }
    *out = nullptr;
    pf_db *db = new pf_db();
    db->device = device;
    if (const char *m = getenv("PF_MODE")) db->mode = !strcmp(m, "pair") ? 1 : (!strcmp(m, "sliced") ? 2 : 0);
    if (const char *m = getenv("PF_SLICED_HANDOVER")) db->handover = atoi(m) ? 1 : 0;
    int rc = db_open_impl(db, db_path, search_depth);
    if (rc != PF_OK) {
        std::string keep = g_error;
        db_free(db);
        g_error = keep;
        return rc;
    }
    *out = db;
    return PF_OK;
}

int pf_db_info(const pf_db *db, pf_db_info_t *o) {
    if (!db || !o) {
        set_error("pf_db_info: null argument");
        return PF_ERR_ARG;
    }
    memset(o, 0, sizeof *o);
    o->kmer_size = db->tree.kmer_size;
    o->num_bits = db->geom.num_bits;
    o->words_per_filter = db->wpf;
    o->n_nodes = db->n_nodes;
    o->n_leaves = db->n_leaves;
    o->n_filters = db->n_slots;
    o->n_levels = db->level_start.size() - 1;
    o->filter_bytes = db->n_slots * db->wpf * 8;
    o->seed1 = db->geom.seed1;
    o->seed2 = db->geom.seed2;
    o->num_hashes = db->geom.num_hashes;
    o->largest_genome = db->tree.largest_genome;
    o->false_pos_rate = db->tree.false_pos_rate;
    o->device = db->device;
    o->hash_rot = (int32_t)db->hp.rot;
    o->fast_path = fast_path_ok(db->hp) ? 1 : 0;
    o->n_internal = db->n_internal;
    o->n_monotone = db->n_monotone;
    return PF_OK;
}

const char *pf_db_leaf_id(const pf_db *db, uint64_t i) {
    if (!db || i >= db->n_leaves) return nullptr;
    return db->leaf_ids[i].c_str();
}

int pf_db_set_hash_rot(pf_db *db, int rot) {
    if (!db || rot < 0 || rot > 63) {
        set_error("pf_db_set_hash_rot: bad argument");
        return PF_ERR_ARG;
    }
    db->hp.rot = (uint32_t)rot;
    return PF_OK;
}

int pf_db_set_exhaustive(pf_db *db, int on) {
    if (!db) return PF_ERR_ARG;
    db->exhaustive = on ? 1 : 0;
    return PF_OK;
}
int pf_db_set_hash_cache_bytes(pf_db *db, uint64_t bytes) {
    if (!db || bytes < 8) return PF_ERR_ARG;
    db->hash_cache_bytes = bytes;
    return PF_OK;
}
int pf_db_set_memo(pf_db *db, int on, uint64_t budget_bytes) {
    if (!db) return PF_ERR_ARG;
    db->memo = on ? 1 : 0;
    if (budget_bytes) db->memo_budget_bytes = budget_bytes;
    db->steps_mode = -1;  // regions are part of the plan
    return PF_OK;
}
int pf_db_set_lazy(pf_db *db, int on) {
    if (!db) return PF_ERR_ARG;
    db->lazy = on ? 1 : 0;
    return PF_OK;
}
int pf_db_set_frontier_cap(pf_db *db, uint64_t pairs) {
    if (!db || pairs < 1 || pairs > 0xFF000000ULL) {
        set_error("pf_db_set_frontier_cap: 1 .. 0xFF000000 pairs");
        return PF_ERR_ARG;
    }
    db->frontier_cap = pairs;
    return PF_OK;
}
int pf_db_set_handover(pf_db *db, int handover) {
    if (!db || handover < -1 || handover > 1) {
        set_error("pf_db_set_handover: -1, 0 or 1");
        return PF_ERR_ARG;
    }
    db->handover = handover;
    return PF_OK;
}
int pf_db_set_tile_cols(pf_db *db, int cols) {
    if (!db || (cols != 32 && cols != 64 && cols != 128 && cols != 256)) {
        set_error("pf_db_set_tile_cols: 32, 64, 128 or 256");
        return PF_ERR_ARG;
    }
    db->tile_cols = (uint32_t)cols;
    return PF_OK;
}
int pf_db_set_mode(pf_db *db, int mode) {
    if (!db || mode < 0 || mode > 2) {
        set_error("pf_db_set_mode: bad argument");
        return PF_ERR_ARG;
    }
    db->mode = mode;
    return PF_OK;
}
int pf_db_node_steps(pf_db *db, float threshold, uint64_t nominal_kmers, uint32_t *steps) {
    return pf_db_node_plan(db, threshold, nominal_kmers, steps, nullptr);
}
int pf_db_node_plan(pf_db *db, float threshold, uint64_t nominal_kmers, uint32_t *steps, uint32_t *strides) {
    if (!db || (!steps && !strides)) return PF_ERR_ARG;
    PF_CUDA_OK(cudaSetDevice(db->device));
    int rc = update_steps(db, threshold, nominal_kmers ? nominal_kmers : 1);
    if (rc != PF_OK) return rc;
    if (steps) memcpy(steps, db->h_steps.data(), db->n_nodes * 4);
    if (strides) memcpy(strides, db->h_stride.data(), db->n_nodes * 4);
    return PF_OK;
}

int pf_db_detect_hash_rot(pf_db *db, uint64_t dfs_leaf, const uint8_t *genome, uint64_t len, int *rot_out) {
    if (!db || !genome || !rot_out || dfs_leaf >= db->n_leaves) {
        set_error("pf_db_detect_hash_rot: bad argument");
        return PF_ERR_ARG;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    *rot_out = -1;
    uint32_t node = NONE32;
    for (uint32_t u = 0; u < db->n_nodes; ++u)
        if (db->h_leaf[u] == (int32_t)dfs_leaf) node = u;
    if (node == NONE32) return PF_ERR_STATE;
    uint64_t *scratch = nullptr;
    PF_CUDA_OK(cudaMalloc(&scratch, db->wpf * 8));
    int rc = PF_OK;
    for (int rot : {26, 20}) {
        HashParams hp = db->hp;
        hp.rot = (uint32_t)rot;
        bool eq = false;
        if ((rc = build_leaf_filter(genome, len, hp, scratch, db->wpf, db->stream))) break;
        if ((rc = filters_equal(scratch, db->d_filters + (uint64_t)db->h_slot[node] * db->wpf, db->wpf, db->stream, &eq)))
            break;
        if (eq) {
            *rot_out = rot;
            break;
        }
    }
    cudaFree(scratch);
    return rc;
}

void pf_db_close(pf_db *db) { db_free(db); }

int pf_batch_upload(pf_db *db, const pf_read_batch *in, pf_dev_batch **out) {
    if (!db || !in || !out) {
        set_error("pf_batch_upload: null argument");
        return PF_ERR_ARG;
    }
    pf_dev_batch *b = new pf_dev_batch();
    int rc = batch_upload_impl(db, in, b, db->stream);
    if (rc == PF_OK && cudaStreamSynchronize(db->stream) != cudaSuccess) {
        set_error("CUDA error in pf_batch_upload: %s", cudaGetErrorString(cudaGetLastError()));
        rc = PF_ERR_CUDA;
    }
    if (rc != PF_OK) {
        b->release();
        delete b;
        return rc;
    }
    *out = b;
    return PF_OK;
}

void pf_batch_free(pf_db *db, pf_dev_batch *b) {
    if (!b) return;
    if (db) cudaSetDevice(db->device);
    b->release();
    delete b;
}

int pf_query_device(pf_db *db, pf_dev_batch *batch, float threshold, int want_hits, pf_hits *out) {
    if (!db || !batch) {
        set_error("pf_query_device: null argument");
        return PF_ERR_ARG;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    if (batch->ready) PF_CUDA_OK(cudaStreamWaitEvent(db->stream, batch->ready, 0));  // asynchronous upload
    return query_impl(db, batch, threshold, want_hits, out);
}

int pf_batch_upload_async(pf_db *db, const pf_read_batch *in, pf_dev_batch **inout) {
    if (!db || !in || !inout) {
        set_error("pf_batch_upload_async: null argument");
        return PF_ERR_ARG;
    }
    pf_dev_batch *b = *inout ? *inout : new pf_dev_batch();
    int rc = batch_upload_impl(db, in, b, db->copy_stream);
    if (rc != PF_OK) {
        if (!*inout) {
            b->release();
            delete b;
        }
        return rc;
    }
    *inout = b;
    return PF_OK;
}

int pf_query_block(pf_db *db, const pf_read_batch *in, float threshold, int want_hits, pf_hits *out) {
    if (!db || !in) {
        set_error("pf_query_block: null argument");
        return PF_ERR_ARG;
    }
    static const bool timing = getenv("PF_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = batch_upload_impl(db, in, &db->own_batch, db->stream);
    if (rc != PF_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    rc = query_impl(db, &db->own_batch, threshold, want_hits, out);
    if (timing && rc == PF_OK) {
        float dev_ms = 0.f;
        cudaEventElapsedTime(&dev_ms, db->ev_begin, db->ev_end);
        fprintf(stderr, "[pf_query_block] reads=%u upload_call=%.2f ms query_call=%.2f ms (device %.2f ms)\n", in->n_reads,
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), dev_ms);
    }
    return rc;
}

int pf_leaf_counts(pf_db *db, uint64_t *counts) {
    if (!db || !counts) {
        set_error("pf_leaf_counts: null argument");
        return PF_ERR_ARG;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    PF_CUDA_OK(cudaMemcpyAsync(counts, db->d_counts, db->n_leaves * 8, cudaMemcpyDeviceToHost, db->stream));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    return PF_OK;
}

int pf_reset_counts(pf_db *db) {
    if (!db) return PF_ERR_ARG;
    PF_CUDA_OK(cudaSetDevice(db->device));
    PF_CUDA_OK(cudaMemsetAsync(db->d_counts, 0, std::max<uint64_t>(db->n_leaves, 1) * 8, db->stream));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    return PF_OK;
}

int pf_save_leaf_counts(pf_db *db, const char *csv_path) {
    if (!db || !csv_path) {
        set_error("pf_save_leaf_counts: null argument");
        return PF_ERR_ARG;
    }
    std::vector<uint64_t> c(std::max<uint64_t>(db->n_leaves, 1));
    int rc = pf_leaf_counts(db, c.data());
    if (rc != PF_OK) return rc;
    FILE *fp = fopen(csv_path, "wb");
    if (!fp) {
        set_error("cannot create %s", csv_path);
        return PF_ERR_IO;
    }
    for (uint64_t i = 0; i < db->n_leaves; ++i)
        if (c[i] > 0) fprintf(fp, "%s,%llu\n", db->leaf_ids[i].c_str(), (unsigned long long)c[i]);
    if (fclose(fp) != 0) {
        set_error("problem writing to output file %s", csv_path);
        return PF_ERR_IO;
    }
    return PF_OK;
}

int pf_get_stats(pf_db *db, pf_stats_t *o) {
    if (!db || !o) return PF_ERR_ARG;
    *o = db->stats;
    return PF_OK;
}
void *pf_db_stream(pf_db *db) { return db ? (void *)db->stream : nullptr; }
int pf_reset_stats(pf_db *db) {
    if (!db) return PF_ERR_ARG;
    db->stats = pf_stats_t{};
    return PF_OK;
}

int pf_nccl_unique_id(void *id128) {
    if (!id128) return PF_ERR_ARG;
    int rc = load_nccl();
    if (rc != PF_OK) return rc;
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) {
        set_error("ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        return PF_ERR_NCCL;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return PF_OK;
}

int pf_comm_init(pf_db *db, int nranks, int rank, const void *id128) {
    if (!db || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
        set_error("pf_comm_init: bad argument");
        return PF_ERR_ARG;
    }
    if (db->comm) {
        set_error("pf_comm_init: the handle already has a communicator");
        return PF_ERR_STATE;
    }
    return comm_init_impl(db, nranks, rank, id128);
}

}  // extern "C"

int pf::comm_init_impl(pf_db *db, int nranks, int rank, const void *id128) {
    int rc = load_nccl();
    if (rc != PF_OK) return rc;
    PF_CUDA_OK(cudaSetDevice(db->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&db->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        db->comm = nullptr;
        return PF_ERR_NCCL;
    }
    return PF_OK;
}

extern "C" {

int pf_allreduce_counts(pf_db *db) {
    if (!db) return PF_ERR_ARG;
    if (!db->comm) {
        set_error("pf_allreduce_counts: call pf_comm_init first");
        return PF_ERR_STATE;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    ncclResult_t r = g_nccl.AllReduce(db->d_counts, db->d_counts, db->n_leaves, ncclUint64, ncclSum, db->comm, db->stream);
    if (r != ncclSuccess) {
        set_error("ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        return PF_ERR_NCCL;
    }
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    return PF_OK;
}

}  // extern "C"
