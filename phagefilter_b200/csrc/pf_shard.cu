// pf_shard.cu -- subtree-sharded gSBT for trees larger than one GPU's HBM (north_star).
//
// The flattened tree is cut at one level L: the nodes of levels < L (the "top") are replicated on every rank,
// every node of level L roots a subtree owned by exactly one rank, and a rank keeps only the filters of the top
// and of its own subtrees resident.  A query is collective over the ranks of the communicator:
//
//   1. all ranks' 2-bit read batches are gathered over NVLink (40 B per 150 bp read), so a read has the same
//      global index on every rank;
//   2. phase A: every rank descends the top with ITS OWN reads (query.rs:99-158, levels < L);
//   3. the surviving (read, node) pairs at level L are node-major, and an owner's nodes are contiguous inside
//      the level, so each destination's pairs are one slice of the frontier: ONE all-to-all (ncclSend/ncclRecv
//      group) moves them to the owners of the subtrees;
//   4. phase B: every rank descends its subtrees with the pairs it received, hashing only the reads they name;
//      where the step plan skips the whole top above a subtree (saturated filters), the owner pairs that
//      subtree's entry nodes with the reads of every rank itself and nothing is exchanged for them;
//   5. (read, leaf) hits go back to the rank that owns the read with a second all-to-all; per-leaf counters
//      stay partial per rank and are combined by the usual single all-reduce (pf_allreduce_counts).
//
// Results are identical to the replicated tree's and to the reference's: the same nodes see the same reads.
#include <cstring>

#include "pf_db.h"

namespace pf {

struct ShardState {
    pf_dev_batch gathered;                 // every rank's reads, rank-major
    unsigned long long *d_mine = nullptr;  // [max(8, nranks + 1)] this rank's header / counts / slice bounds
    unsigned long long *d_all = nullptr;   // [nranks * max(8, nranks)] gathered headers / count matrix
    unsigned long long *h_all = nullptr;   // pinned copy
    uint32_t *d_cut_lo = nullptr;          // [nranks + 1]
    uint32_t *d_rbase = nullptr;           // [nranks + 1] first global read index of every rank
    unsigned int *d_cursor = nullptr;      // [nranks] partition cursors
    DevBuf<uint8_t> need_hash;             // [n_total] reads named by the pairs received for phase B
    DevBuf<uint32_t> send_read, send_leaf; // hits on their way back to the reads' owners
    pf_shard_stats_t stats{};
};

constexpr int HDR_WORDS = 8;  // n_reads, n_words, n_exc, exc_nbytes, max_length, total_bases, 2 spare

// ---- kernels -------------------------------------------------------------------------------------------
__global__ void rebase_reads_kernel(uint64_t *word_off, uint32_t *exc_index, uint32_t n, uint64_t word_base,
                                    uint32_t exc_base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    word_off[i] += word_base;
    if (exc_index && exc_index[i] != NONE32_D) exc_index[i] += exc_base;
}
__global__ void add_u64_kernel(uint64_t *a, uint32_t n, uint64_t base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += base;
}
__global__ void set_u64_kernel(uint64_t *p, uint64_t v) { *p = v; }

// the sharded descent does not cut chunks: a frontier beyond the pair index is the caller's to split
static int shard_rc(int rc) {
    if (rc == PF_SPLIT_CHUNK) {
        pf::set_error("frontier exceeds the 32-bit pair index: use smaller read blocks with a sharded tree");
        return PF_ERR_NOMEM;
    }
    return rc;
}


// The frontier at the cut level is sorted by node; rank d owns the nodes [cut_lo[d], cut_lo[d+1]).
// out[d] = number of pairs going to rank d; out[nranks + 1 + d] = first pair of that slice.
__global__ void frontier_bounds_kernel(const uint32_t *__restrict__ fr_node, uint32_t n, const uint32_t *__restrict__ cut_lo,
                                       uint32_t nranks, unsigned long long *out) {
    __shared__ uint32_t lb[64];
    const uint32_t d = threadIdx.x;
    if (d <= nranks) {
        const uint32_t key = cut_lo[d];
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (fr_node[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        lb[d] = lo;
    }
    __syncthreads();
    if (d < nranks) {
        out[d] = lb[d + 1] - lb[d];
        out[nranks + 1 + d] = lb[d];
    }
}
__global__ void mark_reads_kernel(const uint32_t *__restrict__ fr_read, uint32_t n, uint8_t *flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[fr_read[i]] = 1;
}
PF_D uint32_t owner_of_read(uint32_t r, const uint32_t *__restrict__ rbase, uint32_t nranks) {
    uint32_t d = 0;
    while (d + 1 < nranks && r >= rbase[d + 1]) ++d;
    return d;
}
__global__ void hit_owner_count_kernel(const uint32_t *__restrict__ hit_read, uint32_t n, const uint32_t *__restrict__ rbase,
                                       uint32_t nranks, unsigned long long *cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n;
    const uint32_t act = __ballot_sync(0xFFFFFFFFu, on);
    if (!on) return;
    const uint32_t d = owner_of_read(hit_read[i], rbase, nranks);
    const uint32_t peers = __match_any_sync(act, d);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(cnt + d, (unsigned long long)__popc(peers));
}
// send slices are rank-major: slice d starts at off[d]
__global__ void hit_partition_kernel(const uint32_t *__restrict__ hit_read, const uint32_t *__restrict__ hit_leaf, uint32_t n,
                                     const uint32_t *__restrict__ rbase, uint32_t nranks,
                                     const unsigned long long *__restrict__ off, unsigned int *cursor, uint32_t *send_read,
                                     uint32_t *send_leaf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n;
    const uint32_t act = __ballot_sync(0xFFFFFFFFu, on);
    if (!on) return;
    const uint32_t r = hit_read[i], lane = threadIdx.x & 31u;
    const uint32_t d = owner_of_read(r, rbase, nranks);
    const uint32_t peers = __match_any_sync(act, d);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor + d, (unsigned)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const unsigned long long p = off[d] + base + __popc(peers & ((1u << lane) - 1u));
    send_read[p] = r;
    send_leaf[p] = hit_leaf[i];
}
__global__ void hit_offsets_kernel(const unsigned long long *cnt, uint32_t nranks, unsigned long long *off) {
    if (threadIdx.x == 0) {
        unsigned long long a = 0;
        for (uint32_t d = 0; d < nranks; ++d) {
            off[d] = a;
            a += cnt[d];
        }
    }
}
__global__ void count_hits_kernel(const uint32_t *__restrict__ hit_read, uint32_t n, uint32_t *read_hits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(read_hits + hit_read[i], 1u);
}

// ---- planning (host only; no CUDA) ------------------------------------------------------------------------
// Cut level: the level whose cut minimises  (#top nodes) + (largest owned node count), i.e. the filters one rank
// has to hold.  Owners take contiguous ranges of the cut level's nodes (balanced by subtree size), which is what
// makes every destination's pairs one slice of the node-major frontier.
void plan_shards(const std::vector<uint32_t> &level_start, const std::vector<uint32_t> &left,
                 const std::vector<uint32_t> &right, int nranks, int64_t cut_req, ShardPlan &out) {
    const size_t nn = left.size(), n_levels = level_start.size() - 1;
    std::vector<uint64_t> size(nn, 1);
    for (size_t u = nn; u-- > 0;) {
        if (left[u] != NONE32) size[u] += size[left[u]];
        if (right[u] != NONE32) size[u] += size[right[u]];
    }
    auto partition = [&](size_t l, std::vector<uint32_t> &lo) -> uint64_t {  // returns the largest owned size
        const uint32_t s = level_start[l], e = level_start[l + 1];
        uint64_t total = 0;
        for (uint32_t u = s; u < e; ++u) total += size[u];
        lo.assign(nranks + 1, e);
        lo[0] = s;
        uint64_t acc = 0, worst = 0, cur = 0;
        int d = 0;
        for (uint32_t u = s; u < e; ++u) {
            // close rank d's range before u when the boundary (d+1)/nranks of the total is nearer to the nodes
            // taken so far than to those including u
            while (d + 1 < nranks && cur > 0 && ((2 * acc + size[u]) * (uint64_t)nranks > 2 * total * (uint64_t)(d + 1))) {
                worst = std::max(worst, cur);
                cur = 0;
                lo[++d] = u;
            }
            acc += size[u];
            cur += size[u];
        }
        worst = std::max(worst, cur);
        for (int q = d + 1; q <= nranks; ++q) lo[q] = e;
        return worst;
    };
    size_t best_l = std::min<size_t>(1, n_levels);
    if (cut_req >= 0) {
        best_l = (size_t)std::min<int64_t>(std::max<int64_t>(cut_req, 1), (int64_t)n_levels);
    } else {
        uint64_t best_cost = ~0ULL;
        std::vector<uint32_t> lo;
        for (size_t l = 1; l < n_levels; ++l) {
            const uint64_t cost = level_start[l] + partition(l, lo);
            if (cost < best_cost) {
                best_cost = cost;
                best_l = l;
            }
        }
    }
    out.cut_level = (uint32_t)best_l;
    out.owner.assign(nn, -1);
    out.cut_lo.assign(nranks + 1, (uint32_t)nn);
    out.max_owned = 0;
    out.top_nodes = best_l < level_start.size() ? level_start[best_l] : nn;
    if (best_l >= n_levels) return;  // everything replicated (single-level tree)
    out.max_owned = partition(best_l, out.cut_lo);
    for (int d = 0; d < nranks; ++d)
        for (uint32_t u = out.cut_lo[d]; u < out.cut_lo[d + 1]; ++u) out.owner[u] = d;
    for (size_t u = level_start[best_l]; u < nn; ++u) {  // parents precede children in level order
        if (left[u] != NONE32) out.owner[left[u]] = out.owner[u];
        if (right[u] != NONE32) out.owner[right[u]] = out.owner[u];
    }
}

int shard_plan(pf_db *db, int64_t cut_level_req) {
    ShardPlan p;
    plan_shards(db->level_start, db->h_left, db->h_right, db->nranks, cut_level_req, p);
    db->cut_level = p.cut_level;
    db->h_owner = p.owner;
    db->cut_lo = p.cut_lo;
    // resident filters: top + owned; slots are renumbered over the resident nodes only
    std::vector<uint32_t> remap(db->n_slots, NONE32);
    uint32_t n_res = 0;
    for (size_t u = 0; u < db->n_nodes; ++u)
        if (db->h_owner[u] < 0 || db->h_owner[u] == db->rank)
            if (remap[db->h_slot[u]] == NONE32) remap[db->h_slot[u]] = n_res++;
    for (size_t u = 0; u < db->n_nodes; ++u) db->h_slot[u] = remap[db->h_slot[u]];  // shared paths stay shared
    db->n_slots = n_res;
    return PF_OK;
}

void shard_free(pf_db *db) {
    ShardState *S = db->shard;
    if (!S) return;
    S->gathered.release();
    cudaFree(S->d_mine);
    cudaFree(S->d_all);
    if (S->h_all) cudaFreeHost(S->h_all);
    cudaFree(S->d_cut_lo);
    cudaFree(S->d_rbase);
    cudaFree(S->d_cursor);
    S->need_hash.release();
    S->send_read.release();
    S->send_leaf.release();
    delete S;
    db->shard = nullptr;
}

static int shard_state_init(pf_db *db) {
    ShardState *S = new ShardState();
    db->shard = S;
    const size_t nr = (size_t)db->nranks, row = std::max<size_t>(HDR_WORDS, 2 * nr + 2);
    PF_CUDA_OK(cudaMalloc(&S->d_mine, row * 8));
    PF_CUDA_OK(cudaMalloc(&S->d_all, nr * row * 8));
    PF_CUDA_OK(cudaMallocHost(&S->h_all, nr * row * 8));
    PF_CUDA_OK(cudaMalloc(&S->d_cut_lo, (nr + 1) * 4));
    PF_CUDA_OK(cudaMalloc(&S->d_rbase, (nr + 1) * 4));
    PF_CUDA_OK(cudaMalloc(&S->d_cursor, nr * 4));
    PF_CUDA_OK(cudaMemcpyAsync(S->d_cut_lo, db->cut_lo.data(), (nr + 1) * 4, cudaMemcpyHostToDevice, db->stream));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    return PF_OK;
}

// all-gather of `words` u64 per rank: d_mine[0, words) -> h_all[rank * words + i]
static int gather_words(pf_db *db, size_t words) {
    ShardState *S = db->shard;
    PF_NCCL_OK(g_nccl.AllGather(S->d_mine, S->d_all, words, ncclUint64, db->comm, db->stream));
    PF_CUDA_OK(cudaMemcpyAsync(S->h_all, S->d_all, (size_t)db->nranks * words * 8, cudaMemcpyDeviceToHost, db->stream));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    S->stats.collectives++;
    return PF_OK;
}

// all-to-all of u32 slices: send[a] + s_off[p], cnt[me][p] elements to rank p; received into recv[a] + r_off[p]
static int exchange_slices(pf_db *db, int n_arrays, const uint32_t *const *send, uint32_t *const *recv,
                           const std::vector<uint64_t> &s_off, const std::vector<uint64_t> &r_off,
                           const std::vector<uint64_t> &cnt /* [nranks * nranks], row = sender */) {
    const int me = db->rank, nr = db->nranks;
    cudaStream_t s = db->stream;
    ShardState *S = db->shard;
    const uint64_t self = cnt[(size_t)me * nr + me];
    for (int a = 0; a < n_arrays && self; ++a)
        PF_CUDA_OK(cudaMemcpyAsync(recv[a] + r_off[me], send[a] + s_off[me], self * 4, cudaMemcpyDeviceToDevice, s));
    bool any = false;
    for (int p = 0; p < nr; ++p)
        if (p != me && (cnt[(size_t)me * nr + p] || cnt[(size_t)p * nr + me])) any = true;
    if (!any) return PF_OK;
    PF_NCCL_OK(g_nccl.GroupStart());
    for (int p = 0; p < nr; ++p) {
        if (p == me) continue;
        const uint64_t out_n = cnt[(size_t)me * nr + p], in_n = cnt[(size_t)p * nr + me];
        for (int a = 0; a < n_arrays; ++a) {
            if (out_n) PF_NCCL_OK(g_nccl.Send(send[a] + s_off[p], out_n, ncclUint32, p, db->comm, s));
            if (in_n) PF_NCCL_OK(g_nccl.Recv(recv[a] + r_off[p], in_n, ncclUint32, p, db->comm, s));
        }
        S->stats.bytes_sent += out_n * 4 * n_arrays;
    }
    PF_NCCL_OK(g_nccl.GroupEnd());
    S->stats.collectives++;
    return PF_OK;
}

// Step 1: gather every rank's batch into S->gathered (global read index = rank base + local index).
static int gather_reads(pf_db *db, const pf_dev_batch *local, std::vector<uint64_t> &rbase) {
    ShardState *S = db->shard;
    pf_dev_batch &g = S->gathered;
    cudaStream_t s = db->stream;
    const int nr = db->nranks, me = db->rank;
    int rc;
    unsigned long long hdr[HDR_WORDS] = {local->n_reads, local->n_reads ? local->n_words : 0,
                                         local->n_reads ? local->n_exc : 0, local->n_reads && local->n_exc ? local->exc_nbytes : 0,
                                         local->n_reads ? local->max_length : 0, local->n_reads ? local->total_bases : 0, 0, 0};
    PF_CUDA_OK(cudaMemcpyAsync(S->d_mine, hdr, sizeof hdr, cudaMemcpyHostToDevice, s));
    if ((rc = gather_words(db, HDR_WORDS))) return rc;
    std::vector<uint64_t> wbase(nr + 1, 0), ebase(nr + 1, 0), bbase(nr + 1, 0);
    rbase.assign(nr + 1, 0);
    uint64_t max_len = 0, total_bases = 0;
    for (int r = 0; r < nr; ++r) {
        const unsigned long long *h = S->h_all + (size_t)r * HDR_WORDS;
        rbase[r + 1] = rbase[r] + h[0];
        wbase[r + 1] = wbase[r] + ((h[1] + 1) & ~1ULL);  // every rank's words start 8-byte aligned
        ebase[r + 1] = ebase[r] + h[2];
        bbase[r + 1] = bbase[r] + h[3];
        max_len = std::max<uint64_t>(max_len, h[4]);
        total_bases += h[5];
    }
    const uint64_t nT = rbase[nr], wT = wbase[nr], eT = ebase[nr], bT = bbase[nr];
    if (nT > 0xFFFFFFF0ULL || eT > 0xFFFFFFF0ULL) {
        set_error("%llu reads over all ranks exceed the 32-bit read index: use smaller read blocks", (unsigned long long)nT);
        return PF_ERR_NOMEM;
    }
    g.n_reads = (uint32_t)nT;
    g.n_exc = (uint32_t)eT;
    g.n_words = wT;
    g.exc_nbytes = bT;
    g.kmer_size = db->tree.kmer_size;
    g.max_length = (uint32_t)max_len;
    g.total_bases = g.total_bases_bound = total_bases;
    const uint32_t k = (uint32_t)db->tree.kmer_size;
    g.max_kmers = kmers_of((uint32_t)max_len, k);
    g.nominal_kmers = nT ? std::max<uint64_t>(1, kmers_of((uint32_t)std::min<uint64_t>(total_bases / nT, 0xFFFFFFFFu), k)) : 1;
    g.h_kmer_off.clear();
    if (nT == 0) return PF_OK;
    if ((rc = g.lengths.ensure(nT)) || (rc = g.word_off.ensure(nT)) || (rc = g.packed.ensure(wT + 4)) ||
        (rc = g.kmer_off.ensure(nT + 1)))
        return rc;
    if (eT && ((rc = g.exc_index.ensure(nT)) || (rc = g.exc_off.ensure(eT + 1)) || (rc = g.exc_bytes.ensure(std::max<uint64_t>(bT, 1)))))
        return rc;
    PF_NCCL_OK(g_nccl.GroupStart());
    for (int r = 0; r < nr; ++r) {
        const unsigned long long *h = S->h_all + (size_t)r * HDR_WORDS;
        if (!h[0]) continue;
        const bool root = r == me;
        uint32_t *len_dst = g.lengths.p + rbase[r];
        uint64_t *wo_dst = g.word_off.p + rbase[r];
        uint32_t *pk_dst = g.packed.p + wbase[r];
        PF_NCCL_OK(g_nccl.Broadcast(root ? (const void *)local->lengths.p : len_dst, len_dst, h[0], ncclUint32, r, db->comm, s));
        PF_NCCL_OK(g_nccl.Broadcast(root ? (const void *)local->word_off.p : wo_dst, wo_dst, h[0], ncclUint64, r, db->comm, s));
        PF_NCCL_OK(g_nccl.Broadcast(root ? (const void *)local->packed.p : pk_dst, pk_dst, h[1], ncclUint32, r, db->comm, s));
        if (h[2]) {
            uint32_t *xi_dst = g.exc_index.p + rbase[r];
            uint64_t *xo_dst = g.exc_off.p + ebase[r];
            uint8_t *xb_dst = g.exc_bytes.p + bbase[r];
            PF_NCCL_OK(g_nccl.Broadcast(root ? (const void *)local->exc_index.p : xi_dst, xi_dst, h[0], ncclUint32, r, db->comm, s));
            PF_NCCL_OK(g_nccl.Broadcast(root ? (const void *)local->exc_off.p : xo_dst, xo_dst, h[2], ncclUint64, r, db->comm, s));
            if (h[3]) PF_NCCL_OK(g_nccl.Broadcast(root ? (const void *)local->exc_bytes.p : xb_dst, xb_dst, h[3], ncclUint8, r, db->comm, s));
        }
        if (!root) S->stats.bytes_received += h[0] * 12 + h[1] * 4 + (h[2] ? h[0] * 4 + h[2] * 8 + h[3] : 0);
    }
    PF_NCCL_OK(g_nccl.GroupEnd());
    S->stats.collectives++;
    PF_CUDA_OK(cudaMemsetAsync(g.packed.p + wT, 0, 16, s));
    for (int r = 0; r < nr; ++r) {
        const unsigned long long *h = S->h_all + (size_t)r * HDR_WORDS;
        if (!h[0]) continue;
        const uint32_t n = (uint32_t)h[0];
        if (eT && !h[2]) PF_CUDA_OK(cudaMemsetAsync(g.exc_index.p + rbase[r], 0xFF, (size_t)n * 4, s));
        if (wbase[r] || (eT && h[2] && ebase[r]))
            rebase_reads_kernel<<<(n + 255) / 256, 256, 0, s>>>(g.word_off.p + rbase[r], eT && h[2] ? g.exc_index.p + rbase[r] : nullptr,
                                                               n, wbase[r], (uint32_t)ebase[r]);
        if (h[2] && bbase[r]) add_u64_kernel<<<((uint32_t)h[2] + 255) / 256, 256, 0, s>>>(g.exc_off.p + ebase[r], (uint32_t)h[2], bbase[r]);
    }
    if (eT) set_u64_kernel<<<1, 1, 0, s>>>(g.exc_off.p + eT, bT);
    // k-mer offsets of the gathered reads (file_parser.rs:136-139), prefix sum on the device
    const uint32_t n = (uint32_t)nT, nb = (n + 1023u) / 1024u;
    if ((rc = g.kcnt.ensure(n)) || (rc = g.kbsum.ensure(nb))) return rc;
    kmer_counts_kernel<<<(n + 255) / 256, 256, 0, s>>>(g.lengths.p, n, k, g.kcnt.p);
    csr_block_sums_kernel<<<nb, 1024, 0, s>>>(g.kcnt.p, n, g.kbsum.p);
    csr_scan_sums_kernel<<<1, 1024, 0, s>>>(g.kbsum.p, nb);
    csr_offsets_kernel<<<nb, 1024, 0, s>>>(g.kcnt.p, n, g.kbsum.p, reinterpret_cast<unsigned long long *>(g.kmer_off.p));
    PF_CUDA_OK(cudaGetLastError());
    return PF_OK;
}

static void fill_hash_args(pf_db *db, const pf_dev_batch &g, HashArgs &h, uint32_t read0, uint32_t n, const uint8_t *flags) {
    h = HashArgs{};
    h.lengths = g.lengths.p;
    h.word_off = g.word_off.p;
    h.packed = g.packed.p;
    h.exc_index = g.n_exc ? g.exc_index.p : nullptr;
    h.exc_off = g.exc_off.p;
    h.exc_bytes = g.exc_bytes.p;
    h.kmer_off = g.kmer_off.p;
    h.hb = db->hb.p;
    h.idx0 = db->hp.small_m ? db->idx0.p : nullptr;
    h.hp = db->hp;
    h.kmer_base = 0;
    h.read0 = read0;
    h.n_reads = n;
    h.k = db->hp.k;
    h.work_ctr = db->d_work + (db->level_start.size() - 1);
    h.flags = flags;
    h.grab = hash_grab(g.max_kmers);
}
static int hash_range(pf_db *db, const pf_dev_batch &g, uint32_t read0, uint32_t n, const uint8_t *flags, Descent &st) {
    if (!n) return PF_OK;
    HashArgs h;
    fill_hash_args(db, g, h, read0, n, flags);
    PF_CUDA_OK(cudaMemsetAsync(h.work_ctr, 0, 4, db->stream));
    launch_hash(h, db->sm_count * 8, db->stream);
    st.other_launches++;
    return PF_OK;
}

static int query_sharded_impl(pf_db *db, const pf_dev_batch *local, float threshold, int want_hits, pf_hits *out) {
    PF_CUDA_OK(cudaSetDevice(db->device));
    ShardState *S = db->shard;
    cudaStream_t s = db->stream;
    const int nr = db->nranks, me = db->rank;
    const size_t n_levels = db->level_start.size() - 1, Lc = std::min<size_t>(db->cut_level, n_levels);
    int rc;
    if (out) *out = pf_hits{};
    if (local->n_reads && local->kmer_size != db->tree.kmer_size) {
        set_error("batch was uploaded for k=%llu but the database has k=%llu", (unsigned long long)local->kmer_size,
                  (unsigned long long)db->tree.kmer_size);
        return PF_ERR_STATE;
    }
    PF_CUDA_OK(cudaEventRecord(db->ev_begin, s));
    std::vector<uint64_t> rbase;
    if ((rc = gather_reads(db, local, rbase))) return rc;
    const pf_dev_batch &g = S->gathered;
    const uint32_t nT = g.n_reads, r_me = (uint32_t)rbase[me], n_me = local->n_reads;
    db->out_off.assign((size_t)n_me + 1, 0);
    if (nT == 0) {
        if (out) out->read_off = db->out_off.data();
        return PF_OK;
    }
    if (g.total_bases > db->hash_cache_bytes / 12) {
        set_error("the k-mer hash cache of %llu reads over all ranks (%.1f GB) exceeds the budget: use smaller read blocks "
                  "or raise pf_db_set_hash_cache_bytes",
                  (unsigned long long)nT, g.total_bases * 12 / 1e9);
        return PF_ERR_NOMEM;
    }
    if ((rc = update_steps(db, threshold, g.nominal_kmers))) return rc;  // same plan on every rank
    {
        std::vector<uint32_t> rb32(rbase.begin(), rbase.end());
        PF_CUDA_OK(cudaMemcpyAsync(S->d_rbase, rb32.data(), (size_t)(nr + 1) * 4, cudaMemcpyHostToDevice, s));
        PF_CUDA_OK(cudaStreamSynchronize(s));  // rb32 is a pageable temporary
    }
    PF_CUDA_OK(cudaMemsetAsync(db->d_blk_counts, 0, std::max<uint64_t>(db->n_leaves, 1) * 8, s));
    PF_CUDA_OK(cudaMemsetAsync(db->d_probes, 0, 24, s));
    PF_CUDA_OK(cudaMemsetAsync(db->d_totals, 0, sizeof(LevelTotals), s));
    PF_CUDA_OK(cudaMemsetAsync(db->d_node_pass, 0, (2 + NODE_PASS_COPIES) * db->n_nodes * 4, s));
    PF_CUDA_OK(cudaMemsetAsync(db->d_work, 0, db->level_start.size() * 4, s));
    if (want_hits) {
        if ((rc = db->read_hits.ensure(nT))) return rc;
        PF_CUDA_OK(cudaMemsetAsync(db->read_hits.p, 0, (size_t)nT * 4, s));
    }
    const uint64_t kmers_bound = std::max<uint64_t>(g.total_bases, 1);
    if ((rc = db->hb.ensure(kmers_bound))) return rc;
    if (db->hp.small_m && (rc = db->idx0.ensure(kmers_bound))) return rc;
    Descent st;
    const uint32_t G = group_rounds_for(g.max_kmers, db->hp.small_m != 0);
    db->stats.group_rounds = G;

    // ---- phase A: the replicated top, this rank's reads ---------------------------------------------------
    if ((rc = hash_range(db, g, r_me, n_me, nullptr, st))) return rc;
    if ((rc = run_levels(db, &g, threshold, want_hits ? 1 : 0, G, 0, 0, Lc, r_me, n_me, st))) return shard_rc(rc);
    const uint64_t hits_a = st.hits_total, pairs_a = st.pairs;
    if (Lc >= n_levels) st.n = 0;  // nothing below the cut

    // ---- frontier exchange: slice d of the node-major frontier goes to the owner of those subtrees ---------
    PF_CUDA_OK(cudaMemsetAsync(S->d_mine, 0, (size_t)(2 * nr + 2) * 8, s));
    if (st.n)
        frontier_bounds_kernel<<<1, 64, 0, s>>>(db->fr_node[st.cur].p, (uint32_t)st.n, S->d_cut_lo, (uint32_t)nr, S->d_mine);
    if ((rc = gather_words(db, (size_t)(2 * nr + 2)))) return rc;
    std::vector<uint64_t> cnt((size_t)nr * nr), s_off(nr), r_off(nr);
    const size_t row = (size_t)(2 * nr + 2);
    for (int a = 0; a < nr; ++a)
        for (int b = 0; b < nr; ++b) cnt[(size_t)a * nr + b] = S->h_all[a * row + b];
    for (int p = 0; p < nr; ++p) s_off[p] = S->h_all[me * row + nr + 1 + p];
    uint64_t n_recv = 0;
    for (int p = 0; p < nr; ++p) {
        r_off[p] = n_recv;
        n_recv += cnt[(size_t)p * nr + me];
    }
    if (n_recv > 0xFFFFFFF0ULL) {
        set_error("frontier of %llu pairs exceeds the 32-bit pair index: use smaller read blocks", (unsigned long long)n_recv);
        return PF_ERR_NOMEM;
    }
    {
        const int nxt = st.cur ^ 1;
        if ((rc = db->fr_read[nxt].ensure(std::max<uint64_t>(n_recv, 1))) || (rc = db->fr_node[nxt].ensure(std::max<uint64_t>(n_recv, 1))))
            return rc;
        const uint32_t *snd[2] = {db->fr_read[st.cur].p, db->fr_node[st.cur].p};
        uint32_t *rcv[2] = {db->fr_read[nxt].p, db->fr_node[nxt].p};
        if ((rc = exchange_slices(db, 2, snd, rcv, s_off, r_off, cnt))) return rc;
        S->stats.pairs_sent += st.n - cnt[(size_t)me * nr + me];
        S->stats.pairs_received += n_recv - cnt[(size_t)me * nr + me];
        st.cur = nxt;
        st.n = n_recv;
    }

    // ---- phase B: this rank's subtrees, the reads of every rank ---------------------------------------------
    if (Lc < n_levels) {
        const bool entries_below = db->entry_start[n_levels] > db->entry_start[Lc];
        if (entries_below) {  // the plan skips the top above some owned subtree: every read is paired with its entry nodes
            if ((rc = hash_range(db, g, 0, r_me, nullptr, st)) || (rc = hash_range(db, g, r_me + n_me, nT - r_me - n_me, nullptr, st)))
                return rc;
        } else if (st.n) {  // hash only the foreign reads the received pairs name
            if ((rc = S->need_hash.ensure(nT))) return rc;
            PF_CUDA_OK(cudaMemsetAsync(S->need_hash.p, 0, nT, s));
            mark_reads_kernel<<<(uint32_t)((st.n + 255) / 256), 256, 0, s>>>(db->fr_read[st.cur].p, (uint32_t)st.n, S->need_hash.p);
            if (n_me) PF_CUDA_OK(cudaMemsetAsync(S->need_hash.p + r_me, 0, n_me, s));  // own reads are hashed already
            st.other_launches++;
            if ((rc = hash_range(db, g, 0, nT, S->need_hash.p, st))) return rc;
        }
        if ((rc = run_levels(db, &g, threshold, want_hits ? 2 : 0, G, 0, Lc, n_levels, 0, nT, st))) return shard_rc(rc);
    }
    add_counts_kernel<<<(uint32_t)((db->n_leaves + 255) / 256), 256, 0, s>>>(db->d_counts, db->d_blk_counts, (uint32_t)db->n_leaves);
    st.other_launches++;
    uint64_t d2h = st.levels * sizeof(LevelTotals);

    // ---- hits go home: (read, leaf) to the rank that owns the read ----------------------------------------
    uint64_t hits_final = hits_a;
    if (want_hits) {
        const uint64_t hits_b = st.hits_total - hits_a;
        if (hits_b > 0xFFFFFFF0ULL) {
            set_error("%llu hits in one block exceed the 32-bit hit index: use smaller read blocks", (unsigned long long)hits_b);
            return PF_ERR_NOMEM;
        }
        PF_CUDA_OK(cudaMemsetAsync(S->d_mine, 0, (size_t)(2 * nr + 2) * 8, s));
        PF_CUDA_OK(cudaMemsetAsync(S->d_cursor, 0, (size_t)nr * 4, s));
        if (hits_b) {
            if ((rc = S->send_read.ensure(hits_b)) || (rc = S->send_leaf.ensure(hits_b))) return rc;
            const uint32_t nbk = (uint32_t)((hits_b + 255) / 256);
            hit_owner_count_kernel<<<nbk, 256, 0, s>>>(db->hit_read.p + hits_a, (uint32_t)hits_b, S->d_rbase, (uint32_t)nr, S->d_mine);
            hit_offsets_kernel<<<1, 32, 0, s>>>(S->d_mine, (uint32_t)nr, S->d_mine + nr + 1);
            hit_partition_kernel<<<nbk, 256, 0, s>>>(db->hit_read.p + hits_a, db->hit_leaf.p + hits_a, (uint32_t)hits_b, S->d_rbase,
                                                    (uint32_t)nr, S->d_mine + nr + 1, S->d_cursor, S->send_read.p, S->send_leaf.p);
            st.other_launches += 3;
        }
        if ((rc = gather_words(db, row))) return rc;
        for (int a = 0; a < nr; ++a)
            for (int b = 0; b < nr; ++b) cnt[(size_t)a * nr + b] = S->h_all[a * row + b];
        for (int p = 0; p < nr; ++p) s_off[p] = S->h_all[me * row + nr + 1 + p];
        uint64_t h_recv = 0;
        for (int p = 0; p < nr; ++p) {
            r_off[p] = hits_a + h_recv;
            h_recv += cnt[(size_t)p * nr + me];
        }
        hits_final = hits_a + h_recv;
        if (hits_final > 0xFFFFFFF0ULL) {
            set_error("%llu hits in one block exceed the 32-bit hit index: use smaller read blocks", (unsigned long long)hits_final);
            return PF_ERR_NOMEM;
        }
        if (hits_final && ((rc = db->hit_read.grow_keep(hits_final, hits_a, s)) || (rc = db->hit_leaf.grow_keep(hits_final, hits_a, s))))
            return rc;
        const uint32_t *snd[2] = {S->send_read.p, S->send_leaf.p};
        uint32_t *rcv[2] = {db->hit_read.p, db->hit_leaf.p};
        if ((rc = exchange_slices(db, 2, snd, rcv, s_off, r_off, cnt))) return rc;
        S->stats.hits_sent += hits_b - cnt[(size_t)me * nr + me];
        if (h_recv) {
            count_hits_kernel<<<(uint32_t)((h_recv + 255) / 256), 256, 0, s>>>(db->hit_read.p + hits_a, (uint32_t)h_recv, db->read_hits.p);
            st.other_launches++;
        }
        if ((rc = finish_csr(db, nT, hits_final, 1, r_me, n_me, out, &st.other_launches, &d2h))) return rc;
    }
    PF_CUDA_OK(cudaEventRecord(db->ev_end, s));
    PF_CUDA_OK(cudaStreamSynchronize(s));
    PF_CUDA_OK(cudaGetLastError());
    if (out) {
        if (want_hits) {
            out->n_hits = hits_final;
            out->read_off = db->pin_off.p;
            out->leaf = db->pin_leaf.p;
        } else {
            out->read_off = db->out_off.data();
        }
    }
    account_stats(db, st, n_me, d2h);
    S->stats.queries++;
    S->stats.pairs_top += pairs_a;
    S->stats.pairs_subtrees += st.pairs - pairs_a;
    S->stats.reads_gathered += nT;
    return PF_OK;
}

}  // namespace pf

// =================================== C ABI ======================================================
extern "C" {

int pf_shard_plan(const char *db_path, int64_t search_depth, int nranks, int64_t cut_level, uint32_t *cut_level_out,
                  int32_t *owner_out, uint64_t owner_cap, uint64_t *n_nodes_out) {
    if (!db_path || nranks < 1 || nranks > 32) {
        set_error("pf_shard_plan: bad argument");
        return PF_ERR_ARG;
    }
    pf_db tmp;  // host fields only; no CUDA call is made
    std::string err;
    if (!read_tree_bin(join_path(db_path, "tree.bin"), tmp.tree, err)) {
        set_error("%s", err.c_str());
        return err.rfind("cannot open", 0) == 0 ? PF_ERR_IO : PF_ERR_FORMAT;
    }
    flatten(&tmp, search_depth);
    if (tmp.n_nodes == 0) {
        set_error("database has no root node");
        return PF_ERR_FORMAT;
    }
    ShardPlan p;
    plan_shards(tmp.level_start, tmp.h_left, tmp.h_right, nranks, cut_level, p);
    if (cut_level_out) *cut_level_out = p.cut_level;
    if (n_nodes_out) *n_nodes_out = tmp.n_nodes;
    if (owner_out) {
        if (owner_cap < tmp.n_nodes) {
            set_error("pf_shard_plan: owner_out holds %llu entries, the tree has %llu nodes", (unsigned long long)owner_cap,
                      (unsigned long long)tmp.n_nodes);
            return PF_ERR_ARG;
        }
        memcpy(owner_out, p.owner.data(), tmp.n_nodes * 4);
    }
    return PF_OK;
}

int pf_db_open_sharded(const char *db_path, int device, int64_t search_depth, int nranks, int rank, const void *id128,
                       int64_t cut_level, pf_db **out) {
    if (!db_path || !out || !id128 || nranks < 1 || nranks > 32 || rank < 0 || rank >= nranks) {
        set_error("pf_db_open_sharded: bad argument");
        return PF_ERR_ARG;
    }
    *out = nullptr;
    pf_db *db = new pf_db();
    db->device = device;
    db->sharded = 1;
    db->rank = rank;
    db->nranks = nranks;
    db->cut_level_req = cut_level;
    db->nccl_id.assign((const uint8_t *)id128, (const uint8_t *)id128 + 128);
    int rc = db_open_impl(db, db_path, search_depth);
    if (rc == PF_OK) rc = shard_state_init(db);
    if (rc != PF_OK) {
        std::string keep = pf_last_error();
        db_free(db);
        set_error("%s", keep.c_str());
        return rc;
    }
    *out = db;
    return PF_OK;
}

int pf_shard_info(const pf_db *db, pf_shard_info_t *o) {
    if (!db || !o) {
        set_error("pf_shard_info: null argument");
        return PF_ERR_ARG;
    }
    memset(o, 0, sizeof *o);
    o->sharded = db->sharded;
    o->rank = db->rank;
    o->nranks = db->nranks;
    o->cut_level = db->cut_level;
    for (size_t u = 0; u < db->n_nodes; ++u) {
        const int32_t ow = db->h_owner.empty() ? -1 : db->h_owner[u];
        if (ow < 0) o->top_nodes++;
        else if (ow == db->rank) o->owned_nodes++;
    }
    o->resident_filters = db->n_slots;
    o->resident_bytes = db->n_slots * db->wpf * 8;
    return PF_OK;
}

int pf_shard_stats(pf_db *db, pf_shard_stats_t *o) {
    if (!db || !o || !db->shard) {
        set_error("pf_shard_stats: not a sharded handle");
        return PF_ERR_ARG;
    }
    *o = db->shard->stats;
    return PF_OK;
}

int pf_query_sharded(pf_db *db, const pf_read_batch *in, float threshold, int want_hits, pf_hits *out) {
    if (!db || !in) {
        set_error("pf_query_sharded: null argument");
        return PF_ERR_ARG;
    }
    if (!db->sharded || !db->shard) {
        set_error("pf_query_sharded: the handle was not opened with pf_db_open_sharded");
        return PF_ERR_STATE;
    }
    int rc = batch_upload_impl(db, in, &db->own_batch, db->stream);
    if (rc != PF_OK) return rc;
    return query_sharded_impl(db, &db->own_batch, threshold, want_hits, out);
}

int pf_query_sharded_device(pf_db *db, pf_dev_batch *batch, float threshold, int want_hits, pf_hits *out) {
    if (!db || !batch) {
        set_error("pf_query_sharded_device: null argument");
        return PF_ERR_ARG;
    }
    if (!db->sharded || !db->shard) {
        set_error("pf_query_sharded_device: the handle was not opened with pf_db_open_sharded");
        return PF_ERR_STATE;
    }
    PF_CUDA_OK(cudaSetDevice(db->device));
    if (batch->ready) PF_CUDA_OK(cudaStreamWaitEvent(db->stream, batch->ready, 0));
    return query_sharded_impl(db, batch, threshold, want_hits, out);
}

}  // extern "C"
