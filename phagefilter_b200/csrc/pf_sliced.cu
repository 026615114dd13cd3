// pf_sliced.cu -- host side of the bit-sliced evaluation (pf_sliced.cuh): cutting the tree into tiles of up to 256
// nodes, building the tiles' transposed tables from the resident filters, choosing between the node-at-a-time
// descent (pf_query.cu) and the sliced one, and the tile-level descent itself.
//
// Reference semantics reproduced (paths relative to the reference root): query::_query_batch (query.rs:99-158) --
// a node is evaluated for a read iff every ancestor passed, a leaf that passes counts the read (query.rs:143) and
// records (read, genome) (query.rs:149-153); query::query_passes (query.rs:38-49); BloomFilter::contains
// (bloom_filter.rs:312-332).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <deque>

#include "pf_db.h"
#include "pf_sliced.cuh"

namespace pf {

struct SlicedState {
    // plan (host)
    std::vector<uint8_t> skip;             // per node: not evaluated (verified superset chain below, passes ~always)
    std::vector<SlicedTileDev> tiles;
    std::vector<uint32_t> col_slot;        // [tiles][256] filter slot per column
    std::vector<uint32_t> child_tile, child_mask;
    std::vector<uint32_t> entry_tiles;
    std::vector<int32_t> tile_parent;
    std::vector<std::vector<uint32_t>> tile_nodes;  // node of every column
    std::vector<int32_t> tile_group;       // per tile: the group of entry tiles it shares 128-byte lines with, or -1
    uint32_t n_line_tiles = 0;             // leading entries of entry_tiles laid out that way (SL_QUAD per group)
    uint64_t table_words = 0;
    uint64_t entry_bytes = 0;
    double est_sectors_per_read = 0.0;     // cost model: expected sector loads of a read unrelated to the database
    double est_seconds_per_read = 0.0;     // ... and the time they take at the measured random-row rates
    double est_seconds_related = 0.0;      // the same for a read that belongs to a genome of the database
    // device
    SlicedTileDev *d_tiles = nullptr;
    uint32_t *d_tables = nullptr, *d_child_tile = nullptr, *d_child_mask = nullptr, *d_entry = nullptr;
    uint32_t *d_tile_count = nullptr, *d_tile_cursor = nullptr;  // contiguous [2 * n_tiles]
    unsigned long long *d_tile_base = nullptr, *d_counters = nullptr, *d_hit_cursor = nullptr;
    unsigned int *d_work = nullptr;
    uint32_t *h_tile_count = nullptr;            // pinned [n_tiles]
    unsigned long long *h_counters = nullptr;    // pinned [5]
    unsigned long long *h_tile_base = nullptr;   // pinned [n_tiles]
    bool tables_ready = false;
    bool entry_lean = false;  // every entry tile is filter-only with a pre-test of at most 2 steps
    uint32_t n_quad_now = 0;  // leading entry positions the line kernel takes under the current threshold (whole groups)
    // hybrid: tiles for the cut only; what survives them is handed to the node-at-a-time descent as (read, node) pairs
    bool hybrid = true;
    uint32_t *d_node_inj_count = nullptr, *d_node_inj_cursor = nullptr;  // contiguous [2 * n_nodes]
    unsigned long long *d_node_inj_base = nullptr, *d_inj_bsum = nullptr;  // [n_nodes + 1], block sums of the scan
    unsigned long long *h_node_inj_base = nullptr;                        // pinned copy
    DevBuf<uint32_t> inj_read, inj_node;
    DevBuf<uint32_t> fr_read[2], fr_tile[2], fr_src[2], reach[2], alive;
    // what the plan was made for
    float theta = -1.f;
    uint64_t n_nominal = 0;
    int decided_mode = 0;  // 1 pair, 2 sliced (for theta / n_nominal above)
    int decided_under = -1;  // pf_db_set_mode value the decision was taken under
    double decided_rho = 0.5; // related share the decision was taken for
    int decided_handover = -2; // pf_db_set_handover value the decision was taken under
    uint32_t tile_cols = 0;    // pf_db_set_tile_cols value the tiling was made with
    bool failed = false;   // tables could not be built (memory): stay with the node-at-a-time path
    bool deep_failed = false;  // the full set of tables did not fit: cut-only tiles with hand-over from now on
};

static void sliced_release_device(SlicedState *s) {
    cudaFree(s->d_tiles);
    cudaFree(s->d_tables);
    cudaFree(s->d_child_tile);
    cudaFree(s->d_child_mask);
    cudaFree(s->d_entry);
    cudaFree(s->d_tile_count);
    cudaFree(s->d_tile_base);
    cudaFree(s->d_counters);
    cudaFree(s->d_work);
    cudaFree(s->d_node_inj_count);
    cudaFree(s->d_node_inj_base);
    cudaFree(s->d_inj_bsum);
    if (s->h_node_inj_base) cudaFreeHost(s->h_node_inj_base);
    s->d_node_inj_count = s->d_node_inj_cursor = nullptr;
    s->d_node_inj_base = s->d_inj_bsum = nullptr;
    s->h_node_inj_base = nullptr;
    if (s->h_tile_count) cudaFreeHost(s->h_tile_count);
    if (s->h_counters) cudaFreeHost(s->h_counters);
    if (s->h_tile_base) cudaFreeHost(s->h_tile_base);
    s->d_tiles = nullptr;
    s->d_tables = s->d_child_tile = s->d_child_mask = s->d_entry = s->d_tile_count = s->d_tile_cursor = nullptr;
    s->d_tile_base = s->d_counters = s->d_hit_cursor = nullptr;
    s->d_work = nullptr;
    s->h_tile_count = nullptr;
    s->h_counters = s->h_tile_base = nullptr;
    s->tables_ready = false;
}

void sliced_free(pf_db *db) {
    if (!db->sliced) return;
    SlicedState *s = db->sliced;
    sliced_release_device(s);
    for (int i = 0; i < 2; i++) {
        s->fr_read[i].release();
        s->fr_tile[i].release();
        s->fr_src[i].release();
        s->reach[i].release();
    }
    s->alive.release();
    s->inj_read.release();
    s->inj_node.release();
    delete s;
    db->sliced = nullptr;
}

// ---- plan ------------------------------------------------------------------------------------------------------------
static double phi_tab(double z) {  // standard normal CDF, coarse: only steers cost decisions, never results
    return 0.5 * erfc(-z / 1.4142135623730951);
}
static uint64_t need_host(float threshold, uint64_t n) {  // query.rs:48
    volatile float prod = threshold * (float)n;
    const float c = ceilf(prod);
    if (!(c > 0.0f)) return 0;
    if (c >= 18446744073709551616.0f) return ~0ULL;
    return (uint64_t)c;
}
// measured on a B200 (profiles/r2a_mb_gather_b200.csv): random row loads per second against the bytes they range over
static double sector_rate(double footprint_bytes) {
    static const double mb[] = {64, 80, 96, 115, 128, 160, 230, 460, 920, 2048, 8192};
    static const double g[] = {289, 262, 159, 128, 117, 102, 80, 51.6, 44.4, 39.6, 37.1};
    const double f = footprint_bytes / 1048576.0;
    if (f <= mb[0]) return g[0] * 1e9;
    for (int i = 1; i < 11; ++i)
        if (f <= mb[i]) return (g[i - 1] + (g[i] - g[i - 1]) * (f - mb[i - 1]) / (mb[i] - mb[i - 1])) * 1e9;
    return 36.5e9;
}
static uint32_t width_for(size_t cols) { return cols <= 32 ? 32u : (cols <= 64 ? 64u : (cols <= 128 ? 128u : 256u)); }

// Per-node facts the tiling works from (host only).
struct TreeFacts {
    std::vector<uint32_t> parent, sz, leaves;  // subtree size in nodes / in tree leaves
    std::vector<uint8_t> vb;                   // every (node, child) pair below was verified a bitwise superset
    std::vector<double> fill, p, q;            // p: a k-mer unrelated to the node passes it; q: an unrelated read passes it
    double n = 0, allowed = 0;
    uint64_t need = 0;
};
static void tree_facts(const pf_db *db, float threshold, uint64_t n_nominal, TreeFacts &F) {
    const size_t nn = db->n_nodes;
    const uint32_t K = db->geom.num_hashes;
    const double mbits = (double)db->geom.num_bits;
    F.parent.assign(nn, NONE32);
    for (size_t u = 0; u < nn; ++u) {
        if (db->h_left[u] != NONE32) F.parent[db->h_left[u]] = (uint32_t)u;
        if (db->h_right[u] != NONE32) F.parent[db->h_right[u]] = (uint32_t)u;
    }
    F.need = need_host(threshold, n_nominal);
    F.n = (double)n_nominal;
    F.allowed = F.need > n_nominal ? 0.0 : (double)(n_nominal - F.need);
    F.vb.assign(nn, 1);
    F.sz.assign(nn, 1);
    F.leaves.assign(nn, 0);
    F.p.assign(nn, 0.0);
    F.fill.assign(nn, 0.0);
    F.q.assign(nn, 0.0);
    for (size_t u = nn; u-- > 0;) {
        const double fill = (double)db->h_pop[u] / mbits;
        const double p = pow(fill, (double)K), mean = F.n * p, var = F.n * p * (1.0 - p);
        F.p[u] = p;
        F.fill[u] = fill;
        F.q[u] = var < 1e-9 ? (mean + 0.5 >= (double)F.need ? 1.0 : 0.0) : phi_tab((mean - (double)F.need + 0.5) / sqrt(var));
        if (db->h_leaf[u] >= 0) {
            F.leaves[u] = 1;
            continue;
        }
        F.vb[u] = db->h_mono[u];
        for (uint32_t c : {db->h_left[u], db->h_right[u]})
            if (c != NONE32) {
                F.vb[u] = F.vb[u] && F.vb[c];
                F.sz[u] += F.sz[c];
                F.leaves[u] += F.leaves[c];
            }
    }
}

// Where the tiles' tables lie.  Entry tiles that can serve as pure filters for some threshold (no leaf column, every
// column's subtree verified -- a property of the tiling, not of the threshold) come first in entry order and, SL_QUAD at a
// time, share one table whose rows are 128-byte lines: 32 bytes (one sector, 256 columns) per tile, so that one line load
// answers a probe step for all of them (sliced_entry_quad_kernel).  A group needs at least two tiles; every other tile
// keeps a table of its own with rows of row_words words.
static void layout_tables(const pf_db *db, const TreeFacts &F, SlicedState &S) {
    const uint64_t rows = 64ULL * db->wpf;
    const size_t nt = S.tiles.size();
    std::vector<uint8_t> filt_ok(nt, 0);
    for (uint32_t t : S.entry_tiles) {
        bool ok = true;
        for (uint32_t u : S.tile_nodes[t]) ok = ok && F.vb[u] && db->h_leaf[u] < 0;
        filt_ok[t] = ok;
    }
    std::stable_partition(S.entry_tiles.begin(), S.entry_tiles.end(), [&](uint32_t t) { return filt_ok[t] != 0; });
    size_t n_ok = 0;
    while (n_ok < S.entry_tiles.size() && filt_ok[S.entry_tiles[n_ok]]) ++n_ok;
    if (n_ok % SL_QUAD == 1) --n_ok;  // a tile alone in its group gains nothing from a 128-byte row
    static const bool no_lines = getenv("PF_SLICED_NO_LINES") != nullptr;
    if (no_lines) n_ok = 0;
    S.n_line_tiles = (uint32_t)n_ok;
    S.tile_group.assign(nt, -1);
    S.table_words = 0;
    S.entry_bytes = 0;
    for (size_t e = 0; e < n_ok; ++e) {
        SlicedTileDev &tm = S.tiles[S.entry_tiles[e]];
        S.tile_group[S.entry_tiles[e]] = (int32_t)(e / SL_QUAD);
        tm.table_off = S.table_words + 8ULL * (e % SL_QUAD);
        tm.row_stride = 32u;
        if (e % SL_QUAD == SL_QUAD - 1 || e + 1 == n_ok) S.table_words += rows * 32ULL;
    }
    S.entry_bytes = S.table_words * 4ULL;
    for (size_t t = 0; t < nt; ++t) {
        SlicedTileDev &tm = S.tiles[t];
        if (S.tile_group[t] >= 0) continue;
        tm.table_off = S.table_words;
        tm.row_stride = tm.row_words;
        S.table_words += rows * tm.row_words;
        if (tm.entry) S.entry_bytes += rows * tm.row_words * 4ULL;
    }
}

// Cuts the (pruned, level-ordered) tree into tiles given the set of nodes that are not evaluated.
//  * Skipped nodes: interior, connected to the root, and every (node, child) pair below them was verified at load time
//    to be a bitwise superset (analyse_tree), so whatever passes a leaf below them passes them too: evaluating the
//    nodes of the cut exactly selects exactly the reads the reference's descent would let through.
//  * Every other node gets a column in exactly one tile.  A tile's roots all hang below columns of ONE parent tile (or
//    below skipped nodes: entry tiles, evaluated for every read).  Entry tiles hold the nodes of the cut side by side,
//    256 per tile (the rest in the narrowest width that fits, free columns filled with the next level), so that every
//    read touches as few tiles as possible; below them whole subtrees are packed, several per tile, so that a read that
//    survives needs one more tile, not one per level.  entry_only: no tiles below the cut -- the children of the
//    entry tiles' terminal columns are recorded for the hand-over to the node-at-a-time descent instead.
// Also fills the cost model: expected seconds of sector loads for a read unrelated to the database.
static void tile_tree(const pf_db *db, const TreeFacts &F, SlicedState &S, bool entry_only, double pair_related_s) {
    const size_t nn = db->n_nodes;
    const uint32_t K = db->geom.num_hashes;
    const size_t MAXC = std::min<size_t>(std::max<uint32_t>(db->tile_cols, 32u), SL_MAX_COLS);  // columns per tile
    // tiles below the cut: a narrower tile is a smaller table (PF_SLICED_SUB_COLS, experiment)
    static const size_t sub_env = getenv("PF_SLICED_SUB_COLS") ? (size_t)atoi(getenv("PF_SLICED_SUB_COLS")) : 0;
    const size_t MAXS = sub_env >= 32 ? std::min(MAXC, sub_env) : MAXC;
    S.tiles.clear();
    S.col_slot.clear();
    S.child_tile.clear();
    S.child_mask.clear();
    S.entry_tiles.clear();
    S.tile_parent.clear();
    S.tile_nodes.clear();
    std::vector<int32_t> node_tile(nn, -1);
    std::vector<uint32_t> node_col(nn, 0);
    struct Job {
        int32_t parent_tile;
        std::vector<uint32_t> roots;
    };
    std::deque<Job> jobs;
    {
        Job j0{-1, {}};
        for (size_t u = 0; u < nn; ++u)
            if (!S.skip[u] && (u == 0 || S.skip[F.parent[u]])) j0.roots.push_back((uint32_t)u);
        jobs.push_back(std::move(j0));
    }
    auto bfs_subtree = [&](uint32_t r, size_t cap, std::vector<uint32_t> &out) {  // level order, parents first
        const size_t b = out.size();
        out.push_back(r);
        for (size_t i = b; i < out.size(); ++i)
            for (uint32_t c : {db->h_left[out[i]], db->h_right[out[i]]})
                if (c != NONE32 && out.size() - b < cap) out.push_back(c);
    };
    while (!jobs.empty()) {
        Job job = std::move(jobs.front());
        jobs.pop_front();
        std::vector<std::vector<uint32_t>> made;  // column lists of the tiles of this job
        std::vector<uint32_t> RA, RB;
        // entry tiles see every read: roots side by side, as few tiles as possible.  Below them come the reads that
        // (mostly) belong there: whole subtrees, several per tile, so that such a read needs one more tile only.
        for (uint32_t r : job.roots) (job.parent_tile < 0 ? RA : RB).push_back(r);
        // nodes that can serve as pure filters (interior, verified subtree) first: the entry tiles made of them alone can
        // share 128-byte lines (layout_tables); leaves and unverified nodes of the cut gather in the last tiles
        auto filt_node = [&](uint32_t u) { return db->h_leaf[u] < 0 && F.vb[u]; };
        std::stable_partition(RA.begin(), RA.end(), filt_node);
        for (size_t c0 = 0; c0 < RA.size(); c0 += MAXC) {
            std::vector<uint32_t> cols(RA.begin() + c0, RA.begin() + std::min(RA.size(), c0 + MAXC));
            const size_t cap = std::min<size_t>(width_for(cols.size()), MAXC);
            bool pure = true;
            for (uint32_t u : cols) pure = pure && filt_node(u);
            for (size_t i = 0; i < cols.size() && cols.size() < cap; ++i)
                for (uint32_t c : {db->h_left[cols[i]], db->h_right[cols[i]]})
                    if (c != NONE32 && cols.size() < cap && (!pure || filt_node(c))) cols.push_back(c);
            made.push_back(std::move(cols));
        }
        std::vector<uint32_t> cur;
        for (uint32_t r : RB) {
            if (F.sz[r] <= (uint32_t)MAXS) {
                if (cur.size() + F.sz[r] > MAXS) {
                    made.push_back(std::move(cur));
                    cur.clear();
                }
                bfs_subtree(r, MAXS, cur);
            } else {
                if (!cur.empty()) {
                    made.push_back(std::move(cur));
                    cur.clear();
                }
                std::vector<uint32_t> cols;
                bfs_subtree(r, MAXS, cols);
                made.push_back(std::move(cols));
            }
        }
        if (!cur.empty()) made.push_back(std::move(cur));
        const uint32_t first_link = (uint32_t)S.child_tile.size();
        for (auto &cols : made) {
            const uint32_t t = (uint32_t)S.tiles.size();
            SlicedTileDev tm;
            memset(&tm, 0, sizeof tm);
            tm.n_cols = (uint32_t)cols.size();
            tm.row_words = width_for(cols.size()) / 32u;
            tm.entry = job.parent_tile < 0;
            S.col_slot.resize((size_t)(t + 1) * SL_MAX_COLS, NONE32);
            for (size_t c = 0; c < cols.size(); ++c) {
                node_tile[cols[c]] = (int32_t)t;
                node_col[cols[c]] = (uint32_t)c;
            }
            std::vector<uint32_t> depth(cols.size(), 0);
            uint32_t link_mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (size_t c = 0; c < cols.size(); ++c) {
                const uint32_t u = cols[c];
                S.col_slot[(size_t)t * SL_MAX_COLS + c] = db->h_slot[u];
                tm.valid[c >> 5] |= 1u << (c & 31);
                tm.leaf[c] = db->h_leaf[u];
                if (db->h_leaf[u] >= 0) tm.leafmask[c >> 5] |= 1u << (c & 31);
                const uint32_t p = F.parent[u];
                if (p != NONE32 && node_tile[p] == (int32_t)t && node_col[p] < c) {
                    tm.parent[c] = (uint16_t)node_col[p];
                    depth[c] = depth[node_col[p]] + 1;
                    tm.prop_iters = std::max(tm.prop_iters, depth[c]);
                } else if (job.parent_tile < 0) {
                    tm.parent[c] = 0xFFFFu;
                } else {
                    tm.parent[c] = (uint16_t)(0x8000u | node_col[p]);
                    link_mask[node_col[p] >> 5] |= 1u << (node_col[p] & 31);
                }
            }
            for (size_t c = cols.size(); c < (size_t)SL_MAX_COLS; ++c) {
                tm.parent[c] = 0xFFFFu;
                tm.leaf[c] = -1;
            }
            for (size_t c = 0; c < (size_t)SL_MAX_COLS; ++c) tm.child_node[c][0] = tm.child_node[c][1] = NONE32;
            if (job.parent_tile >= 0) {
                S.child_tile.push_back(t);
                S.child_mask.insert(S.child_mask.end(), link_mask, link_mask + 8);
            } else {
                S.entry_tiles.push_back(t);
            }
            S.tiles.push_back(tm);
            S.tile_parent.push_back(job.parent_tile);
            S.tile_nodes.push_back(cols);
        }
        if (job.parent_tile >= 0) {
            S.tiles[job.parent_tile].first_child = first_link;
            S.tiles[job.parent_tile].n_children = (uint32_t)S.child_tile.size() - first_link;
        }
        // the tiles of this job are complete: their out-of-tile children are the roots of the next jobs
        const uint32_t t_first = (uint32_t)(S.tiles.size() - made.size());
        for (size_t k = 0; k < made.size(); ++k) {
            const uint32_t t = t_first + (uint32_t)k;
            Job nj{(int32_t)t, {}};
            for (size_t c = 0; c < made[k].size(); ++c) {
                const uint32_t u = made[k][c];
                bool term = db->h_leaf[u] >= 0;
                int k = 0;
                S.tiles[t].child_node[c][0] = S.tiles[t].child_node[c][1] = NONE32;
                for (uint32_t ch : {db->h_left[u], db->h_right[u]})
                    if (ch != NONE32 && node_tile[ch] != (int32_t)t) {
                        term = true;
                        nj.roots.push_back(ch);
                        if (entry_only) S.tiles[t].child_node[c][k++] = ch;
                    }
                if (term) S.tiles[t].terminal[c >> 5] |= 1u << (c & 31);
            }
            std::sort(nj.roots.begin(), nj.roots.end());
            if (!nj.roots.empty() && !entry_only) jobs.push_back(std::move(nj));
        }
    }
    layout_tables(db, F, S);
    // Cost model (steers choices only, never results).  A read unrelated to the database leaves a tile once every
    // terminal column has more than `allowed` k-mers proven absent.  With s probe steps per k-mer, a terminal of fill f
    // proves an unrelated k-mer absent with probability r = 1 - f^s, so it needs about (allowed + 1) / r k-mers (plus two
    // standard deviations), in rounds of 32 k-mers of s row loads.  Per tile the pre-test depth s (0 = none) is chosen
    // to minimise the expected row loads, assuming half of the reads that arrive belong below the tile (they pay the
    // pre-test on top of the exact pass).  A tile below another one is reached by an unrelated read with the
    // probability that one of the columns above its roots passes.
    {
        const size_t nt = S.tiles.size();
        std::vector<double> pr(nn, 0.0);  // per node: probability that an unrelated read reaches and passes it
        for (size_t u = 0; u < nn; ++u) pr[u] = S.skip[u] ? 1.0 : (u == 0 ? 1.0 : pr[F.parent[u]]) * F.q[u];
        double total_s = 0.0, total_sectors = 0.0;
        std::vector<double> group_unrel((S.n_line_tiles + SL_QUAD - 1) / SL_QUAD, 0.0);
        const double line_bytes = (double)(64ULL * db->wpf) * 128.0;
        uint32_t max_depth = 0;
        std::vector<uint32_t> tdepth(nt, 0);
        const double a1 = F.allowed + 1.0, n = F.n;
        for (size_t t = 0; t < nt; ++t) {
            SlicedTileDev &tm = S.tiles[t];
            if (S.tile_parent[t] >= 0) tdepth[t] = tdepth[S.tile_parent[t]] + 1;
            max_depth = std::max(max_depth, tdepth[t]);
            double reach = 1.0;
            if (S.tile_parent[t] >= 0) {
                double none = 1.0;
                for (uint32_t c = 0; c < tm.n_cols; ++c)
                    if (tm.parent[c] & 0x8000u) none *= 1.0 - pr[F.parent[S.tile_nodes[t][c]]];
                reach = 1.0 - none;
            }
            double fmax = 0.0;  // the terminal that holds out longest is the fullest one
            std::vector<double> tf;
            for (uint32_t c = 0; c < tm.n_cols; ++c)
                if ((tm.terminal[c >> 5] >> (c & 31)) & 1u) {
                    fmax = std::max(fmax, F.fill[S.tile_nodes[t][c]]);
                    tf.push_back(F.fill[S.tile_nodes[t][c]]);
                }
            fmax = std::min(fmax, 0.999999);
            std::sort(tf.begin(), tf.end());
            const double f90 = tf.empty() ? fmax : std::min(tf[(tf.size() * 9) / 10 == tf.size() ? tf.size() - 1 : (tf.size() * 9) / 10], 0.999999);
            auto kmers_to_die = [&](uint32_t s) {  // k-mers until the fullest terminal is ruled out, or -1: it survives
                const double r = 1.0 - pow(fmax, (double)s);
                const double j = (a1 + 2.0 * sqrt(a1 * (1.0 - r))) / r;
                if (j > n) return -1.0;
                return F.allowed == 0.0 ? std::max(16.0, j) : ceil(j / 32.0) * 32.0;
            };
            // Filter-only tile: no leaf column and every column's subtree verified (child passes => node passes all the
            // way down).  Such a tile only has to weed out reads: what its pre-test cannot rule out goes on to the tiles
            // below, which are evaluated exactly -- a leaf that passes there implies that these columns pass too.
            bool filt_ok = true;
            for (uint32_t c = 0; c < tm.n_cols && filt_ok; ++c)
                filt_ok = F.vb[S.tile_nodes[t][c]] && db->h_leaf[S.tile_nodes[t][c]] < 0;
            const double jK = kmers_to_die(K);
            const double exact_unrel = (jK < 0 ? n : jK) * (double)K, exact_rel = n * (double)K;
            double best = 0.5 * exact_unrel + 0.5 * exact_rel, best_unrel = exact_unrel;
            uint32_t best_s = 0;
            const bool pretest_useful = tm.entry || reach > 0.5;
            for (uint32_t s = 1; pretest_useful && s < K && s <= (uint32_t)SL_STEP_BATCH; ++s) {
                const double j = kmers_to_die(s);
                const double unrel = j < 0 ? n * (double)s + exact_unrel : j * (double)s;
                const double c = 0.5 * unrel + 0.5 * (n * (double)s + (filt_ok ? 0.0 : exact_rel));
                if (c < best) {
                    best = c;
                    best_unrel = unrel;
                    best_s = s;
                }
            }
            if (const char *f = getenv("PF_SLICED_FORCE_PRE")) {  // tests: a given pre-test depth on every entry tile
                const uint32_t v = (uint32_t)atoi(f);
                if (tm.entry && v < K && v <= (uint32_t)SL_STEP_BATCH) best_s = v;
            }
            tm.pre_steps = best_s;
            tm.filter_only = best_s != 0 && filt_ok;
            {   // rounds in flight before the first look at the columns: when 90 % of the terminals are expected to be settled
                double j90 = 0.0;
                if (best_s) {
                    const double r = 1.0 - pow(f90, (double)best_s);
                    j90 = (a1 + sqrt(a1 * (1.0 - r))) / r;
                }
                tm.pre_rounds = (uint32_t)std::min(4.0, std::max(1.0, floor(j90 / 32.0 + 0.75)));
            }
            const double bytes = (double)(64ULL * db->wpf) * tm.row_stride * 4.0;
            // entry tiles are worked tile-major, one table hot at a time; deeper tiles are touched at random
            const double rate = tm.entry ? sector_rate(bytes) : sector_rate(1e12);
            const bool in_lines = S.tile_group[t] >= 0 && best_s >= 1 && best_s <= 2 && filt_ok;
            if (in_lines) {
                // tiles that share lines are worked together: one line load per k-mer and step answers all of them, the
                // tile that holds out longest decides
                group_unrel[S.tile_group[t]] = std::max(group_unrel[S.tile_group[t]], best_unrel);
            } else {
                total_s += reach * best_unrel / rate;
                total_sectors += reach * best_unrel;
            }
            if (entry_only) {
                // what the tile cannot rule out goes to the node-at-a-time descent: two (read, child) pairs per surviving
                // terminal column, each a cheap sampled test there (~32 probes at the L2 rate)
                const double r_s = best_s ? 1.0 : 0.0;
                for (uint32_t c = 0; c < tm.n_cols; ++c) {
                    if (!((tm.terminal[c >> 5] >> (c & 31)) & 1u)) continue;
                    const uint32_t u = S.tile_nodes[t][c];
                    double surv = F.q[u];
                    if (r_s > 0.0) {
                        const double r = 1.0 - pow(std::min(F.fill[u], 0.999999), (double)best_s);
                        const double mean = n * r, var = n * r * (1.0 - r);
                        surv = var < 1e-9 ? (mean > F.allowed ? 0.0 : 1.0) : phi_tab((F.allowed + 0.5 - mean) / sqrt(var));
                        if (!tm.filter_only) surv = std::min(surv, F.q[u]);
                    }
                    total_s += reach * surv * 2.0 * 32.0 / 240e9;
                }
            }
        }
        for (double g : group_unrel) {
            total_s += g / sector_rate(line_bytes);  // a line costs what a sector costs (scripts/mb/mb_coop.cu)
            total_sectors += g;
        }
        S.est_sectors_per_read = total_sectors;
        S.est_seconds_per_read = total_s;
        // a read that belongs to a genome of the database: the entry tile's pre-test (plus its exact pass unless the tile
        // only filters), then one exactly evaluated tile per tile-tree level below
        {
            double worst_entry = 0.0;
            for (uint32_t t : S.entry_tiles) {
                const SlicedTileDev &tm = S.tiles[t];
                const double bytes = (double)(64ULL * db->wpf) * tm.row_stride * 4.0;
                const double share = S.tile_group[t] >= 0 && tm.filter_only && tm.pre_steps <= 2 ? 1.0 / SL_QUAD : 1.0;
                worst_entry += (share * n * (double)tm.pre_steps + (tm.filter_only ? 0.0 : n * (double)K)) / sector_rate(bytes);
            }
            S.est_seconds_related = worst_entry + (entry_only ? pair_related_s : (double)max_depth * n * (double)K / sector_rate(1e12));
        }
    }
}

// Chooses the cut: the skipped top is "every verified interior node with more than G leaves below it", for the G that
// makes an unrelated read cheapest under the cost model (deeper cuts have emptier filters, so reads die after fewer
// k-mers, but need more columns and hence more tiles).  G = infinity (nothing skipped) is always a candidate and the
// only one when the top of the tree holds an unverified node.
void plan_tiles(const pf_db *db, float threshold, uint64_t n_nominal, SlicedState &S) {
    TreeFacts F;
    tree_facts(db, threshold, n_nominal, F);
    // a related read in the node-at-a-time descent: one exact leaf plus two cheap sampled tests per level
    const double pair_related_s = ((double)n_nominal * (double)db->geom.num_hashes + 32.0 * (double)(db->level_start.size() - 1)) / 240e9;
    const size_t nn = db->n_nodes;
    std::vector<uint64_t> cand{~0ULL};
    for (double g = (double)std::max<uint64_t>(db->n_leaves, 1); g >= 1.0; g /= 1.4142135623730951) {
        const uint64_t G = (uint64_t)g;
        if (cand.back() != G) cand.push_back(G);
    }
    SlicedState best;
    bool have = false;
    std::vector<uint8_t> prev_skip;
    // second family of cuts: by fill instead of by leaf count ("skip while the filter is fuller than phi") -- the fullest
    // node of the cut decides how many k-mers an unrelated read needs, and subtrees of equal size differ in fill
    const size_t n_by_leaves = cand.size();
    for (int i = 19; i >= 2; --i) cand.push_back((uint64_t)i);  // phi = i / 20
    size_t ci0 = 0, ci1 = cand.size();
    if (const char *f = getenv("PF_SLICED_FORCE_G")) {  // tests: a given cut ("skip every verified node with more than G leaves")
        cand.assign(1, (uint64_t)strtoull(f, nullptr, 10));
        ci0 = 0, ci1 = 1;
    }
    for (size_t ci = ci0; ci < ci1; ++ci) {
        const uint64_t G = cand[ci];
        const bool by_fill = ci1 > 1 && ci >= n_by_leaves;
        const double phi = by_fill ? (double)G / 20.0 : 0.0;
        SlicedState T;
        T.skip.assign(nn, 0);
        for (size_t u = 0; u < nn; ++u) {
            if (db->h_leaf[u] >= 0 || !F.vb[u]) continue;
            if (by_fill ? F.fill[u] <= phi : (uint64_t)F.leaves[u] <= G) continue;
            if (u == 0 || T.skip[F.parent[u]]) T.skip[u] = 1;
        }
        if (have && T.skip == prev_skip) continue;
        prev_skip = T.skip;
        T.hybrid = S.hybrid;
        tile_tree(db, F, T, S.hybrid, pair_related_s);
        if (getenv("PF_SLICED_DEBUG"))
            fprintf(stderr, "[sliced plan] %s=%llu skipped=%zu entry_tiles=%zu tiles=%zu est=%.2f ns/read (%.0f sectors)\n",
                    by_fill ? "G=phi*20" : "G", (unsigned long long)G, (size_t)std::count(T.skip.begin(), T.skip.end(), 1), T.entry_tiles.size(),
                    T.tiles.size(), T.est_seconds_per_read * 1e9, T.est_sectors_per_read);
        if (!have || T.est_seconds_per_read < best.est_seconds_per_read) {
            best = std::move(T);
            have = true;
        }
    }
    S.skip = std::move(best.skip);
    S.tiles = std::move(best.tiles);
    S.col_slot = std::move(best.col_slot);
    S.child_tile = std::move(best.child_tile);
    S.child_mask = std::move(best.child_mask);
    S.entry_tiles = std::move(best.entry_tiles);
    S.tile_parent = std::move(best.tile_parent);
    S.tile_nodes = std::move(best.tile_nodes);
    S.tile_group = std::move(best.tile_group);
    S.n_line_tiles = best.n_line_tiles;
    S.table_words = best.table_words;
    S.entry_bytes = best.entry_bytes;
    S.est_sectors_per_read = best.est_sectors_per_read;
    S.est_seconds_per_read = best.est_seconds_per_read;
    S.est_seconds_related = best.est_seconds_related;
    if (getenv("PF_SLICED_DEBUG")) {
        fprintf(stderr, "[sliced plan] chosen: %zu tiles, %zu entry, tables %.2f GB, est %.2f ns/read\n", S.tiles.size(),
                S.entry_tiles.size(), S.table_words * 4.0 / 1e9, S.est_seconds_per_read * 1e9);
        for (uint32_t t : S.entry_tiles)
            fprintf(stderr, "[sliced plan]   entry tile %u: %u columns, row %u B, stride %u B, %u children, pre-test %u steps, %u rounds first%s\n", t,
                    S.tiles[t].n_cols, S.tiles[t].row_words * 4, S.tiles[t].row_stride * 4, S.tiles[t].n_children, S.tiles[t].pre_steps,
                    S.tiles[t].pre_rounds, S.tiles[t].filter_only ? " (filter only)" : "");
    }
}

// ---- tables ----------------------------------------------------------------------------------------------------------
static int build_tables(pf_db *db, SlicedState &S) {
    sliced_release_device(&S);
    const size_t nt = S.tiles.size();
    size_t free_b = 0, total_b = 0;
    PF_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t want = S.table_words * 4ULL;
    if (want + (8ULL << 30) > free_b) {  // leave room for the batch, its hash cache and the frontier
        set_error("sliced tables need %.1f GB, %.1f GB free", want / 1e9, free_b / 1e9);
        return PF_ERR_NOMEM;
    }
    PF_CUDA_OK(cudaMalloc(&S.d_tables, std::max<uint64_t>(want, 4)));
    PF_CUDA_OK(cudaMalloc(&S.d_tiles, nt * sizeof(SlicedTileDev)));
    PF_CUDA_OK(cudaMalloc(&S.d_child_tile, std::max<size_t>(S.child_tile.size(), 1) * 4));
    PF_CUDA_OK(cudaMalloc(&S.d_child_mask, std::max<size_t>(S.child_mask.size(), 1) * 4));
    PF_CUDA_OK(cudaMalloc(&S.d_entry, std::max<size_t>(S.entry_tiles.size(), 1) * 4));
    PF_CUDA_OK(cudaMalloc(&S.d_tile_count, 2 * nt * 4));
    S.d_tile_cursor = S.d_tile_count + nt;
    PF_CUDA_OK(cudaMalloc(&S.d_tile_base, nt * 8));
    PF_CUDA_OK(cudaMalloc(&S.d_counters, 5 * 8));
    S.d_hit_cursor = S.d_counters + 4;
    PF_CUDA_OK(cudaMalloc(&S.d_work, 8));
    {
        const size_t nn = db->n_nodes;
        PF_CUDA_OK(cudaMalloc(&S.d_node_inj_count, 2 * nn * 4));
        S.d_node_inj_cursor = S.d_node_inj_count + nn;
        PF_CUDA_OK(cudaMalloc(&S.d_node_inj_base, (nn + 1) * 8));
        PF_CUDA_OK(cudaMalloc(&S.d_inj_bsum, ((nn + 1023) / 1024 + 1) * 8));
        PF_CUDA_OK(cudaMallocHost(&S.h_node_inj_base, (nn + 1) * 8));
    }
    PF_CUDA_OK(cudaMallocHost(&S.h_tile_count, nt * 4));
    PF_CUDA_OK(cudaMallocHost(&S.h_counters, 5 * 8));
    PF_CUDA_OK(cudaMallocHost(&S.h_tile_base, nt * 8));
    cudaStream_t s = db->stream;
    uint32_t *d_col_slot = nullptr;
    PF_CUDA_OK(cudaMalloc(&d_col_slot, S.col_slot.size() * 4));
    PF_CUDA_OK(cudaMemcpyAsync(d_col_slot, S.col_slot.data(), S.col_slot.size() * 4, cudaMemcpyHostToDevice, s));
    PF_CUDA_OK(cudaMemcpyAsync(S.d_tiles, S.tiles.data(), nt * sizeof(SlicedTileDev), cudaMemcpyHostToDevice, s));
    if (!S.child_tile.empty()) {
        PF_CUDA_OK(cudaMemcpyAsync(S.d_child_tile, S.child_tile.data(), S.child_tile.size() * 4, cudaMemcpyHostToDevice, s));
        PF_CUDA_OK(cudaMemcpyAsync(S.d_child_mask, S.child_mask.data(), S.child_mask.size() * 4, cudaMemcpyHostToDevice, s));
    }
    PF_CUDA_OK(cudaMemcpyAsync(S.d_entry, S.entry_tiles.data(), S.entry_tiles.size() * 4, cudaMemcpyHostToDevice, s));
    const uint32_t gx = (uint32_t)(2 * db->wpf / 8);
    for (size_t t0 = 0; t0 < nt; t0 += 32768) {
        const uint32_t ny = (uint32_t)std::min<size_t>(32768, nt - t0);
        slice_kernel<<<dim3(gx, ny), 256, 0, s>>>(db->d_filters, db->wpf, d_col_slot, S.d_tiles, (uint32_t)t0, S.d_tables);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(d_col_slot);
    PF_CUDA_OK(e);
    PF_CUDA_OK(cudaGetLastError());
    S.tables_ready = true;
    db->stats.sliced_tiles = nt;
    db->stats.sliced_table_bytes = want;
    return PF_OK;
}

// Decides how a batch with this threshold and nominal read length is evaluated and, for the sliced path, makes sure the
// tiles exist.  Returns PF_OK with *use_sliced set.  Auto mode compares the two cost models: expected bit probes of the
// node-at-a-time plan (L2-resident filter, ~240 G probes/s measured) against expected sector loads of the tiles at the
// measured random-row rate for their footprint.
int sliced_prepare(pf_db *db, float threshold, uint64_t n_nominal, bool *use_sliced) {
    *use_sliced = false;
    if (db->mode == 1 || db->sharded || db->exhaustive || !db->lazy) return PF_OK;
    if (!db->sliced) db->sliced = new SlicedState();
    SlicedState &S = *db->sliced;
    if (S.failed && db->mode != 2) return PF_OK;
    // the share of reads that belong to the database decides which path is cheaper; it is learnt from the blocks
    // already answered (hits per read, capped at 1) and starts at 1/2
    const double rho = db->related_share < 0 ? 0.5 : db->related_share;
    if (S.decided_mode && S.theta == threshold && S.n_nominal == n_nominal && S.decided_under == db->mode &&
        S.decided_handover == db->handover && S.tile_cols == db->tile_cols &&
        (db->mode != 0 || fabs(rho - S.decided_rho) <= 0.1)) {
        *use_sliced = S.decided_mode == 2;
        if (*use_sliced) {
            db->stats.sliced_tiles = S.tiles.size();
            db->stats.sliced_table_bytes = S.table_words * 4ULL;
        }
        return PF_OK;
    }
    SlicedState P;  // plan into a scratch state first: the tables are rebuilt only if the tiling changed
    // Below the cut: tiles all the way down (default), or -- pf_db_set_handover / PF_SLICED_HANDOVER=1, and automatically
    // when the full set of tables does not fit the free HBM -- tiles for the cut only, whose survivors are handed to the
    // node-at-a-time descent.  Measured equal on cfg3 (47 ms either way: a read that belongs touches a cold 1.8 MB leaf
    // filter or a cold table, HBM-bound both ways) and slower on 10 kb reads (114 vs 79 ms), at 1 GB instead of 36 GB.
    P.hybrid = db->handover == 1 || S.deep_failed;
    plan_tiles(db, threshold, n_nominal, P);
    bool sliced = db->mode == 2;
    if (db->mode == 0) {
        // Node-at-a-time: the step plan's expected probes per k-mer for an unrelated read (L2-resident filters, ~240 G
        // probes/s measured); a related read walks root -> leaf, two cheap sampled tests per level and one exact leaf.
        const double n = (double)n_nominal, K = (double)db->geom.num_hashes;
        const double levels = (double)(db->level_start.size() - 1);
        const double p_unrel = db->plan_cost * n / 240e9, p_rel = (n * K + 32.0 * levels) / 240e9;
        const double t_pair = (1.0 - rho) * p_unrel + rho * p_rel;
        const double t_sliced = (1.0 - rho) * P.est_seconds_per_read + rho * P.est_seconds_related;
        // hysteresis: leave the current path only for a clear gain
        sliced = S.decided_mode == 2 ? !(t_pair < 0.8 * t_sliced) : t_sliced < 0.8 * t_pair;
        if (getenv("PF_SLICED_DEBUG"))
            fprintf(stderr, "[sliced plan] auto (related share %.2f): node-at-a-time %.1f ns/read (unrelated %.1f, related %.1f), "
                            "sliced %.1f ns/read (unrelated %.1f, related %.1f) -> %s\n",
                    rho, t_pair * 1e9, p_unrel * 1e9, p_rel * 1e9, t_sliced * 1e9, P.est_seconds_per_read * 1e9,
                    P.est_seconds_related * 1e9, sliced ? "sliced" : "node-at-a-time");
    }
    const bool cols_same = S.tile_cols == db->tile_cols;
    S.tile_cols = db->tile_cols;
    S.theta = threshold;
    S.n_nominal = n_nominal;
    S.decided_mode = sliced ? 2 : 1;
    S.decided_under = db->mode;
    S.decided_rho = rho;
    S.decided_handover = db->handover;
    if (!sliced) return PF_OK;
    const bool same = S.tables_ready && S.skip == P.skip && S.tiles.size() == P.tiles.size() && S.hybrid == P.hybrid &&
                      cols_same;
    if (same) {
        // same tiling, possibly other pre-test depths (they follow the threshold and the read length): refresh the tile records
        for (size_t t = 0; t < S.tiles.size(); ++t) {
            S.tiles[t].pre_steps = P.tiles[t].pre_steps;
            S.tiles[t].filter_only = P.tiles[t].filter_only;
            S.tiles[t].pre_rounds = P.tiles[t].pre_rounds;
        }
        PF_CUDA_OK(cudaMemcpyAsync(S.d_tiles, S.tiles.data(), S.tiles.size() * sizeof(SlicedTileDev), cudaMemcpyHostToDevice, db->stream));
        PF_CUDA_OK(cudaStreamSynchronize(db->stream));
        S.est_sectors_per_read = P.est_sectors_per_read;
        S.est_seconds_per_read = P.est_seconds_per_read;
        S.est_seconds_related = P.est_seconds_related;
    }
    if (!same) {
        S.skip = std::move(P.skip);
        S.tiles = std::move(P.tiles);
        S.col_slot = std::move(P.col_slot);
        S.child_tile = std::move(P.child_tile);
        S.child_mask = std::move(P.child_mask);
        S.entry_tiles = std::move(P.entry_tiles);
        S.tile_parent = std::move(P.tile_parent);
        S.tile_nodes = std::move(P.tile_nodes);
        S.tile_group = std::move(P.tile_group);
        S.n_line_tiles = P.n_line_tiles;
        S.hybrid = P.hybrid;
        S.table_words = P.table_words;
        S.entry_bytes = P.entry_bytes;
        S.est_sectors_per_read = P.est_sectors_per_read;
        S.est_seconds_per_read = P.est_seconds_per_read;
        S.est_seconds_related = P.est_seconds_related;
        int rc = build_tables(db, S);
        if (rc == PF_ERR_NOMEM && !S.hybrid && db->handover != 0) {
            // not enough HBM for tiles all the way down: keep the tiles of the cut only and hand their survivors over
            sliced_release_device(&S);
            S.deep_failed = true;
            S.decided_mode = 0;
            return sliced_prepare(db, threshold, n_nominal, use_sliced);
        }
        if (rc != PF_OK) {
            sliced_release_device(&S);
            S.failed = true;
            S.decided_mode = 1;
            if (db->mode == 2) return rc;  // explicitly requested: report
            return PF_OK;                  // auto: stay with the node-at-a-time path
        }
    }
    // Under this threshold: the leading line groups whose tiles all filter with a 1- or 2-step pre-test are taken by the line
    // kernel; the rest of the entry depth goes pair by pair (lean instantiation when every entry tile is of that kind).
    {
        auto groupable = [&](uint32_t t) {
            return S.tiles[t].filter_only && S.tiles[t].pre_steps >= 1 && S.tiles[t].pre_steps <= 2 &&
                   S.tiles[t].pre_steps < db->geom.num_hashes;
        };
        S.n_quad_now = 0;
        for (uint32_t e0 = 0; e0 < S.n_line_tiles; e0 += SL_QUAD) {
            const uint32_t e1 = std::min<uint32_t>(e0 + SL_QUAD, S.n_line_tiles);
            bool all = true;
            for (uint32_t e = e0; e < e1; ++e) all = all && groupable(S.entry_tiles[e]);
            if (!all) break;
            S.n_quad_now = e1;
        }
        size_t n_groupable = 0;
        for (uint32_t t : S.entry_tiles) n_groupable += groupable(t) ? 1u : 0u;
        S.entry_lean = !S.entry_tiles.empty() && n_groupable == S.entry_tiles.size();
    }
    db->stats.sliced_tiles = S.tiles.size();
    db->stats.sliced_table_bytes = S.table_words * 4ULL;
    *use_sliced = true;
    return PF_OK;
}

uint64_t sliced_entry_tiles(const pf_db *db) { return db->sliced ? db->sliced->entry_tiles.size() : 0; }
bool sliced_hybrid(const pf_db *db) { return db->sliced && db->sliced->hybrid; }

template <int PW>
static void launch_sliced(const SlicedArgs &a, int sm_count, bool lean, cudaStream_t s) {
    static const int lean_ctas = getenv("PF_SLICED_LEAN_CTAS") ? atoi(getenv("PF_SLICED_LEAN_CTAS")) : 3;
    if (lean && lean_ctas == 3) {
        if (a.hp.small_m) sliced_probe_kernel<PW, true, true, 3><<<sm_count * 3, SL_THREADS, 0, s>>>(a);
        else sliced_probe_kernel<PW, false, true, 3><<<sm_count * 3, SL_THREADS, 0, s>>>(a);
    } else if (lean) {
        if (a.hp.small_m) sliced_probe_kernel<PW, true, true, 2><<<sm_count * 2, SL_THREADS, 0, s>>>(a);
        else sliced_probe_kernel<PW, false, true, 2><<<sm_count * 2, SL_THREADS, 0, s>>>(a);
    } else {
        if (a.hp.small_m) sliced_probe_kernel<PW, true, false, 2><<<sm_count * 2, SL_THREADS, 0, s>>>(a);
        else sliced_probe_kernel<PW, false, false, 2><<<sm_count * 2, SL_THREADS, 0, s>>>(a);
    }
}

template <bool FUSE, int CTAS>
static void launch_quad_c(const SlicedArgs &a, uint64_t max_kmers, uint32_t n_entry, uint32_t n_groups, int sm_count, cudaStream_t s) {
    const int grid = sm_count * CTAS;
    if (max_kmers < 256) {
        if (a.hp.small_m) sliced_entry_quad_kernel<8, true, FUSE, CTAS><<<grid, SL_THREADS, 0, s>>>(a, n_entry, n_groups);
        else sliced_entry_quad_kernel<8, false, FUSE, CTAS><<<grid, SL_THREADS, 0, s>>>(a, n_entry, n_groups);
    } else if (max_kmers < 65536) {
        if (a.hp.small_m) sliced_entry_quad_kernel<16, true, FUSE, 3><<<sm_count * 3, SL_THREADS, 0, s>>>(a, n_entry, n_groups);
        else sliced_entry_quad_kernel<16, false, FUSE, 3><<<sm_count * 3, SL_THREADS, 0, s>>>(a, n_entry, n_groups);
    } else {
        if (a.hp.small_m) sliced_entry_quad_kernel<32, true, FUSE, 3><<<sm_count * 3, SL_THREADS, 0, s>>>(a, n_entry, n_groups);
        else sliced_entry_quad_kernel<32, false, FUSE, 3><<<sm_count * 3, SL_THREADS, 0, s>>>(a, n_entry, n_groups);
    }
}
// Short reads (8 count planes): 4 CTAs per SM.  The kernel is latency-bound (ncu on cfg3: 24 warps per SM at 3 CTAs, 7 of
// them waiting on the gathers per issue, DRAM pipe a third busy), so resident warps win over registers: 28.2 ms per cfg3
// step at 4 CTAs (64 registers, ~100 B spilled), 30.0 at 3, 35.6 at 2 (PF_QUAD_CTAS=3 for the A/B).  Working out the next
// round's row indices while the lines are on their way was tried and is slower (30.4 ms: more live registers).
template <bool FUSE>
static void launch_quad(const SlicedArgs &a, uint64_t max_kmers, uint32_t n_entry, uint32_t n_groups, int sm_count, cudaStream_t s) {
    static const int ctas = getenv("PF_QUAD_CTAS") ? atoi(getenv("PF_QUAD_CTAS")) : 4;
    if (ctas == 3) launch_quad_c<FUSE, 3>(a, max_kmers, n_entry, n_groups, sm_count, s);
    else launch_quad_c<FUSE, 4>(a, max_kmers, n_entry, n_groups, sm_count, s);
}

// The entry depth can make its hash values itself when every entry tile goes through the line kernel and k fits the 2-bit
// register path: nothing downstream needs the values of the reads that do not survive (with tiles below the cut,
// run_sliced has the survivors hashed before the next depth; with a hand-over to the node-at-a-time descent, hand_over
// does, together with their step-0 indices).
bool sliced_fused(const pf_db *db, const pf_dev_batch *bt) {
    static const bool off = getenv("PF_SLICED_NO_FUSE") != nullptr || getenv("PF_SLICED_NO_QUAD") != nullptr;
    if (off || !db->sliced) return false;
    const SlicedState &S = *db->sliced;
    (void)bt;
    return S.n_quad_now >= 2 && S.n_quad_now == S.entry_tiles.size() && db->hp.k >= 17 && db->hp.k <= 32;
}

// The tile-level descent for reads [r0, r0 + n_chunk) of the batch whose hash values are cached in db->hb.
int run_sliced(pf_db *db, const pf_dev_batch *bt, float threshold, int want_hits, uint64_t kmer_base, uint32_t r0,
               uint32_t n_chunk, Descent &st, const HashArgs *hf) {
    SlicedState &S = *db->sliced;
    cudaStream_t s = db->stream;
    const size_t nt = S.tiles.size();
    int rc;
    uint64_t n = (uint64_t)n_chunk * S.entry_tiles.size();
    if (n > db->frontier_cap) return PF_SPLIT_CHUNK;
    int cur = 0;
    bool entry = true;
    while (n > 0) {
        if ((rc = S.reach[cur].ensure(n * 8)) || (rc = S.alive.ensure(n))) return rc;
        PF_CUDA_OK(cudaMemsetAsync(S.d_counters, 0, 4 * 8, s));
        PF_CUDA_OK(cudaMemsetAsync(S.d_work, 0, 8, s));
        PF_CUDA_OK(cudaMemsetAsync(S.d_tile_count, 0, 2 * nt * 4, s));
        SlicedArgs a{};
        a.fr_read = entry ? nullptr : S.fr_read[cur].p;
        a.fr_tile = entry ? nullptr : S.fr_tile[cur].p;
        a.fr_src = entry ? nullptr : S.fr_src[cur].p;
        a.n_pairs = (uint32_t)n;
        a.entry_tiles = S.d_entry;
        a.read0 = r0;
        a.n_chunk = n_chunk;
        a.lengths = bt->lengths.p;
        a.kmer_off = bt->kmer_off.p;
        a.hb = db->hb.p;
        a.kmer_base = kmer_base;
        a.tiles = S.d_tiles;
        a.tables = S.d_tables;
        a.child_tile = S.d_child_tile;
        a.child_mask = S.d_child_mask;
        a.src_reach = S.reach[cur ^ 1].p;
        a.reach = S.reach[cur].p;
        a.alive = S.alive.p;
        a.tile_count = S.d_tile_count;
        a.counters = S.d_counters;
        a.work_ctr = S.d_work;
        a.hp = db->hp;
        a.threshold = threshold;
        a.grab = bt->max_kmers <= 256 ? 4u : 1u;
        if (S.hybrid) {
            PF_CUDA_OK(cudaMemsetAsync(S.d_node_inj_count, 0, 2 * db->n_nodes * 4, s));
            a.node_inj_count = S.d_node_inj_count;
        }
        while (db->ev_probe.size() < st.n_ev + 2) {
            cudaEvent_t e;
            PF_CUDA_OK(cudaEventCreate(&e));
            db->ev_probe.push_back(e);
        }
        PF_CUDA_OK(cudaEventRecord(db->ev_probe[st.n_ev], s));
        // Entry depth.  Groups of tiles that only filter with a 1- or 2-step pre-test (the usual plan) and share 128-byte
        // lines go through the line kernel, one pass per read and group (sliced_entry_quad_kernel); the rest of the depth --
        // or all of it -- goes pair by pair, with the lean instantiation when every tile is of that kind.
        const bool lean = entry && S.entry_lean && !getenv("PF_SLICED_NO_LEAN");
        uint32_t n_grouped = 0;
        if (entry && S.n_quad_now >= 2 && !getenv("PF_SLICED_NO_QUAD")) {
            n_grouped = S.n_quad_now;
            const uint32_t n_groups = (n_grouped + SL_QUAD - 1) / SL_QUAD;
            SlicedArgs g = a;
            g.grab = bt->max_kmers <= 256 ? 4u : 1u;
            if (hf) {
                if ((rc = db->read_flag.ensure(bt->n_reads))) return rc;
                PF_CUDA_OK(cudaMemsetAsync(db->read_flag.p + r0, 0, n_chunk, s));
                g.packed = hf->packed;
                g.word_off = hf->word_off;
                g.exc_index = hf->exc_index;
                g.surv_flags = db->read_flag.p;
                launch_quad<true>(g, bt->max_kmers, n_grouped, n_groups, db->sm_count, s);
            } else {
                launch_quad<false>(g, bt->max_kmers, n_grouped, n_groups, db->sm_count, s);
            }
            st.probe_launches++;
        }
        if (!entry || n_grouped < S.entry_tiles.size()) {
            if (n_grouped) {
                a.pair0 = n_grouped * n_chunk;
                a.n_pairs = (uint32_t)(n - (uint64_t)n_grouped * n_chunk);
                a.work_ctr = S.d_work + 1;
            }
            if (bt->max_kmers < 256) launch_sliced<8>(a, db->sm_count, lean, s);
            else if (bt->max_kmers < 65536) launch_sliced<16>(a, db->sm_count, lean, s);
            else launch_sliced<32>(a, db->sm_count, lean, s);
            st.probe_launches++;
        }
        PF_CUDA_OK(cudaEventRecord(db->ev_probe[st.n_ev + 1], s));
        st.n_ev += 2;
        st.ev_sliced.push_back(entry ? 2 : 1);  // 2: the entry depth (line kernel and / or entry pairs)
        st.pairs += n;
        st.sliced_pairs += n;
        st.levels++;
        PF_CUDA_OK(cudaMemcpyAsync(S.h_counters, S.d_counters, 4 * 8, cudaMemcpyDeviceToHost, s));
        PF_CUDA_OK(cudaMemcpyAsync(S.h_tile_count, S.d_tile_count, nt * 4, cudaMemcpyDeviceToHost, s));
        PF_CUDA_OK(cudaStreamSynchronize(s));
        st.sectors += S.h_counters[0];
        st.lines += S.h_counters[3];
        const uint64_t n_alive = S.h_counters[1], hits = S.h_counters[2];
        uint64_t next_n = 0;
        for (size_t t = 0; t < nt; ++t) {
            S.h_tile_base[t] = next_n;
            next_n += S.h_tile_count[t];
        }
        if (next_n > db->frontier_cap) return PF_SPLIT_CHUNK;
        const int nxt = cur ^ 1;
        uint64_t n_inj = 0;
        if (S.hybrid && n_alive) {
            // exclusive scan of the per-node hand-over counts: nodes are numbered level by level, so the offsets at the
            // level boundaries cut the (read, node) list into per-level, node-major slices
            const uint32_t nn = (uint32_t)db->n_nodes, nb = (nn + 1023u) / 1024u;
            csr_block_sums_kernel<<<nb, 1024, 0, s>>>(S.d_node_inj_count, nn, S.d_inj_bsum);
            csr_scan_sums_kernel<<<1, 1024, 0, s>>>(S.d_inj_bsum, nb);
            csr_offsets_kernel<<<nb, 1024, 0, s>>>(S.d_node_inj_count, nn, S.d_inj_bsum, S.d_node_inj_base);
            st.other_launches += 3;
            PF_CUDA_OK(cudaMemcpyAsync(S.h_node_inj_base, S.d_node_inj_base, ((size_t)nn + 1) * 8, cudaMemcpyDeviceToHost, s));
            PF_CUDA_OK(cudaStreamSynchronize(s));
            n_inj = S.h_node_inj_base[nn];
            if (n_inj > db->frontier_cap) return PF_SPLIT_CHUNK;
            if (n_inj && ((rc = S.inj_read.ensure(n_inj)) || (rc = S.inj_node.ensure(n_inj)))) return rc;
        }
        db->inj_level_off.assign(db->level_start.size(), 0);
        if (n_inj) {
            for (size_t l = 0; l < db->level_start.size(); ++l) db->inj_level_off[l] = S.h_node_inj_base[db->level_start[l]];
            db->inj_read = S.inj_read.p;
            db->inj_node = S.inj_node.p;
        }
        if (n_alive) {
            if (next_n && ((rc = S.fr_read[nxt].ensure(next_n)) || (rc = S.fr_tile[nxt].ensure(next_n)) ||
                           (rc = S.fr_src[nxt].ensure(next_n))))
                return rc;
            if (want_hits && hits &&
                ((rc = db->hit_read.grow_keep(st.hits_total + hits, st.hits_total, s)) ||
                 (rc = db->hit_leaf.grow_keep(st.hits_total + hits, st.hits_total, s))))
                return rc;
            PF_CUDA_OK(cudaMemcpyAsync(S.d_tile_base, S.h_tile_base, nt * 8, cudaMemcpyHostToDevice, s));
            SlicedEmitArgs e{};
            e.fr_read = a.fr_read;
            e.fr_tile = a.fr_tile;
            e.entry_tiles = S.d_entry;
            e.read0 = r0;
            e.n_chunk = n_chunk;
            e.alive = S.alive.p;
            e.n_alive = (uint32_t)n_alive;
            e.reach = S.reach[cur].p;
            e.tiles = S.d_tiles;
            e.child_tile = S.d_child_tile;
            e.child_mask = S.d_child_mask;
            e.tile_base = S.d_tile_base;
            e.tile_cursor = S.d_tile_cursor;
            e.nx_read = S.fr_read[nxt].p;
            e.nx_tile = S.fr_tile[nxt].p;
            e.nx_src = S.fr_src[nxt].p;
            e.hit_read = db->hit_read.p;
            e.hit_leaf = db->hit_leaf.p;
            e.read_hits = db->read_hits.p;
            e.hit_cursor = S.d_hit_cursor;
            e.blk_counts = db->d_blk_counts;
            e.want_hits = want_hits;
            if (n_inj) {
                e.node_inj_base = S.d_node_inj_base;
                e.node_inj_cursor = S.d_node_inj_cursor;
                e.inj_read = S.inj_read.p;
                e.inj_node = S.inj_node.p;
            }
            const uint32_t blocks = (uint32_t)std::min<uint64_t>((n_alive + 7) / 8, (uint64_t)db->sm_count * 8);
            sliced_emit_kernel<<<blocks, 256, 0, s>>>(e);
            st.other_launches++;
        }
        if (entry && hf && next_n) {
            // the entry depth made its hash values on the fly: the reads that go on get theirs cached now
            HashArgs h = *hf;
            h.flags = db->read_flag.p;
            PF_CUDA_OK(cudaMemsetAsync(h.work_ctr, 0, 4, s));
            launch_hash(h, db->sm_count * 8, s);
            st.other_launches++;
        }
        st.hits_total += hits;
        st.hits_before = st.hits_total;
        n = next_n;
        cur = nxt;
        entry = false;
    }
    return PF_OK;
}

// hit cursor of the emit kernel: hits of earlier chunks of the same block stay in front
int sliced_begin_block(pf_db *db) {
    SlicedState &S = *db->sliced;
    PF_CUDA_OK(cudaMemsetAsync(S.d_hit_cursor, 0, 8, db->stream));
    return PF_OK;
}
int sliced_set_hit_cursor(pf_db *db, uint64_t hits) {  // a chunk starts over (query_impl)
    SlicedState &S = *db->sliced;
    S.h_counters[4] = hits;
    PF_CUDA_OK(cudaMemcpyAsync(S.d_hit_cursor, S.h_counters + 4, 8, cudaMemcpyHostToDevice, db->stream));
    PF_CUDA_OK(cudaStreamSynchronize(db->stream));
    return PF_OK;
}

}  // namespace pf

// Host-only view of the tiling for tests (no device needed): the tree is given as level-ordered arrays.
static void host_db_from_arrays(pf_db &db, uint64_t n_nodes, const uint32_t *left, const uint32_t *right, const int32_t *leaf,
                                const uint64_t *pop, const uint8_t *mono, uint64_t num_bits, uint32_t num_hashes) {
    db.n_nodes = n_nodes;
    db.h_left.assign(left, left + n_nodes);
    db.h_right.assign(right, right + n_nodes);
    db.h_leaf.assign(leaf, leaf + n_nodes);
    db.h_pop.assign(pop, pop + n_nodes);
    db.h_mono.assign(mono, mono + n_nodes);
    db.h_slot.resize(n_nodes);
    db.n_leaves = 0;
    for (uint64_t u = 0; u < n_nodes; ++u) {
        db.h_slot[u] = (uint32_t)u;
        db.n_leaves += leaf[u] >= 0;
    }
    db.geom.num_bits = num_bits;
    db.geom.num_hashes = num_hashes;
    db.wpf = ((num_bits + 63) / 64 + 15) / 16 * 16;
    // level boundaries from the child links (nodes are level-ordered: children of level l form level l + 1)
    db.level_start.assign(1, 0);
    for (uint32_t lo = 0, hi = 1; lo < hi && hi <= n_nodes;) {
        uint32_t nh = hi;
        for (uint32_t u = lo; u < hi; ++u) nh += (left[u] != NONE32) + (right[u] != NONE32);
        db.level_start.push_back(hi);
        lo = hi;
        hi = nh;
    }
}

extern "C" int pf_plan_tiles(uint64_t n_nodes, const uint32_t *left, const uint32_t *right, const int32_t *leaf,
                             const uint64_t *pop, const uint8_t *mono, uint64_t num_bits, uint32_t num_hashes,
                             float threshold, uint64_t nominal_kmers, int handover, uint8_t *skip_out, int32_t *node_tile_out,
                             uint32_t *node_col_out, int32_t *tile_parent_out, uint32_t *tile_width_out, uint64_t tile_cap,
                             uint64_t *n_tiles_out, uint64_t *n_entry_out) {
    if (!left || !right || !leaf || !pop || !mono || !n_tiles_out || n_nodes == 0 || num_bits == 0) {
        pf::set_error("pf_plan_tiles: bad argument");
        return PF_ERR_ARG;
    }
    pf_db db;
    host_db_from_arrays(db, n_nodes, left, right, leaf, pop, mono, num_bits, num_hashes);
    pf::SlicedState S;
    S.hybrid = handover == 1;
    pf::plan_tiles(&db, threshold, nominal_kmers ? nominal_kmers : 1, S);
    if (skip_out) memcpy(skip_out, S.skip.data(), n_nodes);
    if (node_tile_out)
        for (uint64_t u = 0; u < n_nodes; ++u) node_tile_out[u] = -1;
    for (size_t t = 0; t < S.tiles.size(); ++t) {
        for (size_t c = 0; c < S.tile_nodes[t].size(); ++c) {
            if (node_tile_out) node_tile_out[S.tile_nodes[t][c]] = (int32_t)t;
            if (node_col_out) node_col_out[S.tile_nodes[t][c]] = (uint32_t)c;
        }
        if (t < tile_cap) {
            if (tile_parent_out) tile_parent_out[t] = S.tile_parent[t];
            if (tile_width_out) tile_width_out[t] = S.tiles[t].row_words * 32u;
        }
    }
    *n_tiles_out = S.tiles.size();
    if (n_entry_out) *n_entry_out = S.entry_tiles.size();
    return PF_OK;
}

// The same plan with narrower tiles if asked, and where its tables lie: per tile the first u32 word of its table, the
// words from one row to the next, the words of a row, the group of entry tiles it shares 128-byte lines with (-1: a table
// of its own), its pre-test depth and whether the pre-test is all it does; the entry tiles in the order the kernels walk
// them (the first n_line_tiles share lines, SL_QUAD per group); the words all tables take.
extern "C" int pf_plan_tile_layout(uint64_t n_nodes, const uint32_t *left, const uint32_t *right, const int32_t *leaf,
                                   const uint64_t *pop, const uint8_t *mono, uint64_t num_bits, uint32_t num_hashes,
                                   float threshold, uint64_t nominal_kmers, int handover, int tile_cols, uint64_t tile_cap,
                                   uint64_t *table_off_out, uint32_t *row_stride_out, uint32_t *row_words_out,
                                   int32_t *tile_group_out, uint32_t *pre_steps_out, uint8_t *filter_only_out,
                                   uint32_t *entry_order_out, uint64_t *n_tiles_out, uint64_t *n_entry_out,
                                   uint64_t *n_line_tiles_out, uint64_t *table_words_out, uint64_t *rows_out) {
    if (!left || !right || !leaf || !pop || !mono || !n_tiles_out || n_nodes == 0 || num_bits == 0 ||
        (tile_cols != 32 && tile_cols != 64 && tile_cols != 128 && tile_cols != 256)) {
        pf::set_error("pf_plan_tile_layout: bad argument");
        return PF_ERR_ARG;
    }
    pf_db db;
    host_db_from_arrays(db, n_nodes, left, right, leaf, pop, mono, num_bits, num_hashes);
    db.tile_cols = (uint32_t)tile_cols;
    pf::SlicedState S;
    S.hybrid = handover == 1;
    pf::plan_tiles(&db, threshold, nominal_kmers ? nominal_kmers : 1, S);
    for (size_t t = 0; t < S.tiles.size() && t < tile_cap; ++t) {
        if (table_off_out) table_off_out[t] = S.tiles[t].table_off;
        if (row_stride_out) row_stride_out[t] = S.tiles[t].row_stride;
        if (row_words_out) row_words_out[t] = S.tiles[t].row_words;
        if (tile_group_out) tile_group_out[t] = S.tile_group[t];
        if (pre_steps_out) pre_steps_out[t] = S.tiles[t].pre_steps;
        if (filter_only_out) filter_only_out[t] = (uint8_t)S.tiles[t].filter_only;
    }
    for (size_t e = 0; e < S.entry_tiles.size() && e < tile_cap; ++e)
        if (entry_order_out) entry_order_out[e] = S.entry_tiles[e];
    *n_tiles_out = S.tiles.size();
    if (n_entry_out) *n_entry_out = S.entry_tiles.size();
    if (n_line_tiles_out) *n_line_tiles_out = S.n_line_tiles;
    if (table_words_out) *table_words_out = S.table_words;
    if (rows_out) *rows_out = 64ULL * db.wpf;
    return PF_OK;
}
