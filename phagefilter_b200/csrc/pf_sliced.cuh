// pf_sliced.cuh -- bit-sliced evaluation of the gSBT (sm_100a): one 32-byte sector answers one bloom probe for
// up to 256 tree nodes at once.
//
// The node-at-a-time path (pf_kernels.cuh) reads one BIT per sector it touches.  Here the filters of a group of
// up to 256 nodes (a "tile") are stored transposed: row i of the tile's table holds bit i of every node's filter
// (BitVec<usize, Lsb0>, bloom_filter.rs:84-93), one column per node, row width 4..32 bytes.  For a k-mer,
// BloomFilter::contains (bloom_filter.rs:312-332) at ALL nodes of the tile is then the AND of the K rows
// g_i mod m (hash_iter.rs:13-27): K sector loads instead of K per node.  query_passes (query.rs:38-49) for all
// columns is a bit-sliced count of those masks over the read's k-mers compared with ceil(theta * n_k), and
// _query_batch's descent (query.rs:99-158: a node is reached iff every ancestor passed) is applied to the
// resulting pass bits inside the tile and across tiles.  A node that is evaluated is tested exactly, with its own
// filter, so the result is the reference's for any tree (including the non-superset filters its u16 name collisions
// produce).  Two shortcuts leave nodes out, both only where the load-time analysis VERIFIED that every (node, child)
// pair below is a bitwise superset, i.e. "a leaf below passes => this node passes" (pf_sliced.cu: the skipped top of the
// tree, and filter-only tiles whose sound pre-test is all they do).
//
// Rows that are lines: the tiles every read has to meet (the entry depth) lie four to a 128-byte line, because past the
// TLB reach a random access costs this part the same whether it brings a sector or a line (scripts/mb/mb_coop.cu); the
// kernel for that depth (sliced_entry_quad_kernel) also makes the k-mer hashes on the fly.  See DESIGN.md 4.2 and 5.
#pragma once
#include "pf_kernels.cuh"

namespace pf {

constexpr int SL_MAX_COLS = 256;
constexpr int SL_THREADS = 256;
constexpr int SL_STEP_BATCH = 5;  // row loads a lane keeps in flight per k-mer

struct SlicedTileDev {
    uint64_t table_off;   // first u32 word of the tile's table
    uint32_t n_cols;
    uint32_t row_words;   // 1, 2, 4 or 8 u32 words per row (32..256 columns)
    uint32_t prop_iters;  // depth of the deepest column below the tile's roots
    uint32_t first_child, n_children;  // links to the tiles whose roots hang below this tile's columns
    uint32_t entry;       // 1: every root is reached unconditionally (nothing evaluated above it)
    uint32_t pre_steps;   // probe steps of the sound pre-test pass (0 = none), chosen by the cost model
    uint32_t filter_only; // 1: the pre-test is all this tile does (its columns are implied by what passes below them)
    uint32_t pre_rounds;  // rounds of 32 k-mers the shallow pre-test keeps in flight between two looks at the columns
    uint32_t row_stride;  // u32 words from one row of the table to the next: row_words, or 32 when the tile shares 128-byte
                          // lines with up to three other entry tiles (its slot: 8 words at table_off)
    uint32_t valid[8];    // columns in use
    uint32_t terminal[8]; // columns that are tree leaves or have a child in another tile
    uint32_t leafmask[8]; // columns that are tree leaves
    uint16_t parent[SL_MAX_COLS];  // in-tile parent column; roots: 0x8000 | column in the parent tile (entry: 0xFFFF)
    int32_t leaf[SL_MAX_COLS];     // left-first DFS leaf index (query.rs:197-218) or -1
    uint32_t child_node[SL_MAX_COLS][2];  // tree nodes below a column that are NOT in this tile (node-at-a-time hand-over), or NONE
};

struct SlicedArgs {
    // frontier of (read, tile) pairs; entry depth (fr_read == null): pair i = (read0 + i % n_chunk, entry_tiles[i / n_chunk])
    const uint32_t *fr_read, *fr_tile, *fr_src;
    uint32_t n_pairs;
    uint32_t pair0;              // entry depth: index of the first pair this launch works on (the depth may be split in two launches)
    const uint32_t *entry_tiles;
    uint32_t read0, n_chunk;
    // reads
    const uint32_t *lengths;
    const uint64_t *kmer_off;
    const uint64_t *hb;
    uint64_t kmer_base;
    // tiles
    const SlicedTileDev *tiles;
    const uint32_t *tables;
    const uint32_t *child_tile;  // [links]
    const uint32_t *child_mask;  // [links][8] columns of the parent tile above the child tile's roots
    const uint32_t *src_reach;   // [pairs of the previous depth][8]
    // outputs
    uint32_t *reach;             // [n_pairs][8] columns reached AND passed (written for pairs listed in `alive`)
    uint32_t *alive;             // indices of the pairs with a hit or a successor, any order
    uint32_t *tile_count;        // per tile: pairs the next depth will hold
    uint32_t *node_inj_count;    // hand-over to the node-at-a-time descent: per tree node, (read, node) pairs to inject; or null
    unsigned long long *counters;  // [0] sectors loaded, [1] alive pairs, [2] leaf hits, [3] 128-byte lines loaded (entry groups)
    unsigned int *work_ctr;
    HashParams hp;
    float threshold;
    uint32_t grab;
    // entry line kernel hashing on the fly (FUSE): the 2-bit reads themselves; hash values are then cached only for the
    // reads that survive the entry depth (flagged here, hashed by hash_kernel restricted to the flags)
    const uint32_t *packed;
    const uint64_t *word_off;
    const uint32_t *exc_index;   // reads with bytes other than ACGT keep the cached path (hashed beforehand); may be null
    uint8_t *surv_flags;         // [batch reads] set for every read with a recorded pair; null when not fused
};

template <int RW>
PF_D void sl_load_row(const uint32_t *p, uint32_t (&w)[RW]) {
    if constexpr (RW == 8) {
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                     : "l"(p));
    } else if constexpr (RW == 4) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
        w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
    } else if constexpr (RW == 2) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
        w[0] = v.x, w[1] = v.y;
    } else {
        w[0] = __ldg(p);
    }
}
// streaming 8-byte load: the cached hash values are read once per (read, tile) pair and must not displace table rows
PF_D uint64_t sl_ld_stream(const uint64_t *p) {
    return __ldcs(reinterpret_cast<const unsigned long long *>(p));  // ld.global.cs: evict-first in L1 and L2
}
PF_D uint32_t sl_xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
PF_D uint32_t sl_maj(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (c & (a | b)); }

template <int RW>
struct SlLog2 { static constexpr int v = RW == 8 ? 3 : (RW == 4 ? 2 : (RW == 2 ? 1 : 0)); };
// After the reduce-scatter below lane l owns the row word with this index ...
template <int RW>
PF_D uint32_t sl_word_of_lane(uint32_t lane) {
    uint32_t j = 0;
#pragma unroll
    for (int b = 0; b < SlLog2<RW>::v; ++b) j |= ((lane >> b) & 1u) << (SlLog2<RW>::v - 1 - b);
    return j;
}
// ... and this is the lowest lane that owns word j.
template <int RW>
PF_D uint32_t sl_lane_of_word(uint32_t j) {
    uint32_t l = 0;
#pragma unroll
    for (int b = 0; b < SlLog2<RW>::v; ++b) l |= ((j >> (SlLog2<RW>::v - 1 - b)) & 1u) << b;
    return l;
}

// Per column, the number of lanes whose mask has the column's bit set, as a 6-plane bit-sliced number for the row word
// this lane owns (sl_word_of_lane): a reduce-scatter over the low lane bits (each stage halves the words a lane keeps
// and adds the partner's planes with full adders), then an all-reduce over the remaining lane bits.
template <int RW>
PF_D void sl_count_columns(const uint32_t (&m)[RW], uint32_t lane, uint32_t (&cnt)[6]) {
    constexpr int L = SlLog2<RW>::v;
    uint32_t a[6][RW];
#pragma unroll
    for (int w = 0; w < RW; ++w) a[0][w] = m[w];
#pragma unroll
    for (int b = 0; b < L; ++b) {
        const int P = 1 + b, H = RW >> (b + 1);
        const bool hi = (lane >> b) & 1u;
#pragma unroll
        for (int w = 0; w < H; ++w) {
            uint32_t carry = 0;
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
                const uint32_t lo_v = a[pl][w], hi_v = a[pl][w + H];
                const uint32_t keep = hi ? hi_v : lo_v, send = hi ? lo_v : hi_v;
                const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, send, 1 << b);
                a[pl][w] = sl_xor3(keep, recv, carry);
                carry = sl_maj(keep, recv, carry);
            }
            a[P][w] = carry;
        }
    }
#pragma unroll
    for (int b = L; b < 5; ++b) {
        const int P = 1 + b;
        uint32_t carry = 0;
#pragma unroll
        for (int pl = 0; pl < P; ++pl) {
            const uint32_t keep = a[pl][0];
            const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, keep, 1 << b);
            a[pl][0] = sl_xor3(keep, recv, carry);
            carry = sl_maj(keep, recv, carry);
        }
        a[P][0] = carry;
    }
#pragma unroll
    for (int pl = 0; pl < 6; ++pl) cnt[pl] = a[pl][0];
}

// columns whose bit-sliced counter is >= c
template <int PW>
PF_D uint32_t sl_ge(const uint32_t (&acc)[PW], uint32_t c) {
    if (PW < 32 && (c >> PW) != 0u) return 0u;
    uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int pl = PW - 1; pl >= 0; --pl) {
        const uint32_t cb = ((c >> pl) & 1u) ? 0xFFFFFFFFu : 0u;
        gt |= eq & acc[pl] & ~cb;
        eq &= ~(acc[pl] ^ cb);
    }
    return gt | eq;
}

// One pass over the k-mers of a read against one tile, `steps` probe steps per k-mer (steps = K: exact;
// fewer: a k-mer with no clear bit among its first `steps` rows only counts as POSSIBLY present, so the column counts are
// upper bounds).  In: alive_mine / live = columns not yet proven unable to pass (this lane's word / all words).
// Out: the same, narrowed; acc = per-column counts of this lane's word.  Returns false as soon as no terminal column can
// pass any more (read-level early exit; the rest of the read could not change that).
template <int RW, int PW, bool SMALL_M>
PF_D bool sl_scan(const HashParams &hp, const uint32_t *__restrict__ table, uint32_t stride, const uint64_t *__restrict__ hbp,
                  uint32_t n_k, uint32_t need, uint32_t steps, uint32_t lane, const uint32_t (&term)[RW], uint32_t term_mine,
                  uint32_t &alive_mine, uint32_t (&live)[RW], uint32_t (&acc)[PW], uint32_t &sectors) {
    const bool allowed0 = need == n_k;
    const uint32_t M0 = (uint32_t)hp.M, M1 = (uint32_t)(hp.M >> 32), m32 = (uint32_t)hp.m;
#pragma unroll
    for (int pl = 0; pl < PW; ++pl) acc[pl] = 0u;
    for (uint32_t base = 0; base < n_k; base += 32u) {
        const bool have = base + lane < n_k;
        const uint64_t hbv = have ? sl_ld_stream(hbp + base + lane) : 0ULL;
        const uint64_t h1 = fx_finish(hp.c1, hbv, hp.rot), h2 = fx_finish(hp.c2, hbv, hp.rot);
        uint32_t m[RW];
#pragma unroll
        for (int w = 0; w < RW; ++w) m[w] = 0xFFFFFFFFu;  // lanes without a k-mer stay neutral for the AND below
        bool on = have;
        uint64_t g = h1;  // g_0 = h1, g_1 = h2, g_i = (h1 + i) * h2 = g_{i-1} + h2 from g_2 on (hash_iter.rs:17-24)
        for (uint32_t s = 0; s < steps; s += SL_STEP_BATCH) {
            uint32_t rows[SL_STEP_BATCH][RW];
#pragma unroll
            for (int b = 0; b < SL_STEP_BATCH; ++b) {
#pragma unroll
                for (int w = 0; w < RW; ++w) rows[b][w] = 0xFFFFFFFFu;
                if (s + b < steps) {
                    if (on) {
                        uint64_t idx;
                        if (SMALL_M) idx = mod_small(g, M0, M1, m32);
                        else idx = mod_any(g, hp.m, hp.M);
                        sl_load_row<RW>(table + idx * stride, rows[b]);
                        ++sectors;
                    }
                    const uint32_t i = s + b;
                    g = i == 0u ? h2 : (i == 1u ? (h1 + 2ULL) * h2 : g + h2);
                }
            }
            uint32_t any_live = 0u;
#pragma unroll
            for (int w = 0; w < RW; ++w) {
#pragma unroll
                for (int b = 0; b < SL_STEP_BATCH; ++b) m[w] &= rows[b][w];
                any_live |= m[w] & live[w];
            }
            // BloomFilter::contains stops at the first clear bit; here a k-mer stops once it is absent from every
            // column that can still matter
            on = on && any_live != 0u;
            if (allowed0) {
                // theta = 1.0: one absent k-mer settles a column, so the columns alive are the AND over the lanes
                uint32_t t = 0u;
#pragma unroll
                for (int w = 0; w < RW; ++w) {
                    live[w] &= __reduce_and_sync(0xFFFFFFFFu, m[w]);
                    t |= live[w] & term[w];
                }
                if (t == 0u) return false;
            }
            if (!__any_sync(0xFFFFFFFFu, on)) break;
        }
        if (!have) {
#pragma unroll
            for (int w = 0; w < RW; ++w) m[w] = 0u;
        }
        uint32_t cnt[6];
        sl_count_columns<RW>(m, lane, cnt);
        {  // acc += cnt
            uint32_t carry = 0u;
#pragma unroll
            for (int pl = 0; pl < PW; ++pl) {
                const uint32_t x = pl < 6 ? cnt[pl] : 0u;
                const uint32_t sum = sl_xor3(acc[pl], x, carry);
                carry = sl_maj(acc[pl], x, carry);
                acc[pl] = sum;
            }
        }
        // a column cannot pass any more once its count plus all remaining k-mers stays below `need`
        const uint32_t done = min(base + 32u, n_k), rest = n_k - done;
        if (need > rest) {
            alive_mine &= sl_ge<PW>(acc, need - rest);
            if (!__any_sync(0xFFFFFFFFu, (alive_mine & term_mine) != 0u)) return false;
            if (rest) {
#pragma unroll
                for (int w = 0; w < RW; ++w) live[w] &= __shfl_sync(0xFFFFFFFFu, alive_mine, sl_lane_of_word<RW>(w));
            }
        }
    }
    if (allowed0) {  // columns settled by the AND inside the rounds
        uint32_t mine = 0u;
#pragma unroll
        for (int w = 0; w < RW; ++w)
            if (sl_word_of_lane<RW>(lane) == (uint32_t)w) mine = live[w];
        alive_mine &= mine;
    }
    return true;
}

// The same pass for a shallow pre-test (ST = 1 or 2 probe steps): RB rounds of 32 k-mers are in flight together, so a
// lane still has RB * ST independent row loads outstanding; the columns are re-examined every RB rounds.
template <int RW, int PW, bool SMALL_M, int RB, int ST>
PF_D bool sl_scan_shallow(const HashParams &hp, const uint32_t *__restrict__ table, uint32_t stride,
                          const uint64_t *__restrict__ hbp, uint32_t n_k, uint32_t need, uint32_t rounds, uint32_t lane,
                          const uint32_t (&term)[RW],
                          uint32_t term_mine, uint32_t &alive_mine, uint32_t (&live)[RW], uint32_t (&acc)[PW],
                          uint32_t &sectors) {
    // the plan expects a typical unrelated read to be settled after `rounds` rounds: that many are in flight first, then
    // the columns are re-examined after every further round (the gather is throughput-bound, not latency-bound)
    rounds = min(max(rounds, 1u), (uint32_t)RB);
    const bool allowed0 = need == n_k;
    const uint32_t M0 = (uint32_t)hp.M, M1 = (uint32_t)(hp.M >> 32), m32 = (uint32_t)hp.m;
#pragma unroll
    for (int pl = 0; pl < PW; ++pl) acc[pl] = 0u;
    uint64_t hb_pre = 0ULL;  // hash value of the NEXT round, fetched while this batch's rows are in flight
    bool pre_valid = false;
    for (uint32_t base = 0; base < n_k; base += 32u * rounds, rounds = 1u) {
        uint64_t hbv[RB];
        bool have[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            have[j] = (uint32_t)j < rounds && base + 32u * j + lane < n_k;
            hbv[j] = (j == 0 && pre_valid) ? hb_pre : (have[j] ? sl_ld_stream(hbp + base + 32u * j + lane) : 0ULL);
        }
        uint32_t rows[RB][ST][RW];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const uint64_t h1 = fx_finish(hp.c1, hbv[j], hp.rot), h2 = fx_finish(hp.c2, hbv[j], hp.rot);
#pragma unroll
            for (int st = 0; st < ST; ++st) {
#pragma unroll
                for (int w = 0; w < RW; ++w) rows[j][st][w] = 0xFFFFFFFFu;
                if (have[j]) {
                    const uint64_t g = st == 0 ? h1 : h2;  // g_0 = h1, g_1 = h2 (hash_iter.rs:17-24)
                    uint64_t idx;
                    if (SMALL_M) idx = mod_small(g, M0, M1, m32);
                    else idx = mod_any(g, hp.m, hp.M);
                    sl_load_row<RW>(table + idx * stride, rows[j][st]);
                    ++sectors;
                }
            }
        }
        {
            const uint32_t nb = base + 32u * rounds;  // the batches after the first hold one round
            hb_pre = nb + lane < n_k ? sl_ld_stream(hbp + nb + lane) : 0ULL;
            pre_valid = true;
        }
        if (allowed0) {
            uint32_t t = 0u;
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                uint32_t all = 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < RB; ++j)
#pragma unroll
                    for (int st = 0; st < ST; ++st) all &= rows[j][st][w];
                live[w] &= __reduce_and_sync(0xFFFFFFFFu, all);
                t |= live[w] & term[w];
            }
            if (t == 0u) return false;
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            if ((uint32_t)j >= rounds || base + 32u * j >= n_k) break;  // warp-uniform
            uint32_t m[RW];
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                m[w] = have[j] ? 0xFFFFFFFFu : 0u;
#pragma unroll
                for (int st = 0; st < ST; ++st) m[w] &= rows[j][st][w];
            }
            uint32_t cnt[6];
            sl_count_columns<RW>(m, lane, cnt);
            uint32_t carry = 0u;
#pragma unroll
            for (int pl = 0; pl < PW; ++pl) {
                const uint32_t x = pl < 6 ? cnt[pl] : 0u;
                const uint32_t sum = sl_xor3(acc[pl], x, carry);
                carry = sl_maj(acc[pl], x, carry);
                acc[pl] = sum;
            }
        }
        const uint32_t done = min(base + 32u * rounds, n_k), rest = n_k - done;
        if (need > rest) {
            alive_mine &= sl_ge<PW>(acc, need - rest);
            if (!__any_sync(0xFFFFFFFFu, (alive_mine & term_mine) != 0u)) return false;
        }
    }
#pragma unroll
    for (int w = 0; w < RW; ++w) live[w] &= __shfl_sync(0xFFFFFFFFu, alive_mine, sl_lane_of_word<RW>(w));
    if (allowed0) {
        uint32_t mine = 0u;
#pragma unroll
        for (int w = 0; w < RW; ++w)
            if (sl_word_of_lane<RW>(lane) == (uint32_t)w) mine = live[w];
        alive_mine &= mine;
    }
    return true;
}

// From the pass bits of one (read, tile) pair (this lane's row word) to the columns REACHED and passed, replicated in
// every lane.  _query_batch (query.rs:99-158): a column is reached iff its parent was reached and passed.  Roots take
// that from the pair one tile up (or unconditionally in an entry tile); inside the tile the bits are relaxed prop_iters
// times.  Returns false when no terminal column is left (nothing to emit).
template <int RW>
PF_D bool sl_reach(const SlicedArgs &a, const SlicedTileDev *__restrict__ tm, uint32_t src, uint32_t lane, uint32_t *s_bits,
                   uint32_t pass_mine, uint32_t (&reach_out)[RW]) {
    uint32_t term[RW];
#pragma unroll
    for (int w = 0; w < RW; ++w) term[w] = tm->terminal[w];
    uint32_t pass[RW], t = 0u;
#pragma unroll
    for (int w = 0; w < RW; ++w) {
        pass[w] = __shfl_sync(0xFFFFFFFFu, pass_mine, sl_lane_of_word<RW>(w));
        t |= pass[w] & term[w];
    }
    if (t == 0u) return false;
    // _query_batch (query.rs:99-158): a column is reached iff its parent was reached and passed.  Roots take that from
    // the pair one tile up (or unconditionally in an entry tile); inside the tile the bits are relaxed prop_iters times.
    uint32_t par[RW];
    uint32_t reach[RW];
#pragma unroll
    for (int w = 0; w < RW; ++w) {
        const uint32_t c = 32u * w + lane;
        par[w] = tm->parent[c];
        bool bit = (pass[w] >> lane) & 1u;
        if (bit && (par[w] & 0x8000u)) {
            if (par[w] != 0xFFFFu) {
                const uint32_t pc = par[w] & 0x7FFFu;
                bit = (ldg32(a.src_reach + (size_t)src * 8u + (pc >> 5)) >> (pc & 31u)) & 1u;
            }
        } else if (bit) {
            bit = false;  // decided by the relaxation below
        }
        reach[w] = __ballot_sync(0xFFFFFFFFu, bit);
    }
    const uint32_t iters = tm->prop_iters;
    for (uint32_t it = 0; it < iters; ++it) {
        __syncwarp();
        if (lane < RW) {
#pragma unroll
            for (int w = 0; w < RW; ++w)
                if (lane == (uint32_t)w) s_bits[w] = reach[w];
        }
        __syncwarp();
        uint32_t changed = 0u;
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            bool bit = (reach[w] >> lane) & 1u;
            if (!bit && ((pass[w] >> lane) & 1u) && !(par[w] & 0x8000u))
                bit = (s_bits[par[w] >> 5] >> (par[w] & 31u)) & 1u;
            const uint32_t nw = __ballot_sync(0xFFFFFFFFu, bit);
            changed |= nw ^ reach[w];
            reach[w] = nw;
        }
        if (!changed) break;
    }
    t = 0u;
#pragma unroll
    for (int w = 0; w < RW; ++w) {
        reach_out[w] = reach[w];
        t |= reach[w] & term[w];
    }
    return t != 0u;
}

// One (read, tile) pair.  Returns true when the pair has an output (a leaf hit or a successor tile); the columns
// reached and passed are then in reach_out (replicated in every lane).  s_bits: 8 words of shared memory of this warp.
// LEAN: the instantiation for depths whose tiles are all filter-only with a shallow (1- or 2-step) pre-test -- the exact
// pass and the deep scan are compiled out, which halves the registers and doubles the warps resident per SM.
template <int RW, int PW, bool SMALL_M, bool LEAN>
PF_D bool sl_pair(const SlicedArgs &a, const SlicedTileDev *__restrict__ tm, uint32_t r, uint32_t src, uint32_t lane,
                  uint32_t *s_bits, uint32_t (&reach_out)[RW], uint32_t &sectors) {
    const HashParams &hp = a.hp;
    const uint32_t n_k = kmers_of(ldg32(a.lengths + r), hp.k);
    const uint32_t need = need_of(a.threshold, n_k);
    const uint32_t *__restrict__ table = a.tables + tm->table_off;
    const uint32_t stride = tm->row_stride;
    const uint64_t *__restrict__ hbp = a.hb + (__ldg(a.kmer_off + r) - a.kmer_base);
    const uint32_t my_word = sl_word_of_lane<RW>(lane);
    const uint32_t valid_mine = tm->valid[my_word], term_mine = tm->terminal[my_word];
    uint32_t live[RW], term[RW];
#pragma unroll
    for (int w = 0; w < RW; ++w) {
        live[w] = tm->valid[w];
        term[w] = tm->terminal[w];
    }
    uint32_t alive_mine = valid_mine;
    uint32_t acc[PW];
#pragma unroll
    for (int pl = 0; pl < PW; ++pl) acc[pl] = 0u;
    if (need > n_k) return false;  // theta > 1: nothing can pass
    if (n_k != 0u && need != 0u) {  // need == 0 (theta = 0, or no k-mers at all): every column passes (query.rs:48)
        // Sound pre-test with the first few probe steps only: most reads that do not belong below this tile are
        // settled here at a fraction of the row loads; the others are then evaluated exactly, from scratch, with the
        // columns the pre-test has already ruled out switched off.
        const uint32_t pre = tm->pre_steps;
        if (pre != 0u && pre < hp.K) {
            bool ok;
            if (pre == 1u)
                ok = sl_scan_shallow<RW, PW, SMALL_M, 4, 1>(hp, table, stride, hbp, n_k, need, tm->pre_rounds, lane, term, term_mine, alive_mine, live,
                                                            acc, sectors);
            else if (pre == 2u)
                ok = sl_scan_shallow<RW, PW, SMALL_M, 2, 2>(hp, table, stride, hbp, n_k, need, tm->pre_rounds, lane, term, term_mine, alive_mine, live,
                                                            acc, sectors);
            else if constexpr (!LEAN)
                ok = sl_scan<RW, PW, SMALL_M>(hp, table, stride, hbp, n_k, need, pre, lane, term, term_mine, alive_mine, live, acc,
                                              sectors);
            else
                ok = true;  // not reached: the host launches the lean kernel only where pre <= 2
            if (!ok) return false;
        }
        if (LEAN || (pre != 0u && pre < hp.K && tm->filter_only)) {
            // columns the pre-test could not rule out count as passed: the exact tiles below decide (see plan)
#pragma unroll
            for (int pl = 0; pl < PW; ++pl) acc[pl] = 0xFFFFFFFFu;
        } else if constexpr (!LEAN) {
            if (!sl_scan<RW, PW, SMALL_M>(hp, table, stride, hbp, n_k, need, hp.K, lane, term, term_mine, alive_mine, live, acc, sectors))
                return false;
        }
    }
    // query_passes (query.rs:48): hits >= ceil(theta * n_k).  Columns ruled out earlier may hold stale counts (their
    // k-mers stop being probed), hence the mask.
    const uint32_t pass_mine = sl_ge<PW>(acc, need) & alive_mine;
    return sl_reach<RW>(a, tm, src, lane, s_bits, pass_mine, reach_out);
}

template <int RW>
PF_D void sl_record(const SlicedArgs &a, const SlicedTileDev *__restrict__ tm, uint32_t pair, uint32_t lane,
                    const uint32_t (&reach)[RW]);

template <int RW, int PW, bool SMALL_M, bool LEAN>
PF_D void sl_pair_and_record(const SlicedArgs &a, const SlicedTileDev *__restrict__ tm, uint32_t pair, uint32_t r,
                             uint32_t src, uint32_t lane, uint32_t *s_bits, uint32_t &sectors) {
    uint32_t reach[RW];
    if (!sl_pair<RW, PW, SMALL_M, LEAN>(a, tm, r, src, lane, s_bits, reach, sectors)) return;
    sl_record<RW>(a, tm, pair, lane, reach);
}

// A pair with an output: its reach vector, its place in the list of such pairs, the pairs it makes one depth down (counted
// per tile here, written by sliced_emit_kernel) and its leaf hits.
template <int RW>
PF_D void sl_record(const SlicedArgs &a, const SlicedTileDev *__restrict__ tm, uint32_t pair, uint32_t lane,
                    const uint32_t (&reach)[RW]) {
    if (lane < 8u) {
        uint32_t v = 0u;
#pragma unroll
        for (int w = 0; w < RW; ++w)
            if (lane == (uint32_t)w) v = reach[w];
        a.reach[(size_t)pair * 8u + lane] = v;
    }
    uint32_t hits = 0u;
#pragma unroll
    for (int w = 0; w < RW; ++w) hits += __popc(reach[w] & tm->leafmask[w]);
    if (lane == 0u) {
        const unsigned long long slot = atomicAdd(a.counters + 1, 1ULL);
        a.alive[slot] = pair;
        if (a.surv_flags) a.surv_flags[a.fr_read ? ldg32(a.fr_read + pair) : a.read0 + pair % a.n_chunk] = 1;
        if (hits) atomicAdd(a.counters + 2, (unsigned long long)hits);
    }
    for (uint32_t c = lane; c < tm->n_children; c += 32u) {
        const uint32_t *cm = a.child_mask + (size_t)(tm->first_child + c) * 8u;
        uint32_t t = 0u;
#pragma unroll
        for (int w = 0; w < RW; ++w) t |= reach[w] & ldg32(cm + w);
        if (t) atomicAdd(a.tile_count + ldg32(a.child_tile + tm->first_child + c), 1u);
    }
    if (a.node_inj_count) {
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            if (!((reach[w] & tm->terminal[w]) >> lane & 1u)) continue;
            const uint32_t c = 32u * w + lane;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t ch = tm->child_node[c][k];
                if (ch != NONE32_D) atomicAdd(a.node_inj_count + ch, 1u);
            }
        }
    }
}

// Persistent grid; a warp takes `grab` consecutive pairs per ticket and works them one after the other.
template <int PW, bool SMALL_M, bool LEAN, int CTAS>
static __global__ void __launch_bounds__(SL_THREADS, CTAS) sliced_probe_kernel(const SlicedArgs a) {
    __shared__ uint32_t s_bits_all[SL_THREADS / 32][8];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t *const s_bits = s_bits_all[threadIdx.x >> 5];
    uint32_t sectors = 0u;
    unsigned long long sectors_total = 0ULL;
    for (;;) {
        uint32_t g0 = 0;
        if (lane == 0) g0 = atomicAdd(a.work_ctr, a.grab);
        g0 = __shfl_sync(0xFFFFFFFFu, g0, 0);
        if (g0 >= a.n_pairs) break;
        const uint32_t g1 = min(g0 + a.grab, a.n_pairs);
        for (uint32_t ii = g0; ii < g1; ++ii) {
            const uint32_t i = a.pair0 + ii;
            uint32_t r, t, src = NONE32_D;
            if (a.fr_read) {
                r = ldg32(a.fr_read + i);
                t = ldg32(a.fr_tile + i);
                src = ldg32(a.fr_src + i);
            } else {
                const uint32_t e = i / a.n_chunk;
                r = a.read0 + (i - e * a.n_chunk);
                t = ldg32(a.entry_tiles + e);
            }
            const SlicedTileDev *tm = a.tiles + t;
            switch (tm->row_words) {
                case 8: sl_pair_and_record<8, PW, SMALL_M, LEAN>(a, tm, i, r, src, lane, s_bits, sectors); break;
                case 4: sl_pair_and_record<4, PW, SMALL_M, LEAN>(a, tm, i, r, src, lane, s_bits, sectors); break;
                case 2: sl_pair_and_record<2, PW, SMALL_M, LEAN>(a, tm, i, r, src, lane, s_bits, sectors); break;
                default: sl_pair_and_record<1, PW, SMALL_M, LEAN>(a, tm, i, r, src, lane, s_bits, sectors); break;
            }
        }
        sectors_total += __reduce_add_sync(0xFFFFFFFFu, sectors);
        sectors = 0u;
    }
    if (lane == 0 && sectors_total) atomicAdd(a.counters, sectors_total);
}

// ---- entry depth on 128-byte lines ------------------------------------------------------------------------------------
// Every read meets every entry tile, and a k-mer's row index is the same in every table (seeds and m are tree-global).  Past
// the reach of the first-level TLB a random load costs the same per (instruction, 128-byte line) whether it brings 4 or 128
// bytes (scripts/mb/mb_coop.cu: 36.7 G lines/s = 4.7 TB/s when four adjacent lanes take a line's four sectors in one
// instruction, against 36.7 G sectors/s = 1.2 TB/s for a sector per lane).  So up to SL_QUAD entry tiles that only filter
// (filter_only, pre-test of 1 or 2 steps -- the usual plan) have their tables interleaved row by row: row i of the shared
// table is one line holding row i of each tile (32 bytes per tile, row_stride = 32 words), and a warp takes one read
// through all of them at once.  Lane l serves tile l & 3; per round of 32 k-mers the warp issues four loads, load j
// bringing the lines of k-mers 8 j + (l >> 2).  Each lane then holds four masks of ITS tile, adds them in place (3 planes)
// and a reduce-scatter over the eight lanes of the tile (lane bits 2..4) leaves every lane with the 6-plane count of one
// 32-column word: 32 lanes = 4 tiles x 8 words.  State per lane is what sl_scan_shallow keeps, for that one word: the
// bit-sliced count of possibly present k-mers and the columns not yet ruled out; a tile drops out once none of its terminal
// columns is left, the read once every tile has.  Outputs are those of the per-tile pairs (pair index = entry position *
// n_chunk + read), so everything downstream is unchanged.
constexpr int SL_QUAD = 4;

// x: this lane's four masks (k-mers q, 8 + q, 16 + q, 24 + q of the round, q = lane >> 2) of the tile lane & 3.  Out: per
// column of word sl_word_of_lane<8>(lane >> 2) of that tile, the number of the round's 32 k-mers whose mask has the bit set.
PF_D void sl_quad_count(const uint32_t (&x)[4][8], uint32_t lane, uint32_t (&cnt)[6]) {
    uint32_t a[6][8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s1 = sl_xor3(x[0][w], x[1][w], x[2][w]), c1 = sl_maj(x[0][w], x[1][w], x[2][w]);
        const uint32_t c2 = s1 & x[3][w];
        a[0][w] = s1 ^ x[3][w];
        a[1][w] = c1 ^ c2;
        a[2][w] = c1 & c2;
    }
#pragma unroll
    for (int st = 0; st < 3; ++st) {
        const int P = 3 + st, H = 8 >> (st + 1), bit = 2 + st;
        const bool hi = (lane >> bit) & 1u;
#pragma unroll
        for (int w = 0; w < H; ++w) {
            uint32_t carry = 0;
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
                const uint32_t lo_v = a[pl][w], hi_v = a[pl][w + H];
                const uint32_t keep = hi ? hi_v : lo_v, send = hi ? lo_v : hi_v;
                const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, send, 1 << bit);
                a[pl][w] = sl_xor3(keep, recv, carry);
                carry = sl_maj(keep, recv, carry);
            }
            a[P][w] = carry;
        }
    }
#pragma unroll
    for (int pl = 0; pl < 6; ++pl) cnt[pl] = a[pl][0];
}

// n_entry: entry positions covered (the first n_entry entries of a.entry_tiles, SL_QUAD per group, the last group
// possibly fewer); all tiles of a group share the table that starts at the first one's table_off.
// FUSE: the hash values are not read from the cache but made on the spot from the 2-bit read (17 <= k <= 32; reads with
// other bytes than ACGT were hashed beforehand and keep the cached path) -- nine reads in ten leave after two or three
// rounds, so most of the batch is never hashed completely nor written to and read back from HBM.
template <int PW, bool SMALL_M, bool FUSE, int CTAS>
static __global__ void __launch_bounds__(SL_THREADS, CTAS) sliced_entry_quad_kernel(const SlicedArgs a, uint32_t n_entry,
                                                                                  uint32_t n_groups) {
    __shared__ uint32_t s_bits_all[SL_THREADS / 32][8];
    const uint32_t lane = threadIdx.x & 31u, sub = lane & 3u, q = lane >> 2;
    const uint32_t my_word = sl_word_of_lane<8>(q);
    const uint32_t sub_lanes = 0x11111111u << sub;  // the lanes serving my tile
    uint32_t *const s_bits = s_bits_all[threadIdx.x >> 5];
    const HashParams &hp = a.hp;
    const uint32_t M0 = (uint32_t)hp.M, M1 = (uint32_t)(hp.M >> 32), m32 = (uint32_t)hp.m;
    unsigned long long lines_total = 0ULL;
    const uint32_t n_items = a.n_chunk * n_groups;  // item = group * n_chunk + read: group-major like the tile-major pairs
    // what this lane keeps of its group's tile records (items are group-major: reloaded only when the group changes)
    uint32_t cur_grp = NONE32_D, valid_mine = 0u, term_mine = 0u;
    bool two = false;
    const uint32_t *__restrict__ slot = nullptr;
    for (;;) {
        uint32_t g0 = 0;
        if (lane == 0) g0 = atomicAdd(a.work_ctr, a.grab);
        g0 = __shfl_sync(0xFFFFFFFFu, g0, 0);
        if (g0 >= n_items) break;
        const uint32_t g1 = min(g0 + a.grab, n_items);
        for (uint32_t item = g0; item < g1; ++item) {
            const uint32_t grp = item / a.n_chunk, ri = item - grp * a.n_chunk, r = a.read0 + ri;
            const uint32_t e0 = grp * SL_QUAD, ne = min((uint32_t)SL_QUAD, n_entry - e0);
            const uint32_t n_k = kmers_of(ldg32(a.lengths + r), hp.k);
            const uint32_t need = need_of(a.threshold, n_k);
            if (need > n_k) continue;  // theta > 1: nothing can pass
            bool on_the_fly = false;
            if (FUSE) on_the_fly = !a.exc_index || ldg32(a.exc_index + r) == NONE32_D;
            // cached hash values (not FUSE, or a read with bytes other than ACGT) ...
            const uint64_t *__restrict__ hbp = nullptr;
            if (!on_the_fly) hbp = a.hb + (__ldg(a.kmer_off + r) - a.kmer_base);
            // ... or the 2-bit read itself: lane l keeps 64-bit word wb + l of it, rounds take their two words by shuffle
            const uint64_t *w64 = nullptr;
            uint64_t wv = 0ULL;
            uint32_t n_w = 0u;  // words the rounds will ask for: 0 .. ((n_k - 1) >> 5) + 1
            if (FUSE && on_the_fly && n_k != 0u) {
                w64 = reinterpret_cast<const uint64_t *>(a.packed + __ldg(a.word_off + r));
                n_w = ((n_k - 1u) >> 5) + 2u;
                if (lane < n_w) wv = __ldg(w64 + lane);
            }
            if (grp != cur_grp) {
                cur_grp = grp;
                const bool tile_on = sub < ne;
                const SlicedTileDev *tm_mine = a.tiles + ldg32(a.entry_tiles + e0 + (tile_on ? sub : 0u));
                valid_mine = tile_on ? tm_mine->valid[my_word] : 0u;
                term_mine = tile_on ? tm_mine->terminal[my_word] : 0u;
                // a second probe step for every tile of the group as soon as one asks for it (more steps are never unsound)
                two = __any_sync(0xFFFFFFFFu, tile_on && tm_mine->pre_steps > 1u);
                slot = a.tables + (a.tiles + ldg32(a.entry_tiles + e0))->table_off + sub * 8u;
            }
            uint32_t alive = valid_mine;
            uint32_t acc[PW];
#pragma unroll
            for (int pl = 0; pl < PW; ++pl) acc[pl] = 0u;
            uint32_t live = __ballot_sync(0xFFFFFFFFu, (alive & term_mine) != 0u);  // lanes whose word still holds a terminal
            if (n_k != 0u && need != 0u) {  // need == 0: every column passes (query.rs:48)
                // row indices of the k-mer this lane owns in the round that starts at k-mer `base` (warp-uniform call)
                auto round_indices = [&](uint32_t base, uint64_t &i0, uint64_t &i1) {
                    uint64_t hbv;
                    if (FUSE && on_the_fly) {  // as hash_kernel: the 32 bases from position base + lane on, canonical, hashed
                        const uint32_t wi = base >> 5, wb = (wi / 31u) * 31u;  // lanes hold words wb .. wb + 31; a round needs wi, wi + 1
                        if (wi == wb && wi != 0u) wv = wb + lane < n_w ? __ldg(w64 + wb + lane) : 0ULL;
                        const uint64_t lo = __shfl_sync(0xFFFFFFFFu, (unsigned long long)wv, wi - wb);
                        const uint64_t hi = __shfl_sync(0xFFFFFFFFu, (unsigned long long)wv, wi - wb + 1u);
                        const uint32_t sh = 2u * lane;
                        hbv = canonical_hash_2bit_rt((lo >> sh) | ((hi << 1) << (63u - sh)), hp.k);
                    } else {
                        hbv = base + lane < n_k ? sl_ld_stream(hbp + base + lane) : 0ULL;
                    }
                    const uint64_t h1 = fx_finish(hp.c1, hbv, hp.rot);
                    if (SMALL_M) i0 = mod_small(h1, M0, M1, m32);
                    else i0 = mod_any(h1, hp.m, hp.M);
                    i1 = 0ULL;
                    if (two) {
                        const uint64_t h2 = fx_finish(hp.c2, hbv, hp.rot);
                        if (SMALL_M) i1 = mod_small(h2, M0, M1, m32);
                        else i1 = mod_any(h2, hp.m, hp.M);
                    }
                };
                uint64_t i0 = 0ULL, i1 = 0ULL;
                for (uint32_t base = 0; base < n_k && live; base += 32u) {
                    round_indices(base, i0, i1);
                    const bool mine_live = (live & sub_lanes) != 0u;
                    uint32_t x[4][8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t kk = 8u * j + q;
                        uint64_t r0i, r1i = 0ULL;
                        if (SMALL_M) {
                            r0i = __shfl_sync(0xFFFFFFFFu, (uint32_t)i0, kk);
                            if (two) r1i = __shfl_sync(0xFFFFFFFFu, (uint32_t)i1, kk);
                        } else {
                            r0i = __shfl_sync(0xFFFFFFFFu, (unsigned long long)i0, kk);
                            if (two) r1i = __shfl_sync(0xFFFFFFFFu, (unsigned long long)i1, kk);
                        }
#pragma unroll
                        for (int w = 0; w < 8; ++w) x[j][w] = 0u;
                        if (mine_live && base + kk < n_k) {
                            sl_load_row<8>(slot + r0i * 32u, x[j]);
                            if (two) {
                                uint32_t y[8];
                                sl_load_row<8>(slot + r1i * 32u, y);
#pragma unroll
                                for (int w = 0; w < 8; ++w) x[j][w] &= y[w];
                            }
                        }
                    }
                    const uint32_t done = min(base + 32u, n_k), rest = n_k - done;
                    if (lane == 0u) lines_total += (unsigned long long)(done - base) * (two ? 2u : 1u);
                    uint32_t cnt[6];
                    sl_quad_count(x, lane, cnt);
                    uint32_t carry = 0u;
#pragma unroll
                    for (int pl = 0; pl < PW; ++pl) {
                        const uint32_t v = pl < 6 ? cnt[pl] : 0u;
                        const uint32_t sum = sl_xor3(acc[pl], v, carry);
                        carry = sl_maj(acc[pl], v, carry);
                        acc[pl] = sum;
                    }
                    // a column cannot pass any more once its count plus all remaining k-mers stays below `need`
                    if (need > rest) {
                        alive &= sl_ge<PW>(acc, need - rest);
                        live = __ballot_sync(0xFFFFFFFFu, (alive & term_mine) != 0u);
                    }
                }
            }
            // columns the pre-test could not rule out count as passed (filter-only tiles: the exact tiles below decide)
            for (uint32_t t = 0; t < ne; ++t) {
                // this tile's words to where sl_reach expects them: lane l holds word sl_word_of_lane<8>(l)
                const uint32_t pass_t = __shfl_sync(0xFFFFFFFFu, alive, 4u * (lane & 7u) + t);
                if (!(live & (0x11111111u << t))) continue;  // warp-uniform
                const SlicedTileDev *tm = a.tiles + ldg32(a.entry_tiles + e0 + t);
                uint32_t reach[8];
                if (sl_reach<8>(a, tm, NONE32_D, lane, s_bits, pass_t, reach)) sl_record<8>(a, tm, (e0 + t) * a.n_chunk + ri, lane, reach);
            }
        }
    }
    if (lane == 0 && lines_total) atomicAdd(a.counters + 3, lines_total);
}

// Second pass over the pairs that have an output: leaf hits go to the hit list and the per-leaf histogram
// (mapped_reads, query.rs:143; ResultMap::add_read_map, result_map.rs:20-22), successor tiles get their (read, tile,
// source pair) entries, tile-major.
struct SlicedEmitArgs {
    const uint32_t *fr_read, *fr_tile;  // null at the entry depth (see SlicedArgs)
    const uint32_t *entry_tiles;
    uint32_t read0, n_chunk;
    const uint32_t *alive;
    uint32_t n_alive;
    const uint32_t *reach;
    const SlicedTileDev *tiles;
    const uint32_t *child_tile, *child_mask;
    const unsigned long long *tile_base;  // per tile: first slot in the next frontier
    uint32_t *tile_cursor;
    uint32_t *nx_read, *nx_tile, *nx_src;
    uint32_t *hit_read, *hit_leaf, *read_hits;
    unsigned long long *hit_cursor;  // next free slot of the hit list
    unsigned long long *blk_counts;
    int want_hits;
    // hand-over to the node-at-a-time descent: (read, node) pairs, node-major (null: tiles all the way down)
    const unsigned long long *node_inj_base;
    uint32_t *node_inj_cursor;
    uint32_t *inj_read, *inj_node;
};
static __global__ void sliced_emit_kernel(const SlicedEmitArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t x = warp; x < a.n_alive; x += n_warps) {
        const uint32_t i = ldg32(a.alive + x);
        uint32_t r, t;
        if (a.fr_read) {
            r = ldg32(a.fr_read + i);
            t = ldg32(a.fr_tile + i);
        } else {
            const uint32_t e = i / a.n_chunk;
            r = a.read0 + (i - e * a.n_chunk);
            t = ldg32(a.entry_tiles + e);
        }
        const SlicedTileDev *tm = a.tiles + t;
        uint32_t reach[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) reach[w] = ldg32(a.reach + (size_t)i * 8u + w);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t bits = reach[w] & tm->leafmask[w];
            if (!bits) continue;  // warp-uniform
            const uint32_t n = __popc(bits);
            const bool mine = (bits >> lane) & 1u;
            const int32_t lf = mine ? tm->leaf[32 * w + lane] : -1;
            if (mine) atomicAdd(a.blk_counts + lf, 1ULL);
            if (a.want_hits) {
                unsigned long long p0 = 0;
                if (lane == 0) p0 = atomicAdd(a.hit_cursor, (unsigned long long)n);
                p0 = __shfl_sync(0xFFFFFFFFu, p0, 0);
                if (mine) {
                    const unsigned long long p = p0 + __popc(bits & ((1u << lane) - 1u));
                    a.hit_read[p] = r;
                    a.hit_leaf[p] = (uint32_t)lf;
                }
                if (lane == 0 && a.want_hits == 1) atomicAdd(a.read_hits + r, n);
            }
        }
        for (uint32_t c = lane; c < tm->n_children; c += 32u) {
            const uint32_t *cm = a.child_mask + (size_t)(tm->first_child + c) * 8u;
            uint32_t any = 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w) any |= reach[w] & ldg32(cm + w);
            if (any) {
                const uint32_t ct = ldg32(a.child_tile + tm->first_child + c);
                const unsigned long long p = a.tile_base[ct] + atomicAdd(a.tile_cursor + ct, 1u);
                a.nx_read[p] = r;
                a.nx_tile[p] = ct;
                a.nx_src[p] = i;
            }
        }
        if (a.node_inj_base) {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (!((reach[w] & tm->terminal[w]) >> lane & 1u)) continue;
                const uint32_t c = 32u * w + lane;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint32_t ch = tm->child_node[c][k];
                    if (ch == NONE32_D) continue;
                    const unsigned long long p = a.node_inj_base[ch] + atomicAdd(a.node_inj_cursor + ch, 1u);
                    a.inj_read[p] = r;
                    a.inj_node[p] = ch;
                }
            }
        }
    }
}

// Builds one tile's table from the row-major filters: block (bx, tile) transposes the 256 bit positions
// [256 bx, 256 bx + 256) of up to 256 filters with warp ballots.  Warp j handles the columns 32 j .. 32 j + 31.
static __global__ void __launch_bounds__(256) slice_kernel(const uint64_t *__restrict__ filters, uint64_t wpf,
                                                           const uint32_t *__restrict__ col_slot,  // [tiles][256]
                                                           const SlicedTileDev *__restrict__ tiles, uint32_t tile0,
                                                           uint32_t *tables) {
    const uint32_t t = tile0 + blockIdx.y;
    const SlicedTileDev *tm = tiles + t;
    const uint32_t stride = tm->row_stride;
    const uint32_t rw = stride == 32u ? 8u : tm->row_words;  // a slot in a shared line is written whole (unused columns: zero)
    const uint32_t wj = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (wj >= rw) return;
    const uint32_t slot = col_slot[(size_t)t * SL_MAX_COLS + 32u * wj + lane];
    const uint64_t w0 = (uint64_t)blockIdx.x * 8u;  // first u32 word of this block's bit range
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0u;
    if (slot != NONE32_D && w0 < 2u * wpf) {
        const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(filters + (uint64_t)slot * wpf) + w0);
        const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        x[0] = v0.x, x[1] = v0.y, x[2] = v0.z, x[3] = v0.w, x[4] = v1.x, x[5] = v1.y, x[6] = v1.z, x[7] = v1.w;
    }
    uint32_t *out = tables + tm->table_off + (w0 * 32u) * stride + wj;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t mine = 0u;
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const uint32_t v = __ballot_sync(0xFFFFFFFFu, (x[i] >> b) & 1u);
            if (lane == (uint32_t)b) mine = v;
        }
        out[((uint64_t)i * 32u + lane) * stride] = mine;  // row 32 (w0 + i) + lane, word wj
    }
}

}  // namespace pf
