// pf_kernels.cuh -- device code of the query path (sm_100a).
//
//   probe_kernel    one warp owns one (read,node) pair of the frontier: lanes take consecutive k-mers
//                   in rounds of 32, re-create the reference's canonical-k-mer hashes in registers
//                   from 2-bit codes, and gather filter bits with k-mer-level early exit
//                   (BloomFilter::contains, bloom_filter.rs:312-332; query_passes, query.rs:38-49).
//   level_scan_kernel / scatter_kernel
//                   prune the frontier at the threshold and expand the survivors to both children,
//                   node-major, with warp-aggregated atomics and a prefix sum (_query_batch,
//                   query.rs:99-158); leaf passes go to the per-leaf histogram and the hit list
//                   (mapped_reads, query.rs:143; ResultMap::add_read_map, result_map.rs:20-22).
#pragma once
#include "pf_hash.cuh"

namespace pf {

constexpr uint32_t NONE32_D = 0xFFFFFFFFu;  // "no child" / "not an exception read"
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_CHUNK = 8;  // pairs fetched per warp per work-counter atomic

struct ProbeArgs {
    // frontier
    const uint32_t *fr_read;
    const uint32_t *fr_node;
    uint32_t n_pairs;
    // read batch
    const uint32_t *lengths;
    const uint64_t *word_off;
    const uint32_t *packed;
    const uint32_t *exc_index;  // may be null
    const uint64_t *exc_off;
    const uint8_t *exc_bytes;
    // tree
    const uint32_t *node_slot;
    const uint64_t *filters;
    uint64_t words_per_filter;
    // outputs
    uint8_t *pass;
    uint32_t *node_pass;
    unsigned int *work_ctr;
    unsigned long long *probes;
    HashParams hp;
    float threshold;
    int exhaustive;
};

// (threshold * n_k as f32).ceil() as usize   (query.rs:48): f32 product, ceil, saturating cast.
PF_D uint32_t need_of(float threshold, uint32_t n_k) {
    float c = ceilf(__fmul_rn(threshold, __uint2float_rn(n_k)));
    if (!(c > 0.0f)) return 0u;  // negative and NaN saturate to 0
    if (c >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)c;
}

PF_D uint32_t ldg32(const uint32_t *p) { return __ldg(p); }

// One round: every active lane owns one k-mer (h1,h2).  Step i makes each still-alive lane test bit
// g_i mod m.  Returns the mask of lanes whose k-mer is contained.  `failed` is set as soon as the
// pair can no longer reach its bound (read-level early exit; result-identical to counting all).
template <bool SMALL_M>
PF_D uint32_t probe_round(const uint32_t *__restrict__ filt, const HashParams &hp, uint64_t h1, uint64_t h2, bool active,
                          uint32_t misses_before, uint32_t allowed, bool exhaustive, uint32_t &probes, bool &failed) {
    const uint32_t act_mask = __ballot_sync(0xFFFFFFFFu, active);
    uint32_t alive_mask = act_mask;
    bool alive = active;
    uint64_t g = h1;
    const uint32_t M0 = (uint32_t)hp.M, M1 = (uint32_t)(hp.M >> 32), m32 = (uint32_t)hp.m;
    for (uint32_t i = 0; i < hp.K; ++i) {
        probes += __popc(alive_mask);
        if (alive) {
            uint32_t w, bit;
            if (SMALL_M) {
                uint32_t idx = mod_small(g, M0, M1, m32);
                w = ldg32(filt + (idx >> 5));
                bit = idx & 31u;
            } else {
                uint64_t idx = mod_any(g, hp.m, hp.M);
                w = ldg32(filt + (idx >> 5));
                bit = (uint32_t)idx & 31u;
            }
            alive = (w >> bit) & 1u;
        }
        alive_mask = __ballot_sync(0xFFFFFFFFu, alive);
        if (!exhaustive && misses_before + __popc(act_mask & ~alive_mask) > allowed) {
            failed = true;
            return alive_mask;
        }
        if (alive_mask == 0u) break;
        // g_{i+1}: g1 = h2, g2 = (h1+2)*h2, then g_{i+1} = g_i + h2   (hash_iter.rs:17-24)
        g = i == 0 ? h2 : (i == 1 ? (h1 + 2ULL) * h2 : g + h2);
    }
    return alive_mask;
}

struct ByteSrc {
    const uint8_t *ascii;    // exception read: raw bytes
    const uint32_t *packed;  // otherwise 2-bit codes
};
PF_D uint8_t src_byte(const ByteSrc &s, uint32_t j) {
    if (s.ascii) return s.ascii[j];
    uint32_t w = ldg32(s.packed + (j >> 4));
    uint32_t c = (w >> (2u * (j & 15u))) & 3u;
    return (uint8_t)(0x54474341u >> (8u * c));
}

// Evaluate one (read,node) pair; warp-uniform control flow.  Returns pass/fail (query_passes).
template <int KM>
PF_D bool probe_pair(const ProbeArgs &a, uint32_t r, uint32_t u, uint32_t lane, uint32_t &probes) {
    const HashParams &hp = a.hp;
    const uint32_t len = ldg32(a.lengths + r);
    const uint32_t k = hp.k;
    const uint32_t n_k = (k == 0u || k > len) ? 0u : len - k + 1u;  // file_parser.rs:136-139
    const uint32_t need = need_of(a.threshold, n_k);
    const bool exhaustive = a.exhaustive != 0;
    if (!exhaustive) {
        if (need == 0u) return true;   // hits >= 0 always
        if (need > n_k) return false;  // hits <= n_k < need
    }
    const uint32_t allowed = need > n_k ? 0u : n_k - need;
    const uint32_t *filt =
        reinterpret_cast<const uint32_t *>(a.filters + (uint64_t)ldg32(a.node_slot + u) * a.words_per_filter);
    const uint64_t woff = __ldg(a.word_off + r);
    uint32_t e = a.exc_index ? ldg32(a.exc_index + r) : NONE32_D;
    uint32_t hits = 0, misses = 0;
    bool failed = false;

    if (KM != 0 && e == NONE32_D) {
        const uint64_t *w64 = reinterpret_cast<const uint64_t *>(a.packed + woff);
        for (uint32_t base = 0; base < n_k; base += 32u) {
            const uint64_t lo = __ldg(w64 + (base >> 5)), hi = __ldg(w64 + (base >> 5) + 1);
            const uint32_t sh = 2u * lane;
            const uint64_t x = (lo >> sh) | ((hi << 1) << (63u - sh));
            const uint64_t hb = canonical_hash_2bit<(KM ? KM : 17)>(x);
            const uint64_t h1 = fx_finish(hp.c1, hb, hp.rot), h2 = fx_finish(hp.c2, hb, hp.rot);
            const bool active = base + lane < n_k;
            const uint32_t cnt = min(32u, n_k - base);
            const uint32_t ok = probe_round<true>(filt, hp, h1, h2, active, misses, allowed, exhaustive, probes, failed);
            if (failed) return false;
            const uint32_t h = __popc(ok);
            hits += h;
            misses += cnt - h;
            if (!exhaustive && hits >= need) return true;
        }
    } else {
        ByteSrc s;
        s.ascii = e == NONE32_D ? nullptr : a.exc_bytes + __ldg(a.exc_off + e);
        s.packed = a.packed + woff;
        for (uint32_t base = 0; base < n_k; base += 32u) {
            const uint32_t pos = base + lane;
            const bool active = pos < n_k;
            uint64_t h1 = 0, h2 = 0;
            if (active) {
                const uint64_t hb = canonical_hash_bytes([&](uint32_t j) { return src_byte(s, pos + j); }, k);
                h1 = fx_finish(hp.c1, hb, hp.rot);
                h2 = fx_finish(hp.c2, hb, hp.rot);
            }
            const uint32_t cnt = min(32u, n_k - base);
            uint32_t ok;
            if (hp.small_m) ok = probe_round<true>(filt, hp, h1, h2, active, misses, allowed, exhaustive, probes, failed);
            else ok = probe_round<false>(filt, hp, h1, h2, active, misses, allowed, exhaustive, probes, failed);
            if (failed) return false;
            const uint32_t h = __popc(ok);
            hits += h;
            misses += cnt - h;
            if (!exhaustive && hits >= need) return true;
        }
    }
    return hits >= need;
}

// Persistent grid; warps pull PROBE_CHUNK consecutive pairs at a time from a global counter.
template <int KM>
__global__ void __launch_bounds__(PROBE_THREADS, 4) probe_kernel(const ProbeArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t probes = 0;
    unsigned long long probes_total = 0;
    for (;;) {
        uint32_t i0 = 0;
        if (lane == 0) i0 = atomicAdd(a.work_ctr, (unsigned)PROBE_CHUNK);
        i0 = __shfl_sync(0xFFFFFFFFu, i0, 0);
        if (i0 >= a.n_pairs) break;
        const uint32_t i1 = min(i0 + (uint32_t)PROBE_CHUNK, a.n_pairs);
        for (uint32_t i = i0; i < i1; ++i) {
            const uint32_t r = ldg32(a.fr_read + i), u = ldg32(a.fr_node + i);
            const bool pass = probe_pair<KM>(a, r, u, lane, probes);
            if (lane == 0) {
                a.pass[i] = pass ? 1 : 0;
                if (pass) atomicAdd(a.node_pass + u, 1u);
            }
        }
        probes_total += probes;
        probes = 0;
    }
    if (lane == 0 && probes_total) atomicAdd(a.probes, probes_total);
}

// ---- frontier bookkeeping ------------------------------------------------------------------
__global__ void init_frontier_kernel(uint32_t *fr_read, uint32_t *fr_node, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        fr_read[i] = i;
        fr_node[i] = 0u;  // root has level-order id 0
    }
}

struct LevelTotals {
    unsigned long long next_pairs;
    unsigned long long hits_total;  // running total of (read,leaf) hits in this block
    unsigned long long probes;
    unsigned long long pad;
};

// One block.  For the nodes [lo,hi) of the current level: exclusive scan of the surviving pair counts
// gives every child its slice of the next frontier (children are numbered level-order, left before
// right, so the slices are node-major) and every leaf its slice of the hit list; leaf counts are added
// to the block histogram.
__global__ void level_scan_kernel(uint32_t lo, uint32_t hi, const uint32_t *__restrict__ node_pass,
                                  const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
                                  const int32_t *__restrict__ leaf, unsigned long long *next_base,
                                  unsigned long long *hit_base, unsigned long long *blk_counts, LevelTotals *totals,
                                  const unsigned long long *probes) {
    __shared__ unsigned long long s_next[1024], s_hit[1024];
    const uint32_t t = threadIdx.x, nt = blockDim.x;
    const uint32_t n = hi - lo;
    const uint32_t per = (n + nt - 1) / nt;
    const uint32_t b = lo + min(n, t * per), e = lo + min(n, (t + 1) * per);
    unsigned long long sn = 0, sh = 0;
    for (uint32_t u = b; u < e; ++u) {
        const unsigned long long c = node_pass[u];
        if (leaf[u] >= 0) sh += c;
        else sn += c * ((left[u] != NONE32_D) + (right[u] != NONE32_D));
    }
    s_next[t] = sn;
    s_hit[t] = sh;
    __syncthreads();
    if (t == 0) {
        unsigned long long an = 0, ah = totals->hits_total;
        for (uint32_t i = 0; i < nt; ++i) {
            unsigned long long x = s_next[i], y = s_hit[i];
            s_next[i] = an;
            s_hit[i] = ah;
            an += x;
            ah += y;
        }
        totals->next_pairs = an;
        totals->hits_total = ah;
        totals->probes = *probes;
    }
    __syncthreads();
    unsigned long long an = s_next[t], ah = s_hit[t];
    for (uint32_t u = b; u < e; ++u) {
        const unsigned long long c = node_pass[u];
        if (leaf[u] >= 0) {
            hit_base[u] = ah;
            ah += c;
            if (c) blk_counts[leaf[u]] += c;  // each leaf is one node: no race
        } else {
            next_base[u] = an;
            an += c * ((left[u] != NONE32_D) + (right[u] != NONE32_D));
        }
    }
}

// Thread per pair: survivors take a rank inside their node (warp-aggregated atomic) and are written
// to both children's slices, or to the hit list when the node is a leaf.
__global__ void scatter_kernel(const uint32_t *__restrict__ fr_read, const uint32_t *__restrict__ fr_node,
                               const uint8_t *__restrict__ pass, uint32_t n, const uint32_t *__restrict__ node_pass,
                               uint32_t *cursor, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
                               const int32_t *__restrict__ leaf, const unsigned long long *__restrict__ next_base,
                               const unsigned long long *__restrict__ hit_base, uint32_t *nx_read, uint32_t *nx_node,
                               uint32_t *hit_read, uint32_t *hit_leaf, int want_hits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n && pass[i];
    const uint32_t act = __ballot_sync(0xFFFFFFFFu, on);
    if (!on) return;
    const uint32_t u = fr_node[i], r = fr_read[i];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t peers = __match_any_sync(act, u);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor + u, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t rank = base + __popc(peers & ((1u << lane) - 1u));
    const int32_t lf = leaf[u];
    if (lf >= 0) {
        if (want_hits) {
            const unsigned long long p = hit_base[u] + rank;
            hit_read[p] = r;
            hit_leaf[p] = (uint32_t)lf;
        }
    } else {
        unsigned long long p = next_base[u] + rank;
        const uint32_t c = node_pass[u];
        const uint32_t l = left[u], rr = right[u];
        if (l != NONE32_D) {
            nx_read[p] = r;
            nx_node[p] = l;
            p += c;
        }
        if (rr != NONE32_D) {
            nx_read[p] = r;
            nx_node[p] = rr;
        }
    }
}

__global__ void add_counts_kernel(unsigned long long *dst, const unsigned long long *src, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

}  // namespace pf
