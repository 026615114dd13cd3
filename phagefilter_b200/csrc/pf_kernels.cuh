// pf_kernels.cuh -- device code of the query path (sm_100a).
//
//   hash_kernel<k>  once per batch: one warp per read, lanes take 32 consecutive k-mers per round and
//                   re-create the reference's canonical k-mer bytes in registers from 2-bit codes
//                   (file_parser.rs:114-148), then rustc-hash's hash_bytes over them.  The 64-bit result is
//                   seed- and node-independent, so it is cached (8 B per k-mer) for the whole descent; the
//                   reference re-hashes every k-mer at every node (hash_iter.rs:31-45).
//   probe_kernel    one warp owns one (read,node) pair of the frontier: h1,h2 from the cached value, then
//                   step i tests bit g_i mod m of the node's filter -- BloomFilter::contains
//                   (bloom_filter.rs:312-332) with its k-mer-level early exit, query_passes (query.rs:38-49)
//                   with a result-identical read-level early exit.
//   level_scan_kernel / scatter_kernel
//                   prune the frontier at the threshold and expand the survivors to both children,
//                   node-major, with warp-aggregated atomics and a prefix sum (_query_batch,
//                   query.rs:99-158); leaf passes go to the per-leaf histogram and the hit list
//                   (mapped_reads, query.rs:143; ResultMap::add_read_map, result_map.rs:20-22).
#pragma once
#include "pf_hash.cuh"

namespace pf {

constexpr uint32_t NONE32_D = 0xFFFFFFFFu;  // "no child" / "not an exception read"
constexpr int NODE_PASS_COPIES = 16;  // copies of the per-node survivor counters (spreads same-address atomics)
constexpr int PROBE_THREADS = 256;
constexpr int PROBE_CHUNK = 8;   // pairs whose first round is batched together
constexpr int PROBE_GRAB = 32;   // most pairs fetched per warp per work-counter atomic (one per lane); ProbeArgs::grab
constexpr int HASH_THREADS = 256;

PF_D uint32_t ldg32(const uint32_t *p) { return __ldg(p); }
// Streamed once per level (frontier, cached hash values and step-0 indices, pass flags): evict-first loads / stores
// (ld.global.cs / st.global.cs) so that they do not displace the filters being probed from L2.
PF_D uint32_t lds32(const uint32_t *p) { return __ldcs(p); }
PF_D uint64_t lds64(const uint64_t *p) { return __ldcs(reinterpret_cast<const unsigned long long *>(p)); }

// number of k-mers of a read (file_parser.rs:136-139)
PF_HD uint32_t kmers_of(uint32_t len, uint32_t k) { return (k == 0u || k > len) ? 0u : len - k + 1u; }
// Slot of k-mer p in a read's idx0 cache: the k-mers are stored in four residue classes (p mod 4), each class
// contiguous, so a stride-4 sample (the usual sampled pre-test) is one coalesced run -- 4 sectors per 32 lanes instead
// of 16 -- while 32 consecutive k-mers still fall into four runs of 8 (4-8 sectors).
// Classes c < r hold q4 k-mers, the others q4 - 1 (r = number of full classes), so the n_k slots are used exactly.
PF_HD uint32_t idx0_slot(uint32_t p, uint32_t n_k) {
    const uint32_t q4 = (n_k + 3u) >> 2, r = ((n_k - 1u) & 3u) + 1u, c = p & 3u;
    return c * q4 - (c > r ? c - r : 0u) + (p >> 2);
}

// ---- hash_kernel ---------------------------------------------------------------------------------
struct HashArgs {
    const uint32_t *lengths;
    const uint64_t *word_off;
    const uint32_t *packed;
    const uint32_t *exc_index;  // may be null
    const uint64_t *exc_off;
    const uint8_t *exc_bytes;
    const uint64_t *kmer_off;   // [n_reads + 1] prefix sum of k-mer counts
    uint64_t *hb;               // out: hash_bytes(canonical k-mer), index kmer_off[r] - kmer_base + pos
    uint32_t *idx0;             // out (m < 2^31 only): bit index of probe step 0, h1 mod m; else null
    HashParams hp;
    uint64_t kmer_base;         // kmer_off[read0]
    uint32_t read0, n_reads;    // reads [read0, read0 + n_reads) of the batch
    uint32_t k;
    unsigned int *work_ctr;
    const uint8_t *flags;       // optional [batch reads]: only reads with a non-zero flag are hashed
    uint32_t grab;              // reads per ticket of the work counter (>= 1)
};

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

// KM in 17..32: 2-bit register path for reads of pure upper-case ACGT; KM == 0: byte path for every read.
// Exception reads (any other byte) always take the byte path, which is what the reference hashes.
template <int KM>
static __global__ void __launch_bounds__(HASH_THREADS) hash_kernel(const HashArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t M0 = (uint32_t)a.hp.M, M1 = (uint32_t)(a.hp.M >> 32), m32 = (uint32_t)a.hp.m;
    for (;;) {
        uint32_t i0 = 0;
        if (lane == 0) i0 = atomicAdd(a.work_ctr, a.grab);  // same-address atomic: as few tickets as balance allows
        i0 = __shfl_sync(0xFFFFFFFFu, i0, 0);
        if (i0 >= a.n_reads) break;
        const uint32_t i1 = min(i0 + a.grab, a.n_reads);
        for (uint32_t i = i0; i < i1; ++i) {
            const uint32_t r = a.read0 + i;
            if (a.flags && !a.flags[r]) continue;
            const uint32_t n_k = kmers_of(ldg32(a.lengths + r), a.k);
            if (n_k == 0) continue;
            const uint64_t kofs = __ldg(a.kmer_off + r) - a.kmer_base;
            uint64_t *out = a.hb + kofs;
            uint32_t *out0 = a.idx0 ? a.idx0 + kofs : nullptr;
            const uint64_t woff = __ldg(a.word_off + r);
            const uint32_t e = a.exc_index ? ldg32(a.exc_index + r) : NONE32_D;
            if (KM != 0 && e == NONE32_D) {
                const uint64_t *w64 = reinterpret_cast<const uint64_t *>(a.packed + woff);
                uint64_t lo = __ldg(w64);
                for (uint32_t base = 0; base < n_k; base += 32u) {
                    const uint64_t hi = __ldg(w64 + (base >> 5) + 1);
                    const uint32_t sh = 2u * lane;
                    const uint64_t x = (lo >> sh) | ((hi << 1) << (63u - sh));
                    const uint64_t hb = canonical_hash_2bit<(KM ? KM : 17)>(x);
                    if (base + lane < n_k) {
                        __stcs(reinterpret_cast<unsigned long long *>(out + base + lane), (unsigned long long)hb);
                        if (out0) __stcs(out0 + idx0_slot(base + lane, n_k), mod_small(fx_finish(a.hp.c1, hb, a.hp.rot), M0, M1, m32));
                    }
                    lo = hi;
                }
            } else {
                ByteSrc s;
                s.ascii = e == NONE32_D ? nullptr : a.exc_bytes + __ldg(a.exc_off + e);
                s.packed = a.packed + woff;
                for (uint32_t pos = lane; pos < n_k; pos += 32u) {
                    const uint64_t hb = canonical_hash_bytes([&](uint32_t j) { return src_byte(s, pos + j); }, a.k);
                    out[pos] = hb;
                    if (out0) out0[idx0_slot(pos, n_k)] = mod_small(fx_finish(a.hp.c1, hb, a.hp.rot), M0, M1, m32);
                }
            }
        }
    }
}

// ---- probe_kernel --------------------------------------------------------------------------------
struct ProbeArgs {
    // frontier
    const uint32_t *fr_read;
    const uint32_t *fr_node;
    uint32_t n_pairs;
    // reads
    const uint32_t *lengths;
    const uint64_t *kmer_off;
    const uint64_t *hb;
    const uint32_t *idx0;  // step-0 bit indices (m < 2^31), else null
    uint64_t kmer_base;
    // tree
    const uint32_t *node_slot;
    const uint32_t *node_steps;  // probe steps per k-mer at this node: K (exact) or fewer (sound pre-test)
    const uint64_t *filters;
    uint64_t words_per_filter;
    // outputs
    uint8_t *pass;
    uint32_t *node_pass;  // [NODE_PASS_COPIES][n_nodes]
    uint32_t n_nodes;
    unsigned int *work_ctr;
    unsigned long long *probes;
    HashParams hp;
    float threshold;
    int exhaustive;
    // k-mer memo of the exact nodes of this level (null = off): one power-of-two region of entries per exact node,
    // node_memo[u] = (first entry / 4096) << 5 | log2(entries), or NONE32.  An entry holds hash_bytes(k-mer) of a k-mer that passed all K probes at
    // that node; contains() depends on the k-mer only through that value, so a match is exact, never probabilistic.
    unsigned long long *memo;
    const uint32_t *node_memo;
    uint32_t grab;  // pairs per ticket: PROBE_GRAB, or PROBE_CHUNK when pairs are long or few (load balance)
    uint32_t order_streams, order_span;  // chunk order: ticket c -> chunk (c % streams) * span + c / streams
};

// (threshold * n_k as f32).ceil() as usize   (query.rs:48): f32 product, ceil, saturating cast.
PF_D uint32_t need_of(float threshold, uint32_t n_k) {
    float c = ceilf(__fmul_rn(threshold, __uint2float_rn(n_k)));
    if (!(c > 0.0f)) return 0u;  // negative and NaN saturate to 0
    if (c >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)c;
}

// One group = up to G rounds of 32 consecutive k-mers; lane L owns k-mers gbase + j*32 + L, j < G, so up to
// G independent gathers per lane are in flight in every step.  Steps follow BloomFilter::contains: step i
// tests bit g_i mod m for every k-mer that is still alive (g0 = h1, g1 = h2, g_i = (h1+i)*h2 = g_{i-1} + h2;
// hash_iter.rs:17-24).  Phases: step 0 of the first round alone (cheap rejection of reads that do not
// belong below this node), step 0 of the other rounds, then steps 1..n_steps-1 of all rounds.  After every
// phase the pair fails as soon as more than `limit` k-mers of the group are known to be absent
// (read-level early exit; counting on could not change the outcome).
template <int G, bool SMALL_M>
struct GroupState {
    uint64_t h1[G], h2[G], g[G];
    uint32_t alive;  // bit j: k-mer j of this lane has had no clear bit so far
};

template <int G, bool SMALL_M, int J0, int J1>
PF_D uint32_t probe_phase(const uint32_t *__restrict__ filt, const HashParams &hp, GroupState<G, SMALL_M> &st) {
    const uint32_t M0 = (uint32_t)hp.M, M1 = (uint32_t)(hp.M >> 32), m32 = (uint32_t)hp.m;
    uint32_t w[G], bit[G];
#pragma unroll
    for (int j = J0; j < J1; ++j) {  // issue every gather of the phase before using any
        w[j] = 0xFFFFFFFFu;
        bit[j] = 0;
        if ((st.alive >> j) & 1u) {
            if (SMALL_M) {
                const uint32_t idx = mod_small(st.g[j], M0, M1, m32);
                w[j] = ldg32(filt + (idx >> 5));
                bit[j] = idx & 31u;
            } else {
                const uint64_t idx = mod_any(st.g[j], hp.m, hp.M);
                w[j] = ldg32(filt + (idx >> 5));
                bit[j] = (uint32_t)idx & 31u;
            }
        }
    }
    uint32_t died = 0;
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        if (!((w[j] >> bit[j]) & 1u)) {
            st.alive &= ~(1u << j);
            ++died;
        }
    }
    return died;  // this lane's k-mers found absent in this phase
}

// one gather per alive k-mer of slots [J0,J1) at precomputed bit indices; returns this lane's newly absent k-mers
template <int G, int J0, int J1>
PF_D uint32_t probe_idx_phase(const uint32_t *__restrict__ filt, const uint32_t (&idx)[G], uint32_t &alive) {
    uint32_t w[G];
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        w[j] = 0xFFFFFFFFu;
        if ((alive >> j) & 1u) w[j] = ldg32(filt + (idx[j] >> 5));
    }
    uint32_t died = 0;
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        if (!((w[j] >> (idx[j] & 31u)) & 1u)) {
            alive &= ~(1u << j);
            ++died;
        }
    }
    return died;
}

// Returns true when the pair's outcome is decided (`pass` set); hits/misses updated otherwise.
// SMALL_M: step 0 uses the cached bit indices (no hashing, no modulo); the cached hash_bytes values are only
// fetched when the node needs more than one step -- speculatively, together with the step-0 gathers.
// The group's k-mer slots q = gbase + j*32 + lane map to the read's k-mers off + q*stride (stride 1, off 0 unless the
// node samples: see probe_pair); n_k is the number of slots.
template <int G, bool SMALL_M>
PF_D bool probe_group(const uint32_t *__restrict__ filt, const HashParams &hp, uint32_t n_steps,
                      const uint64_t *__restrict__ hbp, const uint32_t *__restrict__ i0p, uint32_t gbase, uint32_t n_k,
                      uint32_t lane, uint32_t need, uint32_t allowed, bool exhaustive, uint32_t stride, uint32_t off,
                      int pre, uint32_t nk_full, unsigned long long *memo, uint32_t memo_mask, unsigned long long *stage,
                      uint32_t &hits, uint32_t &misses, uint32_t &probes, uint32_t &my_probes, uint32_t &memo_hits,
                      uint32_t &memo_lookups, bool &pass) {
    // pre >= 0: step 0 of the first round was already done for the whole chunk of pairs (probe_kernel); pre is this
    // lane's result (1 = its k-mer's bit was clear)
    const uint32_t kidx = off + (gbase + lane) * stride, kstep = 32u * stride;  // k-mer of slot j: kidx + j*kstep
    GroupState<G, SMALL_M> st;
    st.alive = 0;
    const uint32_t cnt = min(32u * G, n_k - gbase);
    const uint32_t limit = allowed - misses;  // only used when !exhaustive (misses <= allowed then)
    uint32_t dead = 0;
    uint64_t hb[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        hb[j] = 0;
        if (gbase + (uint32_t)j * 32u + lane < n_k) st.alive |= 1u << j;
    }
    if (SMALL_M) {
        uint32_t idx[G];
#pragma unroll
        for (int j = 0; j < G; ++j) idx[j] = 0;
        if (pre < 0 && (st.alive & 1u)) idx[0] = lds32(i0p + idx0_slot(kidx, nk_full));
        if (n_steps > 1u) {
#pragma unroll
            for (int j = 0; j < G; ++j)
                if ((st.alive >> j) & 1u) hb[j] = lds64(hbp + kidx + (uint32_t)j * kstep);
        }
        // step 0, first round
        probes += min(32u, cnt);
        if (pre >= 0) {
            if (pre) st.alive &= ~1u;
            dead += __popc(__ballot_sync(0xFFFFFFFFu, pre != 0));
        } else {
            dead += __reduce_add_sync(0xFFFFFFFFu, probe_idx_phase<G, 0, 1>(filt, idx, st.alive));
        }
        if (!exhaustive && dead > limit) {
            pass = false;
            return true;
        }
        // step 0, remaining rounds
        if (G > 1 && cnt > 32u) {
#pragma unroll
            for (int j = 1; j < G; ++j)
                if ((st.alive >> j) & 1u) idx[j] = lds32(i0p + idx0_slot(kidx + (uint32_t)j * kstep, nk_full));
            probes += cnt - 32u;
            dead += __reduce_add_sync(0xFFFFFFFFu, probe_idx_phase<G, (G > 1 ? 1 : 0), G>(filt, idx, st.alive));
            if (!exhaustive && dead > limit) {
                pass = false;
                return true;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < G; ++j)
            if ((st.alive >> j) & 1u) hb[j] = lds64(hbp + kidx + (uint32_t)j * kstep);
#pragma unroll
        for (int j = 0; j < G; ++j) st.g[j] = fx_finish(hp.c1, hb[j], hp.rot);
        probes += min(32u, cnt);
        dead += __reduce_add_sync(0xFFFFFFFFu, probe_phase<G, SMALL_M, 0, 1>(filt, hp, st));
        if (!exhaustive && dead > limit) {
            pass = false;
            return true;
        }
        if (G > 1 && cnt > 32u) {
            probes += cnt - 32u;
            dead += __reduce_add_sync(0xFFFFFFFFu, probe_phase<G, SMALL_M, (G > 1 ? 1 : 0), G>(filt, hp, st));
            if (!exhaustive && dead > limit) {
                pass = false;
                return true;
            }
        }
    }
    if (n_steps > 1u && cnt - dead != 0u && memo != nullptr) {
        // Exact node with a memo: a k-mer that survived step 0 and whose hash_bytes value is in the node's memo has
        // already passed all K probes here for another read of the batch (sequencing depth) -- it is a hit without
        // further probes.
        uint32_t found = 0;
        {
            unsigned long long e[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                e[j] = 0;
                if ((st.alive >> j) & 1u) e[j] = __ldcg(memo + ((uint32_t)(hb[j] >> 20) & memo_mask));
            }
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (e[j] == hb[j] && hb[j] != 0ULL && ((st.alive >> j) & 1u)) {
                    st.alive &= ~(1u << j);
                    ++found;
                }
        }
        const uint32_t known = __reduce_add_sync(0xFFFFFFFFu, found);
        memo_hits += known;
        memo_lookups += cnt - dead;
        // The k-mers still unknown are few and scattered over lanes and rounds; running the per-slot step loop for
        // them would issue G slots of index arithmetic per step for a handful of live lanes.  Instead their values are
        // compacted through shared memory, one k-mer per lane, and each lane probes its k-mer's remaining steps three
        // at a time (all three gathers in flight).  No per-k-mer identity is needed afterwards, only the counts.
        uint32_t n_left = 0;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const bool a = (st.alive >> j) & 1u;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, a);
            if (a) stage[n_left + __popc(b & ((1u << lane) - 1u))] = hb[j];
            n_left += __popc(b);
        }
        __syncwarp();
        uint32_t my_dead = 0;
        for (uint32_t t = lane; t < n_left; t += 32u) {
            const unsigned long long v = stage[t];
            const uint64_t h1 = fx_finish(hp.c1, v, hp.rot), h2 = fx_finish(hp.c2, v, hp.rot);
            uint64_t g = h2;  // g_1
            bool ok = true;
            for (uint32_t i = 1; i < n_steps && ok; i += 3u) {
                uint32_t w[3], bit[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    w[c] = 0xFFFFFFFFu;
                    bit[c] = 0;
                    if (i + c < n_steps) {
                        if (SMALL_M) {
                            const uint32_t idx = mod_small(g, (uint32_t)hp.M, (uint32_t)(hp.M >> 32), (uint32_t)hp.m);
                            w[c] = ldg32(filt + (idx >> 5));
                            bit[c] = idx & 31u;
                        } else {
                            const uint64_t idx = mod_any(g, hp.m, hp.M);
                            w[c] = ldg32(filt + (idx >> 5));
                            bit[c] = (uint32_t)idx & 31u;
                        }
                        my_probes += 1u;
                        g = (i + c == 1u) ? (h1 + 2ULL) * h2 : g + h2;  // g_2 = (h1+2)*h2, g_i = g_{i-1} + h2
                    }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) ok = ok && ((w[c] >> bit[c]) & 1u);
            }
            if (!ok) ++my_dead;
            else if (v != 0ULL) __stcg(memo + ((uint32_t)(v >> 20) & memo_mask), v);  // passed all K probes: remember
        }
        dead += __reduce_add_sync(0xFFFFFFFFu, my_dead);
        if (!exhaustive && dead > limit) {
            pass = false;
            return true;
        }
    } else if (n_steps > 1u && cnt - dead != 0u) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
            st.h1[j] = fx_finish(hp.c1, hb[j], hp.rot);
            st.h2[j] = fx_finish(hp.c2, hb[j], hp.rot);
        }
        for (uint32_t i = 1; i < n_steps; ++i) {
            const uint32_t n_alive = cnt - dead;
            if (n_alive == 0u) break;
            probes += n_alive;
#pragma unroll
            for (int j = 0; j < G; ++j)
                st.g[j] = i == 1 ? st.h2[j] : (i == 2 ? (st.h1[j] + 2ULL) * st.h2[j] : st.g[j] + st.h2[j]);
            dead += __reduce_add_sync(0xFFFFFFFFu, probe_phase<G, SMALL_M, 0, G>(filt, hp, st));
            if (!exhaustive && dead > limit) {
                pass = false;
                return true;
            }
        }
    }
    misses += dead;
    hits += cnt - dead;
    if (!exhaustive && hits >= need) {
        pass = true;
        return true;
    }
    return false;
}

struct PairMeta {
    uint32_t r, u, len, slot, steps, memo;
    uint64_t koff;
};

// Evaluate one (read,node) pair; warp-uniform control flow.  Returns pass/fail (query_passes).
template <int G, bool SMALL_M>
PF_D bool probe_pair(const ProbeArgs &a, const PairMeta &pm, uint32_t lane, int pre, unsigned long long *stage,
                     uint32_t &probes, uint32_t &my_probes, uint32_t &memo_hits, uint32_t &memo_lookups) {
    const HashParams &hp = a.hp;
    const uint32_t n_k = kmers_of(pm.len, hp.k);
    const uint32_t need = need_of(a.threshold, n_k);
    const bool exhaustive = a.exhaustive != 0;
    const uint32_t n_steps = pm.steps & 0xFFu, stride = pm.steps >> 8;  // node plan: steps | stride << 8
    if (!exhaustive) {
        if (n_steps == 0u) return true;  // skipped interior node: its children decide (verified superset)
        if (need == 0u) return true;     // hits >= 0 always
        if (need > n_k) return false;    // hits <= n_k < need
    }
    const uint32_t allowed = need > n_k ? 0u : n_k - need;
    // Sampled pre-test (verified-monotone interior nodes only): every stride-th k-mer, centred in the read.  A
    // k-mer that is not probed is not proven absent, so the test stays sound; it just prunes a little less.
    uint32_t n_s = n_k, off = 0;
    if (stride > 1u) {
        n_s = max(1u, n_k / stride);
        off = (n_k - 1u - (n_s - 1u) * stride) >> 1;
    }
    const uint32_t *filt = reinterpret_cast<const uint32_t *>(a.filters + (uint64_t)pm.slot * a.words_per_filter);
    const uint64_t *hbp = a.hb + (pm.koff - a.kmer_base);
    const uint32_t *i0p = a.idx0 + (pm.koff - a.kmer_base);
    uint32_t hits = n_k - n_s, misses = 0;
    bool pass = false;
    // node_memo: (first entry / 4096) << 5 | log2(entries) of the node's region
    unsigned long long *memo = (a.memo && pm.memo != NONE32_D) ? a.memo + ((size_t)(pm.memo >> 5) << 12) : nullptr;
    const uint32_t memo_mask = (1u << (pm.memo & 31u)) - 1u;
    for (uint32_t gbase = 0; gbase < n_s; gbase += 32u * G)
        if (probe_group<G, SMALL_M>(filt, hp, n_steps, hbp, i0p, gbase, n_s, lane, need, allowed, exhaustive, stride, off,
                                    gbase == 0u ? pre : -1, n_k, memo, memo_mask, stage, hits, misses, probes, my_probes,
                                    memo_hits, memo_lookups, pass))
            return pass;
    return hits >= need;
}

// Persistent grid; warps pull PROBE_GRAB = 32 consecutive pairs at a time from a global counter.  The pairs' records
// and per-read / per-node metadata are fetched by the 32 lanes in parallel (two dependent memory round trips per
// ticket instead of per pair) and broadcast with shuffles.
template <int G, bool SMALL_M>
static __global__ void __launch_bounds__(PROBE_THREADS, (G <= 2 ? 6 : (G <= 5 ? 4 : 3))) probe_kernel(const ProbeArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    // per-warp staging of the k-mer values the memo could not answer (probe_group)
    __shared__ unsigned long long stage_all[PROBE_THREADS / 32][32 * G];
    unsigned long long *const stage = stage_all[threadIdx.x >> 5];
    uint32_t *const np_mine = a.node_pass + (size_t)(blockIdx.x % NODE_PASS_COPIES) * a.n_nodes;
    uint32_t probes = 0;      // warp-uniform: probes of the per-pair path
    uint32_t my_probes = 0;   // per lane: probes of the pairs this lane settled after the batched first round
    uint32_t memo_hits = 0, memo_lookups = 0;  // warp-uniform: k-mers answered by / looked up in the memo
    unsigned long long probes_total = 0, memo_total = 0, lookup_total = 0;
    for (;;) {
        // One ticket = PROBE_GRAB consecutive pairs: every lane fetches the metadata of one pair (two dependent round
        // trips per 32 pairs), and the ticket counter -- a same-address atomic with a return value -- is hit once per
        // 32 pairs.  The pairs are then worked in sub-chunks of PROBE_CHUNK.
        uint32_t g0 = 0;
        if (lane == 0) g0 = atomicAdd(a.work_ctr, a.grab);
        g0 = __shfl_sync(0xFFFFFFFFu, g0, 0);
        if (a.order_streams > 1u) {
            // memo levels: consecutive tickets go to pairs far apart in the node-major frontier, so the pairs of one
            // node are spread over time instead of all being in flight at once (the memo can only answer what an
            // EARLIER pair of the node has stored); order_streams bounds how many nodes are then live in L2 together
            const uint32_t c = g0 / a.grab;
            if (c >= a.order_streams * a.order_span) break;
            g0 = ((c % a.order_streams) * a.order_span + c / a.order_streams) * a.grab;
            if (g0 >= a.n_pairs) continue;
        } else if (g0 >= a.n_pairs) break;
        const uint32_t n_grab = min(a.grab, a.n_pairs - g0);
        PairMeta mine{};
        if (lane < n_grab) {
            mine.r = lds32(a.fr_read + g0 + lane);
            mine.u = lds32(a.fr_node + g0 + lane);
            mine.len = ldg32(a.lengths + mine.r);
            mine.koff = __ldg(a.kmer_off + mine.r);
            mine.slot = ldg32(a.node_slot + mine.u);
            mine.steps = ldg32(a.node_steps + mine.u);
            mine.memo = a.memo ? ldg32(a.node_memo + mine.u) : NONE32_D;
        }
        bool my_pass = false;  // outcome of this lane's pair
        for (uint32_t sb = 0; sb < n_grab; sb += PROBE_CHUNK) {  // pairs [sb, sb + n_here) of the ticket: lanes sb + p
        const uint32_t n_here = min((uint32_t)PROBE_CHUNK, n_grab - sb);
        const bool holder = lane >= sb && lane < sb + n_here;  // this lane holds the metadata of a pair of the sub-chunk
        // Step 0 of the first round (up to 32 k-mers) of EVERY pair of the sub-chunk, batched: PROBE_CHUNK index loads,
        // then PROBE_CHUNK gathers in flight per lane, instead of two dependent round trips per pair.  That round
        // alone decides most pairs that fail and all of a sampled single-round pre-test; those pairs are settled
        // here by the lane that holds their metadata and never enter the per-pair path below.
        uint32_t undecided = (1u << n_here) - 1u;  // bit p: pair sb + p
        uint32_t miss_bits = 0;                    // bit p: this lane's k-mer of pair sb + p had its step-0 bit clear
        if (SMALL_M) {
            // per pair: first-round slot count | stride << 8, or 0 when the pair needs no probe at all; and the k-mer
            // index of slot 0
            uint32_t my_r0 = 0, my_nk = 0, my_need = 0, my_ns = 0, my_off = 0;
            uint64_t my_k0 = 0;
            bool my_decided = false;
            if (holder) {
                my_nk = kmers_of(mine.len, a.hp.k);
                my_need = need_of(a.threshold, my_nk);
                const uint32_t n_steps = mine.steps & 0xFFu, stride = mine.steps >> 8;
                if (!a.exhaustive && (n_steps == 0u || my_need == 0u || my_need > my_nk)) {  // as probe_pair
                    my_decided = true;
                    my_pass = n_steps == 0u || my_need == 0u;
                }
                uint32_t off = 0;
                my_ns = my_nk;
                if (stride > 1u) {
                    my_ns = max(1u, my_nk / stride);
                    off = (my_nk - 1u - (my_ns - 1u) * stride) >> 1;
                }
                if (!my_decided && my_nk) my_r0 = min(32u, my_ns) | (stride << 8);
                my_k0 = mine.koff - a.kmer_base;
                my_off = off;  // first sampled k-mer (< 8); n_k travels in its own word (reads may have >= 2^24 k-mers)
            }
            uint32_t idxv[PROBE_CHUNK], wv[PROBE_CHUNK];
#pragma unroll
            for (int p = 0; p < PROBE_CHUNK; ++p) {
                const uint32_t r0 = __shfl_sync(0xFFFFFFFFu, my_r0, sb + p);
                const uint64_t k0 = __shfl_sync(0xFFFFFFFFu, my_k0, sb + p);
                const uint32_t o4 = __shfl_sync(0xFFFFFFFFu, my_off, sb + p);
                const uint32_t nk4 = __shfl_sync(0xFFFFFFFFu, my_nk, sb + p);
                idxv[p] = 0xFFFFFFFFu;
                if (lane < (r0 & 0xFFu)) idxv[p] = lds32(a.idx0 + k0 + idx0_slot(o4 + lane * (r0 >> 8), nk4));
            }
#pragma unroll
            for (int p = 0; p < PROBE_CHUNK; ++p) {
                const uint32_t slot = __shfl_sync(0xFFFFFFFFu, mine.slot, sb + p);
                wv[p] = 0xFFFFFFFFu;
                if (idxv[p] != 0xFFFFFFFFu)
                    wv[p] = ldg32(reinterpret_cast<const uint32_t *>(a.filters + (uint64_t)slot * a.words_per_filter) + (idxv[p] >> 5));
            }
            uint32_t my_dead0 = 0;
#pragma unroll
            for (int p = 0; p < PROBE_CHUNK; ++p) {
                const bool miss = !((wv[p] >> (idxv[p] & 31u)) & 1u);
                if (miss) miss_bits |= 1u << p;
                const uint32_t b = __ballot_sync(0xFFFFFFFFu, miss);
                if (lane == sb + (uint32_t)p) my_dead0 = __popc(b);
            }
            if ((my_r0 & 0xFFu) != 0u) {  // this lane's pair was probed: can its first round settle it?
                const uint32_t cnt0 = my_r0 & 0xFFu, allowed = my_need > my_nk ? 0u : my_nk - my_need;
                if (!a.exhaustive && my_dead0 > allowed) {
                    my_decided = true;
                    my_pass = false;
                    my_probes += cnt0;
                } else if (my_ns <= 32u && (mine.steps & 0xFFu) == 1u) {
                    my_decided = true;
                    my_pass = (my_nk - my_ns) + (cnt0 - my_dead0) >= my_need;
                    my_probes += cnt0;
                }
            }
            undecided = (__ballot_sync(0xFFFFFFFFu, holder && !my_decided) >> sb) & ((1u << PROBE_CHUNK) - 1u);
        }
        uint32_t pass_bits = 0;
        for (uint32_t rest = undecided; rest; rest &= rest - 1u) {
            const uint32_t p = __ffs(rest) - 1u;
            PairMeta pm;
            pm.r = __shfl_sync(0xFFFFFFFFu, mine.r, sb + p);
            pm.u = __shfl_sync(0xFFFFFFFFu, mine.u, sb + p);
            pm.len = __shfl_sync(0xFFFFFFFFu, mine.len, sb + p);
            pm.slot = __shfl_sync(0xFFFFFFFFu, mine.slot, sb + p);
            pm.steps = __shfl_sync(0xFFFFFFFFu, mine.steps, sb + p);
            pm.koff = __shfl_sync(0xFFFFFFFFu, mine.koff, sb + p);
            pm.memo = __shfl_sync(0xFFFFFFFFu, mine.memo, sb + p);
            if (probe_pair<G, SMALL_M>(a, pm, lane, SMALL_M ? (int)((miss_bits >> p) & 1u) : -1, stage, probes, my_probes,
                                       memo_hits, memo_lookups))
                pass_bits |= 1u << p;
        }
        if (holder && ((undecided >> (lane - sb)) & 1u)) my_pass = ((pass_bits >> (lane - sb)) & 1u) != 0u;
        }  // sub-chunks
        // survivors per node: the frontier is node-major, so at any moment most warps of the GPU count into the same
        // node; same-address atomics serialise in L2 (measured: 1 M of them cost ~0.5 ms), so the counter is kept in
        // NODE_PASS_COPIES copies, one per group of CTAs, summed by level_scan_kernel
        const bool pass = lane < n_grab && my_pass;
        if (lane < n_grab) __stcs(a.pass + g0 + lane, (uint8_t)(pass ? 1 : 0));
        if (pass) atomicAdd(np_mine + mine.u, 1u);
        probes_total += probes + __reduce_add_sync(0xFFFFFFFFu, my_probes);
        memo_total += memo_hits;
        lookup_total += memo_lookups;
        probes = 0;
        my_probes = 0;
        memo_hits = memo_lookups = 0;
    }
    if (lane == 0 && probes_total) atomicAdd(a.probes, probes_total);
    if (lane == 0 && memo_total) atomicAdd(a.probes + 1, memo_total);
    if (lane == 0 && lookup_total) atomicAdd(a.probes + 2, lookup_total);
}

// ---- frontier bookkeeping ------------------------------------------------------------------
// Appends (read, entry node) pairs for every read of the chunk and every entry node of this level, node-major.
// Entry nodes are the first nodes below the root that are actually tested (the root itself unless the plan
// skips it): the skipped region above them never enters the frontier.
static __global__ void inject_frontier_kernel(uint32_t *fr_read, uint32_t *fr_node, uint32_t read0, uint32_t n_reads,
                                       const uint32_t *__restrict__ entry, uint32_t n_entry) {
    const uint64_t total = (uint64_t)n_reads * n_entry;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t e = (uint32_t)(i / n_reads), r = (uint32_t)(i - (uint64_t)e * n_reads);
        fr_read[i] = read0 + r;
        fr_node[i] = entry[e];
    }
}

struct LevelTotals {
    unsigned long long memo_lookups;  // k-mers looked up in the memo (one 8-byte read each)
    unsigned long long next_pairs;
    unsigned long long hits_total;  // running total of (read,leaf) hits in this block
    unsigned long long probes;
    unsigned long long memo_hits;  // k-mers answered by the memo instead of K - 1 probes
};

// One block.  For the nodes [lo,hi) of the current level: exclusive scan of the surviving pair counts
// gives every child its slice of the next frontier (children are numbered level-order, left before
// right, so the slices are node-major) and every leaf its slice of the hit list; leaf counts are added
// to the block histogram.
static __global__ void level_scan_kernel(uint32_t lo, uint32_t hi, const uint32_t *__restrict__ node_pass_copies,
                                  uint32_t n_nodes, uint32_t *node_pass,
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
        uint32_t tot = 0;
#pragma unroll
        for (int cpy = 0; cpy < NODE_PASS_COPIES; ++cpy) tot += node_pass_copies[(size_t)cpy * n_nodes + u];
        node_pass[u] = tot;  // the survivors of node u: read again below and by scatter_kernel
        const unsigned long long c = tot;
        if (leaf[u] >= 0) sh += c;
        else sn += c * ((left[u] != NONE32_D) + (right[u] != NONE32_D));
    }
    const unsigned long long hits_before = totals->hits_total;  // every thread reads it before the first barrier;
                                                                // it is rewritten only after that barrier
    // exclusive scan of (sn, sh) over the block's threads: warp scans, then a scan of the 32 warp totals
    const uint32_t lane = t & 31u, wid = t >> 5;
    unsigned long long in_n = sn, in_h = sh;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long yn = __shfl_up_sync(0xFFFFFFFFu, in_n, o), yh = __shfl_up_sync(0xFFFFFFFFu, in_h, o);
        if (lane >= (uint32_t)o) {
            in_n += yn;
            in_h += yh;
        }
    }
    if (lane == 31u) {
        s_next[wid] = in_n;
        s_hit[wid] = in_h;
    }
    __syncthreads();
    if (wid == 0) {
        const unsigned long long xn = lane < (nt >> 5) ? s_next[lane] : 0ULL, xh = lane < (nt >> 5) ? s_hit[lane] : 0ULL;
        unsigned long long zn = xn, zh = xh;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long yn = __shfl_up_sync(0xFFFFFFFFu, zn, o), yh = __shfl_up_sync(0xFFFFFFFFu, zh, o);
            if (lane >= (uint32_t)o) {
                zn += yn;
                zh += yh;
            }
        }
        s_next[lane] = zn - xn;  // exclusive warp bases
        s_hit[lane] = zh - xh;
        if (lane == 31u) {
            totals->next_pairs = zn;
            totals->hits_total = hits_before + zh;
            totals->probes = probes[0];
            totals->memo_hits = probes[1];
            totals->memo_lookups = probes[2];
        }
    }
    __syncthreads();
    unsigned long long an = s_next[wid] + in_n - sn, ah = hits_before + s_hit[wid] + in_h - sh;
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

// Thread per pair: survivors take a rank inside their node and are written to both children's slices, or to the hit
// list when the node is a leaf.  The frontier is node-major, so a block's pairs mostly belong to the node of its
// first pair: those survivors are ranked with ONE atomic per block (per-warp counts in shared memory); pairs of the
// block's other nodes use a warp-aggregated atomic (`__match_any_sync`).  (Consecutive blocks hit the same cursor,
// and same-address atomics with a return value serialise in L2.)
static __global__ void scatter_kernel(const uint32_t *__restrict__ fr_read, const uint32_t *__restrict__ fr_node,
                               const uint8_t *__restrict__ pass, uint32_t n, const uint32_t *__restrict__ node_pass,
                               uint32_t *cursor, const uint32_t *__restrict__ left, const uint32_t *__restrict__ right,
                               const int32_t *__restrict__ leaf, const unsigned long long *__restrict__ next_base,
                               const unsigned long long *__restrict__ hit_base, uint32_t *nx_read, uint32_t *nx_node,
                               uint32_t *hit_read, uint32_t *hit_leaf, uint32_t *read_hits, int want_hits) {
    __shared__ uint32_t s_cnt[32], s_base;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const bool on = i < n && __ldcs(pass + i);
    const uint32_t u = on ? lds32(fr_node + i) : NONE32_D, r = on ? lds32(fr_read + i) : 0u;
    const uint32_t u0 = fr_node[blockIdx.x * blockDim.x];  // the block's first pair exists: grid = ceil(n / block)
    const bool is0 = on && u == u0;
    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, is0);
    if (lane == 0) s_cnt[wid] = __popc(b0);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (uint32_t w = 0; w < n_warps; ++w) {
            const uint32_t c = s_cnt[w];
            s_cnt[w] = tot;
            tot += c;
        }
        s_base = tot ? atomicAdd(cursor + u0, tot) : 0u;
    }
    __syncthreads();
    uint32_t rank = 0;
    if (is0) rank = s_base + s_cnt[wid] + __popc(b0 & ((1u << lane) - 1u));
    const bool other = on && !is0;
    const uint32_t act = __ballot_sync(0xFFFFFFFFu, other);
    if (other) {
        const uint32_t peers = __match_any_sync(act, u);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(cursor + u, (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        rank = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (!on) return;
    const int32_t lf = leaf[u];
    if (lf >= 0) {
        if (want_hits) {
            const unsigned long long p = hit_base[u] + rank;
            hit_read[p] = r;
            hit_leaf[p] = (uint32_t)lf;
            // per-read hit count for the CSR built at the end of the block; mode 2 (subtree shards: the hit is
            // still to be routed to the rank that owns the read) leaves the counting to the receiver
            if (want_hits == 1) atomicAdd(read_hits + r, 1u);
        }
    } else {
        unsigned long long p = next_base[u] + rank;
        const uint32_t c = node_pass[u];
        const uint32_t l = left[u], rr = right[u];
        if (l != NONE32_D) {
            __stcs(nx_read + p, r);
            __stcs(nx_node + p, l);
            p += c;
        }
        if (rr != NONE32_D) {
            __stcs(nx_read + p, r);
            __stcs(nx_node + p, rr);
        }
    }
}

// ---- load-time analysis of the tree --------------------------------------------------------------
// pop[slot] += popcount(filter[slot]); grid = (blocks, n_slots_in_chunk)
static __global__ void fill_kernel(const uint64_t *__restrict__ filters, uint64_t wpf, uint32_t slot0, unsigned long long *pop) {
    const uint32_t slot = slot0 + blockIdx.y;
    const uint64_t *f = filters + (uint64_t)slot * wpf;
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < wpf; i += (uint64_t)gridDim.x * blockDim.x)
        c += __popcll(f[i]);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(pop + slot, c);
}
// viol[u] += popcount(filter(child) & ~filter(u)) over both children; grid = (blocks, nodes_in_chunk)
static __global__ void subset_kernel(const uint64_t *__restrict__ filters, uint64_t wpf, const uint32_t *__restrict__ slot,
                              const uint32_t *__restrict__ left, const uint32_t *__restrict__ right, uint32_t node0,
                              unsigned long long *viol) {
    const uint32_t u = node0 + blockIdx.y;
    const uint32_t l = left[u], r = right[u];
    if (l == NONE32_D && r == NONE32_D) return;
    const uint64_t *fu = filters + (uint64_t)slot[u] * wpf;
    const uint64_t *fl = l != NONE32_D ? filters + (uint64_t)slot[l] * wpf : nullptr;
    const uint64_t *fr = r != NONE32_D ? filters + (uint64_t)slot[r] * wpf : nullptr;
    unsigned long long c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < wpf; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = ~fu[i];
        if (fl) c += __popcll(fl[i] & p);
        if (fr) c += __popcll(fr[i] & p);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(viol + u, c);
}

// k-mer count per read (file_parser.rs:136-139); its exclusive scan (csr_* kernels below) is kmer_off
static __global__ void kmer_counts_kernel(const uint32_t *__restrict__ lengths, uint32_t n, uint32_t k, uint32_t *cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = kmers_of(lengths[i], k);
}

// ---- per-read hit lists (ResultMap, result_map.rs:9-46) as CSR, built on the device ------------------
// exclusive scan of cnt[0..n) into off[0..n] (u64), 1024 elements per block: block sums, scan of the sums,
// then the in-block scan with the block's base.
static __global__ void csr_block_sums_kernel(const uint32_t *__restrict__ cnt, uint32_t n, unsigned long long *bsum) {
    __shared__ unsigned long long s[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
    unsigned long long v = i < n ? cnt[i] : 0;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = s[threadIdx.x];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (threadIdx.x == 0) bsum[blockIdx.x] = v;
    }
}
static __global__ void csr_scan_sums_kernel(unsigned long long *bsum, uint32_t nb) {  // one block, in place, exclusive
    __shared__ unsigned long long s[1024];
    const uint32_t t = threadIdx.x, per = (nb + 1023u) / 1024u;
    const uint32_t b = min(nb, t * per), e = min(nb, (t + 1) * per);
    unsigned long long sum = 0;
    for (uint32_t i = b; i < e; ++i) sum += bsum[i];
    s[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long a = 0;
        for (uint32_t i = 0; i < 1024; ++i) {
            unsigned long long x = s[i];
            s[i] = a;
            a += x;
        }
    }
    __syncthreads();
    unsigned long long a = s[t];
    for (uint32_t i = b; i < e; ++i) {
        unsigned long long x = bsum[i];
        bsum[i] = a;
        a += x;
    }
}
static __global__ void csr_offsets_kernel(const uint32_t *__restrict__ cnt, uint32_t n, const unsigned long long *__restrict__ bsum,
                                   unsigned long long *off) {
    __shared__ unsigned long long s[32];
    const uint32_t i = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const unsigned long long c = i < n ? cnt[i] : 0;
    unsigned long long v = c;
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= (uint32_t)o) v += y;
    }
    if (lane == 31) s[w] = v;
    __syncthreads();
    if (w == 0) {
        unsigned long long x = s[lane], z = x;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, z, o);
            if (lane >= (uint32_t)o) z += y;
        }
        s[lane] = z - x;  // exclusive warp bases
    }
    __syncthreads();
    const unsigned long long excl = bsum[blockIdx.x] + s[w] + v - c;
    if (i < n) off[i] = excl;
    if (i == n - 1) off[n] = excl + c;
}
// place every hit in its read's segment (any order), then sort each segment ascending by DFS leaf index
static __global__ void csr_fill_kernel(const uint32_t *__restrict__ hit_read, const uint32_t *__restrict__ hit_leaf,
                                unsigned long long n_hits, const unsigned long long *__restrict__ off, uint32_t *cnt,
                                uint32_t *out_leaf) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    const uint32_t r = hit_read[i];
    const uint32_t k = atomicSub(cnt + r, 1u) - 1u;
    out_leaf[off[r] + k] = hit_leaf[i];
}
static __global__ void csr_sort_kernel(const unsigned long long *__restrict__ off, uint32_t n, uint32_t *out_leaf) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const unsigned long long b = off[r], e = off[r + 1];
    for (unsigned long long i = b + 1; i < e; ++i) {
        const uint32_t x = out_leaf[i];
        unsigned long long j = i;
        while (j > b && out_leaf[j - 1] > x) {
            out_leaf[j] = out_leaf[j - 1];
            --j;
        }
        out_leaf[j] = x;
    }
}

// flag[r] = 1 for every read that owns a pair of the list
static __global__ void flag_reads_kernel(const uint32_t *__restrict__ reads, unsigned long long n, uint8_t *flag) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        flag[reads[i]] = 1;
}

// flag[r] = 1 for the exception reads (bytes other than upper-case ACGT) among reads [r0, r0 + n), 0 for the others
static __global__ void flag_exc_kernel(const uint32_t *__restrict__ exc_index, uint32_t r0, uint32_t n, uint8_t *flag) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flag[r0 + i] = exc_index[r0 + i] != NONE32_D;
}

static __global__ void zero_kernel(uint4 *p, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}

static __global__ void add_counts_kernel(unsigned long long *dst, const unsigned long long *src, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

}  // namespace pf
