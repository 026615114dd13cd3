// pf_hash.cuh -- the reference's k-mer hashing, bit for bit, for device (and host) code.
//
// Reference chain (paths relative to the reference root):
//   file_parser.rs:114-148   canonical k-mer = bytewise min(kmer, revcomp(kmer)), forward on ties
//   hasher.rs:12-21          HashSeed::build_hasher: FxHasher::default(); write_usize(seed)
//   hash_iter.rs:31-45       h1 = hash_one(item) with seed one, h2 with seed two
//   hash_iter.rs:13-27       g0 = h1, g1 = h2, gi = (h1 + i) * h2 (wrapping)
//   bloom_filter.rs:312-332  idx = g % bits.len(); bit idx of BitVec<usize, Lsb0>
// rustc-hash 2.x FxHasher (crate not vendored with the reference): add_to_hash(x): h = (h + x) * K;
// write(bytes) = add_to_hash(hash_bytes(bytes)); <[u8] as Hash> writes a usize length prefix first;
// finish() = rotate_left(h, ROT) with ROT = 26 (2.1.1) or 20 (earlier 2.x) -- a runtime parameter.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define PF_HD __host__ __device__ __forceinline__
#define PF_D __device__ __forceinline__

namespace pf {

constexpr uint64_t FX_K = 0xf1357aea2e62a9c5ULL;
constexpr uint64_t FX_SEED1 = 0x243f6a8885a308d3ULL;
constexpr uint64_t FX_SEED2 = 0x13198a2e03707344ULL;
constexpr uint64_t FX_PREVENT = 0xa4093822299f31d0ULL;

// Everything the kernels need about one database's hashing and filter geometry.
struct HashParams {
    uint64_t c1, c2;   // (seed_j * K + k) * K: hasher state after write_usize(seed), write_usize(len)
    uint64_t m;        // bits per filter
    uint64_t M;        // Barrett reciprocal floor(2^64 / m)
    uint32_t K;        // probes per k-mer (num_hashes)
    uint32_t k;        // k-mer size
    uint32_t rot;      // finish() rotate
    uint32_t small_m;  // 1 if m < 2^31 (32-bit remainder path)
};

PF_HD uint64_t mulmix(uint64_t x, uint64_t y) {
#ifdef __CUDA_ARCH__
    return (x * y) ^ __umul64hi(x, y);
#else
    __uint128_t p = (__uint128_t)x * (__uint128_t)y;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
#endif
}
PF_HD uint64_t rotl64(uint64_t x, uint32_t r) {
    r &= 63u;
    return r ? (x << r) | (x >> (64u - r)) : x;
}
// h_j from the seed-independent hash_bytes value: state after the two usize writes is c_j.
PF_HD uint64_t fx_finish(uint64_t c, uint64_t hb, uint32_t rot) { return rotl64((c + hb) * FX_K, rot); }

inline HashParams make_hash_params(uint64_t seed1, uint64_t seed2, uint64_t k, uint64_t m, uint32_t K, int rot) {
    HashParams p{};
    p.c1 = (seed1 * FX_K + k) * FX_K;
    p.c2 = (seed2 * FX_K + k) * FX_K;
    p.m = m;
    // floor(2^64 / m); with this M, q = hi64(g*M) is floor(g/m) or one less, so one conditional
    // subtraction finishes the remainder.  (m == 1: M = 2^64-1 also satisfies that bound.)
    bool pow2 = m > 1 && (m & (m - 1)) == 0;
    p.M = m ? (~0ULL / m) + (pow2 ? 1ULL : 0ULL) : 0;
    p.K = K;
    p.k = (uint32_t)k;
    p.rot = (uint32_t)rot;
    p.small_m = m < (1ULL << 31) ? 1u : 0u;
    return p;
}

// g mod m for m < 2^31: only the low 32 bits of q = hi64(g*M) are needed because the remainder
// before the final correction is < 2m < 2^32.
PF_D uint32_t mod_small(uint64_t g, uint32_t M0, uint32_t M1, uint32_t m32) {
    uint32_t g0 = (uint32_t)g, g1 = (uint32_t)(g >> 32);
    uint64_t p = (uint64_t)g0 * M0;
    uint64_t S = (uint64_t)g1 * M0 + (p >> 32);
    S += (uint64_t)g0 * M1;  // wraps mod 2^64: bits [32,64) are still exact
    uint32_t q = g1 * M1 + (uint32_t)(S >> 32);
    uint32_t r = g0 - q * m32;
    return min(r, r - m32);
}
PF_HD uint64_t mod_any(uint64_t g, uint64_t m, uint64_t M) {
#ifdef __CUDA_ARCH__
    uint64_t q = __umul64hi(g, M);
#else
    uint64_t q = (uint64_t)(((__uint128_t)g * M) >> 64);
#endif
    uint64_t r = g - q * m;
    return r >= m ? r - m : r;
}

// ---- byte-exact generic path -------------------------------------------------------------
// bio::alphabets::dna::complement: identity except the IUPAC pairs (and their lower-case forms).
PF_HD uint8_t complement(uint8_t b) {
    uint8_t u = b & 0xDFu;  // upper-case form if b is a letter
    bool letter = (u >= 'A' && u <= 'Z') && (b == u || b == (uint8_t)(u | 0x20u));
    if (!letter) return b;
    uint8_t c;
    switch (u) {
        case 'A': c = 'T'; break;
        case 'T': c = 'A'; break;
        case 'G': c = 'C'; break;
        case 'C': c = 'G'; break;
        case 'Y': c = 'R'; break;
        case 'R': c = 'Y'; break;
        case 'K': c = 'M'; break;
        case 'M': c = 'K'; break;
        case 'D': c = 'H'; break;
        case 'H': c = 'D'; break;
        case 'V': c = 'B'; break;
        case 'B': c = 'V'; break;
        default: return b;  // W, S, N map to themselves; other letters untouched
    }
    return (uint8_t)(c | (b & 0x20u));
}

// hash_bytes over an arbitrary byte accessor b(j), j in [0,n).
template <class ByteFn>
PF_HD uint64_t hash_bytes_fn(ByteFn b, uint32_t n) {
    auto le64 = [&](uint32_t o) {
        uint64_t v = 0;
        for (int t = 0; t < 8; ++t) v |= (uint64_t)b(o + t) << (8 * t);
        return v;
    };
    auto le32 = [&](uint32_t o) {
        uint64_t v = 0;
        for (int t = 0; t < 4; ++t) v |= (uint64_t)b(o + t) << (8 * t);
        return v;
    };
    uint64_t s0 = FX_SEED1, s1 = FX_SEED2;
    if (n <= 16) {
        if (n >= 8) {
            s0 ^= le64(0);
            s1 ^= le64(n - 8);
        } else if (n >= 4) {
            s0 ^= le32(0);
            s1 ^= le32(n - 4);
        } else if (n > 0) {
            s0 ^= (uint64_t)b(0);
            s1 ^= ((uint64_t)b(n - 1) << 8) | (uint64_t)b(n / 2);
        }
    } else {
        uint32_t off = 0;
        while (off < n - 16) {
            uint64_t t = mulmix(s0 ^ le64(off), FX_PREVENT ^ le64(off + 8));
            s0 = s1;
            s1 = t;
            off += 16;
        }
        s0 ^= le64(n - 16);
        s1 ^= le64(n - 8);
    }
    return mulmix(s0, s1) ^ (uint64_t)n;
}

// Canonical k-mer hash_bytes for a k-mer given by a forward byte accessor f(j), j in [0,k).
template <class ByteFn>
PF_HD uint64_t canonical_hash_bytes(ByteFn f, uint32_t k) {
    int c = 0;
    for (uint32_t i = 0; i < k && c == 0; ++i) {
        uint8_t a = f(i), r = complement(f(k - 1 - i));
        c = a < r ? -1 : (a > r ? 1 : 0);
    }
    if (c <= 0) return hash_bytes_fn(f, k);
    return hash_bytes_fn([&](uint32_t j) { return complement(f(k - 1 - j)); }, k);
}

// ---- 2-bit register path (pure upper-case ACGT, 17 <= k <= 32) ---------------------------
// Codes A=0,C=1,G=2,T=3, base j at bits [2j,2j+2).  ASCII order A<C<G<T equals code order, so
// the bytewise comparison of the reference is an integer comparison on packed codes.

// reverse the order of the 32 two-bit groups of x
PF_D uint64_t rev2(uint64_t x) {
    uint64_t y = __brevll(x);
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
}
// 8 two-bit codes (low 16 bits of x) -> 8 ASCII bytes, base 0 in the least significant byte
PF_D uint64_t expand8(uint32_t x) {
    uint32_t t = x & 0xFFFFu;
    t = (t | (t << 8)) & 0x00FF00FFu;
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    t = (t | (t << 2)) & 0x33333333u;  // nibble i = code i: a PRMT selector
    uint32_t lo = __byte_perm(0x54474341u, 0u, t);        // "ACGT" little-endian
    uint32_t hi = __byte_perm(0x54474341u, 0u, t >> 16);
    return ((uint64_t)hi << 32) | lo;
}

// x holds >= 2*KM valid low bits: the forward k-mer.  Returns hash_bytes(canonical k-mer bytes).
template <int KM>
PF_D uint64_t canonical_hash_2bit(uint64_t x) {
    static_assert(KM >= 17 && KM <= 32, "2-bit path covers 17 <= k <= 32");
    constexpr uint64_t mask = KM == 32 ? ~0ULL : ((1ULL << (2 * (KM & 31))) - 1ULL);
    uint64_t F = x & mask;
    uint64_t R = rev2(~x) >> (64 - 2 * KM);
    // key(F) = mask & ~R and key(R) = mask & ~F (first base most significant), so
    // forward < revcomp lexicographically  <=>  F < R as integers; ties pick the same bytes.
    uint64_t C = F < R ? F : R;
    uint32_t c0 = (uint32_t)C, c1 = (uint32_t)(C >> 32);
    uint64_t W[4];
    W[0] = expand8(c0);
    W[1] = expand8(c0 >> 16);
    W[2] = expand8(c1);
    W[3] = KM > 24 ? expand8(c1 >> 16) : 0ULL;
    auto bytes_at = [&](int o) -> uint64_t {  // le64 of bytes [o, o+8); o is a compile-time constant
        int q = o >> 3, r = o & 7;
        if (r == 0) return W[q];
        return (W[q] >> (8 * r)) | (W[q + 1] << (64 - 8 * r));
    };
    // hash_bytes, len in 17..32: one bulk iteration (off = 0), then the 16-byte suffix
    uint64_t t = mulmix(FX_SEED1 ^ W[0], FX_PREVENT ^ W[1]);
    uint64_t s0 = FX_SEED2 ^ bytes_at(KM - 16);
    uint64_t s1 = t ^ bytes_at(KM - 8);
    return mulmix(s0, s1) ^ (uint64_t)KM;
}

// The same with k (17..32) known only at run time: for kernels that hash on the fly next to other work and cannot afford
// one instantiation per k (the entry line kernel of the sliced path).  A few selects and variable shifts more.
PF_D uint64_t canonical_hash_2bit_rt(uint64_t x, uint32_t k) {
    const uint64_t mask = k >= 32u ? ~0ULL : ((1ULL << (2u * k)) - 1ULL);
    const uint64_t F = x & mask;
    const uint64_t R = rev2(~x) >> (64u - 2u * k);
    const uint64_t C = F < R ? F : R;
    const uint32_t c0 = (uint32_t)C, c1 = (uint32_t)(C >> 32);
    const uint64_t W0 = expand8(c0), W1 = expand8(c0 >> 16), W2 = expand8(c1), W3 = expand8(c1 >> 16);
    auto bytes_at = [&](uint32_t o) -> uint64_t {  // le64 of bytes [o, o + 8), 1 <= o <= 24; o + 8 <= k, so W3 is only read for k > 24
        const uint32_t q = o >> 3, sh = 8u * (o & 7u);
        const uint64_t a = q == 0u ? W0 : (q == 1u ? W1 : (q == 2u ? W2 : W3));
        const uint64_t b = q == 0u ? W1 : (q == 1u ? W2 : W3);
        return sh ? (a >> sh) | (b << (64u - sh)) : a;
    };
    const uint64_t t = mulmix(FX_SEED1 ^ W0, FX_PREVENT ^ W1);
    const uint64_t s0 = FX_SEED2 ^ bytes_at(k - 16u);
    const uint64_t s1 = t ^ bytes_at(k - 8u);
    return mulmix(s0, s1) ^ (uint64_t)k;
}

}  // namespace pf
