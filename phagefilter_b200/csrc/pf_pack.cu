// pf_pack.cu -- host-side packing of raw reads into the 2-bit batch layout of pf_read_batch.
//
// The reference materialises one Vec<u8> per canonical k-mer (file_parser.rs:135-148); here a read
// travels as 2 bits per base and k-mers are re-created on the GPU.  Reads holding any byte other than
// upper-case A/C/G/T are hashed verbatim by the reference, so they are carried as raw bytes too
// (exception side channel) and take the byte-exact path on the device.
//
// One pass over the bases on several host threads (table lookup, 16 bases per word); the pinned buffers of
// a previous batch can be recycled, because page-locking memory costs more than packing it.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "pf_common.h"

namespace pf {  // pf_pack_simd.cpp
bool cpu_has_avx2();
bool pack_groups_avx2(const uint8_t *s, uint64_t n32, uint32_t *dst);
}  // namespace pf

namespace {

struct HostBuf {  // pinned when a CUDA device is present, pageable otherwise (packing is host logic)
    void *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        release();
        const size_t want = bytes + bytes / 4 + 64;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
            if (cudaMallocHost(&p, want) == cudaSuccess) pinned = true;
            else {
                cudaGetLastError();
                p = nullptr;
            }
        } else {
            cudaGetLastError();
        }
        if (!p) {
            p = aligned_alloc(64, (want + 63) / 64 * 64);
            pinned = false;
        }
        cap = p ? want : 0;
        return p != nullptr;
    }
    void release() {
        if (p) {
            if (pinned) cudaFreeHost(p);
            else free(p);
        }
        p = nullptr;
        cap = 0;
    }
};

struct Lut {
    uint8_t code[256];
    Lut() {
        memset(code, 0xFF, sizeof code);
        code[(uint8_t)'A'] = 0;
        code[(uint8_t)'C'] = 1;
        code[(uint8_t)'G'] = 2;
        code[(uint8_t)'T'] = 3;
    }
};
const Lut kLut;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline uint64_t words_of(uint64_t len) { return ((len + 15) / 16 + 1) & ~1ULL; }  // even: 8-byte aligned reads

// packs one read into nw zero-padded words; returns false (all zeros) if it holds a byte outside {A,C,G,T}
inline bool pack_read(const uint8_t *s, uint64_t len, uint32_t *dst, uint64_t nw, bool avx2) {
    const uint64_t full = len / 16;
    uint32_t bad = 0;
    uint64_t w0 = 0;
    if (avx2 && len >= 32) {  // 32 bases (two words) per step; the table loop below finishes the read
        const uint64_t n32 = len / 32;
        if (!pf::pack_groups_avx2(s, n32, dst)) bad = 0x80u;
        w0 = 2 * n32;
    }
    for (uint64_t w = w0; w < full; ++w) {
        const uint8_t *q = s + w * 16;
        uint32_t v = 0;
#pragma GCC unroll 16
        for (int j = 0; j < 16; ++j) {
            const uint32_t c = kLut.code[q[j]];
            bad |= c;
            v |= (c & 3u) << (2 * j);
        }
        dst[w] = v;
    }
    uint64_t used = full;
    if (len % 16) {
        uint32_t v = 0;
        for (uint64_t j = 0; j < len % 16; ++j) {
            const uint32_t c = kLut.code[s[full * 16 + j]];
            bad |= c;
            v |= (c & 3u) << (2 * j);
        }
        dst[used++] = v;
    }
    for (; used < nw; ++used) dst[used] = 0;
    if (bad & 0x80u) {
        memset(dst, 0, nw * 4);
        return false;
    }
    return true;
}

}  // namespace

struct pf_packed {
    pf_read_batch b{};
    HostBuf main, exc;
};

extern "C" {

void *pf_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        pf::set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void pf_free_pinned(void *p) {
    if (p) cudaFreeHost(p);
}
int pf_thread_set_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return PF_OK;  // no device: the packer uses pageable memory (host logic only)
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        pf::set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
        return PF_ERR_CUDA;
    }
    return PF_OK;
}

// seq(r) / len(r) accessors: contiguous (seqs + offs) or scattered (one pointer per read)
struct ReadSrcView {
    const uint8_t *seqs;
    const uint64_t *offs;
    const uint8_t *const *ptrs;
    const uint32_t *lens;
    const uint8_t *seq(uint32_t r) const { return ptrs ? ptrs[r] : seqs + offs[r]; }
    uint64_t len(uint32_t r) const { return ptrs ? lens[r] : offs[r + 1] - offs[r]; }
};
static int pack_impl(const ReadSrcView &src, uint32_t n_reads, pf_packed **out);

int pf_pack_reads(const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, pf_packed **out) {
    if (!out || !offs || (!seqs && n_reads && offs[n_reads] > offs[0])) {
        pf::set_error("pf_pack_reads: null argument");
        return PF_ERR_ARG;
    }
    return pack_impl(ReadSrcView{seqs, offs, nullptr, nullptr}, n_reads, out);
}

int pf_pack_reads_ptrs(const uint8_t *const *seq_ptrs, const uint32_t *lengths, uint32_t n_reads, pf_packed **out) {
    if (!out || (n_reads && (!seq_ptrs || !lengths))) {
        pf::set_error("pf_pack_reads_ptrs: null argument");
        return PF_ERR_ARG;
    }
    return pack_impl(ReadSrcView{nullptr, nullptr, seq_ptrs, lengths}, n_reads, out);
}

static int pack_impl(const ReadSrcView &src, uint32_t n_reads, pf_packed **out) {
    pf_packed *p = *out ? *out : new pf_packed();
    auto fail = [&](int rc) {
        if (!*out) {
            p->main.release();
            p->exc.release();
            delete p;
        }
        return rc;
    };
    unsigned n_thr = std::thread::hardware_concurrency();
    if (const char *e = getenv("PF_PACK_THREADS")) n_thr = (unsigned)atoi(e);
    n_thr = std::max(1u, std::min(n_thr, 32u));
    // small batches are not worth the thread start-up (estimate: first read's length x reads; ranges are by read count)
    if (n_reads < 64 || (uint64_t)src.len(0) * n_reads < (1u << 20)) n_thr = 1;
    auto parallel = [&](const std::function<void(unsigned)> &fn) {
        if (n_thr == 1) return fn(0);
        std::vector<std::thread> th;
        for (unsigned t = 1; t < n_thr; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto &x : th) x.join();
    };
    auto range = [&](unsigned t, uint32_t &r0, uint32_t &r1) {
        r0 = (uint32_t)((uint64_t)n_reads * t / n_thr);
        r1 = (uint32_t)((uint64_t)n_reads * (t + 1) / n_thr);
    };
    // sizes from the lengths alone: per-thread sums over contiguous ranges of reads, then a prefix over the threads
    struct Part {
        uint64_t bases = 0, words = 0;
        uint32_t max_len = 0, too_long = pf::NONE32;
    };
    std::vector<Part> part(n_thr);
    parallel([&](unsigned t) {
        uint32_t r0, r1;
        range(t, r0, r1);
        Part q;
        for (uint32_t r = r0; r < r1; ++r) {
            const uint64_t len = src.len(r);
            if (len > 0xFFFFFFFFull) {
                if (q.too_long == pf::NONE32) q.too_long = r;
                continue;
            }
            q.bases += len;
            q.max_len = std::max<uint32_t>(q.max_len, (uint32_t)len);
            q.words += words_of(len);
        }
        part[t] = q;
    });
    uint64_t total_bases = 0, n_words = 0;
    uint32_t max_length = 0;
    std::vector<uint64_t> first_word(n_thr);
    for (unsigned t = 0; t < n_thr; ++t) {
        if (part[t].too_long != pf::NONE32) {
            pf::set_error("read %u longer than 2^32-1 bases", part[t].too_long);
            return fail(PF_ERR_ARG);
        }
        first_word[t] = n_words;
        total_bases += part[t].bases;
        max_length = std::max(max_length, part[t].max_len);
        n_words += part[t].words;
    }
    const uint64_t pad_words = 4;
    const size_t o_len = 0;
    const size_t o_woff = align_up(o_len + (size_t)n_reads * 4, 16);
    const size_t o_packed = align_up(o_woff + (size_t)n_reads * 8, 16);
    const size_t o_excidx = align_up(o_packed + (size_t)(n_words + pad_words) * 4, 16);
    const size_t total = o_excidx + (size_t)n_reads * 4 + 16;
    if (!p->main.ensure(total)) {
        pf::set_error("pf_pack_reads: out of host memory (%zu bytes)", total);
        return fail(PF_ERR_NOMEM);
    }
    uint8_t *base = static_cast<uint8_t *>(p->main.p);
    uint32_t *lengths = reinterpret_cast<uint32_t *>(base + o_len);
    uint64_t *word_off = reinterpret_cast<uint64_t *>(base + o_woff);
    uint32_t *packed = reinterpret_cast<uint32_t *>(base + o_packed);
    uint32_t *exc_index = reinterpret_cast<uint32_t *>(base + o_excidx);
    memset(packed + n_words, 0, pad_words * 4);
    // one pass over the bases, in parallel over the same ranges: lengths, word offsets and the 2-bit codes
    const bool avx2 = pf::cpu_has_avx2() && !getenv("PF_PACK_SCALAR");
    std::vector<std::vector<uint32_t>> exc_lists(n_thr);
    parallel([&](unsigned t) {
        uint32_t r0, r1;
        range(t, r0, r1);
        uint64_t w = first_word[t];
        for (uint32_t r = r0; r < r1; ++r) {
            const uint64_t len = src.len(r), nw = words_of(len);
            lengths[r] = (uint32_t)len;
            word_off[r] = w;
            exc_index[r] = pf::NONE32;
            if (!pack_read(src.seq(r), len, packed + w, nw, avx2)) exc_lists[t].push_back(r);
            w += nw;
        }
    });
    // exception side channel (rare): raw bytes of the reads that could not be packed
    uint32_t n_exc = 0;
    uint64_t exc_bytes = 0;
    for (auto &l : exc_lists) {
        n_exc += (uint32_t)l.size();
        for (uint32_t r : l) exc_bytes += lengths[r];
    }
    uint64_t *exc_off = nullptr;
    uint8_t *exc_b = nullptr;
    if (n_exc) {
        const size_t o_b = align_up(((size_t)n_exc + 1) * 8, 16);
        if (!p->exc.ensure(o_b + exc_bytes + 16)) {
            pf::set_error("pf_pack_reads: out of host memory");
            return fail(PF_ERR_NOMEM);
        }
        exc_off = reinterpret_cast<uint64_t *>(p->exc.p);
        exc_b = static_cast<uint8_t *>(p->exc.p) + o_b;
        uint32_t e = 0;
        uint64_t eb = 0;
        for (auto &l : exc_lists)  // thread ranges are contiguous and ascending: indices stay sorted
            for (uint32_t r : l) {
                exc_index[r] = e;
                exc_off[e++] = eb;
                memcpy(exc_b + eb, src.seq(r), lengths[r]);
                eb += lengths[r];
            }
        exc_off[n_exc] = eb;
    }
    p->b.n_reads = n_reads;
    p->b.n_exc = n_exc;
    p->b.lengths = lengths;
    p->b.word_off = word_off;
    p->b.packed = packed;
    p->b.n_words = n_words + pad_words;
    p->b.exc_index = n_exc ? exc_index : nullptr;
    p->b.exc_off = exc_off;
    p->b.exc_bytes = exc_b;
    p->b.max_length = max_length;
    p->b.total_bases = total_bases;
    *out = p;
    return PF_OK;
}

const pf_read_batch *pf_packed_batch(const pf_packed *p) { return p ? &p->b : nullptr; }

int pf_packed_reserve_like(pf_packed **inout, const pf_packed *model) {
    if (!inout || !model) {
        pf::set_error("pf_packed_reserve_like: null argument");
        return PF_ERR_ARG;
    }
    pf_packed *p = *inout ? *inout : new pf_packed();
    if (!p->main.ensure(model->main.cap) || (model->exc.cap && !p->exc.ensure(model->exc.cap))) {
        if (!*inout) pf_packed_free(p);
        pf::set_error("pf_packed_reserve_like: out of host memory");
        return PF_ERR_NOMEM;
    }
    *inout = p;
    return PF_OK;
}

void pf_packed_free(pf_packed *p) {
    if (!p) return;
    p->main.release();
    p->exc.release();
    delete p;
}

}  // extern "C"
