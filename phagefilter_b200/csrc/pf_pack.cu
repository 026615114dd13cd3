// pf_pack.cu -- host-side packing of raw reads into the 2-bit batch layout of pf_read_batch.
//
// The reference materialises one Vec<u8> per canonical k-mer (file_parser.rs:135-148); here a read
// travels as 2 bits per base and k-mers are re-created on the GPU.  Reads holding any byte other than
// upper-case A/C/G/T are hashed verbatim by the reference, so they are carried as raw bytes too
// (exception side channel) and take the byte-exact path on the device.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pf_common.h"

struct pf_packed {
    pf_read_batch b{};
    void *pinned = nullptr;  // one allocation carved into the arrays below
    size_t pinned_bytes = 0;
    bool is_pinned = false;  // false: pageable memory (no CUDA device present; packing is host logic)
};

namespace {
inline int code_of(uint8_t c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return -1;
    }
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace

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

int pf_pack_reads(const uint8_t *seqs, const uint64_t *offs, uint32_t n_reads, pf_packed **out) {
    if (!out || !offs || (!seqs && n_reads && offs[n_reads] > 0)) {
        pf::set_error("pf_pack_reads: null argument");
        return PF_ERR_ARG;
    }
    *out = nullptr;
    // pass 1: sizes
    uint64_t n_words = 0, exc_bytes = 0, total_bases = 0;
    uint32_t n_exc = 0, max_length = 0;
    std::vector<uint8_t> is_exc(n_reads, 0);
    for (uint32_t r = 0; r < n_reads; ++r) {
        const uint64_t len = offs[r + 1] - offs[r];
        if (len > 0xFFFFFFFFull) {
            pf::set_error("read %u longer than 2^32-1 bases", r);
            return PF_ERR_ARG;
        }
        const uint8_t *s = seqs + offs[r];
        total_bases += len;
        if (len > max_length) max_length = (uint32_t)len;
        bool exc = false;
        for (uint64_t j = 0; j < len; ++j)
            if (code_of(s[j]) < 0) {
                exc = true;
                break;
            }
        is_exc[r] = exc;
        if (exc) {
            n_exc++;
            exc_bytes += len;
        }
        n_words += align_up((len + 15) / 16, 2);  // every read starts on an 8-byte boundary
    }
    const uint64_t pad_words = 4;
    size_t o_len = 0;
    size_t o_woff = align_up(o_len + (size_t)n_reads * 4, 16);
    size_t o_packed = align_up(o_woff + (size_t)n_reads * 8, 16);
    size_t o_excidx = align_up(o_packed + (size_t)(n_words + pad_words) * 4, 16);
    size_t o_excoff = align_up(o_excidx + (n_exc ? (size_t)n_reads * 4 : 0), 16);
    size_t o_excbytes = align_up(o_excoff + (n_exc ? ((size_t)n_exc + 1) * 8 : 0), 16);
    size_t total = o_excbytes + (size_t)exc_bytes + 16;
    pf_packed *p = new pf_packed();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
        p->pinned = pf_alloc_pinned(total);
        p->is_pinned = p->pinned != nullptr;
    } else {
        cudaGetLastError();
    }
    if (!p->pinned) p->pinned = aligned_alloc(64, align_up(total, 64));
    if (!p->pinned) {
        delete p;
        pf::set_error("pf_pack_reads: out of host memory (%zu bytes)", total);
        return PF_ERR_NOMEM;
    }
    p->pinned_bytes = total;
    uint8_t *base = static_cast<uint8_t *>(p->pinned);
    uint32_t *lengths = reinterpret_cast<uint32_t *>(base + o_len);
    uint64_t *word_off = reinterpret_cast<uint64_t *>(base + o_woff);
    uint32_t *packed = reinterpret_cast<uint32_t *>(base + o_packed);
    uint32_t *exc_index = n_exc ? reinterpret_cast<uint32_t *>(base + o_excidx) : nullptr;
    uint64_t *exc_off = n_exc ? reinterpret_cast<uint64_t *>(base + o_excoff) : nullptr;
    uint8_t *exc_b = n_exc ? base + o_excbytes : nullptr;
    memset(packed, 0, (size_t)(n_words + pad_words) * 4);
    // pass 2: fill
    uint64_t w = 0, eb = 0;
    uint32_t e = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        const uint64_t len = offs[r + 1] - offs[r];
        const uint8_t *s = seqs + offs[r];
        lengths[r] = (uint32_t)len;
        word_off[r] = w;
        if (!is_exc[r]) {
            for (uint64_t j = 0; j < len; ++j) packed[w + (j >> 4)] |= (uint32_t)code_of(s[j]) << (2 * (j & 15));
        }
        if (exc_index) exc_index[r] = is_exc[r] ? e : pf::NONE32;
        if (is_exc[r]) {
            exc_off[e++] = eb;
            memcpy(exc_b + eb, s, len);
            eb += len;
        }
        w += align_up((len + 15) / 16, 2);
    }
    if (exc_off) exc_off[n_exc] = eb;
    p->b.n_reads = n_reads;
    p->b.n_exc = n_exc;
    p->b.lengths = lengths;
    p->b.word_off = word_off;
    p->b.packed = packed;
    p->b.n_words = n_words + pad_words;
    p->b.exc_index = exc_index;
    p->b.exc_off = exc_off;
    p->b.exc_bytes = exc_b;
    p->b.max_length = max_length;
    p->b.total_bases = total_bases;
    *out = p;
    return PF_OK;
}

const pf_read_batch *pf_packed_batch(const pf_packed *p) { return p ? &p->b : nullptr; }

void pf_packed_free(pf_packed *p) {
    if (!p) return;
    if (p->is_pinned) pf_free_pinned(p->pinned);
    else free(p->pinned);
    delete p;
}

}  // extern "C"
