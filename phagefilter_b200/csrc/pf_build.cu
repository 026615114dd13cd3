// pf_build.cu -- GPU database builder producing the reference's on-disk format.
//
//   BloomTree::new            bloom_tree.rs:100-119     pf_builder_create
//   BloomTree::insert         bloom_tree.rs:128-145     pf_builder_insert
//     init_leaf_node          bloom_tree.rs:154-170     insert_kernel (canonical k-mers -> atomicOr)
//     add_to_tree             bloom_tree.rs:187-214     descend_kernel: one cooperative launch walks root -> leaf on the
//     init_internal_node      bloom_tree.rs:226-246     device (union + two Hamming distances per level, grid barrier)
//   BloomFilter::union        bloom_filter.rs:275-278   union_kernel
//   BloomFilter::distance     bloom_filter.rs:142-150   distance2_kernel (popcount of xor)
//   BloomTree::save           bloom_tree.rs:339-355     pf_builder_save (tree.bin + one .bf per node)
#include <sys/stat.h>

#include <cooperative_groups.h>

#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pf_common.h"
#include "pf_format.h"
#include "pf_hash.cuh"

namespace pf {

// One thread per k-mer start; byte-exact canonicalisation and hashing on the raw genome bytes.
__global__ void insert_kernel(const uint8_t *__restrict__ seq, uint64_t n_k, HashParams hp, uint32_t *filt) {
    for (uint64_t pos = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < n_k;
         pos += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t *p = seq + pos;
        const uint64_t hb = canonical_hash_bytes([&](uint32_t j) { return p[j]; }, hp.k);
        const uint64_t h1 = fx_finish(hp.c1, hb, hp.rot), h2 = fx_finish(hp.c2, hb, hp.rot);
        uint64_t g = h1;
        for (uint32_t i = 0; i < hp.K; ++i) {
            const uint64_t idx = mod_any(g, hp.m, hp.M);
            atomicOr(filt + (idx >> 5), 1u << (idx & 31u));
            g = i == 0 ? h2 : (i == 1 ? (h1 + 2ULL) * h2 : g + h2);
        }
    }
}

__global__ void union_kernel(uint64_t *__restrict__ dst, const uint64_t *__restrict__ src, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] |= src[i];
}

// out[0] += hamming(a, x), out[1] += hamming(b, x)
__global__ void distance2_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b,
                                 const uint64_t *__restrict__ x, uint64_t n, unsigned long long *out) {
    unsigned long long da = 0, db = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = x[i];
        da += __popcll(a[i] ^ v);
        db += __popcll(b[i] ^ v);
    }
    for (int o = 16; o; o >>= 1) {
        da += __shfl_xor_sync(0xFFFFFFFFu, da, o);
        db += __shfl_xor_sync(0xFFFFFFFFu, db, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (da) atomicAdd(out, da);
        if (db) atomicAdd(out + 1, db);
    }
}

// add_to_tree + init_internal_node (bloom_tree.rs:187-246) for one new leaf, entirely on the device: one cooperative
// launch walks from the root to the leaf the new genome is paired with.  At a two-child node the new leaf's filter is
// OR-ed into the node (:194) while its Hamming distances to both children are summed (:197-199); after a grid-wide
// barrier every thread reads the two sums and steps right iff right < left (:200-206, ties go left).  At the leaf it
// stops at, the fresh internal node `in` becomes union(new leaf, that leaf) (:237-238) with the existing node as its
// left and the new one as its right child (:242-243), and takes the old node's place under its parent (or as root).
// The tree lives in device arrays (left / right / filter pointer per node, node 0.. in creation order), so the host
// neither waits for a distance nor launches per level: ~50 launches and ~15 round trips per insert became one launch.
struct DescendArgs {
    int32_t *left, *right;       // children per node, -1 = none
    uint64_t *const *fptr;       // filter per node
    int32_t *root;               // device copy of the root index
    unsigned long long *dist;    // [2 * MAX_LEVELS], zeroed before the launch: (right, left) distance sums per level
    int32_t leaf, in;            // the new leaf and the internal node created above the leaf it is paired with
    uint64_t wpf;
    int32_t *err;                // set to 1 on a one-child node ("should not happen", bloom_tree.rs:209-213)
};
constexpr int BUILD_MAX_LEVELS = 4096;

__global__ void __launch_bounds__(256) descend_kernel(const DescendArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t *__restrict__ x = a.fptr[a.leaf];
    int32_t cur = *a.root, parent = -1;
    bool parent_left = false;
    for (int level = 0;; ++level) {
        const int32_t l = a.left[cur], r = a.right[cur];
        if (l < 0 && r < 0) break;
        if (l < 0 || r < 0 || level >= BUILD_MAX_LEVELS) {
            if (tid == 0) *a.err = (l < 0 || r < 0) ? 1 : 2;
            return;  // uniform: every thread sees the same tree
        }
        uint64_t *__restrict__ c = a.fptr[cur];
        const uint64_t *__restrict__ fl = a.fptr[l], *__restrict__ fr = a.fptr[r];
        unsigned long long dr = 0, dl = 0;
        for (uint64_t i = tid; i < a.wpf; i += nthr) {
            const uint64_t v = x[i];
            c[i] |= v;
            dr += __popcll(fr[i] ^ v);
            dl += __popcll(fl[i] ^ v);
        }
        for (int o = 16; o; o >>= 1) {
            dr += __shfl_xor_sync(0xFFFFFFFFu, dr, o);
            dl += __shfl_xor_sync(0xFFFFFFFFu, dl, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (dr) atomicAdd(a.dist + 2 * level, dr);
            if (dl) atomicAdd(a.dist + 2 * level + 1, dl);
        }
        grid.sync();
        const unsigned long long right_d = __ldcg(a.dist + 2 * level), left_d = __ldcg(a.dist + 2 * level + 1);
        parent = cur;
        if (right_d < left_d) {
            parent_left = false;
            cur = r;
        } else {
            parent_left = true;
            cur = l;
        }
    }
    uint64_t *__restrict__ fin = a.fptr[a.in];
    const uint64_t *__restrict__ fc = a.fptr[cur];
    for (uint64_t i = tid; i < a.wpf; i += nthr) fin[i] = x[i] | fc[i];
    if (tid == 0) {
        a.left[a.in] = cur;
        a.right[a.in] = a.leaf;
        if (parent < 0) *a.root = a.in;
        else if (parent_left) a.left[parent] = a.in;
        else a.right[parent] = a.in;
    }
}

__global__ void init_nodes_kernel(int32_t *left, int32_t *right, uint64_t **fptr, int32_t i0, uint64_t *p0, int32_t i1, uint64_t *p1) {
    left[i0] = right[i0] = -1;
    fptr[i0] = p0;
    if (i1 >= 0) {
        left[i1] = right[i1] = -1;
        fptr[i1] = p1;
    }
}

__global__ void diff_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, uint64_t n,
                            unsigned long long *out) {
    unsigned long long d = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        d += a[i] != b[i];
    if (d) atomicAdd(out, d);
}

// Zero d_filter[0..wpf) and insert every canonical k-mer of the genome (init_leaf_node).
int build_leaf_filter(const uint8_t *h_seq, uint64_t len, const HashParams &hp, uint64_t *d_filter, uint64_t wpf,
                      cudaStream_t s) {
    PF_CUDA_OK(cudaMemsetAsync(d_filter, 0, wpf * 8, s));
    const uint64_t k = hp.k;
    const uint64_t n_k = (k == 0 || k > len) ? 0 : len - k + 1;  // file_parser.rs:136-139
    if (n_k == 0) return PF_OK;
    uint8_t *d_seq = nullptr;
    PF_CUDA_OK(cudaMalloc(&d_seq, len));
    cudaError_t e = cudaMemcpyAsync(d_seq, h_seq, len, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        const int grid = (int)std::min<uint64_t>((n_k + 255) / 256, 148 * 16);
        insert_kernel<<<grid, 256, 0, s>>>(d_seq, n_k, hp, reinterpret_cast<uint32_t *>(d_filter));
        e = cudaGetLastError();
    }
    cudaStreamSynchronize(s);
    cudaFree(d_seq);
    PF_CUDA_OK(e);
    return PF_OK;
}

int filters_equal(const uint64_t *a, const uint64_t *b, uint64_t n_words, cudaStream_t s, bool *equal) {
    unsigned long long *d = nullptr, h = 0;
    PF_CUDA_OK(cudaMalloc(&d, 8));
    cudaMemsetAsync(d, 0, 8, s);
    diff_kernel<<<148 * 4, 256, 0, s>>>(a, b, n_words, d);
    cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(d);
    PF_CUDA_OK(e);
    *equal = h == 0;
    return PF_OK;
}

// filter geometry in f32, exactly as bloom_filter.rs:342-357 (f32::ln -> logf, f32::round -> roundf,
// `as usize` / `as u32` saturating).  volatile keeps every intermediate in f32.
static uint64_t needed_bits(float fpr, uint32_t n) {
    volatile float ln22 = 0.693147180559945309417232121458176568f * 0.693147180559945309417232121458176568f;
    volatile float inv = 1.0f / fpr;
    volatile float l = logf(inv);
    volatile float q = l / ln22;
    volatile float v = (float)n * q;
    float r = roundf(v);
    if (!(r > 0.0f)) return 0;
    if (r >= 18446744073709551616.0f) return ~0ULL;
    return (uint64_t)r;
}
static uint32_t optimal_num_hashes(uint64_t bits, uint32_t n) {
    volatile float a = (float)bits / (float)n;
    volatile float b = a * 0.693147180559945309417232121458176568f;
    float r = roundf(b);
    uint32_t k = !(r > 0.0f) ? 0u : (r >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)r);
    return k < 2 ? 2 : (k > 200 ? 200 : k);
}

static uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

}  // namespace pf

using namespace pf;

struct pf_builder {
    int device = 0;
    cudaStream_t stream = nullptr;
    HostTree tree;                    // nodes in creation order; links are refreshed from the device before saving
    std::vector<uint64_t *> filters;  // device filter per node (index = node index), carved out of `chunks`
    std::vector<uint64_t *> chunks;   // filter pool: one allocation per POOL_FILTERS filters instead of one per node
    size_t pool_used = 0;             // filters handed out of the last chunk
    uint64_t m = 0, n_words = 0, wpf = 0;
    uint32_t K = 0;
    HashParams hp{};
    int name_mode = 0;
    uint64_t name_state = 0, name_counter = 0;
    std::vector<uint8_t> name_used;
    // device copy of the tree the descent kernel walks and updates
    int32_t *d_left = nullptr, *d_right = nullptr, *d_root = nullptr, *d_err = nullptr;
    uint64_t **d_fptr = nullptr;
    size_t dev_cap = 0;
    unsigned long long *d_dist = nullptr;
    bool links_on_device = false;     // host links are stale
    int coop_grid = 0;
    // genome staging: two pinned buffers alternate so that the copy of genome i+1 overlaps the descent of genome i
    uint8_t *h_seq[2] = {nullptr, nullptr}, *d_seq[2] = {nullptr, nullptr};
    size_t seq_cap[2] = {0, 0};
    cudaEvent_t seq_done[2] = {nullptr, nullptr};
    int ring = 0;
};
constexpr size_t POOL_FILTERS = 256;

static int ensure_dev_tree(pf_builder *b, size_t n_nodes) {
    if (n_nodes <= b->dev_cap) return PF_OK;
    const size_t cap = std::max<size_t>(n_nodes, std::max<size_t>(1024, b->dev_cap * 2));
    int32_t *l = nullptr, *r = nullptr;
    uint64_t **f = nullptr;
    PF_CUDA_OK(cudaMalloc(&l, cap * 4));
    PF_CUDA_OK(cudaMalloc(&r, cap * 4));
    PF_CUDA_OK(cudaMalloc(&f, cap * 8));
    if (b->dev_cap) {
        PF_CUDA_OK(cudaMemcpyAsync(l, b->d_left, b->dev_cap * 4, cudaMemcpyDeviceToDevice, b->stream));
        PF_CUDA_OK(cudaMemcpyAsync(r, b->d_right, b->dev_cap * 4, cudaMemcpyDeviceToDevice, b->stream));
        PF_CUDA_OK(cudaMemcpyAsync(f, b->d_fptr, b->dev_cap * 8, cudaMemcpyDeviceToDevice, b->stream));
        PF_CUDA_OK(cudaStreamSynchronize(b->stream));
        cudaFree(b->d_left);
        cudaFree(b->d_right);
        cudaFree(b->d_fptr);
    }
    b->d_left = l;
    b->d_right = r;
    b->d_fptr = f;
    b->dev_cap = cap;
    return PF_OK;
}

// host links <- device (after inserts the descent kernel is the only one who knows where the leaves went)
static int pull_links(pf_builder *b) {
    if (!b->links_on_device) return PF_OK;
    const size_t n = b->tree.nodes.size();
    std::vector<int32_t> l(n), r(n);
    int32_t root = -1, err = 0;
    PF_CUDA_OK(cudaMemcpyAsync(l.data(), b->d_left, n * 4, cudaMemcpyDeviceToHost, b->stream));
    PF_CUDA_OK(cudaMemcpyAsync(r.data(), b->d_right, n * 4, cudaMemcpyDeviceToHost, b->stream));
    PF_CUDA_OK(cudaMemcpyAsync(&root, b->d_root, 4, cudaMemcpyDeviceToHost, b->stream));
    PF_CUDA_OK(cudaMemcpyAsync(&err, b->d_err, 4, cudaMemcpyDeviceToHost, b->stream));
    PF_CUDA_OK(cudaStreamSynchronize(b->stream));
    PF_CUDA_OK(cudaGetLastError());
    if (err) {
        set_error(err == 1 ? "Node with only one child encountered - should not happen."
                           : "tree deeper than 4096 levels: not supported by the device-side insert");
        return PF_ERR_STATE;
    }
    for (size_t i = 0; i < n; ++i) {
        b->tree.nodes[i].left = l[i];
        b->tree.nodes[i].right = r[i];
    }
    b->tree.root = root;
    b->links_on_device = false;
    return PF_OK;
}

static int builder_device_state(pf_builder *b) {
    PF_CUDA_OK(cudaMalloc(&b->d_root, 4));
    PF_CUDA_OK(cudaMalloc(&b->d_err, 4));
    PF_CUDA_OK(cudaMemsetAsync(b->d_err, 0, 4, b->stream));
    PF_CUDA_OK(cudaMalloc(&b->d_dist, 2 * BUILD_MAX_LEVELS * 8));
    for (int i = 0; i < 2; ++i) PF_CUDA_OK(cudaEventCreateWithFlags(&b->seq_done[i], cudaEventDisableTiming));
    int per_sm = 0, sms = 0;
    PF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, descend_kernel, 256, 0));
    PF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device));
    b->coop_grid = sms * std::max(1, std::min(per_sm, 4));
    return PF_OK;
}

static int new_node(pf_builder *b, const std::string &id, int32_t *out) {
    if (b->chunks.empty() || b->pool_used == POOL_FILTERS) {
        uint64_t *c = nullptr;
        cudaError_t e = cudaMalloc(&c, POOL_FILTERS * b->wpf * 8);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cannot allocate filters beyond %zu (%.1f MB each)", b->filters.size(), b->wpf * 8 / 1e6);
            return PF_ERR_NOMEM;
        }
        b->chunks.push_back(c);
        b->pool_used = 0;
    }
    uint64_t *f = b->chunks.back() + b->pool_used++ * b->wpf;
    HostNode n;
    n.bf_path = id + ".bf";  // make_bloom_node, bloom_tree.rs:281
    n.has_tax = true;
    n.tax_id = id;
    b->tree.nodes.push_back(n);
    b->filters.push_back(f);
    *out = (int32_t)b->tree.nodes.size() - 1;
    return ensure_dev_tree(b, b->tree.nodes.size());
}

extern "C" {

uint64_t pf_needed_bits(float fpr, uint32_t n) { return needed_bits(fpr, n); }
uint32_t pf_optimal_num_hashes(uint64_t bits, uint32_t n) { return optimal_num_hashes(bits, n); }

int pf_builder_create(uint64_t kmer_size, float fpr, uint32_t largest_genome, uint64_t seed1, uint64_t seed2, int device,
                      int name_mode, uint64_t name_seed, pf_builder **out) {
    if (!out) return PF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libpfgpu has no CPU fallback");
        return PF_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range", device);
        return PF_ERR_ARG;
    }
    pf_builder *b = new pf_builder();
    b->device = device;
    b->m = needed_bits(fpr, largest_genome);  // with_rate, bloom_filter.rs:229-240
    b->K = optimal_num_hashes(b->m, largest_genome);
    if (b->m == 0) {
        delete b;
        set_error("filter with zero bits");
        return PF_ERR_ARG;
    }
    b->n_words = (b->m + 63) / 64;
    b->wpf = (b->n_words + 15) / 16 * 16;
    b->tree.false_pos_rate = fpr;
    b->tree.largest_genome = largest_genome;
    b->tree.kmer_size = kmer_size;
    b->tree.seed1 = seed1;
    b->tree.seed2 = seed2;
    b->hp = make_hash_params(seed1, seed2, kmer_size, b->m, b->K, 26);
    b->name_mode = name_mode;
    b->name_state = name_seed;
    if (name_mode == 1) b->name_used.assign(65536, 0);
    cudaSetDevice(device);
    if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess || builder_device_state(b) != PF_OK) {
        set_error("CUDA error creating builder: %s", cudaGetErrorString(cudaGetLastError()));
        pf_builder_free(b);
        return PF_ERR_CUDA;
    }
    *out = b;
    return PF_OK;
}

// BloomTree::load for the `add` subcommand (main.rs:202-247): every node's filter goes back to the device.
int pf_builder_open(const char *db_path, int device, pf_builder **out) {
    if (!db_path || !out) return PF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: libpfgpu has no CPU fallback");
        return PF_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("device %d out of range", device);
        return PF_ERR_ARG;
    }
    std::string dir(db_path), err;
    pf_builder *b = new pf_builder();
    b->device = device;
    if (!read_tree_bin(join_path(dir, "tree.bin"), b->tree, err)) {
        set_error("%s", err.c_str());
        delete b;
        return err.rfind("cannot open", 0) == 0 ? PF_ERR_IO : PF_ERR_FORMAT;
    }
    if (b->tree.root < 0) {
        set_error("tree.bin holds no root");
        delete b;
        return PF_ERR_FORMAT;
    }
    BfHeader h;
    if (!read_bf_header(join_path(dir, b->tree.nodes[b->tree.root].bf_path), h, err)) {
        set_error("%s", err.c_str());
        delete b;
        return PF_ERR_FORMAT;
    }
    b->m = h.num_bits;
    b->K = h.num_hashes;
    b->n_words = h.n_words;
    b->wpf = (b->n_words + 15) / 16 * 16;
    b->tree.seed1 = h.seed1;
    b->tree.seed2 = h.seed2;
    b->hp = make_hash_params(h.seed1, h.seed2, b->tree.kmer_size, b->m, b->K, 26);
    b->name_mode = 1;
    b->name_state = h.seed1 ^ h.seed2 ^ (uint64_t)b->tree.nodes.size();
    b->name_used.assign(65536, 0);
    cudaSetDevice(device);
    uint64_t *stage = nullptr;
    if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess || builder_device_state(b) != PF_OK ||
        cudaMallocHost(&stage, b->wpf * 8) != cudaSuccess) {
        set_error("CUDA error opening builder: %s", cudaGetErrorString(cudaGetLastError()));
        pf_builder_free(b);
        return PF_ERR_CUDA;
    }
    int rc = PF_OK;
    for (size_t i = 0; i < b->tree.nodes.size() && rc == PF_OK; ++i) {
        const HostNode &n = b->tree.nodes[i];
        unsigned v = 0;
        if (sscanf(n.tax_id.c_str(), "Internal_Node_%u", &v) == 1) {
            if (v < 65536) b->name_used[v] = 1;
            b->name_counter++;
        }
        if (b->chunks.empty() || b->pool_used == POOL_FILTERS) {
            uint64_t *c = nullptr;
            if (cudaMalloc(&c, POOL_FILTERS * b->wpf * 8) != cudaSuccess) {
                cudaGetLastError();
                set_error("cannot allocate filter %zu", i);
                rc = PF_ERR_NOMEM;
                break;
            }
            b->chunks.push_back(c);
            b->pool_used = 0;
        }
        uint64_t *f = b->chunks.back() + b->pool_used++ * b->wpf;
        b->filters.push_back(f);
        memset(stage, 0, b->wpf * 8);
        BfHeader hh;
        if (!read_bf(join_path(dir, n.bf_path), hh, stage, b->wpf, err) || hh.num_bits != b->m || hh.num_hashes != b->K) {
            set_error("%s", err.empty() ? "filter geometry differs inside the database" : err.c_str());
            rc = PF_ERR_FORMAT;
            break;
        }
        cudaMemcpyAsync(f, stage, b->wpf * 8, cudaMemcpyHostToDevice, b->stream);
        cudaStreamSynchronize(b->stream);
    }
    cudaFreeHost(stage);
    if (rc == PF_OK && (rc = ensure_dev_tree(b, b->tree.nodes.size())) == PF_OK) {
        // the tree the descent kernel walks: links and filter pointers of the loaded nodes
        const size_t n = b->tree.nodes.size();
        std::vector<int32_t> l(n), r(n);
        for (size_t i = 0; i < n; ++i) {
            l[i] = b->tree.nodes[i].left;
            r[i] = b->tree.nodes[i].right;
        }
        const int32_t root = b->tree.root;
        if (cudaMemcpy(b->d_left, l.data(), n * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(b->d_right, r.data(), n * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(b->d_fptr, b->filters.data(), n * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(b->d_root, &root, 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("CUDA error uploading the tree: %s", cudaGetErrorString(cudaGetLastError()));
            rc = PF_ERR_CUDA;
        }
    }
    if (rc != PF_OK) {
        pf_builder_free(b);
        return rc;
    }
    *out = b;
    return PF_OK;
}

int pf_builder_set_hash_rot(pf_builder *b, int rot) {
    if (!b || rot < 0 || rot > 63) return PF_ERR_ARG;
    b->hp.rot = (uint32_t)rot;
    return PF_OK;
}

int pf_builder_insert(pf_builder *b, const char *id, const uint8_t *seq, uint64_t len) {
    if (!b || !id || (!seq && len)) {
        set_error("pf_builder_insert: null argument");
        return PF_ERR_ARG;
    }
    PF_CUDA_OK(cudaSetDevice(b->device));
    cudaStream_t s = b->stream;
    int32_t leaf, in = -1;
    int rc = new_node(b, id, &leaf);
    if (rc != PF_OK) return rc;
    const bool first = b->tree.nodes.size() == 1;
    if (!first) {
        // init_internal_node (bloom_tree.rs:226-246): every insert but the first creates exactly one internal node
        char name[64];
        if (b->name_mode == 1 && b->name_counter < 65536) {
            // a random u16 like the reference (bloom_tree.rs:232-234), redrawn until unused; once all 65,536
            // are taken the names continue above the u16 range instead of colliding
            uint16_t n2;
            do {
                n2 = (uint16_t)splitmix64(b->name_state);
            } while (b->name_used[n2]);
            b->name_used[n2] = 1;
            snprintf(name, sizeof name, "Internal_Node_%u", (unsigned)n2);
        } else {
            snprintf(name, sizeof name, "Internal_Node_%llu", (unsigned long long)b->name_counter);
        }
        b->name_counter++;
        if ((rc = new_node(b, name, &in))) return rc;
    }
    init_nodes_kernel<<<1, 1, 0, s>>>(b->d_left, b->d_right, b->d_fptr, leaf, b->filters[leaf], in, in >= 0 ? b->filters[in] : nullptr);
    // init_leaf_node (bloom_tree.rs:154-170): the genome goes up through one of two pinned buffers, so that this copy
    // overlaps the previous genome's descent
    const int slot = b->ring;
    b->ring ^= 1;
    PF_CUDA_OK(cudaEventSynchronize(b->seq_done[slot]));
    if (len > b->seq_cap[slot]) {
        if (b->h_seq[slot]) cudaFreeHost(b->h_seq[slot]);
        if (b->d_seq[slot]) cudaFree(b->d_seq[slot]);
        b->h_seq[slot] = nullptr;
        b->d_seq[slot] = nullptr;
        const size_t cap = std::max<size_t>(len + len / 4, 1 << 16);
        PF_CUDA_OK(cudaMallocHost(&b->h_seq[slot], cap));
        PF_CUDA_OK(cudaMalloc(&b->d_seq[slot], cap));
        b->seq_cap[slot] = cap;
    }
    PF_CUDA_OK(cudaMemsetAsync(b->filters[leaf], 0, b->wpf * 8, s));
    const uint64_t k = b->hp.k;
    const uint64_t n_k = (k == 0 || k > len) ? 0 : len - k + 1;  // file_parser.rs:136-139
    if (n_k) {
        memcpy(b->h_seq[slot], seq, len);
        PF_CUDA_OK(cudaMemcpyAsync(b->d_seq[slot], b->h_seq[slot], len, cudaMemcpyHostToDevice, s));
        const int grid = (int)std::min<uint64_t>((n_k + 255) / 256, 148 * 16);
        insert_kernel<<<grid, 256, 0, s>>>(b->d_seq[slot], n_k, b->hp, reinterpret_cast<uint32_t *>(b->filters[leaf]));
    }
    PF_CUDA_OK(cudaEventRecord(b->seq_done[slot], s));
    if (first) {
        const int32_t root = leaf;
        PF_CUDA_OK(cudaMemcpyAsync(b->d_root, &root, 4, cudaMemcpyHostToDevice, s));
        PF_CUDA_OK(cudaStreamSynchronize(s));
        b->tree.root = leaf;
        return PF_OK;
    }
    // add_to_tree (bloom_tree.rs:187-214) + init_internal_node: one cooperative launch, no host round trip
    PF_CUDA_OK(cudaMemsetAsync(b->d_dist, 0, 2 * BUILD_MAX_LEVELS * 8, s));
    DescendArgs a{};
    a.left = b->d_left;
    a.right = b->d_right;
    a.fptr = b->d_fptr;
    a.root = b->d_root;
    a.dist = b->d_dist;
    a.leaf = leaf;
    a.in = in;
    a.wpf = b->wpf;
    a.err = b->d_err;
    void *args[] = {&a};
    PF_CUDA_OK(cudaLaunchCooperativeKernel((const void *)descend_kernel, dim3(b->coop_grid), dim3(256), args, 0, s));
    b->links_on_device = true;
    return PF_OK;
}

int pf_builder_save(pf_builder *b, const char *db_path) {
    if (!b || !db_path) return PF_ERR_ARG;
    PF_CUDA_OK(cudaSetDevice(b->device));
    {
        int prc = pull_links(b);
        if (prc != PF_OK) return prc;
    }
    mkdir(db_path, 0777);
    std::string dir(db_path), err;
    // tree.bin wants pre-order with inline children: re-index from creation order
    HostTree out = b->tree;
    out.nodes.clear();
    out.root = -1;
    if (b->tree.root >= 0) {
        std::vector<int32_t> map(b->tree.nodes.size(), -1), st{b->tree.root};
        std::vector<int32_t> order;
        while (!st.empty()) {
            int32_t u = st.back();
            st.pop_back();
            map[u] = (int32_t)order.size();
            order.push_back(u);
            if (b->tree.nodes[u].right >= 0) st.push_back(b->tree.nodes[u].right);
            if (b->tree.nodes[u].left >= 0) st.push_back(b->tree.nodes[u].left);
        }
        for (int32_t u : order) {
            HostNode n = b->tree.nodes[u];
            n.left = n.left >= 0 ? map[n.left] : -1;
            n.right = n.right >= 0 ? map[n.right] : -1;
            out.nodes.push_back(n);
        }
        out.root = 0;
    }
    if (!write_tree_bin(join_path(dir, "tree.bin"), out, err)) {
        set_error("%s", err.c_str());
        return PF_ERR_IO;
    }
    BfHeader h;
    h.num_bits = b->m;
    h.n_words = b->n_words;
    h.num_hashes = b->K;
    h.seed1 = b->tree.seed1;
    h.seed2 = b->tree.seed2;
    // one .bf per node: several writer threads, each with its own pinned buffer and stream (device -> host copy of one
    // filter overlaps the file writes of the others)
    PF_CUDA_OK(cudaStreamSynchronize(b->stream));
    const size_t n = b->tree.nodes.size();
    const unsigned hc = std::thread::hardware_concurrency();
    const size_t n_thr = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hc ? hc : 4, 8), n));
    std::atomic<size_t> next{0};
    std::atomic<int> status{PF_OK};
    std::vector<std::string> errs(n_thr);
    auto worker = [&](size_t t) {
        cudaSetDevice(b->device);
        uint64_t *stage = nullptr;
        cudaStream_t ws = nullptr;
        if (cudaMallocHost(&stage, b->wpf * 8) != cudaSuccess || cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking) != cudaSuccess) {
            errs[t] = std::string("CUDA error preparing a writer: ") + cudaGetErrorString(cudaGetLastError());
            status = PF_ERR_CUDA;
        }
        for (size_t i; status == PF_OK && (i = next.fetch_add(1)) < n;) {
            if (cudaMemcpyAsync(stage, b->filters[i], b->wpf * 8, cudaMemcpyDeviceToHost, ws) != cudaSuccess ||
                cudaStreamSynchronize(ws) != cudaSuccess) {
                errs[t] = std::string("CUDA error copying filter: ") + cudaGetErrorString(cudaGetLastError());
                status = PF_ERR_CUDA;
                break;
            }
            const std::string p = join_path(dir, b->tree.nodes[i].bf_path);
            // Drop writes each filter with file_path = directory.join(name) (bloom_filter.rs:105-117, bloom_tree.rs:282)
            std::string e;
            if (!write_bf(p, h, stage, p, e)) {
                errs[t] = e;
                status = PF_ERR_IO;
            }
        }
        if (ws) cudaStreamDestroy(ws);
        if (stage) cudaFreeHost(stage);
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < n_thr; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th) t.join();
    if (status != PF_OK)
        for (auto &e : errs)
            if (!e.empty()) {
                set_error("%s", e.c_str());
                break;
            }
    return status;
}

void pf_builder_free(pf_builder *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (auto c : b->chunks) cudaFree(c);
    cudaFree(b->d_left);
    cudaFree(b->d_right);
    cudaFree(b->d_fptr);
    cudaFree(b->d_root);
    cudaFree(b->d_err);
    cudaFree(b->d_dist);
    for (int i = 0; i < 2; ++i) {
        if (b->h_seq[i]) cudaFreeHost(b->h_seq[i]);
        if (b->d_seq[i]) cudaFree(b->d_seq[i]);
        if (b->seq_done[i]) cudaEventDestroy(b->seq_done[i]);
    }
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

// ---- roofline micro-benchmark: independent random 32-byte-sector loads --------------------------
__global__ void sector_gather_kernel(const uint32_t *__restrict__ buf, uint64_t n_sectors, int iters, uint32_t *sink) {
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ULL + 12345;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 8 independent loads in flight per thread
            x ^= x << 13;
            x ^= x >> 7;
            x ^= x << 17;
            const uint64_t sec = (uint64_t)(((unsigned __int128)x * n_sectors) >> 64);
            acc += __ldg(buf + sec * 8 + (x & 7));
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int pf_microbench_sectors(int device, uint64_t bytes, int iters, double *sectors_per_s) {
    if (!sectors_per_s || bytes < 32 || iters < 1) return PF_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device");
        return PF_ERR_CUDA;
    }
    PF_CUDA_OK(cudaSetDevice(device));
    uint32_t *buf = nullptr, *sink = nullptr;
    const uint64_t n_sectors = bytes / 32;
    PF_CUDA_OK(cudaMalloc(&buf, n_sectors * 32));
    PF_CUDA_OK(cudaMalloc(&sink, 4));
    PF_CUDA_OK(cudaMemset(buf, 0x5a, n_sectors * 32));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    sector_gather_kernel<<<grid, block>>>(buf, n_sectors, 4, sink);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        sector_gather_kernel<<<grid, block>>>(buf, n_sectors, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double rate = (double)grid * block * iters * 8 / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    PF_CUDA_OK(e);
    *sectors_per_s = best;
    return PF_OK;
}

}  // extern "C"
