// pf_format.h -- the reference's on-disk database format.
//   <db>/tree.bin  = bincode(BloomTree)   (bloom_tree.rs:28-61, 339-355)
//   <db>/<name>.bf = bincode(BloomFilter) (bloom_filter.rs:84-93, 176-207) with bitvec's serde layout
// bincode 1.3 defaults: little-endian, fixed-width integers, u64 lengths, Option = u8 tag.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pf {

struct HostNode {
    int32_t left = -1, right = -1;  // indices into HostTree::nodes
    std::string bf_path;            // BloomNode.bloom_filter_path (relative file name)
    bool has_tax = false;
    std::string tax_id;             // BloomNode.tax_id
    uint64_t mapped_reads = 0;
};

struct HostTree {
    std::vector<HostNode> nodes;  // pre-order
    int32_t root = -1;
    float false_pos_rate = 0.f;
    uint32_t largest_genome = 0;
    uint64_t kmer_size = 0;
    uint64_t seed1 = 0, seed2 = 0;
};

struct BfHeader {
    uint64_t num_bits = 0, n_words = 0;
    uint32_t num_hashes = 0;
    uint64_t seed1 = 0, seed2 = 0;
};

bool read_tree_bin(const std::string &path, HostTree &out, std::string &err);
bool write_tree_bin(const std::string &path, const HostTree &t, std::string &err);
// Decodes one .bf; the n_words payload words are written to dst (capacity cap_words).
bool read_bf(const std::string &path, BfHeader &hdr, uint64_t *dst, uint64_t cap_words, std::string &err);
// Geometry and seeds only (skips the payload).
bool read_bf_header(const std::string &path, BfHeader &hdr, std::string &err);
bool write_bf(const std::string &path, const BfHeader &hdr, const uint64_t *words, const std::string &recorded_path,
              std::string &err);
std::string join_path(const std::string &dir, const std::string &name);

}  // namespace pf
