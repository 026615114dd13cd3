// pf_format.cpp -- reader/writer for tree.bin and .bf (see pf_format.h).
#include "pf_format.h"

#include <cerrno>
#include <cstdio>
#include <cstring>

namespace pf {

namespace {
const char kBitOrder[] = "bitvec::order::Lsb0";  // core::any::type_name::<Lsb0>()

struct Reader {
    FILE *fp = nullptr;
    bool ok = true;
    explicit Reader(const std::string &p) { fp = fopen(p.c_str(), "rb"); }
    ~Reader() {
        if (fp) fclose(fp);
    }
    template <class T>
    T get() {
        T v{};
        if (ok && fread(&v, sizeof v, 1, fp) != 1) ok = false;
        return v;
    }
    std::string str() {
        uint64_t n = get<uint64_t>();
        if (!ok || n > (1u << 20)) {
            ok = false;
            return {};
        }
        std::string s(n, '\0');
        if (n && fread(&s[0], 1, n, fp) != n) ok = false;
        return s;
    }
};
struct Writer {
    FILE *fp = nullptr;
    bool ok = true;
    explicit Writer(const std::string &p) { fp = fopen(p.c_str(), "wb"); }
    ~Writer() {
        if (fp) fclose(fp);
    }
    template <class T>
    void put(T v) {
        if (ok && fwrite(&v, sizeof v, 1, fp) != 1) ok = false;
    }
    void str(const std::string &s) {
        put<uint64_t>(s.size());
        if (ok && !s.empty() && fwrite(s.data(), 1, s.size(), fp) != s.size()) ok = false;
    }
    bool close() {
        if (fp && fclose(fp) != 0) ok = false;
        fp = nullptr;
        return ok;
    }
};
}  // namespace

std::string join_path(const std::string &dir, const std::string &name) {
    if (dir.empty() || dir.back() == '/') return dir + name;
    return dir + "/" + name;
}

// BloomNode = { left_child: Option<Box<BloomNode>>, right_child: Option<Box<..>>, bloom_filter_path: PathBuf,
//               tax_id: Option<String>, mapped_reads: usize }  -- serialised pre-order, children inline.
// Iterative decoder (greedy insertion can produce very deep trees).
bool read_tree_bin(const std::string &path, HostTree &t, std::string &err) {
    Reader r(path);
    if (!r.fp) {
        err = "cannot open " + path + ": " + strerror(errno);
        return false;
    }
    t = HostTree{};
    uint8_t tag = r.get<uint8_t>();
    if (!r.ok || tag > 1) {
        err = "tree.bin: bad root tag in " + path;
        return false;
    }
    if (tag == 1) {
        struct Frame {
            int32_t node;
            int stage;
        };
        std::vector<Frame> st;
        t.nodes.emplace_back();
        t.root = 0;
        st.push_back({0, 0});
        while (!st.empty() && r.ok) {
            Frame &f = st.back();
            int32_t me = f.node;
            if (f.stage == 0 || f.stage == 1) {
                uint8_t ct = r.get<uint8_t>();
                if (!r.ok || ct > 1) {
                    r.ok = false;
                    break;
                }
                int stage = f.stage;
                f.stage++;
                if (ct == 1) {
                    int32_t child = (int32_t)t.nodes.size();
                    t.nodes.emplace_back();
                    if (stage == 0) t.nodes[me].left = child;
                    else t.nodes[me].right = child;
                    st.push_back({child, 0});  // invalidates f
                }
            } else {
                HostNode &n = t.nodes[me];
                n.bf_path = r.str();
                uint8_t tt = r.get<uint8_t>();
                if (!r.ok || tt > 1) {
                    r.ok = false;
                    break;
                }
                n.has_tax = tt == 1;
                if (n.has_tax) n.tax_id = r.str();
                n.mapped_reads = r.get<uint64_t>();
                st.pop_back();
            }
        }
    }
    t.false_pos_rate = r.get<float>();
    t.largest_genome = r.get<uint32_t>();
    t.kmer_size = r.get<uint64_t>();
    t.seed1 = r.get<uint64_t>();
    t.seed2 = r.get<uint64_t>();
    if (!r.ok) {
        err = "tree.bin does not decode: " + path;
        return false;
    }
    return true;
}

bool write_tree_bin(const std::string &path, const HostTree &t, std::string &err) {
    Writer w(path);
    if (!w.fp) {
        err = "cannot create " + path + ": " + strerror(errno);
        return false;
    }
    if (t.root < 0) {
        w.put<uint8_t>(0);
    } else {
        w.put<uint8_t>(1);
        struct Frame {
            int32_t node;
            int stage;
        };
        std::vector<Frame> st{{t.root, 0}};
        while (!st.empty()) {
            Frame &f = st.back();
            const HostNode &n = t.nodes[f.node];
            if (f.stage == 0 || f.stage == 1) {
                int32_t c = f.stage == 0 ? n.left : n.right;
                f.stage++;
                if (c >= 0) {
                    w.put<uint8_t>(1);
                    st.push_back({c, 0});
                } else {
                    w.put<uint8_t>(0);
                }
            } else {
                w.str(n.bf_path);
                w.put<uint8_t>(n.has_tax ? 1 : 0);
                if (n.has_tax) w.str(n.tax_id);
                w.put<uint64_t>(n.mapped_reads);
                st.pop_back();
            }
        }
    }
    w.put<float>(t.false_pos_rate);
    w.put<uint32_t>(t.largest_genome);
    w.put<uint64_t>(t.kmer_size);
    w.put<uint64_t>(t.seed1);
    w.put<uint64_t>(t.seed2);
    if (!w.close()) {
        err = "write error on " + path;
        return false;
    }
    return true;
}

// BloomFilter = { bits: BitVec<usize,Lsb0>, num_hashes: u32, hash_builder_one: {seed}, hash_builder_two: {seed},
//                 file_path: Option<PathBuf> }.  bitvec serde: { order: str, head: {width: u8, index: u8},
//                 bits: u64, data: [u64] }.
bool read_bf(const std::string &path, BfHeader &h, uint64_t *dst, uint64_t cap_words, std::string &err) {
    Reader r(path);
    if (!r.fp) {
        err = "Failed to open Bloom filter file: " + path;
        return false;
    }
    std::string order = r.str();
    uint8_t width = r.get<uint8_t>(), index = r.get<uint8_t>();
    h.num_bits = r.get<uint64_t>();
    h.n_words = r.get<uint64_t>();
    if (!r.ok || order != kBitOrder || width != 64 || index != 0 || h.n_words != (h.num_bits + 63) / 64) {
        err = "Failed to deserialize Bloom filter from file (bitvec header): " + path;
        return false;
    }
    if (h.n_words > cap_words) {
        err = "Bloom filter " + path + " has a different size from the rest of the database";
        return false;
    }
    if (h.n_words && fread(dst, 8, h.n_words, r.fp) != h.n_words) r.ok = false;
    h.num_hashes = r.get<uint32_t>();
    h.seed1 = r.get<uint64_t>();
    h.seed2 = r.get<uint64_t>();
    uint8_t tag = r.get<uint8_t>();
    if (r.ok && tag == 1) (void)r.str();  // stored file_path is ignored on load (bloom_filter.rs:171)
    if (!r.ok || tag > 1) {
        err = "Failed to deserialize Bloom filter from file: " + path;
        return false;
    }
    return true;
}

bool read_bf_header(const std::string &path, BfHeader &h, std::string &err) {
    Reader r(path);
    if (!r.fp) {
        err = "Failed to open Bloom filter file: " + path;
        return false;
    }
    std::string order = r.str();
    uint8_t width = r.get<uint8_t>(), index = r.get<uint8_t>();
    h.num_bits = r.get<uint64_t>();
    h.n_words = r.get<uint64_t>();
    if (!r.ok || order != kBitOrder || width != 64 || index != 0 || h.n_words != (h.num_bits + 63) / 64 ||
        fseek(r.fp, (long)(h.n_words * 8), SEEK_CUR) != 0) {
        err = "Failed to deserialize Bloom filter from file (bitvec header): " + path;
        return false;
    }
    h.num_hashes = r.get<uint32_t>();
    h.seed1 = r.get<uint64_t>();
    h.seed2 = r.get<uint64_t>();
    if (!r.ok) {
        err = "Failed to deserialize Bloom filter from file: " + path;
        return false;
    }
    return true;
}

bool write_bf(const std::string &path, const BfHeader &h, const uint64_t *words, const std::string &recorded_path,
              std::string &err) {
    Writer w(path);
    if (!w.fp) {
        err = "Failed to create Bloom filter file: " + path;
        return false;
    }
    w.str(kBitOrder);
    w.put<uint8_t>(64);
    w.put<uint8_t>(0);
    w.put<uint64_t>(h.num_bits);
    w.put<uint64_t>(h.n_words);
    if (w.ok && h.n_words && fwrite(words, 8, h.n_words, w.fp) != h.n_words) w.ok = false;
    w.put<uint32_t>(h.num_hashes);
    w.put<uint64_t>(h.seed1);
    w.put<uint64_t>(h.seed2);
    w.put<uint8_t>(1);
    w.str(recorded_path);
    if (!w.close()) {
        err = "write error on " + path;
        return false;
    }
    return true;
}

}  // namespace pf
