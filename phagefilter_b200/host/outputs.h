// outputs.h -- POS_FILTERING / NEG_FILTERING writers of the query driver (src/main.rs:345-364, 394-404) on several
// host threads.
//
// The reference keeps a ResultMap keyed by read id that is cleared after every block of `-b` reads
// (main.rs:345-364, result_map.rs:20-41): a read is "mapped" when ANY record of its block carrying the same id has
// a hit, and its header lists the union of those records' genomes.  Blocks are independent, so whole blocks are
// formatted concurrently into per-task byte buffers, which are then written in input order.
#pragma once
#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <set>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "seq_reader.h"

namespace pfhost {

// main.rs:380-391: an existing output directory is removed with everything in it, then created again (the parent must
// exist, like fs::create_dir)
inline void create_and_overwrite_directory(const std::string &dir) {
    std::error_code ec;
    if (std::filesystem::is_directory(dir, ec)) {
        std::filesystem::remove_all(dir, ec);
        if (ec) die("cannot remove '" + dir + "': " + ec.message());
    }
    std::filesystem::create_directory(dir, ec);
    if (ec) die("cannot create '" + dir + "': " + ec.message());
}

// POS_FILTERING.<ext> / NEG_FILTERING.<ext> in the output directory; only the files whose flag was given exist
// (main.rs:311-321).  Plain file descriptors: the writer places every task's bytes with pwrite.
class FilterFiles {
  public:
    FilterFiles(const std::string &out_dir, const char *ext, bool pos, bool neg) {
        if (pos) pos_fd_ = create(out_dir + "/POS_FILTERING." + ext);
        if (neg) neg_fd_ = create(out_dir + "/NEG_FILTERING." + ext);
    }
    ~FilterFiles() { close_all(); }
    FilterFiles(const FilterFiles &) = delete;
    FilterFiles &operator=(const FilterFiles &) = delete;
    int pos_fd() const { return pos_fd_; }
    int neg_fd() const { return neg_fd_; }
    void close_all() {
        if (pos_fd_ >= 0 && ::close(pos_fd_) != 0) die("cannot write POS_FILTERING");
        pos_fd_ = -1;
        if (neg_fd_ >= 0 && ::close(neg_fd_) != 0) die("cannot write NEG_FILTERING");
        neg_fd_ = -1;
    }

  private:
    static int create(const std::string &path) {
        const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) die("cannot create " + path);
        return fd;
    }
    int pos_fd_ = -1, neg_fd_ = -1;
};

class FilterWriter {
  public:
    // pos_fd / neg_fd may be -1 (flag not given: such records are written nowhere, main.rs:350-359).
    FilterWriter(int pos_fd, int neg_fd, size_t block, const std::vector<std::string> *leaf_ids, Pool *pool)
        : pos_fd_(pos_fd), neg_fd_(neg_fd), block_(block), leaf_ids_(leaf_ids), pool_(pool) {}

    // recs[0..n): the records of whole blocks (only the last block of the input may be short);
    // read_off[n+1], leaf[]: the per-read hit lists of pf_hits (DFS leaf indices).
    void write(const Record *recs, size_t n, const uint64_t *read_off, const uint32_t *leaf) {
        if (!n || (pos_fd_ < 0 && neg_fd_ < 0)) return;
        const size_t n_blocks = (n + block_ - 1) / block_;
        const size_t tasks = std::min<size_t>(n_blocks, (size_t)(pool_ ? pool_->size() : 1) * 4);
        if (pos_.size() < tasks) pos_.resize(tasks), neg_.resize(tasks);
        auto job = [&](int t) {
            std::string &pos = pos_[(size_t)t], &neg = neg_[(size_t)t];
            pos.clear();
            neg.clear();
            const size_t blk0 = n_blocks * (size_t)t / tasks, blk1 = n_blocks * ((size_t)t + 1) / tasks;
            std::unordered_map<std::string_view, std::set<uint32_t>> result_map;  // keys: views into this block's records
            for (size_t blk = blk0; blk < blk1; ++blk) {
                const size_t b0 = blk * block_, b1 = std::min(n, b0 + block_);
                result_map.clear();
                for (size_t i = b0; i < b1; ++i)
                    for (uint64_t j = read_off[i]; j < read_off[i + 1]; ++j)
                        result_map[std::string_view(recs[i].id, recs[i].id_len)].insert(leaf[j]);
                for (size_t i = b0; i < b1; ++i) {
                    const Record &r = recs[i];
                    auto it = result_map.empty() ? result_map.end() : result_map.find(std::string_view(r.id, r.id_len));
                    const bool mapped = it != result_map.end();
                    if ((mapped && pos_fd_ < 0) || (!mapped && neg_fd_ < 0)) continue;
                    std::string &o = mapped ? pos : neg;
                    o.push_back(r.has_qual ? '@' : '>');  // main.rs:394-404
                    o.append(r.id, r.id_len);
                    if (mapped) {  // get_ext_id, result_map.rs:24-37
                        o.append(" |");
                        // ResultMap holds a HashSet<String> of genome ids (result_map.rs:10): two leaves with the same
                        // tax_id (a FASTA id built or added twice) are listed once
                        bool first = true;
                        const size_t ids_at = o.size();
                        for (uint32_t l : it->second) {
                            const std::string &gid = (*leaf_ids_)[l];
                            bool seen = false;
                            for (size_t p = ids_at; !seen && p < o.size();) {
                                const size_t e = std::min(o.find(',', p), o.size());
                                seen = o.compare(p, e - p, gid) == 0;
                                p = e + 1;
                            }
                            if (seen) continue;
                            if (!first) o.push_back(',');
                            o.append(gid);
                            first = false;
                        }
                    }
                    o.push_back('\n');
                    const size_t at = o.size();
                    o.append(r.seq, r.seq_len);
                    for (size_t c = at; c < o.size(); ++c)
                        if (o[c] >= 'a' && o[c] <= 'z') o[c] = (char)(o[c] - 32);  // to_ascii_uppercase, main.rs:347-349
                    if (r.has_qual) {
                        o.append("\n+\n");
                        o.append(r.qual, r.qual_len);
                    }
                    o.push_back('\n');
                }
            }
        };
        if (pool_) pool_->run((int)tasks, job);
        else
            for (size_t t = 0; t < tasks; ++t) job((int)t);
        // input order = task order: every task's bytes go to their place in the file, all tasks at once (one thread
        // copying 1.3 GB into the page cache was 3/4 of the output time)
        std::vector<uint64_t> pos_at(tasks), neg_at(tasks);
        for (size_t t = 0; t < tasks; ++t) {
            pos_at[t] = pos_size_;
            neg_at[t] = neg_size_;
            pos_size_ += pos_[t].size();
            neg_size_ += neg_[t].size();
        }
        auto put = [&](int t) {
            if (pos_fd_ >= 0) pwrite_all(pos_fd_, pos_[(size_t)t], pos_at[(size_t)t], "cannot write POS_FILTERING");
            if (neg_fd_ >= 0) pwrite_all(neg_fd_, neg_[(size_t)t], neg_at[(size_t)t], "cannot write NEG_FILTERING");
        };
        if (pool_) pool_->run((int)tasks, put);
        else
            for (size_t t = 0; t < tasks; ++t) put((int)t);
    }

  private:
    static void pwrite_all(int fd, const std::string &buf, uint64_t at, const char *what) {
        size_t done = 0;
        while (done < buf.size()) {
            const ssize_t w = ::pwrite(fd, buf.data() + done, buf.size() - done, (off_t)(at + done));
            if (w <= 0) die(what);
            done += (size_t)w;
        }
    }
    int pos_fd_, neg_fd_;
    uint64_t pos_size_ = 0, neg_size_ = 0;
    size_t block_;
    const std::vector<std::string> *leaf_ids_;
    Pool *pool_;
    std::vector<std::string> pos_, neg_;
};

}  // namespace pfhost
