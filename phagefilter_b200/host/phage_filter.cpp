// phage_filter.cpp -- host driver with the reference's command line on top of the C ABI (include/pfgpu.h).
//
// Mirrors src/main.rs of the reference: subcommands build | add | query with the same flags (main.rs:38-136),
// the same stdout lines (main.rs:181,199,285-297,375), the same on-disk DB and the same output files
// (CLASSIFICATION.csv, POS_FILTERING.{fa,fq}, NEG_FILTERING.{fa,fq}; main.rs:311-374, 394-404).
// The reference is Rust; no Rust toolchain exists in this image, so the host side is C++ (INTEGRATION.md shows
// the Rust binding).  File parsing stays on the host (file_parser.rs); everything else runs in libpfgpu.
//
// Differences that cannot change results: --threads and --cache-size are accepted and ignored (all filters are
// resident in HBM; reads are packed by the host thread); reads go to the GPU in batches larger than
// --block-size-reads, while ResultMap's per-block, id-keyed semantics (main.rs:345-364) are kept on the host.
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pfgpu.h"

namespace {

[[noreturn]] void die(const std::string &msg) {
    // the reference panics (exit code 101) on every error on this path
    fprintf(stderr, "thread 'main' panicked: %s\n", msg.c_str());
    exit(101);
}
void check(int rc, const char *what) {
    if (rc != PF_OK) die(std::string(what) + ": " + pf_last_error());
}

// ---- file_parser.rs ----------------------------------------------------------------------------------
enum class Fmt { Auto, Fasta, Fastq };

struct Record {
    std::string id, seq, qual;
    bool has_qual = false;
};

class LineReader {  // transparent gzip (open_reader, file_parser.rs:89-101): gzread also passes plain files through
  public:
    explicit LineReader(const std::string &path) {
        gz_ = gzopen(path.c_str(), "rb");
        if (!gz_) die("Failed to open '" + path + "'");
        gzbuffer(gz_, 1 << 20);
    }
    ~LineReader() {
        if (gz_) gzclose(gz_);
    }
    bool getline(std::string &out) {
        out.clear();
        for (;;) {
            if (pos_ == len_) {
                int n = gzread(gz_, buf_, sizeof buf_);
                if (n <= 0) return !out.empty() || had_partial_();
                len_ = (size_t)n;
                pos_ = 0;
            }
            const char *nl = (const char *)memchr(buf_ + pos_, '\n', len_ - pos_);
            if (nl) {
                out.append(buf_ + pos_, nl - (buf_ + pos_));
                pos_ = (size_t)(nl - buf_) + 1;
                if (!out.empty() && out.back() == '\r') out.pop_back();
                return true;
            }
            out.append(buf_ + pos_, len_ - pos_);
            pos_ = len_;
            partial_ = true;
        }
    }
    int peek() {
        if (pos_ == len_) {
            int n = gzread(gz_, buf_, sizeof buf_);
            if (n <= 0) return -1;
            len_ = (size_t)n;
            pos_ = 0;
        }
        return (unsigned char)buf_[pos_];
    }

  private:
    bool had_partial_() {
        bool p = partial_;
        partial_ = false;
        return p;
    }
    gzFile gz_ = nullptr;
    char buf_[1 << 16];
    size_t pos_ = 0, len_ = 0;
    bool partial_ = false;
};

std::string lower_ext(const std::string &name) {
    size_t p = name.rfind('.');
    return p == std::string::npos ? "" : name.substr(p + 1);
}
std::string stem(const std::string &name) {
    size_t p = name.rfind('.');
    return p == std::string::npos ? name : name.substr(0, p);
}
const std::set<std::string> kSeqExt = {"fa", "fasta", "fna", "fsa", "fas", "fq", "fastq"};  // file_parser.rs:303
bool is_gz_ext(const std::string &e) { return e == "gz" || e == "gzip"; }

bool has_supported_extension(const std::string &path) {  // file_parser.rs:323-344
    std::string base = path.substr(path.find_last_of('/') + 1);
    std::string e = lower_ext(base);
    if (e.empty()) return false;
    if (kSeqExt.count(e)) return true;
    if (is_gz_ext(e)) return kSeqExt.count(lower_ext(stem(base))) > 0;
    return false;
}
Fmt format_from_extension(const std::string &path) {  // file_parser.rs:69-86
    std::string base = path.substr(path.find_last_of('/') + 1);
    std::string e = lower_ext(base);
    if (is_gz_ext(e)) e = lower_ext(stem(base));
    return (e == "fq" || e == "fastq") ? Fmt::Fastq : Fmt::Fasta;
}
Fmt detect_format(const std::string &path, Fmt override_) {  // file_parser.rs:33-66
    if (override_ != Fmt::Auto) return override_;
    LineReader r(path);
    int c = r.peek();
    if (c == '>') return Fmt::Fasta;
    if (c == '@') return Fmt::Fastq;
    return format_from_extension(path);
}
std::string first_token(const std::string &header) {  // bio: id = header up to the first whitespace
    size_t b = 1, e = b;
    while (e < header.size() && header[e] != ' ' && header[e] != '\t') ++e;
    return header.substr(b, e - b);
}

class RecordStream {
  public:
    RecordStream(const std::string &path, Fmt fmt) : r_(path), fmt_(fmt) {}
    bool next(Record &rec) {
        rec = Record{};
        std::string line;
        if (fmt_ == Fmt::Fastq) {
            do {
                if (!r_.getline(line)) return false;
            } while (line.empty());
            if (line[0] != '@') die("Expected @ at record start");
            rec.id = first_token(line);
            rec.has_qual = true;
            while (r_.getline(line) && (line.empty() || line[0] != '+')) rec.seq += line;
            while (rec.qual.size() < rec.seq.size() && r_.getline(line)) rec.qual += line;
            return true;
        }
        if (pending_.empty()) {
            do {
                if (!r_.getline(line)) return false;
            } while (line.empty());
            pending_ = line;
        }
        if (pending_[0] != '>') die("Expected > at record start.");
        rec.id = first_token(pending_);
        pending_.clear();
        while (r_.getline(line)) {
            if (!line.empty() && line[0] == '>') {
                pending_ = line;
                break;
            }
            rec.seq += line;
        }
        return true;
    }

  private:
    LineReader r_;
    Fmt fmt_;
    std::string pending_;
};

class ReadQueue {  // file_parser.rs:227-301
  public:
    ReadQueue(const std::string &path, Fmt fmt) : fmt_(fmt) {
        struct stat st;
        if (stat(path.c_str(), &st) != 0) die("No such file or directory: " + path);
        if (S_ISREG(st.st_mode)) {
            files_.push_back(path);
        } else {
            DIR *d = opendir(path.c_str());
            if (!d) die("cannot read directory " + path);
            while (dirent *e = readdir(d)) {
                std::string p = path + (path.back() == '/' ? "" : "/") + e->d_name;
                struct stat s2;
                if (stat(p.c_str(), &s2) == 0 && S_ISREG(s2.st_mode) && has_supported_extension(p)) files_.push_back(p);
            }
            closedir(d);
            std::sort(files_.begin(), files_.end());  // read_dir order is unspecified in the reference
        }
    }
    Fmt peek_format() const { return files_.empty() ? Fmt::Fasta : detect_format(files_.back(), fmt_); }
    bool next(Record &rec) {
        for (;;) {
            if (!cur_) {
                if (files_.empty()) return false;
                std::string f = files_.back();  // popped from the END (file_parser.rs:238)
                files_.pop_back();
                cur_ = new RecordStream(f, detect_format(f, fmt_));
            }
            if (cur_->next(rec)) return true;
            delete cur_;
            cur_ = nullptr;
        }
    }
    ~ReadQueue() { delete cur_; }

  private:
    std::vector<std::string> files_;
    Fmt fmt_;
    RecordStream *cur_ = nullptr;
};

// ---- argument parsing (clap surface of main.rs:38-136) -------------------------------------------------
struct Args {
    std::map<std::string, std::string> kv;
    std::set<std::string> flags;
    std::string get(const std::string &k, const std::string &def = "") const {
        auto it = kv.find(k);
        return it == kv.end() ? def : it->second;
    }
    bool has(const std::string &k) const { return kv.count(k) > 0; }
};
Args parse(int argc, char **argv, int start, const std::map<std::string, std::string> &short_to_long,
           const std::set<std::string> &bool_flags) {
    Args a;
    for (int i = start; i < argc; ++i) {
        std::string t = argv[i], key, val;
        bool have_val = false;
        if (t.rfind("--", 0) == 0) {
            size_t eq = t.find('=');
            key = t.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) {
                val = t.substr(eq + 1);
                have_val = true;
            }
        } else if (t.size() >= 2 && t[0] == '-') {
            auto it = short_to_long.find(t.substr(1, 1));
            if (it == short_to_long.end()) {
                if (t == "-v" || t == "-q" || t.rfind("-vv", 0) == 0) continue;  // clap-verbosity-flag
                die("unexpected argument '" + t + "'");
            }
            key = it->second;
            if (t.size() > 2) {
                val = t.substr(2);
                have_val = true;
            }
        } else {
            die("unexpected argument '" + t + "'");
        }
        if (bool_flags.count(key)) {
            a.flags.insert(key);
            continue;
        }
        if (key == "verbose" || key == "quiet") continue;
        if (!have_val) {
            if (i + 1 >= argc) die("a value is required for '--" + key + "'");
            val = argv[++i];
        }
        a.kv[key] = val;
    }
    return a;
}
Fmt parse_fmt(const std::string &s) {
    if (s == "auto" || s.empty()) return Fmt::Auto;
    if (s == "fasta") return Fmt::Fasta;
    if (s == "fastq") return Fmt::Fastq;
    die("invalid value '" + s + "' for '--format'");
}

void create_and_overwrite_directory(const std::string &dir) {  // main.rs:380-391
    struct stat st;
    if (stat(dir.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) {
        std::string cmd = "rm -rf -- '" + dir + "'";
        if (system(cmd.c_str()) != 0) die("cannot remove " + dir);
    }
    mkdir(dir.c_str(), 0777);
}
void write_record(FILE *fp, const std::string &id, const std::string &seq, const Record &r) {  // main.rs:394-404
    if (r.has_qual) fprintf(fp, "@%s\n%s\n+\n%s\n", id.c_str(), seq.c_str(), r.qual.c_str());
    else fprintf(fp, ">%s\n%s\n", id.c_str(), seq.c_str());
}

// ---- build / add (main.rs:148-247) -----------------------------------------------------------------------
int cmd_build(const Args &a, bool add) {
    if (!a.has("genomes") || !a.has("db-path")) die("the following required arguments were not provided: --genomes --db-path");
    const int device = atoi(a.get("device", "0").c_str());
    pf_builder *b = nullptr;
    if (add) {
        puts("Adding new genomes to the SBT...");
        check(pf_builder_open(a.get("db-path").c_str(), device, &b), "BloomTree::load");
    } else {
        puts("Building the SBT...");
        uint64_t s1 = strtoull(a.get("seed-one", "0").c_str(), nullptr, 0), s2 = strtoull(a.get("seed-two", "0").c_str(), nullptr, 0);
        if (!a.has("seed-one") || !a.has("seed-two")) {  // HashSeed::new(): rand::thread_rng().gen() (hasher.rs:23-29)
            FILE *ur = fopen("/dev/urandom", "rb");
            if (!ur || fread(&s1, 8, 1, ur) != 1 || fread(&s2, 8, 1, ur) != 1) die("cannot read /dev/urandom");
            fclose(ur);
        }
        check(pf_builder_create(strtoull(a.get("kmer-size", "20").c_str(), nullptr, 10), strtof(a.get("false-pos-rate", "0.001").c_str(), nullptr),
                                (uint32_t)strtoul(a.get("largest-genome", "1000000").c_str(), nullptr, 10), s1, s2, device,
                                a.get("node-names", "u16") == "counter" ? 0 : 1, s1 ^ s2, &b),
              "BloomTree::new");
    }
    ReadQueue q(a.get("genomes"), parse_fmt(a.get("format", "auto")));
    Record rec;
    while (q.next(rec))  // one leaf per record (main.rs:173-195)
        check(pf_builder_insert(b, rec.id.c_str(), (const uint8_t *)rec.seq.data(), rec.seq.size()), "BloomTree::insert");
    check(pf_builder_save(b, a.get("db-path").c_str()), "BloomTree::save");
    pf_builder_free(b);
    puts("Finished.");
    return 0;
}

// ---- query (main.rs:249-376) -------------------------------------------------------------------------------
int cmd_query(const Args &a) {
    if (!a.has("reads") || !a.has("out") || !a.has("db-path"))
        die("the following required arguments were not provided: --reads --out --db-path");
    const std::string out = a.get("out");
    const size_t block = strtoull(a.get("block-size-reads", "100").c_str(), nullptr, 10);
    const float theta = strtof(a.get("filter-threshold", "1.0").c_str(), nullptr);
    const bool pos = a.flags.count("pos-filter") > 0, neg = a.flags.count("neg-filter") > 0;
    const bool filtering = pos || neg;
    const int64_t depth = a.has("search-depth") ? strtoll(a.get("search-depth").c_str(), nullptr, 10) : -1;
    const int device = atoi(a.get("device", "0").c_str());
    const size_t gpu_batch = strtoull(a.get("gpu-batch-reads", "1000000").c_str(), nullptr, 10);
    if (block == 0) die("block size must be positive");

    pf_db *db = nullptr;
    check(pf_db_open(a.get("db-path").c_str(), device, depth, &db), "BloomTree::load");
    pf_db_info_t info;
    check(pf_db_info(db, &info), "pf_db_info");
    if (a.has("hash-rot")) check(pf_db_set_hash_rot(db, atoi(a.get("hash-rot").c_str())), "pf_db_set_hash_rot");

    puts("Querying reads...");
    printf("Filtering settings: positive=%s; negative=%s\n", pos ? "true" : "false", neg ? "true" : "false");
    if (depth >= 0) {
        if (!filtering) puts("If using a search depth, use a filtering flag (--pos-filter or --neg-filter, or both!)");
        printf("Search depth settings: %lld\n", (long long)depth);
    }
    ReadQueue q(a.get("reads"), parse_fmt(a.get("format", "auto")));
    create_and_overwrite_directory(out);
    const char *ext = q.peek_format() == Fmt::Fastq ? "fq" : "fa";
    FILE *pos_fp = nullptr, *neg_fp = nullptr;
    if (pos && !(pos_fp = fopen((out + "/POS_FILTERING." + ext).c_str(), "wb"))) die("cannot create POS_FILTERING");
    if (neg && !(neg_fp = fopen((out + "/NEG_FILTERING." + ext).c_str(), "wb"))) die("cannot create NEG_FILTERING");

    std::vector<Record> recs;
    std::vector<uint64_t> offs;
    std::string blob;
    // GPU batches are whole multiples of the reference's block so block boundaries stay where they were
    const size_t batch_reads = std::max(block, gpu_batch / block * block);
    bool more = true;
    while (more) {
        recs.clear();
        offs.assign(1, 0);
        blob.clear();
        Record rec;
        while (recs.size() < batch_reads && (more = q.next(rec))) {
            blob += rec.seq;  // k-mers come from the raw bytes (file_parser.rs:203-205)
            offs.push_back(blob.size());
            if (!filtering) {
                rec.seq.clear();
                rec.qual.clear();
            }
            recs.push_back(std::move(rec));
        }
        if (recs.empty()) break;
        pf_packed *packed = nullptr;
        check(pf_pack_reads((const uint8_t *)blob.data(), offs.data(), (uint32_t)recs.size(), &packed), "pf_pack_reads");
        pf_hits hits{};
        check(pf_query_block(db, pf_packed_batch(packed), theta, filtering ? 1 : 0, &hits), "query_batch");
        if (filtering) {
            // ResultMap is keyed by read id and cleared after every block of `block` reads (main.rs:345-364)
            for (size_t b0 = 0; b0 < recs.size(); b0 += block) {
                const size_t b1 = std::min(recs.size(), b0 + block);
                std::unordered_map<std::string, std::set<uint32_t>> result_map;
                for (size_t i = b0; i < b1; ++i)
                    for (uint64_t j = hits.read_off[i]; j < hits.read_off[i + 1]; ++j) result_map[recs[i].id].insert(hits.leaf[j]);
                for (size_t i = b0; i < b1; ++i) {
                    std::string seq = recs[i].seq;
                    for (auto &c : seq) c = (char)toupper((unsigned char)c);  // to_ascii_uppercase, main.rs:347-349
                    auto it = result_map.find(recs[i].id);
                    if (it != result_map.end()) {
                        if (pos_fp) {
                            std::string id = recs[i].id + " |";  // get_ext_id, result_map.rs:24-37
                            bool first = true;
                            for (uint32_t leaf : it->second) {
                                if (!first) id += ",";
                                id += pf_db_leaf_id(db, leaf);
                                first = false;
                            }
                            write_record(pos_fp, id, seq, recs[i]);
                        }
                    } else if (neg_fp) {
                        write_record(neg_fp, recs[i].id, seq, recs[i]);
                    }
                }
            }
        }
        pf_packed_free(packed);
    }
    if (pos_fp) fclose(pos_fp);
    if (neg_fp) fclose(neg_fp);
    check(pf_save_leaf_counts(db, (out + "/CLASSIFICATION.csv").c_str()), "save_leaf_counts");
    pf_db_close(db);
    puts("Finished.");
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) die("usage: phage_filter <build|add|query> [options]");
    const std::string cmd = argv[1];
    if (cmd == "build")
        return cmd_build(parse(argc, argv, 2,
                               {{"g", "genomes"}, {"d", "db-path"}, {"t", "threads"}, {"k", "kmer-size"}, {"c", "cache-size"},
                                {"f", "false-pos-rate"}, {"l", "largest-genome"}, {"F", "format"}},
                               {}),
                         false);
    if (cmd == "add")
        return cmd_build(parse(argc, argv, 2, {{"g", "genomes"}, {"d", "db-path"}, {"t", "threads"}, {"c", "cache-size"}, {"F", "format"}}, {}),
                         true);
    if (cmd == "query")
        return cmd_query(parse(argc, argv, 2,
                               {{"r", "reads"}, {"o", "out"}, {"d", "db-path"}, {"t", "threads"}, {"b", "block-size-reads"},
                                {"f", "filter-threshold"}, {"c", "cache-size"}, {"F", "format"}},
                               {"pos-filter", "neg-filter"}));
    if (cmd == "--version" || cmd == "-V") {
        printf("PhageFilter 2.0 (%s)\n", pf_version());
        return 0;
    }
    die("unrecognized subcommand '" + cmd + "'");
}
