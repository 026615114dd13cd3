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
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <string_view>
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

// A record as slices of the parse buffer (or of the carry arena): nothing is copied per read.
struct Record {
    const char *id = nullptr, *seq = nullptr, *qual = nullptr;
    uint32_t id_len = 0, seq_len = 0;
    bool has_qual = false;
    std::string id_str() const { return std::string(id, id_len); }
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
Fmt detect_format(const std::string &path, Fmt override_) {  // file_parser.rs:33-66 (gzread passes plain files through)
    if (override_ != Fmt::Auto) return override_;
    gzFile gz = gzopen(path.c_str(), "rb");
    if (!gz) die("Failed to open '" + path + "'");
    char c = 0;
    int n = gzread(gz, &c, 1);
    gzclose(gz);
    if (n == 1 && c == '>') return Fmt::Fasta;
    if (n == 1 && c == '@') return Fmt::Fastq;
    return format_from_extension(path);
}

// One input file read in large blocks; records are parsed in place (memchr line scanning; multi-line
// sequences are compacted inside the buffer) and handed out as slices.  open_reader: file_parser.rs:89-101.
class SeqFile {
  public:
    SeqFile(const std::string &path, Fmt fmt, size_t buf_bytes) : fmt_(fmt) {
        gz_ = gzopen(path.c_str(), "rb");
        if (!gz_) die("Failed to open '" + path + "'");
        gzbuffer(gz_, 1 << 20);
        buf_.resize(buf_bytes);
    }
    ~SeqFile() {
        if (gz_) gzclose(gz_);
    }
    // Parses up to max_records complete records into out (appending).  Slices stay valid until the next call.
    // Returns false when the file is exhausted and nothing was appended.
    bool next_records(std::vector<Record> &out, size_t max_records) {
        const size_t before = out.size();
        refill();
        for (;;) {
            while (out.size() - before < max_records) {
                Record r;
                const size_t used = fmt_ == Fmt::Fastq ? parse_fastq(r) : parse_fasta(r);
                if (!used) break;
                pos_ += used;
                out.push_back(r);
            }
            if (out.size() > before || eof_) break;
            // not even one complete record fits: grow the buffer and read more
            if (len_ - pos_ >= buf_.size() / 2) buf_.resize(buf_.size() * 2);
            if (!refill_more()) break;
        }
        return out.size() > before;
    }

  private:
    void refill() {  // move the unparsed tail to the front, then fill the rest of the buffer
        if (pos_ > 0) {
            memmove(buf_.data(), buf_.data() + pos_, len_ - pos_);
            len_ -= pos_;
            pos_ = 0;
        }
        refill_more();
    }
    bool refill_more() {
        bool got = false;
        while (!eof_ && len_ < buf_.size()) {
            const size_t want = std::min<size_t>(buf_.size() - len_, 1u << 30);
            const int n = gzread(gz_, buf_.data() + len_, (unsigned)want);
            if (n <= 0) {
                eof_ = true;
                break;
            }
            len_ += (size_t)n;
            got = true;
        }
        return got;
    }
    // [p, end of line) ; returns pointer past the newline, or nullptr if the line is not complete in the buffer
    const char *line_end(const char *p, const char *&eol) const {
        const char *lim = buf_.data() + len_;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(lim - p));
        if (!nl) {
            if (!eof_) return nullptr;
            eol = lim;  // last line without a newline
            return lim;
        }
        eol = nl;
        return nl + 1;
    }
    static void strip_cr(const char *b, const char *&e) {
        if (e > b && e[-1] == '\r') --e;
    }
    static void set_id(Record &r, const char *b, const char *e) {  // bio: id = header up to the first whitespace
        const char *q = b + 1;
        while (q < e && *q != ' ' && *q != '\t') ++q;
        r.id = b + 1;
        r.id_len = (uint32_t)(q - (b + 1));
    }
    size_t parse_fastq(Record &r) {
        char *base = buf_.data() + pos_, *lim = buf_.data() + len_;
        const char *p = base;
        while (p < lim && (*p == '\n' || *p == '\r')) ++p;  // blank lines between records
        if (p >= lim) return eof_ ? (size_t)(p - base) * 0 : 0;
        if (*p != '@') die("Expected @ at record start");
        const char *eol, *nx = line_end(p, eol);
        if (!nx) return 0;
        const char *hb = p, *he = eol;
        strip_cr(hb, he);
        // sequence lines up to the '+' separator (one line in practice)
        const char *sb = nx, *q = nx;
        std::vector<std::pair<const char *, const char *>> extra;  // multi-line pieces beyond the first
        const char *s0e = nullptr;
        for (;;) {
            if (q >= lim) {
                if (!eof_) return 0;
                die("Incomplete FASTQ record");
            }
            if (*q == '+') break;
            const char *e2, *n2 = line_end(q, e2);
            if (!n2) return 0;
            strip_cr(q, e2);
            if (!s0e) s0e = e2;
            else extra.emplace_back(q, e2);
            q = n2;
        }
        if (!s0e) s0e = sb;
        const char *e3, *n3 = line_end(q, e3);  // '+' line
        if (!n3) return 0;
        size_t seq_len = (size_t)(s0e - sb);
        for (auto &pc : extra) seq_len += (size_t)(pc.second - pc.first);
        // quality lines until they cover the sequence
        const char *qb = n3, *qq = n3, *q0e = nullptr;
        std::vector<std::pair<const char *, const char *>> qextra;
        size_t qual_len = 0;
        while (qual_len < seq_len) {
            if (qq >= lim) {
                if (!eof_) return 0;
                break;
            }
            const char *e4, *n4 = line_end(qq, e4);
            if (!n4) return 0;
            strip_cr(qq, e4);
            if (!q0e) q0e = e4;
            else qextra.emplace_back(qq, e4);
            qual_len += (size_t)(e4 - qq);
            qq = n4;
        }
        if (seq_len == 0 && qq < lim && *qq != '@') {  // empty sequence still has an (empty) quality line
            const char *e4, *n4 = line_end(qq, e4);
            if (!n4) return 0;
            qq = n4;
        }
        // the record is complete: compact multi-line pieces in place (rare)
        char *w = const_cast<char *>(s0e);
        for (auto &pc : extra) {
            memmove(w, pc.first, (size_t)(pc.second - pc.first));
            w += pc.second - pc.first;
        }
        char *wq = const_cast<char *>(q0e ? q0e : qb);
        for (auto &pc : qextra) {
            memmove(wq, pc.first, (size_t)(pc.second - pc.first));
            wq += pc.second - pc.first;
        }
        set_id(r, hb, he);
        r.seq = sb;
        r.seq_len = (uint32_t)seq_len;
        r.qual = qb;
        r.has_qual = true;
        return (size_t)(qq - base);
    }
    size_t parse_fasta(Record &r) {
        char *base = buf_.data() + pos_, *lim = buf_.data() + len_;
        const char *p = base;
        while (p < lim && (*p == '\n' || *p == '\r')) ++p;
        if (p >= lim) return 0;
        if (*p != '>') die("Expected > at record start.");
        const char *eol, *nx = line_end(p, eol);
        if (!nx) return 0;
        const char *hb = p, *he = eol;
        strip_cr(hb, he);
        // find the next header ("\n>") or the end of the input
        const char *q = nx, *rec_end = nullptr;
        while (q < lim) {
            const char *gt = (const char *)memchr(q, '>', (size_t)(lim - q));
            if (!gt) break;
            if (gt == nx || gt[-1] == '\n') {
                rec_end = gt;
                break;
            }
            q = gt + 1;
        }
        if (!rec_end) {
            if (!eof_) return 0;
            rec_end = lim;
        }
        // compact the sequence lines in place
        char *w = const_cast<char *>(nx);
        const char *rd = nx;
        while (rd < rec_end) {
            const char *nl = (const char *)memchr(rd, '\n', (size_t)(rec_end - rd));
            const char *le = nl ? nl : rec_end;
            const char *e2 = le;
            strip_cr(rd, e2);
            if (w != rd) memmove(w, rd, (size_t)(e2 - rd));
            w += e2 - rd;
            rd = nl ? nl + 1 : rec_end;
        }
        set_id(r, hb, he);
        r.seq = nx;
        r.seq_len = (uint32_t)(w - nx);
        r.has_qual = false;
        return (size_t)(rec_end - base);
    }

    gzFile gz_ = nullptr;
    Fmt fmt_;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false;
};

class ReadQueue {  // file_parser.rs:227-301, block-wise
  public:
    ReadQueue(const std::string &path, Fmt fmt, size_t buf_bytes) : fmt_(fmt), buf_bytes_(buf_bytes) {
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
    ~ReadQueue() { delete cur_; }
    Fmt peek_format() const { return files_.empty() ? Fmt::Fasta : detect_format(files_.back(), fmt_); }
    // Appends up to max_records records of the CURRENT file (slices valid until the next call).  Returns false
    // when every file is exhausted; `file_done` tells the caller that the current file ended (so the next call
    // moves to another buffer and anything it still needs from this one must be copied).
    bool next_records(std::vector<Record> &out, size_t max_records, bool &file_done) {
        file_done = false;
        for (;;) {
            if (!cur_) {
                if (files_.empty()) return false;
                const std::string f = files_.back();  // popped from the END (file_parser.rs:238)
                files_.pop_back();
                cur_ = new SeqFile(f, detect_format(f, fmt_), buf_bytes_);
            }
            if (cur_->next_records(out, max_records)) return true;
            delete cur_;
            cur_ = nullptr;
            file_done = true;
            return true;
        }
    }

  private:
    std::vector<std::string> files_;
    Fmt fmt_;
    size_t buf_bytes_;
    SeqFile *cur_ = nullptr;
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
    fputc(r.has_qual ? '@' : '>', fp);
    fwrite(id.data(), 1, id.size(), fp);
    fputc('\n', fp);
    fwrite(seq.data(), 1, seq.size(), fp);
    if (r.has_qual) {
        fwrite("\n+\n", 1, 3, fp);
        fwrite(r.qual, 1, r.seq_len, fp);
    }
    fputc('\n', fp);
}

constexpr size_t kParseBufBytes = 256u << 20;

struct PhaseTimer {  // --stats: where the wall time of a run goes
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    std::map<std::string, double> ms;
    void lap(const char *name) {
        auto n = std::chrono::steady_clock::now();
        ms[name] += std::chrono::duration<double, std::milli>(n - t).count();
        t = n;
    }
    void report(uint64_t reads) const {
        double total = 0;
        for (auto &kv : ms) total += kv.second;
        fprintf(stderr, "[stats] reads=%llu total=%.1f ms (%.2f M reads/s)", (unsigned long long)reads, total,
                total > 0 ? reads / total / 1e3 : 0.0);
        for (auto &kv : ms) fprintf(stderr, " %s=%.1f", kv.first.c_str(), kv.second);
        fputc('\n', stderr);
    }
};

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
    ReadQueue q(a.get("genomes"), parse_fmt(a.get("format", "auto")), 64u << 20);
    std::vector<Record> recs;
    bool file_done;
    while (q.next_records(recs, 256, file_done)) {  // one leaf per record (main.rs:173-195)
        for (const Record &r : recs)
            check(pf_builder_insert(b, r.id_str().c_str(), (const uint8_t *)r.seq, r.seq_len), "BloomTree::insert");
        recs.clear();
    }
    check(pf_builder_save(b, a.get("db-path").c_str()), "BloomTree::save");
    pf_builder_free(b);
    puts("Finished.");
    return 0;
}

// ---- query (main.rs:249-376) -------------------------------------------------------------------------------
struct OwnedRecord {
    std::string id, seq, qual;
};

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

    PhaseTimer timer;
    const bool stats = a.flags.count("stats") > 0;
    uint64_t total_reads = 0;
    pf_db *db = nullptr;
    check(pf_db_open(a.get("db-path").c_str(), device, depth, &db), "BloomTree::load");
    timer.lap("db_open");
    if (a.has("hash-rot")) check(pf_db_set_hash_rot(db, atoi(a.get("hash-rot").c_str())), "pf_db_set_hash_rot");

    puts("Querying reads...");
    printf("Filtering settings: positive=%s; negative=%s\n", pos ? "true" : "false", neg ? "true" : "false");
    if (depth >= 0) {
        if (!filtering) puts("If using a search depth, use a filtering flag (--pos-filter or --neg-filter, or both!)");
        printf("Search depth settings: %lld\n", (long long)depth);
    }
    ReadQueue q(a.get("reads"), parse_fmt(a.get("format", "auto")), kParseBufBytes);
    create_and_overwrite_directory(out);
    const char *ext = q.peek_format() == Fmt::Fastq ? "fq" : "fa";
    FILE *pos_fp = nullptr, *neg_fp = nullptr;
    if (pos && !(pos_fp = fopen((out + "/POS_FILTERING." + ext).c_str(), "wb"))) die("cannot create POS_FILTERING");
    if (neg && !(neg_fp = fopen((out + "/NEG_FILTERING." + ext).c_str(), "wb"))) die("cannot create NEG_FILTERING");

    // GPU batches are whole multiples of the reference's block, so block boundaries stay where they were
    const size_t batch_reads = std::max(block, gpu_batch / block * block);
    std::vector<Record> recs;
    std::vector<OwnedRecord> carry;  // records of an unfinished block, copied out of the parse buffer
    std::vector<const uint8_t *> ptrs;
    std::vector<uint32_t> lens;
    pf_packed *packed = nullptr;  // recycled: its pinned buffers are reused by every batch
    std::string seq_up, ext_id;
    std::unordered_map<std::string_view, std::set<uint32_t>> result_map;

    auto process = [&](size_t n) {
        ptrs.resize(n);
        lens.resize(n);
        for (size_t i = 0; i < n; ++i) {
            ptrs[i] = (const uint8_t *)recs[i].seq;  // k-mers come from the raw bytes (file_parser.rs:203-205)
            lens[i] = recs[i].seq_len;
        }
        total_reads += n;
        check(pf_pack_reads_ptrs(ptrs.data(), lens.data(), (uint32_t)n, &packed), "pf_pack_reads");
        timer.lap("pack");
        pf_hits hits{};
        check(pf_query_block(db, pf_packed_batch(packed), theta, filtering ? 1 : 0, &hits), "query_batch");
        timer.lap("gpu");
        if (!filtering) return;
        // ResultMap is keyed by read id and cleared after every block of `block` reads (main.rs:345-364)
        for (size_t b0 = 0; b0 < n; b0 += block) {
            const size_t b1 = std::min(n, b0 + block);
            result_map.clear();  // keys are views into the records of this block
            for (size_t i = b0; i < b1; ++i)
                for (uint64_t j = hits.read_off[i]; j < hits.read_off[i + 1]; ++j)
                    result_map[std::string_view(recs[i].id, recs[i].id_len)].insert(hits.leaf[j]);
            for (size_t i = b0; i < b1; ++i) {
                const Record &r = recs[i];
                auto it = result_map.empty() ? result_map.end() : result_map.find(std::string_view(r.id, r.id_len));
                const bool mapped = it != result_map.end();
                if ((mapped && !pos_fp) || (!mapped && !neg_fp)) continue;
                seq_up.assign(r.seq, r.seq_len);
                for (auto &c : seq_up)
                    if (c >= 'a' && c <= 'z') c = (char)(c - 32);  // to_ascii_uppercase, main.rs:347-349
                if (mapped) {
                    ext_id.assign(r.id, r.id_len);  // get_ext_id, result_map.rs:24-37
                    ext_id += " |";
                    bool first = true;
                    for (uint32_t leaf : it->second) {
                        if (!first) ext_id += ",";
                        ext_id += pf_db_leaf_id(db, leaf);
                        first = false;
                    }
                    write_record(pos_fp, ext_id, seq_up, r);
                } else {
                    ext_id.assign(r.id, r.id_len);
                    write_record(neg_fp, ext_id, seq_up, r);
                }
            }
        }
    };

    for (bool final = false; !final;) {
        bool file_done = false;
        if (!q.next_records(recs, batch_reads - recs.size(), file_done)) final = true;
        timer.lap("read_parse");
        const size_t n = final ? recs.size() : recs.size() / block * block;
        if (n) process(n);
        timer.lap("outputs");
        // what is left (< one block) must outlive the parse buffer: copy it
        std::vector<OwnedRecord> keep(recs.size() - n);
        for (size_t i = n; i < recs.size(); ++i) {
            OwnedRecord &o = keep[i - n];
            o.id.assign(recs[i].id, recs[i].id_len);
            o.seq.assign(recs[i].seq, recs[i].seq_len);
            if (recs[i].has_qual) o.qual.assign(recs[i].qual, recs[i].seq_len);
        }
        std::vector<Record> rest(keep.size());
        for (size_t i = 0; i < keep.size(); ++i) {
            rest[i].id = keep[i].id.data();
            rest[i].id_len = (uint32_t)keep[i].id.size();
            rest[i].seq = keep[i].seq.data();
            rest[i].seq_len = (uint32_t)keep[i].seq.size();
            rest[i].has_qual = recs[n + i].has_qual;
            rest[i].qual = keep[i].qual.data();
        }
        carry.swap(keep);
        recs.swap(rest);
    }
    pf_packed_free(packed);
    if (pos_fp) fclose(pos_fp);
    if (neg_fp) fclose(neg_fp);
    check(pf_save_leaf_counts(db, (out + "/CLASSIFICATION.csv").c_str()), "save_leaf_counts");
    pf_db_close(db);
    timer.lap("finish");
    if (stats) timer.report(total_reads);
    puts("Finished.");
    return 0;
}

// ---- parse: dump the records the reader sees (id<TAB>sequence<TAB>quality); used to test the parser on CPU
int cmd_parse(const Args &a) {
    if (!a.has("reads")) die("the following required arguments were not provided: --reads");
    const size_t buf = strtoull(a.get("buf-bytes", "1048576").c_str(), nullptr, 10);
    const size_t chunk = strtoull(a.get("chunk", "1000").c_str(), nullptr, 10);
    ReadQueue q(a.get("reads"), parse_fmt(a.get("format", "auto")), buf);
    std::vector<Record> recs;
    bool file_done;
    while (q.next_records(recs, chunk, file_done)) {
        for (const Record &r : recs) {
            fwrite(r.id, 1, r.id_len, stdout);
            fputc('\t', stdout);
            fwrite(r.seq, 1, r.seq_len, stdout);
            fputc('\t', stdout);
            if (r.has_qual) fwrite(r.qual, 1, r.seq_len, stdout);
            else fputc('-', stdout);
            fputc('\n', stdout);
        }
        recs.clear();
    }
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
                               {"pos-filter", "neg-filter", "stats"}));
    if (cmd == "parse") return cmd_parse(parse(argc, argv, 2, {{"r", "reads"}, {"F", "format"}}, {}));
    if (cmd == "--version" || cmd == "-V") {
        printf("PhageFilter 2.0 (%s)\n", pf_version());
        return 0;
    }
    die("unrecognized subcommand '" + cmd + "'");
}
