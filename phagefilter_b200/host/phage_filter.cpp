// phage_filter.cpp -- host driver with the reference's command line on top of the C ABI (include/pfgpu.h).
//
// Mirrors src/main.rs of the reference: subcommands build | add | query with the same flags (main.rs:38-136),
// the same stdout lines (main.rs:181,199,285-297,375), the same on-disk DB and the same output files
// (CLASSIFICATION.csv, POS_FILTERING.{fa,fq}, NEG_FILTERING.{fa,fq}; main.rs:311-374, 394-404).
// The reference is Rust; no Rust toolchain exists in this image, so the host side is C++ (INTEGRATION.md shows
// the Rust binding).  File parsing stays on the host (file_parser.rs); everything else runs in libpfgpu.
//
// Differences that cannot change results: --threads and --cache-size are accepted and ignored (all filters are
// resident in HBM; host threads are set with --host-threads / PF_HOST_THREADS); reads go to the GPU in batches
// larger than --block-size-reads, while ResultMap's per-block, id-keyed semantics (main.rs:345-364) are kept on the
// host.  An ingest thread (seq_reader.h: parallel read + parse, then the 2-bit packer) runs one chunk ahead of the
// query thread; POS/NEG records are formatted on several threads (outputs.h).
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pfgpu.h"
#include "outputs.h"
#include "seq_reader.h"

namespace {

using namespace pfhost;

void check(int rc, const char *what) {
    if (rc != PF_OK) die(std::string(what) + ": " + pf_last_error());
}

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

constexpr size_t kParseBufBytes = 256u << 20;

struct PhaseTimer {  // --stats: where the time of a run goes (one timer per host thread; stages overlap)
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    std::map<std::string, double> ms;
    void lap(const char *name) {
        auto n = std::chrono::steady_clock::now();
        ms[name] += std::chrono::duration<double, std::milli>(n - t).count();
        t = n;
    }
    void print(const char *who) const {
        fprintf(stderr, " | %s:", who);
        for (auto &kv : ms) fprintf(stderr, " %s=%.1f", kv.first.c_str(), kv.second);
    }
};

// ---- ingest pipeline ---------------------------------------------------------------------------------------
// A producer thread reads, parses (seq_reader.h) and 2-bit-packs (pf_pack_reads_ptrs) one chunk of the input
// while the caller's thread queries and writes the outputs of the chunk before: file_parser.rs's ReadQueue with
// the reading taken off the query thread.  A chunk is the content of one parse buffer (more if a block of
// `block` reads does not fit in one), cut to whole blocks; the rest (< block records) is carried into the next
// chunk, so block boundaries stay where the reference puts them (main.rs:322-364), also across files.
struct Chunk {
    std::vector<RawBuf> bufs;  // parse buffers the records are slices of
    size_t used_bufs = 0;
    std::vector<char> carry;         // bytes of the records carried over from the chunk before
    std::vector<Record> recs;        // carried records first, then this chunk's
    size_t n = 0;                    // recs[0, n) are processed with this chunk
    std::vector<pf_packed *> packed;  // one per GPU batch of <= batch_reads reads; pinned buffers recycled
    size_t n_batches = 0;
    bool last = false;
    ~Chunk() {
        for (pf_packed *p : packed) pf_packed_free(p);
    }
};

template <class T>
class Channel {
  public:
    void push(T v) {
        {
            std::lock_guard<std::mutex> g(mu_);
            q_.push_back(v);
        }
        cv_.notify_one();
    }
    T pop() {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return !q_.empty(); });
        T v = q_.front();
        q_.pop_front();
        return v;
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<T> q_;
};

class Ingest {
  public:
    // block: the reference's --block-size-reads; batch_reads: reads per GPU call (a multiple of block);
    // pack: also build the 2-bit batches (the `parse` test hook can switch it off)
    Ingest(const std::string &path, Fmt fmt, size_t buf_bytes, size_t block, size_t batch_reads, bool pack, int threads,
           int device = 0, size_t min_segment = 256u << 10)
        : pool_(threads), q_(path, fmt, &pool_), buf_bytes_(buf_bytes), block_(block), batch_reads_(batch_reads), pack_(pack), device_(device) {
        q_.set_min_segment(min_segment);
        first_format_ = q_.peek_format();  // before the producer starts popping files
        for (auto &c : chunks_) free_.push(&c);
    }
    ~Ingest() {
        if (thread_.joinable()) thread_.join();
    }
    Fmt peek_format() const { return first_format_; }  // format of the first file read (main.rs:314-315)
    void start() {
        thread_ = std::thread([this] { produce(); });
    }
    Chunk *next() { return ready_.pop(); }       // blocks until the producer has a chunk; the last one has ->last set
    void release(Chunk *c) { free_.push(c); }    // hand the chunk back once its records are no longer needed
    const PhaseTimer &timer() const { return timer_; }  // read after the last chunk was released

  private:
    void produce() {
        if (pack_) check(pf_thread_set_device(device_), "pf_thread_set_device");  // pinned batches belong to the handle's GPU
        std::vector<char> pend_bytes;  // records of an unfinished block, copied out of the chunk they came from
        std::vector<Record> pend_recs;
        std::vector<const uint8_t *> ptrs;
        std::vector<uint32_t> lens;
        for (bool more = true; more;) {
            Chunk *c = free_.pop();
            c->used_bufs = 0;
            c->carry.swap(pend_bytes);  // swapping keeps the data pointers the records hold
            c->recs.swap(pend_recs);
            pend_bytes.clear();
            pend_recs.clear();
            timer_.lap("wait");
            // at least one block; beyond that, about one parse buffer's worth of input (small files are gathered)
            size_t chunk_bytes = 0;
            do {
                if (c->bufs.size() <= c->used_bufs) c->bufs.emplace_back();
                const size_t before = c->recs.size();
                more = q_.next_records(c->bufs[c->used_bufs], buf_bytes_, c->recs);
                chunk_bytes += q_.last_bytes();
                if (c->recs.size() > before) ++c->used_bufs;
            } while (more && (c->recs.size() < block_ || (c->recs.size() < batch_reads_ && chunk_bytes < buf_bytes_)));
            timer_.lap("read_parse");
            const size_t n = more ? c->recs.size() / block_ * block_ : c->recs.size();
            c->n = n;
            c->last = !more;
            // what is left (< one block) must outlive this chunk's buffers: copy it
            size_t bytes = 0;
            for (size_t i = n; i < c->recs.size(); ++i) bytes += (size_t)c->recs[i].id_len + c->recs[i].seq_len + (c->recs[i].has_qual ? c->recs[i].qual_len : 0);
            pend_bytes.resize(bytes);
            char *w = pend_bytes.data();
            for (size_t i = n; i < c->recs.size(); ++i) {
                const Record &r = c->recs[i];
                Record o = r;
                memcpy(w, r.id, r.id_len), o.id = w, w += r.id_len;
                memcpy(w, r.seq, r.seq_len), o.seq = w, w += r.seq_len;
                if (r.has_qual) memcpy(w, r.qual, r.qual_len), o.qual = w, w += r.qual_len;
                pend_recs.push_back(o);
            }
            c->recs.resize(n);
            c->n_batches = 0;
            if (pack_) {
                for (size_t lo = 0; lo < n; lo += batch_reads_) {
                    const size_t cnt = std::min(batch_reads_, n - lo);
                    ptrs.resize(cnt);
                    lens.resize(cnt);
                    const size_t T = std::min<size_t>((size_t)pool_.size(), cnt / 65536 + 1);
                    pool_.run((int)T, [&](int t) {
                        for (size_t i = cnt * (size_t)t / T; i < cnt * ((size_t)t + 1) / T; ++i) {
                            ptrs[i] = (const uint8_t *)c->recs[lo + i].seq;  // k-mers come from the raw bytes (file_parser.rs:203-205)
                            lens[i] = c->recs[lo + i].seq_len;
                        }
                    });
                    if (c->packed.size() <= c->n_batches) c->packed.push_back(nullptr);
                    check(pf_pack_reads_ptrs(ptrs.data(), lens.data(), (uint32_t)cnt, &c->packed[c->n_batches]), "pf_pack_reads");
                    if (!reserved_) {  // page-lock the other chunks' first batch now, before any query runs beside it
                        reserved_ = true;
                        for (auto &o : chunks_)
                            if (&o != c) {
                                o.packed.resize(1, nullptr);
                                check(pf_packed_reserve_like(&o.packed[0], c->packed[0]), "pf_packed_reserve_like");
                            }
                    }
                    ++c->n_batches;
                }
                timer_.lap("pack");
            }
            ready_.push(c);
        }
    }

    Pool pool_;
    ReadQueue q_;
    size_t buf_bytes_, block_, batch_reads_;
    bool pack_, reserved_ = false;
    int device_;
    Fmt first_format_ = Fmt::Fasta;
    Chunk chunks_[3];
    Channel<Chunk *> free_, ready_;
    std::thread thread_;
    PhaseTimer timer_;
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
    Pool pool(a.has("host-threads") ? std::max(1, atoi(a.get("host-threads").c_str())) : default_host_threads());
    ReadQueue q(a.get("genomes"), parse_fmt(a.get("format", "auto")), &pool);
    RawBuf buf;
    std::vector<Record> recs;
    while (q.next_records(buf, 64u << 20, recs)) {  // one leaf per record (main.rs:173-195)
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
    const int host_threads = a.has("host-threads") ? std::max(1, atoi(a.get("host-threads").c_str())) : default_host_threads();
    if (block == 0) die("block size must be positive");
    setenv("PF_PACK_THREADS", std::to_string(host_threads).c_str(), 0);  // the packer's threads follow --host-threads

    const auto t_start = std::chrono::steady_clock::now();
    PhaseTimer timer;
    const bool stats = a.flags.count("stats") > 0;
    uint64_t total_reads = 0;
    // GPU batches are whole multiples of the reference's block, so block boundaries stay where they were
    const size_t batch_reads = std::max(block, gpu_batch / block * block);
    const size_t buf_bytes = getenv("PF_PARSE_BUF_BYTES") ? strtoull(getenv("PF_PARSE_BUF_BYTES"), nullptr, 10) : kParseBufBytes;  // tests: many chunks
    auto ingest = std::make_unique<Ingest>(a.get("reads"), parse_fmt(a.get("format", "auto")), buf_bytes, block, batch_reads, true, host_threads, device);
    // CUDA start-up and the database load run alone: started before them, the ingest threads' page faults and pinned
    // allocations contend with the driver's own memory mapping (measured: start-up up to 3.8 s instead of ~1 s)
    pf_db *db = nullptr;
    check(pf_db_open(a.get("db-path").c_str(), device, depth, &db), "BloomTree::load");
    timer.lap("db_open");
    ingest->start();
    if (a.has("hash-rot")) check(pf_db_set_hash_rot(db, atoi(a.get("hash-rot").c_str())), "pf_db_set_hash_rot");
    pf_db_info_t info{};
    check(pf_db_info(db, &info), "pf_db_info");
    std::vector<std::string> leaf_ids(info.n_leaves);
    for (uint64_t l = 0; l < info.n_leaves; ++l) leaf_ids[l] = pf_db_leaf_id(db, l);

    puts("Querying reads...");
    printf("Filtering settings: positive=%s; negative=%s\n", pos ? "true" : "false", neg ? "true" : "false");
    if (depth >= 0) {
        if (!filtering) puts("If using a search depth, use a filtering flag (--pos-filter or --neg-filter, or both!)");
        printf("Search depth settings: %lld\n", (long long)depth);
    }
    create_and_overwrite_directory(out);
    const char *ext = ingest->peek_format() == Fmt::Fastq ? "fq" : "fa";
    FilterFiles files(out, ext, pos, neg);
    Pool out_pool(filtering ? host_threads : 1);
    FilterWriter writer(files.pos_fd(), files.neg_fd(), block, &leaf_ids, &out_pool);
    timer.lap("setup");

    for (bool last = false; !last;) {
        Chunk *c = ingest->next();
        timer.lap("wait_ingest");
        for (size_t j = 0; j < c->n_batches; ++j) {
            const size_t lo = j * batch_reads, cnt = std::min(batch_reads, c->n - lo);
            pf_hits hits{};
            check(pf_query_block(db, pf_packed_batch(c->packed[j]), theta, filtering ? 1 : 0, &hits), "query_batch");
            timer.lap("gpu");
            // ResultMap is keyed by read id and cleared after every block of `block` reads (main.rs:345-364)
            if (filtering) writer.write(c->recs.data() + lo, cnt, hits.read_off, hits.leaf);
            timer.lap("outputs");
        }
        total_reads += c->n;
        last = c->last;
        ingest->release(c);
    }
    files.close_all();
    check(pf_save_leaf_counts(db, (out + "/CLASSIFICATION.csv").c_str()), "save_leaf_counts");
    const PhaseTimer ingest_timer = ingest->timer();
    timer.lap("finish");
    if (stats) {
        const double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
        fprintf(stderr, "[stats] reads=%llu total=%.1f ms (%.2f M reads/s) host_threads=%d", (unsigned long long)total_reads, total,
                total > 0 ? total_reads / total / 1e3 : 0.0, host_threads);
        timer.print("query thread");
        ingest_timer.print("ingest thread");
        fputc('\n', stderr);
    }
    puts("Finished.");
    // Every output is on disk.  Unpinning host buffers and freeing device memory one allocation at a time costs
    // 0.05-2 s (measured) for nothing: the process ends here and the driver reclaims everything at once.
    fflush(stdout);
    fflush(stderr);
    _exit(0);
}

// ---- parse: run the ingest pipeline alone and dump the records it hands out (id<TAB>sequence<TAB>quality);
// tests the reader, the chunking and the packer's view of the records on a machine without a GPU
int cmd_parse(const Args &a) {
    if (!a.has("reads")) die("the following required arguments were not provided: --reads");
    const size_t buf = strtoull(a.get("buf-bytes", "1048576").c_str(), nullptr, 10);
    const size_t block = strtoull(a.get("block-size-reads", "100").c_str(), nullptr, 10);
    const size_t gpu_batch = strtoull(a.get("gpu-batch-reads", "1000000").c_str(), nullptr, 10);
    const size_t min_segment = strtoull(a.get("min-segment", "262144").c_str(), nullptr, 10);
    const int host_threads = a.has("host-threads") ? std::max(1, atoi(a.get("host-threads").c_str())) : default_host_threads();
    const bool count_only = a.flags.count("count") > 0, pack = a.flags.count("pack") > 0;
    if (block == 0) die("block size must be positive");
    const size_t batch_reads = std::max(block, gpu_batch / block * block);
    const auto t0 = std::chrono::steady_clock::now();
    Ingest ingest(a.get("reads"), parse_fmt(a.get("format", "auto")), buf, block, batch_reads, pack, host_threads, 0, min_segment);
    ingest.start();
    uint64_t n_reads = 0, n_bases = 0, n_chunks = 0;
    for (bool last = false; !last;) {
        Chunk *c = ingest.next();
        ++n_chunks;
        if (!c->last && (c->n == 0 || c->n % block)) die("a chunk that is not the last must hold whole blocks");
        if (pack) {  // the packed batches must describe exactly the records handed out
            size_t at = 0;
            for (size_t j = 0; j < c->n_batches; ++j) {
                const pf_read_batch *b = pf_packed_batch(c->packed[j]);
                for (uint32_t i = 0; i < b->n_reads; ++i, ++at)
                    if (at >= c->n || b->lengths[i] != c->recs[at].seq_len) die("packed batch does not match the records");
            }
            if (at != c->n) die("packed batches do not cover the chunk");
        }
        for (size_t i = 0; i < c->n; ++i) {
            const Record &r = c->recs[i];
            n_bases += r.seq_len;
            if (count_only) continue;
            fwrite(r.id, 1, r.id_len, stdout);
            fputc('\t', stdout);
            fwrite(r.seq, 1, r.seq_len, stdout);
            fputc('\t', stdout);
            if (r.has_qual) fwrite(r.qual, 1, r.qual_len, stdout);
            else fputc('-', stdout);
            fputc('\n', stdout);
        }
        n_reads += c->n;
        last = c->last;
        ingest.release(c);
    }
    if (count_only) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        printf("reads=%llu bases=%llu chunks=%llu ms=%.1f host_threads=%d\n", (unsigned long long)n_reads, (unsigned long long)n_bases,
               (unsigned long long)n_chunks, ms, host_threads);
        ingest.timer().print("ingest thread");
        fputc('\n', stderr);
    }
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    // clap-verbosity-flag is a global argument of the reference (main.rs:38-44): `phage_filter -vv query ...` is valid
    auto is_verbosity = [](const std::string &t) {
        if (t == "--verbose" || t == "--quiet") return true;
        if (t.size() < 2 || t[0] != '-' || t[1] == '-') return false;
        return t.find_first_not_of(t[1] == 'v' ? 'v' : 'q', 1) == std::string::npos && (t[1] == 'v' || t[1] == 'q');
    };
    while (argc >= 2 && is_verbosity(argv[1])) {
        ++argv;
        --argc;
    }
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
    if (cmd == "parse") return cmd_parse(parse(argc, argv, 2, {{"r", "reads"}, {"F", "format"}, {"b", "block-size-reads"}}, {"count", "pack"}));
    if (cmd == "--version" || cmd == "-V") {
        printf("PhageFilter 2.0 (%s)\n", pf_version());
        return 0;
    }
    die("unrecognized subcommand '" + cmd + "'");
}
