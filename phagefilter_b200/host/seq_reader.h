// seq_reader.h -- host ingest: FASTA/FASTQ/gzip reading and parsing on several host threads.
//
// Mirrors src/file_parser.rs of the reference (format detection :33-86, open_reader :89-101, the record readers
// :191-225 and ReadQueue :227-301) block-wise: a file is read in large stretches, every stretch is cut at record
// boundaries into one segment per thread, the segments are parsed concurrently, and records are handed out as
// slices of the stretch's buffer (nothing is copied per read).
//
// The result is by construction the one a single left-to-right pass gives:
//  * FASTA: every '>' at the start of a line starts a record for the serial parser too, so the cut points are exact.
//  * FASTQ: a cut point is a GUESS ('@' line whose line+2 starts with '+'; exact for 4-line records).  Segments are
//    parsed speculatively (read-only, nothing is reported) and accepted only while each one ends exactly where the
//    next one starts; from the first disagreement (multi-line records, a malformed record, ...) the rest of the
//    stretch is parsed serially by the full parser, which also reports errors at the place a serial pass would.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <dirent.h>
#include <functional>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace pfhost {

[[noreturn]] inline void die(const std::string &msg) {
    // the reference panics (exit code 101) on every error on this path.  _exit: other host threads may be inside
    // CUDA calls or file reads; nothing of theirs is worth flushing after a panic.
    fflush(stdout);
    fprintf(stderr, "thread 'main' panicked: %s\n", msg.c_str());
    fflush(stderr);
    _exit(101);
}

// ---- a small fork-join pool: run(n, fn) executes fn(0..n-1) on the workers and the caller ---------------------
class Pool {
  public:
    explicit Pool(int threads) {
        for (int i = 1; i < threads; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    Pool(const Pool &) = delete;
    Pool &operator=(const Pool &) = delete;
    int size() const { return (int)workers_.size() + 1; }
    // Not re-entrant: one run at a time per pool.  Every worker takes part in every run and run() returns only
    // after all of them have left it, so a late wake-up can never meet the next run's counters.
    void run(int n, const std::function<void(int)> &fn) {
        if (n <= 0) return;
        if (n == 1 || workers_.empty()) {
            for (int i = 0; i < n; ++i) fn(i);
            return;
        }
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            busy_ = (int)workers_.size();
            ++gen_;
        }
        cv_.notify_all();
        work(n, fn);
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work(int n, const std::function<void(int)> &fn) {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n) break;
            fn(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            int n;
            const std::function<void(int)> *fn;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                n = n_;
                fn = fn_;
            }
            work(n, *fn);
            {
                std::lock_guard<std::mutex> g(mu_);
                if (--busy_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, busy_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

inline int default_host_threads() {
    if (const char *e = getenv("PF_HOST_THREADS")) {
        const int v = atoi(e);
        if (v > 0) return std::min(v, 64);
    }
    // two cores stay free for the query thread (kernel launches, per-level syncs) and the CUDA driver's own threads
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min<unsigned>(hc ? hc : 4, 16) - (hc > 4 ? 2 : 0));
}

// ---- file_parser.rs ----------------------------------------------------------------------------------------
enum class Fmt { Auto, Fasta, Fastq };

// A record as slices of a parse buffer (or of a carry arena): nothing is copied per read.
struct Record {
    const char *id = nullptr, *seq = nullptr, *qual = nullptr;
    uint32_t id_len = 0, seq_len = 0, qual_len = 0;  // qual_len != seq_len only in malformed records (kept as read, like bio's reader)
    bool has_qual = false;
    std::string id_str() const { return std::string(id, id_len); }
};

// mmap-backed byte buffer on transparent huge pages where the kernel allows it (madvise): a parse buffer is
// first-touched by every parser thread at once, and 65,536 small-page faults per 256 MB contend for the process's
// mmap lock with the CUDA driver's allocations on the query thread (measured: GPU calls stalled for 100s of ms).
// Growing keeps the contents and never value-initialises.
struct RawBuf {
    char *p = nullptr;
    size_t cap = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr, o.cap = 0; }
    ~RawBuf() {
        if (p) munmap(p, cap);
    }
    void reserve(size_t n) {
        if (n <= cap) return;
        const size_t huge = 2u << 20;
        const size_t want = n >= huge ? (n + huge - 1) / huge * huge : (n + 4095) / 4096 * 4096;
        void *q = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (q == MAP_FAILED) die("out of memory growing a parse buffer");
        if (want >= huge) madvise(q, want, MADV_HUGEPAGE);  // advisory: failure only means small pages
        if (p) {
            memcpy(q, p, cap);
            munmap(p, cap);
        }
        p = (char *)q;
        cap = want;
    }
};

inline std::string lower_ext(const std::string &name) {
    size_t p = name.rfind('.');
    return p == std::string::npos ? "" : name.substr(p + 1);
}
inline std::string stem(const std::string &name) {
    size_t p = name.rfind('.');
    return p == std::string::npos ? name : name.substr(0, p);
}
inline const std::set<std::string> &seq_extensions() {  // file_parser.rs:303
    static const std::set<std::string> k = {"fa", "fasta", "fna", "fsa", "fas", "fq", "fastq"};
    return k;
}
inline bool is_gz_ext(const std::string &e) { return e == "gz" || e == "gzip"; }

inline bool has_supported_extension(const std::string &path) {  // file_parser.rs:323-344
    std::string base = path.substr(path.find_last_of('/') + 1);
    std::string e = lower_ext(base);
    if (e.empty()) return false;
    if (seq_extensions().count(e)) return true;
    if (is_gz_ext(e)) return seq_extensions().count(lower_ext(stem(base))) > 0;
    return false;
}
inline Fmt format_from_extension(const std::string &path) {  // file_parser.rs:69-86
    std::string base = path.substr(path.find_last_of('/') + 1);
    std::string e = lower_ext(base);
    if (is_gz_ext(e)) e = lower_ext(stem(base));
    return (e == "fq" || e == "fastq") ? Fmt::Fastq : Fmt::Fasta;
}
inline Fmt detect_format(const std::string &path, Fmt override_) {  // file_parser.rs:33-66 (gzread passes plain files through)
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

// ---- record parsers over a stretch [base, lim) ---------------------------------------------------------------
struct Span {
    char *base;       // first byte of the stretch
    const char *lim;  // one past its last byte
    bool eof;         // the stretch ends the file (a last line without newline is complete)
};
constexpr size_t kBail = ~(size_t)0;  // speculative parse: "the full parser has to look at this record"

namespace detail {
// [p, end of line); returns the pointer past the newline, or nullptr if the line is not complete in the stretch
inline const char *line_end(const Span &s, const char *p, const char *&eol) {
    const char *nl = (const char *)memchr(p, '\n', (size_t)(s.lim - p));
    if (!nl) {
        if (!s.eof) return nullptr;
        eol = s.lim;  // last line without a newline
        return s.lim;
    }
    eol = nl;
    return nl + 1;
}
inline void strip_cr(const char *b, const char *&e) {
    if (e > b && e[-1] == '\r') --e;
}
inline void set_id(Record &r, const char *b, const char *e) {  // bio: id = header up to the first whitespace
    const char *q = b + 1;
    while (q < e && *q != ' ' && *q != '\t') ++q;
    r.id = b + 1;
    r.id_len = (uint32_t)(q - (b + 1));
}
}  // namespace detail

// One FASTQ record starting at p (first non-blank byte).  Returns the bytes used from p, 0 if the record is not
// complete in the stretch, kBail (speculative only) where the full parser would report an error or has to
// compact a multi-line record in place.  Non-speculative: errors panic like the reference's reader.
template <bool kSpeculative>
inline size_t parse_fastq(const Span &s, char *p0, Record &r) {
    using namespace detail;
    const char *lim = s.lim, *p = p0;
    if (p >= lim) return 0;
    if (*p != '@') {
        if (kSpeculative) return kBail;
        die("Expected @ at record start");
    }
    const char *eol, *nx = line_end(s, p, eol);
    if (!nx) return 0;
    const char *hb = p, *he = eol;
    strip_cr(hb, he);
    // sequence lines up to the '+' separator (one line in practice)
    const char *sb = nx, *q = nx;
    std::vector<std::pair<const char *, const char *>> extra;  // multi-line pieces beyond the first
    const char *s0e = nullptr;
    for (;;) {
        if (q >= lim) {
            if (!s.eof) return 0;
            if (kSpeculative) return kBail;
            die("Incomplete FASTQ record");
        }
        if (*q == '+') break;
        const char *e2, *n2 = line_end(s, q, e2);
        if (!n2) return 0;
        strip_cr(q, e2);
        if (!s0e) s0e = e2;
        else if (kSpeculative) return kBail;
        else extra.emplace_back(q, e2);
        q = n2;
    }
    if (!s0e) s0e = sb;
    const char *e3, *n3 = line_end(s, q, e3);  // '+' line
    if (!n3) return 0;
    size_t seq_len = (size_t)(s0e - sb);
    for (auto &pc : extra) seq_len += (size_t)(pc.second - pc.first);
    // quality lines until they cover the sequence
    const char *qb = n3, *qq = n3, *q0e = nullptr;
    std::vector<std::pair<const char *, const char *>> qextra;
    size_t qual_len = 0;
    while (qual_len < seq_len) {
        if (qq >= lim) {
            if (!s.eof) return 0;
            break;
        }
        const char *e4, *n4 = line_end(s, qq, e4);
        if (!n4) return 0;
        strip_cr(qq, e4);
        if (!q0e) q0e = e4;
        else if (kSpeculative) return kBail;
        else qextra.emplace_back(qq, e4);
        qual_len += (size_t)(e4 - qq);
        qq = n4;
    }
    if (kSpeculative && qual_len != seq_len) return kBail;  // short or over-long quality: the full parser decides
    if (seq_len == 0 && qq < lim && *qq != '@') {  // empty sequence still has an (empty) quality line
        const char *e4, *n4 = line_end(s, qq, e4);
        if (!n4) return 0;
        qq = n4;
    }
    // the record is complete: compact multi-line pieces in place (rare; never in speculative mode)
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
    r.qual_len = (uint32_t)qual_len;
    r.has_qual = true;
    return (size_t)(qq - p0);
}

// One FASTA record starting at p (first non-blank byte); the sequence lines are compacted in place.
// `err` (when given) receives the message instead of panicking -- worker threads hand errors to their caller.
inline size_t parse_fasta(const Span &s, char *p0, Record &r, std::string *err = nullptr) {
    using namespace detail;
    const char *lim = s.lim, *p = p0;
    if (p >= lim) return 0;
    if (*p != '>') {
        if (err) {
            *err = "Expected > at record start.";
            return kBail;
        }
        die("Expected > at record start.");
    }
    const char *eol, *nx = line_end(s, p, eol);
    if (!nx) return 0;
    const char *hb = p, *he = eol;
    strip_cr(hb, he);
    // find the next header (a '>' at the start of a line) or the end of the input
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
        if (!s.eof) return 0;
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
    return (size_t)(rec_end - p0);
}

// First position >= from (from >= 1) that looks like the start of a record, or len if there is none.
inline size_t find_record_start(const char *b, size_t from, size_t len, Fmt fmt) {
    const char *lim = b + len;
    if (from >= len) return len;
    if (fmt == Fmt::Fasta) {
        const char *q = b + from;
        while (q < lim) {
            const char *gt = (const char *)memchr(q, '>', (size_t)(lim - q));
            if (!gt) break;
            if (gt[-1] == '\n') return (size_t)(gt - b);
            q = gt + 1;
        }
        return len;
    }
    const char *nl = (const char *)memchr(b + from - 1, '\n', (size_t)(lim - (b + from - 1)));
    const char *p = nl ? nl + 1 : lim;
    while (p < lim) {
        const char *l1 = (const char *)memchr(p, '\n', (size_t)(lim - p));
        if (!l1) break;
        if (*p == '@') {
            const char *l2 = (const char *)memchr(l1 + 1, '\n', (size_t)(lim - (l1 + 1)));
            if (!l2) break;
            if (l2 + 1 < lim && l2[1] == '+') return (size_t)(p - b);
        }
        p = l1 + 1;
    }
    return len;
}

// Parses every complete record of the stretch (appending to out) and returns the position where parsing stopped
// (start of the first incomplete record, or the end).  pool == nullptr or one thread: plain serial pass.
struct ParseScratch {  // per-segment record lists, kept between stretches (fresh vectors page-fault on every stretch)
    std::vector<std::vector<Record>> seg;
};
inline size_t parse_stretch(const Span &s, Fmt fmt, Pool *pool, std::vector<Record> &out, size_t min_segment = 256u << 10,
                            ParseScratch *scratch = nullptr) {
    char *b = s.base;
    const size_t len = (size_t)(s.lim - s.base);
    auto skip_blank = [&](size_t pos, size_t end) {
        while (pos < end && (b[pos] == '\n' || b[pos] == '\r')) ++pos;  // blank lines between records
        return pos;
    };
    auto serial_from = [&](size_t pos) {
        for (;;) {
            pos = skip_blank(pos, len);
            Record r;
            const size_t used = fmt == Fmt::Fastq ? parse_fastq<false>(s, b + pos, r) : parse_fasta(s, b + pos, r);
            if (!used) return pos;
            pos += used;
            out.push_back(r);
        }
    };
    const int T = pool ? (int)std::min<size_t>((size_t)pool->size(), len / std::max<size_t>(min_segment, 1) + 1) : 1;
    if (T <= 1) return serial_from(0);

    std::vector<size_t> start((size_t)T + 1), stop((size_t)T);
    start[0] = 0;
    start[(size_t)T] = len;
    for (int i = 1; i < T; ++i) start[(size_t)i] = std::max(start[(size_t)i - 1], find_record_start(b, std::max<size_t>(len / (size_t)T * (size_t)i, 1), len, fmt));
    ParseScratch local;
    std::vector<std::vector<Record>> &seg = (scratch ? scratch : &local)->seg;
    if (seg.size() < (size_t)T) seg.resize((size_t)T);
    std::vector<std::string> errs((size_t)T);
    pool->run(T, [&](int i) {
        size_t pos = start[(size_t)i];
        const size_t end = start[(size_t)i + 1];
        auto &recs = seg[(size_t)i];
        recs.clear();
        recs.reserve((end - pos) / 128 + 16);
        for (;;) {
            pos = skip_blank(pos, end);
            if (pos >= end) break;
            Record r;
            const size_t used = fmt == Fmt::Fastq ? parse_fastq<true>(s, b + pos, r) : parse_fasta(s, b + pos, r, &errs[(size_t)i]);
            if (!used || used == kBail) break;
            pos += used;
            recs.push_back(r);
        }
        stop[(size_t)i] = pos;
    });
    // accept segments while each one ends exactly where the next one starts, then copy them out concurrently
    int good = 0;
    size_t cur = 0, total = 0;
    bool handover = false;
    std::vector<size_t> at((size_t)T);
    for (int i = 0; i < T; ++i) {
        if (start[(size_t)i] != cur) {  // the previous segment ran past this guess
            handover = true;
            break;
        }
        at[(size_t)i] = total;
        total += seg[(size_t)i].size();
        cur = stop[(size_t)i];
        good = i + 1;
        if (!errs[(size_t)i].empty() || cur < start[(size_t)i + 1]) {  // error, bailed or incomplete: the full parser takes over
            handover = true;
            break;
        }
    }
    const size_t base = out.size();
    out.resize(base + total);
    pool->run(good, [&](int i) {
        if (!seg[(size_t)i].empty()) memcpy(out.data() + base + at[(size_t)i], seg[(size_t)i].data(), seg[(size_t)i].size() * sizeof(Record));
    });
    for (int i = 0; i < good; ++i)
        if (!errs[(size_t)i].empty()) die(errs[(size_t)i]);
    return handover ? serial_from(cur) : cur;
}

// ---- gzip (and anything that is not a plain regular file): inflated by zlib on a thread of its own ------------------
// zlib inflates one stream at 0.3-0.4 GB/s and cannot be split, so (a) the current file is inflated while the
// stretch before it is parsed, packed and queried, and (b) ReadQueue starts the next few queued gzip files early:
// a directory of N gzip files is inflated on up to kAheadFiles cores.  The consumer sees the same byte stream.
inline bool is_gzip_or_special(const std::string &path) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode) || st.st_size == 0) return true;
    unsigned char magic[2] = {0, 0};
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return true;
    const bool gz = pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    close(fd);
    return gz;
}

class GzAhead {
  public:
    static constexpr size_t kMaxBlocks = 16;  // at most 16 blocks (256 MB) inflated ahead per file
    static size_t block_bytes() {             // inflated bytes per queue entry (PF_GZ_BLOCK: tests use tiny blocks)
        static const size_t v = [] {
            const char *e = getenv("PF_GZ_BLOCK");
            const size_t x = e ? strtoull(e, nullptr, 10) : 0;
            return x ? x : (size_t)(16u << 20);
        }();
        return v;
    }
    explicit GzAhead(const std::string &path) : path_(path) {
        thread_ = std::thread([this] { run(); });
    }
    ~GzAhead() {
        {
            std::lock_guard<std::mutex> g(mu_);
            cancel_ = true;
        }
        space_.notify_all();
        thread_.join();
    }
    GzAhead(const GzAhead &) = delete;
    GzAhead &operator=(const GzAhead &) = delete;
    const std::string &path() const { return path_; }
    // Copies up to want bytes of the inflated stream to dst; less than want only at the end of the file.
    size_t read(char *dst, size_t want) {
        size_t done = 0;
        while (done < want) {
            std::unique_lock<std::mutex> g(mu_);
            data_.wait(g, [this] { return !q_.empty() || finished_; });
            if (q_.empty()) break;  // finished and drained
            Block &b = q_.front();
            const size_t n = std::min(want - done, b.len - b.pos);
            g.unlock();  // the front block is only ever touched by this (single) consumer
            memcpy(dst + done, b.p.get() + b.pos, n);
            done += n;
            g.lock();
            b.pos += n;
            if (b.pos == b.len) {
                spare_.push_back(std::move(b.p));  // recycled: fresh 16 MB allocations page-fault every time
                q_.pop_front();
                space_.notify_one();
            }
        }
        return done;
    }

  private:
    struct Block {
        std::unique_ptr<char[]> p;
        size_t len = 0, pos = 0;
    };
    void run() {
        gzFile gz = gzopen(path_.c_str(), "rb");
        if (!gz) die("Failed to open '" + path_ + "'");
        gzbuffer(gz, 1 << 20);
        for (;;) {
            {
                std::unique_lock<std::mutex> g(mu_);
                space_.wait(g, [this] { return cancel_ || q_.size() < kMaxBlocks; });
                if (cancel_) break;
            }
            const size_t cap = block_bytes();
            Block b;
            {
                std::lock_guard<std::mutex> g(mu_);
                if (!spare_.empty()) {
                    b.p = std::move(spare_.back());
                    spare_.pop_back();
                }
            }
            if (!b.p) b.p.reset(new char[cap]);
            bool end = false;
            while (b.len < cap) {
                const int n = gzread(gz, b.p.get() + b.len, (unsigned)(cap - b.len));
                if (n < 0) die("corrupt gzip input '" + path_ + "'");
                if (n == 0) {
                    end = true;
                    break;
                }
                b.len += (size_t)n;
            }
            std::lock_guard<std::mutex> g(mu_);
            if (b.len) q_.push_back(std::move(b));
            if (end) finished_ = true;
            data_.notify_one();
            if (end) break;
        }
        gzclose(gz);
        std::lock_guard<std::mutex> g(mu_);
        finished_ = true;
        data_.notify_one();
    }
    std::string path_;
    std::mutex mu_;
    std::condition_variable data_, space_;
    std::deque<Block> q_;
    std::vector<std::unique_ptr<char[]>> spare_;
    bool finished_ = false, cancel_ = false;
    std::thread thread_;
};

// ---- one input file, read in large stretches (open_reader: file_parser.rs:89-101) -------------------------------
class SeqFile {
  public:
    // ahead: an inflater ReadQueue already started for this file (gzip), or null
    SeqFile(const std::string &path, Fmt fmt, Pool *pool, ParseScratch *scratch = nullptr, std::unique_ptr<GzAhead> ahead = nullptr)
        : fmt_(fmt), pool_(pool), scratch_(scratch), gz_(std::move(ahead)) {
        // plain regular files are read with pread (several threads); gzip and everything else through zlib
        if (gz_ || is_gzip_or_special(path)) {
            if (!gz_) gz_.reset(new GzAhead(path));
            // first buffer size: sequence text inflates 3-6x; a pipe or device has no size, take the full buffer
            struct stat st;
            gz_guess_ = stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode) ? (size_t)st.st_size * 8 : ~(size_t)0 >> 1;
            return;
        }
        fd_ = open(path.c_str(), O_RDONLY);
        if (fd_ < 0) die("Failed to open '" + path + "'");
        struct stat st;
        if (fstat(fd_, &st) != 0) die("Failed to open '" + path + "'");
        fsize_ = (size_t)st.st_size;
    }
    ~SeqFile() {
        if (fd_ >= 0) close(fd_);
    }
    SeqFile(const SeqFile &) = delete;
    SeqFile &operator=(const SeqFile &) = delete;

    // Reads the next stretch of the file into buf (at least buf_bytes; grown while not even one record fits),
    // parses every complete record in it and appends them to out as slices of buf.  Returns false when the file
    // is exhausted and nothing was appended.
    bool next_records(RawBuf &buf, size_t buf_bytes, std::vector<Record> &out) {
        const size_t before = out.size();
        buf_bytes = std::max<size_t>(buf_bytes, 64);
        // small inputs get small buffers (a directory of many small files must not reserve 256 MB for each)
        size_t want = gz_ ? std::min<size_t>(buf_bytes, std::max<size_t>(8u << 20, gz_guess_)) : std::min(buf_bytes, tail_.size() + (fsize_ - foff_));
        buf.reserve(std::max<size_t>(std::max<size_t>(want, 64), tail_.size() * 2));
        size_t len = tail_.size();
        if (len) memcpy(buf.p, tail_.data(), len);
        tail_.clear();
        last_bytes_ = 0;
        for (;;) {
            const size_t had = len;
            fill(buf, len);
            last_bytes_ += len - had;
            if (gz_ && !eof_ && buf.cap < buf_bytes) {  // gzip: the size is unknown, grow towards buf_bytes first
                buf.reserve(std::min(buf_bytes, buf.cap * 2));
                continue;
            }
            const Span s{buf.p, buf.p + len, eof_};
            const size_t pos = parse_stretch(s, fmt_, pool_, out, min_segment_, scratch_);
            if (out.size() > before || eof_) {
                if (!eof_) tail_.assign(buf.p + pos, buf.p + len);  // at the end only blank lines can be left
                break;
            }
            buf.reserve(buf.cap * 2);  // not even one complete record fits: grow and read more
        }
        return out.size() > before;
    }
    size_t last_bytes() const { return last_bytes_; }  // file bytes the last call consumed
    void set_min_segment(size_t bytes) { min_segment_ = bytes; }  // tests: force many segments on small inputs

  private:
    bool fill(RawBuf &buf, size_t &len) {
        if (eof_ || len >= buf.cap) return false;
        if (gz_) {
            const size_t want = buf.cap - len;
            const size_t n = gz_->read(buf.p + len, want);
            len += n;
            if (n < want) eof_ = true;
            return n > 0;
        }
        const size_t want = std::min(buf.cap - len, fsize_ - foff_);
        const int T = pool_ ? (int)std::min<size_t>((size_t)pool_->size(), want / (4u << 20) + 1) : 1;
        std::atomic<bool> bad{false};
        auto slice = [&](int i) {
            size_t lo = want / (size_t)T * (size_t)i, hi = i == T - 1 ? want : want / (size_t)T * ((size_t)i + 1);
            while (lo < hi) {
                const ssize_t n = pread(fd_, buf.p + len + lo, hi - lo, (off_t)(foff_ + lo));
                if (n <= 0) {
                    bad = true;
                    return;
                }
                lo += (size_t)n;
            }
        };
        if (T > 1) pool_->run(T, slice);
        else slice(0);
        if (bad) die("short read (the input file changed while it was being read)");
        foff_ += want;
        len += want;
        if (foff_ >= fsize_) eof_ = true;
        return want > 0;
    }

    Fmt fmt_;
    Pool *pool_;
    ParseScratch *scratch_;
    int fd_ = -1;
    std::unique_ptr<GzAhead> gz_;
    size_t fsize_ = 0, foff_ = 0, gz_guess_ = 0, min_segment_ = 256u << 10;
    bool eof_ = false;
    size_t last_bytes_ = 0;
    std::vector<char> tail_;  // the incomplete record at the end of the previous stretch
};

// ---- ReadQueue (file_parser.rs:227-301), stretch-wise ----------------------------------------------------------
class ReadQueue {
  public:
    ReadQueue(const std::string &path, Fmt fmt, Pool *pool) : fmt_(fmt), pool_(pool) {
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
    ReadQueue(const ReadQueue &) = delete;
    ReadQueue &operator=(const ReadQueue &) = delete;
    Fmt peek_format() const { return files_.empty() ? Fmt::Fasta : detect_format(files_.back(), fmt_); }
    void set_min_segment(size_t bytes) { min_segment_ = bytes; }
    // Appends the records of the next stretch of the current file to out (slices of buf, valid as long as buf is
    // left alone).  Returns false when every file is exhausted.  A call may append nothing (a file just ended).
    bool next_records(RawBuf &buf, size_t buf_bytes, std::vector<Record> &out) {
        if (!cur_) {
            if (files_.empty()) return false;
            const std::string f = files_.back();  // popped from the END (file_parser.rs:238)
            files_.pop_back();
            std::unique_ptr<GzAhead> mine;
            if (!ahead_.empty() && ahead_.front()->path() == f) {
                mine = std::move(ahead_.front());
                ahead_.pop_front();
            }
            // start inflating the next queued gzip files: ahead_[j] belongs to files_[size - 1 - j]
            const size_t want_ahead = pool_ ? std::min<size_t>((size_t)std::max(pool_->size() - 1, 0), kAheadFiles) : 0;
            for (size_t j = ahead_.size(); j < want_ahead && j < files_.size(); ++j) {
                const std::string &next = files_[files_.size() - 1 - j];
                if (!is_gzip_or_special(next)) break;  // plain files are read in place; the run of inflaters stays contiguous
                ahead_.emplace_back(new GzAhead(next));
            }
            cur_ = new SeqFile(f, detect_format(f, fmt_), pool_, &scratch_, std::move(mine));
            cur_->set_min_segment(min_segment_);
        }
        const bool got = cur_->next_records(buf, buf_bytes, out);
        last_bytes_ = cur_->last_bytes();
        if (!got) {
            delete cur_;
            cur_ = nullptr;
        }
        return true;
    }
    size_t last_bytes() const { return last_bytes_; }

  private:
    std::vector<std::string> files_;
    Fmt fmt_;
    Pool *pool_;
    static constexpr size_t kAheadFiles = 8;
    std::deque<std::unique_ptr<GzAhead>> ahead_;  // inflaters of files_[size-1], files_[size-2], ... in that order
    SeqFile *cur_ = nullptr;
    ParseScratch scratch_;
    size_t min_segment_ = 256u << 10, last_bytes_ = 0;
};

}  // namespace pfhost
