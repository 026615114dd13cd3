// CPU harness for the host driver's POS/NEG writer (phagefilter_b200/host/outputs.h): reads records with the
// product's reader, fabricates per-read hit lists by a fixed rule (the GPU is not involved), and writes
// POS_FILTERING / NEG_FILTERING through FilterWriter.  tests/test_host_outputs_cpu.py restates main.rs:345-364
// in Python with the same rule and compares the files byte for byte.
//   usage: harness <reads> <out_dir> <block> <threads> <pos 0|1> <neg 0|1> <n_leaves> <batch_reads>
#include "../../phagefilter_b200/host/outputs.h"

using namespace pfhost;

int main(int argc, char **argv) {
    if (argc != 9) return 2;
    const std::string reads = argv[1], out = argv[2];
    const size_t block = strtoull(argv[3], nullptr, 10);
    const int threads = atoi(argv[4]);
    const bool pos = atoi(argv[5]) != 0, neg = atoi(argv[6]) != 0;
    const uint32_t n_leaves = (uint32_t)atoi(argv[7]);
    const size_t batch_reads = strtoull(argv[8], nullptr, 10);  // a multiple of block, like the driver's GPU batches
    std::vector<std::string> leaf_ids;
    for (uint32_t l = 0; l < n_leaves; ++l) leaf_ids.push_back("genome_" + std::to_string(l));
    create_and_overwrite_directory(out);  // the driver's own helper: stale content must be gone afterwards
    Pool pool(threads);
    ReadQueue q(reads, Fmt::Auto, &pool);
    const char *ext = q.peek_format() == Fmt::Fastq ? "fq" : "fa";
    FilterFiles files(out, ext, pos, neg);
    FilterWriter w(files.pos_fd(), files.neg_fd(), block, &leaf_ids, &pool);
    RawBuf buf;
    std::vector<Record> recs;
    struct stat st;
    if (stat(reads.c_str(), &st) != 0) return 3;
    while (q.next_records(buf, (size_t)st.st_size + 4096, recs)) {}  // one plain file: everything lands in one buffer
    size_t g = 0;  // global read index: the rule depends on it, not on the batch
    for (size_t lo = 0; lo < recs.size(); lo += batch_reads) {
        const size_t n = std::min(batch_reads, recs.size() - lo);
        std::vector<uint64_t> off(n + 1, 0);
        std::vector<uint32_t> leaf;
        for (size_t i = 0; i < n; ++i, ++g) {
            const size_t cnt = g % 4 == 0 ? 0 : g % 3;  // 0, 1 or 2 hits
            std::set<uint32_t> s;
            for (size_t j = 0; j < cnt; ++j) s.insert((uint32_t)((g * 7 + j * 5) % n_leaves));
            for (uint32_t l : s) leaf.push_back(l);  // ascending within a read, like pf_hits
            off[i + 1] = leaf.size();
        }
        w.write(recs.data() + lo, n, off.data(), leaf.data());
    }
    files.close_all();
    return 0;
}
