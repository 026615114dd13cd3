// Test stub of the database/query entry points of include/pfgpu.h, preloaded (LD_PRELOAD) in front of libpfgpu.so so that
// the `phage_filter query` driver can be run end to end on a machine without a GPU: it checks the driver's glue (argument
// handling, ingest thread, GPU batches, block semantics, POS/NEG/CLASSIFICATION files, stdout, exit) -- never results of the
// real query path.  Hits are fabricated by the rule of tests/host/filter_writer_harness.cpp (global read index g: g % 4 == 0
// -> no hit, else g % 3 hits at leaves (7g + 5j) % 9).  The packer entry points are NOT stubbed: they come from libpfgpu.so.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "../../include/pfgpu.h"

namespace {
constexpr uint32_t kLeaves = 9;
uint64_t g_read = 0, g_counts[kLeaves];
std::vector<uint64_t> g_off;
std::vector<uint32_t> g_leaf;
std::string g_names[kLeaves];
}  // namespace

extern "C" {
const char *pf_last_error(void) { return "stub"; }
int pf_db_open(const char *db_path, int, int64_t, pf_db **out) {
    if (!db_path || !*db_path || std::string(db_path).find("missing") != std::string::npos) return PF_ERR_IO;
    for (uint32_t l = 0; l < kLeaves; ++l) g_names[l] = "genome_" + std::to_string(l);
    if (getenv("PF_STUB_DUP_NAMES")) g_names[1] = g_names[5];  // two leaves with one tax_id (a FASTA id added twice)
    *out = reinterpret_cast<pf_db *>(&g_read);
    return PF_OK;
}
int pf_db_info(const pf_db *, pf_db_info_t *o) {
    memset(o, 0, sizeof *o);
    o->n_leaves = kLeaves;
    return PF_OK;
}
const char *pf_db_leaf_id(const pf_db *, uint64_t l) { return l < kLeaves ? g_names[l].c_str() : nullptr; }
int pf_db_set_hash_rot(pf_db *, int) { return PF_OK; }
int pf_query_block(pf_db *, const pf_read_batch *in, float, int want_hits, pf_hits *out) {
    g_off.assign((size_t)in->n_reads + 1, 0);
    g_leaf.clear();
    for (uint32_t i = 0; i < in->n_reads; ++i, ++g_read) {
        const uint64_t g = g_read, cnt = g % 4 == 0 ? 0 : g % 3;
        std::set<uint32_t> s;
        for (uint64_t j = 0; j < cnt; ++j) s.insert((uint32_t)((g * 7 + j * 5) % kLeaves));
        for (uint32_t l : s) {
            ++g_counts[l];
            if (want_hits) g_leaf.push_back(l);
        }
        g_off[i + 1] = g_leaf.size();
    }
    out->n_hits = g_leaf.size();
    out->read_off = g_off.data();
    out->leaf = g_leaf.data();
    return PF_OK;
}
int pf_save_leaf_counts(pf_db *, const char *csv_path) {
    FILE *fp = fopen(csv_path, "wb");
    if (!fp) return PF_ERR_IO;
    for (uint32_t l = 0; l < kLeaves; ++l)
        if (g_counts[l]) fprintf(fp, "%s,%llu\n", g_names[l].c_str(), (unsigned long long)g_counts[l]);
    fclose(fp);
    return PF_OK;
}
void pf_db_close(pf_db *) {}

// build / add: the stub "database" is a log of what the driver asked for, written at save time
static std::string g_log;
int pf_builder_create(uint64_t k, float fpr, uint32_t largest, uint64_t s1, uint64_t s2, int, int name_mode, uint64_t, pf_builder **out) {
    char b[256];
    snprintf(b, sizeof b, "create k=%llu fpr=%g largest=%u seeds=%llu,%llu names=%d\n", (unsigned long long)k, (double)fpr, largest,
             (unsigned long long)s1, (unsigned long long)s2, name_mode);
    g_log = b;
    *out = reinterpret_cast<pf_builder *>(&g_log);
    return PF_OK;
}
int pf_builder_open(const char *db_path, int, pf_builder **out) {
    if (std::string(db_path).find("missing") != std::string::npos) return PF_ERR_IO;
    g_log = std::string("open ") + db_path + "\n";
    *out = reinterpret_cast<pf_builder *>(&g_log);
    return PF_OK;
}
int pf_builder_insert(pf_builder *, const char *id, const uint8_t *seq, uint64_t len) {
    g_log += std::string("insert ") + id + " " + std::string(reinterpret_cast<const char *>(seq), (size_t)len) + "\n";
    return PF_OK;
}
int pf_builder_save(pf_builder *, const char *db_path) {
    FILE *fp = fopen(db_path, "wb");  // the test passes a file path as --db-path
    if (!fp) return PF_ERR_IO;
    fwrite(g_log.data(), 1, g_log.size(), fp);
    fclose(fp);
    return PF_OK;
}
void pf_builder_free(pf_builder *) {}
}
