// Test harness for libbin3c_io's sources under ThreadSanitizer / AddressSanitizer+UBSan (tests/test_io.py builds it
// together with io_bam.cpp and io_edges.cpp, so the library code itself is instrumented).
//   io_sanitize <bam> <threads> <min_mapq> <strong> <edges_out>     prints: <status> <n_records> <xor of records> <n_edges_bytes>
// Reads every pair record in odd-sized chunks, packs / unpacks them in 5-byte records when the table allows, and
// writes a small edge file from them; a malformed file must end with a negative status, not with a sanitizer report.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/bin3c_io.h"

int main(int argc, char **argv) {
    if (argc < 6) return 2;
    b3c_bam *h = nullptr;
    int rc = b3c_bam_open(argv[1], atoi(argv[2]), 0, &h);
    if (rc != 0) {
        printf("%d 0 0 0\n", rc);
        return 0;
    }
    b3c_bam_set_filter(h, atoi(argv[3]), atoi(argv[4]), 0, nullptr, 0);
    const int32_t n_refs = b3c_bam_n_refs(h);
    std::vector<uint64_t> all, buf(1237);
    int64_t n;
    while ((n = b3c_bam_read_pairs(h, buf.data(), (int64_t)buf.size())) > 0) all.insert(all.end(), buf.begin(), buf.begin() + n);
    int64_t stats[8];
    b3c_bam_stats(h, stats, 8);
    b3c_bam_close(h);
    if (n < 0) {
        printf("%" PRId64 " %zu 0 0\n", n, all.size());
        return 0;
    }
    uint64_t x = 0;
    for (uint64_t r : all) x ^= r * 0x9e3779b97f4a7c15ull;
    const int B = b3c_records_bytes(n_refs);
    std::vector<uint8_t> packed((all.size() * B + 7) / 8 * 8 + 8);
    std::vector<uint64_t> back(all.size() + 1);
    if (b3c_records_pack(all.data(), (int64_t)all.size(), B, packed.data(), 3) < 0) return 3;
    if (b3c_records_unpack(packed.data(), (int64_t)all.size(), B, back.data()) < 0) return 3;
    for (size_t i = 0; i < all.size(); ++i)
        if (back[i] != all[i]) return 4;
    std::vector<int32_t> u(all.size()), v(all.size());
    std::vector<double> w(all.size());
    for (size_t i = 0; i < all.size(); ++i) {
        u[i] = (int32_t)(all[i] & 0x7fffffff);
        v[i] = (int32_t)((all[i] >> 32) & 0x7fffffff);
        w[i] = 1.0 / (double)(i + 3);
    }
    const int64_t nb = b3c_edges_write_fmt(argv[5], u.data(), v.data(), w.data(), (int64_t)all.size(), ' ', (int)(all.size() & 1), 4);
    printf("0 %zu %" PRIu64 " %" PRId64 "\n", all.size(), x, nb);
    return 0;
}
