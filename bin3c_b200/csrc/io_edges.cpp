// bin3c_io, part 2: edge arrays -> the text file the reference hands to Infomap (include/bin3c_io.h).
//
// The step AFTER the contact-map hot path.  Replaces nx.write_edgelist(g, path,
// data=['weight'], delimiter=' ') (cluster.py:139-151): one "u v weight" line per undirected edge.
// networkx prints the weight with str(): under the Python 2.7 the reference pins that is '%.12g' with
// '.0' appended to integer-looking values (B3C_FLOAT_STR12); under Python 3 it is repr(), the shortest
// decimal that reads back to the same double (B3C_FLOAT_REPR).  Both layouts are produced here from the
// correctly rounded digits (std::to_chars / glibc printf), following CPython's format_float_short rules.
// Lines are formatted by a pool of threads, a chunk of edges each, and written in order.
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bin3c_io.h"
#include "io_common.h"

namespace {

// digits (no dot, no trailing zeros beyond the first) and the position of the decimal point relative to them
struct Digits {
    char d[24];
    int n, decpt;
};

void strip_zeros(Digits &g) {
    while (g.n > 1 && g.d[g.n - 1] == '0') --g.n;
}

// parse "d.ddddde[+-]xx"
void from_sci(const char *s, const char *e, Digits &g) {
    g.n = 0;
    const char *p = s;
    for (; p < e && *p != 'e'; ++p)
        if (*p != '.') g.d[g.n++] = *p;
    int ex = 0;
    if (p < e) {
        ++p;
        const bool neg = *p == '-';
        if (*p == '-' || *p == '+') ++p;
        for (; p < e; ++p) ex = ex * 10 + (*p - '0');
        if (neg) ex = -ex;
    }
    g.decpt = ex + 1;
    strip_zeros(g);
}

// CPython format_float_short: exponent form when decpt <= -4 or decpt > max_fixed; '.0' on integer-looking values
int layout(const Digits &g, bool neg, int max_fixed, char *out) {
    char *o = out;
    if (neg) *o++ = '-';
    if (g.decpt <= -4 || g.decpt > max_fixed) {
        *o++ = g.d[0];
        if (g.n > 1) {
            *o++ = '.';
            memcpy(o, g.d + 1, g.n - 1);
            o += g.n - 1;
        }
        int ex = g.decpt - 1;
        *o++ = 'e';
        *o++ = ex < 0 ? '-' : '+';
        if (ex < 0) ex = -ex;
        if (ex >= 100) *o++ = (char)('0' + ex / 100);
        *o++ = (char)('0' + ex / 10 % 10);
        *o++ = (char)('0' + ex % 10);
    } else if (g.decpt <= 0) {
        *o++ = '0';
        *o++ = '.';
        for (int i = 0; i < -g.decpt; ++i) *o++ = '0';
        memcpy(o, g.d, g.n);
        o += g.n;
    } else if (g.decpt >= g.n) {
        memcpy(o, g.d, g.n);
        o += g.n;
        for (int i = g.n; i < g.decpt; ++i) *o++ = '0';
        *o++ = '.';
        *o++ = '0';
    } else {
        memcpy(o, g.d, g.decpt);
        o += g.decpt;
        *o++ = '.';
        memcpy(o, g.d + g.decpt, g.n - g.decpt);
        o += g.n - g.decpt;
    }
    return (int)(o - out);
}

int format_weight(double w, int style, char *out) {
    if (std::isnan(w)) {
        memcpy(out, "nan", 3);
        return 3;
    }
    if (std::isinf(w)) {
        const int n = w < 0 ? 4 : 3;
        memcpy(out, w < 0 ? "-inf" : "inf", n);
        return n;
    }
    const bool neg = std::signbit(w);
    const double a = neg ? -w : w;
    Digits g;
    if (a == 0.0) {
        g.d[0] = '0';
        g.n = 1;
        g.decpt = 1;
    } else if (style == B3C_FLOAT_STR12) {
        char tmp[40];
        const int n = snprintf(tmp, sizeof(tmp), "%.11e", a);
        from_sci(tmp, tmp + n, g);
    } else {
        char tmp[40];
        const auto r = std::to_chars(tmp, tmp + sizeof(tmp), a, std::chars_format::scientific);
        from_sci(tmp, r.ptr, g);
    }
    return layout(g, neg, style == B3C_FLOAT_STR12 ? 12 : 16, out);
}

inline char *put_int(char *o, int32_t v) { return std::to_chars(o, o + 12, v).ptr; }

void format_range(const int32_t *u, const int32_t *v, const double *w, int64_t lo, int64_t hi, char sep, int style,
                  std::vector<char> &buf) {
    buf.resize((size_t)(hi - lo) * 60 + 64);             // 11 + 1 + 11 + 1 + <= 26 + 1 per line
    char *o = buf.data();
    for (int64_t i = lo; i < hi; ++i) {
        o = put_int(o, u[i]);
        *o++ = sep;
        o = put_int(o, v[i]);
        *o++ = sep;
        o += format_weight(w[i], style, o);
        *o++ = '\n';
    }
    buf.resize((size_t)(o - buf.data()));
}

}  // namespace

extern "C" {

int32_t b3c_format_weight_fmt(double w, int32_t float_style, char *h_buf, int32_t capacity) {
    if (!h_buf || capacity < 32 || (float_style != B3C_FLOAT_REPR && float_style != B3C_FLOAT_STR12)) return B3C_IO_ERR_ARG;
    const int n = format_weight(w, float_style, h_buf);
    h_buf[n] = '\0';
    return n;
}

int32_t b3c_format_weight(double w, char *h_buf, int32_t capacity) {
    return b3c_format_weight_fmt(w, B3C_FLOAT_REPR, h_buf, capacity);
}

int64_t b3c_edges_write_fmt(const char *path, const int32_t *h_u, const int32_t *h_v, const double *h_w, int64_t n_edges,
                            char sep, int32_t float_style, int32_t n_threads) {
    if (!path || n_edges < 0 || (n_edges > 0 && (!h_u || !h_v || !h_w)) ||
        (float_style != B3C_FLOAT_REPR && float_style != B3C_FLOAT_STR12)) {
        b3cio::set_err("b3c_edges_write: bad argument");
        return B3C_IO_ERR_ARG;
    }
    FILE *fp = fopen(path, "wb");
    if (!fp) {
        b3cio::set_err("%s: cannot open for writing", path);
        return B3C_IO_ERR_OPEN;
    }
    setvbuf(fp, nullptr, _IOFBF, 1 << 22);
    if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    constexpr int64_t CHUNK = 1 << 17;
    const int64_t n_chunks = (n_edges + CHUNK - 1) / CHUNK;
    if (n_chunks < n_threads) n_threads = n_chunks > 0 ? (int32_t)n_chunks : 1;
    std::vector<std::vector<char>> bufs(n_threads);
    int64_t written = 0;
    bool io_ok = true;
    for (int64_t c0 = 0; c0 < n_chunks && io_ok; c0 += n_threads) {
        const int nt = (int)((n_chunks - c0 < n_threads) ? (n_chunks - c0) : n_threads);
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) {
            const int64_t lo = (c0 + t) * CHUNK, hi = lo + CHUNK < n_edges ? lo + CHUNK : n_edges;
            th.emplace_back(format_range, h_u, h_v, h_w, lo, hi, sep, (int)float_style, std::ref(bufs[t]));
        }
        {
            const int64_t lo = c0 * CHUNK, hi = lo + CHUNK < n_edges ? lo + CHUNK : n_edges;
            format_range(h_u, h_v, h_w, lo, hi, sep, (int)float_style, bufs[0]);
        }
        for (auto &t : th) t.join();
        for (int t = 0; t < nt; ++t) {
            if (fwrite(bufs[t].data(), 1, bufs[t].size(), fp) != bufs[t].size()) io_ok = false;
            written += (int64_t)bufs[t].size();
        }
    }
    if (fclose(fp) != 0) io_ok = false;
    if (!io_ok) {
        b3cio::set_err("%s: write failed", path);
        return B3C_IO_ERR_OPEN;
    }
    return written;
}

// ---- narrow pair records (the host side of b3c_accum_add_pairs_packed) ---------------------------------------
int32_t b3c_records_bytes(int64_t n_refs) {
    if (n_refs < (1ll << 19) - 1) return 5;
    if (n_refs < (1ll << 23) - 1) return 6;
    return 8;
}

static void pack_range(const uint64_t *in, int64_t lo, int64_t hi, int B, uint8_t *out) {
    const int tb = (8 * B - 1) / 2;
    const uint64_t tmask = (1ull << tb) - 1ull;
    for (int64_t i = lo; i < hi; ++i) {
        const uint64_t r = in[i];
        uint64_t t1 = r & 0x7fffffffull, t2 = (r >> 32) & 0x7fffffffull;
        if (t1 > tmask) t1 = tmask;
        if (t2 > tmask) t2 = tmask;
        const uint64_t v = t1 | (((r >> 31) & 1ull) << tb) | (t2 << (tb + 1));
        uint8_t *o = out + i * B;
        for (int k = 0; k < B; ++k) o[k] = (uint8_t)(v >> (8 * k));
    }
}

int64_t b3c_records_pack(const uint64_t *h_records, int64_t n, int32_t B, uint8_t *h_out, int32_t n_threads) {
    if (n < 0 || (n > 0 && (!h_records || !h_out)) || (B != 5 && B != 6 && B != 8)) {
        b3cio::set_err("b3c_records_pack: bad argument");
        return B3C_IO_ERR_ARG;
    }
    const int64_t total = (n * B + 7) / 8 * 8;
    if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n < (1 << 16)) n_threads = 1;
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t)
        th.emplace_back(pack_range, h_records, n * t / n_threads, n * (t + 1) / n_threads, (int)B, h_out);
    pack_range(h_records, 0, n / n_threads, (int)B, h_out);
    for (auto &t : th) t.join();
    for (int64_t k = n * B; k < total; ++k) h_out[k] = 0;
    return total;
}

int32_t b3c_records_same_bytes(int64_t n_refs) {
    if (n_refs < (1ll << 23) - 1) return 3;
    if (n_refs < (1ll << 31) - 1) return 4;
    return 0;
}

// a record whose mates lie on the same reference (ids compared after the clamp a narrow record applies)
static inline bool rec_is_same(uint64_t r) { return (r & 0x7fffffffull) == ((r >> 32) & 0x7fffffffull); }

static void split_count(const uint64_t *in, int64_t lo, int64_t hi, int64_t *n_same) {
    int64_t c = 0;
    for (int64_t i = lo; i < hi; ++i) c += rec_is_same(in[i]) ? 1 : 0;
    *n_same = c;
}

static void split_fill(const uint64_t *in, int64_t lo, int64_t hi, int B, int Bs, uint8_t *out_same, int64_t at_same,
                       uint8_t *out_pair, int64_t at_pair) {
    const int tb = (8 * B - 1) / 2, ts = 8 * Bs - 1;
    const uint64_t tmask = (1ull << tb) - 1ull, smask = (1ull << ts) - 1ull;
    for (int64_t i = lo; i < hi; ++i) {
        const uint64_t r = in[i];
        uint64_t t1 = r & 0x7fffffffull, t2 = (r >> 32) & 0x7fffffffull;
        const uint64_t pass = (r >> 31) & 1ull;
        if (t1 == t2) {
            if (t1 > smask) t1 = smask;
            const uint64_t v = t1 | (pass << ts);
            uint8_t *o = out_same + at_same * Bs;
            for (int k = 0; k < Bs; ++k) o[k] = (uint8_t)(v >> (8 * k));
            ++at_same;
        } else {
            if (t1 > tmask) t1 = tmask;
            if (t2 > tmask) t2 = tmask;
            const uint64_t v = t1 | (pass << tb) | (t2 << (tb + 1));
            uint8_t *o = out_pair + at_pair * B;
            for (int k = 0; k < B; ++k) o[k] = (uint8_t)(v >> (8 * k));
            ++at_pair;
        }
    }
}

int64_t b3c_records_split(const uint64_t *h_records, int64_t n, int32_t B, int32_t Bs, uint8_t *h_out_same,
                          uint8_t *h_out_pair, int64_t *h_counts, int32_t n_threads) {
    if (n < 0 || (n > 0 && !h_records) || !h_counts || (B != 5 && B != 6 && B != 8) || (Bs != 3 && Bs != 4) ||
        ((h_out_same == nullptr) != (h_out_pair == nullptr))) {
        b3cio::set_err("b3c_records_split: bad argument");
        return B3C_IO_ERR_ARG;
    }
    if (n_threads <= 0) n_threads = (int32_t)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n < (1 << 16)) n_threads = 1;
    std::vector<int64_t> same(n_threads, 0);
    {
        std::vector<std::thread> th;
        for (int t = 1; t < n_threads; ++t)
            th.emplace_back(split_count, h_records, n * t / n_threads, n * (t + 1) / n_threads, &same[t]);
        split_count(h_records, 0, n / n_threads, &same[0]);
        for (auto &t : th) t.join();
    }
    int64_t n_same = 0;
    for (int t = 0; t < n_threads; ++t) n_same += same[t];
    h_counts[0] = n_same;
    h_counts[1] = n - n_same;
    if (!h_out_same) return n;                          // count only: the caller sizes its buffers
    {
        std::vector<std::thread> th;
        int64_t at_same = 0, at_pair = 0;
        int64_t a0 = 0, p0 = 0;
        for (int t = 0; t < n_threads; ++t) {
            const int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
            if (t == 0) {
                a0 = at_same;
                p0 = at_pair;
            } else {
                th.emplace_back(split_fill, h_records, lo, hi, (int)B, (int)Bs, h_out_same, at_same, h_out_pair, at_pair);
            }
            at_same += same[t];
            at_pair += (hi - lo) - same[t];
        }
        split_fill(h_records, 0, n / n_threads, (int)B, (int)Bs, h_out_same, a0, h_out_pair, p0);
        for (auto &t : th) t.join();
    }
    for (int64_t k = n_same * Bs; k < (n_same * Bs + 7) / 8 * 8; ++k) h_out_same[k] = 0;
    for (int64_t k = (n - n_same) * B; k < ((n - n_same) * B + 7) / 8 * 8; ++k) h_out_pair[k] = 0;
    return n;
}

int64_t b3c_records_unpack(const uint8_t *h_bytes, int64_t n, int32_t B, uint64_t *h_records) {
    if (n < 0 || (n > 0 && (!h_bytes || !h_records)) || (B != 5 && B != 6 && B != 8)) {
        b3cio::set_err("b3c_records_unpack: bad argument");
        return B3C_IO_ERR_ARG;
    }
    const int tb = (8 * B - 1) / 2;
    const uint64_t tmask = (1ull << tb) - 1ull;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t v = 0;
        for (int k = 0; k < B; ++k) v |= (uint64_t)h_bytes[i * B + k] << (8 * k);
        uint64_t t1 = v & tmask, t2 = (v >> (tb + 1)) & tmask;
        if (t1 == tmask) t1 = 0x7fffffffull;
        if (t2 == tmask) t2 = 0x7fffffffull;
        h_records[i] = t1 | (((v >> tb) & 1ull) << 31) | (t2 << 32);
    }
    return n;
}

int64_t b3c_edges_write(const char *path, const int32_t *h_u, const int32_t *h_v, const double *h_w, int64_t n_edges,
                        char sep, int32_t n_threads) {
    return b3c_edges_write_fmt(path, h_u, h_v, h_w, n_edges, sep, B3C_FLOAT_REPR, n_threads);
}

}  // extern "C"
