// Slate ingestion (host code only): svmlight / LETOR text and the reference's baseline-score CSV -> padded, query-major
// arrays the loss kernels consume ([Q, n, F] features, [Q, n] labels with -1 padding, [Q, n, S] baseline scores).
// Replaces utils/dataset.py:33-101 of the reference (get_data: sklearn's load_svmlight_file + a Python loop that starts a
// new query whenever the qid of consecutive lines changes; load_baseline_file: "qid,score,score,..." lines grouped the
// same way).  The file is mapped once; a scan pass finds the line and query boundaries, then the lines are parsed by
// several host threads straight into the caller's (pinned) buffers -- query-sharded callers pass their own query range
// and only those lines are parsed.  No CUDA in this file: the H2D copy is the caller's (torch), like every allocation.
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <thread>
#include <vector>

#include "ltr_common.cuh"
#include "ltr_host.h"

namespace ltr {

struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    char* heap = nullptr;    // a file whose size is a whole number of pages is read into a NUL-terminated heap copy instead:
                             // strtod / strtoll on its last line must find a terminator (a mapping's tail page is zero-filled)
    bool open_file(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        size = (size_t)st.st_size;
        if (size == 0) return true;
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        if (size % page == 0) {
            heap = static_cast<char*>(malloc(size + 1));
            if (!heap) return false;
            size_t got = 0;
            while (got < size) {
                const ssize_t r = ::read(fd, heap + got, size - got);
                if (r <= 0) return false;
                got += (size_t)r;
            }
            heap[size] = 0;
            data = heap;
            return true;
        }
        void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) return false;
        data = static_cast<const char*>(p);
        madvise(p, size, MADV_SEQUENTIAL);
        return true;
    }
    ~MappedFile() {
        if (heap) free(heap);
        else if (data) munmap(const_cast<char*>(data), size);
        if (fd >= 0) ::close(fd);
    }
};

struct Line {
    size_t begin, end;       // [begin, end): without the newline, without a trailing comment
};

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// strtod for the numbers these files are made of.  A decimal with at most 15 significant digits and a power of ten within
// 10^+-22 is w * 10^e or w / 10^-e with BOTH operands exact doubles, so the one IEEE operation returns the correctly rounded
// value -- the same bits strtod returns (Clinger's fast path) -- at a twentieth of its cost (no locale, no big-number
// arithmetic).  Everything else (more digits, larger exponents, nan / inf / hex, no digits at all) goes to strtod itself.
static inline double fast_strtod(const char* s, char** endp) {
    static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* p = s;
    bool neg = false;
    if (*p == '-') {
        neg = true;
        ++p;
    } else if (*p == '+') {
        ++p;
    }
    unsigned long long w = 0;
    int nd = 0, e10 = 0;
    bool any = false;
    while (*p >= '0' && *p <= '9') {
        any = true;
        if (w != 0 || *p != '0') {
            if (nd >= 15) return strtod(s, endp);
            w = w * 10 + (unsigned)(*p - '0');
            ++nd;
        }
        ++p;
    }
    if (*p == '.') {
        ++p;
        while (*p >= '0' && *p <= '9') {
            any = true;
            if (w != 0 || *p != '0') {
                if (nd >= 15) return strtod(s, endp);
                w = w * 10 + (unsigned)(*p - '0');
                ++nd;
            }
            --e10;
            ++p;
        }
    }
    if (!any || *p == 'x' || *p == 'X') return strtod(s, endp);
    if (*p == 'e' || *p == 'E') {
        const char* q = p + 1;
        bool eneg = false;
        if (*q == '-') {
            eneg = true;
            ++q;
        } else if (*q == '+') {
            ++q;
        }
        if (*q >= '0' && *q <= '9') {
            int ev = 0;
            while (*q >= '0' && *q <= '9') {
                if (ev < 10000) ev = ev * 10 + (*q - '0');
                ++q;
            }
            e10 += eneg ? -ev : ev;
            p = q;
        }
    }
    if (e10 < -22 || e10 > 22) return strtod(s, endp);
    double v = (double)w;
    if (e10 < 0) v /= P10[-e10];
    else v *= P10[e10];
    *endp = const_cast<char*>(p);
    return neg ? -v : v;
}
// strtoll for a feature index / query id at a non-space position
static inline long long fast_strtoll(const char* s, char** endp) {
    const char* p = s;
    bool neg = false;
    if (*p == '-') {
        neg = true;
        ++p;
    } else if (*p == '+') {
        ++p;
    }
    if (!(*p >= '0' && *p <= '9')) return strtoll(s, endp, 10);
    unsigned long long v = 0;
    int nd = 0;
    while (*p >= '0' && *p <= '9') {
        if (++nd > 18) return strtoll(s, endp, 10);
        v = v * 10 + (unsigned)(*p - '0');
        ++p;
    }
    *endp = const_cast<char*>(p);
    return neg ? -(long long)v : (long long)v;
}

// non-empty lines of the byte range [from, to) of the file: a line belongs to the range its first byte lies in
// (comment-only and blank lines are skipped, as sklearn's parser does)
static void split_range(const MappedFile& f, size_t from, size_t to, std::vector<Line>& lines) {
    size_t pos = from;
    if (from > 0) {                                        // the line that straddles `from` belongs to the range before
        const char* nl = static_cast<const char*>(memchr(f.data + from - 1, '\n', f.size - (from - 1)));
        pos = nl ? (size_t)(nl - f.data) + 1 : f.size;
    }
    while (pos < to && pos < f.size) {
        const char* nl = static_cast<const char*>(memchr(f.data + pos, '\n', f.size - pos));
        const size_t eol = nl ? (size_t)(nl - f.data) : f.size;
        size_t b = pos, e = eol;
        const char* hash = static_cast<const char*>(memchr(f.data + b, '#', e - b));
        if (hash) e = (size_t)(hash - f.data);
        while (b < e && is_space(f.data[b])) ++b;
        while (e > b && is_space(f.data[e - 1])) --e;
        if (e > b) lines.push_back(Line{b, e});
        pos = eol + 1;
    }
}
static void split_lines(const MappedFile& f, std::vector<Line>& lines, int threads = 1) {
    const int T = (threads > 1 && f.size > (1u << 20)) ? (threads > 64 ? 64 : threads) : 1;
    if (T == 1) {
        split_range(f, 0, f.size, lines);
        return;
    }
    std::vector<std::vector<Line>> part(T);
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
        pool.emplace_back([&, t] { split_range(f, f.size * t / T, f.size * (t + 1) / T, part[t]); });
    for (auto& th : pool) th.join();
    size_t total = 0;
    for (auto& v : part) total += v.size();
    lines.reserve(total);
    for (auto& v : part) lines.insert(lines.end(), v.begin(), v.end());
}

// "label qid:ID idx:val ..." -> the qid (or -1 when the line has none); *after_label receives the offset after the label
static long long svm_qid(const char* s, size_t len, double* label, size_t* feat_at) {
    char* endp = nullptr;
    *label = fast_strtod(s, &endp);
    size_t p = (size_t)(endp - s);
    while (p < len && is_space(s[p])) ++p;
    long long qid = -1;
    if (p + 4 <= len && memcmp(s + p, "qid:", 4) == 0) {
        qid = fast_strtoll(s + p + 4, &endp);
        p = (size_t)(endp - s);
    }
    *feat_at = p;
    return qid;
}

struct SvmIndex {
    MappedFile file;
    std::vector<Line> lines;
    std::vector<long long> qid;          // per line
    std::vector<size_t> qstart;          // first line of every query (+ sentinel): a new query starts where the qid changes
    int max_docs = 0;
    long long min_index = -1, max_index = -1;
};

// grouping by CHANGE of qid between consecutive lines (utils/dataset.py:52-65), not by distinct qid
static void group_queries(const std::vector<long long>& qid, std::vector<size_t>& qstart, int* max_docs) {
    qstart.clear();
    *max_docs = 0;
    for (size_t i = 0; i < qid.size(); ++i)
        if (i == 0 || qid[i] != qid[i - 1]) qstart.push_back(i);
    qstart.push_back(qid.size());
    for (size_t q = 0; q + 1 < qstart.size(); ++q) {
        const int d = (int)(qstart[q + 1] - qstart[q]);
        if (d > *max_docs) *max_docs = d;
    }
}

static int host_threads(int requested, size_t work_items) {
    int t = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    if ((size_t)t > work_items) t = work_items ? (int)work_items : 1;
    return t;
}

template <typename Fn>
static void parallel_ranges(size_t n, int threads, Fn fn) {
    if (threads <= 1 || n < 2) {
        fn(0, n, 0);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        pool.emplace_back([=] { fn(lo, hi, t); });
    }
    for (auto& th : pool) th.join();
}

}  // namespace ltr

// Pass 1: how many lines / queries the file holds, the longest query and the feature index range.
//   counts[0] = lines, [1] = queries, [2] = docs in the longest query, [3] = smallest feature index, [4] = largest
LTR_API int ltr_svmlight_scan(const char* path, int threads, long long* counts) {
    using namespace ltr;
    if (!path || !counts) return set_error(LTR_EINVAL, "ltr_svmlight_scan: null argument");
    SvmIndex ix;
    if (!ix.file.open_file(path)) return set_error(LTR_EINVAL, "ltr_svmlight_scan: cannot open / map the file");
    split_lines(ix.file, ix.lines, host_threads(threads, (size_t)1 << 30));
    const size_t L = ix.lines.size();
    ix.qid.resize(L);
    const int T = host_threads(threads, L);
    std::vector<long long> tmin(T, -1), tmax(T, -1);
    const char* d = ix.file.data;
    parallel_ranges(L, T, [&](size_t lo, size_t hi, int t) {
        long long mn = -1, mx = -1;
        for (size_t i = lo; i < hi; ++i) {
            const char* s = d + ix.lines[i].begin;
            const size_t len = ix.lines[i].end - ix.lines[i].begin;
            double label;
            size_t p;
            ix.qid[i] = svm_qid(s, len, &label, &p);
            while (p < len) {                                  // idx:val pairs
                while (p < len && is_space(s[p])) ++p;
                if (p >= len) break;
                char* endp = nullptr;
                const long long idx = fast_strtoll(s + p, &endp);
                p = (size_t)(endp - s);
                if (p < len && s[p] == ':') {
                    if (mn < 0 || idx < mn) mn = idx;
                    if (idx > mx) mx = idx;
                }
                while (p < len && !is_space(s[p])) ++p;        // skip the value
            }
        }
        tmin[t] = mn;
        tmax[t] = mx;
    });
    long long mn = -1, mx = -1;
    for (int t = 0; t < T; ++t) {
        if (tmin[t] >= 0 && (mn < 0 || tmin[t] < mn)) mn = tmin[t];
        if (tmax[t] > mx) mx = tmax[t];
    }
    group_queries(ix.qid, ix.qstart, &ix.max_docs);
    counts[0] = (long long)L;
    counts[1] = (long long)ix.qstart.size() - 1;
    counts[2] = ix.max_docs;
    counts[3] = mn;
    counts[4] = mx;
    return 0;
}

// Pass 2: queries [first_query, first_query + num_queries) -> X [num_queries, n_docs, n_features] (zero-filled, then the
// sparse entries), y [num_queries, n_docs] (pad_value beyond a query's docs), qids [num_queries], lens [num_queries]
// (docs of the query BEFORE truncation to n_docs).  index_base: 1 (LETOR / MSLR files), 0, or -1 = sklearn's "auto"
// (zero-based iff an index 0 occurs).  Entries with a feature index outside [0, n_features) are dropped.
LTR_API int ltr_svmlight_load(const char* path, long long first_query, long long num_queries, int n_docs, int n_features,
                              int index_base, float pad_value, float* X, float* y, long long* qids, int* lens, int threads) {
    using namespace ltr;
    if (!path || !X || !y || first_query < 0 || num_queries < 0 || n_docs <= 0 || n_features <= 0)
        return set_error(LTR_EINVAL, "ltr_svmlight_load: bad arguments");
    SvmIndex ix;
    if (!ix.file.open_file(path)) return set_error(LTR_EINVAL, "ltr_svmlight_load: cannot open / map the file");
    split_lines(ix.file, ix.lines, host_threads(threads, (size_t)1 << 30));
    const size_t L = ix.lines.size();
    ix.qid.resize(L);
    const char* d = ix.file.data;
    const int T = host_threads(threads, L);
    std::vector<int> saw_zero(T, 0);
    parallel_ranges(L, T, [&](size_t lo, size_t hi, int t) {
        for (size_t i = lo; i < hi; ++i) {
            double label;
            size_t p;
            ix.qid[i] = svm_qid(d + ix.lines[i].begin, ix.lines[i].end - ix.lines[i].begin, &label, &p);
            if (index_base < 0 && !saw_zero[t]) {              // "auto": look for a literal index 0
                const char* s = d + ix.lines[i].begin;
                const size_t len = ix.lines[i].end - ix.lines[i].begin;
                for (size_t k = p; k + 1 < len; ++k)
                    if (s[k] == '0' && s[k + 1] == ':' && (k == 0 || is_space(s[k - 1]))) saw_zero[t] = 1;
            }
        }
    });
    int base = index_base;
    if (base < 0) {
        base = 1;
        for (int t = 0; t < T; ++t)
            if (saw_zero[t]) base = 0;
    }
    group_queries(ix.qid, ix.qstart, &ix.max_docs);
    const long long Q = (long long)ix.qstart.size() - 1;
    if (first_query + num_queries > Q) return set_error(LTR_EINVAL, "ltr_svmlight_load: query range beyond the file");
    const size_t row = (size_t)n_features, slate = (size_t)n_docs * row;
    parallel_ranges((size_t)num_queries, host_threads(threads, (size_t)num_queries), [&](size_t lo, size_t hi, int) {
        for (size_t q = lo; q < hi; ++q) {
            const size_t l0 = ix.qstart[first_query + q], l1 = ix.qstart[first_query + q + 1];
            float* Xq = X + q * slate;
            float* yq = y + q * (size_t)n_docs;
            memset(Xq, 0, slate * sizeof(float));
            for (int j = 0; j < n_docs; ++j) yq[j] = pad_value;
            if (qids) qids[q] = ix.qid[l0];
            if (lens) lens[q] = (int)(l1 - l0);
            for (size_t li = l0; li < l1 && (int)(li - l0) < n_docs; ++li) {
                const char* s = d + ix.lines[li].begin;
                const size_t len = ix.lines[li].end - ix.lines[li].begin;
                double label;
                size_t p;
                svm_qid(s, len, &label, &p);
                yq[li - l0] = (float)label;
                float* xr = Xq + (li - l0) * row;
                while (p < len) {
                    while (p < len && is_space(s[p])) ++p;
                    if (p >= len) break;
                    char* endp = nullptr;
                    const long long idx = fast_strtoll(s + p, &endp) - base;
                    p = (size_t)(endp - s);
                    if (p < len && s[p] == ':') {
                        const double v = fast_strtod(s + p + 1, &endp);
                        p = (size_t)(endp - s);
                        if (idx >= 0 && idx < n_features) xr[idx] = (float)v;
                    } else {
                        while (p < len && !is_space(s[p])) ++p;
                    }
                }
            }
        }
    });
    return 0;
}

// Baseline-score CSV ("qid,score_1,...,score_S" per doc, utils/dataset.py:68-101).
//   counts[0] = lines, [1] = queries (by change of qid), [2] = docs in the longest query, [3] = scores per line
LTR_API int ltr_baseline_csv_scan(const char* path, long long* counts) {
    using namespace ltr;
    if (!path || !counts) return set_error(LTR_EINVAL, "ltr_baseline_csv_scan: null argument");
    MappedFile f;
    if (!f.open_file(path)) return set_error(LTR_EINVAL, "ltr_baseline_csv_scan: cannot open / map the file");
    std::vector<Line> lines;
    split_lines(f, lines);
    std::vector<long long> qid(lines.size());
    long long S = 0;
    for (size_t i = 0; i < lines.size(); ++i) {
        qid[i] = strtoll(f.data + lines[i].begin, nullptr, 10);
        if (i == 0)
            for (size_t k = lines[i].begin; k < lines[i].end; ++k) S += f.data[k] == ',';
    }
    std::vector<size_t> qstart;
    int max_docs = 0;
    group_queries(qid, qstart, &max_docs);
    counts[0] = (long long)lines.size();
    counts[1] = (long long)qstart.size() - 1;
    counts[2] = max_docs;
    counts[3] = S;
    return 0;
}

// scores [num_queries, n_docs, S] (zero beyond a query's docs), for queries [first_query, first_query + num_queries)
LTR_API int ltr_baseline_csv_load(const char* path, long long first_query, long long num_queries, int n_docs, int S,
                                  float* scores, int* lens) {
    using namespace ltr;
    if (!path || !scores || first_query < 0 || num_queries < 0 || n_docs <= 0 || S <= 0)
        return set_error(LTR_EINVAL, "ltr_baseline_csv_load: bad arguments");
    MappedFile f;
    if (!f.open_file(path)) return set_error(LTR_EINVAL, "ltr_baseline_csv_load: cannot open / map the file");
    std::vector<Line> lines;
    split_lines(f, lines);
    std::vector<long long> qid(lines.size());
    for (size_t i = 0; i < lines.size(); ++i) qid[i] = strtoll(f.data + lines[i].begin, nullptr, 10);
    std::vector<size_t> qstart;
    int max_docs = 0;
    group_queries(qid, qstart, &max_docs);
    if (first_query + num_queries > (long long)qstart.size() - 1)
        return set_error(LTR_EINVAL, "ltr_baseline_csv_load: query range beyond the file");
    memset(scores, 0, (size_t)num_queries * n_docs * S * sizeof(float));
    for (long long q = 0; q < num_queries; ++q) {
        const size_t l0 = qstart[first_query + q], l1 = qstart[first_query + q + 1];
        if (lens) lens[q] = (int)(l1 - l0);
        for (size_t li = l0; li < l1 && (int)(li - l0) < n_docs; ++li) {
            const char* s = f.data + lines[li].begin;
            const char* end = f.data + lines[li].end;
            const char* p = static_cast<const char*>(memchr(s, ',', (size_t)(end - s)));
            float* dst = scores + ((size_t)q * n_docs + (li - l0)) * S;
            for (int j = 0; j < S && p && p < end; ++j) {
                char* endp = nullptr;
                dst[j] = (float)fast_strtod(p + 1, &endp);
                p = static_cast<const char*>(memchr(endp, ',', (size_t)(end - endp)));
            }
        }
    }
    return 0;
}
