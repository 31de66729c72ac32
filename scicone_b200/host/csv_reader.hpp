// Parallel reader for the reference's count matrices.
//
// The reference parses its inputs with getline + stod, one field at a time on one thread (Utils::read_counts,
// Utils.cpp:74-102): ~80 ns per value, i.e. 15 s for the 10 000 x 18 000 bins matrix of configs[2].  Here the file is
// mapped, cut into pieces at line boundaries, and the pieces are parsed in parallel by whatever `run` is handed in (the
// scoring library's worker pool in the host).  Values come out exactly as strtod would give them: integers and short
// decimals take Clinger's fast path (mantissa < 2^53, at most 22 fraction digits: one correctly rounded division),
// everything else (exponents, long mantissas, inf / nan) goes through strtod itself.
//
// Line semantics are those of the reference's loop: every '\n'-terminated line is a row (an empty line is a row of the
// matrix's fill value), a last line without '\n' counts, fields beyond n_cols are ignored, missing fields keep the fill.
#ifndef SCICONE_HOST_CSV_READER_HPP
#define SCICONE_HOST_CSV_READER_HPP

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <functional>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace scicone_host {

namespace csv_detail {
inline bool is_sep(char c) { return c == ',' || c == ' ' || c == '\r' || c == '\t'; }

// One field starting at p (p < end, *p is not a separator or '\n').  Returns the position behind the field.
inline const char* parse_field(const char* p, const char* end, double* out) {
    static const double pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                     1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* q = p;
    bool neg = false;
    if (*q == '-' || *q == '+') {
        neg = *q == '-';
        ++q;
    }
    uint64_t m = 0;
    int digits = 0, frac = 0;
    bool fast = true;
    while (q < end && *q >= '0' && *q <= '9') {
        if (digits < 19) m = m * 10 + static_cast<uint64_t>(*q - '0');
        ++digits;
        ++q;
    }
    if (q < end && *q == '.') {
        ++q;
        while (q < end && *q >= '0' && *q <= '9') {
            if (digits < 19) m = m * 10 + static_cast<uint64_t>(*q - '0');
            ++digits;
            ++frac;
            ++q;
        }
    }
    if (digits == 0 || digits > 19 || m > (1ull << 53) || frac > 22) fast = false;
    if (fast && q < end && !(is_sep(*q) || *q == '\n')) fast = false;  // exponent, "inf", "nan", hex, garbage
    if (fast) {
        double v = static_cast<double>(m);
        if (frac) v /= pow10[frac];
        *out = neg ? -v : v;
        return q;
    }
    // strtod on a terminated copy of the field (the mapping is not NUL terminated)
    const char* e = p;
    while (e < end && !is_sep(*e) && *e != '\n') ++e;
    char buf[128];
    std::string big;
    const size_t n = static_cast<size_t>(e - p);
    const char* src;
    if (n < sizeof buf) {
        std::memcpy(buf, p, n);
        buf[n] = 0;
        src = buf;
    } else {
        big.assign(p, n);
        src = big.c_str();
    }
    char* stop = nullptr;
    const double v = std::strtod(src, &stop);
    if (stop == src) return nullptr;  // not a number: the rest of the line is ignored (as the serial loop did)
    *out = v;
    return p + (stop - src);
}

// Rows of [p, end) into mat starting at row `row`; returns the number of rows seen.
inline long parse_piece(const char* p, const char* end, double* mat, long row, long n_rows, long n_cols) {
    long rows = 0;
    while (p < end) {
        const char* nl = static_cast<const char*>(std::memchr(p, '\n', static_cast<size_t>(end - p)));
        const char* le = nl ? nl : end;
        if (row + rows < n_rows) {
            double* dst = mat + static_cast<size_t>(row + rows) * n_cols;
            long j = 0;
            const char* q = p;
            while (q < le) {
                while (q < le && is_sep(*q)) ++q;
                if (q >= le) break;
                double v;
                const char* next = parse_field(q, le, &v);
                if (!next) break;
                if (j < n_cols) dst[j] = v;
                ++j;
                if (next == q) break;
                q = next;
            }
        }
        ++rows;
        p = nl ? nl + 1 : end;
    }
    return rows;
}
}  // namespace csv_detail

// run(n, fn): call fn(i) for i in [0, n), possibly in parallel.
typedef std::function<void(int, const std::function<void(int)>&)> CsvRunner;

inline void serial_runner(int n, const std::function<void(int)>& fn) {
    for (int i = 0; i < n; ++i) fn(i);
}

// Fill the pre-sized row-major matrix `mat` (n_rows x n_cols) from the comma separated file `path`.  Throws if the file
// does not have exactly n_rows lines.
inline void read_counts_csv(std::vector<double>& mat, long n_rows, long n_cols, const std::string& path, const CsvRunner& run = serial_runner,
                            int n_workers = 1) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) {
        ::close(fd);
        throw std::runtime_error("cannot stat " + path);
    }
    const size_t size = static_cast<size_t>(st.st_size);
    long total = 0;
    if (size > 0) {
        void* base = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (base == MAP_FAILED) throw std::runtime_error("cannot map " + path);
        madvise(base, size, MADV_SEQUENTIAL);
        const char* data = static_cast<const char*>(base);
        // pieces of >= 1 MB that start behind a line break
        int P = static_cast<int>(std::min<size_t>(static_cast<size_t>(std::max(1, n_workers)) * 4, size / (1u << 20) + 1));
        std::vector<size_t> start(P + 1, size);
        start[0] = 0;
        for (int p = 1; p < P; ++p) {
            size_t b = size / P * p;
            if (b < start[p - 1]) b = start[p - 1];
            if (b > 0 && data[b - 1] != '\n') {
                const char* nl = static_cast<const char*>(std::memchr(data + b, '\n', size - b));
                b = nl ? static_cast<size_t>(nl - data) + 1 : size;
            }
            start[p] = b;
        }
        // pass 1: lines per piece -> first row of every piece
        std::vector<long> rows(P, 0);
        run(P, [&](int p) {
            long n = 0;
            const char* q = data + start[p];
            const char* e = data + start[p + 1];
            while (q < e) {
                const char* nl = static_cast<const char*>(std::memchr(q, '\n', static_cast<size_t>(e - q)));
                ++n;
                q = nl ? nl + 1 : e;
            }
            rows[p] = n;
        });
        std::vector<long> first(P + 1, 0);
        for (int p = 0; p < P; ++p) first[p + 1] = first[p] + rows[p];
        total = first[P];
        if (total > n_rows) {
            munmap(base, size);
            throw std::runtime_error("more rows in " + path + " than --n_cells");
        }
        // pass 2: parse
        run(P, [&](int p) { csv_detail::parse_piece(data + start[p], data + start[p + 1], mat.data(), first[p], n_rows, n_cols); });
        munmap(base, size);
    } else
        ::close(fd);
    if (total != n_rows) throw std::runtime_error("the counts file has " + std::to_string(total) + " rows, --n_cells says " + std::to_string(n_rows));
}

}  // namespace scicone_host
#endif
