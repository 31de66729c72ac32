// Test driver for scicone_b200/host/csv_reader.hpp: parse <file> as <rows> x <cols> with the parallel reader (threads of
// this process) and with the serial getline + strtod loop the host used before (the reference's Utils::read_counts
// semantics), compare bit for bit, print both times.   csv_check <file> <rows> <cols> [threads]
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

#include "../../scicone_b200/host/csv_reader.hpp"

static void serial_read(std::vector<double>& mat, long n_rows, long n_cols, const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::string line;
    long i = 0;
    while (std::getline(in, line)) {
        if (i >= n_rows) throw std::runtime_error("more rows in " + path + " than --n_cells");
        const char* p = line.c_str();
        long j = 0;
        while (*p) {
            char* end = nullptr;
            double v = std::strtod(p, &end);
            if (end == p) break;
            if (j < n_cols) mat[static_cast<size_t>(i) * n_cols + j] = v;
            ++j;
            p = end;
            while (*p == ',' || *p == ' ' || *p == '\r') ++p;
        }
        ++i;
    }
    if (i != n_rows) throw std::runtime_error("the counts file has " + std::to_string(i) + " rows, --n_cells says " + std::to_string(n_rows));
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string path = argv[1];
    const long rows = std::atol(argv[2]), cols = std::atol(argv[3]);
    const int threads = argc > 4 ? std::atoi(argv[4]) : 4;
    typedef std::chrono::steady_clock Clock;
    std::vector<double> a(static_cast<size_t>(rows) * cols, -7.0), b(a);
    std::string err_a, err_b;
    auto runner = [threads](int n, const std::function<void(int)>& fn) {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([&, t] {
                for (int i = t; i < n; i += threads) fn(i);
            });
        for (std::thread& t : pool) t.join();
    };
    Clock::time_point t0 = Clock::now();
    try {
        scicone_host::read_counts_csv(a, rows, cols, path, runner, threads);
    } catch (const std::exception& e) {
        err_a = e.what();
    }
    Clock::time_point t1 = Clock::now();
    try {
        serial_read(b, rows, cols, path);
    } catch (const std::exception& e) {
        err_b = e.what();
    }
    Clock::time_point t2 = Clock::now();
    std::printf("parallel %.1f ms, serial %.1f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(t2 - t1).count());
    if (err_a != err_b) {
        std::printf("error mismatch: '%s' vs '%s'\n", err_a.c_str(), err_b.c_str());
        return 1;
    }
    if (!err_a.empty()) {
        std::printf("both failed: %s\n", err_a.c_str());
        return 0;
    }
    if (std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) != 0) {
        for (size_t i = 0; i < a.size(); ++i)
            if (std::memcmp(&a[i], &b[i], 8) != 0) {
                std::printf("mismatch at row %zu col %zu: %.17g vs %.17g\n", i / cols, i % cols, a[i], b[i]);
                break;
            }
        return 1;
    }
    std::printf("identical\n");
    return 0;
}
