// Writer of <postfix>_inferred_cnvs.csv: cells x bins copy numbers as comma separated text,
// what Utils::regions_to_bins_cnvs (Utils.cpp:165-188) + the loop of infer_trees.cpp:321-331 produce one value at a time
// (2 G values = 4 GB of text at configs[4]).
//
// A region contributes its copy number once per bin, so a row is a concatenation of runs "v,v,v,...": every run is one
// memcpy from a prebuilt pattern of the widest region.  The length of every row follows from the region widths and the
// digit counts alone, so the file is sized first and blocks of rows are formatted and written in parallel (pwrite at
// their offsets) by whatever `run` is handed in.
#ifndef SCICONE_HOST_CNV_WRITER_HPP
#define SCICONE_HOST_CNV_WRITER_HPP

#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <unistd.h>
#include <vector>

namespace scicone_host {

// cn: [n_cells][n_regions] row-major; run(n, fn): fn(i) for i in [0, n), possibly in parallel
inline void write_bins_cnvs(const std::string& path, const int32_t* cn, long n_cells, int n_regions, const std::vector<int>& region_sizes,
                            const std::function<void(int, const std::function<void(int)>&)>& run) {
    int widest = 0;
    for (int k = 0; k < n_regions; ++k) widest = std::max(widest, region_sizes[k]);
    // "v," repeated for the widest region, per distinct copy number
    std::map<int, std::string> pattern;
    for (size_t i = 0; i < static_cast<size_t>(n_cells) * n_regions; ++i)
        if (!pattern.count(cn[i])) {
            const std::string tok = std::to_string(cn[i]) + ",";
            std::string rep;
            rep.reserve(tok.size() * static_cast<size_t>(std::max(widest, 0)));
            for (int b = 0; b < widest; ++b) rep += tok;
            pattern.emplace(cn[i], std::move(rep));
        }
    auto tok_len = [&](int v) { return widest > 0 ? pattern.find(v)->second.size() / static_cast<size_t>(widest) : size_t(0); };
    // bytes of every row: its runs, the last comma replaced by the line break (an empty row is just the line break)
    const long rows_per_block = 64;
    const int n_blocks = static_cast<int>((n_cells + rows_per_block - 1) / rows_per_block);
    std::vector<uint64_t> block_bytes(static_cast<size_t>(n_blocks), 0);
    auto row_bytes = [&](long j) {
        uint64_t n = 0;
        for (int k = 0; k < n_regions; ++k) n += tok_len(cn[static_cast<size_t>(j) * n_regions + k]) * static_cast<uint64_t>(region_sizes[k]);
        return n == 0 ? uint64_t(1) : n;
    };
    run(n_blocks, [&](int b) {
        uint64_t n = 0;
        for (long j = b * rows_per_block; j < std::min(n_cells, (b + 1) * rows_per_block); ++j) n += row_bytes(j);
        block_bytes[b] = n;
    });
    std::vector<uint64_t> offset(static_cast<size_t>(n_blocks) + 1, 0);
    for (int b = 0; b < n_blocks; ++b) offset[b + 1] = offset[b] + block_bytes[b];
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) throw std::runtime_error("cannot create " + path);
    bool failed = false;
    run(n_blocks, [&](int b) {
        std::string buf;
        buf.reserve(block_bytes[b]);
        for (long j = b * rows_per_block; j < std::min(n_cells, (b + 1) * rows_per_block); ++j) {
            const size_t start = buf.size();
            for (int k = 0; k < n_regions; ++k) {
                const std::string& rep = pattern.find(cn[static_cast<size_t>(j) * n_regions + k])->second;
                buf.append(rep.data(), rep.size() / static_cast<size_t>(widest) * static_cast<size_t>(region_sizes[k]));
            }
            if (buf.size() > start) buf.pop_back();
            buf.push_back('\n');
        }
        size_t done = 0;
        while (done < buf.size()) {
            const ssize_t w = ::pwrite(fd, buf.data() + done, buf.size() - done, static_cast<off_t>(offset[b] + done));
            if (w <= 0) {
                failed = true;
                return;
            }
            done += static_cast<size_t>(w);
        }
    });
    ::close(fd);
    if (failed) throw std::runtime_error("writing " + path + " failed");
}

}  // namespace scicone_host
#endif
