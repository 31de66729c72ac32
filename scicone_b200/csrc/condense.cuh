// K0: bins -> regions.  Utils::condense_matrix (reference scicone/src/Utils.cpp:104-138; pyscicone does the same in
// scicone.py:276-299): out[i][r] = sum of the bins of region r of row i, added one bin after the other in bin order,
// which this kernel keeps (bit-identical sums also for non-integer inputs).
//
// The only multi-GB stream of the pipeline (10 000 x 18 000 doubles = 1.44 GB, 100 000 x 20 000 = 16 GB): HBM bound,
// algorithmic traffic 8 * n_bins read + 8 * n_regions written per row.  What limits a streaming kernel on this part is
// the number of bytes in flight per SM (7.7 TB/s x ~2 us of loaded latency / 148 SMs = ~100 KB), so ONE persistent CTA
// per SM owns almost all of the SM's shared memory as a ring of kK0Stages tiles of kK0Rows rows x kK0Tile bins (32 KB
// each, 5 of them in flight while one is reduced), filled by a dedicated producer warp through the TMA unit
// (cp.async.bulk + mbarrier transaction counts, one lane per row; rows land at a padded pitch so that the reduction
// reads shared memory without bank conflicts - which is why these are 32 row copies and not one 2-D tensor copy, whose
// box would be dense) and drained by 8 consumer warps.  Full / empty mbarriers per stage: the producer runs ahead of the
// consumers by the depth of the ring and never waits for their tile barrier.  A consumer thread owns one (row, region)
// piece of a tile and adds its bins sequentially; the region that straddles a tile boundary hands its partial sum to
// the next tile through a per-row carry.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace scicone {

constexpr int kK0Rows = 32;      // rows per group
constexpr int kK0Tile = 128;     // bins per tile and row
constexpr int kK0Consumers = 256;               // threads that reduce
constexpr int kK0Threads = kK0Consumers + 32;   // + the producer warp
constexpr int kK0Stages = 6;     // ring depth
constexpr int kK0MaxIndex = 1536;  // region offsets + tile table entries kept in shared memory (else read from global memory)
constexpr int kK0Stride = kK0Tile + 2;  // doubles per row in a stage: one leading element of slack for 16-byte alignment
constexpr size_t kK0SmemBytes = sizeof(double) * kK0Stages * kK0Rows * kK0Stride + sizeof(double) * 2 * kK0Rows + 8 * 2 * kK0Stages + 64 + sizeof(int) * kK0MaxIndex;
static_assert(kK0Rows == 32, "the issuing lanes (one warp) reduce their byte counts with full-warp shuffles");
static_assert(kK0SmemBytes <= 227 * 1024, "one CTA per SM: the ring must fit the SM's shared memory");

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    while (!done)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit; completion is counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// bins [n_rows][row_stride] row-major, of which the first n_bins = sum of the region sizes are used (the allocation is
// padded by 16 bytes), off [K+1] prefix sums of the region sizes,
// tile_first [n_tiles+1]: first region that reaches into each tile, out [n_rows][K]
__global__ void __launch_bounds__(kK0Threads, 1)
    condense_rows_kernel(const double* __restrict__ bins, long long n_rows, long long row_stride, long long n_bins, const int* __restrict__ off,
                         const int* __restrict__ tile_first, int K, int n_tiles, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);                                  // [kK0Stages][kK0Rows][kK0Stride]
    double* carry = stage_buf + (size_t)kK0Stages * kK0Rows * kK0Stride;                      // [2][kK0Rows]
    uint64_t* full = reinterpret_cast<uint64_t*>(carry + 2 * kK0Rows);                        // [kK0Stages] tile landed
    uint64_t* empty = full + kK0Stages;                                                       // [kK0Stages] tile consumed
    int* sm_index = reinterpret_cast<int*>(empty + kK0Stages + 1);                            // off[K+1], tile_first[n_tiles+1]
    // region offsets and the tile table are read for every piece of every tile: keep them in shared memory when they fit
    const bool index_in_smem = (K + 1) + (n_tiles + 1) <= kK0MaxIndex;
    if (index_in_smem) {
        for (int q = threadIdx.x; q <= K; q += kK0Threads) sm_index[q] = off[q];
        for (int q = threadIdx.x; q <= n_tiles; q += kK0Threads) sm_index[K + 1 + q] = tile_first[q];
    }
    const int* offs = index_in_smem ? sm_index : off;
    const int* firsts = index_in_smem ? sm_index + K + 1 : tile_first;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kK0Stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // The CTA's work is one flat sequence of (row group, tile) items, so the ring also runs across group boundaries.
    const long long n_groups = (n_rows + kK0Rows - 1) / kK0Rows;
    const long long my_groups = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_items = my_groups * n_tiles;

    if (threadIdx.x >= kK0Consumers) {
        // ---- producer warp: item i -> stage i % kK0Stages, as soon as the consumers have released it.  Lane 0 arms the
        // barrier with the byte count, every lane issues one row's bulk copy (aligned down to 16 bytes).
        const int r = threadIdx.x - kK0Consumers;
        for (long long i = 0; i < n_items; ++i) {
            const int s = (int)(i % kK0Stages);
            if (i >= kK0Stages) mbar_wait(&empty[s], (unsigned)(((i / kK0Stages) - 1) & 1));
            const long long row0 = (blockIdx.x + (i / n_tiles) * gridDim.x) * kK0Rows;
            const int rows = (int)min((long long)kK0Rows, n_rows - row0);
            const long long lo = (i % n_tiles) * kK0Tile;
            const int width = (int)min((long long)kK0Tile, n_bins - lo);
            const long long first = (row0 + r) * row_stride + lo;  // element index of the tile's first bin in this row
            const int lead = (int)(first & 1);
            const unsigned bytes = r < rows ? (unsigned)(((width + lead + 1) & ~1) * sizeof(double)) : 0u;
            unsigned total = bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (r == 0) mbar_expect_tx(&full[s], total);
            __syncwarp();
            if (r < rows) tma_load_1d(stage_buf + ((size_t)s * kK0Rows + r) * kK0Stride, bins + first - lead, bytes, &full[s]);
        }
        return;
    }
    // ---- consumer warps
    for (long long i = 0; i < n_items; ++i) {
        const int s = (int)(i % kK0Stages);
        mbar_wait(&full[s], (unsigned)((i / kK0Stages) & 1));
        const int t = (int)(i % n_tiles);
        const long long row0 = (blockIdx.x + (i / n_tiles) * gridDim.x) * kK0Rows;
        const int rows = (int)min((long long)kK0Rows, n_rows - row0);
        const long long lo = (long long)t * kK0Tile;
        const long long hi = min(lo + kK0Tile, n_bins);
        const int r_lo = firsts[t], r_hi = firsts[t + 1];  // regions r_lo .. r_hi (inclusive) reach into this tile
        const int n_reg = r_hi - r_lo + 1;
        for (int item = threadIdx.x; item < n_reg * rows; item += kK0Consumers) {
            const int row = item / n_reg, r = r_lo + item % n_reg;
            if (r >= K) continue;
            const long long s_r = offs[r], e_r = offs[r + 1];
            if (e_r <= lo || s_r >= hi) continue;
            const long long first = (row0 + row) * row_stride + lo;
            const double* src = stage_buf + ((size_t)s * kK0Rows + row) * kK0Stride + (first & 1) - lo;
            double acc = s_r < lo ? carry[(t & 1) * kK0Rows + row] : 0.0;
            const long long b1 = min(e_r, hi);
            // bin order, as the reference adds them: the additions form one dependent chain, the shared-memory
            // loads feeding it are independent and are issued eight at a time
            long long b = max(s_r, lo);
            for (; b + 8 <= b1; b += 8) {
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = src[b + q];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc += v[q];
            }
            for (; b < b1; ++b) acc += src[b];
            if (e_r > hi)
                carry[((t + 1) & 1) * kK0Rows + row] = acc;
            else
                out[(row0 + row) * K + r] = acc;
        }
        // every consumer is done with stage s (and with this tile's carries): hand the stage back to the producer
        asm volatile("bar.sync 1, %0;" ::"n"(kK0Consumers) : "memory");
        if (threadIdx.x == 0) mbar_arrive(&empty[s]);
    }
}

// Cluster condensation (reference scripts/condense_clusters.py:28-38, pyscicone scicone.py:331-335): the mean of the rows of
// every cluster, rows added in row order (numpy's axis-0 reduction order).  One thread per (cluster, column); `order`
// lists the rows sorted by cluster (stable), `first[c] .. first[c+1]` are cluster c's entries.
__global__ void cluster_means_kernel(const double* __restrict__ D, int K, const int* __restrict__ order, const int* __restrict__ first,
                                     int n_clusters, double* __restrict__ means) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (k >= K || c >= n_clusters) return;
    const int b = first[c], e = first[c + 1];
    double acc = 0.0;
    int i = b;
    for (; i + 4 <= e; i += 4) {  // the loads are independent of the running sum: four in flight
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __ldg(D + (long long)order[i + q] * K + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += v[q];
    }
    for (; i < e; ++i) acc += __ldg(D + (long long)order[i] * K + k);
    means[(long long)c * K + k] = e > b ? acc / (double)(e - b) : 0.0;
}

}  // namespace scicone
