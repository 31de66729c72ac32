// K0: bins -> regions.  Utils::condense_matrix (reference scicone/src/Utils.cpp:104-138; pyscicone does the same in
// scicone.py:276-299): out[i][r] = sum of the bins of region r of row i, added one bin after the other in bin order,
// which this kernel keeps (bit-identical sums also for non-integer inputs).
//
// The only multi-GB stream of the pipeline (10 000 x 18 000 doubles = 1.44 GB, 100 000 x 20 000 = 16 GB): HBM bound,
// algorithmic traffic 8 * n_bins read + 8 * n_regions written per row.  The sum of a region is one dependent chain of
// additions per row (~30 cycles per addition with its shared-memory load: fp64 additions of one warp do not pipeline),
// so the parallelism is ROWS x REGION BLOCKS: a work item is a group of 32 rows and a block of consecutive regions; a warp
// takes an item, lane = row, and streams the block's bins through a private shared-memory ring of tiles of 32 rows x
// kTile bins.  All lanes step through the same bins - region boundaries are the same for every row - so there is no
// divergence, no barrier between threads other than the ring's mbarriers, and the running sum of the region that
// straddles a tile boundary simply stays in its register.  Blocks start on region boundaries, so no sum is ever split.
//
// The tiles are filled by the TMA unit.  Measured on this part: an SM retires about one bulk copy per 45 ns whatever its
// size, so a tile must be ONE copy - a 2-D tensor-map copy (cp.async.bulk.tensor.2d, one elected lane per tile, 32 rows x
// 34 bins = 8.7 KB) - to reach the HBM rate; 32 one-row copies of 1 KB cap a kernel at ~3.5 TB/s and 256-byte ones at
// 1.4 TB/s (profiles/r2_summary.md).  A tile's rows are 34 x 8 = 272 bytes apart in shared memory: 68 words, i.e. lane l
// starts at bank 4 l mod 32 - the 16 lanes of a 64-bit access phase fall on 8 distinct bank pairs (two-way conflict)
// instead of on one, which a power-of-two tile width would give.  What else bounds a streaming kernel is the bytes in
// flight per SM (7.7 TB/s x ~2 us of loaded latency / 148 SMs = ~100 KB): the launch picks warps per CTA and ring depth so
// that the active warps of an SM together keep that much in flight (few row groups: few warps with deep rings).
// Matrices whose row pitch is not a multiple of 16 bytes (odd bin counts) cannot be described by a tensor map and take
// the same kernel with one-row 1-D bulk copies.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace scicone {

constexpr int kK0Rows = 32;                 // rows per group = lanes of the warp that owns it
constexpr int kK0TileTensor = 34;           // bins per tile, 2-D tensor copies (dense rows of 272 bytes)
constexpr int kK0TileRows = 128;            // bins per tile, one-row 1-D copies (odd row pitch): rows at a pitch of kK0TileRows + 2
constexpr int kK0MaxWarps = 8;
constexpr int kK0MaxStages = 8;
constexpr int kK0MaxBlocks = 64;            // region blocks a row group is cut into (more warps than row groups)
constexpr int kK0MaxIndex = 2048;           // region offsets kept in shared memory (else read from global memory)
constexpr size_t kK0SmemBudget = 214 * 1024;  // 8 warps x 3 stages x 8.7 KB fit
__host__ __device__ constexpr size_t k0_stage_bytes(bool tensor) {
    return sizeof(double) * kK0Rows * (tensor ? (size_t)kK0TileTensor : (size_t)kK0TileRows + 2);
}
static_assert(k0_stage_bytes(true) % 128 == 0, "tensor copies land on 128-byte aligned stages");
static_assert(kK0TileTensor % 2 == 0 && kK0TileRows % 2 == 0, "tiles start on even bins");

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
// 2-D tensor-map copy: the box at (bin c0, row c1) of the matrix -> shared; out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// bins [n_rows][row_stride] row-major, of which the first n_bins = sum of the region sizes are used (the allocation is
// padded by 16 bytes), off [K+1] prefix sums of the region sizes, out [n_rows][K].  blockDim.x = 32 * warps per CTA;
// dynamic shared memory: [warps][n_stages] stages, then [warps][n_stages] mbarriers, then the region offsets.
// kTensor: `map` describes the matrix as (row_stride, n_rows) doubles with boxes of kK0TileTensor x 32.
template <bool kTensor>
__global__ void __launch_bounds__(32 * kK0MaxWarps, 1)
    condense_rows_kernel(const __grid_constant__ CUtensorMap map, const double* __restrict__ bins, long long n_rows, long long row_stride,
                         long long n_bins, const int* __restrict__ off, int K, int n_stages, const int* __restrict__ blocks, int n_blocks,
                         double* __restrict__ out) {
    constexpr int kTile = kTensor ? kK0TileTensor : kK0TileRows;
    constexpr int kPitch = kTensor ? kK0TileTensor : kK0TileRows + 2;  // doubles between the rows of a stage
    constexpr size_t kStageBytes = k0_stage_bytes(kTensor);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int wpc = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* stage_buf = reinterpret_cast<double*>(smem_raw + (size_t)warp * n_stages * kStageBytes);  // this warp's ring
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)wpc * n_stages * kStageBytes) + warp * n_stages;
    int* sm_off = reinterpret_cast<int*>(smem_raw + (size_t)wpc * n_stages * (kStageBytes + 8));
    const bool off_in_smem = K + 1 <= kK0MaxIndex;
    if (off_in_smem)
        for (int q = threadIdx.x; q <= K; q += blockDim.x) sm_off[q] = off[q];
    const int* offs = off_in_smem ? sm_off : off;
    if (lane == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long n_groups = (n_rows + kK0Rows - 1) / kK0Rows;
    const long long n_work = n_groups * n_blocks;  // item w = (row group w / n_blocks, region block w % n_blocks)
    const long long warps_total = (long long)gridDim.x * wpc;
    unsigned long long it_issue = 0, it_wait = 0;  // tiles issued / consumed by this warp so far (stage and parity of the ring)
    for (long long w = (long long)blockIdx.x * wpc + warp; w < n_work; w += warps_total) {
        const long long g = w / n_blocks;
        const int blk = (int)(w - g * n_blocks);
        const int r_first = blocks[blk], r_last = blocks[blk + 1];  // regions [r_first, r_last)
        const long long bin0 = offs[r_first], bin1 = offs[r_last];
        const long long tile0 = bin0 & ~1ll;  // tiles start on even bins: a box must start on a 16-byte boundary of its row
        const long long n_tiles = (bin1 - tile0 + kTile - 1) / kTile;
        const long long row0 = g * kK0Rows;
        const int rows = (int)min((long long)kK0Rows, n_rows - row0);
        const bool valid = lane < rows;
        const long long row = row0 + lane;
        // tile t of this group -> next stage of the ring
        auto issue = [&](long long t) {
            const int s = (int)(it_issue % n_stages);
            const long long lo = tile0 + t * kTile;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' reads of this stage are ordered before the refill
            if (kTensor) {  // one copy for the whole tile; the unit always delivers the full box (zeros beyond the matrix)
                if (lane == 0) {
                    mbar_expect_tx(&full[s], (unsigned)kStageBytes);
                    tma_load_2d(stage_buf + (size_t)s * kK0Rows * kPitch, &map, (int)lo, (int)row0, &full[s]);
                }
            } else {  // lane 0 arms the barrier with the byte count, every lane issues its row's copy (source aligned down to
                      // 16 bytes; `lead` = 1 when the row's first bin of the tile sits at an odd element)
                const int width = (int)min((long long)kTile, bin1 - lo);
                const long long first = row * row_stride + lo;
                const int lead = (int)(first & 1);
                const unsigned bytes = valid ? (unsigned)(((width + lead + 1) & ~1) * sizeof(double)) : 0u;
                unsigned total = bytes;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (lane == 0) mbar_expect_tx(&full[s], total);
                __syncwarp();
                if (valid) tma_load_1d(stage_buf + ((size_t)s * kK0Rows + lane) * kPitch, bins + first - lead, bytes, &full[s]);
            }
            ++it_issue;
        };
        for (long long t = 0; t < n_stages - 1 && t < n_tiles; ++t) issue(t);
        int r = r_first;
        long long e_r = offs[r_first + 1];  // end of the current region
        double acc = 0.0;
        for (long long t = 0; t < n_tiles; ++t) {
            if (t + n_stages - 1 < n_tiles) issue(t + n_stages - 1);  // its stage was consumed in the previous iteration
            const int s = (int)(it_wait % n_stages);
            mbar_wait(&full[s], (unsigned)((it_wait / n_stages) & 1));
            ++it_wait;
            const long long lo = tile0 + t * kTile;
            const long long hi = min(lo + kTile, bin1);
            const double* src = stage_buf + ((size_t)s * kK0Rows + lane) * kPitch + (kTensor ? 0 : (int)((row * row_stride + lo) & 1)) - lo;
            long long b = max(lo, bin0);
            while (b < hi) {
                const long long end = min(e_r, hi);
                // bin order, as the reference adds them: the additions form one dependent chain, the shared-memory loads
                // feeding it are independent and are issued eight at a time
                for (; b + 8 <= end; b += 8) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = src[b + q];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc += v[q];
                }
                for (; b < end; ++b) acc += src[b];
                if (end == e_r) {  // the region is complete (uniform over the warp: region boundaries are the same for every row)
                    if (valid) out[row * K + r] = acc;
                    acc = 0.0;
                    ++r;
                    e_r = r < r_last ? offs[r + 1] : bin1 + 1;
                }
            }
            __syncwarp();  // every lane is done with stage s before it is refilled
        }
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
