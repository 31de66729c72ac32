// Device side of the B200 tree-scoring engine (sm_100a, fp64, no tensor cores:
// the path is a sparse gather + scan + reduce, latency / issue bound).
//
// Layout in HBM (cells are the fastest-varying dimension everywhere, so a warp
// touches one 256-byte segment per row it reads):
//   Dt        [K][Mp]   region-major copy of the count matrix D
//   S         [Mp]      per-cell read total  sum_k D[j][k]   (k ascending, as std::accumulate)
//   w         [Mp]      cluster sizes (int32, 0 in the padding)
//   lgamma rows  [Mp]   L(k,cn)[j] = lgamma(D[j][k] + (nu*cn)*r_k), built once per (region, copy number, nu) by
//                       build_rows_kernel, then only gathered; an event contributes L(k,cn_node) - L(k,cn_parent)
//   z-term rows  [Mp]   lgamma(S[j] + nu*z_node), built by build_rows_kernel for every rescored node of a launch
//   score rows   [Mp]   one row per tree node: attachment score of every cell to that node
//   aggregates   [Mp]   per-cell log-sum-exp (or max) over the nodes, + argmax id in max mode
//   lut       [K+1][blocks][2]  count-range table: {min, range} of every row of counts per block of kUnitCells cells
//                       (0 where the counts are not integers in a narrow range): lets K1 evaluate lgamma once per count
//   gather buffer (multi-GPU) flags + per-rank chunk sums, mapped into every rank: written by finish_reduction of ALL
//                       ranks over NVLink, read by wait_flags_kernel + one copy to the host
//
// Two kernels per batch of proposals:
//   build_rows_kernel       (K1) every lgamma of the path (Tree::compute_score, reference scicone/src/Tree.h:214-231;
//                           Tree::get_od_root_score, Tree.h:1792-1797): lgamma rows, z-term rows, region slices of the
//                           root score of proposals that change nu.  Persistent CTAs over (item, cell block) units.
//   score_proposals_kernel  (K2-K4) per proposal (grid.y) and per 128-cell tile (grid.x), one thread per cell: walks
//                           the rescored nodes parent-before-child (Tree::compute_stack, Tree.h:260-286) with the
//                           gathers of the next four nodes in flight (cp.async ring), folds the new scores into the
//                           cell's aggregate exactly like Inference::compute_t_prime_sums
//                           (scicone/src/Inference.h:1069-1196) / MathOp::log_replace_sum
//                           (scicone/src/MathOp.cpp:307-361), and the CTA reduces aggregate * cluster_size into a
//                           partial of the tree score (Inference.h:292-294).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "lgam_core.h"

namespace scicone {

constexpr int kCells = 128;         // cells per CTA of the proposal kernel: one thread per cell
constexpr int kThreads = kCells;
constexpr int kTaskChunk = 32;      // task records staged in shared memory at a time
constexpr int kEvStage = 64;        // event records staged with them (the rest is read from global memory)
constexpr int kDepth = 4;           // nodes whose gathers are in flight per thread (cp.async ring stages)
constexpr int kSlots = 6;           // gathers per node and stage: L rows of two events (node, parent side), two z-term rows
constexpr int kScRing = 4;          // most recent node scores a thread keeps in registers (parent lookups)
constexpr int kMaxOdSlices = 32;    // region slices of a root score
constexpr int kCtasPerChunk = 8;    // 8 CTAs = 1024 cells: fixed-order reduction unit (multi-GPU invariant)
constexpr int kBuildBlock = 256;    // threads per CTA of the build kernel
constexpr int kBuildCells = 4;      // cells per thread and step of the build kernel
constexpr int kUnitSub = 2;         // steps per unit of the build kernel
constexpr int kUnitCells = kBuildBlock * kBuildCells * kUnitSub;  // 2048 cells per (item, block) unit
constexpr int kLutBatch = 8;        // cells per thread that a tabulated unit gathers at a time
constexpr int kLutMax = 2048;       // largest range of counts that is tabulated (host side: the dataset's count-range table)
constexpr int kLutCap = 4096;       // entries of the shared-memory value table of a CTA (several regions of a slice share it)
constexpr int kLutGroup = 32;       // regions of a slice tabulated together at most
static_assert(kLutMax <= kLutCap, "one region's table must fit");
constexpr int kPad = 256;           // row padding (cells)

enum : unsigned {
    TASK_ROOT = 1u,      // tree root: score 0 (Tree::compute_root_score, Tree.h:145-160)
    TASK_SKIP_AGG = 4u,  // a later task rewrites the same node: only that one enters the aggregate
    TASK_PS_FAR = 8u,    // parent score from global memory, read one node ahead: the committed row of a subtree root's parent,
                         // or the row this thread wrote more than kScRing nodes ago
    TASK_BIG = 16u,      // more events than the staging buffer holds: read from global memory
    PROP_MAX = 1u,       // max scoring (Inference::max_scoring)
    PROP_OD = 2u,        // overdispersed likelihood (globals is_overdispersed)
    PROP_FULL = 4u,      // full table pass (Inference::compute_t_table): no committed state is read
    PROP_SCRATCH = 8u,   // to_add.size() >= unchanged_vals.size(): recompute the log-sum from scratch
    PROP_WITH_OD = 16u,  // also evaluate the root ("overdispersed") score from its region slices
    PROP_OD_ONLY = 32u,  // no tasks, no aggregate: only the root score
    BUILD_OD_SLICE = 1u,
    BUILD_OD_SLICE_LOG = 2u,  // non-overdispersed root score slice: sum_k D_k log r_k
    BUILD_G = 3u              // dst = lgamma(src + c1): lgamma rows L(k, cn) (src = Dt[k]) and z-term rows (src = S)
};

struct alignas(16) DevTask {
    double* dst;                 // score row being written
    const double* parent_row;    // TASK_PS_FAR: where the parent's score lives in global memory
    double lg_nz, lg_nzp;        // OD: lgamma(nu*z), lgamma(nu*z_parent);  non-OD: log(z), log(z_parent)
    const double* g_row;         // OD: z-term row lgamma(S + nu*z) built by K1
    const double* gp_row;        // OD: z-term row lgamma(S + nu*z_parent) (the parent task's g_row when the arguments are equal)
    int ev_begin;                // first event record (index into the proposal's event array)
    int node_id;
    unsigned short n_ev;         // number of event records (Node::c_change entries)
    unsigned short ev_rel;       // first event record relative to the chunk's staged events
    unsigned char flags;
    unsigned char parent_back;   // 1..kScRing: the parent is the task this many positions back (score still in registers)
    unsigned char pad[2];
};
static_assert(sizeof(DevTask) == 64, "DevTask is staged as four 16-byte pieces");

struct alignas(16) DevChunk {  // tasks [first_task, first_task + n_tasks) and events [ev_base, ev_base + n_ev) are staged together
    int first_task, n_tasks, ev_base, n_ev;
};

struct alignas(16) DevEvent {
    const double* row;    // OD: lgamma row L(k, cn_node);  non-OD: Dt[k]
    const double* row_p;  // OD: lgamma row L(k, cn_parent)
    double a, b;          // OD: lgamma((nu*cn)*r) of node / parent;  non-OD: a = log(cn_node) - log(cn_parent)
};
static_assert(sizeof(DevEvent) == 32, "DevEvent is staged as two 16-byte pieces");

struct alignas(16) DevUnch {
    const double* row;
    int id;
    int pad;
};

// One per proposal at the start of the arena: where its blob is and what its first chunk stages, so that a CTA can issue
// the header read and the first task / event copies together instead of chasing offsets through the header.
struct alignas(16) LaunchEntry {
    unsigned blob_off;   // byte offset of the proposal's blob (DevHeader first, DevTask array right behind it)
    unsigned n_tasks0;   // tasks of the first chunk
    unsigned n_ev0;      // staged events of the first chunk
    unsigned ev_off0;    // byte offset (inside the blob) of the first chunk's first event record
};

struct alignas(16) DevHeader {
    int n_tasks, n_old, n_unch, deleted_id;
    unsigned flags;
    int n_cells;  // M of this shard
    int n_od_slices;
    int n_chunks;
    const double* S;
    const int* w;
    const double* sums_old;
    double* sums_new;
    const int* maxid_old;
    int* maxid_new;
    double* partial;      // [2][gridDim.x] per-CTA partial sums (0: tree score, 1: root score)
    double* chunk_sums;   // [2][ceil(gridDim.x / kCtasPerChunk)]
    double* total;        // [2]
    unsigned* counter;    // arrival counter for the last-CTA reduction
    unsigned* status;     // bit 0: NaN aggregate
    double od_lg_nz, od_nz;  // OD: lgamma(nu*z_root), nu*z_root;  non-OD: od_lg_nz = log(sum_r)
    unsigned off_tasks, off_events, off_old, off_unch, off_od_rows, off_chunks;
    const double* pad4;
    const double* od_g_row;          // PROP_WITH_OD, OD: row lgamma(S + nu*z_root) built by K1
    // cell-sharded datasets: the chunk sums go straight into every rank's gather buffer over NVLink (peer memory), see
    // finish_reduction.  xchg_world == 0: single GPU.
    double* const* xchg_data;               // [world] base of this rank's block in rank p's buffer (lane and parity chosen by the host)
    unsigned long long* const* xchg_flag;   // [world] this rank's flag row in rank p's buffer
    unsigned long long xchg_seq;            // number of this launch on its lane
    int xchg_world, xchg_slot, xchg_stride, xchg_pad;  // stride = chunk sums per rank and value (max_chunks_rank)
};
constexpr int kXchgSlots = 128;  // proposals per launch of a cell-sharded dataset (SCICONE_MAX_SHARDED_BATCH in the header)

struct alignas(16) DevBuild {
    double* dst;
    const double* src;     // BUILD_G: Dt[k] or S;  slices: Dt
    double c1;             // BUILD_G: (nu*cn)*r or nu*z
    unsigned off_rows;     // OD slices: arena offset of rows[K]: where lgamma(D_k + arg[k]) of region k is also stored (or null)
    unsigned lut_row;      // BUILD_G: 1 + row of the count-range table that describes src (region k, or K for S); 0: none
    unsigned kind;
    int k0, k1;            // slices: region range
    unsigned off_arg;      // slices: arena offsets of arg[K], cc[K]
    unsigned off_c;
    int n_cells;
    long long row_stride;  // Mp
};

// lgamma for positive arguments: from 16 up the Stirling series of lgam_core.h (table-driven logarithm, <= 2 ulp from
// glibc's lgamma and bit-identical for 99.9 % of arguments: tests/test_lgam_cpu.py); below 16 the CUDA library routine.
// The large branch is what the z-terms lgamma(S + nu*z) always take, and what the data terms take unless a region has
// (almost) no reads.  `tbl` is the shared-memory copy of the logarithm's reduction table.
__device__ const double g_log_table[kLogTableDoubles] = {SCICONE_LOG_TABLE_VALUES};

template <int kBlock>  // compile-time block size: a run-time stride costs an integer division in every CTA
__device__ __forceinline__ void load_log_table(double* tbl) {
#pragma unroll
    for (int i = 0; i < (kLogTableDoubles + kBlock - 1) / kBlock; ++i) {
        const int q = i * kBlock + threadIdx.x;
        if (q < kLogTableDoubles) tbl[q] = g_log_table[q];
    }
}

__device__ __forceinline__ double lgam(double x, const double* tbl) {
    if (x >= 16.0) return lgam_large(x, tbl);
    return lgamma(x);
}

__device__ __forceinline__ double block_sum(double v, double* warp_buf) {
    // fixed-shape reduction: xor-shuffle tree inside each warp, warps added in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_buf[wid] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kThreads / 32; ++i) s += warp_buf[i];
    return s;  // valid in thread 0
}

// Last CTA to arrive adds the per-CTA partials in a fixed order: first inside
// 1024-cell chunks, then the chunks.  The result does not depend on the CTA
// schedule, and (because shards are whole chunks) not on the number of GPUs.
// Two sums travel together: [0] the tree score, [1] the root score (0 when not requested).
//
// Cell-sharded datasets: the same CTA stores the chunk sums into the gather buffer of EVERY rank (peer memory, NVLink
// stores) and then publishes the launch number in the flag word of this (rank, proposal) on every rank - a release at
// system scope after a CTA barrier, so a rank that sees the flag sees the sums.  No collective kernel runs, and no GPU
// waits for another one here: only the reader of the sums does (wait_flags_kernel).
__device__ __forceinline__ void finish_reduction(double cta_sum0, double cta_sum1, double* partial, double* chunk_sums,
                                                 double* total, unsigned* counter, const DevHeader& H) {
    __shared__ bool is_last;
    __shared__ double smem_chunks[kThreads];
    const int n_cta = gridDim.x;
    const int n_chunks = (n_cta + kCtasPerChunk - 1) / kCtasPerChunk;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = cta_sum0;
        partial[n_cta + blockIdx.x] = cta_sum1;
        __threadfence();
        unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int v = 0; v < 2; ++v) {
        const double* part = partial + (size_t)v * n_cta;
        double* chunks = chunk_sums + (size_t)v * n_chunks;
        double grand = 0.0;
        for (int base = 0; base < n_chunks; base += kThreads) {
            int c = base + threadIdx.x;
            double cs = 0.0;
            if (c < n_chunks) {
                const int b0 = c * kCtasPerChunk;
                const int b1 = min(b0 + kCtasPerChunk, n_cta);
                for (int b = b0; b < b1; ++b) cs += __ldcg(part + b);
                chunks[c] = cs;
                for (int p = 0; p < H.xchg_world; ++p)
                    H.xchg_data[p][((size_t)H.xchg_slot * 2 + v) * H.xchg_stride + c] = cs;
            }
            smem_chunks[threadIdx.x] = cs;
            __syncthreads();
            if (threadIdx.x == 0) {
                const int lim = min(kThreads, n_chunks - base);
                for (int i = 0; i < lim; ++i) grand += smem_chunks[i];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) total[v] = grand;
    }
    if (threadIdx.x == 0) *counter = 0u;
    if (H.xchg_world > 0) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < H.xchg_world) {
            unsigned long long* f = H.xchg_flag[threadIdx.x] + H.xchg_slot;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(H.xchg_seq) : "memory");
        }
    }
}

// Reader side of the exchange: one small CTA waits until every rank has published launch `seq` for the first n proposals
// (flags of this lane: [world][kXchgSlots]); the copy of the gathered sums to the host follows it on the stream.  A rank
// that never arrives is reported through the status word (bit 1) after `timeout_ns` instead of hanging the GPU.
__global__ void wait_flags_kernel(const unsigned long long* flags, int world, int n, unsigned long long seq, unsigned* status,
                                  unsigned long long timeout_ns) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int idx = threadIdx.x; idx < world * n; idx += blockDim.x) {
        const unsigned long long* f = flags + (size_t)(idx / n) * kXchgSlots + idx % n;
        unsigned spins = 0;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= seq) break;
            if ((++spins & 0x3ffu) == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t - t0 > timeout_ns) {
                    atomicOr(status, 2u);
                    return;
                }
            }
            __nanosleep(200);
        }
    }
}

// A slice of the root score of kBuildCells cells: sum over the regions [k0, k1) (Tree::get_od_root_score, Tree.h:1792-1805).
// The cells of a thread advance together, so their loads are in flight together and their lgamma chains interleave.
__device__ __forceinline__ void build_slice(const DevBuild& it, const unsigned char* __restrict__ arena, int j0, const double* tbl) {
    const double* arg = reinterpret_cast<const double*>(arena + it.off_arg);
    const double* cc = reinterpret_cast<const double*>(arena + it.off_c);
    double acc[kBuildCells];
    bool ok[kBuildCells];
#pragma unroll
    for (int q = 0; q < kBuildCells; ++q) {
        acc[q] = 0.0;
        ok[q] = j0 + q * kBuildBlock < it.n_cells;
    }
    if (it.kind == BUILD_OD_SLICE) {  // Tree.h:1792-1797
        // lgamma(D_k + nu*neutral_k*r_k) is also the lgamma row L(k, neutral_k) of the tree's events: where the proposal
        // needs that row, the value is stored on the way instead of being evaluated a second time by a row item
        double* const* keep = reinterpret_cast<double* const*>(arena + it.off_rows);
        for (int k = it.k0; k < it.k1; ++k) {
            const double* row = it.src + (long long)k * it.row_stride + j0;
            const double a = arg[k], c = cc[k];
            double* const out = keep[k];
            double d[kBuildCells];
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) d[q] = ok[q] ? __ldg(row + q * kBuildBlock) : 32.0;
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) {
                const double v = lgam(d[q] + a, tbl);
                if (out && ok[q]) out[j0 + q * kBuildBlock] = v;
                acc[q] += v;
                acc[q] -= c;
            }
        }
    } else {  // Tree.h:1803-1805
        for (int k = it.k0; k < it.k1; ++k) {
            const double* row = it.src + (long long)k * it.row_stride + j0;
            const double a = arg[k];
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q)
                if (ok[q]) acc[q] += __ldg(row + q * kBuildBlock) * a;
        }
    }
#pragma unroll
    for (int q = 0; q < kBuildCells; ++q)
        if (ok[q]) it.dst[j0 + q * kBuildBlock] = acc[q];
}

// The same slice for a unit whose counts are integers in a narrow range (the normal case: read counts): the CTA tabulates
// lgamma(v + arg_k) once for every value v the unit's cells can hold in region k, for as many regions of the slice as
// fit the table, and the cells pick their entries - the same function of the same argument as the direct evaluation,
// bit for bit, at a fraction of the lgamma work.  `info` = {min, range} pairs of this unit's block, one per region
// (stride info_stride ints); the caller has checked that every region of the slice has a table there.  A slice whose
// tables do not fit at once is done in groups of regions, the partial sums waiting in dst in between (the order of
// the additions of a cell is that of build_slice: regions ascending).
struct SliceStage {  // per-region constants of the group of regions whose tables are resident (shared memory)
    int offs[kLutGroup + 1];  // first table entry of region g
    int vmin[kLutGroup];
    int range[kLutGroup];
    double arg[kLutGroup];
    double c[kLutGroup];
    double* keep[kLutGroup];
    int n;
};

__device__ __forceinline__ void build_slice_lut(const DevBuild& it, const unsigned char* __restrict__ arena, int base, const double* tbl,
                                                double* vals, SliceStage& st, const int* __restrict__ info, int info_stride) {
    const double* arg = reinterpret_cast<const double*>(arena + it.off_arg);
    const double* cc = reinterpret_cast<const double*>(arena + it.off_c);
    double* const* keep = reinterpret_cast<double* const*>(arena + it.off_rows);
    int k = it.k0;
    while (k < it.k1) {
        // the constants of up to kLutGroup regions, fetched side by side ...
        if (threadIdx.x < kLutGroup && k + (int)threadIdx.x < it.k1) {
            const int g = threadIdx.x;
            st.vmin[g] = info[(long long)(k + g) * info_stride];
            st.range[g] = info[(long long)(k + g) * info_stride + 1];
            st.arg[g] = arg[k + g];
            st.c[g] = cc[k + g];
            st.keep[g] = keep[k + g];
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // ... then as many of them as fit the value table: offs[g] = first entry of region k + g
            int tot = 0, n = 0;
            st.offs[0] = 0;
            while (k + n < it.k1 && n < kLutGroup && tot + st.range[n] <= kLutCap) {
                tot += st.range[n];
                ++n;
                st.offs[n] = tot;
            }
            st.n = n;
        }
        __syncthreads();
        const int n = st.n;
        const int tot = st.offs[n];
        for (int e = threadIdx.x; e < tot; e += kBuildBlock) {
            int g = 0;
            while (st.offs[g + 1] <= e) ++g;
            vals[e] = lgam((double)(st.vmin[g] + (e - st.offs[g])) + st.arg[g], tbl);
        }
        __syncthreads();
#pragma unroll 1
        for (int b = 0; b < kUnitCells / (kBuildBlock * kLutBatch); ++b) {
            const int j0 = base + b * kBuildBlock * kLutBatch + threadIdx.x;
            if (j0 - (int)threadIdx.x >= it.n_cells) break;  // uniform; no barrier inside this loop anyway
            double acc[kLutBatch];
            bool ok[kLutBatch];
#pragma unroll
            for (int q = 0; q < kLutBatch; ++q) {
                ok[q] = j0 + q * kBuildBlock < it.n_cells;
                acc[q] = (k == it.k0 || !ok[q]) ? 0.0 : it.dst[j0 + q * kBuildBlock];
            }
            // the counts of region g + 1 are in flight while region g is looked up
            double d[kLutBatch], dn[kLutBatch];
            const double* row = it.src + (long long)k * it.row_stride + j0;
#pragma unroll
            for (int q = 0; q < kLutBatch; ++q) dn[q] = ok[q] ? __ldg(row + q * kBuildBlock) : 0.0;
            for (int g = 0; g < n; ++g) {
#pragma unroll
                for (int q = 0; q < kLutBatch; ++q) d[q] = dn[q];
                if (g + 1 < n) {
                    row += it.row_stride;
#pragma unroll
                    for (int q = 0; q < kLutBatch; ++q) dn[q] = ok[q] ? __ldg(row + q * kBuildBlock) : 0.0;
                }
                const double c = st.c[g];
                double* const out = st.keep[g];
                const int shift = st.offs[g] - st.vmin[g];
#pragma unroll
                for (int q = 0; q < kLutBatch; ++q) {
                    const double v = vals[ok[q] ? (int)d[q] + shift : st.offs[g]];
                    if (out && ok[q]) out[j0 + q * kBuildBlock] = v;
                    acc[q] += v;
                    acc[q] -= c;
                }
            }
#pragma unroll
            for (int q = 0; q < kLutBatch; ++q)
                if (ok[q]) it.dst[j0 + q * kBuildBlock] = acc[q];
        }
        __syncthreads();  // the table and the constants are rewritten by the next group / unit
        k += n;
    }
}

// ---------------------------------------------------------------------------------
// K1: data-only lgamma work.  Persistent grid (up to 4 CTAs per SM) over n_items x ceil(M / kUnitCells) units.
//
// `lut` (or null) is the dataset's count-range table: for every row of counts (regions 0..K-1, then the row sums S) and
// every block of kUnitCells cells, {smallest count, number of distinct values to tabulate} - the second is 0 where the
// counts are not integers, or spread too widely for a table to pay.  A unit with a table evaluates lgamma once per
// possible count (typically a few hundred) instead of once per cell (2048).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBuildBlock, 4)
    build_rows_kernel(const unsigned char* __restrict__ arena, unsigned off_items, int n_items, int blocks_per_row,
                      const int* __restrict__ lut) {
    __shared__ double tbl[kLogTableDoubles];
    __shared__ double vals[kLutCap];
    __shared__ SliceStage stage;
    load_log_table<kBuildBlock>(tbl);
    const DevBuild* items = reinterpret_cast<const DevBuild*>(arena + off_items);
    const int total = n_items * blocks_per_row;
    __syncthreads();  // table loaded
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        const DevBuild& it = items[u / blocks_per_row];
        const int blk = u % blocks_per_row;
        const int base = blk * kUnitCells;
        const unsigned kind = it.kind;
        if (kind == BUILD_G) {  // one lgamma per cell: L(k, cn) (Tree.h:214-218) or the z-term of a node (Tree.h:229-231)
            double* const dst = it.dst;
            const double* const src = it.src;
            const double c1 = it.c1;
            const int n_cells = it.n_cells;
            int vmin = 0, range = 0;
            if (lut != nullptr && it.lut_row != 0) {
                const int* info = lut + ((long long)(it.lut_row - 1) * blocks_per_row + blk) * 2;
                vmin = info[0];
                range = info[1];
            }
            if (range > 0) {  // uniform over the CTA: tabulate, then gather kLutBatch cells per thread at a time
                int idx[kLutBatch];
                auto load = [&](int b) {
#pragma unroll
                    for (int q = 0; q < kLutBatch; ++q) {
                        const int j = base + (b * kLutBatch + q) * kBuildBlock + threadIdx.x;
                        idx[q] = j < n_cells ? (int)__ldg(src + j) - vmin : 0;
                    }
                };
                load(0);
                for (int i = threadIdx.x; i < range; i += kBuildBlock) vals[i] = lgam((double)(vmin + i) + c1, tbl);
                __syncthreads();
#pragma unroll 1
                for (int b = 0; b < kUnitCells / (kBuildBlock * kLutBatch); ++b) {
                    if (base + b * kLutBatch * kBuildBlock >= n_cells) break;  // uniform
                    double v[kLutBatch];
#pragma unroll
                    for (int q = 0; q < kLutBatch; ++q) v[q] = vals[idx[q]];
                    const int jb = base + b * kLutBatch * kBuildBlock + threadIdx.x;
                    if (b + 1 < kUnitCells / (kBuildBlock * kLutBatch) && jb + kLutBatch * kBuildBlock - threadIdx.x < n_cells) load(b + 1);
#pragma unroll
                    for (int q = 0; q < kLutBatch; ++q) {
                        const int j = jb + q * kBuildBlock;
                        if (j < n_cells) dst[j] = v[q];
                    }
                }
                __syncthreads();
            } else {
#pragma unroll 1
                for (int h = 0; h < kUnitSub; ++h) {  // kBuildCells lgamma chains at a time
                    const int j0 = base + h * kBuildCells * kBuildBlock + threadIdx.x;
                    if (j0 >= n_cells) break;
                    double d[kBuildCells];
#pragma unroll
                    for (int q = 0; q < kBuildCells; ++q) {
                        const int j = j0 + q * kBuildBlock;
                        d[q] = j < n_cells ? __ldg(src + j) : 32.0;
                    }
#pragma unroll
                    for (int q = 0; q < kBuildCells; ++q) {
                        const int j = j0 + q * kBuildBlock;
                        if (j < n_cells) dst[j] = lgam(d[q] + c1, tbl);
                    }
                }
            }
        } else {
            bool tabulated = kind == BUILD_OD_SLICE && lut != nullptr;
            if (tabulated)
                for (int k = it.k0; k < it.k1; ++k) tabulated = tabulated && lut[((long long)k * blocks_per_row + blk) * 2 + 1] > 0;
            if (tabulated) {
                build_slice_lut(it, arena, base, tbl, vals, stage, lut + (long long)blk * 2, 2 * blocks_per_row);
                continue;
            }
#pragma unroll 1
            for (int sub = 0; sub < kUnitSub; ++sub) {
                const int j0 = base + sub * kBuildBlock * kBuildCells + threadIdx.x;
                if (j0 - threadIdx.x >= it.n_cells) break;
                build_slice(it, arena, j0, tbl);
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// K2-K4: grid = (ceil(M / kCells), n_proposals), one THREAD per cell.  The arena starts with one 32-bit blob offset
// per proposal.
//
// A thread walks the rescored nodes of its proposal parent-before-child (Tree::compute_stack, Tree.h:260-286):
//   stage   (CTA) task + event records of the next kTaskChunk nodes -> shared memory, one coalesced 16-byte copy per thread
//   score   delta-row gathers issued together and added in the reference's order, z-terms by the table-driven lgamma
//           (lgamma(S + nu z_parent) is the previous node's lgamma(S + nu z) when the parent was visited just before),
//           score = parent score + increment: the parent's score is the previous node's (register) or was prefetched one
//           node ahead from the row this thread wrote earlier / from the committed table; row written, running max / LSE
//   C       aggregate (Inference::compute_t_prime_sums): the unchanged rows are only read by cells that need them,
//           four loads in flight
//   D       root score from its region slices (nu proposals)
//   reduce  aggregate * cluster_size -> warp shuffle -> CTA partial -> fixed-order chunk / grand totals (last CTA)
// There is no per-cell state in shared memory, so occupancy is bounded by registers only and the only barriers are the
// two around each staging step; with B proposals x M cells >= 300 k threads the SMs stay full (LPT order of the
// proposals in grid.y handles the long full-tree rescorings of nu proposals).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void copy16(void* dst, const void* src) {
    *reinterpret_cast<int4*>(dst) = __ldg(reinterpret_cast<const int4*>(src));
}
// 8-byte asynchronous copy global -> this thread's own shared-memory slot (LDGSTS): no register is tied up while the
// load is in flight, so a thread keeps kDepth nodes x kSlots gathers outstanding
__device__ __forceinline__ void cp8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

template <bool kMax, bool kOd, int kOcc>
__global__ void __launch_bounds__(kThreads, kOcc) score_proposals_kernel(const unsigned char* __restrict__ arena) {
    __shared__ __align__(16) DevTask sm_task[kTaskChunk];
    __shared__ __align__(16) DevEvent sm_ev[kEvStage];
    __shared__ double ring[kDepth * kSlots][kThreads];  // [stage * kSlots + slot][thread]: every thread owns one column
    __shared__ double red_buf[kThreads / 32];

    const LaunchEntry le = reinterpret_cast<const LaunchEntry*>(arena)[blockIdx.y];
    const unsigned char* blob = arena + le.blob_off;
    const DevHeader& H = *reinterpret_cast<const DevHeader*>(blob);
    // first chunk: its records follow from the launch entry alone, so these copies go out together with the header reads
    for (int q = threadIdx.x; q < (int)le.n_tasks0 * 4; q += kThreads)
        copy16(reinterpret_cast<unsigned char*>(sm_task) + 16 * q, blob + sizeof(DevHeader) + 16 * q);
    for (int q = threadIdx.x; q < (int)le.n_ev0 * 2; q += kThreads)
        copy16(reinterpret_cast<unsigned char*>(sm_ev) + 16 * q, blob + le.ev_off0 + 16 * q);

    const int tid = threadIdx.x;
    const int j = blockIdx.x * kCells + tid;
    const bool live = j < H.n_cells;
    const unsigned pflags = H.flags;
    constexpr bool od = kOd;
    const double S = live ? __ldg(H.S + j) : 1.0;

    // running aggregate over the rescored nodes
    double new_max = -DBL_MAX;  // max over the new values (both modes)
    int new_id = 0;             // its node id (lowest id among equals = first in ascending-id order)
    double new_sumexp = 0.0;    // sum mode: sum_i exp(v_i - new_max)
    bool prev_in_new = (pflags & PROP_FULL) != 0;
    const bool incremental = live && !(pflags & (PROP_FULL | PROP_OD_ONLY));
    const int pm = (kMax && incremental) ? H.maxid_old[j] : -1;
    const double agg_old = incremental ? H.sums_old[j] : 0.0;  // fetched now, needed in phase C

    double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;  // scores of the last kScRing nodes of the walk (s1 = the node just before)
    static_assert(kScRing == 4, "the register ring is written out for four entries");
    const DevChunk* chunks = reinterpret_cast<const DevChunk*>(blob + H.off_chunks);
    const int n_chunks = H.n_chunks;
    __syncthreads();  // first chunk staged
    for (int c = 0; c < n_chunks; ++c) {
        DevChunk ck;
        if (c == 0) {
            ck.first_task = 0;
            ck.n_tasks = (int)le.n_tasks0;
            ck.ev_base = 0;
            ck.n_ev = (int)le.n_ev0;
        } else {
            ck = chunks[c];
            const DevTask* tasks = reinterpret_cast<const DevTask*>(blob + sizeof(DevHeader));
            const DevEvent* events = reinterpret_cast<const DevEvent*>(blob + H.off_events);
            // ---- stage the task and event records of this chunk
            __syncthreads();
            for (int q = threadIdx.x; q < ck.n_tasks * 4; q += kThreads)
                copy16(reinterpret_cast<unsigned char*>(sm_task) + 16 * q, reinterpret_cast<const unsigned char*>(tasks + ck.first_task) + 16 * q);
            for (int q = threadIdx.x; q < ck.n_ev * 2; q += kThreads)
                copy16(reinterpret_cast<unsigned char*>(sm_ev) + 16 * q, reinterpret_cast<const unsigned char*>(events + ck.ev_base) + 16 * q);
            __syncthreads();
        }
        if (!live) continue;
        const int nt = ck.n_tasks;
        // issue the gathers of staged task t into ring stage t % kDepth (one commit group per task, also when empty)
        auto issue = [&](int t) {
            const DevTask& tk = sm_task[t];
            const unsigned tf = tk.flags;
            double* col = &ring[(t % kDepth) * kSlots][tid];
            if (!(tf & TASK_ROOT)) {
                if (!(tf & TASK_BIG)) {
                    const DevEvent* ev = sm_ev + tk.ev_rel;
                    const int ne = tk.n_ev;
                    if (ne > 0) {
                        cp8(col + 0 * kThreads, ev[0].row + j);
                        if (od) cp8(col + 1 * kThreads, ev[0].row_p + j);
                    }
                    if (ne > 1) {
                        cp8(col + 2 * kThreads, ev[1].row + j);
                        if (od) cp8(col + 3 * kThreads, ev[1].row_p + j);
                    }
                }
                if (od) {
                    cp8(col + 4 * kThreads, tk.g_row + j);
                    cp8(col + 5 * kThreads, tk.gp_row + j);
                }
            }
            cp_commit();
        };
#pragma unroll
        for (int q = 0; q < kDepth; ++q) {
            if (q < nt)
                issue(q);
            else
                cp_commit();
        }
        double ps_far = 0.0;  // parent score that is not in the register ring: plain load, one node ahead
        if (sm_task[0].flags & TASK_PS_FAR) ps_far = __ldcg(sm_task[0].parent_row + j);
        for (int t = 0; t < nt; ++t) {
            const DevTask& tk = sm_task[t];
            const unsigned tf = tk.flags;
            const double ps_far_cur = ps_far;
            if (t + 1 < nt && (sm_task[t + 1].flags & TASK_PS_FAR)) ps_far = __ldcg(sm_task[t + 1].parent_row + j);
            cp_wait<kDepth - 1>();  // this node's gathers have landed (groups complete in order)
            const double* col = &ring[(t % kDepth) * kSlots][tid];
            double sc = 0.0;
            // ---- increment  score(node) - score(parent)   (Tree::compute_score, Tree.h:177-238)
            if (!(tf & TASK_ROOT)) {
                double x = 0.0;
                const int ne = tk.n_ev;
                const bool big = tf & TASK_BIG;  // more events than the staging buffer holds (rare): read from global memory
                const DevEvent* ev = big ? reinterpret_cast<const DevEvent*>(blob + H.off_events) + tk.ev_begin : sm_ev + tk.ev_rel;
                if (od) {  // added in event order (Tree.h:214-218): lgamma rows of the node's and the parent's copy number
                    if (!big) {
                        if (ne > 0) {
                            x += col[0 * kThreads] - col[1 * kThreads];
                            x -= ev[0].a;
                            x += ev[0].b;
                        }
                        if (ne > 1) {
                            x += col[2 * kThreads] - col[3 * kThreads];
                            x -= ev[1].a;
                            x += ev[1].b;
                        }
                    }
                    for (int q = big ? 0 : 2; q < ne; ++q) {
                        x += __ldg(ev[q].row + j) - __ldg(ev[q].row_p + j);
                        x -= ev[q].a;
                        x += ev[q].b;
                    }
                    x += tk.lg_nz;
                    x -= tk.lg_nzp;
                    x -= col[4 * kThreads];  // lgamma(S + nu z)
                    x += col[5 * kThreads];  // lgamma(S + nu z_parent)
                } else {  // Tree.h:221,236-237
                    if (!big) {
                        if (ne > 0) x += col[0 * kThreads] * ev[0].a;
                        if (ne > 1) x += col[2 * kThreads] * ev[1].a;
                    }
                    for (int q = big ? 0 : 2; q < ne; ++q) x += __ldg(ev[q].row + j) * ev[q].a;
                    x -= S * tk.lg_nz;
                    x += S * tk.lg_nzp;
                }
                const int back = tk.parent_back;
                const double ps = back == 1 ? s1 : back == 2 ? s2 : back == 3 ? s3 : back == 4 ? s4 : ps_far_cur;
                sc = ps + x;
            }
            s4 = s3;
            s3 = s2;
            s2 = s1;
            s1 = sc;
            tk.dst[j] = sc;
            // refill this stage with the node kDepth positions ahead
            if (t + kDepth < nt)
                issue(t + kDepth);
            else
                cp_commit();
            if (tf & TASK_SKIP_AGG) continue;
            const int id = tk.node_id;
            if (kMax && id == pm) prev_in_new = true;
            if (kMax) {
                if (sc > new_max || (sc == new_max && id < new_id)) {
                    new_max = sc;
                    new_id = id;
                }
            } else {  // online log-sum-exp of the new values
                if (sc > new_max) {
                    new_sumexp = new_sumexp * exp(new_max - sc) + 1.0;
                    new_max = sc;
                } else
                    new_sumexp += exp(sc - new_max);
            }
        }
        cp_wait<0>();
    }

    double contrib = 0.0, contrib_od = 0.0;
    if (live && !(pflags & PROP_OD_ONLY)) {
        // ---- per-cell aggregate: Inference::compute_t_prime_sums ----
        const int n_unch = H.n_unch;
        const DevUnch* unch = reinterpret_cast<const DevUnch*>(blob + H.off_unch);
        double res;
        if (kMax) {  // Inference.h:1086-1152
            int prev_id = pm;
            double prev_val = agg_old;
            if (prev_in_new || pm == H.deleted_id) {
                // the previous argmax was rescored or deleted: first maximum of the unchanged rows in ascending id order.
                // The rows are fetched through the thread's ring, kDepth * kSlots of them in flight.
                constexpr int kBatch = kDepth * kSlots;
                double cur = -DBL_MAX;
                int uid = 0;
                for (int u0 = 0; u0 < n_unch; u0 += kBatch) {
                    const int nb = min(kBatch, n_unch - u0);
                    for (int q = 0; q < nb; ++q) cp8(&ring[q][tid], unch[u0 + q].row + j);
                    cp_commit();
                    cp_wait<0>();
                    for (int q = 0; q < nb; ++q) {
                        const double v = ring[q][tid];
                        if (v > cur) {
                            cur = v;
                            uid = unch[u0 + q].id;
                        }
                    }
                }
                prev_id = uid;
                prev_val = cur > -DBL_MAX ? cur : 0.0;  // value-initialised std::pair when nothing is unchanged
            }
            res = new_max;
            int rid = new_id;
            if (prev_val > new_max) {
                res = prev_val;
                rid = prev_id;
            }
            H.maxid_new[j] = rid;
        } else {  // MathOp::log_replace_sum, MathOp.cpp:307-361
            // rows are fetched through the thread's ring, kBatch at a time; maxima and sums run over them in list order
            constexpr int kBatch = kDepth * kSlots;
            bool scratch = (pflags & (PROP_SCRATCH | PROP_FULL)) != 0;
            double sub = 0.0;
            if (!scratch) {  // incremental branch: take the old values out of the old sum
                const double* const* old_rows = reinterpret_cast<const double* const*>(blob + H.off_old);
                const double sum_old = agg_old;
                const int n_old = H.n_old;
                double mx = sum_old;
                for (int o0 = 0; o0 < n_old; o0 += kBatch) {
                    const int nb = min(kBatch, n_old - o0);
                    for (int q = 0; q < nb; ++q) cp8(&ring[q][tid], old_rows[o0 + q] + j);
                    cp_commit();
                    cp_wait<0>();
                    for (int q = 0; q < nb; ++q) {
                        const double v = ring[q][tid];
                        if (v > mx) mx = v;
                    }
                }
                double s = 0.0;
                s += exp(sum_old - mx);
                for (int o0 = 0; o0 < n_old; o0 += kBatch) {
                    const int nb = min(kBatch, n_old - o0);
                    if (n_old > kBatch) {  // more rows than the ring holds: fetch this batch again
                        for (int q = 0; q < nb; ++q) cp8(&ring[q][tid], old_rows[o0 + q] + j);
                        cp_commit();
                        cp_wait<0>();
                    }
                    for (int q = 0; q < nb; ++q) s -= exp(ring[q][tid] - mx);
                }
                if (s <= 1e-6)
                    scratch = true;  // possible precision loss: recompute from every node
                else
                    sub = log(s) + mx;
            }
            double mx = new_max;
            if (scratch) {
                for (int u0 = 0; u0 < n_unch; u0 += kBatch) {
                    const int nb = min(kBatch, n_unch - u0);
                    for (int q = 0; q < nb; ++q) cp8(&ring[q][tid], unch[u0 + q].row + j);
                    cp_commit();
                    cp_wait<0>();
                    for (int q = 0; q < nb; ++q) {
                        const double v = ring[q][tid];
                        if (v > mx) mx = v;
                    }
                }
            } else if (sub > mx)
                mx = sub;
            double s = new_sumexp * exp(new_max - mx);
            if (scratch) {
                for (int u0 = 0; u0 < n_unch; u0 += kBatch) {
                    const int nb = min(kBatch, n_unch - u0);
                    if (n_unch > kBatch) {
                        for (int q = 0; q < nb; ++q) cp8(&ring[q][tid], unch[u0 + q].row + j);
                        cp_commit();
                        cp_wait<0>();
                    }
                    for (int q = 0; q < nb; ++q) s += exp(ring[q][tid] - mx);
                }
            } else
                s += exp(sub - mx);
            res = log(s) + mx;
        }
        if (res != res) atomicOr(H.status, 1u);
        H.sums_new[j] = res;
        contrib = res * (double)H.w[j];
    }
    if (live && (pflags & PROP_WITH_OD)) {
        // Tree::get_od_root_score (Tree.h:1774-1810) from the region slices K1 prepared;
        // Inference::compute_t_prime_od_scores, Inference.h:1421-1433
        const double* const* slices = reinterpret_cast<const double* const*>(blob + H.off_od_rows);
        const int ns = H.n_od_slices;
        double v = 0.0;
        if (od) {
            v += H.od_lg_nz;
            v -= __ldcg(H.od_g_row + j);
        } else
            v += -S * H.od_lg_nz;
        constexpr int kBatch = kDepth * kSlots;
        for (int s0 = 0; s0 < ns; s0 += kBatch) {
            const int nb = min(kBatch, ns - s0);
            for (int q = 0; q < nb; ++q) cp8(&ring[q][tid], slices[s0 + q] + j);
            cp_commit();
            cp_wait<0>();
            for (int q = 0; q < nb; ++q) v += ring[q][tid];
        }
        contrib_od = v * (double)H.w[j];
    }
    // ---- weighted sum over cells: Inference::comparison, Inference.h:292-294 ----
    __syncthreads();
    const double cta_sum = block_sum(contrib, red_buf);
    __syncthreads();
    double cta_od = 0.0;
    if (pflags & PROP_WITH_OD) {  // uniform per proposal
        cta_od = block_sum(contrib_od, red_buf);
        __syncthreads();
    }
    finish_reduction(cta_sum, cta_od, H.partial, H.chunk_sums, H.total, H.counter, H);
}

// Test hook: the device lgamma on caller-provided arguments (scicone_test_lgamma).
__global__ void lgam_probe_kernel(const double* __restrict__ x, double* __restrict__ out, int n) {
    __shared__ double tbl[kLogTableDoubles];
    load_log_table<256>(tbl);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lgam(x[i], tbl);
}

// Multi-GPU: copy the per-chunk sums of each proposal into the fixed-width, zero-padded send buffer of
// the all-gather (one CTA per proposal).
__global__ void pack_chunks_kernel(const double* __restrict__ chunk_sums, int n_chunks, double* __restrict__ send,
                                   int width) {
    // per proposal: [2][n_chunks] -> [2][width]
    for (int c = threadIdx.x; c < 2 * width; c += blockDim.x) {
        const int v = c / width, q = c % width;
        send[(long long)blockIdx.x * 2 * width + c] = q < n_chunks ? chunk_sums[((long long)blockIdx.x * 2 + v) * n_chunks + q] : 0.0;
    }
}

// ---------------------------------------------------------------------------------
// Dataset set-up: D arrives row-major [M][K] (cells x regions, as the reference reads it);
// transpose into Dt [K][Mp] through a padded shared-memory tile, then row sums.
// ---------------------------------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ D, double* __restrict__ Dt, int M, int K, long long Mp) {
    __shared__ double tile[32][33];
    const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, k = k0 + threadIdx.x;
        if (j < M && k < K) tile[r][threadIdx.x] = D[(long long)j * K + k];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, j = j0 + threadIdx.x;
        if (j < M && k < K) Dt[(long long)k * Mp + j] = tile[threadIdx.x][r];
    }
}

__global__ void row_sum_kernel(const double* __restrict__ Dt, double* __restrict__ S, int M, int K, long long Mp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = s + Dt[(long long)k * Mp + j];  // std::accumulate order (Tree.h:253)
    S[j] = s;
}

// ---------------------------------------------------------------------------------
// Inference::assign_cells_to_nodes, Inference.h:1346-1362: first maximum in ascending id order,
// then the cells x regions copy-number matrix of the chosen nodes.
// ---------------------------------------------------------------------------------
__global__ void argmax_kernel(const double* const* __restrict__ rows, const int* __restrict__ ids, int n_rows, int M,
                              int* __restrict__ out_id, int* __restrict__ out_col) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    int best = 0;
    double bv = rows[0][j];
    for (int c = 1; c < n_rows; ++c) {
        const double v = rows[c][j];
        if (bv < v) {
            bv = v;
            best = c;
        }
    }
    out_id[j] = ids[best];
    out_col[j] = best;
}

__global__ void expand_cn_kernel(const int* __restrict__ col, const int* __restrict__ cn /*[n_rows][K]*/, int M, int K,
                                 int* __restrict__ out /*[M][K]*/) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * K) return;
    const int j = (int)(i / K), k = (int)(i % K);
    out[i] = cn[(long long)col[j] * K + k];
}

__global__ void gather_rows_kernel(const double* const* __restrict__ rows, int n_rows, int M, double* __restrict__ out /*[M][n_rows]*/) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, j = j0 + threadIdx.x;
        if (c < n_rows && j < M) tile[r][threadIdx.x] = rows[c][j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, c = c0 + threadIdx.x;
        if (c < n_rows && j < M) out[(long long)j * n_rows + c] = tile[threadIdx.x][r];
    }
}

}  // namespace scicone
