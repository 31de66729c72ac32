// Device side of the B200 tree-scoring engine (sm_100a, fp64, no tensor cores: the path is a sparse gather + scan +
// reduce - it is bound by load latency and issue slots, not by a dense contraction).
//
// Layout in HBM (cells are the fastest-varying dimension everywhere, so a warp touches one contiguous segment per row):
//   Dt   [K][Mp] f64   region-major copy of the count matrix D
//   It   [K][Mp] u16   D[j][k] - min_k where the counts of region k over the shard's cells are integers in a narrow
//                      range (read counts are): the index into the per-event value tables of K1
//   S    [Mp]    f64   per-cell read total sum_k D[j][k] (k ascending, as std::accumulate);  Si [Mp] u16 the same as index
//   w    [Mp]    i32   cluster sizes (0 in the padding)
//   lut  [K+1][2] i32  {min, range} of every row of counts (regions, then S) over the shard's cells
//                      (range 0: not tabulated - non-integer counts, or spread too widely for a table to pay)
//   tables (per launch lane) f64  the value tables K1a builds for the launch's events / z-terms / root-score regions
//   row pool  [rows][Mp] f64   ONE contiguous virtual range per rank, addressed by 32-bit row indices: score rows
//                      (one per tree node and chain), per-cell aggregates and argmax ids, root-score slices.  Proposals
//                      refer to rows by index only, so a compiled proposal is valid on every rank of a cell-sharded run.
//   gather buffer (multi-GPU)  flags + per-rank chunk sums, mapped into every rank, written over NVLink by the last
//                      CTA of a launch (finish_reduction), read and added up by gather_totals_kernel, which writes the totals into pinned host memory
//
// Three kernels per batch of proposals:
//   build_tables_kernel (K1a) + build_increments_kernel (K1b)  for every rescored (cell, node): score(node) -
//                           score(parent), i.e. all of Tree::compute_score (reference scicone/src/Tree.h:163-244) but the
//                           final addition - every pair is independent of every other, so this part is embarrassingly
//                           parallel.  The data terms lgamma(D + a_node) - lgamma(D + a_parent) of an event are looked
//                           up in a table over the possible counts (a few hundred lgamma instead of one per cell); also
//                           the root score of proposals that change nu (Tree::get_od_root_score, Tree.h:1774-1810):
//                           from the dataset's count histograms (one number per few regions, K1a only) or, for
//                           non-integer data, as per-cell region slices.
//   score_proposals_kernel  (K2-K4) per proposal (grid.y) and per 128-cell tile (grid.x), one thread per cell: walks the
//                           rescored nodes parent-before-child (Tree::compute_stack, Tree.h:260-286), score = parent
//                           score + increment (the increments of the next eight nodes are in flight in registers),
//                           folds the new scores into the cell's aggregate exactly like Inference::compute_t_prime_sums
//                           (scicone/src/Inference.h:1069-1196) / MathOp::log_replace_sum (scicone/src/MathOp.cpp:307-361),
//                           and the CTA reduces aggregate * cluster_size into a partial of the tree score
//                           (Inference.h:292-294).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "lgam_core.h"

namespace scicone {

constexpr int kCells = 128;         // cells per CTA of the proposal kernel: one thread per cell
constexpr int kThreads = kCells;
constexpr int kTaskChunk = 64;      // task records staged in shared memory at a time
constexpr int kGroup = 8;           // nodes whose increments are in flight per thread (two register buffers of this size)
constexpr int kScRing = 16;         // most recent node scores of a thread's walk that stay in its column of a shared-memory ring (parent lookups)
constexpr int kMaxOdSlices = 32;    // region slices of a root score
constexpr int kCtasPerChunk = 8;    // 8 CTAs = 1024 cells: fixed-order reduction unit (multi-GPU invariant)
constexpr int kTablesBlock = 256;   // threads per CTA of the table kernel (K1a)
constexpr int kBuildBlock = 256;    // threads per CTA of the build kernel
constexpr int kBuildCells = 4;      // cells per thread and step of the build kernel
constexpr int kLutMax = 4096;       // largest range of counts that is tabulated (host side: the dataset's count-range table)
constexpr int kPad = 256;           // row padding (cells)
constexpr unsigned kNoRow = 0xffffffffu;

enum : unsigned {
    TASK_ROOT = 1u,      // tree root: score 0 (Tree::compute_root_score, Tree.h:145-160)
    TASK_SKIP_AGG = 4u,  // a later task rewrites the same node: only that one enters the aggregate
    TASK_PS_FAR = 8u,    // parent score from global memory, read one node ahead: the committed row of a subtree root's parent,
                         // or the row this thread wrote more than kScRing nodes ago
    PROP_MAX = 1u,       // max scoring (Inference::max_scoring)
    PROP_OD = 2u,        // overdispersed likelihood (globals is_overdispersed)
    PROP_FULL = 4u,      // full table pass (Inference::compute_t_table): no committed state is read
    PROP_SCRATCH = 8u,   // to_add.size() >= unchanged_vals.size(): recompute the log-sum from scratch
    PROP_WITH_OD = 16u,  // also evaluate the root ("overdispersed") score from its region slices
    PROP_OD_ONLY = 32u,  // no tasks, no aggregate: only the root score
    PROP_OD_HIST = 64u,  // the region part of the root score comes from the dataset's count histograms (ITEM_OD_HIST): one
                         // number per item in the table arena instead of one slice row per cell
    ITEM_NODE_OD = 1u,   // increment of one node, Dirichlet-multinomial branch (Tree.h:212-231)
    ITEM_NODE_MN = 2u,   // increment of one node, multinomial branch (Tree.h:221,236-237)
    ITEM_SLICE_OD = 3u,  // slice of the root score (Tree.h:1792-1797)
    ITEM_SLICE_MN = 4u,  // non-overdispersed root score slice: sum_k D_k log r_k (Tree.h:1803-1805)
    ITEM_G_ROW = 5u,     // dst = lgamma(S + c): the cell-level term of the root score
    ITEM_OD_HIST = 6u    // K1a only: sum over the regions [k0, k1) and the possible counts v of
                         // hist_k[v] * (lgamma(v + nu*neutral_k*r_k) - lgamma(nu*neutral_k*r_k)) -> one number (g_tab)
};

// Everything that differs between the ranks of a cell-sharded run (addresses, shard size, reduction scratch) travels as
// kernel parameter; the staged proposals themselves are rank independent.
struct LaunchParams {
    const unsigned char* arena;  // this launch: LaunchEntry[n_props], proposal blobs, K1 items, their events / slice constants
    double* rows;                // base of the row pool
    long long stride;            // doubles per row (Mp)
    const double* Dt;
    const unsigned short* It;
    const double* S;
    const unsigned short* Si;
    const int* w;
    const int* lut;   // [K + 1][2]
    double* tables;   // table arena of this launch's lane
    int n_cells, n_regions;
    // reduction scratch of the launch's lane, indexed by slot (position of the proposal in the caller's list)
    double* partial;     // [slots][2][n_cta]   (0: tree score, 1: root score)
    double* chunk_sums;  // [slots][2][cs_stride]
    double* total;       // [slots][2]
    unsigned* counter;   // [slots] arrival counters of the last-CTA reduction
    unsigned* status;    // [slots] bit 0: NaN aggregate
    int n_cta, n_red_chunks, cs_stride, n_props;
    // single-GPU datasets: the results go straight into pinned host memory (no copy, no event): totals [slots][2], status
    // [slots], then the launch's ticket in *h_flag once every proposal of the launch is in
    double* h_total;
    unsigned* h_status;
    unsigned long long* h_flag;
    unsigned long long ticket;
    // cell-sharded datasets (xchg_world > 0): the chunk sums of the whole launch go straight into every rank's gather
    // buffer over NVLink (peer memory), see finish_reduction
    unsigned* launch_counter;               // proposals of this launch whose reduction has finished
    double* const* xchg_data;               // [world] this rank's block in rank p's buffer (lane and parity chosen by the host)
    unsigned long long* const* xchg_flag;   // [world] this rank's flag word in rank p's buffer
    unsigned long long xchg_seq;            // number of this launch on its lane
    int xchg_world;
    // K1
    unsigned off_items;
    int n_items;
    // Root score from count histograms (datasets whose counts are integers in a narrow range in every region): hist holds,
    // region after region, sum_j cluster_size_j [D_jk == gmin_k + i] over ALL cells of the run (every rank of a cell-sharded
    // run holds the same global histogram, so every rank computes the same number and only `od_owner` adds it to its sums)
    const double* hist;
    const int* glut;           // [K][2] {gmin_k, range_k}
    const unsigned* hist_off;  // [K + 1] first entry of region k
    int od_owner;
};

struct alignas(16) DevTask {
    unsigned dst;         // score row being written
    unsigned inc;         // row that holds the increment (== dst unless the node is rescored twice by the proposal)
    unsigned parent_row;  // TASK_PS_FAR: row that holds the parent's score
    int node_id;
    unsigned flags;
    unsigned parent_back;  // 1..kScRing: the parent is the task this many positions back (1: register, else the thread's ring column)
    unsigned pad[2];
};
static_assert(sizeof(DevTask) == 32, "DevTask is staged as two 16-byte pieces");

struct alignas(16) DevUnch {
    unsigned row;
    int id;
};

// One per proposal at the start of the arena.
struct alignas(16) LaunchEntry {
    unsigned blob_off;  // byte offset of the proposal's blob (DevHeader first, DevTask array right behind it)
    unsigned n_tasks0;  // tasks of the first chunk
    unsigned slot;      // reduction slot = position in the caller's list
    unsigned pad;
};

struct alignas(16) DevHeader {
    int n_tasks, n_old, n_unch, deleted_id;
    unsigned flags;
    int n_od_slices;
    unsigned sums_old, sums_new;    // rows: per-cell aggregates (committed / primed)
    unsigned maxid_old, maxid_new;  // rows holding int32 argmax ids
    unsigned od_g_row;              // PROP_WITH_OD, OD: row lgamma(S + nu*z_root) built by K1
    unsigned off_old;               // u32 rows whose values leave the aggregate
    unsigned off_unch, off_od_rows;
    unsigned od_dot_tab;            // PROP_OD_HIST: rank-local, filled at submit: first of the n_od_slices numbers of the ITEM_OD_HIST items
    unsigned pad1;
    double od_lg_nz, od_nz;  // OD: lgamma(nu*z_root), nu*z_root;  non-OD: od_lg_nz = log(sum_r)
};
static_assert(sizeof(DevHeader) % 16 == 0, "the task records follow the header");

struct alignas(16) DevEvent {
    double argN, argP;  // OD: (nu*cn)*r of node / parent
    double a, b;        // OD: lgamma(argN), lgamma(argP);  non-OD: a = log(cn_node) - log(cn_parent)
    int k;              // region
    unsigned tab;       // rank-local, filled at submit: first entry of the event's table in the launch's table arena (kNoRow: none)
    int pad[2];
};
static_assert(sizeof(DevEvent) == 48, "");

struct alignas(16) DevItem {
    unsigned kind;
    unsigned dst;      // row written
    unsigned aux_off;  // arena offset: NODE: its DevEvent array;  SLICE: arg[K] then c[K] (doubles)
    int n_ev;
    double nz, nzp;        // NODE_OD: nu*z, nu*z_parent;  G_ROW: nz = c
    double lg_nz, lg_nzp;  // NODE_OD: lgamma(nz), lgamma(nzp);  NODE_MN: log(z), log(z_parent)
    int k0, k1;            // SLICE: region range
    unsigned g_tab;        // rank-local, filled at submit: NODE_OD: the two z-term tables;  G_ROW: its table (kNoRow: none)
    unsigned pad;
};
static_assert(sizeof(DevItem) == 64, "");

// lgamma for positive arguments: from 16 up the Stirling series of lgam_core.h (table-driven logarithm, <= 2 ulp from
// glibc's lgamma and bit-identical for 99.9 % of arguments: tests/test_lgam.py); below 16 the CUDA library routine.
// `tbl` is the shared-memory copy of the logarithm's reduction table.
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

// Programmatic dependent launch (sm_90+): K1a -> K1b -> proposal kernel (-> gather kernel) are launched back to back on one
// stream with programmatic stream serialization, so the next kernel's CTAs are scheduled, and run their prologue (which
// only reads the staged arena and committed rows), while the previous kernel drains; `pdl_wait` blocks until the previous
// kernel has completed and its writes are visible.  Both are no-ops in a kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// Last CTA of a proposal to arrive adds the per-CTA partials in a fixed order: first inside 1024-cell chunks, then the
// chunks.  The result does not depend on the CTA schedule, and (because shards are whole chunks) not on the number of
// GPUs.  Two sums travel together: [0] the tree score, [1] the root score (0 when not requested).
//
// Cell-sharded datasets: the CTA that finishes the LAST proposal of the launch copies the chunk sums of the whole launch
// into the gather buffer of every rank (peer memory: 16-byte NVLink stores, all threads) and then publishes the launch
// number in this rank's flag word on every rank - one release at system scope per peer after a CTA barrier, so a rank
// that sees the flag sees the sums.  No collective kernel runs and no GPU waits for another one here: only the reader of
// the sums does (gather_totals_kernel).
__device__ __forceinline__ void finish_reduction(double cta_sum0, double cta_sum1, bool with_od, unsigned slot, const LaunchParams& P) {
    __shared__ int role;  // 0: nothing left to do, 1: last CTA of the proposal, 2: ... and of the launch
    __shared__ double smem_chunks[kThreads];
    const int n_cta = P.n_cta;
    const int n_chunks = P.n_red_chunks;
    double* partial = P.partial + (size_t)slot * 2 * n_cta;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = cta_sum0;
        if (with_od) partial[n_cta + blockIdx.x] = cta_sum1;
        __threadfence();
        const unsigned ticket = atomicAdd(P.counter + slot, 1u);
        role = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (role == 0) return;
    __threadfence();
    for (int v = 0; v < 2; ++v) {
        double* chunks = P.chunk_sums + ((size_t)slot * 2 + v) * P.cs_stride;
        if (v == 1 && !with_od) {  // no root score was asked for: zeros, without the round of loads and barriers
            for (int c = threadIdx.x; c < P.cs_stride; c += kThreads) chunks[c] = 0.0;
            if (threadIdx.x == 0) P.total[(size_t)slot * 2 + 1] = 0.0;
            break;
        }
        const double* part = partial + (size_t)v * n_cta;
        double grand = 0.0;
        for (int base = 0; base < P.cs_stride; base += kThreads) {
            const int c = base + threadIdx.x;
            double cs = 0.0;
            if (c < n_chunks) {
                const int b0 = c * kCtasPerChunk;
                const int b1 = min(b0 + kCtasPerChunk, n_cta);
                for (int b = b0; b < b1; ++b) cs += __ldcg(part + b);
            }
            if (c < P.cs_stride) chunks[c] = cs;  // zero beyond this rank's chunks: every rank sends the same width
            smem_chunks[threadIdx.x] = cs;
            __syncthreads();
            if (threadIdx.x == 0) {
                const int lim = min(kThreads, n_chunks - base);
                for (int i = 0; i < lim; ++i) grand += smem_chunks[i];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) P.total[(size_t)slot * 2 + v] = grand;
    }
    if (threadIdx.x == 0) P.counter[slot] = 0u;
    if (P.h_flag) {
        // host-visible results: the CTA that finishes the LAST proposal of the launch copies every total and status word of
        // the launch into pinned host memory, then - one fence at system scope later - the launch's ticket
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();  // this proposal's totals are visible before the launch counter moves
            const unsigned done = atomicAdd(P.launch_counter, 1u);
            role = (done == (unsigned)P.n_props - 1) ? 2 : 1;
        }
        __syncthreads();
        if (role != 2) return;
        __threadfence();
        for (int q = threadIdx.x; q < 2 * P.n_props; q += kThreads) P.h_total[q] = __ldcg(P.total + q);
        for (int q = threadIdx.x; q < P.n_props; q += kThreads) P.h_status[q] = __ldcg(P.status + q);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            *P.launch_counter = 0u;
            *reinterpret_cast<volatile unsigned long long*>(P.h_flag) = P.ticket;
        }
        return;
    }
    if (P.xchg_world == 0) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // this proposal's chunk sums are visible before the launch counter moves
        const unsigned done = atomicAdd(P.launch_counter, 1u);
        role = (done == (unsigned)P.n_props - 1) ? 2 : 1;
    }
    __syncthreads();
    if (role != 2) return;
    __threadfence();
    // every 16-byte piece is loaded once and stored to all ranks: the stores to peer memory are fire-and-forget, the loads
    // (four in flight per thread) are the only round trips of this loop
    const int n16 = P.n_props * P.cs_stride;  // 16-byte pieces: 2 * n_props * cs_stride doubles, contiguous from slot 0
    const double2* src = reinterpret_cast<const double2*>(P.chunk_sums);
    constexpr int kPieces = 4;
    for (int q0 = threadIdx.x; q0 < n16; q0 += kThreads * kPieces) {
        double2 v[kPieces];
#pragma unroll
        for (int i = 0; i < kPieces; ++i) {
            const int q = q0 + i * kThreads;
            v[i] = q < n16 ? __ldcg(src + q) : make_double2(0.0, 0.0);
        }
        for (int p = 0; p < P.xchg_world; ++p) {
            double2* dst = reinterpret_cast<double2*>(P.xchg_data[p]);
#pragma unroll
            for (int i = 0; i < kPieces; ++i) {
                const int q = q0 + i * kThreads;
                if (q < n16) dst[q] = v[i];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *P.launch_counter = 0u;
    if ((int)threadIdx.x < P.xchg_world) {
        unsigned long long* f = P.xchg_flag[threadIdx.x];
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(P.xchg_seq) : "memory");
    }
}

// Reader side of the exchange: one CTA waits until every rank has published launch `seq` on this lane (flags[world]), then
// adds, per proposal, the chunk sums of all ranks in GLOBAL chunk order (rank 0's chunks, rank 1's, ... - shards are whole
// chunks in cell order, so this is the order a single GPU adds them in and the total does not depend on the number of
// GPUs) and writes totals, status words and the launch's ticket straight into pinned host memory.
// A rank that never arrives is reported through status bit 1 after `timeout_ns` instead of hanging the GPU.
__global__ void __launch_bounds__(128) gather_totals_kernel(const unsigned long long* flags, int world, unsigned long long seq, const double* gathered,
                                                            long long rank_stride, int n_props, int chunks_per_rank, const unsigned* status_dev,
                                                            double* h_total, unsigned* h_status, unsigned long long* h_flag,
                                                            unsigned long long ticket, unsigned long long timeout_ns) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    pdl_wait();  // this rank's proposal kernel has completed (status words, its own flag)
    __syncthreads();
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int idx = threadIdx.x; idx < world; idx += blockDim.x) {
        const unsigned long long* f = flags + idx;
        unsigned spins = 0;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= seq) break;
            if ((++spins & 0x3ffu) == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t - t0 > timeout_ns) {
                    timed_out = 1;
                    break;
                }
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    __threadfence_system();
    for (int idx = threadIdx.x; idx < 2 * n_props; idx += blockDim.x) {
        double v = 0.0;
        for (int rk = 0; rk < world; ++rk) {
            const volatile double* c = gathered + rk * rank_stride + (long long)idx * chunks_per_rank;
            for (int q = 0; q < chunks_per_rank; ++q) v += c[q];
        }
        h_total[idx] = v;
    }
    for (int idx = threadIdx.x; idx < n_props; idx += blockDim.x) h_status[idx] = __ldcg(status_dev + idx) | (timed_out && idx == 0 ? 2u : 0u);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(h_flag) = ticket;
}

// ---------------------------------------------------------------------------------
// K1: every lgamma of the path, in two kernels without a barrier between threads that do different work.
//
// Read counts are integers, and the counts of one region over a shard's cells take a few hundred distinct values: the
// dataset's count-range table `lut` says, for every row of counts (regions 0..K-1, then the row sums S), {smallest count,
// number of values to tabulate} - 0 where the counts are not integers or spread too widely.  For such a row a lgamma
// term is a function of a few hundred possible arguments only:
//   build_tables_kernel (K1a)     one CTA per item evaluates its tables over the possible counts into the launch's table
//                                 arena: per event lgamma(v + a_node) - lgamma(v + a_parent), per node the two z-term
//                                 tables lgamma(s + nu z), lgamma(s + nu z_parent), per region of a root-score slice
//                                 lgamma(v + nu neutral_k r_k).  A few hundred lgamma per table instead of one per cell.
//   build_increments_kernel (K1b) one thread per (item, 4 cells): picks the entries through the 16-bit index matrix It
//                                 and adds them in the reference's order - the same function of the same argument as a
//                                 direct evaluation, bit for bit.  Rows without a table (non-integer data: cluster means)
//                                 are evaluated directly from Dt in the same loop (kDirect).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTablesBlock) build_tables_kernel(const __grid_constant__ LaunchParams P) {
    __shared__ double tbl[kLogTableDoubles];
    __shared__ double red[kTablesBlock / 32];
    pdl_launch_dependents();
    load_log_table<kTablesBlock>(tbl);
    const unsigned char* __restrict__ arena = P.arena;
    const DevItem& it = reinterpret_cast<const DevItem*>(arena + P.off_items)[blockIdx.x];
    const int* __restrict__ lut = P.lut;
    double* __restrict__ T = P.tables;
    const unsigned kind = it.kind;
    __syncthreads();  // logarithm table loaded
    if (kind == ITEM_NODE_OD) {
        const DevEvent* ev = reinterpret_cast<const DevEvent*>(arena + it.aux_off);
        for (int e = 0; e < it.n_ev; ++e) {
            const unsigned off = ev[e].tab;
            if (off == kNoRow) continue;
            const int vmin = lut[2 * ev[e].k], range = lut[2 * ev[e].k + 1];
            const double aN = ev[e].argN, aP = ev[e].argP;
            for (int i = threadIdx.x; i < range; i += kTablesBlock) {
                const double v = (double)(vmin + i);
                T[off + i] = lgam(v + aN, tbl) - lgam(v + aP, tbl);  // Tree.h:214-218, one rounding of the difference as there
            }
        }
        if (it.g_tab != kNoRow) {  // Tree.h:229-231
            const int smin = lut[2 * P.n_regions], srange = lut[2 * P.n_regions + 1];
            const double nz = it.nz, nzp = it.nzp;
            for (int i = threadIdx.x; i < srange; i += kTablesBlock) {
                const double v = (double)(smin + i);
                T[it.g_tab + i] = lgam(v + nz, tbl);
                T[it.g_tab + srange + i] = lgam(v + nzp, tbl);
            }
        }
    } else if (kind == ITEM_SLICE_OD) {  // Tree.h:1792-1797
        const double* arg = reinterpret_cast<const double*>(arena + it.aux_off);
        const unsigned* taboff = reinterpret_cast<const unsigned*>(arg + 2 * P.n_regions);
        for (int k = it.k0; k < it.k1; ++k) {
            const unsigned off = taboff[k];
            if (off == kNoRow) continue;
            const int vmin = lut[2 * k], range = lut[2 * k + 1];
            const double a = arg[k];
            for (int i = threadIdx.x; i < range; i += kTablesBlock) T[off + i] = lgam((double)(vmin + i) + a, tbl);
        }
    } else if (kind == ITEM_OD_HIST) {
        // Region part of Tree::get_od_root_score (Tree.h:1792-1797) summed over the cells FIRST: the cells only enter through
        // how often each count occurs, sum_j w_j (lgamma(D_jk + a_k) - lgamma(a_k)) = sum_v hist_k[v] (lgamma(v + a_k) - lgamma(a_k)),
        // a few hundred lgamma per region instead of a look-up per cell.  Fixed order: thread t takes the entries t, t + 256, ...
        // of each region in turn, then a fixed-shape reduction - the number does not depend on the schedule.
        const double* arg = reinterpret_cast<const double*>(arena + it.aux_off);
        const double* cc = arg + P.n_regions;
        double acc = 0.0;
        for (int k = it.k0; k < it.k1; ++k) {
            const int vmin = P.glut[2 * k], range = P.glut[2 * k + 1];
            const double* h = P.hist + P.hist_off[k];
            const double a = arg[k], c = cc[k];
            for (int i = threadIdx.x; i < range; i += kTablesBlock) {
                const double w = h[i];
                if (w != 0.0) acc += w * (lgam((double)(vmin + i) + a, tbl) - c);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = 0; i < kTablesBlock / 32; ++i) s += red[i];
            T[it.g_tab] = s;
        }
    } else if (kind == ITEM_G_ROW) {
        if (it.g_tab != kNoRow) {
            const int smin = lut[2 * P.n_regions], srange = lut[2 * P.n_regions + 1];
            const double c1 = it.nz;
            for (int i = threadIdx.x; i < srange; i += kTablesBlock) T[it.g_tab + i] = lgam((double)(smin + i) + c1, tbl);
        }
    }
}

// Entries of the value tables K1a wrote: an ordinary cached load, not the read-only path - with programmatic launches this
// kernel's lifetime overlaps K1a's, so the tables are not read-only data of this kernel.
__device__ __forceinline__ double ld_table(const double* p) { return __ldca(p); }

template <bool kDirect>  // kDirect: some row of this shard has no table - the logarithm table is needed
__global__ void __launch_bounds__(kBuildBlock) build_increments_kernel(const __grid_constant__ LaunchParams P) {
    __shared__ double tbl[kDirect ? kLogTableDoubles : 1];
    if (kDirect) {
        load_log_table<kBuildBlock>(tbl);
        __syncthreads();
    }
    pdl_launch_dependents();
    const unsigned char* __restrict__ arena = P.arena;
    const DevItem it = reinterpret_cast<const DevItem*>(arena + P.off_items)[blockIdx.y];  // by value: read before the wait
    const double* T = P.tables;
    const long long stride = P.stride;
    const unsigned kind = it.kind;
    double* const dst = P.rows + (long long)it.dst * stride;
    const int j0 = blockIdx.x * (kBuildBlock * kBuildCells) + threadIdx.x;
    bool ok[kBuildCells];
#pragma unroll
    for (int q = 0; q < kBuildCells; ++q) ok[q] = j0 + q * kBuildBlock < P.n_cells;
    pdl_wait();  // the value tables of K1a are complete from here on
    if (kind == ITEM_NODE_OD) {
        // score(node) - score(parent) in the reference's operation order (Tree::compute_score, Tree.h:212-231): for every
        // event x += lgamma(D_k + a_node) - lgamma(D_k + a_parent); x -= lgamma(a_node); x += lgamma(a_parent); then the
        // four z-terms
        const DevEvent* ev = reinterpret_cast<const DevEvent*>(arena + it.aux_off);
        double x[kBuildCells];
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q) x[q] = 0.0;
        const int n_ev = it.n_ev;
        for (int e = 0; e < n_ev; ++e) {
            const unsigned off = ev[e].tab;
            const long long row = (long long)ev[e].k * stride + j0;
            double d[kBuildCells];
            if (off != kNoRow) {
                const unsigned short* ir = P.It + row;
                unsigned idx[kBuildCells];
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) idx[q] = ok[q] ? (unsigned)__ldg(ir + q * kBuildBlock) : 0u;
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) d[q] = ld_table(T + off + idx[q]);
            } else if (kDirect) {
                const double* dr = P.Dt + row;
                const double aN = ev[e].argN, aP = ev[e].argP;
                double c[kBuildCells];
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) c[q] = ok[q] ? __ldg(dr + q * kBuildBlock) : 32.0;
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) d[q] = lgam(c[q] + aN, tbl) - lgam(c[q] + aP, tbl);
            }
            const double a = ev[e].a, b = ev[e].b;
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) {
                x[q] += d[q];
                x[q] -= a;
                x[q] += b;
            }
        }
        double gn[kBuildCells], gp[kBuildCells];
        if (it.g_tab != kNoRow) {
            const int srange = P.lut[2 * P.n_regions + 1];
            unsigned idx[kBuildCells];
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) idx[q] = ok[q] ? (unsigned)__ldg(P.Si + j0 + q * kBuildBlock) : 0u;
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) {
                gn[q] = ld_table(T + it.g_tab + idx[q]);
                gp[q] = ld_table(T + it.g_tab + srange + idx[q]);
            }
        } else if (kDirect) {
            const double nz = it.nz, nzp = it.nzp;
            double s[kBuildCells];
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) s[q] = ok[q] ? __ldg(P.S + j0 + q * kBuildBlock) : 32.0;
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) {
                gn[q] = lgam(s[q] + nz, tbl);
                gp[q] = lgam(s[q] + nzp, tbl);
            }
        }
        const double lg_nz = it.lg_nz, lg_nzp = it.lg_nzp;
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q) {  // Tree.h:227-231
            x[q] += lg_nz;
            x[q] -= lg_nzp;
            x[q] -= gn[q];
            x[q] += gp[q];
            if (ok[q]) dst[j0 + q * kBuildBlock] = x[q];
        }
    } else if (kind == ITEM_NODE_MN) {  // Tree.h:221,236-237
        const DevEvent* ev = reinterpret_cast<const DevEvent*>(arena + it.aux_off);
        const double lg_nz = it.lg_nz, lg_nzp = it.lg_nzp;
        double x[kBuildCells];
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q) x[q] = 0.0;
        for (int e = 0; e < it.n_ev; ++e) {
            const double* dr = P.Dt + (long long)ev[e].k * stride + j0;
            const double a = ev[e].a;
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q)
                if (ok[q]) x[q] += __ldg(dr + q * kBuildBlock) * a;
        }
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q)
            if (ok[q]) {
                const double s = __ldg(P.S + j0 + q * kBuildBlock);
                x[q] -= s * lg_nz;
                x[q] += s * lg_nzp;
                dst[j0 + q * kBuildBlock] = x[q];
            }
    } else if (kind == ITEM_SLICE_OD) {
        // sum over the regions [k0, k1) of lgamma(D_k + nu*neutral_k*r_k) - lgamma(nu*neutral_k*r_k), regions ascending
        // (Tree::get_od_root_score, Tree.h:1792-1797)
        const double* arg = reinterpret_cast<const double*>(arena + it.aux_off);
        const double* cc = arg + P.n_regions;
        const unsigned* taboff = reinterpret_cast<const unsigned*>(arg + 2 * P.n_regions);
        double acc[kBuildCells];
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q) acc[q] = 0.0;
        for (int k = it.k0; k < it.k1; ++k) {
            const unsigned off = taboff[k];
            const long long row = (long long)k * stride + j0;
            double v[kBuildCells];
            if (off != kNoRow) {
                const unsigned short* ir = P.It + row;
                unsigned idx[kBuildCells];
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) idx[q] = ok[q] ? (unsigned)__ldg(ir + q * kBuildBlock) : 0u;
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) v[q] = ld_table(T + off + idx[q]);
            } else if (kDirect) {
                const double* dr = P.Dt + row;
                const double a = arg[k];
                double c[kBuildCells];
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) c[q] = ok[q] ? __ldg(dr + q * kBuildBlock) : 32.0;
#pragma unroll
                for (int q = 0; q < kBuildCells; ++q) v[q] = lgam(c[q] + a, tbl);
            }
            const double c = cc[k];
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q) {
                acc[q] += v[q];
                acc[q] -= c;
            }
        }
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q)
            if (ok[q]) dst[j0 + q * kBuildBlock] = acc[q];
    } else if (kind == ITEM_SLICE_MN) {  // Tree.h:1803-1805
        const double* arg = reinterpret_cast<const double*>(arena + it.aux_off);
        double acc[kBuildCells];
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q) acc[q] = 0.0;
        for (int k = it.k0; k < it.k1; ++k) {
            const double* dr = P.Dt + (long long)k * stride + j0;
            const double a = arg[k];
#pragma unroll
            for (int q = 0; q < kBuildCells; ++q)
                if (ok[q]) acc[q] += __ldg(dr + q * kBuildBlock) * a;
        }
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q)
            if (ok[q]) dst[j0 + q * kBuildBlock] = acc[q];
    } else {  // ITEM_G_ROW: dst = lgamma(S + c)
        const double c1 = it.nz;
#pragma unroll
        for (int q = 0; q < kBuildCells; ++q) {
            const int j = j0 + q * kBuildBlock;
            if (j < P.n_cells) {
                if (it.g_tab != kNoRow)
                    dst[j] = ld_table(T + it.g_tab + __ldg(P.Si + j));
                else if (kDirect)
                    dst[j] = lgam(__ldg(P.S + j) + c1, tbl);
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// K2-K4: grid = (ceil(M / kCells), n_proposals), one THREAD per cell.  The arena starts with one LaunchEntry per proposal.
//
// A thread walks the rescored nodes of its proposal parent-before-child (Tree::compute_stack, Tree.h:260-286):
//   stage   (CTA) task records of the next kTaskChunk nodes -> shared memory, one coalesced 16-byte copy per thread
//   score   score = parent score + increment (K1 left the increment in the node's own row): the increments of the next
//           kGroup nodes are loaded into registers while the current kGroup are consumed; the parent's score is the previous
//           node's (register), one of the thread's last kScRing scores (its column of a shared-memory ring) or was read one
//           node ahead from the row this thread wrote earlier / from the committed table; row written, running max / online log-sum-exp
//   C       aggregate (Inference::compute_t_prime_sums): the unchanged rows are only read by cells that need them,
//           eight loads in flight
//   D       root score (nu proposals): the cell-level term, plus the region slices where the dataset has no count histograms
//   reduce  aggregate * cluster_size -> warp shuffle -> CTA partial -> fixed-order chunk / grand totals (last CTA)
// The only per-cell state in shared memory is the thread's own column of the score ring (16 KB per CTA: six CTAs per SM
// still fit, registers stay the bound), so the only barriers are the two around each staging step; with B proposals x M
// cells >= 300 k threads the SMs stay full (LPT order of the proposals in grid.y handles the long full-tree rescorings
// of nu proposals).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void copy16(void* dst, const void* src) {
    *reinterpret_cast<int4*>(dst) = __ldg(reinterpret_cast<const int4*>(src));
}

template <bool kMax>
__global__ void __launch_bounds__(kThreads, 6) score_proposals_kernel(const __grid_constant__ LaunchParams P) {
    __shared__ __align__(16) DevTask sm_task[kTaskChunk];
    __shared__ double red_buf[kThreads / 32];
    // the last kScRing scores of every thread's walk, one column per thread (no thread reads another one's column: no barrier)
    __shared__ double sc_ring[kScRing][kThreads];

    pdl_launch_dependents();
    const unsigned char* __restrict__ arena = P.arena;
    const LaunchEntry le = reinterpret_cast<const LaunchEntry*>(arena)[blockIdx.y];
    const unsigned char* blob = arena + le.blob_off;
    const DevHeader& H = *reinterpret_cast<const DevHeader*>(blob);
    const DevTask* tasks = reinterpret_cast<const DevTask*>(blob + sizeof(DevHeader));
    // first chunk: its records follow from the launch entry alone, so these copies go out together with the header reads
    for (int q = threadIdx.x; q < (int)le.n_tasks0 * 2; q += kThreads)
        copy16(reinterpret_cast<unsigned char*>(sm_task) + 16 * q, reinterpret_cast<const unsigned char*>(tasks) + 16 * q);

    const int tid = threadIdx.x;
    const int j = blockIdx.x * kCells + tid;
    const bool live = j < P.n_cells;
    const unsigned pflags = H.flags;
    const long long stride = P.stride;
    double* const rows = P.rows + j;  // this cell's element of row r: rows[r * stride]

    // running aggregate over the rescored nodes
    double new_max = -DBL_MAX;  // max over the new values (both modes)
    int new_id = 0;             // its node id (lowest id among equals = first in ascending-id order)
    double new_sumexp = 0.0;    // sum mode: sum_i exp(v_i - new_max)
    bool prev_in_new = (pflags & PROP_FULL) != 0;
    const bool incremental = live && !(pflags & (PROP_FULL | PROP_OD_ONLY));
    const int pm = (kMax && incremental) ? reinterpret_cast<const int*>(P.rows + (long long)H.maxid_old * stride)[j] : -1;
    const double agg_old = incremental ? rows[(long long)H.sums_old * stride] : 0.0;  // fetched now, needed in phase C

    // Row addresses: base + row * (bytes per row) with a 32-bit row index and a 32-bit row pitch is ONE multiply-add
    // (IMAD.WIDE.U32) - the walk below is bound by the instructions one warp issues, not by memory.
    const unsigned pitch = (unsigned)(stride * (long long)sizeof(double));  // < 2^32: scicone_dataset_create bounds the cell count
    unsigned char* const col = reinterpret_cast<unsigned char*>(rows);
    auto at = [&](unsigned row) -> double* { return reinterpret_cast<double*>(col + (unsigned long long)row * pitch); };
    double prev = 0.0;  // score of the task just before (the parent of a first child: about half of a depth-first walk)
    static_assert((kScRing & (kScRing - 1)) == 0, "ring slots are taken modulo a power of two");
    const int n_tasks = H.n_tasks;
    // everything above reads the staged arena and committed rows only; the increments K1b wrote are read from here on
    pdl_wait();
    __syncthreads();  // first chunk staged
    for (int c0 = 0; c0 < n_tasks; c0 += kTaskChunk) {
        const int nt = min(kTaskChunk, n_tasks - c0);
        if (c0 > 0) {  // ---- stage the task records of this chunk
            __syncthreads();
            for (int q = threadIdx.x; q < nt * 2; q += kThreads)
                copy16(reinterpret_cast<unsigned char*>(sm_task) + 16 * q, reinterpret_cast<const unsigned char*>(tasks + c0) + 16 * q);
            __syncthreads();
        }
        if (!live) continue;
        auto load_inc = [&](int t) -> double {
            const DevTask& tk = sm_task[t];
            return (t < nt && !(tk.flags & TASK_ROOT)) ? __ldcg(at(tk.inc)) : 0.0;
        };
        double nxt[kGroup];
#pragma unroll
        for (int q = 0; q < kGroup; ++q) nxt[q] = load_inc(q);
        double ps_far = 0.0;  // parent score that is not in the ring: plain load, one node ahead
        if (sm_task[0].flags & TASK_PS_FAR) ps_far = __ldcg(at(sm_task[0].parent_row));
        for (int t0 = 0; t0 < nt; t0 += kGroup) {
            double cur[kGroup];
#pragma unroll
            for (int q = 0; q < kGroup; ++q) cur[q] = nxt[q];
            if (t0 + kGroup < nt) {
#pragma unroll
                for (int q = 0; q < kGroup; ++q) nxt[q] = load_inc(t0 + kGroup + q);
            }
#pragma unroll
            for (int q = 0; q < kGroup; ++q) {
                const int t = t0 + q;
                if (t < nt) {
                    const DevTask& tk = sm_task[t];
                    const unsigned tf = tk.flags;
                    const double ps_far_cur = ps_far;
                    if (t + 1 < nt && (sm_task[t + 1].flags & TASK_PS_FAR)) ps_far = __ldcg(at(sm_task[t + 1].parent_row));
                    // ring slot = position of the task in the proposal's walk modulo the ring size (chunks start on multiples of it)
                    static_assert(kTaskChunk % kScRing == 0, "ring slots of the chunked walk");
                    const int slot = (t0 + q) & (kScRing - 1);
                    double sc = 0.0;
                    if (!(tf & TASK_ROOT)) {  // Tree.h:241: score = parent score + increment
                        const unsigned back = tk.parent_back;
                        const double ps = back == 1 ? prev : back != 0 ? sc_ring[(slot - (int)back) & (kScRing - 1)][tid] : ps_far_cur;
                        sc = ps + cur[q];
                    }
                    prev = sc;
                    sc_ring[slot][tid] = sc;
                    *at(tk.dst) = sc;
                    if (!(tf & TASK_SKIP_AGG)) {
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
                }
            }
        }
    }

    constexpr int kBatch = 8;  // rows in flight per thread in the aggregate phases
    double contrib = 0.0, contrib_od = 0.0;
    const int n_unch = H.n_unch;
    const DevUnch* unch = reinterpret_cast<const DevUnch*>(blob + H.off_unch);
    // Max scoring, Inference.h:1086-1152: a cell whose previous argmax was rescored or deleted needs the first maximum of
    // the UNCHANGED rows in ascending id order.  Few cells do (those attached inside the moved subtree), but most warps
    // hold one: the warp scans for them together - lane l reads rows l, l + 32, ... of the cell's column (independent
    // loads, four cells per round), then a shuffle reduction picks the maximum, equal values going to the lower id (the
    // rows are listed in ascending id order, so that is the first one).
    double rescan_val = 0.0;
    int rescan_id = 0;
    if (kMax && !(pflags & (PROP_OD_ONLY | PROP_FULL)) && n_unch > 0) {  // uniform over the CTA
        const bool need = live && (prev_in_new || pm == H.deleted_id);
        unsigned todo = __ballot_sync(0xffffffffu, need);
        const int lane = tid & 31;
        const double* const warp_rows = P.rows + (blockIdx.x * kCells + (tid & ~31));
        constexpr int kCellsPerRound = 4;
        while (todo) {
            int src[kCellsPerRound];
#pragma unroll
            for (int c = 0; c < kCellsPerRound; ++c) {
                src[c] = todo ? __ffs(todo) - 1 : -1;
                if (todo) todo &= todo - 1;
            }
            double best[kCellsPerRound];
            int bid[kCellsPerRound];
#pragma unroll
            for (int c = 0; c < kCellsPerRound; ++c) {
                best[c] = -DBL_MAX;
                bid[c] = 0x7fffffff;
            }
            for (int u = lane; u < n_unch; u += 32) {
                const DevUnch un = unch[u];
                const long long at = (long long)un.row * stride;
                double v[kCellsPerRound];
#pragma unroll
                for (int c = 0; c < kCellsPerRound; ++c) v[c] = src[c] >= 0 ? __ldcg(warp_rows + at + src[c]) : -DBL_MAX;
#pragma unroll
                for (int c = 0; c < kCellsPerRound; ++c)
                    if (v[c] > best[c]) {
                        best[c] = v[c];
                        bid[c] = un.id;
                    }
            }
#pragma unroll
            for (int c = 0; c < kCellsPerRound; ++c) {
                if (src[c] < 0) continue;  // uniform over the warp
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best[c], o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bid[c], o);
                    if (ov > best[c] || (ov == best[c] && oi < bid[c])) {
                        best[c] = ov;
                        bid[c] = oi;
                    }
                }
                if (lane == src[c] && best[c] > -DBL_MAX) {
                    rescan_val = best[c];
                    rescan_id = bid[c];
                }
            }
        }
    }
    if (live && !(pflags & PROP_OD_ONLY)) {
        // ---- per-cell aggregate: Inference::compute_t_prime_sums ----
        double res;
        if (kMax) {  // Inference.h:1086-1152
            int prev_id = pm;
            double prev_val = agg_old;
            if (prev_in_new || pm == H.deleted_id) {  // value-initialised std::pair when nothing is unchanged
                prev_id = rescan_id;
                prev_val = rescan_val;
            }
            res = new_max;
            int rid = new_id;
            if (prev_val > new_max) {
                res = prev_val;
                rid = prev_id;
            }
            reinterpret_cast<int*>(P.rows + (long long)H.maxid_new * stride)[j] = rid;
        } else {  // MathOp::log_replace_sum, MathOp.cpp:307-361
            bool scratch = (pflags & (PROP_SCRATCH | PROP_FULL)) != 0;
            double sub = 0.0;
            if (!scratch) {  // incremental branch: take the old values out of the old sum
                const unsigned* old_rows = reinterpret_cast<const unsigned*>(blob + H.off_old);
                const double sum_old = agg_old;
                const int n_old = H.n_old;
                double mx = sum_old;
                for (int o0 = 0; o0 < n_old; o0 += kBatch) {
                    double v[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) v[q] = o0 + q < n_old ? __ldcg(rows + (long long)old_rows[o0 + q] * stride) : -DBL_MAX;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q)
                        if (v[q] > mx) mx = v[q];
                }
                double s = 0.0;
                s += exp(sum_old - mx);
                for (int o0 = 0; o0 < n_old; o0 += kBatch) {
                    double v[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) v[q] = o0 + q < n_old ? __ldcg(rows + (long long)old_rows[o0 + q] * stride) : 0.0;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q)
                        if (o0 + q < n_old) s -= exp(v[q] - mx);
                }
                if (s <= 1e-6)
                    scratch = true;  // possible precision loss: recompute from every node
                else
                    sub = log(s) + mx;
            }
            double mx = new_max;
            if (scratch) {
                for (int u0 = 0; u0 < n_unch; u0 += kBatch) {
                    double v[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) v[q] = u0 + q < n_unch ? __ldcg(rows + (long long)unch[u0 + q].row * stride) : -DBL_MAX;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q)
                        if (v[q] > mx) mx = v[q];
                }
            } else if (sub > mx)
                mx = sub;
            double s = new_sumexp * exp(new_max - mx);
            if (scratch) {
                for (int u0 = 0; u0 < n_unch; u0 += kBatch) {
                    double v[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) v[q] = u0 + q < n_unch ? __ldcg(rows + (long long)unch[u0 + q].row * stride) : 0.0;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q)
                        if (u0 + q < n_unch) s += exp(v[q] - mx);
                }
            } else
                s += exp(sub - mx);
            res = log(s) + mx;
        }
        if (res != res) atomicOr(P.status + le.slot, 1u);
        rows[(long long)H.sums_new * stride] = res;
        contrib = res * (double)P.w[j];
    }
    if (live && (pflags & PROP_WITH_OD)) {
        // Tree::get_od_root_score (Tree.h:1774-1810) from the region slices K1 prepared;
        // Inference::compute_t_prime_od_scores, Inference.h:1421-1433
        const unsigned* slices = reinterpret_cast<const unsigned*>(blob + H.off_od_rows);
        const int ns = (pflags & PROP_OD_HIST) ? 0 : H.n_od_slices;  // PROP_OD_HIST: the region part is added once per proposal, below
        double v = 0.0;
        if (pflags & PROP_OD) {
            v += H.od_lg_nz;
            v -= __ldcg(rows + (long long)H.od_g_row * stride);
        } else
            v += -__ldg(P.S + j) * H.od_lg_nz;
        for (int s0 = 0; s0 < ns; s0 += kBatch) {
            double x[kBatch];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) x[q] = s0 + q < ns ? __ldcg(rows + (long long)slices[s0 + q] * stride) : 0.0;
#pragma unroll
            for (int q = 0; q < kBatch; ++q)
                if (s0 + q < ns) v += x[q];
        }
        contrib_od = v * (double)P.w[j];
    }
    // ---- weighted sum over cells: Inference::comparison, Inference.h:292-294 ----
    __syncthreads();
    const double cta_sum = block_sum(contrib, red_buf);
    __syncthreads();
    double cta_od = 0.0;
    if (pflags & PROP_WITH_OD) {  // uniform per proposal
        cta_od = block_sum(contrib_od, red_buf);
        __syncthreads();
        if ((pflags & PROP_OD_HIST) && blockIdx.x == 0 && threadIdx.x == 0 && P.od_owner) {
            // region part of the root score: the numbers the ITEM_OD_HIST items of K1a left in the table arena, added in item
            // order into the partial of the first CTA of the run's first shard (the histograms cover ALL cells of the run)
            const double* dots = P.tables + H.od_dot_tab;
            double r = 0.0;
            for (int s = 0; s < H.n_od_slices; ++s) r += __ldcg(dots + s);
            cta_od += r;
        }
    }
    finish_reduction(cta_sum, cta_od, (pflags & PROP_WITH_OD) != 0, le.slot, P);
}

// Dataset set-up: the 16-bit index matrix of the tabulated rows.  Row K of the table describes S.
__global__ void index_kernel(const double* __restrict__ Dt, const double* __restrict__ S, const int* __restrict__ lut, int M, int K,
                             long long Mp, unsigned short* __restrict__ It, unsigned short* __restrict__ Si) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (j >= M) return;
    const int vmin = lut[2 * k], range = lut[2 * k + 1];
    const double v = k < K ? Dt[(long long)k * Mp + j] : S[j];
    const unsigned short idx = range > 0 ? (unsigned short)((int)v - vmin) : (unsigned short)0;
    if (k < K)
        It[(long long)k * Mp + j] = idx;
    else
        Si[j] = idx;
}

// Dataset set-up: weighted count histograms of the regions, hist[hist_off[k] + D_jk - gmin_k] += cluster_size_j.  The sums
// are integers held in doubles (exact, so the order of the atomic additions does not matter).
__global__ void hist_kernel(const double* __restrict__ Dt, const int* __restrict__ w, const int* __restrict__ glut,
                            const unsigned* __restrict__ hist_off, int M, long long Mp, double* __restrict__ hist) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (j >= M) return;
    const int wj = w[j];
    if (wj == 0) return;
    const int i = (int)Dt[(long long)k * Mp + j] - glut[2 * k];
    if (i >= 0 && i < glut[2 * k + 1]) atomicAdd(hist + hist_off[k] + i, (double)wj);
}

// Test hook: the device lgamma on caller-provided arguments (scicone_test_lgamma).
__global__ void lgam_probe_kernel(const double* __restrict__ x, double* __restrict__ out, int n) {
    __shared__ double tbl[kLogTableDoubles];
    load_log_table<256>(tbl);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lgam(x[i], tbl);
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
__global__ void argmax_kernel(const double* __restrict__ pool, long long stride, const unsigned* __restrict__ rows, const int* __restrict__ ids,
                              int n_rows, int M, int* __restrict__ out_id, int* __restrict__ out_col) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    int best = 0;
    double bv = pool[(long long)rows[0] * stride + j];
    for (int c = 1; c < n_rows; ++c) {
        const double v = pool[(long long)rows[c] * stride + j];
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

__global__ void gather_rows_kernel(const double* __restrict__ pool, long long stride, const unsigned* __restrict__ rows, int n_rows, int M,
                                   double* __restrict__ out /*[M][n_rows]*/) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, j = j0 + threadIdx.x;
        if (c < n_rows && j < M) tile[r][threadIdx.x] = pool[(long long)rows[c] * stride + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, c = c0 + threadIdx.x;
        if (c < n_rows && j < M) out[(long long)j * n_rows + c] = tile[threadIdx.x][r];
    }
}

}  // namespace scicone
