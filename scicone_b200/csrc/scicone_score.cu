// libscicone_score.so - host side of the B200 tree-scoring engine behind include/scicone_score.h.
//
// What lives where
//   dataset  device copy of the count matrix (region-major Dt[K][Mp], 16-bit index matrix It, count-range table), read
//            totals S, cluster sizes w, the ROW POOL - one contiguous virtual address range per rank whose rows are
//            addressed by 32-bit indices (physical memory is mapped in as the pool grows, so the base never moves) -
//            the launch streams, the pinned staging arena through which proposals travel, the reduction scratch, the
//            NCCL communicator and the peer-mapped gather buffers of a cell-sharded run.
//   chain    what one reference `Inference` object owns for scoring (scicone/src/Inference.h:39-44): the committed
//            table t_scores as {node id -> score row, Node::z}, t_sums / t_maxs, the primed rows of the queued proposal.
//
// A proposal is compiled on the host into a small position-independent "blob" (header, one task per rescored node in
// parent-before-child order, aggregate bookkeeping - all rows by index) plus the K1 work it needs (one item per rescored
// node with its events, the slices of a root score) and evaluated by ONE launch of build_increments_kernel +
// score_proposals_kernel (kernels.cuh) - for any number of chains at once (grid.y).  Because nothing in a compiled
// proposal is an address, the owner of a chain in a cell-sharded run compiles it ONCE and ships the bytes
// (scicone_export_t_prime / scicone_import_t_prime); the other ranks only stage them.  Accept / reject are index swaps;
// no score is ever copied.
//
// There is no CPU implementation of any scoring step in this file: without a CUDA device every
// compute entry point fails with SCICONE_ERR_CUDA.
#include "../../include/scicone_score.h"
#include "kernels.cuh"
#include "condense.cuh"
#include "parallel_for.hpp"

#include <cuda.h>  // types of the virtual-memory-management API only: the entry points are resolved at run time
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

using namespace scicone;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return fail(SCICONE_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                             \
    } while (0)

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time so that single-GPU users need no NCCL at all.
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommSplit)(void*, int, int, void**, void*) = nullptr;  // NCCL >= 2.18
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handle) return false;
        GetUniqueId = (int (*)(UniqueId*))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (int (*)(void**, int, UniqueId, int))dlsym(handle, "ncclCommInitRank");
        CommDestroy = (int (*)(void*))dlsym(handle, "ncclCommDestroy");
        AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(handle, "ncclAllGather");
        CommSplit = (int (*)(void*, int, int, void**, void*))dlsym(handle, "ncclCommSplit");
        GetErrorString = (const char* (*)(int))dlsym(handle, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllGather;
    }
};
NcclApi g_nccl;
constexpr int kNcclDouble = 8;  // ncclFloat64

// ------------------------------------------------------------------------------------------------
// CUDA virtual memory management (driver API), resolved through the runtime so that the library links against
// nothing but libcudart and still loads on a machine without a driver (tests/test_abi_cpu.py).
// ------------------------------------------------------------------------------------------------
struct DriverApi {
    CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*Granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool loaded = false;
    bool load() {
        if (loaded) return true;
        auto get = [](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult st;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess && *fn;
        };
        loaded = get("cuMemAddressReserve", (void**)&AddressReserve) && get("cuMemAddressFree", (void**)&AddressFree) &&
                 get("cuMemCreate", (void**)&Create) && get("cuMemRelease", (void**)&Release) && get("cuMemMap", (void**)&Map) &&
                 get("cuMemUnmap", (void**)&Unmap) && get("cuMemSetAccess", (void**)&SetAccess) &&
                 get("cuMemGetAllocationGranularity", (void**)&Granularity);
        get("cuTensorMapEncodeTiled", (void**)&TensorMapEncodeTiled);  // optional: K0 falls back to one-row copies without it
        return loaded;
    }
};
DriverApi g_drv;
std::mutex g_drv_mu;

// ------------------------------------------------------------------------------------------------
// Pool of device rows (Mp doubles each), addressed by index: row r lives at base + r * Mp.  The index space is cut into
// one partition per rank of a cell-sharded run; a rank hands out rows of its own partition only, so the rows of a chain
// (allocated by the chain's owner rank) never collide with those of a chain another rank owns, and a compiled proposal
// names the same rows on every rank.  Physical memory is mapped chunk by chunk into the reserved range as a partition
// grows - on the owner when it allocates, on the other ranks when a proposal that uses the rows arrives.
// ------------------------------------------------------------------------------------------------
struct RowPool {
    size_t Mp = 0, row_bytes = 0;
    int device = 0;
    CUdeviceptr base = 0;
    size_t va_bytes = 0;
    static constexpr size_t kPartRows = 65536;  // rows of the index space per rank: the same on every rank
    size_t part_rows = 0, chunk_rows = 0, part_cap = 0;
    int n_parts = 1, my_part = 0;
    struct Part {
        size_t mapped_rows = 0;
        std::vector<CUmemGenericAllocationHandle> handles;
    };
    std::vector<Part> parts;
    std::vector<unsigned> free_rows;
    size_t next_row = 0;  // rows of this rank's partition that were ever handed out
    size_t in_use = 0;
    std::mutex mu;  // chains of one dataset may be compiled / committed from several host threads

    double* base_ptr() const { return reinterpret_cast<double*>(base); }
    void release_all() {
        for (int p = 0; p < (int)parts.size(); ++p) {
            Part& pt = parts[p];
            if (pt.mapped_rows) g_drv.Unmap(base + (CUdeviceptr)((size_t)p * part_rows * row_bytes), pt.mapped_rows * row_bytes);
            for (CUmemGenericAllocationHandle h : pt.handles) g_drv.Release(h);
            pt.handles.clear();
            pt.mapped_rows = 0;
        }
        if (base) g_drv.AddressFree(base, va_bytes);
        base = 0;
        va_bytes = 0;
    }
    int init(size_t mp, int dev, int parts_n, int mine) {
        {
            std::lock_guard<std::mutex> lk(g_drv_mu);
            if (!g_drv.load()) return fail(SCICONE_ERR_CUDA, "the CUDA driver does not offer the virtual memory management API");
        }
        if (in_use) return fail(SCICONE_ERR_STATE, "the row pool cannot be re-partitioned while rows are in use (attach the communicator before creating chains)");
        if (base) release_all();
        device = dev;
        Mp = mp;
        row_bytes = Mp * sizeof(double);
        n_parts = parts_n;
        my_part = mine;
        free_rows.clear();
        next_row = 0;
        CUmemAllocationProp prop;
        memset(&prop, 0, sizeof prop);
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        size_t gran = 0;
        if (g_drv.Granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
            return fail(SCICONE_ERR_CUDA, "cuMemGetAllocationGranularity failed");
        // chunk = the smallest multiple of 64 rows whose byte size is a multiple of the granularity and at least 32 MB
        chunk_rows = 64;
        while ((chunk_rows * row_bytes) % gran != 0 || chunk_rows * row_bytes < (size_t(32) << 20)) chunk_rows += 64;
        // The partition size is part of the row INDEX (row = partition * part_rows + r) and compiled proposals travel between
        // ranks, so it must not depend on anything rank-local (the chunk size does: it follows from this shard's padded cell
        // count).  Rows are padded to 256 cells, so partition bases are 2 MB aligned.  A partition is usable up to its last
        // whole chunk.
        part_rows = kPartRows;
        part_cap = part_rows / chunk_rows * chunk_rows;
        if (part_cap == 0) return fail(SCICONE_ERR_CUDA, "row pool: a mapping chunk of %zu rows exceeds the partition of %zu rows", chunk_rows, part_rows);
        va_bytes = part_rows * row_bytes * n_parts;
        if (g_drv.AddressReserve(&base, va_bytes, gran, 0, 0) != CUDA_SUCCESS) {
            base = 0;
            return fail(SCICONE_ERR_CUDA, "cuMemAddressReserve of %zu bytes failed", va_bytes);
        }
        parts.assign(n_parts, Part());
        return SCICONE_OK;
    }
    // the first `rows` rows of partition `part` are backed by memory (callers hold `mu`)
    int ensure_mapped_locked(int part, size_t rows) {
        if (part < 0 || part >= n_parts) return fail(SCICONE_ERR_ARG, "row pool: partition %d out of range", part);
        Part& pt = parts[part];
        if (rows <= pt.mapped_rows) return SCICONE_OK;
        if (rows > part_cap) return fail(SCICONE_ERR_CUDA, "row pool: partition of %zu rows exhausted", part_cap);
        CUDA_TRY(cudaSetDevice(device));
        CUmemAllocationProp prop;
        memset(&prop, 0, sizeof prop);
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        CUmemAccessDesc acc;
        memset(&acc, 0, sizeof acc);
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        while (pt.mapped_rows < rows) {
            const size_t bytes = chunk_rows * row_bytes;
            const CUdeviceptr at = base + (CUdeviceptr)(((size_t)part * part_rows + pt.mapped_rows) * row_bytes);
            CUmemGenericAllocationHandle h;
            CUresult r = g_drv.Create(&h, bytes, &prop, 0);
            if (r != CUDA_SUCCESS) return fail(SCICONE_ERR_CUDA, "row pool: cuMemCreate of %zu bytes failed (%d)", bytes, (int)r);
            r = g_drv.Map(at, bytes, 0, h, 0);
            if (r == CUDA_SUCCESS) r = g_drv.SetAccess(at, bytes, &acc, 1);
            if (r != CUDA_SUCCESS) {
                g_drv.Release(h);
                return fail(SCICONE_ERR_CUDA, "row pool: mapping %zu bytes failed (%d)", bytes, (int)r);
            }
            pt.handles.push_back(h);
            pt.mapped_rows += chunk_rows;
        }
        return SCICONE_OK;
    }
    int ensure_mapped(int part, size_t rows) {
        std::lock_guard<std::mutex> lk(mu);
        return ensure_mapped_locked(part, rows);
    }
    // rows move in batches, one lock per batch (the per-chain stashes use these)
    int get_many(size_t n, std::vector<unsigned>& out) {
        std::lock_guard<std::mutex> lk(mu);
        while (free_rows.size() < n) {
            int rc = ensure_mapped_locked(my_part, next_row + chunk_rows);
            if (rc) return rc;
            const size_t upto = std::min(parts[my_part].mapped_rows, next_row + chunk_rows);
            for (size_t r = upto; r-- > next_row;) free_rows.push_back((unsigned)((size_t)my_part * part_rows + r));
            next_row = upto;
        }
        out.insert(out.end(), free_rows.end() - n, free_rows.end());
        free_rows.resize(free_rows.size() - n);
        in_use += n;
        return SCICONE_OK;
    }
    void put_many(std::vector<unsigned>& rows, size_t keep) {
        if (rows.size() <= keep) return;
        std::lock_guard<std::mutex> lk(mu);
        in_use -= rows.size() - keep;
        free_rows.insert(free_rows.end(), rows.begin() + keep, rows.end());
        rows.resize(keep);
    }
    void destroy() {
        if (g_drv.loaded) release_all();
        free_rows.clear();
    }
};

struct NodeRow {
    unsigned row;
    int z;  // Node::z, an int in the reference (scicone/src/Node.h:25)
};

// Host copy of a scicone_tree (the caller's buffers are only valid during the call).
struct TreeCopy {
    int n = 0;
    double nu = 0.0;
    std::vector<int> id, parent, chg_off, chg_region, chg_value, gen_off, gen_region, gen_value;
    void assign(const scicone_tree* t) {
        n = t->n_nodes;
        nu = t->nu;
        id.assign(t->node_id, t->node_id + n);
        parent.assign(t->parent, t->parent + n);
        chg_off.assign(t->chg_off, t->chg_off + n + 1);
        chg_region.assign(t->chg_region, t->chg_region + chg_off[n]);
        chg_value.assign(t->chg_value, t->chg_value + chg_off[n]);
        gen_off.assign(t->gen_off, t->gen_off + n + 1);
        gen_region.assign(t->gen_region, t->gen_region + gen_off[n]);
        gen_value.assign(t->gen_value, t->gen_value + gen_off[n]);
    }
    // node->c.count(k) ? node->c[k] : 0   (scicone/src/Tree.h:189-190)
    int genotype_at(int node, int k) const {
        const int* b = gen_region.data() + gen_off[node];
        const int* e = gen_region.data() + gen_off[node + 1];
        const int* it = std::lower_bound(b, e, k);
        return (it != e && *it == k) ? gen_value[it - gen_region.data()] : 0;
    }
};

int validate_tree(const scicone_tree* t, int K) {
    if (!t || t->n_nodes <= 0 || !t->node_id || !t->parent || !t->chg_off || !t->gen_off)
        return fail(SCICONE_ERR_ARG, "tree: null or empty");
    if (t->parent[0] != -1) return fail(SCICONE_ERR_ARG, "tree: entry 0 must be the root");
    for (int i = 1; i < t->n_nodes; ++i)
        if (t->parent[i] < 0 || t->parent[i] >= i)
            return fail(SCICONE_ERR_ARG, "tree: entries must be parent-before-child (entry %d has parent %d)", i, t->parent[i]);
    for (int i = 0; i < t->n_nodes; ++i) {
        if (t->node_id[i] < 0) return fail(SCICONE_ERR_ARG, "tree: negative node id");
        if (t->chg_off[i + 1] < t->chg_off[i] || t->gen_off[i + 1] < t->gen_off[i])
            return fail(SCICONE_ERR_ARG, "tree: CSR offsets not monotone");
        for (int e = t->chg_off[i]; e < t->chg_off[i + 1]; ++e)
            if (t->chg_region[e] < 0 || t->chg_region[e] >= K)  // reference: D.size() != r.size() -> std::logic_error
                return fail(SCICONE_ERR_ARG, "tree: event region %d outside [0,%d)", t->chg_region[e], K);
        for (int e = t->gen_off[i]; e < t->gen_off[i + 1]; ++e)
            if (t->gen_region[e] < 0 || t->gen_region[e] >= K)
                return fail(SCICONE_ERR_ARG, "tree: genotype region %d outside [0,%d)", t->gen_region[e], K);
    }
    return SCICONE_OK;
}

// Ordered map on a sorted vector (find / emplace / erase / operator[] / ascending iteration, the subset of std::map used
// here).  The tables have tens to hundreds of entries and are rebuilt for every proposal: contiguous storage keeps the
// host side of a proposal free of per-entry heap allocations.
template <class K, class V>
class FlatMap {
public:
    typedef std::pair<K, V> value_type;
    typedef typename std::vector<value_type>::iterator iterator;
    iterator begin() { return v_.begin(); }
    iterator end() { return v_.end(); }
    size_t size() const { return v_.size(); }
    bool empty() const { return v_.empty(); }
    void clear() { v_.clear(); }
    iterator find(const K& k) {
        iterator it = lower(k);
        return (it != v_.end() && it->first == k) ? it : v_.end();
    }
    size_t count(const K& k) { return find(k) != v_.end() ? 1 : 0; }
    std::pair<iterator, bool> emplace(const K& k, const V& val) {
        iterator it = lower(k);
        if (it != v_.end() && it->first == k) return std::make_pair(it, false);
        return std::make_pair(v_.insert(it, value_type(k, val)), true);
    }
    V& operator[](const K& k) { return emplace(k, V()).first->second; }
    void erase(iterator it) { v_.erase(it); }
    void swap(FlatMap& o) { v_.swap(o.v_); }

private:
    iterator lower(const K& k) {
        return std::lower_bound(v_.begin(), v_.end(), k, [](const value_type& e, const K& key) { return e.first < key; });
    }
    std::vector<value_type> v_;
};

}  // namespace

// ================================================================================================
struct scicone_dataset {
    int M = 0, K = 0;
    size_t Mp = 0;
    int device = 0;
    int n_sm = 148;
    double eta = 0;
    bool od = false, maxs = false;
    long long cell_offset = 0, M_global = 0;
    std::vector<int> r, neutral;
    int root_z = 0;  // Tree::compute_root_score, Tree.h:145-160
    long long sum_r = 0;

    double* dDt = nullptr;
    double* dS = nullptr;
    unsigned short* dIt = nullptr;  // [K][Mp] 16-bit count indices of the tabulated (region, block) pairs
    unsigned short* dSi = nullptr;  // [Mp] the same for the read totals
    // count-range table of K1 (kernels.cuh): [K + 1 rows: regions, then S]{min, range}; the host keeps a copy to lay out the
    // value tables of a launch
    int* d_lut = nullptr;
    std::vector<int> lut;
    bool all_tab = false;    // every row has a table: K1b needs no direct lgamma evaluation
    // Root score from count histograms (kernels.cuh, ITEM_OD_HIST): when every region's counts are integers in a narrow
    // range, the region part of Tree::get_od_root_score summed over the cells is a function of how often each count occurs.
    // The histograms cover ALL cells of the run: a cell-sharded dataset sums its shard's histograms with the other ranks'
    // at scicone_dataset_attach_comm (integers: exact), so every rank holds the same table and the root score does not
    // depend on the number of GPUs.  raw_*: per row of counts (regions, then S) {min, max, all integers} of this shard.
    std::vector<double> raw_lo, raw_hi;
    std::vector<char> raw_int;
    bool od_hist_allowed = true;  // SCICONE_OD_HIST=0: per-cell slices as for non-integer data (test hook)
    bool od_hist = false;
    std::vector<int> glut;            // [K][2] {gmin_k, range_k} over all cells of the run
    std::vector<unsigned> hist_off;   // [K + 1]
    int* d_glut = nullptr;
    unsigned* d_hist_off = nullptr;
    double* d_hist = nullptr;
    double* d_tables[4] = {nullptr, nullptr, nullptr, nullptr};  // per lane: value tables of the launch in flight
    size_t tables_cap = 0;                     // doubles per lane
    int* dW = nullptr;
    cudaStream_t stream = nullptr;
    RowPool pool;

    // staging: a ring of pinned segments (proposal blobs out, totals back) so that launches can be queued
    // without a host round trip; one device arena per lane (copies and kernels are ordered by the stream)
    struct Segment {
        unsigned char* h_arena = nullptr;
        double* h_total = nullptr;     // pinned [slots][2]
        unsigned* h_status = nullptr;  // pinned [slots]
        unsigned long long* h_flag = nullptr;  // pinned: ticket of the launch whose results are in h_total / h_status
        double* h_gather = nullptr;    // pinned [world][slots][2][max_chunks_rank]
        cudaEvent_t evb = nullptr, ev0 = nullptr, ev1 = nullptr, done = nullptr;
        int n = 0, n_items = 0;
        std::vector<char> with_od;
        bool busy = false;
        bool direct = false;  // the kernel writes totals, status and the ticket into this segment's pinned buffers
        bool summed = false;  // cell-sharded: h_total already holds the sums over all ranks (gather_totals_kernel)
        unsigned long long ticket = 0;
        int lane = 0;
        std::vector<unsigned> temp_rows;  // rows that only live for this launch: back to the pool when it is collected
    };
    static constexpr int kRing = 8;
    Segment seg[kRing];
    unsigned long long next_ticket = 1;
    // Launch lanes: lane l has its own stream, device arena and range of reduction slots.  One lane (default): launches
    // are serialised by the stream, so accept / reject may be issued before the wait.  Two lanes (scicone_set_streams):
    // consecutive launches alternate lanes and overlap on the device; the caller must then collect a launch before it
    // settles its chains.
    static constexpr int kLanes = 4;
    int n_lanes = 1;
    cudaStream_t lane_stream[kLanes] = {nullptr, nullptr, nullptr, nullptr};
    unsigned char* d_arena[kLanes] = {nullptr, nullptr, nullptr, nullptr};
    size_t arena_cap = 0;
    // per batch slot: reduction scratch and results
    int n_cta = 0, n_chunks = 0;
    int cs_stride = 0;  // chunk sums per slot and value: n_chunks, or the widest shard's count on a cell-sharded dataset
    int slots = 0;
    double* d_partial = nullptr;  // [lanes][slots][2][n_cta]   (0: tree score, 1: root score)
    double* d_chunk = nullptr;    // [lanes][slots][2][cs_stride]
    double* d_total = nullptr;    // per lane: [slots][2] doubles, then [slots] status words
    size_t lane_result_bytes = 0;  // stride of one lane's result block
    unsigned* d_counter = nullptr;         // [lanes][slots]
    unsigned* d_launch_counter = nullptr;  // [lanes]
    // multi-GPU
    void* comm = nullptr;
    void* comm_lane[4] = {nullptr, nullptr, nullptr, nullptr};  // lanes 1..: a communicator of their own (ncclCommSplit of `comm`), see scicone_set_streams
    int rank = 0, world = 1;
    int max_chunks_rank = 0;
    double* d_gather = nullptr;  // NCCL variant, per lane: [world][slots][2][max_chunks_rank]
    size_t gather_cap = 0;
    // Exchange of the chunk sums through peer memory (default on cell-sharded datasets; SCICONE_EXCHANGE=nccl selects the
    // all-gather instead): every rank owns one buffer - flags [kLanes][world], then sums
    // [kLanes][2 (launch parity)][world][kXchgSlots][2][max_chunks_rank] - mapped into all the others (CUDA IPC), and the
    // last CTA of a launch stores the launch's sums and its flag into all of them (finish_reduction, kernels.cuh).
    bool xchg_on = false;
    unsigned char* xchg = nullptr;
    std::vector<void*> xchg_peer;                     // base of rank p's buffer in this process (own: xchg)
    double** d_xchg_data_tab = nullptr;               // device: [kLanes][2][world] where this rank writes in rank p's buffer
    unsigned long long** d_xchg_flag_tab = nullptr;   // device: [kLanes][world]
    unsigned long long lane_seq[kLanes] = {0, 0, 0, 0};     // launches submitted per lane (the value published in the flags)
    size_t xchg_flag_bytes() const { return round_up(sizeof(unsigned long long) * kLanes * world, 16); }
    size_t xchg_block() const { return (size_t)SCICONE_MAX_SHARDED_BATCH * 2 * max_chunks_rank; }  // doubles of one rank in one (lane, parity) block
    unsigned long long* xchg_flags(void* base, int lane, int rk) const {
        return reinterpret_cast<unsigned long long*>(base) + (size_t)lane * world + rk;
    }
    double* xchg_data(void* base, int lane, int parity, int rk) const {
        return reinterpret_cast<double*>(static_cast<unsigned char*>(base) + xchg_flag_bytes()) + (((size_t)lane * 2 + parity) * world + rk) * xchg_block();
    }
    // profiling
    bool profiling = false;
    bool direct_results = true;  // SCICONE_DIRECT_RESULTS=0: results travel by a stream-ordered copy + event instead
    bool pdl = true;             // SCICONE_PDL=0: K1a, K1b and the proposal kernel are launched the ordinary, fully serialised way
    double last_kernel_us = 0.0, sum_kernel_us = 0.0, sum_build_us = 0.0;
    unsigned long long n_timed = 0, n_build_timed = 0;
    unsigned long long n_launches = 0;
    // bookkeeping for the bench: bytes staged per direction, algorithmic bytes / scores of the launched proposals
    unsigned long long h2d_bytes = 0, d2h_bytes = 0;
    double alg_bytes = 0.0, n_scores = 0.0;
    double k1_bytes = 0.0, k1_lgamma = 0.0;  // K1a / K1b: algorithmic bytes and lgamma evaluations
    cudaEvent_t region_ev[2] = {nullptr, nullptr};

    int set_device() const {
        CUDA_TRY(cudaSetDevice(device));
        return SCICONE_OK;
    }
    int sync_lanes() {  // every queued launch on every lane has finished
        for (int l = 0; l < kLanes; ++l)
            if (lane_stream[l]) CUDA_TRY(cudaStreamSynchronize(lane_stream[l]));
        return SCICONE_OK;
    }
    int ensure_arena(size_t bytes) {
        if (bytes <= arena_cap) return SCICONE_OK;
        size_t cap = std::max<size_t>(bytes * 2, size_t(4) << 20);
        // every queued launch has finished after the synchronize; the totals of tickets that were not collected yet sit in
        // their own pinned buffers (h_total), which are not touched here
        int rc0 = sync_lanes();
        if (rc0) return rc0;
        for (Segment& g : seg) {
            if (g.h_arena) cudaFreeHost(g.h_arena);
            g.h_arena = nullptr;
        }
        for (int l = 0; l < kLanes; ++l) {
            if (d_arena[l]) cudaFree(d_arena[l]);
            d_arena[l] = nullptr;
        }
        arena_cap = 0;
        for (Segment& g : seg) CUDA_TRY(cudaMallocHost(&g.h_arena, cap));
        for (int l = 0; l < kLanes; ++l) CUDA_TRY(cudaMalloc(&d_arena[l], cap));
        arena_cap = cap;
        return SCICONE_OK;
    }
    int ensure_tables(size_t doubles) {
        if (doubles <= tables_cap) return SCICONE_OK;
        int rc0 = sync_lanes();
        if (rc0) return rc0;
        const size_t cap = std::max<size_t>(doubles * 2, size_t(1) << 20);
        for (int l = 0; l < kLanes; ++l) {
            if (d_tables[l]) cudaFree(d_tables[l]);
            d_tables[l] = nullptr;
        }
        tables_cap = 0;
        for (int l = 0; l < kLanes; ++l) CUDA_TRY(cudaMalloc(&d_tables[l], sizeof(double) * cap));
        tables_cap = cap;
        return SCICONE_OK;
    }
    double* lane_total(int lane) const { return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(d_total) + lane * lane_result_bytes); }
    unsigned* lane_status(int lane) const { return reinterpret_cast<unsigned*>(lane_total(lane) + (size_t)slots * 2); }
    double* lane_partial(int lane) const { return d_partial + (size_t)lane * slots * 2 * n_cta; }
    double* lane_chunk(int lane) const { return d_chunk + (size_t)lane * slots * 2 * cs_stride; }
    int ensure_slots(int n) {
        if (n <= slots) return SCICONE_OK;
        int ns = std::max(n, std::max(4, slots * 2));
        if (world > 1) ns = std::min(std::max(n, ns), std::max(n, SCICONE_MAX_SHARDED_BATCH));
        int rc0 = sync_lanes();
        if (rc0) return rc0;
        for (Segment& g : seg)
            if (g.busy) return fail(SCICONE_ERR_STATE, "the batch size cannot grow while launches are outstanding (scicone_dataset_reserve sets it up front)");
        cudaFree(d_partial); cudaFree(d_chunk); cudaFree(d_total); cudaFree(d_counter);
        d_partial = d_chunk = d_total = nullptr;
        d_counter = nullptr;
        for (Segment& g : seg) {
            if (g.h_total) cudaFreeHost(g.h_total);
            if (g.h_gather) cudaFreeHost(g.h_gather);
            g.h_total = nullptr;
            g.h_status = nullptr;
            g.h_gather = nullptr;
        }
        slots = 0;
        const size_t all = size_t(ns) * kLanes;  // every lane has its own range of ns slots
        lane_result_bytes = round_up((sizeof(double) * 2 + sizeof(unsigned)) * ns, 16);
        CUDA_TRY(cudaMalloc(&d_partial, sizeof(double) * all * 2 * n_cta));
        CUDA_TRY(cudaMalloc(&d_chunk, sizeof(double) * all * 2 * cs_stride));
        CUDA_TRY(cudaMalloc(&d_total, lane_result_bytes * kLanes));  // per lane: totals, then the status words: one copy back
        CUDA_TRY(cudaMalloc(&d_counter, sizeof(unsigned) * all));
        CUDA_TRY(cudaMemset(d_counter, 0, sizeof(unsigned) * all));
        CUDA_TRY(cudaMemset(d_total, 0, lane_result_bytes * kLanes));
        CUDA_TRY(cudaMemset(d_chunk, 0, sizeof(double) * all * 2 * cs_stride));
        for (Segment& g : seg) {
            CUDA_TRY(cudaMallocHost(&g.h_total, lane_result_bytes + 16));
            g.h_status = reinterpret_cast<unsigned*>(g.h_total + (size_t)ns * 2);
            g.h_flag = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(g.h_total) + lane_result_bytes);
            *g.h_flag = 0ull;
        }
        slots = ns;
        if (world > 1) {
            cudaFree(d_gather);
            d_gather = nullptr;
            size_t per_rank = size_t(ns) * 2 * max_chunks_rank;
            CUDA_TRY(cudaMalloc(&d_gather, sizeof(double) * per_rank * world * kLanes));
            for (Segment& g : seg) CUDA_TRY(cudaMallocHost(&g.h_gather, sizeof(double) * per_rank * world));
            gather_cap = per_rank;
        }
        return SCICONE_OK;
    }
};

static_assert(SCICONE_MAX_SHARDED_BATCH == 128, "documented in the header");
constexpr size_t kStashRefill = 32, kStashMax = 160;

struct scicone_chain {
    scicone_dataset* ds = nullptr;
    bool follower = false;  // cell-sharded runs: this rank does not own the chain, it only stages the owner's compiled proposals
    FlatMap<int, NodeRow> t;  // committed table, ascending node id (std::map order of the reference)
    unsigned sums[2] = {kNoRow, kNoRow};   // rows: [0] committed t_sums, [1] primed
    unsigned maxid[2] = {kNoRow, kNoRow};  // rows holding the int32 argmax ids
    double t_nu = 0.0;  // Tree::nu of the committed tree
    bool od_p_valid = false;  // compute_t_prime_od_scores was already called for od_p_nu
    double od_p_nu = 0.0;
    bool with_od = false;     // the compiled proposal carries the root-score part
    // queued proposal
    bool pending = false;
    bool compiled = false;  // the queued proposal's blob is built (scicone_prepare_t_prime, an import, or the batch call)
    bool evaluated = false;
    TreeCopy tp;
    std::vector<int> roots;
    FlatMap<int, NodeRow> p;  // rescored node id -> primed row
    int deleted_id = -1;
    bool p_full = false;
    // compiled proposal: the blob, the K1 work it needs first (items + their events / per-region constants), and rows that
    // only live for this launch (root-score slices, increments of nodes that are rescored twice)
    std::vector<unsigned char> blob;
    std::vector<DevItem> items;
    std::vector<unsigned char> aux;
    std::vector<unsigned> temp_rows;
    double alg_bytes = 0.0, n_scores = 0.0;
    struct Scratch {
        std::vector<int> child_off, child, fill, ptask, task_of_node, stack;
        std::vector<DevTask> tasks;
        std::vector<unsigned> prow, old_rows, slice_rows;
        std::vector<DevUnch> unch;
    } scratch;
    // rows handed out by the dataset's pool in bulk: the chain takes and returns rows here without a lock, so chains
    // compiled / committed on different host threads do not meet at the pool's mutex on every row
    std::vector<unsigned> stash;
    int row_get(unsigned* out) {
        if (stash.empty()) {
            int rc = ds->pool.get_many(kStashRefill, stash);
            if (rc) return rc;
        }
        *out = stash.back();
        stash.pop_back();
        return SCICONE_OK;
    }
    void row_put(unsigned r) {
        if (r == kNoRow) return;
        stash.push_back(r);
        if (stash.size() > kStashMax) ds->pool.put_many(stash, kStashMax / 2);
    }
    LaunchEntry entry = {0u, 0u, 0u, 0u};  // first-chunk descriptor of the compiled blob (blob_off, slot are filled in at submit)
    double cost = 0.0;  // relative device time of the compiled proposal (orders the blobs inside a launch)
};

namespace {

void drop_pending(scicone_chain* ch) {
    for (auto& kv : ch->p) ch->row_put(kv.second.row);
    ch->p.clear();
    ch->roots.clear();
    ch->pending = false;
    ch->compiled = false;
    ch->evaluated = false;
    ch->deleted_id = -1;
    ch->p_full = false;
    for (unsigned r : ch->temp_rows) ch->row_put(r);
    ch->temp_rows.clear();
    ch->od_p_valid = false;
    ch->with_od = false;
}

template <class T>
size_t bytes_append(std::vector<unsigned char>& b, const T* v, size_t n) {
    size_t off = round_up(b.size(), 16);
    b.resize(off + std::max<size_t>(round_up(n * sizeof(T), 16), 16));
    if (n) memcpy(b.data() + off, v, n * sizeof(T));
    return off;
}
template <class T>
size_t bytes_append(std::vector<unsigned char>& b, const std::vector<T>& v) {
    return bytes_append(b, v.data(), v.size());
}

constexpr int kOdRegionsPerSlice = 6;  // 200 regions -> 32 slices: enough units for one nu proposal to fill the GPU
constexpr int kOdRegionsPerHistItem = 3, kMaxOdHistItems = 96;  // histogram form: ~7 lgamma per thread of a K1a item at 200 regions

// Root-score request (Tree::get_od_root_score, Tree.h:1774-1810): K1 items that reduce slices of the regions into
// temporary rows; the proposal kernel adds the slices and the two cell-level terms.
int add_od_request(scicone_chain* ch, double nu, DevHeader& H, std::vector<unsigned>& slice_rows) {
    scicone_dataset* ds = ch->ds;
    const int K = ds->K;
    std::vector<double> arg(2 * (size_t)K + ((size_t)K + 1) / 2, 0.0);  // arg[K], c[K], then room for K 32-bit table offsets (filled at submit)
    double* cc = arg.data() + K;
    if (ds->od) {  // Tree.h:1783-1797 (no eta substitution here, quirk Q2)
        H.od_nz = nu * ds->root_z;
        H.od_lg_nz = lgamma(H.od_nz);
        {   // lgamma(S + nu z_root) as a K1 row
            unsigned row = kNoRow;
            int rc = ch->row_get(&row);
            if (rc) return rc;
            ch->temp_rows.push_back(row);
            DevItem it;
            memset(&it, 0, sizeof it);
            it.kind = ITEM_G_ROW;
            it.dst = row;
            it.nz = H.od_nz;
            ch->items.push_back(it);
            H.od_g_row = row;
        }
        for (int k = 0; k < K; ++k) {
            arg[k] = nu * ds->neutral[k] * ds->r[k];
            cc[k] = lgamma(arg[k]);
        }
    } else {  // Tree.h:1799-1806
        H.od_lg_nz = log((double)ds->sum_r);
        for (int k = 0; k < K; ++k) arg[k] = log((double)ds->r[k]);
    }
    const unsigned aux_off = (unsigned)bytes_append(ch->aux, arg);
    if (ds->od_hist) {
        // The region part summed over the cells first (kernels.cuh, ITEM_OD_HIST): K1a items of a few regions each leave one
        // number apiece in the table arena; no slice rows, no per-cell look-ups.  The cell-level term lgamma(S + nu z_root)
        // stays per cell (the row requested above), so the sum over the cells keeps its chunk order.
        const int n_hist = std::max(1, std::min(kMaxOdHistItems, (K + kOdRegionsPerHistItem - 1) / kOdRegionsPerHistItem));
        for (int s = 0; s < n_hist; ++s) {
            DevItem it;
            memset(&it, 0, sizeof it);
            it.kind = ITEM_OD_HIST;
            it.dst = kNoRow;
            it.aux_off = aux_off;
            it.k0 = (int)((long long)K * s / n_hist);
            it.k1 = (int)((long long)K * (s + 1) / n_hist);
            ch->items.push_back(it);
        }
        H.n_od_slices = n_hist;
        H.flags |= PROP_WITH_OD | PROP_OD_HIST;
        return SCICONE_OK;
    }
    const int n_slices = std::max(1, std::min(kMaxOdSlices, K / kOdRegionsPerSlice));
    for (int s = 0; s < n_slices; ++s) {
        unsigned row = kNoRow;
        int rc = ch->row_get(&row);
        if (rc) return rc;
        ch->temp_rows.push_back(row);
        slice_rows.push_back(row);
        DevItem it;
        memset(&it, 0, sizeof it);
        it.kind = ds->od ? ITEM_SLICE_OD : ITEM_SLICE_MN;
        it.dst = row;
        it.aux_off = aux_off;  // relative to the chain's aux block; rebased at submit
        it.k0 = (int)((long long)K * s / n_slices);
        it.k1 = (int)((long long)K * (s + 1) / n_slices);
        ch->items.push_back(it);
    }
    H.n_od_slices = n_slices;
    H.flags |= PROP_WITH_OD;
    return SCICONE_OK;
}

// Compile the queued proposal of `ch` into ch->blob (+ ch->items, ch->aux).
int compile_proposal(scicone_chain* ch, bool full, bool full_with_od) {
    scicone_dataset* ds = ch->ds;
    const TreeCopy& T = ch->tp;
    const double nu = T.nu;
    const bool od = ds->od;
    ch->items.clear();
    ch->aux.clear();

    // scratch vectors live in the chain: a proposal is compiled every ~100 us per chain, so no heap traffic here
    scicone_chain::Scratch& W = ch->scratch;
    std::vector<int>&child_off = W.child_off, &child = W.child, &fill = W.fill;
    child_off.assign(T.n + 1, 0);
    child.assign(T.n > 1 ? T.n - 1 : 0, 0);
    for (int i = 1; i < T.n; ++i) child_off[T.parent[i] + 1]++;
    for (int i = 0; i < T.n; ++i) child_off[i + 1] += child_off[i];
    fill.assign(child_off.begin(), child_off.end() - 1);
    for (int i = 1; i < T.n; ++i) child[fill[T.parent[i]]++] = i;

    std::vector<DevTask>& tasks = W.tasks;
    std::vector<unsigned>& prow = W.prow;  // per task: row that holds its parent's score
    std::vector<int>& ptask = W.ptask;     // per task: task index of the parent, -1 when the parent is not rescored before
    std::vector<int>& task_of_node = W.task_of_node;
    std::vector<int>& stack = W.stack;
    tasks.clear();
    prow.clear();
    ptask.clear();
    stack.clear();
    task_of_node.assign(T.n, -1);
    size_t n_events = 0;
    std::vector<DevEvent> evs;  // events of the node being compiled
    for (size_t ri = 0; ri < ch->roots.size(); ++ri) {
        const int root = ch->roots[ri];
        stack.push_back(root);
        while (!stack.empty()) {
            const int n = stack.back();
            stack.pop_back();
            const int id = T.id[n];
            const int pidx = T.parent[n];
            if (tasks.size() >= 32000) return fail(SCICONE_ERR_ARG, "more than 32000 rescored nodes in one proposal");
            DevTask tk;
            memset(&tk, 0, sizeof tk);
            tk.node_id = id;
            int parent_task = -1;
            unsigned parent_row = kNoRow;
            DevItem item;
            memset(&item, 0, sizeof item);
            int z_int;
            if (pidx < 0) {  // Tree::compute_root_score, Tree.h:145-160
                tk.flags = TASK_ROOT;
                z_int = ds->root_z;
            } else {
                const int pid = T.id[pidx];
                int pz;
                auto pit = ch->p.find(pid);
                if (n == root) {
                    // Inference.h:1034: the parent's score is restored from the COMMITTED table;
                    // its Node::z is whatever the tree copy carries (primed if rescored before, else committed).
                    auto cit = ch->t.find(pid);
                    if (cit == ch->t.end())
                        return fail(SCICONE_ERR_STATE, "parent id %d of the rescored subtree is not in the committed table", pid);
                    parent_row = cit->second.row;
                    pz = (pit != ch->p.end()) ? pit->second.z : cit->second.z;
                } else {
                    parent_task = task_of_node[pidx];
                    parent_row = tasks[parent_task].dst;
                    pz = pit->second.z;
                }
                // Tree::compute_score, Tree.h:177-198: z = parent.z + sum_k r_k (cn_node - cn_parent)
                double zd = (double)pz;
                const double zp = (double)pz;
                evs.clear();
                for (int e = T.chg_off[n]; e < T.chg_off[n + 1]; ++e) {
                    const int k = T.chg_region[e];
                    const int cn_i = T.genotype_at(n, k) + ds->neutral[k];
                    const int cp_i = T.genotype_at(pidx, k) + ds->neutral[k];
                    const double node_cn = cn_i == 0 ? ds->eta : (double)cn_i;
                    const double parent_cn = cp_i == 0 ? ds->eta : (double)cp_i;
                    zd += ds->r[k] * (node_cn - parent_cn);
                    DevEvent ev;
                    memset(&ev, 0, sizeof ev);
                    ev.k = k;
                    if (od) {  // Tree.h:214-218: x += lgamma(D_k + argN) - lgamma(D_k + argP); x -= lgamma(argN); x += lgamma(argP)
                        ev.argN = nu * node_cn * ds->r[k];
                        ev.argP = nu * parent_cn * ds->r[k];
                        ev.a = lgamma(ev.argN);
                        ev.b = lgamma(ev.argP);
                    } else  // Tree.h:221
                        ev.a = log(node_cn) - log(parent_cn);
                    evs.push_back(ev);
                }
                item.kind = od ? ITEM_NODE_OD : ITEM_NODE_MN;
                item.n_ev = (int)evs.size();
                item.aux_off = evs.empty() ? 0u : (unsigned)bytes_append(ch->aux, evs);
                n_events += evs.size();
                if (od) {  // Tree.h:227-231
                    item.nz = nu * zd;
                    item.nzp = nu * zp;
                    item.lg_nz = lgamma(item.nz);
                    item.lg_nzp = lgamma(item.nzp);
                } else {  // Tree.h:236-237
                    item.lg_nz = log(zd);
                    item.lg_nzp = log(zp);
                }
                z_int = (int)zd;  // Tree.h:241 into the int Node::z (quirk Q1)
            }
            auto own = ch->p.find(id);
            if (own == ch->p.end()) {
                unsigned row = kNoRow;
                int rc = ch->row_get(&row);
                if (rc) return rc;
                own = ch->p.emplace(id, NodeRow{row, z_int}).first;
                tk.inc = row;
            } else {
                // the node is rescored a second time by this proposal (overlapping subtree roots of a label swap): the score
                // goes to the same row, the increment of this visit waits in a row of its own
                own->second.z = z_int;
                unsigned row = kNoRow;
                int rc = ch->row_get(&row);
                if (rc) return rc;
                ch->temp_rows.push_back(row);
                tk.inc = row;
            }
            tk.dst = own->second.row;
            if (!(tk.flags & TASK_ROOT)) {
                item.dst = tk.inc;
                ch->items.push_back(item);
            }
            task_of_node[n] = (int)tasks.size();
            tasks.push_back(tk);
            prow.push_back(parent_row);
            ptask.push_back(parent_task);
            for (int c = child_off[n + 1] - 1; c >= child_off[n]; --c) stack.push_back(child[c]);
        }
    }

    // aggregate bookkeeping: Inference::compute_t_prime_sums, Inference.h:1069-1196
    if (ch->roots.size() > 1) {  // a node written twice (overlapping subtree roots) enters the aggregate once, with its last value
        std::map<int, int> last;
        for (size_t i = 0; i < tasks.size(); ++i) last[tasks[i].node_id] = (int)i;
        for (size_t i = 0; i < tasks.size(); ++i)
            if (last[tasks[i].node_id] != (int)i) tasks[i].flags |= TASK_SKIP_AGG;
    }
    // Walk bookkeeping for the one-thread-per-cell kernel.  Parent scores: a parent visited at most kScRing nodes ago is
    // still in the thread's column of the shared-memory ring (the node just before: in a register); the committed row of a
    // subtree root's parent, or a parent rescored further back, is read from its row (plain load, one node ahead).
    for (size_t i = 0; i < tasks.size(); ++i) {
        if (tasks[i].flags & TASK_ROOT) continue;
        const int back = ptask[i] >= 0 ? (int)i - ptask[i] : 0;
        if (back >= 1 && back <= kScRing)
            tasks[i].parent_back = (unsigned)back;
        else {
            tasks[i].parent_row = prow[i];
            tasks[i].flags |= TASK_PS_FAR;
        }
    }
    std::vector<unsigned>& old_rows = W.old_rows;
    std::vector<DevUnch>& unch = W.unch;
    old_rows.clear();
    unch.clear();
    ch->deleted_id = -1;
    if (!full) {
        // Inference::deleted_node_idx, Inference.h:1270-1289
        if ((size_t)T.n < ch->t.size()) {
            for (auto& kv : ch->t) {
                bool found = false;
                for (int i = 0; i < T.n && !found; ++i) found = (T.id[i] == kv.first);
                if (!found) ch->deleted_id = kv.first;
            }
        }
        for (auto& kv : ch->p) {
            auto it = ch->t.find(kv.first);
            if (it != ch->t.end()) old_rows.push_back(it->second.row);
        }
        if (ch->deleted_id != -1) old_rows.push_back(ch->t[ch->deleted_id].row);
        for (auto& kv : ch->t)
            if (!ch->p.count(kv.first) && kv.first != ch->deleted_id) {
                DevUnch u;
                u.row = kv.second.row;
                u.id = kv.first;
                unch.push_back(u);
            }
    }

    DevHeader H;
    memset(&H, 0, sizeof H);
    H.n_tasks = (int)tasks.size();
    H.n_old = (int)old_rows.size();
    H.n_unch = (int)unch.size();
    H.deleted_id = ch->deleted_id;
    H.flags = (ds->maxs ? PROP_MAX : 0u) | (od ? PROP_OD : 0u) | (full ? PROP_FULL : 0u);
    if (!full && ch->p.size() >= unch.size()) H.flags |= PROP_SCRATCH;  // MathOp.cpp:315
    H.sums_old = ch->sums[0];
    H.sums_new = ch->sums[1];
    H.maxid_old = ch->maxid[0];
    H.maxid_new = ch->maxid[1];
    H.od_g_row = kNoRow;
    // a proposal that changes nu also needs Inference::compute_t_prime_od_scores (Inference.h:942): evaluated by
    // the same launch unless the caller already asked for it separately
    std::vector<unsigned>& slice_rows = W.slice_rows;
    slice_rows.clear();
    ch->with_od = full ? full_with_od : (nu != ch->t_nu && !(ch->od_p_valid && ch->od_p_nu == nu));
    const bool whole_tree = full || (ch->roots.size() == 1 && T.parent[ch->roots[0]] < 0);
    {   // SURVEY.md section 8d.  Delta move: 16 + 8 E_n bytes per rescored cell x node score, 8 N + 12 per cell aggregate.
        // Full pass / nu move: 8 K (the cell's row of D) + 8 N (scores written) + 8 N + 12 (aggregate).
        const double n_table = full ? (double)ch->p.size() : (double)(unch.size() + ch->p.size());
        if (whole_tree)
            ch->alg_bytes = (double)ds->M * (8.0 * ds->K + 8.0 * tasks.size() + 8.0 * n_table + 12.0);
        else
            ch->alg_bytes = (double)ds->M * (16.0 * tasks.size() + 8.0 * n_events + 8.0 * n_table + 12.0);
        ch->n_scores = (double)ds->M * tasks.size();
        ch->cost = 2.0 * tasks.size() + (full ? 0.0 : 0.5 * unch.size());
    }
    if (ch->with_od) {
        int rc = add_od_request(ch, nu, H, slice_rows);
        if (rc) return rc;
        if (!whole_tree) ch->alg_bytes += (double)ds->M * 8.0 * ds->K;
    }
    std::vector<unsigned char>& b = ch->blob;
    b.clear();
    b.resize(sizeof(DevHeader));
    const size_t off_tasks = bytes_append(b, tasks);
    H.off_old = (unsigned)bytes_append(b, old_rows);
    H.off_unch = (unsigned)bytes_append(b, unch);
    H.off_od_rows = (unsigned)bytes_append(b, slice_rows);
    memcpy(b.data(), &H, sizeof H);
    if (off_tasks != sizeof(DevHeader)) return fail(SCICONE_ERR_STATE, "internal: task records must follow the header");
    ch->entry = LaunchEntry{0u, (unsigned)std::min<size_t>(tasks.size(), kTaskChunk), 0u, 0u};
    return SCICONE_OK;
}

// A blob that only evaluates the root score at `nu` (Inference::compute_t_od_scores, Inference.h:1406-1419).
int compile_od_only(scicone_chain* ch, double nu) {
    ch->items.clear();
    ch->aux.clear();
    DevHeader H;
    memset(&H, 0, sizeof H);
    H.deleted_id = -1;
    H.flags = (ch->ds->od ? PROP_OD : 0u) | PROP_OD_ONLY;
    H.sums_old = H.sums_new = ch->sums[0];
    H.maxid_old = H.maxid_new = ch->maxid[0];
    H.od_g_row = kNoRow;
    std::vector<unsigned> slice_rows;
    int rc = add_od_request(ch, nu, H, slice_rows);
    if (rc) return rc;
    std::vector<unsigned char>& b = ch->blob;
    b.clear();
    b.resize(sizeof(DevHeader));
    H.off_old = H.off_unch = (unsigned)b.size();
    H.off_od_rows = (unsigned)bytes_append(b, slice_rows);
    memcpy(b.data(), &H, sizeof H);
    ch->alg_bytes = (double)ch->ds->M * 8.0 * ch->ds->K;
    ch->n_scores = 0.0;
    ch->cost = 1.0;
    ch->with_od = true;
    ch->entry = LaunchEntry{0u, 0u, 0u, 0u};
    return SCICONE_OK;
}

// Queue the compiled proposals of `n` chains (all on one dataset): K1 for the increments and slices they need, then ONE
// launch of the proposal kernel; returns a ticket.
int submit_proposals(scicone_dataset* ds, scicone_chain* const* chains, int n, unsigned long long* ticket) {
    const size_t head = sizeof(LaunchEntry) * n;
    size_t bytes = head, n_items = 0, aux_bytes = 0;
    for (int i = 0; i < n; ++i) {
        bytes += chains[i]->blob.size();
        n_items += chains[i]->items.size();
        aux_bytes += round_up(chains[i]->aux.size(), 16);
    }
    const size_t off_items = bytes;
    const size_t off_aux = off_items + n_items * sizeof(DevItem);
    bytes = off_aux + aux_bytes;
    int rc = ds->ensure_arena(bytes);
    if (rc) return rc;
    const unsigned long long tk = ds->next_ticket;
    scicone_dataset::Segment& g = ds->seg[tk % scicone_dataset::kRing];
    if (g.busy) return fail(SCICONE_ERR_STATE, "more than %d launches outstanding: collect a ticket first", scicone_dataset::kRing);
    // A lane's slots, arena and table arena belong to one launch at a time.  Cell-sharded datasets: launch k goes to lane
    // k mod n_lanes on every rank (the gather buffers are per lane), so launches must be collected in order; one GPU: any
    // free lane, so a short launch can be collected, and its lane reused, while a long one is still running.
    int lane = 0;
    if (ds->n_lanes > 1) {
        bool used[scicone_dataset::kLanes] = {false, false, false, false};
        for (scicone_dataset::Segment& o : ds->seg)
            if (o.busy) used[o.lane] = true;
        lane = (int)(tk % ds->n_lanes);
        if (ds->world == 1)
            for (int l = 0; l < ds->n_lanes && used[lane]; ++l) lane = l;
        if (used[lane]) return fail(SCICONE_ERR_STATE, "every launch lane has a launch outstanding: collect a ticket first");
    }
    cudaStream_t st = ds->lane_stream[lane];
    unsigned char* d_arena = ds->d_arena[lane];
    g.lane = lane;
    if (ds->world > 1 && n > SCICONE_MAX_SHARDED_BATCH)
        return fail(SCICONE_ERR_ARG, "at most %d proposals per launch on a cell-sharded dataset", SCICONE_MAX_SHARDED_BATCH);
    if (ds->xchg_on) ds->lane_seq[lane]++;
    g.with_od.assign((size_t)n, 0);
    LaunchEntry* offs = reinterpret_cast<LaunchEntry*>(g.h_arena);
    size_t at = head, item_at = off_items;
    // grid.y order = longest proposal first (CTAs are dispatched in block-id order): the heavy full-tree
    // rescorings of nu proposals start at once and the short ones fill in behind them.  Results are addressed by
    // the slot in each launch entry, so the order of the blobs is free.
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return chains[a]->cost > chains[b]->cost; });
    std::vector<unsigned> blob_at(n);
    for (int i = 0; i < n; ++i) {
        scicone_chain* ch = chains[i];
        blob_at[i] = (unsigned)at;
        memcpy(g.h_arena + at, ch->blob.data(), ch->blob.size());
        at += ch->blob.size();
    }
    // K1 items: the root-score slices (~6 look-ups per cell) first, the node increments behind them.  The value tables of
    // the launch are laid out here, per rank: which rows of counts have a table, and how long it is, depends on the shard.
    const int* lut = ds->lut.data();
    const int K = ds->K;
    const int s_range = lut[2 * K + 1];
    size_t tab_at = 0;
    // SCICONE_TABLES=0: every lgamma term is evaluated directly in K1b (2 per event and cell) and K1a does not run - measured
    // on configs[2] this is a wash for launches of small proposals (one kernel less, 8 x the lgamma work: 142 against 144 us
    // per MCMC iteration of 32 chains) and a loss for nu moves, so tables are the default; the switch stays as a test hook.
    static const int tables_mode = [] {
        const char* e = getenv("SCICONE_TABLES");
        return e ? atoi(e) : -1;
    }();
    std::vector<char> direct_chain(n, 0);
    bool any_direct = false;
    for (int i = 0; i < n; ++i) {
        size_t n_ev = 0;
        for (const DevItem& it : chains[i]->items)
            if (it.kind == ITEM_NODE_OD) n_ev += (size_t)it.n_ev;
        direct_chain[i] = tables_mode == 0;
        any_direct = any_direct || (direct_chain[i] && n_ev > 0);
    }
    size_t n_k1b_items = 0;  // the ITEM_OD_HIST items (pass 2) are K1a work only: they sit behind everything K1b looks at
    for (int pass = 0; pass < 3; ++pass) {
        size_t chain_aux = off_aux;
        for (int i = 0; i < n; ++i) {
            scicone_chain* ch = chains[i];
            if (pass == 0 && !ch->aux.empty()) memcpy(g.h_arena + chain_aux, ch->aux.data(), ch->aux.size());
            bool slices_laid_out = false, dots_laid_out = false;
            for (const DevItem& src : ch->items) {
                const bool slice = src.kind == ITEM_SLICE_OD || src.kind == ITEM_SLICE_MN;
                if ((slice ? 0 : src.kind == ITEM_OD_HIST ? 2 : 1) != pass) continue;
                if (pass < 2) n_k1b_items++;
                DevItem it = src;
                it.aux_off += (unsigned)chain_aux;
                it.g_tab = kNoRow;
                if (it.kind == ITEM_OD_HIST) {  // one number per item, consecutive per chain; the header learns where they start
                    if (!dots_laid_out) reinterpret_cast<DevHeader*>(g.h_arena + blob_at[i])->od_dot_tab = (unsigned)tab_at;
                    dots_laid_out = true;
                    it.g_tab = (unsigned)tab_at;
                    tab_at += 1;
                } else if (it.kind == ITEM_NODE_OD) {
                    DevEvent* ev = reinterpret_cast<DevEvent*>(g.h_arena + it.aux_off);
                    for (int e = 0; e < it.n_ev; ++e) {
                        const int range = direct_chain[i] ? 0 : lut[2 * ev[e].k + 1];
                        ev[e].tab = range > 0 ? (unsigned)tab_at : kNoRow;
                        tab_at += (size_t)range;
                    }
                    if (s_range > 0 && !direct_chain[i]) {
                        it.g_tab = (unsigned)tab_at;
                        tab_at += 2 * (size_t)s_range;
                    }
                } else if (it.kind == ITEM_G_ROW) {
                    if (s_range > 0) {
                        it.g_tab = (unsigned)tab_at;
                        tab_at += (size_t)s_range;
                    }
                } else if (it.kind == ITEM_SLICE_OD && !slices_laid_out) {  // one table per region, shared by the chain's slices
                    unsigned* taboff = reinterpret_cast<unsigned*>(g.h_arena + it.aux_off + 2 * (size_t)K * sizeof(double));
                    for (int k = 0; k < K; ++k) {
                        const int range = lut[2 * k + 1];
                        taboff[k] = range > 0 ? (unsigned)tab_at : kNoRow;
                        tab_at += (size_t)range;
                    }
                    slices_laid_out = true;
                }
                {   // bookkeeping for the bench (scicone_build_counters): bytes K1b moves per cell, lgamma evaluations of K1a / K1b
                    const double M = (double)ds->M;
                    if (it.kind == ITEM_NODE_OD) {
                        const DevEvent* ev = reinterpret_cast<const DevEvent*>(g.h_arena + it.aux_off);
                        for (int e = 0; e < it.n_ev; ++e) {
                            const int range = ev[e].tab != kNoRow ? lut[2 * ev[e].k + 1] : 0;
                            ds->k1_bytes += M * (range > 0 ? 2.0 : 8.0);
                            ds->k1_lgamma += 2.0 * (range > 0 ? range : M);
                        }
                        ds->k1_bytes += M * (8.0 + (it.g_tab != kNoRow ? 2.0 : 8.0));
                        ds->k1_lgamma += 2.0 * (it.g_tab != kNoRow ? s_range : M);
                    } else if (it.kind == ITEM_NODE_MN)
                        ds->k1_bytes += M * (16.0 + 8.0 * it.n_ev);
                    else if (it.kind == ITEM_G_ROW) {
                        ds->k1_bytes += M * (8.0 + (s_range > 0 ? 2.0 : 8.0));
                        ds->k1_lgamma += s_range > 0 ? s_range : M;
                    } else if (it.kind == ITEM_OD_HIST) {
                        for (int k = it.k0; k < it.k1; ++k) {
                            ds->k1_bytes += 8.0 * ds->glut[2 * (size_t)k + 1];
                            ds->k1_lgamma += ds->glut[2 * (size_t)k + 1];
                        }
                    } else {
                        ds->k1_bytes += M * 8.0;
                        for (int k = it.k0; k < it.k1; ++k) {
                            const int range = it.kind == ITEM_SLICE_OD ? lut[2 * k + 1] : 0;
                            ds->k1_bytes += M * (range > 0 ? 2.0 : 8.0);
                            if (it.kind == ITEM_SLICE_OD) ds->k1_lgamma += range > 0 ? range : M;  // one table per region and chain, or one evaluation per cell
                        }
                    }
                }
                memcpy(g.h_arena + item_at, &it, sizeof it);
                item_at += sizeof it;
            }
            chain_aux += round_up(ch->aux.size(), 16);
        }
    }
    if (tab_at >= 0xffffff00ull) return fail(SCICONE_ERR_ARG, "the value tables of one launch exceed 2^32 entries");
    rc = ds->ensure_tables(tab_at);
    if (rc) return rc;
    for (int q = 0; q < n; ++q) {
        offs[q] = chains[order[q]]->entry;
        offs[q].blob_off = blob_at[order[q]];
        offs[q].slot = (unsigned)order[q];
    }
    CUDA_TRY(cudaMemcpyAsync(d_arena, g.h_arena, bytes, cudaMemcpyHostToDevice, st));
    LaunchParams P;
    memset(&P, 0, sizeof P);
    P.arena = d_arena;
    P.rows = ds->pool.base_ptr();
    P.stride = (long long)ds->Mp;
    P.Dt = ds->dDt;
    P.It = ds->dIt;
    P.S = ds->dS;
    P.Si = ds->dSi;
    P.w = ds->dW;
    P.lut = ds->d_lut;
    P.tables = ds->d_tables[lane];
    P.n_cells = ds->M;
    P.n_regions = ds->K;
    P.partial = ds->lane_partial(lane);
    P.chunk_sums = ds->lane_chunk(lane);
    P.total = ds->lane_total(lane);
    P.counter = ds->d_counter + (size_t)lane * ds->slots;
    P.status = ds->lane_status(lane);
    P.n_cta = ds->n_cta;
    P.n_red_chunks = ds->n_chunks;
    P.cs_stride = ds->cs_stride;
    P.n_props = n;
    P.launch_counter = ds->d_launch_counter + lane;
    if (ds->xchg_on) {
        const unsigned long long seq = ds->lane_seq[lane];
        P.xchg_data = ds->d_xchg_data_tab + ((size_t)lane * 2 + (seq & 1)) * ds->world;
        P.xchg_flag = ds->d_xchg_flag_tab + (size_t)lane * ds->world;
        P.xchg_seq = seq;
        P.xchg_world = ds->world;
    }
    const bool direct_results = ds->world == 1 && ds->direct_results;
    if (direct_results) {
        P.h_total = g.h_total;
        P.h_status = g.h_status;
        P.h_flag = g.h_flag;
        P.ticket = tk;
    }
    P.off_items = (unsigned)off_items;
    P.n_items = (int)n_items;
    P.hist = ds->d_hist;
    P.glut = ds->d_glut;
    P.hist_off = ds->d_hist_off;
    P.od_owner = ds->rank == 0 ? 1 : 0;
    // K1a -> K1b -> proposal kernel (-> gather kernel): each of them may be scheduled while its predecessor drains and waits
    // (griddepcontrol.wait) before it reads what the predecessor wrote.  Only a kernel that directly follows another kernel
    // of this launch gets the attribute; with kernel events on, the kernels are serialised the ordinary way so that the
    // events bracket one kernel each.
    const bool pdl = ds->pdl && !ds->profiling;
    bool after_kernel = false;  // the previous operation on the stream is a kernel of this launch
    auto launch = [&](auto kernel, dim3 grid, unsigned block) -> cudaError_t {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid;
        cfg.blockDim = dim3(block);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl && after_kernel) ? 1 : 0;
        after_kernel = true;
        ds->n_launches++;
        return cudaLaunchKernelEx(&cfg, kernel, P);
    };
    if (n_items) {
        if (ds->profiling) CUDA_TRY(cudaEventRecord(g.evb, st));
        if (n_items > 65535) return fail(SCICONE_ERR_ARG, "more than 65535 rescored nodes in one launch");
        if (tab_at > 0) CUDA_TRY(launch(build_tables_kernel, dim3((unsigned)n_items), kTablesBlock));
        if (n_k1b_items) {
            const dim3 bgrid((ds->M + kBuildBlock * kBuildCells - 1) / (kBuildBlock * kBuildCells), (unsigned)n_k1b_items);
            if (ds->all_tab && !any_direct)
                CUDA_TRY(launch(build_increments_kernel<false>, bgrid, kBuildBlock));
            else
                CUDA_TRY(launch(build_increments_kernel<true>, bgrid, kBuildBlock));
        }
        CUDA_TRY(cudaGetLastError());
    }
    g.n_items = (int)n_items;
    if (ds->profiling) CUDA_TRY(cudaEventRecord(g.ev0, st));
    {   // the scoring mode is a property of the dataset: one specialisation per launch
        const dim3 grid(ds->n_cta, n);
        if (ds->maxs)
            CUDA_TRY(launch(score_proposals_kernel<true>, grid, kThreads));
        else
            CUDA_TRY(launch(score_proposals_kernel<false>, grid, kThreads));
    }
    if (ds->profiling) CUDA_TRY(cudaEventRecord(g.ev1, st));
    CUDA_TRY(cudaGetLastError());
    bool gathered_direct = false;
    if (ds->world > 1 && ds->xchg_on) {
        // The proposal kernel has stored this rank's chunk sums into every rank's gather buffer and published its flag; one
        // small kernel waits for the flags of all ranks, adds all chunks in global chunk order (so the total does not depend
        // on the number of GPUs) and writes the n totals, the status words and the ticket into this
        // launch's pinned host buffers: no copy, no event.
        const unsigned long long seq = ds->lane_seq[lane];
        {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(1);
            cfg.blockDim = dim3(128);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl ? 1 : 0;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, gather_totals_kernel, (const unsigned long long*)ds->xchg_flags(ds->xchg, lane, 0), (int)ds->world, seq,
                                        (const double*)ds->xchg_data(ds->xchg, lane, (int)(seq & 1), 0), (long long)ds->xchg_block(), n,
                                        (int)ds->max_chunks_rank, (const unsigned*)ds->lane_status(lane), g.h_total, g.h_status, g.h_flag,
                                        (unsigned long long)tk, 60ull * 1000000000ull));
        }
        CUDA_TRY(cudaGetLastError());
        ds->n_launches++;
        gathered_direct = true;
    } else if (ds->world > 1) {
        // the same exchange as an NCCL all-gather of the chunk sums (every rank's slots are zero-padded to the same width)
        double* d_gather = ds->d_gather + (size_t)lane * ds->gather_cap * ds->world;
        const size_t cnt = (size_t)n * 2 * ds->max_chunks_rank;
        int nrc = g_nccl.AllGather(ds->lane_chunk(lane), d_gather, cnt, kNcclDouble, lane == 0 ? ds->comm : ds->comm_lane[lane], st);
        if (nrc != 0) return fail(SCICONE_ERR_COMM, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "?");
        CUDA_TRY(cudaMemcpyAsync(g.h_gather, d_gather, sizeof(double) * cnt * ds->world, cudaMemcpyDeviceToHost, st));
    } else if (!direct_results)
        CUDA_TRY(cudaMemcpyAsync(g.h_total, ds->lane_total(lane), ds->lane_result_bytes, cudaMemcpyDeviceToHost, st));
    if (ds->world > 1 && !gathered_direct)
        CUDA_TRY(cudaMemcpyAsync(g.h_status, ds->lane_status(lane), sizeof(unsigned) * n, cudaMemcpyDeviceToHost, st));
    g.direct = direct_results || gathered_direct;
    g.summed = gathered_direct;
    if (!g.direct) CUDA_TRY(cudaEventRecord(g.done, st));
    ds->h2d_bytes += bytes;
    ds->d2h_bytes += (ds->world > 1 && !gathered_direct ? sizeof(double) * 2 * n * ds->max_chunks_rank * ds->world : sizeof(double) * 2 * n) + sizeof(unsigned) * n;
    for (int i = 0; i < n; ++i) {
        scicone_chain* ch = chains[i];
        ds->alg_bytes += ch->alg_bytes;
        ds->n_scores += ch->n_scores;
        g.with_od[i] = ch->with_od;
        // rows that only live for this launch go back to the pool when it is collected (with two lanes a later launch is
        // not ordered behind this one)
        g.temp_rows.insert(g.temp_rows.end(), ch->temp_rows.begin(), ch->temp_rows.end());
        ch->temp_rows.clear();
        ch->items.clear();
        ch->aux.clear();
    }
    g.busy = true;
    g.n = n;
    g.ticket = tk;
    ds->next_ticket++;
    *ticket = tk;
    return SCICONE_OK;
}

// Wait for a queued launch and fetch its totals.
int wait_proposals(scicone_dataset* ds, unsigned long long ticket, double* totals, double* od_scores) {
    scicone_dataset::Segment& g = ds->seg[ticket % scicone_dataset::kRing];
    if (!g.busy || g.ticket != ticket) return fail(SCICONE_ERR_STATE, "unknown or already collected ticket %llu", ticket);
    if (g.direct) {
        // the last CTA of the launch (or, cell-sharded, the gather kernel) publishes the ticket after the totals (system-scope
        // fences): spin on it, and look at the stream from time to time so that a failed launch is reported, not waited for
        volatile unsigned long long* flag = g.h_flag;
        unsigned spins = 0;
        while (*flag != ticket) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if ((++spins & 0xffffu) == 0) {
                cudaError_t e = cudaStreamQuery(ds->lane_stream[g.lane]);
                if (e != cudaSuccess && e != cudaErrorNotReady) return fail(SCICONE_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
                if (e == cudaSuccess && *flag != ticket) return fail(SCICONE_ERR_CUDA, "the launch finished without publishing its results");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else
        CUDA_TRY(cudaEventSynchronize(g.done));
    g.busy = false;
    ds->pool.put_many(g.temp_rows, 0);
    const int n = g.n;
    if (ds->profiling) {
        float ms = 0.f;
        if (g.direct) CUDA_TRY(cudaEventSynchronize(g.ev1));  // the results are published before the kernel has retired
        CUDA_TRY(cudaEventElapsedTime(&ms, g.ev0, g.ev1));
        ds->last_kernel_us = 1e3 * ms;
        ds->sum_kernel_us += ds->last_kernel_us;
        ds->n_timed++;
        if (g.n_items) {
            CUDA_TRY(cudaEventElapsedTime(&ms, g.evb, g.ev0));
            ds->sum_build_us += 1e3 * ms;
            ds->n_build_timed++;
        }
    }
    bool nan = false;
    for (int i = 0; i < n; ++i) {
        double v[2] = {0.0, 0.0};
        if (ds->world > 1 && !g.summed) {
            for (int q2 = 0; q2 < 2; ++q2)
                for (int rk = 0; rk < ds->world; ++rk) {
                    const double* c = g.h_gather + (((size_t)rk * n + i) * 2 + q2) * ds->max_chunks_rank;
                    for (int q = 0; q < ds->max_chunks_rank; ++q) v[q2] += c[q];
                }
        } else {
            v[0] = g.h_total[2 * i];
            v[1] = g.h_total[2 * i + 1];
        }
        if (totals) totals[i] = v[0];
        if (od_scores) od_scores[i] = g.with_od[i] ? v[1] : NAN;
        if (g.h_status[i] & 1u) nan = true;
    }
    if (ds->world > 1 && (g.h_status[0] & 2u)) {
        CUDA_TRY(cudaMemsetAsync(ds->lane_status(g.lane), 0, sizeof(unsigned) * ds->slots, ds->lane_stream[g.lane]));
        return fail(SCICONE_ERR_COMM, "a rank did not publish its chunk sums within 60 s (launches must be submitted in the same order on all ranks)");
    }
    if (nan) {
        CUDA_TRY(cudaMemsetAsync(ds->lane_status(g.lane), 0, sizeof(unsigned) * ds->slots, ds->lane_stream[g.lane]));
        return fail(SCICONE_ERR_NAN, "a per-cell aggregate is NaN (reference: assert(!isnan), Inference.h:1150,1191)");
    }
    return SCICONE_OK;
}

int launch_proposals(scicone_dataset* ds, scicone_chain* const* chains, int n, double* totals, double* ods) {
    unsigned long long tk = 0;
    int rc = submit_proposals(ds, chains, n, &tk);
    return rc ? rc : wait_proposals(ds, tk, totals, ods);
}

int od_scores(scicone_chain* ch, double nu, double* out) {
    scicone_dataset* ds = ch->ds;
    if (ch->follower) return fail(SCICONE_ERR_STATE, "a follower chain only stages imported proposals");
    if (ch->compiled && !ch->evaluated) return fail(SCICONE_ERR_STATE, "root scores must be requested before scicone_prepare_t_prime");
    int rc = ds->set_device();
    if (rc) return rc;
    rc = ds->ensure_slots(1);
    if (rc) return rc;
    // the queued proposal of this chain (if any) is compiled later and rebuilds its own blob
    std::vector<unsigned char> keep;
    keep.swap(ch->blob);
    const bool keep_with_od = ch->with_od;
    const LaunchEntry keep_entry = ch->entry;
    std::vector<unsigned> keep_temp;
    keep_temp.swap(ch->temp_rows);
    rc = compile_od_only(ch, nu);
    double od = 0.0;
    unsigned long long tk = 0;
    if (!rc) rc = submit_proposals(ds, &ch, 1, &tk);
    if (!rc) rc = wait_proposals(ds, tk, nullptr, &od);
    for (unsigned r : ch->temp_rows) ch->row_put(r);  // only after a failed compile / submit: a submitted launch took them over
    ch->temp_rows.swap(keep_temp);
    keep.swap(ch->blob);
    ch->with_od = keep_with_od;
    ch->entry = keep_entry;
    if (rc) return rc;
    if (out) *out = od;
    return SCICONE_OK;
}

int gather_rows(scicone_dataset* ds, const std::vector<unsigned>& rows, double* out) {
    for (int l = 1; l < scicone_dataset::kLanes; ++l)
        if (ds->lane_stream[l]) CUDA_TRY(cudaStreamSynchronize(ds->lane_stream[l]));
    const int n = (int)rows.size();
    if (n == 0 || ds->M == 0) return SCICONE_OK;
    unsigned* d_rows = nullptr;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_rows, sizeof(unsigned) * n));
    cudaError_t e = cudaMalloc(&d_out, sizeof(double) * (size_t)n * ds->M);
    if (e != cudaSuccess) {
        cudaFree(d_rows);
        return fail(SCICONE_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    }
    cudaMemcpyAsync(d_rows, rows.data(), sizeof(unsigned) * n, cudaMemcpyHostToDevice, ds->stream);
    dim3 grid((n + 31) / 32, (ds->M + 31) / 32), block(32, 8);
    gather_rows_kernel<<<grid, block, 0, ds->stream>>>(ds->pool.base_ptr(), (long long)ds->Mp, d_rows, n, ds->M, d_out);
    cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n * ds->M, cudaMemcpyDeviceToHost, ds->stream);
    e = cudaStreamSynchronize(ds->stream);
    cudaFree(d_rows);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(SCICONE_ERR_CUDA, "gather_rows: %s", cudaGetErrorString(e));
    return SCICONE_OK;
}

template <class T>
int read_row(scicone_dataset* ds, unsigned row, T* out, size_t n) {
    if (row == kNoRow) return fail(SCICONE_ERR_STATE, "the chain holds no such table on this rank");
    for (int l = 1; l < scicone_dataset::kLanes; ++l)
        if (ds->lane_stream[l]) CUDA_TRY(cudaStreamSynchronize(ds->lane_stream[l]));
    CUDA_TRY(cudaMemcpyAsync(out, ds->pool.base_ptr() + (size_t)row * ds->Mp, sizeof(T) * n, cudaMemcpyDeviceToHost, ds->stream));
    CUDA_TRY(cudaStreamSynchronize(ds->stream));
    return SCICONE_OK;
}

// Wire format of an exported proposal (scicone_export_t_prime): this head, then the blob, the K1 items and their aux
// block.  Everything in it is rank independent: rows are indices into the row pool, and the rows of a chain come from
// its owner's partition of the pool.
struct ExportHead {
    unsigned magic, blob_bytes, n_items, aux_bytes;
    unsigned with_od, part, rows_needed, full;
    double alg_bytes_per_cell, n_scores_per_cell, cost, pad;
    LaunchEntry entry;
};
constexpr unsigned kExportMagic = 0x53433242u;  // "SC2B"

// all-gather of `cnt` doubles per rank through the communicator (set-up plumbing only)
int gather_doubles(scicone_dataset* ds, const double* mine, int cnt, std::vector<double>& all) {
    double *d_one = nullptr, *d_all = nullptr;
    all.assign((size_t)cnt * ds->world, 0.0);
    CUDA_TRY(cudaMalloc(&d_one, sizeof(double) * cnt));
    CUDA_TRY(cudaMalloc(&d_all, sizeof(double) * cnt * ds->world));
    CUDA_TRY(cudaMemcpyAsync(d_one, mine, sizeof(double) * cnt, cudaMemcpyHostToDevice, ds->stream));
    int nrc = g_nccl.AllGather(d_one, d_all, cnt, kNcclDouble, ds->comm, ds->stream);
    if (nrc != 0) return fail(SCICONE_ERR_COMM, "ncclAllGather failed (%d)", nrc);
    CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, sizeof(double) * cnt * ds->world, cudaMemcpyDeviceToHost, ds->stream));
    CUDA_TRY(cudaStreamSynchronize(ds->stream));
    cudaFree(d_one);
    cudaFree(d_all);
    return SCICONE_OK;
}


// Count histograms for the root score (scicone_dataset::od_hist).  One GPU: from the shard's own count ranges, at
// scicone_dataset_create.  Cell-sharded: again at scicone_dataset_attach_comm - the ranks agree on the global range of every
// region, histogram their shards over it and add the histograms up (integers, exact), so every rank ends up with the
// same table.  Not used (od_hist stays false, the per-cell slices do the work) when a region holds non-integer counts or
// counts spread over more than kLutMax values, or with the multinomial likelihood.
int build_od_hist(scicone_dataset* ds) {
    const int K = ds->K;
    ds->od_hist = false;
    cudaFree(ds->d_glut);
    cudaFree(ds->d_hist_off);
    cudaFree(ds->d_hist);
    ds->d_glut = nullptr;
    ds->d_hist_off = nullptr;
    ds->d_hist = nullptr;
    const bool have = ds->raw_int.size() == (size_t)K + 1;
    std::vector<double> st(3 * (size_t)K + 1, 0.0);
    for (int k = 0; k < K && have; ++k) {
        st[k] = ds->raw_lo[k];
        st[(size_t)K + k] = ds->raw_hi[k];
        st[2 * (size_t)K + k] = ds->raw_int[k] ? 1.0 : 0.0;
    }
    st[3 * (size_t)K] = (have && ds->od && ds->od_hist_allowed) ? 1.0 : 0.0;
    if (ds->world > 1) {
        std::vector<double> all;
        int rc = gather_doubles(ds, st.data(), (int)st.size(), all);
        if (rc) return rc;
        for (int p = 0; p < ds->world; ++p) {
            const double* o = all.data() + (size_t)p * st.size();
            for (int k = 0; k < K; ++k) {
                st[k] = std::min(st[k], o[k]);
                st[(size_t)K + k] = std::max(st[(size_t)K + k], o[(size_t)K + k]);
                st[2 * (size_t)K + k] = std::min(st[2 * (size_t)K + k], o[2 * (size_t)K + k]);
            }
            st[3 * (size_t)K] = std::min(st[3 * (size_t)K], o[3 * (size_t)K]);
        }
    }
    bool ok = st[3 * (size_t)K] == 1.0;
    for (int k = 0; k < K && ok; ++k) ok = st[2 * (size_t)K + k] == 1.0 && st[(size_t)K + k] - st[k] + 1.0 <= (double)kLutMax;
    if (!ok) return SCICONE_OK;
    ds->glut.assign((size_t)K * 2, 0);
    ds->hist_off.assign((size_t)K + 1, 0u);
    for (int k = 0; k < K; ++k) {
        ds->glut[2 * (size_t)k] = (int)st[k];
        ds->glut[2 * (size_t)k + 1] = (int)(st[(size_t)K + k] - st[k] + 1.0);
        ds->hist_off[(size_t)k + 1] = ds->hist_off[k] + (unsigned)ds->glut[2 * (size_t)k + 1];
    }
    const size_t n = ds->hist_off[K];
    CUDA_TRY(cudaMalloc(&ds->d_glut, sizeof(int) * ds->glut.size()));
    CUDA_TRY(cudaMalloc(&ds->d_hist_off, sizeof(unsigned) * ds->hist_off.size()));
    CUDA_TRY(cudaMalloc(&ds->d_hist, sizeof(double) * n));
    CUDA_TRY(cudaMemcpyAsync(ds->d_glut, ds->glut.data(), sizeof(int) * ds->glut.size(), cudaMemcpyHostToDevice, ds->stream));
    CUDA_TRY(cudaMemcpyAsync(ds->d_hist_off, ds->hist_off.data(), sizeof(unsigned) * ds->hist_off.size(), cudaMemcpyHostToDevice, ds->stream));
    CUDA_TRY(cudaMemsetAsync(ds->d_hist, 0, sizeof(double) * n, ds->stream));
    hist_kernel<<<dim3((ds->M + 255) / 256, K), 256, 0, ds->stream>>>(ds->dDt, ds->dW, ds->d_glut, ds->d_hist_off, ds->M, (long long)ds->Mp, ds->d_hist);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ds->stream));
    if (ds->world > 1) {
        std::vector<double> mine(n), all;
        CUDA_TRY(cudaMemcpy(mine.data(), ds->d_hist, sizeof(double) * n, cudaMemcpyDeviceToHost));
        int rc = gather_doubles(ds, mine.data(), (int)n, all);
        if (rc) return rc;
        std::fill(mine.begin(), mine.end(), 0.0);
        for (int p = 0; p < ds->world; ++p)
            for (size_t i = 0; i < n; ++i) mine[i] += all[(size_t)p * n + i];
        CUDA_TRY(cudaMemcpy(ds->d_hist, mine.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    ds->od_hist = true;
    return SCICONE_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int scicone_abi_version(void) { return SCICONE_ABI_VERSION; }
const char* scicone_last_error(void) { return g_err.c_str(); }

int scicone_dataset_create(scicone_dataset** out, const scicone_dataset_desc* d) {
    if (!out || !d) return fail(SCICONE_ERR_ARG, "dataset_create: null argument");
    *out = nullptr;
    if (d->n_cells <= 0 || d->n_regions <= 0 || !d->D || !d->region_sizes || !d->neutral_states || !d->cluster_sizes)
        return fail(SCICONE_ERR_ARG, "dataset_create: empty matrix or null buffer (n_cells=%d, n_regions=%d)", d->n_cells, d->n_regions);
    if (d->n_cells > 400000000) return fail(SCICONE_ERR_ARG, "dataset_create: at most 400 000 000 cells per dataset (the kernels address a row with a 32-bit byte pitch)");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(SCICONE_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (d->device < 0 || d->device >= n_dev) return fail(SCICONE_ERR_ARG, "device %d out of range (%d devices)", d->device, n_dev);
    scicone_dataset* ds = new scicone_dataset();
    ds->M = d->n_cells;
    ds->K = d->n_regions;
    ds->Mp = round_up(ds->M, kPad);
    ds->device = d->device;
    ds->eta = d->eta;
    ds->od = d->is_overdispersed != 0;
    ds->maxs = d->max_scoring != 0;
    ds->cell_offset = d->cell_offset;
    ds->M_global = d->n_cells_global > 0 ? d->n_cells_global : d->n_cells;
    ds->r.assign(d->region_sizes, d->region_sizes + ds->K);
    ds->neutral.assign(d->neutral_states, d->neutral_states + ds->K);
    for (int k = 0; k < ds->K; ++k) {
        ds->root_z += ds->r[k] * ds->neutral[k];
        ds->sum_r += ds->r[k];
    }
    ds->n_cta = (ds->M + kCells - 1) / kCells;
    ds->n_chunks = (ds->n_cta + kCtasPerChunk - 1) / kCtasPerChunk;
    ds->cs_stride = ds->n_chunks;
    auto cleanup = [&](int rc) {
        scicone_dataset_destroy(ds);
        return rc;
    };
#define DS_TRY(expr)                                                                                         \
    do {                                                                                                     \
        cudaError_t e_ = (expr);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return cleanup(fail(SCICONE_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    DS_TRY(cudaSetDevice(ds->device));
    DS_TRY(cudaFree(0));  // the primary context exists before the driver API is used
    DS_TRY(cudaDeviceGetAttribute(&ds->n_sm, cudaDevAttrMultiProcessorCount, ds->device));
    {
        int rc = ds->pool.init(ds->Mp, ds->device, 1, 0);
        if (rc) return cleanup(rc);
    }
    DS_TRY(cudaStreamCreateWithFlags(&ds->stream, cudaStreamNonBlocking));
    ds->lane_stream[0] = ds->stream;
    for (scicone_dataset::Segment& g : ds->seg) {
        DS_TRY(cudaEventCreate(&g.evb));
        DS_TRY(cudaEventCreate(&g.ev0));
        DS_TRY(cudaEventCreate(&g.ev1));
        DS_TRY(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
    }
    DS_TRY(cudaMalloc(&ds->dDt, sizeof(double) * ds->Mp * ds->K));
    DS_TRY(cudaMalloc(&ds->dIt, sizeof(unsigned short) * ds->Mp * ds->K));
    DS_TRY(cudaMalloc(&ds->dS, sizeof(double) * ds->Mp));
    DS_TRY(cudaMalloc(&ds->dSi, sizeof(unsigned short) * ds->Mp));
    DS_TRY(cudaMalloc(&ds->dW, sizeof(int) * ds->Mp));
    DS_TRY(cudaMalloc(&ds->d_launch_counter, sizeof(unsigned) * scicone_dataset::kLanes));
    DS_TRY(cudaMemsetAsync(ds->d_launch_counter, 0, sizeof(unsigned) * scicone_dataset::kLanes, ds->stream));
    DS_TRY(cudaMemsetAsync(ds->dDt, 0, sizeof(double) * ds->Mp * ds->K, ds->stream));
    DS_TRY(cudaMemsetAsync(ds->dIt, 0, sizeof(unsigned short) * ds->Mp * ds->K, ds->stream));
    DS_TRY(cudaMemsetAsync(ds->dW, 0, sizeof(int) * ds->Mp, ds->stream));
    DS_TRY(cudaMemsetAsync(ds->dS, 0, sizeof(double) * ds->Mp, ds->stream));
    DS_TRY(cudaMemsetAsync(ds->dSi, 0, sizeof(unsigned short) * ds->Mp, ds->stream));
    {
        double* dD = nullptr;
        DS_TRY(cudaMalloc(&dD, sizeof(double) * (size_t)ds->M * ds->K));
        cudaError_t e1 = cudaMemcpyAsync(dD, d->D, sizeof(double) * (size_t)ds->M * ds->K, cudaMemcpyHostToDevice, ds->stream);
        dim3 grid((ds->K + 31) / 32, (ds->M + 31) / 32), block(32, 8);
        transpose_kernel<<<grid, block, 0, ds->stream>>>(dD, ds->dDt, ds->M, ds->K, (long long)ds->Mp);
        row_sum_kernel<<<(ds->M + 127) / 128, 128, 0, ds->stream>>>(ds->dDt, ds->dS, ds->M, ds->K, (long long)ds->Mp);
        // Count-range table: read counts are integers, and the counts of one region over the cells of a shard take a few
        // hundred distinct values - K1 then evaluates lgamma(count + c) once per possible count instead of once per cell
        // (identical results: the same function of the same argument).  A row of counts is tabulated when all its values are
        // non-negative integers and the table is at most half as long as the row; SCICONE_LUT=0 turns it off.
        const int K = ds->K, M = ds->M;
        std::vector<int>& lut = ds->lut;
        lut.assign((size_t)(K + 1) * 2, 0);
        const char* lut_env = getenv("SCICONE_LUT");
        if (!(lut_env && atoi(lut_env) == 0)) {
            std::vector<double> lo((size_t)K + 1, 1e300), hi((size_t)K + 1, -1e300);
            std::vector<char> integral((size_t)K + 1, 1);
            for (int j = 0; j < M; ++j) {
                const double* row = d->D + (size_t)j * K;
                double sum = 0.0;
                bool all_int = true;
                for (int k = 0; k < K; ++k) {
                    const double v = row[k];
                    sum = sum + v;  // the order of row_sum_kernel; exact for integer counts anyway
                    if (!(v >= 0.0 && v < 1073741824.0 && v == std::floor(v))) {
                        integral[k] = 0;
                        all_int = false;
                    }
                    lo[k] = std::min(lo[k], v);
                    hi[k] = std::max(hi[k], v);
                }
                if (!all_int || !(sum < 1073741824.0)) integral[K] = 0;
                lo[K] = std::min(lo[K], sum);
                hi[K] = std::max(hi[K], sum);
            }
            ds->raw_lo = lo;
            ds->raw_hi = hi;
            ds->raw_int = integral;
            ds->all_tab = true;
            for (int k = 0; k <= K; ++k) {
                const double range = hi[k] - lo[k] + 1.0;
                if (!integral[k] || range > (double)kLutMax || 2.0 * range > (double)std::max(M, 2)) {
                    ds->all_tab = false;
                    continue;
                }
                lut[(size_t)k * 2] = (int)lo[k];
                lut[(size_t)k * 2 + 1] = (int)range;
            }
        }
        DS_TRY(cudaMalloc(&ds->d_lut, sizeof(int) * lut.size()));
        DS_TRY(cudaMemcpyAsync(ds->d_lut, lut.data(), sizeof(int) * lut.size(), cudaMemcpyHostToDevice, ds->stream));
        index_kernel<<<dim3((M + 255) / 256, K + 1), 256, 0, ds->stream>>>(ds->dDt, ds->dS, ds->d_lut, M, K, (long long)ds->Mp, ds->dIt, ds->dSi);
        cudaError_t e2 = cudaMemcpyAsync(ds->dW, d->cluster_sizes, sizeof(int) * ds->M, cudaMemcpyHostToDevice, ds->stream);
        cudaError_t e3 = cudaStreamSynchronize(ds->stream);  // `lut` and the caller's buffers are read until here
        cudaFree(dD);
        DS_TRY(e1);
        DS_TRY(e2);
        DS_TRY(e3);
        DS_TRY(cudaGetLastError());
    }
#undef DS_TRY
    if (const char* e = getenv("SCICONE_DIRECT_RESULTS")) ds->direct_results = atoi(e) != 0;
    if (const char* e = getenv("SCICONE_OD_HIST")) ds->od_hist_allowed = atoi(e) != 0;
    if (const char* e = getenv("SCICONE_PDL")) ds->pdl = atoi(e) != 0;
    int rc = build_od_hist(ds);
    if (rc) return cleanup(rc);
    rc = ds->ensure_slots(64);
    if (rc) return cleanup(rc);
    rc = ds->ensure_arena(size_t(4) << 20);
    if (rc) return cleanup(rc);
    rc = ds->ensure_tables(size_t(1) << 20);
    if (rc) return cleanup(rc);
    *out = ds;
    return SCICONE_OK;
}

void scicone_dataset_destroy(scicone_dataset* ds) {
    if (!ds) return;
    cudaSetDevice(ds->device);
    if (ds->stream) cudaStreamSynchronize(ds->stream);
    for (int l = 1; l < scicone_dataset::kLanes; ++l)
        if (ds->lane_stream[l]) cudaStreamSynchronize(ds->lane_stream[l]);
    for (int p = 0; p < (int)ds->xchg_peer.size(); ++p)
        if (p != ds->rank && ds->xchg_peer[p]) cudaIpcCloseMemHandle(ds->xchg_peer[p]);
    // the exported buffer itself is NOT freed here: another rank may still have it mapped, and freeing exported memory
    // before every importer has closed it is undefined; it goes away with the process (a sharded dataset lives as long
    // as its host process does)
    cudaFree(ds->d_xchg_data_tab); cudaFree(ds->d_xchg_flag_tab);
    for (void* c : ds->comm_lane)
        if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    if (ds->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ds->comm);
    ds->pool.destroy();
    cudaFree(ds->dDt); cudaFree(ds->dS); cudaFree(ds->dW); cudaFree(ds->d_lut); cudaFree(ds->dIt); cudaFree(ds->dSi);
    cudaFree(ds->d_glut); cudaFree(ds->d_hist_off); cudaFree(ds->d_hist);
    for (unsigned char* a : ds->d_arena) cudaFree(a);
    cudaFree(ds->d_partial); cudaFree(ds->d_chunk); cudaFree(ds->d_total);
    cudaFree(ds->d_counter); cudaFree(ds->d_gather); cudaFree(ds->d_launch_counter);
    for (double* t : ds->d_tables) cudaFree(t);
    for (scicone_dataset::Segment& g : ds->seg) {
        if (g.h_arena) cudaFreeHost(g.h_arena);
        if (g.h_total) cudaFreeHost(g.h_total);
        if (g.h_gather) cudaFreeHost(g.h_gather);
        if (g.evb) cudaEventDestroy(g.evb);
        if (g.ev0) cudaEventDestroy(g.ev0);
        if (g.ev1) cudaEventDestroy(g.ev1);
        if (g.done) cudaEventDestroy(g.done);
    }
    for (cudaEvent_t e : ds->region_ev)
        if (e) cudaEventDestroy(e);
    for (int l = 1; l < scicone_dataset::kLanes; ++l)
        if (ds->lane_stream[l]) cudaStreamDestroy(ds->lane_stream[l]);
    if (ds->stream) cudaStreamDestroy(ds->stream);
    delete ds;
}

int scicone_dataset_reserve(scicone_dataset* ds, int32_t max_batch, int64_t rows) {
    if (!ds || max_batch < 0 || rows < 0) return fail(SCICONE_ERR_ARG, "dataset_reserve: bad argument");
    int rc = ds->set_device();
    if (rc) return rc;
    if (max_batch > 0) {
        if (ds->world > 1) max_batch = std::min<int32_t>(max_batch, SCICONE_MAX_SHARDED_BATCH);
        rc = ds->ensure_slots(max_batch);
        if (rc) return rc;
    }
    for (int p = 0; p < ds->pool.n_parts && rows > 0; ++p) {
        rc = ds->pool.ensure_mapped(p, std::min<size_t>((size_t)rows, ds->pool.part_cap));
        if (rc) return rc;
    }
    return SCICONE_OK;
}

int scicone_max_batch(scicone_dataset* ds, int32_t* n) {
    if (!ds || !n) return fail(SCICONE_ERR_ARG, "null argument");
    *n = ds->slots;
    return SCICONE_OK;
}

static int chain_create(scicone_chain** out, scicone_dataset* ds, bool follower) {
    if (!out || !ds) return fail(SCICONE_ERR_ARG, "chain_create: null argument");
    *out = nullptr;
    int rc = ds->set_device();
    if (rc) return rc;
    scicone_chain* ch = new scicone_chain();
    ch->ds = ds;
    ch->follower = follower;
    if (!follower)
        for (int i = 0; i < 2 && !rc; ++i) {
            rc = ch->row_get(&ch->sums[i]);
            if (!rc) rc = ch->row_get(&ch->maxid[i]);
        }
    if (rc) {
        scicone_chain_destroy(ch);
        return rc;
    }
    *out = ch;
    return SCICONE_OK;
}
int scicone_chain_create(scicone_chain** out, scicone_dataset* ds) { return chain_create(out, ds, false); }
int scicone_chain_create_follower(scicone_chain** out, scicone_dataset* ds) { return chain_create(out, ds, true); }

void scicone_chain_destroy(scicone_chain* ch) {
    if (!ch) return;
    scicone_dataset* ds = ch->ds;
    cudaSetDevice(ds->device);
    cudaStreamSynchronize(ds->stream);
    for (int l = 1; l < scicone_dataset::kLanes; ++l)
        if (ds->lane_stream[l]) cudaStreamSynchronize(ds->lane_stream[l]);
    if (!ch->follower) {
        drop_pending(ch);
        for (auto& kv : ch->t) ch->row_put(kv.second.row);
        for (int i = 0; i < 2; ++i) {
            ch->row_put(ch->sums[i]);
            ch->row_put(ch->maxid[i]);
        }
        ds->pool.put_many(ch->stash, 0);
    }
    delete ch;
}

int scicone_prepare_t_table(scicone_chain* ch, const scicone_tree* t, int32_t with_od) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    if (ch->follower) return fail(SCICONE_ERR_STATE, "a follower chain only stages imported proposals");
    scicone_dataset* ds = ch->ds;
    int rc = validate_tree(t, ds->K);
    if (rc) return rc;
    rc = ds->set_device();
    if (rc) return rc;
    drop_pending(ch);
    for (auto& kv : ch->t) ch->row_put(kv.second.row);
    ch->t.clear();
    ch->tp.assign(t);
    ch->roots.assign(1, 0);
    ch->pending = true;
    ch->p_full = true;
    rc = compile_proposal(ch, true, with_od != 0);
    if (rc) {
        drop_pending(ch);
        return rc;
    }
    ch->compiled = true;
    return SCICONE_OK;
}

int scicone_compute_t_table(scicone_chain* ch, const scicone_tree* t, double* t_sum) {
    int rc = scicone_prepare_t_table(ch, t, 0);
    if (rc) return rc;
    scicone_dataset* ds = ch->ds;
    rc = ds->ensure_slots(1);
    double total = 0.0;
    if (!rc) rc = launch_proposals(ds, &ch, 1, &total, nullptr);
    if (rc && rc != SCICONE_ERR_NAN) {
        drop_pending(ch);
        return rc;
    }
    ch->evaluated = true;
    int rc2 = scicone_accept(ch);
    if (t_sum) *t_sum = total;
    return rc ? rc : rc2;
}

int scicone_compute_t_od_scores(scicone_chain* ch, double nu, double* od_score) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    return od_scores(ch, nu, od_score);
}
int scicone_compute_t_prime_od_scores(scicone_chain* ch, double nu, double* od_score) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    int rc = od_scores(ch, nu, od_score);
    if (!rc) {
        ch->od_p_valid = true;
        ch->od_p_nu = nu;
    }
    return rc;
}

int scicone_compute_t_prime_scores(scicone_chain* ch, const scicone_tree* t_prime, int32_t node_idx) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    if (ch->follower) return fail(SCICONE_ERR_STATE, "a follower chain only stages imported proposals");
    if (ch->t.empty()) return fail(SCICONE_ERR_STATE, "compute_t_prime_scores before compute_t_table");
    if (ch->evaluated) return fail(SCICONE_ERR_STATE, "the previous proposal was neither accepted nor rejected");
    if (ch->compiled) return fail(SCICONE_ERR_STATE, "compute_t_prime_scores after scicone_prepare_t_prime");
    if (!ch->pending) {
        int rc = validate_tree(t_prime, ch->ds->K);
        if (rc) return rc;
        ch->tp.assign(t_prime);
        ch->pending = true;
        ch->p_full = false;
    }
    if (node_idx < 0 || node_idx >= ch->tp.n) return fail(SCICONE_ERR_ARG, "node index %d outside the tree", node_idx);
    ch->roots.push_back(node_idx);
    return SCICONE_OK;
}

int scicone_prepare_t_prime(scicone_chain* ch) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    if (!ch->pending || ch->roots.empty()) return fail(SCICONE_ERR_STATE, "prepare_t_prime without compute_t_prime_scores");
    if (ch->evaluated) return fail(SCICONE_ERR_STATE, "proposal already evaluated");
    if (ch->compiled) return SCICONE_OK;
    int rc = compile_proposal(ch, false, false);
    if (rc) {
        drop_pending(ch);
        return rc;
    }
    ch->compiled = true;
    return SCICONE_OK;
}

int scicone_export_t_prime(scicone_chain* ch, void* buf, uint64_t capacity, uint64_t* bytes) {
    if (!ch || !bytes) return fail(SCICONE_ERR_ARG, "export_t_prime: null argument");
    if (ch->follower) return fail(SCICONE_ERR_STATE, "export_t_prime on a follower chain");
    if (!ch->compiled || ch->evaluated) return fail(SCICONE_ERR_STATE, "export_t_prime needs a compiled proposal that has not been submitted");
    const size_t need = sizeof(ExportHead) + ch->blob.size() + ch->items.size() * sizeof(DevItem) + ch->aux.size();
    *bytes = need;
    if (!buf || capacity < need) return fail(SCICONE_ERR_ARG, "export_t_prime: the buffer holds %llu bytes, the proposal needs %zu", (unsigned long long)capacity, need);
    scicone_dataset* ds = ch->ds;
    ExportHead h;
    memset(&h, 0, sizeof h);
    h.magic = kExportMagic;
    h.blob_bytes = (unsigned)ch->blob.size();
    h.n_items = (unsigned)ch->items.size();
    h.aux_bytes = (unsigned)ch->aux.size();
    h.with_od = ch->with_od ? 1u : 0u;
    h.part = (unsigned)ds->pool.my_part;
    h.rows_needed = (unsigned)ds->pool.next_row;
    h.full = ch->p_full ? 1u : 0u;
    h.alg_bytes_per_cell = ch->alg_bytes / (double)ds->M;  // the importer scales by its own shard
    h.n_scores_per_cell = ch->n_scores / (double)ds->M;
    h.cost = ch->cost;
    h.entry = ch->entry;
    unsigned char* o = static_cast<unsigned char*>(buf);
    memcpy(o, &h, sizeof h);
    o += sizeof h;
    memcpy(o, ch->blob.data(), ch->blob.size());
    o += ch->blob.size();
    if (!ch->items.empty()) memcpy(o, ch->items.data(), ch->items.size() * sizeof(DevItem));
    o += ch->items.size() * sizeof(DevItem);
    if (!ch->aux.empty()) memcpy(o, ch->aux.data(), ch->aux.size());
    return SCICONE_OK;
}

int scicone_import_t_prime(scicone_chain* ch, const void* buf, uint64_t bytes) {
    if (!ch || !buf) return fail(SCICONE_ERR_ARG, "import_t_prime: null argument");
    if (!ch->follower) return fail(SCICONE_ERR_STATE, "import_t_prime needs a follower chain (scicone_chain_create_follower)");
    if (ch->compiled && !ch->evaluated) return fail(SCICONE_ERR_STATE, "import_t_prime: the previous proposal was not submitted");
    if (bytes < sizeof(ExportHead)) return fail(SCICONE_ERR_ARG, "import_t_prime: short message");
    ExportHead h;
    memcpy(&h, buf, sizeof h);
    if (h.magic != kExportMagic || bytes < sizeof h + (size_t)h.blob_bytes + (size_t)h.n_items * sizeof(DevItem) + h.aux_bytes)
        return fail(SCICONE_ERR_ARG, "import_t_prime: malformed message");
    scicone_dataset* ds = ch->ds;
    int rc = ds->pool.ensure_mapped((int)h.part, h.rows_needed);  // the owner may have grown its partition
    if (rc) return rc;
    const unsigned char* in = static_cast<const unsigned char*>(buf) + sizeof h;
    ch->blob.assign(in, in + h.blob_bytes);
    in += h.blob_bytes;
    ch->items.resize(h.n_items);
    if (h.n_items) memcpy(ch->items.data(), in, (size_t)h.n_items * sizeof(DevItem));
    in += (size_t)h.n_items * sizeof(DevItem);
    ch->aux.assign(in, in + h.aux_bytes);
    ch->with_od = h.with_od != 0;
    ch->p_full = h.full != 0;
    ch->alg_bytes = h.alg_bytes_per_cell * ds->M;
    ch->n_scores = h.n_scores_per_cell * ds->M;
    ch->cost = h.cost;
    ch->entry = h.entry;
    ch->roots.assign(1, 0);
    ch->pending = true;
    ch->compiled = true;
    ch->evaluated = false;
    return SCICONE_OK;
}

int scicone_batch_submit(scicone_chain* const* chains, const scicone_tree* const* t_primes, int32_t n, uint64_t* ticket) {
    if (!chains || n <= 0 || !ticket) return fail(SCICONE_ERR_ARG, "batch: null or empty");
    scicone_dataset* ds = chains[0]->ds;
    for (int i = 0; i < n; ++i) {
        if (!chains[i] || chains[i]->ds != ds) return fail(SCICONE_ERR_ARG, "batch: all chains must share one dataset");
        if (!chains[i]->pending || chains[i]->roots.empty())
            return fail(SCICONE_ERR_STATE, "compute_t_prime_sums without compute_t_prime_scores (chain %d)", i);
        if (chains[i]->evaluated) return fail(SCICONE_ERR_STATE, "proposal of chain %d already evaluated", i);
        if (t_primes && t_primes[i] && !chains[i]->follower && t_primes[i]->n_nodes != chains[i]->tp.n)
            return fail(SCICONE_ERR_ARG, "batch: t_prime of chain %d differs from the tree given to compute_t_prime_scores", i);
    }
    int rc = ds->set_device();
    if (rc) return rc;
    rc = ds->ensure_slots(n);
    if (rc) return rc;
    {   // blobs not yet built by scicone_prepare_t_prime are independent of each other: compile them on the pool
        struct Job {
            scicone_chain* const* chains;
            std::vector<int> todo, rc;
            std::vector<std::string> msg;
        } job{chains, {}, std::vector<int>(n, 0), std::vector<std::string>(n)};
        for (int i = 0; i < n; ++i)
            if (!chains[i]->compiled) job.todo.push_back(i);
        scicone::ForkJoinPool::instance().run((int)job.todo.size(), [](int q, void* ctx) {
            Job& j = *static_cast<Job*>(ctx);
            const int i = j.todo[q];
            j.rc[i] = compile_proposal(j.chains[i], false, false);
            if (j.rc[i]) j.msg[i] = g_err;  // thread-local in the worker
        }, &job);
        for (int i = 0; i < n; ++i)
            if (job.rc[i]) {
                for (int q = 0; q < n; ++q) drop_pending(chains[q]);
                g_err = job.msg[i];
                return job.rc[i];
            }
        for (int i = 0; i < n; ++i) chains[i]->compiled = true;
    }
    unsigned long long tk = 0;
    rc = submit_proposals(ds, chains, n, &tk);
    if (rc) {
        for (int q = 0; q < n; ++q) drop_pending(chains[q]);
        return rc;
    }
    for (int i = 0; i < n; ++i) chains[i]->evaluated = true;
    *ticket = tk;
    return SCICONE_OK;
}

int scicone_batch_wait(scicone_dataset* ds, uint64_t ticket, double* t_prime_sums, double* od_scores) {
    if (!ds || !t_prime_sums) return fail(SCICONE_ERR_ARG, "batch_wait: null argument");
    int rc = ds->set_device();
    return rc ? rc : wait_proposals(ds, ticket, t_prime_sums, od_scores);
}

int scicone_batch_poll(scicone_dataset* ds, uint64_t ticket, int32_t* done) {
    if (!ds || !done) return fail(SCICONE_ERR_ARG, "batch_poll: null argument");
    scicone_dataset::Segment& g = ds->seg[ticket % scicone_dataset::kRing];
    if (!g.busy || g.ticket != ticket) return fail(SCICONE_ERR_STATE, "unknown or already collected ticket %llu", (unsigned long long)ticket);
    if (g.direct) {
        *done = *static_cast<volatile unsigned long long*>(g.h_flag) == ticket ? 1 : 0;
        return SCICONE_OK;
    }
    cudaError_t e = cudaEventQuery(g.done);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(SCICONE_ERR_CUDA, "batch_poll: %s", cudaGetErrorString(e));
    *done = e == cudaSuccess ? 1 : 0;
    return SCICONE_OK;
}

int scicone_batch_compute_t_prime_sums(scicone_chain* const* chains, const scicone_tree* const* t_primes, int32_t n,
                                       double* t_prime_sums, double* od_scores) {
    if (!t_prime_sums) return fail(SCICONE_ERR_ARG, "batch: null output");
    uint64_t tk = 0;
    int rc = scicone_batch_submit(chains, t_primes, n, &tk);
    return rc ? rc : scicone_batch_wait(chains[0]->ds, tk, t_prime_sums, od_scores);
}

int scicone_compute_t_prime_sums(scicone_chain* ch, const scicone_tree* t_prime, double* t_prime_sum) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    double total = 0.0;
    int rc = scicone_batch_compute_t_prime_sums(&ch, &t_prime, 1, &total, nullptr);
    if (t_prime_sum) *t_prime_sum = total;
    return rc;
}

int scicone_accept(scicone_chain* ch) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    if (!ch->pending || !ch->evaluated) return fail(SCICONE_ERR_STATE, "accept without an evaluated proposal");
    if (ch->follower) {  // the owner's index swaps are all there is: the next imported proposal names the new rows
        ch->pending = ch->compiled = ch->evaluated = false;
        return SCICONE_OK;
    }
    // Inference::update_t_scores, Inference.h:996-1012 - as index swaps
    for (auto& kv : ch->p) {
        auto it = ch->t.find(kv.first);
        if (it != ch->t.end()) {
            ch->row_put(it->second.row);
            it->second = kv.second;
        } else
            ch->t.emplace(kv.first, kv.second);
    }
    ch->p.clear();
    if (ch->deleted_id != -1) {
        auto it = ch->t.find(ch->deleted_id);
        if (it != ch->t.end()) {
            ch->row_put(it->second.row);
            ch->t.erase(it);
        }
    }
    ch->t_nu = ch->tp.nu;
    std::swap(ch->sums[0], ch->sums[1]);    // t_sums = t_prime_sums, Inference.h:867
    std::swap(ch->maxid[0], ch->maxid[1]);  // t_maxs = t_prime_maxs, Inference.h:868
    drop_pending(ch);
    return SCICONE_OK;
}

int scicone_reject(scicone_chain* ch) {
    if (!ch) return fail(SCICONE_ERR_ARG, "null chain");
    if (ch->follower) {
        ch->pending = ch->compiled = ch->evaluated = false;
        return SCICONE_OK;
    }
    drop_pending(ch);
    return SCICONE_OK;
}

namespace {
struct PerChainJob {
    scicone_chain* const* chains;
    const scicone_tree* const* trees;
    const int32_t* arg;
    std::vector<int> rc;
    std::vector<std::string> msg;
};
int finish_job(PerChainJob& job) {
    for (size_t i = 0; i < job.rc.size(); ++i)
        if (job.rc[i]) {
            g_err = job.msg[i];
            return job.rc[i];
        }
    return SCICONE_OK;
}
}  // namespace

int scicone_batch_compute_t_prime_scores(scicone_chain* const* chains, const scicone_tree* const* t_primes,
                                         const int32_t* node_idx, int32_t n) {
    if (!chains || !t_primes || !node_idx || n <= 0) return fail(SCICONE_ERR_ARG, "batch: null or empty");
    PerChainJob job{chains, t_primes, node_idx, std::vector<int>(n, 0), std::vector<std::string>(n)};
    scicone::ForkJoinPool::instance().run(n, [](int i, void* ctx) {
        PerChainJob& j = *static_cast<PerChainJob*>(ctx);
        j.rc[i] = scicone_compute_t_prime_scores(j.chains[i], j.trees[i], j.arg[i]);
        if (j.rc[i]) j.msg[i] = g_err;
    }, &job);
    return finish_job(job);
}

int scicone_batch_settle(scicone_chain* const* chains, const int32_t* accept, int32_t n) {
    if (!chains || !accept || n <= 0) return fail(SCICONE_ERR_ARG, "batch: null or empty");
    PerChainJob job{chains, nullptr, accept, std::vector<int>(n, 0), std::vector<std::string>(n)};
    scicone::ForkJoinPool::instance().run(n, [](int i, void* ctx) {
        PerChainJob& j = *static_cast<PerChainJob*>(ctx);
        j.rc[i] = j.arg[i] ? scicone_accept(j.chains[i]) : scicone_reject(j.chains[i]);
        if (j.rc[i]) j.msg[i] = g_err;
    }, &job);
    return finish_job(job);
}

int scicone_t_n_nodes(scicone_chain* ch, int32_t* n_nodes) {
    if (!ch || !n_nodes) return fail(SCICONE_ERR_ARG, "null argument");
    *n_nodes = (int32_t)ch->t.size();
    return SCICONE_OK;
}
int scicone_t_node_ids(scicone_chain* ch, int32_t* ids) {
    if (!ch || !ids) return fail(SCICONE_ERR_ARG, "null argument");
    int i = 0;
    for (auto& kv : ch->t) ids[i++] = kv.first;
    return SCICONE_OK;
}

int scicone_read_t_scores(scicone_chain* ch, double* out) {
    if (!ch || !out) return fail(SCICONE_ERR_ARG, "null argument");
    int rc = ch->ds->set_device();
    if (rc) return rc;
    std::vector<unsigned> rows;
    for (auto& kv : ch->t) rows.push_back(kv.second.row);
    return gather_rows(ch->ds, rows, out);
}
int scicone_read_t_sums(scicone_chain* ch, double* out) {
    if (!ch || !out) return fail(SCICONE_ERR_ARG, "null argument");
    return read_row(ch->ds, ch->sums[0], out, ch->ds->M);
}
int scicone_read_t_maxs(scicone_chain* ch, int32_t* ids, double* vals) {
    if (!ch || !ids || !vals) return fail(SCICONE_ERR_ARG, "null argument");
    if (!ch->ds->maxs) return fail(SCICONE_ERR_STATE, "t_maxs only exists with max_scoring");
    int rc = read_row(ch->ds, ch->maxid[0], ids, ch->ds->M);
    return rc ? rc : read_row(ch->ds, ch->sums[0], vals, ch->ds->M);
}
int scicone_read_t_prime_sums(scicone_chain* ch, double* out) {
    if (!ch || !out) return fail(SCICONE_ERR_ARG, "null argument");
    if (!ch->evaluated) return fail(SCICONE_ERR_STATE, "no evaluated proposal");
    return read_row(ch->ds, ch->sums[1], out, ch->ds->M);
}
int scicone_read_t_prime_maxs(scicone_chain* ch, int32_t* ids, double* vals) {
    if (!ch || !ids || !vals) return fail(SCICONE_ERR_ARG, "null argument");
    if (!ch->evaluated) return fail(SCICONE_ERR_STATE, "no evaluated proposal");
    if (!ch->ds->maxs) return fail(SCICONE_ERR_STATE, "t_prime_maxs only exists with max_scoring");
    int rc = read_row(ch->ds, ch->maxid[1], ids, ch->ds->M);
    return rc ? rc : read_row(ch->ds, ch->sums[1], vals, ch->ds->M);
}
int scicone_t_prime_n_rescored(scicone_chain* ch, int32_t* n) {
    if (!ch || !n) return fail(SCICONE_ERR_ARG, "null argument");
    *n = (int32_t)ch->p.size();
    return SCICONE_OK;
}
int scicone_t_prime_cost(scicone_chain* ch, double* cost) {
    if (!ch || !cost) return fail(SCICONE_ERR_ARG, "null argument");
    if (!ch->compiled) return fail(SCICONE_ERR_STATE, "no compiled proposal");
    *cost = ch->cost;
    return SCICONE_OK;
}
int scicone_read_t_prime_scores(scicone_chain* ch, int32_t* ids, double* out) {
    if (!ch || !out) return fail(SCICONE_ERR_ARG, "null argument");
    if (!ch->evaluated) return fail(SCICONE_ERR_STATE, "no evaluated proposal");
    int rc = ch->ds->set_device();
    if (rc) return rc;
    std::vector<unsigned> rows;
    int i = 0;
    for (auto& kv : ch->p) {
        if (ids) ids[i++] = kv.first;
        rows.push_back(kv.second.row);
    }
    return gather_rows(ch->ds, rows, out);
}

int scicone_assign_cells_to_nodes(scicone_chain* ch, const scicone_tree* t, int32_t* cell_node_ids, int32_t* cell_region_cn) {
    if (!ch || !cell_node_ids) return fail(SCICONE_ERR_ARG, "null argument");
    scicone_dataset* ds = ch->ds;
    int rc = validate_tree(t, ds->K);
    if (rc) return rc;
    if (ch->t.empty()) return fail(SCICONE_ERR_STATE, "assign_cells_to_nodes before compute_t_table");
    rc = ds->set_device();
    if (rc) return rc;
    for (int l = 1; l < scicone_dataset::kLanes; ++l)
        if (ds->lane_stream[l]) CUDA_TRY(cudaStreamSynchronize(ds->lane_stream[l]));
    const int n = (int)ch->t.size(), M = ds->M, K = ds->K;
    std::vector<unsigned> rows;
    std::vector<int> ids;
    std::vector<int> cn((size_t)n * K);
    for (auto& kv : ch->t) {
        int idx = -1;
        for (int i = 0; i < t->n_nodes; ++i)
            if (t->node_id[i] == kv.first) idx = i;
        if (idx < 0) return fail(SCICONE_ERR_ARG, "node id %d of the committed table is not in the tree", kv.first);
        int* dst = cn.data() + rows.size() * K;
        for (int k = 0; k < K; ++k) dst[k] = ds->neutral[k];
        for (int e = t->gen_off[idx]; e < t->gen_off[idx + 1]; ++e) dst[t->gen_region[e]] += t->gen_value[e];
        rows.push_back(kv.second.row);
        ids.push_back(kv.first);
    }
    unsigned* d_rows = nullptr;
    int *d_ids = nullptr, *d_cn = nullptr, *d_out_id = nullptr, *d_out_col = nullptr, *d_out_cn = nullptr;
    cudaError_t e = cudaMalloc(&d_rows, sizeof(unsigned) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_ids, sizeof(int) * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_cn, sizeof(int) * (size_t)n * K);
    if (e == cudaSuccess) e = cudaMalloc(&d_out_id, sizeof(int) * M);
    if (e == cudaSuccess) e = cudaMalloc(&d_out_col, sizeof(int) * M);
    if (e == cudaSuccess && cell_region_cn) e = cudaMalloc(&d_out_cn, sizeof(int) * (size_t)M * K);
    if (e == cudaSuccess) {
        cudaMemcpyAsync(d_rows, rows.data(), sizeof(unsigned) * n, cudaMemcpyHostToDevice, ds->stream);
        cudaMemcpyAsync(d_ids, ids.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ds->stream);
        cudaMemcpyAsync(d_cn, cn.data(), sizeof(int) * (size_t)n * K, cudaMemcpyHostToDevice, ds->stream);
        argmax_kernel<<<(M + 127) / 128, 128, 0, ds->stream>>>(ds->pool.base_ptr(), (long long)ds->Mp, d_rows, d_ids, n, M, d_out_id, d_out_col);
        cudaMemcpyAsync(cell_node_ids, d_out_id, sizeof(int) * M, cudaMemcpyDeviceToHost, ds->stream);
        if (cell_region_cn) {
            const long long tot = (long long)M * K;
            expand_cn_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ds->stream>>>(d_out_col, d_cn, M, K, d_out_cn);
            cudaMemcpyAsync(cell_region_cn, d_out_cn, sizeof(int) * (size_t)M * K, cudaMemcpyDeviceToHost, ds->stream);
        }
        e = cudaStreamSynchronize(ds->stream);
    }
    cudaFree(d_rows); cudaFree(d_ids); cudaFree(d_cn); cudaFree(d_out_id); cudaFree(d_out_col); cudaFree(d_out_cn);
    if (e != cudaSuccess) return fail(SCICONE_ERR_CUDA, "assign_cells_to_nodes: %s", cudaGetErrorString(e));
    return SCICONE_OK;
}

int scicone_condense_matrix(int32_t device, const double* D_bins, int64_t n_cells, int64_t n_bins, const int32_t* region_sizes,
                            int32_t n_regions, double* D_regions, double* kernel_us) {
    if (!D_bins || !region_sizes || !D_regions || n_cells <= 0 || n_bins <= 0 || n_regions <= 0)
        return fail(SCICONE_ERR_ARG, "condense_matrix: null buffer or empty matrix");
    std::vector<int> off(n_regions + 1, 0);
    for (int r = 0; r < n_regions; ++r) {
        if (region_sizes[r] <= 0) return fail(SCICONE_ERR_ARG, "condense_matrix: region %d has size %d", r, region_sizes[r]);
        off[r + 1] = off[r] + region_sizes[r];
    }
    const long long n_used = off[n_regions];
    if (n_used > n_bins) return fail(SCICONE_ERR_ARG, "condense_matrix: region sizes add up to %lld bins, the matrix has %lld", n_used, (long long)n_bins);
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return fail(SCICONE_ERR_CUDA, "condense_matrix: no such CUDA device; this library has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    // rows travel in chunks of at most 2 GB: upload, condense, download
    const long long chunk_rows = std::max<long long>(kK0Rows, std::min<long long>(n_cells, (2048ll << 20) / (n_bins * (long long)sizeof(double))));
    double *d_bins = nullptr, *d_out = nullptr;
    int *d_off = nullptr, *d_blocks = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaMalloc(&d_bins, sizeof(double) * chunk_rows * n_bins + 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(double) * chunk_rows * n_regions);
    if (e == cudaSuccess) e = cudaMalloc(&d_off, sizeof(int) * off.size());
    if (e == cudaSuccess) e = cudaMalloc(&d_blocks, sizeof(int) * (kK0MaxBlocks + 2));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(condense_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kK0SmemBudget + 8 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(condense_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kK0SmemBudget + 8 * 1024));
    {
        std::lock_guard<std::mutex> lk(g_drv_mu);
        g_drv.load();
    }
    // a tensor map needs a row pitch that is a multiple of 16 bytes (and coordinates that fit 32 bits)
    const bool tensor = g_drv.TensorMapEncodeTiled && n_bins % 2 == 0 && n_bins < (1ll << 31) && !(getenv("SCICONE_K0_ROWCOPIES") && atoi(getenv("SCICONE_K0_ROWCOPIES")));
    double total_us = 0.0;
    for (long long r0 = 0; e == cudaSuccess && r0 < n_cells; r0 += chunk_rows) {
        const long long rows = std::min(chunk_rows, (long long)n_cells - r0);
        e = cudaMemcpyAsync(d_bins, D_bins + r0 * n_bins, sizeof(double) * rows * n_bins, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        const long long groups = (rows + kK0Rows - 1) / kK0Rows;
        // Work items = row groups x region blocks: enough of them to give every warp slot of the GPU a few, but none shorter
        // than ~16 tiles (the ring has to fill and drain per item).  Blocks hold about the same number of bins each.
        int n_blocks = (int)std::max<long long>(1, std::min<long long>({(long long)kK0MaxBlocks, (long long)n_regions, n_used / 600 + 1,
                                                                         (3ll * n_sm * kK0MaxWarps + groups - 1) / groups}));
        std::vector<int> blocks(1, 0);
        for (int b = 1; b < n_blocks; ++b) {
            const long long want = n_used * b / n_blocks;
            int r = blocks.back() + 1;
            while (r < n_regions && off[r] < want) ++r;
            if (r >= n_regions) break;
            blocks.push_back(r);
        }
        blocks.push_back(n_regions);
        n_blocks = (int)blocks.size() - 1;
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_blocks, blocks.data(), sizeof(int) * blocks.size(), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) break;
        const long long n_work = groups * n_blocks;
        // warps per CTA and ring depth: the active warps of an SM together keep ~100 KB or more in flight
        const int wpc = (int)std::max<long long>(1, std::min<long long>(kK0MaxWarps, (n_work + n_sm - 1) / n_sm));
        const size_t index_bytes = (size_t)(n_regions + 1 <= kK0MaxIndex ? n_regions + 1 : 0) * sizeof(int);
        const size_t stage_bytes = k0_stage_bytes(tensor);
        const int n_stages = (int)std::max<size_t>(2, std::min<size_t>(kK0MaxStages, (kK0SmemBudget - index_bytes) / ((size_t)wpc * (stage_bytes + 8))));
        const size_t smem = (size_t)wpc * n_stages * (stage_bytes + 8) + index_bytes + 16;
        const unsigned grid = (unsigned)std::min<long long>(n_sm, (n_work + wpc - 1) / wpc);
        CUtensorMap map;
        memset(&map, 0, sizeof map);
        bool use_tensor = tensor;
        if (use_tensor) {
            const cuuint64_t dims[2] = {(cuuint64_t)n_bins, (cuuint64_t)rows};
            const cuuint64_t strides[1] = {(cuuint64_t)n_bins * sizeof(double)};
            const cuuint32_t box[2] = {(cuuint32_t)kK0TileTensor, (cuuint32_t)kK0Rows};
            const cuuint32_t estr[2] = {1, 1};
            CUresult cr = g_drv.TensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d_bins, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            use_tensor = cr == CUDA_SUCCESS;
        }
        cudaEventRecord(e0, st);
        if (use_tensor)
            condense_rows_kernel<true><<<grid, 32 * wpc, smem, st>>>(map, d_bins, rows, n_bins, n_used, d_off, n_regions, n_stages, d_blocks, n_blocks, d_out);
        else {  // one-row copies: 33 KB stages, so fewer warps per CTA
            const size_t sb = k0_stage_bytes(false);
            const int wpc1 = (int)std::max<size_t>(1, std::min<size_t>((size_t)wpc, (kK0SmemBudget - index_bytes) / (2 * (sb + 8))));
            const int ns = (int)std::max<size_t>(2, std::min<size_t>(kK0MaxStages, (kK0SmemBudget - index_bytes) / ((size_t)wpc1 * (sb + 8))));
            const unsigned grid1 = (unsigned)std::min<long long>(n_sm, (n_work + wpc1 - 1) / wpc1);
            condense_rows_kernel<false><<<grid1, 32 * wpc1, (size_t)wpc1 * ns * (sb + 8) + index_bytes + 16, st>>>(map, d_bins, rows, n_bins, n_used, d_off, n_regions,
                                                                                                               ns, d_blocks, n_blocks, d_out);
        }
        cudaEventRecord(e1, st);
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(D_regions + r0 * n_regions, d_out, sizeof(double) * rows * n_regions, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            total_us += 1e3 * ms;
        }
    }
    cudaFree(d_bins); cudaFree(d_out); cudaFree(d_off); cudaFree(d_blocks);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (e != cudaSuccess) return fail(SCICONE_ERR_CUDA, "condense_matrix: %s", cudaGetErrorString(e));
    if (kernel_us) *kernel_us = total_us;
    return SCICONE_OK;
}

int scicone_condense_clusters(int32_t device, const double* D, int64_t n_cells, int32_t n_cols, const int32_t* assignments,
                              int32_t n_clusters, double* means, int32_t* sizes) {
    if (!D || !assignments || !means || !sizes || n_cells <= 0 || n_cols <= 0 || n_clusters <= 0)
        return fail(SCICONE_ERR_ARG, "condense_clusters: null buffer or empty matrix");
    std::vector<int> first(n_clusters + 1, 0), order((size_t)n_cells);
    for (int64_t i = 0; i < n_cells; ++i) {
        if (assignments[i] < 0 || assignments[i] >= n_clusters)
            return fail(SCICONE_ERR_ARG, "condense_clusters: assignment %d of row %lld outside [0, %d)", assignments[i], (long long)i, n_clusters);
        first[assignments[i] + 1]++;
    }
    for (int c = 0; c < n_clusters; ++c) {
        sizes[c] = first[c + 1];
        first[c + 1] += first[c];
    }
    {
        std::vector<int> fill(first.begin(), first.end() - 1);
        for (int64_t i = 0; i < n_cells; ++i) order[fill[assignments[i]]++] = (int)i;  // stable: row order inside a cluster
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return fail(SCICONE_ERR_CUDA, "condense_clusters: no such CUDA device; this library has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    double *dD = nullptr, *dM = nullptr;
    int *dO = nullptr, *dF = nullptr;
    cudaError_t e = cudaMalloc(&dD, sizeof(double) * n_cells * n_cols);
    if (e == cudaSuccess) e = cudaMalloc(&dM, sizeof(double) * (size_t)n_clusters * n_cols);
    if (e == cudaSuccess) e = cudaMalloc(&dO, sizeof(int) * n_cells);
    if (e == cudaSuccess) e = cudaMalloc(&dF, sizeof(int) * first.size());
    if (e == cudaSuccess) e = cudaMemcpy(dD, D, sizeof(double) * n_cells * n_cols, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dO, order.data(), sizeof(int) * n_cells, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dF, first.data(), sizeof(int) * first.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        cluster_means_kernel<<<dim3((n_cols + 127) / 128, n_clusters), 128>>>(dD, n_cols, dO, dF, n_clusters, dM);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(means, dM, sizeof(double) * (size_t)n_clusters * n_cols, cudaMemcpyDeviceToHost);
    cudaFree(dD); cudaFree(dM); cudaFree(dO); cudaFree(dF);
    if (e != cudaSuccess) return fail(SCICONE_ERR_CUDA, "condense_clusters: %s", cudaGetErrorString(e));
    return SCICONE_OK;
}

int scicone_test_lgamma(int32_t device, const double* x, int32_t n, double* out) {
    if (!x || !out || n <= 0) return fail(SCICONE_ERR_ARG, "test_lgamma: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
        return fail(SCICONE_ERR_CUDA, "test_lgamma: no such CUDA device; this library has no CPU fallback");
    CUDA_TRY(cudaSetDevice(device));
    double *dx = nullptr, *dout = nullptr;
    CUDA_TRY(cudaMalloc(&dx, sizeof(double) * n));
    cudaError_t e = cudaMalloc(&dout, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        lgam_probe_kernel<<<(n + 255) / 256, 256>>>(dx, dout, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(dout);
    if (e != cudaSuccess) return fail(SCICONE_ERR_CUDA, "test_lgamma: %s", cudaGetErrorString(e));
    return SCICONE_OK;
}

int scicone_parallel_for(int32_t n, void (*fn)(int32_t index, void* ctx), void* ctx) {
    if (n < 0 || !fn) return fail(SCICONE_ERR_ARG, "parallel_for: bad argument");
    scicone::ForkJoinPool::instance().run(n, fn, ctx);
    return SCICONE_OK;
}
int scicone_host_threads(void) { return scicone::ForkJoinPool::instance().n_threads(); }

int scicone_comm_unique_id(uint8_t out[128]) {
    if (!out) return fail(SCICONE_ERR_ARG, "null argument");
    if (!g_nccl.load()) return fail(SCICONE_ERR_COMM, "libnccl.so.2 not found: %s", dlerror());
    NcclApi::UniqueId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail(SCICONE_ERR_COMM, "ncclGetUniqueId failed (%d)", rc);
    memcpy(out, &id, 128);
    return SCICONE_OK;
}

namespace {
// Map every rank's gather buffer into every process (CUDA IPC over NVLink peer access).  If any rank cannot (no peer
// access between two devices, IPC unavailable in the container), ALL ranks stay with the NCCL all-gather.
int setup_peer_exchange(scicone_dataset* ds) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle travels as 8 doubles");
    const int W = ds->world;
    const size_t bytes = ds->xchg_flag_bytes() + sizeof(double) * scicone_dataset::kLanes * 2 * W * ds->xchg_block();
    bool ok = cudaMalloc(&ds->xchg, bytes) == cudaSuccess && cudaMemset(ds->xchg, 0, bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
    double mine[9] = {0};
    cudaIpcMemHandle_t h;
    if (ok) ok = cudaIpcGetMemHandle(&h, ds->xchg) == cudaSuccess;
    if (ok) memcpy(mine, &h, 64);
    mine[8] = ok ? 1.0 : 0.0;
    std::vector<double> all;
    int rc = gather_doubles(ds, mine, 9, all);
    if (rc) return rc;
    for (int p = 0; p < W; ++p) ok = ok && all[(size_t)p * 9 + 8] == 1.0;
    ds->xchg_peer.assign(W, nullptr);
    if (ok)
        for (int p = 0; p < W && ok; ++p) {
            if (p == ds->rank) {
                ds->xchg_peer[p] = ds->xchg;
                continue;
            }
            cudaIpcMemHandle_t hp;
            memcpy(&hp, &all[(size_t)p * 9], 64);
            ok = cudaIpcOpenMemHandle(&ds->xchg_peer[p], hp, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        }
    cudaGetLastError();  // a failed probe must not poison later calls
    // every rank must have mapped every buffer
    double mapped = ok ? 1.0 : 0.0;
    rc = gather_doubles(ds, &mapped, 1, all);
    if (rc) return rc;
    for (double v : all) ok = ok && v == 1.0;
    if (!ok) {
        fprintf(stderr, "libscicone_score: peer-memory exchange unavailable on rank %d, using the NCCL all-gather\n", ds->rank);
        for (int p = 0; p < W; ++p)
            if (p != ds->rank && ds->xchg_peer[p]) cudaIpcCloseMemHandle(ds->xchg_peer[p]);
        ds->xchg_peer.clear();
        cudaFree(ds->xchg);
        ds->xchg = nullptr;
        cudaGetLastError();
        return SCICONE_OK;
    }
    std::vector<double*> data_tab((size_t)scicone_dataset::kLanes * 2 * W);
    std::vector<unsigned long long*> flag_tab((size_t)scicone_dataset::kLanes * W);
    for (int l = 0; l < scicone_dataset::kLanes; ++l)
        for (int p = 0; p < W; ++p) {
            flag_tab[(size_t)l * W + p] = ds->xchg_flags(ds->xchg_peer[p], l, ds->rank);
            for (int par = 0; par < 2; ++par) data_tab[((size_t)l * 2 + par) * W + p] = ds->xchg_data(ds->xchg_peer[p], l, par, ds->rank);
        }
    CUDA_TRY(cudaMalloc(&ds->d_xchg_data_tab, sizeof(double*) * data_tab.size()));
    CUDA_TRY(cudaMalloc(&ds->d_xchg_flag_tab, sizeof(unsigned long long*) * flag_tab.size()));
    CUDA_TRY(cudaMemcpy(ds->d_xchg_data_tab, data_tab.data(), sizeof(double*) * data_tab.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ds->d_xchg_flag_tab, flag_tab.data(), sizeof(unsigned long long*) * flag_tab.size(), cudaMemcpyHostToDevice));
    ds->xchg_on = true;
    return SCICONE_OK;
}
}  // namespace

int scicone_dataset_attach_comm(scicone_dataset* ds, const uint8_t unique_id[128], int32_t rank, int32_t world_size) {
    if (!ds || !unique_id || world_size < 1 || rank < 0 || rank >= world_size) return fail(SCICONE_ERR_ARG, "attach_comm: bad argument");
    if (ds->comm) return fail(SCICONE_ERR_STATE, "attach_comm: communicator already attached");
    if (!g_nccl.load()) return fail(SCICONE_ERR_COMM, "libnccl.so.2 not found");
    int rc = ds->set_device();
    if (rc) return rc;
    // one partition of the row pool per rank: the rows of a chain come from its owner's partition on every rank
    rc = ds->pool.init(ds->Mp, ds->device, world_size, rank);
    if (rc) return rc;
    NcclApi::UniqueId id;
    memcpy(&id, unique_id, 128);
    int nrc = g_nccl.CommInitRank(&ds->comm, world_size, id, rank);
    if (nrc != 0) return fail(SCICONE_ERR_COMM, "ncclCommInitRank failed (%d)", nrc);
    ds->rank = rank;
    ds->world = world_size;
    // every rank pads to the same chunk count: the largest shard a rank can hold
    const long long chunk_cells = (long long)kCells * kCtasPerChunk;
    const long long global_chunks = (ds->M_global + chunk_cells - 1) / chunk_cells;
    ds->max_chunks_rank = (int)std::max<long long>((global_chunks + world_size - 1) / world_size + 1, ds->n_chunks);
    // all ranks must agree on max_chunks_rank: exchange through the communicator
    {
        double mine = (double)ds->max_chunks_rank;
        std::vector<double> all;
        rc = gather_doubles(ds, &mine, 1, all);
        if (rc) return rc;
        for (double v : all) ds->max_chunks_rank = std::max(ds->max_chunks_rank, (int)v);
    }
    ds->cs_stride = ds->max_chunks_rank;
    rc = build_od_hist(ds);  // the histograms of the root score cover all cells of the run: summed over the ranks here
    if (rc) return rc;
    const int keep = ds->slots;
    ds->slots = 0;  // force reallocation with the gather buffers and the common chunk width
    rc = ds->ensure_slots(std::max(keep, 4));
    if (rc) return rc;
    const char* how = getenv("SCICONE_EXCHANGE");
    if (how && strcmp(how, "nccl") == 0) return SCICONE_OK;
    return setup_peer_exchange(ds);
}

int scicone_set_streams(scicone_dataset* ds, int32_t n) {
    if (!ds || n < 1 || n > scicone_dataset::kLanes) return fail(SCICONE_ERR_ARG, "set_streams: 1 to %d streams", scicone_dataset::kLanes);
    int rc = ds->set_device();
    if (rc) return rc;
    rc = ds->sync_lanes();
    if (rc) return rc;
    for (scicone_dataset::Segment& g : ds->seg)
        if (g.busy) return fail(SCICONE_ERR_STATE, "set_streams with launches outstanding");
    for (int l = 1; l < n; ++l)
        if (!ds->lane_stream[l]) CUDA_TRY(cudaStreamCreateWithFlags(&ds->lane_stream[l], cudaStreamNonBlocking));
    if (n > 1 && ds->world > 1 && !ds->xchg_on) {
        // cell-sharded: the all-gathers of the lanes run concurrently, each lane needs its own communicator.  Collective:
        // every rank calls scicone_set_streams at the same point, and all ranks submit their launches in the same order.
        if (!g_nccl.CommSplit) return fail(SCICONE_ERR_COMM, "set_streams: this NCCL has no ncclCommSplit (needs >= 2.18)");
        for (int l = 1; l < n; ++l) {
            if (ds->comm_lane[l]) continue;
            int nrc = g_nccl.CommSplit(ds->comm, 0, ds->rank, &ds->comm_lane[l], nullptr);
            if (nrc != 0 || !ds->comm_lane[l]) return fail(SCICONE_ERR_COMM, "ncclCommSplit failed (%d)", nrc);
        }
    }
    ds->n_lanes = n;
    return SCICONE_OK;
}

int scicone_set_profiling(scicone_dataset* ds, int32_t on) {
    if (!ds) return fail(SCICONE_ERR_ARG, "null dataset");
    ds->profiling = on != 0;
    return SCICONE_OK;
}
int scicone_last_kernel_time_us(scicone_dataset* ds, double* us) {
    if (!ds || !us) return fail(SCICONE_ERR_ARG, "null argument");
    *us = ds->last_kernel_us;
    return SCICONE_OK;
}
int scicone_kernel_time_stats(scicone_dataset* ds, double* sum_us, uint64_t* n_timed, int32_t reset) {
    if (!ds) return fail(SCICONE_ERR_ARG, "null dataset");
    if (sum_us) *sum_us = ds->sum_kernel_us;
    if (n_timed) *n_timed = ds->n_timed;
    if (reset) {
        ds->sum_kernel_us = ds->sum_build_us = 0.0;
        ds->n_timed = ds->n_build_timed = 0;
    }
    return SCICONE_OK;
}
int scicone_build_time_stats(scicone_dataset* ds, double* sum_us, uint64_t* n_timed) {
    if (!ds) return fail(SCICONE_ERR_ARG, "null dataset");
    if (sum_us) *sum_us = ds->sum_build_us;
    if (n_timed) *n_timed = ds->n_build_timed;
    return SCICONE_OK;
}
int scicone_counters(scicone_dataset* ds, double out[4], int32_t reset) {
    if (!ds || !out) return fail(SCICONE_ERR_ARG, "null argument");
    out[0] = (double)ds->h2d_bytes;
    out[1] = (double)ds->d2h_bytes;
    out[2] = ds->alg_bytes;
    out[3] = ds->n_scores;
    if (reset) {
        ds->h2d_bytes = ds->d2h_bytes = 0;
        ds->alg_bytes = ds->n_scores = 0.0;
    }
    return SCICONE_OK;
}
int scicone_build_counters(scicone_dataset* ds, double out[2], int32_t reset) {
    if (!ds || !out) return fail(SCICONE_ERR_ARG, "null argument");
    out[0] = ds->k1_bytes;
    out[1] = ds->k1_lgamma;
    if (reset) ds->k1_bytes = ds->k1_lgamma = 0.0;
    return SCICONE_OK;
}
int scicone_region_mark(scicone_dataset* ds, int32_t which) {
    if (!ds || which < 0 || which > 1) return fail(SCICONE_ERR_ARG, "region_mark: bad argument");
    int rc = ds->set_device();
    if (rc) return rc;
    if (!ds->region_ev[which]) CUDA_TRY(cudaEventCreate(&ds->region_ev[which]));
    for (int l = 1; l < scicone_dataset::kLanes; ++l)  // the mark covers every lane
        if (ds->lane_stream[l]) CUDA_TRY(cudaStreamSynchronize(ds->lane_stream[l]));
    CUDA_TRY(cudaEventRecord(ds->region_ev[which], ds->stream));
    return SCICONE_OK;
}
int scicone_region_elapsed_ms(scicone_dataset* ds, double* ms) {
    if (!ds || !ms || !ds->region_ev[0] || !ds->region_ev[1]) return fail(SCICONE_ERR_ARG, "region_elapsed: marks missing");
    CUDA_TRY(cudaEventSynchronize(ds->region_ev[1]));
    float f = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&f, ds->region_ev[0], ds->region_ev[1]));
    *ms = f;
    return SCICONE_OK;
}
int scicone_device_bytes(scicone_dataset* ds, uint64_t* bytes) {
    if (!ds || !bytes) return fail(SCICONE_ERR_ARG, "null argument");
    *bytes = (uint64_t)ds->pool.in_use * ds->Mp * sizeof(double) + (uint64_t)ds->Mp * ds->K * (sizeof(double) + sizeof(unsigned short));
    return SCICONE_OK;
}
int scicone_kernel_launches(scicone_dataset* ds, uint64_t* n) {
    if (!ds || !n) return fail(SCICONE_ERR_ARG, "null argument");
    *n = ds->n_launches;
    return SCICONE_OK;
}

}  // extern "C"
