// Device side of the B200 tree-scoring engine (sm_100a, fp64, no tensor cores:
// the path is a sparse gather + scan + reduce, HBM/latency bound).
//
// Layout in HBM (cells are the fastest-varying dimension everywhere, so a warp
// touches one 256-byte segment per row it reads):
//   Dt        [K][Mp]   region-major copy of the count matrix D
//   S         [Mp]      per-cell read total  sum_k D[j][k]   (k ascending, as std::accumulate)
//   w         [Mp]      cluster sizes (int32, 0 in the padding)
//   lgamma rows  [Mp]   L(k,cn)[j] = lgamma(D[j][k] + (nu*cn)*r_k), built lazily per (region, copy number, nu)
//   score rows   [Mp]   one row per tree node: attachment score of every cell to that node
//   aggregates   [Mp]   per-cell log-sum-exp (or max) over the nodes, + argmax id in max mode
//
// One thread owns one cell for the whole proposal: it walks the rescored nodes
// parent-before-child (Tree::compute_stack, reference scicone/src/Tree.h:260-286),
// evaluates Tree::compute_score (Tree.h:163-244) as gathers from the lgamma rows,
// then folds the new scores into the cell's aggregate exactly like
// Inference::compute_t_prime_sums (scicone/src/Inference.h:1069-1196) /
// MathOp::log_replace_sum (scicone/src/MathOp.cpp:307-361), and the CTA reduces
// aggregate * cluster_size into a partial of the tree score (Inference.h:292-294).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace scicone {

constexpr int kBlock = 128;         // cells per CTA (one thread per cell)
constexpr int kSlots = 16;          // per-cell values of rescored ancestors kept in shared memory
constexpr int kCtasPerChunk = 8;    // 8 CTAs = 1024 cells: fixed-order reduction unit (multi-GPU invariant)

enum : unsigned {
    TASK_ROOT = 1u,      // tree root: score 0 (Tree::compute_root_score, Tree.h:145-160)
    TASK_REUSE_G = 2u,   // lgamma(S + nu*z_parent) equals the value the parent task kept in its slot
    EV_BUILD_N = 1u,     // the node-side lgamma row does not exist yet: compute and store it
    EV_BUILD_P = 2u,     // same for the parent-side row
    PROP_MAX = 1u,       // max scoring (Inference::max_scoring)
    PROP_OD = 2u,        // overdispersed likelihood (globals is_overdispersed)
    PROP_FULL = 4u,      // full table pass (Inference::compute_t_table): no committed state is read
    PROP_SCRATCH = 8u,   // to_add.size() >= unchanged_vals.size(): recompute the log-sum from scratch
    PROP_WITH_OD = 16u   // the proposal changes nu: also evaluate the root ("overdispersed") score at the new nu
};

struct alignas(16) DevTask {
    double* dst;               // score row being written
    const double* parent_row;  // committed score row of the parent (subtree roots only)
    double lg_nz, lg_nzp;      // OD: lgamma(nu*z), lgamma(nu*z_parent);  non-OD: log(z), log(z_parent)
    double nz, nzp;            // OD: nu*z, nu*z_parent
    int ev_begin, ev_end;
    int node_id;
    short parent_slot, own_slot;  // shared-memory slots (-1: none)
    unsigned flags;
    int pad;
};

struct alignas(16) DevEvent {
    double* rowN;        // L(k, cn_node)
    double* rowP;        // L(k, cn_parent)
    const double* drow;  // Dt[k]
    double cN, cP;       // OD: lgamma((nu*cn)*r) node / parent;  non-OD: cN = log(cn_node) - log(cn_parent)
    double argN, argP;   // (nu*cn)*r_k
    unsigned flags;
    int pad;
};

struct alignas(16) DevUnch {
    const double* row;
    int id;
    int pad;
};

struct alignas(16) DevHeader {
    int n_tasks, n_old, n_unch, deleted_id;
    unsigned flags;
    int n_cells;  // M of this shard
    int n_add;    // distinct rescored node ids (== entries of addorder)
    int pad1;
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
    unsigned off_tasks, off_events, off_addorder, off_old, off_unch;
    unsigned off_od;      // DevOdHeader of the root-score part (PROP_WITH_OD)
};

struct alignas(16) DevOdHeader {
    int n_regions, n_cells;
    unsigned flags;
    int pad0;
    const double* S;
    const int* w;
    const double* Dt;
    long long row_stride;  // Mp
    double* cell_out;      // per-cell root score or nullptr
    double* partial;
    double* chunk_sums;
    double* total;
    unsigned* counter;
    double lg_nz, nz;      // OD: lgamma(nu*z_root), nu*z_root;  non-OD: lg_nz = log(sum_r)
    unsigned off_rows, off_arg, off_c, off_build;
};

__device__ __forceinline__ double block_sum(double v, double* warp_buf) {
    // fixed-shape reduction: xor-shuffle tree inside each warp, warps added in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_buf[wid] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kBlock / 32; ++i) s += warp_buf[i];
    return s;  // valid in thread 0
}

// Last CTA to arrive adds the per-CTA partials in a fixed order: first inside
// 1024-cell chunks, then the chunks.  The result does not depend on the CTA
// schedule, and (because shards are whole chunks) not on the number of GPUs.
// Two sums travel together: [0] the tree score, [1] the root score (0 when not requested).
__device__ __forceinline__ void finish_reduction(double cta_sum0, double cta_sum1, double* partial, double* chunk_sums,
                                                 double* total, unsigned* counter, double* smem_chunks) {
    __shared__ bool is_last;
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
        for (int base = 0; base < n_chunks; base += kBlock) {
            int c = base + threadIdx.x;
            double cs = 0.0;
            if (c < n_chunks) {
                const int b0 = c * kCtasPerChunk;
                const int b1 = min(b0 + kCtasPerChunk, n_cta);
                for (int b = b0; b < b1; ++b) cs += __ldcg(part + b);
                chunks[c] = cs;
            }
            smem_chunks[threadIdx.x] = cs;
            __syncthreads();
            if (threadIdx.x == 0) {
                const int lim = min(kBlock, n_chunks - base);
                for (int i = 0; i < lim; ++i) grand += smem_chunks[i];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) total[v] = grand;
    }
    if (threadIdx.x == 0) *counter = 0u;
}

// Tree::get_od_root_score, Tree.h:1774-1810, for one cell (also materialises the neutral-state lgamma rows it
// computes anyway).
__device__ __forceinline__ double od_root_cell(const DevOdHeader& H, const unsigned char* blob, int j) {
    double* const* rows = reinterpret_cast<double* const*>(blob + H.off_rows);
    const double* arg = reinterpret_cast<const double*>(blob + H.off_arg);
    const double* cc = reinterpret_cast<const double*>(blob + H.off_c);
    const unsigned char* build = blob + H.off_build;
    const double S = H.S[j];
    const int K = H.n_regions;
    double od = 0.0;
    if (H.flags & PROP_OD) {
        od += H.lg_nz;
        od -= lgamma(S + H.nz);
        for (int k = 0; k < K; ++k) {
            double v;
            if (build[k]) {
                v = lgamma(H.Dt[(long long)k * H.row_stride + j] + arg[k]);
                if (rows[k]) rows[k][j] = v;
            } else
                v = rows[k][j];
            od += v;
            od -= cc[k];
        }
    } else {
        const double term1 = -S * H.lg_nz;
        double term2 = 0.0;
        for (int k = 0; k < K; ++k) term2 += H.Dt[(long long)k * H.row_stride + j] * arg[k];
        od += term1 + term2;
    }
    return od;
}

// ---------------------------------------------------------------------------------
// Proposal kernel: grid = (ceil(M / kBlock), n_proposals).
// ---------------------------------------------------------------------------------
// The arena starts with one 32-bit blob offset per proposal.
__global__ void __launch_bounds__(kBlock) score_proposals_kernel(const unsigned char* __restrict__ arena) {
    const unsigned* blob_off = reinterpret_cast<const unsigned*>(arena);
    __shared__ double slot_sc[kSlots][kBlock];
    __shared__ double slot_g[kSlots][kBlock];
    __shared__ double red_buf[kBlock];

    const unsigned char* blob = arena + blob_off[blockIdx.y];
    const DevHeader& H = *reinterpret_cast<const DevHeader*>(blob);
    const DevTask* tasks = reinterpret_cast<const DevTask*>(blob + H.off_tasks);
    const DevEvent* events = reinterpret_cast<const DevEvent*>(blob + H.off_events);
    const int* addorder = reinterpret_cast<const int*>(blob + H.off_addorder);
    const double* const* old_rows = reinterpret_cast<const double* const*>(blob + H.off_old);
    const DevUnch* unch = reinterpret_cast<const DevUnch*>(blob + H.off_unch);

    const int tid = threadIdx.x;
    const int j = blockIdx.x * kBlock + tid;
    const bool live = j < H.n_cells;
    const unsigned pflags = H.flags;
    const bool od = pflags & PROP_OD;
    double contrib = 0.0, contrib_od = 0.0;

    if (live) {
        const double S = H.S[j];
        // ---- node scores: Tree::compute_stack / compute_score over the rescored subtree(s) ----
        const int n_tasks = H.n_tasks;
        for (int t = 0; t < n_tasks; ++t) {
            const DevTask tk = tasks[t];
            double ll;
            if (tk.flags & TASK_ROOT) {
                ll = 0.0;
                if (tk.own_slot >= 0) {
                    slot_sc[tk.own_slot][tid] = 0.0;
                    if (od) slot_g[tk.own_slot][tid] = lgamma(S + tk.nz);
                }
                tk.dst[j] = 0.0;
                continue;
            }
            double g_parent = 0.0;
            bool have_gp = false;
            if (tk.parent_row != nullptr)
                ll = tk.parent_row[j];
            else {
                ll = slot_sc[tk.parent_slot][tid];
                if (tk.flags & TASK_REUSE_G) {
                    g_parent = slot_g[tk.parent_slot][tid];
                    have_gp = true;
                }
            }
            if (od) {
                for (int e = tk.ev_begin; e < tk.ev_end; ++e) {
                    const DevEvent ev = events[e];
                    double vn, vp;
                    if (ev.flags & EV_BUILD_N) {
                        vn = lgamma(ev.drow[j] + ev.argN);
                        ev.rowN[j] = vn;
                    } else
                        vn = ev.rowN[j];
                    if (ev.flags & EV_BUILD_P) {
                        vp = lgamma(ev.drow[j] + ev.argP);
                        ev.rowP[j] = vp;
                    } else
                        vp = ev.rowP[j];
                    ll += vn;
                    ll -= vp;
                    ll -= ev.cN;
                    ll += ev.cP;
                }
                ll += tk.lg_nz;
                ll -= tk.lg_nzp;
                const double g = lgamma(S + tk.nz);
                ll -= g;
                if (!have_gp) g_parent = lgamma(S + tk.nzp);
                ll += g_parent;
                if (tk.own_slot >= 0) slot_g[tk.own_slot][tid] = g;
            } else {
                for (int e = tk.ev_begin; e < tk.ev_end; ++e) {
                    const DevEvent ev = events[e];
                    ll += ev.drow[j] * ev.cN;
                }
                ll -= S * tk.lg_nz;
                ll += S * tk.lg_nzp;
            }
            tk.dst[j] = ll;
            if (tk.own_slot >= 0) slot_sc[tk.own_slot][tid] = ll;
        }

        // ---- per-cell aggregate: Inference::compute_t_prime_sums ----
        const int n_unch = H.n_unch;
        const int n_add = H.n_add;
        double res;
        if (pflags & PROP_MAX) {
            double newbest = -DBL_MAX;
            int newid = 0;
            bool prev_in_new = (pflags & PROP_FULL) != 0;
            const int pm = (pflags & PROP_FULL) ? -1 : H.maxid_old[j];
            for (int a = 0; a < n_add; ++a) {
                const DevTask& tk = tasks[addorder[a]];
                const double v = tk.dst[j];
                if (tk.node_id == pm) prev_in_new = true;
                if (v > newbest) {
                    newbest = v;
                    newid = tk.node_id;
                }
            }
            int prev_id = pm;
            double prev_val = (pflags & PROP_FULL) ? 0.0 : H.sums_old[j];
            if (prev_in_new || pm == H.deleted_id) {
                double cur = -DBL_MAX;
                int uid = 0;
                double uval = 0.0;  // value-initialised std::pair when nothing is unchanged
                for (int u = 0; u < n_unch; ++u) {
                    const double v = unch[u].row[j];
                    if (v > cur) {
                        uid = unch[u].id;
                        uval = v;
                        cur = v;
                    }
                }
                prev_id = uid;
                prev_val = uval;
            }
            res = newbest;
            int rid = newid;
            if (prev_val > newbest) {
                res = prev_val;
                rid = prev_id;
            }
            H.maxid_new[j] = rid;
        } else {
            bool scratch = (pflags & (PROP_SCRATCH | PROP_FULL)) != 0;
            double sub = 0.0;
            if (!scratch) {  // MathOp::log_replace_sum, incremental branch
                const double sum_old = H.sums_old[j];
                const int n_old = H.n_old;
                double mx = sum_old;
                for (int o = 0; o < n_old; ++o) {
                    const double v = old_rows[o][j];
                    if (v > mx) mx = v;
                }
                double s = 0.0;
                s += exp(sum_old - mx);
                for (int o = 0; o < n_old; ++o) s -= exp(old_rows[o][j] - mx);
                if (s <= 1e-6)
                    scratch = true;  // possible precision loss: recompute from every node
                else
                    sub = log(s) + mx;
            }
            double mx = -DBL_MAX;
            for (int a = 0; a < n_add; ++a) {
                const double v = tasks[addorder[a]].dst[j];
                if (v > mx) mx = v;
            }
            if (scratch) {
                for (int u = 0; u < n_unch; ++u) {
                    const double v = unch[u].row[j];
                    if (v > mx) mx = v;
                }
            } else if (sub > mx)
                mx = sub;
            double s = 0.0;
            for (int a = 0; a < n_add; ++a) s += exp(tasks[addorder[a]].dst[j] - mx);
            if (scratch) {
                for (int u = 0; u < n_unch; ++u) s += exp(unch[u].row[j] - mx);
            } else
                s += exp(sub - mx);
            res = log(s) + mx;
        }
        if (res != res) atomicOr(H.status, 1u);
        H.sums_new[j] = res;
        contrib = res * (double)H.w[j];
        if (pflags & PROP_WITH_OD)  // Inference::compute_t_prime_od_scores, Inference.h:1421-1433
            contrib_od = od_root_cell(*reinterpret_cast<const DevOdHeader*>(blob + H.off_od), blob + H.off_od, j) * (double)H.w[j];
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
    finish_reduction(cta_sum, cta_od, H.partial, H.chunk_sums, H.total, H.counter, red_buf);
}

// ---------------------------------------------------------------------------------
// Root ("overdispersed root") score: Tree::get_od_root_score, Tree.h:1774-1810, and the
// weighted sum of Inference::compute_t_od_scores, Inference.h:1406-1419.
// Also materialises the neutral-state lgamma rows L(k, neutral_k) it computes anyway.
// grid = (ceil(M / kBlock), n_requests).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) od_root_kernel(const unsigned char* __restrict__ arena) {
    const unsigned* blob_off = reinterpret_cast<const unsigned*>(arena);
    __shared__ double red_buf[kBlock];
    const unsigned char* blob = arena + blob_off[blockIdx.y];
    const DevOdHeader& H = *reinterpret_cast<const DevOdHeader*>(blob);
    const int j = blockIdx.x * kBlock + threadIdx.x;
    double contrib = 0.0;
    if (j < H.n_cells) {
        const double od = od_root_cell(H, blob, j);
        if (H.cell_out) H.cell_out[j] = od;
        contrib = od * (double)H.w[j];
    }
    __syncthreads();
    const double cta_sum = block_sum(contrib, red_buf);
    __syncthreads();
    finish_reduction(0.0, cta_sum, H.partial, H.chunk_sums, H.total, H.counter, red_buf);
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
