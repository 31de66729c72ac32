/*
 * scicone_score.h - C ABI of the B200 tree-scoring engine (libscicone_score.so).
 *
 * The reference (cbg-ethz/SCICoNE) has no plugin / FFI boundary: its MCMC driver
 * `Inference` (scicone/src/Inference.h) calls the scoring code in `Tree`
 * (scicone/src/Tree.h) and `MathOp` (scicone/src/MathOp.cpp) directly and owns the
 * cells x nodes score tables as std::vector<std::map<int,double>>.  This header
 * cuts the seam at exactly those `Inference` methods.  Each entry point names
 * the reference member it replaces (file:line relative to /root/reference).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, int return codes (0 = ok, <0 = error;
 *     scicone_last_error() gives the message).  No exception crosses the ABI.
 *     Where the reference throws std::logic_error (Tree.h:170-171, size
 *     mismatch) the call returns SCICONE_ERR_ARG; where it asserts !isnan
 *     (Inference.h:1150,1191) the call returns SCICONE_ERR_NAN.
 *   - All host buffers are caller-owned and only read during the call.
 *   - A dataset owns the device copy of D (cells x regions), region sizes,
 *     neutral states and cluster sizes; it is immutable after creation and may
 *     be shared by any number of chains.
 *   - A chain owns what one reference `Inference` object owns for scoring:
 *     t_scores / t_sums / t_maxs and their primed counterparts
 *     (Inference.h:39-44), device resident.  A chain is not thread-safe.
 *   - There is no CPU fallback: every compute entry point needs a CUDA device
 *     and fails with SCICONE_ERR_CUDA otherwise.
 */
#ifndef SCICONE_SCORE_H
#define SCICONE_SCORE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCICONE_ABI_VERSION 2
#define SCICONE_MAX_SHARDED_BATCH 128 /* proposals per launch on a cell-sharded dataset (slots of the gather buffer) */

enum {
    SCICONE_OK = 0,
    SCICONE_ERR_ARG = -1,   /* bad argument / inconsistent sizes (reference: std::logic_error) */
    SCICONE_ERR_CUDA = -2,  /* CUDA runtime failure or no device */
    SCICONE_ERR_STATE = -3, /* call sequence violated (e.g. sums without scores) */
    SCICONE_ERR_NAN = -4,   /* a per-cell aggregate became NaN (reference: assert) */
    SCICONE_ERR_COMM = -5   /* multi-GPU communicator failure */
};

typedef struct scicone_dataset scicone_dataset;
typedef struct scicone_chain scicone_chain;

/*
 * Flat encoding of a reference `Tree` (Tree.h:36-118) / `Node` (Node.h:18-81).
 * Nodes are listed parent-before-child; entry 0 is the root.  Only the fields
 * that feed the score are carried: Node::id, parent link, Node::c_change and
 * Node::c (both std::map<u_int,int>, i.e. region-ascending), and Tree::nu.
 * Node::z is re-derived by the library with the reference's int truncation
 * (Tree.h:178-179,197,241; Node.h:25).
 */
typedef struct scicone_tree {
    int32_t n_nodes;           /* number of entries, root included */
    const int32_t* node_id;    /* [n_nodes] Node::id (root = 0), unique */
    const int32_t* parent;     /* [n_nodes] index of the parent entry, -1 for the root */
    const int32_t* chg_off;    /* [n_nodes+1] CSR offsets of Node::c_change */
    const int32_t* chg_region; /* region index, ascending inside a node */
    const int32_t* chg_value;  /* event value (may be 0, see Node.h:389-418) */
    const int32_t* gen_off;    /* [n_nodes+1] CSR offsets of Node::c (genotype relative to neutral) */
    const int32_t* gen_region; /* region index, ascending inside a node */
    const int32_t* gen_value;  /* non-zero copy-number offset */
    double nu;                 /* Tree::nu */
} scicone_tree;

typedef struct scicone_dataset_desc {
    int32_t n_cells;               /* rows of D handed to this dataset (this rank's shard) */
    int32_t n_regions;             /* columns of D */
    const double* D;               /* [n_cells][n_regions] row-major, as Utils::read_counts fills it (Utils.cpp:74-102) */
    const int32_t* region_sizes;   /* [n_regions]  r */
    const int32_t* neutral_states; /* [n_regions]  Tree::region_neutral_states */
    const int32_t* cluster_sizes;  /* [n_cells]    weights of the sum over cells (Inference.h:269,294) */
    double eta;                    /* globals.cpp eta: replaces copy number 0 (Tree.h:194-195) */
    int32_t is_overdispersed;      /* globals.cpp is_overdispersed (Tree.h:212) */
    int32_t max_scoring;           /* Inference::max_scoring (Inference.h:46) */
    int32_t device;                /* CUDA device ordinal */
    int64_t cell_offset;           /* global index of this shard's first cell (0 on a single GPU) */
    int64_t n_cells_global;        /* total cells over all ranks (== n_cells on a single GPU) */
} scicone_dataset_desc;

int scicone_abi_version(void);
const char* scicone_last_error(void);

/*
 * Uploads the matrix (transposed to region-major on the device) and the per-cell read totals.  D may hold any doubles;
 * where the counts of a region over a block of 2048 cells are non-negative integers in a narrow range (read counts are),
 * the lgamma rows of that block are evaluated once per possible count instead of once per cell - the results are
 * bit-identical either way (environment SCICONE_LUT=0 turns the tabulation off, for A/B runs).
 */
int scicone_dataset_create(scicone_dataset** out, const scicone_dataset_desc* desc);
void scicone_dataset_destroy(scicone_dataset* ds);

/*
 * Optional, before the first launch: set up the reduction scratch for launches of up to max_batch proposals and back
 * the first `rows` rows of (every rank's partition of) the row pool with memory, so that neither grows inside a timed
 * loop.  A chain holds one row of n_cells doubles per tree node plus one per node of a queued proposal, four rows of
 * per-cell aggregates and up to ~35 temporary rows.  No reference counterpart (the reference's tables are std::maps on
 * the heap).  scicone_max_batch returns the current capacity; the batch size cannot grow while launches are outstanding.
 */
int scicone_dataset_reserve(scicone_dataset* ds, int32_t max_batch, int64_t rows);
int scicone_max_batch(scicone_dataset* ds, int32_t* n);

/* One chain == the score tables of one reference `Inference` object (Inference.h:183-198). */
int scicone_chain_create(scicone_chain** out, scicone_dataset* ds);
/*
 * Cell-sharded runs with one owner rank per chain: the handle the OTHER ranks hold for that chain.  It owns no tables;
 * it only stages the proposals the owner compiled (scicone_import_t_prime -> scicone_batch_submit -> scicone_accept /
 * scicone_reject, which merely clear it).
 */
int scicone_chain_create_follower(scicone_chain** out, scicone_dataset* ds);
void scicone_chain_destroy(scicone_chain* ch);

/*
 * Inference::compute_t_table (Inference.h:224-279): full pass over every cell and
 * every node of `t`; fills t_scores, t_sums (per-cell log-sum-exp, or max with
 * argmax in t_maxs when max_scoring) and returns
 * t_sum = sum_j aggregate_j * cluster_sizes[j].  Replaces any committed table
 * (as assign_cells_to_nodes does at Inference.h:1320-1323).
 */
int scicone_compute_t_table(scicone_chain* ch, const scicone_tree* t, double* t_sum);
/*
 * The same full pass as a queued proposal: compiled now, launched by scicone_batch_submit (alone or together with other
 * chains), committed by scicone_accept.  with_od != 0 also evaluates the root score of compute_t_od_scores in that
 * launch (returned in od_scores[i] by scicone_batch_wait).  This is the form a cell-sharded run uses: the owner of the
 * chain compiles, every rank launches.
 */
int scicone_prepare_t_table(scicone_chain* ch, const scicone_tree* t, int32_t with_od);

/*
 * Inference::compute_t_od_scores (Inference.h:1406-1419) for the committed tree and
 * Inference::compute_t_prime_od_scores (Inference.h:1421-1433) for the proposal:
 * sum_j cluster_sizes[j] * Tree::get_od_root_score(cell j) (Tree.h:1774-1810) at `nu`.
 */
int scicone_compute_t_od_scores(scicone_chain* ch, double nu, double* od_score);
int scicone_compute_t_prime_od_scores(scicone_chain* ch, double nu, double* od_score);

/*
 * Inference::compute_t_prime_scores (Inference.h:1015-1052): rescore the subtree of
 * `t_prime` rooted at entry `node_idx`, relative to the committed table (the
 * parent's score is re-read from t_scores, Inference.h:1034).  May be called
 * several times before one compute_t_prime_sums (label swap, Inference.h:1059-1063).
 * The work is queued; it runs fused with the aggregate in compute_t_prime_sums.
 */
int scicone_compute_t_prime_scores(scicone_chain* ch, const scicone_tree* t_prime, int32_t node_idx);

/*
 * Inference::compute_t_prime_sums (Inference.h:1069-1196) followed by the weighted
 * sum over cells of Inference::comparison (Inference.h:292-294).  A node deleted by
 * the move is found as Inference::deleted_node_idx does (Inference.h:1270-1289):
 * the id present in the committed tree and absent from `t_prime`.
 * Sum scoring reproduces MathOp::log_replace_sum (MathOp.cpp:307-361), max scoring
 * the incremental argmax rule (Inference.h:1086-1152).
 */
int scicone_compute_t_prime_sums(scicone_chain* ch, const scicone_tree* t_prime, double* t_prime_sum);

/*
 * Optional, between the last scicone_compute_t_prime_scores and the sums call: build the launch description of the
 * queued proposal now, on the calling thread (Node::z re-derivation, host-side lgamma constants, row bookkeeping).
 * Thread-safe for different chains, so a multi-chain host spreads this work over its threads instead of leaving it
 * to the (serial) batch call.  No reference counterpart.
 */
int scicone_prepare_t_prime(scicone_chain* ch);

/*
 * Cell-sharded runs: a compiled proposal names every device row by its index in the row pool and nothing by address,
 * and the rows of a chain come from its owner rank's partition of the pool, so the compiled bytes are valid on every
 * rank.  The owner exports them after scicone_prepare_t_prime / scicone_prepare_t_table (*bytes receives the size; the
 * call fails with SCICONE_ERR_ARG if the buffer is smaller), ships them by its own means (the native host uses a
 * shared-memory mailbox), and the other ranks import them into their follower handle of the chain.  All ranks then
 * submit the same launches in the same order.  No reference counterpart (the reference is one process).
 */
int scicone_export_t_prime(scicone_chain* ch, void* buf, uint64_t capacity, uint64_t* bytes);
int scicone_import_t_prime(scicone_chain* ch, const void* buf, uint64_t bytes);

/*
 * The same two steps for `n` chains in ONE kernel launch (independent restarts or
 * proposals evaluated in lock-step).  Every chain must have queued its subtree
 * roots with scicone_compute_t_prime_scores.  All chains must share one dataset.
 * A proposal whose nu differs from the committed tree's (Inference::apply_overdispersion_change,
 * Inference.h:916-954) also gets its root score at the new nu from the same launch, returned in
 * od_scores[i] (NaN for the other proposals) - unless scicone_compute_t_prime_od_scores was already
 * called for that nu.
 */
int scicone_batch_compute_t_prime_sums(scicone_chain* const* chains, const scicone_tree* const* t_primes,
                                       int32_t n, double* t_prime_sums, double* od_scores /* [n] or NULL */);

/*
 * Pipelined form of the batch call: submit queues the copy-in and the kernels on the dataset's stream and
 * returns at once; the last CTA of the launch writes the totals and the ticket into pinned host memory, and
 * wait spins on that ticket (no device-to-host copy, no event) and returns the totals.
 * Accept / reject of the submitted proposals are host-side pointer swaps and may be issued before the
 * wait (later launches are ordered behind earlier ones by the stream).  At most 8 tickets may be
 * outstanding.  Reads (scicone_read_*) always see the finished state.
 */
int scicone_batch_submit(scicone_chain* const* chains, const scicone_tree* const* t_primes, int32_t n, uint64_t* ticket);
int scicone_batch_wait(scicone_dataset* ds, uint64_t ticket, double* t_prime_sums, double* od_scores /* [n] or NULL */);

/*
 * Batched forms of the per-chain calls for hosts that drive many chains (one call instead of n, run on the worker
 * pool): queue subtree root node_idx[i] of t_primes[i] on chains[i] (scicone_compute_t_prime_scores), and accept
 * (accept[i] != 0) or reject chains[i].  The chains must be distinct.
 */
int scicone_batch_compute_t_prime_scores(scicone_chain* const* chains, const scicone_tree* const* t_primes,
                                         const int32_t* node_idx, int32_t n);
int scicone_batch_settle(scicone_chain* const* chains, const int32_t* accept, int32_t n);

/*
 * Launch lanes of a dataset (default 1, at most 4).  A lane is a stream with its own device arena, table arena and
 * reduction slots; launches on different lanes overlap on the device.  One lane serialises launches, which is what allows
 * scicone_accept / scicone_reject before scicone_batch_wait; with several lanes a launch must be collected
 * (scicone_batch_wait) before its chains are settled or queued again, and at most one launch per lane may be outstanding.
 * On one GPU a launch takes any free lane and launches may be collected in any order; on a cell-sharded dataset launch k
 * uses lane k mod n on every rank, all ranks must submit their launches in the same order and collect them in order, and
 * the call is collective when the exchange runs over NCCL (every further lane gets a communicator of its own,
 * ncclCommSplit).
 */
int scicone_set_streams(scicone_dataset* ds, int32_t n);

/* Non-blocking probe of a ticket: *done = 1 when scicone_batch_wait would return at once. */
int scicone_batch_poll(scicone_dataset* ds, uint64_t ticket, int32_t* done);

/*
 * Accept (Inference.h:867-870: t_sums = t_prime_sums, t_maxs = t_prime_maxs,
 * update_t_scores(), and the proposal's od score / nu become the committed ones)
 * or reject (Inference.h:881,890) the queued proposal.  After a delete move the
 * deleted id leaves the table (Inference.h:1001-1009).
 */
int scicone_accept(scicone_chain* ch);
int scicone_reject(scicone_chain* ch);

/* Number of nodes in the committed table and their ids in ascending order. */
int scicone_t_n_nodes(scicone_chain* ch, int32_t* n_nodes);
int scicone_t_node_ids(scicone_chain* ch, int32_t* ids /* [n_nodes] */);

/*
 * Read-back of the committed tables (row-major [n_cells][n_nodes], node columns in
 * ascending id order = std::map iteration order; infer_trees.cpp:336-349).
 */
int scicone_read_t_scores(scicone_chain* ch, double* out);
int scicone_read_t_sums(scicone_chain* ch, double* out /* [n_cells] */);
int scicone_read_t_maxs(scicone_chain* ch, int32_t* ids /* [n_cells] */, double* vals /* [n_cells] */);
/* Primed tables of the queued proposal after compute_t_prime_sums. */
int scicone_read_t_prime_sums(scicone_chain* ch, double* out /* [n_cells] */);
int scicone_read_t_prime_maxs(scicone_chain* ch, int32_t* ids, double* vals);
/* t_prime_scores: rescored node ids (ascending) and their [n_cells][n_rescored] matrix. */
int scicone_t_prime_n_rescored(scicone_chain* ch, int32_t* n);
/* Relative device cost of the compiled (or imported) proposal of `ch`: about 2 per rescored node + 0.5 per unchanged node.
   A scheduler can use it to keep whole-tree rescorings (nu moves) out of the launches of the small proposals. */
int scicone_t_prime_cost(scicone_chain* ch, double* cost);
int scicone_read_t_prime_scores(scicone_chain* ch, int32_t* ids, double* out);

/*
 * Inference::assign_cells_to_nodes (Inference.h:1346-1362): per-cell argmax over the
 * committed table (std::max_element: first maximum in ascending id order) and the
 * cells x regions copy-number matrix neutral + c(argmax node).  `t` must be the
 * committed tree (its genotypes are expanded on the device).
 */
int scicone_assign_cells_to_nodes(scicone_chain* ch, const scicone_tree* t, int32_t* cell_node_ids /* [n_cells] */,
                                  int32_t* cell_region_cn /* [n_cells][n_regions] or NULL */);

/*
 * Utils::condense_matrix (Utils.cpp:104-138; the same collapse in pyscicone, scicone.py:276-299): cells x bins ->
 * cells x regions, out[i][r] = sum of the region_sizes[r] consecutive bins of region r, added in bin order (so the sums
 * are bit-identical to the reference's).  Bins beyond sum(region_sizes) are ignored, as in the reference.  Both
 * matrices are caller-owned host buffers; the work runs on `device` (TMA-staged streaming kernel, csrc/condense.cuh).
 * kernel_us (optional) receives the device time of the condensation kernel(s).
 */
int scicone_condense_matrix(int32_t device, const double* D_bins /* [n_cells][n_bins] */, int64_t n_cells, int64_t n_bins,
                            const int32_t* region_sizes, int32_t n_regions, double* D_regions /* [n_cells][n_regions] */,
                            double* kernel_us);

/*
 * Cluster condensation (reference scripts/condense_clusters.py:28-38; pyscicone scicone.py:331-335): means[c][k] = mean of
 * column k over the rows with assignments[i] == c, rows added in row order (numpy's axis-0 order, so the means equal the
 * reference's), sizes[c] = number of such rows.  assignments must lie in [0, n_clusters).  Host buffers, device work.
 */
int scicone_condense_clusters(int32_t device, const double* D /* [n_cells][n_cols] */, int64_t n_cells, int32_t n_cols,
                              const int32_t* assignments /* [n_cells] */, int32_t n_clusters,
                              double* means /* [n_clusters][n_cols] */, int32_t* sizes /* [n_clusters] */);

/*
 * Test hook: the device log-gamma used for every data and z term (reference: libm lgamma called from
 * Tree::compute_score, Tree.h:214-231) evaluated on `n` caller-provided positive arguments.
 */
int scicone_test_lgamma(int32_t device, const double* x, int32_t n, double* out);

/*
 * Host-side parallel loop on the library's worker pool: fn(i, ctx) for i in [0, n), returns when all are done.
 * No reference counterpart (the reference is one chain per process: global RNG singleton and `extern` tunables,
 * SURVEY.md section 8b).  The per-chain calls above - scicone_compute_t_prime_scores, scicone_accept,
 * scicone_reject, scicone_t_prime_n_rescored - may be issued concurrently for DIFFERENT chains of one dataset, so
 * a multi-chain host draws, queues and commits its chains inside this loop and launches from the calling thread;
 * scicone_batch_submit compiles its n proposals on the same pool.  Pool size: SCICONE_THREADS, else the cores (<= 16).
 */
int scicone_parallel_for(int32_t n, void (*fn)(int32_t index, void* ctx), void* ctx);
int scicone_host_threads(void);

/*
 * Multi-GPU (cells sharded over ranks).  The host side
 * exchanges the 128-byte NCCL unique id by its own means (the bench uses
 * torch.distributed), then every rank attaches its dataset.  Afterwards every
 * t_sum / od_score returned by the calls above is the sum over ALL ranks,
 * combined in a fixed order so that it does not depend on the number of GPUs.
 */
int scicone_comm_unique_id(uint8_t out[128]);
int scicone_dataset_attach_comm(scicone_dataset* ds, const uint8_t unique_id[128], int32_t rank, int32_t world_size);

/*
 * Measurement hooks for bench.py (no reference counterpart; the reference's only timer is the
 * wall clock around infer_mcmc, infer_trees.cpp:291-304).  With profiling on, every proposal
 * launch is bracketed by CUDA events on the dataset's stream.
 */
int scicone_set_profiling(scicone_dataset* ds, int32_t on);
int scicone_last_kernel_time_us(scicone_dataset* ds, double* us); /* device time of the last proposal kernel */
int scicone_kernel_time_stats(scicone_dataset* ds, double* sum_us, uint64_t* n_timed, int32_t reset);
int scicone_build_time_stats(scicone_dataset* ds, double* sum_us, uint64_t* n_timed); /* same for build_rows_kernel */
int scicone_kernel_launches(scicone_dataset* ds, uint64_t* n);
/* out = {bytes staged host->device, device->host, algorithmic bytes, cell x node scores} of all launches */
int scicone_counters(scicone_dataset* ds, double out[4], int32_t reset);
/* out = {algorithmic bytes, lgamma evaluations} of the table / increment kernels (K1a, K1b) of all launches: per rescored node
   of a tabulated proposal and cell 8 bytes written + 2 (count index) or 8 (count) read per event; per root-score slice
   and cell 8 written + 2 or 8 per region */
int scicone_build_counters(scicone_dataset* ds, double out[2], int32_t reset);
/* CUDA events on the dataset's stream bracketing a timed region (which = 0 start, 1 end). */
int scicone_region_mark(scicone_dataset* ds, int32_t which);
int scicone_region_elapsed_ms(scicone_dataset* ds, double* ms);
/* Device bytes in use: count matrix + score / lgamma rows handed out by the row pool. */
int scicone_device_bytes(scicone_dataset* ds, uint64_t* bytes);    /* kernels launched on this dataset so far */

#ifdef __cplusplus
}
#endif
#endif /* SCICONE_SCORE_H */
