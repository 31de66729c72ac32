// cpu_backend.cpp - the C ABI of include/scicone_score.h implemented on top of the oracle port.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/cpu_backend/libscicone_score.so and put on
// LD_LIBRARY_PATH by tests/ so that the MCMC host (scicone_b200/host, a product binary that normally links the CUDA
// library) can be exercised in the GPU-less container: move logic, RNG consumption and file formats are compared
// with the unmodified reference `inference`.  Nothing in the product loads this library; scoring numbers produced
// through it are the oracle's, not the engine's, and are never reported as results.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "scicone_oracle.h"
#include "../scicone_b200/csrc/parallel_for.hpp"  // the same pool as the CUDA library, so the host's threading is exercised

struct scicone_dataset {
    int M, K;
    std::vector<double> D;
    std::vector<int32_t> r, neutral, w;
    double eta;
    int od, maxs;
    struct Result { std::vector<double> sums, ods; int rc; };
    std::map<uint64_t, Result> results;
    uint64_t next_ticket = 1;
};
struct scicone_chain {
    scicone_dataset* ds;
    orc_ctx* ctx;
    bool pending = false;
    bool full_pending = false;  // scicone_prepare_t_table: the queued "proposal" is a full pass
    int full_with_od = 0;
    std::vector<int32_t> roots;  // queued subtree roots (what an exported proposal carries)
    // copy of the queued tree
    std::vector<int32_t> id, parent, co, cr, cv, go, gr, gv;
    scicone_tree pod;
    double t_nu = 0.0;
    void keep(const scicone_tree* t) {
        int n = t->n_nodes;
        id.assign(t->node_id, t->node_id + n);
        parent.assign(t->parent, t->parent + n);
        co.assign(t->chg_off, t->chg_off + n + 1);
        cr.assign(t->chg_region, t->chg_region + co[n]); cr.push_back(0);
        cv.assign(t->chg_value, t->chg_value + co[n]); cv.push_back(0);
        go.assign(t->gen_off, t->gen_off + n + 1);
        gr.assign(t->gen_region, t->gen_region + go[n]); gr.push_back(0);
        gv.assign(t->gen_value, t->gen_value + go[n]); gv.push_back(0);
        pod.n_nodes = n; pod.node_id = id.data(); pod.parent = parent.data(); pod.chg_off = co.data(); pod.chg_region = cr.data();
        pod.chg_value = cv.data(); pod.gen_off = go.data(); pod.gen_region = gr.data(); pod.gen_value = gv.data(); pod.nu = t->nu;
    }
};
static thread_local std::string g_err;
static int fail(int rc, const char* m) { g_err = m; return rc; }

extern "C" {
int scicone_abi_version(void) { return SCICONE_ABI_VERSION; }
const char* scicone_last_error(void) { return g_err.c_str(); }
int scicone_dataset_create(scicone_dataset** out, const scicone_dataset_desc* d) {
    scicone_dataset* ds = new scicone_dataset();
    ds->M = d->n_cells; ds->K = d->n_regions;
    ds->D.assign(d->D, d->D + (size_t)ds->M * ds->K);
    ds->r.assign(d->region_sizes, d->region_sizes + ds->K);
    ds->neutral.assign(d->neutral_states, d->neutral_states + ds->K);
    ds->w.assign(d->cluster_sizes, d->cluster_sizes + ds->M);
    ds->eta = d->eta; ds->od = d->is_overdispersed; ds->maxs = d->max_scoring;
    *out = ds;
    return 0;
}
void scicone_dataset_destroy(scicone_dataset* ds) { delete ds; }
int scicone_chain_create(scicone_chain** out, scicone_dataset* ds) {
    scicone_chain* ch = new scicone_chain();
    ch->ds = ds;
    ch->ctx = orc_create(ds->M, ds->K, ds->D.data(), ds->r.data(), ds->neutral.data(), ds->w.data(), ds->eta, ds->od, ds->maxs);
    *out = ch;
    return 0;
}
void scicone_chain_destroy(scicone_chain* ch) { if (ch) { orc_destroy(ch->ctx); delete ch; } }
int scicone_compute_t_table(scicone_chain* ch, const scicone_tree* t, double* s) { ch->pending = false; ch->t_nu = t->nu; return orc_compute_t_table(ch->ctx, t, s); }
int scicone_compute_t_od_scores(scicone_chain* ch, double nu, double* od) { return orc_compute_od_scores(ch->ctx, nu, od, nullptr); }
int scicone_compute_t_prime_od_scores(scicone_chain* ch, double nu, double* od) { return orc_compute_od_scores(ch->ctx, nu, od, nullptr); }
int scicone_compute_t_prime_scores(scicone_chain* ch, const scicone_tree* t, int32_t idx) {
    if (!ch->pending) { ch->keep(t); ch->pending = true; ch->full_pending = false; ch->roots.clear(); }
    ch->roots.push_back(idx);
    return orc_compute_t_prime_scores(ch->ctx, &ch->pod, idx) ? fail(-3, "oracle: compute_t_prime_scores failed") : 0;
}
int scicone_chain_create_follower(scicone_chain** out, scicone_dataset* ds) { return scicone_chain_create(out, ds); }
int scicone_prepare_t_table(scicone_chain* ch, const scicone_tree* t, int32_t with_od) {
    ch->keep(t);
    ch->pending = true;
    ch->full_pending = true;
    ch->full_with_od = with_od;
    ch->roots.clear();
    return 0;
}
// "compiled proposal" of this backend: the queued tree and what to do with it (each process evaluates it on its own copy)
int scicone_export_t_prime(scicone_chain* ch, void* buf, uint64_t cap, uint64_t* bytes) {
    if (!ch->pending) return fail(-3, "export without a queued proposal");
    std::vector<int32_t> m;
    const int n = ch->pod.n_nodes;
    m.push_back(ch->full_pending); m.push_back(ch->full_with_od); m.push_back(n); m.push_back((int32_t)ch->roots.size());
    int64_t nu_bits; std::memcpy(&nu_bits, &ch->pod.nu, 8);
    m.push_back((int32_t)(nu_bits & 0xffffffff)); m.push_back((int32_t)(nu_bits >> 32));
    auto put = [&](const std::vector<int32_t>& v, size_t cnt) { m.insert(m.end(), v.begin(), v.begin() + cnt); };
    put(ch->id, n); put(ch->parent, n); put(ch->co, n + 1); put(ch->go, n + 1);
    put(ch->cr, ch->co[n]); put(ch->cv, ch->co[n]); put(ch->gr, ch->go[n]); put(ch->gv, ch->go[n]);
    put(ch->roots, ch->roots.size());
    *bytes = 4 * m.size();
    if (!buf || cap < *bytes) return fail(-1, "export: buffer too small");
    std::memcpy(buf, m.data(), *bytes);
    return 0;
}
int scicone_import_t_prime(scicone_chain* ch, const void* buf, uint64_t bytes) {
    const int32_t* m = static_cast<const int32_t*>(buf);
    if (bytes < 24) return fail(-1, "import: short message");
    std::vector<int32_t> v(m, m + bytes / 4);
    const int full = v[0], with_od = v[1], n = v[2], n_roots = v[3];
    int64_t nu_bits = (int64_t)(uint32_t)v[4] | ((int64_t)v[5] << 32);
    double nu; std::memcpy(&nu, &nu_bits, 8);
    size_t at = 6;
    auto take = [&](std::vector<int32_t>& dst, size_t cnt) { dst.assign(v.begin() + at, v.begin() + at + cnt); dst.push_back(0); at += cnt; };
    std::vector<int32_t> id, parent, co, go, cr, cv, gr, gv, roots;
    take(id, n); take(parent, n); take(co, n + 1); take(go, n + 1);
    take(cr, co[n]); take(cv, co[n]); take(gr, go[n]); take(gv, go[n]); take(roots, n_roots);
    scicone_tree t;
    t.n_nodes = n; t.node_id = id.data(); t.parent = parent.data(); t.chg_off = co.data(); t.chg_region = cr.data(); t.chg_value = cv.data();
    t.gen_off = go.data(); t.gen_region = gr.data(); t.gen_value = gv.data(); t.nu = nu;
    ch->pending = false;
    if (full) return scicone_prepare_t_table(ch, &t, with_od);
    for (int i = 0; i < n_roots; ++i)
        if (int rc = scicone_compute_t_prime_scores(ch, &t, roots[i])) return rc;
    return 0;
}
int scicone_dataset_reserve(scicone_dataset*, int32_t, int64_t) { return 0; }
int scicone_max_batch(scicone_dataset*, int32_t* n) { *n = 1 << 20; return 0; }
int scicone_batch_compute_t_prime_sums(scicone_chain* const* chains, const scicone_tree* const*, int32_t n, double* sums, double* ods) {
    for (int i = 0; i < n; ++i)
        if (!chains[i]->pending) return fail(-3, "compute_t_prime_sums without compute_t_prime_scores");
    struct Job { scicone_chain* const* chains; double* sums; double* ods; std::vector<int> rc; } job{chains, sums, ods, std::vector<int>(n, 0)};
    scicone::ForkJoinPool::instance().run(n, [](int i, void* p) {
        Job& j = *static_cast<Job*>(p);
        scicone_chain* ch = j.chains[i];
        if (ch->full_pending) {  // full pass queued by scicone_prepare_t_table (commits inside the oracle)
            if (orc_compute_t_table(ch->ctx, &ch->pod, &j.sums[i])) j.rc[i] = -4;
            if (j.ods) {
                j.ods[i] = NAN;
                if (ch->full_with_od) orc_compute_od_scores(ch->ctx, ch->pod.nu, &j.ods[i], nullptr);
            }
            return;
        }
        if (orc_compute_t_prime_sums(ch->ctx, &ch->pod, &j.sums[i])) j.rc[i] = -4;
        if (j.ods) {
            j.ods[i] = NAN;
            if (ch->pod.nu != ch->t_nu) orc_compute_od_scores(ch->ctx, ch->pod.nu, &j.ods[i], nullptr);
        }
    }, &job);
    int rc = 0;
    for (int r : job.rc) if (r) rc = r;
    return rc ? fail(rc, "NaN aggregate") : 0;
}
int scicone_compute_t_prime_sums(scicone_chain* ch, const scicone_tree* t, double* s) { return scicone_batch_compute_t_prime_sums(&ch, &t, 1, s, nullptr); }
int scicone_accept(scicone_chain* ch) {
    ch->pending = false;
    ch->t_nu = ch->pod.nu;
    if (ch->full_pending) { ch->full_pending = false; return 0; }
    return orc_accept(ch->ctx, &ch->pod);
}
int scicone_reject(scicone_chain* ch) { ch->pending = false; ch->full_pending = false; return orc_reject(ch->ctx); }
int scicone_t_n_nodes(scicone_chain* ch, int32_t* n) { *n = orc_t_n_nodes(ch->ctx); return 0; }
int scicone_t_node_ids(scicone_chain* ch, int32_t* ids) { orc_t_node_ids(ch->ctx, ids); return 0; }
int scicone_read_t_scores(scicone_chain* ch, double* out) { orc_read_t_scores(ch->ctx, out); return 0; }
int scicone_read_t_sums(scicone_chain* ch, double* out) { orc_read_t_sums(ch->ctx, out); return 0; }
int scicone_read_t_maxs(scicone_chain* ch, int32_t* ids, double* v) { orc_read_t_maxs(ch->ctx, ids, v); return 0; }
int scicone_read_t_prime_sums(scicone_chain* ch, double* out) { orc_read_t_prime_sums(ch->ctx, out); return 0; }
int scicone_read_t_prime_maxs(scicone_chain* ch, int32_t* ids, double* v) { orc_read_t_prime_maxs(ch->ctx, ids, v); return 0; }
int scicone_t_prime_n_rescored(scicone_chain* ch, int32_t* n) { *n = orc_t_prime_n_rescored(ch->ctx); return 0; }
int scicone_t_prime_cost(scicone_chain*, double* cost) { *cost = 1.0; return 0; }
int scicone_read_t_prime_scores(scicone_chain* ch, int32_t* ids, double* out) { orc_read_t_prime_scores(ch->ctx, ids, out); return 0; }
int scicone_assign_cells_to_nodes(scicone_chain* ch, const scicone_tree* t, int32_t* ids, int32_t* cn) { orc_assign_cells_to_nodes(ch->ctx, t, ids, cn); return 0; }
int scicone_comm_unique_id(uint8_t*) { return fail(-5, "cpu backend: no communicator"); }
int scicone_dataset_attach_comm(scicone_dataset*, const uint8_t*, int32_t, int32_t) { return fail(-5, "cpu backend: no communicator"); }
// "pipelined" form: evaluated at submit, handed out at wait (keeps the host's two-group schedule testable without a GPU)
int scicone_batch_submit(scicone_chain* const* chains, const scicone_tree* const* tp, int32_t n, uint64_t* ticket) {
    scicone_dataset* ds = chains[0]->ds;
    scicone_dataset::Result r;
    r.sums.assign(n, 0.0);
    r.ods.assign(n, NAN);
    r.rc = scicone_batch_compute_t_prime_sums(chains, tp, n, r.sums.data(), r.ods.data());
    if (r.rc && r.rc != -4) return r.rc;
    *ticket = ds->next_ticket++;
    ds->results[*ticket] = r;
    return 0;
}
int scicone_batch_wait(scicone_dataset* ds, uint64_t ticket, double* sums, double* ods) {
    auto it = ds->results.find(ticket);
    if (it == ds->results.end()) return fail(-3, "unknown ticket");
    for (size_t i = 0; i < it->second.sums.size(); ++i) {
        sums[i] = it->second.sums[i];
        if (ods) ods[i] = it->second.ods[i];
    }
    int rc = it->second.rc;
    ds->results.erase(it);
    return rc ? fail(rc, "NaN aggregate") : 0;
}
int scicone_set_profiling(scicone_dataset*, int32_t) { return 0; }
int scicone_last_kernel_time_us(scicone_dataset*, double* us) { *us = 0; return 0; }
int scicone_kernel_time_stats(scicone_dataset*, double* s, uint64_t* n, int32_t) { if (s) *s = 0; if (n) *n = 0; return 0; }
int scicone_build_time_stats(scicone_dataset*, double* s, uint64_t* n) { if (s) *s = 0; if (n) *n = 0; return 0; }
int scicone_build_counters(scicone_dataset*, double out[2], int32_t) { out[0] = out[1] = 0; return 0; }
int scicone_kernel_launches(scicone_dataset*, uint64_t* n) { *n = 0; return 0; }
int scicone_parallel_for(int32_t n, void (*fn)(int32_t, void*), void* ctx) { scicone::ForkJoinPool::instance().run(n, fn, ctx); return 0; }
int scicone_test_lgamma(int32_t, const double*, int32_t, double*) { return fail(-2, "cpu backend: no device lgamma"); }
int scicone_prepare_t_prime(scicone_chain*) { return 0; }
int scicone_batch_compute_t_prime_scores(scicone_chain* const* chains, const scicone_tree* const* t, const int32_t* idx, int32_t n) {
    int rc = 0;
    for (int i = 0; i < n; ++i) if (int r = scicone_compute_t_prime_scores(chains[i], t[i], idx[i])) rc = r;
    return rc;
}
int scicone_batch_settle(scicone_chain* const* chains, const int32_t* accept, int32_t n) {
    int rc = 0;
    for (int i = 0; i < n; ++i) if (int r = accept[i] ? scicone_accept(chains[i]) : scicone_reject(chains[i])) rc = r;
    return rc;
}
int scicone_condense_matrix(int32_t, const double* bins, int64_t n_cells, int64_t n_bins, const int32_t* sizes, int32_t n_regions, double* out, double* us) {
    if (us) *us = 0.0;
    return orc_condense_matrix(bins, n_cells, n_bins, sizes, n_regions, out) ? fail(-1, "oracle: condense_matrix failed") : 0;
}
int scicone_batch_poll(scicone_dataset* ds, uint64_t ticket, int32_t* done) { *done = ds->results.count(ticket) ? 1 : 0; return *done ? 0 : fail(-3, "unknown ticket"); }
int scicone_set_streams(scicone_dataset*, int32_t) { return 0; }
int scicone_condense_clusters(int32_t, const double*, int64_t, int32_t, const int32_t*, int32_t, double*, int32_t*) { return fail(-2, "cpu backend: no device"); }
int scicone_host_threads(void) { return scicone::ForkJoinPool::instance().n_threads(); }
int scicone_counters(scicone_dataset*, double out[4], int32_t) { out[0] = out[1] = out[2] = out[3] = 0; return 0; }
int scicone_region_mark(scicone_dataset*, int32_t) { return 0; }
int scicone_region_elapsed_ms(scicone_dataset*, double* ms) { *ms = 0; return 0; }
int scicone_device_bytes(scicone_dataset*, uint64_t* b) { *b = 0; return 0; }
}
