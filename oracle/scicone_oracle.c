/*
 * scicone_oracle.c - plain-C restatement of SCICoNE's attachment-score path.
 *
 * TEST INFRASTRUCTURE ONLY (see scicone_oracle.h).  Parity status: PINNED against
 * the reference's known-answer vectors and against the unmodified reference built
 * in oracle/_ref (tests/test_oracle_*.py, fixtures in tests/golden/).
 *
 * Every function names the reference lines it follows (paths relative to
 * /root/reference).  The arithmetic order is the reference's: no reassociation,
 * compiled with -ffp-contract=off, libm lgamma/exp/log as the reference uses.
 * Score tables are dense arrays indexed by node id standing in for the
 * reference's std::map<int,double> per cell (Inference.h:39-44); "ascending id"
 * loops reproduce std::map iteration order.
 */
#include "scicone_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_ctx {
    int M, K;
    double* D;    /* [M][K] */
    double* sumD; /* [M] std::accumulate left to right (Tree.h:253, Inference.h:1031) */
    int *r, *neutral, *w;
    double eta;
    int od, maxs;
    /* tables indexed by node id */
    int cap;
    char *t_has, *p_has;
    double **t_sc, **p_sc; /* [id] -> [M] */
    int *t_z, *p_z;        /* Node::z (int!) of t and of t_prime */
    char* p_zset;
    double *t_sums, *p_sums;
    int *t_max_id, *p_max_id;
    double *t_max_val, *p_max_val;
    int p_deleted;
};

static void grow(orc_ctx* c, int id) {
    if (id < c->cap) return;
    int ncap = c->cap ? c->cap : 16;
    while (ncap <= id) ncap *= 2;
    c->t_has = realloc(c->t_has, ncap);
    c->p_has = realloc(c->p_has, ncap);
    c->p_zset = realloc(c->p_zset, ncap);
    c->t_sc = realloc(c->t_sc, sizeof(double*) * ncap);
    c->p_sc = realloc(c->p_sc, sizeof(double*) * ncap);
    c->t_z = realloc(c->t_z, sizeof(int) * ncap);
    c->p_z = realloc(c->p_z, sizeof(int) * ncap);
    for (int i = c->cap; i < ncap; ++i) {
        c->t_has[i] = c->p_has[i] = c->p_zset[i] = 0;
        c->t_sc[i] = c->p_sc[i] = NULL;
        c->t_z[i] = c->p_z[i] = 0;
    }
    c->cap = ncap;
}

orc_ctx* orc_create(int32_t M, int32_t K, const double* D, const int32_t* r, const int32_t* neutral,
                    const int32_t* w, double eta, int32_t od, int32_t maxs) {
    orc_ctx* c = calloc(1, sizeof(orc_ctx));
    c->M = M;
    c->K = K;
    c->D = malloc(sizeof(double) * (size_t)M * K);
    memcpy(c->D, D, sizeof(double) * (size_t)M * K);
    c->sumD = malloc(sizeof(double) * M);
    for (int j = 0; j < M; ++j) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s = s + D[(size_t)j * K + k];
        c->sumD[j] = s;
    }
    c->r = malloc(sizeof(int) * K);
    c->neutral = malloc(sizeof(int) * K);
    c->w = malloc(sizeof(int) * M);
    memcpy(c->r, r, sizeof(int) * K);
    memcpy(c->neutral, neutral, sizeof(int) * K);
    memcpy(c->w, w, sizeof(int) * M);
    c->eta = eta;
    c->od = od;
    c->maxs = maxs;
    c->t_sums = calloc(M, sizeof(double));
    c->p_sums = calloc(M, sizeof(double));
    c->t_max_id = calloc(M, sizeof(int));
    c->p_max_id = calloc(M, sizeof(int));
    c->t_max_val = calloc(M, sizeof(double));
    c->p_max_val = calloc(M, sizeof(double));
    c->p_deleted = -1;
    grow(c, 15);
    return c;
}

void orc_destroy(orc_ctx* c) {
    if (!c) return;
    for (int i = 0; i < c->cap; ++i) {
        free(c->t_sc[i]);
        free(c->p_sc[i]);
    }
    free(c->t_sc); free(c->p_sc); free(c->t_has); free(c->p_has); free(c->p_zset);
    free(c->t_z); free(c->p_z); free(c->D); free(c->sumD); free(c->r); free(c->neutral); free(c->w);
    free(c->t_sums); free(c->p_sums); free(c->t_max_id); free(c->p_max_id); free(c->t_max_val); free(c->p_max_val);
    free(c);
}

/* node->c.count(k) ? node->c[k] : 0   (Tree.h:189-190) */
static int genotype_at(const scicone_tree* t, int n, int k) {
    for (int e = t->gen_off[n]; e < t->gen_off[n + 1]; ++e)
        if (t->gen_region[e] == k) return t->gen_value[e];
    return 0;
}

/* MathOp::log_sum(const vector<double>&), MathOp.cpp:284-303 (the map overload :262-282 is the same loop) */
double orc_log_sum(const double* v, int n) {
    double max = -DBL_MAX; /* numeric_limits<double>::lowest() */
    for (int i = 0; i < n; ++i)
        if (v[i] > max) max = v[i];
    double sum_in_normal_space = 0.0;
    for (int i = 0; i < n; ++i) sum_in_normal_space += exp(v[i] - max);
    return log(sum_in_normal_space) + max;
}

/* MathOp::log_replace_sum, MathOp.cpp:307-361 */
double orc_log_replace_sum(double sum, const double* to_subtract, int n_sub, const double* to_add, int n_add,
                           const double* unchanged, int n_unch) {
    double res;
    double* buf = malloc(sizeof(double) * (size_t)(n_add + n_unch + 1));
    if (n_add >= n_unch) {
        memcpy(buf, to_add, sizeof(double) * n_add);
        memcpy(buf + n_add, unchanged, sizeof(double) * n_unch);
        res = orc_log_sum(buf, n_add + n_unch);
    } else {
        double max = sum;
        for (int i = 0; i < n_sub; ++i)
            if (to_subtract[i] > max) max = to_subtract[i];
        double sum_in_normal_space = 0.0;
        sum_in_normal_space += exp(sum - max);
        for (int i = 0; i < n_sub; ++i) {
            double temp = exp(to_subtract[i] - max);
            sum_in_normal_space -= temp;
        }
        if (sum_in_normal_space <= 1e-6) {
            memcpy(buf, to_add, sizeof(double) * n_add);
            memcpy(buf + n_add, unchanged, sizeof(double) * n_unch);
            res = orc_log_sum(buf, n_add + n_unch);
        } else {
            double subtracted_res = log(sum_in_normal_space) + max;
            memcpy(buf, to_add, sizeof(double) * n_add);
            buf[n_add] = subtracted_res;
            res = orc_log_sum(buf, n_add + 1);
        }
    }
    free(buf);
    return res;
}

/*
 * Tree::compute_score for one node and one cell, Tree.h:163-244.
 * parent_score / parent_z are node->parent->attachment_score and node->parent->z
 * (int).  Returns the node's attachment score and stores the truncated z.
 */
static double compute_score(const orc_ctx* c, const scicone_tree* t, int n, const double* Dj, double sum_D,
                            double parent_score, int parent_z, int* z_out) {
    const int p_idx = t->parent[n];
    const double nu = t->nu;
    double log_likelihood = parent_score;
    double z = parent_z;
    double z_parent = parent_z;
    for (int e = t->chg_off[n]; e < t->chg_off[n + 1]; ++e) { /* Tree.h:187-198 */
        int k = t->chg_region[e];
        int cf = genotype_at(t, n, k);
        int cp_f = genotype_at(t, p_idx, k);
        int p = c->neutral[k];
        double node_cn = (cf + p) == 0 ? c->eta : (double)(cf + p);
        double parent_cn = (cp_f + p) == 0 ? c->eta : (double)(cp_f + p);
        z += c->r[k] * (node_cn - parent_cn);
    }
    for (int e = t->chg_off[n]; e < t->chg_off[n + 1]; ++e) { /* Tree.h:200-222 */
        int k = t->chg_region[e];
        int cf = genotype_at(t, n, k);
        int cp_f = genotype_at(t, p_idx, k);
        int p = c->neutral[k];
        double node_cn = (cf + p) == 0 ? c->eta : (double)(cf + p);
        double parent_cn = (cp_f + p) == 0 ? c->eta : (double)(cp_f + p);
        if (c->od) {
            log_likelihood += lgamma(Dj[k] + nu * (node_cn)*c->r[k]);
            log_likelihood -= lgamma(Dj[k] + nu * (parent_cn)*c->r[k]);
            log_likelihood -= lgamma(nu * (node_cn)*c->r[k]);
            log_likelihood += lgamma(nu * (parent_cn)*c->r[k]);
        } else
            log_likelihood += Dj[k] * (log(node_cn) - log(parent_cn));
    }
    if (c->od) { /* Tree.h:224-232 */
        log_likelihood += lgamma(nu * z);
        log_likelihood -= lgamma(nu * z_parent);
        log_likelihood -= lgamma(sum_D + nu * z);
        log_likelihood += lgamma(sum_D + nu * z_parent);
    } else { /* Tree.h:233-238; node->parent->z is the int */
        log_likelihood -= sum_D * log(z);
        log_likelihood += sum_D * log((double)parent_z);
    }
    *z_out = (int)z; /* Tree.h:241 with Node::z an int (Node.h:25) */
    return log_likelihood;
}

/* Tree::compute_root_score, Tree.h:145-160 */
static int root_z(const orc_ctx* c) {
    int z = 0;
    for (int i = 0; i < c->K; ++i) z += c->r[i] * c->neutral[i];
    return z;
}

static double* row(double** tab, int id, int M) {
    if (!tab[id]) tab[id] = malloc(sizeof(double) * M);
    return tab[id];
}

static int max_id_of(const scicone_tree* t) {
    int m = 0;
    for (int i = 0; i < t->n_nodes; ++i)
        if (t->node_id[i] > m) m = t->node_id[i];
    return m;
}

/* Inference::compute_t_table, Inference.h:224-279 (+ Tree::compute_tree/compute_stack, Tree.h:247-286) */
int orc_compute_t_table(orc_ctx* c, const scicone_tree* t, double* t_sum_out) {
    grow(c, max_id_of(t));
    for (int i = 0; i < c->cap; ++i) c->t_has[i] = 0;
    double* sc = malloc(sizeof(double) * t->n_nodes);
    int* zz = malloc(sizeof(int) * t->n_nodes);
    /* ascending-id order of the nodes (std::map iteration order) */
    int* order = malloc(sizeof(int) * t->n_nodes);
    for (int i = 0; i < t->n_nodes; ++i) order[i] = i;
    for (int a = 1; a < t->n_nodes; ++a) { /* insertion sort by id */
        int v = order[a], b = a - 1;
        while (b >= 0 && t->node_id[order[b]] > t->node_id[v]) { order[b + 1] = order[b]; --b; }
        order[b + 1] = v;
    }
    double* vals = malloc(sizeof(double) * t->n_nodes);
    double t_sum = 0.0;
    for (int j = 0; j < c->M; ++j) {
        const double* Dj = c->D + (size_t)j * c->K;
        double sum_d = c->sumD[j];
        for (int n = 0; n < t->n_nodes; ++n) { /* parent-before-child == any stack DFS order */
            if (t->parent[n] < 0) {
                sc[n] = 0.0;
                zz[n] = root_z(c);
            } else
                sc[n] = compute_score(c, t, n, Dj, sum_d, sc[t->parent[n]], zz[t->parent[n]], &zz[n]);
        }
        for (int n = 0; n < t->n_nodes; ++n) {
            int id = t->node_id[n];
            row(c->t_sc, id, c->M)[j] = sc[n];
            c->t_has[id] = 1;
            c->t_z[id] = zz[n];
        }
        double aux;
        if (c->maxs) { /* Inference.h:243-255 */
            double currentMax = -DBL_MAX;
            int arg = 0;
            for (int a = 0; a < t->n_nodes; ++a) {
                double v = sc[order[a]];
                if (v > currentMax) { arg = t->node_id[order[a]]; currentMax = v; }
            }
            aux = currentMax;
            c->t_max_id[j] = arg;
            c->t_max_val[j] = currentMax;
        } else {
            for (int a = 0; a < t->n_nodes; ++a) vals[a] = sc[order[a]];
            aux = orc_log_sum(vals, t->n_nodes);
        }
        t_sum = t_sum + aux * c->w[j]; /* Inference.h:269 */
        c->t_sums[j] = aux;
    }
    free(sc); free(zz); free(order); free(vals);
    orc_reject(c);
    if (t_sum_out) *t_sum_out = t_sum;
    return 0;
}

/* Tree::get_od_root_score (Tree.h:1774-1810) per cell, weighted sum of Inference.h:1412-1415 */
int orc_compute_od_scores(orc_ctx* c, double nu, double* od_score, double* per_cell) {
    double total = 0.;
    for (int j = 0; j < c->M; ++j) {
        const double* Dj = c->D + (size_t)j * c->K;
        double sum_D = c->sumD[j];
        double od_root_score = 0.0;
        if (c->od) {
            int z = root_z(c);
            od_root_score += lgamma(nu * z);
            od_root_score -= lgamma(sum_D + nu * z);
            for (int i = 0; i < c->K; ++i) {
                od_root_score += lgamma(Dj[i] + nu * c->neutral[i] * c->r[i]);
                od_root_score -= lgamma(nu * c->neutral[i] * c->r[i]);
            }
        } else {
            int sum_r = 0;
            for (int i = 0; i < c->K; ++i) sum_r += c->r[i];
            double term1 = -sum_D * log((double)sum_r);
            double term2 = 0.0;
            for (int i = 0; i < c->K; ++i) term2 += Dj[i] * log((double)c->r[i]);
            od_root_score += term1 + term2;
        }
        if (per_cell) per_cell[j] = od_root_score;
        total = total + od_root_score * c->w[j];
    }
    if (od_score) *od_score = total;
    return 0;
}

/* Inference::compute_t_prime_scores, Inference.h:1015-1052 */
int orc_compute_t_prime_scores(orc_ctx* c, const scicone_tree* t, int32_t node_idx) {
    if (node_idx < 0 || node_idx >= t->n_nodes) return -1;
    grow(c, max_id_of(t));
    char* in_sub = calloc(t->n_nodes, 1);
    in_sub[node_idx] = 1;
    for (int n = node_idx + 1; n < t->n_nodes; ++n)
        if (t->parent[n] >= 0 && in_sub[t->parent[n]]) in_sub[n] = 1;
    double* sc = calloc(t->n_nodes, sizeof(double));
    int* zz = calloc(t->n_nodes, sizeof(int));
    int rc = 0;
    const int par = t->parent[node_idx];
    for (int j = 0; j < c->M && rc == 0; ++j) {
        const double* Dj = c->D + (size_t)j * c->K;
        double sum_d = c->sumD[j];
        if (par >= 0) {
            int pid = t->node_id[par];
            /* attached_node->parent->attachment_score = t_scores[j][parent->id], Inference.h:1034 */
            if (!c->t_has[pid]) { rc = -3; break; } /* the reference would default-insert 0.0 here */
            sc[par] = c->t_sc[pid][j];
            zz[par] = c->p_zset[pid] ? c->p_z[pid] : c->t_z[pid]; /* Node::z travels with the tree copy */
        }
        for (int n = node_idx; n < t->n_nodes; ++n) {
            if (!in_sub[n]) continue;
            if (t->parent[n] < 0) {
                sc[n] = 0.0;
                zz[n] = root_z(c);
            } else
                sc[n] = compute_score(c, t, n, Dj, sum_d, sc[t->parent[n]], zz[t->parent[n]], &zz[n]);
            row(c->p_sc, t->node_id[n], c->M)[j] = sc[n];
        }
    }
    if (rc == 0)
        for (int n = node_idx; n < t->n_nodes; ++n)
            if (in_sub[n]) {
                int id = t->node_id[n];
                c->p_has[id] = 1;
                c->p_z[id] = zz[n];
                c->p_zset[id] = 1;
            }
    free(in_sub); free(sc); free(zz);
    return rc;
}

/* Inference::deleted_node_idx, Inference.h:1270-1289 */
static int deleted_node_idx(const orc_ctx* c, const scicone_tree* tp) {
    int n_t = 0;
    for (int i = 0; i < c->cap; ++i) n_t += c->t_has[i];
    if (tp->n_nodes >= n_t) return -1;
    int deleted = -1;
    for (int id = 0; id < c->cap; ++id) {
        if (!c->t_has[id]) continue;
        int found = 0;
        for (int n = 0; n < tp->n_nodes; ++n)
            if (tp->node_id[n] == id) { found = 1; break; }
        if (!found) deleted = id;
    }
    return deleted;
}

/* Inference::compute_t_prime_sums, Inference.h:1069-1196, then the sum of Inference.h:292-294 */
int orc_compute_t_prime_sums(orc_ctx* c, const scicone_tree* tp, double* t_prime_sum) {
    const int deleted_index = deleted_node_idx(c, tp);
    c->p_deleted = deleted_index;
    double* old_vals = malloc(sizeof(double) * (c->cap + 1));
    double* new_vals = malloc(sizeof(double) * (c->cap + 1));
    double* unch = malloc(sizeof(double) * (c->cap + 1));
    int rc = 0;
    for (int i = 0; i < c->M; ++i) {
        if (c->maxs) { /* Inference.h:1086-1152 */
            double currentMax_newval = -DBL_MAX;
            int new_first = 0;
            int prev_max_in_new = 0;
            for (int id = 0; id < c->cap; ++id) {
                if (!c->p_has[id]) continue;
                if (c->t_max_id[i] == id) prev_max_in_new = 1;
                if (c->p_sc[id][i] > currentMax_newval) {
                    currentMax_newval = c->p_sc[id][i];
                    new_first = id;
                }
            }
            int prev_first = c->t_max_id[i];
            double prev_second = c->t_max_val[i];
            if (prev_max_in_new || c->t_max_id[i] == deleted_index) {
                double currentMax = -DBL_MAX;
                int u_first = 0;       /* value-initialised std::pair<int,double> */
                double u_second = 0.0;
                for (int id = 0; id < c->cap; ++id) {
                    if (!c->t_has[id] || c->p_has[id] || id == deleted_index) continue; /* unchanged_vals */
                    if (c->t_sc[id][i] > currentMax) {
                        u_first = id;
                        u_second = c->t_sc[id][i];
                        currentMax = u_second;
                    }
                }
                prev_first = u_first;
                prev_second = u_second;
            }
            int nm_first = new_first;
            double nm_second = currentMax_newval;
            if (prev_second > currentMax_newval) {
                nm_first = prev_first;
                nm_second = prev_second;
            }
            c->p_max_id[i] = nm_first;
            c->p_max_val[i] = nm_second;
            if (isnan(nm_second)) rc = -4;
            c->p_sums[i] = nm_second;
        } else { /* Inference.h:1157-1194 */
            int n_old = 0, n_new = 0, n_unch = 0;
            for (int id = 0; id < c->cap; ++id) {
                if (!c->p_has[id]) continue;
                if (c->t_has[id]) old_vals[n_old++] = c->t_sc[id][i];
                new_vals[n_new++] = c->p_sc[id][i];
            }
            if (deleted_index != -1) old_vals[n_old++] = c->t_sc[deleted_index][i];
            for (int id = 0; id < c->cap; ++id)
                if (c->t_has[id] && !c->p_has[id] && id != deleted_index) unch[n_unch++] = c->t_sc[id][i];
            double res = orc_log_replace_sum(c->t_sums[i], old_vals, n_old, new_vals, n_new, unch, n_unch);
            if (isnan(res)) rc = -4;
            c->p_sums[i] = res;
        }
    }
    free(old_vals); free(new_vals); free(unch);
    double s = 0.;
    for (int i = 0; i < c->M; ++i) s = s + c->p_sums[i] * c->w[i]; /* Inference.h:292-294 */
    if (t_prime_sum) *t_prime_sum = s;
    return rc;
}

/* accept branch of infer_mcmc, Inference.h:867-870 (update_t_scores :996-1012; t = t_prime copies Node::z) */
int orc_accept(orc_ctx* c, const scicone_tree* tp) {
    (void)tp;
    memcpy(c->t_sums, c->p_sums, sizeof(double) * c->M);
    memcpy(c->t_max_id, c->p_max_id, sizeof(int) * c->M);
    memcpy(c->t_max_val, c->p_max_val, sizeof(double) * c->M);
    for (int id = 0; id < c->cap; ++id) {
        if (!c->p_has[id]) continue;
        memcpy(row(c->t_sc, id, c->M), c->p_sc[id], sizeof(double) * c->M);
        c->t_has[id] = 1;
        c->t_z[id] = c->p_z[id];
    }
    if (c->p_deleted != -1) c->t_has[c->p_deleted] = 0;
    return orc_reject(c);
}

/* reject branch, Inference.h:881 (t_prime = t) and :890 (t_prime_scores.clear()) */
int orc_reject(orc_ctx* c) {
    for (int id = 0; id < c->cap; ++id) c->p_has[id] = c->p_zset[id] = 0;
    c->p_deleted = -1;
    return 0;
}

int orc_t_n_nodes(orc_ctx* c) {
    int n = 0;
    for (int i = 0; i < c->cap; ++i) n += c->t_has[i];
    return n;
}
void orc_t_node_ids(orc_ctx* c, int32_t* ids) {
    int n = 0;
    for (int i = 0; i < c->cap; ++i)
        if (c->t_has[i]) ids[n++] = i;
}
void orc_read_t_scores(orc_ctx* c, double* out) {
    int nn = orc_t_n_nodes(c), col = 0;
    for (int id = 0; id < c->cap; ++id) {
        if (!c->t_has[id]) continue;
        for (int j = 0; j < c->M; ++j) out[(size_t)j * nn + col] = c->t_sc[id][j];
        ++col;
    }
}
void orc_read_t_sums(orc_ctx* c, double* out) { memcpy(out, c->t_sums, sizeof(double) * c->M); }
void orc_read_t_maxs(orc_ctx* c, int32_t* ids, double* vals) {
    memcpy(ids, c->t_max_id, sizeof(int) * c->M);
    memcpy(vals, c->t_max_val, sizeof(double) * c->M);
}
void orc_read_t_prime_sums(orc_ctx* c, double* out) { memcpy(out, c->p_sums, sizeof(double) * c->M); }
void orc_read_t_prime_maxs(orc_ctx* c, int32_t* ids, double* vals) {
    memcpy(ids, c->p_max_id, sizeof(int) * c->M);
    memcpy(vals, c->p_max_val, sizeof(double) * c->M);
}
int orc_t_prime_n_rescored(orc_ctx* c) {
    int n = 0;
    for (int i = 0; i < c->cap; ++i) n += c->p_has[i];
    return n;
}
void orc_read_t_prime_scores(orc_ctx* c, int32_t* ids, double* out) {
    int nn = orc_t_prime_n_rescored(c), col = 0;
    for (int id = 0; id < c->cap; ++id) {
        if (!c->p_has[id]) continue;
        if (ids) ids[col] = id;
        for (int j = 0; j < c->M; ++j) out[(size_t)j * nn + col] = c->p_sc[id][j];
        ++col;
    }
}

/* Inference::assign_cells_to_nodes, Inference.h:1341-1362 (the full pass before it is orc_compute_t_table) */
void orc_assign_cells_to_nodes(orc_ctx* c, const scicone_tree* t, int32_t* cell_node_ids, int32_t* cell_region_cn) {
    for (int j = 0; j < c->M; ++j) {
        int best = -1;
        double bv = 0.0;
        for (int id = 0; id < c->cap; ++id) { /* std::max_element: first maximum in ascending id order */
            if (!c->t_has[id]) continue;
            if (best < 0 || bv < c->t_sc[id][j]) { best = id; bv = c->t_sc[id][j]; }
        }
        if (cell_node_ids) cell_node_ids[j] = best;
        if (cell_region_cn) {
            int32_t* out = cell_region_cn + (size_t)j * c->K;
            for (int k = 0; k < c->K; ++k) out[k] = c->neutral[k];
            for (int n = 0; n < t->n_nodes; ++n)
                if (t->node_id[n] == best)
                    for (int e = t->gen_off[n]; e < t->gen_off[n + 1]; ++e)
                        out[t->gen_region[e]] = t->gen_value[e] + c->neutral[t->gen_region[e]];
        }
    }
}

/* Utils::condense_matrix (reference scicone/src/Utils.cpp:104-138): walk the first sum(region_sizes) bins of every row and
 * add each to the current region's cell, moving to the next region when region_sizes[region] bins have been taken. */
int orc_condense_matrix(const double* D_bins, int64_t n_cells, int64_t n_bins, const int32_t* region_sizes, int32_t n_regions,
                        double* D_regions) {
    int64_t total = 0;
    for (int32_t r = 0; r < n_regions; ++r) total += region_sizes[r];
    if (total > n_bins) return -1;
    for (int64_t i = 0; i < n_cells; ++i) {
        double* out = D_regions + i * n_regions;
        const double* in = D_bins + i * n_bins;
        for (int32_t r = 0; r < n_regions; ++r) out[r] = 0.0;
        int32_t region = 0, taken = 0;
        for (int64_t j = 0; j < total; ++j) {
            out[region] += in[j];
            if (++taken == region_sizes[region]) {
                ++region;
                taken = 0;
            }
        }
    }
    return 0;
}

/* Cluster condensation (reference scripts/condense_clusters.py:28-38: np.mean(input_data[rows of the cluster], axis=0), i.e.
 * the cluster's rows added one after the other in row order, divided by their number; pyscicone scicone.py:331-335). */
int orc_condense_clusters(const double* D, int64_t n_cells, int32_t n_cols, const int32_t* assignments, int32_t n_clusters,
                          double* means, int32_t* sizes) {
    for (int32_t c = 0; c < n_clusters; ++c) {
        sizes[c] = 0;
        for (int32_t k = 0; k < n_cols; ++k) means[(int64_t)c * n_cols + k] = 0.0;
    }
    for (int64_t i = 0; i < n_cells; ++i) {
        const int32_t c = assignments[i];
        if (c < 0 || c >= n_clusters) return -1;
        sizes[c]++;
        for (int32_t k = 0; k < n_cols; ++k) means[(int64_t)c * n_cols + k] += D[i * n_cols + k];
    }
    for (int32_t c = 0; c < n_clusters; ++c)
        for (int32_t k = 0; k < n_cols; ++k)
            if (sizes[c] > 0) means[(int64_t)c * n_cols + k] /= (double)sizes[c];
    return 0;
}
