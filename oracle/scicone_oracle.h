/*
 * scicone_oracle.h - CPU restatement of the reference's attachment-score path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.  The product (libscicone_score.so)
 * never links, loads or calls it and has no CPU fallback.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py)
 * against the reference's own known-answer vectors (scicone/tests/validation.h,
 * SURVEY.md section 8c / App. E) and against outputs of the unmodified reference
 * compiled here (oracle/Makefile -> oracle/_ref/{score,inference,ref_dump}),
 * committed as fixtures under tests/golden/.
 */
#ifndef SCICONE_ORACLE_H
#define SCICONE_ORACLE_H

#include <stdint.h>
#include "../include/scicone_score.h" /* only for the scicone_tree POD */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(int32_t n_cells, int32_t n_regions, const double* D, const int32_t* region_sizes,
                    const int32_t* neutral_states, const int32_t* cluster_sizes, double eta,
                    int32_t is_overdispersed, int32_t max_scoring);
void orc_destroy(orc_ctx* c);

/* Inference::compute_t_table, Inference.h:224-279 */
int orc_compute_t_table(orc_ctx* c, const scicone_tree* t, double* t_sum);
/* Inference::get_tree_od_root_scores + compute_t_od_scores, Inference.h:1389-1433 */
int orc_compute_od_scores(orc_ctx* c, double nu, double* od_score, double* per_cell /* [n_cells] or NULL */);
/* Inference::compute_t_prime_scores, Inference.h:1015-1052 */
int orc_compute_t_prime_scores(orc_ctx* c, const scicone_tree* t_prime, int32_t node_idx);
/* Inference::compute_t_prime_sums, Inference.h:1069-1196, + comparison's sum, :292-294 */
int orc_compute_t_prime_sums(orc_ctx* c, const scicone_tree* t_prime, double* t_prime_sum);
/* Inference.h:867-870 / :881,890 */
int orc_accept(orc_ctx* c, const scicone_tree* t_prime);
int orc_reject(orc_ctx* c);

int orc_t_n_nodes(orc_ctx* c);
void orc_t_node_ids(orc_ctx* c, int32_t* ids);
void orc_read_t_scores(orc_ctx* c, double* out); /* [n_cells][n_nodes], ids ascending */
void orc_read_t_sums(orc_ctx* c, double* out);
void orc_read_t_maxs(orc_ctx* c, int32_t* ids, double* vals);
void orc_read_t_prime_sums(orc_ctx* c, double* out);
void orc_read_t_prime_maxs(orc_ctx* c, int32_t* ids, double* vals);
int orc_t_prime_n_rescored(orc_ctx* c);
void orc_read_t_prime_scores(orc_ctx* c, int32_t* ids, double* out);
/* Inference::assign_cells_to_nodes, Inference.h:1346-1362 */
void orc_assign_cells_to_nodes(orc_ctx* c, const scicone_tree* t, int32_t* cell_node_ids, int32_t* cell_region_cn);

/* MathOp::log_sum (vector), MathOp.cpp:284-303, and log_replace_sum, :307-361; exposed for unit tests */
double orc_log_sum(const double* v, int n);
double orc_log_replace_sum(double sum, const double* to_subtract, int n_sub, const double* to_add, int n_add,
                           const double* unchanged, int n_unch);

/* Utils::condense_matrix, Utils.cpp:104-138: bins -> regions, one bin after the other (row-major in and out) */
int orc_condense_matrix(const double* D_bins, int64_t n_cells, int64_t n_bins, const int32_t* region_sizes, int32_t n_regions,
                        double* D_regions);

/* scripts/condense_clusters.py:28-38: per-cluster column means (rows added in row order) and sizes */
int orc_condense_clusters(const double* D, int64_t n_cells, int32_t n_cols, const int32_t* assignments, int32_t n_clusters,
                          double* means, int32_t* sizes);

#ifdef __cplusplus
}
#endif
#endif
