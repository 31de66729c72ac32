// ref_dump - golden-vector generator.  TEST INFRASTRUCTURE ONLY.
//
// This driver is ours; everything it calls is the UNMODIFIED reference
// (Inference / Tree / Node / MathOp from /root/reference/scicone/src, compiled by
// oracle/Makefile with the RNG/NLopt shims).  It sequences the reference's public
// methods the way Inference::infer_mcmc does (Inference.h:544-913) but with ONE
// attempt per move, and dumps at full precision (%.17g) everything the hot path
// produces, as JSON, so that oracle/scicone_oracle.c and the CUDA engine can be
// replayed against it:
//   init   : tree t (flat), t_scores, t_sums, t_maxs, total, od_score
//   moves[]: move id, tree t_prime (flat), rescored ids, t_prime_scores,
//            t_prime_sums, t_prime_maxs, total (and od_score for move 13),
//            accepted flag from Inference::comparison.
//
// Usage: ref_dump --d_matrix_file F --region_sizes_file F --n_cells M --n_regions K
//        [--cluster_sizes_file F] [--tree_file F | --n_nodes N | --worked_example 1]
//        [--nu X] [--is_overdispersed 0|1] [--max_scoring 0|1] [--seed S]
//        [--n_moves N] [--move_probs p0,...,p14] [--cf X] [--eta X] > out.json
#include <cstdio>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "MathOp.h"
#include "Tree.h"
#include "Inference.h"
#include <cxxopts.hpp>
#include "globals.cpp"

int print_precision;
int copy_number_limit;
double lambda_s;
double lambda_r;
double lambda_c;
double cf;
double c_penalise;
unsigned is_overdispersed;
string f_name_posfix;
int verbosity;
double eta;

static void jd(double v) {
    if (std::isnan(v)) printf("NaN");
    else if (std::isinf(v)) printf(v > 0 ? "Infinity" : "-Infinity");
    else printf("%.17g", v);
}

static void dump_tree(Tree& t) {
    std::vector<Node*> nodes = t.root->get_descendents(true);
    printf("{\"nu\":");
    jd(t.nu);
    printf(",\"nodes\":[");
    for (size_t i = 0; i < nodes.size(); ++i) {
        Node* n = nodes[i];
        printf("%s{\"id\":%d,\"parent\":%d,\"z\":%d,\"c_change\":[", i ? "," : "", n->id,
               n->parent ? n->parent->id : -1, n->z);
        bool first = true;
        for (auto const& x : n->c_change) { printf("%s[%u,%d]", first ? "" : ",", x.first, x.second); first = false; }
        printf("],\"c\":[");
        first = true;
        for (auto const& x : n->c) { printf("%s[%u,%d]", first ? "" : ",", x.first, x.second); first = false; }
        printf("]}");
    }
    printf("]}");
}

static void dump_score_maps(const std::vector<std::map<int, double>>& s) {
    // ids from the first cell (all cells share the key set)
    printf("{\"ids\":[");
    if (!s.empty()) {
        bool first = true;
        for (auto const& kv : s[0]) { printf("%s%d", first ? "" : ",", kv.first); first = false; }
    }
    printf("],\"values\":[");
    for (size_t j = 0; j < s.size(); ++j) {
        printf("%s[", j ? "," : "");
        bool first = true;
        for (auto const& kv : s[j]) { if (!first) printf(","); jd(kv.second); first = false; }
        printf("]");
    }
    printf("]}");
}

static void dump_vec(const std::vector<double>& v) {
    printf("[");
    for (size_t i = 0; i < v.size(); ++i) { if (i) printf(","); jd(v[i]); }
    printf("]");
}

static void dump_maxs(const std::vector<std::pair<int, double>>& v) {
    printf("{\"ids\":[");
    for (size_t i = 0; i < v.size(); ++i) printf("%s%d", i ? "," : "", v[i].first);
    printf("],\"values\":[");
    for (size_t i = 0; i < v.size(); ++i) { if (i) printf(","); jd(v[i].second); }
    printf("]}");
}

int main(int argc, char* argv[]) {
    int n_cells = 0, n_nodes = 0, ploidy = 2, seed = 42, n_moves = 0, worked_example = 0;
    size_t n_regions = 0;
    string region_sizes_file, d_matrix_file, cluster_sizes_file, tree_file, neutral_file;
    double nu = 1.0, gamma = 1.0;
    bool max_scoring = false;
    print_precision = 15; lambda_s = 0.5; lambda_r = 0.1; lambda_c = 0.2; cf = 0.0; c_penalise = 10.0;
    copy_number_limit = 5; is_overdispersed = 1; verbosity = 0; eta = 1e-4;
    vector<float> move_probs;
    string move_probs_str = "0.0,1.0,0.0,1.0,0.0,1.0,0.0,1.0,0.0,1.0,0.01,0.1,0.01,1.0,0.01";
    int max_regions_per_node = 1;

    cxxopts::Options options("ref_dump", "golden vector generator on top of the unmodified reference");
    options.add_options()
        ("region_sizes_file", "", cxxopts::value(region_sizes_file))
        ("d_matrix_file", "", cxxopts::value(d_matrix_file))
        ("cluster_sizes_file", "", cxxopts::value(cluster_sizes_file))
        ("region_neutral_states_file", "", cxxopts::value(neutral_file))
        ("tree_file", "", cxxopts::value(tree_file))
        ("n_regions", "", cxxopts::value(n_regions))
        ("n_cells", "", cxxopts::value(n_cells))
        ("n_nodes", "", cxxopts::value(n_nodes))
        ("max_regions_per_node", "", cxxopts::value(max_regions_per_node))
        ("worked_example", "", cxxopts::value(worked_example))
        ("ploidy", "", cxxopts::value(ploidy))
        ("seed", "", cxxopts::value(seed))
        ("n_moves", "", cxxopts::value(n_moves))
        ("copy_number_limit", "", cxxopts::value(copy_number_limit))
        ("lambda_r", "", cxxopts::value(lambda_r))
        ("lambda_c", "", cxxopts::value(lambda_c))
        ("lambda_s", "", cxxopts::value(lambda_s))
        ("cf", "", cxxopts::value(cf))
        ("c_penalise", "", cxxopts::value(c_penalise))
        ("is_overdispersed", "", cxxopts::value(is_overdispersed))
        ("nu", "", cxxopts::value(nu))
        ("gamma", "", cxxopts::value(gamma))
        ("eta", "", cxxopts::value(eta))
        ("move_probs", "", cxxopts::value(move_probs)->default_value(move_probs_str))
        ("max_scoring", "", cxxopts::value<bool>(max_scoring)->default_value("false"));
    auto result = options.parse(argc, argv);
    if (!result.count("move_probs"))
        move_probs = {0.0f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.01f, 0.1f, 0.01f, 1.0f, 0.01f};

    SingletonRandomGenerator::get_instance(seed);

    vector<vector<double>> D(n_cells, vector<double>(n_regions));
    Utils::read_counts(D, d_matrix_file);
    vector<int> r;
    Utils::read_vector(r, region_sizes_file);
    vector<int> cluster_sizes;
    if (!cluster_sizes_file.empty()) Utils::read_vector(cluster_sizes, cluster_sizes_file);
    else cluster_sizes = vector<int>(n_cells, 1);
    vector<int> neutral;
    if (!neutral_file.empty()) Utils::read_vector(neutral, neutral_file);
    else neutral = vector<int>(n_regions, ploidy);

    Inference mcmc(n_regions, n_cells, neutral, ploidy, verbosity, max_scoring);
    if (worked_example) mcmc.initialize_worked_example(worked_example);
    else if (!tree_file.empty()) { mcmc.initialize_from_file(tree_file); }
    else mcmc.random_initialize(n_nodes, n_regions, 10000, max_regions_per_node);
    mcmc.t.nu = mcmc.t_prime.nu = nu;

    mcmc.compute_t_table(D, r, cluster_sizes);
    mcmc.compute_t_od_scores(D, r, cluster_sizes);
    mcmc.update_t_prime();

    printf("{\"n_cells\":%d,\"n_regions\":%zu,\"is_overdispersed\":%u,\"max_scoring\":%d,\"eta\":", n_cells, n_regions,
           is_overdispersed, max_scoring ? 1 : 0);
    jd(eta);
    printf(",\"cf\":"); jd(cf);
    printf(",\"seed\":%d,\"region_sizes\":[", seed);
    for (size_t i = 0; i < r.size(); ++i) printf("%s%d", i ? "," : "", r[i]);
    printf("],\"neutral_states\":[");
    for (size_t i = 0; i < neutral.size(); ++i) printf("%s%d", i ? "," : "", neutral[i]);
    printf("],\"cluster_sizes\":[");
    for (size_t i = 0; i < cluster_sizes.size(); ++i) printf("%s%d", i ? "," : "", cluster_sizes[i]);
    printf("],\"D\":[");
    for (int j = 0; j < n_cells; ++j) { printf("%s", j ? "," : ""); dump_vec(D[j]); }
    printf("],\n\"init\":{\"tree\":");
    dump_tree(mcmc.t);
    printf(",\"t_scores\":"); dump_score_maps(mcmc.t_scores);
    printf(",\"t_sums\":"); dump_vec(mcmc.t_sums);
    if (max_scoring) { printf(",\"t_maxs\":"); dump_maxs(mcmc.t_maxs); }
    printf(",\"total\":"); jd(mcmc.t.total_attachment_score);
    printf(",\"od_score\":"); jd(mcmc.t.od_score);
    printf(",\"posterior\":"); jd(mcmc.t.posterior_score);
    printf("},\n\"moves\":[");

    const int m = n_cells;
    unsigned size_limit = std::numeric_limits<unsigned>::max();
    bool first_move = true;
    for (int it = 0; it < n_moves; ++it) {
        std::mt19937& gen = SingletonRandomGenerator::get_instance().generator;
        boost::random::discrete_distribution<> d(move_probs.begin(), move_probs.end());
        unsigned move_id = d(gen);
        if (move_id == 10) continue;  // genotype-preserving PR never rescores cells (Inference.h:731-757)
        bool ok = false;
        try {
            switch (move_id) {
                case 0: ok = mcmc.apply_prune_reattach(D, r, false, false); break;
                case 1: ok = mcmc.apply_prune_reattach(D, r, true, false); break;
                case 2: ok = mcmc.apply_swap(D, r, false, false); break;
                case 3: ok = mcmc.apply_swap(D, r, true, false); break;
                case 4: ok = mcmc.apply_add_remove_events(D, r, false, false); break;
                case 5: ok = mcmc.apply_add_remove_events(D, r, true, false); break;
                case 6: ok = mcmc.apply_insert_delete_node(D, r, size_limit, false); break;
                case 7: ok = mcmc.apply_insert_delete_node(D, r, size_limit, true); break;
                case 8: ok = mcmc.apply_condense_split(D, r, size_limit, false); break;
                case 9: ok = mcmc.apply_condense_split(D, r, size_limit, true); break;
                case 11: ok = mcmc.apply_expand_shrink_blocks(D, r, false, false); break;
                case 12: ok = mcmc.apply_add_common_ancestor(D, r, false); break;
                case 13: ok = mcmc.apply_overdispersion_change(D, r, cluster_sizes); break;
                case 14: ok = mcmc.apply_delete_leaf(D, r); break;
                default: break;
            }
        } catch (const std::exception& e) { ok = false; }
        if (!ok) {  // failed attempt: undo (Inference.h:1489,1495)
            mcmc.t_prime = mcmc.t;
            mcmc.t_prime_scores.clear();
            continue;
        }
        Tree* accepted = &mcmc.t;
        try { accepted = mcmc.comparison(m, gamma, move_id, cluster_sizes); } catch (const std::exception& e) { accepted = &mcmc.t; }
        bool acc = (accepted == &mcmc.t_prime);

        printf("%s\n{\"move_id\":%u,\"accepted\":%d,\"tree\":", first_move ? "" : ",", move_id, acc ? 1 : 0);
        first_move = false;
        dump_tree(mcmc.t_prime);
        printf(",\"t_prime_scores\":"); dump_score_maps(mcmc.t_prime_scores);
        printf(",\"t_prime_sums\":"); dump_vec(mcmc.t_prime_sums);
        if (max_scoring) { printf(",\"t_prime_maxs\":"); dump_maxs(mcmc.t_prime_maxs); }
        printf(",\"total\":"); jd(mcmc.t_prime.total_attachment_score);
        printf(",\"od_score\":"); jd(mcmc.t_prime.od_score);
        printf(",\"posterior\":"); jd(mcmc.t_prime.posterior_score);
        printf("}");

        if (acc) {  // Inference.h:867-870
            mcmc.t_sums = mcmc.t_prime_sums;
            mcmc.t_maxs = mcmc.t_prime_maxs;
            mcmc.update_t_scores();
            mcmc.t = mcmc.t_prime;
        } else
            mcmc.t_prime = mcmc.t;  // Inference.h:881
        mcmc.t_prime_scores.clear();  // Inference.h:890
    }
    printf("],\n\"final\":{\"tree\":");
    dump_tree(mcmc.t);
    printf(",\"t_scores\":"); dump_score_maps(mcmc.t_scores);
    printf(",\"t_sums\":"); dump_vec(mcmc.t_sums);
    printf("}}\n");
    return 0;
}
