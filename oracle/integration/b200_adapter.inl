// b200_adapter.inl - the reference-side binding of libscicone_score.so.
//
// TEST / INTEGRATION INFRASTRUCTURE (lives under oracle/, never shipped in the product).
// This is the glue a SCICoNE maintainer would add to scicone/src/Inference.h to run the MCMC host
// (Tree / Node / Inference, moves, priors, accept/reject, CLI, file formats - all unmodified) on the
// B200 scoring engine: the seven seam methods of `Inference` listed in SURVEY.md section 8b are
// forwarded to the C ABI of include/scicone_score.h; everything else is the reference's code.
// oracle/integration/patch_reference.py renames the reference's own definitions of those methods to
// ref_<name> and includes this file at the end of Inference.h (see INTEGRATION.md for the diff).
#include <stdexcept>
#include <unordered_map>
#include "scicone_score.h"

namespace b200 {

struct Flat {
    std::vector<int32_t> id, parent, chg_off, chg_region, chg_value, gen_off, gen_region, gen_value;
    std::unordered_map<const Node*, int> index;
    scicone_tree pod;
    void build(const Tree& t) {
        std::vector<Node*> nodes = t.root->get_descendents(true);  // stack DFS: parent before child, root first
        id.clear(); parent.clear(); chg_off.assign(1, 0); chg_region.clear(); chg_value.clear();
        gen_off.assign(1, 0); gen_region.clear(); gen_value.clear(); index.clear();
        for (size_t i = 0; i < nodes.size(); ++i) index[nodes[i]] = (int)i;
        for (Node* n : nodes) {
            id.push_back(n->id);
            parent.push_back(n->parent ? index.at(n->parent) : -1);
            for (auto const& x : n->c_change) { chg_region.push_back((int32_t)x.first); chg_value.push_back(x.second); }
            chg_off.push_back((int32_t)chg_region.size());
            for (auto const& x : n->c) { gen_region.push_back((int32_t)x.first); gen_value.push_back(x.second); }
            gen_off.push_back((int32_t)gen_region.size());
        }
        chg_region.push_back(0); chg_value.push_back(0); gen_region.push_back(0); gen_value.push_back(0);  // never empty
        pod.n_nodes = (int32_t)nodes.size();
        pod.node_id = id.data(); pod.parent = parent.data();
        pod.chg_off = chg_off.data(); pod.chg_region = chg_region.data(); pod.chg_value = chg_value.data();
        pod.gen_off = gen_off.data(); pod.gen_region = gen_region.data(); pod.gen_value = gen_value.data();
        pod.nu = t.nu;
    }
};

struct State {
    scicone_dataset* ds = nullptr;
    scicone_chain* ch = nullptr;
    int phase = 0;  // 0 idle, 1 subtree roots queued, 2 proposal evaluated (awaiting accept or implicit reject)
    double prime_total = 0.0, t_total = 0.0;
    Flat flat;
};

inline State& state(const Inference* inf) {
    static std::unordered_map<const Inference*, State> all;
    return all[inf];
}

inline void check(int rc) {
    if (rc != SCICONE_OK) throw std::runtime_error(std::string("libscicone_score: ") + scicone_last_error());
}

inline void settle(State& s) {  // Inference.h:881,890: a proposal that was not accepted is dropped
    if (s.phase != 0) {
        check(scicone_reject(s.ch));
        s.phase = 0;
    }
}

inline void open(Inference* inf, State& s, const vector<vector<double>>& D, const vector<int>& r, const vector<int>& cluster_sizes) {
    if (s.ds) return;
    const int M = (int)D.size(), K = (int)r.size();
    std::vector<double> flatD((size_t)M * K);
    for (int j = 0; j < M; ++j) {
        if ((int)D[j].size() != K) throw std::logic_error("Size of the counts per cell needs to be equal to the number of regions!");
        std::copy(D[j].begin(), D[j].end(), flatD.begin() + (size_t)j * K);
    }
    std::vector<int32_t> rr(r.begin(), r.end()), w(cluster_sizes.begin(), cluster_sizes.end()),
        neutral(inf->region_neutral_states.begin(), inf->region_neutral_states.end());
    scicone_dataset_desc d;
    d.n_cells = M; d.n_regions = K; d.D = flatD.data(); d.region_sizes = rr.data(); d.neutral_states = neutral.data();
    d.cluster_sizes = w.data(); d.eta = eta; d.is_overdispersed = (int32_t)is_overdispersed; d.max_scoring = inf->max_scoring ? 1 : 0;
    d.device = 0; d.cell_offset = 0; d.n_cells_global = M;
    check(scicone_dataset_create(&s.ds, &d));
    check(scicone_chain_create(&s.ch, s.ds));
}

inline double prime_total(const Inference* inf) { return state(inf).prime_total; }
inline double t_total(const Inference* inf) { return state(inf).t_total; }

// fills Inference::t_scores / t_sums from the device table (only for the writers at the end of a run)
inline void materialise_t_scores(Inference* inf) {
    State& s = state(inf);
    int32_t n = 0;
    check(scicone_t_n_nodes(s.ch, &n));
    std::vector<int32_t> ids(n);
    check(scicone_t_node_ids(s.ch, ids.data()));
    const int M = inf->n_cells;
    std::vector<double> sc((size_t)M * n), sums(M);
    check(scicone_read_t_scores(s.ch, sc.data()));
    check(scicone_read_t_sums(s.ch, sums.data()));
    inf->t_scores.assign(M, std::map<int, double>());
    inf->t_sums.assign(sums.begin(), sums.end());
    for (int j = 0; j < M; ++j)
        for (int c = 0; c < n; ++c) inf->t_scores[j][ids[c]] = sc[(size_t)j * n + c];
}

}  // namespace b200

// ---- the seam: same signatures as the reference methods they replace --------------------------------------------

void Inference::compute_t_table(const vector<vector<double>> &D, const vector<int> &r, const vector<int> &cluster_sizes) {
    b200::State& s = b200::state(this);
    b200::open(this, s, D, r, cluster_sizes);
    s.phase = 0;
    s.flat.build(this->t);
    double t_sum = 0.0;
    b200::check(scicone_compute_t_table(s.ch, &s.flat.pod, &t_sum));
    s.t_total = t_sum;
    int m = D.size();
    int n = t.get_n_nodes();
    t.total_attachment_score = t_sum;
    t.prior_score = log_tree_prior(m, n);
    t.posterior_score = log_tree_posterior(t_sum, m, t);
}

void Inference::compute_t_od_scores(const vector<vector<double>> &D, const vector<int> &r, const vector<int> &cluster_sizes) {
    b200::State& s = b200::state(this);
    b200::open(this, s, D, r, cluster_sizes);
    double od = 0.0;
    b200::check(scicone_compute_t_od_scores(s.ch, this->t.nu, &od));
    this->t.od_score = od;
}

void Inference::compute_t_prime_od_scores(const vector<vector<double>> &D, const vector<int> &r, const vector<int> &cluster_sizes) {
    b200::State& s = b200::state(this);
    b200::settle(s);
    double od = 0.0;
    b200::check(scicone_compute_t_prime_od_scores(s.ch, this->t_prime.nu, &od));
    this->t_prime.od_score = od;
}

void Inference::compute_t_prime_scores(Node *attached_node, const vector<vector<double>> &D, const vector<int> &r) {
    b200::State& s = b200::state(this);
    if (s.phase == 2) b200::settle(s);
    if (s.phase == 0) s.flat.build(this->t_prime);
    int rc = scicone_compute_t_prime_scores(s.ch, &s.flat.pod, s.flat.index.at(attached_node));
    if (rc != SCICONE_OK) {
        scicone_reject(s.ch);
        s.phase = 0;
        b200::check(rc);
    }
    s.phase = 1;
}

void Inference::compute_t_prime_sums(const vector<vector<double>> &D) {
    b200::State& s = b200::state(this);
    double total = 0.0;
    int rc = scicone_compute_t_prime_sums(s.ch, &s.flat.pod, &total);
    if (rc != SCICONE_OK) {
        scicone_reject(s.ch);
        s.phase = 0;
        b200::check(rc);
    }
    s.prime_total = total;
    s.phase = 2;
}

void Inference::update_t_scores() {
    b200::State& s = b200::state(this);
    b200::check(scicone_accept(s.ch));
    s.t_total = s.prime_total;
    s.phase = 0;
}
