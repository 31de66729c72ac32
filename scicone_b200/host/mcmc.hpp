// MCMC driver of the host: move dispatch, retry policy, Metropolis-Hastings comparison, commit / rollback.
//
// Mirrors the reference's `Inference` class (reference: scicone/src/Inference.h) with the same member names at the
// seam, but every cell-dependent number comes from libscicone_score.so (include/scicone_score.h): the class
// keeps no score table on the host.  An iteration is split in two halves so that any number of chains can be
// advanced in lock-step with ONE kernel launch per step:
//     propose()  - draw the move, mutate t_prime (up to 50 attempts, Inference::apply_multiple_times), queue the
//                  subtree(s) to rescore with scicone_compute_t_prime_scores;
//     [driver]   - scicone_batch_compute_t_prime_sums over all chains that have something queued;
//     finish()   - Inference::comparison with the returned totals, then accept (pointer swaps on the device) or reject.
#ifndef SCICONE_B200_HOST_MCMC_HPP
#define SCICONE_B200_HOST_MCMC_HPP

#include <algorithm>
#include <array>
#include <cfloat>
#include <chrono>
#include <atomic>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <deque>
#include <iostream>
#include <thread>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>

#include "../../include/scicone_score.h"
#include "tree.hpp"

namespace scicone_host {

inline void abi_check(int rc) {
    if (rc != SCICONE_OK) throw std::runtime_error(std::string("libscicone_score: ") + scicone_last_error());
}

// Flat C-ABI view of a Tree (parent-before-child, root first).
struct FlatTree {
    std::vector<int32_t> id, parent, chg_off, chg_region, chg_value, gen_off, gen_region, gen_value;
    std::vector<int> index_of_node;  // pool index -> entry
    scicone_tree pod;
    void build(const Tree& t) {
        std::vector<int> walk = t.descendents(t.root, true);
        id.clear(); parent.clear(); chg_region.clear(); chg_value.clear(); gen_region.clear(); gen_value.clear();
        chg_off.assign(1, 0);
        gen_off.assign(1, 0);
        index_of_node.assign(t.pool.size(), -1);
        for (size_t i = 0; i < walk.size(); ++i) index_of_node[walk[i]] = static_cast<int>(i);
        for (int n : walk) {
            const Node& nd = t.pool[n];
            id.push_back(nd.id);
            parent.push_back(nd.parent == -1 ? -1 : index_of_node[nd.parent]);
            for (auto const& e : nd.c_change) {
                chg_region.push_back(static_cast<int32_t>(e.first));
                chg_value.push_back(e.second);
            }
            chg_off.push_back(static_cast<int32_t>(chg_region.size()));
            for (auto const& e : nd.c) {
                gen_region.push_back(static_cast<int32_t>(e.first));
                gen_value.push_back(e.second);
            }
            gen_off.push_back(static_cast<int32_t>(gen_region.size()));
        }
        chg_region.push_back(0); chg_value.push_back(0); gen_region.push_back(0); gen_value.push_back(0);
        pod.n_nodes = static_cast<int32_t>(walk.size());
        pod.node_id = id.data();
        pod.parent = parent.data();
        pod.chg_off = chg_off.data();
        pod.chg_region = chg_region.data();
        pod.chg_value = chg_value.data();
        pod.gen_off = gen_off.data();
        pod.gen_region = gen_region.data();
        pod.gen_value = gen_value.data();
        pod.nu = t.nu;
    }
};

class Inference {
public:
    Params P;
    Rng rng;
    Tree t, t_prime, best_tree;
    unsigned n_regions;
    int n_cells;  // cells on all ranks
    int ploidy;
    bool max_scoring;
    std::vector<int> region_neutral_states;
    scicone_dataset* ds = nullptr;
    scicone_chain* chain = nullptr;
    FlatTree flat;

    // state of the iteration in flight
    unsigned move_id = 0;
    bool queued = false;            // t_prime has subtrees queued on the device
    bool tp_pristine = false;       // t_prime is exactly the copy of t made by the last restore (and t has not changed since)
    bool nu_move = false;
    double t_total = 0.0;           // committed sum_j aggregate_j * cluster_size_j
    double gamma = 1.0, alpha = 0.0;
    unsigned n_accepted = 0, n_rejected = 0;
    double t_finish_us = 0, t_propose_us = 0, t_prepare_us = 0, t_max_visit_us = 0;  // time spent in this chain's host work
    unsigned long long n_attempts = 0;
    unsigned long long n_rescored_nodes = 0;  // sum over scored proposals of the rescored subtree sizes
    bool have_first_score = false;
    double first_score = 0.0;
    std::ofstream f_chain, f_rel, f_acc, f_gamma;
    bool tracing = false;
    int n_iters_total = 0;

    Inference(const Params& params, int seed, unsigned n_regions_, int n_cells_, const std::vector<int>& neutral, int ploidy_, bool max_scoring_)
        : P(params), rng(seed), n_regions(n_regions_), n_cells(n_cells_), ploidy(ploidy_), max_scoring(max_scoring_),
          region_neutral_states(neutral) {
        t = Tree(ploidy, n_regions, neutral, &P, &rng);
        t_prime = Tree(ploidy, n_regions, neutral, &P, &rng);
        best_tree = Tree(ploidy, n_regions, neutral, &P, &rng);
    }
    Inference(const Inference&) = delete;
    ~Inference() {
        if (chain) scicone_chain_destroy(chain);
    }
    // follower: a cell-sharded run in which another rank owns this chain - the handle only stages the owner's compiled
    // proposals (scicone_import_t_prime); no tree search runs here
    void attach_chain(scicone_dataset* dataset, bool follower = false) {
        ds = dataset;
        is_follower = follower;
        abi_check(follower ? scicone_chain_create_follower(&chain, ds) : scicone_chain_create(&chain, ds));
    }
    bool is_follower = false;

    // ---- initial tree (Inference::random_initialize, Inference.h:99-156) ------------------------------------------
    void random_initialize(unsigned n_nodes, int max_iters, int max_regions_per_node = 1) {
        if (n_nodes == 0) return;
        std::unique_ptr<Tree> cand;
        for (int i = 1;; ++i) {
            cand.reset(new Tree(ploidy, n_regions, region_neutral_states, &P, &rng));
            for (unsigned j = 0; j < n_nodes; ++j) {
                EventMap labels;
                try {
                    labels = cand->random_labels(max_regions_per_node, P.lambda_r);
                } catch (const std::out_of_range&) {
                    cand.reset(new Tree(ploidy, n_regions, region_neutral_states, &P, &rng));
                    break;
                }
                cand->new_child(cand->uniform_sample(true), labels);
            }
            if (i > max_iters)
                throw std::runtime_error("a valid tree cannot be found after " + std::to_string(max_iters) +
                                         " iterations. Please re-set the lambda_r, lambda_c and n_nodes variables.");
            if (cand->n_nodes == 0) continue;
            if (cand->is_valid_subtree(cand->root) && !cand->is_redundant()) break;
        }
        t.copy_from(*cand);
        t.compute_weights();
    }
    void initialize_from_file(const std::string& path) { t.load_from_file(path); }
    void update_t_prime() {
        t_prime.copy_from(t);
        tp_pristine = true;
    }
    // Tree::operator= of the reference after a failed attempt / a rejection (Inference.h:1489,1495,881).  The copy is a
    // pure function of `t` (sibling order included), so it is skipped when t_prime already is that copy, untouched.
    void restore_t_prime() {
        if (tp_pristine && !t_prime.dirty) return;
        t_prime.copy_from(t);
        tp_pristine = true;
    }

    // ---- priors (Inference::log_tree_prior / log_tree_posterior, Inference.h:956-994) -----------------------------
    double log_tree_prior(int m, int n) const {
        double combinatorial = -P.cf * m * n * std::log(2);
        if (max_scoring) m = 0;
        return -(n - 1 + m) * std::log(n + 1) + combinatorial;
    }
    double log_tree_posterior(double tree_sum, int m, const Tree& tree) const {
        double lp = tree_sum + log_tree_prior(m, static_cast<int>(tree.n_nodes));
        lp += tree.event_prior();
        return lp;
    }

    // ---- full passes on the device ----------------------------------------------------------------------------------
    void compute_t_table() {  // Inference::compute_t_table, Inference.h:224-279
        flat.build(t);
        double sum = 0.0;
        abi_check(scicone_compute_t_table(chain, &flat.pod, &sum));
        t_total = sum;
        t.total_attachment_score = sum;
        t.prior_score = log_tree_prior(n_cells, static_cast<int>(t.n_nodes));
        t.posterior_score = log_tree_posterior(sum, n_cells, t);
    }
    // The same full pass in two halves, for cell-sharded runs: the owner compiles it (with the root score of
    // compute_t_od_scores in the same launch), every rank launches it, the owner commits the result.
    void prepare_t_table() {
        flat.build(t);
        abi_check(scicone_prepare_t_table(chain, &flat.pod, 1));
    }
    void finish_t_table(double sum, double od) {
        abi_check(scicone_accept(chain));
        t_total = sum;
        t.total_attachment_score = sum;
        t.prior_score = log_tree_prior(n_cells, static_cast<int>(t.n_nodes));
        t.posterior_score = log_tree_posterior(sum, n_cells, t);
        t.od_score = od;
    }
    void compute_t_od_scores() {  // Inference::compute_t_od_scores, Inference.h:1406-1419
        double od = 0.0;
        abi_check(scicone_compute_t_od_scores(chain, t.nu, &od));
        t.od_score = od;
    }

    // ---- one attempt of each move: mutate t_prime, queue the rescoring (Inference::apply_*) -------------------------
    void queue(int node) {
        if (!queued) flat.build(t_prime);
        abi_check(scicone_compute_t_prime_scores(chain, &flat.pod, flat.index_of_node[node]));
        queued_roots.push_back(flat.index_of_node[node]);
        queued = true;
    }
    bool attempt(unsigned size_limit) {
        switch (move_id) {
            case 0: case 1: {  // Inference::apply_prune_reattach, :204
                int n = t_prime.prune_reattach(move_id == 1);
                if (n < 0) return false;
                queue(n);
                return true;
            }
            case 2: case 3: {  // apply_swap, :1054
                std::vector<int> nodes = t_prime.swap_labels(move_id == 3);
                if (nodes.empty()) return false;  // rejected draw (tree.hpp: rejected_nodes)
                for (int n : nodes) queue(n);
                return true;
            }
            case 4: case 5: {  // apply_add_remove_events, :1198
                int n = t_prime.add_remove_events(move_id == 5);
                if (n < 0) return false;
                queue(n);
                return true;
            }
            case 6: case 7: {  // apply_insert_delete_node, :1247
                int n = t_prime.insert_delete_node(size_limit, move_id == 7, max_scoring);
                if (n < 0) return false;
                queue(n);
                return true;
            }
            case 8: case 9: {  // apply_condense_split, :1291
                int n = t_prime.condense_split_node(size_limit, move_id == 9, max_scoring);
                if (n < 0) return false;
                queue(n);
                return true;
            }
            case 11: {  // apply_expand_shrink_blocks, :1223
                int n = t_prime.expand_shrink_blocks(false);
                if (n < 0) return false;
                queue(n);
                return true;
            }
            case 12: {  // apply_add_common_ancestor, :1516
                int n = t_prime.add_common_ancestor();
                if (n < 0) return false;
                queue(n);
                return true;
            }
            case 13: {  // apply_overdispersion_change, :916-954: Gaussian random walk on log nu
                double step = rng.normal(0.0, 0.02);
                double log_nu = std::log(t_prime.nu) + step;
                if (log_nu > 10.0) return false;
                t_prime.dirty = true;
                t_prime.nu = std::exp(log_nu);
                nu_move = true;  // the root score at the new nu comes back with the proposal's launch
                queue(t_prime.root);
                return true;
            }
            case 14: {  // apply_delete_leaf, :1500
                queue(t_prime.delete_leaf());
                return true;
            }
            default:
                throw std::logic_error("undefined move index");
        }
    }

    // First half of an iteration.  Returns true when a proposal is queued on the device and finish() must be given its totals.
    bool propose(const std::vector<float>& move_probs, unsigned size_limit) {
        move_id = static_cast<unsigned>(rng.discrete(move_probs.begin(), move_probs.end()));
        queued = false;
        queued_roots.clear();
        nu_move = false;
        if (move_id == 10) return false;  // handled entirely on the host in finish()
        for (unsigned a = 0; a < 50; ++a) {  // Inference::apply_multiple_times, Inference.h:1462-1498
            bool ok = false;
            ++n_attempts;
            try {
                ok = attempt(size_limit);
            } catch (const InvalidMove&) {
                break;  // no point in retrying
            } catch (const std::exception&) {
                drop_queue();
                restore_t_prime();
                continue;
            }
            if (ok) return true;
            drop_queue();
            restore_t_prime();
        }
        return false;
    }
    void drop_queue() {
        if (queued) {
            scicone_reject(chain);
            queued = false;
        }
        queued_roots.clear();
    }

    // Inference::comparison (Inference.h:285-516) given the weighted total of t_prime from the device.
    bool comparison(double t_prime_sum, int m) {
        const unsigned t_n = t.n_nodes, tp_n = t_prime.n_nodes;
        t_prime.posterior_score = log_tree_posterior(t_prime_sum, m, t_prime);
        t_prime.prior_score = log_tree_prior(m, static_cast<int>(tp_n));
        t_prime.total_attachment_score = t_prime_sum;
        const bool weighted = (move_id % 2 == 1) && (move_id <= 9);
        double score_diff;
        if (move_id == 13)
            score_diff = (t_prime.od_score - t.od_score) + (t_prime.posterior_score - t.posterior_score);
        else
            score_diff = t_prime.posterior_score - t.posterior_score;
        double nbd = 0.0;
        if (!max_scoring) {  // neighbourhood corrections of the posterior sampler
            nbd = 1.0;
            double sum_chi = 0.0, sum_chi_p = 0.0, sum_omega = 0.0, sum_omega_p = 0.0;
            if (move_id == 1) {
                double c = t.cost() / t_prime.cost();
                if (std::isinf(c)) return false;
                nbd *= c;
            } else if (move_id == 6 || move_id == 7) {
                std::vector<double> v = t.chi_insert_delete(weighted);
                sum_chi = std::accumulate(v.begin(), v.end(), 0.0);
                v = t_prime.chi_insert_delete(weighted);
                sum_chi_p = std::accumulate(v.begin(), v.end(), 0.0);
                if (std::isinf(sum_chi)) throw std::out_of_range("sum_chi is infinity, insert/delete move will be rejected");
                v = t.omega_insert_delete(weighted, max_scoring);
                sum_omega = std::accumulate(v.begin(), v.end(), 0.0);
                v = t_prime.omega_insert_delete(weighted, max_scoring);
                sum_omega_p = std::accumulate(v.begin(), v.end(), 0.0);
            } else if (move_id == 8 || move_id == 9) {
                std::vector<double> v = t.chi_condense_split(weighted);
                sum_chi = std::accumulate(v.begin(), v.end(), 0.0);
                v = t_prime.chi_condense_split(weighted);
                sum_chi_p = std::accumulate(v.begin(), v.end(), 0.0);
                v = t.omega_condense_split(weighted, max_scoring);
                sum_omega = std::accumulate(v.begin(), v.end(), 0.0);
                v = t_prime.omega_condense_split(weighted, max_scoring);
                sum_omega_p = std::accumulate(v.begin(), v.end(), 0.0);
            }
            if (move_id >= 6 && move_id <= 9) nbd *= (t_n < tp_n) ? sum_chi / sum_omega_p : sum_omega / sum_chi_p;
            if (move_id == 7 || move_id == 9) {
                auto ids_of = [](const Tree& tr) {
                    std::set<int> s;
                    for (int n : tr.descendents(tr.root, true)) s.insert(tr.pool[n].id);
                    return s;
                };
                if (t_n > tp_n) {  // a node of t is gone
                    std::set<int> kept = ids_of(t_prime);
                    int gone = -1;
                    for (int n : t.descendents(t.root, false))
                        if (!kept.count(t.pool[n].id)) {
                            gone = n;
                            break;
                        }
                    if (gone < 0) throw std::logic_error("The deleted node could not be found in t!");
                    unsigned d_t = t.pool[gone].n_desc, d_tp = 0;
                    int pid = t.pool[t.pool[gone].parent].id;
                    for (int n : t_prime.descendents(t_prime.root, true))
                        if (t_prime.pool[n].id == pid) {
                            d_tp = t_prime.pool[n].n_desc;
                            break;
                        }
                    if (d_t == 0) throw std::logic_error("The deleted node's parent could not be found in t_prime!");
                    double num = d_tp + 1, den = 2 * d_t;
                    nbd *= num / den;
                } else {  // a node was added
                    std::set<int> had = ids_of(t);
                    int added = -1;
                    for (int n : t_prime.descendents(t_prime.root, false))
                        if (!had.count(t_prime.pool[n].id)) {
                            added = n;
                            break;
                        }
                    if (added < 0) throw std::logic_error("The inserted node could not be found in t_prime!");
                    unsigned d_tp = t_prime.pool[added].n_desc, d_t = 0;
                    int pid = t_prime.pool[t_prime.pool[added].parent].id;
                    for (int n : t.descendents(t.root, true))
                        if (t.pool[n].id == pid) {
                            d_t = t.pool[n].n_desc;
                            break;
                        }
                    if (d_t == 0) throw std::logic_error("The inserted node's parent could not be found in t!");
                    double num = 2 * d_tp, den = d_t + 1;
                    nbd *= num / den;
                }
            }
        }
        const double log_nbd = nbd == 0.0 ? 0.0 : std::log(nbd);
        const double log_acc = log_nbd + gamma * score_diff;
        if (std::isnan(log_acc) || std::isinf(log_acc)) return false;
        if (!t_prime.is_valid_subtree(t_prime.root)) return false;
        if (log_acc > 0) return true;
        return log_acc > std::log(rng.uniform_real01());
    }

    // Second half of an iteration.  `scored`: propose() returned true and (total, od) are the device results.
    void finish(bool scored, double total, double od, int iter) {
        const int m = n_cells;
        Tree* accepted = &t;
        if (move_id == 10) {  // Gibbs move on the committed tree (Inference.h:731-757)
            last_decision = -1;
            bool ok = true;
            try {
                t.genotype_preserving_prune_reattach(gamma);
            } catch (const std::exception&) {
                ok = false;
            }
            if (!ok) {
                ++n_rejected;
                trace_move(-1, iter);
            } else {
                ++n_accepted;
                trace_move(static_cast<int>(move_id), iter);
                t.posterior_score = log_tree_posterior(t_total, m, t);
                t_prime.copy_from(t);
                tp_pristine = true;
            }
        } else {
            bool take = false;
            last_decision = -1;
            if (scored) {
                if (nu_move) t_prime.od_score = od;
                n_rescored_nodes += rescored_last;
                try {
                    take = comparison(total, m);
                } catch (const std::exception&) {
                    take = false;
                }
            }
            const double score_diff = t_prime.posterior_score - t.posterior_score;
            if (scored) last_decision = take ? 1 : 0;
            if (take) {
                accepted = &t_prime;
                if (move_id != 13 && std::abs(score_diff) > 0.1) gamma *= std::exp((1.0 - alpha) * alpha);
                trace_move(static_cast<int>(move_id), iter);
                ++n_accepted;
                abi_check(scicone_accept(chain));  // t_sums = t_prime_sums; t_maxs = t_prime_maxs; update_t_scores()
                queued = false;
                t_total = total;
                t.copy_from(t_prime);
                tp_pristine = false;  // t changed: t_prime is the original, not the copy
                if ((t_prime.posterior_score + t_prime.od_score) > (best_tree.posterior_score + best_tree.od_score)) best_tree.copy_from(t_prime);
            } else {
                if (move_id != 13 && std::abs(score_diff) > 0.1) gamma *= std::exp((0.0 - alpha) * alpha);
                trace_move(-1, iter);
                ++n_rejected;
                drop_queue();
                restore_t_prime();
            }
            if (gamma > 100.0) gamma = 100.0;
            if (gamma < 1.0 / 100.0) gamma = 1.0 / 100.0;
        }
        if (tracing) {
            const double s = accepted->posterior_score + accepted->od_score;
            if (!have_first_score) {
                first_score = s;
                have_first_score = true;
            }
            const bool last = iter == n_iters_total - 1;
            f_chain << std::setprecision(P.print_precision) << s;
            f_rel << std::setprecision(P.print_precision) << s - first_score;
            f_gamma << gamma;
            if (!last) {
                f_chain << ',';
                f_rel << ',';
                f_gamma << ',';
            }
        }
    }
    unsigned rescored_last = 0;
    std::vector<int32_t> queued_roots;  // entries of flat (the compiled t_prime) queued for rescoring by the current proposal
    int last_decision = -1;             // outcome of the last scored proposal: 1 accepted, 0 rejected, -1 nothing was scored
    uint64_t msg_seq = 0;               // cell-sharded runs: messages published (owner) / consumed (follower) for this chain
    std::vector<unsigned char> msg_buf;
    size_t export_cap = 0;              // bytes the last exported proposal needed (sizes the next message buffer)

    void open_traces(int n_iters) {  // Inference.h:532-538
        n_iters_total = n_iters;
        if (P.verbosity > 1) {
            tracing = true;
            f_chain.open(P.postfix + "_markov_chain.csv");
            f_rel.open(P.postfix + "_rel_markov_chain.csv");
            f_acc.open(P.postfix + "_acceptance_ratio.csv");
            f_gamma.open(P.postfix + "_gamma_values.csv");
        }
    }
    void trace_move(int id, int iter) {
        if (tracing) f_acc << std::setprecision(P.print_precision) << id << ((iter == n_iters_total - 1) ? "" : ",");
    }
};

// Cell-sharded runs (one host process per GPU on one node): every chain has ONE owner rank that runs its tree search;
// the other ranks only need the proposal it drew (to queue and compile it against their own shard's tables) and, one
// visit later, whether it was accepted.  Both travel through a shared-memory mailbox, one slot per chain: the owner
// writes the payload and then bumps the sequence number, followers wait for the number they expect and acknowledge it
// once they have copied the message out; the owner does not overwrite a slot before every follower has acknowledged the
// previous message.  (On the GPU the collective inside the launch that scored the previous proposal already orders the
// two; a visit in which nothing of a group was scored, or a run without collectives, has no such launch.)
class Mailbox {
public:
    static constexpr size_t kSlotBytes = 1u << 20;
    static constexpr int kMaxRanks = 16;
    struct Slot {
        std::atomic<uint64_t> seq;
        uint32_t bytes;
        uint32_t pad;
        std::atomic<uint64_t> ack[kMaxRanks];  // last message each rank has consumed
        unsigned char payload[kSlotBytes - 16 - 8 * kMaxRanks];
    };
    ~Mailbox() {
        if (base_) munmap(base_, map_bytes_);
        if (owner_ && !name_.empty()) shm_unlink(name_.c_str());
    }
    // Launch log (dynamic schedule of a cell-sharded run): rank 0 decides which chains go into launch k and writes the
    // list here; the other ranks submit the same launches in the same order (every launch contains a collective).
    static constexpr int kLogRing = 64;
    struct LogHead {
        std::atomic<uint64_t> consumed[kMaxRanks];  // launches each rank has submitted
    };
    struct LogRec {
        std::atomic<uint64_t> seq;  // k + 1 once the members of launch k are in place
        int32_t n;
        int32_t pad;
        // int32_t members[n_chains] follow
    };
    // Rank 0 creates the segment (replacing a stale one that a crashed run may have left under the same name) and zeroes
    // it; every other rank maps it, says hello with its pid and waits for rank 0's echo - a rank that mapped a stale
    // segment gets no echo, notices that the name now leads to another segment, and maps that one.
    struct Hello {
        std::atomic<uint64_t> said[kMaxRanks], echoed[kMaxRanks];
    };
    void open(const std::string& name, int n_chains, int rank, int world) {
        if (world > kMaxRanks) throw std::runtime_error("mailbox: too many ranks");
        name_ = name;
        owner_ = rank == 0;
        n_chains_ = n_chains;
        rec_bytes_ = (sizeof(LogRec) + 4 * static_cast<size_t>(n_chains) + 63) / 64 * 64;
        const size_t slot_bytes = sizeof(Slot) * static_cast<size_t>(n_chains);
        map_bytes_ = slot_bytes + sizeof(LogHead) + rec_bytes_ * kLogRing + sizeof(Hello);
        const uint64_t me = static_cast<uint64_t>(getpid());
        const auto t0 = std::chrono::steady_clock::now();
        auto expired = [&] { return std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120); };
        auto map_fd = [&](int fd) {
            base_ = mmap(nullptr, map_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (base_ == MAP_FAILED) {
                base_ = nullptr;
                close(fd);
                throw std::runtime_error("mmap of the mailbox failed");
            }
            slots_ = static_cast<Slot*>(base_);
            log_head_ = reinterpret_cast<LogHead*>(static_cast<unsigned char*>(base_) + slot_bytes);
            log_recs_ = static_cast<unsigned char*>(base_) + slot_bytes + sizeof(LogHead);
            hello_ = reinterpret_cast<Hello*>(log_recs_ + rec_bytes_ * kLogRing);
        };
        if (rank == 0) {
            shm_unlink(name.c_str());
            int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0) throw std::runtime_error("shm_open(" + name + ") failed");
            if (ftruncate(fd, static_cast<off_t>(map_bytes_)) != 0) {  // a fresh segment is zero filled: every seq / ack / hello is 0
                close(fd);
                throw std::runtime_error("ftruncate on the mailbox failed");
            }
            map_fd(fd);
            close(fd);
            for (int r = 1; r < world; ++r) {
                uint64_t who = 0;
                while ((who = hello_->said[r].load(std::memory_order_acquire)) == 0) {
                    if (expired()) throw std::runtime_error("mailbox: rank " + std::to_string(r) + " did not arrive");
                    std::this_thread::sleep_for(std::chrono::microseconds(200));
                }
                hello_->echoed[r].store(who, std::memory_order_release);
            }
            return;
        }
        for (;;) {
            if (expired()) throw std::runtime_error("mailbox: rank 0 did not create " + name);
            int fd = shm_open(name.c_str(), O_RDWR, 0600);
            struct stat st_mine;
            if (fd < 0 || fstat(fd, &st_mine) != 0 || static_cast<size_t>(st_mine.st_size) < map_bytes_) {
                if (fd >= 0) close(fd);
                std::this_thread::sleep_for(std::chrono::microseconds(500));
                continue;
            }
            map_fd(fd);
            close(fd);
            hello_->said[rank].store(me, std::memory_order_release);
            bool stale = false;
            while (hello_->echoed[rank].load(std::memory_order_acquire) != me && !stale) {
                if (expired()) throw std::runtime_error("mailbox: no answer from rank 0");
                std::this_thread::sleep_for(std::chrono::microseconds(500));
                int fd2 = shm_open(name.c_str(), O_RDWR, 0600);  // does the name still lead to the segment that is mapped here?
                struct stat st_now;
                stale = fd2 < 0 || fstat(fd2, &st_now) != 0 || st_now.st_ino != st_mine.st_ino;
                if (fd2 >= 0) close(fd2);
            }
            if (!stale) return;
            munmap(base_, map_bytes_);
            base_ = nullptr;
        }
    }
    LogRec* log_rec(uint64_t k) const { return reinterpret_cast<LogRec*>(log_recs_ + rec_bytes_ * (k % kLogRing)); }
    // rank 0: members of the next launch; returns its index
    uint64_t log_publish(const std::vector<int>& members, int world) {
        const uint64_t k = next_launch_++;
        if (k >= kLogRing)  // the record about to be overwritten must have been submitted everywhere
            for (int r = 1; r < world; ++r)
                spin_until([&] { return log_head_->consumed[r].load(std::memory_order_acquire) + kLogRing > k; }, "the launch log", r);
        LogRec* rec = log_rec(k);
        rec->n = static_cast<int32_t>(members.size());
        std::memcpy(reinterpret_cast<unsigned char*>(rec) + sizeof(LogRec), members.data(), 4 * members.size());
        rec->seq.store(k + 1, std::memory_order_release);
        return k;
    }
    // other ranks: members of the next launch if rank 0 has published it
    bool log_peek(std::vector<int>& members) const {
        LogRec* rec = log_rec(next_launch_);
        const uint64_t seq = rec->seq.load(std::memory_order_acquire);
        if (seq < next_launch_ + 1) return false;
        if (seq > next_launch_ + 1) throw std::runtime_error("launch log overrun");
        members.resize(static_cast<size_t>(rec->n));
        std::memcpy(members.data(), reinterpret_cast<const unsigned char*>(rec) + sizeof(LogRec), 4 * members.size());
        return true;
    }
    void log_consumed(int rank) { log_head_->consumed[rank].store(++next_launch_, std::memory_order_release); }
    // non-blocking forms for the dynamic schedule, where a worker thread must never wait for another rank
    bool can_publish(int chain, uint64_t seq, int rank, int world) const {
        const Slot& s = slots_[chain];
        for (int r = 0; r < world; ++r)
            if (r != rank && s.ack[r].load(std::memory_order_acquire) + 1 < seq) return false;
        return true;
    }
    const unsigned char* try_receive(int chain, uint64_t seq, uint32_t* bytes) const {
        const Slot& s = slots_[chain];
        if (s.seq.load(std::memory_order_acquire) < seq) return nullptr;
        *bytes = s.bytes;
        return s.payload;
    }
    template <class Pred>
    static void spin_until(Pred&& done, const char* what, int chain) {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned spins = 0;
        while (!done()) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if ((++spins & 0xfffu) == 0) {
                std::this_thread::yield();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                    throw std::runtime_error(std::string("timed out waiting for ") + what + " of chain " + std::to_string(chain));
            }
        }
    }
    // owner: message `seq` of `chain`; `rank` / `world` name the followers whose acknowledgement of seq-1 is awaited
    void publish(int chain, uint64_t seq, const std::vector<unsigned char>& msg, int rank, int world) {
        Slot& s = slots_[chain];
        if (msg.size() > sizeof(s.payload)) throw std::runtime_error("proposal too large for the mailbox slot");
        if (world > kMaxRanks) throw std::runtime_error("mailbox: too many ranks");
        for (int r = 0; r < world; ++r)
            if (r != rank) spin_until([&] { return s.ack[r].load(std::memory_order_acquire) + 1 >= seq; }, "a follower", chain);
        std::memcpy(s.payload, msg.data(), msg.size());
        s.bytes = static_cast<uint32_t>(msg.size());
        s.seq.store(seq, std::memory_order_release);
    }
    const unsigned char* receive(int chain, uint64_t seq, uint32_t* bytes) {
        Slot& s = slots_[chain];
        spin_until([&] { return s.seq.load(std::memory_order_acquire) >= seq; }, "the owner", chain);
        *bytes = s.bytes;
        return s.payload;
    }
    // follower: message `seq` has been copied out of the slot
    void acknowledge(int chain, uint64_t seq, int rank) { slots_[chain].ack[rank].store(seq, std::memory_order_release); }

private:
    std::string name_;
    bool owner_ = false;
    void* base_ = nullptr;
    size_t map_bytes_ = 0, rec_bytes_ = 0;
    int n_chains_ = 0;
    Slot* slots_ = nullptr;
    LogHead* log_head_ = nullptr;
    Hello* hello_ = nullptr;
    unsigned char* log_recs_ = nullptr;
    uint64_t next_launch_ = 0;  // launches published (rank 0) / submitted (other ranks) so far
};

// message = {u8 has_decision, u8 accept, u8 scored, u8 pad, u32 payload bytes, payload}.  The payload is the owner's COMPILED
// proposal (scicone_export_t_prime): it names device rows by index only and is valid on every rank, so a follower neither
// sees the tree nor repeats any of the owner's work - it stages the bytes for the next launch.
inline void pack_message(Inference& c, bool scored, std::vector<unsigned char>& out) {
    uint64_t need = 0;
    if (scored) {
        size_t cap = std::max<size_t>(c.export_cap, 8192);
        out.resize(8 + cap);
        int rc = scicone_export_t_prime(c.chain, out.data() + 8, cap, &need);
        if (rc != SCICONE_OK && need > cap) {
            cap = static_cast<size_t>(need) * 2;
            out.resize(8 + cap);
            rc = scicone_export_t_prime(c.chain, out.data() + 8, cap, &need);
        }
        abi_check(rc);
        c.export_cap = cap;
    }
    out.resize(8 + static_cast<size_t>(need));
    out[0] = static_cast<unsigned char>(c.last_decision >= 0);
    out[1] = static_cast<unsigned char>(c.last_decision == 1);
    out[2] = static_cast<unsigned char>(scored);
    out[3] = 0;
    const uint32_t n32 = static_cast<uint32_t>(need);
    std::memcpy(out.data() + 4, &n32, 4);
}

// follower side: settle the previous proposal of the chain, stage the next one; returns its `scored` flag
inline bool apply_message(Inference& c, const unsigned char* p, uint32_t bytes) {
    if (bytes < 8) throw std::runtime_error("short mailbox message");
    const bool has_decision = p[0] != 0, accept = p[1] != 0, scored = p[2] != 0;
    uint32_t n32;
    std::memcpy(&n32, p + 4, 4);
    if (has_decision) abi_check(accept ? scicone_accept(c.chain) : scicone_reject(c.chain));
    if (!scored) return false;
    if (bytes < 8 + n32) throw std::runtime_error("truncated mailbox message");
    abi_check(scicone_import_t_prime(c.chain, p + 8, n32));
    return true;
}

// fn(i) for i in [0, n) on the scoring library's worker pool; the first exception is rethrown on the caller.
template <class F>
inline void parallel_chains(int n, F&& fn) {
    struct Ctx {
        F* fn;
        std::exception_ptr err;
        std::mutex mu;
    } ctx{&fn, nullptr, {}};
    abi_check(scicone_parallel_for(n, [](int32_t i, void* p) {
        Ctx& c = *static_cast<Ctx*>(p);
        try {
            (*c.fn)(i);
        } catch (...) {
            std::lock_guard<std::mutex> lk(c.mu);
            if (!c.err) c.err = std::current_exception();
        }
    }, &ctx));
    if (ctx.err) std::rethrow_exception(ctx.err);
}

// Advance `chains` (all on one dataset) by n_iters iterations.  The chains are split into `n_groups` groups (default:
// two when there are at least 8 chains); a group moves in lock-step - its chains finish the previous step, draw the
// next proposal and build its launch description in parallel on the host pool, then ONE kernel launch scores them
// all - and while the launch of one group is in flight the host works on the other group.  Every chain still sees exactly the sequence
// propose -> score -> compare of the reference's loop (Inference::infer_mcmc, Inference.h:518-913), so its trajectory
// does not depend on the grouping or on the number of threads.
struct HostPhaseTimes {  // wall time of the driver thread per phase, microseconds (for --stats_file)
    double host_us = 0, submit_us = 0, wait_us = 0;  // parallel finish+propose+prepare region / launch / waiting for the GPU
    unsigned long long launches = 0, launched_chains = 0;  // dynamic schedule: number of launches and proposals in them
    // --timeline=1: {microseconds since the mark, 0 = submit / 1 = collected, proposals in the launch} per event of the driver thread
    bool want_timeline = false;
    std::vector<std::array<double, 3>> timeline;
};

inline void infer_mcmc(std::vector<Inference*>& chains, const std::vector<float>& move_probs, int n_iters, unsigned size_limit,
                       int n_groups = 0, HostPhaseTimes* times = nullptr, int rank = 0, int world = 1, Mailbox* mailbox = nullptr) {
    typedef std::chrono::steady_clock Clock;
    auto us_since = [](Clock::time_point t0) { return std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); };
    HostPhaseTimes local_times;
    HostPhaseTimes& T = times ? *times : local_times;
    const int B = static_cast<int>(chains.size());
    if (B == 0 || n_iters <= 0) return;
    scicone_dataset* ds = chains[0]->ds;
    int G = n_groups > 0 ? n_groups : (B >= 8 ? 2 : 1);
    if (G > B) G = B;
    if (G > 4) G = 4;  // the library keeps at most 8 launches in flight
    if (world > 1 && (B + G - 1) / G > SCICONE_MAX_SHARDED_BATCH)
        throw std::runtime_error("static schedule: more than " + std::to_string(SCICONE_MAX_SHARDED_BATCH) +
                                 " chains per group on a cell-sharded dataset - use the dynamic schedule or fewer chains per process");
    abi_check(scicone_dataset_reserve(ds, (B + G - 1) / G, 0));  // the reduction scratch cannot grow while a launch is outstanding
    parallel_chains(B, [&](int b) {
        chains[b]->best_tree.copy_from(chains[b]->t);  // Inference.h:540
        if (mailbox == nullptr || b % world == rank) chains[b]->open_traces(n_iters);  // a follower writes no traces
    });
    struct Group {
        std::vector<int> members;
        std::vector<char> scored;
        std::vector<scicone_chain*> handles;
        std::vector<int> who;  // member position of each handle
        std::vector<double> totals, ods, tot_m, od_m;
        uint64_t ticket = 0;
        bool in_flight = false;
    };
    std::vector<Group> groups(G);
    for (int g = 0; g < G; ++g) {
        for (int b = static_cast<int>(static_cast<long long>(B) * g / G); b < static_cast<int>(static_cast<long long>(B) * (g + 1) / G); ++b)
            groups[g].members.push_back(b);
        const size_t n = groups[g].members.size();
        groups[g].scored.assign(n, 0);
        groups[g].totals.assign(n, 0.0);
        groups[g].ods.assign(n, 0.0);
        groups[g].tot_m.assign(n, 0.0);
        groups[g].od_m.assign(n, 0.0);
    }
    // One visit of a group: collect the launch of its previous step, then ONE parallel region in which every chain
    // finishes that step (compare, accept / reject), draws its next proposal and builds the launch description, then
    // one launch for the whole group.
    auto visit = [&](Group& g, int it) {  // `it` = step to propose; step it-1 (if any) is finished first
        const int n = static_cast<int>(g.members.size());
        Clock::time_point t0 = Clock::now();
        if (g.in_flight) {
            int rc = scicone_batch_wait(ds, g.ticket, g.totals.data(), g.ods.data());
            if (rc != SCICONE_OK && rc != SCICONE_ERR_NAN) abi_check(rc);
            g.in_flight = false;
            for (size_t h = 0; h < g.who.size(); ++h) {
                g.tot_m[g.who[h]] = g.totals[h];
                g.od_m[g.who[h]] = g.ods[h];
            }
        }
        T.wait_us += us_since(t0);
        t0 = Clock::now();
        // Cell-sharded run with a mailbox: a chain's tree search runs on its owner rank only.  Owned chains first (nobody
        // waits for anybody), then the followers of the other ranks' chains.
        auto owned = [&](int b) { return mailbox == nullptr || b % world == rank; };
        parallel_chains(n, [&](int i) {
            if (!owned(g.members[i])) return;
            Inference* c = chains[g.members[i]];
            const Clock::time_point c0 = Clock::now();
            Clock::time_point c1 = c0, c2 = c0;
            if (it > 0) {
                if (g.scored[i]) {
                    int32_t nr = 0;
                    scicone_t_prime_n_rescored(c->chain, &nr);
                    c->rescored_last = static_cast<unsigned>(nr);
                }
                c->finish(g.scored[i] != 0, g.tot_m[i], g.od_m[i], it - 1);
                c1 = c2 = Clock::now();
            } else
                c->last_decision = -1;
            if (it < n_iters) {
                g.scored[i] = c->propose(move_probs, size_limit) ? 1 : 0;
                c2 = Clock::now();
                if (g.scored[i]) abi_check(scicone_prepare_t_prime(c->chain));
            } else
                g.scored[i] = 0;
            if (mailbox) {  // the decision about the previous proposal and the next proposal, in one message
                std::vector<unsigned char>& msg = c->msg_buf;
                pack_message(*c, g.scored[i] != 0, msg);
                mailbox->publish(g.members[i], ++c->msg_seq, msg, rank, world);
            }
            const Clock::time_point c3 = Clock::now();
            auto us = [](Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
            c->t_finish_us += us(c0, c1);
            c->t_propose_us += us(c1, c2);
            c->t_prepare_us += us(c2, c3);
            c->t_max_visit_us = std::max(c->t_max_visit_us, us(c0, c3));
        });
        if (mailbox)
            parallel_chains(n, [&](int i) {
                if (owned(g.members[i])) return;
                Inference* c = chains[g.members[i]];
                uint32_t bytes = 0;
                const unsigned char* msg = mailbox->receive(g.members[i], ++c->msg_seq, &bytes);
                g.scored[i] = apply_message(*c, msg, bytes) ? 1 : 0;
                mailbox->acknowledge(g.members[i], c->msg_seq, rank);
            });
        T.host_us += us_since(t0);
        if (it >= n_iters) return;
        t0 = Clock::now();
        g.handles.clear();
        g.who.clear();
        for (int i = 0; i < n; ++i)
            if (g.scored[i]) {
                g.handles.push_back(chains[g.members[i]]->chain);
                g.who.push_back(i);
            }
        g.in_flight = !g.handles.empty();
        if (g.in_flight) abi_check(scicone_batch_submit(g.handles.data(), nullptr, static_cast<int32_t>(g.handles.size()), &g.ticket));
        T.submit_us += us_since(t0);
    };
    for (int it = 0; it <= n_iters; ++it)
        for (Group& g : groups) visit(g, it);
}

// The same chains under a dynamic schedule: no group ever waits for its slowest member.  Worker threads take chains whose
// launch has come back, finish that iteration, draw and compile the next proposal and hand the chain to the driver
// thread; the driver launches whatever is ready as soon as `min_batch` chains are (or nothing else is coming), with at
// most two launches in flight, and feeds completed launches back to the workers.  Batch composition varies from run to
// run; a chain's own sequence propose -> score -> compare does not, so trajectories are those of infer_mcmc.
//
// Cell-sharded runs (world > 1, with a mailbox): every rank must submit the same launches in the same order.  A chain's
// search runs on its owner rank, which publishes each scored proposal (with the decision about the previous one); the
// other ranks compile it against their shard as the message arrives.  Rank 0 composes the launches from the chains that
// are compiled on rank 0 and writes each list to the launch log; the other ranks replay the log, submitting launch k as
// soon as its chains are compiled locally.  Nothing ever waits for a later launch, and no worker thread blocks on
// another rank: a chain whose message (or whose followers' acknowledgement) is not there yet goes back to the queue.
//
// `n_warmup` > 0 (timing runs): the first n_warmup iterations of every chain are a separate phase - when ALL chains have
// finished them and every launch has been collected (host and device idle, the same state as between two calls) the
// driver calls `at_mark` (which starts the caller's clock) and releases the chains for the n_iters measured iterations.
// Worker threads, streams and buffers stay as they are across the mark, as they would in a long run.
inline void infer_mcmc_dynamic(std::vector<Inference*>& chains, const std::vector<float>& move_probs, int n_iters, unsigned size_limit,
                               int min_batch = 0, HostPhaseTimes* times = nullptr, int n_streams = 2, int rank = 0, int world = 1,
                               Mailbox* mailbox = nullptr, int n_warmup = 0, const std::function<void()>& at_mark = nullptr) {
    typedef std::chrono::steady_clock Clock;
    const int B = static_cast<int>(chains.size());
    if (B == 0 || n_iters <= 0) return;
    const int n_total = n_warmup + n_iters;
    std::atomic<int> limit{n_warmup > 0 ? n_warmup : n_total};  // iterations the chains may run in the current phase
    if (world > 1 && mailbox == nullptr) throw std::runtime_error("the dynamic schedule of a multi-process run needs the mailbox");
    scicone_dataset* ds = chains[0]->ds;
    // Measured (us per iteration of all chains).  32 chains x 10 k cells, one GPU, 300 iterations: 1 -> 110, 2 -> 104-119,
    // 4 -> 106, 8 -> 110 (a launch of few small proposals costs little more than its latency, so waiting for company does
    // not pay).  16 chains x 100 k cells: 3 -> 488, 5 -> 458 (launches are bandwidth work there).  Cell-sharded runs keep the
    // larger batch: every launch also costs a round of messages.
    if (min_batch <= 0) min_batch = std::max(1, (world == 1 && chains[0]->n_cells <= 20000) ? B / 8 : B / 4);
    // two launch streams: K1 of one launch overlaps the proposal kernel of the other (a launch is always collected here
    // before its chains are settled, which is what the second stream requires)
    n_streams = std::max(1, std::min(n_streams, 4));
    if (n_streams > 1) abi_check(scicone_set_streams(ds, n_streams));
    const size_t max_flights = static_cast<size_t>(std::max(2, n_streams));
    const bool split_heavy = getenv("SCICONE_SPLIT_HEAVY") ? atoi(getenv("SCICONE_SPLIT_HEAVY")) != 0 : false;  // measured: 191 us per iteration of 32 chains with the split, 176 without
    // the reduction scratch is sized once, before the first launch: a launch never has more proposals than it has slots
    abi_check(scicone_dataset_reserve(ds, world > 1 ? std::min(B, SCICONE_MAX_SHARDED_BATCH) : B, 0));
    int32_t max_batch = 0;
    abi_check(scicone_max_batch(ds, &max_batch));
    if (world > 1 && max_batch > SCICONE_MAX_SHARDED_BATCH) max_batch = SCICONE_MAX_SHARDED_BATCH;
    auto owned = [&](int b) { return mailbox == nullptr || b % world == rank; };
    parallel_chains(B, [&](int b) {
        chains[b]->best_tree.copy_from(chains[b]->t);  // Inference.h:540
        if (owned(b)) chains[b]->open_traces(n_total);
    });
    struct Slot {
        int iter = 0;             // iteration whose proposal is queued / in flight
        bool has_result = false;  // total / od of that proposal are in
        bool to_publish = false;  // owner: the message is packed, the followers have not acknowledged the previous one yet
        bool queued = false;      //        ... and it carries a scored proposal
        double total = 0.0, od = 0.0;
        bool heavy = false;  // the queued proposal rescores many nodes (a nu move): it goes into a launch of its own kind
        Clock::time_point waiting_since{};  // first unsuccessful poll of the other ranks
        bool polling = false;
    };
    std::vector<Slot> st(B);
    std::mutex mu;  // guards work, ready, n_done, n_host
    std::deque<int> work;     // chains that need host work
    std::vector<int> ready;   // chains with a compiled proposal, waiting for a launch
    int n_done = 0, n_host = B;  // finished chains; chains currently owned by the host side (queued or being worked on)
    std::atomic<bool> stop{false};
    std::atomic<int> n_work{0};  // size of `work`, readable without the mutex
    std::exception_ptr err;
    for (int b = 0; b < B; ++b) work.push_back(b);
    n_work.store(B, std::memory_order_release);

    Clock::time_point t_mark = Clock::now();
    std::vector<std::array<double, 3>> done_at;  // --timeline: {microseconds since the mark, 2, chain} when a chain has finished its phase (guarded by mu)
    // the chain goes back to the queue: the other ranks are not there yet
    auto poll_again = [&](int b, const char* what) {
        Slot& s = st[b];
        if (!s.polling) {
            s.polling = true;
            s.waiting_since = Clock::now();
        } else if (Clock::now() - s.waiting_since > std::chrono::seconds(120))
            throw std::runtime_error(std::string("timed out waiting for ") + what + " of chain " + std::to_string(b));
        // a short breath, not a long one: a rank follows up to B - B/world chains through this queue, and the time one lap
        // takes is the delay with which it notices a proposal the owner has published
        for (int i = 0; i < 4; ++i) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        std::lock_guard<std::mutex> lk(mu);
        work.push_back(b);
        n_work.fetch_add(1, std::memory_order_release);
    };
    auto hand_over = [&](int b, bool queued) {
        st[b].polling = false;
        if (queued) {
            double cost = 0.0;
            abi_check(scicone_t_prime_cost(chains[b]->chain, &cost));
            st[b].heavy = cost >= 40.0;
        }
        std::lock_guard<std::mutex> lk(mu);
        --n_host;
        if (queued)
            ready.push_back(b);
        else {
            ++n_done;
            if (times && times->want_timeline) done_at.push_back({std::chrono::duration<double, std::micro>(Clock::now() - t_mark).count(), 2.0, static_cast<double>(b)});
        }
    };
    auto host_work = [&](int b) {
        Inference* c = chains[b];
        Slot& s = st[b];
        if (!owned(b)) {  // follower: the owner's decision about the previous proposal and its next one
            uint32_t bytes = 0;
            const unsigned char* msg = mailbox->try_receive(b, c->msg_seq + 1, &bytes);
            if (!msg) return poll_again(b, "the owner");
            ++c->msg_seq;
            const bool scored = apply_message(*c, msg, bytes);
            mailbox->acknowledge(b, c->msg_seq, rank);
            return hand_over(b, scored);
        }
        if (!s.to_publish) {
            const Clock::time_point c0 = Clock::now();
            int decision = -1;
            if (s.has_result) {
                int32_t nr = 0;
                scicone_t_prime_n_rescored(c->chain, &nr);
                c->rescored_last = static_cast<unsigned>(nr);
                c->finish(true, s.total, s.od, s.iter);
                decision = c->last_decision;
                s.has_result = false;
                ++s.iter;
            }
            s.queued = false;
            const int lim = limit.load(std::memory_order_acquire);
            double dbg_prep = 0.0;
            const unsigned long long dbg_att0 = c->n_attempts;
            while (s.iter < lim) {
                if (c->propose(move_probs, size_limit)) {
                    const Clock::time_point p0 = Clock::now();
                    abi_check(scicone_prepare_t_prime(c->chain));
                    dbg_prep = std::chrono::duration<double, std::micro>(Clock::now() - p0).count();
                    s.queued = true;
                    break;
                }
                c->finish(false, 0.0, 0.0, s.iter);  // nothing to score (move 10, or every attempt failed)
                ++s.iter;
            }
            if (mailbox) {  // one message per scored proposal, and a last one that only carries the decision
                c->last_decision = decision;
                pack_message(*c, s.queued, c->msg_buf);
                s.to_publish = true;
            }
            const double us = std::chrono::duration<double, std::micro>(Clock::now() - c0).count();
            // SCICONE_DEBUG_VISITS=1: note visits whose tree work (everything but the compile) took more than 40 us
            static const bool debug_visits = getenv("SCICONE_DEBUG_VISITS") != nullptr;
            if (debug_visits && us - dbg_prep > 40.0)
                fprintf(stderr, "visit chain %d iter %d move %u us %.1f prep %.1f attempts %llu\n", b, s.iter, c->move_id, us - dbg_prep, dbg_prep, c->n_attempts - dbg_att0);
            c->t_propose_us += us;
            c->t_max_visit_us = std::max(c->t_max_visit_us, us);
        }
        if (s.to_publish) {
            if (!mailbox->can_publish(b, c->msg_seq + 1, rank, world)) return poll_again(b, "a follower");
            mailbox->publish(b, ++c->msg_seq, c->msg_buf, rank, world);
            s.to_publish = false;
        }
        hand_over(b, s.queued);
    };
    auto worker = [&]() {
        int idle = 0;
        while (!stop.load(std::memory_order_acquire)) {
            int b = -1;
            if (n_work.load(std::memory_order_acquire) > 0) {  // idle workers watch the counter, not the mutex
                std::lock_guard<std::mutex> lk(mu);
                if (!work.empty()) {
                    b = work.front();
                    work.pop_front();
                    n_work.fetch_sub(1, std::memory_order_release);
                }
            }
            if (b < 0) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
                if (++idle > 20000) std::this_thread::yield();
                continue;
            }
            idle = 0;
            try {
                host_work(b);
            } catch (...) {
                std::lock_guard<std::mutex> lk(mu);
                if (!err) err = std::current_exception();
                stop.store(true, std::memory_order_release);
            }
        }
    };
    const int n_workers = std::max(1, scicone_host_threads() - 1);
    std::vector<std::thread> pool;
    for (int i = 0; i < n_workers; ++i) pool.emplace_back(worker);

    struct Flight {
        uint64_t ticket;
        std::vector<int> members;
    };
    std::deque<Flight> flights;
    std::vector<scicone_chain*> handles;
    std::vector<double> totals(B), ods(B);
    unsigned long long n_launches = 0, n_launched_chains = 0;
    auto note = [&](double kind, size_t n) {
        if (times && times->want_timeline)
            times->timeline.push_back({std::chrono::duration<double, std::micro>(Clock::now() - t_mark).count(), kind, static_cast<double>(n)});
    };
    auto collect = [&](Flight& f) {
        int rc = scicone_batch_wait(ds, f.ticket, totals.data(), ods.data());
        if (rc != SCICONE_OK && rc != SCICONE_ERR_NAN) abi_check(rc);
        note(1, f.members.size());
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < f.members.size(); ++i) {
            Slot& s = st[f.members[i]];
            s.total = totals[i];
            s.od = ods[i];
            s.has_result = true;
            work.push_back(f.members[i]);
            n_work.fetch_add(1, std::memory_order_release);
            ++n_host;
        }
    };
    auto submit = [&](std::vector<int>& take) {
        handles.clear();
        for (int b : take) handles.push_back(chains[b]->chain);
        Flight f;
        f.members = take;
        abi_check(scicone_batch_submit(handles.data(), nullptr, static_cast<int32_t>(handles.size()), &f.ticket));
        note(0, take.size());
        flights.push_back(std::move(f));
        ++n_launches;
        n_launched_chains += take.size();
    };
    const bool replay = world > 1 && rank != 0;
    std::vector<int> wanted;  // replaying ranks: members of the next launch of the log
    bool have_wanted = false;
    try {
        for (;;) {
            if (stop.load(std::memory_order_acquire)) break;
            std::vector<int> take;
            bool all_done = false;
            if (replay && !have_wanted && flights.size() < max_flights) have_wanted = mailbox->log_peek(wanted);
            {
                std::lock_guard<std::mutex> lk(mu);
                all_done = n_done == B;
                if (replay) {
                    if (have_wanted && flights.size() < max_flights) {  // launch k goes out once all of its chains are compiled here
                        size_t found = 0;
                        for (int b : wanted)
                            if (std::find(ready.begin(), ready.end(), b) != ready.end()) ++found;
                        if (found == wanted.size()) {
                            for (int b : wanted) ready.erase(std::find(ready.begin(), ready.end(), b));
                            take = wanted;
                        }
                    }
                } else {
                    const bool nothing_coming = n_host == 0 && flights.empty();
                    if (!ready.empty() && flights.size() < max_flights && (static_cast<int>(ready.size()) >= min_batch || nothing_coming || n_host == 0))
                    {
                        take.swap(ready);
                        // Whole-tree rescorings (nu moves) need the table / increment kernels in front of the proposal
                        // kernel; a launch of small proposals is ONE kernel.  Keep the two kinds apart: the heavy ones go
                        // first, the small ones follow in the next launch (on another lane).
                        if (split_heavy) {
                            size_t n_heavy = 0;
                            for (int b : take) n_heavy += st[b].heavy ? 1 : 0;
                            if (n_heavy > 0 && n_heavy < take.size()) {
                                std::vector<int> heavy;
                                for (int b : take) (st[b].heavy ? heavy : ready).push_back(b);
                                take.swap(heavy);
                            }
                        }
                        if (static_cast<int>(take.size()) > max_batch) {  // the rest waits for the next launch
                            ready.insert(ready.end(), take.begin() + max_batch, take.end());
                            take.resize(static_cast<size_t>(max_batch));
                        }
                    }
                }
            }
            if (all_done) {
                if (limit.load(std::memory_order_acquire) >= n_total) break;
                // end of the warm-up phase: every chain has stopped at the mark and nothing is in flight
                if (at_mark) at_mark();
                if (times) times->timeline.clear();
                n_launches = n_launched_chains = 0;
                std::lock_guard<std::mutex> lk(mu);
                t_mark = Clock::now();
                done_at.clear();
                limit.store(n_total, std::memory_order_release);
                n_done = 0;
                n_host = B;
                for (int b = 0; b < B; ++b) work.push_back(b);
                n_work.store(B, std::memory_order_release);
                continue;
            }
            if (!take.empty()) {
                if (replay) {
                    submit(take);
                    mailbox->log_consumed(rank);
                    have_wanted = false;
                } else {
                    std::sort(take.begin(), take.end());
                    if (world > 1) mailbox->log_publish(take, world);
                    submit(take);
                }
                continue;
            }
            if (!flights.empty()) {
                // one GPU: whichever launch has come back (a launch of small proposals overtakes a whole-tree rescoring that
                // was submitted before it); cell-sharded: in order - launch k uses lane k mod n on every rank
                const size_t n_look = world == 1 ? flights.size() : 1;
                bool got = false;
                for (size_t f = 0; f < n_look && !got; ++f) {
                    int32_t done = 0;
                    abi_check(scicone_batch_poll(ds, flights[f].ticket, &done));
                    if (done) {
                        collect(flights[f]);
                        flights.erase(flights.begin() + static_cast<std::ptrdiff_t>(f));
                        got = true;
                    }
                }
                if (!got && flights.size() >= max_flights) {
                    collect(flights.front());
                    flights.pop_front();
                }
            }
        }
    } catch (...) {
        std::lock_guard<std::mutex> lk(mu);
        if (!err) err = std::current_exception();
    }
    stop.store(true, std::memory_order_release);
    for (std::thread& t : pool) t.join();
    if (times) {
        times->launches = n_launches;
        times->launched_chains = n_launched_chains;
        times->timeline.insert(times->timeline.end(), done_at.begin(), done_at.end());
    }
    if (err) std::rethrow_exception(err);
}

}  // namespace scicone_host
#endif
