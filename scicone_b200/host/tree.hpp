// Event tree of the MCMC host: nodes, genotypes, move set, validity rules and priors.
//
// This is the host half of the north-star design: the tree search itself (what to propose, whether a
// proposal is legal, the structural priors) stays on the CPU and keeps the behaviour of the reference's
// Tree / Node classes (reference: scicone/src/Tree.h, scicone/src/Node.h); every cell-dependent number comes
// from libscicone_score.so.  The implementation is index based (one contiguous node pool per tree, no
// pointers, no per-node heap objects except the event maps) and carries its own random stream and parameter
// block, so many chains can live in one process - the reference has one tree engine per process
// (global parameters, singleton generator).
//
// Trajectory parity with the reference requires more than the same distributions: the order in which nodes
// are enumerated decides which node a random index selects.  The places where that order is defined are
// marked "ORDER" below and follow the reference: children are appended at the tail of the sibling list
// (Tree.h:348-361), subtree walks use an explicit stack and therefore visit the LAST child first
// (Node.h:129-152), and copying a tree re-inserts nodes in that walk order, which reverses every sibling
// list (Tree.h:523-542).
#ifndef SCICONE_B200_HOST_TREE_HPP
#define SCICONE_B200_HOST_TREE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <type_traits>
#include <iomanip>
#include <limits>
#include <map>
#include <numeric>
#include <ostream>
#include <random>
#include <set>
#include <sstream>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rng_dist.hpp"

namespace scicone_host {

// Ordered map on a sorted array with inline storage for the first N entries: the event maps of a node hold a handful
// of regions, and a chain copies its whole tree (every node, three maps each) at least twice per iteration, so a
// node-per-entry std::map would spend the host's time in malloc.  Iteration order, lookup semantics and comparison are
// those of std::map (the reference's std::map<u_int,int>), which the move set and the redundancy check rely on.
template <class K, class V, unsigned N>
class SmallMap {
public:
    typedef std::pair<K, V> value_type;
    typedef value_type* iterator;
    typedef const value_type* const_iterator;
    typedef std::reverse_iterator<const_iterator> const_reverse_iterator;

    SmallMap() {}
    SmallMap(const SmallMap& o) { assign(o); }
    SmallMap(SmallMap&& o) noexcept { steal(o); }
    SmallMap& operator=(const SmallMap& o) {
        if (this != &o) assign(o);
        return *this;
    }
    SmallMap& operator=(SmallMap&& o) noexcept {
        if (this != &o) {
            release();
            steal(o);
        }
        return *this;
    }
    ~SmallMap() { release(); }

    iterator begin() { return data(); }
    iterator end() { return data() + n_; }
    const_iterator begin() const { return data(); }
    const_iterator end() const { return data() + n_; }
    const_reverse_iterator rbegin() const { return const_reverse_iterator(end()); }
    const_reverse_iterator rend() const { return const_reverse_iterator(begin()); }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    void clear() { n_ = 0; }

    iterator find(const K& k) {
        iterator it = lower(k);
        return (it != end() && it->first == k) ? it : end();
    }
    const_iterator find(const K& k) const {
        const_iterator it = const_cast<SmallMap*>(this)->lower(k);
        return (it != end() && it->first == k) ? it : end();
    }
    size_t count(const K& k) const { return find(k) != end() ? 1 : 0; }
    V& operator[](const K& k) {
        iterator it = lower(k);
        if (it != end() && it->first == k) return it->second;
        const size_t pos = static_cast<size_t>(it - data());
        grow(n_ + 1);
        value_type* d = data();
        std::memmove(static_cast<void*>(d + pos + 1), static_cast<const void*>(d + pos), (n_ - pos) * sizeof(value_type));
        d[pos] = value_type(k, V());
        ++n_;
        return d[pos].second;
    }
    V& at(const K& k) {
        iterator it = find(k);
        if (it == end()) throw std::out_of_range("map::at");
        return it->second;
    }
    const V& at(const K& k) const {
        const_iterator it = find(k);
        if (it == end()) throw std::out_of_range("map::at");
        return it->second;
    }
    iterator erase(iterator it) {
        std::memmove(static_cast<void*>(it), static_cast<const void*>(it + 1), static_cast<size_t>(end() - (it + 1)) * sizeof(value_type));
        --n_;
        return it;
    }
    size_t erase(const K& k) {
        iterator it = find(k);
        if (it == end()) return 0;
        erase(it);
        return 1;
    }
    void swap(SmallMap& o) {
        SmallMap tmp(std::move(o));
        o = std::move(*this);
        *this = std::move(tmp);
    }
    friend bool operator==(const SmallMap& a, const SmallMap& b) { return a.n_ == b.n_ && std::equal(a.begin(), a.end(), b.begin()); }
    friend bool operator!=(const SmallMap& a, const SmallMap& b) { return !(a == b); }
    friend bool operator<(const SmallMap& a, const SmallMap& b) { return std::lexicographical_compare(a.begin(), a.end(), b.begin(), b.end()); }

private:
    static_assert(std::is_trivially_copy_constructible<K>::value && std::is_trivially_copy_constructible<V>::value &&
                      std::is_trivially_destructible<K>::value && std::is_trivially_destructible<V>::value,
                  "entries are moved with memmove");
    value_type* data() { return heap_ ? heap_ : reinterpret_cast<value_type*>(inline_); }
    const value_type* data() const { return heap_ ? heap_ : reinterpret_cast<const value_type*>(inline_); }
    iterator lower(const K& k) {
        return std::lower_bound(begin(), end(), k, [](const value_type& e, const K& key) { return e.first < key; });
    }
    void grow(size_t want) {
        if (want <= cap_) return;
        size_t cap = std::max<size_t>(want, 2 * static_cast<size_t>(cap_));
        value_type* fresh = static_cast<value_type*>(::operator new(cap * sizeof(value_type)));
        std::memcpy(static_cast<void*>(fresh), static_cast<const void*>(data()), n_ * sizeof(value_type));
        if (heap_) ::operator delete(heap_);
        heap_ = fresh;
        cap_ = static_cast<unsigned>(cap);
    }
    void assign(const SmallMap& o) {
        n_ = 0;
        grow(o.n_);
        std::memcpy(static_cast<void*>(data()), static_cast<const void*>(o.data()), o.n_ * sizeof(value_type));
        n_ = o.n_;
    }
    void steal(SmallMap& o) {
        if (o.heap_) {
            heap_ = o.heap_;
            cap_ = o.cap_;
            o.heap_ = nullptr;
            o.cap_ = N;
        } else {
            heap_ = nullptr;
            cap_ = N;
            std::memcpy(static_cast<void*>(inline_), static_cast<const void*>(o.inline_), o.n_ * sizeof(value_type));
        }
        n_ = o.n_;
        o.n_ = 0;
    }
    void release() {
        if (heap_) ::operator delete(heap_);
        heap_ = nullptr;
        cap_ = N;
        n_ = 0;
    }
    alignas(value_type) unsigned char inline_[N * sizeof(value_type)];
    value_type* heap_ = nullptr;
    unsigned n_ = 0, cap_ = N;
};

typedef SmallMap<unsigned, int, 8> EventMap;  // region -> value; ordered like the reference's std::map<u_int,int>

// A move that cannot be made on this tree at all: the driver stops retrying (reference: CustomExceptions.h InvalidMove).
struct InvalidMove : std::exception {
    std::string msg;
    explicit InvalidMove(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
struct InvalidTree : std::runtime_error {
    explicit InvalidTree(const std::string& m) : std::runtime_error(m) {}
};

// The reference's process-wide tunables (globals.cpp), one block per chain here.
struct Params {
    int print_precision = 15;
    int copy_number_limit = 5;
    double lambda_s = 0.5, lambda_r = 0.1, lambda_c = 0.2;
    double cf = 0.0, c_penalise = 10.0;
    unsigned is_overdispersed = 1;
    double eta = 1e-4;
    int verbosity = 0;
    std::string postfix;
};

// One random stream per chain: std::mt19937 consumed through the Boost-compatible distributions of
// rng_dist.hpp (reference: SingletonRandomGenerator.h:21-47 and the call sites of SURVEY.md App. B).
struct Rng {
    std::mt19937 gen;
    explicit Rng(int seed) {
        if (seed == -1) {
            std::random_device rd;
            gen.seed(rd());
        } else
            gen.seed(static_cast<std::mt19937::result_type>(seed));
    }
    int uniform_int(int lo, int hi) { return scicone_rng::uniform_int_distribution<>(lo, hi)(gen); }
    double uniform_real01() { return scicone_rng::uniform_real_distribution<double>(0.0, 1.0)(gen); }
    bool coin() { return scicone_rng::bernoulli_distribution<double>(0.5)(gen); }
    int poisson(double mean) { return scicone_rng::poisson_distribution<int>(mean)(gen); }
    template <class It>
    int discrete(It first, It last) { return scicone_rng::discrete_distribution<>(first, last)(gen); }
    double normal(double mean, double sigma) { return scicone_rng::normal_distribution<double>(mean, sigma)(gen); }
};

// lgamma of small integers by the running sum the reference memoises (Lgamma.cpp:10-33): lg(0)=inf, lg(1)=lg(2)=0,
// lg(n) = lg(n-1) + log(n-1).  Kept as a running sum (not ::lgamma) so that priors round identically.  One table per
// thread: chains are advanced by several host threads at once, and a shared table that grows under a reader is a heap
// corruption (seen as a rare "double free or corruption" in sum-scoring runs); the running sum is the same in every thread.
inline double lgamma_int(unsigned long n) {
    thread_local std::vector<double> table = {std::numeric_limits<double>::infinity(), 0.0, 0.0};
    while (table.size() - 1 < n) table.push_back(table.back() + std::log(static_cast<double>(table.size() - 1)));
    return table[n];
}
inline double log_n_choose_k(unsigned long n, unsigned long k) { return lgamma_int(n + 1) - lgamma_int(k + 1) - lgamma_int(n - k + 1); }

inline bool all_zero(const EventMap& m) {
    for (auto const& kv : m)
        if (kv.second != 0) return false;
    return true;
}

// key-wise a - b over the union of keys, zeros dropped (reference: Utils::map_diff, Utils.cpp:190-213)
inline EventMap map_diff(const EventMap& a, const EventMap& b) {
    EventMap out;
    auto ia = a.begin();
    auto ib = b.begin();
    while (ia != a.end() || ib != b.end()) {
        unsigned k;
        int va = 0, vb = 0;
        if (ib == b.end() || (ia != a.end() && ia->first < ib->first)) {
            k = ia->first;
            va = (ia++)->second;
        } else if (ia == a.end() || ib->first < ia->first) {
            k = ib->first;
            vb = (ib++)->second;
        } else {
            k = ia->first;
            va = (ia++)->second;
            vb = (ib++)->second;
        }
        if (va - vb != 0) out[k] = va - vb;
    }
    return out;
}

struct Node {
    int id = 0;
    EventMap c;                   // genotype relative to the neutral state (Node::c)
    mutable EventMap c_change;    // events relative to the parent (Node::c_change)
    mutable SmallMap<unsigned, std::pair<int, int>, 4> event_blocks;  // filled by event_prior(), read by expand/shrink
    unsigned n_desc = 1;          // subtree size including the node (Node::n_descendents); may be stale, as in the reference
    int first_child = -1, next = -1, parent = -1;  // indices into Tree::pool
};

class Tree {
public:
    std::vector<Node> pool;      // node storage; indices are stable while the tree lives
    std::vector<int> order;      // ORDER: the reference's all_nodes_vec (insertion order, root first)
    int root = 0;
    unsigned n_nodes = 0;        // nodes besides the root
    int counter = 0;             // last id handed out
    unsigned n_regions = 0;
    int ploidy = 2;
    std::vector<int> neutral;
    const Params* P = nullptr;
    Rng* rng = nullptr;
    double posterior_score = 0.0, prior_score = 0.0, od_score = 0.0, total_attachment_score = 0.0;
    double nu = 1.0;
    bool dirty = false;          // set by a move right before its first change: an untouched copy need not be re-made after a failed attempt

    Tree() {}
    Tree(int ploidy_, unsigned n_regions_, const std::vector<int>& neutral_, const Params* p, Rng* r)
        : n_regions(n_regions_), ploidy(ploidy_), neutral(neutral_), P(p), rng(r) {
        pool.emplace_back();
        order.push_back(0);
        nu = 1.0 / static_cast<double>(n_regions_);  // Tree.h:311
    }

    Node& at(int i) { return pool[i]; }
    const Node& at(int i) const { return pool[i]; }
    unsigned size() const { return static_cast<unsigned>(order.size()); }

    // ---- structure -------------------------------------------------------------------------------------------
    int n_children(int n) const {
        int c = 0;
        for (int t = pool[n].first_child; t != -1; t = pool[t].next) ++c;
        return c;
    }
    std::vector<int> children(int n) const {
        std::vector<int> out;
        for (int t = pool[n].first_child; t != -1; t = pool[t].next) out.push_back(t);
        return out;
    }
    // ORDER: stack walk, parent before child, LAST child first (Node::get_descendents, Node.h:129-152)
    std::vector<int> descendents(int n, bool with_n = true) const {
        std::vector<int> out, stk(1, n);
        while (!stk.empty()) {
            int top = stk.back();
            stk.pop_back();
            for (int t = pool[top].first_child; t != -1; t = pool[t].next) stk.push_back(t);
            out.push_back(top);
        }
        if (!with_n) out.erase(out.begin());
        return out;
    }
    bool is_ancestor(int target, int curr) const {  // Tree.h:857-875 (the root is never reported)
        for (int p = curr; pool[p].parent != -1; p = pool[p].parent)
            if (p == target) return true;
        return false;
    }
    int find_node(int id) const {
        for (int n : descendents(root, true))
            if (pool[n].id == id) return n;
        return -1;
    }

    // genotype of `node` from its parent's (Tree::update_label, Tree.h:460-488)
    void update_label(int node) {
        Node& nd = pool[node];
        nd.c = pool[nd.parent].c;
        for (auto const& ev : nd.c_change) {
            auto it = nd.c.find(ev.first);
            int v = (it == nd.c.end() ? 0 : it->second) + ev.second;
            if (v == 0) {
                if (it != nd.c.end()) nd.c.erase(it);
            } else if (it == nd.c.end())
                nd.c[ev.first] = v;
            else
                it->second = v;
        }
    }
    void update_desc_labels(int node) {  // Tree.h:943-964
        for (int n : descendents(node, true))
            if (n != root) update_label(n);
    }
    // ORDER: appended behind the last sibling and at the end of `order` (Tree::insert_child, Tree.h:348-383)
    int attach(int pos, int node) {
        Node& nd = pool[node];
        nd.parent = pos;
        if (nd.id == 0) nd.id = ++counter;
        order.push_back(node);
        ++n_nodes;
        if (pool[pos].first_child == -1)
            pool[pos].first_child = node;
        else {
            int t = pool[pos].first_child;
            while (pool[t].next != -1) t = pool[t].next;
            pool[t].next = node;
        }
        update_label(node);
        return node;
    }
    int new_child(int pos, const EventMap& events) {  // Tree::insert_child(Node*, map&&), Tree.h:405-420
        pool.emplace_back();
        int n = static_cast<int>(pool.size()) - 1;
        pool[n].c_change = events;
        pool[n].c = events;
        return attach(pos, n);
    }
    int detach(int pos) {  // Tree::prune, Tree.h:624-664: unlinks, keeps the node's own subtree
        if (pos == root) throw std::logic_error("root cannot be pruned!");
        Node& nd = pool[pos];
        Node& par = pool[nd.parent];
        if (par.first_child == pos)
            par.first_child = nd.next;
        else
            for (int t = par.first_child; t != -1; t = pool[t].next)
                if (pool[t].next == pos) {
                    pool[t].next = nd.next;
                    break;
                }
        order.erase(std::remove(order.begin(), order.end(), pos), order.end());
        nd.next = -1;
        --n_nodes;
        return pos;
    }
    int remove_node(int node) {  // Tree::delete_node, Tree.h:1625-1651: children move up, in sibling order
        detach(node);
        int par = pool[node].parent;
        for (int ch : children(node)) {
            detach(ch);
            attach(par, ch);
        }
        pool[node].first_child = -1;
        return par;
    }
    void compute_weights() {  // Tree.h:784-819
        std::vector<int> walk = descendents(root, true);
        for (auto it = walk.rbegin(); it != walk.rend(); ++it) {
            unsigned n = 1;
            for (int t = pool[*it].first_child; t != -1; t = pool[t].next) n += pool[t].n_desc;
            pool[*it].n_desc = n;
        }
    }

    // ORDER: a copy re-inserts the nodes in stack order (last child first), so every sibling list comes out
    // reversed and `order` becomes that walk (Tree::copy_tree / copy_tree_nodes, Tree.h:490-542).
    void copy_from(const Tree& src) {
        if (this == &src) return;
        dirty = false;
        neutral = src.neutral;
        ploidy = src.ploidy;
        total_attachment_score = src.total_attachment_score;
        prior_score = src.prior_score;
        posterior_score = src.posterior_score;
        od_score = src.od_score;
        n_regions = src.n_regions;
        nu = src.nu;
        P = src.P;
        rng = src.rng;
        pool.clear();
        order.clear();
        pool.reserve(src.order.size() + 8);
        n_nodes = 0;
        counter = src.counter;
        auto clone = [&](int s) {
            pool.emplace_back();
            Node& d = pool.back();
            const Node& n = src.pool[s];
            d.id = n.id;
            d.c = n.c;
            d.c_change = n.c_change;
            d.event_blocks = n.event_blocks;
            d.n_desc = n.n_desc;
            return static_cast<int>(pool.size()) - 1;
        };
        root = clone(src.root);
        order.push_back(root);
        std::vector<std::pair<int, int>> todo;  // (destination parent, source node)
        for (int t = src.pool[src.root].first_child; t != -1; t = src.pool[t].next) todo.emplace_back(root, t);
        while (!todo.empty()) {
            std::pair<int, int> top = todo.back();
            todo.pop_back();
            int d = attach(top.first, clone(top.second));
            for (int t = src.pool[top.second].first_child; t != -1; t = src.pool[t].next) todo.emplace_back(d, t);
        }
        n_nodes = src.n_nodes;
    }

    // ---- validity --------------------------------------------------------------------------------------------
    bool subtree_out_of_bound(int n) const {  // Tree.h:1136-1153
        for (int d : descendents(n, true))
            for (auto const& kv : pool[d].c)
                if (kv.second < -neutral[kv.first] || kv.second > P->copy_number_limit) return true;
        return false;
    }
    bool region_changes(int n, unsigned region) const {  // Tree.h:1183-1195 (a stored zero counts)
        for (int d : descendents(n, true))
            if (pool[d].c_change.count(region)) return true;
        return false;
    }
    bool zero_ploidy_changes(int n) const {  // Tree.h:1156-1180
        for (int d : descendents(n, true)) {
            if (pool[d].id == 0) continue;
            for (auto const& kv : pool[pool[d].parent].c)
                if (kv.second <= -neutral[kv.first] && region_changes(d, kv.first)) return true;
        }
        return false;
    }
    bool is_valid_subtree(int n) const { return !subtree_out_of_bound(n) && !zero_ploidy_changes(n); }
    // two nodes with the same genotype (Tree::is_redundant, Tree.h:1198-1223; the reference compares 64-bit hashes of
    // the genotype first and the maps on a hit - equal maps always hit, so the exact comparison decides)
    bool is_redundant() const {
        std::vector<const EventMap*> g;
        g.reserve(order.size());
        for (int n : order) g.push_back(&pool[n].c);
        std::sort(g.begin(), g.end(), [](const EventMap* a, const EventMap* b) { return *a < *b; });
        for (size_t i = 1; i < g.size(); ++i)
            if (*g[i - 1] == *g[i]) return true;
        return false;
    }

    // ---- sampling --------------------------------------------------------------------------------------------
    int uniform_sample(bool with_root) const {  // Tree.h:322-346
        int pick = 0;
        if (order.empty()) throw std::length_error("length of nodes must be bigger than zero, in order to sample from the tree");
        if (order.size() == 1) {
            if (!with_root) throw std::logic_error("root is the only node and cannot be sampled");
        } else
            pick = rng->uniform_int(with_root ? 0 : 1, static_cast<int>(order.size()) - 1);
        return order[pick];
    }
    int weighted_sample() const {  // Tree.h:822-854: weight 1/n_desc in single precision
        if (order.empty()) throw std::length_error("length of nodes must be bigger than zero, in order to sample from the tree");
        if (order.size() == 1) throw std::length_error("there is only 1 element which is the root and root cannot be sampled");
        std::vector<float> w;
        for (size_t i = 1; i < order.size(); ++i) w.push_back(1.0f / pool[order[i]].n_desc);
        return order[1 + rng->discrete(w.begin(), w.end())];
    }
    double cost() const {  // Tree.h:1754-1771 (float weights)
        double zeta = 0.0;
        for (int n : descendents(root, false)) {
            float w = static_cast<float>(1.0 / pool[n].n_desc);
            zeta += w;
        }
        return zeta;
    }

    // ---- priors ----------------------------------------------------------------------------------------------
    // Event prior of one node: -v log(2K) - c_penalise * (undoing blocks), where v counts the copy-number steps needed
    // to draw the node's events as blocks over adjacent regions; also records those blocks for the expand/shrink move
    // (Node::compute_event_prior, Node.h:183-354).
    double node_event_prior(int n) const {
        const Node& nd = pool[n];
        const EventMap& ev = nd.c_change;
        const EventMap& parent_c = pool[nd.parent].c;
        EventMap undoing(ev);
        nd.event_blocks.clear();
        std::vector<int> open;  // start regions of the blocks currently open (a stack)
        int block = 0, v = 0, prev_val = 0, prev_region = -1;
        auto close_all = [&](int end_region) {
            while (!open.empty()) {
                nd.event_blocks[block++] = std::make_pair(open.back(), end_region);
                open.pop_back();
            }
        };
        const unsigned last_region = ev.rbegin()->first;
        for (auto const& e : ev) {
            auto pc = parent_c.find(e.first);
            if (pc == parent_c.end() || std::signbit(e.second) == std::signbit(pc->second)) undoing.erase(e.first);
            if (static_cast<int>(e.first) - 1 != prev_region) {  // not adjacent to the previous event region
                if (0 - prev_val > 0) v += 0 - prev_val;
                close_all(prev_region);
                prev_val = 0;
            }
            if (e.second - prev_val > 0) v += e.second - prev_val;
            int carried = prev_val;
            if (carried != 0 && std::signbit(carried * e.second)) {  // sign flip between adjacent regions
                carried = 0;
                close_all(prev_region);
            }
            int remaining = std::abs(e.second) - std::abs(carried);
            for (; remaining > 0; --remaining) open.push_back(static_cast<int>(e.first));
            for (; remaining < 0; ++remaining) {
                nd.event_blocks[block++] = std::make_pair(open.back(), static_cast<int>(e.first) - 1);
                open.pop_back();
            }
            prev_val = e.second;
            prev_region = static_cast<int>(e.first);
            if (e.first == last_region) {
                if (0 - prev_val > 0) v += 0 - prev_val;
                close_all(prev_region);
            }
        }
        // blocks that undo what the parent carries, counted regardless of magnitude
        int repeats = 0, r_prev = 0, r_prev_region = -1;
        const unsigned last_undo = undoing.empty() ? 0u : undoing.rbegin()->first;
        for (auto const& e : undoing) {
            (void)parent_c.at(e.first);
            if (static_cast<int>(e.first) - 1 != r_prev_region) {
                if (0 - r_prev > 0) repeats += 1;
                r_prev = 0;
            }
            if (e.second - r_prev > 0) repeats += 1;
            r_prev = e.second;
            r_prev_region = static_cast<int>(e.first);
            if (e.first == last_undo && 0 - r_prev > 0) repeats += 1;
        }
        double pv = 0.0;
        const int K = static_cast<int>(n_regions);
        pv -= v * std::log(2 * K);
        pv -= P->c_penalise * repeats;
        return pv;
    }
    double event_prior() const {  // Tree::event_prior, Tree.h:1813-1839
        std::vector<double> pv;
        for (size_t i = 1; i < order.size(); ++i) pv.push_back(node_event_prior(order[i]));
        if (static_cast<size_t>(n_nodes) != pv.size())
            throw std::logic_error("number of nodes needs to be equal to the size of the event prior vector");
        double total = 0.0;
        total += std::accumulate(pv.begin(), pv.end(), 0.0);
        total -= lgamma_int(n_nodes + 1);
        return total;
    }

    // ---- move weights (Tree.h:1654-1751, MathOp.cpp:401-437, :545-600) ---------------------------------------------
    double omega_insert_delete_node(int n) const {
        const EventMap& ev = pool[n].c_change;
        const int r_i = static_cast<int>(ev.size());
        if (r_i == 0) return 0.0;
        if (static_cast<unsigned long>(r_i) > n_regions) throw std::logic_error("There are not r distinct regions to sample");
        double w = 0.0;
        w += std::log(P->lambda_r) * (r_i - 1) + (-1 * P->lambda_r);
        double sum_minus = 0.0;
        for (auto const& e : ev) sum_minus += std::abs(e.second) - 1;
        w += std::log(P->lambda_c) * sum_minus;
        w += (-1 * r_i * P->lambda_c);
        w -= std::log(2) * r_i;
        w -= log_n_choose_k(n_regions, r_i);
        w -= lgamma_int(r_i);
        double lg = 0.0;
        for (auto const& e : ev) lg += lgamma_int(std::abs(e.second));
        w -= lg;
        return std::exp(w);
    }
    double omega_condense_split_node(int n) const {
        const EventMap& child = pool[n].c_change;
        const EventMap& par = pool[pool[n].parent].c_change;
        double w = 1.0;
        std::vector<int> merged;
        for (unsigned i = 0; i < n_regions; ++i) {
            if (all_zero(par)) return 0;  // the parent is the root
            auto ic = child.find(i);
            auto ip = par.find(i);
            const int c_val = ic == child.end() ? 0 : ic->second;
            const int p_val = ip == par.end() ? 0 : ip->second;
            if (p_val == 0 && c_val == 0) continue;
            if (p_val + c_val == 0) return 0;  // the two nodes cancel in this region
            const int c = std::abs(c_val - p_val);
            double f = 1.0;
            if (c == 0)
                f *= std::exp(-1 * P->lambda_s);
            else if (c % 2 == 0) {
                f *= std::pow(P->lambda_s, c / 2);
                f *= std::exp(-1 * P->lambda_s);
                f /= 2 * std::exp(lgamma_int(c / 2 + 1));
            } else {
                f *= std::pow(P->lambda_s, (c - 1) / 2);
                f *= std::exp(-1 * P->lambda_s);
                f /= 2 * std::exp(lgamma_int((c - 1) / 2 + 1));
            }
            w *= f;
            merged.push_back(c_val + p_val);
        }
        if (merged.size() == 1 && merged[0] == 1) return 0;
        return w;
    }
    std::vector<double> chi_insert_delete(bool weighted) const {
        std::vector<double> chi;
        for (int n : order)
            chi.push_back(weighted ? std::pow(2, n_children(n) + 1) / (pool[n].n_desc + 1) : std::pow(2, n_children(n)));
        return chi;
    }
    std::vector<double> omega_insert_delete(bool weighted, bool max_scoring) const {
        std::vector<double> omega;
        for (int n : descendents(root, false)) {
            double w = 1.0;
            if (!max_scoring) w = omega_insert_delete_node(n);
            if (weighted) w = w / pool[n].n_desc;
            omega.push_back(w);
        }
        return omega;
    }
    std::vector<double> chi_condense_split(bool weighted) const {  // nodes with a single event are skipped (quirk Q8)
        std::vector<double> chi;
        for (int n : descendents(root, false)) {
            if (pool[n].c_change.size() <= 1) continue;
            double v = std::pow(2, n_children(n));
            if (weighted) v /= (pool[n].n_desc + 1);
            chi.push_back(v);
        }
        return chi;
    }
    std::vector<double> omega_condense_split(bool weighted, bool max_scoring) const {
        std::vector<double> omega;
        for (int n : descendents(root, false)) {
            double w = 1.0;
            if (!max_scoring) w = omega_condense_split_node(n);
            if (weighted) w /= (pool[pool[n].parent].n_desc - 1);
            omega.push_back(w);
        }
        return omega;
    }

    // ---- moves: each returns the node(s) whose subtree must be rescored, -1 / empty when the proposal is void ----
    // Random event set for a new node (Utils::random_initialize_labels_map, Utils.cpp:31-72)
    EventMap random_labels(int max_regions_per_node, double lambda_r) {
        EventMap out;
        int r = std::min(max_regions_per_node, rng->poisson(lambda_r) + 1);
        if (r > static_cast<int>(n_regions)) throw std::out_of_range("There cannot be more than n_regions amount of distinct regions");
        for (int got = 0; got < r;) {
            int k = rng->uniform_int(0, static_cast<int>(n_regions) - 1);
            if (!out.count(k)) {
                out[k] = rng->coin() ? 1 : -1;
                ++got;
            }
        }
        return out;
    }

    // A move that cannot be applied to the nodes it drew.  The reference throws std::logic_error at these points and its
    // retry loop (Inference::apply_multiple_times, Inference.h:1462-1498) catches it, undoes t_prime and tries again -
    // exactly what it does when a move returns false, so the failure is RETURNED here (same draws, same control flow):
    // unwinding an exception costs ~5 us per attempt, and a chain whose 50 attempts all fail sits that much longer
    // outside the launch pipeline.  The reference's message stays as the argument.
    static int rejected(const char* /*why*/) { return -1; }
    static std::vector<int> rejected_nodes(const char* /*why*/) { return std::vector<int>(); }

    int prune_reattach(bool weighted) {  // Tree.h:544-621
        if (order.size() <= 2) throw InvalidMove("prune and reattach move does not make sense when there is only one node besides the root");
        int pruned = weighted ? weighted_sample() : uniform_sample(false);
        std::vector<int> dest = order;
        for (int d : descendents(pruned, true)) dest.erase(std::remove(dest.begin(), dest.end(), d), dest.end());
        int target = dest[rng->uniform_int(1, static_cast<int>(dest.size())) - 1];
        if (pool[pool[pruned].parent].id == pool[target].id) return -1;
        dirty = true;
        detach(pruned);
        attach(target, pruned);
        update_desc_labels(pruned);
        if (!is_valid_subtree(pruned) || is_redundant()) return -1;
        compute_weights();
        return pruned;
    }

    std::vector<int> swap_labels(bool weighted) {  // Tree.h:878-940
        if (order.size() <= 2) throw InvalidMove("swapping labels does not make sense when they is only one node besides the root");
        int a, b;
        if (weighted) {
            a = weighted_sample();
            do b = weighted_sample();
            while (a == b);
        } else {
            a = uniform_sample(false);
            do b = uniform_sample(false);
            while (a == b);
        }
        if (pool[a].c_change == pool[b].c_change)
            return rejected_nodes("swapping 2 nodes with the same labels does not make sense, the move will be rejected");
        if (pool[a].parent == pool[b].parent && pool[a].first_child == -1 && pool[b].first_child == -1)
            return rejected_nodes("swapping the sibling leaves will have no impact, move will be rejected.");
        dirty = true;
        pool[a].c_change.swap(pool[b].c_change);
        std::vector<int> out;
        if (is_ancestor(a, b))
            out.push_back(a);
        else if (is_ancestor(b, a))
            out.push_back(b);
        else {
            out.push_back(a);
            out.push_back(b);
        }
        for (int n : out) {
            update_desc_labels(n);
            if (!is_valid_subtree(n)) throw InvalidTree("Swap labels move created an invalid tree.");
        }
        if (is_redundant()) throw InvalidTree("Swap labels move created a redundant tree");
        return out;
    }

    int add_remove_events(bool weighted) {  // Tree.h:999-1079
        if (order.size() <= 1) throw InvalidMove("Adding or removing events does not make sense when there is 1 node or less. Root has to be neutral.");
        int node = weighted ? weighted_sample() : uniform_sample(false);
        int want = rng->poisson(P->lambda_r) + 1;
        if (want > static_cast<int>(n_regions))
            throw std::logic_error("the number of distinct regions to sample cannot be bigger than the total number of distinct regions available");
        std::set<unsigned> regions;
        while (static_cast<int>(regions.size()) < want) regions.insert(static_cast<unsigned>(rng->uniform_int(0, static_cast<int>(n_regions) - 1)));
        EventMap& ev = pool[node].c_change;
        dirty = true;
        for (unsigned k : regions) {  // ascending region order
            int copies = rng->poisson(P->lambda_c) + 1;
            bool up = rng->coin();
            ev[k] += up ? copies : -copies;
            if (ev.at(k) == 0) ev.erase(k);
        }
        if (all_zero(ev)) return -1;
        update_desc_labels(node);
        if (!is_valid_subtree(node) || is_redundant()) return -1;
        return node;
    }

    // Grow or shrink one recorded event block by one region at either end (Node::expand_shrink_block, Node.h:383-436).
    // The reference indexes the maps with operator[], which plants zero-valued entries; the same accesses are made here
    // because those entries take part in later decisions (quirk Q7).
    void expand_shrink_block(int n, int block_id, bool expand, bool from_end) {
        Node& nd = pool[n];
        EventMap& ev = nd.c_change;
        const int start = nd.event_blocks[block_id].first;
        const int end = nd.event_blocks[block_id].second;
        if (expand) {
            if (from_end) {
                if (static_cast<unsigned>(end) == n_regions - 1) throw std::logic_error("Can not expand beyond the final region");
                (void)ev[end + 1];  // touched by the reference's (never firing) neighbour test
                (void)ev[end];
                int current = ev[end + 1];
                int step = std::signbit(ev[end]) ? -1 : 1;
                ev[end + 1] = current + step;
            } else {
                if (start == 0) throw std::logic_error("Can not expand backwards beyond the first region");
                if (ev[start - 1] * ev[start] == -1) throw std::logic_error("Can not expand onto the previous block");
                ev[start - 1] = ev[start];
            }
        } else {
            const int region = from_end ? end : start;
            int current = ev[region];
            int step = std::signbit(ev[region]) ? 1 : -1;
            ev[region] = current + step;
            if (ev.at(region) == 0) ev.erase(region);
        }
    }
    int expand_shrink_blocks(bool weighted) {  // Tree.h:1082-1132
        if (order.size() <= 1) throw InvalidMove("Adding or removing events does not make sense when there is 1 node or less. Root has to be neutral.");
        int node = weighted ? weighted_sample() : uniform_sample(false);
        // with no recorded blocks the upper bound wraps to -1 and the draw returns a raw engine word, as in the reference
        int block = rng->uniform_int(0, static_cast<int>(pool[node].event_blocks.size() - 1));
        bool from_end = rng->coin();
        bool expand = rng->coin();
        dirty = true;
        expand_shrink_block(node, block, expand, from_end);
        if (all_zero(pool[node].c_change)) return -1;
        update_desc_labels(node);
        if (!is_valid_subtree(node) || is_redundant()) return -1;
        return node;
    }

    int insert_delete_node(unsigned size_limit, bool weighted, bool max_scoring) {  // Tree.h:1226-1322
        std::vector<double> chi = chi_insert_delete(weighted);
        std::vector<double> omega = omega_insert_delete(weighted, max_scoring);
        int rescore;
        if (rng->uniform_real01() < 0.5) {  // insert
            if (order.size() >= size_limit) return rejected("Tree size limit is reached, insert node move will be rejected!");
            int parent = order[rng->discrete(chi.begin(), chi.end())];
            EventMap labels = random_labels(1, 0.1);  // default arguments of the reference, not the chain's lambda_r (quirk Q14)
            std::vector<int> kids = children(parent);
            dirty = true;
            int fresh = new_child(parent, labels);
            for (int k : kids)
                if (rng->coin()) {
                    detach(k);
                    attach(fresh, k);
                }
            rescore = fresh;
        } else {  // delete
            if (order.size() <= 1) return rejected("Root cannot be deleted, delete move will be rejected");
            int idx = rng->discrete(omega.begin(), omega.end());
            dirty = true;
            rescore = remove_node(descendents(root, false)[idx]);
        }
        update_desc_labels(rescore);
        if (!is_valid_subtree(rescore) || is_redundant()) return -1;
        compute_weights();
        return rescore;
    }

    int condense_split_node(unsigned size_limit, bool weighted, bool max_scoring) {  // Tree.h:1325-1462
        if (order.size() <= 1) throw InvalidMove("condense or split does not make sense when there is 1 node or less. ");
        std::vector<double> chi = chi_condense_split(weighted);
        std::vector<double> omega = omega_condense_split(weighted, max_scoring);
        double u = rng->uniform_real01();
        std::vector<int> below_root = descendents(root, false);
        int rescore;
        if (u < 0.5) {  // split
            if (order.size() >= size_limit) return rejected("Tree size limit is reached, split move will be rejected!");
            int node = below_root[rng->discrete(chi.begin(), chi.end())];  // index into the unfiltered walk (quirk Q8)
            if (pool[node].c_change.size() == 1) return rejected("Nodes with single events in a single region cannot be split");
            EventMap upper, lower;
            for (auto const& e : pool[node].c_change) {
                bool sign = rng->coin();
                int beta = 2 * sign - 1;
                int lam = rng->poisson(P->lambda_s);
                int v = e.second;
                if (v % 2 == 0)
                    upper[e.first] = static_cast<int>(v / 2 + beta * (0.5 + lam));
                else
                    upper[e.first] = v / 2 + beta * lam;
                lower[e.first] = v - upper[e.first];
                if (upper.at(e.first) == 0) upper.erase(e.first);
                if (lower.at(e.first) == 0) lower.erase(e.first);
            }
            dirty = true;
            pool[node].c_change = upper;
            std::vector<int> kids = children(node);
            int fresh = new_child(node, lower);
            for (int k : kids)
                if (rng->coin()) {
                    detach(k);
                    attach(fresh, k);
                }
            rescore = node;
        } else {  // condense: delete a node and push its events into the parent
            if (order.size() <= 2) return rejected("condensing nodes does not make sense when there are 2 or less nodes. Root has to be neutral.");
            int node = below_root[rng->discrete(omega.begin(), omega.end())];
            int par = pool[node].parent;
            if (all_zero(pool[node].c_change) || all_zero(pool[par].c_change))
                return rejected("cannot condense a node with an empty node (root)");
            EventMap merged = pool[par].c_change;
            for (auto const& e : pool[node].c_change) {
                int v = e.second;
                auto it = merged.find(e.first);
                if (it != merged.end()) v += it->second;
                if (v == 0) {
                    if (it != merged.end()) merged.erase(it);
                } else
                    merged[e.first] = v;
            }
            if (merged.size() == 1 && merged.begin()->second == 1)
                return rejected("Nodes cannot be combined if they result in a single event");
            dirty = true;
            pool[par].c_change = merged;
            rescore = remove_node(node);
        }
        update_desc_labels(rescore);
        if (!is_valid_subtree(rescore) || is_redundant()) return -1;
        compute_weights();
        return rescore;
    }

    // Events two siblings share (same region, same sign, smaller magnitude) - Tree::get_event_intersection, Tree.h:1587-1622
    EventMap shared_events(int a, int b) const {
        EventMap out;
        for (auto const& e : pool[a].c_change) {
            auto it = pool[b].c_change.find(e.first);
            if (it == pool[b].c_change.end() || std::signbit(it->second) != std::signbit(e.second)) continue;
            int mag = std::min(std::abs(e.second), std::abs(it->second));
            out[e.first] = (std::signbit(e.second) ? -1 : 1) * mag;
        }
        return out;
    }
    int add_common_ancestor() {  // Tree.h:1464-1585
        if (order.size() <= 1) throw InvalidMove("add common ancestor does not make sense when there is 1 node or less. ");
        int parent = uniform_sample(true);
        std::vector<int> sibs = children(parent);
        if (sibs.size() < 2) return rejected("Can not create common ancestor from less than 2 siblings.");
        std::vector<int> idx(sibs.size());
        std::iota(idx.begin(), idx.end(), 0);  // sibling positions double as sampling weights (quirk Q9)
        int ia = rng->discrete(idx.begin(), idx.end());
        int a = sibs[ia];
        idx.erase(idx.begin() + ia);
        sibs.erase(sibs.begin() + ia);
        int b = sibs[rng->discrete(idx.begin(), idx.end())];
        EventMap common = shared_events(a, b);
        if (common.empty()) return rejected("Can not create common ancestor from siblings with no intersection.");
        dirty = true;
        for (auto const& e : common)
            for (int n : {a, b}) {
                EventMap& ev = pool[n].c_change;
                ev[e.first] = ev[e.first] - e.second;
                if (ev.at(e.first) == 0) ev.erase(e.first);
            }
        for (int n : {a, b})
            if (pool[n].c_change.empty()) return rejected("Can not create common ancestor if a child would become empty.");
        int anc = new_child(parent, common);
        for (int n : {a, b}) {
            detach(n);
            attach(anc, n);
        }
        if (!is_valid_subtree(parent) || is_redundant()) return -1;
        return anc;  // weights are deliberately left stale, as in the reference
    }

    int delete_leaf() {  // Tree.h:2011-2038
        if (order.size() <= 2) throw InvalidMove("delete leaf move does not make sense when there is only one node besides the root");
        std::vector<int> leaves;
        for (int n : order)
            if (pool[n].first_child == -1) leaves.push_back(n);
        const int victim = leaves[rng->uniform_int(0, static_cast<int>(leaves.size()) - 1)];
        dirty = true;
        return remove_node(victim);
    }

    // Gibbs move over all re-attachments that keep every genotype (Tree.h:1842-1999).  Acts on the committed tree and
    // needs no rescoring: only event priors change.
    void genotype_preserving_prune_reattach(double gamma) {
        if (order.size() <= 2) throw InvalidMove("prune and reattach move does not make sense when there is only one node besides the root");
        std::vector<double> scores(1, 0.0);
        std::vector<std::pair<int, int>> moves(1, std::make_pair(0, 0));  // (id to prune, id to attach to)
        for (size_t i = 1; i < order.size(); ++i) {
            const int cand = order[i];
            const double before = node_event_prior(cand);
            std::vector<int> dest = order;
            for (int d : descendents(cand, true)) dest.erase(std::remove(dest.begin(), dest.end(), d), dest.end());
            for (int target : dest) {
                if (pool[pool[cand].parent].id == pool[target].id) continue;
                EventMap rel = map_diff(pool[cand].c, pool[target].c);
                if (rel.empty()) continue;
                // a stand-in leaf with the candidate's genotype below the target: prior and the zero-ploidy rule
                pool.emplace_back();
                const int tmp = static_cast<int>(pool.size()) - 1;
                pool[tmp].id = pool[cand].id;
                pool[tmp].c = pool[cand].c;
                pool[tmp].c_change = rel;
                pool[tmp].parent = target;
                const double after = node_event_prior(tmp);
                attach(target, tmp);
                const bool bad = zero_ploidy_changes(target);
                detach(tmp);
                pool.pop_back();
                if (bad) continue;
                moves.emplace_back(pool[cand].id, pool[target].id);
                scores.push_back((after - before) * gamma);
            }
        }
        const double top = *std::max_element(scores.begin(), scores.end());
        for (double& s : scores) s = std::exp(s - top);
        const int pick = rng->discrete(scores.begin(), scores.end());
        if (pick == 0) throw std::logic_error("The node is reattached to the same position.");
        int node = find_node(moves[pick].first);
        if (node < 0) throw std::logic_error("Node to prune could not be found in the tree. Move will be rejected.");
        detach(node);
        int target = find_node(moves[pick].second);
        if (target < 0) throw std::logic_error("Node to attach could not be found in the tree. Move will be rejected.");
        pool[node].c_change = map_diff(pool[node].c, pool[target].c);
        if (pool[node].c_change.empty()) throw std::logic_error("Empty c_label node created. Move will be rejected.");
        attach(target, node);
        compute_weights();
    }

    // ---- text form (Tree.h:121-142, Node.h:83-106; loader Tree.h:686-782) ----------------------------------------------
    void write(std::ostream& os) const {
        const int pp = P->print_precision;
        os << "Tree posterior: " << std::setprecision(pp) << posterior_score << std::endl;
        os << "Tree prior: " << std::setprecision(pp) << prior_score << std::endl;
        os << "Event prior: " << std::setprecision(pp) << event_prior() << std::endl;
        os << "Log likelihood: " << std::setprecision(pp) << total_attachment_score << std::endl;
        os << "Root score: " << std::setprecision(pp) << od_score << std::endl;
        os << "Tree score: " << std::setprecision(pp) << posterior_score + od_score << std::endl;
        os << "Nu: " << std::setprecision(pp) << nu << std::endl;
        for (int n : descendents(root, true)) {
            const Node& nd = pool[n];
            os << "node " << nd.id << ": ";
            if (nd.parent == -1)
                os << "p_id:NULL";
            else
                os << "p_id:" << pool[nd.parent].id;
            os << ",[";
            bool first = true;
            for (auto const& e : nd.c_change) {
                if (!first) os << ',';
                os << e.first << ":" << e.second;
                first = false;
            }
            os << ']' << std::endl;
        }
    }
    void load_from_file(const std::string& path) {
        if (pool[root].first_child != -1) throw std::logic_error("Tree to read into must be empty!");
        std::ifstream in(path);
        std::string line;
        for (int i = 0; i < 7; ++i) std::getline(in, line);  // six score lines, then "Nu: x"
        nu = std::stod(line.substr(line.find(" "), line.size()));
        std::getline(in, line);  // the root
        while (std::getline(in, line)) {
            std::istringstream iss(line);
            std::string a, b, c;
            if (!(iss >> a >> b >> c)) break;
            b.pop_back();
            int node_id = std::stoi(b);
            if (node_id > counter) counter = node_id + 1;
            std::string head = c.substr(0, c.find(",["));
            std::string rest = c.substr(c.find(",[") + 1, c.length());
            if (rest.front() != '[' || rest.back() != ']')
                throw std::runtime_error(" Incorrect tree format to parse! \nThe rows should be space separated.");
            int parent_id = std::stoi(head.substr(5, head.size()));
            std::string body = rest.substr(1, rest.find("]") - 1);
            pool.emplace_back();
            int n = static_cast<int>(pool.size()) - 1;
            pool[n].id = node_id;
            std::stringstream ss(body);
            std::string pair;
            while (std::getline(ss, pair, ',')) {
                unsigned key = static_cast<unsigned>(std::stoi(pair.substr(0, pair.find(':'))));
                pool[n].c_change[key] = std::stoi(pair.substr(pair.find(':') + 1, pair.length()));
            }
            int parent = -1;
            for (int q : order)
                if (pool[q].id == parent_id) {
                    parent = q;
                    break;
                }
            if (parent < 0) throw std::runtime_error("tree file: parent id not seen before its child");
            attach(parent, n);
        }
        if (!is_valid_subtree(root)) throw InvalidTree("The loaded tree is invalid!");
        if (is_redundant()) throw InvalidTree("The loaded tree is redundant!");
        compute_weights();
    }
};

}  // namespace scicone_host
#endif
