// Host-level known answers for scicone_b200/host/tree.hpp, mirroring the reference's own unit tests that do not touch
// the scoring path (scicone/tests/validation.h): tree-file validity fixtures (test_tree_validation, :999-1063),
// subtree sizes (test_n_descendents_computation, :1065-1080), genotype bookkeeping and copy semantics.
// Driven by tests/test_host_units.py; prints one line per check and exits non-zero on the first failure.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../scicone_b200/host/tree.hpp"

using namespace scicone_host;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);    \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

static Tree load(const std::string& path, Params* P, Rng* rng, bool* threw_invalid) {
    std::vector<int> neutral;  // n_regions = 0 in the reference's test: every region falls back to the ploidy... here 8 regions of 2
    neutral.assign(8, 2);
    Tree t(2, 8, neutral, P, rng);
    *threw_invalid = false;
    try {
        t.load_from_file(path);
    } catch (const InvalidTree&) {
        *threw_invalid = true;
    }
    return t;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    Params P;
    P.copy_number_limit = 5;  // run_tests.cpp:36-45 sets the globals the same way
    Rng rng(42);
    bool bad = false;

    Tree v1 = load(dir + "/valid_tree_1.txt", &P, &rng, &bad);
    CHECK(!bad);
    CHECK(v1.is_valid_subtree(v1.root));
    CHECK(!v1.is_redundant());
    Tree v2 = load(dir + "/valid_tree_2.txt", &P, &rng, &bad);
    CHECK(!bad);
    CHECK(v2.is_valid_subtree(v2.root));
    CHECK(!v2.is_redundant());

    Tree i2 = load(dir + "/invalid_tree_2.txt", &P, &rng, &bad);  // two nodes carry the same genotype
    CHECK(bad);
    CHECK(i2.is_valid_subtree(i2.root));
    CHECK(i2.is_redundant());
    Tree i3 = load(dir + "/invalid_tree_3.txt", &P, &rng, &bad);  // a region comes back from zero copies
    CHECK(bad);
    CHECK(i3.zero_ploidy_changes(i3.root));
    CHECK(!i3.is_redundant());
    Tree i4 = load(dir + "/invalid_tree_4.txt", &P, &rng, &bad);  // out of bounds and back from zero
    CHECK(bad);
    CHECK(i4.subtree_out_of_bound(i4.root));
    CHECK(i4.zero_ploidy_changes(i4.root));
    CHECK(!i4.is_redundant());

    // subtree sizes after compute_weights: every node counts itself and its descendents
    v1.compute_weights();
    for (int n : v1.descendents(v1.root, true)) CHECK(v1.pool[n].n_desc == v1.descendents(n, true).size());

    // a copy reverses every sibling list (Tree.h:523-542, quirk Q6); copying twice restores the child order, and the
    // genotypes / events / ids are preserved either way
    Tree c1(2, 8, std::vector<int>(8, 2), &P, &rng), c2(2, 8, std::vector<int>(8, 2), &P, &rng);
    c1.copy_from(v1);
    c2.copy_from(c1);
    CHECK(c1.n_nodes == v1.n_nodes && c2.n_nodes == v1.n_nodes);
    {
        std::vector<int> a = v1.children(v1.root), b = c1.children(c1.root), c = c2.children(c2.root);
        CHECK(a.size() == b.size() && a.size() == c.size());
        for (size_t i = 0; i < a.size(); ++i) {
            CHECK(v1.pool[a[i]].id == c1.pool[b[b.size() - 1 - i]].id);
            CHECK(v1.pool[a[i]].id == c2.pool[c[i]].id);
        }
    }
    for (int n : v1.descendents(v1.root, true)) {
        const int m = c1.find_node(v1.pool[n].id);
        CHECK(m >= 0);
        if (m >= 0) {
            CHECK(c1.pool[m].c == v1.pool[n].c);
            CHECK(c1.pool[m].c_change == v1.pool[n].c_change);
        }
    }
    // genotype = sum of the events on the path from the root (Tree::update_label, Tree.h:460-488)
    for (int n : v1.descendents(v1.root, false)) {
        EventMap g;
        for (int p = n; p != v1.root; p = v1.pool[p].parent)
            for (auto const& e : v1.pool[p].c_change) g[e.first] += e.second;
        for (auto it = g.begin(); it != g.end();) it = it->second == 0 ? g.erase(it) : it + 1;
        CHECK(g == v1.pool[n].c);
    }
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "ok", failures);
    return failures ? 1 : 0;
}
