#!/usr/bin/env python3
"""Write a patched copy of the reference's Inference.h (seam methods forwarded to libscicone_score.so).

TEST / INTEGRATION INFRASTRUCTURE.  Usage: patch_reference.py <reference scicone/src dir> <build dir>
The patched header goes to <build dir>/Inference.h (a scratch directory outside the repo: reference
sources are never copied into the repository); only the linked binaries land in oracle/_ref/.
The edits are exactly the diff shown in INTEGRATION.md:
  1. the reference's own definitions of the seam methods are renamed ref_<name> (kept for A/B use);
  2. Inference::comparison takes the proposal's weighted total from the engine instead of re-summing
     t_prime_sums (Inference.h:292-294); the move-10 branch likewise for t_sums (:750-752);
  3. assign_cells_to_nodes materialises t_scores from the device table after its full pass;
  4. b200_adapter.inl is included at the end of the header.
"""
import os
import re
import sys

SEAM = ["compute_t_table", "compute_t_od_scores", "compute_t_prime_od_scores", "compute_t_prime_scores",
        "compute_t_prime_sums", "update_t_scores"]


def main():
    src, build = sys.argv[1], sys.argv[2]
    text = open(os.path.join(src, "Inference.h")).read()
    n_total = 0
    for name in SEAM:
        text, n = re.subn(r"^void Inference::%s\(" % name, "void Inference::ref_%s(" % name, text, flags=re.M)
        assert n == 1, (name, n)
        n_total += n
    # declarations of the renamed originals
    decl = "\n".join([
        "    void ref_compute_t_od_scores(const vector<vector<double>> &D, const vector<int> &r, const vector<int> &cluster_sizes);",
        "    void ref_compute_t_prime_od_scores(const vector<vector<double>> &D, const vector<int> &r, const vector<int> &cluster_sizes);",
        "    void ref_compute_t_table(const vector<vector<double>> &D, const vector<int> &r, const vector<int> &cluster_sizes);",
        "    void ref_compute_t_prime_scores(Node *attached_node, const vector<vector<double>> &D, const vector<int> &r);",
        "    void ref_compute_t_prime_sums(const vector<vector<double>> &D);",
        "    void ref_update_t_scores();",
    ])
    text, n = re.subn(r"^(    void update_t_scores\(\);)", r"\1\n" + decl, text, count=1, flags=re.M)
    assert n == 1
    # forward declarations for the adapter's helpers used inside member functions
    text, n = re.subn(r"^class Inference \{", "class Inference;\nnamespace b200 { inline double prime_total(const Inference*); "
                      "inline double t_total(const Inference*); inline void materialise_t_scores(Inference*); }\nclass Inference {",
                      text, count=1, flags=re.M)
    assert n == 1
    # 2. totals come from the engine
    text, n = re.subn(r"for \(int i=0; i<m;i\+\+\)\s*\n\s*t_prime_sum = t_prime_sum \+ t_prime_sums\[i\]\*cluster_sizes\[i\];",
                      "t_prime_sum = b200::prime_total(this);", text)
    assert n == 1, n
    text, n = re.subn(r"for \(int i=0; i<m;i\+\+\)\s*\n\s*t_sum = t_sum \+ t_sums\[i\]\*cluster_sizes\[i\];",
                      "t_sum = b200::t_total(this);", text)
    assert n == 1, n
    # 3. writers at the end of a run read t_scores
    text, n = re.subn(r"(    this->compute_t_table\(D,r,cluster_sizes\);\n)(\n    std::ofstream cell_node_ids_file;)",
                      r"\1    b200::materialise_t_scores(this);\n\2", text)
    assert n == 1, n
    # 4. the adapter
    text, n = re.subn(r"#endif //SC_DNA_INFERENCE_H\s*$", '#include "b200_adapter.inl"\n\n#endif //SC_DNA_INFERENCE_H\n', text)
    assert n == 1
    os.makedirs(build, exist_ok=True)
    open(os.path.join(build, "Inference.h"), "w").write(text)
    for main_cpp, out in (("infer_trees.cpp", "inference_b200.cpp"), ("score_tree.cpp", "score_b200.cpp")):
        with open(os.path.join(build, out), "w") as f:
            f.write('#include "Inference.h"  // patched copy in this directory (same include guard as the original)\n')
            f.write('#include "%s"\n' % os.path.join(src, main_cpp))


if __name__ == "__main__":
    main()
