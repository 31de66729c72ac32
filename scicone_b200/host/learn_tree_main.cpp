// `learn_tree` - restarts of the tree search spread over the GPUs of one box, best-of-N selection, robustness score and
// retries: what pyscicone's SCICoNE.learn_tree_parallel / the cluster-tree loop of SCICoNE.learn_tree do around the
// `inference` binary (reference: pyscicone/scicone/scicone.py:349-375 and :491-503).
//
//   reference                                                       here
//   Pool(n_reps), one `inference` process per restart, seed 42 + i   one `inference` process PER GPU, its share of the restarts
//                                                                    as chains of that process (--n_chains, seeds 42 + i):
//                                                                    replicas only - no collective, no data exchanged
//   tree.score = the "Tree score:" line of *_tree_inferred.txt       the same line of every restart's file
//   scores[scores == 0.] = -inf; best = argmax                       the same (first maximum)
//   robustness = count(isclose(scores, best, atol=5)) / n_reps       the same: |s - best| <= 5 + 1e-5 |best| (numpy's rtol)
//   while robustness < robustness_thr and tries < max_tries:         the same loop; a retry starts every restart from the best
//       rerun from the best tree, nu = best tree's nu                tree so far (--tree_file, --nu of that file)
//
// Usage: learn_tree --n_reps=10 --n_gpus=8 --max_tries=2 --robustness_thr=0.5 --postfix=P <flags of `inference` ...>
// Every other flag goes to `inference` unchanged (matrix, region sizes, cluster sizes, n_iters, move_probs ...).
// Output: P_learn_tree.json (scores of every try, best restart, robustness), and the files of the best restart copied to
// P_<name> (P_tree_inferred.txt, P_cell_node_ids.tsv, ...).
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Flag {
    std::string key, value;
    bool has_value;
};

std::vector<Flag> parse(int argc, char** argv) {
    std::vector<Flag> out;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        if (s.rfind("--", 0) != 0) continue;
        s = s.substr(2);
        const size_t eq = s.find('=');
        if (eq != std::string::npos)
            out.push_back({s.substr(0, eq), s.substr(eq + 1), true});
        else if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0)
            out.push_back({s, argv[++i], true});
        else
            out.push_back({s, "", false});
    }
    return out;
}

std::string self_dir() {
    char buf[4096];
    const ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    if (n <= 0) return ".";
    buf[n] = 0;
    std::string p(buf);
    const size_t slash = p.rfind('/');
    return slash == std::string::npos ? "." : p.substr(0, slash);
}

// the "Tree score:" and "Nu:" lines of a tree file (Tree::operator<<, scicone/src/Tree.h; parsed by pyscicone tree.py:199-214)
bool read_tree_file(const std::string& path, double* score, double* nu) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    bool have = false;
    while (std::getline(in, line)) {
        if (line.rfind("Tree score:", 0) == 0) {
            *score = std::stod(line.substr(line.rfind(' ') + 1));
            have = true;
        } else if (line.rfind("Nu:", 0) == 0)
            *nu = std::stod(line.substr(line.rfind(' ') + 1));
    }
    return have;
}

void copy_file(const std::string& from, const std::string& to) {
    std::ifstream in(from, std::ios::binary);
    std::ofstream out(to, std::ios::binary | std::ios::trunc);
    out << in.rdbuf();
}

}  // namespace

int main(int argc, char** argv) {
    try {
        std::vector<Flag> flags = parse(argc, argv);
        int n_reps = 10, n_gpus = 1, max_tries = 2, seed_base = 42, first_device = 0;
        double robustness_thr = 0.5;
        std::string postfix, binary = self_dir() + "/inference";
        std::vector<std::string> pass;  // flags of `inference`
        for (const Flag& f : flags) {
            if (f.key == "n_reps") n_reps = std::stoi(f.value);
            else if (f.key == "n_gpus") n_gpus = std::stoi(f.value);
            else if (f.key == "first_device") first_device = std::stoi(f.value);
            else if (f.key == "max_tries") max_tries = std::stoi(f.value);
            else if (f.key == "robustness_thr") robustness_thr = std::stod(f.value);
            else if (f.key == "seed_base") seed_base = std::stoi(f.value);
            else if (f.key == "postfix") postfix = f.value;
            else if (f.key == "inference_binary") binary = f.value;
            else if (f.key == "seed" || f.key == "n_chains" || f.key == "device" || f.key == "rank" || f.key == "world")
                throw std::runtime_error("--" + f.key + " is set by learn_tree itself (restart i runs with seed seed_base + i)");
            else
                pass.push_back(f.has_value ? "--" + f.key + "=" + f.value : "--" + f.key);
        }
        if (n_reps < 1 || n_gpus < 1 || max_tries < 1) throw std::runtime_error("n_reps, n_gpus and max_tries must be positive");
        if (postfix.empty()) throw std::runtime_error("--postfix is required");
        n_gpus = std::min(n_gpus, n_reps);
        const unsigned cores = std::max(1u, std::thread::hardware_concurrency());

        std::ostringstream js;
        js << std::setprecision(17) << "{\"n_reps\":" << n_reps << ",\"n_gpus\":" << n_gpus << ",\"robustness_thr\":" << robustness_thr
           << ",\"max_tries\":" << max_tries << ",\"tries\":[";
        std::string best_prefix, start_tree;
        double best_score = 0.0, best_nu = 1.0, robustness = 0.0;
        int best_rep = -1, cnt = 0;
        while (robustness < robustness_thr) {  // scicone.py:491-503
            if (cnt >= max_tries) break;
            // restart i -> GPU i * n_gpus / n_reps: contiguous blocks, one process per GPU
            std::vector<int> first(n_gpus + 1, 0);
            for (int g = 0; g <= n_gpus; ++g) first[g] = static_cast<int>(static_cast<long long>(n_reps) * g / n_gpus);
            std::vector<pid_t> pids(n_gpus, -1);
            std::vector<std::string> prefix(n_reps);
            for (int g = 0; g < n_gpus; ++g) {
                const int n_here = first[g + 1] - first[g];
                const std::string pf = postfix + "_try" + std::to_string(cnt) + "_g" + std::to_string(g);
                for (int c = 0; c < n_here; ++c) prefix[first[g] + c] = n_here > 1 ? pf + "_c" + std::to_string(c) : pf;
                std::vector<std::string> args{binary};
                args.insert(args.end(), pass.begin(), pass.end());
                args.push_back("--postfix=" + pf);
                args.push_back("--seed=" + std::to_string(seed_base + first[g]));
                args.push_back("--n_chains=" + std::to_string(n_here));
                args.push_back("--device=" + std::to_string(first_device + g));
                if (!start_tree.empty()) {  // a retry: every restart continues from the best tree so far (scicone.py:499, tree.py:330-345)
                    args.push_back("--tree_file=" + start_tree);
                    std::ostringstream nu;
                    nu << std::setprecision(17) << best_nu;
                    args.push_back("--nu=" + nu.str());
                }
                const pid_t pid = fork();
                if (pid < 0) throw std::runtime_error("fork failed");
                if (pid == 0) {
                    if (!getenv("SCICONE_THREADS")) setenv("SCICONE_THREADS", std::to_string(std::max(1u, cores / n_gpus)).c_str(), 1);
                    std::vector<char*> av;
                    for (std::string& s : args) av.push_back(const_cast<char*>(s.c_str()));
                    av.push_back(nullptr);
                    execv(binary.c_str(), av.data());
                    perror("learn_tree: execv of the inference binary failed");
                    _exit(127);
                }
                pids[g] = pid;
            }
            bool failed = false;
            for (int g = 0; g < n_gpus; ++g) {
                int status = 0;
                if (waitpid(pids[g], &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) failed = true;
            }
            if (failed) throw std::runtime_error("an inference process of try " + std::to_string(cnt) + " failed");
            // scicone.py:361-373
            std::vector<double> scores(n_reps, 0.0), nus(n_reps, 1.0);
            for (int i = 0; i < n_reps; ++i)
                if (!read_tree_file(prefix[i] + "_tree_inferred.txt", &scores[i], &nus[i]))
                    throw std::runtime_error("no tree file for restart " + std::to_string(i) + " (" + prefix[i] + "_tree_inferred.txt)");
            std::vector<double> sel = scores;
            for (double& s : sel)
                if (s == 0.0) s = -std::numeric_limits<double>::infinity();  // "just in case"
            int arg = 0;
            for (int i = 1; i < n_reps; ++i)
                if (sel[i] > sel[arg]) arg = i;  // np.argmax: first maximum
            const double best = scores[arg];    // best_tree.score is the file's value, 0 included
            int close = 0;
            for (int i = 0; i < n_reps; ++i)
                if (std::abs(sel[i] - best) <= 5.0 + 1e-5 * std::abs(best)) ++close;  // np.isclose(scores, best, atol=5), rtol = 1e-5
            robustness = static_cast<double>(close) / n_reps;
            best_rep = arg;
            best_score = best;
            best_nu = nus[arg];
            best_prefix = prefix[arg];
            start_tree = best_prefix + "_tree_inferred.txt";
            js << (cnt ? "," : "") << "{\"try\":" << cnt << ",\"scores\":[";
            for (int i = 0; i < n_reps; ++i) js << (i ? "," : "") << scores[i];
            js << "],\"best_rep\":" << arg << ",\"best_score\":" << best << ",\"best_nu\":" << best_nu << ",\"robustness\":" << robustness
               << ",\"best_prefix\":\"" << best_prefix << "\"}";
            ++cnt;
            std::cout << "Cluster tree finished with a robustness score of " << robustness << " after " << cnt << " tries" << std::endl;
        }
        js << "],\"n_tries\":" << cnt << ",\"best_rep\":" << best_rep << ",\"best_score\":" << best_score << ",\"best_nu\":" << best_nu
           << ",\"robustness\":" << robustness << ",\"best_prefix\":\"" << best_prefix << "\"}";
        {
            std::ofstream out(postfix + "_learn_tree.json");
            out << js.str() << std::endl;
        }
        // the files of the best restart under the caller's postfix
        DIR* d = opendir(".");
        if (d) {
            const std::string want = best_prefix + "_";
            std::vector<std::string> names;
            while (dirent* e = readdir(d)) {
                const std::string name = e->d_name;
                if (name.rfind(want, 0) == 0) names.push_back(name);
            }
            closedir(d);
            for (const std::string& name : names) copy_file(name, postfix + "_" + name.substr(want.size()));
        }
        return EXIT_SUCCESS;
    } catch (const std::exception& e) {
        std::cerr << "learn_tree: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
}
