// `inference` - tree inference by MCMC on the B200 scoring engine.
//
// Keeps the command line, input files and output files of the reference's `inference` binary
// (reference: scicone/src/infer_trees.cpp; formats in SURVEY.md App. D) so that pyscicone and the tutorial
// commands work unchanged.  Additions (all optional): --n_chains=B runs B independent restarts (seeds seed, seed+1, ...)
// in lock-step on one GPU with one kernel launch per step, --device, --stats_file (JSON timing for bench.py),
// and --rank/--world/--nccl_id for cell-sharded multi-GPU runs.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "mcmc.hpp"
#include "csv_reader.hpp"
#include "cnv_writer.hpp"

using namespace scicone_host;

namespace {

struct Args {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string str(const std::string& k, const std::string& d = "") const { return has(k) ? kv.at(k) : d; }
    double num(const std::string& k, double d) const { return has(k) ? std::stod(kv.at(k)) : d; }
    long integer(const std::string& k, long d) const { return has(k) ? std::stol(kv.at(k)) : d; }
    bool flag(const std::string& k, bool d) const {  // cxxopts: (t|T)(rue)?|1 is true, (f|F)(alse)?|0 is false
        if (!has(k)) return d;
        const std::string& v = kv.at(k);
        if (v.empty() || v == "1" || v == "t" || v == "T" || v == "true" || v == "True") return true;
        if (v == "0" || v == "f" || v == "F" || v == "false" || v == "False") return false;
        throw std::runtime_error("Argument '" + v + "' failed to parse for option --" + k);
    }
};

Args parse(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        if (s.rfind("--", 0) != 0) continue;
        s = s.substr(2);
        size_t eq = s.find('=');
        if (eq != std::string::npos)
            a.kv[s.substr(0, eq)] = s.substr(eq + 1);
        else if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0)
            a.kv[s] = argv[++i];
        else
            a.kv[s] = "";
    }
    return a;
}

// Utils::read_counts (Utils.cpp:74-102): comma separated doubles into a pre-sized cells x regions matrix
void read_counts(std::vector<double>& mat, int n_cells, int n_regions, const std::string& path) {
    if (path.size() > 4 && path.compare(path.size() - 4, 4, ".f64") == 0) {
        // binary side format for large matrices: n_cells x n_regions little-endian doubles, row-major (no reference
        // counterpart; the CSV parser below is the reference's format)
        std::ifstream bin(path, std::ios::binary);
        if (!bin) throw std::runtime_error("cannot open " + path);
        bin.read(reinterpret_cast<char*>(mat.data()), static_cast<std::streamsize>(mat.size() * sizeof(double)));
        if (static_cast<size_t>(bin.gcount()) != mat.size() * sizeof(double))
            throw std::runtime_error(path + " holds fewer than n_cells x n_regions doubles");
        return;
    }
    // the reference's format; parsed in parallel on the scoring library's worker pool (csv_reader.hpp)
    read_counts_csv(mat, n_cells, n_regions, path, [](int n, const std::function<void(int)>& fn) { parallel_chains(n, [&fn](int i) { fn(i); }); },
                    scicone_host_threads());
}

// Utils::read_vector (Utils.cpp:141-163): one number per line
void read_vector(std::vector<int>& vec, const std::string& path) {
    std::ifstream in(path);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        vec.push_back(static_cast<int>(std::stod(line)));
    }
    if (vec.empty()) throw std::logic_error("The size of the vector should not be zero!");
}

std::vector<float> parse_floats(const std::string& s) {
    std::vector<float> out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ',')) out.push_back(std::stof(tok));
    return out;
}

void hex_to_bytes(const std::string& hex, uint8_t* out, size_t n) {
    if (hex.size() != 2 * n) throw std::runtime_error("--nccl_id must be 256 hex characters");
    for (size_t i = 0; i < n; ++i) out[i] = static_cast<uint8_t>(std::stoi(hex.substr(2 * i, 2), nullptr, 16));
}

}  // namespace

extern "C" int scicone_inference_main(int argc, char** argv) {
    try {
        Args a = parse(argc, argv);
        Params P;
        int n_iters = static_cast<int>(a.integer("n_iters", 5000));
        int ploidy = static_cast<int>(a.integer("ploidy", 2));
        P.verbosity = static_cast<int>(a.integer("verbosity", 0));
        int seed = static_cast<int>(a.integer("seed", 0));
        unsigned size_limit = a.has("size_limit") ? static_cast<unsigned>(a.integer("size_limit", 0)) : std::numeric_limits<unsigned>::max();
        int n_nodes = static_cast<int>(a.integer("n_nodes", 3));
        P.lambda_r = a.num("lambda_r", 0.1);
        P.lambda_c = a.num("lambda_c", 0.2);
        P.cf = a.num("cf", 0.0);
        P.c_penalise = a.num("c_penalise", 10.0);
        P.is_overdispersed = 1;
        P.eta = a.num("eta", 1e-4);
        P.print_precision = static_cast<int>(a.integer("print_precision", 15));
        P.lambda_s = 0.5;
        P.copy_number_limit = static_cast<int>(a.integer("copy_number_limit", 5));
        P.postfix = a.str("postfix");
        double nu = a.num("nu", 1.0);
        double alpha = a.num("alpha", 0.0);
        double gamma = a.num("gamma", 1.0);
        bool max_scoring = a.flag("max_scoring", true);
        std::vector<float> move_probs = a.has("move_probs")
                                            ? parse_floats(a.str("move_probs"))
                                            : std::vector<float>{0.0f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.01f, 0.1f, 0.01f, 1.0f, 0.01f};
        if (move_probs.size() != 15) throw std::runtime_error("--move_probs needs 15 comma separated values");
        const int n_chains = static_cast<int>(a.integer("n_chains", 1));
        const int device = static_cast<int>(a.integer("device", 0));
        const int rank = static_cast<int>(a.integer("rank", 0)), world = static_cast<int>(a.integer("world", 1));

        if (!a.has("region_sizes_file")) { std::cerr << "the region sizes file is not provided." << std::endl; return EXIT_FAILURE; }
        if (!a.has("d_matrix_file") && !a.has("bins_matrix_file")) { std::cerr << "the D matrix file is not provided." << std::endl; return EXIT_FAILURE; }
        if (!a.has("n_regions")) { std::cerr << "the number of regions is not provided." << std::endl; return EXIT_FAILURE; }
        if (!a.has("n_cells")) { std::cerr << "the number of cells is not provided." << std::endl; return EXIT_FAILURE; }
        const int n_cells = static_cast<int>(a.integer("n_cells", 0));
        const int n_regions = static_cast<int>(a.integer("n_regions", 0));
        if (!a.has("seed")) seed = -1;  // the reference seeds from std::random_device when --seed is absent
        bool random_init = a.flag("random_init", true);
        if (a.has("tree_file")) random_init = false;

        if (P.verbosity > 0) {
            if (a.has("alpha")) std::cout << "the learning rate alpha constant is specified to be " << alpha << std::endl;
            else std::cout << "the learning rate alpha constant is " << alpha << " since it is not specified" << std::endl;
            if (a.has("gamma")) std::cout << "initial gamma is specified to be " << gamma << std::endl;
            else std::cout << "gamma value is not specified, the default value is: " << gamma << std::endl;
            if (a.has("cf")) std::cout << "Cluster fraction in tree prior is: " << P.cf << std::endl;
            std::cout << "Reading the input matrix..." << std::endl;
        }
        std::vector<double> D(static_cast<size_t>(n_cells) * n_regions, 0.0);
        if (!a.has("bins_matrix_file")) read_counts(D, n_cells, n_regions, a.str("d_matrix_file"));
        if (P.verbosity > 0) std::cout << "Reading the region_sizes file..." << std::endl;
        std::vector<int> region_sizes;
        read_vector(region_sizes, a.str("region_sizes_file"));
        if (static_cast<int>(region_sizes.size()) != n_regions)
            throw std::logic_error("Size of the counts per cell needs to be equal to the number of regions!");  // Tree.h:170-171
        if (a.has("bins_matrix_file")) {
            // cells x bins input (what pyscicone holds before scicone.py:276-299 collapses it): the bins of every region are
            // summed on the device (K0, scicone_condense_matrix = Utils::condense_matrix, Utils.cpp:104-138)
            if (!a.has("n_bins")) throw std::runtime_error("--bins_matrix_file needs --n_bins");
            const long n_bins = a.integer("n_bins", 0);
            std::vector<double> bins(static_cast<size_t>(n_cells) * n_bins, 0.0);
            read_counts(bins, n_cells, static_cast<int>(n_bins), a.str("bins_matrix_file"));
            std::vector<int32_t> sizes32(region_sizes.begin(), region_sizes.end());
            double k0_us = 0.0;
            abi_check(scicone_condense_matrix(device, bins.data(), n_cells, n_bins, sizes32.data(), n_regions, D.data(), &k0_us));
            if (P.verbosity > 0) std::cout << "Condensed " << n_bins << " bins into " << n_regions << " regions (" << k0_us << " us on the device)" << std::endl;
        }
        std::vector<int> cluster_sizes;
        if (a.has("cluster_sizes_file") && !a.str("cluster_sizes_file").empty()) {
            if (P.verbosity > 0) std::cout << "Reading the cluster_sizes file..." << std::endl;
            read_vector(cluster_sizes, a.str("cluster_sizes_file"));
        } else {
            cluster_sizes.assign(n_cells, 1);
            if (n_cells < 20)
                std::cout << "Warning: there are only " << n_cells
                          << " observations. If these are clusters, the cluster_sizes_file parameter should be specified for accurate tree scoring." << std::endl;
        }
        if (!max_scoring) {  // infer_trees.cpp:198-210
            if (P.verbosity > 0) std::cout << "Will perform sum scoring." << std::endl;
            P.cf = 1.0;
            move_probs[11] = 0.0f;
            move_probs[12] = 0.0f;
        } else {
            if (P.verbosity > 0) std::cout << "Will perform maximum scoring." << std::endl;
            move_probs.back() = 0.0f;
        }
        std::vector<int> neutral;
        if (a.has("region_neutral_states_file") && !a.str("region_neutral_states_file").empty()) {
            if (P.verbosity > 0) std::cout << "Reading the region_neutral_states file..." << std::endl;
            read_vector(neutral, a.str("region_neutral_states_file"));
        } else {
            if (P.verbosity > 0) std::cout << "Assuming root to have copy number state " << ploidy << " in all regions" << std::endl;
            neutral.assign(n_regions, ploidy);
        }

        // nu / likelihood branch (infer_trees.cpp:250-281): without --nu and without the nu move the model is the plain multinomial
        const bool learn_nu = static_cast<bool>(move_probs[13]);
        if (!learn_nu && !a.has("nu") && !a.has("tree_file")) P.is_overdispersed = 0;
        if (!learn_nu && !a.has("nu") && a.has("tree_file")) P.is_overdispersed = 0;

        // this rank's block of cells (chunk aligned so that totals do not depend on the number of GPUs)
        // --no_shard (tests): several host processes with the ownership protocol but every one holding all the cells
        const bool no_shard = a.has("no_shard");
        long lo = 0, hi = n_cells;
        if (world > 1 && !no_shard) {
            const long chunk = 1024, n_chunks = (n_cells + chunk - 1) / chunk;
            const long base = n_chunks / world, extra = n_chunks % world;
            long c0 = 0;
            for (int r = 0; r < rank; ++r) c0 += base + (r < extra ? 1 : 0);
            long c1 = c0 + base + (rank < extra ? 1 : 0);
            lo = std::min<long>(c0 * chunk, n_cells);
            hi = std::min<long>(c1 * chunk, n_cells);
            if (hi <= lo) throw std::runtime_error("this rank has no cells: use fewer GPUs");
        }
        if (static_cast<int>(cluster_sizes.size()) != n_cells) throw std::runtime_error("the cluster sizes file does not have n_cells lines");
        if (static_cast<int>(neutral.size()) != n_regions) throw std::runtime_error("the region neutral states file does not have n_regions lines");
        std::vector<int32_t> r32(region_sizes.begin(), region_sizes.end()), n32(neutral.begin(), neutral.end()),
            w32(cluster_sizes.begin() + lo, cluster_sizes.begin() + hi);
        scicone_dataset_desc desc;
        desc.n_cells = static_cast<int32_t>(hi - lo);
        desc.n_regions = n_regions;
        desc.D = D.data() + static_cast<size_t>(lo) * n_regions;
        desc.region_sizes = r32.data();
        desc.neutral_states = n32.data();
        desc.cluster_sizes = w32.data();
        desc.eta = P.eta;
        desc.is_overdispersed = static_cast<int32_t>(P.is_overdispersed);
        desc.max_scoring = max_scoring ? 1 : 0;
        desc.device = device;
        desc.cell_offset = lo;
        desc.n_cells_global = n_cells;
        scicone_dataset* ds = nullptr;
        abi_check(scicone_dataset_create(&ds, &desc));
        if (world > 1 && !no_shard) {
            uint8_t id[128];
            hex_to_bytes(a.str("nccl_id"), id, 128);
            abi_check(scicone_dataset_attach_comm(ds, id, rank, world));
        }
        const bool stats = a.has("stats_file");
        if (stats && a.flag("profile_kernels", true)) scicone_set_profiling(ds, 1);  // CUDA events around every kernel (per-kernel times of the stats file)

        // cell-sharded run: every chain's tree search runs on ONE rank (chain c -> rank c mod world), which compiles each
        // proposal once and ships the compiled bytes through a shared-memory mailbox (mcmc.hpp); the other ranks hold a
        // follower handle of the chain and only stage those bytes.  --shm_name names the segment; by default it is derived
        // from the NCCL id of the run.  --replicated_host: every rank runs every chain's search itself (no mailbox).
        std::unique_ptr<Mailbox> mailbox;
        if (world > 1 && !a.has("replicated_host")) {
            std::string shm = a.str("shm_name");
            if (shm.empty() && a.has("nccl_id")) shm = "/scicone_" + a.str("nccl_id").substr(0, 24);
            if (shm.empty()) throw std::runtime_error("--shm_name is needed for a multi-process run without --nccl_id");
            if (shm[0] != '/') shm = "/" + shm;
            mailbox.reset(new Mailbox());
            mailbox->open(shm, n_chains, rank, world);
        }
        if (world > 1 && !mailbox && seed == -1)
            throw std::runtime_error("--replicated_host needs --seed: every rank must draw the same chains");
        auto owner_of = [&](int c) { return mailbox ? c % world : rank; };
        {   // the row pool and the reduction scratch are sized before the first launch, not inside the timed loop
            const int per_rank = (n_chains + world - 1) / world;
            const long rows_per_chain = 3L * (n_nodes + 24) + 48;
            abi_check(scicone_dataset_reserve(ds, world > 1 ? std::min(n_chains, SCICONE_MAX_SHARDED_BATCH) : n_chains,
                                              static_cast<int64_t>(mailbox || world == 1 ? per_rank : n_chains) * rows_per_chain));
        }

        std::vector<std::unique_ptr<Inference>> owned;
        std::vector<Inference*> chains;
        for (int c = 0; c < n_chains; ++c) {
            Params Pc = P;
            if (n_chains > 1) Pc.postfix = P.postfix + "_c" + std::to_string(c);
            const bool mine = owner_of(c) == rank;
            owned.emplace_back(new Inference(Pc, seed == -1 ? -1 : seed + c, static_cast<unsigned>(n_regions), n_cells, neutral, ploidy, max_scoring));
            Inference* inf = owned.back().get();
            inf->attach_chain(ds, !mine);
            inf->gamma = gamma;
            inf->alpha = alpha;
            chains.push_back(inf);
            if (!mine) continue;  // the owner draws the start tree; a follower never sees a tree
            if (random_init) {
                if (P.verbosity > 0 && c == 0) std::cout << "Trying to randomly initialise a valid tree with " << 10000 << " iterations." << std::endl;
                try {
                    inf->random_initialize(static_cast<unsigned>(n_nodes), 10000);
                } catch (const std::runtime_error& e) {
                    std::cerr << "A runtime error was caught during the random tree initialize function, with message '" << e.what() << '\'' << std::endl;
                    return EXIT_FAILURE;
                }
            }
            if (a.has("tree_file")) {
                inf->initialize_from_file(a.str("tree_file"));
                nu = inf->t.nu;
            }
            if (a.has("nu")) inf->t.nu = inf->t_prime.nu = nu;  // with --tree_file, nu == the file's value (infer_trees.cpp:247)
            if (!mailbox) {
                inf->compute_t_table();
                inf->compute_t_od_scores();
                inf->update_t_prime();
            }
        }
        if (mailbox) {
            // Full passes of the start trees (Inference::compute_t_table + compute_t_od_scores, infer_trees.cpp:285-286) as
            // launches of all ranks: the owner compiles and publishes, everybody stages and launches, the owner commits.
            int32_t max_batch = 0;
            abi_check(scicone_max_batch(ds, &max_batch));
            max_batch = std::min<int32_t>(max_batch, SCICONE_MAX_SHARDED_BATCH);
            for (int c0 = 0; c0 < n_chains; c0 += max_batch) {
                const int c1 = std::min(n_chains, c0 + max_batch);
                std::vector<scicone_chain*> handles;
                for (int c = c0; c < c1; ++c) {
                    Inference* inf = chains[c];
                    if (owner_of(c) == rank) {
                        inf->prepare_t_table();
                        inf->last_decision = -1;
                        pack_message(*inf, true, inf->msg_buf);
                        mailbox->publish(c, ++inf->msg_seq, inf->msg_buf, rank, world);
                    } else {
                        uint32_t bytes = 0;
                        const unsigned char* msg = mailbox->receive(c, ++inf->msg_seq, &bytes);
                        apply_message(*inf, msg, bytes);
                        mailbox->acknowledge(c, inf->msg_seq, rank);
                    }
                    handles.push_back(inf->chain);
                }
                std::vector<double> sums(handles.size()), ods(handles.size());
                abi_check(scicone_batch_compute_t_prime_sums(handles.data(), nullptr, static_cast<int32_t>(handles.size()), sums.data(), ods.data()));
                for (int c = c0; c < c1; ++c) {
                    Inference* inf = chains[c];
                    if (owner_of(c) == rank) {
                        inf->finish_t_table(sums[c - c0], ods[c - c0]);
                        inf->update_t_prime();
                    } else
                        abi_check(scicone_accept(inf->chain));
                }
            }
        }

        if (P.verbosity > 0) std::cout << "Inferring the tree using MCMC with " << n_iters << " iterations..." << std::endl;
        HostPhaseTimes phase;
        phase.want_timeline = a.flag("timeline", false);
        // schedule: "dynamic" (default for several chains on one GPU) launches whatever is ready - in a cell-sharded run
        // rank 0 composes the launches and the other ranks replay its launch log; "static" moves groups of chains in lock-step
        const std::string schedule = a.str("schedule", ((world == 1 || mailbox) && n_chains >= 4) ? "dynamic" : "static");
        const bool dynamic = schedule == "dynamic" && (world == 1 || mailbox);
        // --warmup_iters (bench.py): untimed iterations before the measured run; the chains simply continue afterwards
        const int warmup_iters = static_cast<int>(a.integer("warmup_iters", 0));
        double k0[4];
        uint64_t launches0 = 0;
        std::chrono::high_resolution_clock::time_point start;
        auto mark = [&]() {  // everything before this point is warm-up: counters back to zero, clocks started
            for (Inference* c : chains) {
                c->n_accepted = c->n_rejected = 0;
                c->n_rescored_nodes = c->n_attempts = 0;
                c->t_finish_us = c->t_propose_us = c->t_prepare_us = c->t_max_visit_us = 0;
            }
            scicone_counters(ds, k0, 1);
            scicone_build_counters(ds, k0, 1);
            scicone_kernel_time_stats(ds, nullptr, nullptr, 1);
            scicone_kernel_launches(ds, &launches0);
            scicone_region_mark(ds, 0);
            start = std::chrono::high_resolution_clock::now();
        };
        if (dynamic) {
            // one call: the warm-up iterations are a phase of their own (all chains stop at the mark, nothing in flight),
            // then the clock starts; worker threads and launch lanes are those of a long run
            if (warmup_iters <= 0) mark();
            infer_mcmc_dynamic(chains, move_probs, n_iters, size_limit, static_cast<int>(a.integer("min_batch", 0)), &phase,
                               static_cast<int>(a.integer("n_streams", 2)), rank, world, mailbox.get(), std::max(0, warmup_iters), mark);
        } else {
            const int n_groups = static_cast<int>(a.integer("n_groups", 0));
            if (warmup_iters > 0) infer_mcmc(chains, move_probs, warmup_iters, size_limit, n_groups, nullptr, rank, world, mailbox.get());
            mark();
            infer_mcmc(chains, move_probs, n_iters, size_limit, n_groups, &phase, rank, world, mailbox.get());
        }
        if (P.verbosity > 0 && n_chains > 1)
            std::cout << "driver thread: chains " << phase.host_us << " us, submit " << phase.submit_us << " us, wait " << phase.wait_us
                      << " us (" << scicone_host_threads() << " host threads)" << std::endl;
        scicone_region_mark(ds, 1);
        double region_ms = 0.0;
        scicone_region_elapsed_ms(ds, &region_ms);
        auto stop = std::chrono::high_resolution_clock::now();
        auto duration = std::chrono::duration_cast<std::chrono::microseconds>(stop - start);
        if (P.verbosity > 0) std::cout << "Time taken by infer_mcmc function: " << duration.count() << " microseconds" << std::endl;
        if (stats && (rank == 0 || a.has("stats_all_ranks"))) {
            double k[4], kern_us = 0, build_us = 0;
            uint64_t n_timed = 0, n_build = 0, launches = 0;
            scicone_counters(ds, k, 0);
            double kb[2] = {0, 0};
            scicone_build_counters(ds, kb, 0);
            scicone_kernel_time_stats(ds, &kern_us, &n_timed, 0);
            scicone_build_time_stats(ds, &build_us, &n_build);
            scicone_kernel_launches(ds, &launches);
            uint64_t dev_bytes = 0;
            scicone_device_bytes(ds, &dev_bytes);
            unsigned long long acc = 0, rej = 0, resc = 0, attempts = 0;
            double cf = 0, cp = 0, cq = 0, cmax = 0;
            for (Inference* c : chains) {
                acc += c->n_accepted; rej += c->n_rejected; resc += c->n_rescored_nodes; attempts += c->n_attempts;
                cf += c->t_finish_us; cp += c->t_propose_us; cq += c->t_prepare_us; cmax = std::max(cmax, c->t_max_visit_us);
            }
            std::ofstream sf(a.str("stats_file"));
            sf << std::setprecision(17) << "{\"n_chains\":" << n_chains << ",\"n_iters\":" << n_iters << ",\"n_cells\":" << n_cells
               << ",\"wall_us\":" << duration.count() << ",\"device_region_ms\":" << region_ms << ",\"score_kernel_us\":" << kern_us
               << ",\"score_kernel_launches\":" << n_timed << ",\"build_kernel_us\":" << build_us << ",\"build_kernel_launches\":" << n_build
               << ",\"gpu_launches\":" << (launches - launches0) << ",\"h2d_bytes\":" << k[0] << ",\"d2h_bytes\":" << k[1]
               << ",\"alg_bytes\":" << k[2] << ",\"n_scores\":" << k[3] << ",\"build_alg_bytes\":" << kb[0] << ",\"build_lgamma\":" << kb[1] << ",\"device_bytes\":" << dev_bytes << ",\"accepted\":" << acc
               << ",\"rejected\":" << rej << ",\"rescored_nodes\":" << resc << ",\"host_threads\":" << scicone_host_threads() << ",\"attempts\":" << attempts << ",\"chain_finish_us\":" << cf << ",\"chain_propose_us\":" << cp
               << ",\"chain_prepare_us\":" << cq << ",\"chain_max_visit_us\":" << cmax << ",\"schedule\":\"" << schedule << "\",\"launches\":" << phase.launches << ",\"launched_chains\":" << phase.launched_chains
               << ",\"host_us\":" << phase.host_us
               << ",\"submit_us\":" << phase.submit_us << ",\"wait_us\":" << phase.wait_us;
            if (phase.want_timeline) {
                sf << ",\"timeline\":[";
                for (size_t i = 0; i < phase.timeline.size(); ++i)
                    sf << (i ? "," : "") << "[" << std::setprecision(9) << phase.timeline[i][0] << "," << phase.timeline[i][1] << "," << phase.timeline[i][2] << "]";
                sf << "]";
            }
            sf << "}" << std::endl;
        }

        if (P.verbosity > 0) std::cout << "Writing the inferred tree and cnvs..." << std::endl;
        for (size_t ci = 0; ci < chains.size(); ++ci) {
            Inference* inf = chains[ci];
            if (world > 1) {
                // cell-sharded run: the owner of a chain writes its tree; per-cell outputs would have to be merged over the
                // ranks and are not produced here
                const bool owner = mailbox ? static_cast<int>(ci) % world == rank : rank == 0;
                if (owner) {
                    std::ofstream f_tree("./" + inf->P.postfix + "_tree_inferred.txt");
                    inf->best_tree.write(f_tree);
                }
                continue;
            }
            // Inference::assign_cells_to_nodes, Inference.h:1311-1377
            inf->t.copy_from(inf->best_tree);
            inf->compute_t_table();
            const int M = desc.n_cells;
            std::vector<int32_t> ids(M), cn(static_cast<size_t>(M) * n_regions);
            abi_check(scicone_assign_cells_to_nodes(inf->chain, &inf->flat.pod, ids.data(), cn.data()));
            const std::string& pf = inf->P.postfix;
            if (P.verbosity > 1) {
                std::ofstream f_ids(pf + "_cell_node_ids.tsv"), f_cn(pf + "_cell_region_cnvs.csv");
                for (int j = 0; j < M; ++j) f_ids << j << '\t' << ids[j] << '\n';
                for (int j = 0; j < M; ++j) {
                    for (int k = 0; k < n_regions; ++k) f_cn << cn[static_cast<size_t>(j) * n_regions + k] << (k == n_regions - 1 ? "" : ",");
                    f_cn << '\n';
                }
            }
            std::ofstream f_tree("./" + pf + "_tree_inferred.txt");
            inf->best_tree.write(f_tree);
            // Utils::regions_to_bins_cnvs (Utils.cpp:165-188) + the writer of infer_trees.cpp:321-331: cells x bins as text,
            // formatted and written in blocks of rows on the worker pool (cnv_writer.hpp).  Timing runs skip it: 4 GB per
            // chain at configs[4].
            if (!a.flag("no_cnv_file", false))
                write_bins_cnvs("./" + pf + "_inferred_cnvs.csv", cn.data(), M, n_regions, region_sizes,
                                [](int n, const std::function<void(int)>& fn) { parallel_chains(n, [&fn](int i) { fn(i); }); });
            if (P.verbosity > 1) {
                int32_t nn = 0;
                abi_check(scicone_t_n_nodes(inf->chain, &nn));
                std::vector<double> sc(static_cast<size_t>(M) * nn);
                abi_check(scicone_read_t_scores(inf->chain, sc.data()));
                std::ofstream f_sc("./" + pf + "_attachment_scores.csv");
                for (int j = 0; j < M; ++j) {
                    for (int c = 0; c < nn; ++c) f_sc << sc[static_cast<size_t>(j) * nn + c] << (c == nn - 1 ? "" : ",");
                    f_sc << std::endl;
                }
            }
        }
        owned.clear();
        scicone_dataset_destroy(ds);
        if (P.verbosity > 0) std::cout << "Tree inference is successfully completed!" << std::endl;
        return EXIT_SUCCESS;
    } catch (const std::exception& e) {
        std::cerr << "inference: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
}

// The reference's second executable, `score` (scicone/src/score_tree.cpp:31-125): load a tree file, set nu, one full pass
// (compute_t_table, score_tree.cpp:114) and the root scores (compute_t_od_scores, :115), print the rescored tree.  Same
// flags, defaults (cf = 1, lambda_r = 2, lambda_c = 1, c_penalise = 1, sum scoring) and output; plus --device.
extern "C" int scicone_score_main(int argc, char** argv) {
    try {
        Args a = parse(argc, argv);
        Params P;
        P.print_precision = static_cast<int>(a.integer("print_precision", 15));
        P.lambda_s = 0.5;
        P.lambda_r = 2.0;
        P.lambda_c = 1.0;
        P.cf = a.num("cf", 1.0);
        P.c_penalise = 1.0;
        P.copy_number_limit = 5;
        P.is_overdispersed = static_cast<unsigned>(a.integer("is_overdispersed", 1));
        P.eta = 1e-4;
        P.verbosity = 0;
        P.postfix = a.str("postfix");
        const int n_cells = static_cast<int>(a.integer("n_cells", 0));
        int n_regions = static_cast<int>(a.integer("n_regions", 0));
        const int ploidy = static_cast<int>(a.integer("ploidy", 2));
        const double nu = a.num("nu", 1.0);
        if (n_cells <= 0 || n_regions <= 0) throw std::runtime_error("--n_cells and --n_regions are needed");
        std::cout << "Reading the input matrix..." << std::endl;
        std::vector<double> D(static_cast<size_t>(n_cells) * n_regions, 0.0);
        read_counts(D, n_cells, n_regions, a.str("d_matrix_file"));
        std::cout << "Reading the region_sizes file..." << std::endl;
        std::vector<int> region_sizes;
        read_vector(region_sizes, a.str("region_sizes_file"));
        if (static_cast<int>(region_sizes.size()) != n_regions)
            throw std::logic_error("Size of the counts per cell needs to be equal to the number of regions!");  // Tree.h:170-171
        std::vector<int> cluster_sizes;
        if (a.has("cluster_sizes_file")) {
            std::cout << "Reading the cluster_sizes file..." << std::endl;
            read_vector(cluster_sizes, a.str("cluster_sizes_file"));
        } else {
            if (n_cells < 20)
                std::cout << "Warning: there are only " << n_cells
                          << " observations. If these are clusters, the cluster_sizes_file parameter should be specified for accurate tree scoring.";
            cluster_sizes.assign(n_cells, 1);
        }
        if (static_cast<int>(cluster_sizes.size()) != n_cells) throw std::runtime_error("the cluster sizes file does not have n_cells lines");
        std::vector<int> neutral;
        if (a.has("region_neutral_states_file")) {
            std::cout << "Reading the region_neutral_states file..." << std::endl;
            read_vector(neutral, a.str("region_neutral_states_file"));
        } else {
            std::cout << "Assuming root to have copy number state " << ploidy << " in all regions" << std::endl;
            neutral.assign(n_regions, ploidy);
        }
        if (static_cast<int>(cluster_sizes.size()) != n_cells) throw std::runtime_error("the cluster sizes file does not have n_cells lines");
        if (static_cast<int>(neutral.size()) != n_regions) throw std::runtime_error("the region neutral states file does not have n_regions lines");
        std::vector<int32_t> r32(region_sizes.begin(), region_sizes.end()), n32(neutral.begin(), neutral.end()), w32(cluster_sizes.begin(), cluster_sizes.end());
        scicone_dataset_desc desc;
        desc.n_cells = n_cells;
        desc.n_regions = n_regions;
        desc.D = D.data();
        desc.region_sizes = r32.data();
        desc.neutral_states = n32.data();
        desc.cluster_sizes = w32.data();
        desc.eta = P.eta;
        desc.is_overdispersed = static_cast<int32_t>(P.is_overdispersed);
        desc.max_scoring = 0;  // Inference's default (Inference.h:51)
        desc.device = static_cast<int32_t>(a.integer("device", 0));
        desc.cell_offset = 0;
        desc.n_cells_global = n_cells;
        scicone_dataset* ds = nullptr;
        abi_check(scicone_dataset_create(&ds, &desc));
        {
            Inference inf(P, 0, static_cast<unsigned>(n_regions), n_cells, neutral, ploidy, false);
            inf.attach_chain(ds);
            inf.initialize_from_file(a.str("tree_file"));
            inf.t.nu = nu;
            inf.compute_t_table();
            inf.compute_t_od_scores();
            inf.update_t_prime();
            inf.t.write(std::cout);
        }
        scicone_dataset_destroy(ds);
        return EXIT_SUCCESS;
    } catch (const std::exception& e) {
        std::cerr << "score: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
}

#ifndef SCICONE_HOST_NO_MAIN
#ifdef SCICONE_SCORE_MAIN
int main(int argc, char** argv) { return scicone_score_main(argc, argv); }
#else
int main(int argc, char** argv) { return scicone_inference_main(argc, argv); }
#endif
#endif
