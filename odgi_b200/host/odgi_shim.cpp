// odgi_shim.cpp — the odgi-side of the drop-in: the reference's own C++ entry points, implemented on top of the
// C-ABI (include/pgsgd.h).  Compiled against the odgi headers (never copied here); a maintainer builds this file
// INSTEAD of src/cuda/layout.cu and links libpgsgd_b200.so.
//
//   void cuda::gpu_layout(cuda::layout_config_t, const odgi::graph_t&, std::vector<std::atomic<double>>& X,
//                         std::vector<std::atomic<double>>& Y)                 — same signature as src/cuda/layout.h:80,
//        so algorithms::path_linear_sgd_layout_gpu (src/algorithms/path_sgd_layout.cpp:470-504) and
//        `odgi layout --gpu` (src/subcommand/layout_main.cpp:333-358) work unmodified;
//   std::vector<double> odgi::algorithms::path_linear_sgd_gpu(...)             — same parameter list as
//        path_linear_sgd (src/algorithms/path_sgd.cpp:12-31); the body of `odgi sort --gpu`: path_linear_sgd_order
//        (path_sgd.cpp:503-684) calls it where it calls path_linear_sgd today.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "odgi.hpp"
#include "algorithms/xp.hpp"
#include "cuda/layout.h"

#include "pgsgd_flatten.hpp"

namespace {

[[noreturn]] void die(const char* where) {
    // the reference's GPU path reports CUDA failures with printf + exit(EXIT_FAILURE) (layout.cu:6-13)
    std::printf("Failed: %s: %s\n", where, pgsgd_last_error());
    std::exit(EXIT_FAILURE);
}

void check_abi() {
    // a shim compiled against another include/pgsgd.h than the library it loaded would pass structs of the wrong size
    if (pgsgd_version() != PGSGD_VERSION) {
        std::printf("Failed: libpgsgd_b200 ABI %d, this binary was built against %d: rebuild\n", pgsgd_version(), PGSGD_VERSION);
        std::exit(EXIT_FAILURE);
    }
}

pgsgd::FlatGraph flatten_or_exit(const odgi::graph_t& graph, uint64_t nthreads) {
    try {
        // config.nthreads is used for exactly this in the reference too (OpenMP path walk, layout.cu:371)
        return pgsgd::flatten_handle_graph<odgi::graph_t, handlegraph::path_handle_t, handlegraph::step_handle_t>(graph, (unsigned) (nthreads ? nthreads : 1));
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());   // same message and exit code as layout.cu:320-323
        std::exit(1);
    }
}

}  // namespace

namespace cuda {

void gpu_layout(layout_config_t config, const odgi::graph_t& graph, std::vector<std::atomic<double>>& X,
                std::vector<std::atomic<double>>& Y) {
    std::cout << "===== Use GPU to compute odgi-layout =====" << std::endl;  // layout.cu:293
    const auto t_begin = std::chrono::steady_clock::now();
    const pgsgd::FlatGraph fg = flatten_or_exit(graph, (uint64_t) (config.nthreads > 0 ? config.nthreads : 1));
    const double flatten_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    check_abi();
    pgsgd_config c{};
    c.iter_max = config.iter_max;
    c.iter_with_max_learning_rate = (uint64_t) config.iter_with_max_learning_rate;
    c.min_term_updates = config.min_term_updates;
    c.delta = 0;                       // the reference GPU path has no early stop (layout.cu:179)
    c.eps = config.eps;
    c.eta_max = config.eta_max;
    c.theta = config.theta;
    c.space = config.space;
    c.space_max = config.space_max;
    c.space_quantization_step = config.space_quantization_step;
    // layout_config_t carries first_cooling_iteration = floor(cooling_start * iter_max) (path_sgd_layout.cpp:496)
    c.cooling_start = ((double) config.first_cooling_iteration + 0.5) / (double) config.iter_max;
    c.seed = 9399220;                  // path_sgd_layout.cpp:168
    const uint64_t n2 = 2 * (uint64_t) graph.get_node_count();
    std::vector<double> x(n2), y(n2);
    for (uint64_t i = 0; i < n2; ++i) { x[i] = X[i].load(); y[i] = Y[i].load(); }
    pgsgd_stats st;
    const pgsgd_graph_view v = fg.view();
    if (pgsgd_layout_2d(&v, &c, x.data(), y.data(), &st) != PGSGD_OK) die("pgsgd_layout_2d");
    for (uint64_t i = 0; i < n2; ++i) {
        if (!std::isfinite(x[i]) || !std::isfinite(y[i])) std::cout << "WARNING: invalid coordiate" << std::endl;  // layout.cu:455-459
        X[i].store(x[i]);
        Y[i].store(y[i]);
    }
    if (std::getenv("PGSGD_SHIM_TIMING")) {   // where the time of the reference-side call goes (one JSON line on stderr)
        const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        std::fprintf(stderr, "{\"shim\": \"cuda::gpu_layout\", \"threads\": %d, \"graph_walk_flatten_s\": %.4f, \"upload_s\": %.4f, "
                             "\"iterations_s\": %.4f, \"download_s\": %.4f, \"total_s\": %.4f, \"term_updates\": %llu}\n",
                     (int) config.nthreads, flatten_s, st.seconds_upload, st.seconds_iterations, st.seconds_download, total_s,
                     (unsigned long long) st.term_updates);
    }
}

}  // namespace cuda

namespace odgi {
namespace algorithms {

std::vector<double> path_linear_sgd_gpu(const graph_t& graph, const xp::XP& /*path_index: not needed on the GPU path*/,
                                        const std::vector<path_handle_t>& /*path_sgd_use_paths*/, const uint64_t& iter_max,
                                        const uint64_t& iter_with_max_learning_rate, const uint64_t& min_term_updates,
                                        const double& delta, const double& eps, const double& eta_max, const double& theta,
                                        const uint64_t& space, const uint64_t& space_max, const uint64_t& space_quantization_step,
                                        const double& cooling_start, const uint64_t& nthreads, const bool& /*progress*/,
                                        const bool& /*snapshot*/, std::vector<std::string>& /*snapshots*/,
                                        const bool& target_sorting, std::vector<bool>& target_nodes) {
    const pgsgd::FlatGraph fg = flatten_or_exit(graph, nthreads);
    check_abi();
    pgsgd_config c{};
    c.iter_max = iter_max;
    c.iter_with_max_learning_rate = iter_with_max_learning_rate;
    c.min_term_updates = min_term_updates;
    c.delta = delta;
    c.eps = eps;
    c.eta_max = eta_max;
    c.theta = theta;
    c.space = space;
    c.space_max = space_max;
    c.space_quantization_step = space_quantization_step;
    c.cooling_start = cooling_start;
    c.seed = 9399220;
    std::vector<uint8_t> frozen;
    if (target_sorting) {  // odgi sort -H: reference nodes stay put (path_sgd.cpp:290-302)
        frozen.resize(graph.get_node_count());
        for (size_t i = 0; i < frozen.size(); ++i) frozen[i] = target_nodes[i] ? 1 : 0;
    }
    std::vector<double> X(graph.get_node_count());
    pgsgd_stats st;
    const pgsgd_graph_view v = fg.view();
    // X is initialised to the cumulative node length by the engine, as path_linear_sgd does itself (path_sgd.cpp:63-69)
    if (pgsgd_sort_1d(&v, &c, target_sorting ? frozen.data() : nullptr, 0, X.data(), &st) != PGSGD_OK) die("pgsgd_sort_1d");
    return X;
}

}  // namespace algorithms
}  // namespace odgi
