// shim_driver.cpp — TEST INFRASTRUCTURE: `odgi layout --gpu` / `odgi sort --gpu` through the reference's own call
// chain with our shim linked in place of src/cuda/layout.cu.
//   layout: graph_t (reference GFA ingest) -> algorithms::path_linear_sgd_layout_gpu (UNMODIFIED reference code,
//           path_sgd_layout.cpp:470-504, compiled -DUSE_GPU) -> cuda::gpu_layout (odgi_b200/host/odgi_shim.cpp) -> C-ABI
//   sort:   graph_t -> algorithms::path_linear_sgd_gpu (shim) -> C-ABI
// usage: shim_driver layout|sort <in.gfa> <out.arr|-> [iter_max] [threads]      (PGSGD_SHIM_TIMING=1: phase times on stderr)
#include <atomic>
#include <chrono>
#include <cmath>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "odgi.hpp"
#include "gfa_to_handle.hpp"
#include "algorithms/xp.hpp"
#include "algorithms/path_sgd_layout.hpp"
#include "../../odgi_b200/host/pgsgd_arrays.hpp"

using namespace odgi;
namespace odgi { namespace algorithms {
std::vector<double> path_linear_sgd_gpu(const graph_t&, const xp::XP&, const std::vector<path_handle_t>&, const uint64_t&, const uint64_t&,
                                        const uint64_t&, const double&, const double&, const double&, const double&, const uint64_t&,
                                        const uint64_t&, const uint64_t&, const double&, const uint64_t&, const bool&, const bool&,
                                        std::vector<std::string>&, const bool&, std::vector<bool>&);
} }

// path_sgd.cpp references utils::graph_deep_copy on its snapshot branch only (see ref_driver.cpp)
namespace utils { void graph_deep_copy(const odgi::graph_t&, odgi::graph_t*) { std::abort(); } }

int main(int argc, char** argv) {
    if (argc < 4) { std::cerr << "usage: shim_driver layout|sort <in.gfa> <out.arr> [iter_max]" << std::endl; return 2; }
    const std::string mode = argv[1];
    const uint64_t threads = argc > 5 ? std::stoull(argv[5]) : 1;
    graph_t graph;
    const auto t_load = std::chrono::steady_clock::now();
    gfa_to_handle(argv[2], &graph, false, threads, false);
    graph.set_number_of_threads(threads);
    const double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load).count();
    xp::XP path_index;  // the GPU path never reads it; left empty on purpose (skip XP when --gpu)
    std::vector<path_handle_t> paths;
    uint64_t sum_steps = 0, max_steps = 0, max_bp = 0;
    graph.for_each_path_handle([&](const path_handle_t& p) {
        paths.push_back(p);
        uint64_t c = graph.get_step_count(p), bp = 0;
        graph.for_each_step_in_path(p, [&](const step_handle_t& s) { bp += graph.get_length(graph.get_handle_of_step(s)); });
        sum_steps += c; max_steps = std::max(max_steps, c); max_bp = std::max(max_bp, bp);
    });
    const uint64_t N = graph.get_node_count();
    const bool keep = std::string(argv[3]) != "-";
    pgsgd::ArrayWriter w(keep ? argv[3] : "/dev/null");
    double call_s = 0;
    if (mode == "layout") {
        const uint64_t iter_max = argc > 4 ? std::stoull(argv[4]) : 30;
        std::vector<std::atomic<double>> X(2 * N), Y(2 * N);
        std::mt19937 rng(42);
        std::normal_distribution<double> gaussian_noise(0, std::sqrt((double) N * 2));
        uint64_t len = 0;
        for (uint64_t r = 0; r < N; ++r) {  // layout_main.cpp:322-328
            X[2 * r].store(len); Y[2 * r].store(gaussian_noise(rng));
            len += graph.get_length(graph.get_handle(r + 1, false));
            X[2 * r + 1].store(len); Y[2 * r + 1].store(gaussian_noise(rng));
        }
        std::vector<double> x0(2 * N), y0(2 * N);
        for (uint64_t i = 0; i < 2 * N; ++i) { x0[i] = X[i].load(); y0[i] = Y[i].load(); }
        const auto t_call = std::chrono::steady_clock::now();
        algorithms::path_linear_sgd_layout_gpu(graph, path_index, paths, iter_max, 0, 10 * sum_steps, 0, 0.01, (double) max_steps * max_steps,
                                               0.99, max_steps, 1000, 100, 0.5, threads, false, false, "", X, Y);
        call_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_call).count();
        std::vector<double> x(2 * N), y(2 * N);
        for (uint64_t i = 0; i < 2 * N; ++i) { x[i] = X[i].load(); y[i] = Y[i].load(); }
        w.add("X0", x0); w.add("Y0", y0); w.add("X", x); w.add("Y", y);
    } else {
        const uint64_t iter_max = argc > 4 ? std::stoull(argv[4]) : 100;
        const uint64_t space_max = 100, max_dists = 101;
        const uint64_t q = std::max<uint64_t>(2, (uint64_t) std::ceil((double) (max_bp - space_max) / (double) (max_dists - space_max)));
        std::vector<std::string> snapshots;
        std::vector<bool> target_nodes;
        const auto t_call = std::chrono::steady_clock::now();
        std::vector<double> x = algorithms::path_linear_sgd_gpu(graph, path_index, paths, iter_max, 0, sum_steps, 0, 0.01,
                                                                (double) max_steps * max_steps, 0.99, max_bp, space_max, q, 0.5, threads, false, false,
                                                                snapshots, false, target_nodes);
        call_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_call).count();
        w.add("X", x);
    }
    w.close();
    std::cout << "{\"mode\": \"" << mode << "\", \"nodes\": " << N << ", \"steps\": " << sum_steps << ", \"threads\": " << threads
              << ", \"gfa_load_s\": " << load_s << ", \"reference_call_s\": " << call_s << "}" << std::endl;
    return 0;
}
