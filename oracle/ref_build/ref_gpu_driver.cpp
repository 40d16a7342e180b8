// ref_gpu_driver.cpp — TEST/BENCH INFRASTRUCTURE ONLY.
// Times the reference's own CUDA path, cuda::gpu_layout (src/cuda/layout.cu:290-476, compiled unmodified for
// sm_100a by oracle/ref_build/Makefile `refgpu`), as the REPORTED GPU baseline (not the optimisation target).
//
//   ref_gpu_driver <in.gfa> <out.arr|-> [iter_max=30] [threads=8] [updates_x=10] [init.arr|-] [repeats=0]
//
// With init.arr (X, Y: the injected initial layout the scale bands use) and repeats > 0 it instead writes the final layouts of
// `repeats` complete runs of iter_max iterations to <out.arr>.run<k>.arr: the reference CUDA path's own distribution of the
// final stress (its worker seeds are fixed, layout.cu:29, so the runs differ by GPU timing only).
//
// The reference gives no hook around its iteration loop (layout.cu:442-447), so the loop time is obtained by
// difference: the call is run with iter_max and with 2*iter_max; (t2 - t1) is the time of iter_max iterations
// (flattening, managed-memory first touch, RNG init and copy-back cancel).
#include <atomic>
#include <chrono>
#include <cmath>
#include <iostream>
#include <string>
#include <vector>

#include "odgi.hpp"
#include "gfa_to_handle.hpp"
#include "cuda/layout.h"
#include "../../odgi_b200/host/pgsgd_arrays.hpp"

using namespace odgi;

static std::vector<double> g_init_x, g_init_y;   // injected initial layout (optional)

static double run_once(const graph_t& graph, uint64_t iter_max, uint64_t U, uint64_t max_steps, int threads,
                       std::vector<std::atomic<double>>& X, std::vector<std::atomic<double>>& Y) {
    uint64_t N = graph.get_node_count();
    uint64_t len = 0;
    if (g_init_x.size() == 2 * N) {
        for (uint64_t i = 0; i < 2 * N; ++i) { X[i].store(g_init_x[i]); Y[i].store(g_init_y[i]); }
    } else
    for (uint64_t r = 0; r < N; ++r) {   // 'd' initialisation without noise (layout_main.cpp:322-328)
        X[2 * r].store(len); Y[2 * r].store(0.01 * (double) (r % 97));
        len += graph.get_length(graph.get_handle(r + 1, false));
        X[2 * r + 1].store(len); Y[2 * r + 1].store(0.01 * (double) (r % 89));
    }
    cuda::layout_config_t config;   // as path_linear_sgd_layout_gpu fills it (path_sgd_layout.cpp:490-501) with layout_main.cpp defaults
    config.iter_max = iter_max;
    config.min_term_updates = U;
    config.eta_max = (double) max_steps * (double) max_steps;
    config.eps = 0.01;
    config.iter_with_max_learning_rate = 0;
    config.first_cooling_iteration = std::floor(0.5 * (double) iter_max);
    config.theta = 0.99;
    config.space = uint32_t(max_steps);
    config.space_max = 1000;
    config.space_quantization_step = 100;
    config.nthreads = threads;
    auto t0 = std::chrono::steady_clock::now();
    cuda::gpu_layout(config, graph, X, Y);
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char** argv) {
    if (argc < 3) { std::cerr << "usage: ref_gpu_driver <in.gfa> <out.arr|-> [iter_max] [threads] [updates_x]" << std::endl; return 2; }
    uint64_t iter_max = argc > 3 ? std::stoull(argv[3]) : 30;
    int threads = argc > 4 ? std::stoi(argv[4]) : 8;
    double updates_x = argc > 5 ? std::stod(argv[5]) : 10.0;
    graph_t graph;
    auto tl = std::chrono::steady_clock::now();
    gfa_to_handle(argv[1], &graph, false, threads, false);
    graph.set_number_of_threads(threads);
    double load_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - tl).count();
    uint64_t sum_steps = 0, max_steps = 0;
    graph.for_each_path_handle([&](const path_handle_t& p) {
        uint64_t c = graph.get_step_count(p);
        sum_steps += c; max_steps = std::max(max_steps, c);
    });
    uint64_t U = (uint64_t) (updates_x * sum_steps);
    uint64_t N = graph.get_node_count();
    std::vector<std::atomic<double>> X(2 * N), Y(2 * N);
    if (argc > 6 && std::string(argv[6]) != "-") {
        auto a = pgsgd::read_arrays(argv[6]);
        g_init_x = a.at("X").vec<double>(); g_init_y = a.at("Y").vec<double>();
        if (g_init_x.size() != 2 * N || g_init_y.size() != 2 * N) { std::cerr << "init.arr: X, Y must hold 2N doubles" << std::endl; return 2; }
    }
    const int repeats = argc > 7 ? std::stoi(argv[7]) : 0;
    run_once(graph, 2, U, max_steps, threads, X, Y);  // warm-up: CUDA context creation, module load
    if (repeats > 0) {
        for (int k = 0; k < repeats; ++k) {
            const double t = run_once(graph, iter_max, U, max_steps, threads, X, Y);
            std::vector<double> x(2 * N), y(2 * N);
            for (uint64_t i = 0; i < 2 * N; ++i) { x[i] = X[i].load(); y[i] = Y[i].load(); }
            pgsgd::ArrayWriter w(std::string(argv[2]) + ".run" + std::to_string(k) + ".arr");
            w.add("X", x); w.add("Y", y);
            w.close();
            std::cout << "{\"impl\": \"reference src/cuda/layout.cu (sm_100a)\", \"run\": " << k << ", \"iter_max\": " << iter_max << ", \"call_s\": " << t << "}" << std::endl;
        }
        return 0;
    }
    double t1 = run_once(graph, iter_max, U, max_steps, threads, X, Y);
    double t2 = run_once(graph, 2 * iter_max, U, max_steps, threads, X, Y);
    double loop_s = t2 - t1;
    uint64_t Ur = ((U + 1023) / 1024) * 1024;  // the reference rounds the launch up to whole 1024-thread blocks (layout.cu:431)
    if (std::string(argv[2]) != "-") {
        std::vector<double> x(2 * N), y(2 * N);
        for (uint64_t i = 0; i < 2 * N; ++i) { x[i] = X[i].load(); y[i] = Y[i].load(); }
        pgsgd::ArrayWriter w(argv[2]);
        w.add("X", x); w.add("Y", y);
        w.close();
    }
    std::cout << "{\"impl\": \"reference src/cuda/layout.cu (sm_100a)\", \"nodes\": " << N << ", \"steps\": " << sum_steps << ", \"iter_max\": " << iter_max
              << ", \"updates_per_iter\": " << Ur << ", \"gfa_load_s\": " << load_s << ", \"call_s_iter\": " << t1 << ", \"call_s_2iter\": " << t2
              << ", \"loop_s\": " << loop_s << ", \"updates_per_sec\": " << (double) (iter_max * Ur) / loop_s << "}" << std::endl;
    return 0;
}
