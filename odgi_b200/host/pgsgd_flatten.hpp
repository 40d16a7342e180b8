// pgsgd_flatten.hpp — host-side flattening of a path handle graph into the path-major SoA the C-ABI takes.
//
// Replaces the walk cuda::gpu_layout does for itself (src/cuda/layout.cu:325-410) and, for the GPU path, the whole
// XP index (src/algorithms/xp.cpp:49-175 — minutes and a disk-backed multimap at chr6 scale, never read by the GPU).
// Templated on the graph type so that it compiles against odgi::graph_t (libhandlegraph interface names:
// get_node_count, for_each_path_handle, for_each_step_in_path, get_handle_of_step, get_id, get_is_reverse, get_length)
// without this header including any odgi header.
#pragma once
#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pgsgd.h"

namespace pgsgd {

struct FlatGraph {
    std::vector<uint32_t> node_len;
    std::vector<uint64_t> path_first_step{0};
    std::vector<uint32_t> step_node;
    std::vector<uint8_t> step_rev;
    std::vector<uint64_t> step_pos;
    std::vector<std::string> path_names;
    uint64_t max_path_steps = 0, max_path_bp = 0;

    pgsgd_graph_view view() const {
        pgsgd_graph_view v;
        v.node_count = node_len.size();
        v.path_count = path_first_step.size() - 1;
        v.step_count = step_node.size();
        v.node_len = node_len.data();
        v.path_first_step = path_first_step.data();
        v.step_node = step_node.data();
        v.step_rev = step_rev.data();
        v.step_pos = step_pos.data();
        return v;
    }
    uint64_t steps() const { return step_node.size(); }

    void begin_path(const std::string& name) { path_names.push_back(name); cur_bp_ = 0; cur_steps_ = 0; }
    void add_step(uint32_t node_rank, bool is_rev) {
        step_node.push_back(node_rank);
        step_rev.push_back(is_rev ? 1 : 0);
        step_pos.push_back(cur_bp_);   // == XP positions[rank] (xp.cpp:607-616)
        cur_bp_ += node_len[node_rank];
        ++cur_steps_;
    }
    void end_path() {
        path_first_step.push_back(step_node.size());
        if (cur_steps_ > max_path_steps) max_path_steps = cur_steps_;
        if (cur_bp_ > max_path_bp) max_path_bp = cur_bp_;
    }

private:
    uint64_t cur_bp_ = 0, cur_steps_ = 0;
};

// Graph -> FlatGraph.  The graph must be "optimized" (node ids exactly 1..N), as odgi layout/sort require
// (layout_main.cpp:148, layout.cu:320-323); throws std::runtime_error with the reference's message otherwise.
// PathHandle / StepHandle are the graph's handle types (handlegraph::path_handle_t, handlegraph::step_handle_t):
// libhandlegraph's iteratee wrappers need concrete lambda parameter types.
// nthreads > 1 walks the paths concurrently (graph reads are thread-safe in odgi; the reference's GPU host code walks them
// under OpenMP for the same reason, layout.cu:371): step counts give every path its slice up front, then worker threads
// take whole paths from a shared counter.  The result does not depend on nthreads.
template <typename Graph, typename PathHandle, typename StepHandle>
FlatGraph flatten_handle_graph(const Graph& graph, unsigned nthreads = 1) {
    FlatGraph fg;
    const uint64_t N = graph.get_node_count();
    if ((uint64_t) graph.min_node_id() != 1 || (uint64_t) graph.max_node_id() != N) {
        throw std::runtime_error("[odgi::layout] error: the node IDs are not compacted. Please run 'odgi sort' using -O, --optimize to optimize the graph.");
    }
    fg.node_len.resize(N);
    for (uint64_t r = 0; r < N; ++r) fg.node_len[r] = (uint32_t) graph.get_length(graph.get_handle(r + 1, false));
    std::vector<PathHandle> paths;
    graph.for_each_path_handle([&](const PathHandle& path) { paths.push_back(path); });
    fg.path_first_step.assign(1, 0);
    for (const PathHandle& path : paths) {
        fg.path_names.push_back(graph.get_path_name(path));
        fg.path_first_step.push_back(fg.path_first_step.back() + (uint64_t) graph.get_step_count(path));
    }
    const uint64_t S = fg.path_first_step.back();
    fg.step_node.resize(S);
    fg.step_rev.resize(S);
    fg.step_pos.resize(S);
    std::vector<uint64_t> path_bp(paths.size(), 0);
    std::atomic<uint64_t> next{0};
    std::atomic<bool> count_mismatch{false};
    auto work = [&]() {
        for (uint64_t p = next.fetch_add(1); p < paths.size(); p = next.fetch_add(1)) {
            uint64_t i = fg.path_first_step[p], bp = 0;
            const uint64_t end = fg.path_first_step[p + 1];
            graph.for_each_step_in_path(paths[p], [&](const StepHandle& step) {
                if (i >= end) { count_mismatch.store(true); return; }
                const auto h = graph.get_handle_of_step(step);
                const uint32_t rank = (uint32_t) (graph.get_id(h) - 1);
                fg.step_node[i] = rank;
                fg.step_rev[i] = graph.get_is_reverse(h) ? 1 : 0;
                fg.step_pos[i] = bp;   // == XP positions[rank] (xp.cpp:607-616)
                bp += fg.node_len[rank];
                ++i;
            });
            if (i != end) count_mismatch.store(true);
            path_bp[p] = bp;
        }
    };
    if (nthreads <= 1 || paths.size() <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads && t < paths.size(); ++t) pool.emplace_back(work);
        for (auto& th : pool) th.join();
    }
    if (count_mismatch.load()) throw std::runtime_error("[pgsgd::flatten] a path's walk disagrees with its step count");
    for (uint64_t p = 0; p < paths.size(); ++p) {
        const uint64_t c = fg.path_first_step[p + 1] - fg.path_first_step[p];
        if (c > fg.max_path_steps) fg.max_path_steps = c;
        if (path_bp[p] > fg.max_path_bp) fg.max_path_bp = path_bp[p];
    }
    return fg;
}

}  // namespace pgsgd
