// ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A small driver around the UNMODIFIED reference implementation (compiled in place from
// /root/reference by oracle/ref_build/Makefile).  It is used to
//   (1) pin the CPU restatement in oracle/pgsgd_oracle.c against the reference itself
//       (XP integer tables, per-term traces via the reference's own -Deval_path_sgd hook, final coordinates),
//   (2) generate the golden fixtures under tests/golden/ (scripts/make_golden.py),
//   (3) serve as the "reference" CPU baseline timed by bench.py (cpu_baseline.kind == "reference").
//
// It calls only the reference's public entry points:
//   odgi::gfa_to_handle                      (src/gfa_to_handle.hpp:33)
//   xp::XP::from_handle_graph                (src/algorithms/xp.hpp:69)
//   algorithms::path_linear_sgd_layout       (src/algorithms/path_sgd_layout.hpp:37, 2D)
//   algorithms::path_linear_sgd              (src/algorithms/path_sgd.cpp:12, 1D)
//   algorithms::path_linear_sgd_order        (src/algorithms/path_sgd.hpp:66, 1D order)
//
// usage:
//   ref_driver dump   <in.gfa> <out.arr>                        flattened graph + XP tables
//   ref_driver layout <in.gfa> <init.arr|-> <out.arr> [k=v...]  2D PG-SGD (X,Y injected from init.arr)
//   ref_driver sort   <in.gfa> <out.arr> [k=v...]               1D PG-SGD (+ order)
//   ref_driver schedule <eta_max> <iter_max> <iter_lr> <eps>     the reference schedule as hex floats
//   ref_driver lay_write <xy.arr> <out.lay>                      algorithms::layout::Layout(X, Y).serialize (layout.cpp:43-61)
//   ref_driver lay_read  <in.lay> <out.arr>                      Layout::load + get_X / get_Y (layout.cpp:63-113)
// keys: threads iter_max iter_lr updates_x (U = updates_x * sum steps) updates (absolute U) delta eps
//       eta_max theta space space_max space_q cooling order(0/1) freeze_mod (sort: freeze every k-th node rank, as -H does for target paths)
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <unistd.h>
#include <vector>

#include "odgi.hpp"
#include "gfa_to_handle.hpp"
#include "algorithms/xp.hpp"
#include "algorithms/path_sgd_layout.hpp"
#include "algorithms/layout.hpp"
#include <fstream>

#include "../../odgi_b200/host/pgsgd_arrays.hpp"
#include "../../odgi_b200/host/pgsgd_flatten.hpp"

using namespace odgi;

// The reference header still declares a stale 18-argument path_linear_sgd (path_sgd.hpp:40-57); the
// function actually defined (and called by path_linear_sgd_order) takes 20 (path_sgd.cpp:12-31).
namespace odgi { namespace algorithms {
std::vector<double> path_linear_sgd(const graph_t&, const xp::XP&, const std::vector<path_handle_t>&, const uint64_t&, const uint64_t&,
                                    const uint64_t&, const double&, const double&, const double&, const double&, const uint64_t&,
                                    const uint64_t&, const uint64_t&, const double&, const uint64_t&, const bool&, const bool&,
                                    std::vector<std::string>&, const bool&, std::vector<bool>&);
// path_sgd.hpp:66-88 (not included: it and path_sgd_layout.hpp both pull in an unguarded zipfian header)
std::vector<handle_t> path_linear_sgd_order(const graph_t&, const xp::XP&, const std::vector<path_handle_t>&, const uint64_t&, const uint64_t&,
                                            const uint64_t&, const double&, const double&, const double&, const double&, const uint64_t&,
                                            const uint64_t&, const uint64_t&, const double&, const uint64_t&, const bool&, const std::string&,
                                            const bool&, const std::string&, const bool&, const std::string&, const bool&, std::vector<bool>&);
} }

// path_sgd.cpp references utils::graph_deep_copy (only on the snapshot branch, never taken here);
// utils.cpp would drag in the GFAz codec, so the symbol is satisfied by this aborting stand-in.
namespace utils {
void graph_deep_copy(const odgi::graph_t&, odgi::graph_t*) {
    std::cerr << "[ref_driver] graph_deep_copy is not available in the oracle build" << std::endl;
    std::abort();
}
}

static std::map<std::string, std::string> parse_kv(int argc, char** argv, int from) {
    std::map<std::string, std::string> kv;
    for (int i = from; i < argc; ++i) {
        std::string a(argv[i]);
        auto eq = a.find('=');
        if (eq == std::string::npos) { std::cerr << "bad arg " << a << std::endl; std::exit(2); }
        kv[a.substr(0, eq)] = a.substr(eq + 1);
    }
    return kv;
}
static double getd(const std::map<std::string, std::string>& kv, const char* k, double d) {
    auto it = kv.find(k); return it == kv.end() ? d : std::stod(it->second);
}
static uint64_t getu(const std::map<std::string, std::string>& kv, const char* k, uint64_t d) {
    auto it = kv.find(k); return it == kv.end() ? d : std::stoull(it->second);
}

struct Loaded {
    graph_t graph;
    xp::XP xp;
    std::vector<path_handle_t> paths;
    uint64_t sum_steps = 0, max_steps = 0, max_path_bp = 0;
};

static void load(const std::string& gfa, uint64_t threads, Loaded& L, bool build_xp) {
    gfa_to_handle(gfa, &L.graph, false, threads, false);
    L.graph.set_number_of_threads(threads);
    if (!L.graph.is_optimized()) { std::cerr << "[ref_driver] graph is not optimized" << std::endl; std::exit(1); }
    char cwd[512];
    if (!getcwd(cwd, sizeof(cwd))) std::exit(1);
    xp::temp_file::set_dir(std::string(cwd));
    if (build_xp) L.xp.from_handle_graph(L.graph, threads);
    L.graph.for_each_path_handle([&](const path_handle_t& p) {
        L.paths.push_back(p);
        uint64_t c = L.graph.get_step_count(p);
        L.sum_steps += c;
        L.max_steps = std::max(L.max_steps, c);
        if (build_xp) L.max_path_bp = std::max<uint64_t>(L.max_path_bp, L.xp.get_path_length(p));
    });
}

// graph_t walk -> path-major arrays (the same walk the reference GPU host code does, src/cuda/layout.cu:371-410)
static void flatten(const graph_t& g, const std::vector<path_handle_t>& paths, std::vector<uint32_t>& node_len,
                    std::vector<uint64_t>& path_first, std::vector<uint32_t>& step_node, std::vector<uint8_t>& step_rev,
                    std::vector<uint64_t>& step_pos) {
    uint64_t N = g.get_node_count();
    node_len.resize(N);
    for (uint64_t r = 0; r < N; ++r) node_len[r] = g.get_length(g.get_handle(r + 1, false));
    path_first.assign(1, 0);
    for (auto& p : paths) {
        uint64_t pos = 0;
        g.for_each_step_in_path(p, [&](const step_handle_t& s) {
            handle_t h = g.get_handle_of_step(s);
            step_node.push_back((uint32_t)(g.get_id(h) - 1));
            step_rev.push_back(g.get_is_reverse(h) ? 1 : 0);
            step_pos.push_back(pos);
            pos += g.get_length(h);
        });
        path_first.push_back(step_node.size());
    }
}

static int cmd_dump(int argc, char** argv) {
    if (argc < 4) return 2;
    Loaded L;
    load(argv[2], 1, L, true);  // 1 thread: the per-node step slots (hence XP's node-major order) depend on loader interleaving
    std::vector<uint32_t> node_len, step_node;
    std::vector<uint64_t> path_first, step_pos;
    std::vector<uint8_t> step_rev;
    flatten(L.graph, L.paths, node_len, path_first, step_node, step_rev, step_pos);
    // XP's view of the same data: the accessors the CPU workers use (path_sgd_layout.cpp:186,199,242-249)
    const sdsl::int_vector<>& nr = L.xp.get_nr_iv();
    const sdsl::int_vector<>& npi = L.xp.get_npi_iv();
    std::vector<uint64_t> xp_nr(nr.size()), xp_npi(npi.size());
    for (uint64_t i = 0; i < nr.size(); ++i) { xp_nr[i] = nr[i]; xp_npi[i] = npi[i]; }
    std::vector<uint64_t> xp_pos, xp_handle, xp_path_id, xp_path_len;
    std::vector<uint8_t> path_names;  // '\n'-joined, in path order
    for (auto& p : L.paths) {
        for (char ch : L.graph.get_path_name(p)) path_names.push_back((uint8_t) ch);
        path_names.push_back((uint8_t) '\n');
        uint64_t c = L.xp.get_path_step_count(p);
        xp_path_id.push_back(as_integer(p));
        xp_path_len.push_back(L.xp.get_path_length(p));
        for (uint64_t r = 0; r < c; ++r) {
            step_handle_t s;
            as_integers(s)[0] = as_integer(p);
            as_integers(s)[1] = r;
            xp_pos.push_back(L.xp.get_position_of_step(s));
            xp_handle.push_back(as_integer(L.xp.get_handle_of_step(s)));
        }
    }
    // the product's own walk over the reference's graph_t (the template the odgi shim instantiates): must equal both
    const pgsgd::FlatGraph shim = pgsgd::flatten_handle_graph<graph_t, path_handle_t, step_handle_t>(L.graph);
    // ... and does not depend on the number of walker threads
    const pgsgd::FlatGraph shim4 = pgsgd::flatten_handle_graph<graph_t, path_handle_t, step_handle_t>(L.graph, 4);
    if (shim4.step_node != shim.step_node || shim4.step_rev != shim.step_rev || shim4.step_pos != shim.step_pos ||
        shim4.path_first_step != shim.path_first_step || shim4.node_len != shim.node_len || shim4.path_names != shim.path_names ||
        shim4.max_path_steps != shim.max_path_steps || shim4.max_path_bp != shim.max_path_bp)
        throw std::runtime_error("flatten_handle_graph: the 4-thread walk differs from the 1-thread walk");
    pgsgd::ArrayWriter w(argv[3]);
    w.add("shim_node_len", shim.node_len);
    w.add("shim_path_first_step", shim.path_first_step);
    w.add("shim_step_node", shim.step_node);
    w.add("shim_step_rev", shim.step_rev);
    w.add("shim_step_pos", shim.step_pos);
    w.add("node_len", node_len);
    w.add("path_first_step", path_first);
    w.add("step_node", step_node);
    w.add("step_rev", step_rev);
    w.add("step_pos", step_pos);
    w.add("xp_nr_iv", xp_nr);
    w.add("xp_npi_iv", xp_npi);
    w.add("xp_position_of_step", xp_pos);
    w.add("xp_handle_of_step", xp_handle);
    w.add("xp_path_id", xp_path_id);
    w.add("xp_path_length", xp_path_len);
    w.add("path_names", path_names);
    w.close();
    std::cout << "{\"nodes\": " << node_len.size() << ", \"paths\": " << L.paths.size() << ", \"steps\": " << step_node.size() << "}" << std::endl;
    return 0;
}

static int cmd_layout(int argc, char** argv) {
    if (argc < 5) return 2;
    auto kv = parse_kv(argc, argv, 5);
    uint64_t threads = getu(kv, "threads", 1);
    Loaded L;
    load(argv[2], threads, L, true);
    uint64_t N = L.graph.get_node_count();
    // defaults: src/subcommand/layout_main.cpp:198-266
    uint64_t iter_max = getu(kv, "iter_max", 30);
    uint64_t iter_lr = getu(kv, "iter_lr", 0);
    uint64_t U = kv.count("updates") ? getu(kv, "updates", 0) : (uint64_t)(getd(kv, "updates_x", 10.0) * L.sum_steps);
    double delta = getd(kv, "delta", 0), eps = getd(kv, "eps", 0.01);
    double eta_max = getd(kv, "eta_max", (double) L.max_steps * L.max_steps);
    double theta = getd(kv, "theta", 0.99);
    uint64_t space = getu(kv, "space", L.max_steps);
    uint64_t space_max = getu(kv, "space_max", 1000);
    uint64_t space_q = getu(kv, "space_q", 100);
    double cooling = getd(kv, "cooling", 0.5);

    std::vector<std::atomic<double>> X(2 * N), Y(2 * N);
    std::string init = argv[3];
    if (init == "-") {
        // 'd' initialisation without the noise (layout_main.cpp:322-328 with Y = 0)
        uint64_t len = 0;
        for (uint64_t r = 0; r < N; ++r) {
            X[2 * r].store(len); Y[2 * r].store(0);
            len += L.graph.get_length(L.graph.get_handle(r + 1, false));
            X[2 * r + 1].store(len); Y[2 * r + 1].store(0);
        }
    } else {
        auto arrs = pgsgd::read_arrays(init);
        const double* x = arrs.at("X").as<double>();
        const double* y = arrs.at("Y").as<double>();
        if (arrs.at("X").count != 2 * N) { std::cerr << "init size mismatch" << std::endl; return 1; }
        for (uint64_t i = 0; i < 2 * N; ++i) { X[i].store(x[i]); Y[i].store(y[i]); }
    }
    auto t0 = std::chrono::steady_clock::now();
    algorithms::path_linear_sgd_layout(L.graph, L.xp, L.paths, iter_max, iter_lr, U, delta, eps, eta_max, theta, space,
                                       space_max, space_q, cooling, threads, false, false, "", X, Y);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<double> x(2 * N), y(2 * N);
    for (uint64_t i = 0; i < 2 * N; ++i) { x[i] = X[i].load(); y[i] = Y[i].load(); }
    pgsgd::ArrayWriter w(argv[4]);
    w.add("X", x);
    w.add("Y", y);
    w.add_scalar<double>("seconds", secs);
    w.add_scalar<uint64_t>("nominal_updates", iter_max * U);
    w.close();
    std::cout << "{\"mode\": \"layout\", \"threads\": " << threads << ", \"iter_max\": " << iter_max << ", \"updates_per_iter\": " << U
              << ", \"nominal_updates\": " << iter_max * U << ", \"seconds\": " << secs
              << ", \"updates_per_sec\": " << (double)(iter_max * U) / secs << "}" << std::endl;
    return 0;
}

static int cmd_sort(int argc, char** argv) {
    if (argc < 4) return 2;
    auto kv = parse_kv(argc, argv, 4);
    uint64_t threads = getu(kv, "threads", 1);
    Loaded L;
    load(argv[2], threads, L, true);
    uint64_t N = L.graph.get_node_count();
    // defaults: src/subcommand/sort_main.cpp:313-414
    uint64_t iter_max = getu(kv, "iter_max", 100);
    uint64_t iter_lr = getu(kv, "iter_lr", 0);
    uint64_t U = kv.count("updates") ? getu(kv, "updates", 0) : (uint64_t)(getd(kv, "updates_x", 1.0) * L.sum_steps);
    double delta = getd(kv, "delta", 0), eps = getd(kv, "eps", 0.01);
    double eta_max = getd(kv, "eta_max", (double) L.max_steps * L.max_steps);
    double theta = getd(kv, "theta", 0.99);
    uint64_t space = getu(kv, "space", L.max_path_bp);
    uint64_t space_max = getu(kv, "space_max", 100);
    uint64_t space_q;
    if (kv.count("space_q")) {
        space_q = getu(kv, "space_q", 100);
    } else {  // sort_main.cpp:390-412
        uint64_t max_dists = std::max<uint64_t>(space_max + 1, 100);
        space_q = std::max<uint64_t>(2, (uint64_t) std::ceil((double)(space - space_max) / (double)(max_dists - space_max)));
    }
    double cooling = getd(kv, "cooling", 0.5);
    bool want_order = getu(kv, "order", 0) != 0;
    // `odgi sort -H`: nodes of the target paths are frozen (sort_main.cpp:266-311 fills target_nodes); here every
    // freeze_mod-th node rank is frozen so the frozen branch (path_sgd.cpp:290-302,387-392) can be pinned
    const uint64_t freeze_mod = getu(kv, "freeze_mod", 0);
    std::vector<bool> target_nodes;
    if (freeze_mod) {
        target_nodes.resize(N);
        for (uint64_t i = 0; i < N; ++i) target_nodes[i] = (i % freeze_mod) == 0;
    }
    const bool target_sorting = freeze_mod != 0;
    std::vector<std::string> snapshots;
    std::vector<double> x;
    std::vector<uint64_t> order;
    auto t0 = std::chrono::steady_clock::now();
    if (want_order) {
        // The order is derived from an X the function does not return; the iteration boundaries depend on a 1 ms poll, so a
        // second run would not reproduce it even single-threaded.  The reference's own 1D .lay output (path_sgd.cpp:660-678:
        // sorted position i holds the start coordinate of the i-th node of the order) gives the X of THIS run.
        const std::string lay_path = std::string(argv[3]) + ".lay";
        std::vector<handle_t> o = algorithms::path_linear_sgd_order(L.graph, L.xp, L.paths, iter_max, iter_lr, U, delta, eps, eta_max, theta,
                                                                    space, space_max, space_q, cooling, threads, false, "", false, "",
                                                                    true, lay_path, target_sorting, target_nodes);
        for (auto& h : o) order.push_back(as_integer(h));
        algorithms::layout::Layout lay;
        std::ifstream lf(lay_path, std::ios::binary);
        lay.load(lf);
        const std::vector<double> sx = lay.get_X();
        x.assign(N, 0.0);
        for (uint64_t i = 0; i < o.size(); ++i) x[as_integer(o[i]) >> 1] = sx[2 * i];
        unlink(lay_path.c_str());
    } else {
        x = algorithms::path_linear_sgd(L.graph, L.xp, L.paths, iter_max, iter_lr, U, delta, eps, eta_max, theta, space, space_max,
                                        space_q, cooling, threads, false, false, snapshots, target_sorting, target_nodes);
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    pgsgd::ArrayWriter w(argv[3]);
    w.add("X", x);
    w.add("order", order);
    w.add_scalar<double>("seconds", secs);
    w.add_scalar<uint64_t>("nominal_updates", (iter_max + 1) * U);
    w.add_scalar<uint64_t>("space", space);
    w.add_scalar<uint64_t>("space_q", space_q);
    w.close();
    std::cout << "{\"mode\": \"sort\", \"threads\": " << threads << ", \"iter_max\": " << iter_max << ", \"updates_per_iter\": " << U
              << ", \"nominal_updates\": " << (iter_max + 1) * U << ", \"seconds\": " << secs
              << ", \"updates_per_sec\": " << (double)((iter_max + 1) * U) / secs << ", \"space\": " << space << ", \"space_q\": " << space_q
              << ", \"N\": " << N << "}" << std::endl;
    return 0;
}

// prints the reference's learning-rate schedule (path_sgd_layout.cpp:433-468) as exact hex floats
static int cmd_schedule(int argc, char** argv) {
    if (argc < 6) return 2;
    double eta_max = std::stod(argv[2]);
    uint64_t iter_max = std::stoull(argv[3]), iter_lr = std::stoull(argv[4]);
    double eps = std::stod(argv[5]);
    double w_min = (double) 1.0 / (double) (eta_max);  // as the caller computes it, path_sgd_layout.cpp:75
    std::vector<double> etas = algorithms::path_linear_sgd_layout_schedule(w_min, 1.0, iter_max, iter_lr, eps);
    for (double e : etas) std::printf("%a\n", e);
    return 0;
}

// the .lay container exactly as `odgi layout -o` writes it / `odgi draw` reads it
static int cmd_lay_write(int argc, char** argv) {
    if (argc < 4) return 2;
    auto arrs = pgsgd::read_arrays(argv[2]);
    algorithms::layout::Layout lay(arrs.at("X").vec<double>(), arrs.at("Y").vec<double>());
    std::ofstream f(argv[3], std::ios::binary);
    lay.serialize(f);
    return 0;
}

static int cmd_lay_read(int argc, char** argv) {
    if (argc < 4) return 2;
    algorithms::layout::Layout lay;
    std::ifstream f(argv[2], std::ios::binary);
    lay.load(f);
    pgsgd::ArrayWriter w(argv[3]);
    w.add("X", lay.get_X());
    w.add("Y", lay.get_Y());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: ref_driver dump|layout|sort ..." << std::endl; return 2; }
    std::string cmd = argv[1];
    try {
        if (cmd == "dump") return cmd_dump(argc, argv);
        if (cmd == "layout") return cmd_layout(argc, argv);
        if (cmd == "sort") return cmd_sort(argc, argv);
        if (cmd == "schedule") return cmd_schedule(argc, argv);
        if (cmd == "lay_write") return cmd_lay_write(argc, argv);
        if (cmd == "lay_read") return cmd_lay_read(argc, argv);
    } catch (const std::exception& e) {
        std::cerr << "[ref_driver] error: " << e.what() << std::endl;
        return 1;
    }
    std::cerr << "unknown command " << cmd << std::endl;
    return 2;
}
