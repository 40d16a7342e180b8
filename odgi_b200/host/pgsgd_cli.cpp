// pgsgd_cli.cpp — `pgsgd layout|sort ... --gpu`: the `odgi layout` / `odgi sort -Y` command-line surface for the PG-SGD
// path (same short/long flags and defaults: src/subcommand/layout_main.cpp:28-107,198-266; sort_main.cpp:313-414),
// standalone: GFA in -> flattened graph -> C-ABI (include/pgsgd.h) -> .lay / TSV layout / node order out.
// Everything that is not the PG-SGD path (the .og container, other sort pipelines, drawing) stays in odgi.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "gfa_lite.hpp"
#include "lay_format.hpp"
#include "pgsgd_arrays.hpp"

namespace {

struct Args {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const { return kv.count(k) != 0; }
    std::string str(const std::string& k, const std::string& d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    double num(const std::string& k, double d) const { auto it = kv.find(k); return it == kv.end() ? d : std::stod(it->second); }
    uint64_t u64(const std::string& k, uint64_t d) const { auto it = kv.find(k); return it == kv.end() ? d : std::stoull(it->second); }
};

// flag table: short, long, takes value
struct Flag { const char* s; const char* l; bool val; };
const Flag LAYOUT_FLAGS[] = {
    {"i", "idx", true}, {"o", "out", true}, {"T", "tsv", true}, {"N", "layout-initialization", true},
    {"G", "path-sgd-min-term-updates-paths", true}, {"U", "path-sgd-min-term-updates-nodes", true}, {"j", "path-sgd-delta", true},
    {"g", "path-sgd-eps", true}, {"v", "path-sgd-eta-max", true}, {"a", "path-sgd-zipf-theta", true}, {"x", "path-sgd-iter-max", true},
    {"K", "path-sgd-cooling", true}, {"F", "path-sgd-iteration-max-learning-rate", true}, {"k", "path-sgd-zipf-space", true},
    {"I", "path-sgd-zipf-space-max", true}, {"l", "path-sgd-zipf-space-quantization-step", true}, {"t", "threads", true},
    {"", "gpu", false}, {"P", "progress", false}, {"h", "help", false}, {"", "seed", true}, {"", "init-seed", true}, {"", "sampling", true},
    {"u", "path-sgd-snapshot", true}, {"f", "path-sgd-use-paths", true}, {"", "device-ingest", false}, {"", "timing", false}};
const Flag SORT_FLAGS[] = {
    {"i", "idx", true}, {"o", "out", true}, {"Y", "path-sgd", false}, {"G", "path-sgd-min-term-updates-paths", true},
    {"U", "path-sgd-min-term-updates-nodes", true}, {"j", "path-sgd-delta", true}, {"g", "path-sgd-eps", true},
    {"v", "path-sgd-eta-max", true}, {"a", "path-sgd-zipf-theta", true}, {"x", "path-sgd-iter-max", true}, {"K", "path-sgd-cooling", true},
    {"F", "path-sgd-iteration-max-learning-rate", true}, {"k", "path-sgd-zipf-space", true}, {"I", "path-sgd-zipf-space-max", true},
    {"l", "path-sgd-zipf-space-quantization-step", true}, {"t", "threads", true}, {"", "gpu", false}, {"P", "progress", false},
    {"h", "help", false}, {"", "seed", true}, {"", "sampling", true}, {"", "layout-out", true}, {"e", "path-sgd-layout", true},
    {"f", "path-sgd-use-paths", true}, {"H", "target-paths", true}, {"", "prepared-out", true}};

template <size_t NF>
bool parse(int argc, char** argv, const Flag (&flags)[NF], Args& a, const char* sub) {
    for (int i = 2; i < argc; ++i) {
        std::string tok = argv[i], name, value;
        bool have_value = false;
        if (tok.rfind("--", 0) == 0) {
            name = tok.substr(2);
            auto eq = name.find('=');
            if (eq != std::string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); have_value = true; }
        } else if (tok.size() >= 2 && tok[0] == '-') {
            name = tok.substr(1, 1);
            if (tok.size() > 2) { value = tok.substr(tok[2] == '=' ? 3 : 2); have_value = true; }
        } else {
            std::cerr << "[odgi::" << sub << "] error: unexpected argument " << tok << std::endl;
            return false;
        }
        const Flag* f = nullptr;
        for (const Flag& c : flags) if (name == c.l || (c.s[0] && name == c.s)) { f = &c; break; }
        if (!f) { std::cerr << "[odgi::" << sub << "] error: unknown flag " << tok << std::endl; return false; }
        if (f->val && !have_value) {
            if (i + 1 >= argc) { std::cerr << "[odgi::" << sub << "] error: flag " << tok << " needs a value" << std::endl; return false; }
            value = argv[++i];
        }
        a.kv[f->l] = f->val ? value : "1";
    }
    return true;
}

// weakly connected components over path adjacencies and L lines (union-find); component ids in order of first node
struct Components {
    std::vector<uint32_t> parent;
    explicit Components(size_t n) : parent(n) { std::iota(parent.begin(), parent.end(), 0u); }
    uint32_t find(uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
    void unite(uint32_t a, uint32_t b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
};

std::vector<uint32_t> components_of(const pgsgd::FlatGraph& fg, const std::string& gfa) {
    Components uf(fg.node_len.size());
    for (size_t p = 0; p + 1 < fg.path_first_step.size(); ++p)
        for (uint64_t s = fg.path_first_step[p]; s + 1 < fg.path_first_step[p + 1]; ++s) uf.unite(fg.step_node[s], fg.step_node[s + 1]);
    std::ifstream in(gfa);
    std::string line;
    while (std::getline(in, line)) {
        if (line.size() < 2 || line[0] != 'L') continue;
        std::istringstream ss(line);
        std::string tag, a, ao, b;
        ss >> tag >> a >> ao >> b;
        const uint64_t ia = std::strtoull(a.c_str(), nullptr, 10), ib = std::strtoull(b.c_str(), nullptr, 10);
        if (ia >= 1 && ib >= 1 && ia <= fg.node_len.size() && ib <= fg.node_len.size()) uf.unite((uint32_t) ia - 1, (uint32_t) ib - 1);
    }
    std::vector<uint32_t> comp(fg.node_len.size());
    std::map<uint32_t, uint32_t> ids;
    for (uint32_t i = 0; i < comp.size(); ++i) {
        const uint32_t r = uf.find(i);
        auto it = ids.find(r);
        if (it == ids.end()) it = ids.emplace(r, (uint32_t) ids.size()).first;
        comp[i] = it->second;
    }
    return comp;
}

// what the default schedule parameters are derived from: all paths, or the ones named by -f (sort_main.cpp:355-387; the
// sampler itself always draws from every path, path_sgd.cpp:82)
struct PathStats { uint64_t sum_steps = 0, max_steps = 0, max_bp = 0; };

PathStats path_stats(const pgsgd::FlatGraph& fg, const std::vector<uint64_t>* only = nullptr) {
    PathStats ps;
    auto add = [&](uint64_t p) {
        const uint64_t lo = fg.path_first_step[p], hi = fg.path_first_step[p + 1], cnt = hi - lo;
        const uint64_t bp = cnt ? fg.step_pos[hi - 1] + fg.node_len[fg.step_node[hi - 1]] : 0;
        ps.sum_steps += cnt;
        ps.max_steps = std::max(ps.max_steps, cnt);
        ps.max_bp = std::max(ps.max_bp, bp);
    };
    if (only) for (uint64_t p : *only) add(p);
    else for (uint64_t p = 0; p + 1 < fg.path_first_step.size(); ++p) add(p);
    return ps;
}

void common_config(const Args& a, const pgsgd::FlatGraph& fg, bool is_sort, pgsgd_config& c, const PathStats* use = nullptr) {
    std::memset(&c, 0, sizeof(c));
    const PathStats all = use ? *use : path_stats(fg);
    const uint64_t S = all.sum_steps, N = fg.node_len.size();
    c.iter_max = a.u64("path-sgd-iter-max", is_sort ? 100 : 30);
    c.iter_with_max_learning_rate = a.u64("path-sgd-iteration-max-learning-rate", 0);
    c.theta = a.num("path-sgd-zipf-theta", 0.99);
    c.eps = a.num("path-sgd-eps", 0.01);
    c.delta = a.num("path-sgd-delta", 0);
    c.cooling_start = a.num("path-sgd-cooling", 0.5);
    if (a.has("path-sgd-min-term-updates-paths")) c.min_term_updates = (uint64_t) (a.num("path-sgd-min-term-updates-paths", 0) * (double) S);
    else if (a.has("path-sgd-min-term-updates-nodes")) c.min_term_updates = (uint64_t) (a.num("path-sgd-min-term-updates-nodes", 0) * (double) N);
    else c.min_term_updates = (uint64_t) ((is_sort ? 1.0 : 10.0) * (double) S);
    c.eta_max = a.has("path-sgd-eta-max") ? a.num("path-sgd-eta-max", 0) : (double) all.max_steps * (double) all.max_steps;
    if (is_sort) {  // sort_main.cpp:387-412: -k / -I are taken as given (no clamp), q = 100 unless space > space_max
        const uint64_t max_len = all.max_bp;
        c.space = a.has("path-sgd-zipf-space") && a.u64("path-sgd-zipf-space", 0) ? a.u64("path-sgd-zipf-space", 0) : max_len;
        c.space_max = a.has("path-sgd-zipf-space-max") && a.u64("path-sgd-zipf-space-max", 0) ? a.u64("path-sgd-zipf-space-max", 0) : 100;
        if (a.has("path-sgd-zipf-space-quantization-step") && a.u64("path-sgd-zipf-space-quantization-step", 0)) {
            c.space_quantization_step = std::max<uint64_t>(2, a.u64("path-sgd-zipf-space-quantization-step", 0));
        } else {
            const uint64_t max_dists = std::max<uint64_t>(c.space_max + 1, 100);   // MAX_NUMBER_OF_ZIPF_DISTRIBUTIONS
            if (c.space > c.space_max && max_dists > c.space_max)
                c.space_quantization_step = std::max<uint64_t>(2, (uint64_t) std::ceil((double) (c.space - c.space_max) / (double) (max_dists - c.space_max)));
            else
                c.space_quantization_step = 100;
        }
    } else {  // layout_main.cpp:261-266
        c.space = a.has("path-sgd-zipf-space") ? std::min(a.u64("path-sgd-zipf-space", 0), all.max_steps) : all.max_steps;
        c.space_max = a.has("path-sgd-zipf-space-max") ? std::min(c.space, a.u64("path-sgd-zipf-space-max", 0)) : 1000;
        c.space_quantization_step = a.has("path-sgd-zipf-space-quantization-step") ? std::max<uint64_t>(2, a.u64("path-sgd-zipf-space-quantization-step", 0)) : 100;
    }
    c.seed = a.u64("seed", 9399220);
    c.sampling = (uint32_t) a.u64("sampling", 0);
}

int need_gpu(const Args& a, const char* sub) {
    if (pgsgd_version() != PGSGD_VERSION) {
        std::cerr << "[odgi::" << sub << "] error: libpgsgd_b200 ABI " << pgsgd_version() << ", this binary was built against " << PGSGD_VERSION << ": rebuild." << std::endl;
        return 1;
    }
    if (!a.has("gpu")) {
        std::cerr << "[odgi::" << sub << "] error: this build provides only the GPU path of the path-guided SGD; pass --gpu "
                     "(the CPU path is odgi's own: src/algorithms/path_sgd" << (std::string(sub) == "layout" ? "_layout" : "") << ".cpp)." << std::endl;
        return 1;
    }
    if (pgsgd_device_count() < 1) {
        std::cerr << "[odgi::" << sub << "] error: --gpu given but no usable CUDA device was found." << std::endl;
        return 1;
    }
    return 0;
}

// Hilbert curve: distance d along the curve -> cell (x, y).  The reference walks quadrant sizes s = 1, 2, 4, ... < n
// with n = 2 * node count, not rounded to a power of two (src/algorithms/hilbert.hpp:30-41, layout_main.cpp:287,312-318).
void hilbert_d2xy(uint64_t n, uint64_t d, uint64_t& x, uint64_t& y) {
    x = y = 0;
    for (uint64_t s = 1, t = d; s < n; s <<= 1, t >>= 2) {
        const uint64_t rx = (t >> 1) & 1, ry = (t ^ rx) & 1;
        if (ry == 0) {
            if (rx == 1) { x = s - 1 - x; y = s - 1 - y; }
            std::swap(x, y);
        }
        x += s * rx;
        y += s * ry;
    }
}

// the five coordinate initialisations of `odgi layout -N` (layout_main.cpp:268-330); the reference seeds its mt19937 from
// std::random_device, --init-seed makes runs reproducible
bool init_layout(const pgsgd::FlatGraph& fg, char mode, bool seeded, uint64_t seed, std::vector<double>& X, std::vector<double>& Y) {
    const uint64_t N = fg.node_len.size();
    X.assign(2 * N, 0.0);
    Y.assign(2 * N, 0.0);
    std::mt19937 rng(seeded ? (uint32_t) seed : std::random_device{}());
    std::uniform_real_distribution<double> uniform_noise(0, std::sqrt((double) N * 2));
    std::normal_distribution<double> gaussian_noise(0, std::sqrt((double) N * 2));
    uint64_t total_length = 0;
    for (uint32_t l : fg.node_len) total_length += l;
    std::uniform_real_distribution<double> uniform_noise_in_length(0, (double) total_length);
    uint64_t len = 0;
    for (uint64_t r = 0; r < N; ++r) {
        const uint64_t pos = 2 * r;
        switch (mode) {
            case 'g': X[pos] = gaussian_noise(rng); Y[pos] = gaussian_noise(rng); X[pos + 1] = gaussian_noise(rng); Y[pos + 1] = gaussian_noise(rng); break;
            case 'u': X[pos] = (double) len; Y[pos] = uniform_noise(rng); len += fg.node_len[r]; X[pos + 1] = (double) len; Y[pos + 1] = uniform_noise(rng); break;
            case 'r': X[pos] = uniform_noise_in_length(rng); Y[pos] = uniform_noise_in_length(rng); X[pos + 1] = uniform_noise_in_length(rng); Y[pos + 1] = uniform_noise_in_length(rng); break;
            case 'h': {
                uint64_t x, y;
                hilbert_d2xy(2 * N, pos, x, y); X[pos] = (double) x; Y[pos] = (double) y;
                hilbert_d2xy(2 * N, pos + 1, x, y); X[pos + 1] = (double) x; Y[pos + 1] = (double) y;
                break;
            }
            case 'd': X[pos] = (double) len; Y[pos] = gaussian_noise(rng); len += fg.node_len[r]; X[pos + 1] = (double) len; Y[pos + 1] = gaussian_noise(rng); break;
            default:
                std::cerr << "[odgi::layout] error: unknown layout initialization '" << mode << "' (d, r, u, g, h)." << std::endl;
                return false;
        }
    }
    return true;
}

bool read_path_list(const std::string& file, const pgsgd::FlatGraph& fg, bool reject_duplicates, bool unknown_is_error, std::vector<uint64_t>& out);

// `pgsgd layout --device-ingest`: the whole file-to-file path with no per-step and no per-node-coordinate host work —
// the host maps the GFA and finds its lines (scan_gfa), the P-line step lists are parsed and flattened on the GPU
// (pgsgd_engine_create_from_gfa_paths), the layout runs there, and the component stacking + `.lay` encoding
// (layout_main.cpp:402-463) happen there too (pgsgd_engine_encode_lay).  Weak components come from the L lines.
// --timing prints where the wall clock went as one JSON line on stderr.
int layout_device_ingest(const Args& a) {
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    std::thread warm([]() { pgsgd_device_warmup(0); });   // the CUDA context comes up while the host reads the file
    pgsgd::GfaIndex ix;
    try { pgsgd::scan_gfa(a.str("idx"), ix); } catch (const std::exception& e) { warm.join(); std::cerr << e.what() << std::endl; return 1; }
    warm.join();
    const double t_scan = now();
    pgsgd::FlatGraph fg;               // node table only: what the defaults and the initialisation read
    fg.node_len = ix.node_len;
    const uint64_t N = fg.node_len.size();
    std::vector<double> X, Y;
    bool init_ok = false;              // the initial layout needs the node table only: drawn while the device parses the paths
    std::thread init([&]() { init_ok = init_layout(fg, a.str("layout-initialization", "d")[0], a.has("init-seed"), a.u64("init-seed", 0), X, Y); });
    pgsgd_engine* e = nullptr;
    const int rc_create = pgsgd_engine_create_from_gfa_paths(ix.node_len.data(), ix.node_len.size(), ix.text, ix.field_begin.data(), ix.field_end.data(),
                                                             ix.field_begin.size(), 0, &e);
    const double t_engine = now();
    init.join();
    if (rc_create != PGSGD_OK) {
        std::cerr << "[odgi::layout] error: " << pgsgd_last_error() << std::endl;
        return 1;
    }
    PathStats ps;
    pgsgd_engine_graph_stats(e, &ps.sum_steps, &ps.max_steps, &ps.max_bp, nullptr);
    pgsgd_config c;
    common_config(a, fg, false, c, &ps);
    int rc = init_ok ? 0 : 1;
    pgsgd_stats st;
    std::memset(&st, 0, sizeof(st));
    if (!rc) rc = pgsgd_engine_set_coords_2d(e, X.data(), Y.data());
    const double t_init = now();
    if (!rc) rc = pgsgd_engine_run_2d(e, &c, &st);
    const double t_run = now();
    // weak components from the links (union-find over node ranks), numbered by their smallest node as components_of does
    Components uf(N);
    for (auto& l : ix.links) if (l.first < N && l.second < N) uf.unite(l.first, l.second);
    std::vector<uint32_t> comp(N);
    uint32_t n_comp = 0;
    {
        std::map<uint32_t, uint32_t> ids;
        for (uint32_t i = 0; i < N; ++i) {
            const uint32_t r = uf.find(i);
            auto it = ids.find(r);
            if (it == ids.end()) it = ids.emplace(r, (uint32_t) ids.size()).first;
            comp[i] = it->second;
        }
        n_comp = (uint32_t) ids.size();
    }
    uint64_t lay_bytes = 0;
    if (!rc && a.has("out")) {
        rc = pgsgd_engine_encode_lay(e, comp.data(), n_comp, nullptr, 0, &lay_bytes);
        std::vector<uint8_t> buf(lay_bytes);
        if (!rc) rc = pgsgd_engine_encode_lay(e, comp.data(), n_comp, buf.data(), buf.size(), &lay_bytes);
        if (!rc) {
            if (a.str("out") == "-") std::cout.write((const char*) buf.data(), (std::streamsize) buf.size());
            else { std::ofstream f(a.str("out"), std::ios::binary); f.write((const char*) buf.data(), (std::streamsize) buf.size()); }
        }
    }
    if (!rc && a.has("tsv")) {   // the text form needs the coordinates on the host: download + the host stacking, as without the flag
        rc = pgsgd_engine_get_coords_2d(e, X.data(), Y.data());
        if (!rc) {
            const double border = 1000.0, inf = std::numeric_limits<double>::max();
            std::vector<double> min_x(n_comp, inf), min_y(n_comp, inf), max_y(n_comp, std::numeric_limits<double>::lowest());
            for (uint64_t r = 0; r < N; ++r)
                for (uint64_t j = 2 * r; j <= 2 * r + 1; ++j) {
                    min_x[comp[r]] = std::min(min_x[comp[r]], X[j]); min_y[comp[r]] = std::min(min_y[comp[r]], Y[j]); max_y[comp[r]] = std::max(max_y[comp[r]], Y[j]);
                }
            double curr = border;
            std::vector<double> x_off(n_comp), y_off(n_comp);
            for (uint32_t k = 0; k < n_comp; ++k) { x_off[k] = min_x[k] - border; y_off[k] = curr - min_y[k]; curr += (max_y[k] - min_y[k]) + border; }
            std::ofstream fout;
            std::ostream* out = &std::cout;
            if (a.str("tsv") != "-") { fout.open(a.str("tsv")); out = &fout; }
            *out << std::setprecision(std::numeric_limits<double>::digits10 + 1) << "idx\tX\tY\tcomponent" << std::endl;
            for (uint32_t k = 0; k < n_comp; ++k)
                for (uint64_t r = 0; r < N; ++r) {
                    if (comp[r] != k) continue;
                    for (uint64_t j = 2 * r; j <= 2 * r + 1; ++j) *out << j << "\t" << X[j] - x_off[k] << "\t" << Y[j] + y_off[k] << "\t" << k << '\n';
                }
        }
    }
    const double t_out = now();
    if (e) pgsgd_engine_destroy(e);
    if (rc) { std::cerr << "[odgi::layout] error: " << pgsgd_last_error() << std::endl; return 1; }
    if (a.has("progress") || a.has("timing"))
        std::fprintf(stderr, "{\"ingest\": \"device\", \"nodes\": %llu, \"paths\": %llu, \"steps\": %llu, \"gfa_bytes\": %llu, \"scan_lines_s\": %.4f, "
                             "\"upload_parse_flatten_s\": %.4f, \"init_coords_s\": %.4f, \"iterations_s\": %.4f, \"run_call_s\": %.4f, "
                             "\"components_stack_encode_write_s\": %.4f, \"lay_bytes\": %llu, \"total_s\": %.4f, \"term_updates\": %llu}\n",
                     (unsigned long long) N, (unsigned long long) ix.field_begin.size(), (unsigned long long) ps.sum_steps, (unsigned long long) ix.bytes,
                     t_scan - t0, t_engine - t_scan, t_init - t_engine, st.seconds_iterations, t_run - t_init, t_out - t_run,
                     (unsigned long long) lay_bytes, t_out - t0, (unsigned long long) st.term_updates);
    return 0;
}

int main_layout(int argc, char** argv) {
    Args a;
    if (!parse(argc, argv, LAYOUT_FLAGS, a, "layout") || a.has("help") || argc == 2) {
        std::cout << "pgsgd layout -i g.gfa (-o out.lay | -T out.tsv) --gpu [-u snapshot_prefix] [-x N] [-G N|-U N] [-j N] [-g N] [-v N] [-a N] [-K N] [-F N] [-k N] [-I N] [-l N] [-N d|r|u|g|h] [-t N] [-P]\n"
                     "  the `odgi layout` PG-SGD flags with the same defaults; --seed N (worker streams), --init-seed N (layout initialisation)\n";
        return a.has("help") ? 0 : 1;
    }
    if (!a.has("idx")) { std::cerr << "[odgi::layout] error: Please specify an input file from where to load the graph via -i=[FILE], --idx=[FILE]." << std::endl; return 1; }
    // layout_main.cpp:110-114
    if (!a.has("tsv") && !a.has("out")) { std::cerr << "[odgi::layout] error: Please specify an output file to where to store the layout via -o/--out=[FILE], -T/--tsv=[FILE]" << std::endl; return 1; }
    if (int rc = need_gpu(a, "layout")) return rc;
    if (a.has("device-ingest")) return layout_device_ingest(a);
    pgsgd::FlatGraph fg;
    try { fg = pgsgd::read_gfa_flat(a.str("idx")); } catch (const std::exception& e) { std::cerr << e.what() << std::endl; return 1; }
    pgsgd_config c;
    if (a.has("path-sgd-use-paths")) {   // layout_main.cpp:230-263: the default parameters come from these paths only
        std::vector<uint64_t> use;
        if (!read_path_list(a.str("path-sgd-use-paths"), fg, false, true, use)) return 1;
        const PathStats ps = path_stats(fg, &use);
        common_config(a, fg, false, c, &ps);
    } else {
        common_config(a, fg, false, c);
    }
    const uint64_t N = fg.node_len.size();
    std::vector<double> X, Y;
    if (!init_layout(fg, a.str("layout-initialization", "d")[0], a.has("init-seed"), a.u64("init-seed", 0), X, Y)) return 1;
    pgsgd_stats st;
    const pgsgd_graph_view v = fg.view();
    if (!a.has("path-sgd-snapshot")) {
        if (pgsgd_layout_2d(&v, &c, X.data(), Y.data(), &st) != PGSGD_OK) { std::cerr << "[odgi::layout] error: " << pgsgd_last_error() << std::endl; return 1; }
    } else {
        // -u PREFIX: after every iteration but the last, the current coordinates as a .lay in PREFIX<iteration>
        // (snapshot_lambda, path_sgd_layout.cpp:379-409; the reference GPU path ignores the flag).  The graph stays
        // resident in HBM; only the coordinates come back per snapshot.
        pgsgd_engine* e = nullptr;
        int rc = pgsgd_engine_create(&v, 0, &e);
        if (!rc) rc = pgsgd_engine_set_coords_2d(e, X.data(), Y.data());
        std::memset(&st, 0, sizeof(st));
        for (uint64_t it = 0; !rc && it < c.iter_max; ++it) {
            pgsgd_stats one;
            rc = pgsgd_engine_run_range(e, &c, 2, it, it + 1, &one);
            if (rc) break;
            st.term_updates += one.term_updates;
            st.seconds_iterations += one.seconds_iterations;
            st.seconds_upload = one.seconds_upload;
            if (one.iterations_run == 0) break;  // nothing to do (no path with more than one step)
            rc = pgsgd_engine_get_coords_2d(e, X.data(), Y.data());
            if (!rc && it + 1 < c.iter_max) {
                if (a.has("progress")) std::cerr << "[odgi::path_linear_sgd_layout] snapshot thread: Taking snapshot!" << std::endl;
                std::ofstream f(a.str("path-sgd-snapshot") + std::to_string(it + 1), std::ios::binary);
                pgsgd::lay::write_lay(f, X, Y);
            }
            if (!rc && c.delta > 0 && one.last_delta_max <= c.delta) break;  // early stop, as the checker thread decides it
        }
        if (e) pgsgd_engine_destroy(e);
        if (rc) { std::cerr << "[odgi::layout] error: " << pgsgd_last_error() << std::endl; return 1; }
    }
    if (a.has("progress"))
        std::cerr << "[odgi::path_linear_sgd_layout] 2D path-guided SGD: " << st.term_updates << " term updates in " << st.seconds_iterations
                  << " s on the GPU (" << st.term_updates / st.seconds_iterations / 1e6 << " M updates/s), upload " << st.seconds_upload << " s" << std::endl;
    // stack the weakly connected components vertically with a 1000-unit border (layout_main.cpp:402-435)
    const std::vector<uint32_t> comp = components_of(fg, a.str("idx"));
    const uint32_t n_comp = comp.empty() ? 0 : *std::max_element(comp.begin(), comp.end()) + 1;
    const double border = 1000.0, inf = std::numeric_limits<double>::max();
    std::vector<double> min_x(n_comp, inf), min_y(n_comp, inf), max_y(n_comp, std::numeric_limits<double>::lowest());
    for (uint64_t r = 0; r < N; ++r)
        for (uint64_t j = 2 * r; j <= 2 * r + 1; ++j) {
            min_x[comp[r]] = std::min(min_x[comp[r]], X[j]); min_y[comp[r]] = std::min(min_y[comp[r]], Y[j]); max_y[comp[r]] = std::max(max_y[comp[r]], Y[j]);
        }
    std::vector<double> x_off(n_comp), y_off(n_comp);
    double curr_y_offset = border;
    for (uint32_t k = 0; k < n_comp; ++k) {
        x_off[k] = min_x[k] - border;
        y_off[k] = curr_y_offset - min_y[k];
        curr_y_offset += (max_y[k] - min_y[k]) + border;
    }
    for (uint64_t r = 0; r < N; ++r)
        for (uint64_t j = 2 * r; j <= 2 * r + 1; ++j) { X[j] -= x_off[comp[r]]; Y[j] += y_off[comp[r]]; }
    // layout::to_tsv (src/algorithms/layout.cpp:10-34): idx X Y component, two rows per node, grouped by component
    if (a.has("tsv")) {
        std::ofstream fout;
        std::ostream* out = &std::cout;
        if (a.str("tsv") != "-") { fout.open(a.str("tsv")); out = &fout; }
        *out << std::setprecision(std::numeric_limits<double>::digits10 + 1);
        *out << "idx\tX\tY\tcomponent" << std::endl;
        for (uint32_t k = 0; k < n_comp; ++k)
            for (uint64_t r = 0; r < N; ++r) {
                if (comp[r] != k) continue;
                *out << 2 * r << "\t" << X[2 * r] << "\t" << Y[2 * r] << "\t" << k << '\n';
                *out << 2 * r + 1 << "\t" << X[2 * r + 1] << "\t" << Y[2 * r + 1] << "\t" << k << '\n';
            }
    }
    // the binary container `odgi draw` reads (layout_main.cpp:451-463, layout.cpp:43-61)
    if (a.has("out")) {
        try {
            if (a.str("out") == "-") pgsgd::lay::write_lay(std::cout, X, Y);
            else { std::ofstream f(a.str("out"), std::ios::binary); pgsgd::lay::write_lay(f, X, Y); }
        } catch (const std::exception& e) { std::cerr << "[odgi::layout] error: " << e.what() << std::endl; return 1; }
    }
    return 0;
}

// a line-separated list of path names -> path ranks (file order)
bool read_path_list(const std::string& file, const pgsgd::FlatGraph& fg, bool reject_duplicates, bool unknown_is_error, std::vector<uint64_t>& out) {
    std::ifstream in(file);
    if (!in) { std::cerr << "[odgi::sort] error: cannot open " << file << std::endl; return false; }
    std::map<std::string, uint64_t> rank;
    for (uint64_t p = 0; p < fg.path_names.size(); ++p) rank.emplace(fg.path_names[p], p);
    std::vector<bool> seen(fg.path_names.size(), false);
    std::string line;
    uint64_t in_file = 0;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        ++in_file;
        auto it = rank.find(line);
        if (it == rank.end()) {
            if (unknown_is_error) std::cerr << "[odgi::sort] error: path '" << line << "' as was given by -f=[FILE], --path-sgd-use-paths=[FILE] is not present in the graph." << std::endl;
            continue;   // sort_main.cpp:241,361-367: unknown names are skipped (with a message for -f)
        }
        if (seen[it->second] && reject_duplicates) { std::cerr << "[odgi::sort] error: in the path list there are duplicated path names." << std::endl; return false; }
        seen[it->second] = true;
        out.push_back(it->second);
    }
    if (reject_duplicates) std::cerr << "[odgi::sort] found " << out.size() << "/" << in_file << " paths to consider." << std::endl;
    if (out.empty()) { std::cerr << "[odgi::sort] error: no path to consider." << std::endl; return false; }
    return true;
}

// `odgi sort -H` (sort_graph_by_target_paths, sort_main.cpp:266-311): the nodes of the target paths come first, in the order
// the paths first visit them (paths in file order), all other nodes follow in id order; the first `ref_nodes` ranks of the
// reordered graph are the ones PG-SGD keeps fixed.  Returns old rank of every new rank.
std::vector<uint32_t> reorder_by_target_paths(pgsgd::FlatGraph& fg, const std::vector<uint64_t>& targets, uint64_t& ref_nodes) {
    const uint64_t N = fg.node_len.size();
    std::vector<uint32_t> old_of_new;
    old_of_new.reserve(N);
    std::vector<bool> is_ref(N, false);
    for (uint64_t p : targets)
        for (uint64_t s = fg.path_first_step[p]; s < fg.path_first_step[p + 1]; ++s) {
            const uint32_t r = fg.step_node[s];
            if (!is_ref[r]) { is_ref[r] = true; old_of_new.push_back(r); }
        }
    ref_nodes = old_of_new.size();
    for (uint32_t r = 0; r < N; ++r) if (!is_ref[r]) old_of_new.push_back(r);
    std::vector<uint32_t> new_of_old(N);
    for (uint32_t n = 0; n < N; ++n) new_of_old[old_of_new[n]] = n;
    std::vector<uint32_t> len(N);
    for (uint32_t n = 0; n < N; ++n) len[n] = fg.node_len[old_of_new[n]];
    fg.node_len.swap(len);
    for (auto& r : fg.step_node) r = new_of_old[r];   // positions along the paths do not change
    return old_of_new;
}

int main_sort(int argc, char** argv) {
    Args a;
    if (!parse(argc, argv, SORT_FLAGS, a, "sort") || a.has("help") || argc == 2) {
        std::cout << "pgsgd sort -i g.gfa -o order.txt [-e sorted.lay] [-f use_paths.txt] [-H target_paths.txt] -Y --gpu [-x N] [-G N|-U N] [-j N] [-g N] [-v N] [-a N] [-K N] [-F N] [-k N] [-I N] [-l N] [-t N] [-P]\n"
                     "  the `odgi sort -Y` PG-SGD flags with the same defaults; writes the node order (one node id per line) that\n"
                     "  path_linear_sgd_order derives (sorted by position, then handle); odgi applies it with apply_ordering.\n";
        return a.has("help") ? 0 : 1;
    }
    if (!a.has("idx")) { std::cerr << "[odgi::sort] error: please specify an input file from where to load the graph via -i=[FILE], --idx=[FILE]." << std::endl; return 1; }
    if (!a.has("out")) { std::cerr << "[odgi::sort] error: please specify an output file to where to store the node order via -o=[FILE], --out=[FILE]." << std::endl; return 1; }
    if (!a.has("path-sgd")) { std::cerr << "[odgi::sort] error: only the path-guided SGD sort (-Y, --path-sgd) is provided by this build." << std::endl; return 1; }
    if (!a.has("prepared-out")) { if (int rc = need_gpu(a, "sort")) return rc; }
    pgsgd::FlatGraph fg;
    try { fg = pgsgd::read_gfa_flat(a.str("idx")); } catch (const std::exception& e) { std::cerr << e.what() << std::endl; return 1; }
    const uint64_t N = fg.node_len.size();
    // -H: target (reference) paths stay put
    std::vector<uint32_t> old_of_new;
    std::vector<uint8_t> frozen;
    if (a.has("target-paths")) {
        std::vector<uint64_t> targets;
        if (!read_path_list(a.str("target-paths"), fg, true, false, targets)) return 1;
        uint64_t ref_nodes = 0;
        old_of_new = reorder_by_target_paths(fg, targets, ref_nodes);
        frozen.assign(N, 0);
        std::fill_n(frozen.begin(), ref_nodes, (uint8_t) 1);
    }
    // -f: the paths the default parameters are derived from
    pgsgd_config c;
    if (a.has("path-sgd-use-paths")) {
        std::vector<uint64_t> use;
        if (!read_path_list(a.str("path-sgd-use-paths"), fg, false, true, use)) return 1;
        const PathStats ps = path_stats(fg, &use);
        common_config(a, fg, true, c, &ps);
    } else {
        common_config(a, fg, true, c);
    }
    if (a.has("prepared-out")) {   // what goes to the GPU, for inspection and tests: no device needed
        pgsgd::ArrayWriter w(a.str("prepared-out"));
        w.add("node_len", fg.node_len);
        w.add("path_first_step", fg.path_first_step);
        w.add("step_node", fg.step_node);
        w.add("step_rev", fg.step_rev);
        w.add("step_pos", fg.step_pos);
        w.add("frozen", frozen);
        w.add("old_of_new", old_of_new);
        w.add_scalar<uint64_t>("min_term_updates", c.min_term_updates);
        w.add_scalar<uint64_t>("space", c.space);
        w.add_scalar<uint64_t>("space_max", c.space_max);
        w.add_scalar<uint64_t>("space_quantization_step", c.space_quantization_step);
        w.add_scalar<double>("eta_max", c.eta_max);
        w.close();
        if (!a.has("gpu")) return 0;
    }
    std::vector<double> X(N);
    pgsgd_stats st;
    const pgsgd_graph_view v = fg.view();
    if (pgsgd_sort_1d(&v, &c, frozen.empty() ? nullptr : frozen.data(), 0, X.data(), &st) != PGSGD_OK) { std::cerr << "[odgi::sort] error: " << pgsgd_last_error() << std::endl; return 1; }
    if (a.has("progress"))
        std::cerr << "[odgi::path_linear_sgd] 1D path-guided SGD: " << st.term_updates << " term updates in " << st.seconds_iterations << " s on the GPU" << std::endl;
    // path_linear_sgd_order (path_sgd.cpp:638-683): sort by (weak component, pos, handle), components numbered by their
    // average node id (:557-573).  The reference clear()s its component map before reading it (:588), but clear() leaves
    // the storage in place and the unchecked operator[] reads still return the ids: the running binary does order by
    // component (checked against the reference on a graph with interleaved components, tests/golden/order_multi3.json).
    std::vector<uint32_t> comp_key(N, 0);
    {
        std::vector<uint32_t> comp = components_of(fg, a.str("idx"));   // ranks of THIS run's graph (reordered with -H)
        if (!old_of_new.empty()) {   // components_of numbers by the input ids the L lines carry
            std::vector<uint32_t> by_new(N);
            for (uint64_t r = 0; r < N; ++r) by_new[r] = comp[old_of_new[r]];
            comp.swap(by_new);
        }
        const uint32_t n_comp = comp.empty() ? 0 : *std::max_element(comp.begin(), comp.end()) + 1;
        std::vector<double> id_sum(n_comp, 0.0);
        std::vector<uint64_t> cnt(n_comp, 0);
        for (uint64_t r = 0; r < N; ++r) { id_sum[comp[r]] += (double) (r + 1); ++cnt[comp[r]]; }
        std::vector<std::pair<double, uint32_t>> by_avg;
        for (uint32_t k = 0; k < n_comp; ++k) by_avg.emplace_back(id_sum[k] / (double) cnt[k], k);
        std::sort(by_avg.begin(), by_avg.end());
        std::vector<uint32_t> key_of(n_comp);
        for (uint32_t i = 0; i < n_comp; ++i) key_of[by_avg[i].second] = i;
        for (uint64_t r = 0; r < N; ++r) comp_key[r] = key_of[comp[r]];
    }
    std::vector<uint64_t> order(N);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](uint64_t i, uint64_t j) {
        if (comp_key[i] != comp_key[j]) return comp_key[i] < comp_key[j];
        return X[i] < X[j] || (X[i] == X[j] && i < j);
    });
    // node ids of the INPUT graph, in sorted order (with -H the run worked on the reordered graph)
    auto input_id = [&](uint64_t r) { return (uint64_t) (old_of_new.empty() ? r : old_of_new[r]) + 1; };
    std::ofstream f(a.str("out"));
    for (uint64_t r : order) f << input_id(r) << '\n';
    if (a.has("layout-out")) {  // -L in odgi sort: the 1D layout (start, start + node length) per sorted node
        std::ofstream l(a.str("layout-out"));
        l << std::setprecision(std::numeric_limits<double>::digits10 + 1) << "node\tstart\tend\n";
        for (uint64_t r : order) l << input_id(r) << "\t" << X[r] << "\t" << X[r] + (double) fg.node_len[r] << '\n';
    }
    if (a.has("path-sgd-layout")) {  // -e in odgi sort: the same as a .lay, X = (start, start + length) per sorted node, Y = 0 (path_sgd.cpp:659-677)
        std::vector<double> sorted_layout(2 * N), dummy(2 * N, 0.0);
        for (uint64_t i = 0; i < N; ++i) {
            sorted_layout[2 * i] = X[order[i]];
            sorted_layout[2 * i + 1] = X[order[i]] + (double) fg.node_len[order[i]];
        }
        try { std::ofstream l(a.str("path-sgd-layout"), std::ios::binary); pgsgd::lay::write_lay(l, sorted_layout, dummy); }
        catch (const std::exception& e) { std::cerr << "[odgi::sort] error: " << e.what() << std::endl; return 1; }
    }
    return 0;
}

// pgsgd scan -i g.gfa : what the line scan of the device ingest finds (counts and checksums; PGSGD_SCAN_THREADS = host threads)
int main_scan(int argc, char** argv) {
    std::string in;
    for (int i = 2; i + 1 < argc; i += 2) if (!std::strcmp(argv[i], "-i")) in = argv[i + 1];
    if (in.empty()) { std::cerr << "usage: pgsgd scan -i g.gfa" << std::endl; return 1; }
    pgsgd::GfaIndex ix;
    const auto t0 = std::chrono::steady_clock::now();
    try { pgsgd::scan_gfa(in, ix); } catch (const std::exception& e) { std::cerr << e.what() << std::endl; return 1; }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t h = 1469598103934665603ull, bp = 0;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    for (uint32_t l : ix.node_len) { mix(l); bp += l; }
    for (size_t p = 0; p < ix.field_begin.size(); ++p) { mix(ix.field_begin[p]); mix(ix.field_end[p]); for (char c : ix.path_names[p]) mix((unsigned char) c); }
    for (auto& l : ix.links) { mix(l.first); mix(l.second); }
    std::cout << "{\"nodes\": " << ix.node_len.size() << ", \"paths\": " << ix.field_begin.size() << ", \"links\": " << ix.links.size() << ", \"bp\": " << bp
              << ", \"fnv\": " << h << ", \"seconds\": " << dt << "}" << std::endl;
    return 0;
}

// pgsgd flatten -i g.gfa -o g.arr : the flattened graph as a PGSGDARR container (what bench.py / the tests load)
int main_flatten(int argc, char** argv) {
    std::string in, out;
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "-i")) in = argv[i + 1];
        else if (!std::strcmp(argv[i], "-o")) out = argv[i + 1];
    }
    if (in.empty() || out.empty()) { std::cerr << "usage: pgsgd flatten -i g.gfa -o g.arr" << std::endl; return 1; }
    pgsgd::FlatGraph fg;
    try { fg = pgsgd::read_gfa_flat(in); } catch (const std::exception& e) { std::cerr << e.what() << std::endl; return 1; }
    std::vector<uint8_t> names;
    for (auto& n : fg.path_names) { names.insert(names.end(), n.begin(), n.end()); names.push_back((uint8_t) '\n'); }
    pgsgd::ArrayWriter w(out);
    w.add("node_len", fg.node_len);
    w.add("path_first_step", fg.path_first_step);
    w.add("step_node", fg.step_node);
    w.add("step_rev", fg.step_rev);
    w.add("step_pos", fg.step_pos);
    w.add("path_names", names);
    w.close();
    std::cout << "{\"nodes\": " << fg.node_len.size() << ", \"paths\": " << fg.path_names.size() << ", \"steps\": " << fg.steps() << "}" << std::endl;
    return 0;
}

// pgsgd init -i g.gfa -N d|r|u|g|h [--init-seed S] -a out.arr : the coordinates `pgsgd layout` would start from
int main_init(int argc, char** argv) {
    std::string in, arr, mode = "d", seed;
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "-i")) in = argv[i + 1];
        else if (!std::strcmp(argv[i], "-a")) arr = argv[i + 1];
        else if (!std::strcmp(argv[i], "-N")) mode = argv[i + 1];
        else if (!std::strcmp(argv[i], "--init-seed")) seed = argv[i + 1];
    }
    if (in.empty() || arr.empty() || mode.empty()) { std::cerr << "usage: pgsgd init -i g.gfa -N d|r|u|g|h [--init-seed S] -a out.arr" << std::endl; return 1; }
    pgsgd::FlatGraph fg;
    try { fg = pgsgd::read_gfa_flat(in); } catch (const std::exception& e) { std::cerr << e.what() << std::endl; return 1; }
    std::vector<double> X, Y;
    if (!init_layout(fg, mode[0], !seed.empty(), seed.empty() ? 0 : std::stoull(seed), X, Y)) return 1;
    pgsgd::ArrayWriter w(arr);
    w.add("X", X);
    w.add("Y", Y);
    w.close();
    return 0;
}

// pgsgd lay -i in.lay -T out.tsv      a .lay as `idx X Y` rows (Layout::to_tsv, layout.cpp:68-74)
// pgsgd lay -i in.lay -a out.arr      ... or as a PGSGDARR container with X and Y
// pgsgd lay -c xy.arr -o out.lay      a .lay from X and Y arrays (what `pgsgd layout -o` writes after the run)
int main_lay(int argc, char** argv) {
    std::string in, coords, tsv, arr, out;
    for (int i = 2; i + 1 < argc; i += 2) {
        if (!std::strcmp(argv[i], "-i")) in = argv[i + 1];
        else if (!std::strcmp(argv[i], "-c")) coords = argv[i + 1];
        else if (!std::strcmp(argv[i], "-T")) tsv = argv[i + 1];
        else if (!std::strcmp(argv[i], "-a")) arr = argv[i + 1];
        else if (!std::strcmp(argv[i], "-o")) out = argv[i + 1];
    }
    try {
        std::vector<double> X, Y;
        if (!in.empty()) {
            std::ifstream f(in, std::ios::binary);
            if (!f) throw std::runtime_error("cannot open " + in);
            pgsgd::lay::read_lay(f, X, Y);
        } else if (!coords.empty()) {
            auto arrs = pgsgd::read_arrays(coords);
            X = arrs.at("X").vec<double>();
            Y = arrs.at("Y").vec<double>();
        } else {
            std::cerr << "usage: pgsgd lay (-i in.lay | -c xy.arr) [-T out.tsv] [-a out.arr] [-o out.lay]" << std::endl;
            return 1;
        }
        if (!tsv.empty()) {
            std::ofstream f(tsv);
            f << std::setprecision(std::numeric_limits<double>::digits10 + 1);
            f << "idx\tX\tY" << std::endl;
            for (uint64_t i = 0; i < X.size(); ++i) f << i << "\t" << X[i] << "\t" << Y[i] << '\n';
        }
        if (!arr.empty()) { pgsgd::ArrayWriter w(arr); w.add("X", X); w.add("Y", Y); w.close(); }
        if (!out.empty()) { std::ofstream f(out, std::ios::binary); pgsgd::lay::write_lay(f, X, Y); }
    } catch (const std::exception& e) { std::cerr << "[pgsgd::lay] error: " << e.what() << std::endl; return 1; }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: pgsgd layout|sort [flags] --gpu   (the PG-SGD subset of `odgi layout` / `odgi sort -Y`)" << std::endl; return 1; }
    const std::string sub = argv[1];
    if (sub == "layout") return main_layout(argc, argv);
    if (sub == "sort") return main_sort(argc, argv);
    if (sub == "flatten") return main_flatten(argc, argv);
    if (sub == "lay") return main_lay(argc, argv);
    if (sub == "scan") return main_scan(argc, argv);
    if (sub == "init") return main_init(argc, argv);
    std::cerr << "unknown subcommand " << sub << " (layout, sort, flatten, lay, init)" << std::endl;
    return 1;
}
