// gfa_lite.hpp — the minimum of GFA1 the PG-SGD path needs, read straight into the flattened form: S lines give node
// lengths, P lines give the path steps; L lines and everything else are skipped (the reference's ingest also reads only
// S, L and P, src/gfa_to_handle.cpp:62-64, and PG-SGD never looks at edges).  Node ids must be exactly 1..N.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>

#include "pgsgd_flatten.hpp"

namespace pgsgd {

inline FlatGraph read_gfa_flat(const std::string& path) {
    // P lines of a chromosome-scale graph are tens of MB each: read through a large buffer (libstdc++ takes it only before open)
    std::vector<char> iobuf(8u << 20);
    std::ifstream in;
    in.rdbuf()->pubsetbuf(iobuf.data(), (std::streamsize) iobuf.size());
    in.open(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    FlatGraph fg;
    std::string line;
    // pass 1: segments
    std::vector<std::pair<uint64_t, uint32_t>> segs;
    uint64_t max_id = 0;
    bool numeric = true;
    uint64_t steps_total = 0;   // counted here so that pass 2 fills exactly-sized arrays (no regrowth of multi-GB vectors)
    while (std::getline(in, line)) {
        if (line.size() >= 2 && line[0] == 'P' && line[1] == '\t') {
            const size_t b = line.find('\t', 2);
            if (b == std::string::npos) continue;   // reported in pass 2
            size_t c = line.find('\t', b + 1);
            if (c == std::string::npos) c = line.size();
            if (c > b + 1) {
                ++steps_total;
                for (const char* q = line.data() + b + 1, * const e = line.data() + c; (q = (const char*) std::memchr(q, ',', (size_t) (e - q))) != nullptr; ++q) ++steps_total;
            }
            continue;
        }
        if (line.size() < 2 || line[0] != 'S' || line[1] != '\t') continue;
        const size_t a = 2, b = line.find('\t', a);
        if (b == std::string::npos) throw std::runtime_error("malformed S line");
        size_t c = line.find('\t', b + 1);
        if (c == std::string::npos) c = line.size();
        const std::string name = line.substr(a, b - a);
        char* endp = nullptr;
        const uint64_t id = std::strtoull(name.c_str(), &endp, 10);
        if (*endp != '\0' || id == 0) { numeric = false; break; }
        uint32_t len = (uint32_t) (c - b - 1);
        if (len == 1 && line[b + 1] == '*') {   // sequence omitted: LN:i: tag
            len = 0;
            const size_t t = line.find("LN:i:", c);
            if (t != std::string::npos) len = (uint32_t) std::strtoull(line.c_str() + t + 5, nullptr, 10);
        }
        segs.emplace_back(id, len);
        if (id > max_id) max_id = id;
    }
    if (!numeric || max_id != segs.size()) {
        throw std::runtime_error("[odgi::layout] error: the graph is not optimized. Please run 'odgi sort' using -O, --optimize.");
    }
    fg.node_len.assign(max_id, 0);
    for (auto& s : segs) fg.node_len[s.first - 1] = s.second;
    fg.step_node.reserve(steps_total);
    fg.step_rev.reserve(steps_total);
    fg.step_pos.reserve(steps_total);
    // pass 2: paths, in file order (path handles are assigned in file order, gfa_to_handle.cpp:193-199)
    in.clear();
    in.seekg(0);
    while (std::getline(in, line)) {
        if (line.size() < 2 || line[0] != 'P' || line[1] != '\t') continue;
        const size_t a = 2, b = line.find('\t', a);
        if (b == std::string::npos) throw std::runtime_error("malformed P line");
        size_t c = line.find('\t', b + 1);
        if (c == std::string::npos) c = line.size();
        fg.begin_path(line.substr(a, b - a));
        const char* p = line.c_str() + b + 1;
        const char* end = line.c_str() + c;
        while (p < end) {
            uint64_t id = 0;
            const char* q = p;
            while (q < end && (unsigned) (*q - '0') <= 9u) id = id * 10 + (uint64_t) (*q++ - '0');
            if (q == p || q >= end || (*q != '+' && *q != '-') || id == 0 || id > max_id) throw std::runtime_error("bad step in path " + fg.path_names.back());
            fg.add_step((uint32_t) (id - 1), *q == '-');
            p = q + 1;
            if (p < end && *p == ',') ++p;
        }
        fg.end_path();
    }
    return fg;
}


// ---- the host half of the DEVICE ingest (pgsgd_engine_create_from_gfa_paths) ------------------------------------------
// One pass over the mapped file with memchr: node lengths from the S lines, the link endpoints from the L lines (weak
// components for the layout stacking) and, for every P line, only the BYTE RANGE of its step list — the steps themselves
// (the bulk of a pangenome GFA) are parsed on the GPU.
struct GfaIndex {
    std::vector<uint32_t> node_len;
    std::vector<std::string> path_names;
    std::vector<uint64_t> field_begin, field_end;          // step list of path p: text[field_begin[p], field_end[p])
    std::vector<std::pair<uint32_t, uint32_t>> links;      // node ranks of every L line
    const char* text = nullptr;
    size_t bytes = 0;
    int fd = -1;
    GfaIndex() = default;
    GfaIndex(const GfaIndex&) = delete;
    GfaIndex& operator=(const GfaIndex&) = delete;
    ~GfaIndex();
};

}  // namespace pgsgd
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
namespace pgsgd {

inline GfaIndex::~GfaIndex() {
    if (text) munmap((void*) text, bytes);
    if (fd >= 0) close(fd);
}

inline void scan_gfa(const std::string& path, GfaIndex& ix) {
    ix.fd = open(path.c_str(), O_RDONLY);
    if (ix.fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat sb;
    if (fstat(ix.fd, &sb) != 0 || sb.st_size == 0) throw std::runtime_error("cannot stat " + path);
    ix.bytes = (size_t) sb.st_size;
    void* m = mmap(nullptr, ix.bytes, PROT_READ, MAP_PRIVATE, ix.fd, 0);
    if (m == MAP_FAILED) throw std::runtime_error("cannot map " + path);
    ix.text = (const char*) m;
    auto field_end_of = [](const char* b, const char* line_end) {   // next tab, or the end of the line (without a CR)
        const char* t = (const char*) std::memchr(b, '\t', (size_t) (line_end - b));
        const char* e = t ? t : line_end;
        if (!t && e > b && e[-1] == '\r') --e;
        return e;
    };
    auto parse_id = [](const char* b, const char* e, uint64_t& id) {
        id = 0;
        if (b == e) return false;
        for (const char* q = b; q < e; ++q) {
            if ((unsigned) (*q - '0') > 9u) return false;
            id = id * 10 + (uint64_t) (*q - '0');
        }
        return id != 0;
    };
    // The file is cut into pieces at line starts and the pieces are scanned by host threads (a 4 GB GFA is two memchr passes
    // over 4 GB: ~1 s on one core); the pieces' findings are concatenated in file order.
    struct Piece {
        std::vector<std::pair<uint64_t, uint32_t>> segs;
        uint64_t max_id = 0;
        std::vector<std::string> path_names;
        std::vector<uint64_t> field_begin, field_end;
        std::vector<std::pair<uint32_t, uint32_t>> links;
        bool not_optimized = false;
    };
    const char* const text = ix.text;
    const char* const end = ix.text + ix.bytes;
    unsigned n_pieces = std::thread::hardware_concurrency();
    if (n_pieces == 0) n_pieces = 1;
    if (n_pieces > 32) n_pieces = 32;
    if (ix.bytes < (64u << 20)) n_pieces = 1;
    if (const char* sv = std::getenv("PGSGD_SCAN_THREADS")) { const int v = std::atoi(sv); if (v >= 1 && v <= 256) n_pieces = (unsigned) v; }
    std::vector<const char*> cut(n_pieces + 1, end);
    cut[0] = text;
    for (unsigned k = 1; k < n_pieces; ++k) {
        const char* q = text + (ix.bytes / n_pieces) * k;
        if (q < cut[k - 1]) q = cut[k - 1];
        const char* nl = q < end ? (const char*) std::memchr(q, '\n', (size_t) (end - q)) : nullptr;
        cut[k] = nl ? nl + 1 : end;
    }
    std::vector<Piece> pieces(n_pieces);
    auto scan_piece = [&](unsigned k) {
        Piece& pc = pieces[k];
        const char* p = cut[k];
        const char* const pend = cut[k + 1];
        while (p < pend) {
            const char* nl = (const char*) std::memchr(p, '\n', (size_t) (end - p));
            const char* le = nl ? nl : end;
            if (le - p >= 2 && p[1] == '\t') {
                if (p[0] == 'S') {
                    const char* b = p + 2;
                    const char* e = field_end_of(b, le);
                    uint64_t id;
                    if (!parse_id(b, e, id)) { pc.not_optimized = true; return; }
                    const char* sb2 = e < le ? e + 1 : le;
                    const char* se = field_end_of(sb2, le);
                    uint32_t len = (uint32_t) (se - sb2);
                    if (len == 1 && *sb2 == '*') {
                        len = 0;
                        const char* tag = se;
                        while (tag + 5 < le) { if (!std::memcmp(tag, "LN:i:", 5)) { len = (uint32_t) std::strtoull(tag + 5, nullptr, 10); break; } ++tag; }
                    }
                    pc.segs.emplace_back(id, len);
                    if (id > pc.max_id) pc.max_id = id;
                } else if (p[0] == 'P') {
                    const char* b = p + 2;
                    const char* e = field_end_of(b, le);
                    pc.path_names.emplace_back(b, (size_t) (e - b));
                    const char* fb = e < le ? e + 1 : le;
                    const char* fe = field_end_of(fb, le);
                    pc.field_begin.push_back((uint64_t) (fb - text));
                    pc.field_end.push_back((uint64_t) (fe - text));
                } else if (p[0] == 'L') {
                    const char* b = p + 2;
                    const char* e = field_end_of(b, le);
                    uint64_t ia = 0, ib = 0;
                    const bool oka = parse_id(b, e, ia);
                    const char* o = e < le ? e + 1 : le;              // orientation field
                    const char* oe = field_end_of(o, le);
                    const char* b2 = oe < le ? oe + 1 : le;
                    const char* e2 = field_end_of(b2, le);
                    if (oka && parse_id(b2, e2, ib)) pc.links.emplace_back((uint32_t) (ia - 1), (uint32_t) (ib - 1));
                }
            }
            p = nl ? nl + 1 : end;
        }
    };
    if (n_pieces == 1) {
        scan_piece(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned k = 0; k < n_pieces; ++k) th.emplace_back(scan_piece, k);
        for (auto& t : th) t.join();
    }
    std::vector<std::pair<uint64_t, uint32_t>> segs;
    uint64_t max_id = 0;
    for (Piece& pc : pieces) {
        if (pc.not_optimized) throw std::runtime_error("[odgi::layout] error: the graph is not optimized. Please run 'odgi sort' using -O, --optimize.");
        segs.insert(segs.end(), pc.segs.begin(), pc.segs.end());
        if (pc.max_id > max_id) max_id = pc.max_id;
        for (auto& nm : pc.path_names) ix.path_names.emplace_back(std::move(nm));
        ix.field_begin.insert(ix.field_begin.end(), pc.field_begin.begin(), pc.field_begin.end());
        ix.field_end.insert(ix.field_end.end(), pc.field_end.begin(), pc.field_end.end());
        ix.links.insert(ix.links.end(), pc.links.begin(), pc.links.end());
    }
    if (max_id != segs.size()) throw std::runtime_error("[odgi::layout] error: the graph is not optimized. Please run 'odgi sort' using -O, --optimize.");
    ix.node_len.assign(max_id, 0);
    for (auto& sg : segs) ix.node_len[sg.first - 1] = sg.second;
}

}  // namespace pgsgd
