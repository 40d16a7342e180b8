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

}  // namespace pgsgd
