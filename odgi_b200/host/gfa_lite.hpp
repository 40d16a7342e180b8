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
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    FlatGraph fg;
    std::string line;
    // pass 1: segments
    std::vector<std::pair<uint64_t, uint32_t>> segs;
    uint64_t max_id = 0;
    bool numeric = true;
    while (std::getline(in, line)) {
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
            char* q = nullptr;
            const uint64_t id = std::strtoull(p, &q, 10);
            if (q == p || q >= end + 1 || id == 0 || id > max_id) throw std::runtime_error("bad step in path " + fg.path_names.back());
            const bool rev = *q == '-';
            fg.add_step((uint32_t) (id - 1), rev);
            p = q + 1;
            if (p < end && *p == ',') ++p;
        }
        fg.end_path();
    }
    return fg;
}

}  // namespace pgsgd
