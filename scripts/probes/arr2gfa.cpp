// arr2gfa.cpp — measurement plumbing: a flattened synthetic graph (PGSGDARR container written by odgi_b200.graphio
// save_graph_arrays) as GFA1 text, fast enough for the c4 graph (4.2e8 steps -> 3.6 GB of text in about a minute), so that the
// unmodified reference and `pgsgd layout --device-ingest` can be timed on the same file.  Same output as
// odgi_b200/synth.py write_gfa: placeholder sequences (only lengths matter to PG-SGD), the adjacencies the paths use as L lines.
// Build: g++ -O2 -std=c++17 -fopenmp -I odgi_b200/host -o scripts/probes/arr2gfa scripts/probes/arr2gfa.cpp
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <parallel/algorithm>
#include <string>
#include <vector>

#include "pgsgd_arrays.hpp"

static inline char* put_u(char* p, uint64_t v) {
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: arr2gfa in.arr out.gfa\n"); return 1; }
    auto a = pgsgd::read_arrays(argv[1]);
    const auto& nl = a.at("node_len");
    const auto& pf = a.at("path_first_step");
    const auto& sn = a.at("step_node");
    const uint64_t N = nl.count, P = pf.count - 1, S = sn.count;
    const uint32_t* node = sn.as<uint32_t>();
    const uint64_t* first = pf.as<uint64_t>();
    const uint32_t* len = nl.as<uint32_t>();
    const uint8_t* rev = a.count("step_rev") ? a.at("step_rev").as<uint8_t>() : nullptr;
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) { std::perror(argv[2]); return 1; }
    std::vector<char> buf(1 << 24);
    std::setvbuf(f, buf.data(), _IOFBF, buf.size());
    std::fputs("H\tVN:Z:1.0\n", f);
    std::string line;
    for (uint64_t i = 0; i < N; ++i) {
        line.assign("S\t");
        char num[24];
        line.append(num, put_u(num, i + 1));
        line.push_back('\t');
        line.append((size_t) len[i], 'A');
        line.push_back('\n');
        std::fwrite(line.data(), 1, line.size(), f);
    }
    // unique (from handle, to handle) pairs over consecutive steps of a path
    std::vector<uint64_t> pairs;
    pairs.reserve(S);
    for (uint64_t p = 0; p < P; ++p)
        for (uint64_t i = first[p]; i + 1 < first[p + 1]; ++i) {
            const uint64_t u = ((uint64_t) node[i] << 1) | (rev ? rev[i] : 0), v = ((uint64_t) node[i + 1] << 1) | (rev ? rev[i + 1] : 0);
            pairs.push_back((u << 32) | v);
        }
    __gnu_parallel::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    for (uint64_t e : pairs) {
        char l[80], *q = l;
        const uint64_t u = e >> 32, v = e & 0xFFFFFFFFull;
        *q++ = 'L'; *q++ = '\t'; q = put_u(q, (u >> 1) + 1); *q++ = '\t'; *q++ = (u & 1) ? '-' : '+'; *q++ = '\t';
        q = put_u(q, (v >> 1) + 1); *q++ = '\t'; *q++ = (v & 1) ? '-' : '+'; *q++ = '\t'; *q++ = '0'; *q++ = 'M'; *q++ = '\n';
        std::fwrite(l, 1, (size_t) (q - l), f);
    }
    std::vector<uint64_t>().swap(pairs);
    std::vector<char> steps;
    for (uint64_t p = 0; p < P; ++p) {
        const uint64_t b = first[p], e = first[p + 1];
        steps.resize((e - b) * 12 + 64);
        char* q = steps.data();
        q += std::snprintf(q, 48, "P\thap%llu\t", (unsigned long long) p);
        for (uint64_t i = b; i < e; ++i) {
            q = put_u(q, (uint64_t) node[i] + 1);
            *q++ = (rev && rev[i]) ? '-' : '+';
            if (i + 1 < e) *q++ = ',';
        }
        *q++ = '\t'; *q++ = '*'; *q++ = '\n';
        std::fwrite(steps.data(), 1, (size_t) (q - steps.data()), f);
    }
    std::fclose(f);
    std::printf("{\"nodes\": %llu, \"paths\": %llu, \"steps\": %llu}\n", (unsigned long long) N, (unsigned long long) P, (unsigned long long) S);
    return 0;
}
