// pgsgd_gfa.cu — GFA `P`-line step lists parsed ON THE DEVICE.
//
// What it replaces: the per-step host work of getting a graph's paths into the flattened form — in the reference the walk
// of every path through graph_t (src/cuda/layout.cu:371-410, two node spinlocks and a delta decode per step,
// src/odgi.cpp:393-424), in this repo's standalone reader a strtoull per step (odgi_b200/host/gfa_lite.hpp: ~16 s for the
// 4.2e8 steps of c4).  Here the host only finds the line boundaries (memchr) and reads the S lines; the step lists — the
// bulk of the file, "123+,124-,..." — are uploaded as raw text and every step parses itself:
//   1. count:  steps per 4 KB chunk of text (a step ends at its orientation character, '+' or '-'),
//   2. scan:   exclusive sum over the chunks (CUB),
//   3. parse:  a block-wide scan gives every orientation character its step index; its thread reads the decimal node id
//              backwards from it and writes step_node (id - 1) / step_rev.
// The first step of every path is the number of orientation characters before the path's field.
#include "pgsgd_kernels.cuh"

#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

namespace pgsgd {

namespace {

constexpr int GFA_BLOCK = 256;
constexpr int GFA_PER_THREAD = 16;
constexpr int GFA_CHUNK = GFA_BLOCK * GFA_PER_THREAD;   // 4096 bytes of text per CTA

__device__ __forceinline__ bool is_end(char c) { return c == '+' || c == '-'; }

__global__ void __launch_bounds__(GFA_BLOCK) gfa_count_kernel(const char* text, uint64_t n, uint64_t* chunk_count) {
    using Reduce = cub::BlockReduce<uint32_t, GFA_BLOCK>;
    __shared__ typename Reduce::TempStorage tmp;
    const uint64_t base = (uint64_t) blockIdx.x * GFA_CHUNK + (uint64_t) threadIdx.x * GFA_PER_THREAD;
    uint32_t c = 0;
    if (base + GFA_PER_THREAD <= n) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + base);   // the text buffer is 16-byte aligned, base a multiple of 16
        const unsigned char* b = reinterpret_cast<const unsigned char*>(&v);
#pragma unroll
        for (int k = 0; k < GFA_PER_THREAD; ++k) c += is_end((char) b[k]) ? 1u : 0u;
    } else {
        for (uint64_t i = base; i < n; ++i) c += is_end(text[i]) ? 1u : 0u;
    }
    const uint32_t total = Reduce(tmp).Sum(c);
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = total;
}

__global__ void __launch_bounds__(GFA_BLOCK) gfa_parse_kernel(const char* text, uint64_t n, const uint64_t* chunk_first, uint32_t n_nodes,
                                                               uint32_t* step_node, uint8_t* step_rev, int* bad) {
    using Scan = cub::BlockScan<uint32_t, GFA_BLOCK>;
    __shared__ typename Scan::TempStorage tmp;
    const uint64_t base = (uint64_t) blockIdx.x * GFA_CHUNK + (uint64_t) threadIdx.x * GFA_PER_THREAD;
    unsigned char b[GFA_PER_THREAD];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < GFA_PER_THREAD; ++k) {
        b[k] = base + k < n ? (unsigned char) text[base + k] : 0;
        c += is_end((char) b[k]) ? 1u : 0u;
    }
    uint32_t before;
    Scan(tmp).ExclusiveSum(c, before);
    uint64_t idx = chunk_first[blockIdx.x] + before;
#pragma unroll
    for (int k = 0; k < GFA_PER_THREAD; ++k) {
        if (!is_end((char) b[k])) continue;
        // the decimal node id ends right before the orientation character; it may begin in another thread's bytes
        uint64_t id = 0, mult = 1;
        int digits = 0;
        for (int64_t q = (int64_t) (base + k) - 1; q >= 0 && digits < 20; --q) {
            const unsigned char d = (unsigned char) text[q];
            if (d < '0' || d > '9') break;
            id += (uint64_t) (d - '0') * mult;
            mult *= 10;
            ++digits;
        }
        if (digits == 0 || id == 0 || id > n_nodes) { atomicExch(bad, 1); id = 1; }
        step_node[idx] = (uint32_t) (id - 1);
        step_rev[idx] = b[k] == '-' ? 1 : 0;
        ++idx;
    }
}

// first step of every path = orientation characters before the start of its field (fields are packed back to back)
__global__ void gfa_path_first_kernel(const char* text, const uint64_t* field_begin, uint32_t n_fields, const uint64_t* chunk_first, uint64_t* path_first) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_fields) return;
    const uint64_t b = field_begin[p], chunk = b / GFA_CHUNK;
    uint64_t c = chunk_first[chunk];
    for (uint64_t i = chunk * GFA_CHUNK; i < b; ++i) c += is_end(text[i]) ? 1u : 0u;
    path_first[p] = c;
}

}  // namespace

// d_text: n_bytes of packed step lists (16-byte aligned, padded to a multiple of 16 with zeros); d_field_begin: [n_fields] offsets.
// Outputs (device, caller frees): *d_step_node [S], *d_step_rev [S]; d_path_first [n_fields + 1] is filled; *S_out = steps.
cudaError_t launch_parse_gfa_paths(const char* d_text, uint64_t n_bytes, const uint64_t* d_field_begin, uint32_t n_fields, uint32_t n_nodes,
                                   uint64_t* d_path_first, uint32_t** d_step_node, uint8_t** d_step_rev, uint64_t* S_out, int* bad_out,
                                   cudaStream_t stream) {
    *d_step_node = nullptr; *d_step_rev = nullptr; *S_out = 0; *bad_out = 0;
    const uint64_t n_chunks = (n_bytes + GFA_CHUNK - 1) / GFA_CHUNK;
    uint64_t* d_cnt = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int* d_bad = nullptr;
    cudaError_t e = cudaMalloc(&d_cnt, (n_chunks + 1) * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_bad, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, sizeof(int), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt + n_chunks, 0, sizeof(uint64_t), stream);
    if (e == cudaSuccess && n_chunks) { gfa_count_kernel<<<(unsigned) n_chunks, GFA_BLOCK, 0, stream>>>(d_text, n_bytes, d_cnt); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_cnt, d_cnt, n_chunks + 1, stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_cnt, d_cnt, n_chunks + 1, stream);
    uint64_t S = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&S, d_cnt + n_chunks, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaMalloc(d_step_node, (S ? S : 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(d_step_rev, S ? S : 1);
    if (e == cudaSuccess && n_chunks) {
        gfa_parse_kernel<<<(unsigned) n_chunks, GFA_BLOCK, 0, stream>>>(d_text, n_bytes, d_cnt, n_nodes, *d_step_node, *d_step_rev, d_bad);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && n_fields) {
        gfa_path_first_kernel<<<(n_fields + 127) / 128, 128, 0, stream>>>(d_text, d_field_begin, n_fields, d_cnt, d_path_first);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_path_first + n_fields, &S, sizeof(uint64_t), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(bad_out, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp); cudaFree(d_cnt); cudaFree(d_bad);
    if (e != cudaSuccess) { cudaFree(*d_step_node); cudaFree(*d_step_rev); *d_step_node = nullptr; *d_step_rev = nullptr; return e; }
    *S_out = S;
    return cudaSuccess;
}

__global__ void gather_mid_nodes_kernel(const uint32_t* step_node, uint64_t S, uint64_t tile_steps, uint64_t n_tiles, uint32_t* out) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const uint64_t lo = t * tile_steps, hi = lo + tile_steps < S ? lo + tile_steps : S;
    out[t] = step_node[lo + (hi - lo) / 2];
}
cudaError_t launch_gather_mid_nodes(const uint32_t* d_step_node, uint64_t S, uint64_t tile_steps, uint64_t n_tiles, uint32_t* d_out, cudaStream_t stream) {
    if (!n_tiles) return cudaSuccess;
    gather_mid_nodes_kernel<<<(unsigned) ((n_tiles + 255) / 256), 256, 0, stream>>>(d_step_node, S, tile_steps, n_tiles, d_out);
    return cudaGetLastError();
}

}  // namespace pgsgd
