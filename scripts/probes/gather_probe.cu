// gather_probe.cu — the memory-traffic SKELETON of one PG-SGD term on one B200, without the arithmetic: how many terms per
// second does the access pattern alone allow?  Per term:
//   seq   one 16 B step record from a sequential stream (the staged tile; coalesced, evict_first)
//   far   one 16 B step record at a random place of the same array (the partner; a DRAM miss in a 6.7 GB array)
//   ca    the first node's float4 (neighbouring lanes -> neighbouring nodes: coalesced), cb the partner's float4 (random)
//   red   two red.global.add.v2.f32, one into each of those float4s
// Modes: 0 = far only, 1 = seq + far, 2 = seq + far + ca/cb loads, 3 = all of it.  Every thread keeps UNROLL independent
// terms in flight (all loads of a stage issued before the first use), 4 CTAs x 256 threads per SM like the kernel.
// Usage: gather_probe [records_millions] [nodes_millions] [far_share_percent]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)


__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t policy(bool last) {
    uint64_t p;
    if (last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld_rec(const uint4* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_xy(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void red2(float* p, float a, float b, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(p), "f"(a), "f"(b), "l"(pol) : "memory");
}

// every CTA walks "tiles" of 2048 consecutive records at pseudo-random places, like the kernel
template <int UNROLL>
__global__ void __launch_bounds__(256, 4) skeleton(const uint4* rec, uint64_t n_rec, float4* xy, uint32_t n_nodes, uint64_t tiles_per_cta,
                                                   int mode, uint32_t far_share_256, uint64_t ratio_fx, float* sink) {
    const uint64_t n_tiles = n_rec / 2048;
    float acc = 0;
    const uint64_t pol_first = policy(false), pol_last = policy(true);
    uint64_t s = (blockIdx.x * 256ull + threadIdx.x) * 0x9e3779b97f4a7c15ULL + 777;
    for (uint64_t t = 0; t < tiles_per_cta; ++t) {
        const uint64_t tile = __umul64hi(mix(blockIdx.x * 0x10001ull + t * 0x9e3779b97f4a7c15ULL), n_tiles);
        const uint4* base = rec + tile * 2048;
#pragma unroll 1
        for (int r = 0; r < 2048 / 256; r += UNROLL) {
            uint4 a[UNROLL], b[UNROLL];
            uint32_t na[UNROLL], nb[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t j = (r + u) * 256 + threadIdx.x;
                if (mode >= 1) a[u] = ld_rec(base + j, pol_first); else a[u] = make_uint4(j, 0, 0, 0);
                s += 0x9e3779b97f4a7c15ULL;
                const uint64_t h = mix(s);
                // the partner: far_share of them anywhere in the array, the rest inside the tile
                const uint64_t far = ((uint32_t) h & 255u) < far_share_256 ? __umul64hi(h << 8, n_rec) : tile * 2048 + ((h >> 8) & 2047);
                b[u] = ld_rec(rec + far, pol_first);
                na[u] = (uint32_t) (((tile * 2048 + j) * ratio_fx) >> 32);   // steps of a path run through neighbouring nodes
                nb[u] = (uint32_t) ((far * ratio_fx) >> 32);
            }
            if (mode >= 2) {
                float4 ca[UNROLL], cb[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    na[u] = min(na[u] + (a[u].x & 1u), n_nodes - 1);   // depend on the landed records, like the kernel
                    nb[u] = min(nb[u] + (b[u].x & 1u), n_nodes - 1);
                    ca[u] = ld_xy(xy + na[u], pol_last);
                    cb[u] = ld_xy(xy + nb[u], pol_last);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const float d = (ca[u].x - cb[u].x) * 1e-30f;
                    if (mode >= 3) {
                        red2(reinterpret_cast<float*>(xy + na[u]), d, d, pol_last);
                        red2(reinterpret_cast<float*>(xy + nb[u]) + 2, -d, -d, pol_last);
                    } else acc += d;
                }
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) acc += (float) (a[u].x ^ b[u].x);
            }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main(int argc, char** argv) {
    const uint64_t n_rec = (uint64_t) ((argc > 1 ? atof(argv[1]) : 419.0) * 1e6) / 2048 * 2048;
    const uint32_t n_nodes = (uint32_t) ((argc > 2 ? atof(argv[2]) : 5.5) * 1e6);
    const int far_pct = argc > 3 ? atoi(argv[3]) : 76;
    CK(cudaSetDevice(0));
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32));
    uint4* rec; float4* xy; float* sink;
    CK(cudaMalloc(&rec, n_rec * sizeof(uint4))); CK(cudaMemset(rec, 1, n_rec * sizeof(uint4)));
    CK(cudaMalloc(&xy, (size_t) n_nodes * sizeof(float4))); CK(cudaMemset(xy, 0, (size_t) n_nodes * sizeof(float4)));
    CK(cudaMalloc(&sink, 4));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int ctas = sms * 4;
    const uint64_t tiles_per_cta = 256;
    const double terms = (double) ctas * tiles_per_cta * 2048;
    printf("records %.1f M (%.2f GB), nodes %.2f M (%.1f MB of float4), far share %d %%, %d CTAs x 256, 2 / 4 / 8 independent terms per thread\n",
           n_rec / 1e6, n_rec * 16 / 1e9, n_nodes / 1e6, n_nodes * 16 / 1e6, far_pct, ctas);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[] = {"far record only", "seq + far records", "records + coordinate loads", "records + coordinate loads + reds (whole term)"};
    for (int unroll = 2; unroll <= 8; unroll *= 2)
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            const uint32_t fs = (uint32_t) (far_pct * 256 / 100);
            const uint64_t fx = ((uint64_t) n_nodes << 32) / n_rec;
            if (unroll == 2) skeleton<2><<<ctas, 256>>>(rec, n_rec, xy, n_nodes, tiles_per_cta, mode, fs, fx, sink);
            else if (unroll == 4) skeleton<4><<<ctas, 256>>>(rec, n_rec, xy, n_nodes, tiles_per_cta, mode, fs, fx, sink);
            else skeleton<8><<<ctas, 256>>>(rec, n_rec, xy, n_nodes, tiles_per_cta, mode, fs, fx, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        printf("  in flight %d  mode %d  %-48s %7.2f ms  %6.1f G terms/s\n", unroll, mode, names[mode], best, terms / best / 1e6);
    }
    return 0;
}
