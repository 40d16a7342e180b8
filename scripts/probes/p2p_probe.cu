// p2p_probe.cu — measures the rate of fine-grained (8-byte) random peer loads and peer REDs over NVLink between two
// B200s, the access pattern a coordinate array partitioned across GPUs would see.  Single process, two devices.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31);
}
// mode 0: random 8B load; 1: random 8B red.add.v2.f32; 2: load + red (the far-b coordinate access of one PG-SGD term)
__global__ void probe(float2* remote, uint64_t n, uint64_t per_thread, int mode, float* sink) {
    uint64_t tid = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    uint64_t s = tid * 0x9e3779b97f4a7c15ULL + 12345;
    float acc = 0;
    for (uint64_t k = 0; k < per_thread; ++k) {
        s += 0x9e3779b97f4a7c15ULL;
        uint64_t i = mix(s) % n;
        if (mode != 1) {
            float2 v;
            asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(remote + i));
            acc += v.x + v.y;
        }
        if (mode != 0) {
            float d = mode == 2 ? acc * 1e-30f : 1e-30f;
            asm volatile("red.relaxed.sys.global.add.v2.f32 [%0], {%1, %2};" :: "l"(remote + i), "f"(d), "f"(d) : "memory");
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs, have %d\n", nd); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("peer access 0->1: %d\n", can);
    const uint64_t n = 1ull << 24;  // 16M float2 = 128 MB on device 1
    float2* buf1; float2* buf0; float* sink;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&buf1, n * sizeof(float2))); CK(cudaMemset(buf1, 0, n * sizeof(float2)));
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&buf0, n * sizeof(float2))); CK(cudaMemset(buf0, 0, n * sizeof(float2)));
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char* names[3] = {"load 8B", "red.add 8B", "load+red 8B"};
    for (int target = 0; target < 2; ++target) {
        float2* p = target ? buf1 : buf0;
        for (int mode = 0; mode < 3; ++mode) {
            for (int blocks_per_sm = 2; blocks_per_sm <= 8; blocks_per_sm *= 2) {
                int grid = 148 * blocks_per_sm, block = 256; uint64_t per_thread = 2000;
                probe<<<grid, block>>>(p, n, 100, mode, sink);
                CK(cudaEventRecord(e0)); probe<<<grid, block>>>(p, n, per_thread, mode, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                double ops = (double) grid * block * per_thread;
                printf("%s  %-12s blocks/SM=%d  %.2f G ops/s\n", target ? "REMOTE(NVLink)" : "local         ", names[mode], blocks_per_sm, ops / ms / 1e6);
            }
        }
    }
    return 0;
}
