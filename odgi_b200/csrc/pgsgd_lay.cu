// pgsgd_lay.cu — the step right after the 2D hot path, on the device:
//   * the per-component offsetting `odgi layout` applies before writing (src/subcommand/layout_main.cpp:402-435): bounding
//     box of every weak component, components stacked vertically with a 1000-unit border;
//   * odgi's `.lay` container (src/algorithms/layout.cpp:43-61): min_value + sdsl::enc_vector<elias_delta, 128> over the bit
//     patterns of (coordinate - min_value), X/Y interleaved — Elias-delta coding of the deltas, a sample table every 128 values.
// The byte format is restated in odgi_b200/host/lay_format.hpp (pinned byte-for-byte on files the reference wrote); this is
// the same encoder as three data-parallel passes: code lengths -> exclusive scan (bit offsets) -> every value ORs its own
// code into place.  Serial on the host it costs ~0.3 s per 10 M nodes plus the 32-byte-per-node download in fp64; here the
// device ships the finished file (about 35 bytes per node).
#include "pgsgd_kernels.cuh"

#include <cstring>
#include <vector>

#include <cub/device/device_scan.cuh>

namespace pgsgd {

namespace {

__device__ __forceinline__ unsigned long long enc_ordered(double v) {   // order-preserving map double -> u64
    const unsigned long long b = (unsigned long long) __double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ double dec_ordered(unsigned long long u) {
    const unsigned long long b = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFULL) : ~u;
    double d;
#ifdef __CUDA_ARCH__
    d = __longlong_as_double((long long) b);
#else
    memcpy(&d, &b, 8);
#endif
    return d;
}

// the coordinate the reference would serialise for value index j (= 4 * node + 2 * end + isY, the order of d_xy): the fp32
// device coordinate as a double, moved by its component's offsets (X -= x_offset, Y += y_offset)
__device__ __forceinline__ double lay_coord(const float* xy, uint64_t j, const uint32_t* comp, const double* x_off, const double* y_off) {
    double c = (double) xy[j];
    if (comp) {
        const uint32_t k = comp[j >> 2];
        c = (j & 1) ? c + y_off[k] : c - x_off[k];
    }
    return c;
}

// stats[3 * k + {0,1,2}] = ordered(min_x), ordered(min_y), ordered(max_y) of component k
__global__ void component_range_kernel(const float4* xy, const uint32_t* comp, uint64_t n_nodes, unsigned long long* stats) {
    const uint64_t n = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    const float4 v = xy[n];
    const uint32_t k = comp[n];
    atomicMin(stats + 3 * (uint64_t) k, enc_ordered((double) fminf(v.x, v.z)));
    atomicMin(stats + 3 * (uint64_t) k + 1, enc_ordered((double) fminf(v.y, v.w)));
    atomicMax(stats + 3 * (uint64_t) k + 2, enc_ordered((double) fmaxf(v.y, v.w)));
}

__global__ void min_value_kernel(const float* xy, uint64_t n_vals, const uint32_t* comp, const double* x_off, const double* y_off, unsigned long long* out) {
    unsigned long long m = ~0ULL;
    for (uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; j < n_vals; j += (uint64_t) gridDim.x * blockDim.x)
        m = min(m, enc_ordered(lay_coord(xy, j, comp, x_off, y_off)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(out, m);
}

__device__ __forceinline__ unsigned long long lay_bits(const float* xy, uint64_t j, const uint32_t* comp, const double* x_off, const double* y_off, double min_value) {
    return (unsigned long long) __double_as_longlong(__dsub_rn(lay_coord(xy, j, comp, x_off, y_off), min_value));
}

__device__ __forceinline__ unsigned hi_bit64(unsigned long long x) { return 63u - (unsigned) __clzll((long long) x); }

// code length of value j (0 for the sampled ones, which go to the sample table) and the largest sampled value
__global__ void code_length_kernel(const float* xy, uint64_t n_vals, const uint32_t* comp, const double* x_off, const double* y_off, double min_value,
                                   uint64_t* len_out, unsigned long long* max_sample) {
    const uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_vals) return;
    if (j == n_vals) { len_out[j] = 0; return; }      // one extra slot: the scan then leaves the total there
    const unsigned long long v = lay_bits(xy, j, comp, x_off, y_off, min_value);
    if ((j & 127) == 0) {
        len_out[j] = 0;
        atomicMax(max_sample, v);
        return;
    }
    const unsigned long long d = v - lay_bits(xy, j - 1, comp, x_off, y_off, min_value);   // mod 2^64; 0 stands for 2^64
    const unsigned len = d ? hi_bit64(d) + 1 : 65;
    len_out[j] = len + 2 * hi_bit64(len);
}

__device__ __forceinline__ void or_bits(unsigned long long* words, uint64_t bit, unsigned __int128 code) {
    const uint64_t w = bit >> 6;
    const unsigned sh = (unsigned) (bit & 63);
    const unsigned long long lo = (unsigned long long) code, hi = (unsigned long long) (code >> 64);
    unsigned long long a = lo << sh, b = sh ? (lo >> (64 - sh)) | (hi << sh) : hi, c = sh ? hi >> (64 - sh) : 0;
    if (a) atomicOr(words + w, a);
    if (b) atomicOr(words + w + 1, b);
    if (c) atomicOr(words + w + 2, c);
}

__global__ void write_codes_kernel(const float* xy, uint64_t n_vals, const uint32_t* comp, const double* x_off, const double* y_off, double min_value,
                                   const uint64_t* bit_off, unsigned long long* z_words, unsigned long long* sp_words, unsigned width) {
    const uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_vals) return;
    const unsigned long long v = lay_bits(xy, j, comp, x_off, y_off, min_value);
    if ((j & 127) == 0) {   // sample: (absolute value, bit offset into z), `width` bits each (enc_vector.hpp:341-350)
        const uint64_t s = j >> 7;
        or_bits(sp_words, (2 * s) * (uint64_t) width, (unsigned __int128) v);
        or_bits(sp_words, (2 * s + 1) * (uint64_t) width, (unsigned __int128) bit_off[j]);
        return;
    }
    const unsigned long long d = v - lay_bits(xy, j - 1, comp, x_off, y_off, min_value);
    const unsigned len = d ? hi_bit64(d) + 1 : 65;
    const unsigned l2 = hi_bit64(len);
    // LSB first: l2 zeros and a one | the length without its top bit | the value without its top bit (coder_elias_delta.hpp:198-212)
    unsigned __int128 code = (unsigned __int128) 1 << l2;
    if (l2) {
        code |= (unsigned __int128) (len & ((1u << l2) - 1)) << (l2 + 1);
        const unsigned long long low = len >= 65 ? d : (d & ((1ULL << (len - 1)) - 1));
        code |= (unsigned __int128) low << (2 * l2 + 1);
    }
    or_bits(z_words, bit_off[j], code);
}

}  // namespace

// Component ranges -> host: stats_out[3K] doubles {min_x, min_y, max_y}
cudaError_t launch_component_ranges(const float* xy, const uint32_t* d_comp, uint64_t n_nodes, uint32_t K, double* h_stats, cudaStream_t stream) {
    unsigned long long* d_stats = nullptr;
    cudaError_t e = cudaMalloc(&d_stats, 3 * (size_t) K * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    std::vector<unsigned long long> init(3 * (size_t) K);
    for (uint32_t k = 0; k < K; ++k) { init[3 * k] = ~0ULL; init[3 * k + 1] = ~0ULL; init[3 * k + 2] = 0ULL; }
    e = cudaMemcpyAsync(d_stats, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess && n_nodes) {
        component_range_kernel<<<(unsigned) ((n_nodes + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(xy), d_comp, n_nodes, d_stats);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(init.data(), d_stats, init.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_stats);
    if (e == cudaSuccess) for (size_t i = 0; i < init.size(); ++i) h_stats[i] = dec_ordered(init[i]);
    return e;
}

// Encodes into device buffers (z words, sp words) and reports the header fields; the caller assembles the file.
cudaError_t launch_encode_lay(const float* xy, uint64_t n_nodes, const uint32_t* d_comp, const double* d_x_off, const double* d_y_off,
                              LayEncoded* out, cudaStream_t stream) {
    const uint64_t n_vals = 4 * n_nodes;
    memset(out, 0, sizeof(*out));
    if (!n_vals) return cudaErrorInvalidValue;
    unsigned long long* d_scal = nullptr;   // [0] min (ordered), [1] max sample
    uint64_t* d_off = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaMalloc(&d_scal, 2 * sizeof(unsigned long long));
    const unsigned long long init[2] = {~0ULL, 0ULL};
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_scal, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) { min_value_kernel<<<148 * 8, 256, 0, stream>>>(xy, n_vals, d_comp, d_x_off, d_y_off, d_scal); e = cudaGetLastError(); }
    unsigned long long h_scal[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_scal, d_scal, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    const double min_value = dec_ordered(h_scal[0]);
    if (e == cudaSuccess) e = cudaMalloc(&d_off, (n_vals + 1) * sizeof(uint64_t));
    if (e == cudaSuccess) {
        code_length_kernel<<<(unsigned) ((n_vals + 1 + 255) / 256), 256, 0, stream>>>(xy, n_vals, d_comp, d_x_off, d_y_off, min_value, d_off, d_scal + 1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_off, d_off, n_vals + 1, stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_off, d_off, n_vals + 1, stream);
    uint64_t z_bits = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&z_bits, d_off + n_vals, sizeof(uint64_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_scal + 1, d_scal + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    const uint64_t samples = (n_vals + 127) / 128;
    const unsigned long long wmax = h_scal[1] > z_bits + 1 ? h_scal[1] : z_bits + 1;
    unsigned width = 1;
    while (width < 64 && (wmax >> width)) ++width;
    const uint64_t sp_bits = (2 * samples + 2) * (uint64_t) width;
    const uint64_t z_words = (z_bits + 63) / 64, sp_words = (sp_bits + 63) / 64;
    unsigned long long *d_z = nullptr, *d_sp = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&d_z, (z_words + 3) * sizeof(unsigned long long));      // spare words: or_bits may touch w + 2
    if (e == cudaSuccess) e = cudaMalloc(&d_sp, (sp_words + 3) * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_z, 0, (z_words + 3) * sizeof(unsigned long long), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_sp, 0, (sp_words + 3) * sizeof(unsigned long long), stream);
    if (e == cudaSuccess) {
        write_codes_kernel<<<(unsigned) ((n_vals + 255) / 256), 256, 0, stream>>>(xy, n_vals, d_comp, d_x_off, d_y_off, min_value, d_off, d_z, d_sp, width);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) {   // the closing sample (0, |z| + 1)
        const unsigned long long last = z_bits + 1;
        const uint64_t bit = (2 * samples + 1) * (uint64_t) width;
        unsigned long long w2[3] = {0, 0, 0};
        const unsigned sh = (unsigned) (bit & 63);
        w2[0] = last << sh;
        if (sh) w2[1] = last >> (64 - sh);
        // read-modify-write of at most two words after the kernel: done on the host side of the copy below
        out->tail_word = bit >> 6;
        out->tail_bits[0] = w2[0];
        out->tail_bits[1] = w2[1];
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp); cudaFree(d_off); cudaFree(d_scal);
    if (e != cudaSuccess) { cudaFree(d_z); cudaFree(d_sp); return e; }
    out->min_value = min_value;
    out->n_vals = n_vals;
    out->z_bits = z_bits;
    out->sp_bits = sp_bits;
    out->width = width;
    out->d_z = d_z;
    out->d_sp = d_sp;
    return cudaSuccess;
}

}  // namespace pgsgd
