// pgsgd_kernels.cuh — launch-side declarations shared by pgsgd_kernels.cu and pgsgd_capi.cu
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pgsgd_device.cuh"

namespace pgsgd {

// Parameters of one SGD iteration launch (one cooling-schedule step), passed by value.
struct IterParams {
    SamplerParams sp;
    const StepRec* steps;       // [S] 16-byte step records, HBM resident
    float* xy;                  // 2D: [4N] {x0,y0,x1,y1} per node (one float2 per node end)
    double* x1d;                // 1D: [N]
    const uint8_t* frozen;      // 1D: [N] or nullptr
    uint64_t* rng;              // [4][rng_stride] Xoshiro256+ state of every worker stream (SoA)
    uint64_t rng_stride;
    uint64_t n_streams;         // worker streams launched (one per thread)
    uint64_t quota_base;        // each stream performs quota_base (+1 for stream < quota_rem) counted updates
    uint64_t quota_rem;
    double eta;                 // learning rate of this iteration
    unsigned int* delta_max_bits;  // max |Delta| as ordered float bits, or nullptr when not tracked
    unsigned long long* counted;   // total counted updates (atomicAdd once per block)
    uint32_t flags;             // PGSGD_FLAG_* write flavour bits
    uint32_t smem_paths;        // 1: path_first table staged in shared memory
    // ---- tile sampling (pgsgd_tile_kernel) ----
    uint64_t n_visits;          // tile visits of this launch (all ranks); visit v is handled by CTA v % gridDim of rank v % n_ranks
    uint64_t n_tiles;           // ceil(S / TILE_STEPS)
    uint64_t last_visit_terms;  // terms of the final visit (<= TILE_STEPS)
    uint64_t perm_mul[16];      // per pass: 1 = tile = tile_perm(i, n_tiles, key = perm_add[pass]) (keyed pseudo-random permutation);
    uint64_t perm_add[16];      //           0 = sweep order, tile = (i + perm_add[pass]) % n_tiles
    uint32_t visit_rank, visit_nranks;  // this rank handles visits v with v % visit_nranks == visit_rank
    // ---- multi-GPU peer mode: the coordinate array is PARTITIONED by node range over the GPUs of the box and every
    //      GPU reads / reds the owner's slice directly through NVLink peer memory (no replica, no all-reduce) ----
    uint32_t n_parts;               // 0/1: single array (p.xy / p.x1d); else number of partitions
    uint32_t part_lo[9];            // partition q owns node ranks [part_lo[q], part_lo[q+1])
    float*  part_xy[8];             // part_xy[q] + 4*node is node's {x0,y0,x1,y1} (pointers pre-offset by -4*part_lo[q])
    double* part_x1d[8];
    const uint32_t* tile_list;      // tile mode: the tiles this rank owns (nullptr = all tiles 0..n_tiles-1)
    // ---- verification: when trace != nullptr every drawn term is appended as {ia, ib | flip_a << 62 | flip_b << 63} ----
    unsigned long long* trace;
    unsigned long long* trace_count;
    uint64_t trace_cap;
};

// Parameters of one launch of the pipelined tile kernel (pgsgd_tile2.cu): everything the host can fold is folded.
struct Tile2Params {
    const StepRec* steps;
    float* xy;                  // 2D: [4N]
    double* x1d;                // 1D: [N]
    const uint8_t* frozen;      // 1D: [N] or nullptr
    uint64_t* rng;
    uint64_t rng_stride;
    const uint64_t* path_first; // [P+1]
    uint64_t step_count;
    const float2* ztab;         // [zeta table size] {zeta_n, 1 / (1 - zeta_2 / zeta_n)} in fp32 for this iteration's theta
    uint32_t path_count;
    uint32_t smem_paths;        // 1: path_first staged in shared memory
    uint32_t space, space_max, space_q;
    float space_q_rcp;          // 1 / space_q when the reciprocal quotient is exact after one correction step, else 0 (integer division)
    uint32_t cooling;
    uint32_t pos32;             // every end-adjusted bp position fits 32 bits
    float one_minus_theta, alpha_frac, thresh2;
    int alpha_int;
    double eta;
    float eta_f;
    uint32_t flags;
    unsigned int* delta_max_bits;
    unsigned long long* counted;
    uint64_t n_visits, n_tiles, last_visit_terms;
    uint64_t perm_mul[16], perm_add[16];
    uint32_t visit_rank, visit_nranks;
    uint32_t sweeps, rem_sweeps;   // terms per staged step and visit in the full passes / in the one pass after them
    uint64_t full_passes;
    unsigned long long* trace;
    unsigned long long* trace_count;
    uint64_t trace_cap;
    // multi-GPU peer phases (see IterParams): coordinates partitioned by node range, every partition mapped through NVLink
    uint32_t n_parts;
    uint32_t part_lo[9];
    float*  part_xy[8];
    double* part_x1d[8];
    const uint32_t* tile_list;  // the tiles this rank owns (2048-step tiles), or nullptr
};

constexpr int STRESS_STREAMS = 4096;   // generators of the sampled path stress (== ORC_STRESS_STREAMS of the oracle)
#ifndef PGSGD_TILE_STEPS
#define PGSGD_TILE_STEPS 2048      // experiments: PGSGD_TILE_STEPS=4096 python -m odgi_b200.build (1024, 2048 or 4096)
#endif
constexpr int TILE_STEPS = PGSGD_TILE_STEPS;   // steps staged in shared memory per tile visit (2048: 32 KB of 16-byte records)
static_assert(TILE_STEPS == 1024 || TILE_STEPS == 2048 || TILE_STEPS == 4096, "tile size: 1024, 2048 or 4096 steps");

struct LaunchShape {
    int block;           // threads per block
    int grid;            // blocks
    size_t smem;         // dynamic shared memory bytes
};

// seeds worker streams [0, n) with seed_base + global_stream_offset + t
cudaError_t launch_seed_streams(uint64_t* rng, uint64_t rng_stride, uint64_t n, uint64_t seed_base, cudaStream_t stream);

// one iteration of 2D / 1D PG-SGD
cudaError_t launch_iteration(int dims, int batch, const IterParams& p, const LaunchShape& shape, cudaStream_t stream);
// the same with tile sampling (TILE_STEPS consecutive steps staged in shared memory per visit)
cudaError_t launch_tile_iteration(int dims, int batch, const IterParams& p, const LaunchShape& shape, cudaStream_t stream);
cudaError_t tile_occupancy(int dims, int batch, size_t smem, bool smem_paths, bool tma, int* blocks_per_sm);
// occupancy query for the kernel variant (resident blocks per SM for the given block size / smem)
cudaError_t iteration_occupancy(int dims, int batch, int block, size_t smem, bool smem_paths, int* blocks_per_sm);

// the pipelined tile kernel (pgsgd_tile2.cu): tile_steps 1024 | 2048 | 4096, tma = double-buffered TMA bulk staging
struct Tile2Params;
size_t tile2_smem_bytes(int tile_steps, bool tma, uint32_t path_count, bool* smem_paths);
cudaError_t launch_tile2_iteration(int dims, int tile_steps, bool tma, const Tile2Params& p, const LaunchShape& shape, cudaStream_t stream);
cudaError_t tile2_occupancy(int dims, int tile_steps, bool tma, size_t smem, int* blocks_per_sm);
// largest end-adjusted bp position over all paths (pos + len of every path's last step), for the 32-bit position path
cudaError_t launch_max_path_bp(const StepRec* steps, const uint64_t* first, uint32_t P, unsigned long long* out, cudaStream_t stream);

// packs SoA step arrays into StepRec records on the device
cudaError_t launch_pack_steps(StepRec* out, const uint32_t* step_node, const uint8_t* step_rev, const uint64_t* step_pos,
                              const uint32_t* node_len, uint64_t n, uint64_t out_offset, uint32_t* depth, cudaStream_t stream);

// the same with the positions derived on the device (device-wide scan of node lengths); *bad is set to 1 when a step refers
// to a node rank >= n_nodes
cudaError_t launch_flatten_on_device(StepRec* out, const uint32_t* step_node, const uint8_t* step_rev, const uint32_t* node_len,
                                     const uint64_t* first, uint32_t P, uint32_t n_nodes, uint64_t n, uint64_t* scratch_len, int* bad,
                                     uint32_t* depth, cudaStream_t stream);

// repeated node visits inside tiles of TILE_STEPS consecutive steps, summed over all tiles (n must start on a tile boundary)
cudaError_t launch_tile_repeats(const uint32_t* step_node, uint64_t n, unsigned long long* total_dups, cudaStream_t stream);

// 1D node order on the device: node ranks sorted by ([component key,] x, rank), stable radix sorts; d_component may be null
cudaError_t launch_order_1d(const double* x, const uint32_t* d_component, uint64_t* order_out, uint64_t n, cudaStream_t stream);

// sorting-goodness sums (odgi stats -l -g -s -d) of the graph sorted by d_order (device, [N] node ranks; null = as it is);
// flags bit 0: -g (gap links not penalised), bit 1: -d (orientation changes penalised); h_acc9: see goodness_kernel
cudaError_t launch_goodness(const StepRec* steps, const uint64_t* first, const uint64_t* h_first, uint32_t P, uint64_t S, uint64_t N,
                            const uint32_t* d_node_len, const uint64_t* d_order, uint32_t flags, unsigned long long* h_acc9, cudaStream_t stream);

// ---- GFA P-line step lists parsed on the device (pgsgd_gfa.cu) ----
cudaError_t launch_parse_gfa_paths(const char* d_text, uint64_t n_bytes, const uint64_t* d_field_begin, uint32_t n_fields, uint32_t n_nodes,
                                   uint64_t* d_path_first, uint32_t** d_step_node, uint8_t** d_step_rev, uint64_t* S_out, int* bad_out,
                                   cudaStream_t stream);
cudaError_t launch_gather_mid_nodes(const uint32_t* d_step_node, uint64_t S, uint64_t tile_steps, uint64_t n_tiles, uint32_t* d_out, cudaStream_t stream);

// ---- `.lay` on the device (pgsgd_lay.cu) ----
struct LayEncoded {
    double min_value;
    uint64_t n_vals, z_bits, sp_bits;
    unsigned width;
    unsigned long long* d_z;      // Elias-delta code words (device; the caller frees)
    unsigned long long* d_sp;     // sample table words (device; the caller frees) — OR tail_bits into words tail_word, tail_word + 1
    uint64_t tail_word;
    unsigned long long tail_bits[2];
};
// per-component {min_x, min_y, max_y} of the resident 2D coordinates (h_stats: [3K] doubles)
cudaError_t launch_component_ranges(const float* xy, const uint32_t* d_comp, uint64_t n_nodes, uint32_t K, double* h_stats, cudaStream_t stream);
// Layout(X, Y).serialize of the resident coordinates (as doubles, moved by the per-component offsets when d_comp != null)
cudaError_t launch_encode_lay(const float* xy, uint64_t n_nodes, const uint32_t* d_comp, const double* d_x_off, const double* d_y_off,
                              LayEncoded* out, cudaStream_t stream);

// coordinate format conversion: reference X/Y (double, index 2*node+end) <-> device float4-per-node
cudaError_t launch_xy_from_XY(float* xy, const double* X, const double* Y, uint64_t n_nodes, cudaStream_t stream);
cudaError_t launch_XY_from_xy(double* X, double* Y, const float* xy, uint64_t n_nodes, cudaStream_t stream);
// 1D default initialisation: X[rank] = cumulative bp (path_sgd.cpp:63-69) from an exclusive scan done on the host side
// multi-GPU helpers
cudaError_t launch_scale_f32(float* a, uint64_t n, float s, cudaStream_t stream);
cudaError_t launch_scale_f64(double* a, uint64_t n, double s, cudaStream_t stream);
cudaError_t launch_sub_f32(float* out, const float* a, const float* b, uint64_t n, cudaStream_t stream);   // out = a - b
cudaError_t launch_add_f32(float* out, const float* a, const float* b, uint64_t n, cudaStream_t stream);   // out = a + b
cudaError_t launch_sub_f64(double* out, const double* a, const double* b, uint64_t n, cudaStream_t stream);
cudaError_t launch_add_f64(double* out, const double* a, const double* b, uint64_t n, cudaStream_t stream);

// sampled path stress (local = 1: the near-pair variant, orc_local_stress_*): per-stream partial sums (STRESS_STREAMS entries each), `per` pairs per stream
cudaError_t launch_stress(int dims, int local, const uint64_t* first, uint32_t P, uint64_t S, const StepRec* steps, const float* xy,
                          const double* x1d, uint64_t per, uint64_t seed, double* acc_out, unsigned long long* used_out, cudaStream_t stream);

// verification hook: first n_terms draws of one stream, produced by the same device sampler
struct SampleOut {
    uint64_t* step_index; uint32_t* path; uint64_t* rank_a; uint64_t* rank_b; uint32_t* node_a; uint32_t* node_b;
    uint64_t* pos_a; uint64_t* pos_b; uint8_t* end_a; uint8_t* end_b; uint8_t* valid;
};
cudaError_t launch_sample_terms(int dims, const SamplerParams& sp, const StepRec* steps, uint64_t seed, uint64_t n_terms,
                                const SampleOut& out, cudaStream_t stream);

}  // namespace pgsgd
