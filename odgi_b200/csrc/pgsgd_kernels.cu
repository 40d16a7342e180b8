// pgsgd_kernels.cu — hand-written sm_100a kernels of the path-guided SGD hot path.
//
// Replaces: cuda::gpu_layout_kernel / update_pos_gpu / cuda_rnd_zipf (src/cuda/layout.cu:89-287) and the CPU
// worker lambdas (src/algorithms/path_sgd_layout.cpp:165-377, src/algorithms/path_sgd.cpp:205-406).
//
// Two iteration kernels, one launch per cooling-schedule step:
//   pgsgd_iter_kernel  "stream sampling": a persistent grid in which every thread is one reference-style worker stream
//       (register-resident Xoshiro256+, bit-identical draws to a CPU worker thread); two random 16-byte step-record
//       loads, two 8-byte coordinate loads and two 8-byte coordinate reds per term.
//   pgsgd_tile_kernel  "tile sampling": a CTA stages 2048 consecutive step records in shared memory with coalesced
//       128-bit loads and uses each once as a term's first step; partners inside the tile come from shared memory.
// Both: per-path step offsets in shared memory (binary search), L2::evict_first on the step stream and L2::evict_last
// on the coordinates, red.global.add coordinate writes by default.  No tensor cores: the path is an integer/byte
// gather-scatter bound by the random-sector rate of HBM, L2 tag lookups and instruction issue (profiles/).
// Also here: flatten-to-device (record packing, scan-derived positions), the sampled path stress, helpers.
#include "pgsgd_kernels.cuh"

#include <cstdio>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

namespace pgsgd {

namespace {

// Racy (last-writer-wins) coordinate writes, selected by PGSGD_FLAG_EXCH_WRITE / PGSGD_FLAG_PLAIN_STORE: one 64-bit
// atom.exch per node end = the reference kernel's four atomicExch(float*) (layout.cu:184-187) with x and y of an end
// written together, or a plain st.global.  Measured on B200 (profiles/): L2-side atomics sustain 21 G updates/s where
// plain st.global (any cache operator) plateaus at 14 G updates/s.  The default write is red.global.add (see the kernel).
__device__ __forceinline__ void st_coord(float2* p, float2 v, bool plain_store) {
    if (plain_store) {
        __stcg(p, v);
    } else {
        unsigned long long bits = ((unsigned long long) __float_as_uint(v.y) << 32) | __float_as_uint(v.x);
        atomicExch(reinterpret_cast<unsigned long long*>(p), bits);
    }
}
__device__ __forceinline__ void st_coord(double* p, double v, bool plain_store) {
    if (plain_store) __stcg(p, v);
    else atomicExch(reinterpret_cast<unsigned long long*>(p), (unsigned long long) __double_as_longlong(v));
}

// coordinate addressing: one array, or the owner's slice in peer mode (n_parts <= 8: a branch-free range count)
__device__ __forceinline__ float2* coord2_ptr(const IterParams& p, uint32_t node, uint32_t end) {
    if (p.n_parts <= 1) return reinterpret_cast<float2*>(p.xy) + ((uint64_t) node * 2 + end);
    uint32_t q = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) q += (k < (int) p.n_parts && node >= p.part_lo[k]) ? 1u : 0u;
    return reinterpret_cast<float2*>(p.part_xy[q]) + ((uint64_t) node * 2 + end);
}
__device__ __forceinline__ double* coord1_ptr(const IterParams& p, uint32_t node) {
    if (p.n_parts <= 1) return p.x1d + node;
    uint32_t q = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) q += (k < (int) p.n_parts && node >= p.part_lo[k]) ? 1u : 0u;
    return p.part_x1d[q] + node;
}

#ifndef TILE_B1_CTAS
#define TILE_B1_CTAS 4
#endif

template <int BATCH>
struct MinBlocks {
    static constexpr int value = BATCH >= 4 ? 2 : (BATCH == 2 ? 3 : 4);
};

// --------------------------------------------------------------------------------------------------
// the iteration kernel
// --------------------------------------------------------------------------------------------------
template <int DIMS, int BATCH, bool SMEM_PATHS>
__global__ void __launch_bounds__(256, MinBlocks<BATCH>::value) pgsgd_iter_kernel(const __grid_constant__ IterParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long block_counted;
    const uint64_t* first;
    if (SMEM_PATHS) {
        uint64_t* sfirst = reinterpret_cast<uint64_t*>(smem_raw);
        for (uint32_t i = threadIdx.x; i <= p.sp.path_count; i += blockDim.x) sfirst[i] = p.sp.path_first[i];
        first = sfirst;
    } else {
        first = p.sp.path_first;
    }
    if (threadIdx.x == 0) block_counted = 0;
    __syncthreads();

    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t done = 0;
    float delta_max = 0.0f;
    if (tid < p.n_streams) {
        Xoshiro g;
        g.s0 = p.rng[tid];
        g.s1 = p.rng[p.rng_stride + tid];
        g.s2 = p.rng[2 * p.rng_stride + tid];
        g.s3 = p.rng[3 * p.rng_stride + tid];
        const uint64_t quota = p.quota_base + (tid < p.quota_rem ? 1 : 0);
        const float eta_f = __double2float_rn(p.eta);
        const bool atomic_add = (p.flags & 5u) == 0;  // default; PGSGD_FLAG_EXCH_WRITE / PGSGD_FLAG_PLAIN_STORE select a racy write
        const bool st_mode = (p.flags & 4u) != 0;     // PGSGD_FLAG_PLAIN_STORE
        const uint64_t pol_stream = l2_policy_evict_first();  // step records: read once, never reused
        const uint64_t pol_keep = l2_policy_evict_last();     // coordinates: keep resident in L2

        while (done < quota) {
            const uint64_t remaining = quota - done;
            Term t[BATCH];
            uint4 ra[BATCH], rb[BATCH];
            // phase 1: draw the batch (RNG + shared-memory binary search only)
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                t[b].valid = 0;
                if ((uint64_t) b < remaining) draw_term<DIMS>(p.sp, first, g, t[b]);
            }
            // phase 2: all step-record loads of the batch in flight together
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (t[b].valid) {
                    ra[b] = load_step(p.steps, t[b].ia, pol_stream);
                    rb[b] = load_step(p.steps, t[b].ib, pol_stream);
                }
            }
            if (DIMS == 2) {
                float2 ca[BATCH], cb[BATCH];
                float2* pa[BATCH];
                float2* pb[BATCH];
                float dij[BATCH];
                // phase 3: integer path distance + coordinate loads
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    if (t[b].valid) {
                        uint64_t pos_a = step_pos(ra[b]), pos_b = step_pos(rb[b]);
                        const uint32_t rev_a = ra[b].x & 1u, rev_b = rb[b].x & 1u;
                        // end choice (path_sgd_layout.cpp:252-269): flip moves to the far end along the path
                        uint32_t end_a = rev_a, end_b = rev_b;
                        if (t[b].flip_a) { pos_a += ra[b].y; end_a ^= 1u; }
                        if (t[b].flip_b) { pos_b += rb[b].y; end_b ^= 1u; }
                        const uint64_t dpos = pos_a > pos_b ? pos_a - pos_b : pos_b - pos_a;
                        dij[b] = dpos ? __ull2float_rn(dpos) : 1e-9f;  // term_dist == 0 -> 1e-9 (:283-285)
                        pa[b] = coord2_ptr(p, ra[b].x >> 1, end_a);
                        pb[b] = coord2_ptr(p, rb[b].x >> 1, end_b);
                        ca[b] = ld_coord2(pa[b], pol_keep);
                        cb[b] = ld_coord2(pb[b], pol_keep);
                    }
                }
                // phase 4: the update (path_sgd_layout.cpp:294-363) in fp32, applied in draw order
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    if (t[b].valid) {
                        float mu = __fdiv_rn(eta_f, dij[b]);
                        if (mu > 1.0f) mu = 1.0f;
                        float dx = __fsub_rn(ca[b].x, cb[b].x);
                        const float dy = __fsub_rn(ca[b].y, cb[b].y);
                        if (dx == 0.0f) dx = 1e-9f;
                        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                        const float Delta = __fmul_rn(__fmul_rn(mu, __fsub_rn(mag, dij[b])), 0.5f);
                        delta_max = fmaxf(delta_max, fabsf(Delta));
                        const float r = __fdiv_rn(Delta, mag);
                        const float r_x = __fmul_rn(r, dx);
                        const float r_y = __fmul_rn(r, dy);
                        if (atomic_add) {
                            red_coord2(pa[b], -r_x, -r_y, pol_keep);
                            red_coord2(pb[b], r_x, r_y, pol_keep);
                        } else {
                            const float2 na = make_float2(__fsub_rn(ca[b].x, r_x), __fsub_rn(ca[b].y, r_y));
                            st_coord(pa[b], na, st_mode);
                            // the reference re-reads X[j] after storing X[i]: matters only when both ends alias
                            const float2 base = (pa[b] == pb[b]) ? na : cb[b];
                            st_coord(pb[b], make_float2(__fadd_rn(base.x, r_x), __fadd_rn(base.y, r_y)), st_mode);
                        }
                        ++done;
                    }
                }
            } else {
                double xa[BATCH], xb[BATCH];
                double* qa[BATCH];
                double* qb[BATCH];
                double dij[BATCH];
                uint32_t upd[BATCH];  // bit0: move a, bit1: move b, bit2: counted
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    upd[b] = 0;
                    if (t[b].valid) {
                        const uint32_t na = ra[b].x >> 1, nb = rb[b].x >> 1;
                        uint32_t u = 3u;
                        if (p.frozen) {  // odgi sort -H target nodes (path_sgd.cpp:290-297)
                            if (p.frozen[na]) u &= ~1u;
                            if (p.frozen[nb]) u &= ~2u;
                        }
                        // 1D uses node starts only (path_sgd.cpp:305-306)
                        const double d = fabs(__dsub_rn(__ull2double_rn(step_pos(ra[b])), __ull2double_rn(step_pos(rb[b]))));
                        dij[b] = d;
                        if (u == 0) {
                            upd[b] = 4u;  // both frozen: counted, nothing moves (:298-302)
                        } else if (d != 0.0) {  // d == 0: `continue`, not counted (:320-323)
                            upd[b] = u | 4u;
                            qa[b] = coord1_ptr(p, na);
                            qb[b] = coord1_ptr(p, nb);
                            xa[b] = ld_coord1(qa[b], pol_keep);
                            xb[b] = ld_coord1(qb[b], pol_keep);
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    if (upd[b] & 4u) {
                        if (upd[b] & 3u) {  // path_sgd.cpp:332-392 in fp64
                            double mu = __dmul_rn(p.eta, __ddiv_rn(1.0, dij[b]));
                            if (mu > 1.0) mu = 1.0;
                            double dx = __dsub_rn(xa[b], xb[b]);
                            if (dx == 0.0) dx = 1e-9;
                            const double mag = fabs(dx);
                            const double Delta = __dmul_rn(__dmul_rn(mu, __dsub_rn(mag, dij[b])), 0.5);
                            delta_max = fmaxf(delta_max, (float) fabs(Delta));
                            const double r_x = __dmul_rn(__ddiv_rn(Delta, mag), dx);
                            if (atomic_add) {
                                if (upd[b] & 1u) red_coord1(qa[b], -r_x, pol_keep);
                                if (upd[b] & 2u) red_coord1(qb[b], r_x, pol_keep);
                            } else {
                                const double na = __dsub_rn(xa[b], r_x);
                                if (upd[b] & 1u) st_coord(qa[b], na, st_mode);
                                const double base = (qa[b] == qb[b] && (upd[b] & 1u)) ? na : xb[b];
                                if (upd[b] & 2u) st_coord(qb[b], __dadd_rn(base, r_x), st_mode);
                            }
                        }
                        ++done;
                    }
                }
            }
        }
        p.rng[tid] = g.s0;
        p.rng[p.rng_stride + tid] = g.s1;
        p.rng[2 * p.rng_stride + tid] = g.s2;
        p.rng[3 * p.rng_stride + tid] = g.s3;
    }

    // per-block bookkeeping: one global atomic per block
    unsigned long long wsum = done;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&block_counted, wsum);
    if (p.delta_max_bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) delta_max = fmaxf(delta_max, __shfl_xor_sync(0xffffffffu, delta_max, o));
        if ((threadIdx.x & 31) == 0) atomicMax(p.delta_max_bits, __float_as_uint(delta_max));
    }
    __syncthreads();
    if (threadIdx.x == 0 && block_counted) atomicAdd(p.counted, block_counted);
}


// --------------------------------------------------------------------------------------------------
// tile-sampling iteration kernel
//
// The stream kernel above is bound by the RANDOM-ACCESS rate of HBM (~40 G 32-byte sectors/s measured, profiles/):
// two random step-record reads per term.  Here a CTA stages TILE_STEPS consecutive step records in shared memory with
// fully coalesced 128-bit loads (a DRAM-page-friendly stream) and uses every staged step exactly once as the first
// step of a term; the partner is drawn with the reference's rule (coin, direction, dirty Zipf | uniform in path) and
// is served from the tile when it falls inside it (~45 % of Zipf partners), else by one random 16-byte load.
// Over a launch the tile visits follow per-pass bijections of the tile index, so every step is the first step of
// exactly floor(U/S) terms (+1 for a prefix): the reference's "uniform over all steps" (path_sgd_layout.cpp:175-182)
// with the sampling noise of the first pick removed; the conditional law of the partner is unchanged.
// Consecutive lanes hold consecutive steps, so their coordinate reads/reds of the first node coalesce as well.
// --------------------------------------------------------------------------------------------------
template <int DIMS, int BATCH, bool SMEM_PATHS, bool TMA>
__global__ void __launch_bounds__(256, BATCH >= 4 ? 2 : (BATCH == 2 ? 3 : TILE_B1_CTAS)) pgsgd_tile_kernel(const __grid_constant__ IterParams p) {
    constexpr int ROUNDS = TILE_STEPS / 256;
    static_assert(ROUNDS % BATCH == 0, "tile rounds must be a multiple of the batch");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long block_counted;
    __shared__ __align__(8) uint64_t tile_bar[2];
    // Staging flavours (PGSGD_FLAG_TMA_STAGING): TMA = two tile buffers, the TMA engine (cp.async.bulk) fills one while the
    // CTA works on the other; default = one buffer filled with coalesced 128-bit LDG + STS.  Measured on B200 the second is
    // faster (c4: 33-39 vs 29-36 G updates/s): the doubled shared-memory footprint costs more occupancy / L1 than the
    // exposed ~2 us of tile latency per 2048 terms, which the other resident CTAs already hide.
    uint4* const tile_buf0 = reinterpret_cast<uint4*>(smem_raw);
    uint4* const tile_buf1 = tile_buf0 + TILE_STEPS;
    const uint64_t* first;
    if (SMEM_PATHS) {
        uint64_t* sfirst = reinterpret_cast<uint64_t*>(smem_raw + (TMA ? 2 : 1) * (size_t) TILE_STEPS * sizeof(uint4));
        for (uint32_t i = threadIdx.x; i <= p.sp.path_count; i += blockDim.x) sfirst[i] = p.sp.path_first[i];
        first = sfirst;
    } else {
        first = p.sp.path_first;
    }
    if (threadIdx.x == 0) {
        block_counted = 0;
        mbar_init(&tile_bar[0], 1);
        mbar_init(&tile_bar[1], 1);
        fence_proxy_async();
    }
    __syncthreads();

    const uint64_t tid = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    Xoshiro g;
    g.s0 = p.rng[tid];
    g.s1 = p.rng[p.rng_stride + tid];
    g.s2 = p.rng[2 * p.rng_stride + tid];
    g.s3 = p.rng[3 * p.rng_stride + tid];
    const float eta_f = __double2float_rn(p.eta);
    const bool atomic_add = (p.flags & 5u) == 0;
    const bool st_mode = (p.flags & 4u) != 0;
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    uint64_t done = 0;
    float delta_max = 0.0f;

    // visits of this rank: v = visit_rank + k * visit_nranks; CTA c takes k = c, c + gridDim.x, ...
    auto tile_base_of = [&](uint64_t v) -> uint64_t {
        const uint64_t pass = v / p.n_tiles, i = v - pass * p.n_tiles;
        uint64_t t_idx = p.perm_mul[pass & 15] ? tile_perm(i, p.n_tiles, p.perm_add[pass & 15]) : (i + p.perm_add[pass & 15]) % p.n_tiles;
        if (p.tile_list) t_idx = p.tile_list[t_idx];  // peer mode: the k-th tile this rank owns
        return t_idx * (uint64_t) TILE_STEPS;
    };
    auto issue_tile = [&](uint64_t v, uint32_t buf) {   // one elected thread: TMA bulk copy of the tile's step records
        const uint64_t b0 = tile_base_of(v);
        const uint64_t n = b0 + TILE_STEPS <= p.sp.step_count ? (uint64_t) TILE_STEPS : p.sp.step_count - b0;
        const uint32_t bytes = (uint32_t) (n * sizeof(uint4));
        mbar_expect_tx(&tile_bar[buf], bytes);
        tma_load_1d(buf ? tile_buf1 : tile_buf0, p.steps + b0, bytes, &tile_bar[buf], pol_stream);
    };
    const uint64_t v0 = (uint64_t) p.visit_rank + (uint64_t) blockIdx.x * p.visit_nranks;
    const uint64_t v_stride = (uint64_t) gridDim.x * p.visit_nranks;
    if (TMA && threadIdx.x == 0 && v0 < p.n_visits) issue_tile(v0, 0);
    uint32_t it = 0;
    for (uint64_t v = v0; v < p.n_visits; v += v_stride, ++it) {
        const uint32_t buf = TMA ? (it & 1u) : 0u;
        uint4* const tile = buf ? tile_buf1 : tile_buf0;
        const uint64_t base = tile_base_of(v);
        const uint32_t terms = v + 1 == p.n_visits ? (uint32_t) p.last_visit_terms : (uint32_t) TILE_STEPS;
        // TMA: prefetch the next visit's tile into the other buffer (its readers finished at the barrier ending the last visit)
        if (TMA && threadIdx.x == 0 && v + v_stride < p.n_visits) {
            fence_proxy_async();
            issue_tile(v + v_stride, buf ^ 1u);
        }
        // path of the tile's first and last step (one search each per visit instead of one per term)
        const uint64_t last_step = (base + TILE_STEPS <= p.sp.step_count ? base + TILE_STEPS : p.sp.step_count) - 1;
        const uint32_t p_lo = find_path(first, p.sp.path_count, base);
        const bool tile_one_path = first[p_lo + 1] > last_step;
        const uint64_t tile_f = first[p_lo], tile_count = first[p_lo + 1] - tile_f;
        if (TMA) {
            mbar_wait(&tile_bar[buf], (it >> 1) & 1u);   // this visit's tile has landed
        } else {
            __syncthreads();  // the previous visit's readers are done with the tile
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const uint32_t j = r * 256 + threadIdx.x;
                if (base + j < p.sp.step_count) tile[j] = load_step(p.steps, base + j, pol_stream);
            }
            __syncthreads();
        }
#pragma unroll 1
        for (int r0 = 0; r0 < ROUNDS; r0 += BATCH) {
            Term t[BATCH];
            uint4 ra[BATCH], rb[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const uint32_t j = (r0 + b) * 256 + threadIdx.x;
                const uint64_t ia = base + j;
                t[b].valid = 0;
                if (j < terms && ia < p.sp.step_count) {
                    // the tile almost always lies inside one path: its (start, count) were resolved once per visit
                    uint64_t f = tile_f, count = tile_count;
                    if (!tile_one_path) {
                        const uint32_t pp = find_path(first, p.sp.path_count, ia);
                        f = first[pp];
                        count = first[pp + 1] - f;
                    }
                    t[b].ia = ia;
                    if (count > 1) {  // steps of 1-step paths are skipped, not counted (path_sgd_layout.cpp:190-192)
                        t[b].valid = 1;
                        t[b].flip_a = t[b].flip_b = 0;
                        draw_partner_fast<DIMS>(p.sp, g, f, count, ia - f, t[b]);
                        if (p.trace) {
                            const unsigned long long k = atomicAdd(p.trace_count, 1ull);
                            if (k < p.trace_cap) {
                                p.trace[2 * k] = ia;
                                p.trace[2 * k + 1] = t[b].ib | ((unsigned long long) t[b].flip_a << 62) | ((unsigned long long) t[b].flip_b << 63);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                if (t[b].valid) {
                    ra[b] = tile[t[b].ia - base];
                    const uint64_t off = t[b].ib - base;  // wraps to a huge value when ib < base
                    if (off < (uint64_t) TILE_STEPS) rb[b] = tile[off];
                    else rb[b] = load_step(p.steps, t[b].ib, pol_stream);
                }
            }
            if (DIMS == 2) {
                float2 ca[BATCH], cb[BATCH];
                float2* pa[BATCH];
                float2* pb[BATCH];
                float dij[BATCH];
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    if (t[b].valid) {
                        uint64_t pos_a = step_pos(ra[b]), pos_b = step_pos(rb[b]);
                        uint32_t end_a = ra[b].x & 1u, end_b = rb[b].x & 1u;
                        if (t[b].flip_a) { pos_a += ra[b].y; end_a ^= 1u; }
                        if (t[b].flip_b) { pos_b += rb[b].y; end_b ^= 1u; }
                        const uint64_t dpos = pos_a > pos_b ? pos_a - pos_b : pos_b - pos_a;
                        dij[b] = dpos ? __ull2float_rn(dpos) : 1e-9f;
                        pa[b] = coord2_ptr(p, ra[b].x >> 1, end_a);
                        pb[b] = coord2_ptr(p, rb[b].x >> 1, end_b);
                        ca[b] = ld_coord2(pa[b], pol_keep);
                        cb[b] = ld_coord2(pb[b], pol_keep);
                    }
                }
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    if (t[b].valid) {
                        float mu = __fdiv_rn(eta_f, dij[b]);
                        if (mu > 1.0f) mu = 1.0f;
                        float dx = __fsub_rn(ca[b].x, cb[b].x);
                        const float dy = __fsub_rn(ca[b].y, cb[b].y);
                        if (dx == 0.0f) dx = 1e-9f;
                        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                        const float Delta = __fmul_rn(__fmul_rn(mu, __fsub_rn(mag, dij[b])), 0.5f);
                        delta_max = fmaxf(delta_max, fabsf(Delta));
                        const float r = __fdiv_rn(Delta, mag);
                        const float r_x = __fmul_rn(r, dx);
                        const float r_y = __fmul_rn(r, dy);
                        if (atomic_add) {
                            red_coord2(pa[b], -r_x, -r_y, pol_keep);
                            red_coord2(pb[b], r_x, r_y, pol_keep);
                        } else {
                            const float2 na = make_float2(__fsub_rn(ca[b].x, r_x), __fsub_rn(ca[b].y, r_y));
                            st_coord(pa[b], na, st_mode);
                            const float2 bs = (pa[b] == pb[b]) ? na : cb[b];
                            st_coord(pb[b], make_float2(__fadd_rn(bs.x, r_x), __fadd_rn(bs.y, r_y)), st_mode);
                        }
                        ++done;
                    }
                }
            } else {
#pragma unroll
                for (int b = 0; b < BATCH; ++b) {
                    if (t[b].valid) {
                        const uint32_t na = ra[b].x >> 1, nb = rb[b].x >> 1;
                        uint32_t u = 3u;
                        if (p.frozen) {
                            if (p.frozen[na]) u &= ~1u;
                            if (p.frozen[nb]) u &= ~2u;
                        }
                        const double d = fabs(__dsub_rn(__ull2double_rn(step_pos(ra[b])), __ull2double_rn(step_pos(rb[b]))));
                        if (u == 0) { ++done; continue; }
                        if (d == 0.0) continue;
                        double* qa = coord1_ptr(p, na);
                        double* qb = coord1_ptr(p, nb);
                        const double xa = ld_coord1(qa, pol_keep), xb = ld_coord1(qb, pol_keep);
                        double mu = __dmul_rn(p.eta, __ddiv_rn(1.0, d));
                        if (mu > 1.0) mu = 1.0;
                        double dx = __dsub_rn(xa, xb);
                        if (dx == 0.0) dx = 1e-9;
                        const double mag = fabs(dx);
                        const double Delta = __dmul_rn(__dmul_rn(mu, __dsub_rn(mag, d)), 0.5);
                        delta_max = fmaxf(delta_max, (float) fabs(Delta));
                        const double r_x = __dmul_rn(__ddiv_rn(Delta, mag), dx);
                        if (atomic_add) {
                            if (u & 1u) red_coord1(qa, -r_x, pol_keep);
                            if (u & 2u) red_coord1(qb, r_x, pol_keep);
                        } else {
                            const double nav = __dsub_rn(xa, r_x);
                            if (u & 1u) st_coord(qa, nav, st_mode);
                            const double bs = (qa == qb && (u & 1u)) ? nav : xb;
                            if (u & 2u) st_coord(qb, __dadd_rn(bs, r_x), st_mode);
                        }
                        ++done;
                    }
                }
            }
        }
        if (TMA) __syncthreads();  // every reader is done with this buffer before the TMA engine may refill it
    }
    p.rng[tid] = g.s0;
    p.rng[p.rng_stride + tid] = g.s1;
    p.rng[2 * p.rng_stride + tid] = g.s2;
    p.rng[3 * p.rng_stride + tid] = g.s3;

    unsigned long long wsum = done;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&block_counted, wsum);
    if (p.delta_max_bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) delta_max = fmaxf(delta_max, __shfl_xor_sync(0xffffffffu, delta_max, o));
        if ((threadIdx.x & 31) == 0) atomicMax(p.delta_max_bits, __float_as_uint(delta_max));
    }
    __syncthreads();
    if (threadIdx.x == 0 && block_counted) atomicAdd(p.counted, block_counted);
}

__global__ void seed_streams_kernel(uint64_t* rng, uint64_t stride, uint64_t n, uint64_t seed_base) {
    const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    Xoshiro g;
    xoshiro_seed(g, seed_base + t);
    rng[t] = g.s0;
    rng[stride + t] = g.s1;
    rng[2 * stride + t] = g.s2;
    rng[3 * stride + t] = g.s3;
}

__global__ void pack_steps_kernel(StepRec* out, const uint32_t* step_node, const uint8_t* step_rev, const uint64_t* pos,
                                  const uint32_t* node_len, uint64_t n, uint64_t out_offset, uint32_t* depth) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t node = step_node[i];
    if (depth) atomicAdd(depth + node, 1u);
    StepRec r;
    r.handle = (node << 1) | (step_rev ? (uint32_t) (step_rev[i] != 0) : 0u);
    r.len = node_len[node];
    const uint64_t p = pos[i];
    r.pos_lo = (uint32_t) p;
    r.pos_hi = (uint32_t) (p >> 32);
    out[out_offset + i] = r;
}

// flatten-to-device, positions derived on the GPU: len[i] = node_len[step_node[i]] (+ id validation), one device-wide
// exclusive scan, then pos[i] = scan[i] - scan[first step of i's path] (== XP positions, xp.cpp:607-616)
__global__ void gather_len_kernel(uint64_t* len, const uint32_t* step_node, const uint32_t* node_len, uint64_t n, uint32_t n_nodes, int* bad) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t node = step_node[i];
    if (node >= n_nodes) { atomicExch(bad, 1); len[i] = 0; return; }
    len[i] = node_len[node];
}

__global__ void pack_steps_scan_kernel(StepRec* out, const uint32_t* step_node, const uint8_t* step_rev, const uint64_t* scan,
                                       const uint32_t* node_len, const uint64_t* first, uint32_t P, uint64_t n, uint32_t* depth) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t node = step_node[i];
    if (depth) atomicAdd(depth + node, 1u);
    const uint32_t p = find_path(first, P, i);
    const uint64_t pos = scan[i] - scan[first[p]];
    StepRec r;
    r.handle = (node << 1) | (step_rev ? (uint32_t) (step_rev[i] != 0) : 0u);
    r.len = node_len[node];
    r.pos_lo = (uint32_t) pos;
    r.pos_hi = (uint32_t) (pos >> 32);
    out[i] = r;
}

// How often does a path revisit a node within one tile of TILE_STEPS consecutive steps?  One CTA per tile inserts the
// tile's nodes into a shared-memory hash set and counts the repeated visits; the total over all tiles gives the share of
// steps that are repeats (tandem repeats, LPA: high; SNP-dense pangenome graphs: ~0), which the tile-sampling launch
// shape has to respect because the concurrent terms of a tile land on that tile's distinct nodes.
__global__ void tile_repeat_kernel(const uint32_t* step_node, uint64_t n, unsigned long long* total_dups) {
    constexpr uint32_t SLOTS = 2 * TILE_STEPS, EMPTY = 0xFFFFFFFFu;
    __shared__ uint32_t slot[SLOTS];
    __shared__ unsigned int dups;
    for (uint32_t i = threadIdx.x; i < SLOTS; i += blockDim.x) slot[i] = EMPTY;
    if (threadIdx.x == 0) dups = 0;
    __syncthreads();
    const uint64_t base = (uint64_t) blockIdx.x * TILE_STEPS;
    unsigned int mine = 0;
    for (uint32_t j = threadIdx.x; j < (uint32_t) TILE_STEPS; j += blockDim.x) {
        if (base + j >= n) break;
        const uint32_t node = step_node[base + j];
        uint32_t h = (node * 2654435761u) >> 20;
        for (;;) {
            h &= SLOTS - 1;
            const uint32_t old = atomicCAS(&slot[h], EMPTY, node);
            if (old == EMPTY) break;
            if (old == node) { ++mine; break; }
            ++h;
        }
    }
    if (mine) atomicAdd(&dups, mine);
    __syncthreads();
    if (threadIdx.x == 0 && dups) atomicAdd(total_dups, (unsigned long long) dups);
}

__global__ void xy_from_XY_kernel(float4* xy, const double* X, const double* Y, uint64_t n) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xy[i] = make_float4((float) X[2 * i], (float) Y[2 * i], (float) X[2 * i + 1], (float) Y[2 * i + 1]);
}

__global__ void XY_from_xy_kernel(double* X, double* Y, const float4* xy, uint64_t n) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = xy[i];
    X[2 * i] = v.x; Y[2 * i] = v.y; X[2 * i + 1] = v.z; Y[2 * i + 1] = v.w;
}

template <typename T>
__global__ void scale_kernel(T* a, uint64_t n, T s) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] *= s;
}
template <typename T, int SIGN>
__global__ void addsub_kernel(T* out, const T* a, const T* b, uint64_t n) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = SIGN > 0 ? a[i] + b[i] : a[i] - b[i];
}

// verification hook: one thread replays a stream with the SAME draw_term the SGD kernels use
template <int DIMS>
__global__ void sample_terms_kernel(SamplerParams sp, const StepRec* steps, uint64_t seed, uint64_t n_terms, SampleOut out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Xoshiro g;
    xoshiro_seed(g, seed);
    for (uint64_t k = 0; k < n_terms; ++k) {
        Term t;
        draw_term<DIMS>(sp, sp.path_first, g, t);
        const uint64_t f = sp.path_first[t.path];
        uint64_t pos_a = 0, pos_b = 0;
        uint32_t node_a = 0, node_b = 0, end_a = 0, end_b = 0;
        if (t.valid) {
            const uint4 ra = load_step(steps, t.ia), rb = load_step(steps, t.ib);
            pos_a = step_pos(ra); pos_b = step_pos(rb);
            node_a = ra.x >> 1; node_b = rb.x >> 1;
            end_a = ra.x & 1u; end_b = rb.x & 1u;
            if (DIMS == 2) {
                if (t.flip_a) { pos_a += ra.y; end_a ^= 1u; }
                if (t.flip_b) { pos_b += rb.y; end_b ^= 1u; }
            } else {
                end_a = end_b = 0;
            }
        }
        if (out.step_index) out.step_index[k] = t.step_index;
        if (out.path) out.path[k] = t.path;
        if (out.rank_a) out.rank_a[k] = t.ia - f;
        if (out.rank_b) out.rank_b[k] = t.ib - f;
        if (out.node_a) out.node_a[k] = node_a;
        if (out.node_b) out.node_b[k] = node_b;
        if (out.pos_a) out.pos_a[k] = pos_a;
        if (out.pos_b) out.pos_b[k] = pos_b;
        if (out.end_a) out.end_a[k] = (uint8_t) end_a;
        if (out.end_b) out.end_b[k] = (uint8_t) end_b;
        if (out.valid) out.valid[k] = (uint8_t) t.valid;
    }
}

// Sampled path stress of the resident coordinates (definition: SURVEY.md §8d, oracle/pgsgd_oracle.c orc_path_stress_*):
// STRESS_STREAMS generators (stream t seeded seed + t) draw `per` pairs each — step uniform over all steps, partner
// uniform in the same path, ends uniform (2D) — and accumulate ((|p_a - p_b| - d) / d)^2 in fp64 with IEEE operations, so
// the per-stream sums are bit-identical to the oracle's; the host adds them in stream order.
template <int DIMS, bool LOCAL>
__global__ void stress_kernel(const uint64_t* first, uint32_t P, uint64_t S, const StepRec* steps, const float* xy, const double* x1d,
                              uint64_t per, uint64_t seed, double* acc_out, unsigned long long* used_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= STRESS_STREAMS) return;
    Xoshiro g;
    xoshiro_seed(g, seed + t);
    double acc = 0.0;
    unsigned long long used = 0;
    for (uint64_t k = 0; k < per; ++k) {
        const uint64_t ia = draw_uniform(g, S);
        const uint32_t p = find_path(first, P, ia);
        const uint64_t f = first[p], cnt = first[p + 1] - f;
        uint64_t ib;
        bool skip = false;
        if (LOCAL) {   // orc_local_stress_*: partner 1..64 ranks away along the path, pairs beyond 1000 bp skipped
            const uint64_t j = 1 + draw_uniform(g, 64);
            const uint32_t back = draw_flip(g);
            const uint64_t ra = ia - f;
            skip = back ? ra < j : ra + j >= cnt;
            ib = skip ? ia : (back ? ia - j : ia + j);
        } else {
            ib = f + draw_uniform(g, cnt);
        }
        uint32_t fa = 0, fb = 0;
        if (DIMS == 2) { fa = draw_flip(g); fb = draw_flip(g); }   // drawn before any skip, as the oracle does
        if (skip) continue;
        const uint4 ra = load_step(steps, ia), rb = load_step(steps, ib);
        uint64_t pa = step_pos(ra), pb = step_pos(rb);
        if (DIMS == 2) {
            uint32_t ea = ra.x & 1u, eb = rb.x & 1u;
            if (fa) { pa += ra.y; ea ^= 1u; }
            if (fb) { pb += rb.y; eb ^= 1u; }
            if (pa == pb) continue;
            const double d = fabs(__dsub_rn(__ull2double_rn(pa), __ull2double_rn(pb)));
            if (LOCAL && d > 1000.0) continue;
            const float2 ca = __ldcg(reinterpret_cast<const float2*>(xy) + ((uint64_t) (ra.x >> 1) * 2 + ea));
            const float2 cb = __ldcg(reinterpret_cast<const float2*>(xy) + ((uint64_t) (rb.x >> 1) * 2 + eb));
            const double dx = __dsub_rn((double) ca.x, (double) cb.x), dy = __dsub_rn((double) ca.y, (double) cb.y);
            const double e = __ddiv_rn(__dsub_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), d), d);
            acc = __dadd_rn(acc, __dmul_rn(e, e));
        } else {
            if (pa == pb) continue;
            const double d = fabs(__dsub_rn(__ull2double_rn(pa), __ull2double_rn(pb)));
            if (LOCAL && d > 1000.0) continue;
            const double xa = __ldcg(x1d + (ra.x >> 1)), xb = __ldcg(x1d + (rb.x >> 1));
            const double e = __ddiv_rn(__dsub_rn(fabs(__dsub_rn(xa, xb)), d), d);
            acc = __dadd_rn(acc, __dmul_rn(e, e));
        }
        ++used;
    }
    acc_out[t] = acc;
    used_out[t] = used;
}

template <int DIMS, int BATCH>
cudaError_t launch_iter_t(const IterParams& p, const LaunchShape& s, cudaStream_t stream) {
    if (p.smem_paths) {
        auto k = pgsgd_iter_kernel<DIMS, BATCH, true>;
        if (s.smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s.smem);
            if (e != cudaSuccess) return e;
        }
        k<<<s.grid, s.block, s.smem, stream>>>(p);
    } else {
        pgsgd_iter_kernel<DIMS, BATCH, false><<<s.grid, s.block, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

template <int DIMS, int BATCH>
cudaError_t occupancy_t(int block, size_t smem, bool smem_paths, int* out) {
    if (smem_paths) {
        auto k = pgsgd_iter_kernel<DIMS, BATCH, true>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k, block, smem);
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, pgsgd_iter_kernel<DIMS, BATCH, false>, block, 0);
}

inline unsigned grid_for(uint64_t n, int block) { return (unsigned) ((n + block - 1) / block); }

}  // namespace

template <int DIMS, int BATCH, bool SP, bool TMA>
cudaError_t tile_launch_one(const IterParams& p, const LaunchShape& s, cudaStream_t stream) {
    auto k = pgsgd_tile_kernel<DIMS, BATCH, SP, TMA>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s.smem);
    if (e != cudaSuccess) return e;
    k<<<s.grid, s.block, s.smem, stream>>>(p);
    return cudaGetLastError();
}
template <int DIMS, int BATCH, bool SP, bool TMA>
cudaError_t tile_occ_one(size_t smem, int* out) {
    auto k = pgsgd_tile_kernel<DIMS, BATCH, SP, TMA>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k, 256, smem);
}

template <int DIMS, int BATCH>
cudaError_t tile_launch_pick(bool sp, bool tma, const IterParams& p, const LaunchShape& shape, cudaStream_t stream) {
    if (tma) return sp ? tile_launch_one<DIMS, BATCH, true, true>(p, shape, stream) : tile_launch_one<DIMS, BATCH, false, true>(p, shape, stream);
    return sp ? tile_launch_one<DIMS, BATCH, true, false>(p, shape, stream) : tile_launch_one<DIMS, BATCH, false, false>(p, shape, stream);
}
template <int DIMS, int BATCH>
cudaError_t tile_occ_pick(bool sp, bool tma, size_t smem, int* out) {
    if (tma) return sp ? tile_occ_one<DIMS, BATCH, true, true>(smem, out) : tile_occ_one<DIMS, BATCH, false, true>(smem, out);
    return sp ? tile_occ_one<DIMS, BATCH, true, false>(smem, out) : tile_occ_one<DIMS, BATCH, false, false>(smem, out);
}

cudaError_t launch_tile_iteration(int dims, int batch, const IterParams& p, const LaunchShape& shape, cudaStream_t stream) {
    if (shape.block != 256) return cudaErrorInvalidValue;
    const bool sp = p.smem_paths != 0, tma = (p.flags & 8u) != 0;
    if (dims == 2 && batch == 4) return tile_launch_pick<2, 4>(sp, tma, p, shape, stream);
    if (dims == 2 && batch == 2) return tile_launch_pick<2, 2>(sp, tma, p, shape, stream);
    if (dims == 2 && batch == 1) return tile_launch_pick<2, 1>(sp, tma, p, shape, stream);
    if (dims == 1 && batch == 1) return tile_launch_pick<1, 1>(sp, tma, p, shape, stream);
    if (dims == 1 && batch == 4) return tile_launch_pick<1, 4>(sp, tma, p, shape, stream);
    if (dims == 1 && batch == 2) return tile_launch_pick<1, 2>(sp, tma, p, shape, stream);
    return cudaErrorInvalidValue;
}

cudaError_t tile_occupancy(int dims, int batch, size_t smem, bool sp, bool tma, int* out) {
    if (dims == 2 && batch == 4) return tile_occ_pick<2, 4>(sp, tma, smem, out);
    if (dims == 2 && batch == 2) return tile_occ_pick<2, 2>(sp, tma, smem, out);
    if (dims == 2 && batch == 1) return tile_occ_pick<2, 1>(sp, tma, smem, out);
    if (dims == 1 && batch == 1) return tile_occ_pick<1, 1>(sp, tma, smem, out);
    if (dims == 1 && batch == 4) return tile_occ_pick<1, 4>(sp, tma, smem, out);
    if (dims == 1 && batch == 2) return tile_occ_pick<1, 2>(sp, tma, smem, out);
    return cudaErrorInvalidValue;
}

cudaError_t launch_seed_streams(uint64_t* rng, uint64_t rng_stride, uint64_t n, uint64_t seed_base, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    seed_streams_kernel<<<grid_for(n, 256), 256, 0, stream>>>(rng, rng_stride, n, seed_base);
    return cudaGetLastError();
}

cudaError_t launch_iteration(int dims, int batch, const IterParams& p, const LaunchShape& shape, cudaStream_t stream) {
    if (dims == 2) {
        switch (batch) {
            case 1: return launch_iter_t<2, 1>(p, shape, stream);
            case 2: return launch_iter_t<2, 2>(p, shape, stream);
            case 4: return launch_iter_t<2, 4>(p, shape, stream);
        }
    } else if (dims == 1) {
        switch (batch) {
            case 1: return launch_iter_t<1, 1>(p, shape, stream);
            case 2: return launch_iter_t<1, 2>(p, shape, stream);
            case 4: return launch_iter_t<1, 4>(p, shape, stream);
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t iteration_occupancy(int dims, int batch, int block, size_t smem, bool smem_paths, int* blocks_per_sm) {
    if (dims == 2) {
        switch (batch) {
            case 1: return occupancy_t<2, 1>(block, smem, smem_paths, blocks_per_sm);
            case 2: return occupancy_t<2, 2>(block, smem, smem_paths, blocks_per_sm);
            case 4: return occupancy_t<2, 4>(block, smem, smem_paths, blocks_per_sm);
        }
    } else if (dims == 1) {
        switch (batch) {
            case 1: return occupancy_t<1, 1>(block, smem, smem_paths, blocks_per_sm);
            case 2: return occupancy_t<1, 2>(block, smem, smem_paths, blocks_per_sm);
            case 4: return occupancy_t<1, 4>(block, smem, smem_paths, blocks_per_sm);
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_pack_steps(StepRec* out, const uint32_t* step_node, const uint8_t* step_rev, const uint64_t* step_pos,
                              const uint32_t* node_len, uint64_t n, uint64_t out_offset, uint32_t* depth, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    pack_steps_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, step_node, step_rev, step_pos, node_len, n, out_offset, depth);
    return cudaGetLastError();
}

cudaError_t launch_flatten_on_device(StepRec* out, const uint32_t* step_node, const uint8_t* step_rev, const uint32_t* node_len,
                                     const uint64_t* first, uint32_t P, uint32_t n_nodes, uint64_t n, uint64_t* scratch_len, int* bad,
                                     uint32_t* depth, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    gather_len_kernel<<<grid_for(n, 256), 256, 0, stream>>>(scratch_len, step_node, node_len, n, n_nodes, bad);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, scratch_len, scratch_len, n, stream);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, scratch_len, scratch_len, n, stream);
    if (e == cudaSuccess) {
        int h_bad = 0;  // the packing kernel indexes node tables: only run it on validated ids
        e = cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess && !h_bad) pack_steps_scan_kernel<<<grid_for(n, 256), 256, 0, stream>>>(out, step_node, step_rev, scratch_len, node_len, first, P, n, depth);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp);
    return e;
}

__global__ void iota_kernel(uint64_t* v, uint64_t n) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void gather_u32_kernel(uint32_t* out, const uint32_t* table, const uint64_t* idx, uint64_t n) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = table[idx[i]];
}

// 1D node order (path_linear_sgd_order's sort, path_sgd.cpp:650-658): node ranks sorted by (weak component, position,
// handle).  A STABLE radix sort of (x, rank) pairs that start in rank order gives (position, handle); a second stable sort
// of the result by the component key (when the caller supplies one; device pointer, [n]) makes the component the major key.
cudaError_t launch_order_1d(const double* x, const uint32_t* d_component, uint64_t* order_out, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    double* keys_out = nullptr;
    uint64_t* vals_in = nullptr;
    uint64_t* vals_mid = nullptr;
    uint32_t *ck_in = nullptr, *ck_out = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0, tmp2 = 0;
    uint64_t* first_out = d_component ? nullptr : order_out;
    cudaError_t e = cudaMalloc(&keys_out, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&vals_in, n * sizeof(uint64_t));
    if (e == cudaSuccess && d_component) {
        e = cudaMalloc(&vals_mid, n * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaMalloc(&ck_in, n * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&ck_out, n * sizeof(uint32_t));
        first_out = vals_mid;
    }
    if (e == cudaSuccess) { iota_kernel<<<grid_for(n, 256), 256, 0, stream>>>(vals_in, n); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, x, keys_out, vals_in, first_out, n, 0, 64, stream);
    if (e == cudaSuccess && d_component) {
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp2, ck_in, ck_out, vals_mid, order_out, n, 0, 32, stream);
        if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    }
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, x, keys_out, vals_in, first_out, n, 0, 64, stream);
    if (e == cudaSuccess && d_component) {
        gather_u32_kernel<<<grid_for(n, 256), 256, 0, stream>>>(ck_in, d_component, vals_mid, n);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ck_in, ck_out, vals_mid, order_out, n, 0, 32, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp); cudaFree(vals_in); cudaFree(keys_out); cudaFree(vals_mid); cudaFree(ck_in); cudaFree(ck_out);
    return e;
}

cudaError_t launch_tile_repeats(const uint32_t* step_node, uint64_t n, unsigned long long* total_dups, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    const uint64_t tiles = (n + TILE_STEPS - 1) / TILE_STEPS;
    tile_repeat_kernel<<<(unsigned) tiles, 256, 0, stream>>>(step_node, n, total_dups);
    return cudaGetLastError();
}

cudaError_t launch_xy_from_XY(float* xy, const double* X, const double* Y, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    xy_from_XY_kernel<<<grid_for(n, 256), 256, 0, stream>>>(reinterpret_cast<float4*>(xy), X, Y, n);
    return cudaGetLastError();
}
cudaError_t launch_XY_from_xy(double* X, double* Y, const float* xy, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    XY_from_xy_kernel<<<grid_for(n, 256), 256, 0, stream>>>(X, Y, reinterpret_cast<const float4*>(xy), n);
    return cudaGetLastError();
}
cudaError_t launch_scale_f32(float* a, uint64_t n, float s, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    scale_kernel<float><<<grid_for(n, 256), 256, 0, stream>>>(a, n, s);
    return cudaGetLastError();
}
cudaError_t launch_scale_f64(double* a, uint64_t n, double s, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    scale_kernel<double><<<grid_for(n, 256), 256, 0, stream>>>(a, n, s);
    return cudaGetLastError();
}
cudaError_t launch_sub_f32(float* out, const float* a, const float* b, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    addsub_kernel<float, -1><<<grid_for(n, 256), 256, 0, stream>>>(out, a, b, n);
    return cudaGetLastError();
}
cudaError_t launch_add_f32(float* out, const float* a, const float* b, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    addsub_kernel<float, 1><<<grid_for(n, 256), 256, 0, stream>>>(out, a, b, n);
    return cudaGetLastError();
}
cudaError_t launch_sub_f64(double* out, const double* a, const double* b, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    addsub_kernel<double, -1><<<grid_for(n, 256), 256, 0, stream>>>(out, a, b, n);
    return cudaGetLastError();
}
cudaError_t launch_add_f64(double* out, const double* a, const double* b, uint64_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    addsub_kernel<double, 1><<<grid_for(n, 256), 256, 0, stream>>>(out, a, b, n);
    return cudaGetLastError();
}

cudaError_t launch_stress(int dims, int local, const uint64_t* first, uint32_t P, uint64_t S, const StepRec* steps, const float* xy,
                          const double* x1d, uint64_t per, uint64_t seed, double* acc_out, unsigned long long* used_out, cudaStream_t stream) {
    const dim3 grid(STRESS_STREAMS / 128), block(128);
    if (dims == 2 && !local) stress_kernel<2, false><<<grid, block, 0, stream>>>(first, P, S, steps, xy, x1d, per, seed, acc_out, used_out);
    else if (dims == 2) stress_kernel<2, true><<<grid, block, 0, stream>>>(first, P, S, steps, xy, x1d, per, seed, acc_out, used_out);
    else if (dims == 1 && !local) stress_kernel<1, false><<<grid, block, 0, stream>>>(first, P, S, steps, xy, x1d, per, seed, acc_out, used_out);
    else if (dims == 1) stress_kernel<1, true><<<grid, block, 0, stream>>>(first, P, S, steps, xy, x1d, per, seed, acc_out, used_out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_sample_terms(int dims, const SamplerParams& sp, const StepRec* steps, uint64_t seed, uint64_t n_terms,
                                const SampleOut& out, cudaStream_t stream) {
    if (dims == 2) sample_terms_kernel<2><<<1, 32, 0, stream>>>(sp, steps, seed, n_terms, out);
    else if (dims == 1) sample_terms_kernel<1><<<1, 32, 0, stream>>>(sp, steps, seed, n_terms, out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------------
// Sorting-goodness readout on the device: `odgi stats -l [-g] -s [-d]` of the graph sorted by `order`
// (src/subcommand/stats_main.cpp:399-800, 1D branch): every consecutive step pair of every path is a link; all sums are
// 64-bit integers, so the result equals the oracle's (oracle.sort_goodness) exactly.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void scatter_rank_kernel(uint32_t* new_rank, const uint64_t* order, uint64_t n) {
    const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) new_rank[order[k]] = (uint32_t) k;
}
__global__ void gather_len_by_order_kernel(uint64_t* out, const uint32_t* node_len, const uint64_t* order, uint64_t n) {
    const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = node_len[order[k]];
    if (k == n) out[k] = 0;
}
__global__ void step_rank_kernel(uint32_t* out, const StepRec* steps, const uint32_t* new_rank, uint64_t n) {
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = new_rank[steps[i].handle >> 1];
}

// acc: [0] mll node, [1] mll nt, [2] links, [3] gap links, [4] spd node, [5] spd nt, [6] nucleotides, [7] penalties, [8] diff orientation
__global__ void goodness_kernel(const StepRec* steps, const uint64_t* first, uint32_t P, uint64_t S, const uint32_t* new_rank,
                                const uint64_t* pm, const uint32_t* sorted_rank, uint32_t flags, unsigned long long* acc) {
    unsigned long long a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint64_t s = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (uint64_t) gridDim.x * blockDim.x) {
        const uint4 h = load_step(steps, s);
        a[6] += h.y;
        const uint32_t p = find_path(first, P, s);
        const uint64_t lo = first[p], hi = first[p + 1];
        if (s + 1 == hi) {            // end of path: "so the best metric equals 1" (stats_main.cpp:733-735)
            a[4] += 1;
            a[5] += h.y;
            continue;
        }
        const uint4 i = load_step(steps, s + 1);
        const uint32_t uh = new_rank[h.x >> 1], ui = new_rank[i.x >> 1], bh = h.x & 1u, bi = i.x & 1u;
        a[2] += 1;
        bool gap = false;
        if (flags & 1u) {             // -g: the link to the next node of the path's own ordered node set is not penalised (:479-511)
            uint64_t l = lo, r = hi;  // first element > uh in the path's sorted ranks
            while (l < r) {
                const uint64_t m = (l + r) >> 1;
                if (sorted_rank[m] <= uh) l = m + 1; else r = m;
            }
            gap = l < hi && sorted_rank[l] == ui;
        }
        if (gap) {
            a[3] += 1;
        } else {
            uint32_t ia = uh + (1u - bh), ib = ui + bi;
            if (ib < ia) { const uint32_t t = ia; ia = ib; ib = t; }
            a[0] += ib - ia;
            a[1] += pm[ib] - pm[ia];
        }
        uint32_t x = uh, y = ui;
        unsigned long long w = 1;
        if (y < x) { x = ui; y = uh; w = 3; a[7] += 1; }
        const unsigned long long dn = y - x, dt = pm[y] - pm[x];
        a[4] += w * dn;
        a[5] += w * dt;
        if (bh != bi) {
            a[8] += 1;
            if (flags & 2u) { a[4] += 2 * dn; a[5] += 2 * dt; }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        unsigned long long v = a[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(acc + k, v);
    }
}

}  // namespace

cudaError_t launch_goodness(const StepRec* steps, const uint64_t* first, const uint64_t* h_first, uint32_t P, uint64_t S, uint64_t N,
                            const uint32_t* d_node_len, const uint64_t* d_order, uint32_t flags, unsigned long long* h_acc9, cudaStream_t stream) {
    uint32_t *new_rank = nullptr, *sr_in = nullptr, *sr_out = nullptr;
    uint64_t* pm = nullptr;
    unsigned long long* acc = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0, tmp2 = 0;
    uint64_t* own_order = nullptr;
    cudaError_t e = cudaMalloc(&new_rank, (N ? N : 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&pm, (N + 1) * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&acc, 9 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(acc, 0, 9 * sizeof(unsigned long long), stream);
    if (e == cudaSuccess && !d_order) {
        e = cudaMalloc(&own_order, (N ? N : 1) * sizeof(uint64_t));
        if (e == cudaSuccess) { iota_kernel<<<grid_for(N, 256), 256, 0, stream>>>(own_order, N); e = cudaGetLastError(); }
        d_order = own_order;
    }
    if (e == cudaSuccess && N) {
        scatter_rank_kernel<<<grid_for(N, 256), 256, 0, stream>>>(new_rank, d_order, N);
        gather_len_by_order_kernel<<<grid_for(N + 1, 256), 256, 0, stream>>>(pm, d_node_len, d_order, N);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pm, pm, N + 1, stream);
    if ((flags & 1u) && e == cudaSuccess && S) {
        e = cudaMalloc(&sr_in, S * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&sr_out, S * sizeof(uint32_t));
        if (e == cudaSuccess) e = cub::DeviceSegmentedSort::SortKeys(tmp, tmp2, sr_in, sr_out, (int64_t) S, (int64_t) P, first, first + 1, stream);
        if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
    }
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pm, pm, N + 1, stream);
    if ((flags & 1u) && e == cudaSuccess && S) {
        step_rank_kernel<<<grid_for(S, 256), 256, 0, stream>>>(sr_in, steps, new_rank, S);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cub::DeviceSegmentedSort::SortKeys(tmp, tmp_bytes, sr_in, sr_out, (int64_t) S, (int64_t) P, first, first + 1, stream);
    }
    if (e == cudaSuccess && S) {
        const unsigned grid = (unsigned) (S / 256 + 1 < 148 * 16 ? S / 256 + 1 : 148 * 16);
        goodness_kernel<<<grid, 256, 0, stream>>>(steps, first, P, S, new_rank, pm, sr_out, flags, acc);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_acc9, acc, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(tmp); cudaFree(new_rank); cudaFree(pm); cudaFree(acc); cudaFree(sr_in); cudaFree(sr_out); cudaFree(own_order);
    (void) h_first;
    return e;
}

}  // namespace pgsgd
