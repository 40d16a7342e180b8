// pgsgd_tile2.cu — the pipelined tile-sampling iteration kernel (round 2; the bench default).
//
// Replaces on the GPU: the 2D / 1D worker lambdas (src/algorithms/path_sgd_layout.cpp:165-377,
// src/algorithms/path_sgd.cpp:205-406) and cuda::gpu_layout_kernel (src/cuda/layout.cu:190-287).
//
// Same sampling scheme as pgsgd_tile_kernel (pgsgd_kernels.cu): a CTA stages TILE consecutive 16-byte step records in
// shared memory, every staged step is the FIRST step of exactly one term per visit, the partner is drawn by the
// reference's rule (coin, direction, dirty Zipf | uniform in path) and served from the tile when it falls inside it.
// What round 1's ncu capture showed (profiles/r01_ncu_kernels.md): that kernel is neither bandwidth- nor sector-bound but
// ISSUE- and LATENCY-bound — 443 warp instructions per warp-term and ONE dependent chain per thread
// (tile -> far step record -> far coordinate -> red), i.e. one long-latency load in flight per thread.  This kernel
//   * software-pipelines three consecutive terms of a thread: stage A(i) draws the partner and sends the far step
//     record on its way (cp.async 16 B into a per-thread shared-memory slot), stage B(i-1) turns the landed record into
//     the integer path distance and sends both coordinate loads on their way (cp.async 16 B = the node's float4),
//     stage C(i-2) computes the update and issues the two reds.  Two dependent HBM/L2 latencies per term are overlapped
//     with the neighbouring terms' work instead of being waited for; no register is held while a load is in flight;
//   * is written branch-free in 32-bit arithmetic with per-launch constants folded on the host (fp32 {zeta_n,
//     1/(1-zeta_2/zeta_n)} table, reciprocal of the quantisation step, a^100 by a fixed multiplication chain), approximate
//     (MUFU) division / rsqrt in the update — the tile kernel's contract is the LAW of the sampler and the stress band,
//     not bit-equality with a CPU thread (that contract belongs to pgsgd_iter_kernel);
//   * stages tiles either with coalesced LDG.128 + STS.128 into one buffer or (PGSGD_FLAG_TMA_STAGING) with TMA bulk
//     copies (cp.async.bulk + mbarrier, SASS UBLKCP) into two buffers, the next visit's tile in flight while the current
//     one is worked on.
// Peer phases of multi-GPU runs address the owner's coordinate slice through NVLink with the same pipeline (a remote
// round trip is just a longer latency to overlap).  Not here (legacy kernel): paths with >= 2^31 steps.
#include "pgsgd_kernels.cuh"

namespace pgsgd {

namespace {

__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(smem_dst), "l"(gsrc), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ double lds64d(uint32_t addr) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
    return r;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// dirtyzipf::fast_precise_pow restated on floats (see fast_precise_pow_f32 in pgsgd_device.cuh): fractional exponent part
__device__ __forceinline__ float pow_hack_frac(float a, float frac) {
    const int hi = (__float_as_int(a) >> 3) + 0x38000000;                  // high word of (double) a
    const int t = 1072632447 + __float2int_rd(frac * (float) (hi - 1072632447));
    return __int_as_float((t - 0x38000000) << 3);
}
// a^e for the integer part of the Zipf exponent alpha = 1 / (1 - theta) (99 for the reference's theta = 0.99: 1 - 0.99 is
// not exact in binary).  e < 128: seven squarings with the multiplies selected by the (launch-uniform) bits of e, unrolled —
// no loop, no divergence; larger exponents take the generic loop.
__device__ __forceinline__ float pow_int(float a, int e) {
    if (e < 128) {
        float r = (e & 1) ? a : 1.0f;
#pragma unroll
        for (int k = 1; k < 7; ++k) {
            a *= a;
            if ((e >> k) & 1) r *= a;
        }
        return r;
    }
    float r = 1.0f;
    while (e) {
        if (e & 1) r *= a;
        a *= a;
        e >>= 1;
    }
    return r;
}

// owner partition of a node (peer phases; n_parts <= 8: a branch-free range count)
__device__ __forceinline__ uint32_t part_of(const Tile2Params& p, uint32_t node) {
    uint32_t q = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) q += (k < (int) p.n_parts && node >= p.part_lo[k]) ? 1u : 0u;
    return q;
}
__device__ __forceinline__ float* xy_base(const Tile2Params& p, uint32_t node) {
    return p.n_parts <= 1 ? p.xy : p.part_xy[part_of(p, node)];
}
__device__ __forceinline__ double* x1d_base(const Tile2Params& p, uint32_t node) {
    return p.n_parts <= 1 ? p.x1d : p.part_x1d[part_of(p, node)];
}

struct VisitInfo {
    unsigned long long base;    // global index of the tile's first step
    unsigned long long f;       // first step of the path the tile starts in
    uint32_t count;             // steps of that path
    uint32_t one_path;          // the whole tile lies inside that path
    uint32_t terms;             // first steps used in this visit (TILE, less for the last visit)
    uint32_t n_in_tile;         // steps present in the tile (TILE, less at the end of the step array)
    uint32_t sweeps;            // terms per staged step in this visit (p.sweeps in full passes, the rest in the last ones)
    uint32_t pad;
};

template <typename FirstPtr>
__device__ __forceinline__ void make_visit(const Tile2Params& p, FirstPtr first, uint64_t v, uint32_t tile_steps, VisitInfo* out) {
    // passes 0 .. full_passes-1: every tile once, p.sweeps terms per staged step; then (if any) one pass with the left-over
    // sweeps; then the remainder of U mod S as single-sweep visits over the first tiles of one more pass
    const uint64_t pass = v / p.n_tiles, i = v - pass * p.n_tiles;
    out->sweeps = pass < p.full_passes ? p.sweeps : ((pass == p.full_passes && p.rem_sweeps) ? p.rem_sweeps : 1u);
    out->pad = 0;
    uint64_t t_idx = p.perm_mul[pass & 15] ? tile_perm(i, p.n_tiles, p.perm_add[pass & 15]) : (i + p.perm_add[pass & 15]) % p.n_tiles;
    if (p.flags & 8192u) {   // experiments: tiles drawn WITH replacement (a hash of the visit number) instead of a bijection per pass
        uint64_t z = (v + 1) * 0x9e3779b97f4a7c15ULL + p.perm_add[pass & 15];
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; z ^= z >> 31;
        t_idx = __umul64hi(z, p.n_tiles);
    }
    if (p.tile_list) t_idx = p.tile_list[t_idx];   // peer phases: the k-th tile this rank owns
    const uint64_t base = t_idx * (uint64_t) tile_steps;
    const uint64_t end = base + tile_steps <= p.step_count ? base + tile_steps : p.step_count;
    const uint32_t pl = find_path(first, p.path_count, base);
    out->base = base;
    out->f = first[pl];
    out->count = (uint32_t) (first[pl + 1] - first[pl]);
    out->one_path = first[pl + 1] >= end ? 1u : 0u;
    out->terms = v + 1 == p.n_visits ? (uint32_t) p.last_visit_terms : tile_steps;
    out->n_in_tile = (uint32_t) (end - base);
}

// ---------------------------------------------------------------------------------------------------------------------
template <int DIMS, bool TMA, int TILE>
__global__ void __launch_bounds__(256, 4) pgsgd_tile2_kernel(const __grid_constant__ Tile2Params p) {
    constexpr int ROUNDS = TILE / 256;
    constexpr int NBUF = TMA ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ VisitInfo vinfo[2];
    __shared__ __align__(8) uint64_t tile_bar[2];
    __shared__ unsigned long long block_counted;

    uint4* const tile0 = reinterpret_cast<uint4*>(smem_raw);
    const uint32_t tile0_s = smem_u32(tile0);
    const uint32_t slots_s = tile0_s + NBUF * TILE * 16;                 // [3][256] 16-byte per-thread slots
    const uint32_t slot_rb = slots_s + threadIdx.x * 16;
    const uint32_t slot_ca = slot_rb + 256 * 16;
    const uint32_t slot_cb = slot_ca + 256 * 16;
    const uint64_t* first;
    if (p.smem_paths) {
        uint64_t* sfirst = reinterpret_cast<uint64_t*>(smem_raw + NBUF * TILE * 16 + 3 * 256 * 16);
        for (uint32_t i = threadIdx.x; i <= p.path_count; i += 256) sfirst[i] = p.path_first[i];
        first = sfirst;
    } else {
        first = p.path_first;
    }
    if (threadIdx.x == 0) {
        block_counted = 0;
        if (TMA) {
            mbar_init(&tile_bar[0], 1);
            mbar_init(&tile_bar[1], 1);
            fence_proxy_async();
        }
    }
    __syncthreads();

    const uint64_t tid = (uint64_t) blockIdx.x * 256 + threadIdx.x;
    Xoshiro g;
    g.s0 = p.rng[tid];
    g.s1 = p.rng[p.rng_stride + tid];
    g.s2 = p.rng[2 * p.rng_stride + tid];
    g.s3 = p.rng[3 * p.rng_stride + tid];
    // step records: evict_first (a stream with no reuse) — except in sweep order, where the resident CTAs share one window of
    // the step array and a record fetched by one CTA is the far partner of its neighbours: normal priority keeps it around
    const uint64_t pol_stream = (p.flags & 256u) ? l2_policy_evict_normal() : l2_policy_evict_first();
    const uint64_t pol_keep = (p.flags & 512u) ? l2_policy_evict_normal() : l2_policy_evict_last();
    // experiments (PGSGD_FLAG_COORD_LD_FIRST / _NORMAL): the policy of the coordinate LOADS alone; the reds keep evict_last.
    // A load of a line homed on the other die leaves a copy in the local L2 partition: if those copies are what pushes
    // coordinate lines out (0.3-0.4 DRAM sectors per update on c4), a short-lived policy for them should show.
    const uint64_t pol_cld = (p.flags & 1024u) ? l2_policy_evict_first() : ((p.flags & 2048u) ? l2_policy_evict_normal() : pol_keep);
    const bool atomic_add = (p.flags & 5u) == 0;
    const bool st_mode = (p.flags & 4u) != 0;
    uint32_t done = 0;
    float delta_max = 0.0f;

    const uint64_t v0 = (uint64_t) p.visit_rank + (uint64_t) blockIdx.x * p.visit_nranks;
    const uint64_t v_stride = (uint64_t) gridDim.x * p.visit_nranks;
    if (threadIdx.x == 0 && v0 < p.n_visits) {
        make_visit(p, first, v0, TILE, &vinfo[0]);
        if (TMA) {
            const uint32_t bytes = vinfo[0].n_in_tile * 16u;
            mbar_expect_tx(&tile_bar[0], bytes);
            tma_load_1d(tile0, p.steps + vinfo[0].base, bytes, &tile_bar[0], pol_stream);
        }
    }

    uint32_t it = 0;
    for (uint64_t v = v0; v < p.n_visits; v += v_stride, ++it) {
        __syncthreads();   // vinfo[it & 1] is visible; every reader of the buffer about to be refilled is done
        const VisitInfo vi = vinfo[it & 1];
        const uint32_t buf = TMA ? (it & 1u) : 0u;
        const uint32_t tile_s = tile0_s + buf * (TILE * 16);
        if (threadIdx.x == 0 && v + v_stride < p.n_visits) {
            // the NEXT visit is resolved (two 64-bit modulos + a path search) while this one is worked on
            make_visit(p, first, v + v_stride, TILE, &vinfo[(it + 1) & 1]);
            if (TMA) {   // ... and its tile is put in flight into the other buffer (its readers passed the barrier above)
                const VisitInfo& nx = vinfo[(it + 1) & 1];
                const uint32_t bytes = nx.n_in_tile * 16u;
                fence_proxy_async();
                mbar_expect_tx(&tile_bar[buf ^ 1u], bytes);
                tma_load_1d(tile0 + (buf ^ 1u) * TILE, p.steps + nx.base, bytes, &tile_bar[buf ^ 1u], pol_stream);
            }
        }
        if (TMA) {
            mbar_wait(&tile_bar[buf], (it >> 1) & 1u);
        } else {
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const uint32_t j = r * 256 + threadIdx.x;
                if (j < vi.n_in_tile) {
                    const uint4 rec = load_step(p.steps, vi.base + j, pol_stream);
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(tile_s + j * 16), "r"(rec.x), "r"(rec.y), "r"(rec.z), "r"(rec.w) : "memory");
                }
            }
            __syncthreads();
        }
        const uint32_t rank0 = (uint32_t) (vi.base - vi.f);   // path rank of the tile's first step (one-path tiles)

        // ---- the three-stage pipeline over the ROUNDS terms of this thread ----
        // carried A -> B
        uint32_t b_rb = 0;        // shared-memory address of the partner's record (tile or own slot) | flip_b | valid << 1
        uint32_t b_ia = 0;        // 2D: float2 index of the first end (node * 2 + end); 1D: node
        uint32_t b_pa_lo = 0, b_pa_hi = 0;   // end-adjusted bp position of the first end
        // carried B -> C
        uint32_t c_ia = 0, c_ib = 0;
        float c_dij = -1.0f;      // < 0: nothing to do
        double c_dij_d = 0.0;     // 1D
        uint32_t c_upd = 0;       // 1D: bit0 move a, bit1 move b

        const int n_terms = ROUNDS * (int) vi.sweeps;   // this thread's terms in this visit: `sweeps` passes over its ROUNDS steps
#pragma unroll 1
        for (int i = 0; i < n_terms + 2; ++i) {
            // ================= stage C(i-2): the update =================
            if (i >= 2) {
                cp_async_wait<1>();   // coordinates of term i-2 landed (the one younger group is term i-1's step record)
                if (DIMS == 2) {
                    if (c_dij >= 0.0f) {
                        const float2 ca = lds64f(slot_ca + ((c_ia & 1u) << 3));
                        const float2 cb = lds64f(slot_cb + ((c_ib & 1u) << 3));
                        const float dij = c_dij;
                        float mu = p.eta_f * rcp_approx(dij);      // path_sgd_layout.cpp:294-363 in fp32
                        mu = fminf(mu, 1.0f);
                        float dx = ca.x - cb.x;
                        const float dy = ca.y - cb.y;
                        if (dx == 0.0f) dx = 1e-9f;
                        const float s = fmaf(dx, dx, dy * dy);
                        const float rs = rsqrt_approx(s);
                        const float mag = s * rs;
                        const float Delta = 0.5f * mu * (mag - dij);
                        delta_max = fmaxf(delta_max, fabsf(Delta));
                        const float rr = Delta * rs;             // Delta / mag
                        const float r_x = rr * dx, r_y = rr * dy;
                        float2* pa = reinterpret_cast<float2*>(xy_base(p, c_ia >> 1)) + c_ia;
                        float2* pb = reinterpret_cast<float2*>(xy_base(p, c_ib >> 1)) + c_ib;
                        if (atomic_add) {
                            red_coord2(pa, -r_x, -r_y, pol_keep);
                            red_coord2(pb, r_x, r_y, pol_keep);
                        } else {
                            const float2 na = make_float2(ca.x - r_x, ca.y - r_y);
                            const float2 bs = (c_ia == c_ib) ? na : cb;
                            const float2 nb = make_float2(bs.x + r_x, bs.y + r_y);
                            if (st_mode) { __stcg(pa, na); __stcg(pb, nb); }
                            else {
                                atomicExch(reinterpret_cast<unsigned long long*>(pa), ((unsigned long long) __float_as_uint(na.y) << 32) | __float_as_uint(na.x));
                                atomicExch(reinterpret_cast<unsigned long long*>(pb), ((unsigned long long) __float_as_uint(nb.y) << 32) | __float_as_uint(nb.x));
                            }
                        }
                        ++done;
                    }
                } else {
                    if (c_dij >= 0.0f) {
                        if (c_upd) {   // path_sgd.cpp:332-392 in fp64
                            const double xa = lds64d(slot_ca + ((c_ia & 1u) << 3));
                            const double xb = lds64d(slot_cb + ((c_ib & 1u) << 3));
                            const double d = c_dij_d;
                            double mu = p.eta / d;
                            if (mu > 1.0) mu = 1.0;
                            double dx = xa - xb;
                            if (dx == 0.0) dx = 1e-9;
                            const double mag = fabs(dx);
                            const double Delta = mu * (mag - d) * 0.5;
                            delta_max = fmaxf(delta_max, (float) fabs(Delta));
                            const double r_x = Delta / mag * dx;
                            double* qa = x1d_base(p, c_ia) + c_ia;
                            double* qb = x1d_base(p, c_ib) + c_ib;
                            if (atomic_add) {
                                if (c_upd & 1u) red_coord1(qa, -r_x, pol_keep);
                                if (c_upd & 2u) red_coord1(qb, r_x, pol_keep);
                            } else {
                                const double nav = xa - r_x;
                                const double bs = (c_ia == c_ib && (c_upd & 1u)) ? nav : xb;
                                if (st_mode) {
                                    if (c_upd & 1u) __stcg(qa, nav);
                                    if (c_upd & 2u) __stcg(qb, bs + r_x);
                                } else {
                                    if (c_upd & 1u) atomicExch(reinterpret_cast<unsigned long long*>(qa), (unsigned long long) __double_as_longlong(nav));
                                    if (c_upd & 2u) atomicExch(reinterpret_cast<unsigned long long*>(qb), (unsigned long long) __double_as_longlong(bs + r_x));
                                }
                            }
                        }
                        ++done;   // both ends frozen: counted, nothing moves (path_sgd.cpp:298-302)
                    }
                }
                c_dij = -1.0f;
            }
            // ================= stage B(i-1): partner record -> distance, coordinate loads =================
            if (i >= 1 && i <= n_terms) {
                cp_async_wait<0>();   // the partner's step record landed in this thread's slot
                c_dij = -1.0f;
                if (b_rb & 2u) {
                    const uint4 rb = lds128(b_rb & ~15u);
                    if (DIMS == 2) {
                        uint32_t end_b = rb.x & 1u;
                        uint32_t pb_lo = rb.z, pb_hi = rb.w;
                        if (b_rb & 1u) {   // flip: the far end along the path (path_sgd_layout.cpp:252-269)
                            end_b ^= 1u;
                            const uint64_t q = (((uint64_t) pb_hi << 32) | pb_lo) + rb.y;
                            pb_lo = (uint32_t) q; pb_hi = (uint32_t) (q >> 32);
                        }
                        float dij;
                        if (p.pos32) {
                            const uint32_t d = b_pa_lo > pb_lo ? b_pa_lo - pb_lo : pb_lo - b_pa_lo;
                            dij = d ? __uint2float_rn(d) : 1e-9f;       // term_dist == 0 -> 1e-9 (:283-285)
                        } else {
                            const uint64_t a64 = ((uint64_t) b_pa_hi << 32) | b_pa_lo, b64 = ((uint64_t) pb_hi << 32) | pb_lo;
                            const uint64_t d = a64 > b64 ? a64 - b64 : b64 - a64;
                            dij = d ? __ull2float_rn(d) : 1e-9f;
                        }
                        c_dij = dij;
                        c_ia = b_ia;
                        c_ib = (rb.x >> 1) * 2u + end_b;
                        // the node's float4 {x0,y0,x1,y1} (16 bytes, one sector half) -> slot; the end is picked in stage C
                        cp_async_16(slot_ca, reinterpret_cast<const float4*>(xy_base(p, c_ia >> 1)) + (c_ia >> 1), pol_cld);
                        cp_async_16(slot_cb, reinterpret_cast<const float4*>(xy_base(p, c_ib >> 1)) + (c_ib >> 1), pol_cld);
                    } else {
                        const uint32_t na = b_ia, nb = rb.x >> 1;
                        uint32_t u = 3u;
                        if (p.frozen) {   // odgi sort -H target nodes (path_sgd.cpp:290-297)
                            if (p.frozen[na]) u &= ~1u;
                            if (p.frozen[nb]) u &= ~2u;
                        }
                        // 1D uses node starts only (path_sgd.cpp:305-306)
                        const uint64_t a64 = ((uint64_t) b_pa_hi << 32) | b_pa_lo, b64 = ((uint64_t) rb.w << 32) | rb.z;
                        const uint64_t d = a64 > b64 ? a64 - b64 : b64 - a64;
                        if (u == 0) {
                            c_dij = 1.0f; c_upd = 0;
                        } else if (d != 0) {   // d == 0: `continue`, not counted (path_sgd.cpp:320-323)
                            c_dij = 1.0f; c_upd = u;
                            c_dij_d = __ull2double_rn(d);
                            c_ia = na; c_ib = nb;
                            cp_async_16(slot_ca, reinterpret_cast<const double2*>(x1d_base(p, na)) + (na >> 1), pol_keep);
                            cp_async_16(slot_cb, reinterpret_cast<const double2*>(x1d_base(p, nb)) + (nb >> 1), pol_keep);
                        }
                    }
                }
                cp_async_commit();
                b_rb = 0;
            }
            // ================= stage A(i): first step from the tile, partner draw, far record on its way =================
            if (i < n_terms) {
                uint32_t j = (uint32_t) (i % ROUNDS) * 256 + threadIdx.x;
                b_rb = 0;
                if (j < vi.terms && j < vi.n_in_tile) {
                    if (p.flags & 16384u)   // experiments: the first step drawn WITH replacement inside the tile instead of every staged step once
                        j = __umulhi((uint32_t) (xoshiro_next(g) >> 11), min(vi.terms, vi.n_in_tile));
                    if (p.flags & 32768u) {   // experiments: every warp draws a 32-step segment of the tile with replacement (lanes stay neighbours)
                        const uint32_t lim = min(vi.terms, vi.n_in_tile) >> 5;
                        const uint32_t am = __activemask();
                        const uint32_t sgm = __shfl_sync(am, __umulhi((uint32_t) (xoshiro_next(g) >> 11), lim), __ffs(am) - 1);
                        if (lim) j = sgm * 32u + (threadIdx.x & 31u);
                    }
                    if (p.flags & 65536u) {   // experiments: every staged step once, but neighbouring lanes far apart in the tile (odd multiplier mod TILE)
                        // groups of 1 / 2 / 4 neighbouring lanes stay neighbours (flags 131072: 2, 262144: 4): a group of 2 is one 32-byte coordinate sector
                        const uint32_t gs = (p.flags & 262144u) ? 2u : ((p.flags & 131072u) ? 1u : 0u);
                        const uint32_t jj = (((((j >> gs) * 1229u + (uint32_t) (vi.base >> 11) * 7u) & (uint32_t) ((TILE >> gs) - 1)) << gs) | (j & ((1u << gs) - 1u)));
                        if (jj < vi.terms && jj < vi.n_in_tile && vi.n_in_tile == (uint32_t) TILE && vi.terms == (uint32_t) TILE) j = jj;
                    }
                    uint64_t f = vi.f;
                    uint32_t count = vi.count, s_rank = rank0 + j;
                    if (!vi.one_path) {
                        const uint64_t ia = vi.base + j;
                        const uint32_t pp = find_path(first, p.path_count, ia);
                        f = first[pp];
                        count = (uint32_t) (first[pp + 1] - f);
                        s_rank = (uint32_t) (ia - f);
                    }
                    if (count > 1) {   // steps of 1-step paths are skipped, not counted (path_sgd_layout.cpp:190-192)
                        const uint4 ra = lds128(tile_s + j * 16);
                        const uint64_t x = xoshiro_next(g);
                        const uint32_t top = (uint32_t) (x >> 32);
                        // --- partner rank: coin | direction | dirty Zipf  or  uniform in path (path_sgd_layout.cpp:205-262) ---
                        const bool zipf = p.cooling || (top >> 31);
                        const bool backward = (s_rank > 0 && ((top >> 30) & 1u)) || s_rank == count - 1;
                        const uint32_t room = backward ? s_rank : count - s_rank - 1;
                        const uint32_t js = min(p.space, room);
                        uint32_t zi = js;
                        if (js > p.space_max) {
                            const uint32_t n = js - p.space_max;
                            uint32_t k;
                            if (p.space_q_rcp != 0.0f) {  // exact quotient from the fp32 reciprocal + one correction step (host: space < 2^24, q >= 16)
                                k = __float2uint_rz(__uint2float_rz(n) * p.space_q_rcp);
                                const int32_t rem = (int32_t) (n - k * p.space_q);
                                k += (rem >= (int32_t) p.space_q) ? 1u : 0u;
                                k -= (rem < 0) ? 1u : 0u;
                            } else {
                                k = n / p.space_q;
                            }
                            zi = p.space_max + k + 1;
                        }
                        const float2 zt = __ldg(p.ztab + zi);   // {zeta_n, 1 / (1 - zeta_2 / zeta_n)}
                        const float nf = __uint2float_rn(js);
                        const float eta_z = (1.0f - pow_hack_frac(2.0f * rcp_approx(nf), p.one_minus_theta)) * zt.y;
                        const float u = (float) ((top >> 4) & 0xFFFFFFu) * 5.9604644775390625e-8f;  // 2^-24
                        const float uz = u * zt.x;
                        const float bs = fmaf(eta_z, u, 1.0f - eta_z);
                        const float pw = pow_int(bs, p.alpha_int) * pow_hack_frac(bs, p.alpha_frac);
                        uint32_t z = 1u + __float2uint_rz(nf * pw);
                        z = min(z, js);
                        z = uz < p.thresh2 ? 2u : z;
                        z = uz < 1.0f ? 1u : z;
                        const uint32_t rank_z = backward ? s_rank - z : s_rank + z;
                        const uint32_t rank_u = __umulhi((uint32_t) (x >> 4), count);  // bits [35:4]
                        const uint32_t rank_b = zipf ? rank_z : rank_u;
                        const uint64_t ib = f + rank_b;
                        const uint64_t off = ib - vi.base;           // wraps to a huge value when ib < base
                        const uint32_t flip_a = (top >> 29) & 1u, flip_b = (top >> 28) & 1u;
                        if (p.trace) {
                            const unsigned long long kk = atomicAdd(p.trace_count, 1ull);
                            if (kk < p.trace_cap) {
                                p.trace[2 * kk] = vi.base + j;
                                p.trace[2 * kk + 1] = ib | ((unsigned long long) (DIMS == 2 ? flip_a : 0u) << 62) | ((unsigned long long) (DIMS == 2 ? flip_b : 0u) << 63);
                            }
                        }
                        uint32_t src = slot_rb;
                        if (off < (uint64_t) vi.n_in_tile) src = tile_s + (uint32_t) off * 16u;
                        else cp_async_16(slot_rb, p.steps + ib, pol_stream);
                        // first end: orientation-aware end choice, end-adjusted position
                        uint32_t pa_lo = ra.z, pa_hi = ra.w;
                        if (DIMS == 2) {
                            uint32_t end_a = ra.x & 1u;
                            if (flip_a) {
                                end_a ^= 1u;
                                const uint64_t q = (((uint64_t) pa_hi << 32) | pa_lo) + ra.y;
                                pa_lo = (uint32_t) q; pa_hi = (uint32_t) (q >> 32);
                            }
                            b_ia = (ra.x >> 1) * 2u + end_a;
                            b_rb = src | 2u | flip_b;
                        } else {
                            b_ia = ra.x >> 1;
                            b_rb = src | 2u;
                        }
                        b_pa_lo = pa_lo; b_pa_hi = pa_hi;
                    }
                }
            }
            // one group per iteration at the stage-A position (empty in the two drain iterations), so that "all but the
            // youngest group" in stage C always covers the coordinate loads of term i-2
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
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

template <int DIMS, bool TMA, int TILE>
cudaError_t tile2_launch(const Tile2Params& p, const LaunchShape& s, cudaStream_t stream) {
    auto k = pgsgd_tile2_kernel<DIMS, TMA, TILE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) s.smem);
    if (e != cudaSuccess) return e;
    k<<<s.grid, 256, s.smem, stream>>>(p);
    return cudaGetLastError();
}
template <int DIMS, bool TMA, int TILE>
cudaError_t tile2_occ(size_t smem, int* out) {
    auto k = pgsgd_tile2_kernel<DIMS, TMA, TILE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k, 256, smem);
}

__global__ void max_path_bp_kernel(const StepRec* steps, const uint64_t* first, uint32_t P, unsigned long long* out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= P || first[q + 1] == first[q]) return;
    const StepRec r = steps[first[q + 1] - 1];
    atomicMax(out, ((((unsigned long long) r.pos_hi) << 32) | r.pos_lo) + r.len);
}

}  // namespace

cudaError_t launch_max_path_bp(const StepRec* steps, const uint64_t* first, uint32_t P, unsigned long long* out, cudaStream_t stream) {
    if (!P) return cudaSuccess;
    max_path_bp_kernel<<<(P + 255) / 256, 256, 0, stream>>>(steps, first, P, out);
    return cudaGetLastError();
}

size_t tile2_smem_bytes(int tile_steps, bool tma, uint32_t path_count, bool* smem_paths) {
    const size_t base = (size_t) (tma ? 2 : 1) * tile_steps * 16 + 3 * 256 * 16;
    const size_t paths = ((size_t) path_count + 1) * sizeof(uint64_t);
    const bool sp = paths <= 8 * 1024;   // keep 4 CTAs per SM: the table stays in global memory beyond ~1000 paths
    if (smem_paths) *smem_paths = sp;
    return base + (sp ? paths : 0);
}

#define TILE2_DISPATCH(FN, ...)                                                                   \
    if (dims == 2 && !tma && tile_steps == 2048) return FN<2, false, 2048>(__VA_ARGS__);          \
    if (dims == 2 && !tma && tile_steps == 1024) return FN<2, false, 1024>(__VA_ARGS__);          \
    if (dims == 2 && !tma && tile_steps == 4096) return FN<2, false, 4096>(__VA_ARGS__);          \
    if (dims == 2 && tma && tile_steps == 2048) return FN<2, true, 2048>(__VA_ARGS__);            \
    if (dims == 2 && tma && tile_steps == 1024) return FN<2, true, 1024>(__VA_ARGS__);            \
    if (dims == 1 && !tma && tile_steps == 2048) return FN<1, false, 2048>(__VA_ARGS__);          \
    if (dims == 1 && !tma && tile_steps == 1024) return FN<1, false, 1024>(__VA_ARGS__);          \
    if (dims == 1 && tma && tile_steps == 1024) return FN<1, true, 1024>(__VA_ARGS__);            \
    if (dims == 1 && tma && tile_steps == 2048) return FN<1, true, 2048>(__VA_ARGS__);            \
    return cudaErrorInvalidValue;

cudaError_t launch_tile2_iteration(int dims, int tile_steps, bool tma, const Tile2Params& p, const LaunchShape& shape, cudaStream_t stream) {
    TILE2_DISPATCH(tile2_launch, p, shape, stream)
}
cudaError_t tile2_occupancy(int dims, int tile_steps, bool tma, size_t smem, int* blocks_per_sm) {
    TILE2_DISPATCH(tile2_occ, smem, blocks_per_sm)
}

}  // namespace pgsgd
