// pgsgd_device.cuh — device-side building blocks of the B200 PG-SGD kernels (sm_100a).
//
//  * Xoshiro256+ / SplitMix64 worker streams       (reference: deps/Xoshiro-cpp/XoshiroCpp.hpp:684-746)
//  * libstdc++-compatible integer / canonical draws  (bits/uniform_int_dist.h:252-283, bits/random.tcc:3349-3381)
//  * dirty Zipf with the reference's bit-hack pow     (deps/dirtyzipf/dirty_zipfian_int_distribution.h:82-104,230-243)
//  * the term sampler shared by the SGD kernels and the verification hook
//
// Every floating-point operation whose result feeds an integer decision is written with explicit
// round-to-nearest intrinsics (__dmul_rn, __dadd_rn, ...) so that nvcc cannot contract it into an FMA:
// given the same stream seed the device draws the SAME term sequence as a reference CPU worker thread.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pgsgd {

// ---- HBM-resident step record: one 16-byte load per path step -------------------------------------
//   x = (node_rank << 1) | is_reverse      (== the libhandlegraph handle integer, util.hpp:44-62)
//   y = node length in bp
//   z,w = 64-bit bp offset of the node start within its path (== XP positions[rank], xp.cpp:393-397)
struct __align__(16) StepRec {
    uint32_t handle;
    uint32_t len;
    uint32_t pos_lo;
    uint32_t pos_hi;
};
static_assert(sizeof(StepRec) == 16, "StepRec must be one 128-bit load");

struct Xoshiro {
    uint64_t s0, s1, s2, s3;
};

__host__ __device__ __forceinline__ uint64_t splitmix64_next(uint64_t& state) {
    uint64_t z = (state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// Xoshiro256Plus(seed): four SplitMix64 outputs (XoshiroCpp.hpp:729-730)
__host__ __device__ __forceinline__ void xoshiro_seed(Xoshiro& g, uint64_t seed) {
    uint64_t sm = seed;
    g.s0 = splitmix64_next(sm);
    g.s1 = splitmix64_next(sm);
    g.s2 = splitmix64_next(sm);
    g.s3 = splitmix64_next(sm);
}

// Xoshiro256Plus::operator() (XoshiroCpp.hpp:735-746)
__device__ __forceinline__ uint64_t xoshiro_next(Xoshiro& g) {
    const uint64_t result = g.s0 + g.s3;
    const uint64_t t = g.s1 << 17;
    g.s2 ^= g.s0;
    g.s3 ^= g.s1;
    g.s1 ^= g.s2;
    g.s0 ^= g.s3;
    g.s2 ^= t;
    g.s3 = (g.s3 << 45) | (g.s3 >> 19);
    return result;
}

// std::uniform_int_distribution<uint64_t>(0, range-1): Lemire's method on the 128-bit product with
// the rejection loop (taken with probability range / 2^64)
__device__ __forceinline__ uint64_t draw_uniform(Xoshiro& g, uint64_t range) {
    uint64_t x = xoshiro_next(g);
    uint64_t hi = __umul64hi(x, range);
    uint64_t lo = x * range;
    if (lo < range) {
        const uint64_t threshold = (0 - range) % range;
        while (lo < threshold) {
            x = xoshiro_next(g);
            hi = __umul64hi(x, range);
            lo = x * range;
        }
    }
    return hi;
}

// uniform_int_distribution<uint64_t>(0,1): the product's high word is the generator's top bit; the
// rejection threshold (-2 % 2) is 0, so exactly one draw is consumed
__device__ __forceinline__ uint32_t draw_flip(Xoshiro& g) { return (uint32_t)(xoshiro_next(g) >> 63); }

// std::generate_canonical<double,53>: double(x) / 2^64, clamped below 1
__device__ __forceinline__ double draw_canonical(Xoshiro& g) {
    double r = __dmul_rn(__ull2double_rn(xoshiro_next(g)), 5.421010862427522e-20 /* 2^-64, exact */);
    if (r >= 1.0) r = __longlong_as_double(0x3FEFFFFFFFFFFFFFLL);  // nextafter(1, 0)
    return r;
}

// dirtyzipf::fast_precise_pow — the result is defined by these exact operations, not by pow()
__device__ __forceinline__ double fast_precise_pow(double a, double b) {
    int e = __double2int_rz(b);
    const int hi = __double2hiint(a);
    const double t = __dadd_rn(__dmul_rn(__dsub_rn(b, (double) e), (double) (hi - 1072632447)), 1072632447.0);
    const double frac = __hiloint2double(__double2int_rz(t), 0);
    double r = 1.0;
    while (e) {
        if (e & 1) r = __dmul_rn(r, a);
        a = __dmul_rn(a, a);
        e >>= 1;
    }
    return __dmul_rn(r, frac);
}

// constants of one Zipf configuration, computed once per iteration on the host with the same bit-hack pow
struct ZipfConst {
    double theta;          // exponent handed to the draw (1D cooling: 0.001 with zetas of the original theta)
    double one_minus_theta;
    double alpha;          // 1 / (1 - theta)
    double zeta2;          // zeta(2, theta) as the reference recomputes per draw (dirty_zipfian...h:126-128)
    double thresh2;        // 1.0 + fast_precise_pow(0.5, theta)
};

// dirty_zipfian_int_distribution<uint64_t>(1, n, theta, zeta_n)(gen)
__device__ __forceinline__ uint64_t draw_zipf(Xoshiro& g, uint64_t n, const ZipfConst& zc, double zeta_n) {
    const double eta = __ddiv_rn(__dsub_rn(1.0, fast_precise_pow(__ddiv_rn(2.0, (double) n), zc.one_minus_theta)),
                                 __dsub_rn(1.0, __ddiv_rn(zc.zeta2, zeta_n)));
    const double u = draw_canonical(g);
    const double uz = __dmul_rn(u, zeta_n);
    if (uz < 1.0) return 1;
    if (uz < zc.thresh2) return 2;
    const double base = __dadd_rn(__dsub_rn(__dmul_rn(eta, u), eta), 1.0);
    return __double2ull_rz(__dadd_rn(1.0, __dmul_rn((double) n, fast_precise_pow(base, zc.alpha))));
}

// fp32 constants of the same Zipf configuration for the economical sampler of the tile kernel
struct ZipfConstF {
    float one_minus_theta;
    float alpha_frac;      // alpha - floor(alpha)
    int   alpha_int;       // floor(alpha)
    float zeta2;
    float thresh2;
};

// ---- everything the sampler needs, passed by value to the kernels ---------------------------------
struct SamplerParams {
    const uint64_t* path_first;   // [P+1] global copy (used when the table does not fit in shared memory)
    const double* zetas;          // zeta table (path_sgd_layout.cpp:87-97)
    uint64_t step_count;          // S
    uint32_t path_count;          // P
    uint32_t cooling;             // 1: always take the Zipf branch (path_sgd_layout.cpp:205)
    uint64_t space, space_max, space_q;
    ZipfConst zipf;
    ZipfConstF zipf_f;
};

struct Term {
    uint64_t ia, ib;      // global step indices of the two steps
    uint64_t step_index;  // the raw first draw
    uint32_t path;
    uint32_t flip_a, flip_b;
    uint32_t valid;
};

// largest p with first[p] <= idx (first[P] = S > idx)
template <typename FirstPtr>
__device__ __forceinline__ uint32_t find_path(FirstPtr first, uint32_t P, uint64_t idx) {
    uint32_t lo = 0, hi = P;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (first[mid] <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

// The partner draw given the first step (path start f, step count, rank s_rank): [coin], [direction coin],
// Zipf | uniform partner, [end a, end b] — the reference's order of RNG consumption (2D: path_sgd_layout.cpp:205-262,
// 1D: path_sgd.cpp:245-279).  Shared by the stream kernel, the tile kernel and the verification hook.
template <int DIMS>
__device__ __forceinline__ void draw_partner(const SamplerParams& sp, Xoshiro& g, uint64_t f, uint64_t count, uint64_t s_rank, Term& t) {
    uint64_t rank_b;
    if (sp.cooling || draw_flip(g)) {
        const bool backward = (s_rank > 0 && draw_flip(g)) || s_rank == count - 1;
        const uint64_t room = backward ? s_rank : count - s_rank - 1;
        const uint64_t jump_space = sp.space < room ? sp.space : room;
        uint64_t zi = jump_space;
        if (jump_space > sp.space_max) zi = sp.space_max + (jump_space - sp.space_max) / sp.space_q + 1;
        const uint64_t z = draw_zipf(g, jump_space, sp.zipf, __ldg(sp.zetas + zi));
        rank_b = backward ? s_rank - z : s_rank + z;
    } else {
        rank_b = draw_uniform(g, count);
    }
    t.ib = f + rank_b;
    if (DIMS == 2) {
        t.flip_a = draw_flip(g);
        t.flip_b = draw_flip(g);
    }
}

// ---- economical partner draw (tile kernel) ---------------------------------------------------------------------
// Same law as draw_partner, restated for throughput: ONE Xoshiro256+ output feeds a whole term (4 coin bits, a
// 24-bit uniform for the Zipf draw or a 32-bit uniform partner), and the dirty Zipf — including the reference's
// exponent-bit-hack pow — is evaluated in fp32.  The fp32 evaluation can move a partner by one step in rare boundary
// cases; the distribution is checked against the oracle's in tests/test_gpu_parity.py (chi-square on the jump lengths).

// dirtyzipf::fast_precise_pow with the exponent split as (e + frac) by the caller, on floats: the hack only reads the
// top 20 mantissa bits of the double's high word, which a float carries exactly
__device__ __forceinline__ float fast_precise_pow_f32(float a, int e, float frac) {
    const int hi = (__float_as_int(a) >> 3) + 0x38000000;                  // high word of (double) a
    const int t = 1072632447 + __float2int_rd(frac * (float) (hi - 1072632447));
    const float f = __int_as_float((t - 0x38000000) << 3);
    float r = 1.0f;
    while (e) {
        if (e & 1) r *= a;
        a *= a;
        e >>= 1;
    }
    return r * f;
}

__device__ __forceinline__ uint64_t draw_zipf_f32(uint32_t u24, uint64_t n, const ZipfConstF& zc, float zeta_n) {
    const float nf = (float) n;
    const float eta = __fdividef(1.0f - fast_precise_pow_f32(__fdividef(2.0f, nf), 0, zc.one_minus_theta),
                                 1.0f - __fdividef(zc.zeta2, zeta_n));
    const float u = (float) u24 * 5.9604644775390625e-8f;  // 2^-24
    const float uz = u * zeta_n;
    if (uz < 1.0f) return 1;
    if (uz < zc.thresh2) return 2;
    const float base = fmaf(eta, u, 1.0f - eta);
    uint64_t z = 1 + (uint64_t) (nf * fast_precise_pow_f32(base, zc.alpha_int, zc.alpha_frac));
    return z > n ? n : z;
}

template <int DIMS>
__device__ __forceinline__ void draw_partner_fast(const SamplerParams& sp, Xoshiro& g, uint64_t f, uint64_t count, uint64_t s_rank, Term& t) {
    const uint64_t x = xoshiro_next(g);
    const uint32_t top = (uint32_t) (x >> 32);
    uint64_t rank_b;
    if (sp.cooling || (top >> 31)) {
        const bool backward = (s_rank > 0 && ((top >> 30) & 1u)) || s_rank == count - 1;
        const uint64_t room = backward ? s_rank : count - s_rank - 1;
        const uint64_t jump_space = sp.space < room ? sp.space : room;
        uint64_t zi = jump_space;
        if (jump_space > sp.space_max) zi = sp.space_max + (jump_space - sp.space_max) / sp.space_q + 1;
        const uint64_t z = draw_zipf_f32((top >> 4) & 0xFFFFFFu, jump_space, sp.zipf_f, (float) __ldg(sp.zetas + zi));
        rank_b = backward ? s_rank - z : s_rank + z;
    } else if (count <= 0xFFFFFFFFull) {
        rank_b = __umulhi((uint32_t) (x >> 4), (uint32_t) count);  // bits [35:4]
    } else {
        rank_b = draw_uniform(g, count);
    }
    t.ib = f + rank_b;
    if (DIMS == 2) {
        t.flip_a = (top >> 29) & 1u;
        t.flip_b = (top >> 28) & 1u;
    }
}

// Resolve a global step index to its path and draw the partner.  valid = 0 for a step of a 1-step path
// (`continue` without counting, path_sgd_layout.cpp:190-192).
template <int DIMS, typename FirstPtr>
__device__ __forceinline__ void draw_term_at(const SamplerParams& sp, FirstPtr first, Xoshiro& g, uint64_t step_index, Term& t) {
    const uint32_t p = find_path(first, sp.path_count, step_index);
    const uint64_t f = first[p];
    const uint64_t count = first[p + 1] - f;
    t.step_index = step_index;
    t.path = p;
    t.flip_a = t.flip_b = 0;
    t.ia = t.ib = step_index;
    if (count == 1) {
        t.valid = 0;
        return;
    }
    t.valid = 1;
    draw_partner<DIMS>(sp, g, f, count, step_index - f, t);
}

// One worker draw exactly as a reference worker thread makes it: uniform first step over all steps
// (path_sgd_layout.cpp:182), then the partner.
template <int DIMS, typename FirstPtr>
__device__ __forceinline__ void draw_term(const SamplerParams& sp, FirstPtr first, Xoshiro& g, Term& t) {
    draw_term_at<DIMS>(sp, first, g, draw_uniform(g, sp.step_count), t);
}

// ---- tile visiting order: a keyed pseudo-random PERMUTATION of [0, n) per pass ------------------------------------------
// Four-round Feistel network on the smallest even-width power-of-two domain >= n, cycle-walked back into [0, n).  Round 1
// used an affine map i -> (i * mul + add) mod n: a bijection, but the tiles of concurrently resident CTAs then form an
// arithmetic progression, which resonates with the path structure (tiles a whole path apart cover the SAME nodes): on a
// long thin graph two seeds in six ended 15-50 % above the reference's far-stress band, stream sampling none (profiles/).
__host__ __device__ __forceinline__ uint32_t perm_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint64_t tile_perm(uint64_t i, uint64_t n, uint64_t key) {
    if (n <= 1) return 0;
    unsigned bits = 1;
    while ((1ull << bits) < n) ++bits;
    bits += bits & 1u;                       // even: two equal halves
    const unsigned half = bits >> 1;
    const uint64_t mask = (1ull << half) - 1;
    uint64_t x = i;
    do {
        uint64_t L = x >> half, R = x & mask;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint64_t kr = key >> (16 * r);
            const uint64_t f = ((uint64_t) perm_mix((uint32_t) R + (uint32_t) kr) | ((uint64_t) perm_mix((uint32_t) (R >> 32) ^ (uint32_t) (kr >> 7) ^ 0x9e3779b9U) << 32)) & mask;
            const uint64_t t = L ^ f;
            L = R;
            R = t;
        }
        x = (L << half) | R;
    } while (x >= n);
    return x;
}

// ---- L2 residency control ---------------------------------------------------------------------------
// The step records are a multi-GB stream with no reuse; the coordinates (16 B per node) are re-read and re-written
// millions of times per iteration and fit (or nearly fit) the 126 MB L2.  Step-record loads therefore carry an
// L2::evict_first policy and every coordinate access an L2::evict_last policy, so the stream cannot push the
// coordinates out (ncu before/after: profiles/).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint4 load_step(const StepRec* steps, uint64_t i, uint64_t policy) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(reinterpret_cast<const uint4*>(steps) + i), "l"(policy));
    return r;
}
__device__ __forceinline__ uint4 load_step(const StepRec* steps, uint64_t i) {
    return __ldg(reinterpret_cast<const uint4*>(steps) + i);
}

__device__ __forceinline__ float2 ld_coord2(const float2* p, uint64_t policy) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void red_coord2(float2* p, float dx, float dy, uint64_t policy) {
    asm volatile("red.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(p), "f"(dx), "f"(dy), "l"(policy) : "memory");
}
__device__ __forceinline__ double ld_coord1(const double* p, uint64_t policy) {
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void red_coord1(double* p, double d, uint64_t policy) {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(p), "d"(d), "l"(policy) : "memory");
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: the tile kernel's global -> shared staging --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one elected thread: `bytes` (multiple of 16) from global to shared, completion signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t step_pos(const uint4& r) { return ((uint64_t) r.w << 32) | r.z; }

}  // namespace pgsgd
