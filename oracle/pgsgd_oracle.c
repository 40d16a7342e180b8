/* pgsgd_oracle.c — CPU restatement of odgi's path-guided SGD.  TEST INFRASTRUCTURE ONLY
 * (see pgsgd_oracle.h).  Plain C99, IEEE-754 double arithmetic; build with -ffp-contract=off and
 * without -ffast-math so the result is the one the C++ source of the reference defines.
 */
#include "pgsgd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * RNG — deps/Xoshiro-cpp/XoshiroCpp.hpp
 * ---------------------------------------------------------------------------------------------- */

/* SplitMix64::operator() (:684-690) */
static uint64_t splitmix64_next(uint64_t* state) {
    uint64_t z = (*state += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* Xoshiro256Plus(seed) (:729-730): state = SplitMix64{seed}.generateSeedSequence<4>() (:692-703) */
void orc_rng_seed(orc_rng* g, uint64_t seed) {
    uint64_t sm = seed;
    for (int i = 0; i < 4; ++i) g->s[i] = splitmix64_next(&sm);
}

/* Xoshiro256Plus::operator() (:735-746) */
uint64_t orc_rng_next(orc_rng* g) {
    uint64_t* s = g->s;
    const uint64_t result = s[0] + s[3];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = (s[3] << 45) | (s[3] >> 19);
    return result;
}

/* libstdc++ 13 uniform_int_distribution<uint64_t>(0, range-1)(urng) with a full-64-bit URBG:
 * operator() takes the "downscaling" branch (bits/uniform_int_dist.h:300-320) and calls
 * _S_nd<unsigned __int128>(urng, range) (:252-283), Lemire's nearly divisionless method. */
uint64_t orc_uniform(orc_rng* g, uint64_t range) {
    unsigned __int128 product = (unsigned __int128) orc_rng_next(g) * (unsigned __int128) range;
    uint64_t low = (uint64_t) product;
    if (low < range) {
        uint64_t threshold = (0 - range) % range;
        while (low < threshold) {
            product = (unsigned __int128) orc_rng_next(g) * (unsigned __int128) range;
            low = (uint64_t) product;
        }
    }
    return (uint64_t)(product >> 64);
}

/* libstdc++ 13 generate_canonical<double, 53>(urng) (bits/random.tcc:3349-3381) for a URBG with
 * min()=0, max()=2^64-1: __r = 2^64, __log2r = 64, __m = 1, so one draw:
 *   __sum = double(urng()) ; __tmp = 2^64 ; __ret = __sum / __tmp ; if (__ret >= 1) nextafter(1, 0) */
double orc_canonical(orc_rng* g) {
    double sum = (double) orc_rng_next(g); /* u64 -> double, round-to-nearest */
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret;
}

/* ------------------------------------------------------------------------------------------------
 * dirty zipf — deps/dirtyzipf/dirty_zipfian_int_distribution.h
 * ---------------------------------------------------------------------------------------------- */

/* fast_precise_pow (:82-104): exponent-bit hack on the fractional part of b, squaring on the integer part */
double orc_fast_precise_pow(double a, double b) {
    int e = (int) b;
    union { double d; int x[2]; } u;
    u.d = a;
    u.x[1] = (int) ((b - e) * (u.x[1] - 1072632447) + 1072632447);
    u.x[0] = 0;
    double r = 1.0;
    while (e) {
        if (e & 1) r *= a;
        a *= a;
        e >>= 1;
    }
    return r * u.d;
}

/* param_type::zeta (:165-171) */
double orc_zeta(uint64_t n, double theta) {
    double ans = 0.0;
    for (uint64_t i = 1; i <= n; ++i) ans += orc_fast_precise_pow(1.0 / i, theta);
    return ans;
}

/* operator()(urng, param_type(1, n, theta, zeta_n)) (:230-243) */
uint64_t orc_dirty_zipf(orc_rng* g, uint64_t n, double theta, double zeta_n) {
    const uint64_t a = 1, b = n;
    const double zeta2theta = orc_zeta(2, theta); /* recomputed by every param_type ctor (:126-128) */
    double alpha = 1 / (1 - theta);
    double eta = (1 - orc_fast_precise_pow(2.0 / (b - a + 1), 1 - theta)) / (1 - zeta2theta / zeta_n);
    double u = orc_canonical(g);
    double uz = u * zeta_n;
    if (uz < 1.0) return a;
    if (uz < 1.0 + orc_fast_precise_pow(0.5, theta)) return a + 1;
    /* __p.a() + ((__p.b() - __p.a() + 1) * pow(...)): uint64 + (uint64 * double) -> double -> uint64 */
    return (uint64_t) ((double) a + ((double) (b - a + 1) * orc_fast_precise_pow(eta * u - eta + 1, alpha)));
}

/* ------------------------------------------------------------------------------------------------
 * schedule + zeta table
 * ---------------------------------------------------------------------------------------------- */

/* path_linear_sgd_layout_schedule (path_sgd_layout.cpp:433-468) == path_linear_sgd_schedule
 * (path_sgd.cpp:466-501), called with w_min = 1/eta_max, w_max = 1 (path_sgd_layout.cpp:75-84) */
void orc_schedule(double eta_max_in, uint64_t iter_max, uint64_t iter_with_max_learning_rate, double eps, double* etas) {
    double w_min = (double) 1.0 / (double) (eta_max_in);
    double w_max = 1.0;
    double eta_max = 1.0 / w_min;
    double eta_min = eps / w_max;
    double lambda = log(eta_max / eta_min) / ((double) iter_max - 1);
    for (int64_t t = 0; t <= (int64_t) iter_max; t++) {
        int64_t d = t - (int64_t) iter_with_max_learning_rate;
        if (d < 0) d = -d;
        etas[t] = eta_max * exp(-lambda * (double) d);
    }
}

/* zeta cache (path_sgd_layout.cpp:87-97) */
uint64_t orc_zetas(uint64_t space, uint64_t space_max, uint64_t space_q, double theta, double* zetas, uint64_t cap) {
    uint64_t n = (space <= space_max ? space : space_max + (space - space_max) / space_q + 1) + 1;
    if (!zetas) return n;
    for (uint64_t i = 0; i < n && i < cap; ++i) zetas[i] = 0.0;
    double zeta_tmp = 0.0;
    for (uint64_t i = 1; i < space + 1; i++) {
        zeta_tmp += orc_fast_precise_pow(1.0 / i, theta);
        if (i <= space_max) {
            if (i < cap) zetas[i] = zeta_tmp;
        }
        if (i >= space_max && (i - space_max) % space_q == 0) {
            uint64_t k = space_max + 1 + (i - space_max) / space_q;
            if (k < cap) zetas[k] = zeta_tmp;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * sampling one term
 * ---------------------------------------------------------------------------------------------- */

static uint64_t find_path(const orc_graph* g, uint64_t idx) {
    uint64_t lo = 0, hi = g->path_count; /* first[lo] <= idx < first[hi] */
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (g->path_first_step[mid] <= idx) lo = mid; else hi = mid;
    }
    /* skip empty paths that share the same offset */
    return lo;
}

static int sample_term_at(const orc_graph* g, const orc_config* c, const double* zetas, int dims, int cooling,
                          double theta_zipf, orc_rng* rng, uint64_t step_index, orc_term* t);

int orc_sample_term(const orc_graph* g, const orc_config* c, const double* zetas, int dims, int cooling,
                    double theta_zipf, orc_rng* rng, orc_term* t) {
    /* path_sgd_layout.cpp:175,182: dis_step(0, np_bv.size()-1) */
    uint64_t step_index = orc_uniform(rng, g->step_count);
    return sample_term_at(g, c, zetas, dims, cooling, theta_zipf, rng, step_index, t);
}

/* everything after the first pick: the partner and the node ends, given the first step */
static int sample_term_at(const orc_graph* g, const orc_config* c, const double* zetas, int dims, int cooling,
                          double theta_zipf, orc_rng* rng, uint64_t step_index, orc_term* t) {
    memset(t, 0, sizeof(*t));
    uint64_t idx = g->step_perm ? g->step_perm[step_index] : step_index;
    /* :186,199: path_i = npi_iv[step_index]; s_rank = nr_iv[step_index] - 1 */
    uint64_t p = find_path(g, idx);
    uint64_t first = g->path_first_step[p];
    uint64_t path_step_count = g->path_first_step[p + 1] - first;
    uint64_t s_rank = idx - first;
    t->step_index = step_index;
    t->path = p;
    /* :189-192 */
    if (path_step_count == 1) return 0;
    uint64_t rank_b;
    /* :205 cooling.load() || flip(gen)   (flip is not drawn while cooling) */
    if (cooling || orc_uniform(rng, 2)) {
        t->zipf = 1;
        /* :206 s_rank > 0 && flip(gen) || s_rank == path_step_count-1 */
        if ((s_rank > 0 && orc_uniform(rng, 2)) || s_rank == path_step_count - 1) {
            /* go backward :208-218 */
            uint64_t jump_space = c->space < s_rank ? c->space : s_rank;
            uint64_t space = jump_space;
            if (jump_space > c->space_max) space = c->space_max + (jump_space - c->space_max) / c->space_quantization_step + 1;
            uint64_t z_i = orc_dirty_zipf(rng, jump_space, theta_zipf, zetas[space]);
            rank_b = s_rank - z_i;
        } else {
            /* go forward :220-231 */
            uint64_t rem = path_step_count - s_rank - 1;
            uint64_t jump_space = c->space < rem ? c->space : rem;
            uint64_t space = jump_space;
            if (jump_space > c->space_max) space = c->space_max + (jump_space - c->space_max) / c->space_quantization_step + 1;
            uint64_t z_i = orc_dirty_zipf(rng, jump_space, theta_zipf, zetas[space]);
            rank_b = s_rank + z_i;
        }
    } else {
        /* :235-237 rando(0, path_step_count-1) */
        rank_b = orc_uniform(rng, path_step_count);
    }
    t->rank_a = s_rank;
    t->rank_b = rank_b;
    uint64_t ia = first + s_rank, ib = first + rank_b;
    /* :242-249 handles, lengths, positions */
    t->node_a = g->step_node[ia];
    t->node_b = g->step_node[ib];
    t->rev_a = g->step_rev[ia];
    t->rev_b = g->step_rev[ib];
    uint64_t pos_a = g->step_pos[ia];
    uint64_t pos_b = g->step_pos[ib];
    if (dims == 2) {
        /* :252-269 end choice; flip == 1 moves to the far end of the node along the path */
        uint8_t fa = (uint8_t) orc_uniform(rng, 2);
        if (fa) { pos_a += g->node_len[t->node_a]; t->end_a = !t->rev_a; } else { t->end_a = t->rev_a; }
        uint8_t fb = (uint8_t) orc_uniform(rng, 2);
        if (fb) { pos_b += g->node_len[t->node_b]; t->end_b = !t->rev_b; } else { t->end_b = t->rev_b; }
        t->flip_a = fa;
        t->flip_b = fb;
    }
    t->pos_a = pos_a;
    t->pos_b = pos_b;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * applying one term
 * ---------------------------------------------------------------------------------------------- */

/* path_sgd_layout.cpp:280-363 */
double orc_apply_2d(const orc_term* t, double eta, double* X, double* Y) {
    double term_dist = fabs((double) t->pos_a - (double) t->pos_b);
    if (term_dist == 0) term_dist = 1e-9;
    double term_weight = 1.0 / (double) term_dist;
    double w_ij = term_weight;
    double mu = eta * w_ij;
    if (mu > 1) mu = 1;
    double d_ij = term_dist;
    uint64_t i = t->node_a, j = t->node_b;
    uint64_t offset_i = t->end_a ? 1 : 0, offset_j = t->end_b ? 1 : 0;
    double dx = X[2 * i + offset_i] - X[2 * j + offset_j];
    double dy = Y[2 * i + offset_i] - Y[2 * j + offset_j];
    if (dx == 0) dx = 1e-9;
    double mag = sqrt(dx * dx + dy * dy);
    double Delta = mu * (mag - d_ij) / 2;
    double Delta_abs = fabs(Delta);
    double r = Delta / mag;
    double r_x = r * dx;
    double r_y = r * dy;
    X[2 * i + offset_i] = X[2 * i + offset_i] - r_x;
    Y[2 * i + offset_i] = Y[2 * i + offset_i] - r_y;
    X[2 * j + offset_j] = X[2 * j + offset_j] + r_x;
    Y[2 * j + offset_j] = Y[2 * j + offset_j] + r_y;
    return Delta_abs;
}

/* path_sgd.cpp:285-392 */
double orc_apply_1d(const orc_term* t, double eta, double* X, const uint8_t* frozen) {
    int update_i = 1, update_j = 1;
    if (frozen) {
        if (frozen[t->node_a]) update_i = 0;
        if (frozen[t->node_b]) update_j = 0;
    }
    if (!update_i && !update_j) return 0.0; /* counted, nothing moves (:298-302) */
    double term_dist = fabs((double) t->pos_a - (double) t->pos_b);
    if (term_dist == 0) return -1.0; /* :320-323 continue — not counted */
    double term_weight = 1.0 / term_dist;
    double w_ij = term_weight;
    double mu = eta * w_ij;
    if (mu > 1) mu = 1;
    double d_ij = term_dist;
    uint64_t i = t->node_a, j = t->node_b;
    double dx = X[i] - X[j];
    if (dx == 0) dx = 1e-9;
    double mag = fabs(dx);
    double Delta = mu * (mag - d_ij) / 2;
    double Delta_abs = fabs(Delta);
    double r = Delta / mag;
    double r_x = r * dx;
    if (update_i) X[i] = X[i] - r_x;
    if (update_j) X[j] = X[j] + r_x;
    return Delta_abs;
}

/* fp32 device model: same algebra as orc_apply_2d with every operation rounded to fp32
 * (the volatile stores forbid the compiler from keeping wider intermediates or fusing) */
float orc_apply_2d_f32(const orc_term* t, double eta, float* xy) {
    uint64_t dpos = t->pos_a > t->pos_b ? t->pos_a - t->pos_b : t->pos_b - t->pos_a;
    volatile float d_ij = (float) dpos; /* u64 -> f32, round-to-nearest */
    if (dpos == 0) d_ij = 1e-9f;
    volatile float mu = (float) eta / d_ij;
    if (mu > 1.0f) mu = 1.0f;
    float* pa = xy + 4 * (uint64_t) t->node_a + 2 * (t->end_a ? 1 : 0);
    float* pb = xy + 4 * (uint64_t) t->node_b + 2 * (t->end_b ? 1 : 0);
    volatile float dx = pa[0] - pb[0];
    volatile float dy = pa[1] - pb[1];
    if (dx == 0.0f) dx = 1e-9f;
    volatile float dx2 = dx * dx;
    volatile float dy2 = dy * dy;
    volatile float s = dx2 + dy2;
    volatile float mag = sqrtf(s);
    volatile float diff = mag - d_ij;
    volatile float md = mu * diff;
    volatile float Delta = md * 0.5f;
    volatile float r = Delta / mag;
    volatile float r_x = r * dx;
    volatile float r_y = r * dy;
    volatile float ax = pa[0] - r_x, ay = pa[1] - r_y;
    pa[0] = ax; pa[1] = ay;
    volatile float bx = pb[0] + r_x, by = pb[1] + r_y;
    pb[0] = bx; pb[1] = by;
    return fabsf(Delta);
}

/* ------------------------------------------------------------------------------------------------
 * whole runs (deterministic iteration boundaries)
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    double* etas;
    double* zetas;
    uint64_t first_cooling_iteration;
} run_tables;

static int any_path_with_more_than_one_step(const orc_graph* g) {
    for (uint64_t p = 0; p < g->path_count; ++p)
        if (g->path_first_step[p + 1] - g->path_first_step[p] > 1) return 1;
    return 0;
}

static void tables_init(run_tables* rt, const orc_config* c) {
    rt->etas = (double*) malloc(sizeof(double) * (c->iter_max + 1));
    orc_schedule(c->eta_max, c->iter_max, c->iter_with_max_learning_rate, c->eps, rt->etas);
    uint64_t nz = orc_zetas(c->space, c->space_max, c->space_quantization_step, c->theta, NULL, 0);
    rt->zetas = (double*) malloc(sizeof(double) * nz);
    orc_zetas(c->space, c->space_max, c->space_quantization_step, c->theta, rt->zetas, nz);
    rt->first_cooling_iteration = (uint64_t) floor(c->cooling_start * (double) c->iter_max);
}
static void tables_free(run_tables* rt) { free(rt->etas); free(rt->zetas); }

/* mode: 0 = 2D fp64, 1 = 2D fp32 model, 2 = 1D.
 * Runs iterations [iter_begin, iter_end) of the schedule with `updates` counted updates per iteration.  rng_state
 * (nullable, 4*n_streams words) carries the worker streams across calls; streams are seeded seed_base + t when no
 * state is supplied or the supplied state is still all-zero. */
static uint64_t run_streams_range(const orc_graph* g, const orc_config* c, uint64_t n_streams, uint64_t seed_base, uint64_t updates,
                                  uint64_t iter_begin, uint64_t iter_end, int mode, double* X, double* Y, float* xy,
                                  const uint8_t* frozen, uint64_t* rng_state) {
    if (!any_path_with_more_than_one_step(g)) return 0; /* path_sgd_layout.cpp:64-74 */
    run_tables rt;
    tables_init(&rt, c);
    orc_rng* rngs = (orc_rng*) malloc(sizeof(orc_rng) * n_streams);
    uint64_t* remaining = (uint64_t*) malloc(sizeof(uint64_t) * n_streams);
    /* an all-zero state is not a valid xoshiro state: it marks "not seeded yet" */
    if (!rng_state || (rng_state[0] | rng_state[1] | rng_state[2] | rng_state[3]) == 0) {
        for (uint64_t t = 0; t < n_streams; ++t) orc_rng_seed(&rngs[t], seed_base + t);
    } else {
        memcpy(rngs, rng_state, sizeof(orc_rng) * n_streams);
    }
    const int dims = mode == 2 ? 1 : 2;
    uint64_t n_iters = mode == 2 ? c->iter_max + 1 : c->iter_max;
    if (iter_end < n_iters) n_iters = iter_end;
    uint64_t counted = 0;
    for (uint64_t iter = iter_begin; iter < n_iters; ++iter) {
        const double eta = rt.etas[iter];
        int cooling;
        double theta_zipf = c->theta;
        if (mode == 2) {
            cooling = iter > rt.first_cooling_iteration;  /* path_sgd.cpp:194 */
            if (cooling) theta_zipf = 0.001;             /* path_sgd.cpp:195,246 */
        } else {
            cooling = iter >= rt.first_cooling_iteration; /* path_sgd_layout.cpp:153; adj_theta unused in 2D (:213) */
        }
        uint64_t base = updates / n_streams, rem = updates % n_streams;
        uint64_t live = 0;
        for (uint64_t t = 0; t < n_streams; ++t) { remaining[t] = base + (t < rem ? 1 : 0); live += remaining[t] != 0; }
        double delta_max = 0;
        while (live) {
            for (uint64_t t = 0; t < n_streams; ++t) {
                if (!remaining[t]) continue;
                orc_term term;
                if (!orc_sample_term(g, c, rt.zetas, dims, cooling, theta_zipf, &rngs[t], &term)) continue;
                double da;
                if (mode == 0) da = orc_apply_2d(&term, eta, X, Y);
                else if (mode == 1) da = orc_apply_2d_f32(&term, eta, xy);
                else { da = orc_apply_1d(&term, eta, X, frozen); if (da < 0) continue; }
                if (da > delta_max) delta_max = da;
                ++counted;
                if (--remaining[t] == 0) --live;
            }
        }
        /* early stop: checker_lambda path_sgd_layout.cpp:142 / path_sgd.cpp:183 (only reachable with delta > 0
         * in practice; Delta_max is re-armed to delta at every boundary :152) */
        if (c->delta > 0 && iter + 1 < n_iters && delta_max <= c->delta) break;
    }
    if (rng_state) memcpy(rng_state, rngs, sizeof(orc_rng) * n_streams);
    free(rngs);
    free(remaining);
    tables_free(&rt);
    return counted;
}

/* Tile-ORDER model of the device's tile sampling (pgsgd_tile_kernel), sequential and with the exact partner law: per
 * iteration every step is the first step of floor(U/S) terms (+1 for the steps of the leading tiles of a last, partial
 * pass); tiles of tile_steps consecutive steps are visited in a fresh pseudo-random bijection per pass, the steps of a tile
 * in order.  Isolates what the blocked ORDER of terms does to the result (no concurrency, no fp32 sampler).
 * mode 1 = 2D fp32, mode 2 = 1D. */
uint64_t orc_run_tile_order(const orc_graph* g, const orc_config* c, uint64_t tile_steps, int mode, float* xy, double* X) {
    if (!any_path_with_more_than_one_step(g) || tile_steps == 0) return 0;
    run_tables rt;
    tables_init(&rt, c);
    orc_rng rng, perm_rng;
    orc_rng_seed(&rng, c->seed);
    orc_rng_seed(&perm_rng, c->seed ^ 0x9e3779b97f4a7c15ULL);
    const int dims = mode == 2 ? 1 : 2;
    const uint64_t n_iters = mode == 2 ? c->iter_max + 1 : c->iter_max;
    const uint64_t S = g->step_count, n_tiles = (S + tile_steps - 1) / tile_steps;
    uint64_t* order = (uint64_t*) malloc(sizeof(uint64_t) * n_tiles);
    uint64_t counted = 0;
    for (uint64_t iter = 0; iter < n_iters; ++iter) {
        const double eta = rt.etas[iter];
        int cooling;
        double theta_zipf = c->theta;
        if (mode == 2) { cooling = iter > rt.first_cooling_iteration; if (cooling) theta_zipf = 0.001; }
        else cooling = iter >= rt.first_cooling_iteration;
        uint64_t left = c->min_term_updates;
        while (left) {
            for (uint64_t i = 0; i < n_tiles; ++i) order[i] = i;                       /* Fisher-Yates: one bijection per pass */
            for (uint64_t i = n_tiles - 1; i > 0; --i) { uint64_t j = orc_uniform(&perm_rng, i + 1), tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
            for (uint64_t k = 0; k < n_tiles && left; ++k) {
                const uint64_t lo = order[k] * tile_steps, hi = lo + tile_steps < S ? lo + tile_steps : S;
                for (uint64_t s = lo; s < hi && left; ++s) {
                    orc_term term;
                    --left;                                                            /* a visit slot is used whether or not it counts */
                    if (!sample_term_at(g, c, rt.zetas, dims, cooling, theta_zipf, &rng, s, &term)) continue;
                    if (mode == 1) orc_apply_2d_f32(&term, eta, xy);
                    else if (orc_apply_1d(&term, eta, X, NULL) < 0) continue;
                    ++counted;
                }
            }
        }
    }
    free(order);
    tables_free(&rt);
    return counted;
}

/* Hogwild staleness model (a planning tool for the launch-shape caps, DESIGN.md 3.4 — NOT a model of the reference): the
 * n_streams worker streams advance in waves; within a wave every stream reads the coordinates as they were when the wave
 * began, and the writes land together at its end — either summed (write = 0: the device's red.add) or last-writer-wins on
 * top of the stale value (write = 1: atomicExch / plain store).  n_streams terms in flight with the longest possible
 * read-to-write distance: an upper bound on what that many GPU threads in flight do.  mode 1 = 2D fp32, mode 2 = 1D. */
uint64_t orc_run_inflight(const orc_graph* g, const orc_config* c, uint64_t n_streams, int mode, int write, float* xy, double* X) {
    if (!any_path_with_more_than_one_step(g)) return 0;
    run_tables rt;
    tables_init(&rt, c);
    orc_rng* rngs = (orc_rng*) malloc(sizeof(orc_rng) * n_streams);
    for (uint64_t t = 0; t < n_streams; ++t) orc_rng_seed(&rngs[t], c->seed + t);
    uint64_t* ia = (uint64_t*) malloc(sizeof(uint64_t) * n_streams * 2);
    double* dv = (double*) malloc(sizeof(double) * n_streams * 4);   /* per term: rx, ry, stale a.x.., used per mode */
    double* stale = (double*) malloc(sizeof(double) * n_streams * 4);
    const int dims = mode == 2 ? 1 : 2;
    const uint64_t n_iters = mode == 2 ? c->iter_max + 1 : c->iter_max;
    uint64_t counted = 0;
    for (uint64_t iter = 0; iter < n_iters; ++iter) {
        const double eta = rt.etas[iter];
        int cooling;
        double theta_zipf = c->theta;
        if (mode == 2) { cooling = iter > rt.first_cooling_iteration; if (cooling) theta_zipf = 0.001; }
        else cooling = iter >= rt.first_cooling_iteration;
        uint64_t done = 0;
        while (done < c->min_term_updates) {
            uint64_t n = 0;
            for (uint64_t t = 0; t < n_streams && done + n < c->min_term_updates; ++t) {   /* reads of the wave */
                orc_term term;
                if (!orc_sample_term(g, c, rt.zetas, dims, cooling, theta_zipf, &rngs[t], &term)) continue;
                if (mode == 2) {
                    double d = fabs((double) term.pos_a - (double) term.pos_b);
                    if (d == 0) continue;
                    double mu = eta / d; if (mu > 1) mu = 1;
                    double xa = X[term.node_a], xb = X[term.node_b];
                    double dx = xa - xb; if (dx == 0) dx = 1e-9;
                    double mag = fabs(dx);
                    double r_x = mu * (mag - d) / 2 / mag * dx;
                    ia[2 * n] = term.node_a; ia[2 * n + 1] = term.node_b;
                    dv[4 * n] = r_x; stale[4 * n] = xa; stale[4 * n + 1] = xb;
                } else {
                    uint64_t dpos = term.pos_a > term.pos_b ? term.pos_a - term.pos_b : term.pos_b - term.pos_a;
                    float d = dpos ? (float) dpos : 1e-9f;
                    float mu = (float) eta / d; if (mu > 1.0f) mu = 1.0f;
                    uint64_t pa = 4 * (uint64_t) term.node_a + 2 * (term.end_a ? 1 : 0), pb = 4 * (uint64_t) term.node_b + 2 * (term.end_b ? 1 : 0);
                    float dx = xy[pa] - xy[pb], dy = xy[pa + 1] - xy[pb + 1];
                    if (dx == 0.0f) dx = 1e-9f;
                    float mag = sqrtf(dx * dx + dy * dy);
                    float r = mu * (mag - d) * 0.5f / mag;
                    ia[2 * n] = pa; ia[2 * n + 1] = pb;
                    dv[4 * n] = r * dx; dv[4 * n + 1] = r * dy;
                    stale[4 * n] = xy[pa]; stale[4 * n + 1] = xy[pa + 1]; stale[4 * n + 2] = xy[pb]; stale[4 * n + 3] = xy[pb + 1];
                }
                ++n;
            }
            for (uint64_t k = 0; k < n; ++k) {   /* writes of the wave */
                if (mode == 2) {
                    if (write == 0) { X[ia[2 * k]] -= dv[4 * k]; X[ia[2 * k + 1]] += dv[4 * k]; }
                    else { X[ia[2 * k]] = stale[4 * k] - dv[4 * k]; X[ia[2 * k + 1]] = stale[4 * k + 1] + dv[4 * k]; }
                } else {
                    const uint64_t pa = ia[2 * k], pb = ia[2 * k + 1];
                    if (write == 0) {
                        xy[pa] -= (float) dv[4 * k]; xy[pa + 1] -= (float) dv[4 * k + 1];
                        xy[pb] += (float) dv[4 * k]; xy[pb + 1] += (float) dv[4 * k + 1];
                    } else {
                        xy[pa] = (float) (stale[4 * k] - dv[4 * k]); xy[pa + 1] = (float) (stale[4 * k + 1] - dv[4 * k + 1]);
                        xy[pb] = (float) (stale[4 * k + 2] + dv[4 * k]); xy[pb + 1] = (float) (stale[4 * k + 3] + dv[4 * k + 1]);
                    }
                }
            }
            done += n;
            counted += n;
            if (n == 0) break;
        }
    }
    free(rngs); free(ia); free(dv); free(stale);
    tables_free(&rt);
    return counted;
}

static uint64_t run_streams(const orc_graph* g, const orc_config* c, uint64_t n_streams, int mode, double* X, double* Y,
                            float* xy, const uint8_t* frozen) {
    return run_streams_range(g, c, n_streams, c->seed, c->min_term_updates, 0, UINT64_MAX, mode, X, Y, xy, frozen, NULL);
}

uint64_t orc_run_range(const orc_graph* g, const orc_config* c, uint64_t n_streams, uint64_t seed_base, uint64_t updates,
                       uint64_t iter_begin, uint64_t iter_end, int mode, double* X, double* Y, float* xy,
                       const uint8_t* frozen, uint64_t* rng_state) {
    return run_streams_range(g, c, n_streams, seed_base, updates, iter_begin, iter_end, mode, X, Y, xy, frozen, rng_state);
}

uint64_t orc_layout_2d(const orc_graph* g, const orc_config* c, uint64_t n_streams, double* X, double* Y) {
    return run_streams(g, c, n_streams, 0, X, Y, NULL, NULL);
}
uint64_t orc_layout_2d_f32(const orc_graph* g, const orc_config* c, uint64_t n_streams, float* xy) {
    return run_streams(g, c, n_streams, 1, NULL, NULL, xy, NULL);
}
uint64_t orc_sort_1d(const orc_graph* g, const orc_config* c, uint64_t n_streams, const uint8_t* frozen, double* X) {
    return run_streams(g, c, n_streams, 2, X, NULL, NULL, frozen);
}

/* Emulation of the multi-GPU "peer" schedule with STALE REMOTE READS: the coordinate array is partitioned by node range
 * over n_ranks owners; every update is applied to the one true array (as the NVLink red.add does), but a rank reads the
 * coordinates of nodes it does not own from a snapshot that is refreshed `refreshes` times per iteration.  Rank r only
 * draws terms whose first node it owns (tile ownership).  2D fp32 model; iterations [iter_begin, iter_end). */
uint64_t orc_peer_stale_2d_f32(const orc_graph* g, const orc_config* c, uint64_t n_ranks, uint64_t streams_per_rank, uint64_t refreshes,
                               uint64_t iter_begin, uint64_t iter_end, float* xy) {
    run_tables rt;
    tables_init(&rt, c);
    const uint64_t N = g->node_count, chunk = (N + n_ranks - 1) / n_ranks, T = n_ranks * streams_per_rank;
    orc_rng* rngs = (orc_rng*) malloc(sizeof(orc_rng) * T);
    for (uint64_t t = 0; t < T; ++t) orc_rng_seed(&rngs[t], c->seed + 0x51ED27 * (iter_begin + 1) + t);
    float* snap = (float*) malloc(sizeof(float) * 4 * N);
    uint64_t counted = 0;
    uint64_t n_iters = c->iter_max < iter_end ? c->iter_max : iter_end;
    for (uint64_t iter = iter_begin; iter < n_iters; ++iter) {
        const double eta = rt.etas[iter];
        const int cooling = iter >= rt.first_cooling_iteration;
        for (uint64_t k = 0; k < refreshes; ++k) {
            memcpy(snap, xy, sizeof(float) * 4 * N);
            const uint64_t slice = c->min_term_updates / refreshes;
            uint64_t left = slice;
            while (left) {
                for (uint64_t t = 0; t < T && left; ++t) {
                    const uint64_t r = t % n_ranks;   /* ranks interleaved: concurrent ranks have no order among themselves */
                    orc_term term;
                    if (!orc_sample_term(g, c, rt.zetas, 2, cooling, c->theta, &rngs[t], &term)) continue;
                    uint64_t oa = term.node_a / chunk, ob = term.node_b / chunk;
                    if (oa != r) continue;  /* not this rank's tile: redraw */
                    /* the update as the device computes it: a from the true array (own), b from the true array when owned
                     * else from the snapshot; displacement added to the true array for both */
                    float* pa = xy + 4 * (uint64_t) term.node_a + 2 * (term.end_a ? 1 : 0);
                    float* pb_true = xy + 4 * (uint64_t) term.node_b + 2 * (term.end_b ? 1 : 0);
                    const float* pb_read = ob == r ? pb_true : snap + 4 * (uint64_t) term.node_b + 2 * (term.end_b ? 1 : 0);
                    uint64_t dpos = term.pos_a > term.pos_b ? term.pos_a - term.pos_b : term.pos_b - term.pos_a;
                    float d_ij = dpos ? (float) dpos : 1e-9f;
                    float mu = (float) eta / d_ij;
                    if (mu > 1.0f) mu = 1.0f;
                    float dx = pa[0] - pb_read[0], dy = pa[1] - pb_read[1];
                    if (dx == 0.0f) dx = 1e-9f;
                    float mag = sqrtf(dx * dx + dy * dy);
                    float Delta = mu * (mag - d_ij) * 0.5f;
                    float rr = Delta / mag, r_x = rr * dx, r_y = rr * dy;
                    pa[0] -= r_x; pa[1] -= r_y;
                    pb_true[0] += r_x; pb_true[1] += r_y;
                    ++counted;
                    --left;
                }
            }
        }
    }
    free(snap);
    free(rngs);
    tables_free(&rt);
    return counted;
}

uint64_t orc_replay_single(const orc_graph* g, const orc_config* c, int dims, uint64_t n_terms, uint64_t switch_at,
                           double eta0, double eta1, int cooling0, int cooling1, double theta1,
                           double* X, double* Y, orc_term* out_terms) {
    return orc_replay_single_frozen(g, c, dims, n_terms, switch_at, eta0, eta1, cooling0, cooling1, theta1, X, Y, out_terms, NULL);
}

uint64_t orc_replay_single_frozen(const orc_graph* g, const orc_config* c, int dims, uint64_t n_terms, uint64_t switch_at,
                                  double eta0, double eta1, int cooling0, int cooling1, double theta1,
                                  double* X, double* Y, orc_term* out_terms, const uint8_t* frozen) {
    run_tables rt;
    tables_init(&rt, c);
    orc_rng rng;
    orc_rng_seed(&rng, c->seed);
    uint64_t emitted = 0;
    while (emitted < n_terms) {
        int sw = emitted >= switch_at;
        double eta = sw ? eta1 : eta0;
        int cooling = sw ? cooling1 : cooling0;
        double theta_zipf = sw ? theta1 : c->theta;
        orc_term term;
        if (!orc_sample_term(g, c, rt.zetas, dims, cooling, theta_zipf, &rng, &term)) continue;
        if (dims == 2) {
            if (X) orc_apply_2d(&term, eta, X, Y);
        } else {
            /* the reference prints its trace line after the d == 0 `continue` (path_sgd.cpp:320-327):
             * skipped terms are not emitted */
            /* ... and after the both-frozen `continue` (path_sgd.cpp:298-302), which comes first */
            if (frozen && frozen[term.node_a] && frozen[term.node_b]) continue;
            if (term.pos_a == term.pos_b) continue;
            if (X) orc_apply_1d(&term, eta, X, frozen);
        }
        if (out_terms) out_terms[emitted] = term;
        ++emitted;
    }
    tables_free(&rt);
    return emitted;
}

/* ------------------------------------------------------------------------------------------------
 * sampled path stress (our definition)
 * ---------------------------------------------------------------------------------------------- */

/* The K pairs are drawn by ORC_STRESS_STREAMS independent generators (stream t seeded seed + t, ceil(K / streams)
 * pairs each) so that a GPU can evaluate the same definition with one thread per stream. */
#define ORC_STRESS_STREAMS 4096

double orc_path_stress_2d(const orc_graph* g, const double* X, const double* Y, uint64_t n_pairs, uint64_t seed) {
    const uint64_t per = (n_pairs + ORC_STRESS_STREAMS - 1) / ORC_STRESS_STREAMS;
    double acc = 0;
    uint64_t used = 0;
    for (uint64_t t = 0; t < ORC_STRESS_STREAMS; ++t) {
        orc_rng rng;
        orc_rng_seed(&rng, seed + t);
        double a = 0;
        for (uint64_t k = 0; k < per; ++k) {
            uint64_t ia = orc_uniform(&rng, g->step_count);
            uint64_t p = find_path(g, ia);
            uint64_t first = g->path_first_step[p], cnt = g->path_first_step[p + 1] - first;
            uint64_t ib = first + orc_uniform(&rng, cnt);
            uint64_t fa = orc_uniform(&rng, 2), fb = orc_uniform(&rng, 2);
            uint32_t na = g->step_node[ia], nb = g->step_node[ib];
            uint64_t pa = g->step_pos[ia] + (fa ? g->node_len[na] : 0);
            uint64_t pb = g->step_pos[ib] + (fb ? g->node_len[nb] : 0);
            uint64_t ea = fa ? !g->step_rev[ia] : g->step_rev[ia];
            uint64_t eb = fb ? !g->step_rev[ib] : g->step_rev[ib];
            if (pa == pb) continue;
            double d = fabs((double) pa - (double) pb);
            double dx = X[2 * (uint64_t) na + ea] - X[2 * (uint64_t) nb + eb];
            double dy = Y[2 * (uint64_t) na + ea] - Y[2 * (uint64_t) nb + eb];
            double e = (sqrt(dx * dx + dy * dy) - d) / d;
            a += e * e;
            ++used;
        }
        acc += a;
    }
    return used ? acc / (double) used : 0.0;
}

double orc_path_stress_1d(const orc_graph* g, const double* X, uint64_t n_pairs, uint64_t seed) {
    const uint64_t per = (n_pairs + ORC_STRESS_STREAMS - 1) / ORC_STRESS_STREAMS;
    double acc = 0;
    uint64_t used = 0;
    for (uint64_t t = 0; t < ORC_STRESS_STREAMS; ++t) {
        orc_rng rng;
        orc_rng_seed(&rng, seed + t);
        double a = 0;
        for (uint64_t k = 0; k < per; ++k) {
            uint64_t ia = orc_uniform(&rng, g->step_count);
            uint64_t p = find_path(g, ia);
            uint64_t first = g->path_first_step[p], cnt = g->path_first_step[p + 1] - first;
            uint64_t ib = first + orc_uniform(&rng, cnt);
            uint64_t pa = g->step_pos[ia], pb = g->step_pos[ib];
            if (pa == pb) continue;
            double d = fabs((double) pa - (double) pb);
            double e = (fabs(X[g->step_node[ia]] - X[g->step_node[ib]]) - d) / d;
            a += e * e;
            ++used;
        }
        acc += a;
    }
    return used ? acc / (double) used : 0.0;
}

/* Local path stress: the same estimator restricted to NEAR pairs — partner = the step `j` ranks away along the
 * path, j uniform in [1, ORC_LOCAL_WINDOW], random direction, pairs further apart than ORC_LOCAL_MAX_BP skipped.  The
 * far-pair stress above is dominated by pairs megabases apart and is blind to the fine structure of the layout (where
 * limited coordinate precision would show first); this one measures exactly that.  Same stream structure. */
#define ORC_LOCAL_WINDOW 64
#define ORC_LOCAL_MAX_BP 1000

double orc_local_stress_2d(const orc_graph* g, const double* X, const double* Y, uint64_t n_pairs, uint64_t seed) {
    const uint64_t per = (n_pairs + ORC_STRESS_STREAMS - 1) / ORC_STRESS_STREAMS;
    double acc = 0;
    uint64_t used = 0;
    for (uint64_t t = 0; t < ORC_STRESS_STREAMS; ++t) {
        orc_rng rng;
        orc_rng_seed(&rng, seed + t);
        double a = 0;
        for (uint64_t k = 0; k < per; ++k) {
            uint64_t ia = orc_uniform(&rng, g->step_count);
            uint64_t p = find_path(g, ia);
            uint64_t first = g->path_first_step[p], cnt = g->path_first_step[p + 1] - first;
            uint64_t j = 1 + orc_uniform(&rng, ORC_LOCAL_WINDOW);
            uint64_t back = orc_uniform(&rng, 2);
            uint64_t fa = orc_uniform(&rng, 2), fb = orc_uniform(&rng, 2);
            uint64_t ra = ia - first;
            if (back ? ra < j : ra + j >= cnt) continue;
            uint64_t ib = back ? ia - j : ia + j;
            uint32_t na = g->step_node[ia], nb = g->step_node[ib];
            uint64_t pa = g->step_pos[ia] + (fa ? g->node_len[na] : 0);
            uint64_t pb = g->step_pos[ib] + (fb ? g->node_len[nb] : 0);
            uint64_t ea = fa ? !g->step_rev[ia] : g->step_rev[ia];
            uint64_t eb = fb ? !g->step_rev[ib] : g->step_rev[ib];
            if (pa == pb) continue;
            double d = fabs((double) pa - (double) pb);
            if (d > (double) ORC_LOCAL_MAX_BP) continue;
            double dx = X[2 * (uint64_t) na + ea] - X[2 * (uint64_t) nb + eb];
            double dy = Y[2 * (uint64_t) na + ea] - Y[2 * (uint64_t) nb + eb];
            double e = (sqrt(dx * dx + dy * dy) - d) / d;
            a += e * e;
            ++used;
        }
        acc += a;
    }
    return used ? acc / (double) used : 0.0;
}

double orc_local_stress_1d(const orc_graph* g, const double* X, uint64_t n_pairs, uint64_t seed) {
    const uint64_t per = (n_pairs + ORC_STRESS_STREAMS - 1) / ORC_STRESS_STREAMS;
    double acc = 0;
    uint64_t used = 0;
    for (uint64_t t = 0; t < ORC_STRESS_STREAMS; ++t) {
        orc_rng rng;
        orc_rng_seed(&rng, seed + t);
        double a = 0;
        for (uint64_t k = 0; k < per; ++k) {
            uint64_t ia = orc_uniform(&rng, g->step_count);
            uint64_t p = find_path(g, ia);
            uint64_t first = g->path_first_step[p], cnt = g->path_first_step[p + 1] - first;
            uint64_t j = 1 + orc_uniform(&rng, ORC_LOCAL_WINDOW);
            uint64_t back = orc_uniform(&rng, 2);
            uint64_t ra = ia - first;
            if (back ? ra < j : ra + j >= cnt) continue;
            uint64_t ib = back ? ia - j : ia + j;
            uint64_t pa = g->step_pos[ia], pb = g->step_pos[ib];
            if (pa == pb) continue;
            double d = fabs((double) pa - (double) pb);
            if (d > (double) ORC_LOCAL_MAX_BP) continue;
            double e = (fabs(X[g->step_node[ia]] - X[g->step_node[ib]]) - d) / d;
            a += e * e;
            ++used;
        }
        acc += a;
    }
    return used ? acc / (double) used : 0.0;
}
