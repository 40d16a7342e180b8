/* pgsgd_oracle.h — CPU restatement of odgi's path-guided SGD (2D layout + 1D sort).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this.  The product (odgi_b200/) never does.
 *
 * Parity status: PINNED against the reference itself — scripts/pin_oracle.py replays the traces and
 * final coordinates produced by oracle/_ref/ref_driver_trace (the unmodified reference compiled with
 * its own -Deval_path_sgd hook) bit-for-bit; see tests/test_oracle_pinned.py and tests/golden/.
 *
 * Every function cites the reference file:line (relative to the odgi tree) it restates.
 */
#ifndef PGSGD_ORACLE_H
#define PGSGD_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG: deps/Xoshiro-cpp/XoshiroCpp.hpp:684-746 ---- */
typedef struct { uint64_t s[4]; } orc_rng;
void orc_rng_seed(orc_rng* g, uint64_t seed);            /* Xoshiro256Plus(seed): 4 SplitMix64 outputs */
uint64_t orc_rng_next(orc_rng* g);                       /* Xoshiro256Plus::operator() */
/* std::uniform_int_distribution<uint64_t>(0, range-1) on a 64-bit URBG: libstdc++ 13
 * bits/uniform_int_dist.h:252-283 (Lemire, 128-bit product, rejection) */
uint64_t orc_uniform(orc_rng* g, uint64_t range);
/* std::generate_canonical<double,53>: libstdc++ 13 bits/random.tcc:3349-3381 */
double orc_canonical(orc_rng* g);

/* ---- dirty zipf: deps/dirtyzipf/dirty_zipfian_int_distribution.h ---- */
double orc_fast_precise_pow(double a, double b);         /* :82-104 */
double orc_zeta(uint64_t n, double theta);               /* :165-171 */
/* operator()(urng, param_type(1, n, theta, zeta_n)) :230-243; zeta2 = zeta(2, theta) as the
 * param_type constructor computes it (:126-128) */
uint64_t orc_dirty_zipf(orc_rng* g, uint64_t n, double theta, double zeta_n);

/* ---- schedule and zeta table ---- */
/* path_linear_sgd_layout_schedule / path_linear_sgd_schedule (path_sgd_layout.cpp:433-468,
 * path_sgd.cpp:466-501): writes iter_max+1 etas */
void orc_schedule(double eta_max, uint64_t iter_max, uint64_t iter_with_max_learning_rate, double eps, double* etas);
/* zeta table (path_sgd_layout.cpp:87-97, path_sgd.cpp:128-138); returns the entry count,
 * writes at most cap entries when zetas != NULL */
uint64_t orc_zetas(uint64_t space, uint64_t space_max, uint64_t space_q, double theta, double* zetas, uint64_t cap);

/* ---- flattened graph (path-major).  step_perm (nullable) maps the sampler's step_index to a
 * path-major step index: NULL = identity (what the CUDA path uses); XP's node-major table
 * (xp.cpp:143-144 npi_iv/nr_iv) reproduces the reference's own draw order. ---- */
typedef struct {
    uint64_t node_count, path_count, step_count;
    const uint32_t* node_len;        /* [N] */
    const uint64_t* path_first_step; /* [P+1] */
    const uint32_t* step_node;       /* [S] node rank = id-1 */
    const uint8_t*  step_rev;        /* [S] */
    const uint64_t* step_pos;        /* [S] 0-based bp offset of the step's node start in its path */
    const uint64_t* step_perm;       /* [S] or NULL */
} orc_graph;

typedef struct {
    uint64_t iter_max;
    uint64_t iter_with_max_learning_rate;
    uint64_t min_term_updates;       /* U: term updates per iteration */
    double delta, eps, eta_max, theta;
    uint64_t space, space_max, space_quantization_step;
    double cooling_start;
    uint64_t seed;                   /* stream t is seeded with seed + t (reference: 9399220 + tid) */
} orc_config;

/* one sampled term (everything the integer half of the hot path produces) */
typedef struct {
    uint64_t step_index;             /* raw draw in [0,S) */
    uint64_t path;                   /* 0-based path index */
    uint64_t rank_a, rank_b;         /* step ranks within the path */
    uint32_t node_a, node_b;         /* node ranks */
    uint8_t rev_a, rev_b;            /* orientation of the steps */
    uint8_t flip_a, flip_b;          /* raw coin flips for the end choice (2D only) */
    uint8_t end_a, end_b;            /* which end of the node is moved: 0 = start (+0), 1 = end (+1) (2D) */
    uint64_t pos_a, pos_b;           /* end-adjusted bp positions in the path */
    uint8_t zipf;                    /* 1 if the Zipf branch was taken */
} orc_term;

/* Draw one term exactly as a CPU worker does (2D: path_sgd_layout.cpp:182-269; 1D: path_sgd.cpp:222-306).
 * Returns 1 with *t filled, or 0 if the draw hit a 1-step path (the reference `continue`s without
 * counting, path_sgd_layout.cpp:190-192). theta_zipf is the theta given to the Zipf draw (1D cooling
 * passes 0.001 with zetas of the original theta, path_sgd.cpp:195,246). */
int orc_sample_term(const orc_graph* g, const orc_config* c, const double* zetas, int dims, int cooling,
                    double theta_zipf, orc_rng* rng, orc_term* t);

/* Apply one term in fp64 exactly as the reference (2D: path_sgd_layout.cpp:280-363 on X/Y indexed
 * 2*node+end; 1D: path_sgd.cpp:317-392 on X indexed by node).  Returns |Delta|; for 1D returns -1 and
 * leaves X untouched when term_dist == 0 (path_sgd.cpp:320-323, not counted). */
double orc_apply_2d(const orc_term* t, double eta, double* X, double* Y);
double orc_apply_1d(const orc_term* t, double eta, double* X, const uint8_t* frozen);

/* Device arithmetic model of the 2D update: fp32 coordinates, fp32 math, no FMA contraction, IEEE
 * sqrt/div — the arithmetic odgi_b200/csrc implements.  xy is float[4*N] = {x0,y0,x1,y1} per node. */
float orc_apply_2d_f32(const orc_term* t, double eta, float* xy);

/* Whole runs with deterministic iteration boundaries (exactly U counted updates per iteration, the
 * semantics of the GPU paths; reference GPU: layout.cu:442-447).  n_streams independent worker streams
 * (stream t seeded seed+t) are interleaved round-robin one term at a time; stream t performs
 * U/n_streams (+1 for t < U % n_streams) counted updates per iteration.
 *   2D: iterations 0..iter_max-1, cooling when iter >= floor(cooling_start*iter_max) (path_sgd_layout.cpp:140,153)
 *   1D: iterations 0..iter_max,   cooling when iter >  floor(cooling_start*iter_max) (path_sgd.cpp:181,194)
 * Returns the number of counted updates. */
uint64_t orc_layout_2d(const orc_graph* g, const orc_config* c, uint64_t n_streams, double* X, double* Y);
uint64_t orc_layout_2d_f32(const orc_graph* g, const orc_config* c, uint64_t n_streams, float* xy);
uint64_t orc_sort_1d(const orc_graph* g, const orc_config* c, uint64_t n_streams, const uint8_t* frozen, double* X);

/* Iterations [iter_begin, iter_end) with an explicit per-iteration update count and stream seed base: the building
 * block used by the tests to emulate the multi-GPU schedule (each rank: its share of the updates with its own streams,
 * then the coordinates are averaged across ranks).  mode: 0 = 2D fp64, 1 = 2D fp32 model, 2 = 1D.  rng_state
 * (4*n_streams words, nullable) carries the streams across calls. */
uint64_t orc_run_range(const orc_graph* g, const orc_config* c, uint64_t n_streams, uint64_t seed_base, uint64_t updates,
                       uint64_t iter_begin, uint64_t iter_end, int mode, double* X, double* Y, float* xy,
                       const uint8_t* frozen, uint64_t* rng_state);

/* Emulation of the multi-GPU peer schedule with stale remote reads (see the .c file); experiments only. */
uint64_t orc_peer_stale_2d_f32(const orc_graph* g, const orc_config* c, uint64_t n_ranks, uint64_t streams_per_rank, uint64_t refreshes,
                               uint64_t iter_begin, uint64_t iter_end, float* xy);

/* Replay helper for pinning: runs ONE stream (seed+0) for n_terms emitted terms with the cooling flag /
 * eta switching at emitted-term index switch_at (use n_terms for "never"), applying fp64 updates;
 * optionally records the terms.  This is what scripts/pin_oracle.py aligns with the reference trace. */
uint64_t orc_replay_single(const orc_graph* g, const orc_config* c, int dims, uint64_t n_terms, uint64_t switch_at,
                           double eta0, double eta1, int cooling0, int cooling1, double theta1,
                           double* X, double* Y, orc_term* out_terms /* nullable, n_terms */);

/* the same with frozen nodes (1D, `odgi sort -H`): terms with both nodes frozen are counted but not traced */
uint64_t orc_replay_single_frozen(const orc_graph* g, const orc_config* c, int dims, uint64_t n_terms, uint64_t switch_at,
                                  double eta0, double eta1, int cooling0, int cooling1, double theta1,
                                  double* X, double* Y, orc_term* out_terms, const uint8_t* frozen);

/* Sampled path stress (our definition — the reference has none).  K pairs drawn with a
 * fixed seed: step uniform over all steps (=> path ∝ step count), partner uniform in the same path,
 * ends uniform (2D) / node starts (1D); d = |pos_a - pos_b| in bp, d == 0 skipped;
 * stress = mean(((|p_a - p_b| - d)/d)^2).  2D coords: X/Y indexed 2*node+end; 1D: X by node. */
double orc_path_stress_2d(const orc_graph* g, const double* X, const double* Y, uint64_t n_pairs, uint64_t seed);
double orc_path_stress_1d(const orc_graph* g, const double* X, uint64_t n_pairs, uint64_t seed);
/* Local variant: partner 1..64 ranks away along the path, pairs further apart than 1000 bp skipped (fine structure). */
double orc_local_stress_2d(const orc_graph* g, const double* X, const double* Y, uint64_t n_pairs, uint64_t seed);
double orc_local_stress_1d(const orc_graph* g, const double* X, uint64_t n_pairs, uint64_t seed);

#ifdef __cplusplus
}
#endif
/* tile-ORDER model of the device's tile sampling, sequential, exact partner law (isolates the effect of the blocked order) */
uint64_t orc_run_tile_order(const orc_graph* g, const orc_config* c, uint64_t tile_steps, int mode, float* xy, double* X);

/* Hogwild staleness model: waves of n_streams terms that all read before any of them writes (planning tool) */
uint64_t orc_run_inflight(const orc_graph* g, const orc_config* c, uint64_t n_streams, int mode, int write, float* xy, double* X);

#endif
