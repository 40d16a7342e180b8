/* pgsgd.h — C-ABI of the B200-native path-guided SGD (PG-SGD) engine.
 *
 * This is the drop-in boundary for odgi's GPU layout path.  Every entry point is `extern "C"` with
 * plain pointers and sizes; the reference interface each one replaces is cited (paths relative to the
 * odgi tree).  The odgi-side C++ shim that calls these (same signatures as the reference's
 * path_linear_sgd_layout_gpu / a new path_linear_sgd_gpu) is odgi_b200/host/odgi_shim.cpp; the
 * binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Memory: host arrays are caller-owned.  Device memory is owned by the opaque engine handle.
 * Errors: every function returns PGSGD_OK (0) or a negative status; pgsgd_last_error() gives the
 * message (thread-local).  There is NO CPU fallback: without a usable CUDA device every compute entry
 * point fails with PGSGD_ERR_CUDA.
 */
#ifndef PGSGD_H
#define PGSGD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bumped whenever a struct below changes size or meaning.  Callers compare PGSGD_VERSION (what they were compiled against)
 * with pgsgd_version() (what they loaded) before the first call: a stale binary would otherwise pass short structs. */
#define PGSGD_VERSION 105 /* 0.1.4 */

typedef enum pgsgd_status {
    PGSGD_OK = 0,
    PGSGD_ERR_ARG = -1,      /* bad argument / inconsistent graph view */
    PGSGD_ERR_CUDA = -2,     /* CUDA runtime error (incl. "no device") */
    PGSGD_ERR_NCCL = -3,     /* NCCL error */
    PGSGD_ERR_NOMEM = -4,    /* host or device allocation failed */
    PGSGD_ERR_STATE = -5,    /* call sequence error (e.g. run before coordinates were set) */
    PGSGD_ERR_UNOPT = -6     /* graph ids are not compacted 1..N (reference: src/cuda/layout.cu:320-323) */
} pgsgd_status;

/* Flattened, path-major view of an odgi graph.  Replaces what cuda::gpu_layout builds for itself in
 * managed memory (src/cuda/layout.cu:325-410: node_t[], path_t[], path_element_t[]) and what the CPU
 * workers read through xp::XP (src/algorithms/xp.cpp:375-436).
 *   node rank   = node id - 1                              (src/odgi.cpp:35-42, graph must be optimized)
 *   step s of path p lives at index path_first_step[p] + rank_in_path
 *   step_pos    = 0-based bp offset of the step's node start in its path == XP get_position_of_step
 *                 (xp.cpp:393-397); may be NULL, then it is derived from node_len (xp.cpp:607-616)
 *   step_rev    = 1 when the step traverses the node in reverse (handle is_reverse); may be NULL (all forward)
 */
typedef struct pgsgd_graph_view {
    uint64_t node_count;              /* N */
    uint64_t path_count;              /* P */
    uint64_t step_count;              /* S = path_first_step[P] */
    const uint32_t* node_len;         /* [N] sequence length of each node, bp */
    const uint64_t* path_first_step;  /* [P+1] prefix sums of per-path step counts */
    const uint32_t* step_node;        /* [S] node rank of each step */
    const uint8_t*  step_rev;         /* [S] or NULL */
    const uint64_t* step_pos;         /* [S] or NULL */
} pgsgd_graph_view;

/* PG-SGD parameters: the argument list of algorithms::path_linear_sgd_layout[_gpu]
 * (src/algorithms/path_sgd_layout.hpp:37-79) / path_linear_sgd (src/algorithms/path_sgd.cpp:12-31),
 * i.e. a superset of cuda::layout_config_t (src/cuda/layout.h:65-77). */
typedef struct pgsgd_config {
    uint64_t iter_max;                     /* 2D runs iter_max iterations, 1D runs iter_max+1 (path_sgd.cpp:181) */
    uint64_t iter_with_max_learning_rate;
    uint64_t min_term_updates;             /* term updates per iteration (whole job, all GPUs) */
    double   delta;                        /* early stop threshold on max |Delta| per iteration; 0 = never */
    double   eps;
    double   eta_max;
    double   theta;                        /* Zipf exponent */
    uint64_t space;                        /* max Zipf jump, in steps */
    uint64_t space_max;
    uint64_t space_quantization_step;
    double   cooling_start;                /* first cooling iteration = floor(cooling_start * iter_max) */
    uint64_t seed;                         /* worker stream t is seeded seed + t; reference: 9399220 (path_sgd_layout.cpp:168) */
    uint32_t n_streams;                    /* device worker streams (one per GPU thread); 0 = automatic: fill the GPU, but keep
                                              the terms in flight (n_streams * batch) below node_count / 2 (/ 4 with the
                                              racy write flavours) — beyond that Hogwild staleness measurably worsens the
                                              layout of small graphs */
    uint32_t batch;                        /* terms a stream keeps in flight (1, 2 or 4); 0 = default (1).
                                              batch 1 applies a stream's terms strictly in order */
    uint32_t flags;                        /* PGSGD_FLAG_* */
    uint32_t sampling;                     /* PGSGD_SAMPLING_*: how the first step of a term is chosen */
    uint64_t multi_switch_iteration;       /* PGSGD_MULTI_HYBRID: first iteration of the peer phase; 0 = iter_max / 3 */
} pgsgd_config;

/* First-step sampling.  The partner of a term is always drawn by the reference's rule.
 *  STREAM: every worker stream draws its first step i.i.d. uniform over all steps, exactly like a reference worker
 *          thread (path_sgd_layout.cpp:175-182) — the device streams are bit-identical to CPU worker threads with the
 *          same seeds.  Two random 16-byte HBM reads per term.
 *  TILE  : the step array is cut into tiles of 2048 consecutive steps; a CTA stages a tile in shared memory with
 *          coalesced 128-bit loads and uses each of its steps once as a first step; tile visits follow per-pass
 *          bijections, so over an iteration every step is a first step exactly floor(U/S) (+1) times — the same
 *          uniform marginal with the first pick's sampling noise removed.  ~1 random HBM read per term or fewer.
 *  AUTO  : TILE for graphs whose step records exceed the L2 (>= 2^22 steps), that are at least PGSGD_AUTO_TILE_MIN_DEPTH steps
 *          deep per node on average (haplotype depth) and large enough for the in-flight cap, else STREAM.  (Shallow graphs keep
 *          the reference-exact sampler as the conservative choice; over 20 far-apart seeds the two samplers end in the same
 *          distribution of final stress on a 6-haplotype graph: DESIGN.md 5.4.) */
#define PGSGD_AUTO_TILE_MIN_DEPTH 8ull
#define PGSGD_SAMPLING_AUTO   0u
#define PGSGD_SAMPLING_STREAM 1u
#define PGSGD_SAMPLING_TILE   2u

/* Coordinate write flavour.  Default (no flag): red.global.add of the displacement — no update is ever lost; a single
 * worker stream gives bit-identical results to the load/compute/store of the reference (tests/test_gpu_parity.py).
 * The other two reproduce the reference's racy last-writer-wins write (path_sgd_layout.cpp:360-363, layout.cu:184-187). */
#define PGSGD_FLAG_EXCH_WRITE   1u  /* 64-bit atom.exch of the new (x,y) — the reference CUDA kernel's atomicExch semantics */
#define PGSGD_FLAG_PLAIN_STORE  4u  /* st.global of the new (x,y) (slower on B200: 14 vs 21 G updates/s, profiles/) */
#define PGSGD_FLAG_TMA_STAGING  8u  /* tile kernels: stage tiles with double-buffered TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP)
                                       instead of coalesced LDG.128/STS.128 into one buffer (DESIGN.md 3.2 has the measurements) */
#define PGSGD_FLAG_KEEP_ADD    16u  /* experiments only: keep the red.add write even where the hub safeguard would switch to the
                                       exchange write — used to measure where the summed write really turns unstable */
#define PGSGD_FLAG_LEGACY_TILE 32u  /* tile sampling with round 1's unpipelined kernel (pgsgd_tile_kernel) instead of the pipelined one
                                       (pgsgd_tile2_kernel, pgsgd_tile2.cu); NVLink peer phases always use the legacy kernel */
#define PGSGD_FLAG_HALF_TILE   64u  /* pipelined tile kernel: 1024-step tiles (with TMA_STAGING: two 1024-step buffers = the footprint of
                                       one 2048-step buffer) */
#define PGSGD_FLAG_BIG_TILE   128u  /* pipelined tile kernel: 4096-step tiles (experiments: occupancy vs in-tile partner rate) */
#define PGSGD_FLAG_SWEEP_TILES 256u /* pipelined tile kernel: every pass walks the tiles in path order from a random offset (tile = (i + add) mod
                                       n_tiles) instead of a random bijection: the CTAs resident at any time then work on ONE contiguous window of the
                                       step array, whose records and coordinates stay in L2 — partners a megabase away are L2 hits, not random DRAM
                                       sectors (the measured wall: ~69 G DRAM sectors/s) */
#define PGSGD_FLAG_L2_WINDOW  512u /* coordinates pinned in L2 by a persisting access-policy window on the engine's stream instead of
                                       per-instruction evict_last hints */
#define PGSGD_FLAG_COORD_LD_FIRST  1024u /* experiments: coordinate loads with L2::evict_first (reds keep evict_last) */
#define PGSGD_FLAG_COORD_LD_NORMAL 2048u /* experiments: coordinate loads with L2::evict_normal */
#define PGSGD_FLAG_WINDOW_TILES 4096u /* experiments: tiles visited window by window — the tiles of ALL paths over one stretch of the node order
                                         form a window; the resident CTAs work on ~grid/C windows at a time, C tiles (paths) of each (C from
                                         PGSGD_WINDOW_C, default 3), so a window's coordinates are fetched into L2 once and reused by every
                                         path that crosses it, while no node sees more than ~C concurrent tiles */
#define PGSGD_FLAG_X_TILE_REPLACE 8192u  /* experiments (tile-sampling bias): tiles drawn with replacement instead of one bijection per pass */
#define PGSGD_FLAG_X_STEP_RANDOM 16384u  /* experiments: the first step of a term drawn with replacement inside the staged tile instead of every step once */
#define PGSGD_FLAG_X_SEGMENT_RANDOM 32768u /* experiments: every warp draws a 32-step segment of the staged tile with replacement */
#define PGSGD_FLAG_X_STEP_SCRAMBLE 65536u /* experiments: every staged step once, neighbouring lanes far apart in the tile */
#define PGSGD_FLAG_SUM_DELTAS   2u  /* multi-GPU: all-reduce SUM of per-iteration displacements instead of the MEAN of coordinates */

typedef struct pgsgd_stats {
    uint64_t iterations_run;
    uint64_t term_updates;                 /* counted term updates performed by this rank */
    double   seconds_iterations;           /* device time of the iteration loop (CUDA events), this rank */
    double   seconds_upload;               /* flatten-to-device + coordinate upload */
    double   seconds_download;
    double   last_delta_max;               /* max |Delta| of the last iteration (only tracked when delta > 0) */
    uint64_t kernel_launches;              /* SGD kernel launches */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t flags_used;                   /* PGSGD_FLAG_* actually in effect (the engine switches to EXCH_WRITE by itself
                                              when a hub node would see too many concurrent red.adds) */
    uint64_t sampling_used;                /* PGSGD_SAMPLING_STREAM or _TILE: what AUTO resolved to (last phase run) */
    double   seconds_kernels;              /* of seconds_iterations: the SGD kernels (CUDA events per iteration) ... */
    double   seconds_collectives;          /* ... and the collectives after them (transfer + waiting for the slowest rank); 0 on one GPU */
} pgsgd_stats;

typedef struct pgsgd_engine pgsgd_engine;   /* opaque: device-resident graph + coordinates + RNG streams */

const char* pgsgd_last_error(void);
int         pgsgd_version(void);
int         pgsgd_device_count(void);       /* 0 when no usable CUDA device */
int         pgsgd_device_warmup(int device); /* creates the device's CUDA context now (~0.2-0.3 s) instead of inside the first engine call: a caller
                                                with host-side work to do first (reading a file) runs this on another thread meanwhile */

/* ---- one-shot entry points (host buffers in, host buffers out) ------------------------------------
 * 2D: replaces  void cuda::gpu_layout(layout_config_t, const odgi::graph_t&, std::vector<std::atomic<double>>& X,
 *               std::vector<std::atomic<double>>& Y)                       (src/cuda/layout.h:80, layout.cu:290-476)
 *     X, Y: [2N] in/out, index 2*node_rank + end — the reference's graph_X/graph_Y layout
 *     (src/subcommand/layout_main.cpp:268-330; path_sgd_layout.cpp:322-323).
 * 1D: GPU counterpart of  std::vector<double> path_linear_sgd(...)          (src/algorithms/path_sgd.cpp:12-464)
 *     X: [N] in/out (the reference initialises X itself, path_sgd.cpp:63-69; pass x_is_initialised = 0
 *     to get that initialisation here), frozen: [N] or NULL = target_nodes of `odgi sort -H` (path_sgd.cpp:290-302). */
int pgsgd_layout_2d(const pgsgd_graph_view* g, const pgsgd_config* cfg, double* X, double* Y, pgsgd_stats* stats);
int pgsgd_sort_1d(const pgsgd_graph_view* g, const pgsgd_config* cfg, const uint8_t* frozen, int x_is_initialised,
                  double* X, pgsgd_stats* stats);

/* The same on n_gpus devices of ONE process (one host thread per GPU inside the call; multi_mode = PGSGD_MULTI_*): how a
 * single-process host such as odgi uses a whole box.  pgsgd_layout_2d / pgsgd_sort_1d dispatch here when the environment
 * variable PGSGD_GPUS is > 1 (PGSGD_MULTI = auto | hybrid | allreduce | peer, default auto), so `PGSGD_GPUS=8 odgi layout --gpu`
 * needs no source change. */
int pgsgd_layout_2d_multi(const pgsgd_graph_view* g, const pgsgd_config* cfg, int n_gpus, int multi_mode, double* X, double* Y,
                          pgsgd_stats* stats);
int pgsgd_sort_1d_multi(const pgsgd_graph_view* g, const pgsgd_config* cfg, int n_gpus, int multi_mode, const uint8_t* frozen,
                        int x_is_initialised, double* X, pgsgd_stats* stats);

/* ---- engine API (graph stays resident in HBM; used by the one-shot calls, by bench.py and by multi-GPU runs) ---- */
int  pgsgd_engine_create(const pgsgd_graph_view* g, int device, pgsgd_engine** out);
/* The same engine straight from GFA text: the caller finds the `P` lines and hands over the byte range of each
 * line's step list ("12+,13-,..." — text[field_begin[p] .. field_end[p]), no tab, no newline), plus the node lengths from the `S`
 * lines; the step lists are uploaded as text and parsed ON THE DEVICE (one thread per step), then flattened there as above.
 * No per-step host work: replaces the path walk of cuda::gpu_layout (src/cuda/layout.cu:371-410) and this repo's own host
 * parser (odgi_b200/host/gfa_lite.hpp).  Node ids must be the numbers 1..node_count (PGSGD_ERR_UNOPT otherwise). */
int  pgsgd_engine_create_from_gfa_paths(const uint32_t* node_len, uint64_t node_count, const char* text, const uint64_t* field_begin,
                                        const uint64_t* field_end, uint64_t path_count, int device, pgsgd_engine** out);
/* what the default schedule parameters are derived from (layout_main.cpp:251-266, sort_main.cpp:355-412); any pointer may be NULL */
int  pgsgd_engine_graph_stats(const pgsgd_engine* e, uint64_t* step_count, uint64_t* max_path_steps, uint64_t* max_path_bp, uint64_t* max_node_depth);
void pgsgd_engine_destroy(pgsgd_engine* e);
int  pgsgd_engine_device(const pgsgd_engine* e);
uint64_t pgsgd_engine_device_bytes(const pgsgd_engine* e);

int pgsgd_engine_set_coords_2d(pgsgd_engine* e, const double* X, const double* Y);     /* [2N] each */
int pgsgd_engine_get_coords_2d(pgsgd_engine* e, double* X, double* Y);
int pgsgd_engine_set_coords_2d_f32(pgsgd_engine* e, const float* xy);                   /* [4N] {x0,y0,x1,y1} per node */
int pgsgd_engine_get_coords_2d_f32(pgsgd_engine* e, float* xy);
int pgsgd_engine_set_coords_1d(pgsgd_engine* e, const double* X);                       /* [N]; NULL = cumulative-bp init */
int pgsgd_engine_get_coords_1d(pgsgd_engine* e, double* X);
int pgsgd_engine_set_frozen_1d(pgsgd_engine* e, const uint8_t* frozen);                 /* [N] or NULL */

/* Run the whole schedule (all iterations) on the engine's device; multi-GPU when a communicator is attached. */
int pgsgd_engine_run_2d(pgsgd_engine* e, const pgsgd_config* cfg, pgsgd_stats* stats);
int pgsgd_engine_run_1d(pgsgd_engine* e, const pgsgd_config* cfg, pgsgd_stats* stats);
/* Iterations [iter_begin, iter_end) of the schedule cfg defines (dims = 2 or 1): lets a caller interleave its own
 * work (snapshots as `odgi layout -u` takes them, progress, timing) between cooling-schedule steps.  The worker
 * streams are seeded when iter_begin == 0 and continue otherwise. */
int pgsgd_engine_run_range(pgsgd_engine* e, const pgsgd_config* cfg, int dims, uint64_t iter_begin, uint64_t iter_end,
                           pgsgd_stats* stats);

/* Multi-GPU (one process per GPU): replicate coordinates, split each iteration's term updates across
 * ranks, combine with one NCCL all-reduce per iteration.  unique_id is the 128-byte ncclUniqueId obtained
 * by rank 0 from pgsgd_comm_unique_id and distributed by the caller (torch.distributed / MPI / a file). */
int pgsgd_comm_unique_id(uint8_t id_out[128]);
int pgsgd_engine_attach_comm(pgsgd_engine* e, const uint8_t unique_id[128], int n_ranks, int rank);

/* How the ranks of a communicator combine their work (select after attach_comm, before uploading coordinates):
 *  ALLREDUCE: coordinates replicated on every GPU; each rank applies its share of an iteration's updates to its replica;
 *             one ncclAllReduce (mean, or sum of displacements with PGSGD_FLAG_SUM_DELTAS) per iteration.
 *  PEER     : coordinates PARTITIONED by node range over the GPUs of one NVLink domain (<= 8); every update reads and
 *             red.adds the owner's slice directly through NVLink peer memory (CUDA IPC), so all GPUs run one shared
 *             Hogwild exactly like the threads of the reference — no replica drift, no all-reduce of coordinates.  Tiles
 *             are assigned to the rank owning their nodes, which keeps ~80 % of the traffic on the local GPU.
 *  HYBRID   : ALLREDUCE for the first third of the schedule (cfg.multi_switch_iteration), PEER from then on.  The early
 *             iterations saturate every update (mu = 1) and draw half of the partners uniformly over the path (mostly on
 *             another GPU): replicas tolerate them and NVLink does not; the annealing two thirds decide the final stress
 *             and run as one shared Hogwild.  Final stress within the single-GPU band at 4 and 8 ranks. */
#define PGSGD_MULTI_ALLREDUCE 0
#define PGSGD_MULTI_PEER      1
#define PGSGD_MULTI_HYBRID    2
/*  SINGLE   : rank 0 runs the whole job alone (the single-GPU result, exactly), the others receive the coordinates by one
 *             broadcast.  For graphs that no way of sharing the work leaves inside the reference's stress band.
 *  AUTO     : ALLREDUCE when every replica still sees at least PGSGD_AUTO_MIN_UPDATES_PER_NODE updates per node and
 *             iteration (10 * steps / nodes / ranks in 2D, steps / nodes / ranks in 1D: graphs many haplotypes deep, where the
 *             mean of the replicas anneals like one Hogwild), else SINGLE: a shallow graph measured at 8 GPUs ends at a far
 *             stress of 1.0 with replicas, +33 % with HYBRID and +35-40 % with PEER, and is a fraction of a
 *             second of work on one GPU.  Resolved when the coordinates are uploaded. */
#define PGSGD_MULTI_AUTO      3
#define PGSGD_MULTI_SINGLE    4
#define PGSGD_AUTO_MIN_UPDATES_PER_NODE 60.0
int pgsgd_engine_set_multi_mode(pgsgd_engine* e, int mode);
/* the mode in effect (what AUTO resolved to; meaningful once coordinates are uploaded) */
int pgsgd_engine_resolved_multi_mode(const pgsgd_engine* e);

/* Path-sharded step records (graphs whose step records do not fit one GPU): a term always pairs two steps of
 * the SAME path, so the paths of a job can be dealt out over the ranks and every rank's engine is created from a view that
 * holds only ITS paths (node_len stays the whole node table).  global_step_count = the steps of the whole job: the rank
 * then performs min_term_updates * S_local / S_global updates per iteration over its own steps, which keeps every step of
 * the job equally likely to start a term (path_sgd_layout.cpp:175-182).  Coordinates stay replicated and are combined per
 * iteration (PGSGD_MULTI_ALLREDUCE only).  0 switches the mode off. */
int pgsgd_engine_set_shard(pgsgd_engine* e, uint64_t global_step_count);

/* ---- verification hooks (used by tests; they exercise exactly the device code the runs use) ---- */
/* The first n_terms draws of worker stream `stream`, produced by the device sampler.  dims = 2 or 1;
 * cooling / theta_zipf as the iteration would set them.  Outputs are [n_terms] each (any may be NULL);
 * valid[k] = 0 marks a draw that hit a 1-step path (not counted, path_sgd_layout.cpp:190-192). */
int pgsgd_engine_sample_terms(pgsgd_engine* e, const pgsgd_config* cfg, int dims, int cooling, double theta_zipf,
                              uint64_t stream, uint64_t n_terms, uint64_t* step_index, uint32_t* path, uint64_t* rank_a,
                              uint64_t* rank_b, uint32_t* node_a, uint32_t* node_b, uint64_t* pos_a, uint64_t* pos_b,
                              uint8_t* end_a, uint8_t* end_b, uint8_t* valid);

/* Sampled path stress of the resident coordinates, evaluated on the device: n_pairs (rounded up to a multiple of 4096)
 * step pairs — first step uniform over all steps, partner uniform in the same path, node ends uniform (2D) / node
 * starts (1D) — stress = mean(((|p_a - p_b| - d) / d)^2), d = path distance in bp, pairs with d = 0 skipped.  The
 * reference has no layout-quality readout on this path; this is the one the parity tests use. */
int pgsgd_engine_path_stress(pgsgd_engine* e, int dims, uint64_t n_pairs, uint64_t seed, double* stress_out);

/* The near-pair variant of the same estimator: partner 1..64 ranks away along the path, pairs further apart than 1000 bp
 * skipped (oracle: orc_local_stress_2d / _1d).  The far-pair stress is dominated by pairs megabases apart; this one reads
 * the fine structure of the layout, where limited coordinate precision would show first. */
int pgsgd_engine_local_stress(pgsgd_engine* e, int dims, uint64_t n_pairs, uint64_t seed, double* stress_out);

/* The node order `odgi sort -Y` derives from the 1D layout (path_linear_sgd_order, src/algorithms/path_sgd.cpp:638-683):
 * node ranks sorted by (weak component, position, handle), computed on the device (stable radix sorts).  order_out: [N]
 * node ranks.  node_component: [N] the component key of every node rank — components numbered by their average node id as
 * path_sgd.cpp:557-573 does — or NULL for a single-component graph.  (The reference clear()s its component map before
 * reading it, path_sgd.cpp:588; the storage stays in place and the unchecked reads still see the ids, so the running
 * reference does sort by component: tests/golden/order_multi3.json.)  The weak components themselves come from the caller:
 * the view (pgsgd_graph_view) carries no edges; odgi has algorithms::weakly_connected_components for this. */
int pgsgd_engine_order_1d_components(pgsgd_engine* e, const uint32_t* node_component, uint64_t* order_out);
int pgsgd_engine_order_1d(pgsgd_engine* e, uint64_t* order_out);   /* == node_component NULL */

/* The step right after the 2D hot path, on the device: odgi's binary layout container (`.lay`,
 * algorithms::layout::Layout::serialize, src/algorithms/layout.cpp:43-61: min_value + sdsl::enc_vector<elias_delta,128> over the
 * bit patterns of coordinate - min_value) of the resident 2D coordinates, after the per-component stacking `odgi layout` applies
 * first (src/subcommand/layout_main.cpp:402-435: bounding box per weak component, 1000-unit border).
 * node_component: [N] id of every node's weak component, numbered as the caller's weakly_connected_component_vectors returns
 * them, n_components of them — or NULL to serialise the coordinates as they are.  The file is byte-identical to what the host
 * writer (odgi_b200/host/lay_format.hpp, pinned on files written by the reference) produces from the downloaded coordinates.
 * Call with buf == NULL to learn the size (*n_bytes), then with a buffer of at least that many bytes. */
int pgsgd_engine_encode_lay(pgsgd_engine* e, const uint32_t* node_component, uint32_t n_components, uint8_t* buf, uint64_t cap, uint64_t* n_bytes);

/* Sorting goodness of the graph as `order` would sort it: the two metrics of `odgi stats -l [-g] -s [-d]`
 * (src/subcommand/stats_main.cpp:399-800, 1D branch), evaluated on the device over every consecutive step pair of every path.
 * order: [N] node ranks in sorted order (e.g. from pgsgd_engine_order_1d) or NULL for the graph as it is.
 * flags: PGSGD_GOODNESS_GAP_LINKS (-g: a link to the next node of the path's own ordered node set is not penalised),
 *        PGSGD_GOODNESS_ORIENTATION (-d: orientation changes along a path are penalised).  All sums are integers: exact. */
#define PGSGD_GOODNESS_GAP_LINKS   1u
#define PGSGD_GOODNESS_ORIENTATION 2u
typedef struct pgsgd_goodness {
    double   mean_links_length_node, mean_links_length_nt;       /* in_node_space, in_nucleotide_space */
    uint64_t num_links, num_gap_links;
    double   sum_path_node_dist_node, sum_path_node_dist_nt;
    uint64_t nodes, nucleotides, num_penalties, num_penalties_diff_orientation;
} pgsgd_goodness;
int pgsgd_engine_sort_goodness(pgsgd_engine* e, const uint64_t* order, uint32_t flags, pgsgd_goodness* out);

/* Tile-sampling verification: with a trace buffer set, every term the tile kernel draws is recorded (first step,
 * partner step as global step indices, flips = flip_a | flip_b << 1) until the buffer is full. */
int pgsgd_engine_set_trace(pgsgd_engine* e, uint64_t capacity);   /* 0 = off */
int pgsgd_engine_get_trace(pgsgd_engine* e, uint64_t* ia_out, uint64_t* ib_out, uint8_t* flips_out, uint64_t* n_out);

/* ---- host-side helpers that mirror reference host code (pure CPU, no device needed) ---- */
/* learning-rate schedule: path_linear_sgd_layout_schedule (path_sgd_layout.cpp:433-468); writes iter_max+1 */
int pgsgd_schedule(const pgsgd_config* cfg, double* etas_out);
/* Zipf zeta table (path_sgd_layout.cpp:87-97); returns the entry count, fills up to cap entries */
uint64_t pgsgd_zetas(const pgsgd_config* cfg, double* zetas_out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* PGSGD_H */
