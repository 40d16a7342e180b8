// pgsgd_capi.cu — the C-ABI (include/pgsgd.h): engine lifetime, schedule, launches, NCCL combine.
//
// Host-side logic mirrored from the reference (cited inline): the learning-rate schedule
// (path_sgd_layout.cpp:433-468), the Zipf zeta table (path_sgd_layout.cpp:87-97), the iteration / cooling
// state machine (path_sgd_layout.cpp:120-163, path_sgd.cpp:161-203, layout.cu:442-447) and the flattening
// that cuda::gpu_layout does for itself (layout.cu:325-410).
#include "../../include/pgsgd.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <random>
#include <tuple>
#include <thread>
#include <unistd.h>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

#include "pgsgd_kernels.cuh"

using namespace pgsgd;

namespace {

thread_local std::string g_last_error;

// Process-level communicator cache: ncclCommInitRank costs ~1.5 s on an 8-GPU box, more than a whole c4 layout.  A
// communicator is kept after its engine is destroyed and handed to the next engine this process attaches with the same
// (device, n_ranks, rank) — every rank of a job re-attaches together (bench.py, the one-shot multi-GPU calls, a host that
// lays out several graphs), so all ranks hit or miss the cache together.  PGSGD_COMM_CACHE=0 switches it off.
struct CommKey {
    int device, n_ranks, rank;
    bool operator<(const CommKey& o) const { return std::tie(device, n_ranks, rank) < std::tie(o.device, o.n_ranks, o.rank); }
};
std::mutex g_comm_mu;
std::map<CommKey, ncclComm_t> g_comm_cache;
bool comm_cache_enabled() {
    const char* s = getenv("PGSGD_COMM_CACHE");
    return !(s && s[0] == '0');
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t _e = (call);                                                                              \
        if (_e != cudaSuccess)                                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? PGSGD_ERR_NOMEM : PGSGD_ERR_CUDA, "%s:%d: %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(_e));                                   \
    } while (0)

#define NC(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t _r = (call);                                                                        \
        if (_r != ncclSuccess)                                                                           \
            return fail(PGSGD_ERR_NCCL, "%s:%d: %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(_r)); \
    } while (0)

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// dirtyzipf::fast_precise_pow on the host (deps/dirtyzipf/dirty_zipfian_int_distribution.h:82-104); the
// device twin is pgsgd::fast_precise_pow in pgsgd_device.cuh
double host_fast_precise_pow(double a, double b) {
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

uint64_t zeta_table_size(const pgsgd_config& c) {
    return (c.space <= c.space_max ? c.space : c.space_max + (c.space - c.space_max) / c.space_quantization_step + 1) + 1;
}

ZipfConst make_zipf_const(double theta) {
    ZipfConst z;
    z.theta = theta;
    z.one_minus_theta = 1 - theta;
    z.alpha = 1 / (1 - theta);
    z.zeta2 = host_fast_precise_pow(1.0 / 1, theta) + host_fast_precise_pow(1.0 / 2, theta);  // zeta(2, theta)
    z.thresh2 = 1.0 + host_fast_precise_pow(0.5, theta);
    return z;
}

ZipfConstF make_zipf_const_f(const ZipfConst& z) {
    ZipfConstF f;
    f.one_minus_theta = (float) z.one_minus_theta;
    f.alpha_int = (int) z.alpha;
    f.alpha_frac = (float) (z.alpha - (double) f.alpha_int);
    f.zeta2 = (float) z.zeta2;
    f.thresh2 = (float) z.thresh2;
    return f;
}

int check_config(const pgsgd_config* c) {
    if (!c) return fail(PGSGD_ERR_ARG, "config is NULL");
    if (c->iter_max == 0) return fail(PGSGD_ERR_ARG, "iter_max must be > 0");
    if (c->space_quantization_step == 0) return fail(PGSGD_ERR_ARG, "space_quantization_step must be > 0");
    if (!(c->theta > 0.0 && c->theta < 1.0)) return fail(PGSGD_ERR_ARG, "theta must be in (0,1)");
    if (!(c->eta_max > 0.0)) return fail(PGSGD_ERR_ARG, "eta_max must be > 0");
    if (c->batch != 0 && c->batch != 1 && c->batch != 2 && c->batch != 4) return fail(PGSGD_ERR_ARG, "batch must be 0, 1, 2 or 4");
    if (c->sampling > PGSGD_SAMPLING_TILE) return fail(PGSGD_ERR_ARG, "sampling must be PGSGD_SAMPLING_AUTO, _STREAM or _TILE");
    return PGSGD_OK;
}

}  // namespace

struct pgsgd_engine {
    int device = 0;
    int sm_count = 0;
    uint64_t N = 0, P = 0, S = 0;
    uint64_t max_path_steps = 0;
    uint64_t max_node_depth = 0;             // most steps on one node (hub nodes bound the safe Hogwild concurrency)
    uint64_t tile_repeats = 0;               // steps that revisit a node already seen in their tile (tandem repeats; bounds the tile-mode shape)
    bool any_multi_step_path = false;
    StepRec* d_steps = nullptr;
    uint64_t* d_path_first = nullptr;
    uint32_t* d_node_len = nullptr;          // [N] kept for the sorting-goodness readout
    float* d_xy = nullptr;        // 2D coordinates
    float* d_xy_prev = nullptr;   // multi-GPU sum-of-deltas scratch
    double* d_x1d = nullptr;      // 1D coordinates
    double* d_x1d_prev = nullptr;
    uint8_t* d_frozen = nullptr;
    double* d_zetas = nullptr;
    uint64_t zetas_cap = 0;
    // the zeta table costs O(space) bit-hack pows on the host (4.6e6 for c4: ~10 ms): kept across run calls with the same Zipf
    // parameters, so that a caller stepping through the schedule (run_range per iteration) pays it once
    std::vector<double> zetas_host;
    uint64_t zkey_space = 0, zkey_max = 0, zkey_q = 0;
    double zkey_theta = -1;
    bool zetas_on_device = false;
    int ztab_dims = 0;                        // for which dims the fp32 companions were uploaded (0 = not)
    float2* d_ztab[2] = {nullptr, nullptr};  // pipelined tile kernel: fp32 {zeta_n, 1/(1 - zeta_2/zeta_n)} for theta / for the 1D cooling theta
    uint64_t ztab_cap = 0;
    uint64_t max_path_bp = 0;                // largest end-adjusted position (pos + len) of any step: < 2^32 selects 32-bit distances
    uint64_t* d_rng = nullptr;
    uint64_t rng_stride = 0;
    uint64_t rng_streams = 0;     // streams seeded by the run in progress
    unsigned int* d_delta = nullptr;
    unsigned long long* d_counted = nullptr;
    unsigned long long* d_trace = nullptr;        // verification trace of the tile kernel (pgsgd_engine_set_trace)
    unsigned long long* d_trace_count = nullptr;
    uint64_t trace_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<double> h_x1d_default;  // cumulative bp of the node order (path_sgd.cpp:63-69), built on first use
    bool have_2d = false, have_1d = false;
    uint64_t bytes = 0;
    double seconds_upload = 0;
    uint64_t h2d_bytes = 0;
    ncclComm_t comm = nullptr;
    bool comm_cached = false;               // the communicator belongs to the process-level cache (not destroyed with the engine)
    bool comm_warm = false;                 // a full-size collective has run on it (NCCL connects its channels lazily)
    int n_ranks = 1, rank = 0;
    // ---- peer mode (PGSGD_MULTI_PEER): coordinates partitioned by node range, accessed through NVLink peer memory ----
    int multi_mode = 0;                      // what the caller selected (PGSGD_MULTI_*)
    int mode = 0;                            // what AUTO resolved to when the coordinates were uploaded (else == multi_mode)
    cudaEvent_t ev_r0 = nullptr, ev_r1 = nullptr;   // bracket a whole run_engine call (both phases of a hybrid run + the switch)
    std::vector<cudaEvent_t> ev_pool;               // per-iteration events of multi-GPU runs (kernel | collective split)
    int active_mode = 0;                     // what the run in progress uses: ALLREDUCE or PEER (HYBRID switches between them)
    bool coords_in_slices = false;           // the authoritative coordinates are in the peer slices (peer phase), not the replica
    bool peer_ready_2d = false, peer_ready_1d = false;
    uint64_t part_chunk = 0;                 // nodes per partition (last one may be shorter)
    uint32_t part_lo[9] = {0};
    float* d_xy_part = nullptr;              // this rank's slice, part_chunk * 4 floats
    double* d_x1d_part = nullptr;
    float* peer_xy[8] = {nullptr};           // pre-offset base pointers (see IterParams::part_xy)
    double* peer_x1d[8] = {nullptr};
    std::vector<void*> ipc_opened;
    std::vector<uint32_t> tile_mid_node;     // node of the middle step of every tile (tile -> owner rank)
    uint32_t* d_tile_list = nullptr;
    double* d_stage_xy[2] = {nullptr, nullptr};   // fp64 X / Y staging of set/get_coords_2d (2N each), kept: no cudaMalloc / cudaFree per transfer
    uint32_t* d_window_list = nullptr;       // PGSGD_FLAG_WINDOW_TILES: all tiles, window by window (build_window_list)
    uint64_t window_list_key = 0;            // (grid, C) the list was built for
    uint64_t my_tiles = 0, my_tile_steps = 0;
    // ---- path-sharded step records (pgsgd_engine_set_shard): this engine holds only some of the job's paths ----
    uint64_t shard_global_steps = 0;         // S of the whole job; 0 = the view is the whole graph
};

namespace {

template <typename T>
int dev_alloc(pgsgd_engine* e, T** p, uint64_t count) {
    if (count == 0) count = 1;
    cudaError_t err = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
    if (err != cudaSuccess) return fail(PGSGD_ERR_NOMEM, "cudaMalloc of %llu bytes failed: %s", (unsigned long long) (count * sizeof(T)), cudaGetErrorString(err));
    e->bytes += count * sizeof(T);
    return PGSGD_OK;
}

int ensure_zetas(pgsgd_engine* e, const std::vector<double>& z) {
    if (z.size() > e->zetas_cap) {
        if (e->d_zetas) cudaFree(e->d_zetas);
        e->d_zetas = nullptr;
        int rc = dev_alloc(e, &e->d_zetas, z.size());
        if (rc) return rc;
        e->zetas_cap = z.size();
    }
    CU(cudaMemcpyAsync(e->d_zetas, z.data(), z.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    return PGSGD_OK;
}

// fp32 companion of the zeta table for the pipelined tile kernel: {zeta_n, 1 / (1 - zeta_2 / zeta_n)} per table entry, so
// that the per-term eta of the dirty Zipf (dirty_zipfian_int_distribution.h:126-133) costs one multiply instead of two
// divisions; zeta_2 is recomputed from the theta handed to the draw exactly as the reference does (:126-128)
int upload_ztab(pgsgd_engine* e, int which, const std::vector<double>& z, double theta_draw) {
    if (z.size() > e->ztab_cap) {
        for (int k = 0; k < 2; ++k) { if (e->d_ztab[k]) cudaFree(e->d_ztab[k]); e->d_ztab[k] = nullptr; }
        for (int k = 0; k < 2; ++k) { int rc = dev_alloc(e, &e->d_ztab[k], z.size()); if (rc) return rc; }
        e->ztab_cap = z.size();
    }
    const float zeta2 = (float) make_zipf_const(theta_draw).zeta2;
    std::vector<float2> t(z.size());
    for (size_t i = 0; i < z.size(); ++i) {
        const float zn = (float) z[i];
        t[i].x = zn;
        t[i].y = 1.0f / (1.0f - zeta2 / zn);
    }
    CU(cudaMemcpyAsync(e->d_ztab[which], t.data(), t.size() * sizeof(float2), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));   // t is a local
    return PGSGD_OK;
}

int ensure_rng(pgsgd_engine* e, uint64_t n_streams) {
    if (n_streams > e->rng_stride) {
        if (e->d_rng) cudaFree(e->d_rng);
        e->d_rng = nullptr;
        int rc = dev_alloc(e, &e->d_rng, 4 * n_streams);
        if (rc) return rc;
        e->rng_stride = n_streams;
    }
    return PGSGD_OK;
}

std::vector<double> build_zetas(const pgsgd_config& c) {
    // path_sgd_layout.cpp:87-97
    std::vector<double> zetas(zeta_table_size(c), 0.0);
    double zeta_tmp = 0.0;
    for (uint64_t i = 1; i < c.space + 1; i++) {
        zeta_tmp += host_fast_precise_pow(1.0 / i, c.theta);
        if (i <= c.space_max) zetas[i] = zeta_tmp;
        // The reference writes zetas[space_max + 1] when i == space == space_max, one element past its vector
        // (path_sgd_layout.cpp:92-95 with a table of space + 1 entries); that entry is never read (jump spaces above
        // space_max do not exist then), so it is simply not stored here.
        if (i >= c.space_max && (i - c.space_max) % c.space_quantization_step == 0) {
            const uint64_t slot = c.space_max + 1 + (i - c.space_max) / c.space_quantization_step;
            if (slot < zetas.size()) zetas[slot] = zeta_tmp;
        }
    }
    return zetas;
}

std::vector<double> build_schedule(const pgsgd_config& c) {
    // path_linear_sgd_layout_schedule (path_sgd_layout.cpp:433-468) with w_min = 1/eta_max, w_max = 1 (:75-84)
    const double w_min = (double) 1.0 / (double) (c.eta_max);
    const double w_max = 1.0;
    const double eta_max = 1.0 / w_min;
    const double eta_min = c.eps / w_max;
    const double lambda = log(eta_max / eta_min) / ((double) c.iter_max - 1);
    std::vector<double> etas;
    etas.reserve(c.iter_max + 1);
    for (int64_t t = 0; t <= (int64_t) c.iter_max; t++) {
        etas.push_back(eta_max * exp(-lambda * (double) (std::llabs(t - (int64_t) c.iter_with_max_learning_rate))));
    }
    return etas;
}

// the iteration loop shared by 2D and 1D
int comm_barrier(pgsgd_engine* e) {
    if (!e->comm) return PGSGD_OK;
    NC(ncclAllReduce(e->d_delta, e->d_delta, 1, ncclUint32, ncclMax, e->comm, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PGSGD_OK;
}

// Peer mode set-up: partition the node range, allocate this rank's slice, exchange CUDA IPC handles through the NCCL
// communicator and map every peer's slice (NVLink peer memory).  Also assigns every tile to the rank that owns the node
// of its middle step, so that in tile mode both first-node accesses and most partner accesses stay on the local GPU.
int setup_peer(pgsgd_engine* e, int dims) {
    if (!e->comm || e->n_ranks < 2) return fail(PGSGD_ERR_STATE, "peer mode needs an attached communicator with at least 2 ranks");
    if (e->n_ranks > 8) return fail(PGSGD_ERR_ARG, "peer mode supports up to 8 ranks (one NVSwitch domain)");
    const int n = e->n_ranks;
    e->part_chunk = ((e->N + n - 1) / n + 1) & ~1ull;   // even: the pipelined kernel fetches 1D coordinates as aligned 16-byte pairs
    for (int q = 0; q <= n; ++q) {
        const uint64_t lo = (uint64_t) q * e->part_chunk;
        e->part_lo[q] = (uint32_t) (lo < e->N ? lo : e->N);
    }
    for (int q = n + 1; q < 9; ++q) e->part_lo[q] = (uint32_t) e->N;
    void* mine = nullptr;
    const size_t bytes = e->part_chunk * (dims == 2 ? 4 * sizeof(float) : sizeof(double));
    if (dims == 2) {
        if (!e->d_xy_part) { int rc = dev_alloc(e, &e->d_xy_part, e->part_chunk * 4); if (rc) return rc; }
        mine = e->d_xy_part;
    } else {
        if (!e->d_x1d_part) { int rc = dev_alloc(e, &e->d_x1d_part, e->part_chunk); if (rc) return rc; }
        mine = e->d_x1d_part;
    }
    CU(cudaMemsetAsync(mine, 0, bytes, e->stream));
    // what every rank publishes about its slice: the CUDA IPC handle (other processes map it) and, for ranks that are
    // threads of THIS process (single-process multi-GPU, e.g. the odgi shim), the raw pointer + device (peer access)
    struct SliceInfo { cudaIpcMemHandle_t ipc; uint64_t pid /* process nonce */; uint64_t ptr; int32_t device; int32_t pad; };
    static_assert(sizeof(SliceInfo) == 88, "SliceInfo layout");
    SliceInfo info;
    memset(&info, 0, sizeof(info));
    CU(cudaIpcGetMemHandle(&info.ipc, mine));
    // "same process" is decided by a per-process random nonce, not by the pid: ranks in different PID namespaces (one process
    // per container) can share a pid, and a foreign raw pointer must never be dereferenced
    static const uint64_t process_nonce = []() {
        std::random_device rd;
        return ((uint64_t) rd() << 32) ^ (uint64_t) rd() ^ ((uint64_t) getpid() << 17) ^ (uint64_t) std::chrono::steady_clock::now().time_since_epoch().count();
    }();
    info.pid = process_nonce;
    info.ptr = (uint64_t) (uintptr_t) mine;
    info.device = e->device;
    uint8_t* d_h = nullptr;
    CU(cudaMalloc(&d_h, sizeof(SliceInfo) * (size_t) (n + 1)));
    CU(cudaMemcpyAsync(d_h, &info, sizeof(SliceInfo), cudaMemcpyHostToDevice, e->stream));
    NC(ncclAllGather(d_h, d_h + sizeof(SliceInfo), sizeof(SliceInfo), ncclUint8, e->comm, e->stream));
    std::vector<SliceInfo> all(n);
    CU(cudaMemcpyAsync(all.data(), d_h + sizeof(SliceInfo), sizeof(SliceInfo) * (size_t) n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(d_h);
    for (int q = 0; q < n; ++q) {
        void* ptr = mine;
        if (q != e->rank) {
            if (all[q].pid == info.pid) {
                // same process: the allocation is directly addressable once peer access is enabled
                cudaError_t pe = cudaDeviceEnablePeerAccess(all[q].device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(PGSGD_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", all[q].device, cudaGetErrorString(pe));
                cudaGetLastError();
                ptr = (void*) (uintptr_t) all[q].ptr;
            } else {
                CU(cudaIpcOpenMemHandle(&ptr, all[q].ipc, cudaIpcMemLazyEnablePeerAccess));
                e->ipc_opened.push_back(ptr);
            }
        }
        if (dims == 2) e->peer_xy[q] = reinterpret_cast<float*>(ptr) - 4 * (ptrdiff_t) e->part_lo[q];
        else e->peer_x1d[q] = reinterpret_cast<double*>(ptr) - (ptrdiff_t) e->part_lo[q];
    }
    if (!e->d_tile_list) {
        std::vector<uint32_t> mine_tiles;
        uint64_t steps = 0;
        const uint64_t W = TILE_STEPS;
        for (uint64_t t = 0; t < e->tile_mid_node.size(); ++t) {
            uint64_t owner = e->tile_mid_node[t] / e->part_chunk;
            if (owner >= (uint64_t) n) owner = n - 1;
            if ((int) owner == e->rank) {
                mine_tiles.push_back((uint32_t) t);
                const uint64_t lo = t * W, hi = lo + W < e->S ? lo + W : e->S;
                steps += hi - lo;
            }
        }
        e->my_tiles = mine_tiles.size();
        e->my_tile_steps = steps;
        int rc = dev_alloc(e, &e->d_tile_list, mine_tiles.size());
        if (rc) return rc;
        if (!mine_tiles.empty()) CU(cudaMemcpyAsync(e->d_tile_list, mine_tiles.data(), mine_tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    if (dims == 2) e->peer_ready_2d = true; else e->peer_ready_1d = true;
    return comm_barrier(e);
}

// full replica (d_xy / d_x1d) -> this rank's slice; called after every coordinate upload in peer mode
int peer_scatter(pgsgd_engine* e, int dims) {
    if (!(dims == 2 ? e->peer_ready_2d : e->peer_ready_1d)) { int rc = setup_peer(e, dims); if (rc) return rc; }
    const uint64_t lo = e->part_lo[e->rank], hi = e->part_lo[e->rank + 1];
    if (dims == 2) CU(cudaMemcpyAsync(e->d_xy_part, e->d_xy + 4 * lo, (hi - lo) * 4 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
    else CU(cudaMemcpyAsync(e->d_x1d_part, e->d_x1d + lo, (hi - lo) * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    e->coords_in_slices = true;
    return comm_barrier(e);
}

// every rank's slice -> full replica on every rank (one ncclAllGather of equally sized, padded slices)
int peer_gather(pgsgd_engine* e, int dims) {
    const int n = e->n_ranks;
    int rc = comm_barrier(e);
    if (rc) return rc;
    if (dims == 2) {
        float* tmp = nullptr;
        CU(cudaMalloc(&tmp, (size_t) n * e->part_chunk * 4 * sizeof(float)));
        ncclResult_t r = ncclAllGather(e->d_xy_part, tmp, e->part_chunk * 4, ncclFloat, e->comm, e->stream);
        cudaError_t ce = cudaSuccess;
        if (r == ncclSuccess) ce = cudaMemcpyAsync(e->d_xy, tmp, 4 * e->N * sizeof(float), cudaMemcpyDeviceToDevice, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        cudaFree(tmp);
        if (r != ncclSuccess) return fail(PGSGD_ERR_NCCL, "peer_gather: %s", ncclGetErrorString(r));
        if (ce != cudaSuccess) return fail(PGSGD_ERR_CUDA, "peer_gather: %s", cudaGetErrorString(ce));
    } else {
        double* tmp = nullptr;
        CU(cudaMalloc(&tmp, (size_t) n * e->part_chunk * sizeof(double)));
        ncclResult_t r = ncclAllGather(e->d_x1d_part, tmp, e->part_chunk, ncclDouble, e->comm, e->stream);
        cudaError_t ce = cudaSuccess;
        if (r == ncclSuccess) ce = cudaMemcpyAsync(e->d_x1d, tmp, e->N * sizeof(double), cudaMemcpyDeviceToDevice, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        cudaFree(tmp);
        if (r != ncclSuccess) return fail(PGSGD_ERR_NCCL, "peer_gather: %s", ncclGetErrorString(r));
        if (ce != cudaSuccess) return fail(PGSGD_ERR_CUDA, "peer_gather: %s", cudaGetErrorString(ce));
    }
    return PGSGD_OK;
}

// iterations [iter_begin, iter_end) of the schedule cfg defines; iter_end == UINT64_MAX means "to the end"
// PGSGD_FLAG_WINDOW_TILES: a visiting order of ALL tiles in which the CTAs resident at any time work on ~grid/C windows of the
// node order with ~C tiles (paths) of each.  Window = depth (= S/N rounded) consecutive tiles of the tile list sorted by the
// node in the tile's middle, i.e. roughly the tiles of all paths over one tile length of the node order.  The windows are
// visited in a seeded random order, the tiles inside a window likewise; a group of grid/C windows is emitted round-robin.
int build_window_list(pgsgd_engine* e, unsigned grid, unsigned C, uint64_t seed) {
    const uint64_t key = ((uint64_t) grid << 32) | C;
    if (e->d_window_list && e->window_list_key == key) return PGSGD_OK;
    const uint64_t nt = e->tile_mid_node.size();
    std::vector<uint32_t> T(nt);
    for (uint64_t t = 0; t < nt; ++t) T[t] = (uint32_t) t;
    std::stable_sort(T.begin(), T.end(), [&](uint32_t a, uint32_t b) { return e->tile_mid_node[a] < e->tile_mid_node[b]; });
    const uint64_t D = std::max<uint64_t>(1, (e->S + e->N / 2) / e->N);
    const uint64_t nw = (nt + D - 1) / D, A = std::max<uint64_t>(1, grid / std::max(1u, C));
    uint64_t sm = seed ^ 0x5851f42d4c957f2dULL;
    auto shuffle = [&](uint32_t* a, uint64_t n) {
        for (uint64_t i = n; i > 1; --i) { const uint64_t j = splitmix64_next(sm) % i; std::swap(a[i - 1], a[j]); }
    };
    for (uint64_t w = 0; w < nw; ++w) shuffle(T.data() + w * D, std::min(D, nt - w * D));
    std::vector<uint32_t> worder(nw);
    for (uint64_t w = 0; w < nw; ++w) worder[w] = (uint32_t) w;
    shuffle(worder.data(), nw);
    std::vector<uint32_t> out;
    out.reserve(nt);
    for (uint64_t g0 = 0; g0 < nw; g0 += A) {
        const uint64_t g1 = std::min(nw, g0 + A);
        for (uint64_t r = 0; r < D; ++r)
            for (uint64_t k = g0; k < g1; ++k) {
                const uint64_t idx = (uint64_t) worder[k] * D + r;
                if (idx < nt) out.push_back(T[idx]);
            }
    }
    if (out.size() != nt) return fail(PGSGD_ERR_STATE, "window list: %llu of %llu tiles", (unsigned long long) out.size(), (unsigned long long) nt);
    if (!e->d_window_list) { int rc = dev_alloc(e, &e->d_window_list, nt ? nt : 1); if (rc) return rc; }
    CU(cudaMemcpyAsync(e->d_window_list, out.data(), nt * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->window_list_key = key;
    return PGSGD_OK;
}

int run_phase(pgsgd_engine* e, const pgsgd_config* cfg, int dims, uint64_t iter_begin, uint64_t iter_end, pgsgd_stats* stats) {
    int rc = check_config(cfg);
    if (rc) return rc;
    if (dims == 2 && !e->have_2d) return fail(PGSGD_ERR_STATE, "2D coordinates were not set (pgsgd_engine_set_coords_2d)");
    if (dims == 1 && !e->have_1d) return fail(PGSGD_ERR_STATE, "1D coordinates were not set (pgsgd_engine_set_coords_1d)");
    CU(cudaSetDevice(e->device));
    pgsgd_stats st;
    memset(&st, 0, sizeof(st));
    st.seconds_upload = e->seconds_upload;
    st.h2d_bytes = e->h2d_bytes;
    // the reference does nothing when no path has more than one step (path_sgd_layout.cpp:64-74)
    if (!e->any_multi_step_path) {
        if (stats) *stats = st;
        return PGSGD_OK;
    }
    if (cfg->space == 0) return fail(PGSGD_ERR_ARG, "space must be > 0");

    const std::vector<double> etas = build_schedule(*cfg);
    if (e->zetas_host.empty() || e->zkey_space != cfg->space || e->zkey_max != cfg->space_max || e->zkey_q != cfg->space_quantization_step ||
        e->zkey_theta != cfg->theta) {
        e->zetas_host = build_zetas(*cfg);
        e->zkey_space = cfg->space; e->zkey_max = cfg->space_max; e->zkey_q = cfg->space_quantization_step; e->zkey_theta = cfg->theta;
        e->zetas_on_device = false;
        e->ztab_dims = 0;
    }
    const std::vector<double>& zetas = e->zetas_host;
    if (!e->zetas_on_device) {
        rc = ensure_zetas(e, zetas);
        if (rc) return rc;
        CU(cudaStreamSynchronize(e->stream));
        e->zetas_on_device = true;
    }

    const uint64_t first_cooling_iteration = (uint64_t) std::floor(cfg->cooling_start * (double) cfg->iter_max);
    uint64_t n_iters = dims == 1 ? cfg->iter_max + 1 : cfg->iter_max;  // path_sgd.cpp:181 vs path_sgd_layout.cpp:140
    if (iter_end < n_iters) n_iters = iter_end;
    if (iter_begin > n_iters) return fail(PGSGD_ERR_ARG, "iter_begin %llu is past the end of the schedule", (unsigned long long) iter_begin);

    // this rank's share of every iteration's term updates
    const uint64_t U = cfg->min_term_updates;
    // replicated step records: an equal share of U; path-sharded records: in proportion to the steps this rank holds, so
    // that every step of the job stays equally likely to be a term's first step (path_sgd_layout.cpp:175-182)
    const bool sharded = e->shard_global_steps != 0;
    if (sharded && e->active_mode != PGSGD_MULTI_ALLREDUCE) return fail(PGSGD_ERR_STATE, "path-sharded step records need PGSGD_MULTI_ALLREDUCE (peer phases walk tiles by node range)");
    const uint64_t U_rank = sharded ? (uint64_t) ((unsigned __int128) U * e->S / e->shard_global_steps)
                                    : U / e->n_ranks + ((uint64_t) e->rank < U % e->n_ranks ? 1 : 0);
    const bool peer = e->active_mode == PGSGD_MULTI_PEER && e->comm;
    if (peer && !(dims == 2 ? e->peer_ready_2d : e->peer_ready_1d)) return fail(PGSGD_ERR_STATE, "peer mode: coordinates were not set after the mode was selected");
    // in peer mode all ranks update ONE coordinate array: the Hogwild in-flight cap is shared by the ranks
    const uint64_t cap_div = peer ? (uint64_t) e->n_ranks : 1;
    // terms in flight <= N/2 with the lossless red.add write (stress parity measured up to ~2N), N/4 with the racy writes
    const uint64_t cap_frac = (cfg->flags & (PGSGD_FLAG_EXCH_WRITE | PGSGD_FLAG_PLAIN_STORE)) ? 4 : 2;

    // ---- sampling mode and launch shape ----
    const int block = 256;
    const size_t smem_first = (e->P + 1) * sizeof(uint64_t);
    // AUTO: tile sampling for graphs whose step records exceed the L2 (>= 2^22 steps) AND that are at least 8 steps deep per
    // node on average.  On a 6-haplotype, 3.6e6-node graph (5 steps per node) tile sampling ends 10-35 % above the reference's
    // far-stress band with a heavy tail over seeds while stream sampling sits inside it; from 10 steps per node on the two
    // agree (profiles/r02_tile_vs_stream_depth.log).  Shallow graphs therefore keep the reference-exact stream sampler.
    bool tile_mode = cfg->sampling == PGSGD_SAMPLING_TILE ||
                     (cfg->sampling == PGSGD_SAMPLING_AUTO && e->S >= (1ull << 22) && e->S >= PGSGD_AUTO_TILE_MIN_DEPTH * e->N);
    int batch = cfg->batch ? (int) cfg->batch : 1;  // 64 registers, 4 CTAs/SM: occupancy beats per-thread batching (profiles/)
    bool smem_paths = false;
    size_t smem = 0;
    int blocks_per_sm = 0;
    uint64_t n_streams = 0;
    LaunchShape shape;
    // the pipelined tile kernel (pgsgd_tile2.cu) unless the legacy one is asked for or needed (NVLink peer-partitioned
    // coordinates, paths with >= 2^31 steps)
    bool tile2 = false;
    int tile_steps = TILE_STEPS;
    if (tile_mode) {
        const bool tma = (cfg->flags & PGSGD_FLAG_TMA_STAGING) != 0;
        tile2 = !(cfg->flags & PGSGD_FLAG_LEGACY_TILE) && e->max_path_steps < (1ull << 31) && e->N < (1ull << 30);
        if (tile2) {
            if (peer) tile_steps = TILE_STEPS;   // tile ownership lists are built for this size
            else if (cfg->flags & PGSGD_FLAG_HALF_TILE) tile_steps = 1024;
            else if ((cfg->flags & PGSGD_FLAG_BIG_TILE) && !tma) tile_steps = 4096;
            batch = 1;
            smem = tile2_smem_bytes(tile_steps, tma, (uint32_t) e->P, &smem_paths);
            CU(tile2_occupancy(dims, tile_steps, tma, smem, &blocks_per_sm));
        } else {
            const size_t tile_bytes = (tma ? 2 : 1) * (size_t) TILE_STEPS * sizeof(StepRec);  // TMA staging is double-buffered
            smem_paths = tile_bytes + smem_first <= 200 * 1024;
            smem = tile_bytes + (smem_paths ? smem_first : 0);
            CU(tile_occupancy(dims, batch, smem, smem_paths, tma, &blocks_per_sm));
        }
        if (blocks_per_sm < 1) return fail(PGSGD_ERR_CUDA, "tile kernel does not fit on an SM (smem %zu)", smem);
        uint64_t grid = (uint64_t) e->sm_count * blocks_per_sm;
        if (cfg->n_streams) grid = (cfg->n_streams + block - 1) / block;
        // Hogwild staleness cap (see below): terms in flight = grid * block * batch
        // tile sampling concentrates the in-flight first steps on (grid) tiles of consecutive steps: nodes a path revisits
        // within a tile (tandem repeats, LPA) see several concurrent terms, so the cap is 2x tighter than for stream sampling
        // ... scaled by the share of a tile's steps that are first visits of their node (1 for a path without repeats)
        const double distinct = e->S ? 1.0 - (double) e->tile_repeats / (double) e->S : 1.0;
        const uint64_t cap_grid = (uint64_t) ((double) (e->N / (2 * cap_frac)) * distinct) / ((uint64_t) block * batch * cap_div);
        if (!cfg->n_streams && grid > cap_grid) grid = cap_grid;
        if (grid == 0) {
            if (cfg->sampling == PGSGD_SAMPLING_TILE) grid = 1; else tile_mode = false;
        }
        if (tile_mode) {
            n_streams = grid * block;
            shape.block = block; shape.grid = (int) grid; shape.smem = smem;
        }
    }
    if (!tile_mode) {
        batch = cfg->batch ? (int) cfg->batch : 1;
        smem_paths = smem_first <= 200 * 1024;
        smem = smem_paths ? smem_first : 0;
        CU(iteration_occupancy(dims, batch, block, smem, smem_paths, &blocks_per_sm));
        if (blocks_per_sm < 1) return fail(PGSGD_ERR_CUDA, "iteration kernel does not fit on an SM (smem %zu)", smem);
        n_streams = cfg->n_streams;
        if (n_streams == 0) {
            n_streams = (uint64_t) e->sm_count * blocks_per_sm * block;
            // Hogwild staleness: with more than ~N/4 terms in flight the final stress of small graphs drifts away from the
            // reference's (measured: profiles/r01_stream_sweep.md); large graphs are not affected by this cap
            uint64_t cap = (e->N / cap_frac) / (batch * cap_div);
            if (cap < 32) cap = 32;
            if (n_streams > cap) n_streams = cap;
            // keep at least ~64 terms per stream so the launch is not all prologue
            const uint64_t want = (U_rank + 63) / 64;
            if (want < n_streams) n_streams = want;
            if (n_streams >= (uint64_t) block) n_streams = (n_streams / block) * block;
            else n_streams = ((n_streams + 31) / 32) * 32;
            if (n_streams == 0) n_streams = 32;
        }
        shape.block = n_streams < (uint64_t) block ? (int) ((n_streams + 31) / 32 * 32) : block;
        shape.grid = (int) ((n_streams + shape.block - 1) / shape.block);
        shape.smem = smem;
    }

    if (iter_begin == 0 || e->rng_streams != n_streams) {
        rc = ensure_rng(e, n_streams);
        if (rc) return rc;
        // worker stream t of rank r is the reference's worker thread (r * n_streams + t): seed + tid (path_sgd_layout.cpp:168).
        // A continuation with a different launch shape (the hybrid multi-GPU schedule changes phase) starts fresh streams.
        const uint64_t reseed = iter_begin == 0 ? 0 : 0x5851F42D4C957F2DULL * iter_begin;
        CU(launch_seed_streams(e->d_rng, e->rng_stride, n_streams, cfg->seed + reseed + (uint64_t) e->rank * n_streams, e->stream));
        e->rng_streams = n_streams;
    }
    CU(cudaMemsetAsync(e->d_counted, 0, sizeof(unsigned long long), e->stream));

    const bool track_delta = cfg->delta > 0;
    const bool sum_deltas = e->comm && e->active_mode != PGSGD_MULTI_PEER && (cfg->flags & PGSGD_FLAG_SUM_DELTAS);
    if (sum_deltas) {
        if (dims == 2 && !e->d_xy_prev) { rc = dev_alloc(e, &e->d_xy_prev, 4 * e->N); if (rc) return rc; }
        if (dims == 1 && !e->d_x1d_prev) { rc = dev_alloc(e, &e->d_x1d_prev, e->N); if (rc) return rc; }
    }

    IterParams p;
    memset(&p, 0, sizeof(p));
    p.sp.path_first = e->d_path_first;
    p.sp.zetas = e->d_zetas;
    p.sp.step_count = e->S;
    p.sp.path_count = (uint32_t) e->P;
    p.sp.space = cfg->space;
    p.sp.space_max = cfg->space_max;
    p.sp.space_q = cfg->space_quantization_step;
    p.steps = e->d_steps;
    p.xy = e->d_xy;
    p.x1d = e->d_x1d;
    p.frozen = dims == 1 ? e->d_frozen : nullptr;
    p.rng = e->d_rng;
    p.rng_stride = e->rng_stride;
    p.n_streams = n_streams;
    p.quota_base = U_rank / n_streams;
    p.quota_rem = U_rank % n_streams;
    p.delta_max_bits = track_delta ? e->d_delta : nullptr;
    p.counted = e->d_counted;
    p.flags = cfg->flags;
    {
        // Hub safeguard.  red.add accumulates every concurrent (stale) displacement of a node; that is stable while a node
        // end sees a few terms in flight (measured: parity up to ~2.3, divergence from ~9) and unstable beyond.  The node
        // with the most steps sees in_flight * 2 * depth / S of them.  Past the margin the run uses the reference kernel's
        // last-writer-wins exchange instead, which cannot overshoot (it drops concurrent updates, as the reference does).
        const double in_flight = (double) n_streams * batch * (peer ? e->n_ranks : 1);
        const double hub_terms = e->S ? in_flight * 2.0 * (double) e->max_node_depth / (double) e->S : 0.0;
        // Over NVLink a term stays in flight several times longer than on one GPU, and the LPA 1D run on 2 GPUs at 3.7 left
        // the reference band (round 1, tests/test_gpu_multi.py), so peer phases switch earlier.
        // One GPU: measured in band up to ~2.3 and diverging from ~9; the staleness model (oracle/pgsgd_oracle.c
        // orc_run_inflight, the worst-case read-to-write distance) turns unstable at ~2.5, so nothing in between is trusted.
        const double hub_margin = peer ? 2.0 : 2.5;
        if (hub_terms > hub_margin && !(p.flags & (PGSGD_FLAG_EXCH_WRITE | PGSGD_FLAG_PLAIN_STORE | PGSGD_FLAG_KEEP_ADD))) {
            p.flags |= PGSGD_FLAG_EXCH_WRITE;
            st.flags_used |= PGSGD_FLAG_EXCH_WRITE;
        }
    }
    st.flags_used |= p.flags;
    st.sampling_used = tile_mode ? PGSGD_SAMPLING_TILE : PGSGD_SAMPLING_STREAM;
    p.smem_paths = smem_paths ? 1u : 0u;
    p.trace = e->d_trace;
    p.trace_count = e->d_trace_count;
    p.trace_cap = e->trace_cap;
    if (tile_mode) {
        // tile visits of the WHOLE job (all ranks); rank r takes visits v = r (mod n_ranks)
        const uint64_t W = TILE_STEPS;
        p.n_tiles = (e->S + W - 1) / W;
        // (a path-sharded rank walks all visits of its own U_rank over its own steps)
        const uint64_t U_job = sharded ? U_rank : U;
        const uint64_t q = U_job / e->S, rU = U_job % e->S;
        const uint64_t extra = (rU + W - 1) / W;
        p.n_visits = q * p.n_tiles + extra;
        p.last_visit_terms = rU ? rU - (extra - 1) * W : W;
        p.visit_rank = sharded ? 0u : (uint32_t) e->rank;
        p.visit_nranks = sharded ? 1u : (uint32_t) e->n_ranks;
        if (peer) {
            // this rank walks ITS OWN tiles (those whose middle node it owns): U * my_steps / S terms per iteration
            p.tile_list = e->d_tile_list;
            p.n_tiles = e->my_tiles;
            p.visit_rank = 0;
            p.visit_nranks = 1;
            const uint64_t rUr = (uint64_t) ((double) rU * (double) e->my_tile_steps / (double) e->S);
            const uint64_t extra_r = (rUr + W - 1) / W;
            p.n_visits = e->my_tiles ? q * p.n_tiles + extra_r : 0;
            p.last_visit_terms = rUr ? rUr - (extra_r - 1) * W : W;
            if (p.n_tiles == 0) p.n_tiles = 1;
        }
    }
    Tile2Params t2;
    memset(&t2, 0, sizeof(t2));
    if (tile_mode && tile2) {
        // same visit arithmetic with this kernel's tile size; PGSGD_TILE_SWEEPS=T (experiment): T terms per staged step and
        // visit — every step is still a first step floor(U/S) (+1) times per iteration, but a tile is fetched q/T times
        const uint64_t W = (uint64_t) tile_steps;
        const uint64_t U_job = sharded ? U_rank : U;
        const uint64_t q = U_job / e->S, rU = U_job % e->S;
        const uint64_t extra = (rU + W - 1) / W;
        uint64_t T = 1;
        if (const char* sv = getenv("PGSGD_TILE_SWEEPS")) T = (uint64_t) atoi(sv);
        if (T < 1 || T > 64) T = 1;
        t2.n_tiles = (e->S + W - 1) / W;
        t2.sweeps = (uint32_t) T;
        t2.full_passes = q / T;
        t2.rem_sweeps = (uint32_t) (q % T);
        t2.n_visits = (t2.full_passes + (t2.rem_sweeps ? 1 : 0)) * t2.n_tiles + extra;
        t2.last_visit_terms = rU ? rU - (extra - 1) * W : W;
        t2.visit_rank = p.visit_rank;
        t2.visit_nranks = p.visit_nranks;
        t2.steps = e->d_steps; t2.xy = e->d_xy; t2.x1d = e->d_x1d; t2.frozen = p.frozen;
        t2.rng = e->d_rng; t2.rng_stride = e->rng_stride;
        t2.path_first = e->d_path_first; t2.step_count = e->S; t2.path_count = (uint32_t) e->P;
        t2.smem_paths = smem_paths ? 1u : 0u;
        t2.space = (uint32_t) (cfg->space < 0xFFFFFFFFull ? cfg->space : 0xFFFFFFFFull);   // jump spaces never exceed the longest path (< 2^31 steps)
        t2.space_max = (uint32_t) (cfg->space_max < 0xFFFFFFFFull ? cfg->space_max : 0xFFFFFFFFull);
        t2.space_q = (uint32_t) (cfg->space_quantization_step < 0xFFFFFFFFull ? cfg->space_quantization_step : 0xFFFFFFFFull);
        t2.space_q_rcp = (cfg->space < (1ull << 24) && cfg->space_quantization_step >= 16) ? 1.0f / (float) t2.space_q : 0.0f;
        t2.pos32 = e->max_path_bp < (1ull << 32) ? 1u : 0u;
        t2.flags = p.flags;
        t2.delta_max_bits = p.delta_max_bits;
        t2.counted = p.counted;
        t2.trace = p.trace; t2.trace_count = p.trace_count; t2.trace_cap = p.trace_cap;
        if (e->ztab_dims != dims) {
            rc = upload_ztab(e, 0, zetas, cfg->theta);
            if (!rc && dims == 1) rc = upload_ztab(e, 1, zetas, 0.001);   // 1D cooling: adj_theta with the zetas of the original theta (path_sgd.cpp:195,246)
            if (rc) return rc;
            e->ztab_dims = dims;
        }
    }
    if (peer) {
        p.n_parts = (uint32_t) e->n_ranks;
        for (int q = 0; q < 9; ++q) p.part_lo[q] = e->part_lo[q];
        for (int q = 0; q < 8; ++q) { p.part_xy[q] = e->peer_xy[q]; p.part_x1d[q] = e->peer_x1d[q]; }
        if (tile_mode && tile2) {   // the same walk over this rank's own tiles as the legacy kernel's (set above in p)
            t2.n_parts = p.n_parts;
            for (int q = 0; q < 9; ++q) t2.part_lo[q] = p.part_lo[q];
            for (int q = 0; q < 8; ++q) { t2.part_xy[q] = p.part_xy[q]; t2.part_x1d[q] = p.part_x1d[q]; }
            t2.tile_list = p.tile_list;
            t2.n_tiles = p.n_tiles;
            t2.n_visits = p.n_visits;
            t2.last_visit_terms = p.last_visit_terms;
            t2.visit_rank = p.visit_rank;
            t2.visit_nranks = p.visit_nranks;
            t2.sweeps = 1; t2.rem_sweeps = 0; t2.full_passes = UINT64_MAX;   // one term per staged step and visit, every pass
        }
        rc = comm_barrier(e);  // nobody starts before every slice is in place
        if (rc) return rc;
    }

    const bool window_order = (cfg->flags & PGSGD_FLAG_WINDOW_TILES) && tile_mode && tile2 && !peer && tile_steps == TILE_STEPS;
    if (window_order) {
        unsigned C = 3;
        if (const char* sv = getenv("PGSGD_WINDOW_C")) C = (unsigned) atoi(sv);
        if (C < 1 || C > 64) C = 3;
        rc = build_window_list(e, (unsigned) shape.grid * (unsigned) t2.visit_nranks, C, cfg->seed);
        if (rc) return rc;
        t2.tile_list = e->d_window_list;
    }
    if ((cfg->flags & PGSGD_FLAG_L2_WINDOW) && tile_mode && tile2 && !peer) {
        // pin the coordinate array in L2: persisting carve-out + access-policy window on this stream (reset after the loop)
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, e->device));
        const size_t bytes = dims == 2 ? 4 * e->N * sizeof(float) : e->N * sizeof(double);
        const size_t carve = std::min<size_t>((size_t) prop.persistingL2CacheMaxSize, bytes);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = dims == 2 ? (void*) e->d_xy : (void*) e->d_x1d;
        attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t) prop.accessPolicyMaxWindowSize);
        attr.accessPolicyWindow.hitRatio = bytes ? std::min(1.0f, (float) carve / (float) bytes) : 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();
    }
    CU(cudaEventRecord(e->ev0, e->stream));
    uint64_t iter = iter_begin;
    for (; iter < n_iters; ++iter) {
        p.eta = etas[iter];
        double theta_zipf = cfg->theta;
        if (dims == 1) {
            p.sp.cooling = iter > first_cooling_iteration;   // path_sgd.cpp:194
            if (p.sp.cooling) theta_zipf = 0.001;            // adj_theta (path_sgd.cpp:195,246); zetas keep the original theta
        } else {
            p.sp.cooling = iter >= first_cooling_iteration;  // path_sgd_layout.cpp:153 (adj_theta is unused in 2D, :213)
        }
        p.sp.zipf = make_zipf_const(theta_zipf);
        p.sp.zipf_f = make_zipf_const_f(p.sp.zipf);
        if (tile_mode) {
            // one keyed pseudo-random permutation of the tile index per pass (tile_perm, pgsgd_device.cuh); sweep order
            // (experiments): path order from a random offset
            const uint64_t nt = tile2 ? t2.n_tiles : p.n_tiles;
            uint64_t sm = cfg->seed ^ (0x9e3779b97f4a7c15ULL * (iter + 1));
            const bool sweep = tile2 && ((cfg->flags & PGSGD_FLAG_SWEEP_TILES) || window_order);   // window order: the list IS the order
            for (int k = 0; k < 16; ++k) {
                const uint64_t key = splitmix64_next(sm);
                p.perm_mul[k] = t2.perm_mul[k] = sweep ? 0 : 1;
                p.perm_add[k] = t2.perm_add[k] = sweep ? key % nt : key;
            }
            if (tile2) {
                t2.eta = p.eta;
                t2.eta_f = (float) p.eta;
                t2.cooling = p.sp.cooling;
                t2.one_minus_theta = p.sp.zipf_f.one_minus_theta;
                t2.alpha_frac = p.sp.zipf_f.alpha_frac;
                t2.alpha_int = p.sp.zipf_f.alpha_int;
                t2.thresh2 = p.sp.zipf_f.thresh2;
                t2.ztab = e->d_ztab[(dims == 1 && p.sp.cooling) ? 1 : 0];
            }
        }
        if (track_delta) CU(cudaMemsetAsync(e->d_delta, 0, sizeof(unsigned int), e->stream));
        if (sum_deltas) {
            if (dims == 2) CU(cudaMemcpyAsync(e->d_xy_prev, e->d_xy, 4 * e->N * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
            else CU(cudaMemcpyAsync(e->d_x1d_prev, e->d_x1d, e->N * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
        }
        if (tile_mode && tile2) CU(launch_tile2_iteration(dims, tile_steps, (cfg->flags & PGSGD_FLAG_TMA_STAGING) != 0, t2, shape, e->stream));
        else if (tile_mode) CU(launch_tile_iteration(dims, batch, p, shape, e->stream));
        else CU(launch_iteration(dims, batch, p, shape, e->stream));
        ++st.kernel_launches;
        if (e->comm) {   // kernel | collective split of this iteration (multi-GPU runs only)
            while (e->ev_pool.size() < 2 * (iter - iter_begin + 1)) {
                cudaEvent_t ev;
                CU(cudaEventCreate(&ev));
                e->ev_pool.push_back(ev);
            }
            CU(cudaEventRecord(e->ev_pool[2 * (iter - iter_begin)], e->stream));
        }
        if (peer) {
            // no coordinate traffic here: every update already went to its owner through NVLink.  One 4-byte all-reduce per
            // iteration keeps the ranks in the same cooling-schedule step (and carries the early-stop statistic).
            NC(ncclAllReduce(e->d_delta, e->d_delta, 1, ncclUint32, ncclMax, e->comm, e->stream));
        } else if (e->comm) {
            // one collective per cooling-schedule step: coordinates are replicated, term updates are sharded
            if (dims == 2) {
                if (sum_deltas) {
                    CU(launch_sub_f32(e->d_xy, e->d_xy, e->d_xy_prev, 4 * e->N, e->stream));
                    NC(ncclAllReduce(e->d_xy, e->d_xy, 4 * e->N, ncclFloat, ncclSum, e->comm, e->stream));
                    CU(launch_add_f32(e->d_xy, e->d_xy, e->d_xy_prev, 4 * e->N, e->stream));
                } else {
                    NC(ncclAllReduce(e->d_xy, e->d_xy, 4 * e->N, ncclFloat, ncclAvg, e->comm, e->stream));
                }
            } else {
                if (sum_deltas) {
                    CU(launch_sub_f64(e->d_x1d, e->d_x1d, e->d_x1d_prev, e->N, e->stream));
                    NC(ncclAllReduce(e->d_x1d, e->d_x1d, e->N, ncclDouble, ncclSum, e->comm, e->stream));
                    CU(launch_add_f64(e->d_x1d, e->d_x1d, e->d_x1d_prev, e->N, e->stream));
                } else {
                    NC(ncclAllReduce(e->d_x1d, e->d_x1d, e->N, ncclDouble, ncclAvg, e->comm, e->stream));
                }
            }
            if (track_delta) NC(ncclAllReduce(e->d_delta, e->d_delta, 1, ncclUint32, ncclMax, e->comm, e->stream));
        }
        if (e->comm) CU(cudaEventRecord(e->ev_pool[2 * (iter - iter_begin) + 1], e->stream));
        if (track_delta) {
            // early stop (checker_lambda: path_sgd_layout.cpp:142, path_sgd.cpp:183): Delta_max <= delta
            unsigned int bits = 0;
            CU(cudaMemcpyAsync(&bits, e->d_delta, sizeof(bits), cudaMemcpyDeviceToHost, e->stream));
            CU(cudaStreamSynchronize(e->stream));
            float dm;
            memcpy(&dm, &bits, sizeof(dm));
            st.last_delta_max = dm;
            if (iter + 1 < n_iters && (double) dm <= cfg->delta) { ++iter; break; }
        }
    }
    CU(cudaEventRecord(e->ev1, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if ((cfg->flags & PGSGD_FLAG_L2_WINDOW) && tile_mode && tile2 && !peer) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    unsigned long long counted = 0;
    CU(cudaMemcpy(&counted, e->d_counted, sizeof(counted), cudaMemcpyDeviceToHost));
    st.iterations_run = iter - iter_begin;
    st.term_updates = counted;
    st.seconds_iterations = ms * 1e-3;
    st.seconds_kernels = st.seconds_iterations;
    if (e->comm && st.iterations_run) {
        double coll = 0;
        for (uint64_t k = 0; k < st.iterations_run && 2 * k + 1 < e->ev_pool.size(); ++k) {
            float t = 0;
            if (cudaEventElapsedTime(&t, e->ev_pool[2 * k], e->ev_pool[2 * k + 1]) == cudaSuccess) coll += t * 1e-3;
        }
        cudaGetLastError();
        st.seconds_collectives = coll;
        st.seconds_kernels = st.seconds_iterations - coll;
    }
    if (stats) *stats = st;
    return PGSGD_OK;
}

// Picks the multi-GPU phase(s).  HYBRID: the first third of the schedule — where every update saturates (mu = 1) and half
// of the partners are uniform over the path, i.e. mostly on another GPU — runs replicated with one all-reduce per
// iteration; from then on the replica is scattered into node-range slices and all GPUs run ONE shared Hogwild through
// NVLink peer memory.  Final stress equals the single-GPU one (oracle emulation + tests/test_gpu_multi.py), at most
// of the all-reduce mode's throughput in the early phase.
int run_engine(pgsgd_engine* e, const pgsgd_config* cfg, int dims, uint64_t iter_begin, uint64_t iter_end, pgsgd_stats* stats) {
    const int mode = e->comm ? e->mode : PGSGD_MULTI_ALLREDUCE;
    if (e->shard_global_steps && e->comm && mode != PGSGD_MULTI_ALLREDUCE)
        return fail(PGSGD_ERR_STATE, "path-sharded step records need PGSGD_MULTI_ALLREDUCE (peer phases walk tiles by node range)");
    if (e->comm && mode == PGSGD_MULTI_SINGLE) {
        // Rank 0 runs the job as if it were alone (the single-GPU result, exactly); the others wait for the broadcast of the
        // coordinates.  For graphs that do not survive any way of sharing the work (shallow ones) and are, being
        // shallow, a fraction of a second of work on one GPU.
        CU(cudaSetDevice(e->device));
        pgsgd_stats st;
        memset(&st, 0, sizeof(st));
        int rc = PGSGD_OK;
        CU(cudaEventRecord(e->ev_r0, e->stream));
        if (e->rank == 0) {
            ncclComm_t comm = e->comm;
            const int nr = e->n_ranks;
            e->comm = nullptr; e->n_ranks = 1;
            e->active_mode = PGSGD_MULTI_ALLREDUCE;
            rc = run_phase(e, cfg, dims, iter_begin, iter_end, &st);
            e->comm = comm; e->n_ranks = nr;
        }
        // what the others need to know: how many iterations ran (early stop) — travels in the delta word
        unsigned int word = e->rank == 0 ? (rc ? 0xFFFFFFFFu : (unsigned int) st.iterations_run) : 0u;
        CU(cudaMemcpyAsync(e->d_delta, &word, sizeof(word), cudaMemcpyHostToDevice, e->stream));
        NC(ncclBroadcast(e->d_delta, e->d_delta, 1, ncclUint32, 0, e->comm, e->stream));
        if (dims == 2) NC(ncclBroadcast(e->d_xy, e->d_xy, 4 * e->N, ncclFloat, 0, e->comm, e->stream));
        else NC(ncclBroadcast(e->d_x1d, e->d_x1d, e->N, ncclDouble, 0, e->comm, e->stream));
        CU(cudaMemcpyAsync(&word, e->d_delta, sizeof(word), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaEventRecord(e->ev_r1, e->stream));
        CU(cudaEventSynchronize(e->ev_r1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e->ev_r0, e->ev_r1));
        if (rc) return rc;
        if (word == 0xFFFFFFFFu) return fail(PGSGD_ERR_STATE, "rank 0 of a PGSGD_MULTI_SINGLE run failed");
        if (e->rank != 0) { st.iterations_run = word; st.seconds_upload = e->seconds_upload; st.h2d_bytes = e->h2d_bytes; }
        st.seconds_collectives = ms * 1e-3 - (e->rank == 0 ? st.seconds_iterations : 0.0);
        st.seconds_iterations = ms * 1e-3;
        if (stats) *stats = st;
        return PGSGD_OK;
    }
    if (!e->comm || mode != PGSGD_MULTI_HYBRID) {
        e->active_mode = mode;
        return run_phase(e, cfg, dims, iter_begin, iter_end, stats);
    }
    int rc = check_config(cfg);
    if (rc) return rc;
    CU(cudaSetDevice(e->device));
    const uint64_t n_iters = dims == 1 ? cfg->iter_max + 1 : cfg->iter_max;
    if (iter_end > n_iters) iter_end = n_iters;
    const uint64_t sw = cfg->multi_switch_iteration ? cfg->multi_switch_iteration : cfg->iter_max / 3;
    pgsgd_stats a, b;
    memset(&a, 0, sizeof(a));
    memset(&b, 0, sizeof(b));
    bool ran_a = false, ran_b = false;
    // one event pair around BOTH phases: the phase switch (slice scatter + barrier) is part of the job and of its time
    CU(cudaEventRecord(e->ev_r0, e->stream));
    if (iter_begin < sw && !e->coords_in_slices) {
        e->active_mode = PGSGD_MULTI_ALLREDUCE;
        rc = run_phase(e, cfg, dims, iter_begin, iter_end < sw ? iter_end : sw, &a);
        if (rc) return rc;
        ran_a = true;
        if (a.iterations_run < (iter_end < sw ? iter_end : sw) - iter_begin) { if (stats) *stats = a; return PGSGD_OK; }  // early stop
    }
    if (iter_end > sw || e->coords_in_slices) {
        if (!e->coords_in_slices) { rc = peer_scatter(e, dims); if (rc) return rc; }
        e->active_mode = PGSGD_MULTI_PEER;
        rc = run_phase(e, cfg, dims, iter_begin > sw ? iter_begin : (ran_a ? sw : iter_begin), iter_end, &b);
        if (rc) return rc;
        ran_b = true;
    }
    CU(cudaEventRecord(e->ev_r1, e->stream));
    CU(cudaEventSynchronize(e->ev_r1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->ev_r0, e->ev_r1));
    if (stats) {
        pgsgd_stats st = ran_a ? a : b;
        if (ran_a && ran_b) {
            st.iterations_run += b.iterations_run;
            st.term_updates += b.term_updates;
            st.kernel_launches += b.kernel_launches;
            st.last_delta_max = b.last_delta_max;
            st.flags_used |= b.flags_used;
            st.sampling_used = b.sampling_used;
            st.seconds_kernels += b.seconds_kernels;
            st.seconds_collectives += b.seconds_collectives;
        }
        st.seconds_iterations = ms * 1e-3;
        *stats = st;
    }
    return PGSGD_OK;
}

// AUTO: replicas + one all-reduce per iteration where every replica still sees enough updates per node and iteration for the
// mean of the replicas to anneal like one Hogwild (deep graphs: many haplotypes per node; c4 and mid at 8 GPUs end at or
// below the single-GPU stress), else rank 0 alone + a broadcast: on the 6-haplotype longthin graph 8 replicas end at a far
// stress of 1.0 (no layout at all), the hybrid schedule at +33 % and one shared Hogwild over NVLink peer memory at +35-40 %
// (profiles/r02_multi_suite_n8.jsonl) — and such a graph is 0.2 s of work on one GPU.
int resolve_mode(const pgsgd_engine* e, int dims) {
    if (e->multi_mode != PGSGD_MULTI_AUTO) return e->multi_mode;
    if (!e->comm || e->n_ranks < 2 || e->shard_global_steps) return PGSGD_MULTI_ALLREDUCE;
    const double per_replica = (dims == 2 ? 10.0 : 1.0) * (double) e->S / (double) e->N / (double) e->n_ranks;
    return per_replica >= PGSGD_AUTO_MIN_UPDATES_PER_NODE ? PGSGD_MULTI_ALLREDUCE : PGSGD_MULTI_SINGLE;
}

// After a coordinate upload: resolve the mode, build what its phases need (peer slices + IPC mappings + tile ownership)
// and run the first full-size collective now — NCCL connects its channels lazily, and none of this belongs to an iteration.
int multi_prepare(pgsgd_engine* e, int dims) {
    e->mode = resolve_mode(e, dims);
    if (!e->comm) return PGSGD_OK;
    if (e->mode == PGSGD_MULTI_PEER || e->mode == PGSGD_MULTI_HYBRID) {
        if (!(dims == 2 ? e->peer_ready_2d : e->peer_ready_1d)) { int rc = setup_peer(e, dims); if (rc) return rc; }
    }
    if (e->mode == PGSGD_MULTI_PEER) { int rc = peer_scatter(e, dims); if (rc) return rc; }
    if (!e->comm_warm && e->mode != PGSGD_MULTI_PEER) {
        const size_t count = dims == 2 ? 4 * e->N : e->N;
        void* tmp = nullptr;
        CU(cudaMalloc(&tmp, count * (dims == 2 ? sizeof(float) : sizeof(double))));
        cudaMemsetAsync(tmp, 0, count * (dims == 2 ? sizeof(float) : sizeof(double)), e->stream);
        ncclResult_t r = ncclAllReduce(tmp, tmp, count, dims == 2 ? ncclFloat : ncclDouble, ncclSum, e->comm, e->stream);
        cudaError_t ce = cudaStreamSynchronize(e->stream);
        cudaFree(tmp);
        if (r != ncclSuccess) return fail(PGSGD_ERR_NCCL, "warm-up all-reduce: %s", ncclGetErrorString(r));
        if (ce != cudaSuccess) return fail(PGSGD_ERR_CUDA, "warm-up all-reduce: %s", cudaGetErrorString(ce));
        e->comm_warm = true;
    }
    return PGSGD_OK;
}

}  // namespace

extern "C" {

const char* pgsgd_last_error(void) { return g_last_error.c_str(); }
int pgsgd_version(void) { return PGSGD_VERSION; }

int pgsgd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pgsgd_device_warmup(int device) {
    const int ndev = pgsgd_device_count();
    if (ndev == 0) return fail(PGSGD_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(PGSGD_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    CU(cudaFree(nullptr));
    return PGSGD_OK;
}

int pgsgd_schedule(const pgsgd_config* cfg, double* etas_out) {
    if (!cfg || !etas_out || cfg->iter_max == 0) return fail(PGSGD_ERR_ARG, "pgsgd_schedule: bad arguments");
    std::vector<double> etas = build_schedule(*cfg);
    memcpy(etas_out, etas.data(), etas.size() * sizeof(double));
    return PGSGD_OK;
}

uint64_t pgsgd_zetas(const pgsgd_config* cfg, double* zetas_out, uint64_t cap) {
    if (!cfg || cfg->space_quantization_step == 0) return 0;
    const uint64_t n = zeta_table_size(*cfg);
    if (zetas_out) {
        std::vector<double> z = build_zetas(*cfg);
        memcpy(zetas_out, z.data(), (n < cap ? n : cap) * sizeof(double));
    }
    return n;
}

// pre_sn / pre_sr: step arrays that are ALREADY on the device (parsed there from GFA text, pgsgd_gfa.cu); the view's host
// step arrays are then null.  Ownership passes to this function once *pre_consumed is set.
static int engine_create_impl(const pgsgd_graph_view* g, int device, uint32_t* pre_sn, uint8_t* pre_sr, bool* pre_consumed, pgsgd_engine** out) {
    if (!g || !out) return fail(PGSGD_ERR_ARG, "pgsgd_engine_create: NULL argument");
    *out = nullptr;
    if (!g->node_len || !g->path_first_step || (!g->step_node && g->step_count && !pre_sn)) return fail(PGSGD_ERR_ARG, "graph view has NULL arrays");
    if (pre_sn && g->step_pos) return fail(PGSGD_ERR_ARG, "device-resident step arrays come without positions");
    if (g->node_count == 0) return fail(PGSGD_ERR_ARG, "graph has no nodes");
    if (g->node_count >= (1ull << 31)) return fail(PGSGD_ERR_ARG, "more than 2^31-1 nodes are not supported by the 32-bit handle field");
    if (g->path_count >= (1ull << 32)) return fail(PGSGD_ERR_ARG, "too many paths");
    if (g->path_first_step[0] != 0 || g->path_first_step[g->path_count] != g->step_count)
        return fail(PGSGD_ERR_ARG, "path_first_step must start at 0 and end at step_count");
    uint64_t max_steps = 0;
    bool any_multi = false;
    for (uint64_t p = 0; p < g->path_count; ++p) {
        if (g->path_first_step[p + 1] < g->path_first_step[p]) return fail(PGSGD_ERR_ARG, "path_first_step is not monotone at path %llu", (unsigned long long) p);
        const uint64_t c = g->path_first_step[p + 1] - g->path_first_step[p];
        if (c > max_steps) max_steps = c;
        if (c > 1) any_multi = true;
    }
    int ndev = pgsgd_device_count();
    if (ndev == 0) return fail(PGSGD_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(PGSGD_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    // The path gathers single 16/32-byte records at random addresses: ask L2 not to widen its DRAM fetches beyond one
    // 32-byte sector (the default granularity over-fetches ~3.4x on this access pattern, profiles/).  PGSGD_L2_FETCH
    // overrides (bytes: 32, 64 or 128) for experiments.
    {
        size_t gran = 32;
        if (const char* s = getenv("PGSGD_L2_FETCH")) gran = (size_t) atoi(s);
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }

    const double t0 = now_s();
    // PGSGD_TIMING=1: phase times of this call as one JSON line on stderr (where does an engine's creation go?)
    const bool timing = getenv("PGSGD_TIMING") != nullptr;
    double t_prev = t0;
    std::string timing_line;
    auto mark = [&](const char* what) {
        if (!timing) return;
        const double t = now_s();
        char buf[96];
        snprintf(buf, sizeof buf, "%s\"%s_s\": %.4f", timing_line.empty() ? "" : ", ", what, t - t_prev);
        timing_line += buf;
        t_prev = t;
    };
    pgsgd_engine* e = new pgsgd_engine();
    e->device = device;
    e->N = g->node_count; e->P = g->path_count; e->S = g->step_count;
    e->max_path_steps = max_steps;
    e->any_multi_step_path = any_multi;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    e->sm_count = prop.multiProcessorCount;
    int rc = PGSGD_OK;
    auto bail = [&](int code) { pgsgd_engine_destroy(e); return code; };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(PGSGD_ERR_CUDA, "cudaStreamCreate failed"));
    if (cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess || cudaEventCreate(&e->ev_r0) != cudaSuccess ||
        cudaEventCreate(&e->ev_r1) != cudaSuccess) return bail(fail(PGSGD_ERR_CUDA, "cudaEventCreate failed"));

    // With caller-supplied positions the node ranks are validated here; without, validation and the per-path bp offsets
    // (xp.cpp:607-616) are done on the device below — no O(S) host pass at all.
    const uint64_t* pos = g->step_pos;
    if (pos) {
        for (uint64_t s = 0; s < g->step_count; ++s) {
            if (g->step_node[s] >= g->node_count) return bail(fail(PGSGD_ERR_UNOPT, "step %llu refers to node rank %u >= node_count: ids are not compacted 1..N", (unsigned long long) s, g->step_node[s]));
        }
    }
    {
        const uint64_t W = TILE_STEPS, nt = (g->step_count + W - 1) / W;
        e->tile_mid_node.resize(nt);
        if (pre_sn) {   // the steps never were on the host: gather the tile-centre nodes on the device
            uint32_t* d_mid = nullptr;
            cudaError_t ce = cudaMalloc(&d_mid, (nt ? nt : 1) * sizeof(uint32_t));
            if (ce == cudaSuccess) ce = launch_gather_mid_nodes(pre_sn, g->step_count, W, nt, d_mid, e->stream);
            if (ce == cudaSuccess && nt) ce = cudaMemcpyAsync(e->tile_mid_node.data(), d_mid, nt * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
            cudaFree(d_mid);
            if (ce != cudaSuccess) return bail(fail(PGSGD_ERR_CUDA, "tile centre gather: %s", cudaGetErrorString(ce)));
        } else {
            for (uint64_t t = 0; t < nt; ++t) {
                const uint64_t lo = t * W, hi = lo + W < g->step_count ? lo + W : g->step_count;
                e->tile_mid_node[t] = g->step_node[lo + (hi - lo) / 2];
            }
        }
    }
    mark("host_prep");
    if ((rc = dev_alloc(e, &e->d_steps, e->S))) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_path_first, e->P + 1))) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_delta, 1))) return bail(rc);
    if ((rc = dev_alloc(e, &e->d_counted, 1))) return bail(rc);
    auto cu_bail = [&](cudaError_t err, const char* what) {
        return bail(fail(PGSGD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err)));
    };
    cudaError_t err;
    if ((err = cudaMemcpyAsync(e->d_path_first, g->path_first_step, (e->P + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, e->stream)) != cudaSuccess) return cu_bail(err, "upload path_first_step");
    e->h2d_bytes += (e->P + 1) * sizeof(uint64_t);

    // flatten-to-device: SoA chunks are staged and packed into 16-byte step records by a device kernel
    uint32_t* d_node_len = nullptr;
    if ((rc = dev_alloc(e, &e->d_node_len, e->N))) return bail(rc);
    d_node_len = e->d_node_len;
    if ((err = cudaMemcpyAsync(d_node_len, g->node_len, e->N * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream)) != cudaSuccess) { return cu_bail(err, "upload node_len"); }
    e->h2d_bytes += e->N * sizeof(uint32_t);
    uint32_t* d_sn = nullptr; uint8_t* d_sr = nullptr; uint64_t* d_sp = nullptr;
    uint32_t* d_depth = nullptr;  // steps per node, counted while packing
    unsigned long long* d_maxdup = nullptr;
    if (cudaMalloc(&d_depth, e->N * sizeof(uint32_t)) != cudaSuccess || cudaMemsetAsync(d_depth, 0, e->N * sizeof(uint32_t), e->stream) != cudaSuccess ||
        cudaMalloc(&d_maxdup, sizeof(unsigned long long)) != cudaSuccess || cudaMemsetAsync(d_maxdup, 0, sizeof(unsigned long long), e->stream) != cudaSuccess) {
        cudaFree(d_depth); cudaFree(d_maxdup);
        return bail(fail(PGSGD_ERR_NOMEM, "cudaMalloc depth table failed"));
    }
    if (pos) {
        const uint64_t CH = 1ull << 26;  // 64 Mi steps per staging chunk (832 MiB of SoA)
        const uint64_t ch = e->S < CH ? (e->S ? e->S : 1) : CH;
        bool ok = cudaMalloc(&d_sn, ch * 4) == cudaSuccess && cudaMalloc(&d_sp, ch * 8) == cudaSuccess && (!g->step_rev || cudaMalloc(&d_sr, ch) == cudaSuccess);
        if (!ok) { cudaFree(d_sn); cudaFree(d_sp); cudaFree(d_sr); return bail(fail(PGSGD_ERR_NOMEM, "cudaMalloc step staging failed")); }
        for (uint64_t off = 0; off < e->S && err == cudaSuccess; off += ch) {
            const uint64_t n = e->S - off < ch ? e->S - off : ch;
            err = cudaMemcpyAsync(d_sn, g->step_node + off, n * 4, cudaMemcpyHostToDevice, e->stream);
            if (err == cudaSuccess) err = cudaMemcpyAsync(d_sp, pos + off, n * 8, cudaMemcpyHostToDevice, e->stream);
            if (err == cudaSuccess && g->step_rev) err = cudaMemcpyAsync(d_sr, g->step_rev + off, n, cudaMemcpyHostToDevice, e->stream);
            if (err == cudaSuccess) err = launch_pack_steps(e->d_steps, d_sn, d_sr, d_sp, d_node_len, n, off, d_depth, e->stream);
            if (err == cudaSuccess) err = launch_tile_repeats(d_sn, n, d_maxdup, e->stream);  // chunks start on tile boundaries
            if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);  // staging buffers and pageable sources are reused
            e->h2d_bytes += n * (4 + 8 + (g->step_rev ? 1 : 0));
        }
    } else {
        // whole-array staging (5 B per step in, 8 B per step of scan scratch), positions by a device-wide scan
        const uint64_t n = e->S ? e->S : 1;
        int* d_bad = nullptr;
        bool ok;
        if (pre_sn) {   // parsed on the device: nothing to upload
            d_sn = pre_sn; d_sr = pre_sr;
            if (pre_consumed) *pre_consumed = true;
            ok = cudaMalloc(&d_sp, n * 8) == cudaSuccess && cudaMalloc(&d_bad, sizeof(int)) == cudaSuccess;
        } else {
            ok = cudaMalloc(&d_sn, n * 4) == cudaSuccess && cudaMalloc(&d_sp, n * 8) == cudaSuccess && cudaMalloc(&d_bad, sizeof(int)) == cudaSuccess &&
                 (!g->step_rev || cudaMalloc(&d_sr, n) == cudaSuccess);
        }
        if (!ok) { cudaFree(d_sn); cudaFree(d_sp); cudaFree(d_sr); cudaFree(d_bad); return bail(fail(PGSGD_ERR_NOMEM, "cudaMalloc step staging failed")); }
        mark("device_alloc");
        err = cudaMemsetAsync(d_bad, 0, sizeof(int), e->stream);
        if (err == cudaSuccess && !pre_sn) err = cudaMemcpyAsync(d_sn, g->step_node, e->S * 4, cudaMemcpyHostToDevice, e->stream);
        if (err == cudaSuccess && !pre_sn && g->step_rev) err = cudaMemcpyAsync(d_sr, g->step_rev, e->S, cudaMemcpyHostToDevice, e->stream);
        if (timing) { cudaStreamSynchronize(e->stream); mark("step_upload"); }
        if (err == cudaSuccess) err = launch_flatten_on_device(e->d_steps, d_sn, d_sr, d_node_len, e->d_path_first, (uint32_t) e->P, (uint32_t) e->N, e->S, d_sp, d_bad, d_depth, e->stream);
        int bad = 0;
        if (err == cudaSuccess) err = cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
        mark("flatten_on_device");
        if (err == cudaSuccess && !bad) err = launch_tile_repeats(d_sn, e->S, d_maxdup, e->stream);
        cudaFree(d_bad);
        if (!pre_sn) e->h2d_bytes += e->S * (4 + (g->step_rev ? 1 : 0));
        if (err == cudaSuccess && bad) {
            cudaFree(d_sn); cudaFree(d_sp); cudaFree(d_sr); cudaFree(d_depth);
            return bail(fail(PGSGD_ERR_UNOPT, "a step refers to a node rank >= node_count: ids are not compacted 1..N"));
        }
    }
    if (err == cudaSuccess) {
        std::vector<uint32_t> depth(e->N);
        err = cudaMemcpy(depth.data(), d_depth, e->N * sizeof(uint32_t), cudaMemcpyDeviceToHost);
        for (uint32_t d : depth) if (d > e->max_node_depth) e->max_node_depth = d;
        unsigned long long md = 0;
        if (err == cudaSuccess) err = cudaMemcpy(&md, d_maxdup, sizeof(md), cudaMemcpyDeviceToHost);
        e->tile_repeats = md;
    }
    if (err == cudaSuccess) {   // largest end-adjusted bp position (the 32-bit distance path of the pipelined tile kernel needs < 2^32)
        err = cudaMemsetAsync(d_maxdup, 0, sizeof(unsigned long long), e->stream);
        if (err == cudaSuccess) err = launch_max_path_bp(e->d_steps, e->d_path_first, (uint32_t) e->P, d_maxdup, e->stream);
        unsigned long long mb = 0;
        if (err == cudaSuccess) err = cudaMemcpyAsync(&mb, d_maxdup, sizeof(mb), cudaMemcpyDeviceToHost, e->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
        e->max_path_bp = mb;
    }
    mark("graph_statistics");
    cudaFree(d_maxdup);
    cudaFree(d_sn); cudaFree(d_sp); cudaFree(d_sr); cudaFree(d_depth);
    if (err != cudaSuccess) return cu_bail(err, "flatten-to-device");
    if ((err = cudaStreamSynchronize(e->stream)) != cudaSuccess) return cu_bail(err, "engine create sync");
    mark("free_staging");
    e->seconds_upload = now_s() - t0;
    if (timing) fprintf(stderr, "{\"pgsgd_engine_create\": {%s, \"total_s\": %.4f}}\n", timing_line.c_str(), e->seconds_upload);
    *out = e;
    return PGSGD_OK;
}

int pgsgd_engine_create(const pgsgd_graph_view* g, int device, pgsgd_engine** out) {
    return engine_create_impl(g, device, nullptr, nullptr, nullptr, out);
}

// The step lists live in pageable (mmap'ed) host memory: a plain cudaMemcpy from there goes through the driver's own staging at
// 3-4 GB/s.  Here host threads copy 16 MB chunks of the packed destination range into their own pinned buffers (two each) and
// send them on their own streams, so that the host copies, and the PCIe transfers of different chunks, overlap.
static cudaError_t upload_fields_staged(char* d_text, const char* text, const uint64_t* field_begin, const std::vector<uint64_t>& packed_begin,
                                        uint64_t path_count, int device, cudaStream_t stream) {
    const uint64_t n_bytes = packed_begin[path_count];
    if (n_bytes == 0) return cudaSuccess;
    uint64_t CH = 16ull << 20;
    if (const char* sv = getenv("PGSGD_UPLOAD_CHUNK")) { const long long v = atoll(sv); if (v >= 256 && v <= (1ll << 30)) CH = (uint64_t) v; }   // tests: many chunks on a small file
    const uint64_t n_chunks = (n_bytes + CH - 1) / CH;
    unsigned T = std::thread::hardware_concurrency();
    T = T >= 16 ? 6 : (T >= 8 ? 4 : 2);
    if (const char* sv = getenv("PGSGD_UPLOAD_THREADS")) { const int v = atoi(sv); if (v >= 1 && v <= 32) T = (unsigned) v; }
    if (T > n_chunks) T = (unsigned) n_chunks;
    std::vector<cudaError_t> err(T, cudaSuccess);
    auto worker = [&](unsigned t) {
        cudaError_t e = cudaSetDevice(device);
        char* buf[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaStream_t st = nullptr;
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaHostAlloc((void**) &buf[b], CH, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
        }
        unsigned k = 0;
        for (uint64_t c = t; c < n_chunks && e == cudaSuccess; c += T, ++k) {
            const int b = (int) (k & 1);
            if (k >= 2) e = cudaEventSynchronize(ev[b]);   // the transfer that last used this buffer is done
            if (e != cudaSuccess) break;
            const uint64_t lo = c * CH, hi = std::min(n_bytes, lo + CH);
            // fields overlapping [lo, hi): packed_begin is sorted
            uint64_t p = (uint64_t) (std::upper_bound(packed_begin.begin(), packed_begin.end(), lo) - packed_begin.begin()) - 1;
            for (; p < path_count && packed_begin[p] < hi; ++p) {
                const uint64_t a = std::max(lo, packed_begin[p]), z = std::min(hi, packed_begin[p + 1]);
                if (z > a) memcpy(buf[b] + (a - lo), text + field_begin[p] + (a - packed_begin[p]), z - a);
            }
            e = cudaMemcpyAsync(d_text + lo, buf[b], hi - lo, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
        }
        if (st) { const cudaError_t e2 = cudaStreamSynchronize(st); if (e == cudaSuccess) e = e2; }
        for (int b = 0; b < 2; ++b) { if (ev[b]) cudaEventDestroy(ev[b]); if (buf[b]) cudaFreeHost(buf[b]); }
        if (st) cudaStreamDestroy(st);
        err[t] = e;
    };
    cudaError_t e0 = cudaStreamSynchronize(stream);   // the memset of the padding precedes the copies
    if (e0 != cudaSuccess) return e0;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
    for (cudaError_t e : err) if (e != cudaSuccess) return e;
    return cudaSuccess;
}

int pgsgd_engine_create_from_gfa_paths(const uint32_t* node_len, uint64_t node_count, const char* text, const uint64_t* field_begin,
                                       const uint64_t* field_end, uint64_t path_count, int device, pgsgd_engine** out) {
    if (!node_len || !out || (path_count && (!text || !field_begin || !field_end))) return fail(PGSGD_ERR_ARG, "create_from_gfa_paths: NULL argument");
    *out = nullptr;
    if (node_count == 0 || node_count >= (1ull << 31)) return fail(PGSGD_ERR_ARG, "create_from_gfa_paths: node_count out of range");
    if (path_count >= (1ull << 32)) return fail(PGSGD_ERR_ARG, "too many paths");
    int ndev = pgsgd_device_count();
    if (ndev == 0) return fail(PGSGD_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(PGSGD_ERR_ARG, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    const double t0 = now_s();
    // pack the step lists back to back (every list ends with an orientation character, the next one starts with a digit)
    std::vector<uint64_t> packed_begin(path_count + 1, 0);
    for (uint64_t p = 0; p < path_count; ++p) {
        if (field_end[p] < field_begin[p]) return fail(PGSGD_ERR_ARG, "create_from_gfa_paths: field %llu ends before it begins", (unsigned long long) p);
        packed_begin[p + 1] = packed_begin[p] + (field_end[p] - field_begin[p]);
    }
    const uint64_t n_bytes = packed_begin[path_count], n_alloc = (n_bytes + 15) / 16 * 16 + 16;
    char* d_text = nullptr;
    uint64_t *d_fb = nullptr, *d_pf = nullptr;
    uint32_t* d_sn = nullptr;
    uint8_t* d_sr = nullptr;
    cudaStream_t stream = nullptr;
    auto cleanup = [&]() { cudaFree(d_text); cudaFree(d_fb); cudaFree(d_pf); if (stream) cudaStreamDestroy(stream); };
    cudaError_t ce = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_text, n_alloc);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(d_text + (n_alloc - 32 < n_alloc ? n_alloc - 32 : 0), 0, n_alloc < 32 ? n_alloc : 32, stream);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_fb, (path_count + 1) * sizeof(uint64_t));
    if (ce == cudaSuccess) ce = cudaMalloc(&d_pf, (path_count + 1) * sizeof(uint64_t));
    if (ce == cudaSuccess) ce = upload_fields_staged(d_text, text, field_begin, packed_begin, path_count, device, stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_fb, packed_begin.data(), (path_count + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream);
    uint64_t S = 0;
    int bad = 0;
    if (ce == cudaSuccess) ce = launch_parse_gfa_paths(d_text, n_bytes, d_fb, (uint32_t) path_count, (uint32_t) node_count, d_pf, &d_sn, &d_sr, &S, &bad, stream);
    std::vector<uint64_t> path_first(path_count + 1, 0);
    if (ce == cudaSuccess) ce = cudaMemcpy(path_first.data(), d_pf, (path_count + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    cleanup();
    if (ce != cudaSuccess) { cudaFree(d_sn); cudaFree(d_sr); return fail(ce == cudaErrorMemoryAllocation ? PGSGD_ERR_NOMEM : PGSGD_ERR_CUDA, "create_from_gfa_paths: %s", cudaGetErrorString(ce)); }
    if (bad) { cudaFree(d_sn); cudaFree(d_sr); return fail(PGSGD_ERR_UNOPT, "a path step is not `<id>+` / `<id>-` with 1 <= id <= node_count: node ids must be the numbers 1..N"); }
    pgsgd_graph_view v;
    memset(&v, 0, sizeof(v));
    v.node_count = node_count; v.path_count = path_count; v.step_count = S;
    v.node_len = node_len; v.path_first_step = path_first.data();
    bool consumed = false;
    int rc = engine_create_impl(&v, device, d_sn, d_sr, &consumed, out);
    if (!consumed) { cudaFree(d_sn); cudaFree(d_sr); }
    if (rc == PGSGD_OK) {
        (*out)->h2d_bytes += n_bytes;
        (*out)->seconds_upload = now_s() - t0;
    }
    return rc;
}

int pgsgd_engine_graph_stats(const pgsgd_engine* e, uint64_t* step_count, uint64_t* max_path_steps, uint64_t* max_path_bp, uint64_t* max_node_depth) {
    if (!e) return fail(PGSGD_ERR_ARG, "graph_stats: NULL engine");
    if (step_count) *step_count = e->S;
    if (max_path_steps) *max_path_steps = e->max_path_steps;
    if (max_path_bp) *max_path_bp = e->max_path_bp;
    if (max_node_depth) *max_node_depth = e->max_node_depth;
    return PGSGD_OK;
}

void pgsgd_engine_destroy(pgsgd_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* q : e->ipc_opened) cudaIpcCloseMemHandle(q);
    cudaFree(e->d_xy_part); cudaFree(e->d_x1d_part); cudaFree(e->d_tile_list); cudaFree(e->d_window_list); cudaFree(e->d_stage_xy[0]); cudaFree(e->d_stage_xy[1]);
    if (e->comm && !e->comm_cached) ncclCommDestroy(e->comm);
    cudaFree(e->d_steps); cudaFree(e->d_path_first); cudaFree(e->d_node_len); cudaFree(e->d_xy); cudaFree(e->d_xy_prev); cudaFree(e->d_x1d);
    cudaFree(e->d_trace); cudaFree(e->d_trace_count);
    cudaFree(e->d_ztab[0]); cudaFree(e->d_ztab[1]);
    cudaFree(e->d_x1d_prev); cudaFree(e->d_frozen); cudaFree(e->d_zetas); cudaFree(e->d_rng); cudaFree(e->d_delta); cudaFree(e->d_counted);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_r0) cudaEventDestroy(e->ev_r0);
    if (e->ev_r1) cudaEventDestroy(e->ev_r1);
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int pgsgd_engine_device(const pgsgd_engine* e) { return e ? e->device : -1; }
uint64_t pgsgd_engine_device_bytes(const pgsgd_engine* e) { return e ? e->bytes : 0; }

static int stage_xy(pgsgd_engine* e) {
    for (int k = 0; k < 2; ++k)
        if (!e->d_stage_xy[k]) { int rc = dev_alloc(e, &e->d_stage_xy[k], 2 * e->N); if (rc) return rc; }
    return PGSGD_OK;
}

int pgsgd_engine_set_coords_2d(pgsgd_engine* e, const double* X, const double* Y) {
    if (!e || !X || !Y) return fail(PGSGD_ERR_ARG, "set_coords_2d: NULL argument");
    CU(cudaSetDevice(e->device));
    const double t0 = now_s();
    if (!e->d_xy) { int rc = dev_alloc(e, &e->d_xy, 4 * e->N); if (rc) return rc; }
    { int rc = stage_xy(e); if (rc) return rc; }
    double *dX = e->d_stage_xy[0], *dY = e->d_stage_xy[1];
    cudaError_t err = cudaMemcpyAsync(dX, X, 2 * e->N * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dY, Y, 2 * e->N * sizeof(double), cudaMemcpyHostToDevice, e->stream);
    if (err == cudaSuccess) err = launch_xy_from_XY(e->d_xy, dX, dY, e->N, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "set_coords_2d: %s", cudaGetErrorString(err));
    e->coords_in_slices = false;
    { int rc = multi_prepare(e, 2); if (rc) return rc; }
    e->have_2d = true;
    e->h2d_bytes += 4 * e->N * sizeof(double);
    e->seconds_upload += now_s() - t0;
    return PGSGD_OK;
}

int pgsgd_engine_get_coords_2d(pgsgd_engine* e, double* X, double* Y) {
    if (!e || !X || !Y) return fail(PGSGD_ERR_ARG, "get_coords_2d: NULL argument");
    if (!e->have_2d) return fail(PGSGD_ERR_STATE, "no 2D coordinates on the device");
    CU(cudaSetDevice(e->device));
    if (e->coords_in_slices && e->peer_ready_2d) { int rc = peer_gather(e, 2); if (rc) return rc; }
    { int rc = stage_xy(e); if (rc) return rc; }
    double *dX = e->d_stage_xy[0], *dY = e->d_stage_xy[1];
    cudaError_t err = launch_XY_from_xy(dX, dY, e->d_xy, e->N, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(X, dX, 2 * e->N * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(Y, dY, 2 * e->N * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "get_coords_2d: %s", cudaGetErrorString(err));
    return PGSGD_OK;
}

int pgsgd_engine_set_coords_2d_f32(pgsgd_engine* e, const float* xy) {
    if (!e || !xy) return fail(PGSGD_ERR_ARG, "set_coords_2d_f32: NULL argument");
    CU(cudaSetDevice(e->device));
    if (!e->d_xy) { int rc = dev_alloc(e, &e->d_xy, 4 * e->N); if (rc) return rc; }
    CU(cudaMemcpyAsync(e->d_xy, xy, 4 * e->N * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->coords_in_slices = false;
    { int rc = multi_prepare(e, 2); if (rc) return rc; }
    e->have_2d = true;
    e->h2d_bytes += 4 * e->N * sizeof(float);
    return PGSGD_OK;
}

int pgsgd_engine_get_coords_2d_f32(pgsgd_engine* e, float* xy) {
    if (!e || !xy) return fail(PGSGD_ERR_ARG, "get_coords_2d_f32: NULL argument");
    if (!e->have_2d) return fail(PGSGD_ERR_STATE, "no 2D coordinates on the device");
    CU(cudaSetDevice(e->device));
    if (e->coords_in_slices && e->peer_ready_2d) { int rc = peer_gather(e, 2); if (rc) return rc; }
    CU(cudaMemcpyAsync(xy, e->d_xy, 4 * e->N * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PGSGD_OK;
}

int pgsgd_engine_set_coords_1d(pgsgd_engine* e, const double* X) {
    if (!e) return fail(PGSGD_ERR_ARG, "set_coords_1d: NULL engine");
    CU(cudaSetDevice(e->device));
    if (!e->d_x1d) {   // one spare double: the pipelined kernel fetches 1D coordinates as aligned 16-byte pairs, so an odd N reads one past the end
        int rc = dev_alloc(e, &e->d_x1d, e->N + 2); if (rc) return rc;
        CU(cudaMemsetAsync(e->d_x1d + e->N, 0, 2 * sizeof(double), e->stream));
    }
    if (!X && e->h_x1d_default.empty()) {   // default initialisation, built on first use: cumulative bp of the node order (path_sgd.cpp:63-69)
        std::vector<uint32_t> len(e->N);
        CU(cudaMemcpy(len.data(), e->d_node_len, e->N * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        e->h_x1d_default.resize(e->N);
        uint64_t sum = 0;
        for (uint64_t r = 0; r < e->N; ++r) { e->h_x1d_default[r] = (double) sum; sum += len[r]; }
    }
    const double* src = X ? X : e->h_x1d_default.data();
    CU(cudaMemcpyAsync(e->d_x1d, src, e->N * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->coords_in_slices = false;
    { int rc = multi_prepare(e, 1); if (rc) return rc; }
    e->have_1d = true;
    e->h2d_bytes += e->N * sizeof(double);
    return PGSGD_OK;
}

int pgsgd_engine_get_coords_1d(pgsgd_engine* e, double* X) {
    if (!e || !X) return fail(PGSGD_ERR_ARG, "get_coords_1d: NULL argument");
    if (!e->have_1d) return fail(PGSGD_ERR_STATE, "no 1D coordinates on the device");
    CU(cudaSetDevice(e->device));
    if (e->coords_in_slices && e->peer_ready_1d) { int rc = peer_gather(e, 1); if (rc) return rc; }
    CU(cudaMemcpyAsync(X, e->d_x1d, e->N * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PGSGD_OK;
}

int pgsgd_engine_set_frozen_1d(pgsgd_engine* e, const uint8_t* frozen) {
    if (!e) return fail(PGSGD_ERR_ARG, "set_frozen_1d: NULL engine");
    CU(cudaSetDevice(e->device));
    if (!frozen) {
        if (e->d_frozen) { cudaFree(e->d_frozen); e->d_frozen = nullptr; }
        return PGSGD_OK;
    }
    if (!e->d_frozen) { int rc = dev_alloc(e, &e->d_frozen, e->N); if (rc) return rc; }
    CU(cudaMemcpyAsync(e->d_frozen, frozen, e->N, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return PGSGD_OK;
}

int pgsgd_engine_run_2d(pgsgd_engine* e, const pgsgd_config* cfg, pgsgd_stats* stats) {
    if (!e) return fail(PGSGD_ERR_ARG, "run_2d: NULL engine");
    return run_engine(e, cfg, 2, 0, UINT64_MAX, stats);
}
int pgsgd_engine_run_1d(pgsgd_engine* e, const pgsgd_config* cfg, pgsgd_stats* stats) {
    if (!e) return fail(PGSGD_ERR_ARG, "run_1d: NULL engine");
    return run_engine(e, cfg, 1, 0, UINT64_MAX, stats);
}
int pgsgd_engine_run_range(pgsgd_engine* e, const pgsgd_config* cfg, int dims, uint64_t iter_begin, uint64_t iter_end, pgsgd_stats* stats) {
    if (!e) return fail(PGSGD_ERR_ARG, "run_range: NULL engine");
    if (dims != 1 && dims != 2) return fail(PGSGD_ERR_ARG, "dims must be 1 or 2");
    return run_engine(e, cfg, dims, iter_begin, iter_end, stats);
}

int pgsgd_engine_set_trace(pgsgd_engine* e, uint64_t capacity) {
    if (!e) return fail(PGSGD_ERR_ARG, "set_trace: NULL engine");
    CU(cudaSetDevice(e->device));
    if (e->d_trace) { cudaFree(e->d_trace); e->d_trace = nullptr; }
    if (!e->d_trace_count) CU(cudaMalloc(&e->d_trace_count, sizeof(unsigned long long)));
    CU(cudaMemset(e->d_trace_count, 0, sizeof(unsigned long long)));
    e->trace_cap = capacity;
    if (capacity) CU(cudaMalloc(&e->d_trace, capacity * 2 * sizeof(unsigned long long)));
    return PGSGD_OK;
}

int pgsgd_engine_get_trace(pgsgd_engine* e, uint64_t* ia_out, uint64_t* ib_out, uint8_t* flips_out, uint64_t* n_out) {
    if (!e || !n_out) return fail(PGSGD_ERR_ARG, "get_trace: NULL argument");
    if (!e->d_trace) return fail(PGSGD_ERR_STATE, "no trace buffer (pgsgd_engine_set_trace)");
    CU(cudaSetDevice(e->device));
    unsigned long long n = 0;
    CU(cudaMemcpy(&n, e->d_trace_count, sizeof(n), cudaMemcpyDeviceToHost));
    if (n > e->trace_cap) n = e->trace_cap;
    std::vector<unsigned long long> h(2 * n);
    if (n) CU(cudaMemcpy(h.data(), e->d_trace, 2 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (uint64_t k = 0; k < n; ++k) {
        if (ia_out) ia_out[k] = h[2 * k];
        if (ib_out) ib_out[k] = h[2 * k + 1] & 0x3FFFFFFFFFFFFFFFull;
        if (flips_out) flips_out[k] = (uint8_t) (h[2 * k + 1] >> 62);
    }
    *n_out = n;
    CU(cudaMemset(e->d_trace_count, 0, sizeof(unsigned long long)));
    return PGSGD_OK;
}

int pgsgd_comm_unique_id(uint8_t id_out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    NC(ncclGetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return PGSGD_OK;
}

int pgsgd_engine_attach_comm(pgsgd_engine* e, const uint8_t unique_id[128], int n_ranks, int rank) {
    if (!e || !unique_id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(PGSGD_ERR_ARG, "attach_comm: bad arguments");
    CU(cudaSetDevice(e->device));
    if (e->comm && !e->comm_cached) ncclCommDestroy(e->comm);
    e->comm = nullptr;
    e->comm_cached = false;
    e->comm_warm = false;
    e->n_ranks = n_ranks;
    e->rank = rank;
    if (n_ranks == 1) return PGSGD_OK;
    const CommKey key{e->device, n_ranks, rank};
    if (comm_cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        auto it = g_comm_cache.find(key);
        if (it != g_comm_cache.end()) {
            e->comm = it->second;
            e->comm_cached = true;
            e->comm_warm = true;
            return PGSGD_OK;
        }
    }
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    NC(ncclCommInitRank(&e->comm, n_ranks, id, rank));
    if (comm_cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_comm_mu);
        g_comm_cache[key] = e->comm;
        e->comm_cached = true;
    }
    return PGSGD_OK;
}

static int stress_impl(pgsgd_engine* e, int dims, int local, uint64_t n_pairs, uint64_t seed, double* stress_out) {
    if (!e || !stress_out) return fail(PGSGD_ERR_ARG, "path_stress: NULL argument");
    if (dims != 1 && dims != 2) return fail(PGSGD_ERR_ARG, "dims must be 1 or 2");
    if (dims == 2 ? !e->have_2d : !e->have_1d) return fail(PGSGD_ERR_STATE, "no coordinates on the device");
    CU(cudaSetDevice(e->device));
    if (e->coords_in_slices) { int rc = peer_gather(e, dims); if (rc) return rc; }
    double* d_acc = nullptr;
    unsigned long long* d_used = nullptr;
    CU(cudaMalloc(&d_acc, STRESS_STREAMS * sizeof(double)));
    if (cudaMalloc(&d_used, STRESS_STREAMS * sizeof(unsigned long long)) != cudaSuccess) { cudaFree(d_acc); return fail(PGSGD_ERR_NOMEM, "cudaMalloc failed"); }
    // path-sharded records: every rank evaluates its own paths with its share of the pairs (own generator streams) and the
    // sums are combined over the communicator below — a COLLECTIVE call then, like the getters of the peer modes
    const bool collective = e->shard_global_steps != 0 && e->comm != nullptr;
    if (collective) {
        n_pairs = (uint64_t) ((unsigned __int128) n_pairs * e->S / e->shard_global_steps);
        seed += (uint64_t) e->rank * STRESS_STREAMS;
    }
    const uint64_t per = (n_pairs + STRESS_STREAMS - 1) / STRESS_STREAMS;
    cudaError_t err = launch_stress(dims, local, e->d_path_first, (uint32_t) e->P, e->S, e->d_steps, e->d_xy, e->d_x1d, per, seed, d_acc, d_used, e->stream);
    std::vector<double> acc(STRESS_STREAMS);
    std::vector<unsigned long long> used(STRESS_STREAMS);
    if (err == cudaSuccess) err = cudaMemcpyAsync(acc.data(), d_acc, STRESS_STREAMS * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaMemcpyAsync(used.data(), d_used, STRESS_STREAMS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(d_acc); cudaFree(d_used);
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "path_stress: %s", cudaGetErrorString(err));
    double total = 0;
    unsigned long long n = 0;
    for (int t = 0; t < STRESS_STREAMS; ++t) { total += acc[t]; n += used[t]; }
    double pair[2] = {total, (double) n};
    if (collective) {
        double* d_pair = nullptr;
        CU(cudaMalloc(&d_pair, 2 * sizeof(double)));
        cudaError_t ce = cudaMemcpyAsync(d_pair, pair, sizeof(pair), cudaMemcpyHostToDevice, e->stream);
        ncclResult_t r = ce == cudaSuccess ? ncclAllReduce(d_pair, d_pair, 2, ncclDouble, ncclSum, e->comm, e->stream) : ncclSuccess;
        if (ce == cudaSuccess && r == ncclSuccess) ce = cudaMemcpyAsync(pair, d_pair, sizeof(pair), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        cudaFree(d_pair);
        if (r != ncclSuccess) return fail(PGSGD_ERR_NCCL, "path_stress: %s", ncclGetErrorString(r));
        if (ce != cudaSuccess) return fail(PGSGD_ERR_CUDA, "path_stress: %s", cudaGetErrorString(ce));
    }
    *stress_out = pair[1] > 0 ? pair[0] / pair[1] : 0.0;
    return PGSGD_OK;
}

int pgsgd_engine_path_stress(pgsgd_engine* e, int dims, uint64_t n_pairs, uint64_t seed, double* stress_out) {
    return stress_impl(e, dims, 0, n_pairs, seed, stress_out);
}
int pgsgd_engine_local_stress(pgsgd_engine* e, int dims, uint64_t n_pairs, uint64_t seed, double* stress_out) {
    return stress_impl(e, dims, 1, n_pairs, seed, stress_out);
}

int pgsgd_engine_order_1d_components(pgsgd_engine* e, const uint32_t* node_component, uint64_t* order_out) {
    if (!e || !order_out) return fail(PGSGD_ERR_ARG, "order_1d: NULL argument");
    if (!e->have_1d) return fail(PGSGD_ERR_STATE, "no 1D coordinates on the device");
    CU(cudaSetDevice(e->device));
    if (e->coords_in_slices) { int rc = peer_gather(e, 1); if (rc) return rc; }
    uint64_t* d_order = nullptr;
    uint32_t* d_comp = nullptr;
    CU(cudaMalloc(&d_order, e->N * sizeof(uint64_t)));
    cudaError_t err = cudaSuccess;
    if (node_component) {
        err = cudaMalloc(&d_comp, e->N * sizeof(uint32_t));
        if (err == cudaSuccess) err = cudaMemcpy(d_comp, node_component, e->N * sizeof(uint32_t), cudaMemcpyHostToDevice);
    }
    if (err == cudaSuccess) err = launch_order_1d(e->d_x1d, d_comp, d_order, e->N, e->stream);
    if (err == cudaSuccess) err = cudaMemcpy(order_out, d_order, e->N * sizeof(uint64_t), cudaMemcpyDeviceToHost);
    cudaFree(d_order); cudaFree(d_comp);
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "order_1d: %s", cudaGetErrorString(err));
    return PGSGD_OK;
}

int pgsgd_engine_order_1d(pgsgd_engine* e, uint64_t* order_out) { return pgsgd_engine_order_1d_components(e, nullptr, order_out); }

int pgsgd_engine_sort_goodness(pgsgd_engine* e, const uint64_t* order, uint32_t flags, pgsgd_goodness* out) {
    if (!e || !out) return fail(PGSGD_ERR_ARG, "sort_goodness: NULL argument");
    CU(cudaSetDevice(e->device));
    uint64_t* d_order = nullptr;
    if (order) {
        std::vector<uint8_t> seen(e->N, 0);
        for (uint64_t k = 0; k < e->N; ++k) {
            if (order[k] >= e->N || seen[order[k]]) return fail(PGSGD_ERR_ARG, "sort_goodness: order is not a permutation of the node ranks (entry %llu)", (unsigned long long) k);
            seen[order[k]] = 1;
        }
        CU(cudaMalloc(&d_order, (e->N ? e->N : 1) * sizeof(uint64_t)));
        cudaError_t ce = cudaMemcpy(d_order, order, e->N * sizeof(uint64_t), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) { cudaFree(d_order); return fail(PGSGD_ERR_CUDA, "sort_goodness: %s", cudaGetErrorString(ce)); }
    }
    unsigned long long a[9];
    cudaError_t err = launch_goodness(e->d_steps, e->d_path_first, nullptr, (uint32_t) e->P, e->S, e->N, e->d_node_len, d_order, flags, a, e->stream);
    cudaFree(d_order);
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "sort_goodness: %s", cudaGetErrorString(err));
    memset(out, 0, sizeof(*out));
    out->num_links = a[2];
    out->num_gap_links = a[3];
    out->mean_links_length_node = a[2] ? (double) a[0] / (double) a[2] : 0.0;
    out->mean_links_length_nt = a[2] ? (double) a[1] / (double) a[2] : 0.0;
    out->nodes = e->S;
    out->nucleotides = a[6];
    out->sum_path_node_dist_node = e->S ? (double) a[4] / (double) e->S : 0.0;
    out->sum_path_node_dist_nt = a[6] ? (double) a[5] / (double) a[6] : 0.0;
    out->num_penalties = a[7];
    out->num_penalties_diff_orientation = a[8];
    return PGSGD_OK;
}

int pgsgd_engine_set_multi_mode(pgsgd_engine* e, int mode) {
    if (!e) return fail(PGSGD_ERR_ARG, "set_multi_mode: NULL engine");
    if (mode != PGSGD_MULTI_ALLREDUCE && mode != PGSGD_MULTI_PEER && mode != PGSGD_MULTI_HYBRID && mode != PGSGD_MULTI_AUTO && mode != PGSGD_MULTI_SINGLE) return fail(PGSGD_ERR_ARG, "unknown multi-GPU mode %d", mode);
    if (mode != PGSGD_MULTI_ALLREDUCE && mode != PGSGD_MULTI_AUTO && mode != PGSGD_MULTI_SINGLE && (!e->comm || e->n_ranks < 2)) return fail(PGSGD_ERR_STATE, "peer mode needs an attached communicator (pgsgd_engine_attach_comm) with >= 2 ranks");
    if (e->have_2d || e->have_1d) return fail(PGSGD_ERR_STATE, "select the multi-GPU mode before uploading coordinates");
    e->multi_mode = mode;
    e->mode = mode == PGSGD_MULTI_AUTO ? PGSGD_MULTI_ALLREDUCE : mode;
    return PGSGD_OK;
}

int pgsgd_engine_encode_lay(pgsgd_engine* e, const uint32_t* node_component, uint32_t n_components, uint8_t* buf, uint64_t cap, uint64_t* n_bytes) {
    if (!e || !n_bytes) return fail(PGSGD_ERR_ARG, "encode_lay: NULL argument");
    if (!e->have_2d) return fail(PGSGD_ERR_STATE, "no 2D coordinates on the device");
    if (node_component && n_components == 0) return fail(PGSGD_ERR_ARG, "encode_lay: n_components is 0");
    CU(cudaSetDevice(e->device));
    if (e->coords_in_slices && e->peer_ready_2d) { int rc = peer_gather(e, 2); if (rc) return rc; }
    uint32_t* d_comp = nullptr;
    double *d_xo = nullptr, *d_yo = nullptr;
    auto cleanup = [&]() { cudaFree(d_comp); cudaFree(d_xo); cudaFree(d_yo); };
    if (node_component) {
        for (uint64_t r = 0; r < e->N; ++r)
            if (node_component[r] >= n_components) return fail(PGSGD_ERR_ARG, "encode_lay: node %llu has component %u >= n_components", (unsigned long long) r, node_component[r]);
        cudaError_t ce = cudaMalloc(&d_comp, e->N * sizeof(uint32_t));
        if (ce == cudaSuccess) ce = cudaMemcpy(d_comp, node_component, e->N * sizeof(uint32_t), cudaMemcpyHostToDevice);
        // bounding box per component on the device, the stacking itself (K values, sequential) on the host: layout_main.cpp:406-420
        std::vector<double> stats(3 * (size_t) n_components), xo(n_components), yo(n_components);
        if (ce == cudaSuccess) ce = launch_component_ranges(e->d_xy, d_comp, e->N, n_components, stats.data(), e->stream);
        const double border = 1000.0;
        double curr_y_offset = border;
        for (uint32_t k = 0; k < n_components; ++k) {
            const double min_x = stats[3 * k], min_y = stats[3 * k + 1], max_y = stats[3 * k + 2];
            xo[k] = min_x - border;
            yo[k] = curr_y_offset - min_y;
            curr_y_offset += (max_y - min_y) + border;
        }
        if (ce == cudaSuccess) ce = cudaMalloc(&d_xo, n_components * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc(&d_yo, n_components * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMemcpy(d_xo, xo.data(), n_components * sizeof(double), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(d_yo, yo.data(), n_components * sizeof(double), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) { cleanup(); return fail(PGSGD_ERR_CUDA, "encode_lay: %s", cudaGetErrorString(ce)); }
    }
    LayEncoded enc;
    cudaError_t err = launch_encode_lay(e->d_xy, e->N, d_comp, d_xo, d_yo, &enc, e->stream);
    cleanup();
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "encode_lay: %s", cudaGetErrorString(err));
    const uint64_t zw = (enc.z_bits + 63) / 64, sw = (enc.sp_bits + 63) / 64;
    const uint64_t total = 8 + 8 + (8 + 1 + zw * 8) + (8 + 1 + sw * 8);
    *n_bytes = total;
    int rc = PGSGD_OK;
    if (buf) {
        if (cap < total) rc = fail(PGSGD_ERR_ARG, "encode_lay: buffer of %llu bytes, need %llu", (unsigned long long) cap, (unsigned long long) total);
        else {
            uint8_t* q = buf;
            memcpy(q, &enc.min_value, 8); q += 8;
            memcpy(q, &enc.n_vals, 8); q += 8;
            memcpy(q, &enc.z_bits, 8); q += 8;
            *q++ = 1;
            err = cudaMemcpy(q, enc.d_z, zw * 8, cudaMemcpyDeviceToHost); q += zw * 8;
            memcpy(q, &enc.sp_bits, 8); q += 8;
            *q++ = (uint8_t) enc.width;
            if (err == cudaSuccess) err = cudaMemcpy(q, enc.d_sp, sw * 8, cudaMemcpyDeviceToHost);
            for (int k = 0; k < 2; ++k)   // the closing sample (0, |z| + 1)
                if (enc.tail_bits[k] && enc.tail_word + k < sw) {
                    unsigned long long w;
                    memcpy(&w, q + (enc.tail_word + k) * 8, 8);
                    w |= enc.tail_bits[k];
                    memcpy(q + (enc.tail_word + k) * 8, &w, 8);
                }
            if (err != cudaSuccess) rc = fail(PGSGD_ERR_CUDA, "encode_lay: %s", cudaGetErrorString(err));
        }
    }
    cudaFree(enc.d_z); cudaFree(enc.d_sp);
    return rc;
}

int pgsgd_engine_resolved_multi_mode(const pgsgd_engine* e) { return e ? (e->comm ? e->mode : PGSGD_MULTI_ALLREDUCE) : -1; }

int pgsgd_engine_set_shard(pgsgd_engine* e, uint64_t global_step_count) {
    if (!e) return fail(PGSGD_ERR_ARG, "set_shard: NULL engine");
    if (global_step_count != 0 && global_step_count < e->S) return fail(PGSGD_ERR_ARG, "set_shard: the job cannot have fewer steps (%llu) than this shard (%llu)",
                                                                    (unsigned long long) global_step_count, (unsigned long long) e->S);
    e->shard_global_steps = global_step_count;
    return PGSGD_OK;
}

int pgsgd_engine_sample_terms(pgsgd_engine* e, const pgsgd_config* cfg, int dims, int cooling, double theta_zipf,
                              uint64_t stream, uint64_t n_terms, uint64_t* step_index, uint32_t* path, uint64_t* rank_a,
                              uint64_t* rank_b, uint32_t* node_a, uint32_t* node_b, uint64_t* pos_a, uint64_t* pos_b,
                              uint8_t* end_a, uint8_t* end_b, uint8_t* valid) {
    if (!e) return fail(PGSGD_ERR_ARG, "sample_terms: NULL engine");
    int rc = check_config(cfg);
    if (rc) return rc;
    if (dims != 1 && dims != 2) return fail(PGSGD_ERR_ARG, "dims must be 1 or 2");
    CU(cudaSetDevice(e->device));
    rc = ensure_zetas(e, build_zetas(*cfg));
    if (rc) return rc;
    e->zetas_on_device = false;   // the table on the device is this call's now
    e->ztab_dims = 0;
    SamplerParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.path_first = e->d_path_first;
    sp.zetas = e->d_zetas;
    sp.step_count = e->S;
    sp.path_count = (uint32_t) e->P;
    sp.cooling = cooling ? 1u : 0u;
    sp.space = cfg->space; sp.space_max = cfg->space_max; sp.space_q = cfg->space_quantization_step;
    sp.zipf = make_zipf_const(theta_zipf);
    sp.zipf_f = make_zipf_const_f(sp.zipf);
    const uint64_t n = n_terms ? n_terms : 1;
    uint8_t* d_buf = nullptr;
    const uint64_t per = 8 * 5 + 4 * 3 + 3;  // bytes per term over all outputs
    CU(cudaMalloc(&d_buf, n * per + 64));
    SampleOut o;
    uint8_t* q = d_buf;
    o.step_index = (uint64_t*) q; q += 8 * n;
    o.rank_a = (uint64_t*) q; q += 8 * n;
    o.rank_b = (uint64_t*) q; q += 8 * n;
    o.pos_a = (uint64_t*) q; q += 8 * n;
    o.pos_b = (uint64_t*) q; q += 8 * n;
    o.path = (uint32_t*) q; q += 4 * n;
    o.node_a = (uint32_t*) q; q += 4 * n;
    o.node_b = (uint32_t*) q; q += 4 * n;
    o.end_a = q; q += n;
    o.end_b = q; q += n;
    o.valid = q;
    cudaError_t err = launch_sample_terms(dims, sp, e->d_steps, cfg->seed + stream, n_terms, o, e->stream);
    auto back = [&](void* dst, const void* src, size_t bytes) {
        if (dst && err == cudaSuccess) err = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream);
    };
    back(step_index, o.step_index, 8 * n_terms); back(rank_a, o.rank_a, 8 * n_terms); back(rank_b, o.rank_b, 8 * n_terms);
    back(pos_a, o.pos_a, 8 * n_terms); back(pos_b, o.pos_b, 8 * n_terms); back(path, o.path, 4 * n_terms);
    back(node_a, o.node_a, 4 * n_terms); back(node_b, o.node_b, 4 * n_terms); back(end_a, o.end_a, n_terms);
    back(end_b, o.end_b, n_terms); back(valid, o.valid, n_terms);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    cudaFree(d_buf);
    if (err != cudaSuccess) return fail(PGSGD_ERR_CUDA, "sample_terms: %s", cudaGetErrorString(err));
    return PGSGD_OK;
}

// ---- one-shot entry points -----------------------------------------------------------------------

static int pick_device() {
    const char* s = getenv("PGSGD_DEVICE");
    return s ? atoi(s) : 0;
}

// One process, n_gpus devices: one host thread per GPU runs the multi-rank sequence of INTEGRATION.md §5 (this is how a
// single-process host such as odgi uses a whole box).  dims 2: X,Y [2N]; dims 1: X [N] (Y, frozen as in the 1-GPU calls).
static int run_multi_in_process(const pgsgd_graph_view* g, const pgsgd_config* cfg, int dims, int n_gpus, int multi_mode,
                                const uint8_t* frozen, int x_is_initialised, double* X, double* Y, pgsgd_stats* stats) {
    uint8_t id[128];
    int rc = pgsgd_comm_unique_id(id);
    if (rc) return rc;
    const uint64_t n_out = dims == 2 ? 2 * g->node_count : g->node_count;
    std::vector<int> rcs(n_gpus, PGSGD_OK);
    std::vector<std::string> errs(n_gpus);
    std::vector<pgsgd_stats> sts(n_gpus);
    std::vector<std::vector<double>> outX(n_gpus), outY(n_gpus);
    // A rank that fails must not leave its siblings blocked in a collective: (1) every engine is created before any
    // communicator is attached, and nobody attaches unless all creations succeeded; (2) a rank that fails later aborts every
    // communicator of the call (ncclCommAbort unblocks the others' pending collectives, which then fail and return).
    std::vector<pgsgd_engine*> engines(n_gpus, nullptr);
    std::mutex mu;
    std::condition_variable cv;
    int created = 0;
    bool create_failed = false, aborted = false;
    auto abort_all = [&]() {
        std::lock_guard<std::mutex> lk(mu);
        if (aborted) return;
        aborted = true;
        for (pgsgd_engine* q : engines)
            if (q && q->comm) {
                {
                    std::lock_guard<std::mutex> lc(g_comm_mu);
                    for (auto it = g_comm_cache.begin(); it != g_comm_cache.end();) it = it->second == q->comm ? g_comm_cache.erase(it) : std::next(it);
                }
                ncclCommAbort(q->comm);
                q->comm = nullptr;
            }
    };
    std::vector<std::thread> threads;
    for (int r = 0; r < n_gpus; ++r) {
        threads.emplace_back([&, r]() {
            pgsgd_engine* e = nullptr;
            int c = pgsgd_engine_create(g, r, &e);
            {
                std::unique_lock<std::mutex> lk(mu);
                engines[r] = e;
                if (c) create_failed = true;
                ++created;
                cv.notify_all();
                cv.wait(lk, [&]() { return created == n_gpus; });
                if (!c && create_failed) c = PGSGD_ERR_STATE;   // a sibling could not create its engine: nobody attaches
            }
            if (c == PGSGD_ERR_STATE && errs[r].empty()) errs[r] = "another GPU of the call failed to create its engine";
            if (!c) c = pgsgd_engine_attach_comm(e, id, n_gpus, r);
            if (!c) c = pgsgd_engine_set_multi_mode(e, multi_mode);
            if (!c) c = dims == 2 ? pgsgd_engine_set_coords_2d(e, X, Y) : pgsgd_engine_set_coords_1d(e, x_is_initialised ? X : nullptr);
            if (!c && dims == 1) c = pgsgd_engine_set_frozen_1d(e, frozen);
            memset(&sts[r], 0, sizeof(pgsgd_stats));
            if (!c) c = dims == 2 ? pgsgd_engine_run_2d(e, cfg, &sts[r]) : pgsgd_engine_run_1d(e, cfg, &sts[r]);
            if (!c) {
                outX[r].resize(n_out);
                if (dims == 2) { outY[r].resize(n_out); c = pgsgd_engine_get_coords_2d(e, outX[r].data(), outY[r].data()); }
                else c = pgsgd_engine_get_coords_1d(e, outX[r].data());
            }
            if (c && errs[r].empty()) errs[r] = pgsgd_last_error();
            rcs[r] = c;
            if (c && !create_failed) abort_all();
        });
    }
    for (auto& t : threads) t.join();
    for (pgsgd_engine* q : engines) if (q) pgsgd_engine_destroy(q);
    for (int r = 0; r < n_gpus; ++r) if (rcs[r]) return fail(rcs[r], "GPU %d: %s", r, errs[r].c_str());
    memcpy(X, outX[0].data(), n_out * sizeof(double));
    if (dims == 2) memcpy(Y, outY[0].data(), n_out * sizeof(double));
    if (stats) {
        *stats = sts[0];
        for (int r = 1; r < n_gpus; ++r) {
            stats->term_updates += sts[r].term_updates;
            stats->kernel_launches += sts[r].kernel_launches;
            if (sts[r].seconds_iterations > stats->seconds_iterations) stats->seconds_iterations = sts[r].seconds_iterations;
        }
    }
    return PGSGD_OK;
}

static int env_gpus() {
    const char* s = getenv("PGSGD_GPUS");
    return s ? atoi(s) : 1;
}
static int env_multi_mode() {
    const char* s = getenv("PGSGD_MULTI");
    if (!s) return PGSGD_MULTI_AUTO;
    if (!strcmp(s, "allreduce")) return PGSGD_MULTI_ALLREDUCE;
    if (!strcmp(s, "peer")) return PGSGD_MULTI_PEER;
    if (!strcmp(s, "hybrid")) return PGSGD_MULTI_HYBRID;
    return PGSGD_MULTI_AUTO;
}

int pgsgd_layout_2d_multi(const pgsgd_graph_view* g, const pgsgd_config* cfg, int n_gpus, int multi_mode, double* X, double* Y, pgsgd_stats* stats) {
    if (!g || !X || !Y) return fail(PGSGD_ERR_ARG, "pgsgd_layout_2d_multi: NULL argument");
    if (n_gpus < 2 || n_gpus > pgsgd_device_count()) return fail(PGSGD_ERR_ARG, "n_gpus %d out of range (2..%d)", n_gpus, pgsgd_device_count());
    return run_multi_in_process(g, cfg, 2, n_gpus, multi_mode, nullptr, 1, X, Y, stats);
}

int pgsgd_sort_1d_multi(const pgsgd_graph_view* g, const pgsgd_config* cfg, int n_gpus, int multi_mode, const uint8_t* frozen,
                        int x_is_initialised, double* X, pgsgd_stats* stats) {
    if (!g || !X) return fail(PGSGD_ERR_ARG, "pgsgd_sort_1d_multi: NULL argument");
    if (n_gpus < 2 || n_gpus > pgsgd_device_count()) return fail(PGSGD_ERR_ARG, "n_gpus %d out of range (2..%d)", n_gpus, pgsgd_device_count());
    return run_multi_in_process(g, cfg, 1, n_gpus, multi_mode, frozen, x_is_initialised, X, nullptr, stats);
}

int pgsgd_layout_2d(const pgsgd_graph_view* g, const pgsgd_config* cfg, double* X, double* Y, pgsgd_stats* stats) {
    if (!X || !Y) return fail(PGSGD_ERR_ARG, "pgsgd_layout_2d: X/Y are NULL");
    if (env_gpus() > 1) return pgsgd_layout_2d_multi(g, cfg, env_gpus(), env_multi_mode(), X, Y, stats);   // PGSGD_GPUS=8 odgi layout --gpu
    pgsgd_engine* e = nullptr;
    int rc = pgsgd_engine_create(g, pick_device(), &e);
    if (rc) return rc;
    rc = pgsgd_engine_set_coords_2d(e, X, Y);
    pgsgd_stats st;
    memset(&st, 0, sizeof(st));
    if (!rc) rc = pgsgd_engine_run_2d(e, cfg, &st);
    if (!rc) {
        const double t0 = now_s();
        rc = pgsgd_engine_get_coords_2d(e, X, Y);
        st.seconds_download = now_s() - t0;
        st.d2h_bytes = 4 * g->node_count * sizeof(double);
    }
    pgsgd_engine_destroy(e);
    if (stats) *stats = st;
    return rc;
}

int pgsgd_sort_1d(const pgsgd_graph_view* g, const pgsgd_config* cfg, const uint8_t* frozen, int x_is_initialised, double* X,
                  pgsgd_stats* stats) {
    if (!X) return fail(PGSGD_ERR_ARG, "pgsgd_sort_1d: X is NULL");
    if (env_gpus() > 1) return pgsgd_sort_1d_multi(g, cfg, env_gpus(), env_multi_mode(), frozen, x_is_initialised, X, stats);
    pgsgd_engine* e = nullptr;
    int rc = pgsgd_engine_create(g, pick_device(), &e);
    if (rc) return rc;
    rc = pgsgd_engine_set_coords_1d(e, x_is_initialised ? X : nullptr);
    if (!rc) rc = pgsgd_engine_set_frozen_1d(e, frozen);
    pgsgd_stats st;
    memset(&st, 0, sizeof(st));
    if (!rc) rc = pgsgd_engine_run_1d(e, cfg, &st);
    if (!rc) {
        const double t0 = now_s();
        rc = pgsgd_engine_get_coords_1d(e, X);
        st.seconds_download = now_s() - t0;
        st.d2h_bytes = g->node_count * sizeof(double);
    }
    pgsgd_engine_destroy(e);
    if (stats) *stats = st;
    return rc;
}

}  // extern "C"
