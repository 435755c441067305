// plantos_abi.cu -- host side of libplantos_b200.so: the C ABI declared in include/plantos.h.
// No torch / pybind types cross this boundary; build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/plantos.h"
#include "plantos_fast.cuh"
#ifdef PLANTOS_WITH_LANE_KERNEL      // round 1's experimental lane-per-env kernel: a recorded experiment, not built by default
#include "plantos_lane.cuh"
#endif
#include "plantos_tile.cuh"

using namespace plantos_dev;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(PLANTOS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

typedef void (*fast_kernel_t)(const Params, const StepIO);
typedef void (*tile_kernel_t)(const Params, const RollIO);

struct FastVariant { int R, C, keep; fast_kernel_t fn; };

#define FAST_ROW(R_, C_) {R_, C_, 0, k_step_fast<R_, C_, false>}, {R_, C_, 1, k_step_fast<R_, C_, true>}
const FastVariant kFastVariants[] = {
    FAST_ROW(6, 16),  // training preset, A2C_training.py:206-212
    FAST_ROW(2, 10),  // ctor default, plantos_env.py:25-26
    FAST_ROW(4, 16),  // test_environment.py:24 custom env
    FAST_ROW(4, 8),
};

#ifdef PLANTOS_WITH_LANE_KERNEL
#define LANE_ROW(R_, C_) {R_, C_, 0, k_step_lane<R_, C_, false>}, {R_, C_, 1, k_step_lane<R_, C_, true>}
const FastVariant kLaneVariants[] = { LANE_ROW(6, 16), LANE_ROW(2, 10), LANE_ROW(4, 16), LANE_ROW(4, 8) };
#endif

// k_step_tile (plantos_tile.cuh): lane-per-env simulation + byte-coded observation output
struct TileVariant { int R, C; tile_kernel_t step, rollout; };
#define TILE_ROW(R_, C_) {R_, C_, k_step_tile<R_, C_>, k_rollout_tile<R_, C_>}
const TileVariant kTileVariants[] = { TILE_ROW(6, 16), TILE_ROW(2, 10), TILE_ROW(4, 16), TILE_ROW(4, 8) };

// Does a LIDAR offset table equal the compile-time one the lane kernel was built with?
template <int R, int C>
bool lane_offsets_equal(const int8_t* off) {
    for (int i = 0; i < C; ++i)
        for (int r = 0; r < R; ++r)
            if (off[(i * R + r) * 2] != LidarGen<R, C>::dx(i, r) || off[(i * R + r) * 2 + 1] != LidarGen<R, C>::dy(i, r)) return false;
    return true;
}
bool lane_offsets_match(int R, int C, const int8_t* off) {
    if (R == 6 && C == 16) return lane_offsets_equal<6, 16>(off);
    if (R == 2 && C == 10) return lane_offsets_equal<2, 10>(off);
    if (R == 4 && C == 16) return lane_offsets_equal<4, 16>(off);
    if (R == 4 && C == 8) return lane_offsets_equal<4, 8>(off);
    return false;
}

// launch shape of one persistent specialised kernel
struct FastLaunch { fast_kernel_t fn; int grid, threads, smem, q; };

}  // namespace

struct plantos {
    plantos_config_t cfg;
    int device;
    int num_sms;
    Params p;
    // owned device buffers
    void* d_tables;      // one allocation holding every table
    uint8_t* d_map_cells;
    int16_t* d_map_rover;
    // staging for plantos_step_host
    long long* s_actions;
    float* s_obs;
    float* s_reward;
    uint8_t* s_done;
    // launch configuration
    bool use_fast;               // a specialised kernel exists for this shape
    FastLaunch trip;             // k_step_fast (fn == nullptr: none)
    FastLaunch lane;             // k_step_lane
    FastLaunch tile;             // k_step_tile<R, C> (fn unused: see tile_step)
    tile_kernel_t tile_step, tile_rollout;   // k_step_tile<R, C> / k_rollout_tile<R, C>
    int rollout_grid, rollout_smem;
    int impl;                    // 0: k_step_tile when possible, 1: k_step_fast, 2: k_step_lane (experiments)
    const char* last_kernel;     // name of the kernel the latest plantos_step launched
    bool wrc_valid;              // the window ring cache mirrors the planes (k_step_tile keeps it so)
    void* d_sync;                // tickets (8 B) + per-tile step flags of k_step_tile
    Params* d_params;            // Params::self: the global-memory copy of p (see sync_params)
    bool pipelining;             // plantos_set_pipelining
    bool prev_tile_step;         // the handle's latest enqueued operation was a k_step_tile / k_rollout_tile launch ...
    bool prev_multi;             // ... of the multi-step kernel (publishes its tiles after its last store)
    const float* prev_obs;       // ... that wrote this observation range on this stream
    size_t prev_obs_bytes;
    void* prev_stream;
    bool lane_offsets_ok;        // the uploaded LIDAR offsets equal the lane kernel's compile-time table
    bool prefer_lane;
    bool use_pdl;
    uint4* d_table_blob;
    int4* d_lane_tab;
    int generic_grid, generic_smem;
    bool did_reset;
    int64_t launches;
    int64_t steps;               // plantos_step calls so far (episode log's step_seq)
};

// Params::self mirrors h->p in global memory for the tile kernels' reset path.  Called (after the device is idle)
// by everything that changes a field that path reads: create, push_maps, set_curriculum(+reuse_map), upload_tables.
static int sync_params(plantos_t* h);

// every enqueued kernel other than a step launch ends a pipelined sequence of steps
static void note_launch(plantos_t* h) { h->launches += 1; h->prev_tile_step = false; }

// ------------------------------------------------------------------ host tables
extern "C" int plantos_default_config(plantos_config_t* cfg) {
    if (!cfg) return fail(PLANTOS_EINVAL, "cfg is NULL");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (int32_t)sizeof(*cfg);
    cfg->num_envs = 1;
    cfg->grid_size = 21; cfg->num_plants = 8; cfg->num_obstacles = 50;
    cfg->lidar_range = 2; cfg->lidar_channels = 10;
    cfg->max_steps = 1000;
    cfg->thirsty_plant_prob = 0.7f;
    cfg->map_source = PLANTOS_MAPS_PHILOX;
    cfg->seed = 0;
    cfg->r_goal = 20; cfg->r_mistake = -10; cfg->r_invalid = -5; cfg->r_water_empty = -5;
    cfg->r_step = -0.1; cfg->r_exploration = 10; cfg->r_revisit = -1; cfg->r_complete_exploration = 50;
    cfg->kernel = PLANTOS_KERNEL_AUTO;
    return PLANTOS_OK;
}

static int validate(const plantos_config_t* c) {
    if (!c) return fail(PLANTOS_EINVAL, "cfg is NULL");
    if (c->struct_size != (int32_t)sizeof(plantos_config_t))
        return fail(PLANTOS_EINVAL, "plantos_config_t.struct_size does not match this library (ABI mismatch)");
    if (c->num_envs < 1) return fail(PLANTOS_EINVAL, "num_envs must be >= 1");
    if (c->grid_size < 5 || c->grid_size > 128) return fail(PLANTOS_EINVAL, "grid_size must be in [5, 128]");
    if (c->lidar_range < 1 || c->lidar_range > 64) return fail(PLANTOS_EINVAL, "lidar_range must be in [1, 64]");
    if (c->lidar_channels < 1 || c->lidar_channels > 64) return fail(PLANTOS_EINVAL, "lidar_channels must be in [1, 64]");
    if (c->max_steps < 1 || c->max_steps > 65535) return fail(PLANTOS_EINVAL, "max_steps must be in [1, 65535]");
    if (c->num_plants < 0 || c->num_obstacles < 0) return fail(PLANTOS_EINVAL, "num_plants / num_obstacles must be >= 0");
    if (c->num_plants > 255) return fail(PLANTOS_EINVAL, "num_plants must be <= 255");
    // The reference raises ValueError when fewer than P+1 cells are free (plantos_env.py:360-364).
    // Obstacle clusters never touch the border ring (centres in [2, G-3], reach 1), so 4G-4
    // cells are always free: require that bound so a reset can never fail on the device.
    if (4 * c->grid_size - 4 < c->num_plants + 1)
        return fail(PLANTOS_EINVAL, "Not enough guaranteed-free positions (4G-4) to place num_plants plants and 1 rover");
    if (!(c->thirsty_plant_prob >= 0.0f && c->thirsty_plant_prob <= 1.0f))
        return fail(PLANTOS_EINVAL, "thirsty_plant_prob must be in [0, 1]");
    if (c->map_source != PLANTOS_MAPS_PHILOX && c->map_source != PLANTOS_MAPS_INJECTED && c->map_source != PLANTOS_MAPS_MAZE)
        return fail(PLANTOS_EINVAL, "map_source must be PLANTOS_MAPS_PHILOX, PLANTOS_MAPS_INJECTED or PLANTOS_MAPS_MAZE");
    if (c->kernel < PLANTOS_KERNEL_AUTO || c->kernel > PLANTOS_KERNEL_FAST)
        return fail(PLANTOS_EINVAL, "kernel must be one of PLANTOS_KERNEL_*");
    if (c->tune_fast_grid < 0 || c->tune_fast_impl < 0 || c->tune_fast_impl > 2 || c->tune_l2_keep_mb < 0)
        return fail(PLANTOS_EINVAL, "tune_* fields out of range");
    return PLANTOS_OK;
}

extern "C" int plantos_obs_dim(const plantos_config_t* cfg) {
    if (!cfg) return fail(PLANTOS_EINVAL, "cfg is NULL");
    return cfg->lidar_channels * 5 + 2 + 25;
}

extern "C" int plantos_compute_tables(const plantos_config_t* c, int8_t* lidar_off, float* dist_tab,
                                      float* pos_tab, float* visit_tab, double* reward_tab) {
    int rc = validate(c);
    if (rc) return rc;
    const int C = c->lidar_channels, R = c->lidar_range, G = c->grid_size;
    if (lidar_off) {
        for (int i = 0; i < C; ++i) {
            // angle = (2 * math.pi * i) / C, evaluated left to right in double (plantos_env.py:261)
            const double angle = (2.0 * M_PI * (double)i) / (double)C;
            for (int r = 1; r <= R; ++r) {
                lidar_off[(i * R + (r - 1)) * 2 + 0] = (int8_t)(int)((double)r * std::cos(angle));  // :266
                lidar_off[(i * R + (r - 1)) * 2 + 1] = (int8_t)(int)((double)r * std::sin(angle));  // :267
            }
        }
    }
    if (dist_tab) for (int r = 0; r <= R; ++r) dist_tab[r] = (float)((double)r / (double)R);       // :288
    if (pos_tab) for (int x = 0; x < G; ++x) pos_tab[x] = (float)((double)x / (double)G);            // :295-296
    if (visit_tab) for (int k = 0; k <= 10; ++k) visit_tab[k] = (float)((double)k / 10.0);           // :308
    if (reward_tab) {
        const double x[PLANTOS_RW_COUNT] = {c->r_exploration, c->r_revisit, c->r_invalid,
                                            c->r_goal, c->r_water_empty, c->r_mistake};
        for (int k = 0; k < PLANTOS_RW_COUNT; ++k) {
            double rew = c->r_step;                       // reward = self.R_STEP            (:164)
            rew += x[k];                                  // reward += handler result        (:167,169)
            reward_tab[k] = rew;
            rew += c->r_complete_exploration;             // reward += R_COMPLETE_EXPLORATION (:180)
            reward_tab[PLANTOS_RW_COUNT + k] = rew;
        }
    }
    return PLANTOS_OK;
}

// ------------------------------------------------------------------ create / destroy
static int upload_tables_impl(plantos_t* h, const int8_t* lidar_off, const float* dist_tab, const float* pos_tab,
                              const float* visit_tab, const double* reward_tab) {
    const Params& p = h->p;
    if (lidar_off) {
        CUDA_TRY(cudaMemcpy((void*)p.lidar_off, lidar_off, (size_t)p.C * p.R * 2, cudaMemcpyHostToDevice));
        h->lane_offsets_ok = lane_offsets_match(p.R, p.C, lidar_off);
    }
    if (dist_tab) CUDA_TRY(cudaMemcpy((void*)p.dist_tab, dist_tab, (size_t)(p.R + 1) * 4, cudaMemcpyHostToDevice));
    if (pos_tab) CUDA_TRY(cudaMemcpy((void*)p.pos_tab, pos_tab, (size_t)p.G * 4, cudaMemcpyHostToDevice));
    if (visit_tab) CUDA_TRY(cudaMemcpy((void*)p.visit_tab, visit_tab, 11 * 4, cudaMemcpyHostToDevice));
    if (reward_tab) {
        float r32[2 * PLANTOS_RW_COUNT];
        for (int i = 0; i < 2 * PLANTOS_RW_COUNT; ++i) r32[i] = (float)reward_tab[i];
        CUDA_TRY(cudaMemcpy((void*)p.reward64, reward_tab, sizeof(double) * 2 * PLANTOS_RW_COUNT, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy((void*)p.reward32, r32, sizeof(r32), cudaMemcpyHostToDevice));
    }
    // re-pack the fast kernel's table image and lane constants from the device tables
    k_pack_tables<<<1, 128, tables_bytes(p.G, p.R, p.C)>>>(p, h->d_table_blob, h->d_lane_tab);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return PLANTOS_OK;
}

static void free_all(plantos_t* h) {
    if (!h) return;
    cudaFree(h->p.rec); cudaFree(h->p.term_rec); cudaFree(h->p.types); cudaFree(h->p.vis4); cudaFree(h->p.visov);
    cudaFree(h->d_tables); cudaFree(h->d_table_blob); cudaFree(h->d_lane_tab); cudaFree(h->p.stats); cudaFree(h->p.err);
    cudaFree(h->d_map_cells); cudaFree(h->d_map_rover);
    cudaFree(h->p.ep_log); cudaFree(h->p.ep_log_count);
    cudaFree(h->p.cur_thr); cudaFree(h->p.cur_cnt); cudaFree(h->p.expl); cudaFree(h->p.wrc); cudaFree(h->d_sync); cudaFree(h->d_params);
    cudaFree(h->s_actions); cudaFree(h->s_obs); cudaFree(h->s_reward); cudaFree(h->s_done);
    delete h;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and process-wide: a later handle with a smaller
// need must not lower what an earlier live handle relies on.  Only ever raise it (per device).
static cudaError_t raise_dyn_smem(const void* fn, int device, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cur;
    std::lock_guard<std::mutex> lock(mu);
    int& have = cur[{fn, device}];
    if (bytes <= have) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

extern "C" int plantos_create(const plantos_config_t* cfg, int device, plantos_t** out) {
    if (!out) return fail(PLANTOS_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = validate(cfg);
    if (rc) return rc;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PLANTOS_EINVAL, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(PLANTOS_ECUDA, std::string("libplantos_b200 is built for sm_100a only; device is ") + prop.name);

    plantos_t* h = new plantos();
    std::memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    Params& p = h->p;
    p.N = cfg->num_envs; p.env_base = cfg->env_id_base;
    p.G = cfg->grid_size; p.P = cfg->num_plants; p.O = cfg->num_obstacles;
    p.R = cfg->lidar_range; p.C = cfg->lidar_channels; p.D = plantos_obs_dim(cfg);
    p.W = (p.G + 31) / 32;
    // wall-padded type plane, u64 words per env: R+2 wall rows above the grid and R+2 (or one
    // more, to make the row count even = 16-byte env stride) below: the fast kernel fetches
    // rows x-R-1 .. x+R+1 (+1 for alignment) around any rover row x without bounds checks
    p.TP = p.R + 2;
    p.TS = ((p.TP + p.G + p.R + 2 + 1) & ~1) * p.W;
    p.VW = ((p.G + 4 + 7) / 8 + 3) / 4 * 4;              // u32 words per visit-nibble row (16-byte rows)
    p.VE = (p.G + 2 * kVisRowPad) * p.VW;                // u32 words per env, bordered nibble plane
    p.max_steps = cfg->max_steps; p.nclusters = cfg->num_obstacles / 3;
    {
        double th = std::floor((double)cfg->thirsty_plant_prob * 4294967296.0);
        if (th < 0) th = 0;
        if (th > 4294967296.0) th = 4294967296.0;
        p.thirsty_thresh = (unsigned long long)th;
    }
    p.seed_lo = (uint32_t)(cfg->seed & 0xffffffffull); p.seed_hi = (uint32_t)(cfg->seed >> 32);
    p.map_source = cfg->map_source; p.map_episodes = 0;

    const size_t N = (size_t)p.N;
#define ALLOC(ptr, bytes)                                                       \
    do {                                                                        \
        cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                   \
        if (_e != cudaSuccess) {                                                \
            free_all(h);                                                        \
            return fail(PLANTOS_ECUDA, std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(_e)); \
        }                                                                       \
    } while (0)
    ALLOC(p.rec, N * 32);
    ALLOC(p.term_rec, N * 32);
    ALLOC(p.types, N * p.TS * 8);
    ALLOC(p.vis4, N * p.VE * 4);
    ALLOC(p.visov, N * p.G * p.G * 2);
    ALLOC(p.stats, kStatCount * 8);
    ALLOC(p.err, 4);
    // tables: rw64 | rw32 | dist | pos | visit | off
    const size_t tb = 2 * kRwCount * 8 + 2 * kRwCount * 4 + (p.R + 1) * 4 + p.G * 4 + 12 * 4 + (size_t)p.C * p.R * 2 + 64;
    ALLOC(h->d_tables, tb);
    ALLOC(h->d_table_blob, tables_bytes(p.G, p.R, p.C));
    ALLOC(h->d_lane_tab, 32 * kLaneTabVec * sizeof(int4));
#undef ALLOC
    p.table_blob = h->d_table_blob;
    p.lane_tab = h->d_lane_tab;
    {
        unsigned char* b = (unsigned char*)h->d_tables;
        p.reward64 = (const double*)b; b += 2 * kRwCount * 8;
        p.reward32 = (const float*)b; b += 2 * kRwCount * 4;
        p.dist_tab = (const float*)b; b += (p.R + 1) * 4;
        p.pos_tab = (const float*)b; b += p.G * 4;
        p.visit_tab = (const float*)b; b += 12 * 4;
        p.lidar_off = (const int8_t*)b;
    }
    cudaMemset(p.rec, 0, N * 32);
    cudaMemset(p.term_rec, 0, N * 32);
    cudaMemset(p.types, 0x55, N * p.TS * 8);          // every cell = obstacle: the wall padding
    cudaMemset(p.vis4, 0xFF, N * p.VE * 4);           // every nibble = 15: the window border
    cudaMemset(p.visov, 0, N * p.G * p.G * 2);
    cudaMemset(p.stats, 0, kStatCount * 8);
    cudaMemset(p.err, 0, 4);

    // default tables
    {
        std::vector<int8_t> off((size_t)p.C * p.R * 2);
        std::vector<float> dist(p.R + 1), pos(p.G), visit(11);
        double rw[2 * PLANTOS_RW_COUNT];
        plantos_compute_tables(cfg, off.data(), dist.data(), pos.data(), visit.data(), rw);
        rc = upload_tables_impl(h, off.data(), dist.data(), pos.data(), visit.data(), rw);
        if (rc) { free_all(h); return rc; }
    }

    // kernel selection
    const bool fast_ok = (p.W == 1) && (p.VW == 4) && (p.G + p.R <= 32) && (2 * p.R + 1 <= 16) && (p.C <= 16) &&
                         (tables_bytes(p.G, p.R, p.C) >> 4) <= kFastWarps * 32 &&  // one 16-byte load per thread stages the tables
                         (long long)p.N * (p.TS > p.VE ? p.TS : p.VE) < (1LL << 31) &&  // the fast kernels index the planes with
                         (long long)p.N * p.G * p.G < (1LL << 31);                      // 32-bit element offsets
    h->use_fast = false;
    if (cfg->kernel != PLANTOS_KERNEL_GENERIC && fast_ok) {
        // L2 policy: PLANTOS_L2_KEEP=1 tags the state accesses evict_last inside a persisting-L2
        // set-aside.  Off by default: with the compact state layout the plain LRU already keeps
        // the state resident (steady-state DRAM reads ~16 MB per 131 072-env step), and the
        // set-aside measured neutral to slightly negative (profiles/r1_summary.md).
        int keep = cfg->tune_l2_keep_mb > 0 ? 1 : 0;
        p.l2_keep = keep;
        if (keep) {
            // evict_last lines only persist inside the persisting-L2 set-aside, which is 0 by
            // default.  Measured on B200 (132.6 MB L2, max set-aside 82.9 MB): up to 56 MB helps
            // slightly, 64 MB and more slows the step down by 40 %, so stay at 56 MB.
            size_t want = (size_t)56 << 20;
            if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
            want = (size_t)cfg->tune_l2_keep_mb << 20;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        }
        for (const FastVariant& v : kFastVariants)
            if (v.R == p.R && v.C == p.C && v.keep == keep) { h->use_fast = true; h->trip.fn = v.fn; }
#ifdef PLANTOS_WITH_LANE_KERNEL
        for (const FastVariant& v : kLaneVariants)
            if (v.R == p.R && v.C == p.C && v.keep == keep) h->lane.fn = v.fn;
#endif
        for (const TileVariant& v : kTileVariants)
            if (v.R == p.R && v.C == p.C && h->use_fast) {
                h->tile.fn = h->trip.fn;                    // (non-null marks the tile path as available)
                h->tile_step = v.step; h->tile_rollout = v.rollout;
            }
    }
    if (cfg->kernel == PLANTOS_KERNEL_FAST && !h->use_fast) {
        free_all(h);
        return fail(PLANTOS_EINVAL, "PLANTOS_KERNEL_FAST requested but (G, R, C) has no fast-kernel instantiation");
    }
    h->generic_smem = tables_bytes(p.G, p.R, p.C) + kGenericWarps * generic_warp_scratch_bytes(p.G, p.W, p.D);
    if (h->use_fast) {
        // Persistent grids: each warp walks its own contiguous range of q envs (a multiple of 4; at
        // least 8 envs per warp when there are few envs).  PLANTOS_FAST_IMPL=lane selects the
        // experimental lane-per-env kernel (k_step_lane) instead of k_step_fast.
        auto shape = [&](FastLaunch& L, int warps, int blocks_per_sm, int smem) {
            long long blocks = (long long)h->num_sms * blocks_per_sm;
            if (cfg->tune_fast_grid > 0) blocks = cfg->tune_fast_grid;
            const long long need = ((long long)p.N / 8 + warps - 1) / warps;
            if (blocks > need) blocks = need;
            L.grid = (int)(blocks < 1 ? 1 : blocks);
            L.threads = warps * 32; L.smem = smem;
            const long long nwarps = (long long)L.grid * warps, nfull = p.N & ~3;
            L.q = (int)((((nfull + nwarps - 1) / nwarps) + 3) & ~3LL);
        };
        shape(h->trip, kFastWarps, PLANTOS_FAST_MINBLOCKS,
              tables_bytes(p.G, p.R, p.C) + kFastWarps * fast_warp_scratch_bytes(p.R, p.G, p.D));
#ifdef PLANTOS_WITH_LANE_KERNEL
        shape(h->lane, kLaneWarps, 1, tables_bytes(p.G, p.R, p.C) + kLaneWarps * lane_warp_scratch_bytes(p.R, p.D));
        if ((tables_bytes(p.G, p.R, p.C) >> 4) > kLaneWarps * 32) h->lane.fn = nullptr;
#endif
        {
            // k_step_tile: one block of kTileWarps warps per SM, every warp walks 32-env tiles
            FastLaunch& L = h->tile;
            const long long ntiles = (((long long)p.N & ~3LL) + 31) / 32;
            long long blocks = (ntiles + kTileWarps - 1) / kTileWarps;
            if (blocks > (long long)h->num_sms * PLANTOS_TILE_MINBLOCKS) blocks = (long long)h->num_sms * PLANTOS_TILE_MINBLOCKS;
            if (cfg->tune_fast_grid > 0 && cfg->tune_fast_grid < blocks) blocks = cfg->tune_fast_grid;
            L.grid = (int)(blocks < 1 ? 1 : blocks);
            L.threads = kTileWarps * 32; L.smem = tile_block_smem_bytes(p.R, p.G, p.C); L.q = 32;
            h->rollout_smem = tile_block_smem_bytes_multi(p.R, p.G, p.C);
            if (L.smem > (int)prop.sharedMemPerBlockOptin || h->rollout_smem > (int)prop.sharedMemPerBlockOptin) L.fn = nullptr;
            if (L.fn) {                                     // the window ring cache, one slice per 32-env tile
                const size_t bytes = (size_t)((p.N + 31) / 32) * wrc_tile_bytes(p.R);
                if (cudaMalloc((void**)&p.wrc, bytes) != cudaSuccess) { p.wrc = nullptr; L.fn = nullptr; cudaGetLastError(); }
                else cudaMemset(p.wrc, 0x55, bytes);
                const size_t sb = ((size_t)L.grid + (size_t)((p.N + 31) / 32)) * 4;   // per-block launch counters | per-tile step flags
                if (L.fn && cudaMalloc(&h->d_sync, sb) == cudaSuccess) {
                    cudaMemset(h->d_sync, 0, sb);
                    p.tickets = (unsigned int*)h->d_sync;
                    p.tile_flags = (unsigned int*)h->d_sync + L.grid;
                } else { L.fn = nullptr; cudaGetLastError(); }
            }
        }
        h->prefer_lane = false;
        h->impl = 0;
        h->impl = cfg->tune_fast_impl;                       // 0 k_step_tile, 1 k_step_fast, 2 k_step_lane (if built)
        h->prefer_lane = h->impl == 2;
        h->use_pdl = cfg->tune_no_pdl == 0;
    }
    cudaError_t e1 = raise_dyn_smem((const void*)k_step_generic, h->device, h->generic_smem);
    cudaError_t e2 = raise_dyn_smem((const void*)k_reset_all, h->device, h->generic_smem);
    if (e2 == cudaSuccess) e2 = raise_dyn_smem((const void*)k_reset_done, h->device, h->generic_smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        free_all(h);
        return fail(PLANTOS_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_step_generic, kGenericWarps * 32, h->generic_smem);
    if (occ < 1) occ = 1;
    const long long want = ((long long)p.N + kGenericWarps - 1) / kGenericWarps;
    const long long cap = (long long)h->num_sms * occ;
    h->generic_grid = (int)(want < cap ? want : cap);
    for (FastLaunch* L : {&h->trip, &h->lane}) {
        if (!h->use_fast || !L->fn) continue;
        cudaError_t e3 = raise_dyn_smem((const void*)L->fn, h->device, L->smem);
        if (e3 != cudaSuccess) { free_all(h); return fail(PLANTOS_ECUDA, std::string("cudaFuncSetAttribute(fast): ") + cudaGetErrorString(e3)); }
    }
    if (h->use_fast && h->tile.fn) {
        cudaError_t e3 = raise_dyn_smem((const void*)h->tile_step, h->device, h->tile.smem);
        cudaError_t e4 = raise_dyn_smem((const void*)h->tile_rollout, h->device, h->rollout_smem);
        if (e3 != cudaSuccess || e4 != cudaSuccess) {
            free_all(h);
            return fail(PLANTOS_ECUDA, std::string("cudaFuncSetAttribute(tile): ") + cudaGetErrorString(e3 != cudaSuccess ? e3 : e4));
        }
    }
    cudaError_t es = cudaDeviceSynchronize();
    if (es != cudaSuccess) { free_all(h); return fail(PLANTOS_ECUDA, std::string("create: ") + cudaGetErrorString(es)); }
    if (cudaMalloc((void**)&h->d_params, sizeof(Params)) != cudaSuccess || sync_params(h) != PLANTOS_OK) {
        free_all(h);
        return fail(PLANTOS_ECUDA, "create: parameter mirror");
    }
    *out = h;
    return PLANTOS_OK;
}

static int sync_params(plantos_t* h) {
    if (!h->d_params) return PLANTOS_OK;
    h->p.self = h->d_params;
    CUDA_TRY(cudaMemcpy(h->d_params, &h->p, sizeof(Params), cudaMemcpyHostToDevice));
    return PLANTOS_OK;
}

extern "C" int plantos_destroy(plantos_t* h) {
    if (!h) return PLANTOS_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_all(h);
    return PLANTOS_OK;
}

extern "C" int plantos_upload_tables(plantos_t* h, const int8_t* lidar_off, const float* dist_tab,
                                     const float* pos_tab, const float* visit_tab, const double* reward_tab) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    return upload_tables_impl(h, lidar_off, dist_tab, pos_tab, visit_tab, reward_tab);
}

extern "C" int plantos_push_maps(plantos_t* h, const uint8_t* cells, const int16_t* rover, int episodes) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    if (h->cfg.map_source != PLANTOS_MAPS_INJECTED) return fail(PLANTOS_ESTATE, "handle was not created with PLANTOS_MAPS_INJECTED");
    if (!cells || !rover || episodes < 1) return fail(PLANTOS_EINVAL, "cells/rover NULL or episodes < 1");
    Params& p = h->p;
    const size_t gg = (size_t)p.G * p.G, n = (size_t)p.N * episodes;
    for (size_t i = 0; i < n; ++i) {
        const int x = rover[2 * i], y = rover[2 * i + 1];
        if (x < 0 || x >= p.G || y < 0 || y >= p.G) return fail(PLANTOS_EINVAL, "rover start outside the grid");
        if (cells[i * gg + (size_t)x * p.G + y] == 1) return fail(PLANTOS_EINVAL, "rover start on an obstacle");
    }
    for (size_t i = 0; i < n; ++i) {
        int plants = 0;
        for (size_t c = 0; c < gg; ++c) {
            const uint8_t v = cells[i * gg + c];
            if (v > 3) return fail(PLANTOS_EINVAL, "cell code > 3");
            plants += v >= 2;
        }
        if (plants > 255) return fail(PLANTOS_EINVAL, "more than 255 plant cells in one map (the env record counts thirsty plants in 8 bits)");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(h->d_map_cells); cudaFree(h->d_map_rover);
    h->d_map_cells = nullptr; h->d_map_rover = nullptr;
    CUDA_TRY(cudaMalloc((void**)&h->d_map_cells, n * gg));
    CUDA_TRY(cudaMalloc((void**)&h->d_map_rover, n * 2 * sizeof(int16_t)));
    CUDA_TRY(cudaMemcpy(h->d_map_cells, cells, n * gg, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_map_rover, rover, n * 2 * sizeof(int16_t), cudaMemcpyHostToDevice));
    p.map_cells = h->d_map_cells; p.map_rover = h->d_map_rover; p.map_episodes = episodes;
    { int rc = sync_params(h); if (rc) return rc; }
    // restart every env's map cursor (episode index lives in rec.w of the first uint4)
    CUDA_TRY(cudaMemset(p.rec, 0, (size_t)p.N * 32));
    h->did_reset = false;
    return PLANTOS_OK;
}

// ------------------------------------------------------------------ reset / step
extern "C" int plantos_reset(plantos_t* h, float* obs_dev, void* stream) {
    if (!h || !obs_dev) return fail(PLANTOS_EINVAL, "handle/obs is NULL");
    if (h->cfg.map_source == PLANTOS_MAPS_INJECTED && h->p.map_episodes == 0)
        return fail(PLANTOS_ESTATE, "injected-map mode: call plantos_push_maps before plantos_reset");
    CUDA_TRY(cudaSetDevice(h->device));
    k_reset_all<<<h->generic_grid, kGenericWarps * 32, h->generic_smem, (cudaStream_t)stream>>>(h->p, obs_dev);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    h->did_reset = true;
    h->wrc_valid = false;
    return PLANTOS_OK;
}

// K consecutive steps (K = 1: plantos_step).  The tile kernels take them in one launch (K > 1: the
// state-resident k_rollout_tile<R, C>); every other configuration steps K times.
static int launch_steps(plantos_t* h, int K, const int64_t* actions, float* obs, size_t obs_stride, float* reward,
                        uint8_t* done, uint8_t* terminated, uint8_t* truncated, float* terminal_obs, void* stream) {
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t N = (size_t)h->p.N;
    const bool aligned = (((uintptr_t)obs) & 15u) == 0 && (K == 1 || (obs_stride & 3u) == 0);
    const bool use_tile = h->use_fast && aligned && h->impl == 0 && h->tile.fn && h->lane_offsets_ok;
    if ((!use_tile || h->p.map_source == PLANTOS_MAPS_MAZE) && K > 1) {   // no multi-step kernel for this configuration
        for (int k = 0; k < K; ++k) {
            int rc = launch_steps(h, 1, actions + k * N, obs + k * obs_stride, 0, reward + k * N, done + k * N,
                                  terminated ? terminated + k * N : nullptr, truncated ? truncated + k * N : nullptr,
                                  terminal_obs, stream);
            if (rc) return rc;
        }
        return PLANTOS_OK;
    }
    h->p.step_seq = (unsigned)h->steps;
    h->steps += K;
    if (h->use_fast && aligned) {
        // launched with programmatic stream serialization so that back-to-back steps overlap the
        // next step's prologue with this step's tail (the kernel waits on griddepcontrol before it
        // touches any state); PLANTOS_PDL=0 falls back to a plain launch
        cudaLaunchConfig_t lc = {};
        if (use_tile && !h->wrc_valid) {                    // something else changed the state: rebuild the cache
            const size_t threads = N * 32;
            k_wrc_build<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->p);
            CUDA_TRY(cudaGetLastError());
            h->launches += 1;
        }
        h->wrc_valid = use_tile;
        // Pipelined launch (plantos_set_pipelining): allowed when the handle's previous operation was a tile-kernel
        // launch on this stream and there is no ragged tail (its envs bypass the tile flags).  A single-step launch
        // publishes a tile BEFORE its expansion stores, so whatever follows it must write a DIFFERENT observation
        // range; a multi-step launch publishes a tile after its last store, so its successor may reuse the buffers.
        const size_t obs_bytes = ((size_t)(K - 1) * obs_stride + N * h->p.D) * sizeof(float);
        const bool disjoint = h->prev_obs && ((const char*)obs + obs_bytes <= (const char*)h->prev_obs ||
                                              (const char*)h->prev_obs + h->prev_obs_bytes <= (const char*)obs);
        h->p.pipelined = (use_tile && h->pipelining && h->use_pdl && h->prev_tile_step && (h->prev_multi || disjoint) &&
                          h->prev_stream == stream && (h->p.N & 3) == 0) ? 1 : 0;
        h->p.release = h->pipelining ? 1 : 0;
        h->prev_tile_step = use_tile; h->prev_multi = K > 1; h->prev_obs = obs; h->prev_obs_bytes = obs_bytes; h->prev_stream = stream;
        // (the experimental lane kernel has no curriculum path)
        const bool use_lane = h->prefer_lane && h->lane.fn && h->lane_offsets_ok && !h->p.cur_mode;
        const FastLaunch& L = use_tile ? h->tile : (use_lane ? h->lane : h->trip);
        h->last_kernel = use_tile ? (K > 1 ? "k_rollout_tile" : "k_step_tile") : (use_lane ? "k_step_lane" : "k_step_fast");
        h->p.fast_q = L.q;
        lc.gridDim = dim3((unsigned)L.grid); lc.blockDim = dim3((unsigned)L.threads);
        lc.dynamicSmemBytes = (size_t)((use_tile && K > 1) ? h->rollout_smem : L.smem); lc.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        // (not while the stream is being captured into a CUDA graph: measured, graph kernel nodes
        // with programmatic edges run 0.8 us per step slower than plain graph edges, while eager
        // launches gain 1.1 us from PDL -- unless the launch is pipelined, which needs the edge)
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing((cudaStream_t)stream, &capturing);
        lc.attrs = at; lc.numAttrs = (h->use_pdl && (capturing == cudaStreamCaptureStatusNone || h->p.pipelined)) ? 1 : 0;
        if (use_tile) {
            RollIO ro;
            ro.actions = (const long long*)actions; ro.obs = obs; ro.obs_stride = obs_stride; ro.reward = reward;
            ro.done = done; ro.terminated = terminated; ro.truncated = truncated; ro.terminal_obs = terminal_obs; ro.K = K;
            CUDA_TRY(cudaLaunchKernelEx(&lc, K > 1 ? h->tile_rollout : h->tile_step, h->p, ro));
        } else {
            StepIO io;
            io.actions = (const long long*)actions; io.obs = obs; io.reward = reward; io.done = done;
            io.terminated = terminated; io.truncated = truncated; io.terminal_obs = terminal_obs;
            CUDA_TRY(cudaLaunchKernelEx(&lc, L.fn, h->p, io));
            h->wrc_valid = false;
        }
        if (h->p.map_source == PLANTOS_MAPS_MAZE) {
            // maze handles: the specialised kernels leave the maze generator out; the envs they finished get
            // their new episodes from this follow-up launch (scans the done flags)
            CUDA_TRY(cudaGetLastError());
            StepIO io;
            io.actions = (const long long*)actions; io.obs = obs; io.reward = reward; io.done = done;
            io.terminated = terminated; io.truncated = truncated; io.terminal_obs = terminal_obs;
            k_reset_done<<<h->generic_grid, kGenericWarps * 32, h->generic_smem, (cudaStream_t)stream>>>(h->p, io);
            h->launches += 1;
            h->prev_tile_step = false;
        }
    } else {
        if (h->cfg.kernel == PLANTOS_KERNEL_FAST)
            return fail(PLANTOS_EINVAL, "PLANTOS_KERNEL_FAST needs a 16-byte aligned obs buffer");
        StepIO io;
        io.actions = (const long long*)actions; io.obs = obs; io.reward = reward; io.done = done;
        io.terminated = terminated; io.truncated = truncated; io.terminal_obs = terminal_obs;
        k_step_generic<<<h->generic_grid, kGenericWarps * 32, h->generic_smem, (cudaStream_t)stream>>>(h->p, io);
        h->last_kernel = "k_step_generic";
        h->wrc_valid = false;
        h->prev_tile_step = false;
    }
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return PLANTOS_OK;
}

extern "C" int plantos_step(plantos_t* h, const int64_t* actions, float* obs, float* reward, uint8_t* done,
                            uint8_t* terminated, uint8_t* truncated, float* terminal_obs, void* stream) {
    if (!h || !actions || !obs || !reward || !done) return fail(PLANTOS_EINVAL, "handle/actions/obs/reward/done is NULL");
    if (!h->did_reset) return fail(PLANTOS_ESTATE, "plantos_step before plantos_reset");
    return launch_steps(h, 1, actions, obs, 0, reward, done, terminated, truncated, terminal_obs, stream);
}

extern "C" int plantos_rollout(plantos_t* h, int num_steps, const int64_t* actions, float* obs, int64_t obs_step_stride,
                               float* reward, uint8_t* done, uint8_t* terminated, uint8_t* truncated,
                               float* terminal_obs, void* stream) {
    if (!h || !actions || !obs || !reward || !done) return fail(PLANTOS_EINVAL, "handle/actions/obs/reward/done is NULL");
    if (num_steps < 1) return fail(PLANTOS_EINVAL, "num_steps must be >= 1");
    if (obs_step_stride < (int64_t)h->p.N * h->p.D) return fail(PLANTOS_EINVAL, "obs_step_stride must be >= num_envs * obs_dim");
    if (!h->did_reset) return fail(PLANTOS_ESTATE, "plantos_rollout before plantos_reset");
    return launch_steps(h, num_steps, actions, obs, (size_t)obs_step_stride, reward, done, terminated, truncated, terminal_obs, stream);
}

extern "C" int plantos_step_host(plantos_t* h, const int64_t* actions_host, float* obs_host, float* reward_host,
                                 uint8_t* done_host, void* stream) {
    if (!h || !actions_host || !obs_host || !reward_host || !done_host)
        return fail(PLANTOS_EINVAL, "handle or a host buffer is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t N = (size_t)h->p.N, D = (size_t)h->p.D;
    if (!h->s_actions) {
        CUDA_TRY(cudaMalloc((void**)&h->s_actions, N * 8));
        CUDA_TRY(cudaMalloc((void**)&h->s_obs, N * D * 4));
        CUDA_TRY(cudaMalloc((void**)&h->s_reward, N * 4));
        CUDA_TRY(cudaMalloc((void**)&h->s_done, N));
    }
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(h->s_actions, actions_host, N * 8, cudaMemcpyHostToDevice, st));
    int rc = plantos_step(h, (const int64_t*)h->s_actions, h->s_obs, h->s_reward, h->s_done, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(obs_host, h->s_obs, N * D * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(reward_host, h->s_reward, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(done_host, h->s_done, N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PLANTOS_OK;
}

// ------------------------------------------------------------------ state access
extern "C" int plantos_get_scalars(plantos_t* h, int which, int32_t* out_dev, void* stream) {
    if (!h || !out_dev) return fail(PLANTOS_EINVAL, "handle/out is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    k_get_scalars<<<(h->p.N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->p, which, out_dev);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    return PLANTOS_OK;
}

extern "C" int plantos_get_returns(plantos_t* h, int which, double* out_dev, void* stream) {
    if (!h || !out_dev) return fail(PLANTOS_EINVAL, "handle/out is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    k_get_returns<<<(h->p.N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->p, which, out_dev);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    return PLANTOS_OK;
}

extern "C" int plantos_get_state(plantos_t* h, uint8_t* cells_dev, int32_t* visits_dev, void* stream) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    if (!cells_dev && !visits_dev) return PLANTOS_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t total = (size_t)h->p.N * h->p.G * h->p.G;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)h->num_sms * 32) blocks = (size_t)h->num_sms * 32;
    k_get_state<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(h->p, cells_dev, visits_dev);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    return PLANTOS_OK;
}

extern "C" int plantos_set_state(plantos_t* h, const uint8_t* cells_dev, const int32_t* visits_dev,
                                 const int32_t* scalars_dev, void* stream) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    if (h->p.cur_mode) return fail(PLANTOS_ESTATE, "plantos_set_state is not supported while a curriculum is active");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t threads = (size_t)h->p.N * 32;
    k_set_state<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->p, cells_dev, visits_dev, scalars_dev);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    h->did_reset = true;
    h->wrc_valid = false;
    return PLANTOS_OK;
}

extern "C" int plantos_stats(plantos_t* h, double* out_dev, int clear, void* stream) {
    if (!h || !out_dev) return fail(PLANTOS_EINVAL, "handle/out is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    k_stats_out<<<1, 32, 0, (cudaStream_t)stream>>>(h->p.stats, out_dev, clear);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    return PLANTOS_OK;
}

extern "C" int plantos_check(plantos_t* h, void* stream) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, h->p.err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (err == PLANTOS_ENOMAPS) return fail(PLANTOS_ENOMAPS, "an env was reset more often than maps were pushed for it");
    if (err == PLANTOS_EINVAL) return fail(err, "plantos_set_state: more than 255 thirsty plants in one env (8-bit counter)");
    if (err != 0) return fail(err, "device-side error flag set");
    return PLANTOS_OK;
}

namespace {
__global__ void k_fill_f64(double* dst, double v, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = v;
}
}  // namespace

extern "C" int plantos_set_curriculum(plantos_t* h, int mode, double initial_threshold, double max_threshold,
                                      double threshold_increment, int max_episodes_per_maze) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    if (mode < PLANTOS_CURRICULUM_OFF || mode > PLANTOS_CURRICULUM_MARK) return fail(PLANTOS_EINVAL, "unknown curriculum mode");
    if (mode != PLANTOS_CURRICULUM_OFF && (max_episodes_per_maze < 1 || !(initial_threshold == initial_threshold) ||
                                           !(max_threshold == max_threshold) || !(threshold_increment == threshold_increment)))
        return fail(PLANTOS_EINVAL, "curriculum: max_episodes_per_maze >= 1 and finite thresholds required");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    Params& p = h->p;
    cudaFree(p.cur_thr); cudaFree(p.cur_cnt); cudaFree(p.expl);
    p.cur_thr = nullptr; p.cur_cnt = nullptr; p.expl = nullptr; p.cur_mode = PLANTOS_CURRICULUM_OFF;
    if (mode == PLANTOS_CURRICULUM_OFF) return sync_params(h);
    const size_t N = (size_t)p.N;
    CUDA_TRY(cudaMalloc((void**)&p.cur_thr, N * 8));
    CUDA_TRY(cudaMalloc((void**)&p.cur_cnt, N * 8));
    CUDA_TRY(cudaMalloc((void**)&p.expl, N * p.G * p.W * 4));
    CUDA_TRY(cudaMemset(p.cur_cnt, 0, N * 8));
    CUDA_TRY(cudaMemset(p.expl, 0, N * p.G * p.W * 4));
    k_fill_f64<<<256, 256>>>(p.cur_thr, initial_threshold, p.N);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    p.cur_mode = mode; p.cur_max_eps = max_episodes_per_maze; p.cur_reuse_map = 0;
    p.cur_max_thr = max_threshold; p.cur_inc = threshold_increment;
    h->did_reset = false;
    return sync_params(h);
}

extern "C" int plantos_set_curriculum_reuse_map(plantos_t* h, int enable) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    if (!h->p.cur_mode) return fail(PLANTOS_ESTATE, "no curriculum is active (plantos_set_curriculum)");
    h->p.cur_reuse_map = enable ? 1 : 0;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    return sync_params(h);
}

extern "C" int plantos_set_max_steps(plantos_t* h, int max_steps) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    if (max_steps < 1 || max_steps > 65535) return fail(PLANTOS_EINVAL, "max_steps must be in [1, 65535]");
    h->cfg.max_steps = max_steps;
    h->p.max_steps = max_steps;          // read by the next enqueued step (kernel parameter)
    return PLANTOS_OK;
}

extern "C" int plantos_get_curriculum_thresholds(plantos_t* h, double* out_dev, void* stream) {
    if (!h || !out_dev) return fail(PLANTOS_EINVAL, "handle/out is NULL");
    if (!h->p.cur_mode) return fail(PLANTOS_ESTATE, "no curriculum is active (plantos_set_curriculum)");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(out_dev, h->p.cur_thr, (size_t)h->p.N * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PLANTOS_OK;
}

extern "C" int plantos_rollout_policy(plantos_t* h, const float* uniforms_dev, int64_t* actions_dev, void* stream) {
    if (!h || !uniforms_dev || !actions_dev) return fail(PLANTOS_EINVAL, "handle/uniforms/actions is NULL");
    if (!h->did_reset) return fail(PLANTOS_ESTATE, "plantos_rollout_policy before plantos_reset");
    CUDA_TRY(cudaSetDevice(h->device));
    k_policy_heuristic<<<(h->p.N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->p, uniforms_dev, (long long*)actions_dev);
    CUDA_TRY(cudaGetLastError());
    note_launch(h);
    return PLANTOS_OK;
}

static_assert(sizeof(plantos_episode_t) == 32, "episode log entries are two uint4");

extern "C" int plantos_episode_log_enable(plantos_t* h, int capacity) {
    if (!h || capacity < 0) return fail(PLANTOS_EINVAL, "handle is NULL / negative capacity");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(h->p.ep_log); cudaFree(h->p.ep_log_count);
    h->p.ep_log = nullptr; h->p.ep_log_count = nullptr; h->p.ep_log_cap = 0;
    if (capacity == 0) return PLANTOS_OK;
    CUDA_TRY(cudaMalloc((void**)&h->p.ep_log, (size_t)capacity * 32));
    CUDA_TRY(cudaMalloc((void**)&h->p.ep_log_count, 4));
    CUDA_TRY(cudaMemset(h->p.ep_log_count, 0, 4));
    h->p.ep_log_cap = capacity;
    return PLANTOS_OK;
}

extern "C" int plantos_episode_log_drain(plantos_t* h, plantos_episode_t* out, int max_entries, int* n_out,
                                         int64_t* dropped_out, void* stream) {
    if (!h || !n_out || (max_entries > 0 && !out)) return fail(PLANTOS_EINVAL, "handle/out/n_out is NULL");
    if (!h->p.ep_log) return fail(PLANTOS_ESTATE, "episode log is not enabled (plantos_episode_log_enable)");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned count = 0;
    CUDA_TRY(cudaMemcpyAsync(&count, h->p.ep_log_count, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const unsigned stored = count < (unsigned)h->p.ep_log_cap ? count : (unsigned)h->p.ep_log_cap;
    const unsigned n = stored < (unsigned)max_entries ? stored : (unsigned)max_entries;
    if (n) CUDA_TRY(cudaMemcpyAsync(out, h->p.ep_log, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemsetAsync(h->p.ep_log_count, 0, 4, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_out = (int)n;
    if (dropped_out) *dropped_out = (int64_t)count - (int64_t)n;
    return PLANTOS_OK;
}

extern "C" int64_t plantos_launch_count(const plantos_t* h) { return h ? h->launches : 0; }

extern "C" const char* plantos_kernel_name(const plantos_t* h) {
    if (!h) return "";
    if (!h->use_fast) return "generic";
    return "fast";
}

extern "C" int plantos_set_pipelining(plantos_t* h, int enable) {
    if (!h) return fail(PLANTOS_EINVAL, "handle is NULL");
    h->pipelining = enable != 0;
    h->prev_tile_step = false;
    return PLANTOS_OK;
}

extern "C" const char* plantos_last_step_kernel(const plantos_t* h) {
    return (h && h->last_kernel) ? h->last_kernel : "";
}

extern "C" int64_t plantos_state_bytes_per_env(const plantos_t* h) {
    if (!h) return 0;
    const Params& p = h->p;
    return 32 + 32 + (int64_t)p.TS * 8 + (int64_t)p.VE * 4 + (int64_t)p.G * p.G * 2;
}

extern "C" const char* plantos_last_error(void) { return g_last_error.c_str(); }
extern "C" int plantos_abi_version(void) { return PLANTOS_ABI_VERSION; }
