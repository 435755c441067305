// plantos_common.cuh -- device-side state layout, tables and helpers shared by the
// generic (one warp per env) and fast (half-warp per env) PlantOS kernels.
//
// Persistent state in HBM, per env (SoA across envs, all little-endian):
//   rec     32 B   hot scalars, two uint4 (see EnvRec)
//   types   TS*8   2-bit cell codes (0 empty 1 obstacle 2 hydrated 3 thirsty), row-major,
//                  W = ceil(G/32) u64 words per row, cell y of row x at bits [2*(y&31), +2)
//                  of word (x+TP)*W + (y>>5).  The plane is WALL-PADDED so that neither the
//                  LIDAR nor a move needs a bounds check: TP = R+2 all-obstacle rows above the
//                  grid, at least R+2 below, and the columns >= G of the last word of every row
//                  hold the obstacle code (01).
//   vis4    VE*4   visit_counts (plantos_env.py:146) as saturating 4-BIT counters, 8 per
//                  u32 word, rows of VW words (a multiple of 4, i.e. 16-byte rows), with a
//                  border of 3 rows above/below and 2 columns left/right: cell (x,y) is nibble
//                  (y+2) of row (x+3).  Border nibbles hold 15.  The observation only needs
//                  min(v,10)/10 (plantos_env.py:308) and 15 maps to the 1.0 the reference
//                  writes for out-of-bounds window cells (:310-311), so the 5x5 window is FIVE
//                  CONSECUTIVE 16-byte rows (when G <= 28) read without bounds checks, and the
//                  7 rows around the pre-move position are one contiguous 112-byte fetch.
//   visov   G*G*2  exact u16 count of the cells whose nibble has saturated (v >= 15); never
//                  read or written for the others, so it stays out of cache and DRAM traffic.
//                  visit_counts[x,y] = nibble < 15 ? nibble : visov[x*G+y]  (exact up to 65535).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace plantos_dev {

constexpr int kEmpty = 0, kObstacle = 1, kHydrated = 2, kThirsty = 3;
constexpr uint64_t kObstAll = 0x5555555555555555ull;  // obstacle code in every cell of a word
constexpr int kFlagCollided = 1, kFlagBonus = 2;
constexpr int kRwCount = 6;  // PLANTOS_RW_COUNT
constexpr int kStatCount = 8;
constexpr int kScCount = 11;  // PLANTOS_SC_COUNT

constexpr int kLaneTabVec = 5;   // int4 per lane: srcl[8] | shf[8] | vsrc0, vsh0, vsrc1, vsh1

struct Params {
    int N;
    long long env_base;
    int G, P, O, R, C, D;
    int W;          // u64 words per type row
    int TP;         // wall rows above the grid in the type plane (R + 2)
    int TS;         // u64 words per env in the type plane: (TP + G + R + 2, even) * W
    int VW;         // u32 words per nibble row (multiple of 4)
    int VE;         // u32 words per env in the nibble plane: (G + 2 * kVisRowPad) * VW
    int max_steps;
    int nclusters;  // O / 3 (plantos_env.py:341)
    unsigned long long thirsty_thresh;  // floor(prob * 2^32); draw < thresh => thirsty
    uint32_t seed_lo, seed_hi;
    int l2_keep;       // 1: tag state accesses L2::evict_last (fast kernel)
    int map_source;    // 0 philox, 1 injected
    int map_episodes;  // injected maps per env
    // persistent state
    uint4* rec;
    uint64_t* types;
    uint32_t* vis4;
    uint16_t* visov;
    uint4* term_rec;   // snapshot of rec at each env's latest terminal step
    // tables (global memory; staged into shared memory per block)
    const int8_t* lidar_off;  // [C][R][2]
    const float* dist_tab;    // [R+1]
    const float* pos_tab;     // [G]
    const float* visit_tab;   // [11]
    const float* reward32;    // [2*kRwCount]
    const double* reward64;   // [2*kRwCount]
    // the same tables pre-packed for the fast kernel by k_pack_tables: the shared-memory image of
    // load_tables (tables_bytes, copied with 16-byte loads) and its per-lane constants
    const uint4* table_blob;
    const int4* lane_tab;     // [32][kLaneTabVec]
    int fast_q;               // envs per warp of the fast kernel's persistent grid (multiple of 4)
    // injected maps
    const uint8_t* map_cells;   // [N][E][G*G]
    const int16_t* map_rover;   // [N][E][2]
    // episode statistics (fixed point, see plantos_stats) and sticky error word
    unsigned long long* stats;
    int* err;
    // optional episode log (SB3 Monitor's r, l per finished episode; plantos_episode_log_*): entries
    // of two uint4 {env, length, step seq, terminated | truncated << 1} {f64 return, collisions, watered}
    uint4* ep_log;              // nullptr: disabled
    unsigned int* ep_log_count; // entries appended since the last drain (may exceed ep_log_cap: dropped)
    int ep_log_cap;
    unsigned int step_seq;      // number of the step launch, set by the host
    // optional CurriculumWrapper state (plantos_set_curriculum; every step kernel implements it)
    int cur_mode;               // 0 off, 1 reaching the threshold terminates (A2C_training.py), 2 marks only (trainingCode.py)
    int cur_max_eps;            // max_episodes_per_maze
    int cur_reuse_map;          // 1: a kept maze is regenerated identically (plantos_set_curriculum_reuse_map)
    double cur_max_thr, cur_inc;
    double* cur_thr;            // [N] exploration_threshold
    int2* cur_cnt;              // [N] {episodes_on_current_maze, bit0 maze_completed | bit1 persistent_visit_counts is not None | bits 8.. the episode the maze started in}
    uint32_t* expl;             // [N][G][W] explored_map > 0, one bit per cell (restarts every episode)
    // window ring cache of k_step_tile (see wrc_* below); nullptr when the shape has no tile kernel
    unsigned char* wrc;
    // cross-launch ordering of k_step_tile (plantos_tile.cuh): tickets[b] counts the launches of block b
    // (= the launch's ordinal, the same for every block), tile_flags[t] the steps tile t has completed.
    // pipelined = 1: the launch does not wait for the previous grid (no griddepcontrol.wait); a warp
    // starts tile t when tile_flags[t] == ordinal.
    unsigned int* tickets;
    unsigned int* tile_flags;
    // a copy of this struct in global memory (kept in sync by the host): the rarely taken reset path of the tile
    // kernels is a real function call that reads its parameters through this pointer, so that its register needs
    // stay out of the hot loop's allocation
    const struct Params* self;
    int pipelined;
    int release;                // the handle may overlap launches (plantos_set_pipelining): publish tile flags with release semantics
};

// ------------------------------------------------------------ window ring cache (WRC)
// k_step_tile reads, per env, only the rows of the two planes that one step can touch: the padded type
// rows x+1 .. x+2R+3 (grid rows x-R-1 .. x+R+1) and the padded nibble rows x .. x+6 (grid rows
// x-3 .. x+3) around the rover row x.  The WRC keeps exactly those rows in a second, tile-major array
// whose address does not depend on x, so that a warp fetches its 32 envs' windows with ONE bulk copy
// issued at kernel start (no record -> window dependency, no per-env address arithmetic):
//   tile tt = env >> 5 owns wrc_tile_bytes(R) bytes: u64 planes [2R+3][32] (type row with padded index pr
//   sits in ring slot pr % (2R+3), column = env & 31) followed by u32 planes [7][4][32] (word w of the
//   nibble row with padded index pn sits in plane (pn % 7) * 4 + w).  Both are rings: a move of one row
//   replaces exactly the slot of the row that left the window.
// The planes stay the source of truth (every patch goes to both); the WRC is rebuilt from them by
// k_wrc_build whenever something other than k_step_tile has changed the state.
__host__ __device__ constexpr int wrc_type_slots(int R) { return 2 * R + 3; }
__host__ __device__ constexpr int wrc_tile_bytes(int R) { return wrc_type_slots(R) * 256 + 7 * 4 * 128; }

struct StepIO {
    const long long* actions;
    float* obs;
    float* reward;
    uint8_t* done;
    uint8_t* terminated;
    uint8_t* truncated;
    float* terminal_obs;
};

// ---------------------------------------------------------------- env record
struct EnvRec {
    int x, y, flags, thirsty;
    int step, explored;
    int total_free, collisions;
    int episode;
    int watered;
    double ret;
};

__device__ __forceinline__ EnvRec unpack_rec(const uint4& a, const uint4& b) {
    EnvRec r;
    r.x = a.x & 0xff;
    r.y = (a.x >> 8) & 0xff;
    r.flags = (a.x >> 16) & 0xff;
    r.thirsty = a.x >> 24;
    r.step = a.y & 0xffff;
    r.explored = a.y >> 16;
    r.total_free = a.z & 0xffff;
    r.collisions = a.z >> 16;
    r.episode = (int)a.w;
    r.watered = b.x & 0xffff;
    r.ret = __hiloint2double((int)b.w, (int)b.z);
    return r;
}

__device__ __forceinline__ void pack_rec(const EnvRec& r, uint4& a, uint4& b) {
    a.x = (uint32_t)r.x | ((uint32_t)r.y << 8) | ((uint32_t)r.flags << 16) | ((uint32_t)r.thirsty << 24);
    a.y = (uint32_t)r.step | ((uint32_t)r.explored << 16);
    a.z = (uint32_t)r.total_free | ((uint32_t)r.collisions << 16);
    a.w = (uint32_t)r.episode;
    b.x = (uint32_t)r.watered;
    b.y = 0u;
    b.z = (uint32_t)__double2loint(r.ret);
    b.w = (uint32_t)__double2hiint(r.ret);
}

// ------------------------------------------------------------------- helpers
__device__ __forceinline__ int cell_of(uint64_t word, int ylow) { return (int)((word >> (2 * ylow)) & 3ull); }

// nibble plane addressing of grid cell (x, y): u32 word index inside the env, and bit shift
constexpr int kVisRowPad = 3;   // border rows above / below the grid in the nibble plane
__device__ __forceinline__ int nib_word(int x, int y, int VW) { return (x + kVisRowPad) * VW + ((y + 2) >> 3); }
__device__ __forceinline__ int nib_shift(int y) { return 4 * ((y + 2) & 7); }

// valid-column mask (bit0 of each cell) for word w of a row
__device__ __forceinline__ uint64_t col_mask(int G, int w) {
    int n = G - w * 32;
    if (n >= 32) return kObstAll;
    return kObstAll & ((1ull << (2 * n)) - 1ull);
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------ Philox4x32-10 (counter RNG)
// Map of env `genv`, episode `ep` is a pure function of (seed, genv, ep): counter =
// (draw j, ep, genv low 32, (genv >> 32) * 4 + stream), key = seed.  Streams: 0 obstacle
// clusters, 1 plants, 2 rover.  Mirrored on the host in oracle/philox_mapgen.py.
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ void map_draw(const Params& p, long long genv, int ep, int stream, uint32_t j,
                                         uint32_t (&out)[4]) {
    out[0] = j;
    out[1] = (uint32_t)ep;
    out[2] = (uint32_t)((unsigned long long)genv & 0xffffffffull);
    out[3] = (uint32_t)(((unsigned long long)genv >> 32) * 4ull + (unsigned)stream);
    philox4x32_10(out, p.seed_lo, p.seed_hi);
}

__device__ __forceinline__ uint32_t bounded(uint32_t w, uint32_t n) { return __umulhi(w, n); }

// ------------------------------------------------- per-block shared-memory tables
struct Tables {
    const int8_t* off;    // [C][R][2]
    const float* dist;    // [R+1]
    const float* pos;     // [G]
    const float* visit;   // [16]: min(k,10)/10 for every nibble value k
    const float* rw32;    // [12]
    const double* rw64;   // [12]
    const float* onehot;  // [4][4] identity rows: the one-hot entity encoding (plantos_env.py:290-292)
};

__host__ __device__ inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// byte layout of the table block; identical on host (for the launch size) and device
__host__ __device__ inline int tables_bytes(int G, int R, int C) {
    int b = 2 * kRwCount * 8;              // rw64
    b += 16 * 4;                           // onehot (16-byte aligned: 96 B in)
    b += 2 * kRwCount * 4;                 // rw32
    b += (R + 1) * 4 + G * 4 + 16 * 4;     // dist, pos, visit
    b += align_up(C * R * 2, 16);          // offsets
    return align_up(b, 16);
}

// Where each table sits in the shared-memory image.
__device__ __forceinline__ Tables tables_at(unsigned char* smem, int G, int R) {
    double* rw64 = reinterpret_cast<double*>(smem);
    float* onehot = reinterpret_cast<float*>(rw64 + 2 * kRwCount);
    float* rw32 = onehot + 16;
    float* dist = rw32 + 2 * kRwCount;
    float* pos = dist + (R + 1);
    float* visit = pos + G;
    int8_t* off = reinterpret_cast<int8_t*>(visit + 16);
    Tables t;
    t.off = off; t.dist = dist; t.pos = pos; t.visit = visit; t.rw32 = rw32; t.rw64 = rw64; t.onehot = onehot;
    return t;
}

// Cooperative load by the whole block; ends with __syncthreads().
__device__ inline Tables load_tables(const Params& p, unsigned char* smem) {
    double* rw64 = reinterpret_cast<double*>(smem);
    float* onehot = reinterpret_cast<float*>(rw64 + 2 * kRwCount);
    float* rw32 = onehot + 16;
    float* dist = rw32 + 2 * kRwCount;
    float* pos = dist + (p.R + 1);
    float* visit = pos + p.G;
    int8_t* off = reinterpret_cast<int8_t*>(visit + 16);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 2 * kRwCount; i += nt) { rw64[i] = p.reward64[i]; rw32[i] = p.reward32[i]; }
    for (int i = tid; i < 16; i += nt) onehot[i] = ((i >> 2) == (i & 3)) ? 1.0f : 0.0f;
    for (int i = tid; i <= p.R; i += nt) dist[i] = p.dist_tab[i];
    for (int i = tid; i < p.G; i += nt) pos[i] = p.pos_tab[i];
    for (int i = tid; i < 16; i += nt) visit[i] = p.visit_tab[i < 10 ? i : 10];   // min(v, 10) / 10.0
    for (int i = tid; i < p.C * p.R * 2; i += nt) off[i] = p.lidar_off[i];
    __syncthreads();
    Tables t;
    t.off = off; t.dist = dist; t.pos = pos; t.visit = visit; t.rw32 = rw32; t.rw64 = rw64; t.onehot = onehot;
    return t;
}

// ------------------------------------------------------------ L2 residency hints
// The env state (records, type rows, visit nibbles) is re-read every step while observations,
// rewards and flags are written once and never read back by the simulator.  On B200 the
// state's per-step working set (a few 128-byte lines per env) fits the 126 MB L2 only if the
// write-once streams do not push it out, so state accesses can carry an L2::evict_last policy
// (observation stores carry evict-first, see plantos_fast.cuh).
struct PlainMem {
    __device__ __forceinline__ uint64_t ld64(const uint64_t* q) const { return *q; }
    __device__ __forceinline__ uint32_t ld32(const uint32_t* q) const { return *q; }
    __device__ __forceinline__ uint4 ld128(const uint4* q) const { return *q; }
    __device__ __forceinline__ void st64(uint64_t* q, uint64_t v) const { *q = v; }
    __device__ __forceinline__ void st32(uint32_t* q, uint32_t v) const { *q = v; }
    __device__ __forceinline__ void st128(uint4* q, const uint4& v) const { *q = v; }
};

struct KeepMem {   // every access tagged L2::evict_last
    uint64_t pol;
    __device__ __forceinline__ KeepMem() {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    }
    __device__ __forceinline__ uint64_t ld64(const uint64_t* q) const {
        uint64_t v;
        asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(q), "l"(pol) : "memory");
        return v;
    }
    __device__ __forceinline__ uint32_t ld32(const uint32_t* q) const {
        uint32_t v;
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(q), "l"(pol) : "memory");
        return v;
    }
    __device__ __forceinline__ uint4 ld128(const uint4* q) const {
        uint4 v;
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(q), "l"(pol) : "memory");
        return v;
    }
    __device__ __forceinline__ void st64(uint64_t* q, uint64_t v) const {
        asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" :: "l"(q), "l"(v), "l"(pol) : "memory");
    }
    __device__ __forceinline__ void st32(uint32_t* q, uint32_t v) const {
        asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(q), "r"(v), "l"(pol) : "memory");
    }
    __device__ __forceinline__ void st128(uint4* q, const uint4& v) const {
        asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;"
                     :: "l"(q), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    }
};

// ------------------------------------------------------------ env transition
// One PlantOSEnv.step up to (not including) the observation: plantos_env.py:160-222 and the
// termination/bonus logic of :176-181.  Executed by ONE thread for env e.  `row_word` is the
// type word holding the cell the action looks at (move target, or the rover's own cell when
// watering) or kObstAll when the target is out of bounds; `word_ptr` is where it lives (for
// the watering write); `vis_e` / `visov_e` are the env's nibble and overflow planes.
struct StepOut {
    int ridx;        // index into the reward tables
    int terminated;  // exploration >= 100 % (plantos_env.py:176,244-246)
    int truncated;   // step_count >= max_steps (:177)
    int watered;     // a thirsty plant was hydrated this step
    int moved;       // the rover entered (tx, ty): its visit count goes up by one
};

__device__ __forceinline__ void action_target(const EnvRec& r, long long action, int G, int& tx, int& ty, bool& inb) {
    if (action < 4) {
        // directions = N, E, S, W on (x, y) (plantos_env.py:186); negative actions index
        // the list from the end in Python, which is the same two low bits.
        const int d = (int)(action & 3);
        tx = r.x + ((d == 2) - (d == 0));
        ty = r.y + ((d == 1) - (d == 3));
        inb = ((unsigned)tx < (unsigned)G) && ((unsigned)ty < (unsigned)G);
    } else {
        tx = r.x; ty = r.y; inb = true;
    }
}

// The decision part, free of memory accesses: `t` is the code of the cell the action looks at
// (kObstacle when the move leaves the grid), `nib` the visit nibble of that cell (only
// meaningful for a valid move).  Sets o.moved / o.watered; the caller applies the two possible
// state writes (visit count +1 at (tx,ty); cell (tx,ty) thirsty -> hydrated).
// `expl_fresh` (CurriculumWrapper only, where visit counts outlive the episode): 1 / 0 = the
// cell is / is not yet in this episode's explored_map; -1 = no curriculum, same as nib == 0.
__device__ __forceinline__ StepOut transition_core(EnvRec& r, long long action, int tx, int ty, int t,
                                                   unsigned nib, int max_steps, int expl_fresh = -1) {
    StepOut o;
    o.watered = 0;
    o.moved = 0;
    r.step += 1;                                           // :162
    if (action < 4) {
        if (t != kObstacle) {                              // :193-195 (plants are walkable)
            const bool fresh = (nib == 0);                 // :197
            o.moved = 1;                                   // :203 visit_counts[new] += 1
            r.x = tx; r.y = ty;                            // :199
            r.explored += expl_fresh < 0 ? (int)fresh : expl_fresh;   // explored_map>0 count, :198-200,320
            o.ridx = fresh ? 0 : 1;                        // R_EXPLORATION / R_REVISIT
        } else {
            r.flags |= kFlagCollided;                      // :209
            r.collisions += 1;                             // :210
            o.ridx = 2;                                    // R_INVALID
        }
    } else {
        if (t == kThirsty) {                               // :217-219
            r.thirsty -= 1;
            r.watered += 1;
            o.watered = 1;
            o.ridx = 3;                                    // R_GOAL
        } else if (t == kHydrated) {
            o.ridx = 5;                                    // documented R_MISTAKE (see plantos.h)
        } else {
            o.ridx = 4;                                    // R_WATER_EMPTY, :221-222
        }
    }
    o.terminated = r.explored >= r.total_free;             // exploration_percentage >= 100
    o.truncated = r.step >= max_steps;
    if (o.terminated && !(r.flags & kFlagBonus)) {         // :179-181
        o.ridx += kRwCount;
        r.flags |= kFlagBonus;
    }
    return o;
}

// visit count of cell (tx,ty) += 1 given its nibble word `w` (already loaded): the count lives in
// the nibble until it saturates at 15, from then on in the u16 overflow plane.  Returns the new word.
template <class Mem>
__device__ __forceinline__ uint32_t bump_visit(uint32_t* vp, uint32_t w, int sh, uint16_t* ov, const Mem& mem) {
    const unsigned nib = (w >> sh) & 15u;
    if (nib < 15u) {
        w += 1u << sh;
        mem.st32(vp, w);
        if (nib == 14u) *ov = 15;
    } else {
        const unsigned v = *ov;
        if (v < 65535u) *ov = (uint16_t)(v + 1u);
    }
    return w;
}

// Transition against the planes in global memory (generic kernel).
template <class Mem>
__device__ __forceinline__ StepOut apply_action(EnvRec& r, long long action, int tx, int ty, bool inb,
                                                uint64_t row_word, uint64_t* word_ptr,
                                                uint32_t* vis_e, uint16_t* visov_e, int G, int VW,
                                                int max_steps, const Mem& mem, int expl_fresh = -1) {
    const int t = inb ? cell_of(row_word, ty & 31) : kObstacle;
    uint32_t* vp = vis_e + nib_word(tx, ty, VW);
    const int sh = nib_shift(ty);
    uint32_t w = 0;
    if (action < 4 && t != kObstacle) w = mem.ld32(vp);
    const StepOut o = transition_core(r, action, tx, ty, t, (w >> sh) & 15u, max_steps, expl_fresh);
    if (o.moved) bump_visit(vp, w, sh, visov_e + tx * G + ty, mem);
    if (o.watered) mem.st64(word_ptr, row_word ^ (1ull << (2 * (ty & 31))));  // 3 -> 2
    return o;
}

}  // namespace plantos_dev
