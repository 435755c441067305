// plantos_generic.cuh -- one-warp-per-env kernels: any supported (G, P, O, R, C).
//
//   k_step_generic   PlantOSEnv.step + SB3 auto-reset            (plantos_env.py:160-183)
//   k_reset_all      PlantOSEnv.reset for every env              (plantos_env.py:125-158)
//   k_get_state / k_set_state / k_get_scalars / k_get_returns / k_stats_out
//
// The warp-cooperative pieces (step_env_warp, reset_env_warp, build_obs_warp) are also the
// rare-path code of the fast kernel in plantos_fast.cuh (auto-reset, ragged last tile).
#pragma once
#include "plantos_common.cuh"

namespace plantos_dev {

constexpr int kGenericWarps = 8;  // warps (= envs in flight) per block

// scratch of the maze generator (map_source = maze): the DFS stack (one u16 per meta cell) and the visited bits
// of the (G-1)/6 x (G-1)/6 meta grid.  It lives right behind the type plane of a reset's scratch (where the
// observation row sits, which is not in use while a map is generated); every kernel leaves that much room.
__host__ __device__ inline int maze_scratch_bytes(int G) {
    const int m = (G - 1) / 6;
    return align_up(m * m * 2, 4) + ((m * m + 31) / 32) * 4;
}
// per-warp scratch: the env's (unpadded) type plane, G*W u64, followed by one observation row
__host__ __device__ inline int generic_warp_scratch_bytes(int G, int W, int D) {
    const int row = align_up(D * 4, 16), maze = align_up(maze_scratch_bytes(G), 16);
    return align_up(G * W * 8, 16) + (row > maze ? row : maze);
}

// ----------------------------------------------------------------- observation
// plantos_env.py:251-315.  `plane` (shared memory, indexed by grid row) must hold the type
// rows x-R .. x+R that lie inside the grid; visit nibbles are read from global memory.
// `fresh_visits`: show the visit window of a just-reset env (1 under the rover, 0 elsewhere) instead of
// the stored plane -- CurriculumWrapper restores the persistent counts only AFTER env.reset() has
// built its observation (A2C_training.py:82-87).
__device__ __forceinline__ void build_obs_warp(const Params& p, const Tables& t, const uint64_t* plane,
                                               const uint32_t* vis_e, int x, int y, float* obs_s, int lane,
                                               bool fresh_visits = false) {
    const int G = p.G, W = p.W, R = p.R, C = p.C;
    for (int i = lane; i < C; i += 32) {                 // one lane per ray
        int dist = R, kind = kEmpty;                     // :262-263
        const int8_t* o = t.off + i * R * 2;
        for (int r = 1; r <= R; ++r) {                   // integer march over the offset table
            const int cx = x + o[2 * (r - 1)];
            const int cy = y + o[2 * (r - 1) + 1];
            int tt;
            if ((unsigned)cx >= (unsigned)G || (unsigned)cy >= (unsigned)G) tt = kObstacle;  // :271-274
            else tt = cell_of(plane[cx * W + (cy >> 5)], cy & 31);                           // :277-284
            if (tt != kEmpty) { dist = r; kind = tt; break; }
        }
        float* q = obs_s + 5 * i;                        // :286-292
        q[0] = t.dist[dist];
        q[1] = (kind == kEmpty) ? 1.0f : 0.0f;
        q[2] = (kind == kObstacle) ? 1.0f : 0.0f;
        q[3] = (kind == kHydrated) ? 1.0f : 0.0f;
        q[4] = (kind == kThirsty) ? 1.0f : 0.0f;
    }
    if (lane == 0) {                                     // :294-296
        obs_s[5 * C] = t.pos[x];
        obs_s[5 * C + 1] = t.pos[y];
    }
    if (lane < 25) {                                     // :298-313; border nibbles (15) read as 1.0
        const int gx = x + lane / 5 - 2, gy = y + lane % 5 - 2;
        unsigned nib = (vis_e[nib_word(gx, gy, p.VW)] >> nib_shift(gy)) & 15u;
        if (fresh_visits)
            nib = ((unsigned)gx >= (unsigned)G || (unsigned)gy >= (unsigned)G) ? 15u : (lane == 12 ? 1u : 0u);
        obs_s[5 * C + 2 + lane] = t.visit[nib];
    }
    __syncwarp();
}

__device__ __forceinline__ void store_obs_row(const float* obs_s, float* dst, int D, int lane) {
    for (int i = lane; i < D; i += 32) dst[i] = obs_s[i];
}

// ---------------------------------------------------------------- maze generator
// The Gradio fork's 'maze' maps (gradio-app/plantos_env_new.py:408-604) with Philox draws: start from a
// grid full of obstacles, run a randomised depth-first search over the (G-1)/6 x (G-1)/6 meta grid, carve a
// 5x5 room per meta cell (with probability 0.3 each a 2x2 extension to the right / downwards, with 0.4 one
// corner cut) and a 5-wide corridor between consecutive rooms (with 0.2 a 2x2 bulge to one side).  Word j of
// Philox stream 3 is the j-th random decision, taken in the fork's order.  A serial algorithm on at most
// 21 x 21 meta cells; `scratch` = maze_scratch_bytes(G) bytes of shared memory.
struct MazeRng {
    const Params& p; long long genv; int ep; uint32_t j; uint32_t buf[4];
    __device__ __forceinline__ uint32_t next() {
        if ((j & 3u) == 0u) map_draw(p, genv, ep, 3, j >> 2, buf);
        return buf[j++ & 3u];
    }
    __device__ __forceinline__ bool chance(uint32_t thresh) { return next() < thresh; }   // random.random() < prob
};

// Executed by the whole warp with identical arguments: every lane owns the grid rows x with x % 32 == lane, so a
// row word is only ever touched by one lane and the carving needs no synchronisation until the plane is read.
__device__ __forceinline__ void maze_rect(uint64_t* plane, int G, int W, int x0, int x1, int y0, int y1, bool obstacle, int lane) {
    // cells [x0, x1) x [y0, y1) clipped to the grid become empty (or obstacle): one masked update per row word
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, G); y1 = min(y1, G);
    if (x0 >= x1 || y0 >= y1) return;
    for (int wd = y0 >> 5; wd <= (y1 - 1) >> 5; ++wd) {
        const int lo = max(y0, 32 * wd) - 32 * wd, n = min(y1, 32 * wd + 32) - 32 * wd - lo;     // n cells from cell lo of this word
        const uint64_t m = (n >= 32 ? ~0ull : ((1ull << (2 * n)) - 1ull)) << (2 * lo);
        for (int x = x0 + ((lane - x0) & 31); x < x1; x += 32) {
            uint64_t& w = plane[x * W + wd];
            w = obstacle ? ((w & ~m) | (kObstAll & m)) : (w & ~m);
        }
    }
}

__device__ __forceinline__ void maze_room(MazeRng& rng, uint64_t* plane, int G, int W, int mx, int my, int lane) {   // :479-517
    const int bx = mx * 6 + 1, by = my * 6 + 1;
    maze_rect(plane, G, W, bx, bx + 5, by, by + 5, false, lane);
    if (rng.chance(1288490188u)) maze_rect(plane, G, W, bx + 5, bx + 7, by + 2, by + 4, false, lane);   // 0.3: extend right
    if (rng.chance(1288490188u)) maze_rect(plane, G, W, bx + 2, bx + 4, by + 5, by + 7, false, lane);   // 0.3: extend down
    if (rng.chance(1717986918u)) {                                                                 // 0.4: cut one corner
        const int c = (int)bounded(rng.next(), 4u);                  // [(0,0), (4,0), (0,4), (4,4)]
        const int px = bx + ((c & 1) ? 4 : 0), py = by + ((c & 2) ? 4 : 0);
        maze_rect(plane, G, W, px, px + 1, py, py + 1, true, lane);
    }
}

// All 32 lanes run the search with the same draws (the decisions are a serial chain; computing them 32 times costs
// nothing extra) and each lane carves the rows it owns (maze_rect): a rectangle is one masked update instead of a
// loop over its rows.  The stack and the visited bits have one writer (lane 0) and warp barriers around the reads.
__device__ inline void maze_generate(const Params& p, long long genv, int ep, uint64_t* plane, unsigned char* scratch, int lane) {
    const int G = p.G, W = p.W, m = (G - 1) / 6;
    for (int x = lane; x < G; x += 32)
        for (int w = 0; w < W; ++w) plane[x * W + w] = kObstAll;       // (columns >= G are obstacles in any case)
    if (m < 1) return;
    uint16_t* stack = reinterpret_cast<uint16_t*>(scratch);
    uint32_t* visited = reinterpret_cast<uint32_t*>(scratch + align_up(m * m * 2, 4));
    if (lane == 0)
        for (int i = 0; i < (m * m + 31) / 32; ++i) visited[i] = 0u;
    MazeRng rng{p, genv, ep, 0u, {0u, 0u, 0u, 0u}};
    int cx = (int)bounded(rng.next(), (uint32_t)m), cy = (int)bounded(rng.next(), (uint32_t)m);   // randint(0, meta - 1) twice
    int sp = 0;
    if (lane == 0) {                                                   // (one writer; every lane reads after the barrier)
        stack[sp] = (uint16_t)(cx * m + cy);
        visited[(cx * m + cy) >> 5] |= 1u << ((cx * m + cy) & 31);
    }
    ++sp;
    maze_room(rng, plane, G, W, cx, cy, lane);
    while (sp > 0) {                                                   // :434-457
        __syncwarp();
        const int cur = stack[sp - 1];
        cx = cur / m; cy = cur - cx * m;
        int cand[4], nc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                  // [(0,1), (0,-1), (1,0), (-1,0)]
            const int nx = cx + (k == 2) - (k == 3), ny = cy + (k == 0) - (k == 1);
            if ((unsigned)nx < (unsigned)m && (unsigned)ny < (unsigned)m && !((visited[(nx * m + ny) >> 5] >> ((nx * m + ny) & 31)) & 1u))
                cand[nc++] = k;
        }
        if (nc == 0) { --sp; continue; }
        const int k = cand[bounded(rng.next(), (uint32_t)nc)];
        const int dx = (k == 2) - (k == 3), dy = (k == 0) - (k == 1);
        const int nx = cx + dx, ny = cy + dy;
        // corridor (:540-560): 5 wide, over both meta cells
        if (dx == 0) maze_rect(plane, G, W, cx * 6 + 1, cx * 6 + 6, min(cy, ny) * 6 + 1, max(cy, ny) * 6 + 7, false, lane);
        else maze_rect(plane, G, W, min(cx, nx) * 6 + 1, max(cx, nx) * 6 + 7, cy * 6 + 1, cy * 6 + 6, false, lane);
        if (rng.chance(858993459u)) {                                  // 0.2: bulge (:562-580)
            const int mx = (cx + nx) / 2, my = (cy + ny) / 2;
            const int dir = bounded(rng.next(), 2u) ? 1 : -1;          // choice([-1, 1])
            if (dx == 0) maze_rect(plane, G, W, mx * 6 + 2 + dir * 2, mx * 6 + 4 + dir * 2, my * 6 + 2, my * 6 + 4, false, lane);
            else maze_rect(plane, G, W, mx * 6 + 2, mx * 6 + 4, my * 6 + 2 + dir * 2, my * 6 + 4 + dir * 2, false, lane);
        }
        maze_room(rng, plane, G, W, nx, ny, lane);
        __syncwarp();                                                  // (every lane has read this round's stack top and visited bits)
        if (lane == 0) {
            visited[(nx * m + ny) >> 5] |= 1u << ((nx * m + ny) & 31);
            stack[sp] = (uint16_t)(nx * m + ny);
        }
        ++sp;
    }
}

// ----------------------------------------------------------------------- reset
// New episode for env e (local index): builds the map in `plane` (shared), writes the type
// rows and a fresh visit-nibble plane (zeros, rover cell = 1, plantos_env.py:146-147; border
// 15) to global memory and returns the fresh record in all lanes.  `episode` selects the
// injected map / Philox counter and is stored incremented.
// `keep_visits` (curriculum): leave the visit planes as they are (persistent_visit_counts).
// `map_episode` >= 0 (curriculum with reuse_map): the map is the one of that earlier episode again.
// MAZE = false leaves the maze generator out: the specialised step kernels instantiate it that way (its serial
// DFS would sit in their register allocation) and hand the resets of maze handles to k_reset_done.
template <bool MAZE = true>
__device__ __forceinline__ EnvRec reset_env_warp(const Params& p, int e, int episode, uint64_t* plane, int lane,
                                                 bool keep_visits = false, int map_episode = -1) {
    const int G = p.G, W = p.W;
    const int nwords = G * W;
    int rx = 0, ry = 0;
    const int mep = map_episode >= 0 ? map_episode : episode;
    if (p.map_source == 1) {
        // recorded map (replaces plantos_env.py:338-372 for equivalence runs)
        int k = mep;
        if (k >= p.map_episodes) {
            if (lane == 0) atomicExch(p.err, -3);  // PLANTOS_ENOMAPS
            k = k % p.map_episodes;
        }
        const size_t mi = (size_t)e * p.map_episodes + k;
        const uint8_t* mc = p.map_cells + mi * (size_t)(G * G);
        for (int idx = lane; idx < nwords; idx += 32) {
            const int row = idx / W, w = idx - row * W;
            uint64_t word = 0;
            for (int c = 0; c < 32; ++c) {
                const int col = w * 32 + c;
                const uint64_t code = (col < G) ? (uint64_t)(mc[row * G + col] & 3) : (uint64_t)kObstacle;
                word |= code << (2 * c);
            }
            plane[idx] = word;
        }
        rx = p.map_rover[mi * 2];
        ry = p.map_rover[mi * 2 + 1];
        __syncwarp();
    } else {
        const long long genv = p.env_base + e;
        bool clusters = true;
        if (MAZE && p.map_source == 2) {
            // the fork's maze generator; like the fork (:463-467) it falls back to the cluster generator
            // when the maze leaves fewer than P + 1 free cells
            maze_generate(p, genv, mep, plane, reinterpret_cast<unsigned char*>(plane) + align_up(nwords * 8, 16), lane);
            __syncwarp();
            int nfree = 0;
            for (int idx = lane; idx < nwords; idx += 32) {
                const uint64_t word = plane[idx];
                nfree += __popcll(~(word | (word >> 1)) & col_mask(G, idx % W));
            }
            clusters = warp_sum_i(nfree) < p.P + 1;
            __syncwarp();
        }
        // procedural map, same construction as plantos_env.py:338-372 with Philox draws
        for (int idx = lane; idx < (clusters ? nwords : 0); idx += 32) {
            const int w = idx % W;
            plane[idx] = kObstAll & ~col_mask(G, w);     // only the beyond-grid padding
        }
        __syncwarp();
        for (int k = lane; k < (clusters ? p.nclusters : 0); k += 32) {   // :343-354, one cluster per lane
            uint32_t d[4];
            map_draw(p, genv, mep, 0, (uint32_t)k, d);
            const int cx = 2 + (int)bounded(d[0], (uint32_t)(G - 4));   // randint(2, G-3)
            const int cy = 2 + (int)bounded(d[1], (uint32_t)(G - 4));
            const int size = 2 + (int)(d[2] >> 31);                     // choice([2, 3])
            for (int dx = 0; dx < size; ++dx)
                for (int dy = 0; dy < size; ++dy) {
                    const int ox = cx + dx - 1, oy = cy + dy - 1;       // size // 2 == 1
                    if ((unsigned)ox < (unsigned)G && (unsigned)oy < (unsigned)G)
                        atomicOr(reinterpret_cast<unsigned long long*>(&plane[ox * W + (oy >> 5)]),
                                 1ull << (2 * (oy & 31)));
                }
        }
        __syncwarp();
        // plants: uniform sample without replacement from the free cells (:366) by rejection, thirsty with
        // probability thirsty_plant_prob (:368).  Draw j of stream 1 is candidate j; candidates are taken in
        // draw order while their cell is free.  32 candidates are drawn at once (one Philox block per lane):
        // a candidate is taken iff its cell is empty in the map so far and no EARLIER candidate of the batch
        // names the same cell (that one took it, or found it occupied) -- exactly the sequential rule.
        {
            const uint32_t ncell = (uint32_t)(G * G);
            int placed = 0;
            for (uint32_t j0 = 0; placed < p.P && j0 < (1u << 20); j0 += 32) {
                uint32_t d[4];
                map_draw(p, genv, mep, 1, j0 + (uint32_t)lane, d);
                const int cell = (int)bounded(d[0], ncell);
                const int cx = cell / G, cy = cell - cx * G;
                const bool empty = cell_of(plane[cx * W + (cy >> 5)], cy & 31) == kEmpty;
                const unsigned same = __match_any_sync(0xffffffffu, cell);
                const bool cand = empty && (same & ((1u << lane) - 1u)) == 0u;
                const unsigned cmask = __ballot_sync(0xffffffffu, cand);
                if (cand && placed + __popc(cmask & ((1u << lane) - 1u)) < p.P) {
                    const uint64_t code = ((unsigned long long)d[1] < p.thirsty_thresh) ? kThirsty : kHydrated;
                    atomicOr(reinterpret_cast<unsigned long long*>(&plane[cx * W + (cy >> 5)]), code << (2 * (cy & 31)));
                }
                placed = min(p.P, placed + __popc(cmask));
                __syncwarp();
            }
            // rover: uniform over free cells that hold no plant (:372): the first candidate of stream 2 on an empty cell
            for (uint32_t j0 = 0; j0 < (1u << 20); j0 += 32) {
                uint32_t d[4];
                map_draw(p, genv, mep, 2, j0 + (uint32_t)lane, d);
                const int cell = (int)bounded(d[0], ncell);
                const int cx = cell / G, cy = cell - cx * G;
                const unsigned emask = __ballot_sync(0xffffffffu, cell_of(plane[cx * W + (cy >> 5)], cy & 31) == kEmpty);
                if (emask) {
                    const int src = __ffs(emask) - 1;
                    rx = __shfl_sync(0xffffffffu, cx, src);
                    ry = __shfl_sync(0xffffffffu, cy, src);
                    break;
                }
            }
        }
        __syncwarp();
    }
    // counts + write-out of the G grid rows (the wall rows above/below were set at create)
    int n_obst = 0, n_thirsty = 0;
    uint64_t* types_e = p.types + (size_t)e * p.TS + (size_t)p.TP * W;
    for (int idx = lane; idx < nwords; idx += 32) {
        const uint64_t word = plane[idx];
        const uint64_t m = col_mask(G, idx % W);
        n_obst += __popcll(word & ~(word >> 1) & m);
        n_thirsty += __popcll(word & (word >> 1) & m);
        types_e[idx] = word;
    }
    n_obst = warp_sum_i(n_obst);
    n_thirsty = warp_sum_i(n_thirsty);
    // visit nibbles: 0 inside the grid, 15 on the border / row padding, 1 under the rover
    uint32_t* vis_e = p.vis4 + (size_t)e * p.VE;
    const int VW = p.VW;
    if (p.cur_mode) {                                    // explored_map: zeros, 2 under the rover (:144-145)
        uint32_t* ex = p.expl + (size_t)e * nwords;
        for (int idx = lane; idx < nwords; idx += 32) ex[idx] = (idx == rx * W + (ry >> 5)) ? (1u << (ry & 31)) : 0u;
    }
    for (int wi = lane; wi < (keep_visits ? 0 : p.VE); wi += 32) {
        const int vx = wi / VW - kVisRowPad, c0 = (wi % VW) * 8 - 2;
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int vy = c0 + k;
            uint32_t v = 15u;
            if ((unsigned)vx < (unsigned)G && (unsigned)vy < (unsigned)G) v = (vx == rx && vy == ry) ? 1u : 0u;
            word |= v << (4 * k);
        }
        vis_e[wi] = word;
    }
    __syncwarp();
    EnvRec r;
    r.x = rx; r.y = ry; r.flags = 0; r.thirsty = n_thirsty;
    r.step = 0; r.explored = 1;                          // :130-133, :236
    r.total_free = G * G - n_obst; r.collisions = 0;
    r.episode = episode + 1; r.watered = 0; r.ret = 0.0;
    return r;
}

// (Re)build env e's window-ring-cache entry from the planes in global memory for rover row x.
// Warp-cooperative; the caller has made the planes' latest contents visible (__syncwarp after
// reset_env_warp's stores).  No-op without a WRC.
__device__ __forceinline__ void wrc_build_env_warp(const Params& p, size_t e, int x, int lane) {
    if (!p.wrc) return;
    const int ntr = wrc_type_slots(p.R);
    unsigned char* tile = p.wrc + (e >> 5) * (size_t)wrc_tile_bytes(p.R);
    const int j = (int)(e & 31);
    for (int i = lane; i < ntr; i += 32) {                // padded type rows x+1 .. x+ntr (W == 1)
        const int pr = x + 1 + i;
        reinterpret_cast<uint64_t*>(tile)[(pr % ntr) * 32 + j] = p.types[e * p.TS + pr];
    }
    if (lane < 28) {                                      // padded nibble rows x .. x+6, four words each (VW == 4)
        const int pn = x + (lane >> 2), w = lane & 3;
        reinterpret_cast<uint32_t*>(tile + ntr * 256)[((pn % 7) * 4 + w) * 32 + j] = p.vis4[e * p.VE + pn * 4 + w];
    }
}

// The same entry right after reset_env_warp, without reading the planes back: the type rows come from the
// shared-memory `plane` the reset has just built (grid rows; every padded row outside the grid is a wall
// row), the nibble rows are the fresh ones (0, rover cell 1, border 15) unless the visit counts were kept
// (curriculum), in which case they are read from the plane in global memory.  `s_ring` != 0: also store the
// entry into the caller's resident copy of the tile's rings (shared-memory byte address of the tile image).
__device__ __forceinline__ void wrc_build_env_fresh_warp(const Params& p, size_t e, int x, int y, const uint64_t* plane,
                                                         bool keep_visits, int lane, uint32_t s_ring) {
    if (!p.wrc) return;
    const int ntr = wrc_type_slots(p.R), G = p.G;
    unsigned char* tile = p.wrc + (e >> 5) * (size_t)wrc_tile_bytes(p.R);
    const int j = (int)(e & 31);
    for (int i = lane; i < ntr; i += 32) {                // padded type rows x+1 .. x+ntr; grid row = padded row - TP
        const int pr = x + 1 + i, g = pr - p.TP;
        const uint64_t row = ((unsigned)g < (unsigned)G) ? plane[g] : kObstAll;
        reinterpret_cast<uint64_t*>(tile)[(pr % ntr) * 32 + j] = row;
        if (s_ring) asm volatile("st.shared.u64 [%0], %1;" :: "r"(s_ring + 256 * (pr % ntr) + 8 * j), "l"(row) : "memory");
    }
    if (lane < 28) {                                      // padded nibble rows x .. x+6; grid row = padded row - 3
        const int pn = x + (lane >> 2), w = lane & 3, vx = pn - kVisRowPad, c0 = w * 8 - 2;
        uint32_t word = 0;
        if (keep_visits) word = p.vis4[e * p.VE + pn * 4 + w];
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int vy = c0 + k;
                uint32_t v = 15u;
                if ((unsigned)vx < (unsigned)G && (unsigned)vy < (unsigned)G) v = (vx == x && vy == y) ? 1u : 0u;
                word |= v << (4 * k);
            }
        }
        const int pl = (pn % 7) * 4 + w;
        reinterpret_cast<uint32_t*>(tile + ntr * 256)[pl * 32 + j] = word;
        if (s_ring) asm volatile("st.shared.u32 [%0], %1;" :: "r"(s_ring + ntr * 256 + 128 * pl + 4 * j), "r"(word) : "memory");
    }
}

// episode statistics of finished envs -> fixed-point accumulators (one atomic per warp)
// and, when enabled, one episode-log entry per finished env (`env` = the lane's env index;
// slots are handed out with one atomic per warp)
__device__ __forceinline__ void accumulate_stats(const Params& p, bool done, const EnvRec& r, int terminated,
                                                 int truncated, int lane, int env, unsigned seq_add = 0u) {
    const unsigned dm = __ballot_sync(0xffffffffu, done);
    if (dm == 0u) return;
    if (p.ep_log) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(p.ep_log_count, (unsigned)__popc(dm));
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned slot = base + (unsigned)__popc(dm & ((1u << lane) - 1u));
        if (done && slot < (unsigned)p.ep_log_cap) {
            const unsigned long long rb = (unsigned long long)__double_as_longlong(r.ret);
            p.ep_log[2 * (size_t)slot] = make_uint4((unsigned)env, (unsigned)r.step, p.step_seq + seq_add,
                                                     (unsigned)(terminated | (truncated << 1)));
            p.ep_log[2 * (size_t)slot + 1] = make_uint4((unsigned)rb, (unsigned)(rb >> 32),
                                                         (unsigned)r.collisions | ((unsigned)r.watered << 16),
                                                         (unsigned)r.explored | ((unsigned)r.total_free << 16));
        }
    }
    long long v[kStatCount];
    v[0] = done ? 1 : 0;
    v[1] = done ? __double2ll_rn(r.ret * 1e6) : 0;
    v[2] = done ? r.step : 0;
    v[3] = done ? __double2ll_rn(((double)r.explored / (double)r.total_free) * 100.0 * 1e6) : 0;
    v[4] = done ? r.collisions : 0;
    v[5] = done ? r.watered : 0;
    v[6] = (done && terminated) ? 1 : 0;
    v[7] = (done && truncated) ? 1 : 0;
#pragma unroll
    for (int i = 0; i < kStatCount; ++i) {
        const long long s = warp_sum_ll(v[i]);
        if (lane == 0 && s != 0) atomicAdd(&p.stats[i], (unsigned long long)s);
    }
}

// CurriculumWrapper.reset (A2C_training.py:57-90): called by lane 0 before env e is reset for its episode
// number `episode`; returns bit 0 = the visit counts persist into the new episode, bits 1.. = the episode
// whose map the reset draws.  The reference always draws a new map (its `reset(seed=current_maze_seed)`
// never reaches the map generator); with plantos_set_curriculum_reuse_map the maze the wrapper means to keep
// really is kept: the map of the episode in which the current maze started (cur_cnt.y bits 8..).
__device__ __forceinline__ int curriculum_on_reset(const Params& p, int e, int episode) {
    int2 c = p.cur_cnt[e];
    double thr = p.cur_thr[e];
    int maze_ep = (int)((unsigned)c.y >> 8);
    c.x += 1;                                                  // episodes_on_current_maze += 1
    const bool timeout = c.x >= p.cur_max_eps;
    bool keep;
    if ((c.y & 1) || timeout) {
        if (c.y & 1) thr = fmin(thr + p.cur_inc, p.cur_max_thr);   // :68-72
        c.x = 0; c.y = 0;                                      // maze_completed = False; persistent = None
        keep = false;
        maze_ep = episode;                                     // a new maze starts with this episode
    } else {
        keep = (c.y & 2) != 0;                                 // :84-87
        c.y = (c.y & 0xff) | 2;
        if (episode == 0) maze_ep = 0;
    }
    c.y = (c.y & 0xff) | (int)(((unsigned)maze_ep & 0xffffffu) << 8);
    p.cur_cnt[e] = c;
    p.cur_thr[e] = thr;
    const int map_ep = p.cur_reuse_map ? maze_ep : episode;
    return (keep ? 1 : 0) | (map_ep << 1);
}

// -------------------------------------------------- one env, one warp (any config)
// MAZE = false (tails of the specialised kernels): no maze generator in the instantiation; a finished env of a
// maze handle keeps its terminal record and k_reset_done starts its new episode.
template <bool MAZE = true>
__device__ __forceinline__ void step_env_warp(const Params& p, const Tables& t, const StepIO& io, int e,
                                              uint64_t* plane, float* obs_s, int lane) {
    uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
    long long action = 0;
    if (lane == 0) {
        ra = p.rec[2 * (size_t)e];
        rb = p.rec[2 * (size_t)e + 1];
        action = io.actions[e];
    }
    const unsigned posw = __shfl_sync(0xffffffffu, ra.x, 0);
    int x = posw & 0xff, y = (posw >> 8) & 0xff;
    uint64_t* types_e = p.types + (size_t)e * p.TS + (size_t)p.TP * p.W;   // grid row 0
    uint32_t* vis_e = p.vis4 + (size_t)e * p.VE;
    uint16_t* visov_e = p.visov + (size_t)e * p.G * p.G;

    // stage the rows the step can look at: x-R-1 .. x+R+1 (move of one row + LIDAR reach)
    const int lo = max(0, x - p.R - 1), hi = min(p.G - 1, x + p.R + 1);
    // (four loads in flight per lane before the first store: a plain copy loop waits for every load in turn,
    // which on the XL preset -- 134 words per env -- was four dependent global round trips)
    for (int idx = lo * p.W + lane; idx < (hi + 1) * p.W; idx += 128) {
        const int end = (hi + 1) * p.W;
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = idx + 32 * u < end ? types_e[idx + 32 * u] : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (idx + 32 * u < end) plane[idx + 32 * u] = v[u];
    }
    __syncwarp();

    int flagw = 0;
    EnvRec r = {};
    if (lane == 0) {
        r = unpack_rec(ra, rb);
        int tx, ty; bool inb;
        action_target(r, action, p.G, tx, ty, inb);
        const int widx = inb ? tx * p.W + (ty >> 5) : 0;
        const uint64_t word = inb ? plane[widx] : kObstAll;
        uint64_t newword = word;   // (generic pointer to a local: plain accessors only)
        int expl_fresh = -1;
        if (p.cur_mode && inb && action < 4) {                 // this episode's explored_map (:198-200)
            uint32_t* ex = p.expl + ((size_t)e * p.G + tx) * p.W + (ty >> 5);
            const uint32_t bit = 1u << (ty & 31);
            expl_fresh = (*ex & bit) ? 0 : 1;
            if (cell_of(word, ty & 31) != kObstacle) *ex |= bit;
        }
        StepOut o = apply_action(r, action, tx, ty, inb, word, &newword, vis_e, visov_e, p.G, p.VW,
                                 p.max_steps, PlainMem(), expl_fresh);
        if (p.cur_mode) {                                      // CurriculumWrapper.step (:94-100)
            const double pct = ((double)r.explored / (double)r.total_free) * 100.0;    // plantos_env.py:334
            if (pct >= p.cur_thr[e]) {
                p.cur_cnt[e].y |= 1;                           // maze_completed
                if (p.cur_mode == 1) o.terminated = 1;
            }
        }
        if (o.watered) { plane[widx] = newword; types_e[widx] = newword; }
        r.ret += t.rw64[o.ridx];
        io.reward[e] = t.rw32[o.ridx];
        const int done = o.terminated | o.truncated;
        io.done[e] = (uint8_t)done;
        if (io.terminated) io.terminated[e] = (uint8_t)o.terminated;
        if (io.truncated) io.truncated[e] = (uint8_t)o.truncated;
        flagw = r.x | (r.y << 8) | (done << 16) | (o.terminated << 17) | (o.truncated << 18);
    }
    __syncwarp();
    flagw = __shfl_sync(0xffffffffu, flagw, 0);
    x = flagw & 0xff; y = (flagw >> 8) & 0xff;
    const bool done = (flagw >> 16) & 1;

    build_obs_warp(p, t, plane, vis_e, x, y, obs_s, lane);
    float* obs_row = io.obs + (size_t)e * p.D;
    if (!done) {
        store_obs_row(obs_s, obs_row, p.D, lane);
        if (lane == 0) pack_rec(r, ra, rb);
    } else {
        // SB3 auto-reset: terminal observation + info snapshot, then a fresh episode
        if (io.terminal_obs) store_obs_row(obs_s, io.terminal_obs + (size_t)e * p.D, p.D, lane);
        int episode = 0;
        if (lane == 0) {
            pack_rec(r, ra, rb);
            p.term_rec[2 * (size_t)e] = ra;
            p.term_rec[2 * (size_t)e + 1] = rb;
            episode = r.episode;
        }
        accumulate_stats(p, lane == 0, r, (flagw >> 17) & 1, (flagw >> 18) & 1, lane, e);
        episode = __shfl_sync(0xffffffffu, episode, 0);
        __syncwarp();
        if (MAZE || p.map_source != 2) {
            int keep = 0, map_ep = -1;
            if (p.cur_mode) {
                int cr = 0;
                if (lane == 0) cr = curriculum_on_reset(p, e, episode);
                cr = __shfl_sync(0xffffffffu, cr, 0);
                keep = cr & 1; map_ep = cr >> 1;
            }
            const EnvRec nr = reset_env_warp<MAZE>(p, e, episode, plane, lane, keep != 0, map_ep);
            build_obs_warp(p, t, plane, vis_e, nr.x, nr.y, obs_s, lane, keep != 0);
            store_obs_row(obs_s, obs_row, p.D, lane);
            if (lane == 0) pack_rec(nr, ra, rb);
        }                                                 // (else: the terminal record stays; k_reset_done takes over)
    }
    if (lane == 0) {
        p.rec[2 * (size_t)e] = ra;
        p.rec[2 * (size_t)e + 1] = rb;
    }
    __syncwarp();
}

#ifndef PLANTOS_GENERIC_MINBLOCKS
#define PLANTOS_GENERIC_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kGenericWarps * 32, PLANTOS_GENERIC_MINBLOCKS)
k_step_generic(const Params p, const StepIO io) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables t = load_tables(p, smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* scratch = smem + tables_bytes(p.G, p.R, p.C) + warp * generic_warp_scratch_bytes(p.G, p.W, p.D);
    uint64_t* plane = reinterpret_cast<uint64_t*>(scratch);
    float* obs_s = reinterpret_cast<float*>(scratch + align_up(p.G * p.W * 8, 16));
    for (int e = blockIdx.x * kGenericWarps + warp; e < p.N; e += gridDim.x * kGenericWarps)
        step_env_warp(p, t, io, e, plane, obs_s, lane);
}

// ------------------------------------------------------------------ reset (all)
__global__ void __launch_bounds__(kGenericWarps * 32)
k_reset_all(const Params p, float* obs) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables t = load_tables(p, smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* scratch = smem + tables_bytes(p.G, p.R, p.C) + warp * generic_warp_scratch_bytes(p.G, p.W, p.D);
    uint64_t* plane = reinterpret_cast<uint64_t*>(scratch);
    float* obs_s = reinterpret_cast<float*>(scratch + align_up(p.G * p.W * 8, 16));
    for (int e = blockIdx.x * kGenericWarps + warp; e < p.N; e += gridDim.x * kGenericWarps) {
        int episode = 0;
        if (lane == 0) episode = (int)p.rec[2 * (size_t)e].w;
        episode = __shfl_sync(0xffffffffu, episode, 0);
        int keep = 0, map_ep = -1;
        if (p.cur_mode) {
            int cr = 0;
            if (lane == 0) cr = curriculum_on_reset(p, e, episode);
            cr = __shfl_sync(0xffffffffu, cr, 0);
            keep = cr & 1; map_ep = cr >> 1;
        }
        const EnvRec nr = reset_env_warp(p, e, episode, plane, lane, keep != 0, map_ep);
        build_obs_warp(p, t, plane, p.vis4 + (size_t)e * p.VE, nr.x, nr.y, obs_s, lane, keep != 0);
        store_obs_row(obs_s, obs + (size_t)e * p.D, p.D, lane);
        if (lane == 0) {
            uint4 ra, rb;
            pack_rec(nr, ra, rb);
            p.rec[2 * (size_t)e] = ra;
            p.rec[2 * (size_t)e + 1] = rb;
        }
        __syncwarp();
    }
}

// Deferred auto-reset (maze handles on the specialised kernels): the step kernel has written the terminal
// record, flags and terminal observation of every finished env but left its episode running; this kernel
// starts the new episodes (map, fresh observation into io.obs, record, window ring cache).  One warp per 32
// envs scans their done flags.
__global__ void __launch_bounds__(kGenericWarps * 32)
k_reset_done(const Params p, const StepIO io) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Tables t = load_tables(p, smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* scratch = smem + tables_bytes(p.G, p.R, p.C) + warp * generic_warp_scratch_bytes(p.G, p.W, p.D);
    uint64_t* plane = reinterpret_cast<uint64_t*>(scratch);
    float* obs_s = reinterpret_cast<float*>(scratch + align_up(p.G * p.W * 8, 16));
    for (int e0 = (blockIdx.x * kGenericWarps + warp) * 32; e0 < p.N; e0 += gridDim.x * kGenericWarps * 32) {
        unsigned dmask = __ballot_sync(0xffffffffu, e0 + lane < p.N && io.done[e0 + lane] != 0);
        while (dmask) {
            const int e = e0 + __ffs(dmask) - 1;
            dmask &= dmask - 1;
            int episode = 0;
            if (lane == 0) episode = (int)p.rec[2 * (size_t)e].w;
            episode = __shfl_sync(0xffffffffu, episode, 0);
            int keep = 0, map_ep = -1;
            if (p.cur_mode) {
                int cr = 0;
                if (lane == 0) cr = curriculum_on_reset(p, e, episode);
                cr = __shfl_sync(0xffffffffu, cr, 0);
                keep = cr & 1; map_ep = cr >> 1;
            }
            const EnvRec nr = reset_env_warp<true>(p, e, episode, plane, lane, keep != 0, map_ep);
            build_obs_warp(p, t, plane, p.vis4 + (size_t)e * p.VE, nr.x, nr.y, obs_s, lane, true);
            store_obs_row(obs_s, io.obs + (size_t)e * p.D, p.D, lane);
            wrc_build_env_fresh_warp(p, (size_t)e, nr.x, nr.y, plane, keep != 0, lane, 0u);
            if (lane == 0) {
                uint4 ra, rb;
                pack_rec(nr, ra, rb);
                p.rec[2 * (size_t)e] = ra;
                p.rec[2 * (size_t)e + 1] = rb;
            }
            __syncwarp();
        }
    }
}

// one warp per env: the whole window ring cache from the planes (after reset / set_state / a step of
// another kernel)
__global__ void k_wrc_build(const Params p) {
    const int lane = threadIdx.x & 31;
    const size_t e = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    if (e >= (size_t)p.N) return;
    wrc_build_env_warp(p, e, (int)(p.rec[2 * e].x & 0xffu), lane);
}

// --------------------------------------------------------------- state access
__global__ void k_get_state(const Params p, uint8_t* cells, int32_t* visits) {
    const int gg = p.G * p.G;
    const size_t total = (size_t)p.N * gg;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i / gg), c = (int)(i - (size_t)e * gg);
        const int x = c / p.G, y = c - x * p.G;
        if (cells) cells[i] = (uint8_t)cell_of(p.types[(size_t)e * p.TS + (size_t)(x + p.TP) * p.W + (y >> 5)], y & 31);
        if (visits) {
            const unsigned nib = (p.vis4[(size_t)e * p.VE + nib_word(x, y, p.VW)] >> nib_shift(y)) & 15u;
            visits[i] = nib < 15u ? (int32_t)nib : (int32_t)p.visov[(size_t)e * gg + c];
        }
    }
}

// warp per env; recomputes total_cells / thirsty from the cells
__global__ void k_set_state(const Params p, const uint8_t* cells, const int32_t* visits, const int32_t* sc) {
    const int lane = threadIdx.x & 31;
    const int e = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (e >= p.N) return;
    const int G = p.G, W = p.W, nwords = G * W, gg = G * G;
    uint4 ra = p.rec[2 * (size_t)e], rb = p.rec[2 * (size_t)e + 1];
    EnvRec r = unpack_rec(ra, rb);
    uint64_t* types_e = p.types + (size_t)e * p.TS + (size_t)p.TP * W;
    if (cells) {
        int n_obst = 0, n_thirsty = 0;
        for (int idx = lane; idx < nwords; idx += 32) {
            const int row = idx / W, w = idx - row * W;
            uint64_t word = 0;
            for (int c = 0; c < 32; ++c) {
                const int col = w * 32 + c;
                const uint64_t code = (col < G) ? (uint64_t)(cells[(size_t)e * gg + row * G + col] & 3) : (uint64_t)kObstacle;
                word |= code << (2 * c);
            }
            types_e[idx] = word;
            const uint64_t m = col_mask(G, w);
            n_obst += __popcll(word & ~(word >> 1) & m);
            n_thirsty += __popcll(word & (word >> 1) & m);
        }
        r.total_free = gg - warp_sum_i(n_obst);
        r.thirsty = warp_sum_i(n_thirsty);
        if (r.thirsty > 255 && lane == 0) atomicExch(p.err, -1);   // PLANTOS_EINVAL: the record counts thirsty plants in 8 bits
    }
    if (visits) {
        uint32_t* vis_e = p.vis4 + (size_t)e * p.VE;
        uint16_t* visov_e = p.visov + (size_t)e * gg;
        const int VW = p.VW;
        for (int wi = lane; wi < p.VE; wi += 32) {     // one lane owns a whole nibble word
            const int vx = wi / VW - kVisRowPad, c0 = (wi % VW) * 8 - 2;
            uint32_t word = 0;
            for (int k = 0; k < 8; ++k) {
                const int vy = c0 + k;
                uint32_t nib = 15u;
                if ((unsigned)vx < (unsigned)G && (unsigned)vy < (unsigned)G) {
                    int v = visits[(size_t)e * gg + vx * G + vy];
                    v = v < 0 ? 0 : (v > 65535 ? 65535 : v);
                    nib = v < 15 ? (uint32_t)v : 15u;
                    if (v >= 15) visov_e[vx * G + vy] = (uint16_t)v;
                }
                word |= nib << (4 * k);
            }
            vis_e[wi] = word;
        }
    }
    if (sc) {
        const size_t n = p.N;
        r.x = sc[0 * n + e]; r.y = sc[1 * n + e]; r.step = sc[2 * n + e]; r.explored = sc[3 * n + e];
        r.collisions = sc[6 * n + e];
        r.flags = (sc[7 * n + e] ? kFlagCollided : 0) | (sc[8 * n + e] ? kFlagBonus : 0);
        r.episode = sc[9 * n + e]; r.watered = sc[10 * n + e];
    }
    if (lane == 0) {
        pack_rec(r, ra, rb);
        p.rec[2 * (size_t)e] = ra;
        p.rec[2 * (size_t)e + 1] = rb;
    }
}

__global__ void k_get_scalars(const Params p, int which, int32_t* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.N) return;
    const uint4* src = which ? p.term_rec : p.rec;
    const EnvRec r = unpack_rec(src[2 * (size_t)e], src[2 * (size_t)e + 1]);
    const size_t n = p.N;
    out[0 * n + e] = r.x; out[1 * n + e] = r.y; out[2 * n + e] = r.step; out[3 * n + e] = r.explored;
    out[4 * n + e] = r.total_free; out[5 * n + e] = r.thirsty; out[6 * n + e] = r.collisions;
    out[7 * n + e] = (r.flags & kFlagCollided) ? 1 : 0; out[8 * n + e] = (r.flags & kFlagBonus) ? 1 : 0;
    out[9 * n + e] = r.episode; out[10 * n + e] = r.watered;
}

__global__ void k_get_returns(const Params p, int which, double* out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.N) return;
    const uint4* src = which ? p.term_rec : p.rec;
    out[e] = unpack_rec(src[2 * (size_t)e], src[2 * (size_t)e + 1]).ret;
}

__global__ void k_stats_out(unsigned long long* acc, double* out, int clear) {
    const int i = threadIdx.x;
    if (i >= kStatCount) return;
    const long long v = (long long)acc[i];
    out[i] = (i == 1 || i == 3) ? (double)v * 1e-6 : (double)v;
    if (clear) acc[i] = 0ull;
}

// MCTS rollout policy (mcts_custom_trainer.py:168-216): with probability 0.7 the move to the least
// visited valid neighbour (first minimum in N, E, S, W order; plants are walkable), else -- and when
// every move is blocked -- a uniformly random action.  The randomness is supplied by the caller as two
// uniforms in [0, 1) per env: u[2e] < 0.7 selects the heuristic, floor(5 * u[2e+1]) is the random action.
__global__ void k_policy_heuristic(const Params p, const float* u, long long* actions) {
    const int G = p.G, W = p.W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.N; e += gridDim.x * blockDim.x) {
        const uint32_t w0 = p.rec[2 * (size_t)e].x;
        const int x = (int)(w0 & 0xff), y = (int)((w0 >> 8) & 0xff);
        const uint64_t* types_e = p.types + (size_t)e * p.TS + (size_t)p.TP * W;
        const uint32_t* vis_e = p.vis4 + (size_t)e * p.VE;
        int best = -1, min_visits = 0x7fffffff;
        for (int a = 0; a < 4; ++a) {
            const int nx = x + ((a == 2) - (a == 0)), ny = y + ((a == 1) - (a == 3));
            if ((unsigned)nx >= (unsigned)G || (unsigned)ny >= (unsigned)G) continue;
            if (cell_of(types_e[nx * W + (ny >> 5)], ny & 31) == kObstacle) continue;
            const unsigned nib = (vis_e[nib_word(nx, ny, p.VW)] >> nib_shift(ny)) & 15u;
            const int v = nib < 15u ? (int)nib : (int)p.visov[(size_t)e * G * G + nx * G + ny];
            if (v < min_visits) { min_visits = v; best = a; }
        }
        int rnd = (int)(u[2 * (size_t)e + 1] * 5.0f);
        rnd = rnd > 4 ? 4 : (rnd < 0 ? 0 : rnd);
        actions[e] = (u[2 * (size_t)e] < 0.7f && best >= 0) ? best : rnd;
    }
}

// Packs what the fast kernel's prologue needs (one block, after every table upload): the
// shared-memory image of the tables, and for each of the 32 lanes its observation-phase constants
//   srcl[rr] = lane of the half-warp holding window row x+dx of LIDAR sample rr of ray (lane & 15)
//   shf[rr]  = left shift bringing column y+dy of that window word to bits 30, 31
//   vsrc/vsh = source lane and nibble shift of the two 5x5 visit cells the lane converts.
__global__ void k_pack_tables(const Params p, uint4* blob, int4* lane_tab) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nbytes = tables_bytes(p.G, p.R, p.C);
    // bytes between the tables are padding: zero them so that the image is deterministic
    for (int i = threadIdx.x; i < (nbytes >> 4); i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const Tables t = load_tables(p, smem);
    for (int i = threadIdx.x; i < (nbytes >> 4); i += blockDim.x) blob[i] = reinterpret_cast<const uint4*>(smem)[i];
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x, sub = lane & 15, hbase = lane & 16, R = p.R;
        int v[4 * kLaneTabVec];
        for (int i = 0; i < 4 * kLaneTabVec; ++i) v[i] = 0;
        for (int rr = 0; rr < R && rr < 8; ++rr) {
            int dx = 0, dy = 0;
            if (sub < p.C) { dx = t.off[(sub * R + rr) * 2]; dy = t.off[(sub * R + rr) * 2 + 1]; }
            v[rr] = hbase + dx + R;
            v[8 + rr] = 30 - 2 * (dy + R);
        }
        v[16] = hbase + sub / 5; v[17] = 4 * (sub % 5);
        v[18] = hbase + (sub + 16) / 5; v[19] = 4 * ((sub + 16) % 5);
        for (int q = 0; q < kLaneTabVec; ++q)
            lane_tab[lane * kLaneTabVec + q] = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

}  // namespace plantos_dev
