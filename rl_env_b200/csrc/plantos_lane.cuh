// plantos_lane.cuh -- k_step_lane: the whole step with ONE LANE PER ENV.
//
// Same shape limits as k_step_fast (W == 1, VW == 4, G + R <= 32, 2R+1 <= 16, C <= 16), plus: the
// LIDAR sample offsets must be the reference's own (plantos_lidar_gen.cuh holds them as
// compile-time tables; the host checks the uploaded table against them).
//
// Why: k_step_fast spends one half-warp per env on the observation and needs ~70 warp instructions,
// 19 shuffles and ~26 shared-memory wavefronts per env for it, which makes it bound by the SM's ALU
// and load/store pipes (profiles/r1_summary.md).  Here every lane owns one env for the whole step:
//   * the rover-centred window (2R+1 words of 2-bit cells) lives in REGISTERS, and because all
//     lanes march the same ray at the same time the sample offsets are immediates: a sample costs
//     two shifts, a ray ~30 instructions for 32 envs, no shuffle at all;
//   * a lane writes its floats into its own row of a 32-row x 32-column shared-memory tile (row
//     stride 33 words => conflict-free both ways); every time 32 columns are complete the tile is
//     flushed row by row with 128-byte coalesced streaming stores, so the row stores are spread
//     over the whole observation phase instead of arriving in one burst per warp;
//   * window rows arrive by per-lane cp.async (8-byte type rows, 4-byte nibble words) in a
//     [row][lane] layout, so the lane-per-env loads are conflict-free too.
// A warp walks its contiguous env range in macro tiles of 32 envs:
//   records + actions -> target words (prefetched during the previous tile) -> transition
//   (plantos_env.py:160-222) -> own windows -> observation (plantos_env.py:251-315) -> flush ->
//   auto-reset of finished envs (generic warp routines).
// One block of PLANTOS_LANE_WARPS warps per SM (10.5 KB of shared memory per warp).
//
// STATUS (round 1): EXPERIMENTAL, opt-in with PLANTOS_FAST_IMPL=lane.  Bit-exact on the whole GPU
// suite and 61 warp instructions per env instead of k_step_fast's 102, but 34.5 us per 131 072-env
// step against 25.1 us: with 12 warps per SM every warp runs its tile's fetch -> transition ->
// fetch -> observation chain almost alone (IPC 0.28 per scheduler, a quarter of the warp time in
// cp.async waits, the rest in LDS->use and ALU dependency stalls); 16 and 20 warps spill
// (36.6 / 60.5 us), a 13.7 KB full-row tile with 10 warps gives 39 us, staggered first tiles do
// not help.  See profiles/r1_summary.md (rows 38-41) for what a round-2 version would need.
#pragma once
#include "plantos_fast.cuh"
#include "plantos_lidar_gen.cuh"

namespace plantos_dev {

#ifndef PLANTOS_LANE_WARPS
#define PLANTOS_LANE_WARPS 12         // one block per SM (<= 168 registers per thread; 16+ warps spill)
#endif
constexpr int kLaneWarps = PLANTOS_LANE_WARPS;

// per-warp scratch: 32 records | 32 actions | 32+32 target words | type rows [2R+1][32] u64 |
// nibble words [5][2][32] u32 | obs tile [32][33] f32 (phase C borrows it as a D-float row)
constexpr int kLaneTileStride = 33;
__host__ __device__ constexpr int lane_warp_scratch_bytes(int R, int D) {
    return 32 * (32 + 8 + 8 + 4) + (2 * R + 1) * 256 + 5 * 2 * 128 + 32 * kLaneTileStride * 4;
}

template <int R, int C, bool KEEP>
__global__ void __launch_bounds__(kLaneWarps * 32, 1)
k_step_lane(const Params p, const StepIO io) {
    using Gen = LidarGen<R, C>;
    static_assert(Gen::ok, "no generated LIDAR offsets for this (R, C)");
    constexpr int D = 5 * C + 27;
    constexpr int NROW = 2 * R + 1;
    constexpr int VW = 4;                 // nibble words per visit row (G + 4 <= 32)
    constexpr int TP = R + 2;             // wall rows above the grid (== Params.TP)
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NROW <= 16 && C <= 16, "lane kernel shape limits");

    extern __shared__ __align__(16) unsigned char smem[];
    typename std::conditional<KEEP, KeepMem, PlainMem>::type const mem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef PLANTOS_EXP_TIMING
    unsigned ts_[10] = {};
#endif
    TSTAMP(0);
    const int G = p.G, VE = p.VE, TS = p.TS;
    unsigned char* scratch = smem + tables_bytes(G, R, C) + warp * lane_warp_scratch_bytes(R, D);
    uint4* const recb = reinterpret_cast<uint4*>(scratch);                                       // [32][2]
    long long* const actb = reinterpret_cast<long long*>(scratch + 1024);                        // [32]
    uint64_t* const tgt_t = reinterpret_cast<uint64_t*>(scratch + 1280);                         // [32]
    uint32_t* const tgt_v = reinterpret_cast<uint32_t*>(scratch + 1536);                         // [32]
    uint64_t* const twin = reinterpret_cast<uint64_t*>(scratch + 1664);                          // [NROW][32]
    uint32_t* const vwin = reinterpret_cast<uint32_t*>(scratch + 1664 + NROW * 256);             // [5][2][32]
    float* const tile = reinterpret_cast<float*>(scratch + 1664 + NROW * 256 + 1280);            // [32][33]

    // Every warp owns one contiguous range of p.fast_q envs (a multiple of 4, computed on the host).
    const int gwarp = blockIdx.x * kLaneWarps + warp, nwarps = gridDim.x * kLaneWarps;
    const int nfull = p.N & ~3;                                       // envs in whole 4-env groups
    const int wbase = min(nfull, gwarp * p.fast_q), wend = min(nfull, wbase + p.fast_q);

    // ---- prefetch helpers: every lane copies for its own env (cp.async holds no registers)
    const uint32_t s_scr = smem_u32(scratch);
    const uint32_t s_rec = s_scr + 32 * lane, s_act = s_scr + 1024 + 8 * lane;
    const uint32_t s_tgt_t = s_scr + 1280 + 8 * lane, s_tgt_v = s_scr + 1536 + 4 * lane;
    const uint32_t s_twin = s_scr + 1664 + 8 * lane, s_vwin = s_scr + 1664 + NROW * 256 + 4 * lane;
    auto fetch_rec = [&](int es) {                     // records + actions of the macro tile at es
        if (lane < min(32, wend - es)) {
            const size_t e = (size_t)es + lane;
            cp_async16(s_rec, p.rec + 2 * e);
            cp_async16(s_rec + 16, p.rec + 2 * e + 1);
            cp_async8(s_act, io.actions + e);
        }
        cp_async_commit();
    };
    auto issue_target = [&](int es) {                  // the two words the transition will look at
        if (lane < min(32, wend - es)) {
            const size_t e = (size_t)es + lane;
            const uint32_t w0 = recb[2 * lane].x;
            EnvRec q;
            q.x = (int)(w0 & 0xff); q.y = (int)((w0 >> 8) & 0xff);
            int tx, ty; bool inb;
            action_target(q, actb[lane], G, tx, ty, inb);
            // (tx, ty) may be one cell outside the grid: wall rows / border nibbles are there
            cp_async8(s_tgt_t, p.types + e * TS + TP + tx);
            cp_async4(s_tgt_v, p.vis4 + e * VE + nib_word(tx, ty, VW));
        }
        cp_async_commit();
    };

    // Programmatic dependent launch, as in k_step_fast: nothing mutable is touched before
    // griddep_wait() returns; the immutable table image is loaded while the previous launch ends.
    griddep_launch_dependents();
    const int n16 = tables_bytes(G, R, C) >> 4;        // <= blockDim.x (checked on the host)
    uint4 tab16 = make_uint4(0, 0, 0, 0);
    if ((int)threadIdx.x < n16) tab16 = __ldg(p.table_blob + threadIdx.x);
    griddep_wait();
    TSTAMP(1);
    if (wbase < wend) fetch_rec(wbase);                // in flight while the tables are staged
    if ((int)threadIdx.x < n16) reinterpret_cast<uint4*>(smem)[threadIdx.x] = tab16;
    __syncthreads();
    const Tables t = tables_at(smem, G, R);
    const uint32_t s_dist = smem_u32(t.dist), s_pos = smem_u32(t.pos), s_visit = smem_u32(t.visit);
    const uint32_t s_onehot = smem_u32(t.onehot), s_rw32 = smem_u32(t.rw32), s_rw64 = smem_u32(t.rw64);
    constexpr uint64_t LOWPAD = kObstAll & ((1ull << (2 * R)) - 1ull);
    float* const row_out = tile + lane * kLaneTileStride;     // this lane's row of the tile

    TSTAMP(2);
    if (wbase < wend) {
        cp_async_wait_all();
        TSTAMP(3);
        issue_target(wbase);
    }

    // The first tile of a warp is shorter by a warp-dependent amount (8, 16, 24 or 32 envs): all
    // warps of the grid start together with identical work, and without the stagger they would all
    // wait for memory and all compute at the same moments.
#ifndef PLANTOS_LANE_STAGGER
#define PLANTOS_LANE_STAGGER 0
#endif
    const int first_ts = PLANTOS_LANE_STAGGER ? 8 * (1 + (warp & 3)) : 32;
    for (int e0 = wbase, ts = 0; e0 < wend; e0 += ts) {
        ts = min(e0 == wbase ? first_ts : 32, wend - e0);  // envs in this macro tile (a multiple of 4)
        const int e_next = e0 + ts;
        const bool has_next = e_next < wend;
        const bool act = lane < ts;
        const size_t e = (size_t)e0 + lane;
        cp_async_wait_all();                              // own record, action and target words are here
        TSTAMP(4);

        // ---- transition (plantos_env.py:160-222)
        int done = 0, term = 0, trunc = 0;
        EnvRec r = {};
        if (act) {
            uint4 ra = recb[2 * lane], rb = recb[2 * lane + 1];
            const long long action = actb[lane];
            r = unpack_rec(ra, rb);
            int tx, ty; bool inb;
            action_target(r, action, G, tx, ty, inb);
            const uint64_t word = tgt_t[lane];
            const uint32_t vword = tgt_v[lane];
            const int t_cell = inb ? cell_of(word, ty & 31) : kObstacle;
            const int sh = nib_shift(ty);
            const StepOut o = transition_core(r, action, tx, ty, t_cell, (vword >> sh) & 15u, p.max_steps);
            if (o.moved)
                bump_visit(p.vis4 + e * VE + nib_word(tx, ty, VW), vword, sh, p.visov + e * G * G + tx * G + ty, mem);
            if (o.watered) mem.st64(p.types + e * TS + TP + tx, word ^ (1ull << (2 * (ty & 31))));   // 3 -> 2
            r.ret += lds_f64(s_rw64 + 8 * o.ridx);
            io.reward[e] = lds_f32(s_rw32 + 4 * o.ridx);
            term = o.terminated; trunc = o.truncated; done = term | trunc;
            io.done[e] = (uint8_t)done;
            if (io.terminated) io.terminated[e] = (uint8_t)term;
            if (io.truncated) io.truncated[e] = (uint8_t)trunc;
            pack_rec(r, ra, rb);
            mem.st128(p.rec + 2 * e, ra);
            mem.st128(p.rec + 2 * e + 1, rb);
            if (done) {
                p.term_rec[2 * e] = ra;
                p.term_rec[2 * e + 1] = rb;
            }
        }
        accumulate_stats(p, act && done, r, term, trunc, lane, (int)e);
        TSTAMP(5);
        if (has_next) fetch_rec(e_next);                  // own slot of the record buffer is free again

        // ---- own windows, centred on the post-move position (the copies read what this same
        // thread stored above): type rows x-R .. x+R = padded rows x+2 .. x+2R+2; nibble rows
        // x-2 .. x+2 = padded rows x+1 .. x+5, of each the two words holding nibbles y .. y+4
        if (act) {
            const uint64_t* tsrc = p.types + e * TS + (r.x + 2);
#pragma unroll
            for (int i = 0; i < NROW; ++i) cp_async8(s_twin + i * 256, tsrc + i);
            const unsigned w0 = (unsigned)r.y >> 3, w1 = w0 < 3u ? w0 + 1u : 3u;
            const uint32_t* vsrc = p.vis4 + e * VE + (size_t)(r.x + 1) * VW;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                cp_async4(s_vwin + (2 * i) * 128, vsrc + i * VW + w0);
                cp_async4(s_vwin + (2 * i + 1) * 128, vsrc + i * VW + w1);
            }
        }
        cp_async_commit();
        cp_async_wait_all();                              // (also the next tile's records)
        TSTAMP(6);
        if (has_next) issue_target(e_next);               // lands while the observation is built

        // ---- observation (plantos_env.py:251-315).  Every lane builds its env's row, 32 columns at
        // a time; lanes beyond the tile's last env compute on stale windows into rows that are never
        // flushed, which keeps the whole phase convergent.
        {
            const int x = r.x, y = r.y;
            float* const obs_tile = io.obs + (size_t)e0 * D;
            // flush columns c0 .. c0+n-1 of the tile's ts rows: one 4n-byte segment per row
            auto flush_cols = [&](int c0, int n) {
                __syncwarp();
                if (lane < n) {
                    float* dst = obs_tile + c0 + lane;
                    const float* src = tile + lane;
#pragma unroll 8
                    for (int rw = 0; rw < ts; ++rw) __stcs(dst + (size_t)rw * D, src[rw * kLaneTileStride]);
                }
                __syncwarp();
            };
            int col = 0;                               // compile-time after unrolling
            auto put = [&](float v) {
                row_out[col & 31] = v;
                ++col;
                if ((col & 31) == 0) flush_cols(col - 32, 32);
            };
            // rover-centred window words: cells y-R .. y+R of rows x-R .. x+R, walls outside
            unsigned w[NROW];
            const int sft = 2 * y;
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                const uint64_t ext = (twin[i * 32 + lane] << (2 * R)) | LOWPAD;
                w[i] = (unsigned)((ext >> sft) | ((kObstAll << 1) << (63 - sft)));
            }
            // the visit slices are fetched early: nibbles y .. y+4 of the five rows (:298-313)
            unsigned sl[5];
            const int vs = 4 * (y & 7);
#pragma unroll
            for (int i = 0; i < 5; ++i)
                sl[i] = __funnelshift_r(vwin[(2 * i) * 32 + lane], vwin[(2 * i + 1) * 32 + lane], vs);
            // LIDAR (:260-292): far sample first; each step shifts the accumulator left by one cell
            // and funnels the sample's two bits in from the top of the aligned row word, so that
            // sample rr ends up at bits 2rr, 2rr+1; the first non-empty sample is the hit
#pragma unroll
            for (int ray = 0; ray < C; ++ray) {
                unsigned acc = 0;
#pragma unroll
                for (int rr = R - 1; rr >= 0; --rr)
                    acc = __funnelshift_l(w[Gen::dx(ray, rr) + R] << (30 - 2 * (Gen::dy(ray, rr) + R)), acc, 2);
                const unsigned m = (acc | (acc >> 1)) & 0x55555555u;
                const int bit = __ffs(m) - 1;                 // -1 when nothing was hit
                const int dist = m ? (bit >> 1) + 1 : R;
                const int kind = m ? (acc >> bit) & 3 : kEmpty;
                const float fd = lds_f32(s_dist + 4 * dist);
                const float4 oh = lds_f32x4(s_onehot + 16 * kind);
                put(fd); put(oh.x); put(oh.y); put(oh.z); put(oh.w);
            }
            put(lds_f32(s_pos + 4 * x));                      // :294-296
            put(lds_f32(s_pos + 4 * y));
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) put(lds_f32(s_visit + 4 * ((sl[i] >> (4 * j)) & 15u)));
            if (col & 31) flush_cols(col & ~31, col & 31);
        }
        TSTAMP(8);
#ifdef PLANTOS_EXP_TIMING
        { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); ts_[9] = sm; }
        if (lane == 0 && ts >= 8)
            for (int k = 0; k < 5; ++k) {
                uint4 v = p.term_rec[2 * ((size_t)e0 + k) + 1];
                v.z = ts_[2 * k]; v.w = ts_[2 * k + 1];
                p.term_rec[2 * ((size_t)e0 + k) + 1] = v;
            }
#endif

        // ---- auto-reset of finished envs (rare; warp-cooperative generic code; the window and
        // tile areas are free now and serve as its scratch)
        unsigned dmask = __ballot_sync(FULL, act && done);
        uint64_t* plane = twin;
        while (dmask) {
            const int j = __ffs(dmask) - 1;
            dmask &= dmask - 1;
            const size_t ej = (size_t)e0 + j;
            const int episode = __shfl_sync(FULL, r.episode, j);
            const int px = __shfl_sync(FULL, r.x, j), py = __shfl_sync(FULL, r.y, j);
            const uint64_t* types_e = p.types + ej * TS + TP;
            const uint32_t* vis_e = p.vis4 + ej * VE;
            if (io.terminal_obs) {
                for (int idx = lane; idx < G; idx += 32) plane[idx] = types_e[idx];
                __syncwarp();
                build_obs_warp(p, t, plane, vis_e, px, py, tile, lane);
                store_obs_row(tile, io.terminal_obs + ej * D, D, lane);
                __syncwarp();
            }
            const EnvRec nr = reset_env_warp(p, (int)ej, episode, plane, lane);
            build_obs_warp(p, t, plane, vis_e, nr.x, nr.y, tile, lane);
            store_obs_row(tile, io.obs + ej * D, D, lane);
            if (lane == 0) {
                uint4 qa, qb;
                pack_rec(nr, qa, qb);
                p.rec[2 * ej] = qa;
                p.rec[2 * ej + 1] = qb;
            }
            __syncwarp();
        }
    }

    // ragged tail: envs beyond the last 4-env group, one at a time (no copies are in flight here)
    if (gwarp == nwarps - 1)
        for (int e = nfull; e < p.N; ++e) step_env_warp(p, t, io, e, twin, tile, lane);
}

}  // namespace plantos_dev
