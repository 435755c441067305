// plantos_fast.cuh -- the sm_100a hot kernel for the reference presets.
//
// Requirements (checked on the host): W == 1 and VW == 4 (G <= 28), G + R <= 32, 2R+1 <= 16,
// C <= 16.  Both the training preset (G25 R6 C16, D=107; A2C_training.py:206-212) and the ctor
// default (G21 R2 C10, D=77; plantos_env.py:25-26) qualify; everything else runs
// k_step_generic.
//
// Persistent grid (PLANTOS_FAST_MINBLOCKS blocks per SM); every warp owns a contiguous range of
// envs and walks it in TILES of EPW consecutive envs through a two-deep software pipeline, so that the memory latency of the next tile hides under the arithmetic and the
// observation stores of the current one:
//   fetch    the 32-byte records and the actions are copied to shared memory two tiles ahead;
//            everything else the step can touch is fetched one tile ahead, centred on the
//            PRE-move position with a margin of one cell: 2R+4 rows of the wall-padded type
//            plane (128 contiguous bytes for R=6) and 7 rows of the visit-nibble plane (112
//            contiguous bytes) -- 15 asynchronous 16-byte global->shared copies (cp.async)
//            per env, spread over the 32/EPW lanes that share an env.  They are issued right
//            after the current tile's phase A, into the other of the warp's two window
//            buffers, and are the ONLY reads of plane state the step performs.
//   phase A  one lane per env, entirely out of shared memory: the transition
//            (plantos_env.py:160-222) -- target-cell lookup, visit count, watering, reward,
//            termination.  The (at most two) modified words go back to global memory and are
//            patched in the shared copy; reward / done / record stores are coalesced.
//   phase B  one HALF-WARP per env, two envs per half-warp interleaved stage by stage for ILP
//            -- the observation (plantos_env.py:251-315) from the shared windows: 2R+1 lanes
//            shift their type row into a rover-centred window word (the padding makes bounds
//            checks unnecessary); one lane per ray marches the integer offset table with a
//            warp shuffle as the row lookup; five lanes cut the 20-bit slice of their visit
//            row that the 5x5 window needs and the 25 cell lanes read it by shuffle.  Rows are
//            assembled in a 4-env shared-memory tile whose 16*D bytes are 16-byte aligned in
//            the [N, D] fp32 buffer and leave with streaming 128-bit stores
//            (st.global.cs.v4, evict-first), so the write-once observation stream does not
//            evict the env state from L2.
//   phase C  whole warp, rare -- SB3 auto-reset of finished envs (terminal observation,
//            Philox / injected map, fresh observation) via the generic warp routines.
// The last N % 4 envs are stepped one by one with step_env_warp.
#pragma once
#include <type_traits>
#include "plantos_generic.cuh"

namespace plantos_dev {

#ifndef PLANTOS_FAST_MINBLOCKS
#define PLANTOS_FAST_MINBLOCKS 4
#endif
#ifndef PLANTOS_FAST_WARPS
#define PLANTOS_FAST_WARPS 7         // 7 warps x 4 blocks = 28 resident warps per SM (<= 72 registers each)
#endif
constexpr int kFastWarps = PLANTOS_FAST_WARPS;
constexpr int kVisWinRows = 7;       // nibble rows x-3 .. x+3 around the pre-move position
constexpr int kVisWinBytes = kVisWinRows * 16;
// type rows fetched per env: x-R-1 .. x+R+1 (2R+3 rows) plus one because the copy starts on an
// even row (16-byte aligned source and size)
__host__ __device__ constexpr int type_win_rows(int R) { return 2 * R + 4; }

// per-warp scratch: two window buffers [type windows | visit windows] and one 4-env obs tile.
// Phase C / the tail reuse a window buffer as the generic code's type plane.
__host__ __device__ inline int fast_win_bytes(int EPW, int R, int G) {
    int win = EPW * (type_win_rows(R) * 8 + kVisWinBytes);
    return win < align_up(G * 8, 16) ? align_up(G * 8, 16) : win;
}
__host__ __device__ inline int fast_warp_scratch_bytes(int EPW, int R, int G, int D) {
    return 2 * fast_win_bytes(EPW, R, G) + 16 * D + 2 * EPW * 40;   // + two record/action buffers
}

// ---- asynchronous global->shared copies (16-byte cp.async, L2 only) ------------------------
// The per-env windows sit at per-env addresses, so they are fetched with per-lane 16-byte
// cp.async copies (a cp.async.bulk / TMA copy needs warp-uniform operands: issuing one per
// env costs a ~15-instruction elect loop per copy, which measured at 27 instructions per env).
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void cp_async16(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst), "l"(gsrc) : "memory");
}
// table reads through 32-bit shared addresses: the tables never change after load_tables, so the
// asm is not volatile and the compiler may schedule / combine these loads freely
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void cp_async8(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int R, int C, int EPW, bool KEEP>
__global__ void __launch_bounds__(kFastWarps * 32, PLANTOS_FAST_MINBLOCKS)
k_step_fast(const Params p, const StepIO io) {
    constexpr int D = 5 * C + 27;
    constexpr int NROW = 2 * R + 1;
    constexpr int VW = 4;                 // nibble words per visit row (G + 4 <= 32)
    constexpr int TP = R + 2;             // wall rows above the grid (== Params.TP)
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TWR = type_win_rows(R);
    constexpr int kTypeWinBytes = TWR * 8;
    static_assert(NROW <= 16 && C <= 16, "fast kernel shape limits");
    static_assert(EPW % 4 == 0 && EPW <= 32, "tile must be whole 4-env groups");

    extern __shared__ __align__(16) unsigned char smem[];
    const Tables t = load_tables(p, smem);
    typename std::conditional<KEEP, KeepMem, PlainMem>::type const mem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = p.G, VE = p.VE, TS = p.TS;
    const int win_bytes = fast_win_bytes(EPW, R, G);
    unsigned char* scratch = smem + tables_bytes(G, R, C) + warp * fast_warp_scratch_bytes(EPW, R, G, D);
    float* tile = reinterpret_cast<float*>(scratch + 2 * win_bytes);
    auto buf_twin = [&](int b) { return reinterpret_cast<uint64_t*>(scratch + b * win_bytes); };                      // [EPW][TWR]
    auto buf_vwin = [&](int b) { return reinterpret_cast<uint32_t*>(scratch + b * win_bytes + EPW * kTypeWinBytes); };   // [EPW][7][4]

    // Every warp owns one contiguous range of Q envs (Q a multiple of 4, the same for all warps, so
    // the work is balanced to within one 4-env group) and walks it in tiles of EPW; the range's
    // last tile may be shorter.
    const int gwarp = blockIdx.x * kFastWarps + warp, nwarps = gridDim.x * kFastWarps;
    const int nfull = p.N & ~3;                                       // envs in whole 4-env groups
    const int Q = (((nfull + nwarps - 1) / nwarps) + 3) & ~3;
    const int wbase = min(nfull, gwarp * Q), wend = min(nfull, wbase + Q);
    // ---- per-lane constants of phase B
    const int sub = lane & 15, half = lane >> 4, hbase = lane & 16;
    int srcl[R], shf[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        int dx = 0, dy = 0;
        if (sub < C) { dx = t.off[(sub * R + rr) * 2]; dy = t.off[(sub * R + rr) * 2 + 1]; }
        srcl[rr] = hbase + dx + R;     // lane holding window row x+dx
        // left shift that brings column y+dy of the window word (bits 2(dy+R), +1) to bits 30, 31
        shf[rr] = 30 - 2 * (dy + R);
    }
    const bool has_row = sub < NROW, has_ray = sub < C, has_vrow = sub < 5, has_v1 = sub < 9;
    // the two window cells this lane converts: q = sub and q = sub + 16 -> (row lane, nibble shift)
    const int vsrc0 = hbase + sub / 5, vsh0 = 4 * (sub % 5);
    const int vsrc1 = hbase + (sub + 16) / 5, vsh1 = 4 * ((sub + 16) % 5);
    const uint32_t s_dist = smem_u32(t.dist), s_pos = smem_u32(t.pos), s_visit = smem_u32(t.visit);
    const uint32_t s_onehot = smem_u32(t.onehot), s_rw32 = smem_u32(t.rw32), s_rw64 = smem_u32(t.rw64);
    constexpr uint64_t LOWPAD = kObstAll & ((1ull << (2 * R)) - 1ull);
    constexpr int NCH = 2;             // envs per half-warp and trip, interleaved for ILP
    const float4* src4 = reinterpret_cast<const float4*>(tile);

    // ---- pipeline helpers.  Records + actions are prefetched two tiles ahead, windows one tile
    // ahead, all with cp.async into shared memory: the prefetch holds no registers.
    auto rec_buf = [&](int b) { return reinterpret_cast<uint4*>(scratch + 2 * win_bytes + 16 * D + b * EPW * 40); };
    auto act_buf = [&](int b) { return reinterpret_cast<long long*>(scratch + 2 * win_bytes + 16 * D + b * EPW * 40 + EPW * 32); };
    auto fetch_rec = [&](int es, int b) {              // tile starting at env es
        if (lane < min(EPW, wend - es)) {
            const size_t e = (size_t)es + lane;
            cp_async16(smem_u32(rec_buf(b) + 2 * lane), p.rec + 2 * e);
            cp_async16(smem_u32(rec_buf(b) + 2 * lane + 1), p.rec + 2 * e + 1);
            cp_async8(smem_u32(act_buf(b) + lane), io.actions + e);
        }
        cp_async_commit();
    };
    // Window fetch of tile `tl` into buffer b.  LPE = 32/EPW lanes share one env: lane l serves
    // env l % EPW and the 16-byte chunks l / EPW, + LPE, ... of its windows (TCH chunks of type
    // rows, then 7 nibble rows).  The tile's records are already in rec_buf(b).
    constexpr int LPE = 32 / EPW, TCH = TWR / 2;
    auto issue_copies = [&](int es, int b) {
        const int j = lane % EPW, k0 = lane / EPW;
        const bool live = j < wend - es;                // the range's last tile may be short
        const int x = live ? (int)(rec_buf(b)[2 * j].x & 0xff) : 0;
        const size_t e = (size_t)es + j;
        // grid rows x-R-1 .. x+R+1 are padded rows x+1 .. x+2R+3; start on the even row at or
        // just below x+1 so that every chunk is 16-byte aligned
        const int r0 = (x + 1) & ~1;
        const uint64_t* tsrc = p.types + e * TS + r0;                       // TWR rows = TCH chunks
        const uint32_t* vsrc = p.vis4 + e * VE + (size_t)x * VW;            // padded nibble rows x .. x+6
        const uint32_t tdst = smem_u32(buf_twin(b) + j * TWR);
        const uint32_t vdst = smem_u32(buf_vwin(b) + j * kVisWinRows * VW);
#pragma unroll
        for (int k = k0, i = 0; i < (TCH + kVisWinRows + LPE - 1) / LPE; ++i, k += LPE) {
            if (!live) continue;
            if (k < TCH) cp_async16(tdst + 16 * k, tsrc + 2 * k);
            else if (k < TCH + kVisWinRows) cp_async16(vdst + 16 * (k - TCH), vsrc + 4 * (k - TCH));
        }
        cp_async_commit();
    };

    if (wbase < wend) {
        fetch_rec(wbase, 0);
        cp_async_wait_all();
        __syncwarp();
        issue_copies(wbase, 0);
        if (wbase + EPW < wend) fetch_rec(wbase + EPW, 1);
    }

    int it = 0;
    for (int e0 = wbase; e0 < wend; e0 += EPW, ++it) {
        const int b = it & 1;
        const int e_next = e0 + EPW;
        const bool has_next = e_next < wend;
        const int ts = min(EPW, wend - e0);               // envs in this tile (a multiple of 4)
        const bool act = lane < ts;
        uint64_t* twin = buf_twin(b);
        uint32_t* vwin = buf_vwin(b);
        const size_t e = (size_t)e0 + lane;
        cp_async_wait_all();                              // this tile's windows and the next tile's records
        __syncwarp();                                     // have landed, for every lane of the warp

        // ---- phase A: transition out of shared memory, one lane per env
        int done = 0, term = 0, trunc = 0;
        unsigned posw = 0;
        EnvRec r = {};
        if (act) {
            uint4 ra = rec_buf(b)[2 * lane], rb = rec_buf(b)[2 * lane + 1];
            const long long action = act_buf(b)[lane];
            r = unpack_rec(ra, rb);
            const int x0 = r.x, r0 = (r.x + 1) & ~1;      // the windows are centred on the pre-move row
            int tx, ty; bool inb;
            action_target(r, action, G, tx, ty, inb);
            // (tx, ty) is at most one cell away, so it is inside both windows even when it is
            // outside the grid (wall padding / border nibbles)
            uint64_t* tw = twin + lane * TWR + (tx + TP - r0);
            const uint64_t word = *tw;
            const int t_cell = inb ? cell_of(word, ty & 31) : kObstacle;
            uint32_t* vw = vwin + (lane * kVisWinRows + (tx - x0 + 3)) * VW + ((ty + 2) >> 3);
            const int sh = nib_shift(ty);
            const uint32_t vword = *vw;
            const StepOut o = transition_core(r, action, tx, ty, t_cell, (vword >> sh) & 15u, p.max_steps);
            if (o.moved)
                *vw = bump_visit(p.vis4 + e * VE + nib_word(tx, ty, VW), vword, sh,
                                 p.visov + e * G * G + tx * G + ty, mem);
            if (o.watered) {
                const uint64_t nw = word ^ (1ull << (2 * (ty & 31)));          // 3 -> 2
                *tw = nw;
                mem.st64(p.types + e * TS + TP + tx, nw);
            }
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
            // new position + where its windows start inside the fetched ones:
            //   type row of grid row x'-R is padded row x'+2, i.e. fetched row x'+2-r0   (0..3)
            //   nibble row of grid row x'-2 is padded row x'+1, i.e. fetched row x'+1-x0 (0..2)
            // packed as x | y << 5 | first type row (lane * TWR + x'+2-r0) << 10 | first nibble row
            // (lane * 7 + x'+1-x0) << 19, i.e. ready-made indices into the window buffers
            posw = (unsigned)r.x | ((unsigned)r.y << 5) | ((unsigned)(lane * TWR + r.x + 2 - r0) << 10) |
                   ((unsigned)(lane * kVisWinRows + r.x + 1 - x0) << 19);
        }
        accumulate_stats(p, act && done, r, term, trunc, lane);
        __syncwarp();   // window patches are visible to the half-warps below

        // the next tile's copies go out now and land while phase B runs
        if (has_next) issue_copies(e_next, b ^ 1);
        if (e_next + EPW < wend) fetch_rec(e_next + EPW, b);     // rec_buf(b) is free again

        // ---- phase B: observations.  Each trip builds four rows: two independent chains per
        // half-warp, written stage by stage (all shared-memory reads of both chains, then the
        // shuffles, then the table reads, then the stores) so that their latencies overlap.
        float4* const obs4 = reinterpret_cast<float4*>(io.obs) + (size_t)(e0 >> 2) * D;
#pragma unroll 1
        for (int base = 0; base < ts; base += 4) {
            int x[NCH], y[NCH], tb[NCH], vb[NCH];
            unsigned w[NCH], vslice[NCH], acc[NCH], s0[NCH], s1[NCH];
            uint64_t trow[NCH];
            unsigned vlo[NCH], vhi[NCH];
            // stage 1: positions of the two envs this half-warp handles (env base + 2c + half)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const unsigned pw = __shfl_sync(FULL, posw, base + 2 * c + half);
                x[c] = pw & 31; y[c] = (pw >> 5) & 31;
                tb[c] = (pw >> 10) & 511; vb[c] = pw >> 19;
            }
            // stage 2: shared-memory reads: this lane's type row and visit-nibble words
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                trow[c] = kObstAll;
                if (has_row) trow[c] = twin[tb[c] + sub];
                vlo[c] = 0; vhi[c] = 0;
                if (has_vrow) {
                    // the 5 nibbles y .. y+4 start in word y>>3 and may spill into the next one
                    // (when they sit entirely in word 3 the funnel's high half is unused)
                    const unsigned w0 = (unsigned)y[c] >> 3, w1 = w0 < 3u ? w0 + 1u : 3u;
                    const uint32_t* vr = vwin + (vb[c] + sub) * VW;
                    vlo[c] = vr[w0]; vhi[c] = vr[w1];
                }
            }
            // stage 3: rover-centred window word (cells y-R .. y+R of this lane's row, walls
            // outside) and visit slice (nibbles y .. y+4 of this lane's visit row)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int sft = 2 * y[c];
                const uint64_t ext = (trow[c] << (2 * R)) | LOWPAD;
                w[c] = (unsigned)((ext >> sft) | ((kObstAll << 1) << (63 - sft)));
                vslice[c] = __funnelshift_r(vlo[c], vhi[c], 4 * (y[c] & 7));
                acc[c] = 0;
            }
            // stage 4: LIDAR march (plantos_env.py:260-284): sample rr looks at window row
            // srcl[rr], bits shf[rr]; visit cells come from the row lanes' slices
#pragma unroll
            // (far sample first: each step shifts the accumulator left by one cell and funnels the
            // sample's two bits in from the top of the aligned row word -- two shifts per sample --
            // so that sample rr ends up at bits 2rr, 2rr+1)
#pragma unroll
            for (int rr = R - 1; rr >= 0; --rr) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const unsigned wr = __shfl_sync(FULL, w[c], srcl[rr]);
                    acc[c] = __funnelshift_l(wr << shf[rr], acc[c], 2);
                }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                s0[c] = __shfl_sync(FULL, vslice[c], vsrc0);
                s1[c] = __shfl_sync(FULL, vslice[c], vsrc1);
            }
            // stage 5: first hit per ray, then every table read of both chains
            float fd[NCH], fp[NCH], fv0[NCH], fv1[NCH];
            float4 oh[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const unsigned m = (acc[c] | (acc[c] >> 1)) & 0x55555555u;
                const int bit = __ffs(m) - 1;                 // -1 when nothing was hit
                const int dist = m ? (bit >> 1) + 1 : R;
                const int kind = m ? (acc[c] >> bit) & 3 : kEmpty;
                fd[c] = lds_f32(s_dist + 4 * dist);
                oh[c] = lds_f32x4(s_onehot + 16 * kind);
                fp[c] = lds_f32(s_pos + 4 * (sub ? y[c] : x[c]));
                fv0[c] = lds_f32(s_visit + 4 * ((s0[c] >> vsh0) & 15u));
                fv1[c] = lds_f32(s_visit + 4 * ((s1[c] >> vsh1) & 15u));
            }
            // stage 6: stores into the tile (row 2c + half)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float* row = tile + (2 * c + half) * D;
                if (has_ray) {                                // :286-292
                    float* q = row + 5 * sub;
                    q[0] = fd[c]; q[1] = oh[c].x; q[2] = oh[c].y; q[3] = oh[c].z; q[4] = oh[c].w;
                }
                if (sub < 2) row[5 * C + sub] = fp[c];        // :294-296
                row[5 * C + 2 + sub] = fv0[c];                // :298-313
                if (has_v1) row[5 * C + 18 + sub] = fv1[c];
            }
            // flush four env rows = D float4, 16-byte aligned because e0 and base are multiples of 4
            __syncwarp();
            float4* dst4 = obs4 + (size_t)(base >> 2) * D;
#pragma unroll
            for (int q = 0; q < (D + 31) / 32; ++q) {
                const int idx = q * 32 + lane;
                if (idx < D) __stcs(dst4 + idx, src4[idx]);
            }
            __syncwarp();
        }

        // ---- phase C: auto-reset of finished envs (rare; warp-cooperative generic code; the
        // current window buffer is free now and serves as the type-plane scratch)
        unsigned dmask = __ballot_sync(FULL, act && done);
        uint64_t* plane = buf_twin(b);
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

    // ragged tail: envs beyond the last full tile, one at a time (no copies are in flight here)
    if (gwarp == nwarps - 1)
        for (int e = nfull; e < p.N; ++e) step_env_warp(p, t, io, e, buf_twin(0), tile, lane);
}

}  // namespace plantos_dev
