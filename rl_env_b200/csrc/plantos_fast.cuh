// plantos_fast.cuh -- the sm_100a hot kernel for the reference presets.
//
// Requirements (checked on the host): W == 1 and VW == 4 (G <= 28), G + R <= 32, 2R+1 <= 16,
// C <= 16.  Both the training preset (G25 R6 C16, D=107; A2C_training.py:206-212) and the ctor
// default (G21 R2 C10, D=77; plantos_env.py:25-26) qualify; everything else runs
// k_step_generic.
//
// One warp owns a tile of EPW consecutive envs and goes through:
//   fetch    one LANE per env: load the 32-byte record and the action, then ask the TMA unit
//            for everything else the step can touch, centred on the PRE-move position with a
//            margin of one cell: 2R+4 rows of the wall-padded type plane (128 contiguous bytes
//            for R=6) and 7 rows of the visit-nibble plane (112 contiguous bytes), as two
//            cp.async.bulk global->shared copies per env completing on the warp's mbarrier.
//            All 2*EPW copies are in flight at once, cost no registers, and are the ONLY reads
//            of plane state the step performs: there is no dependent second round of loads.
//   phase A  still one lane per env, now entirely out of shared memory: the transition
//            (plantos_env.py:160-222) -- target-cell lookup, visit count, watering, reward,
//            termination.  The (at most two) modified words go back to global memory and are
//            patched in the shared copy; reward / done / record stores are coalesced.
//   phase B  one HALF-WARP per env, two envs per iteration -- the observation
//            (plantos_env.py:251-315) from the shared windows: 2R+1 lanes shift their type row
//            into a rover-centred window word (the padding makes bounds checks unnecessary);
//            one lane per ray marches the integer offset table with a warp shuffle as the row
//            lookup; five lanes cut the 20-bit slice of their visit row that the 5x5 window
//            needs and the 25 cell lanes read it by shuffle.  Rows are assembled in a 4-env
//            shared-memory tile whose 16*D bytes are 16-byte aligned in the [N, D] fp32
//            buffer and leave with streaming 128-bit stores (st.global.cs.v4, evict-first),
//            so the write-once observation stream does not evict the env state from L2.
//   phase C  whole warp, rare   -- SB3 auto-reset of finished envs (terminal observation,
//            Philox / injected map, fresh observation) via the generic warp routines.
// A ragged last tile (N % EPW != 0) is stepped env by env with step_env_warp.
#pragma once
#include <type_traits>
#include "plantos_generic.cuh"

namespace plantos_dev {

#ifndef PLANTOS_FAST_MINBLOCKS
#define PLANTOS_FAST_MINBLOCKS 4
#endif
#ifndef PLANTOS_FAST_CHAINS
#define PLANTOS_FAST_CHAINS 2        // independent observation chains per half-warp and trip
#endif
#ifndef PLANTOS_FAST_WARPS
#define PLANTOS_FAST_WARPS 7         // 7 warps x 4 blocks = 28 resident warps per SM
#endif
constexpr int kFastWarps = PLANTOS_FAST_WARPS;
constexpr int kVisWinRows = 7;       // nibble rows x-3 .. x+3 around the pre-move position
constexpr int kVisWinBytes = kVisWinRows * 16;
// type rows fetched per env: x-R-1 .. x+R+1 (2R+3 rows) plus one because the copy starts on an
// even row (16-byte aligned source and size)
__host__ __device__ constexpr int type_win_rows(int R) { return 2 * R + 4; }

// per-warp scratch: [type windows | visit windows | obs tile (4 envs)] + mbarrier.
// Phase C / the ragged tail reuse the window area as the generic code's type plane.
__host__ __device__ inline int fast_warp_scratch_bytes(int EPW, int R, int G, int D) {
    int win = EPW * (type_win_rows(R) * 8 + kVisWinBytes);
    if (win < align_up(G * 8, 16)) win = align_up(G * 8, 16);
    return win + 8 * PLANTOS_FAST_CHAINS * D + 16;
}

// ---- TMA (bulk async copy) + mbarrier helpers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}

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
    const int e0 = (blockIdx.x * kFastWarps + warp) * EPW;
    if (e0 >= p.N) return;
    unsigned char* scratch = smem + tables_bytes(G, R, C) + warp * fast_warp_scratch_bytes(EPW, R, G, D);
    constexpr int kWinBytes = EPW * (kTypeWinBytes + kVisWinBytes);
    const int win_bytes = kWinBytes < align_up(G * 8, 16) ? align_up(G * 8, 16) : kWinBytes;
    uint64_t* twin = reinterpret_cast<uint64_t*>(scratch);                           // [EPW][TWR]
    uint32_t* vwin = reinterpret_cast<uint32_t*>(scratch + EPW * kTypeWinBytes);     // [EPW][7][4]
    float* tile = reinterpret_cast<float*>(scratch + win_bytes);
    uint64_t* plane = reinterpret_cast<uint64_t*>(scratch);     // generic-path scratch (phase C, tail)
    const uint32_t mbar = smem_u32(scratch + win_bytes + 8 * PLANTOS_FAST_CHAINS * D);

    if (p.N - e0 < EPW) {   // ragged last tile
        for (int e = e0; e < p.N; ++e) step_env_warp(p, t, io, e, plane, tile, lane);
        return;
    }
    if (lane == 0) {
        mbar_init(mbar, 1);
        mbar_arrive_expect_tx(mbar, EPW * (kTypeWinBytes + kVisWinBytes));
    }
    __syncwarp();

    // ---- fetch: record + action, then the two windows around the pre-move position
    const bool act = lane < EPW;
    const int e = e0 + lane;
    uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
    long long action = 0;
    EnvRec r = {};
    int r0 = 0;                           // first padded type row of this env's window
    if (act) {
        ra = mem.ld128(p.rec + 2 * (size_t)e);
        rb = mem.ld128(p.rec + 2 * (size_t)e + 1);
        action = __ldcs(io.actions + e);
        r = unpack_rec(ra, rb);
        // grid rows x-R-1 .. x+R+1 are padded rows x+1 .. x+2R+3; start on the even row at or
        // just below x+1 so that source address and size are 16-byte multiples
        r0 = (r.x + 1) & ~1;
        if (!(p.dbg & 32)) {
        bulk_load(smem_u32(twin + lane * TWR), p.types + (size_t)e * TS + r0, kTypeWinBytes, mbar);
        // grid rows x-3 .. x+3 are padded nibble rows x .. x+6
        bulk_load(smem_u32(vwin + lane * kVisWinRows * VW), p.vis4 + (size_t)e * VE + (size_t)r.x * VW,
                  kVisWinBytes, mbar);
        }
    }

    // per-lane constants of phase B (computed while the copies are in flight)
    const int sub = lane & 15, half = lane >> 4, hbase = lane & 16;
    int srcl[R], shf[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        int dx = 0, dy = 0;
        if (sub < C) { dx = t.off[(sub * R + rr) * 2]; dy = t.off[(sub * R + rr) * 2 + 1]; }
        srcl[rr] = hbase + dx + R;     // lane holding window row x+dx
        shf[rr] = 2 * (dy + R);        // bit offset of column y+dy inside the window word
    }
    const bool has_row = sub < NROW, has_ray = sub < C, has_vrow = sub < 5, has_v1 = sub < 9;
    // the two window cells this lane converts: q = sub and q = sub + 16 -> (row lane, nibble shift)
    const int vsrc0 = hbase + sub / 5, vsh0 = 4 * (sub % 5);
    const int vsrc1 = hbase + (sub + 16) / 5, vsh1 = 4 * ((sub + 16) % 5);
    const float4* onehot = reinterpret_cast<const float4*>(t.onehot);
    constexpr uint64_t LOWPAD = kObstAll & ((1ull << (2 * R)) - 1ull);

    // wait for the windows (bounded spin: a lost completion must trap, not hang the GPU)
    {
        uint32_t spins = 0;
        while (!(p.dbg & 32) && !mbar_try_wait(mbar, 0)) {
            if (++spins > (1u << 24)) __trap();
        }
    }

    // ---- phase A: transition out of shared memory, one lane per env
    int done = 0, term = 0, trunc = 0;
    unsigned posw = 0;
    if (act) {
        const int x0 = r.x;               // pre-move row: the windows are centred on it
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
        if (o.moved && !(p.dbg & 16))
            *vw = bump_visit(p.vis4 + (size_t)e * VE + nib_word(tx, ty, VW), vword, sh,
                             p.visov + (size_t)e * G * G + tx * G + ty, mem);
        if (o.watered && !(p.dbg & 16)) {
            const uint64_t nw = word ^ (1ull << (2 * (ty & 31)));          // 3 -> 2
            *tw = nw;
            mem.st64(p.types + (size_t)e * TS + TP + tx, nw);
        }
        r.ret += t.rw64[o.ridx];
        term = o.terminated; trunc = o.truncated; done = term | trunc;
        pack_rec(r, ra, rb);
        if (!(p.dbg & 16)) {
        io.reward[e] = t.rw32[o.ridx];
        io.done[e] = (uint8_t)done;
        if (io.terminated) io.terminated[e] = (uint8_t)term;
        if (io.truncated) io.truncated[e] = (uint8_t)trunc;
        mem.st128(p.rec + 2 * (size_t)e, ra);
        mem.st128(p.rec + 2 * (size_t)e + 1, rb);
        }
        if (done) {
            p.term_rec[2 * (size_t)e] = ra;
            p.term_rec[2 * (size_t)e + 1] = rb;
        }
        // new position + where its windows start inside the fetched ones:
        //   type row of grid row x'-R is padded row x'+2, i.e. fetched row x'+2-r0   (0..3)
        //   nibble row of grid row x'-2 is padded row x'+1, i.e. fetched row x'+1-x0 (0..2)
        posw = (unsigned)r.x | ((unsigned)r.y << 8) | ((unsigned)(r.x + 2 - r0) << 16) |
               ((unsigned)(r.x + 1 - x0) << 20);
    }
    accumulate_stats(p, act && done, r, term, trunc, lane);
    __syncwarp();   // window patches are visible to the half-warps below

    // ---- phase B: observations, half-warp per env
    // Each trip builds 2*NCH observation rows: NCH independent chains per half-warp, written
    // stage by stage (all shared-memory reads of all chains, then the shuffles, then the table
    // reads, then the stores) so that the chains' latencies overlap instead of adding up --
    // the SM runs only ~7 warps per scheduler here, so the parallelism has to come from ILP.
    // chains per half-warp and trip (falls back to 2 when the tile is too small for more)
    constexpr int NCH = (EPW % (2 * PLANTOS_FAST_CHAINS) == 0) ? PLANTOS_FAST_CHAINS : 2;
    constexpr int ROWS = 2 * NCH;                 // env rows per trip (a multiple of 4)
    static_assert(ROWS % 4 == 0 && EPW % ROWS == 0, "a trip must be whole 4-env groups");
    float4* const obs4 = reinterpret_cast<float4*>(io.obs) + (size_t)(e0 >> 2) * D;
    const float4* src4 = reinterpret_cast<const float4*>(tile);
#pragma unroll 1
    for (int base = 0; base < ((p.dbg & 8) ? 0 : EPW); base += ROWS) {
        if (!(p.dbg & 64)) {
        int x[NCH], y[NCH], tb[NCH], vb[NCH];
        unsigned w[NCH], vslice[NCH], acc[NCH], s0[NCH], s1[NCH];
        uint64_t trow[NCH];
        unsigned vlo[NCH], vhi[NCH];
        // stage 1: positions of the NCH envs this half-warp handles (env base + 2c + half)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const unsigned pw = __shfl_sync(FULL, posw, base + 2 * c + half);
            x[c] = pw & 0xff; y[c] = (pw >> 8) & 0xff;
            tb[c] = (pw >> 16) & 15; vb[c] = pw >> 20;
        }
        // stage 2: shared-memory reads: this lane's type row and visit-nibble words
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int j = base + 2 * c + half;
            trow[c] = kObstAll;
            if (has_row) trow[c] = twin[j * TWR + tb[c] + sub];
            vlo[c] = 0; vhi[c] = 0;
            if (has_vrow) {
                // the 5 nibbles y .. y+4 start in word y>>3 and may spill into the next one
                // (when they sit entirely in word 3 the funnel's high half is unused)
                const unsigned w0 = (unsigned)y[c] >> 3, w1 = w0 < 3u ? w0 + 1u : 3u;
                const uint32_t* vr = vwin + (j * kVisWinRows + vb[c] + sub) * VW;
                vlo[c] = vr[w0]; vhi[c] = vr[w1];
            }
        }
        // stage 3: rover-centred window word (cells y-R .. y+R of this lane's row, walls
        // outside) and visit slice (nibbles y .. y+4 of this lane's visit row)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int s = 2 * y[c];
            const uint64_t ext = (trow[c] << (2 * R)) | LOWPAD;
            w[c] = (unsigned)((ext >> s) | ((kObstAll << 1) << (63 - s)));
            vslice[c] = __funnelshift_r(vlo[c], vhi[c], 4 * (y[c] & 7));
            acc[c] = 0;
        }
        // stage 4: LIDAR march (plantos_env.py:260-284): sample rr looks at window row
        // srcl[rr], bits shf[rr]; visit cells come from the row lanes' slices
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const unsigned wr = __shfl_sync(FULL, w[c], srcl[rr]);
                acc[c] += ((wr >> shf[rr]) & 3u) << (2 * rr);
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            s0[c] = __shfl_sync(FULL, vslice[c], vsrc0);
            s1[c] = __shfl_sync(FULL, vslice[c], vsrc1);
        }
        // stage 5: first hit per ray, then every table read of every chain
        float fd[NCH], fp[NCH], fv0[NCH], fv1[NCH];
        float4 oh[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const unsigned m = (acc[c] | (acc[c] >> 1)) & 0x55555555u;
            const int b = __ffs(m) - 1;                   // -1 when nothing was hit
            const int dist = m ? (b >> 1) + 1 : R;
            const int kind = m ? (acc[c] >> b) & 3 : kEmpty;
            fd[c] = t.dist[dist];
            oh[c] = onehot[kind];
            fp[c] = t.pos[sub ? y[c] : x[c]];
            fv0[c] = t.visit[(s0[c] >> vsh0) & 15u];
            fv1[c] = t.visit[(s1[c] >> vsh1) & 15u];
        }
        // stage 6: stores into the tile (row 2c + half of this trip)
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
        }
        // flush ROWS env rows = ROWS/4 * D float4, 16-byte aligned because e0 and base are
        // multiples of 4
        __syncwarp();
        float4* dst4 = obs4 + (size_t)(base >> 2) * D;
        constexpr int NF4 = (ROWS / 4) * D;
#pragma unroll
        for (int k = 0; k < (NF4 + 31) / 32; ++k) {
            const int idx = k * 32 + lane;
            if (idx < NF4 && !(p.dbg & 1)) __stcs(dst4 + idx, src4[idx]);
        }
        __syncwarp();
    }

    // ---- phase C: auto-reset of finished envs (rare; warp-cooperative generic code)
    unsigned dmask = __ballot_sync(FULL, act && done);
    while (dmask) {
        const int j = __ffs(dmask) - 1;
        dmask &= dmask - 1;
        const int ej = e0 + j;
        const int episode = __shfl_sync(FULL, r.episode, j);
        const int px = __shfl_sync(FULL, r.x, j), py = __shfl_sync(FULL, r.y, j);
        const uint64_t* types_e = p.types + (size_t)ej * TS + TP;
        const uint32_t* vis_e = p.vis4 + (size_t)ej * VE;
        if (io.terminal_obs) {
            for (int idx = lane; idx < G; idx += 32) plane[idx] = types_e[idx];
            __syncwarp();
            build_obs_warp(p, t, plane, vis_e, px, py, tile, lane);
            store_obs_row(tile, io.terminal_obs + (size_t)ej * D, D, lane);
            __syncwarp();
        }
        const EnvRec nr = reset_env_warp(p, ej, episode, plane, lane);
        build_obs_warp(p, t, plane, vis_e, nr.x, nr.y, tile, lane);
        store_obs_row(tile, io.obs + (size_t)ej * D, D, lane);
        if (lane == 0) {
            uint4 qa, qb;
            pack_rec(nr, qa, qb);
            p.rec[2 * (size_t)ej] = qa;
            p.rec[2 * (size_t)ej + 1] = qb;
        }
        __syncwarp();
    }
}

}  // namespace plantos_dev
