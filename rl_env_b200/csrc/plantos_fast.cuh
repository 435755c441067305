// plantos_fast.cuh -- the sm_100a hot kernel for the reference presets.
//
// Requirements (checked on the host): W == 1 and VW == 4 (G <= 28), G + R <= 32, 2R+1 <= 16,
// C <= 16.  Both the training preset (G25 R6 C16, D=107; A2C_training.py:206-212) and the ctor
// default (G21 R2 C10, D=77; plantos_env.py:25-26) qualify; everything else runs
// k_step_generic.
//
// One warp owns a tile of EPW consecutive envs and goes through:
//   phase A  one LANE per env   -- the scalar transition (plantos_env.py:160-222): record
//            load, action, target-cell lookup, visit-nibble read-modify-write, watering,
//            reward / done / record stores.  Outputs are coalesced across the tile.
//   fetch    still one lane per env: as soon as a lane knows its env's new position it asks
//            the TMA unit for the two pieces of state the observation needs -- 2R+2 rows of the
//            wall-padded type plane (112 contiguous bytes for R=6, covering x-R .. x+R) and the five
//            16-byte visit-nibble rows of the 5x5 window (80 contiguous bytes) -- with two
//            cp.async.bulk global->shared copies that complete on the warp's mbarrier.  All
//            2*EPW copies of the tile are in flight at once and cost no registers.
//   phase B  one HALF-WARP per env, two envs per iteration -- the observation
//            (plantos_env.py:251-315) from shared memory: 2R+1 lanes shift their type row into
//            a rover-centred window word (the padding makes bounds checks unnecessary); one
//            lane per ray marches the integer offset table with a warp shuffle as the row
//            lookup; five lanes cut the 20-bit slice of their visit row that the window
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
constexpr int kFastWarps = 7;        // 7 warps x 4 blocks = 28 resident warps per SM
constexpr int kVisWinBytes = 5 * 16; // five nibble rows
// rows per env fetched from the type plane: the 2R+1 window rows plus one, rounded up to even,
// because the copy starts on an even row (16-byte aligned source and size)
__host__ __device__ constexpr int type_win_rows(int R) { return (2 * R + 3) & ~1; }

// per-warp scratch: [type windows | visit windows | obs tile (4 envs)] + mbarrier.
// Phase C / the ragged tail reuse the window area as the generic code's type plane.
__host__ __device__ inline int fast_warp_scratch_bytes(int EPW, int R, int G, int D) {
    int win = EPW * (type_win_rows(R) * 8 + kVisWinBytes);
    if (win < align_up(G * 8, 16)) win = align_up(G * 8, 16);
    return win + 16 * D + 16;
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
// order this thread's earlier global stores before its later async-proxy (TMA) reads
__device__ __forceinline__ void fence_global_to_async() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
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
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kTypeWinRows = type_win_rows(R);
    constexpr int kTypeWinBytes = kTypeWinRows * 8;
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
    const uint64_t* twin = reinterpret_cast<const uint64_t*>(scratch);                        // [EPW][14]
    const uint32_t* vwin = reinterpret_cast<const uint32_t*>(scratch + EPW * kTypeWinBytes);  // [EPW][5][4]
    float* tile = reinterpret_cast<float*>(scratch + win_bytes);
    uint64_t* plane = reinterpret_cast<uint64_t*>(scratch);     // generic-path scratch (phase C, tail)
    const uint32_t mbar = smem_u32(scratch + win_bytes + 16 * D);

    if (p.N - e0 < EPW) {   // ragged last tile
        for (int e = e0; e < p.N; ++e) step_env_warp(p, t, io, e, plane, tile, lane);
        return;
    }
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();

    // ---- phase A: transition, one lane per env
    const bool act = lane < EPW;
    EnvRec r = {};
    int done = 0, term = 0, trunc = 0;
    unsigned posw = 0;
    if (act) {
        const int e = e0 + lane;
        uint4 ra = mem.ld128(p.rec + 2 * (size_t)e), rb = mem.ld128(p.rec + 2 * (size_t)e + 1);
        const long long action = __ldcs(io.actions + e);
        r = unpack_rec(ra, rb);
        int tx, ty; bool inb;
        action_target(r, action, G, tx, ty, inb);
        uint64_t* wp = p.types + (size_t)e * TS + R + (inb ? tx : r.x);
        const uint64_t word = inb ? mem.ld64(wp) : kObstAll;
        const StepOut o = apply_action(r, action, tx, ty, inb, word, wp, p.vis4 + (size_t)e * VE,
                                       p.visov + (size_t)e * G * G, G, VW, p.max_steps, mem);
        r.ret += t.rw64[o.ridx];
        io.reward[e] = t.rw32[o.ridx];
        term = o.terminated; trunc = o.truncated; done = term | trunc;
        io.done[e] = (uint8_t)done;
        if (io.terminated) io.terminated[e] = (uint8_t)term;
        if (io.truncated) io.truncated[e] = (uint8_t)trunc;
        pack_rec(r, ra, rb);
        mem.st128(p.rec + 2 * (size_t)e, ra);
        mem.st128(p.rec + 2 * (size_t)e + 1, rb);
        if (done) {
            p.term_rec[2 * (size_t)e] = ra;
            p.term_rec[2 * (size_t)e + 1] = rb;
        }
        posw = (unsigned)r.x | ((unsigned)r.y << 8);
    }

    // ---- fetch: every lane asks the TMA unit for its env's two windows
    if (lane == 0) mbar_arrive_expect_tx(mbar, EPW * (kTypeWinBytes + kVisWinBytes));
    __syncwarp();
    if (act) {
        const int e = e0 + lane;
        fence_global_to_async();     // the nibble / type words this lane just wrote
        // padded type rows x .. x+2R hold grid rows x-R .. x+R; start on the even row below so
        // that source and size are 16-byte multiples (the plane has one spare row for this)
        const uint64_t* tsrc = p.types + (size_t)e * TS + (r.x & ~1);
        bulk_load(smem_u32(twin + lane * kTypeWinRows), tsrc, kTypeWinBytes, mbar);
        // padded nibble rows x .. x+4 hold grid rows x-2 .. x+2
        const uint32_t* vsrc = p.vis4 + (size_t)e * VE + (size_t)r.x * VW;
        bulk_load(smem_u32(vwin + lane * 5 * VW), vsrc, kVisWinBytes, mbar);
    }
    accumulate_stats(p, act && done, r, term, trunc, lane);

    // ---- phase B: observations, half-warp per env
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
        while (!mbar_try_wait(mbar, 0)) {
            if (++spins > (1u << 24)) __trap();
        }
    }

    auto compute = [&](int j, float* row) {
        const unsigned pw = __shfl_sync(FULL, posw, j);
        const int x = pw & 0xff, y = pw >> 8;
        // rover-centred window word: cells y-R .. y+R of this lane's row, walls outside
        uint64_t trow = kObstAll;
        if (has_row) trow = twin[j * kTypeWinRows + (x & 1) + sub];
        const int s = 2 * y;
        const uint64_t ext = (trow << (2 * R)) | LOWPAD;
        const unsigned w = (unsigned)((ext >> s) | ((kObstAll << 1) << (63 - s)));
        // this lane's visit row: the 5 nibbles y .. y+4 start in word y>>3 and may spill into
        // the next one (when they sit entirely in word 3 the funnel's high half is unused)
        unsigned vslice = 0;
        if (has_vrow) {
            const unsigned w0 = y >> 3, w1 = w0 < 3u ? w0 + 1u : 3u;
            const uint32_t* vr = vwin + (j * 5 + sub) * VW;
            vslice = __funnelshift_r(vr[w0], vr[w1], 4 * (y & 7));
        }
        // LIDAR march (plantos_env.py:260-284): sample rr looks at window row srcl[rr], bits shf[rr]
        unsigned acc = 0;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const unsigned wr = __shfl_sync(FULL, w, srcl[rr]);
            acc += ((wr >> shf[rr]) & 3u) << (2 * rr);
        }
        const unsigned s0 = __shfl_sync(FULL, vslice, vsrc0);
        const unsigned s1 = __shfl_sync(FULL, vslice, vsrc1);
        const unsigned m = (acc | (acc >> 1)) & 0x55555555u;
        const int b = __ffs(m) - 1;                       // -1 when nothing was hit
        const int dist = m ? (b >> 1) + 1 : R;
        const int kind = m ? (acc >> b) & 3 : kEmpty;
        if (has_ray) {                                    // :286-292
            float* q = row + 5 * sub;
            const float4 oh = onehot[kind];
            q[0] = t.dist[dist];
            q[1] = oh.x; q[2] = oh.y; q[3] = oh.z; q[4] = oh.w;
        }
        if (sub < 2) row[5 * C + sub] = t.pos[sub ? y : x];                       // :294-296
        row[5 * C + 2 + sub] = t.visit[(s0 >> vsh0) & 15u];                       // :298-313
        if (has_v1) row[5 * C + 18 + sub] = t.visit[(s1 >> vsh1) & 15u];
    };

    float4* const obs4 = reinterpret_cast<float4*>(io.obs) + (size_t)(e0 >> 2) * D;
    const float4* src4 = reinterpret_cast<const float4*>(tile);
    float* const rowA = tile + half * D;          // env 4g + half
    float* const rowB = tile + (2 + half) * D;    // env 4g + 2 + half
#pragma unroll 1
    for (int g = 0; g < EPW / 4; ++g) {
        compute(4 * g + half, rowA);
        compute(4 * g + 2 + half, rowB);
        // flush four env rows = D float4, 16-byte aligned because e0 and 4g are multiples of 4
        __syncwarp();
        float4* dst4 = obs4 + (size_t)g * D;
#pragma unroll
        for (int k = 0; k < (D + 31) / 32; ++k) {
            const int idx = k * 32 + lane;
            if (idx < D && !(p.dbg & 1)) __stcs(dst4 + idx, src4[idx]);
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
        const uint64_t* types_e = p.types + (size_t)ej * TS + R;
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
            uint4 ra, rb;
            pack_rec(nr, ra, rb);
            p.rec[2 * (size_t)ej] = ra;
            p.rec[2 * (size_t)ej + 1] = rb;
        }
        __syncwarp();
    }
}

}  // namespace plantos_dev
