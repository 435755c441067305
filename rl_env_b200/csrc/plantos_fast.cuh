// plantos_fast.cuh -- the sm_100a hot kernel for the reference presets.
//
// Requirements (checked on the host): W == 1 and VW == 4 (G <= 28), G + R <= 32, 2R+1 <= 16,
// C <= 16.  Both the training preset (G25 R6 C16, D=107; A2C_training.py:206-212) and the ctor
// default (G21 R2 C10, D=77; plantos_env.py:25-26) qualify; everything else runs
// k_step_generic.
//
// Work split inside one warp, which owns a tile of EPW consecutive envs:
//   phase A  one LANE per env   -- the scalar transition (plantos_env.py:160-222): record
//            load, action, target-cell lookup, visit-nibble read-modify-write, watering,
//            reward / done / record stores.  Outputs are coalesced across the tile.
//   phase B  one HALF-WARP per env, two envs per iteration -- the observation
//            (plantos_env.py:251-315): 2R+1 lanes each fetch one 8-byte row of the wall-padded
//            type plane and shift it into a rover-centred window word (no bounds checks
//            anywhere); one lane per ray marches the integer offset table, with a warp
//            shuffle as the row lookup; five lanes fetch the five 16-byte visit-nibble rows
//            of the 5x5 window (80 contiguous bytes) and cut the 20-bit slice the window
//            needs, which the 25 cell lanes read by shuffle.  Rows are assembled in a 4-env
//            shared-memory tile whose 16*D bytes are 16-byte aligned in the [N, D] fp32
//            buffer and leave with streaming 128-bit stores (st.global.cs.v4, evict-first)
//            so that the write-once observation stream does not evict the env state from L2.
//            The loop is unrolled by two with ping-pong prefetch registers: the loads of
//            iteration i+1 are in flight during the arithmetic of iteration i.
//   phase C  whole warp, rare   -- SB3 auto-reset of finished envs (terminal observation,
//            Philox / injected map, fresh observation) via the generic warp routines.
// A ragged last tile (N % EPW != 0) is stepped env by env with step_env_warp.
#pragma once
#include <type_traits>
#include "plantos_generic.cuh"

namespace plantos_dev {

#ifndef PLANTOS_FAST_MINBLOCKS
#define PLANTOS_FAST_MINBLOCKS 8
#endif
constexpr int kFastWarps = 4;

__host__ __device__ inline int fast_warp_scratch_bytes(int G, int D) {
    return 2 * 16 * D + align_up(G * 8, 16);   // two 4-env obs tiles + type plane for phase C / tail
}

// ---- TMA (bulk async copy) helpers: shared -> global stores of finished observation tiles
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// make this thread's shared-memory writes visible to the async proxy (the TMA unit)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(ssrc), "r"(bytes), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {     // <= N groups still reading smem
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct Prefetch {
    uint64_t row;       // this lane's type row of the env's window
    unsigned vlo, vhi;  // the two nibble words of this lane's visit row that hold the window
    unsigned pw;        // x | y << 8 of the env
};

template <int R, int C, int EPW, bool KEEP>
__global__ void __launch_bounds__(kFastWarps * 32, PLANTOS_FAST_MINBLOCKS)
k_step_fast(const Params p, const StepIO io) {
    constexpr int D = 5 * C + 27;
    constexpr int NROW = 2 * R + 1;
    constexpr int VW = 4;                 // nibble words per visit row (G + 4 <= 32)
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NROW <= 16 && C <= 16, "fast kernel shape limits");
    static_assert(EPW % 4 == 0 && EPW <= 32, "tile must be whole 4-env groups");

    extern __shared__ __align__(16) unsigned char smem[];
    const Tables t = load_tables(p, smem);
    typename std::conditional<KEEP, KeepMem, PlainMem>::type const mem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = p.G, VE = p.VE, TS = p.TS;
    const int e0 = (blockIdx.x * kFastWarps + warp) * EPW;
    if (e0 >= p.N) return;
    unsigned char* scratch = smem + tables_bytes(G, R, C) + warp * fast_warp_scratch_bytes(G, D);
    float* tile = reinterpret_cast<float*>(scratch);
    uint64_t* plane = reinterpret_cast<uint64_t*>(scratch + 2 * 16 * D);

    if (p.N - e0 < EPW) {   // ragged last tile
        for (int e = e0; e < p.N; ++e) step_env_warp(p, t, io, e, plane, tile, lane);
        return;
    }

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
    accumulate_stats(p, act && done, r, term, trunc, lane);
    __syncwarp();   // phase A's plane updates are visible to the other lanes' loads below

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
    // lane-constant bases (byte pointers + unsigned 32-bit byte offsets keep the per-iteration
    // address math short): padded type row (x - R + sub) + R = x + sub; visit-nibble row
    // (x - 2 + sub) + 2 = x + sub for sub < 5, 16 bytes each
    const char* trow = reinterpret_cast<const char*>(p.types + (size_t)e0 * TS + sub);
    const char* vrow = reinterpret_cast<const char*>(p.vis4 + (size_t)e0 * VE + (size_t)sub * VW);
    const unsigned TS8 = 8u * TS, VE4 = 4u * VE;
    const bool has_row = sub < NROW, has_ray = sub < C, has_vrow = sub < 5, has_v1 = sub < 9;
    // the two window cells this lane converts: q = sub and q = sub + 16 -> (row lane, nibble shift)
    const int vsrc0 = hbase + sub / 5, vsh0 = 4 * (sub % 5);
    const int vsrc1 = hbase + (sub + 16) / 5, vsh1 = 4 * ((sub + 16) % 5);
    // observation tiles: group g (4 envs) is assembled in buffer g & 1; env 2*it + half of an
    // even iteration goes to row `half`, of an odd iteration to row 2 + half
    const float4* onehot = reinterpret_cast<const float4*>(t.onehot);
    constexpr uint64_t LOWPAD = kObstAll & ((1ull << (2 * R)) - 1ull);

    auto issue = [&](int it, Prefetch& n) {
        const int j = 2 * it + half;
        n.pw = __shfl_sync(FULL, posw, j);
        const unsigned x = n.pw & 0xff, y = n.pw >> 8;
        n.row = kObstAll;
        if (has_row && !(p.dbg & 4))
            n.row = mem.ld64(reinterpret_cast<const uint64_t*>(trow + ((unsigned)j * TS8 + 8u * x)));
        n.vlo = 0; n.vhi = 0;
        if (has_vrow && !(p.dbg & 2)) {
            // window nibbles y .. y+4 of this row start in word y>>3 and may spill into the next
            const unsigned w0 = y >> 3, w1 = w0 < 3u ? w0 + 1u : 3u;
            const char* base = vrow + ((unsigned)j * VE4 + 16u * x);
            n.vlo = mem.ld32(reinterpret_cast<const uint32_t*>(base + 4u * w0));
            n.vhi = mem.ld32(reinterpret_cast<const uint32_t*>(base + 4u * w1));
        }
    };

    auto compute = [&](const Prefetch& c, float* row) {
        const int x = c.pw & 0xff, y = c.pw >> 8;
        // rover-centred window word: cells y-R .. y+R of this lane's row, walls outside
        const int s = 2 * y;
        const uint64_t ext = (c.row << (2 * R)) | LOWPAD;
        const unsigned w = (unsigned)((ext >> s) | ((kObstAll << 1) << (63 - s)));
        // this lane's visit row: the 5 nibbles y .. y+4 (20 bits; when they sit entirely in
        // word 3 the funnel's high half is unused)
        const unsigned vslice = __funnelshift_r(c.vlo, c.vhi, 4 * (y & 7));
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
    const uint64_t pol_stream = policy_evict_first();
    const uint32_t tile_s = smem_u32(tile);
    // A finished group leaves through the TMA unit: one bulk shared->global copy of 16*D bytes
    // (16-byte aligned because e0 and the group start are multiples of 4), tagged evict-first.
    // The copy is asynchronous; the buffer is reused two groups later, after wait_group.read.
    auto flush = [&](int group) {
        fence_smem_to_async();
        __syncwarp();
        if (lane == 0 && !(p.dbg & 1))
            bulk_store(obs4 + (size_t)group * D, tile_s + (group & 1) * 16 * D, 16 * D, pol_stream);
    };

    constexpr int NIT = EPW / 2;
    Prefetch pa, pb;
    issue(0, pa);
#pragma unroll 1
    for (int it = 0; it < NIT; it += 2) {
        const int group = it >> 1;
        float* const tb = tile + (group & 1) * 4 * D;
        issue(it + 1, pb);
        if (group >= 2) {                 // buffer last used by group - 2: its bulk read must be done
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
        }
        compute(pa, tb + half * D);
        if (it + 2 < NIT) issue(it + 2, pa);
        compute(pb, tb + (2 + half) * D);
        flush(group);
    }
    if (lane == 0) bulk_wait_all();       // stores complete before phase C may overwrite rows / exit
    __syncwarp();

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
