// plantos_fast.cuh -- the sm_100a hot kernel for the reference presets.
//
// Requirements (checked on the host): W == 1 and VW == 4 (G <= 28), G + R <= 32, 2R+1 <= 16,
// C <= 16.  Both the training preset (G25 R6 C16, D=107; A2C_training.py:206-212) and the ctor
// default (G21 R2 C10, D=77; plantos_env.py:25-26) qualify; everything else runs
// k_step_generic.
//
// A persistent, warp-specialised pipeline.  Envs are processed in STAGES of 32; a block owns the
// stages blockIdx.x, blockIdx.x + gridDim.x, ... and runs them through a ring of kStages
// shared-memory slots guarded by mbarriers:
//
//   producer warp (warp 0), one LANE per env of the stage
//     fetch    loads the 32-byte record and the action (two stages ahead, in registers), then
//              asks the TMA unit for everything else the step can touch, centred on the
//              PRE-move position with a margin of one cell: 2R+4 rows of the wall-padded type
//              plane (128 contiguous bytes for R=6) and 7 rows of the visit-nibble plane (112
//              contiguous bytes) -- two cp.async.bulk global->shared copies per env that
//              complete on the slot's `tma` mbarrier (one stage ahead).  These are the ONLY
//              reads of plane state; there is no dependent second round of loads.
//     phase A  the transition (plantos_env.py:160-222) out of the slot: target-cell lookup,
//              visit count, watering, reward, termination.  The (at most two) modified words
//              go back to global memory and are patched in the slot; reward / done / record
//              stores are coalesced.  Then the slot is published on its `full` mbarrier.
//   consumer warps (warps 1..8), warp c owns envs 4c .. 4c+3 of every stage
//     phase B  the observation (plantos_env.py:251-315), one HALF-WARP per env, two envs per
//              half-warp interleaved stage by stage for ILP: 2R+1 lanes shift their type row
//              into a rover-centred window word (the padding makes bounds checks unnecessary);
//              one lane per ray marches the integer offset table with a warp shuffle as the row
//              lookup; five lanes cut the 20-bit slice of their visit row that the 5x5 window
//              needs and the 25 cell lanes read it by shuffle.  The slot is released on its
//              `empty` mbarrier, the four rows (16*D bytes, 16-byte aligned in the [N, D] fp32
//              buffer) leave with streaming 128-bit stores (st.global.cs.v4, evict-first) so
//              the write-once observation stream does not evict the env state from L2.
//     phase C  rare: SB3 auto-reset of the warp's finished envs (terminal observation,
//              Philox / injected map, fresh observation) via the generic warp routines.
//
// While the consumers compute and store stage k, the producer's copies for stage k+1 and its
// record loads for stage k+2 are in flight: memory latency, arithmetic and the observation
// write stream overlap instead of adding up.  Envs beyond the last full stage (N % 32) are
// stepped one by one with step_env_warp by the last block's consumer warp 0.
#pragma once
#include <type_traits>
#include "plantos_generic.cuh"

namespace plantos_dev {

#ifndef PLANTOS_FAST_STAGES
#define PLANTOS_FAST_STAGES 3
#endif
#ifndef PLANTOS_FAST_BLOCKS_PER_SM
#define PLANTOS_FAST_BLOCKS_PER_SM 3
#endif
constexpr int kStages = PLANTOS_FAST_STAGES;
constexpr int kStageEnvs = 32;
constexpr int kConsumers = 8;                       // consumer warps: 4 envs each per stage
constexpr int kFastThreads = (1 + kConsumers) * 32;
constexpr int kVisWinRows = 7;                      // nibble rows x-3 .. x+3 around the pre-move position
constexpr int kVisWinBytes = kVisWinRows * 16;
// type rows fetched per env: x-R-1 .. x+R+1 (2R+3 rows) plus one because the copy starts on an
// even row (16-byte aligned source and size)
__host__ __device__ constexpr int type_win_rows(int R) { return 2 * R + 4; }

// shared-memory layout after the tables: [3*kStages mbarriers | kStages slots | kConsumers tiles]
//   slot: type windows [32][TWR] u64 | visit windows [32][7][4] u32 | posw [32] u32 | aux [32] u32
__host__ __device__ inline int fast_slot_bytes(int R) {
    return kStageEnvs * (type_win_rows(R) * 8 + kVisWinBytes + 8);
}
__host__ __device__ inline int fast_tile_bytes(int D) { return align_up(16 * D, 16) + 256; }  // 4 rows + reset scratch
__host__ __device__ inline int fast_smem_bytes(int G, int R, int C, int D) {
    return tables_bytes(G, R, C) + align_up(3 * kStages * 8, 16) + kStages * fast_slot_bytes(R) +
           kConsumers * fast_tile_bytes(D);
}

// ---- TMA (bulk async copy) + mbarrier helpers -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
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
// bounded spin: a lost completion must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(mbar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}

struct RecRegs {       // one env's record + action, held in registers ahead of use
    uint4 ra, rb;
    long long action;
};

template <int R, int C, bool KEEP>
__global__ void __launch_bounds__(kFastThreads, PLANTOS_FAST_BLOCKS_PER_SM)
k_step_fast(const Params p, const StepIO io) {
    constexpr int D = 5 * C + 27;
    constexpr int NROW = 2 * R + 1;
    constexpr int VW = 4;                 // nibble words per visit row (G + 4 <= 32)
    constexpr int TP = R + 2;             // wall rows above the grid (== Params.TP)
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TWR = type_win_rows(R);
    constexpr int kTypeWinBytes = TWR * 8;
    constexpr int kSlotBytes = kStageEnvs * (kTypeWinBytes + kVisWinBytes + 8);
    static_assert(NROW <= 16 && C <= 16, "fast kernel shape limits");

    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = p.G, VE = p.VE, TS = p.TS;
    unsigned char* base = smem + tables_bytes(G, R, C);
    const uint32_t bars = smem_u32(base);                   // [kStages] tma | [kStages] full | [kStages] empty
    unsigned char* slots = base + align_up(3 * kStages * 8, 16);
    unsigned char* tiles = slots + kStages * kSlotBytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bars + 8 * s, 1);                     // tma:   producer's expect_tx arrival + bytes
            mbar_init(bars + 8 * (kStages + s), 1);         // full:  producer, after the transition
            mbar_init(bars + 8 * (2 * kStages + s), kConsumers);   // empty: one arrival per consumer warp
        }
        mbar_fence_init();
    }
    const Tables t = load_tables(p, smem);                  // ends with __syncthreads()
    typename std::conditional<KEEP, KeepMem, PlainMem>::type const mem;

    const int nstages = p.N / kStageEnvs;                   // full stages; the remainder is the tail
    const int first = blockIdx.x, stride = gridDim.x;
    const int nloc = first < nstages ? (nstages - first + stride - 1) / stride : 0;

    auto slot_twin = [&](int s) { return reinterpret_cast<uint64_t*>(slots + s * kSlotBytes); };
    auto slot_vwin = [&](int s) { return reinterpret_cast<uint32_t*>(slots + s * kSlotBytes + kStageEnvs * kTypeWinBytes); };
    auto slot_posw = [&](int s) {
        return reinterpret_cast<uint32_t*>(slots + s * kSlotBytes + kStageEnvs * (kTypeWinBytes + kVisWinBytes));
    };
    auto slot_aux = [&](int s) { return slot_posw(s) + kStageEnvs; };

    if (warp == 0) {
        // =========================== producer ===========================
        auto load_rec = [&](int k, RecRegs& q) {
            const size_t e = (size_t)(first + k * stride) * kStageEnvs + lane;
            q.ra = mem.ld128(p.rec + 2 * e);
            q.rb = mem.ld128(p.rec + 2 * e + 1);
            q.action = __ldcs(io.actions + e);
        };
        auto issue_tma = [&](int k, const RecRegs& q) {
            const int s = k % kStages;
            if (k >= kStages) {      // slot reuse: every consumer warp has released it
                mbar_wait(bars + 8 * (2 * kStages + s), ((k / kStages) & 1) ^ 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic writes vs the TMA's
            }
            const uint32_t tma = bars + 8 * s;
            if (lane == 0) mbar_arrive_expect_tx(tma, kStageEnvs * (kTypeWinBytes + kVisWinBytes));
            __syncwarp();
            const size_t e = (size_t)(first + k * stride) * kStageEnvs + lane;
            const int x = q.ra.x & 0xff;
            // grid rows x-R-1 .. x+R+1 are padded rows x+1 .. x+2R+3; start on the even row at
            // or just below x+1 so that source address and size are 16-byte multiples
            const int r0 = (x + 1) & ~1;
            bulk_load(smem_u32(slot_twin(s) + lane * TWR), p.types + e * TS + r0, kTypeWinBytes, tma);
            // grid rows x-3 .. x+3 are padded nibble rows x .. x+6
            bulk_load(smem_u32(slot_vwin(s) + lane * kVisWinRows * VW), p.vis4 + e * VE + (size_t)x * VW,
                      kVisWinBytes, tma);
        };

        RecRegs cur, nxt, nn;
        cur.ra = cur.rb = nxt.ra = nxt.rb = nn.ra = nn.rb = make_uint4(0, 0, 0, 0);
        cur.action = nxt.action = nn.action = 0;
        if (nloc > 0) { load_rec(0, cur); issue_tma(0, cur); }
        if (nloc > 1) load_rec(1, nxt);
        for (int k = 0; k < nloc; ++k) {
            if (k + 1 < nloc) issue_tma(k + 1, nxt);      // copies for the next stage
            if (k + 2 < nloc) load_rec(k + 2, nn);        // record loads two stages ahead
            const int s = k % kStages;
            const size_t e = (size_t)(first + k * stride) * kStageEnvs + lane;
            mbar_wait(bars + 8 * s, (k / kStages) & 1);   // this stage's windows have landed

            // ---- phase A: transition out of shared memory, one lane per env
            uint4 ra = cur.ra, rb = cur.rb;
            EnvRec r = unpack_rec(ra, rb);
            const int x0 = r.x, r0 = (r.x + 1) & ~1;      // the windows are centred on the pre-move row
            int tx, ty; bool inb;
            action_target(r, cur.action, G, tx, ty, inb);
            // (tx, ty) is at most one cell away, so it is inside both windows even when it is
            // outside the grid (wall padding / border nibbles)
            uint64_t* tw = slot_twin(s) + lane * TWR + (tx + TP - r0);
            const uint64_t word = *tw;
            const int t_cell = inb ? cell_of(word, ty & 31) : kObstacle;
            uint32_t* vw = slot_vwin(s) + (lane * kVisWinRows + (tx - x0 + 3)) * VW + ((ty + 2) >> 3);
            const int sh = nib_shift(ty);
            const uint32_t vword = *vw;
            const StepOut o = transition_core(r, cur.action, tx, ty, t_cell, (vword >> sh) & 15u, p.max_steps);
            if (o.moved)
                *vw = bump_visit(p.vis4 + e * VE + nib_word(tx, ty, VW), vword, sh,
                                 p.visov + e * G * G + tx * G + ty, mem);
            if (o.watered) {
                const uint64_t nw = word ^ (1ull << (2 * (ty & 31)));          // 3 -> 2
                *tw = nw;
                mem.st64(p.types + e * TS + TP + tx, nw);
            }
            r.ret += t.rw64[o.ridx];
            io.reward[e] = t.rw32[o.ridx];
            const int done = o.terminated | o.truncated;
            io.done[e] = (uint8_t)done;
            if (io.terminated) io.terminated[e] = (uint8_t)o.terminated;
            if (io.truncated) io.truncated[e] = (uint8_t)o.truncated;
            pack_rec(r, ra, rb);
            mem.st128(p.rec + 2 * e, ra);
            mem.st128(p.rec + 2 * e + 1, rb);
            if (done) {
                p.term_rec[2 * e] = ra;
                p.term_rec[2 * e + 1] = rb;
            }
            accumulate_stats(p, done, r, o.terminated, o.truncated, lane);
            // new position + where its windows start inside the fetched ones:
            //   type row of grid row x'-R is padded row x'+2, i.e. fetched row x'+2-r0   (0..3)
            //   nibble row of grid row x'-2 is padded row x'+1, i.e. fetched row x'+1-x0 (0..2)
            slot_posw(s)[lane] = (unsigned)r.x | ((unsigned)r.y << 8) | ((unsigned)(r.x + 2 - r0) << 16) |
                                 ((unsigned)(r.x + 1 - x0) << 20);
            slot_aux(s)[lane] = ((unsigned)r.episode << 1) | (unsigned)done;
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (kStages + s));   // publish the slot
            cur = nxt; nxt = nn;
        }
        return;
    }

    // =========================== consumers ===========================
    const int cw = warp - 1;                               // owns envs 4cw .. 4cw+3 of every stage
    float* tile = reinterpret_cast<float*>(tiles + cw * fast_tile_bytes(D));
    uint64_t* plane = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(tile) + align_up(16 * D, 16));
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
    constexpr int NCH = 2;             // envs per half-warp, interleaved stage by stage for ILP
    const float4* src4 = reinterpret_cast<const float4*>(tile);

    for (int k = 0; k < nloc; ++k) {
        const int s = k % kStages;
        const int stage = first + k * stride;
        mbar_wait(bars + 8 * (kStages + s), (k / kStages) & 1);     // slot published by the producer
        const uint64_t* twin = slot_twin(s);
        const uint32_t* vwin = slot_vwin(s);

        int x[NCH], y[NCH], tb[NCH], vb[NCH];
        unsigned w[NCH], vslice[NCH], acc[NCH], s0[NCH], s1[NCH];
        uint64_t trow[NCH];
        unsigned vlo[NCH], vhi[NCH];
        // stage 1: positions of the NCH envs this half-warp handles (env 4cw + 2c + half)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const unsigned pw = slot_posw(s)[4 * cw + 2 * c + half];
            x[c] = pw & 0xff; y[c] = (pw >> 8) & 0xff;
            tb[c] = (pw >> 16) & 15; vb[c] = pw >> 20;
        }
        const unsigned aux = slot_aux(s)[4 * cw + (lane & 3)];      // (episode << 1 | done) of env 4cw + lane%4
        // stage 2: shared-memory reads: this lane's type row and visit-nibble words
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int j = 4 * cw + 2 * c + half;
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
        // everything this warp needs from the slot is in registers: release it
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (2 * kStages + s));
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
        // flush four env rows = D float4: 16-byte aligned because the first env is a multiple of 4
        __syncwarp();
        const size_t e4 = (size_t)stage * kStageEnvs + 4 * cw;      // first env of this warp's group
        float4* dst4 = reinterpret_cast<float4*>(io.obs) + (e4 >> 2) * D;
#pragma unroll
        for (int q = 0; q < (D + 31) / 32; ++q) {
            const int idx = q * 32 + lane;
            if (idx < D && !(p.dbg & 1)) __stcs(dst4 + idx, src4[idx]);
        }
        __syncwarp();

        // ---- phase C: auto-reset of this warp's finished envs (rare; generic warp routines)
        unsigned dmask = __ballot_sync(FULL, (lane < 4) && (aux & 1u));
        while (dmask) {
            const int j = __ffs(dmask) - 1;
            dmask &= dmask - 1;
            const size_t ej = e4 + j;
            const int episode = (int)(__shfl_sync(FULL, aux, j) >> 1);
            // the pre-reset position is in the row just written: re-read it from the slot header
            // is not possible (slot released), so take it from the terminal record
            const uint4 tra = p.term_rec[2 * ej];
            const int px = tra.x & 0xff, py = (tra.x >> 8) & 0xff;
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

    // ragged tail: envs beyond the last full stage, one at a time
    if (cw == 0 && blockIdx.x == gridDim.x - 1)
        for (int e = nstages * kStageEnvs; e < p.N; ++e) step_env_warp(p, t, io, e, plane, tile, lane);
}

}  // namespace plantos_dev
