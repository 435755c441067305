// plantos_tile.cuh -- k_step_tile (one step per launch) and k_rollout_tile (K steps per launch, state resident
// on the SM): the round-2 hot kernels.  ONE LANE PER ENV for the whole simulation part of the step, the WHOLE
// WARP for the observation output, ONE BULK COPY per tile for the state.
//
// Same shape limits as k_step_fast (W == 1, VW == 4, G + R <= 32, C <= 16; R <= 6) plus: the LIDAR
// sample offsets must be the reference's own (plantos_lidar_gen.cuh holds them as compile-time
// tables, the host compares every uploaded table with them and routes to k_step_fast otherwise).
//
// Why a second specialised kernel: k_step_fast builds observations with one half-warp per env, which
// costs ~75 warp instructions and ~31 load/store-pipe wavefronts per env and a three-round-trip
// dependent fetch chain (record -> target words -> post-move windows) that every warp of the GPU walks
// at the same time (profiles/r1_summary.md).  Here, per 32-env TILE of a warp:
//   1. one lane issues ONE cp.async.bulk (TMA, mbarrier completion) for the tile's slice of the window
//      ring cache (plantos_common.cuh: the type rows x-R-1 .. x+R+1 and visit-nibble rows x-3 .. x+3 of
//      all 32 envs, 7.4 KB, at an address that does not depend on the rover positions) while every lane
//      loads its env's 32-byte record and action: a single round trip, no per-env address arithmetic;
//   2. lane-per-env: transition (plantos_env.py:160-222) out of the shared-memory rings (patches go to
//      shared memory, the plane and the cache), then the rover-centred window words of the POST-move
//      position go into registers, the LIDAR (plantos_env.py:251-292) is marched with compile-time
//      offsets (two ALU instructions per sample, no shuffles), and the env's observation is emitted as
//      a BYTE CODE: one byte per float = 4 * index into a 64-entry float table (0.0, 1.0, r/R,
//      min(v,10)/10, x/G, all host-evaluated like every table of this library);
//   3. the 32 byte rows form one flat image of the tile's [32, D] slice (32*D bytes, in the ring buffer,
//      which is dead by then); the whole warp expands it: per lane and iteration one code word, four
//      byte-permutes that splice a code byte into the table's 256-byte aligned address, four table
//      reads, one 128-bit streaming store -- fully coalesced, no float staging tile;
//   4. a rover that changed rows pulls the one type row and the one nibble row that entered its window
//      from the planes (loads issued before the expansion, stored into the cache after it).
// k_step_tile: 28 warps per SM (14 blocks x 2 warps, 72 registers): at the benchmark size every warp owns exactly
// one tile.  Auto-reset and the ragged tail reuse the generic warp routines, as in k_step_fast.
#pragma once
#include "plantos_fast.cuh"
#include "plantos_lidar_gen.cuh"

namespace plantos_dev {

#ifndef PLANTOS_TILE_WARPS
#define PLANTOS_TILE_WARPS 2          // small blocks: a finished block's slot goes to the next launch at once
#endif
#ifndef PLANTOS_TILE_MINBLOCKS
#define PLANTOS_TILE_MINBLOCKS 14     // 28 warps per SM (72 registers, 15.2 KB of shared memory per block)
#endif
constexpr int kTileWarps = PLANTOS_TILE_WARPS;
// -DPLANTOS_TILE_SKIP=<mask> builds timing-only variants that leave work out (tools/gpu_bisect.sh; wrong
// results): 1 no observation stores, 2 no expansion, 4 no encode, 8 no ring maintenance, 16 no phase A
// stores.  Never defined in the shipped build.
#ifndef PLANTOS_TILE_SKIP
#define PLANTOS_TILE_SKIP 0
#endif
constexpr int kTileLutBytes = 256;    // 64 floats on a 256-byte boundary: the first bytes of the dynamic shared memory

constexpr int kTileMbarBytes = 64;    // one 8-byte mbarrier per warp, then the launch ordinal
constexpr int kTileRwBytes = 160;     // the two reward tables (12 f64 + 12 f32), read in the transition
// per-warp scratch: the tile's window-ring-cache slice; reused as the flat byte code of the tile and as
// phase C's plane + row
// offset of the new-row staging area (32 B per lane) inside the scratch: beyond everything the expansion reads
__host__ __device__ constexpr int tile_stage_off(int D) { return ((8 * D + 31) / 32) * 128; }
__host__ __device__ inline int tile_warp_scratch_bytes(int R, int G, int D) {
    int b = wrc_tile_bytes(R);
    if (tile_stage_off(D) + 1024 > b) b = tile_stage_off(D) + 1024;
    const int code = align_up(32 * D, 16), resetscratch = align_up(G * 8, 16) + align_up(D * 4, 16);
    if (code > b) b = code;
    if (resetscratch > b) b = resetscratch;
    return b;
}
__host__ __device__ inline int tile_block_smem_bytes(int R, int G, int C) {
    return kTileLutBytes + kTileMbarBytes + kTileRwBytes + kTileWarps * tile_warp_scratch_bytes(R, G, 5 * C + 27);
}
// the multi-step kernel keeps the rings and has the code image (+ phase C scratch) behind them
__host__ __device__ inline int tile_warp_scratch_bytes_multi(int R, int G, int D) {
    int b = tile_stage_off(D);
    const int resetscratch = align_up(G * 8, 16) + align_up(D * 4, 16);
    if (resetscratch > b) b = resetscratch;
    return wrc_tile_bytes(R) + b;
}
__host__ __device__ inline int tile_block_smem_bytes_multi(int R, int G, int C) {
    return kTileLutBytes + kTileMbarBytes + kTileRwBytes + kTileWarps * tile_warp_scratch_bytes_multi(R, G, 5 * C + 27);
}

// ---- TMA (bulk async copy) + mbarrier
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
// orders this thread's (and, after a __syncwarp, its warp's) earlier generic-proxy accesses to shared
// memory before later async-proxy (TMA) writes to it
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* q) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* q, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(q), "r"(v) : "memory");
}
// one 32-byte record = one 256-bit access (sm_100: LDG/STG.E.ENL2.256), L2-coherent
__device__ __forceinline__ void ld_rec256(const uint4* q, uint4& a, uint4& b) {
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(q) : "memory");
}
__device__ __forceinline__ void st_rec256(uint4* q, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(q), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint4 lds_u128_v(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

__device__ __forceinline__ void sts_u32_v(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64_v(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }

// ---- byte-code emitters: OR five / one code bytes into the env's logical code words at a compile-time
// byte offset (the words are registers: every index is a constant after unrolling)
// ray group: B0 = db (a value < 256), B1..B4 = the bytes of kw
template <int OFF, int NWORDS>
__device__ __forceinline__ void emit_ray(uint32_t (&c)[NWORDS], uint32_t db, uint32_t kw) {
    constexpr int i = OFF >> 2, sh = OFF & 3;
    if (sh == 0) { c[i] |= __byte_perm(db, kw, 0x6540); c[i + 1] |= kw >> 24; }
    if (sh == 1) { c[i] |= __byte_perm(db, kw, 0x5401); c[i + 1] |= __byte_perm(db, kw, 0x1176); }
    if (sh == 2) { c[i] |= __byte_perm(db, kw, 0x4011); c[i + 1] |= __byte_perm(db, kw, 0x1765); }
    if (sh == 3) { c[i] |= db << 24; c[i + 1] |= kw; }
}
// visit group: B0..B3 = the bytes of lo4, B4 = hi1 (a value < 256)
template <int OFF, int NWORDS>
__device__ __forceinline__ void emit_vis(uint32_t (&c)[NWORDS], uint32_t lo4, uint32_t hi1) {
    constexpr int i = OFF >> 2, sh = OFF & 3;
    if (sh == 0) { c[i] |= lo4; c[i + 1] |= hi1; }
    if (sh == 1) { c[i] |= lo4 << 8; c[i + 1] |= __byte_perm(lo4, hi1, 0x5543); }
    if (sh == 2) { c[i] |= lo4 << 16; c[i + 1] |= __byte_perm(lo4, hi1, 0x5432); }
    if (sh == 3) { c[i] |= lo4 << 24; c[i + 1] |= __byte_perm(lo4, hi1, 0x4321); }
}
template <int OFF, int NWORDS>
__device__ __forceinline__ void emit_byte(uint32_t (&c)[NWORDS], uint32_t b) { c[OFF >> 2] |= b << (8 * (OFF & 3)); }

// Words FROM .. TO-1 of the env's code are complete: rotate them into the flat image (the env starts at
// byte D*lane, i.e. s8 bits into word 0 of the lane's span) and store them.  Word 0 waits for the tail
// of the previous env (see the end of the encode).
template <int FROM, int TO, int NWORDS>
__device__ __forceinline__ void flush_code(const uint32_t (&c)[NWORDS], uint32_t s_mycode, int s8, bool on) {
#pragma unroll
    for (int m = (FROM < 1 ? 1 : FROM); m < TO; ++m)
        if (on) sts_u32_v(s_mycode + 4 * m, __funnelshift_l(c[m - 1], c[m], s8));
}


// Outputs of K consecutive steps (K = 1: one plantos_step): step k reads actions + k * N and writes
// obs + k * obs_stride, reward / done / terminated / truncated + k * N; terminal_obs is one [N, D] buffer.
struct RollIO {
    const long long* actions;
    float* obs;
    size_t obs_stride;      // floats between the observation buffers of consecutive steps
    float* reward;
    uint8_t* done;
    uint8_t* terminated;
    uint8_t* truncated;
    float* terminal_obs;
    int K;
};

// MULTI = false: k_step_tile, one step per launch (plantos_step).
// MULTI = true:  k_rollout_tile, K steps per launch with the tile's rings and records RESIDENT on the SM
//                (plantos_rollout; the state-resident multi-step kernel of SURVEY 8f row 4): the rings are
//                fetched once, patched in place for K steps (visit-count patches also go to the plane, the
//                source of the rows that enter a window), and written back once; the code image has its own
//                buffer; half as many warps per SM, each walking its tiles one after the other.
// New episode for env ej (SB3 auto-reset, after the terminal observation has been written): map, fresh
// observation into `obs_row`, fresh rings into the cache (and into the resident copy at shared address s_ring if
// that is non-zero); returns the packed fresh record.  Deliberately NOT inlined and fed from the global-memory
// copy of the parameters: inlined, its register needs made ptxas spill loop invariants of the hot loop (23 % of
// the rollout kernel's stall samples were reloads from local memory).
struct PackedRec { uint4 a, b; };
__device__ __noinline__ PackedRec tile_reset_env(const Params* gp, size_t ej, int ep, uint64_t* plane, float* row_s,
                                                  float* obs_row, uint32_t s_ring, int lane) {
    const Params& p = *gp;
    int keep = 0, map_ep = -1;
    if (p.cur_mode) {                                        // CurriculumWrapper.reset
        int cr = 0;
        if (lane == 0) cr = curriculum_on_reset(p, (int)ej, ep);
        cr = __shfl_sync(0xffffffffu, cr, 0);
        keep = cr & 1; map_ep = cr >> 1;
    }
    const EnvRec nr = reset_env_warp<false>(p, (int)ej, ep, plane, lane, keep != 0, map_ep);
    // (fresh visit window: what the planes hold after a plain reset, and what the wrapper's reset observation
    // shows when the counts are kept)
    const Tables tb = tables_at(const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(p.table_blob)), p.G, p.R);
    build_obs_warp(p, tb, plane, p.vis4 + ej * p.VE, nr.x, nr.y, row_s, lane, true);
    store_obs_row(row_s, obs_row, p.D, lane);
    wrc_build_env_fresh_warp(p, ej, nr.x, nr.y, plane, keep != 0, lane, s_ring);
    PackedRec out;
    pack_rec(nr, out.a, out.b);
    __syncwarp();
    return out;
}

template <int R, int C, bool MULTI>
__device__ __forceinline__ void tile_body(const Params& p, const RollIO& io) {
    using Gen = LidarGen<R, C>;
    static_assert(Gen::ok, "no generated LIDAR offsets for this (R, C)");
    constexpr int D = 5 * C + 27;
    constexpr int NROW = 2 * R + 1;
    constexpr int VW = 4;                 // nibble words per visit row (G + 4 <= 32)
    constexpr int TP = R + 2;             // wall rows above the grid (== Params.TP)
    constexpr int NTR = wrc_type_slots(R), WRCB = wrc_tile_bytes(R);
    constexpr int NW = (D + 3) / 4;       // logical code words per env
    constexpr int TB = D & 3;             // bytes in the last logical word (0 = all four)
    constexpr int LB_DIST = 2, LB_VIS = R + 4, LB_POS = R + 20;   // table sections: 0.0 | 1.0 | r/R (R+2) | nibble (16) | x/G (G)
    constexpr int NVEC = 8 * D;           // float4 per full tile
    constexpr int NIT = (NVEC + 31) / 32; // expansion iterations per full tile
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NROW <= 13 && C <= 16, "tile kernel shape limits");
    static_assert(D >= 8, "flat code packing needs two logical words");

    extern __shared__ __align__(256) unsigned char smem[];
    const PlainMem mem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef PLANTOS_EXP_TIMING
    unsigned ts_[10] = {};
#endif
    TSTAMP(0);
    const int G = p.G, VE = p.VE, TS = p.TS;
    // the decode table sits on a 256-byte boundary (checked below): a code byte then IS the low address byte
    const uint32_t s_smem = smem_u32(smem);
    const uint32_t s_lut = s_smem;
    const uint32_t s_mbar = s_smem + kTileLutBytes + 8 * warp;
    const uint32_t s_rw64 = s_smem + kTileLutBytes + kTileMbarBytes, s_rw32 = s_rw64 + 96;
    unsigned char* const scratch = smem + kTileLutBytes + kTileMbarBytes + kTileRwBytes +
                                   warp * (MULTI ? tile_warp_scratch_bytes_multi(R, G, D) : tile_warp_scratch_bytes(R, G, D));
    // the tables stay in global memory (the host-packed image, plantos_common.cuh: tables_at): the hot
    // loop only needs the two reward tables (L1 hits), the rare generic paths read them as they are
    // (recomputed at every use: seven pointers are not worth registers)
    auto tabs = [&]() { return tables_at(const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(p.table_blob)), G, R); };
    const uint32_t s_win = smem_u32(scratch);
    // the code image reuses the ring buffer (single step) or sits behind the resident rings (multi-step)
    constexpr int CODE_OFF = MULTI ? WRCB : 0;
    const uint32_t s_code = s_win + CODE_OFF;
    uint64_t* const plane = reinterpret_cast<uint64_t*>(scratch + CODE_OFF);                      // phase C scratch
    float* const row_s = reinterpret_cast<float*>(scratch + CODE_OFF + align_up(G * 8, 16));
    const int K = MULTI ? io.K : 1;

    // tiles of 32 envs; tile (round r, block b, warp w) = r * nwarps + w * nblocks + b, so every
    // round of the persistent loop spreads evenly over the SMs
    const int nwarps = gridDim.x * kTileWarps;
    const int nfull = p.N & ~3;                             // envs in whole 4-env groups
    const int ntiles = (nfull + 31) >> 5;

    // Launch ordinal: block b counts its own launches in tickets[b], and it does so BEFORE it lets the next
    // launch start (the next launch's block b can only exist after every block of this one has passed
    // this point), so the value it reads numbers the launches without any contention.
    int t = warp * gridDim.x + blockIdx.x;
    unsigned ticket = 0, flag0 = 0;
    if (threadIdx.x == 0) ticket = atomicAdd(p.tickets + blockIdx.x, 1u);
    // (pipelined launches: a first look at the first tile's flag travels together with the ticket)
    if (lane == 0 && t < ntiles && p.pipelined) flag0 = ld_acquire_u32(p.tile_flags + t);
    if (lane == 0) mbar_init(s_mbar, 1);
    if (threadIdx.x < 64) {                                 // the decode table (immutable inputs)
        const Tables tb = tabs();
        const int i = threadIdx.x;
        float v = 0.0f;
        if (i == 1) v = 1.0f;
        else if (i >= LB_DIST && i < LB_VIS) v = __ldg(tb.dist + min(i - LB_DIST, R));      // entry R+1 repeats r = R: "no hit"
        else if (i >= LB_VIS && i < LB_POS) v = __ldg(tb.visit + (i - LB_VIS));
        else if (i >= LB_POS && i < LB_POS + G) v = __ldg(tb.pos + (i - LB_POS));
        sts_u32_v(s_lut + 4 * i, __float_as_uint(v));
        // (acquire loads invalidate the SM's L1 all the time in pipelined mode: the reward tables live in shared memory too)
        if (i < 2 * kRwCount) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(tb.rw64) + i);
            sts_u64_v(s_rw64 + 8 * i, (uint64_t)q.x | ((uint64_t)q.y << 32));
            sts_u32_v(s_rw32 + 4 * i, __float_as_uint(__ldg(tb.rw32 + i)));
        }
    }
    if (threadIdx.x == 0) {
        sts_u32_v(s_smem + kTileLutBytes + 8 * kTileWarps, ticket);
        if ((s_lut & 255u) != 0u) atomicExch(p.err, -6);    // (the launch configuration guarantees the alignment)
    }
    __syncthreads();
    const unsigned ordinal = lds_u32_v(s_smem + kTileLutBytes + 8 * kTileWarps);
    // Programmatic dependent launch (see k_step_fast): nothing mutable but the ticket is touched before the wait.
    griddep_launch_dependents();
    uint32_t parity = 0;
    // Pipelined launches (plantos_set_pipelining) do NOT wait for the previous grid: the per-tile flags
    // below order a tile's steps, so this launch's loads and simulation overlap the previous launch's
    // observation stores.  Otherwise the usual full dependency.
    if (!p.pipelined) griddep_wait();
    TSTAMP(1);
    auto wait_tile = [&](int tile, unsigned seen) {         // lane 0: until the previous step of this tile is complete
        if (p.pipelined) {
            const unsigned* f = p.tile_flags + tile;
            unsigned spins = 0;
            while (seen != ordinal) {
                __nanosleep(64);
                if (++spins > (1u << 24)) { atomicExch(p.err, -5); break; }   // never hang the GPU on a protocol error
                seen = ld_acquire_u32(f);
            }
            fence_proxy_async_global();                     // the bulk copy below reads what other SMs stored
        }
    };
#ifndef PLANTOS_TILE_RINGCOPY
#define PLANTOS_TILE_RINGCOPY 0       // 0: one TMA bulk copy per tile; 1: 16-byte cp.async copies by all lanes (experiment)
#endif
    auto fetch_rings = [&](int tile, unsigned seen) {
        if (PLANTOS_TILE_RINGCOPY) {
            if (lane == 0) wait_tile(tile, seen);
            __syncwarp();
            const unsigned char* src = p.wrc + (size_t)tile * WRCB + 16 * lane;
#pragma unroll
            for (int k = 0; k < (WRCB + 511) / 512; ++k)
                if (512 * k + 16 * lane < WRCB) cp_async16(s_win + 512 * k + 16 * lane, src + 512 * k);
            cp_async_commit();
        } else if (lane == 0) {
            wait_tile(tile, seen);
            mbar_arrive_expect_tx(s_mbar, WRCB);
            bulk_load(s_win, p.wrc + (size_t)tile * WRCB, WRCB, s_mbar);
        }
    };
    if (t < ntiles) fetch_rings(t, flag0);                  // the first tile's rings are on their way at once
    constexpr uint64_t LOWPAD = kObstAll & ((1ull << (2 * R)) - 1ull);
    // flat code image: env j's D bytes start at byte D*j; this lane owns words a_w .. a_w + M - 1
    const int A = D * lane, s8 = 8 * (A & 3);
    const uint32_t s_mycode = s_code + 4 * (A >> 2);
    const bool own_last = (((A & 3) + D) >> 2) == NW;
    // this lane's column of the ring planes: u64 type planes, then u32 nibble planes
    const uint32_t s_t8 = s_win + 8 * lane, s_n4 = s_win + NTR * 256 + 4 * lane;

    for (bool first = true; t < ntiles; t += nwarps, first = false) {
        const int e0 = t * 32;
        const int ts = min(32, nfull - e0);                  // envs in this tile (a multiple of 4)
        const bool act = lane < ts;
        const unsigned e = (unsigned)(e0 + lane);            // 32-bit element offsets (host check)
        unsigned char* const g_tile = p.wrc + (size_t)t * WRCB;   // this tile's rings in global memory
        if (!first) {                                        // (the whole warp is past its last use of the buffer)
            if (lane == 0) fence_proxy_async_shared();
            fetch_rings(t, (p.pipelined && lane == 0) ? ld_acquire_u32(p.tile_flags + t) : 0u);
        }

        __syncwarp();                                        // (lane 0 has seen the tile's flag)
        // ---- records + actions, one lane per env (L2 loads: another SM may have just written the record)
        uint4 ra = make_uint4(0, 0, 0, 0), rbw = ra;
        long long action = 0;
        if (act) {
            ld_rec256(p.rec + 2 * e, ra, rbw);
            action = __ldcg(io.actions + e);
        }
        EnvRec r = unpack_rec(ra, rbw);
        if (r.x >= 0) TSTAMP(2);
        if (PLANTOS_TILE_RINGCOPY) { cp_async_wait_all(); __syncwarp(); }
        else { while (!mbar_try_wait(s_mbar, parity)) {} parity ^= 1u; }
        TSTAMP(4);

      for (int k = 0; k < K; ++k) {                          // (one pass for k_step_tile)
        const size_t ko = (size_t)k * (size_t)p.N;           // offset of step k in the [K, N] outputs
        long long action_next = 0;
        if (MULTI) {
            cp_async_wait_all();                             // the rows the previous step pulled into this lane's rings
            if (act && k + 1 < K) action_next = __ldcg(io.actions + ko + p.N + e);   // in flight during this step (issued
                                                             // AFTER the wait: both would sit on the same scoreboard;
                                                             // waiting only before the window read measured 2 % slower)
        }
        const int x0 = r.x;
        // ---- transition (plantos_env.py:160-222), one lane per env, out of the rings
        int done = 0, term = 0, trunc = 0;
        if (act) {
            int tx, ty; bool inb;
            action_target(r, action, G, tx, ty, inb);
            // grid row g is padded type row g+TP and padded nibble row g+3; the rings hold padded type
            // rows x0+1 .. x0+NTR and padded nibble rows x0 .. x0+6
            const uint32_t o_tw = 256u * (unsigned)((tx + TP) % NTR);
            const uint32_t o_vw = 128u * (unsigned)(((tx + 3) % 7) * 4 + ((ty + 2) >> 3));
            const uint64_t word = lds_u64_v(s_t8 + o_tw);
            const uint32_t vword = lds_u32_v(s_n4 + o_vw);
            const int t_cell = inb ? cell_of(word, ty & 31) : kObstacle;
            const int sh = nib_shift(ty);
            int expl_fresh = -1;
            if (p.cur_mode && inb && action < 4) {           // CurriculumWrapper: this episode's explored_map
                const uint32_t ew = p.expl[e * G + tx], bit = 1u << (ty & 31);
                expl_fresh = (ew & bit) ? 0 : 1;
                if (t_cell != kObstacle) p.expl[e * G + tx] = ew | bit;
            }
            StepOut o = transition_core(r, action, tx, ty, t_cell, (vword >> sh) & 15u, p.max_steps, expl_fresh);
            if (p.cur_mode) {                                // CurriculumWrapper.step, A2C_training.py:94-100
                const double pct = ((double)r.explored / (double)r.total_free) * 100.0;
                if (pct >= p.cur_thr[e]) {
                    p.cur_cnt[e].y |= 1;
                    if (p.cur_mode == 1) o.terminated = 1;
                }
            }
            if (o.moved) {                                   // visit count + 1: plane, ring (shared), ring (cache)
                const uint32_t nw = bump_visit(p.vis4 + e * VE + nib_word(tx, ty, VW), vword, sh, p.visov + e * G * G + tx * G + ty, mem);
                sts_u32_v(s_n4 + o_vw, nw);
                if (!MULTI) *reinterpret_cast<uint32_t*>(g_tile + NTR * 256 + 4 * lane + o_vw) = nw;   // (MULTI: rings written back at the end)
            }
            if (o.watered) {                                 // 3 -> 2
                const uint64_t nw = word ^ (1ull << (2 * (ty & 31)));
                mem.st64(p.types + e * TS + TP + tx, nw);
                sts_u64_v(s_t8 + o_tw, nw);
                if (!MULTI) *reinterpret_cast<uint64_t*>(g_tile + 8 * lane + o_tw) = nw;
            }
            r.ret += lds_f64(s_rw64 + 8 * o.ridx);
            term = o.terminated; trunc = o.truncated; done = term | trunc;
            pack_rec(r, ra, rbw);
            if (!(PLANTOS_TILE_SKIP & 16) || r.step == 54321) {
            io.reward[ko + e] = lds_f32(s_rw32 + 4 * o.ridx);
            io.done[ko + e] = (uint8_t)done;
            if (io.terminated) io.terminated[ko + e] = (uint8_t)term;
            if (io.truncated) io.truncated[ko + e] = (uint8_t)trunc;
            if (!MULTI) st_rec256(p.rec + 2 * e, ra, rbw);   // (MULTI: the record stays in registers until the last step)
            }
            if (done) st_rec256(p.term_rec + 2 * e, ra, rbw);
        }
        accumulate_stats(p, act && done, r, term, trunc, lane, (int)e, (unsigned)k);
        TSTAMP(5);

        // ---- the POST-move window into registers (idle lanes read their stale column: harmless)
        const int x1 = r.x, y1 = r.y, dxm = x1 - x0;
        const int episode = r.episode;
        unsigned w[NROW], sl[5];
        {
            // needed padded type rows x1+2 .. x1+2R+2: ring slots b, b+1, ... wrapping at NTR
            const int b = (x1 + 2) % NTR, kwrap = NTR - b;
            const uint32_t base_n = s_t8 + 256 * b, base_w = base_n - 256 * NTR;
            const int sft = 2 * y1;
            uint64_t raw[NROW];
#pragma unroll
            for (int i = 0; i < NROW; ++i) raw[i] = lds_u64_v((i >= kwrap ? base_w : base_n) + 256 * i);
            // needed padded nibble rows x1+1 .. x1+5; grid column y1-2+k is nibble y1+k: the five nibbles
            // sit in words y1>>3 and (one further, if any)
            const int bv = (x1 + 1) % 7, kv = 7 - bv, w0 = y1 >> 3;
            const uint32_t vbase_n = s_n4 + 512 * bv + 128 * w0, vbase_w = vbase_n - 512 * 7;
            const uint32_t w1off = w0 < 3 ? 128u : 0u;
            const int vs = 4 * (y1 & 7);
            uint32_t va[5], vb[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const uint32_t a = (i >= kv ? vbase_w : vbase_n) + 512 * i;
                va[i] = lds_u32_v(a); vb[i] = lds_u32_v(a + w1off);
            }
#pragma unroll
            for (int i = 0; i < NROW; ++i) {
                const uint64_t ext = (raw[i] << (2 * R)) | LOWPAD;
                w[i] = (unsigned)((ext >> sft) | ((kObstAll << 1) << (63 - sft)));   // cells y1-R .. y1+R, walls outside
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) sl[i] = __funnelshift_r(va[i], vb[i], vs);
        }
        __syncwarp();                                        // every lane has read its rings: the buffer becomes the code image
        if (w[0] != 0xdeadbeefu) TSTAMP(6);
        // a rover that changed rows: the type row and the nibble row that entered its window come from
        // the planes, by cp.async into a staging area behind the code image while the encode runs
        const bool newrow = act && !done && dxm != 0 && !(PLANTOS_TILE_SKIP & 8);
        const int pr_new = dxm > 0 ? x1 + NTR : x1 + 1, pn_new = dxm > 0 ? x1 + 6 : x1;
        const uint32_t s_stage = s_win + tile_stage_off(D) + 32 * lane;
        if (newrow) {
            if (MULTI) {                                     // straight into the ring slots of the rows that left (resident rings)
                cp_async8(s_t8 + 256 * (pr_new % NTR), p.types + e * TS + pr_new);
                const uint32_t* vsrc = p.vis4 + e * VE + pn_new * VW;
                const uint32_t sd = s_n4 + 512 * (pn_new % 7);
                cp_async4(sd, vsrc); cp_async4(sd + 128, vsrc + 1); cp_async4(sd + 256, vsrc + 2); cp_async4(sd + 384, vsrc + 3);
            } else {
                cp_async16(s_stage, p.vis4 + e * VE + pn_new * VW);
                cp_async8(s_stage + 16, p.types + e * TS + pr_new);
            }
        }
        cp_async_commit();

        // ---- observation as a byte code (plantos_env.py:251-315); complete words are stored as soon as
        // they are final so that they do not occupy registers
        uint32_t c[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; ++i) c[i] = 0u;
        // LIDAR (:260-292): sample rr of a ray lands at bits 2(R-rr), 2(R-rr)+1 of acc, so the nearest
        // non-empty sample is the HIGHEST set bit; bit 0 is the "nothing hit" sentinel
#define PLANTOS_TILE_RAY(ray)                                                                               \
        if (ray < C) {                                                                                      \
            constexpr int ry = ray < C ? ray : 0;                                                           \
            unsigned acc = 1u;                                                                              \
            _Pragma("unroll")                                                                               \
            for (int rr = 0; rr < R; ++rr) {                                                                \
                const int pos = 2 * (Gen::dy(ry, rr) + R), tgt = 2 * (R - rr);                              \
                const unsigned wr = w[Gen::dx(ry, rr) + R];                                                 \
                const unsigned al = pos > tgt ? wr >> (pos - tgt) : (pos < tgt ? wr << (tgt - pos) : wr);   \
                acc |= al & (3u << tgt);                                                                    \
            }                                                                                               \
            const int g2 = (31 - __clz(acc)) & ~1;          /* 2(R-rr) of the hit, 0: none */                \
            const unsigned kk = g2 ? acc >> g2 : 0u;        /* nearer samples are empty: the two bits alone */ \
            const unsigned kw = __funnelshift_l(0u, 4u, kk << 3);      /* one-hot: byte `kind` = 4 -> 1.0 */ \
            const unsigned db = (unsigned)(4 * (LB_DIST + R + 1) - 2 * g2);   /* r/R with r = R+1-g2/2 (R+1: the repeat of R) */ \
            emit_ray<5 * ry>(c, db, kw);                                                                    \
            flush_code<(5 * ry) / 4, (5 * ry + 5) / 4>(c, s_mycode, s8, act);                               \
        }
        if (PLANTOS_TILE_SKIP & 4) {
            unsigned xs = 0;
#pragma unroll
            for (int i = 0; i < NROW; ++i) xs ^= w[i];
#pragma unroll
            for (int i = 0; i < 5; ++i) xs ^= sl[i];
            c[0] = xs;
        } else {
        PLANTOS_TILE_RAY(0) PLANTOS_TILE_RAY(1) PLANTOS_TILE_RAY(2) PLANTOS_TILE_RAY(3)
        PLANTOS_TILE_RAY(4) PLANTOS_TILE_RAY(5) PLANTOS_TILE_RAY(6) PLANTOS_TILE_RAY(7)
        PLANTOS_TILE_RAY(8) PLANTOS_TILE_RAY(9) PLANTOS_TILE_RAY(10) PLANTOS_TILE_RAY(11)
        PLANTOS_TILE_RAY(12) PLANTOS_TILE_RAY(13) PLANTOS_TILE_RAY(14) PLANTOS_TILE_RAY(15)
        }
#undef PLANTOS_TILE_RAY
        emit_byte<5 * C>(c, (unsigned)(4 * (LB_POS + x1)));          // :294-296
        emit_byte<5 * C + 1>(c, (unsigned)(4 * (LB_POS + y1)));
        flush_code<(5 * C) / 4, (5 * C + 2) / 4>(c, s_mycode, s8, act);
        // 5x5 visit window (:298-313): nibble k -> byte 4 * (LB_VIS + k)
#define PLANTOS_TILE_VIS(i)                                                                                 \
        {                                                                                                   \
            unsigned v = sl[i] & 0xffffu;                                                                   \
            v = (v | (v << 8)) & 0x00ff00ffu;                                                               \
            v = (v | (v << 4)) & 0x0f0f0f0fu;                                                               \
            emit_vis<5 * C + 2 + 5 * i>(c, v * 4u + 0x01010101u * (4u * LB_VIS), ((sl[i] >> 16) & 15u) * 4u + 4u * LB_VIS); \
            flush_code<(5 * C + 2 + 5 * i) / 4, (5 * C + 2 + 5 * i + 5) / 4>(c, s_mycode, s8, act);         \
        }
        if (!(PLANTOS_TILE_SKIP & 4)) {
        PLANTOS_TILE_VIS(0) PLANTOS_TILE_VIS(1) PLANTOS_TILE_VIS(2) PLANTOS_TILE_VIS(3) PLANTOS_TILE_VIS(4)
        }
#undef PLANTOS_TILE_VIS
        // the words that are still open: the last one (only if this lane owns it) and word 0, which also
        // carries the last (A & 3) bytes of the previous env
        {
            static_assert((5 * C + 27) / 4 == NW - (TB ? 1 : 0), "flush bookkeeping");
            const uint32_t tail = TB ? __funnelshift_r(c[NW - 2], c[NW - 1], 8 * TB) : c[NW - 1];
            const uint32_t tprev = __shfl_up_sync(FULL, tail, 1);
            if (act) {
                sts_u32_v(s_mycode, __funnelshift_l(tprev, c[0], s8));
                if (TB && own_last) sts_u32_v(s_mycode + 4 * (NW - 1), __funnelshift_l(c[NW - 2], c[NW - 1], s8));
            }
        }
        // the fetched rows of the rovers that changed rows go from the staging area into the cache rings,
        // replacing the rows that left the windows; then the tile's state is complete
        if (!MULTI) cp_async_wait_all();
        if (!MULTI && newrow) {
            const uint4 nrow_new = lds_u128_v(s_stage);
            const uint64_t trow_new = lds_u64_v(s_stage + 16);
            *reinterpret_cast<uint64_t*>(g_tile + 256 * (pr_new % NTR) + 8 * lane) = trow_new;
            uint32_t* const gn = reinterpret_cast<uint32_t*>(g_tile + NTR * 256 + 512 * (pn_new % 7) + 4 * lane);
            gn[0] = nrow_new.x; gn[32] = nrow_new.y; gn[64] = nrow_new.z; gn[96] = nrow_new.w;
        }
        const unsigned dmask0 = __ballot_sync(FULL, act && done);
        __syncwarp();                                        // orders every lane's state stores before lane 0's release
        if (!MULTI && dmask0 == 0u && lane == 0) {           // the tile's state is complete: the next step may start
            if (p.release) st_release_u32(p.tile_flags + t, ordinal + 1u);
            else p.tile_flags[t] = ordinal + 1u;             // (nobody overlaps launches on this handle: the kernel end publishes it)
        }
        __syncwarp();
        TSTAMP(7);

        // ---- expand: the whole warp, one float4 per lane and iteration, coalesced streaming stores
        {
            float* const obs_k = io.obs + (size_t)k * io.obs_stride;
            float4* const dst = reinterpret_cast<float4*>(obs_k) + ((size_t)e0 * D >> 2) + lane;
            const uint32_t s_cw = s_code + 4 * lane;
            if ((PLANTOS_TILE_SKIP & 2) && ts >= 0) {
            } else if (ts == 32) {
                constexpr int LASTN = NVEC - 32 * (NIT - 1);        // lanes of the last iteration
#pragma unroll
                for (int q0 = 0; q0 < NIT; q0 += 3) {
                    uint32_t cw[3];
                    float4 f[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u)
                        if (q0 + u < NIT) {
                            cw[u] = lds_u32_v(s_cw + 128 * (q0 + u));
                            // lanes beyond the image in the last iteration read stale bytes: make them a valid code
                            if (q0 + u == NIT - 1 && LASTN != 32 && lane >= LASTN) cw[u] = 0u;
                        }
#pragma unroll
                    for (int u = 0; u < 3; ++u)
                        if (q0 + u < NIT) {
                            f[u].x = lds_f32(__byte_perm(cw[u], s_lut, 0x7650));
                            f[u].y = lds_f32(__byte_perm(cw[u], s_lut, 0x7651));
                            f[u].z = lds_f32(__byte_perm(cw[u], s_lut, 0x7652));
                            f[u].w = lds_f32(__byte_perm(cw[u], s_lut, 0x7653));
                        }
#pragma unroll
                    for (int u = 0; u < 3; ++u)
                        if (q0 + u < NIT && (q0 + u < NIT - 1 || LASTN == 32 || lane < LASTN) &&
                            (!(PLANTOS_TILE_SKIP & 1) || f[u].x == 123.0f)) __stcs(dst + (q0 + u) * 32, f[u]);
                }
            } else {
                const int nvec = (ts * D) >> 2;
#pragma unroll 1
                for (int q = 0; q * 32 + lane < nvec; ++q) {
                    const uint32_t cw = lds_u32_v(s_cw + 128 * q);
                    float4 f;
                    f.x = lds_f32(__byte_perm(cw, s_lut, 0x7650));
                    f.y = lds_f32(__byte_perm(cw, s_lut, 0x7651));
                    f.z = lds_f32(__byte_perm(cw, s_lut, 0x7652));
                    f.w = lds_f32(__byte_perm(cw, s_lut, 0x7653));
                    __stcs(dst + q * 32, f);
                }
            }
        }
        __syncwarp();                                        // the buffer may be reused
        TSTAMP(8);
#ifdef PLANTOS_EXP_TIMING
        { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); ts_[9] = sm; }
        if (lane == 0 && ts == 32)                            // the stamps of the last four launches, side by side
            for (int k = 0; k < 5; ++k) {
                const size_t slot = (size_t)e0 + 5 * (ordinal & 3u) + k;
                uint4 v = p.term_rec[2 * slot + 1];
                v.z = ts_[2 * k]; v.w = ts_[2 * k + 1];
                p.term_rec[2 * slot + 1] = v;
            }
#endif

        // ---- auto-reset of finished envs (SB3 semantics; rare).  The observation just expanded IS the
        // terminal observation: it is expanded once more, straight out of the code image, into
        // terminal_obs (all finished envs first: the resets below reuse the image as their scratch).
        unsigned dmask = dmask0;
        if (io.terminal_obs) {
            while (dmask) {
                const int j = __ffs(dmask) - 1;
                dmask &= dmask - 1;
                float* const trow = io.terminal_obs + ((size_t)e0 + j) * D;
                for (int idx = lane; idx < D; idx += 32) {
                    const uint32_t a = s_code + D * j + idx;
                    const uint32_t byte = (lds_u32_v(a & ~3u) >> (8 * (a & 3u))) & 0xffu;
                    trow[idx] = lds_f32(s_lut + byte);
                }
            }
            __syncwarp();
            dmask = dmask0;
        }
        if (p.map_source == 2) dmask = 0u;                   // maze handles: k_reset_done starts the new episodes after this launch
        while (dmask) {
            const int j = __ffs(dmask) - 1;
            dmask &= dmask - 1;
            const size_t ej = (size_t)e0 + j;
            const int ep = __shfl_sync(FULL, episode, j);
            const PackedRec fresh = tile_reset_env(p.self, ej, ep, plane, row_s, io.obs + (size_t)k * io.obs_stride + ej * D,
                                                   MULTI ? s_win : 0u, lane);
            if (MULTI) { if (lane == j) r = unpack_rec(fresh.a, fresh.b); }   // the lane's record restarts
            else if (lane == 0) st_rec256(p.rec + 2 * ej, fresh.a, fresh.b);
            __syncwarp();
        }
        if (!MULTI && dmask0 != 0u && lane == 0) {           // tiles with resets complete here
            if (p.release) st_release_u32(p.tile_flags + t, ordinal + 1u);
            else p.tile_flags[t] = ordinal + 1u;
        }
        action = action_next;
      }   // steps of the tile
        if (MULTI) {
            // the K steps of this tile are done: records and rings go back to global memory once
            cp_async_wait_all();
            if (act) {
                pack_rec(r, ra, rbw);
                st_rec256(p.rec + 2 * e, ra, rbw);
            }
            __syncwarp();
            for (int q = lane; q < WRCB / 16; q += 32)
                reinterpret_cast<uint4*>(g_tile)[q] = lds_u128_v(s_win + 16 * q);
            __syncwarp();
            if (lane == 0) {
                if (p.release) st_release_u32(p.tile_flags + t, ordinal + 1u);
                else p.tile_flags[t] = ordinal + 1u;
            }
            __syncwarp();
        }
    }

    // ragged tail: envs beyond the last 4-env group, one at a time (they are never part of a tile)
    if (blockIdx.x == gridDim.x - 1 && warp == kTileWarps - 1)
        for (int k = 0; k < K; ++k) {
            StepIO sio;
            const size_t ko = (size_t)k * (size_t)p.N;
            sio.actions = io.actions + ko; sio.obs = io.obs + (size_t)k * io.obs_stride; sio.reward = io.reward + ko;
            sio.done = io.done + ko; sio.terminated = io.terminated ? io.terminated + ko : nullptr;
            sio.truncated = io.truncated ? io.truncated + ko : nullptr; sio.terminal_obs = io.terminal_obs;
            for (int e = nfull; e < p.N; ++e) step_env_warp<false>(p, tabs(), sio, e, plane, row_s, lane);
        }
}

// The two kernels: 28 warps per SM at 72 registers for the single step; for the state-resident rollout 16 warps
// per SM (the code image needs its own buffer behind the rings: 22 KB per block) at 112 registers -- four warps
// per scheduler; 128 registers measured 1 % slower, 96 registers spill in the step loop and 144 registers leave
// only 7 blocks per SM (profiles/r2_summary.md).
template <int R, int C>
__global__ void __launch_bounds__(kTileWarps * 32, PLANTOS_TILE_MINBLOCKS)
k_step_tile(const Params p, const RollIO io) { tile_body<R, C, false>(p, io); }

#ifndef PLANTOS_ROLLOUT_REGS
#define PLANTOS_ROLLOUT_REGS 112
#endif
template <int R, int C>
__global__ void __maxnreg__(PLANTOS_ROLLOUT_REGS)
k_rollout_tile(const Params p, const RollIO io) { tile_body<R, C, true>(p, io); }

}  // namespace plantos_dev
