// plantos_fast.cuh -- the sm_100a hot kernel for the reference presets.
//
// Requirements (checked on the host): W == 1 and VW == 4 (G <= 28), G + R <= 32, 2R+1 <= 16,
// C <= 16.  Both the training preset (G25 R6 C16, D=107; A2C_training.py:206-212) and the ctor
// default (G21 R2 C10, D=77; plantos_env.py:25-26) qualify; everything else runs
// k_step_generic.
//
// Persistent grid (PLANTOS_FAST_MINBLOCKS blocks per SM); every warp owns a contiguous range of
// envs.  The order of work inside a warp keeps every lane busy in the transition and fetches
// only what is used:
//   * a warp walks its contiguous env range in MACRO TILES of 32 envs.  The transition
//     (plantos_env.py:160-222) runs once per macro tile with one lane per env -- all 32 lanes
//     active instead of 8 -- and needs exactly two words of plane state per env: the type row
//     and the visit-nibble word of the cell the action looks at.  Those two words (and the
//     32-byte record + the action before them) are prefetched with cp.async while the previous
//     macro tile is still building observations.
//   * the observations (plantos_env.py:251-315) are built in TRIPS of 4 envs, one half-warp per
//     env and two envs per half-warp, out of shared-memory windows centred on the POST-move
//     position (no margin rows: 2R+2 type rows and 5 nibble rows = 12 sixteen-byte chunks per env
//     for R = 6).  The windows of trip k+1 are copied (cp.async, 48 chunks over 32 lanes) while
//     trip k computes; they are read after the transition's global stores of the same warp, so
//     no shared-memory patching is needed.
//   * rows are assembled in a 4-env shared-memory tile whose 16*D bytes are 16-byte aligned in
//     the [N, D] fp32 buffer and leave with streaming 128-bit stores (st.global.cs.v4,
//     evict-first), so the write-once observation stream does not evict the env state from L2.
//   * auto-reset of finished envs (SB3 semantics: terminal observation, Philox / injected map,
//     fresh observation) and the ragged N % 4 tail use the generic warp routines.
#pragma once
#include <type_traits>
#include "plantos_generic.cuh"

namespace plantos_dev {

#ifndef PLANTOS_FAST_MINBLOCKS
#define PLANTOS_FAST_MINBLOCKS 4
#endif
#ifndef PLANTOS_FAST_WARPS
#define PLANTOS_FAST_WARPS 7          // 7 warps x 4 blocks = 28 resident warps per SM (<= 72 registers each)
#endif
constexpr int kFastWarps = PLANTOS_FAST_WARPS;
constexpr int kFastEnvs = 32;         // envs per macro tile (one lane each in the transition)
constexpr int kFastTrip = 4;          // envs per observation trip
constexpr int kFastVisRows = 5;       // nibble rows x-2 .. x+2

// type rows fetched per env: x-R .. x+R (2R+1 rows) plus one because the copy starts on an even row
__host__ __device__ constexpr int fast_type_rows(int R) { return 2 * R + 2; }
// A window buffer holds the 16-byte chunks of the 4 envs of a trip: chunk k of env j (k < TCH: type
// rows 2k, 2k+1; then the 5 nibble rows) sits at j*128 + 16k for k < 8 and at
// 512 + j*S1 + 16(k-8) otherwise (S1 = 16 * (chunks beyond 8)).  The 8 lanes that copy one env thus
// write 128 contiguous bytes per round: the cp.async shared-memory writes are bank-conflict free
// (the type-major layout before measured 9.8 wavefronts per copy instruction instead of 4).
__host__ __device__ constexpr int fast_win_s1(int R) { return (R + 1 + 5 > 8 ? R + 1 + 5 - 8 : 0) * 16; }
__host__ __device__ inline int fast_win_bytes(int R, int G) {
    const int win = 512 + kFastTrip * fast_win_s1(R);
    return win < align_up(G * 8, 16) ? align_up(G * 8, 16) : win;    // phase C borrows it as a type plane
}
// per-warp scratch: two window buffers | 4-env obs tile | 32 records | 32 actions | 32+32+32 target words
// (measured: a stride that is a multiple of 128 bytes costs 0.7 us per step -- every warp's buffers then
// start on the same banks; 5040 / 4048 / 4784 / 4144 bytes for the instantiated shapes are not)
__host__ __device__ inline int fast_warp_scratch_bytes(int R, int G, int D) {
    return 2 * fast_win_bytes(R, G) + 16 * D + kFastEnvs * (32 + 8 + 8 + 4 + 4);
}

// ---- asynchronous global->shared copies (16-byte cp.async, L2 only) ------------------------
// Records, target words and windows sit at per-env addresses, so they are fetched with per-lane
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
// window / tile accesses through 32-bit shared addresses (volatile: the data changes every trip).
// Generic-pointer accesses made the compiler rebuild the shared-window base (S2UR SR_CgaCtaId ...)
// and the per-warp scratch address from threadIdx inside the trip loop.
__device__ __forceinline__ uint64_t lds_u64_v(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32_v(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f32x4_v(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0+%1], %2;" :: "r"(a), "n"(OFF), "f"(v)); }
__device__ __forceinline__ void cp_async8(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sdst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// programmatic dependent launch (PTX griddepcontrol): see the kernel prologue
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sdst), "l"(gsrc) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// -DPLANTOS_EXP_TIMING builds an instrumented library (tools/exp_timing.py): lane 0 of every warp
// records %globaltimer at the phase boundaries of its first macro tile and parks the stamps in the
// ret fields of the terminal-record snapshot.  Never defined in the shipped build.
#ifdef PLANTOS_EXP_TIMING
__device__ __forceinline__ unsigned gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (unsigned)t; }
#define TSTAMP(k) do { if (lane == 0) ts_[k] = gtime(); } while (0)
#else
#define TSTAMP(k) do {} while (0)
#endif

template <int R, int C, bool KEEP>
__global__ void __launch_bounds__(kFastWarps * 32, PLANTOS_FAST_MINBLOCKS)
k_step_fast(const Params p, const StepIO io) {
    constexpr int D = 5 * C + 27;
    constexpr int NROW = 2 * R + 1;
    constexpr int VW = 4;                 // nibble words per visit row (G + 4 <= 32)
    constexpr int TP = R + 2;             // wall rows above the grid (== Params.TP)
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TWR = fast_type_rows(R);
    constexpr int TCH = TWR / 2;          // 16-byte chunks of type rows per env
    constexpr int NCHUNK = TCH + kFastVisRows;
    static_assert(NROW <= 16 && C <= 16, "fast kernel shape limits");
    static_assert(NCHUNK <= 16, "two copy rounds of 8 chunk lanes per env");

    extern __shared__ __align__(16) unsigned char smem[];
    typename std::conditional<KEEP, KeepMem, PlainMem>::type const mem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef PLANTOS_EXP_TIMING
    unsigned ts_[10] = {};
#endif
    TSTAMP(0);
    const int G = p.G, VE = p.VE, TS = p.TS;
    const int win_bytes = fast_win_bytes(R, G);
    unsigned char* scratch = smem + tables_bytes(G, R, C) + warp * fast_warp_scratch_bytes(R, G, D);
    float* tile = reinterpret_cast<float*>(scratch + 2 * win_bytes);
    uint4* const recb = reinterpret_cast<uint4*>(scratch + 2 * win_bytes + 16 * D);             // [32][2]
    long long* const actb = reinterpret_cast<long long*>(recb + 2 * kFastEnvs);                   // [32]
    uint64_t* const tgt_t = reinterpret_cast<uint64_t*>(actb + kFastEnvs);                        // [32]
    uint32_t* const tgt_v = reinterpret_cast<uint32_t*>(tgt_t + kFastEnvs);                       // [32]
    uint32_t* const tgt_e = tgt_v + kFastEnvs;                                                    // [32] explored-bit word (curriculum)
    auto win_buf = [&](int b) { return scratch + b * win_bytes; };

    // Every warp owns one contiguous range of Q envs (Q a multiple of 4, the same for all warps).
    // (Q = p.fast_q is computed on the host for this grid.)
    const int gwarp = blockIdx.x * kFastWarps + warp, nwarps = gridDim.x * kFastWarps;
    const int nfull = p.N & ~3;                                       // envs in whole 4-env groups
    const int wbase = min(nfull, gwarp * p.fast_q), wend = min(nfull, wbase + p.fast_q);

    // ---- prefetch helpers (all cp.async into shared memory: a prefetch holds no registers)
    // 32-bit shared-window addresses of this lane's copy destinations, computed once
    const uint32_t s_scr = smem_u32(scratch);
    const uint32_t s_rec = s_scr + 2 * win_bytes + 16 * D + 32 * lane;
    const uint32_t s_act = s_scr + 2 * win_bytes + 16 * D + 32 * kFastEnvs + 8 * lane;
    const uint32_t s_tgt_t = s_scr + 2 * win_bytes + 16 * D + 40 * kFastEnvs + 8 * lane;
    const uint32_t s_tgt_v = s_scr + 2 * win_bytes + 16 * D + 48 * kFastEnvs + 4 * lane;
    const uint32_t s_tgt_e = s_scr + 2 * win_bytes + 16 * D + 52 * kFastEnvs + 4 * lane;
    auto fetch_rec = [&](int es) {                     // records + actions of the macro tile at es
        if (lane < min(kFastEnvs, wend - es)) {
            const unsigned e = (unsigned)(es + lane);   // 32-bit element offsets (N * TS, N * VE, N * G * G < 2^32: host check)
            cp_async16(s_rec, p.rec + 2 * e);
            cp_async16(s_rec + 16, p.rec + 2 * e + 1);
            cp_async8(s_act, io.actions + e);
        }
        cp_async_commit();
    };
    auto issue_target = [&](int es) {                  // the two words the transition will look at
        if (lane < min(kFastEnvs, wend - es)) {
            const unsigned e = (unsigned)(es + lane);   // 32-bit element offsets (N * TS, N * VE, N * G * G < 2^32: host check)
            const uint32_t w0 = recb[2 * lane].x;
            EnvRec q;
            q.x = (int)(w0 & 0xff); q.y = (int)((w0 >> 8) & 0xff);
            int tx, ty; bool inb;
            action_target(q, actb[lane], G, tx, ty, inb);
            // (tx, ty) may be one cell outside the grid: wall rows / border nibbles are there
            cp_async8(s_tgt_t, p.types + e * TS + TP + tx);
            cp_async4(s_tgt_v, p.vis4 + e * VE + nib_word(tx, ty, VW));
            if (p.cur_mode && inb) cp_async4(s_tgt_e, p.expl + e * G + tx);   // this episode's explored_map row (W == 1)
        }
        cp_async_commit();
    };

    // Programmatic dependent launch: the next launch in the stream may become resident as soon as
    // every block of this one has started, and everything above plus the loads of the immutable
    // tables below may run while the previous launch is still finishing.  No env state, action or
    // output buffer is touched before griddep_wait() returns (= the previous launch has completed
    // and its writes are visible).
    griddep_launch_dependents();
    const int n16 = tables_bytes(G, R, C) >> 4;        // <= blockDim.x (checked on the host)
    uint4 tab16 = make_uint4(0, 0, 0, 0);
    if ((int)threadIdx.x < n16) tab16 = __ldg(p.table_blob + threadIdx.x);
    // ---- per-lane constants of the observation phase, precomputed by k_pack_tables:
    //   srcl[rr]  lane holding window row x+dx of LIDAR sample rr of this lane's ray
    //   shf[rr]   left shift that brings column y+dy of the window word to bits 30, 31
    //   vsrc/vsh  the two 5x5 visit cells this lane converts (q = sub, sub + 16): row lane, nibble shift
    int lt[4 * kLaneTabVec];
#pragma unroll
    for (int q = 0; q < kLaneTabVec; ++q) {
        const int4 v = __ldg(p.lane_tab + lane * kLaneTabVec + q);
        lt[4 * q] = v.x; lt[4 * q + 1] = v.y; lt[4 * q + 2] = v.z; lt[4 * q + 3] = v.w;
    }
    griddep_wait();
    TSTAMP(1);
    if (wbase < wend) fetch_rec(wbase);                // in flight while the tables are staged
    if ((int)threadIdx.x < n16) reinterpret_cast<uint4*>(smem)[threadIdx.x] = tab16;
    __syncthreads();
    const Tables t = tables_at(smem, G, R);
    const int sub = lane & 15, half = lane >> 4;
    int srcl[R], shf[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) { srcl[rr] = lt[rr]; shf[rr] = lt[8 + rr]; }
    const bool has_ray = sub < C, has_v1 = sub < 9;
    // the two 5x5 visit cells this lane converts: q = sub and q = sub + 16 -> nibble row q / 5, column q % 5
    const int vrow0 = sub / 5, vcol0 = sub % 5, vrow1 = (sub + 16) / 5 < kFastVisRows ? (sub + 16) / 5 : 0, vcol1 = (sub + 16) % 5;
    const uint32_t s_dist = smem_u32(t.dist), s_pos = smem_u32(t.pos), s_visit = smem_u32(t.visit);
    const uint32_t s_rw32 = smem_u32(t.rw32), s_rw64 = smem_u32(t.rw64);
    constexpr uint64_t LOWPAD = kObstAll & ((1ull << (2 * R)) - 1ull);
    constexpr int NCH = 2;             // envs per half-warp and trip, interleaved for ILP
    // Window copy roles: 8 lanes per env of the trip, lane c8 serves chunks c8 and c8 + 8 of env
    // cj (chunks 0 .. TCH-1 are type rows, TCH .. NCHUNK-1 nibble rows; the second round is always
    // a nibble row because TCH <= 8).  Destination offsets inside a window buffer and the source
    // element offsets relative to the env's first fetched row are per-lane constants.
    const int cj = lane >> 3, c8 = lane & 7;
    constexpr int S1 = fast_win_s1(R);
    const bool cp0_type = c8 < TCH, cp0_on = c8 < NCHUNK, cp1_on = c8 + 8 < NCHUNK;
    const uint32_t cp0_dst = s_scr + cj * 128 + 16 * c8;
    const uint32_t cp1_dst = s_scr + 512 + cj * S1 + 16 * c8;
    // byte offset inside a window buffer of nibble row `sub` (chunk TCH + sub) of the two envs this
    // half-warp handles in a trip (env 2c + half)
    // shared address of the nibble row (chunk TCH + row) each of the lane's two cells lives in, per chain
    // (env 2c + half); cells are read straight out of the window: no row slices, no shuffles
    uint32_t s_vc0[2], s_vc1[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int j = 2 * c + (lane >> 4), k0 = TCH + vrow0, k1 = TCH + vrow1;
        s_vc0[c] = s_scr + (k0 < 8 ? j * 128 + 16 * k0 : 512 + j * S1 + 16 * (k0 - 8));
        s_vc1[c] = s_scr + (k1 < 8 ? j * 128 + 16 * k1 : 512 + j * S1 + 16 * (k1 - 8));
    }
    // 32-bit shared addresses of this lane's type row slot (chain 0, even start row; chain 1 is 256
    // bytes on), of its ray's five floats and of its visit cell in tile row `half` (chain 1 is two
    // rows = 8 * D bytes on), and of its float4 of the flush
    const uint32_t s_trow = s_scr + half * 128 + sub * 8;
    const uint32_t s_tile = s_scr + 2 * win_bytes;
    const uint32_t s_ray = s_tile + (half * D + 5 * sub) * 4;
    const uint32_t s_cell = s_tile + (half * D + 5 * C + 2 + sub) * 4;
    const uint32_t s_flush = s_tile + 16 * lane;

    TSTAMP(2);
    if (wbase < wend) {
        cp_async_wait_all();
        __syncwarp();
        TSTAMP(3);
        issue_target(wbase);
    }

    for (int e0 = wbase; e0 < wend; e0 += kFastEnvs) {
        const int e_next = e0 + kFastEnvs;
        const bool has_next = e_next < wend;
        const int ts = min(kFastEnvs, wend - e0);          // envs in this macro tile (a multiple of 4)
        const bool act = lane < ts;
        const unsigned e = (unsigned)(e0 + lane);
        cp_async_wait_all();                              // records, actions and target words are here
        __syncwarp();
        TSTAMP(4);

        // ---- phase A: transition, one lane per env
        int done = 0, term = 0, trunc = 0;
        unsigned posw = 0;
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
            // CurriculumWrapper (plantos_set_curriculum): visit counts outlive the episode, so "new cell"
            // for the exploration percentage comes from this episode's explored bits instead
            int expl_fresh = -1;
            if (p.cur_mode && inb && action < 4) {
                const uint32_t ew = tgt_e[lane], bit = 1u << (ty & 31);
                expl_fresh = (ew & bit) ? 0 : 1;
                if (t_cell != kObstacle) p.expl[e * G + tx] = ew | bit;
            }
            StepOut o = transition_core(r, action, tx, ty, t_cell, (vword >> sh) & 15u, p.max_steps, expl_fresh);
            if (p.cur_mode) {                                  // CurriculumWrapper.step, A2C_training.py:94-100
                const double pct = ((double)r.explored / (double)r.total_free) * 100.0;
                if (pct >= p.cur_thr[e]) {
                    p.cur_cnt[e].y |= 1;
                    if (p.cur_mode == 1) o.terminated = 1;
                }
            }
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
            posw = (unsigned)r.x | ((unsigned)r.y << 5);
        }
        accumulate_stats(p, act && done, r, term, trunc, lane, (int)e);
        // orders the plane stores above before the window copies that other lanes issue below, and
        // frees the record buffer
        __syncwarp();
        TSTAMP(5);
        if (has_next) fetch_rec(e_next);

        // Window copy of the trip starting at env e0 + base into buffer wb: grid rows x-R .. x+R
        // are padded rows x+2 .. x+2R+2; start on the even row at or just below x+2 so that every
        // chunk is 16-byte aligned.  Nibble rows x-2 .. x+2 are padded rows x+1 .. x+5.
        auto issue_win = [&](int base, int wb) {
            // (32-bit element offsets: N * TS and N * VE stay far below 2^32 for any N that fits HBM)
            const unsigned x = __shfl_sync(FULL, posw, base + cj) & 31u;
            const unsigned ej = (unsigned)(e0 + base + cj);
            const uint64_t* tsrc = p.types + (ej * (unsigned)TS + ((x + 2u) & ~1u) + 2u * (unsigned)c8);
            const unsigned v0 = ej * (unsigned)VE + (x + 1u) * VW;              // nibble row x-2
            const uint32_t* vsrc0 = p.vis4 + (v0 + (unsigned)(c8 > TCH ? c8 - TCH : 0) * VW);   // round 0: row c8 - TCH
            const uint32_t* vsrc1 = p.vis4 + (v0 + (unsigned)(c8 + 8 - TCH) * VW);              // round 1: row c8 + 8 - TCH
            const uint32_t boff = wb * win_bytes;
            if (cp0_on) cp_async16(cp0_dst + boff, cp0_type ? (const void*)tsrc : (const void*)vsrc0);
            if (cp1_on) cp_async16(cp1_dst + boff, vsrc1);
        };
        issue_win(0, 0);
        cp_async_commit();

        // ---- phase B: observations, one trip = four rows: two independent chains per half-warp,
        // written stage by stage so that their latencies overlap
        float4* const obs4 = reinterpret_cast<float4*>(io.obs) + (size_t)(e0 >> 2) * D;
        int trip = 0;
#pragma unroll 1
        for (int base = 0; base < ts; base += kFastTrip, ++trip) {
            const int wb = trip & 1;
            // one commit group per trip (empty on the last one): after wait_group<1> everything
            // but the copies just issued has landed -- this trip's windows, and from the second
            // trip on the next macro tile's target words
            if (base + kFastTrip < ts) issue_win(base + kFastTrip, wb ^ 1);
            cp_async_commit();
            cp_async_wait_group<1>();
            __syncwarp();
            if (trip == 0) TSTAMP(6);
            if (trip == 0 && has_next) issue_target(e_next);     // its records landed with this trip's windows
            const uint32_t wboff = wb * win_bytes;

            int x[NCH], y[NCH], tb[NCH];
            unsigned w[NCH], acc[NCH], s0[NCH], s1[NCH];
            uint64_t trow[NCH];
            unsigned n0[NCH], n1[NCH];
            // stage 1: positions of the two envs this half-warp handles (env base + 2c + half)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const unsigned pw = __shfl_sync(FULL, posw, base + 2 * c + half);
                x[c] = pw & 31; y[c] = pw >> 5;
                // first needed type row inside the fetched window: padded row x+2 minus the even start
                tb[c] = (x[c] & 1) * 8;                         // byte offset of the first needed row
            }
            // stage 2: shared-memory reads: this lane's type row and visit-nibble words
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                // (every lane loads: lanes beyond the 2R+1 type rows / the 5 nibble rows read other
                // parts of the warp's scratch and nobody asks them for the result, which is cheaper
                // than branching around the loads)
                trow[c] = lds_u64_v(s_trow + c * 256 + wboff + tb[c]);
                // grid column y-2+col sits at nibble y+col of the padded row: word (y+col)>>3
                n0[c] = (unsigned)y[c] + vcol0; n1[c] = (unsigned)y[c] + vcol1;
                s0[c] = lds_u32_v(s_vc0[c] + wboff + ((n0[c] >> 3) << 2));
                s1[c] = lds_u32_v(s_vc1[c] + wboff + ((n1[c] >> 3) << 2));
            }
            // stage 3: rover-centred window word (cells y-R .. y+R of this lane's row, walls
            // outside) 
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int sft = 2 * y[c];
                const uint64_t ext = (trow[c] << (2 * R)) | LOWPAD;
                w[c] = (unsigned)((ext >> sft) | ((kObstAll << 1) << (63 - sft)));
                acc[c] = 0;
            }
            // stage 4: LIDAR march (plantos_env.py:260-284): sample rr looks at window row
            // srcl[rr], bits shf[rr].  Far sample first: each step shifts the accumulator left by
            // one cell and funnels the sample's two bits in from the top of the aligned row word,
            // so that sample rr ends up at bits 2rr, 2rr+1.
#pragma unroll
            for (int rr = R - 1; rr >= 0; --rr) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const unsigned wr = __shfl_sync(FULL, w[c], srcl[rr]);
                    acc[c] = __funnelshift_l(wr << shf[rr], acc[c], 2);
                }
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
                // one-hot of the hit kind by compares: a 128-bit table read would cost four
                // shared-memory wavefronts, and this phase is bound by those, not by issue slots
                oh[c].x = kind == 0 ? 1.0f : 0.0f; oh[c].y = kind == 1 ? 1.0f : 0.0f;
                oh[c].z = kind == 2 ? 1.0f : 0.0f; oh[c].w = kind == 3 ? 1.0f : 0.0f;
                fp[c] = lds_f32(s_pos + 4 * (sub ? y[c] : x[c]));
                // 4 * nibble in two instructions: rotate the nibble to bits 2..5, mask
                fv0[c] = lds_f32(s_visit + (__funnelshift_r(s0[c], s0[c], (4 * n0[c] + 30) & 31) & 0x3cu));
                fv1[c] = lds_f32(s_visit + (__funnelshift_r(s1[c], s1[c], (4 * n1[c] + 30) & 31) & 0x3cu));
            }
            // stage 6: stores into the tile (row 2c + half)
            if (has_ray) {                                    // :286-292
                sts_f32<0>(s_ray, fd[0]); sts_f32<4>(s_ray, oh[0].x); sts_f32<8>(s_ray, oh[0].y);
                sts_f32<12>(s_ray, oh[0].z); sts_f32<16>(s_ray, oh[0].w);
                sts_f32<8 * D>(s_ray, fd[1]); sts_f32<8 * D + 4>(s_ray, oh[1].x); sts_f32<8 * D + 8>(s_ray, oh[1].y);
                sts_f32<8 * D + 12>(s_ray, oh[1].z); sts_f32<8 * D + 16>(s_ray, oh[1].w);
            }
            if (sub < 2) { sts_f32<-8>(s_cell, fp[0]); sts_f32<8 * D - 8>(s_cell, fp[1]); }      // :294-296
            sts_f32<0>(s_cell, fv0[0]); sts_f32<8 * D>(s_cell, fv0[1]);                           // :298-313
            if (has_v1) { sts_f32<64>(s_cell, fv1[0]); sts_f32<8 * D + 64>(s_cell, fv1[1]); }
            // flush four env rows = D float4, 16-byte aligned because e0 and base are multiples of 4
            __syncwarp();
            float4* dst4 = obs4 + (size_t)(base >> 2) * D;
#pragma unroll
            for (int q = 0; q < (D + 31) / 32; ++q) {
                const int idx = q * 32 + lane;
                if (idx < D) __stcs(dst4 + idx, lds_f32x4_v(s_flush + 512 * q));
            }
            __syncwarp();   // the tile and this trip's window buffer may be overwritten now
            if (trip == 0) TSTAMP(7);
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

        // ---- phase C: auto-reset of finished envs (rare; warp-cooperative generic code; window
        // buffer 0 is free now and serves as the type-plane scratch)
        unsigned dmask = __ballot_sync(FULL, act && done);
        uint64_t* plane = reinterpret_cast<uint64_t*>(win_buf(0));
        const bool deferred = p.map_source == 2;             // maze handles: k_reset_done starts the new episodes after this launch
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
            if (deferred) continue;
            int keep = 0, map_ep = -1;
            if (p.cur_mode) {                                // CurriculumWrapper.reset
                int cr = 0;
                if (lane == 0) cr = curriculum_on_reset(p, (int)ej, episode);
                cr = __shfl_sync(FULL, cr, 0);
                keep = cr & 1; map_ep = cr >> 1;
            }
            const EnvRec nr = reset_env_warp<false>(p, (int)ej, episode, plane, lane, keep != 0, map_ep);
            build_obs_warp(p, t, plane, vis_e, nr.x, nr.y, tile, lane, keep != 0);
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
        for (int e = nfull; e < p.N; ++e) step_env_warp<false>(p, t, io, e, reinterpret_cast<uint64_t*>(win_buf(0)), tile, lane);
}

}  // namespace plantos_dev
