// ubench_decode.cu -- micro-benchmark of the observation OUTPUT stage in isolation (round 2 design study).
//
// Question it answers: how fast can one B200 expand a per-env byte code (one byte per observation float,
// the byte = 4 * index into a small float table) into the coalesced [N, 107] fp32 observation buffer?
// This is the output half of the lane-per-env step kernel (k_step_tile): codes sit in shared memory as a
// flat byte image of a 32-env tile (32 * 107 B = 856 words), every lane of the warp expands one float4
// per iteration (1 LDS.32 code word, 4 table LDS, 1 STG.128).  Variants:
//   mode 0  pure fill (STG.128 of constants)                     -> the write floor
//   mode 1  code bytes from global (coalesced) -> smem -> LUT decode -> STG.128.cs
//   mode 2  like 1 but the table lives in registers-selectable form (one-hot by arithmetic, LUT only for non 0/1)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_decode tools/ubench_decode.cu
// run:   tools/ubench_decode [N=131072] [warps/block=16] [blocks/SM=1] [iters=200]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int D = 107;
constexpr int TILE = 32;
constexpr int TILE_WORDS = TILE * D / 4;   // 856

__device__ __forceinline__ uint32_t smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

template <int MODE>
__global__ void __launch_bounds__(512) k_decode(const uint32_t* __restrict__ codes, float4* __restrict__ obs, int ntiles,
                                               const float* __restrict__ lut_g) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* lut = reinterpret_cast<float*>(smem);                 // 64 floats
    uint32_t* code = reinterpret_cast<uint32_t*>(smem + 256) + warp * TILE_WORDS;
    if (threadIdx.x < 64) lut[threadIdx.x] = lut_g[threadIdx.x];
    __syncthreads();
    const uint32_t s_lut = smem_u32(lut), s_code = smem_u32(code) + 4 * lane;
    const int nwarps = gridDim.x * nw;
    for (int t = warp * gridDim.x + blockIdx.x; t < ntiles; t += nwarps) {
        float4* dst = obs + (size_t)t * TILE_WORDS + lane;
        if (MODE == 0) {
#pragma unroll 9
            for (int it = 0; it < 27; ++it)
                if (it * 32 + lane < TILE_WORDS) __stcs(dst + it * 32, make_float4(0.f, 1.f, 0.f, 0.5f));
            continue;
        }
        const uint32_t* src = codes + (size_t)t * TILE_WORDS + lane;
#pragma unroll
        for (int it = 0; it < 27; ++it)
            if (it * 32 + lane < TILE_WORDS) code[it * 32 + lane] = __ldg(src + it * 32);
        __syncwarp();
#pragma unroll 9
        for (int it = 0; it < 27; ++it) {
            if (it * 32 + lane < TILE_WORDS) {
                const uint32_t cw = lds_u32(s_code + 128 * it);
                float4 f;
                if (MODE == 1) {
                    f.x = lds_f32(s_lut + (cw & 0xffu));
                    f.y = lds_f32(s_lut + ((cw >> 8) & 0xffu));
                    f.z = lds_f32(s_lut + ((cw >> 16) & 0xffu));
                    f.w = lds_f32(s_lut + (cw >> 24));
                } else {
                    // bytes 0 / 4 are 0.0 / 1.0 without a table read; everything else through the table
                    uint32_t b0 = cw & 0xffu, b1 = (cw >> 8) & 0xffu, b2 = (cw >> 16) & 0xffu, b3 = cw >> 24;
                    f.x = b0 <= 4u ? __uint_as_float(b0 * 0x0fe00000u) : lds_f32(s_lut + b0);
                    f.y = b1 <= 4u ? __uint_as_float(b1 * 0x0fe00000u) : lds_f32(s_lut + b1);
                    f.z = b2 <= 4u ? __uint_as_float(b2 * 0x0fe00000u) : lds_f32(s_lut + b2);
                    f.w = b3 <= 4u ? __uint_as_float(b3 * 0x0fe00000u) : lds_f32(s_lut + b3);
                }
                __stcs(dst + it * 32, f);
            }
        }
        __syncwarp();
    }
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 131072;
    const int warps = argc > 2 ? atoi(argv[2]) : 16;
    const int bps = argc > 3 ? atoi(argv[3]) : 1;
    const int iters = argc > 4 ? atoi(argv[4]) : 200;
    const int ntiles = N / TILE;
    const int RING = 5;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * bps;
    // realistic code bytes: per env 16 rays x (dist, one-hot 4) + 2 pos + 25 visits
    std::vector<uint8_t> h((size_t)N * D);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (int e = 0; e < N; ++e) {
        uint8_t* r = &h[(size_t)e * D];
        for (int i = 0; i < 16; ++i) {
            const int dist = rnd() % 7, kind = rnd() % 4;
            r[5 * i] = 4 * (2 + dist);
            for (int k = 0; k < 4; ++k) r[5 * i + 1 + k] = (k == kind) ? 4 : 0;
        }
        r[80] = 4 * (27 + rnd() % 25); r[81] = 4 * (27 + rnd() % 25);
        for (int k = 0; k < 25; ++k) r[82 + k] = 4 * (10 + (rnd() % 8 < 5 ? 0 : rnd() % 11));
    }
    std::vector<float> lut(64);
    lut[0] = 0.f; lut[1] = 1.f;
    for (int i = 0; i < 8; ++i) lut[2 + i] = (float)((double)(i < 6 ? i : 6) / 6.0);
    for (int i = 0; i < 16; ++i) lut[10 + i] = (float)((double)(i < 10 ? i : 10) / 10.0);
    for (int i = 0; i < 25; ++i) lut[27 + i] = (float)((double)i / 25.0);
    uint32_t* d_codes; float* d_lut; float4* d_obs[RING];
    CK(cudaMalloc(&d_codes, (size_t)N * D)); CK(cudaMalloc(&d_lut, 256));
    CK(cudaMemcpy(d_codes, h.data(), (size_t)N * D, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_lut, lut.data(), 256, cudaMemcpyHostToDevice));
    for (int i = 0; i < RING; ++i) CK(cudaMalloc(&d_obs[i], (size_t)N * D * 4));
    const int smem = 256 + warps * TILE_WORDS * 4;
    CK(cudaFuncSetAttribute(k_decode<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_decode<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_decode<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int mode = 0; mode < 3; ++mode) {
        auto launch = [&](int i) {
            if (mode == 0) k_decode<0><<<grid, warps * 32, smem>>>(d_codes, d_obs[i % RING], ntiles, d_lut);
            if (mode == 1) k_decode<1><<<grid, warps * 32, smem>>>(d_codes, d_obs[i % RING], ntiles, d_lut);
            if (mode == 2) k_decode<2><<<grid, warps * 32, smem>>>(d_codes, d_obs[i % RING], ntiles, d_lut);
        };
        for (int i = 0; i < 20; ++i) launch(i);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        for (int i = 0; i < iters; ++i) launch(i);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        const double us = ms * 1e3 / iters;
        printf("mode %d  N %d  grid %d x %d warps  %.2f us/launch  %.0f GB/s written\n", mode, N, grid, warps, us,
               (double)N * D * 4 / us * 1e-3);
    }
    // spot check of mode 1 against the host expansion
    k_decode<1><<<grid, warps * 32, smem>>>(d_codes, d_obs[0], ntiles, d_lut);
    std::vector<float> out((size_t)N * D);
    CK(cudaMemcpy(out.data(), d_obs[0], (size_t)N * D * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < out.size(); ++i) bad += out[i] != lut[h[i] >> 2];
    printf("mode 1 check: %zu mismatches of %zu\n", bad, out.size());
    return bad != 0;
}
