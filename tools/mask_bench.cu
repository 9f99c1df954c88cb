// Micro-benchmark for the spoke mask kernel: isolates what limits it below the streaming-read ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mask_bench tools/mask_bench.cu
// Variants (all read the same random float32 buffer of N bytes, larger than L2):
//   lin        : whole-grid linear sweep, LDG.128 x4, read only (ceiling)
//   warp-tile  : one warp per 16 KiB tile (8 loads in flight per lane), read only / +compute / +mask stores
//   cta-tile   : one CTA per 32 KiB..128 KiB super tile, warps interleaved at 512 B, read only / +compute / +stores
//   tma-ring   : per-warp TMA ring (8 KiB stages), read only / +compute / +stores
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned set_gt(float a, float b) { unsigned r; asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
struct LaneBits { unsigned b0, b1, b2, b3; };
__device__ __forceinline__ LaneBits lane_bits(unsigned lane) { unsigned sh = 4u * (lane & 7u); return LaneBits{1u << sh, 2u << sh, 4u << sh, 8u << sh}; }
__device__ __forceinline__ unsigned nibble(const float4& v, float thr, const LaneBits& lb) {
    return (set_gt(v.x, thr) & lb.b0) | (set_gt(v.y, thr) & lb.b1) | (set_gt(v.z, thr) & lb.b2) | (set_gt(v.w, thr) & lb.b3);
}
__device__ __forceinline__ unsigned butterfly8(const unsigned (&nib)[8], unsigned lane) {
    unsigned a4[4], a2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bool hi = lane & 1u; unsigned keep = hi ? nib[2*i+1] : nib[2*i], give = hi ? nib[2*i] : nib[2*i+1]; a4[i] = keep | __shfl_xor_sync(~0u, give, 1); }
#pragma unroll
    for (int i = 0; i < 2; ++i) { bool hi = lane & 2u; unsigned keep = hi ? a4[2*i+1] : a4[2*i], give = hi ? a4[2*i] : a4[2*i+1]; a2[i] = keep | __shfl_xor_sync(~0u, give, 2); }
    bool hi = lane & 4u; unsigned keep = hi ? a2[1] : a2[0], give = hi ? a2[0] : a2[1];
    return keep | __shfl_xor_sync(~0u, give, 4);
}

// ---- lin -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lin(const float4* __restrict__ src, int64_t n4, float* sink) {
    float acc = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = ld4((const float*)(src + i)), b = ld4((const float*)(src + i + stride)), c = ld4((const float*)(src + i + 2 * stride)), d = ld4((const float*)(src + i + 3 * stride));
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) *sink = acc;
}

// ---- warp-tile: MODE 0 read only, 1 +compute (count only), 2 +mask stores ---------------------------
template <int MODE, bool DYNAMIC>
__global__ void __launch_bounds__(256, 4) k_warp_tile(const float* __restrict__ src, int64_t tiles, float thr, uint32_t* __restrict__ mask,
                                                      uint32_t* __restrict__ tile_count, unsigned* ticket, float* sink) {
    const unsigned lane = threadIdx.x & 31u;
    const LaneBits lb = lane_bits(lane);
    const unsigned word_slot = (lane & 7u) * 4u + (lane >> 3);
    long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * 8;
    float acc = 0;
    while (tile < tiles) {
        unsigned nxt = 0;
        if (DYNAMIC && lane == 0) nxt = atomicAdd(ticket, 1u);
        const float* p = src + tile * 4096;
        uint32_t* mw = mask + tile * 128;
        unsigned cnt = 0;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ld4(p + b * 1024 + k * 128 + lane * 4);
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += v[k].x;
            } else {
                unsigned nib[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) nib[k] = nibble(v[k], thr, lb);
                unsigned word = butterfly8(nib, lane);
                cnt += __popc(word);
                if (MODE == 2) mw[b * 32 + word_slot] = word;
            }
        }
        if (MODE >= 1) { cnt = __reduce_add_sync(~0u, cnt); if (lane == 0) tile_count[tile] = cnt; }
        tile = DYNAMIC ? nw + (long long)__shfl_sync(~0u, nxt, 0) : tile + nw;
    }
    if (acc == 123.456f) *sink = acc;
}

// ---- cta-tile: a CTA of 8 warps sweeps a super tile of SUPER*4096 cells; at every step the 8 warps read
// 8 consecutive 512-byte rows x 8 loads each = 32 KiB contiguous. Per-tile counts via shared atomics. ----
template <int MODE, int SUPER>
__global__ void __launch_bounds__(256, 4) k_cta_tile(const float* __restrict__ src, int64_t supers, float thr, uint32_t* __restrict__ mask,
                                                     uint32_t* __restrict__ tile_count, unsigned* ticket, float* sink) {
    __shared__ unsigned s_next;
    __shared__ unsigned s_cnt[SUPER];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const LaneBits lb = lane_bits(lane);
    float acc = 0;
    long long sup = blockIdx.x;
    while (sup < supers) {
        if (threadIdx.x == 0) s_next = atomicAdd(ticket, 1u);
        if (threadIdx.x < SUPER) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const float* p = src + sup * (SUPER * 4096);
        uint32_t* mw = mask + sup * (SUPER * 128);
        // step = 8192 cells (32 KiB): warp w reads chunks (k*8 + w), k = 0..7, of the step: chunk = 128 cells
#pragma unroll 1
        for (int st = 0; st < SUPER / 2; ++st) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ld4(p + st * 8192 + (k * 8 + warp) * 128 + lane * 4);
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += v[k].x;
            } else {
                unsigned nib[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) nib[k] = nibble(v[k], thr, lb);
                unsigned word = butterfly8(nib, lane);       // lane holds word (lane>>3) of chunk k = lane & 7
                unsigned c = __popc(word);
                // chunk (k*8+warp) -> words 4*(k*8+warp) + (lane>>3) of the step
                if (MODE == 2) mw[st * 256 + ((lane & 7u) * 8u + warp) * 4u + (lane >> 3)] = word;
                // tile of the chunk inside the step: chunks 0..31 -> first tile, 32..63 -> second
                unsigned lo = __reduce_add_sync(~0u, (lane & 7u) < 4u ? c : 0u), hi = __reduce_add_sync(~0u, (lane & 7u) >= 4u ? c : 0u);
                if (lane == 0) { atomicAdd(&s_cnt[st * 2], lo); atomicAdd(&s_cnt[st * 2 + 1], hi); }
            }
        }
        __syncthreads();
        if (MODE >= 1 && threadIdx.x < SUPER) tile_count[sup * SUPER + threadIdx.x] = s_cnt[threadIdx.x];
        sup = (long long)gridDim.x + s_next;
        __syncthreads();
    }
    if (acc == 123.456f) *sink = acc;
}

// ---- tma-ring ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni D1;\nbra.uni W1;\nD1:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// CTA-cooperative TMA ring: one producer thread streams STAGE_KB stages of consecutive memory (a CTA ticket =
// one stage); the NW consumer warps split every stage in 512-byte rows (warp w takes chunks w, w+NW, ...).
template <int MODE, int NW, int STAGE_KB, int STAGES>
__global__ void __launch_bounds__(NW * 32 + 32, 1) k_tma_cta(const float* __restrict__ src, int64_t n_stages, float thr, uint32_t* __restrict__ mask,
                                                             uint32_t* __restrict__ tile_count, unsigned* ticket, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int STAGE_CELLS = STAGE_KB * 256;
    constexpr int TILES = STAGE_CELLS / 4096;
    float* ring = reinterpret_cast<float*>(smem);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)STAGES * STAGE_KB * 1024);
    unsigned long long* empty = full + STAGES;
    unsigned* s_tile = reinterpret_cast<unsigned*>(empty + STAGES);      // [STAGES] stage id held by the slot
    unsigned* s_cnt = s_tile + STAGES;                                   // [STAGES][TILES]
    const unsigned lane = threadIdx.x & 31u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(smem_u32(full + i), 1); mbar_init(smem_u32(empty + i), NW); }
        for (int i = 0; i < STAGES * TILES; ++i) s_cnt[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == NW) {
        if (lane == 0) {
            long long st = blockIdx.x;
            for (int it = 0;; ++it) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(smem_u32(empty + s), ((it / STAGES) - 1) & 1);
                if (st >= n_stages) { s_tile[s] = 0xffffffffu; asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect_tx(smem_u32(full + s), 0); break; }
                s_tile[s] = (unsigned)st;
                mbar_expect_tx(smem_u32(full + s), STAGE_KB * 1024);
                bulk_g2s(smem_u32(ring + (size_t)s * STAGE_CELLS), src + st * STAGE_CELLS, STAGE_KB * 1024, smem_u32(full + s));
                st = (long long)gridDim.x + atomicAdd(ticket, 1u);
            }
        }
        return;
    }
    const LaneBits lb = lane_bits(lane);
    float acc = 0;
    for (int it = 0;; ++it) {
        const int s = it % STAGES;
        mbar_wait(smem_u32(full + s), (it / STAGES) & 1);
        const unsigned st = *(volatile unsigned*)(s_tile + s);
        if (st == 0xffffffffu) break;
        const float4* b4 = reinterpret_cast<const float4*>(ring + (size_t)s * STAGE_CELLS);
        uint32_t* mw = mask + (size_t)st * (STAGE_CELLS / 32);
        // chunks of the stage: STAGE_CELLS/128; warp handles chunk (j*NW + warp); batches of 8 chunks
        constexpr int CHUNKS = STAGE_CELLS / 128;
        constexpr int PER_WARP = CHUNKS / NW;
#pragma unroll 1
        for (int j0 = 0; j0 < PER_WARP; j0 += 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = b4[((j0 + k) * NW + warp) * 32 + lane];
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += v[k].x;
            } else {
                unsigned nib[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) nib[k] = nibble(v[k], thr, lb);
                unsigned word = butterfly8(nib, lane);
                unsigned c = __popc(word);
                const unsigned chunk = (j0 + (lane & 7u)) * NW + warp;
                if (MODE == 2) mw[chunk * 4u + (lane >> 3)] = word;
                // tile of the chunk = chunk / 32
                if (TILES == 1) { c = __reduce_add_sync(~0u, c); if (lane == 0) atomicAdd(&s_cnt[s * TILES], c); }
                else atomicAdd(&s_cnt[s * TILES + chunk / 32], c);
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
        (void)tile_count;
    }
    if (acc == 123.456f) *sink = acc;
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)1536 << 20;       // 1.5 GiB of echo
    const int64_t cells = bytes / 4, tiles = cells / 4096;
    float* src; float* sink; uint32_t* mask; uint32_t* tcount; unsigned* ticket;
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&mask, bytes / 32)); CK(cudaMalloc(&tcount, tiles * 4)); CK(cudaMalloc(&ticket, 256));
    {   // random integers 0..255 as float (like the synthetic echoes), ~1 % above threshold 10
        std::vector<float> h(cells);
        uint32_t x = 12345;
        for (int64_t i = 0; i < cells; ++i) { x = x * 1664525u + 1013904223u; unsigned r = x >> 8; h[i] = (r % 100 == 0) ? (float)(11 + r % 200) : (float)(r % 10); }
        CK(cudaMemcpy(src, h.data(), bytes, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto timeit = [&](const char* name, auto&& launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemset(ticket, 0, 256));
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-58s %8.1f us  %7.1f GB/s\n", name, best * 1e3, bytes / best / 1e6);
    };
    const float thr = 10.f;
    for (int bps : {2, 4, 8}) { char n[96]; snprintf(n, 96, "lin LDG.128x4 read-only, %d CTA/SM", bps); timeit(n, [&] { k_lin<<<sms * bps, 256>>>((const float4*)src, bytes / 16, sink); }); }
    timeit("warp-tile static  read-only", [&] { k_warp_tile<0, false><<<sms * 4, 256>>>(src, tiles, thr, mask, tcount, ticket, sink); });
    timeit("warp-tile dynamic read-only", [&] { k_warp_tile<0, true><<<sms * 4, 256>>>(src, tiles, thr, mask, tcount, ticket, sink); });
    timeit("warp-tile dynamic +compute", [&] { k_warp_tile<1, true><<<sms * 4, 256>>>(src, tiles, thr, mask, tcount, ticket, sink); });
    timeit("warp-tile dynamic +compute +mask stores", [&] { k_warp_tile<2, true><<<sms * 4, 256>>>(src, tiles, thr, mask, tcount, ticket, sink); });
    timeit("cta-tile SUPER=8 (128 KiB) read-only", [&] { k_cta_tile<0, 8><<<sms * 4, 256>>>(src, tiles / 8, thr, mask, tcount, ticket, sink); });
    timeit("cta-tile SUPER=8 +compute", [&] { k_cta_tile<1, 8><<<sms * 4, 256>>>(src, tiles / 8, thr, mask, tcount, ticket, sink); });
    timeit("cta-tile SUPER=8 +compute +mask stores", [&] { k_cta_tile<2, 8><<<sms * 4, 256>>>(src, tiles / 8, thr, mask, tcount, ticket, sink); });
    timeit("cta-tile SUPER=2 (32 KiB) +compute +mask stores", [&] { k_cta_tile<2, 2><<<sms * 4, 256>>>(src, tiles / 2, thr, mask, tcount, ticket, sink); });
#define TMA_CASE(MODE, NW, KB, ST, label)                                                                                      \
    {                                                                                                                          \
        auto kern = k_tma_cta<MODE, NW, KB, ST>;                                                                               \
        int smem = ST * KB * 1024 + ST * 16 + ST * 4 + ST * (KB / 16 > 0 ? KB / 16 : 1) * 4 + 64;                              \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                                     \
        timeit(label, [&] { kern<<<sms, NW * 32 + 32, smem>>>(src, (int64_t)(bytes / (KB * 1024)), thr, mask, tcount, ticket, sink); }); \
    }
    TMA_CASE(0, 8, 64, 3, "tma-cta 64 KiB x3, 8 warps, read-only");
    TMA_CASE(1, 8, 64, 3, "tma-cta 64 KiB x3, 8 warps, +compute");
    TMA_CASE(2, 8, 64, 3, "tma-cta 64 KiB x3, 8 warps, +compute +mask stores");
    TMA_CASE(2, 16, 64, 3, "tma-cta 64 KiB x3, 16 warps, +compute +mask stores");
    TMA_CASE(2, 8, 32, 6, "tma-cta 32 KiB x6, 8 warps, +compute +mask stores");
    TMA_CASE(2, 16, 32, 6, "tma-cta 32 KiB x6, 16 warps, +compute +mask stores");
    return 0;
}
