// Microbenchmark (diagnostics, not part of the library): how fast can one persistent CTA per SM
// stream HBM -> shared memory with 1-D bulk async copies, as a function of stage size / depth /
// pieces per stage; and a plain LDG.128 streaming kernel for comparison.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni DONE;\nbra.uni LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// stages of `stage_bytes`, each filled by `pieces` bulk copies; consumer = one warp that waits and frees
__global__ void __launch_bounds__(64, 1) tma_stream(const unsigned char* src, int64_t total_tiles, int stage_bytes, int stages, int pieces, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = (unsigned long long*)(smem + (size_t)stage_bytes * stages);
    unsigned long long* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int64_t first = blockIdx.x, step = gridDim.x;
    const int n_my = first < total_tiles ? (int)((total_tiles - first + step - 1) / step) : 0;
    if (threadIdx.x == 0) {
        for (int it = 0; it < n_my; ++it) {
            int s = it % stages; uint32_t round = it / stages;
            if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
            const unsigned char* p = src + (first + (int64_t)it * step) * stage_bytes;
            mbar_expect_tx(&full[s], stage_bytes);
            int pb = stage_bytes / pieces;
            for (int k = 0; k < pieces; ++k) bulk_g2s(smem + (size_t)s * stage_bytes + k * pb, p + k * pb, pb, &full[s]);
        }
    } else if (threadIdx.x == 32) {
        float acc = 0;
        for (int it = 0; it < n_my; ++it) {
            int s = it % stages; uint32_t par = (it / stages) & 1;
            mbar_wait(&full[s], par);
            acc += *(float*)(smem + (size_t)s * stage_bytes);
            mbar_arrive(&empty[s]);
        }
        if (acc == 123.456f) *sink = acc;
    }
}

__global__ void __launch_bounds__(256) ldg_stream(const float4* __restrict__ src, int64_t n4, float* sink, int unroll_dummy) {
    float acc = 0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(src + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(src + i + 3 * stride));
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const size_t bytes = (size_t)2 << 30;
    unsigned char* src; float* sink;
    CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(src, 1, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    struct Cfg { int stage_kb, stages, pieces, ctas_per_sm; } cfgs[] = {
        {64, 3, 1, 1}, {64, 3, 4, 1}, {64, 3, 16, 1}, {32, 6, 1, 1}, {32, 6, 4, 1}, {16, 12, 1, 1}, {16, 12, 4, 1}, {8, 24, 1, 1},
        {4, 48, 1, 1}, {32, 3, 1, 2}, {16, 6, 1, 2}, {16, 3, 1, 4}, {8, 6, 1, 4}, {64, 2, 1, 1}, {32, 2, 1, 1}, {32, 4, 1, 1}, {16, 4, 1, 1}, {16, 8, 1, 1},
    };
    CK(cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (auto c : cfgs) {
        int stage_bytes = c.stage_kb * 1024;
        size_t smem = (size_t)stage_bytes * c.stages + 16 * c.stages + 128;
        int64_t tiles = bytes / stage_bytes;
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            tma_stream<<<sms * c.ctas_per_sm, 64, smem>>>(src, tiles, stage_bytes, c.stages, c.pieces, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("TMA stage %3d KB x %2d stages, %2d pieces, %d CTA/SM (%3zu KB in flight/SM): %7.1f GB/s\n", c.stage_kb, c.stages, c.pieces,
               c.ctas_per_sm, (size_t)c.stage_kb * c.stages * c.ctas_per_sm, bytes / best / 1e6);
    }
    for (int bps : {2, 4, 6, 8}) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            ldg_stream<<<sms * bps, 256>>>((const float4*)src, bytes / 16, sink, 0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("LDG.128 x4 unrolled, %d CTAs/SM x 256 thr: %7.1f GB/s\n", bps, bytes / best / 1e6);
    }
    return 0;
}
