// Device side of the time-sharded driver's bookkeeping (sharded.py; SURVEY.md section 8 e): what the collectives carry is
// assembled by these few kernels instead of dozens of small tensor operations per block, so the host's share of a block is
// a handful of library calls.
//
//   rb_shard_pack_stats    [frames built, points, x/y bounds, capacity] of this rank           -> all-gather 1
//   rb_shard_pack_layout   [points owned, start of the last-h-frames zone, ids and per-frame
//                           point counts of the first h and the last h frames]                  -> all-gather 2
//   rb_shard_local_index   float32 time and int64 GLOBAL index of every point of the local problem
//                          [left halo | owned | right halo]
//   rb_shard_pack_keys     the component keys the stitch needs: the four boundary zones run-length encoded (a dense zone of
//                          one component is a single entry) and the rank's distinct component keys      -> all-gather 3
#include "common.cuh"

namespace {

constexpr int SH_THREADS = 256;

__global__ void __launch_bounds__(SH_THREADS) shard_stats_kernel(const int64_t* __restrict__ off, int64_t n_frames, const float* __restrict__ b4,
                                                                double cap, double* __restrict__ out) {
    __shared__ int s_built;
    if (threadIdx.x == 0) s_built = 0;
    __syncthreads();
    int built = 0;
    for (int64_t f = threadIdx.x; f < n_frames; f += blockDim.x) built += off[f + 1] > off[f];
    if (built) atomicAdd(&s_built, built);
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = (double)s_built;
        out[1] = (double)off[n_frames];
        out[2] = (double)b4[0]; out[3] = (double)b4[1]; out[4] = (double)b4[2]; out[5] = (double)b4[3];
        out[6] = cap;
    }
}

// out = [off[F], off[F - hh], ids[0..hh), per[0..hh), ids[F-hh..F), per[F-hh..F)], per[f] = off[f+1] - off[f];
// edge_ids = the 2 hh ids (first hh, last hh) staged by the host
__global__ void shard_layout_kernel(const int64_t* __restrict__ off, int64_t n_frames, const int64_t* __restrict__ edge_ids, int hh,
                                    int64_t* __restrict__ out) {
    const int t = threadIdx.x;
    if (t == 0) { out[0] = off[n_frames]; out[1] = off[n_frames - hh]; }
    if (t < hh) {
        out[2 + t] = edge_ids[t];
        out[2 + hh + t] = off[t + 1] - off[t];
        out[2 + 2 * hh + t] = edge_ids[hh + t];
        const int64_t f = n_frames - hh + t;
        out[2 + 3 * hh + t] = off[f + 1] - off[f];
    }
}

__global__ void __launch_bounds__(SH_THREADS) shard_gidx_kernel(int64_t n_loc, int64_t nl, int64_t n_own, int64_t lbase, int64_t gbase,
                                                               int64_t rbase, int64_t* __restrict__ gidx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_loc) return;
    gidx[i] = i < nl ? lbase + i : (i < nl + n_own ? gbase + (i - nl) : rbase + (i - nl - n_own));
}

struct KeySegs {
    int64_t src[5];        // first source index of the segment in key[]
    int64_t end[5];        // cumulative length up to and including the segment
};

// virtual candidate vector = the four boundary zones, then all points. An ENTRY starts
//   in a zone: where the key differs from the key of the point before it (and at the zone's first point) - the zones are
//              run-length encoded over ALL their points, -1 = not a core point: a dense zone of millions of core points of
//              one component is ONE entry, and the entries of two ranks for the same zone are merged by position on the host;
//   in the last segment: at every point that is its component's smallest core (key == own global index).
__device__ __forceinline__ int64_t key_source(const KeySegs& sg, int64_t v, int* seg, int64_t* pos_in_seg) {
    int j = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) j += v >= sg.end[k];
    *seg = j;
    *pos_in_seg = v - (j ? sg.end[j - 1] : 0);
    return sg.src[j] + *pos_in_seg;
}

__global__ void __launch_bounds__(SH_THREADS) shard_key_flag_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ gidx, KeySegs sg,
                                                                   int32_t* __restrict__ flag) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= sg.end[4]) return;
    int seg;
    int64_t pos;
    const int64_t i = key_source(sg, v, &seg, &pos);
    const int64_t k = key[i];
    flag[v] = seg < 4 ? (pos == 0 || key[i - 1] != k) : (k == gidx[i]);
}

// vec = [5 running entry counts | 4 zone lengths | keys[cap] | starts[cap]]
__global__ void __launch_bounds__(SH_THREADS) shard_key_scatter_kernel(const int64_t* __restrict__ key, KeySegs sg, const int32_t* __restrict__ flag,
                                                                      const int32_t* __restrict__ pos, int64_t cap, int64_t* __restrict__ vec) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= sg.end[4]) return;
    int seg;
    int64_t in_seg;
    const int64_t i = key_source(sg, v, &seg, &in_seg);
    const int take = flag[v];
    const int64_t p = pos[v];
    if (take && p < cap) {
        vec[9 + p] = key[i];
        vec[9 + cap + p] = seg < 4 ? in_seg : 0;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j)
        if (sg.end[j] - 1 == v) vec[j] = p + take;                   // entries up to the end of segment j
    if (v == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) vec[5 + j] = sg.end[j] - (j ? sg.end[j - 1] : 0);
    }
}

// pinned staging: [0, 1 KiB) belongs to the read-backs of other entry points; the regions below are rewritten only by
// the NEXT block on this context, i.e. after that block's first stream sync
constexpr size_t STAGE_LAYOUT = 1024;       // 2 x 64 int64
constexpr size_t STAGE_LOCAL = 4096;        // head + ids of the local frame list

int ensure_pinned(rb_ctx* ctx, size_t need, cudaStream_t stream) {
    if (ctx->pinned_cap >= need) return RB_OK;
    RB_CUDA(cudaStreamSynchronize(stream));
    if (ctx->pinned) RB_CUDA(cudaFreeHost(ctx->pinned));
    ctx->pinned = nullptr; ctx->pinned_cap = 0;
    RB_CUDA(cudaMallocHost(&ctx->pinned, need + 4096));
    ctx->pinned_cap = need + 4096;
    return RB_OK;
}

}  // namespace

extern "C" int rb_shard_pack_stats(rb_ctx* ctx, const int64_t* frame_off, int64_t n_frames, const float* bounds4, int64_t cap, double* out7,
                                   void* stream_) {
    RB_REQUIRE(ctx && frame_off && bounds4 && out7 && n_frames >= 0, "bad arguments");
    RB_CUDA(rb_launch(ctx, shard_stats_kernel, dim3(1), dim3(SH_THREADS), 0, (cudaStream_t)stream_, frame_off, n_frames, bounds4, (double)cap, out7));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

extern "C" int rb_shard_pack_layout(rb_ctx* ctx, const int64_t* frame_off, int64_t n_frames, const int64_t* frame_ids_host, int hh, int64_t* out,
                                    void* stream_) {
    RB_REQUIRE(ctx && frame_off && out && n_frames >= 0 && hh >= 0 && hh <= 64 && hh <= n_frames && (hh == 0 || frame_ids_host), "bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    void* d_ids;
    RB_TRY(rb_scratch_get(ctx, RB_S_SHARD_IDS, sizeof(int64_t) * 128, &d_ids));
    if (hh > 0) {
        int64_t* h = (int64_t*)((unsigned char*)ctx->pinned + STAGE_LAYOUT);
        for (int t = 0; t < hh; ++t) { h[t] = frame_ids_host[t]; h[hh + t] = frame_ids_host[n_frames - hh + t]; }
        RB_CUDA(cudaMemcpyAsync(d_ids, h, sizeof(int64_t) * 2 * (size_t)hh, cudaMemcpyHostToDevice, stream));
    }
    RB_CUDA(rb_launch(ctx, shard_layout_kernel, dim3(1), dim3(64), 0, stream, frame_off, n_frames, (const int64_t*)d_ids, hh, out));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

extern "C" int rb_shard_local_index(rb_ctx* ctx, const int64_t* head_host, const float* ids_host, int64_t n_local_frames, int64_t nl, int64_t n_own,
                                    int64_t nr, int64_t lbase, int64_t gbase, int64_t rbase, float* times, int64_t* gidx, void* stream_) {
    RB_REQUIRE(ctx && n_local_frames >= 0 && nl >= 0 && n_own >= 0 && nr >= 0, "bad arguments");
    const int64_t n_loc = nl + n_own + nr;
    if (n_loc == 0) return RB_OK;
    RB_REQUIRE(head_host && ids_host && times && gidx && head_host[n_local_frames] == n_loc, "bad local frame list");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t head_bytes = sizeof(int64_t) * (size_t)(n_local_frames + 1), ids_bytes = sizeof(float) * (size_t)n_local_frames;
    RB_TRY(ensure_pinned(ctx, STAGE_LOCAL + head_bytes + ids_bytes + 64, stream));
    void* d_raw;
    RB_TRY(rb_scratch_get(ctx, RB_S_SHARD_LOCAL, head_bytes + ids_bytes + 64, &d_raw));
    unsigned char* h = (unsigned char*)ctx->pinned + STAGE_LOCAL;
    memcpy(h, head_host, head_bytes);
    memcpy(h + head_bytes, ids_host, ids_bytes);
    RB_CUDA(cudaMemcpyAsync(d_raw, h, head_bytes + ids_bytes, cudaMemcpyHostToDevice, stream));
    const int64_t* d_head = (const int64_t*)d_raw;
    const float* d_ids = (const float*)((unsigned char*)d_raw + head_bytes);
    RB_TRY(rb_expand_frame_times(ctx, d_head, d_ids, n_local_frames, n_loc, times, stream_));
    RB_CUDA(rb_launch(ctx, shard_gidx_kernel, dim3((unsigned)rb_div_up(n_loc, SH_THREADS)), dim3(SH_THREADS), 0, stream, n_loc, nl, n_own, lbase, gbase,
                      rbase, gidx));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

extern "C" int rb_shard_pack_keys(rb_ctx* ctx, const int64_t* key, const int64_t* gidx, int64_t n_loc, const int64_t* zones8_host, int64_t cap_keys,
                                  int64_t* vec, void* stream_) {
    RB_REQUIRE(ctx && vec && zones8_host && n_loc >= 0 && cap_keys >= 0, "bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    RB_CUDA(cudaMemsetAsync(vec, 0, sizeof(int64_t) * (size_t)(9 + 2 * cap_keys), stream));
    if (n_loc == 0) return RB_OK;
    RB_REQUIRE(key && gidx, "NULL keys");
    KeySegs sg;
    int64_t run = 0;
    for (int j = 0; j < 4; ++j) {
        const int64_t a = zones8_host[2 * j], b = zones8_host[2 * j + 1];
        RB_REQUIRE(a >= 0 && b >= a && b <= n_loc, "bad zone");
        sg.src[j] = a;
        run += b - a;
        sg.end[j] = run;
    }
    sg.src[4] = 0;
    run += n_loc;
    sg.end[4] = run;
    RB_REQUIRE(run < ((int64_t)1 << 31), "too many points");
    void* raw;
    RB_TRY(rb_scratch_get(ctx, RB_S_SHARD_KEYS, sizeof(int32_t) * 2 * (size_t)(run + 1), &raw));
    int32_t* flag = (int32_t*)raw;
    int32_t* pos = flag + run + 1;
    const unsigned blocks = (unsigned)rb_div_up(run, SH_THREADS);
    RB_CUDA(rb_launch(ctx, shard_key_flag_kernel, dim3(blocks), dim3(SH_THREADS), 0, stream, key, gidx, sg, flag));
    RB_LAUNCH_CHECK(ctx);
    RB_TRY(rb_exclusive_scan_i32(ctx, flag, pos, run, nullptr, stream));
    RB_CUDA(rb_launch(ctx, shard_key_scatter_kernel, dim3(blocks), dim3(SH_THREADS), 0, stream, key, sg, (const int32_t*)flag, (const int32_t*)pos,
                      cap_keys, vec));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}
