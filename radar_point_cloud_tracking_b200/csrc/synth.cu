// Device twin of radar_point_cloud_tracking_b200/synthetic.py::synth_echo — TEST/BENCH INPUT ONLY,
// not part of the detection path. Counter-based hashing, so host and device generate identical data.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

constexpr int RECT_COLS = 10;   // frame, s0, s1, j0, j1, span, base@gain0..3

__global__ void synth_kernel(float* __restrict__ echo, int64_t total, int cells, int n_bins, int gpf, int64_t first_frame,
                             const uint32_t* __restrict__ keys, const uint32_t* __restrict__ thr,
                             const int32_t* __restrict__ rects, const int32_t* __restrict__ rect_off) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t w = i / cells;
        uint32_t cell = (uint32_t)(i - w * cells);
        int s = (int)(cell / (uint32_t)n_bins), j = (int)(cell - (uint32_t)s * n_bins);
        uint32_t h1 = mix32(keys[w] ^ (cell * 0x9E3779B9u));
        uint32_t h2 = mix32(h1 + 0x6A09E667u);
        int val = (int)(h1 % 10u);
        if (h2 < thr[w]) val = 11 + (int)(mix32(h2 ^ 0xBB67AE85u) % 245u);
        int64_t f = first_frame + w / gpf;
        int gi = (int)(w % gpf);
        for (int r = rect_off[f]; r < rect_off[f + 1]; ++r) {
            const int32_t* q = rects + (int64_t)r * RECT_COLS;
            if (s >= q[1] && s < q[2] && j >= q[3] && j < q[4]) {
                int v = q[6 + gi] + (int)(mix32(h1 ^ 0x3C6EF372u) % (uint32_t)q[5]);
                val = v < 255 ? v : 255;                       // later rectangles overwrite earlier ones
            }
        }
        echo[i] = (float)val;
    }
}

}  // namespace

extern "C" int rb_synth_echo(rb_ctx* ctx, float* echo, int64_t n_sweeps, int n_spokes, int n_bins, int gains_per_frame,
                             int64_t first_frame, const uint32_t* sweep_keys, const uint32_t* clutter_thr,
                             const int32_t* rects, const int32_t* rect_off, void* stream_) {
    RB_REQUIRE(ctx && echo && sweep_keys && clutter_thr && rect_off, "NULL argument");
    RB_REQUIRE(gains_per_frame >= 1 && gains_per_frame <= 4, "1..4 gains per frame");
    int64_t cells = (int64_t)n_spokes * n_bins;
    RB_REQUIRE(cells > 0 && cells < ((int64_t)1 << 31), "bad sweep size");
    if (n_sweeps <= 0) return RB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t total = cells * n_sweeps;
    int blocks = (int)(rb_div_up(total, 256 * 8) < (int64_t)ctx->sm_count * 16 ? rb_div_up(total, 256 * 8)
                                                                               : (int64_t)ctx->sm_count * 16);
    RB_CUDA(rb_launch(ctx, synth_kernel, dim3(blocks), dim3(256), 0, stream, echo, total, (int)cells, n_bins, gains_per_frame, first_frame, sweep_keys,
                                             clutter_thr, rects, rect_off));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}
