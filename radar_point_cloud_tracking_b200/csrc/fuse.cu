// Max fusion of gains on a grid (reference 5_gain_fusion_ply_builder.py:222-273, fuse_gains_max).
// cell = trunc((v - v_min) / res) in float32 (T5:258-259), max intensity per cell (T5:263), occupied
// cells emitted in y-major order (T5:267). Cell centres are float64 host arithmetic (T5:269-270).
#include "common.cuh"

namespace {

__global__ void fuse_pool_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ inten,
                                 int64_t n, float x_min, float y_min, float res, int nx, int ny, int* __restrict__ grid) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int ix = (int)__fdiv_rn(__fsub_rn(x[i], x_min), res);
        int iy = (int)__fdiv_rn(__fsub_rn(y[i], y_min), res);
        float v = inten[i];
        if (ix < 0 || ix >= nx || iy < 0 || iy >= ny) continue;     // cannot happen for in-bounds input
        if (v > 0.f) atomicMax(grid + (int64_t)ix * ny + iy, __float_as_int(v));   // max_grid starts at 0 (T5:262)
    }
}

// k = iy*nx + ix (y-major emission order)
__global__ void fuse_flag_kernel(const int* __restrict__ grid, int nx, int ny, int* __restrict__ flags) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)nx * ny) return;
    int iy = (int)(k / nx), ix = (int)(k - (int64_t)iy * nx);
    flags[k] = __int_as_float(grid[(int64_t)ix * ny + iy]) > 0.f;
}

__global__ void fuse_emit_kernel(const int* __restrict__ grid, const int* __restrict__ flags, const int* __restrict__ pos,
                                 int nx, int ny, int64_t cap, int32_t* __restrict__ cix, int32_t* __restrict__ ciy,
                                 float* __restrict__ cmax) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)nx * ny || !flags[k]) return;
    int iy = (int)(k / nx), ix = (int)(k - (int64_t)iy * nx);
    int64_t o = pos[k];
    if (o >= cap) return;
    cix[o] = ix; ciy[o] = iy;
    cmax[o] = __int_as_float(grid[(int64_t)ix * ny + iy]);
}

}  // namespace

extern "C" int rb_fuse_max(rb_ctx* ctx, const float* x, const float* y, const float* inten, int64_t n, float x_min,
                           float y_min, float resolution, int nx, int ny, int32_t* cell_ix, int32_t* cell_iy,
                           float* cell_max, int64_t cap_cells, int64_t* n_cells_out, void* stream_) {
    RB_REQUIRE(ctx && n_cells_out, "NULL argument");
    RB_REQUIRE(nx > 0 && ny > 0 && (int64_t)nx * ny < ((int64_t)1 << 30), "bad grid size");
    RB_REQUIRE(resolution > 0, "resolution must be positive");
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_cells_out = 0;
    if (n <= 0) return RB_OK;
    RB_REQUIRE(x && y && inten && cell_ix && cell_iy && cell_max, "NULL argument");
    const int64_t cells = (int64_t)nx * ny;
    void* buf;
    RB_TRY(rb_scratch_get(ctx, RB_S_FUSE_GRID, sizeof(int) * (size_t)(cells * 3 + 1), &buf));
    int* grid = (int*)buf;
    int* flags = grid + cells;
    int* pos = flags + cells;
    RB_CUDA(cudaMemsetAsync(grid, 0, sizeof(int) * (size_t)cells, stream));
    int blocks = (int)(rb_div_up(n, 256) < (int64_t)ctx->sm_count * 8 ? rb_div_up(n, 256) : (int64_t)ctx->sm_count * 8);
    RB_CUDA(rb_launch(ctx, fuse_pool_kernel, dim3(blocks), dim3(256), 0, stream, x, y, inten, n, x_min, y_min, resolution, nx, ny, grid));
    RB_LAUNCH_CHECK(ctx);
    unsigned cb = (unsigned)rb_div_up(cells, 256);
    RB_CUDA(rb_launch(ctx, fuse_flag_kernel, dim3(cb), dim3(256), 0, stream, grid, nx, ny, flags));
    RB_LAUNCH_CHECK(ctx);
    RB_TRY(rb_exclusive_scan_i32(ctx, flags, pos, cells, pos + cells, stream));
    RB_CUDA(rb_launch(ctx, fuse_emit_kernel, dim3(cb), dim3(256), 0, stream, grid, flags, pos, nx, ny, cap_cells, cell_ix, cell_iy, cell_max));
    RB_LAUNCH_CHECK(ctx);
    int* h = (int*)ctx->pinned;
    RB_CUDA(cudaMemcpyAsync(h, pos + cells, sizeof(int), cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    *n_cells_out = h[0];
    if (h[0] > cap_cells) { rb_set_error("rb_fuse_max: %d cells exceed capacity %lld", h[0], (long long)cap_cells); return RB_ERR_CAPACITY; }
    return RB_OK;
}
