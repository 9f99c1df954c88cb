// Spoke-to-point: threshold + ordered stream compaction + stride + polar->Cartesian + gain concat,
// one pass over the echo tensor (HBM-bound: every echo byte is read exactly once).
//
// Replaces load_radar_csv's numeric part (reference 4_temporal_object_tracker.py:200-232) and the
// per-gain concatenation of build_frame (:322-344) for a batch of W = frames x gains sweeps.
//
// Layout: echo[W][S][E] float32. A tile is SP_TILE consecutive cells of ONE sweep; a warp owns
// SP_VEC consecutive 128-cell chunks, a lane 4 consecutive cells of each chunk (one 128-bit
// load), so the row-major order of survivors falls out of warp ballots without any shuffle scan.
// Tile prefixes come from a decoupled look-back over per-tile descriptors (tiles are handed out
// by an atomic ticket, so a tile only ever waits on tiles that are already running); the output
// base of a sweep (sum of ceil(M_w/stride) of all earlier sweeps) is chained through a second,
// per-sweep descriptor published by each sweep's last tile.
#include "common.cuh"

namespace {

constexpr int SP_THREADS = 256;
constexpr int SP_WARPS = SP_THREADS / 32;
constexpr int SP_VEC = 4;                          // 128-bit loads per lane
constexpr int SP_CHUNK = 128;                      // cells per warp load
constexpr int SP_WARP_CELLS = SP_CHUNK * SP_VEC;   // 512
constexpr int SP_TILE = SP_WARP_CELLS * SP_WARPS;  // 4096 cells = 16 KiB

constexpr unsigned long long ST_INVALID = 0ull;
constexpr unsigned long long ST_AGGREGATE = 1ull << 32;
constexpr unsigned long long ST_INCLUSIVE = 2ull << 32;

struct SpokeArgs {
    const float* echo;
    const float* cos_tab;
    const float* sin_tab;
    const float* range_res;
    const float* ranges;                 // optional [W][S][E] explicit range per cell (NULL: range_res*j)
    const int32_t* sweep_gain;
    float* x;
    float* y;
    float* inten;
    int32_t* gain;
    int64_t* sweep_base;                 // [W+1]
    unsigned long long* tile_status;     // [W * tiles_per_sweep]
    int* sweep_ready;                    // [W+1]
    int* ticket;
    int64_t cap;
    int64_t total_tiles;
    int sweep_cells;                     // S*E
    int tiles_per_sweep;
    int n_bins;
    int stride;
    float threshold;
    int vec_ok;                          // sweep_cells % 4 == 0, bins % 4 == 0 and base 16B aligned
};

// Exclusive prefix (survivors of this sweep before tile t) by warp 0. Descriptors of tiles of the
// same sweep only: idx < 0 means "before the sweep" and counts as an inclusive 0.
__device__ __forceinline__ unsigned lookback(const unsigned long long* status, int t) {
    unsigned excl = 0;
    int look = t - 1;
    const unsigned lane = rb_lane();
    while (true) {
        int idx = look - (int)lane;
        unsigned long long st = ST_INCLUSIVE;
        if (idx >= 0) {
            st = rb_ld_acquire_u64(status + idx);
            while ((st >> 32) == 0) st = rb_ld_acquire_u64(status + idx);
        }
        unsigned incl_mask = __ballot_sync(0xffffffffu, (st >> 32) == 2ull);
        unsigned val = (unsigned)(st & 0xffffffffull);
        if (incl_mask) {
            int first = __ffs(incl_mask) - 1;          // nearest predecessor with a full prefix
            if ((int)lane > first) val = 0;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
        excl += val;
        if (incl_mask) break;
        look -= 32;
    }
    return excl;
}

__global__ void __launch_bounds__(SP_THREADS) spoke_to_points_kernel(const SpokeArgs a) {
    __shared__ int s_tile;
    __shared__ unsigned s_warp_total[SP_WARPS];
    __shared__ unsigned s_excl;
    __shared__ long long s_sweep_base;

    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= a.total_tiles) return;
    const int w = tile / a.tiles_per_sweep;
    const int t = tile - w * a.tiles_per_sweep;
    const unsigned lane = rb_lane();
    const int warp = threadIdx.x >> 5;
    const float* __restrict__ src = a.echo + (int64_t)w * a.sweep_cells;
    const int cell0 = t * SP_TILE + warp * SP_WARP_CELLS + (int)lane * 4;

    // ---- load: SP_VEC independent 128-bit streaming loads per lane -------------------------
    float4 v[SP_VEC];
    const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int k = 0; k < SP_VEC; ++k) {
        int c = cell0 + k * SP_CHUNK;
        if (a.vec_ok && c + 3 < a.sweep_cells) {
            v[k] = rb_ld_stream4(src + c);
        } else {
            v[k].x = c + 0 < a.sweep_cells ? __ldg(src + c + 0) : qnan;
            v[k].y = c + 1 < a.sweep_cells ? __ldg(src + c + 1) : qnan;
            v[k].z = c + 2 < a.sweep_cells ? __ldg(src + c + 2) : qnan;
            v[k].w = c + 3 < a.sweep_cells ? __ldg(src + c + 3) : qnan;
        }
    }

    // ---- threshold + in-warp ranks from ballots ---------------------------------------------
    const unsigned lt = rb_lanemask_lt();
    unsigned pass_bits = 0;            // bit (4k + c): cell c of chunk k survives
    unsigned rank_base[SP_VEC];        // survivors of this warp before my first cell of chunk k
    unsigned running = 0;
#pragma unroll
    for (int k = 0; k < SP_VEC; ++k) {
        bool p0 = v[k].x > a.threshold, p1 = v[k].y > a.threshold;
        bool p2 = v[k].z > a.threshold, p3 = v[k].w > a.threshold;
        unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
        unsigned b2 = __ballot_sync(0xffffffffu, p2), b3 = __ballot_sync(0xffffffffu, p3);
        rank_base[k] = running + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
        running += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
        pass_bits |= ((unsigned)p0 | ((unsigned)p1 << 1) | ((unsigned)p2 << 2) | ((unsigned)p3 << 3)) << (4 * k);
    }
    if (lane == 0) s_warp_total[warp] = running;
    __syncthreads();
    unsigned warp_prefix = 0, tile_total = 0;
#pragma unroll
    for (int i = 0; i < SP_WARPS; ++i) {
        unsigned c = s_warp_total[i];
        if (i < warp) warp_prefix += c;
        tile_total += c;
    }

    // ---- tile prefix (decoupled look-back) and sweep output base -----------------------------
    if (warp == 0) {
        unsigned long long* status = a.tile_status + (int64_t)w * a.tiles_per_sweep;
        unsigned excl = 0;
        if (t == 0) {
            if (lane == 0) rb_st_release_u64(status, ST_INCLUSIVE | tile_total);
        } else {
            if (lane == 0) rb_st_release_u64(status + t, ST_AGGREGATE | tile_total);
            excl = lookback(status, t);
            if (lane == 0) rb_st_release_u64(status + t, ST_INCLUSIVE | (unsigned long long)(excl + tile_total));
        }
        if (lane == 0) {
            long long base = 0;
            if (w > 0) {
                while (rb_ld_acquire_s32(a.sweep_ready + w) == 0) { }
                base = a.sweep_base[w];
            } else if (t == 0) {
                a.sweep_base[0] = 0;
            }
            if (t == a.tiles_per_sweep - 1) {          // last tile of the sweep: publish the next base
                long long m = (long long)excl + tile_total;
                a.sweep_base[w + 1] = base + (m + a.stride - 1) / a.stride;
                __threadfence();
                rb_st_release_s32(a.sweep_ready + w + 1, 1);
            }
            s_excl = excl;
            s_sweep_base = base;
        }
    }
    __syncthreads();
    if (pass_bits == 0) return;

    // ---- write survivors whose rank in the sweep is a multiple of the stride ------------------
    const unsigned rank0 = s_excl + warp_prefix;
    const long long out_base = s_sweep_base;
    const int gain_label = a.sweep_gain[w];
    const float* __restrict__ cos_w = a.cos_tab + (int64_t)w * (a.sweep_cells / a.n_bins);
    const float* __restrict__ sin_w = a.sin_tab + (int64_t)w * (a.sweep_cells / a.n_bins);
    const float* __restrict__ res_w = a.range_res ? a.range_res + (int64_t)w * (a.sweep_cells / a.n_bins) : nullptr;
#pragma unroll
    for (int k = 0; k < SP_VEC; ++k) {
        unsigned bits = (pass_bits >> (4 * k)) & 0xfu;
        if (!bits) continue;
        const int c = cell0 + k * SP_CHUNK;
        unsigned r = rank0 + rank_base[k];
        const float vals[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (!(bits & (1u << i))) continue;
            unsigned q = r / (unsigned)a.stride;
            if (q * (unsigned)a.stride == r) {
                long long pos = out_base + q;
                if (pos < a.cap) {
                    int cell = c + i;
                    int s = cell / a.n_bins;
                    int j = cell - s * a.n_bins;
                    float rng = a.ranges ? a.ranges[(int64_t)w * a.sweep_cells + cell]
                                         : __fmul_rn(res_w[s], (float)j);   // T4:214
                    a.x[pos] = __fmul_rn(rng, cos_w[s]);                // T4:217
                    a.y[pos] = __fmul_rn(rng, sin_w[s]);                // T4:218
                    a.inten[pos] = vals[i];
                    a.gain[pos] = gain_label;
                }
            }
            ++r;
        }
    }
}

__global__ void frame_offsets_kernel(const int64_t* __restrict__ sweep_base, int64_t n_frames, int gpf,
                                     int64_t* __restrict__ frame_off) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f <= n_frames) frame_off[f] = sweep_base[f * gpf];
}

__global__ void expand_times_kernel(const int64_t* __restrict__ frame_off, const float* __restrict__ frame_ids,
                                    int64_t n_frames, float* __restrict__ times) {
    // one block per frame, grid-stride over frames
    for (int64_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
        int64_t b = frame_off[f], e = frame_off[f + 1];
        float id = frame_ids[f];
        for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) times[i] = id;
    }
}

// x = ranges * cos[:, None], y = ranges * sin[:, None] on a full [N][M] grid (PKG transforms.py:13-34)
__global__ void polar_grid_kernel(const float* __restrict__ ranges, const float* __restrict__ cs,
                                  const float* __restrict__ sn, int64_t total, int m, float* __restrict__ x,
                                  float* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / m;
        float rg = ranges[i];
        x[i] = __fmul_rn(rg, cs[r]);
        y[i] = __fmul_rn(rg, sn[r]);
    }
}

}  // namespace

extern "C" int rb_polar_to_cartesian(rb_ctx* ctx, const float* ranges, const float* cos_tab, const float* sin_tab,
                                     int64_t n_rows, int n_cols, float* x, float* y, void* stream_) {
    RB_REQUIRE(ctx, "ctx is NULL");
    RB_REQUIRE(n_rows >= 0 && n_cols >= 0, "negative size");
    int64_t total = n_rows * n_cols;
    if (total == 0) return RB_OK;
    RB_REQUIRE(ranges && cos_tab && sin_tab && x && y, "NULL argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    int blocks = (int)(rb_div_up(total, 256 * 4) < (int64_t)ctx->sm_count * 16 ? rb_div_up(total, 256 * 4)
                                                                               : (int64_t)ctx->sm_count * 16);
    polar_grid_kernel<<<blocks, 256, 0, stream>>>(ranges, cos_tab, sin_tab, total, n_cols, x, y);
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

extern "C" int rb_spoke_to_points(rb_ctx* ctx, const float* echo, const float* cos_tab, const float* sin_tab,
                                  const float* range_res, const float* ranges, const int32_t* sweep_gain,
                                  int64_t n_sweeps,
                                  int n_spokes, int n_bins, float threshold, int stride, float* x, float* y,
                                  float* inten, int32_t* gain, int64_t cap, int64_t* sweep_base, void* stream_) {
    RB_REQUIRE(ctx, "ctx is NULL");
    RB_REQUIRE(n_sweeps >= 0 && n_spokes >= 0 && n_bins >= 0, "negative size");
    RB_REQUIRE(sweep_base, "sweep_base is NULL");
    RB_REQUIRE(cap >= 0, "negative capacity");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (stride < 1) stride = 1;                                   // T4:227 applies the stride only when > 1
    int64_t cells = (int64_t)n_spokes * n_bins;
    RB_REQUIRE(cells < (int64_t)1 << 31, "a sweep must hold fewer than 2^31 cells");
    if (n_sweeps == 0 || cells == 0) {
        RB_CUDA(cudaMemsetAsync(sweep_base, 0, sizeof(int64_t) * (size_t)(n_sweeps + 1), stream));
        return RB_OK;
    }
    RB_REQUIRE(echo && cos_tab && sin_tab && (range_res || ranges) && sweep_gain, "NULL input");
    RB_REQUIRE(cap == 0 || (x && y && inten && gain), "NULL output");
    int64_t tiles_per_sweep = rb_div_up(cells, SP_TILE);
    int64_t total_tiles = tiles_per_sweep * n_sweeps;
    RB_REQUIRE(total_tiles < (int64_t)1 << 31, "too many tiles in one batch; split the batch");

    void* status;
    void* flags;
    RB_TRY(rb_scratch_get(ctx, RB_S_TILE_STATUS, sizeof(unsigned long long) * (size_t)total_tiles, &status));
    size_t flag_bytes = sizeof(int) * (size_t)(n_sweeps + 2);
    RB_TRY(rb_scratch_get(ctx, RB_S_SWEEP_FLAGS, flag_bytes, &flags));
    RB_CUDA(cudaMemsetAsync(status, 0, sizeof(unsigned long long) * (size_t)total_tiles, stream));
    RB_CUDA(cudaMemsetAsync(flags, 0, flag_bytes, stream));

    SpokeArgs a;
    a.echo = echo; a.cos_tab = cos_tab; a.sin_tab = sin_tab; a.range_res = range_res;
    a.ranges = ranges;
    a.sweep_gain = sweep_gain;
    a.x = x; a.y = y; a.inten = inten; a.gain = gain;
    a.sweep_base = sweep_base;
    a.tile_status = (unsigned long long*)status;
    a.sweep_ready = (int*)flags;
    a.ticket = (int*)flags + (n_sweeps + 1);
    a.cap = cap;
    a.total_tiles = total_tiles;
    a.sweep_cells = (int)cells;
    a.tiles_per_sweep = (int)tiles_per_sweep;
    a.n_bins = n_bins;
    a.stride = stride;
    a.threshold = threshold;
    a.vec_ok = (cells % 4 == 0) && (n_bins % 4 == 0) && (((uintptr_t)echo & 15u) == 0);
    spoke_to_points_kernel<<<(unsigned)total_tiles, SP_THREADS, 0, stream>>>(a);
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

extern "C" int rb_frame_offsets(rb_ctx* ctx, const int64_t* sweep_base, int64_t n_frames, int gains_per_frame,
                                int64_t* frame_off, void* stream_) {
    RB_REQUIRE(ctx && sweep_base && frame_off, "NULL argument");
    RB_REQUIRE(n_frames >= 0 && gains_per_frame >= 1, "bad sizes");
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned blocks = (unsigned)rb_div_up(n_frames + 1, 256);
    frame_offsets_kernel<<<blocks, 256, 0, stream>>>(sweep_base, n_frames, gains_per_frame, frame_off);
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

extern "C" int rb_expand_frame_times(rb_ctx* ctx, const int64_t* frame_off, const float* frame_ids,
                                     int64_t n_frames, int64_t n_points, float* times, void* stream_) {
    RB_REQUIRE(ctx && frame_off && frame_ids, "NULL argument");
    if (n_frames <= 0 || n_points <= 0) return RB_OK;
    RB_REQUIRE(times, "times is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned blocks = (unsigned)(n_frames < 4096 ? n_frames : 4096);
    expand_times_kernel<<<blocks, 256, 0, stream>>>(frame_off, frame_ids, n_frames, times);
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}
