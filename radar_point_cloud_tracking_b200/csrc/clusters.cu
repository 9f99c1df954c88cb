// Per-frame cluster records on the device (SURVEY.md section 8 f, rank 1; reference 4_temporal_object_tracker.py:511-534,
// consumed by ObjectTracker.update at :553 and save_tracking_results at :873-886).
//
// For every frame the reference walks `set(frame_labels)` and builds, per cluster id present in the frame,
//     points = frame_coords[mask] ; intensities = frame_intensities[mask] ; centroid = np.mean(points, axis=0)
// and the tracker / the CSV writer read num_points, centroid and mean_intensity = float(np.mean(intensities)).
// Here a SEGMENT is one (frame, label) pair that occurs; the call produces
//   * the segment table: frame, label, index of the segment's first point inside its frame (the order of first
//     occurrences is what decides the iteration order of Python's set - the host replays it), number of points,
//     offset of the segment's points in the grouped arrays, centroid x / y and mean intensity;
//   * the points grouped by segment, each segment in the original point order (a STABLE partition), so that the host's
//     `points` / `intensities` of a cluster are slices (views) of three arrays instead of boolean-mask copies.
// numpy's float32 arithmetic is reproduced bit for bit (probed against numpy 2.3, see tests/test_oracle_golden.py):
//   np.mean(points, axis=0)  = add.reduce over the rows in order, one float32 accumulator per column, then / n
//   np.mean(intensities)     = numpy's PAIRWISE float32 sum of the contiguous 1-D array (blocks of <= 128 with eight
//                              partial sums, halves aligned to 8 above that), then / n
// The noise "label" -1 of a frame is a segment too (count and first occurrence only): it takes a slot in Python's set
// before it is discarded.
//
// How: (1) count + first occurrence per (frame, label) in a dense table (integer atomics: deterministic);
// (2) compaction of the occupied slots (scan) and offsets of the segments (scan); (3) the stable partition: frames are
// cut into tiles of 1024 points, one warp per tile; a tile counts its labels in a small shared-memory hash, takes its
// turn in the frame's tile order (tiles of one frame form a chain; most frames are a single tile or two) to fetch the
// running fill of each label, then ranks its points batch by batch (__match_any_sync) and scatters them;
// (4) one lane per segment sums its contiguous slice in numpy's order.
#include "common.cuh"

namespace {

constexpr int CL_THREADS = 256;
constexpr int CL_TILE = 1024;              // points per tile (one warp)
constexpr int CL_HASH = 2048;              // hash entries per warp (a tile holds at most CL_TILE distinct labels)
constexpr int CL_WARPS = 2;                // warps per block of the partition kernel (40 KB of static shared memory)

__device__ __forceinline__ int frame_of(const int64_t* __restrict__ off, int n_frames, int64_t i) {
    int lo = 0, hi = n_frames;             // last f with off[f] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// (1) per (frame, label) slot: number of points and the smallest in-frame index
__global__ void __launch_bounds__(CL_THREADS) cl_count_kernel(const int32_t* __restrict__ labels, int64_t n, const int64_t* __restrict__ off,
                                                             int n_frames, int width, int32_t* __restrict__ cnt, int32_t* __restrict__ first,
                                                             int32_t* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int f = frame_of(off, n_frames, i);
    const int l = labels[i];
    if (l < -1 || l + 1 >= width) { atomicOr(bad, 1); return; }             // a label outside [-1, n_clusters)
    const int64_t slot = (int64_t)f * width + (l + 1);
    atomicAdd(cnt + slot, 1);
    atomicMin(first + slot, (int32_t)(i - off[f]));
}

__global__ void __launch_bounds__(CL_THREADS) cl_flag_kernel(const int32_t* __restrict__ cnt, int64_t m, int32_t* __restrict__ flag) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < m) flag[s] = cnt[s] > 0;
    if (s == m) flag[s] = 0;               // scan sentinel: rank[m] = number of segments
}

// (2) occupied slots -> segment table (frame-major, label ascending); grouped sizes for the second scan
__global__ void __launch_bounds__(CL_THREADS) cl_compact_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ first,
                                                               const int32_t* __restrict__ rank, int64_t m, int width, int64_t cap,
                                                               int32_t* __restrict__ seg_frame, int32_t* __restrict__ seg_label,
                                                               int32_t* __restrict__ seg_first, int32_t* __restrict__ seg_count,
                                                               int32_t* __restrict__ seg_grouped) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m || cnt[s] <= 0) return;
    const int64_t k = rank[s];
    if (k >= cap) return;
    const int label = (int)(s % width) - 1;
    if (seg_frame) seg_frame[k] = (int)(s / width);
    seg_label[k] = label;
    if (seg_first) seg_first[k] = first[s];
    seg_count[k] = cnt[s];
    seg_grouped[k] = label >= 0 ? cnt[s] : 0;
}

// slot -> offset of its points in the grouped arrays (the noise slots: -1)
__global__ void __launch_bounds__(CL_THREADS) cl_slot_start_kernel(const int32_t* __restrict__ cnt, const int32_t* __restrict__ rank,
                                                                  const int32_t* __restrict__ seg_start32, int64_t m, int width, int64_t cap,
                                                                  int32_t* __restrict__ slot_start, int64_t* __restrict__ seg_start) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    int32_t v = -1;
    if (cnt[s] > 0) {
        const int64_t k = rank[s];
        if (k < cap) {
            const bool noise = (s % width) == 0;
            v = noise ? -1 : seg_start32[k];
            seg_start[k] = v;
        }
    }
    slot_start[s] = v;
}

// tiles per frame (for the tile list): tiles[f] = ceil(len / CL_TILE)
__global__ void __launch_bounds__(CL_THREADS) cl_tiles_kernel(const int64_t* __restrict__ off, int n_frames, int32_t* __restrict__ tiles) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_frames) tiles[f] = (int32_t)((off[f + 1] - off[f] + CL_TILE - 1) / CL_TILE);
    if (f == n_frames) tiles[f] = 0;
}

// (3) stable partition, one warp per tile
struct HashEntry { int key; int val; };    // key = label + 1 (0 = empty), val = count, later the running fill

__device__ __forceinline__ int hash_slot(int key) { return (int)(((unsigned)key * 2654435761u) >> 21) & (CL_HASH - 1); }

__device__ __forceinline__ int hash_find_or_insert(HashEntry* __restrict__ h, int key, int* __restrict__ keys, int* __restrict__ n_keys) {
    int p = hash_slot(key);
    while (true) {
        const int old = atomicCAS(&h[p].key, 0, key);
        if (old == 0) { keys[atomicAdd(n_keys, 1)] = p; return p; }
        if (old == key) return p;
        p = (p + 1) & (CL_HASH - 1);
    }
}
__device__ __forceinline__ int hash_find(const HashEntry* __restrict__ h, int key) {
    int p = hash_slot(key);
    while (h[p].key != key) p = (p + 1) & (CL_HASH - 1);
    return p;
}

__global__ void __launch_bounds__(CL_WARPS * 32) cl_partition_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                    const float* __restrict__ inten, const int32_t* __restrict__ labels,
                                                                    const int64_t* __restrict__ off, int n_frames, int width,
                                                                    const int32_t* __restrict__ tile_base /* [F+1] */,
                                                                    const int32_t* __restrict__ slot_start, int32_t* __restrict__ fill,
                                                                    int32_t* __restrict__ turn, unsigned* __restrict__ ticket,
                                                                    float* __restrict__ gx, float* __restrict__ gy, float* __restrict__ gi) {
    __shared__ HashEntry s_hash[CL_WARPS][CL_HASH];
    __shared__ int s_keys[CL_WARPS][CL_TILE];
    __shared__ int s_nkeys[CL_WARPS];
    const unsigned lane = rb_lane();
    const int w = threadIdx.x >> 5;
    HashEntry* h = s_hash[w];
    int* keys = s_keys[w];
    const int total_tiles = tile_base[n_frames];
    while (true) {
        // tiles are handed out in execution order: the tile a warp waits for below has always started already
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if ((int)t >= total_tiles) return;
        int lo = 0, hi = n_frames;                                   // frame of the tile: last f with tile_base[f] <= t
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tile_base[mid] <= (int)t) lo = mid; else hi = mid; }
        const int f = lo, c = (int)t - tile_base[f];
        const int64_t p0 = off[f] + (int64_t)c * CL_TILE;
        const int len = (int)min((int64_t)CL_TILE, off[f + 1] - p0);
        for (int k = lane; k < CL_HASH; k += 32) { h[k].key = 0; h[k].val = 0; }
        if (lane == 0) s_nkeys[w] = 0;
        __syncwarp();
        // pass 1: how many points of each label in the tile
        for (int b = 0; b < len; b += 32) {
            const int j = b + (int)lane;
            const int key = j < len ? labels[p0 + j] + 1 : 0;      // 0 = noise (or past the end): not grouped
            const unsigned active = __ballot_sync(0xffffffffu, key > 0);
            if (key > 0) {
                const unsigned same = __match_any_sync(active, key);
                if ((int)lane == __ffs(same) - 1) {
                    const int p = hash_find_or_insert(h, key, keys, &s_nkeys[w]);
                    atomicAdd(&h[p].val, __popc(same));
                }
            }
            __syncwarp();
        }
        // the frame's tiles take their turn in order: running fill of every label of this tile
        if (c > 0) {
            if (lane == 0) while (rb_ld_acquire_s32(turn + f) != c) { }
            __syncwarp();
        }
        const int nk = s_nkeys[w];
        for (int k = lane; k < nk; k += 32) {
            const int p = keys[k];
            const int64_t slot = (int64_t)f * width + h[p].key;
            h[p].val = atomicAdd(fill + slot, h[p].val);           // before: the tile's count; after: points of the label in earlier tiles
        }
        __syncwarp();
        if (lane == 0) { __threadfence(); rb_st_release_s32(turn + f, c + 1); }
        // pass 2: rank inside the tile, batch by batch in point order, and scatter
        for (int b = 0; b < len; b += 32) {
            const int j = b + (int)lane;
            const int key = j < len ? labels[p0 + j] + 1 : 0;
            const unsigned active = __ballot_sync(0xffffffffu, key > 0);
            if (key > 0) {
                const unsigned same = __match_any_sync(active, key);
                const int p = hash_find(h, key);
                const int base = h[p].val;
                const int64_t dst = (int64_t)slot_start[(int64_t)f * width + key] + base + __popc(same & rb_lanemask_lt());
                if (gx) gx[dst] = x[p0 + j];
                if (gy) gy[dst] = y[p0 + j];
                gi[dst] = inten[p0 + j];
                __syncwarp(same);
                if ((int)lane == __ffs(same) - 1) h[p].val = base + __popc(same);
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// (4) numpy's pairwise float32 sum of a contiguous array (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum): up to
// 128 elements with eight partial sums, above that the two halves (the first a multiple of 8 long) summed separately.
// The recursion is unrolled into an explicit stack (depth <= log2(n / 128) + 1).
__device__ __forceinline__ float pairwise_leaf_f32(const float* __restrict__ a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
        return r;
    }
    float r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 = __fadd_rn(r0, a[i]);     r1 = __fadd_rn(r1, a[i + 1]); r2 = __fadd_rn(r2, a[i + 2]); r3 = __fadd_rn(r3, a[i + 3]);
        r4 = __fadd_rn(r4, a[i + 4]); r5 = __fadd_rn(r5, a[i + 5]); r6 = __fadd_rn(r6, a[i + 6]); r7 = __fadd_rn(r7, a[i + 7]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3)), __fadd_rn(__fadd_rn(r4, r5), __fadd_rn(r6, r7)));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

__device__ float pairwise_sum_f32(const float* __restrict__ a, int n) {
    int s_off[32], s_n[32];
    float s_left[32];
    unsigned char s_state[32];              // 0 = not expanded, 1 = waiting for the left half, 2 = waiting for the right half
    int sp = 1;
    s_off[0] = 0; s_n[0] = n; s_state[0] = 0;
    float ret = 0.f;
    while (sp > 0) {
        const int t = sp - 1;
        if (s_n[t] > 128) {                 // first visit of an inner node: descend into the left half
            int n2 = s_n[t] / 2;
            n2 -= n2 % 8;
            s_state[t] = 1;
            s_off[sp] = s_off[t]; s_n[sp] = n2; s_state[sp] = 0;
            ++sp;
            continue;
        }
        ret = pairwise_leaf_f32(a + s_off[t], s_n[t]);
        --sp;
        while (sp > 0) {                    // hand the finished sum to the parents
            const int q = sp - 1;
            if (s_state[q] == 1) {          // it was the left half: keep it, descend into the right half
                int n2 = s_n[q] / 2;
                n2 -= n2 % 8;
                s_left[q] = ret;
                s_state[q] = 2;
                s_off[sp] = s_off[q] + n2; s_n[sp] = s_n[q] - n2; s_state[sp] = 0;
                ++sp;
                break;
            }
            ret = __fadd_rn(s_left[q], ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(128) cl_reduce_kernel(const float* __restrict__ gx, const float* __restrict__ gy, const float* __restrict__ gi,
                                                       const int32_t* __restrict__ seg_label, const int32_t* __restrict__ seg_count,
                                                       const int64_t* __restrict__ seg_start, const int32_t* __restrict__ n_seg_dev, int64_t cap,
                                                       float* __restrict__ cx, float* __restrict__ cy, float* __restrict__ mi) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_seg = min((int64_t)*n_seg_dev, cap);
    if (k >= n_seg) return;
    if (seg_label[k] < 0) { cx[k] = 0.f; cy[k] = 0.f; mi[k] = 0.f; return; }
    const int n = seg_count[k];
    const float* __restrict__ px = gx + seg_start[k];
    const float* __restrict__ py = gy + seg_start[k];
    float sx = px[0], sy = py[0];                                    // np.add.reduce(axis=0): rows in order, one accumulator per column
    for (int i = 1; i < n; ++i) { sx = __fadd_rn(sx, px[i]); sy = __fadd_rn(sy, py[i]); }
    const float fn = (float)n;                                       // float32 / python int -> float32 true_divide (NEP 50)
    cx[k] = __fdiv_rn(sx, fn);
    cy[k] = __fdiv_rn(sy, fn);
    mi[k] = __fdiv_rn(pairwise_sum_f32(gi + seg_start[k], n), fn);
}

}  // namespace

// The grouping itself (steps 1-3), shared by rb_cluster_records and by the ordered land accumulation (land.cu, cells as
// labels, one "frame"): segment table + the values grouped by (frame, label) in their original order. Optional outputs
// may be NULL. Syncs once (segment count / capacity check). *n_seg_dev_out: device copy of the segment count.
static int group_core(rb_ctx* ctx, const float* x, const float* y, const float* inten, const int32_t* labels, int64_t n,
                      const int64_t* frame_off, int64_t n_frames, int64_t n_labels, int32_t* seg_frame, int32_t* seg_label,
                      int32_t* seg_first, int32_t* seg_count, int64_t* seg_start, int64_t cap_segments, float* gx, float* gy, float* gi,
                      int64_t* n_segments, int64_t* n_grouped, const int32_t** n_seg_dev_out, cudaStream_t stream) {
    const int64_t width = n_labels + 1;                              // labels -1 .. n_labels-1
    const int64_t m = n_frames * width;
    if (m > ((int64_t)1 << 27)) {
        rb_set_error("cluster records: %lld frames x %lld labels exceed the slot table; pass fewer frames per call (frames are independent)",
                     (long long)n_frames, (long long)width);
        return RB_ERR_CAPACITY;
    }
    // scratch: cnt, first, flag/rank (m + 1 each), slot_start, fill (m), tiles / tile_base (F + 1), segment scans (cap), counters
    const size_t M1 = (size_t)m + 1, F1 = (size_t)n_frames + 1, S1 = (size_t)cap_segments + 1;
    void* raw;
    RB_TRY(rb_scratch_get(ctx, RB_S_CLUSTERS, sizeof(int32_t) * (5 * M1 + 3 * F1 + 2 * S1 + 64), &raw));
    int32_t* cnt = (int32_t*)raw;
    int32_t* first = cnt + M1;
    int32_t* rank = first + M1;
    int32_t* slot_start = rank + M1;
    int32_t* fill = slot_start + M1;
    int32_t* tiles = fill + M1;
    int32_t* tile_base = tiles + F1;
    int32_t* turn = tile_base + F1;
    int32_t* seg_grouped = turn + F1;
    int32_t* seg_start32 = seg_grouped + S1;
    int32_t* misc = seg_start32 + S1;                                // [0] = bad label flag, [1] = ticket, [2] = grouped total, [3] = segments
    RB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * M1, stream));
    RB_CUDA(cudaMemsetAsync(first, 0x7f, sizeof(int32_t) * M1, stream));
    RB_CUDA(cudaMemsetAsync(fill, 0, sizeof(int32_t) * M1, stream));
    RB_CUDA(cudaMemsetAsync(turn, 0, sizeof(int32_t) * F1, stream));
    RB_CUDA(cudaMemsetAsync(seg_grouped, 0, sizeof(int32_t) * S1, stream));
    RB_CUDA(cudaMemsetAsync(misc, 0, sizeof(int32_t) * 64, stream));
    const unsigned pblocks = (unsigned)rb_div_up(n, CL_THREADS), mblocks = (unsigned)rb_div_up(m + 1, CL_THREADS);
    RB_CUDA(rb_launch(ctx, cl_count_kernel, dim3(pblocks), dim3(CL_THREADS), 0, stream, labels, n, frame_off, (int)n_frames, (int)width, cnt, first, misc));
    RB_LAUNCH_CHECK(ctx);
    RB_CUDA(rb_launch(ctx, cl_flag_kernel, dim3(mblocks), dim3(CL_THREADS), 0, stream, (const int32_t*)cnt, m, rank));
    RB_LAUNCH_CHECK(ctx);
    int32_t* n_seg_dev = misc + 3;
    RB_TRY(rb_exclusive_scan_i32(ctx, rank, rank, m + 1, n_seg_dev, stream));
    RB_CUDA(rb_launch(ctx, cl_compact_kernel, dim3(mblocks), dim3(CL_THREADS), 0, stream, (const int32_t*)cnt, (const int32_t*)first,
                      (const int32_t*)rank, m, (int)width, cap_segments, seg_frame, seg_label, seg_first, seg_count, seg_grouped));
    RB_LAUNCH_CHECK(ctx);
    if (cap_segments > 0) RB_TRY(rb_exclusive_scan_i32(ctx, seg_grouped, seg_start32, cap_segments, misc + 2, stream));
    RB_CUDA(rb_launch(ctx, cl_slot_start_kernel, dim3(mblocks), dim3(CL_THREADS), 0, stream, (const int32_t*)cnt, (const int32_t*)rank,
                      (const int32_t*)seg_start32, m, (int)width, cap_segments, slot_start, seg_start));
    RB_LAUNCH_CHECK(ctx);
    // read-back: number of segments (capacity check before anything is scattered), grouped points, bad-label flag
    int32_t* h = (int32_t*)ctx->pinned;
    RB_CUDA(cudaMemcpyAsync(h, misc, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    if (h[0]) { rb_set_error("cluster records: a label outside [-1, n_labels)"); return RB_ERR_ARG; }
    *n_segments = h[3];
    *n_grouped = h[2];
    if (h[3] > cap_segments) {
        rb_set_error("cluster records: %d segments need a table of at least that size (cap = %lld)", h[3], (long long)cap_segments);
        return RB_ERR_CAPACITY;
    }
    RB_CUDA(rb_launch(ctx, cl_tiles_kernel, dim3((unsigned)rb_div_up(n_frames + 1, CL_THREADS)), dim3(CL_THREADS), 0, stream, frame_off, (int)n_frames, tiles));
    RB_LAUNCH_CHECK(ctx);
    RB_TRY(rb_exclusive_scan_i32(ctx, tiles, tile_base, n_frames + 1, nullptr, stream));
    const int64_t max_tiles = n / CL_TILE + n_frames;
    const int64_t want = rb_div_up(max_tiles, CL_WARPS);
    const unsigned blocks = (unsigned)(want < (int64_t)ctx->sm_count * 8 ? (want > 0 ? want : 1) : (int64_t)ctx->sm_count * 8);
    RB_CUDA(rb_launch(ctx, cl_partition_kernel, dim3(blocks), dim3(CL_WARPS * 32), 0, stream, x, y, inten, labels, frame_off, (int)n_frames, (int)width,
                      (const int32_t*)tile_base, (const int32_t*)slot_start, fill, turn, (unsigned*)(misc + 1), gx, gy, gi));
    RB_LAUNCH_CHECK(ctx);
    if (n_seg_dev_out) *n_seg_dev_out = n_seg_dev;
    return RB_OK;
}

extern "C" int rb_cluster_records(rb_ctx* ctx, const float* x, const float* y, const float* inten, const int32_t* labels, int64_t n,
                                  const int64_t* frame_off, int64_t n_frames, int64_t n_clusters, const rb_cluster_table* tab,
                                  int64_t cap_segments, float* gx, float* gy, float* gi, int64_t* n_segments, int64_t* n_grouped,
                                  void* stream_) {
    RB_REQUIRE(ctx && tab && n_segments && n_grouped, "NULL argument");
    RB_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && n_frames >= 0 && n_frames < ((int64_t)1 << 30) && n_clusters >= 0 && cap_segments >= 0,
               "bad sizes");
    cudaStream_t stream = (cudaStream_t)stream_;
    *n_segments = 0;
    *n_grouped = 0;
    if (n == 0 || n_frames == 0) return RB_OK;
    RB_REQUIRE(x && y && inten && labels && frame_off && gx && gy && gi, "NULL buffers");
    RB_REQUIRE(tab->frame && tab->label && tab->first && tab->count && tab->start && tab->cx && tab->cy && tab->mean_intensity,
               "NULL table columns");
    const int32_t* n_seg_dev = nullptr;
    RB_TRY(group_core(ctx, x, y, inten, labels, n, frame_off, n_frames, n_clusters, tab->frame, tab->label, tab->first, tab->count, tab->start,
                      cap_segments, gx, gy, gi, n_segments, n_grouped, &n_seg_dev, stream));
    if (*n_segments > 0) {
        RB_CUDA(rb_launch(ctx, cl_reduce_kernel, dim3((unsigned)rb_div_up(*n_segments, 128)), dim3(128), 0, stream, (const float*)gx, (const float*)gy,
                          (const float*)gi, (const int32_t*)tab->label, (const int32_t*)tab->count, (const int64_t*)tab->start, n_seg_dev,
                          cap_segments, tab->cx, tab->cy, tab->mean_intensity));
        RB_LAUNCH_CHECK(ctx);
    }
    return RB_OK;
}

// values[n] grouped by labels[n] (0 .. n_labels-1) in their original order: one segment per label that occurs
// (ascending), seg_start = its offset in `grouped`. Used by the ordered land accumulation (land.cu). Syncs once.
int rb_stable_group(rb_ctx* ctx, const float* values, const int32_t* labels, int64_t n, int64_t n_labels, int32_t* seg_label,
                    int32_t* seg_count, int64_t* seg_start, int64_t cap_segments, float* grouped, int64_t* n_segments, cudaStream_t stream) {
    void* fo;
    RB_TRY(rb_scratch_get(ctx, RB_S_GROUP_TMP, 64, &fo));
    int64_t* h = (int64_t*)((unsigned char*)ctx->pinned + 256);      // (the first bytes of the staging buffer take the read-back)
    h[0] = 0; h[1] = n;
    RB_CUDA(cudaMemcpyAsync(fo, h, sizeof(int64_t) * 2, cudaMemcpyHostToDevice, stream));
    int64_t n_grouped = 0;
    return group_core(ctx, nullptr, nullptr, values, labels, n, (const int64_t*)fo, 1, n_labels, nullptr, seg_label, nullptr, seg_count, seg_start,
                      cap_segments, nullptr, nullptr, grouped, n_segments, &n_grouped, nullptr, stream);
}
