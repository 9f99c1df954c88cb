// The whole hot path for one block of frames behind ONE C-ABI call (rb_detect_block): what the reference's
// run_pipeline does between "CSV parsed" and "labels known" (4_temporal_object_tracker.py:941-977).
//
// The stage kernels are the ones behind the per-function entry points (spoke.cu, land.cu, dbscan.cu); this
// file is the native host driver around them: it sizes nothing from Python, reads back only what the host
// must know (point count + bounds after the spoke stage; the filtered count; the final counters) and
// reproduces numpy's np.arange edges bit for bit, so a block costs three stream syncs and no interpreter
// time between launches.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

// np.arange(lo, fl32(hi + step), step) with np.float32 scalars lo, hi and a Python float step, as
// build_occupancy_grid calls it (T4:372-373). numpy (multiarray/ctors.c, _calc_length + DOUBLE_fill):
//   length = ceil(fl32(fl32(stop - start) / fl32(step)))     start/stop are float32 scalars, step is "weak"
//   a[0] = start, a[1] = fl32(start + fl32(step)), a[i] = a[0] + i * (a[1] - a[0])   in float64
extern "C" int64_t rb_arange_edges(float lo, float hi, double step, double* out, int64_t cap) {
    const volatile float step32 = (float)step;
    const volatile float stop = hi + step32;
    const volatile float span = stop - lo;
    const volatile float val = span / step32;
    const double len_d = ceil((double)val);
    if (!(len_d > 0)) return 0;
    const int64_t len = (int64_t)len_d;
    const volatile float next32 = lo + step32;
    const volatile double start = (double)lo;
    const volatile double delta = (double)next32 - start;
    for (int64_t i = 0; i < len && i < cap; ++i) {
        if (i == 0) out[i] = start;
        else if (i == 1) out[i] = (double)next32;
        else {
            const volatile double prod = (double)i * delta;        // two roundings, like DOUBLE_fill (no FMA)
            out[i] = start + prod;
        }
    }
    return len;
}

// Host side of the time-sharded path: global cluster numbering from the ranks' local components (sharded.py,
// stitch_components). keys = every rank's distinct component keys (a key = the smallest global core index the rank saw
// in the component), pair_a/pair_b = keys of the same boundary core point as seen by two neighbouring ranks: each pair
// ties two local components together. Union-find over the sorted distinct keys with "smaller index wins", so a set's
// root is its smallest key; the id of a set is the rank of that key among all roots - the reference's numbering
// (SURVEY.md N4). Returns the number of distinct keys (table size), or a negative error code.
extern "C" int64_t rb_stitch_components(const int64_t* keys, int64_t n_keys, const int64_t* pair_a, const int64_t* pair_b,
                                        int64_t n_pairs, int64_t* table_keys, int32_t* table_ids, int64_t cap,
                                        int64_t* n_clusters) {
    if (n_keys < 0 || n_pairs < 0 || (n_keys && !keys) || (n_pairs && (!pair_a || !pair_b)) || !n_clusters) {
        rb_set_error("rb_stitch_components: bad arguments");
        return RB_ERR_ARG;
    }
    std::vector<int64_t> uniq(keys, keys + n_keys);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    const int64_t m = (int64_t)uniq.size();
    *n_clusters = 0;
    if (m > cap || (m && (!table_keys || !table_ids))) {
        rb_set_error("rb_stitch_components: %lld distinct keys need a table of at least that size (cap = %lld)", (long long)m, (long long)cap);
        return RB_ERR_CAPACITY;
    }
    std::vector<int32_t> parent((size_t)m);
    for (int64_t i = 0; i < m; ++i) parent[(size_t)i] = (int32_t)i;
    auto find = [&](int32_t x) {
        while (parent[(size_t)x] != x) { parent[(size_t)x] = parent[(size_t)parent[(size_t)x]]; x = parent[(size_t)x]; }
        return x;
    };
    for (int64_t p = 0; p < n_pairs; ++p) {
        const auto ia = std::lower_bound(uniq.begin(), uniq.end(), pair_a[p]);
        const auto ib = std::lower_bound(uniq.begin(), uniq.end(), pair_b[p]);
        if (ia == uniq.end() || *ia != pair_a[p] || ib == uniq.end() || *ib != pair_b[p]) {
            rb_set_error("rb_stitch_components: pair %lld names a key that no rank listed", (long long)p);
            return RB_ERR_ARG;
        }
        int32_t ra = find((int32_t)(ia - uniq.begin())), rb = find((int32_t)(ib - uniq.begin()));
        if (ra == rb) continue;
        if (ra < rb) parent[(size_t)rb] = ra; else parent[(size_t)ra] = rb;
    }
    std::vector<int32_t> root_rank((size_t)m);
    int32_t n_roots = 0;
    for (int64_t i = 0; i < m; ++i) root_rank[(size_t)i] = parent[(size_t)i] == (int32_t)i ? n_roots++ : -1;
    for (int64_t i = 0; i < m; ++i) {
        table_keys[i] = uniq[(size_t)i];
        table_ids[i] = root_rank[(size_t)find((int32_t)i)];
    }
    *n_clusters = n_roots;
    return m;
}

extern "C" int rb_detect_block(rb_ctx* ctx, const float* echo, const float* cos_tab, const float* sin_tab,
                               const float* range_res, const int32_t* sweep_gain, const float* frame_ids,
                               const rb_detect_params* prm, const rb_detect_buffers* buf, rb_detect_result* res,
                               void* stream_) {
    RB_REQUIRE(ctx && prm && buf && res, "NULL argument");
    RB_REQUIRE(prm->n_frames >= 0 && prm->gains_per_frame >= 1 && prm->n_spokes >= 0 && prm->n_bins >= 0, "bad sizes");
    RB_REQUIRE(buf->cap >= 0 && buf->frame_off && buf->f_frame_off, "bad buffers");
    cudaStream_t stream = (cudaStream_t)stream_;
    memset(res, 0, sizeof *res);
    res->filtered_is_raw = 1;
    const int64_t F = prm->n_frames;
    const int64_t W = F * prm->gains_per_frame;

    // ---- a1 + a2: spoke-to-point, frame offsets, bounds of however many points there are -------------------
    void* sb_v;
    RB_TRY(rb_scratch_get(ctx, RB_S_PIPE, sizeof(int64_t) * (size_t)(W + 1) + 64, &sb_v));
    int64_t* sweep_base = (int64_t*)sb_v;
    float* d_bounds = (float*)(sweep_base + W + 1);
    if (prm->echo_u8)
        RB_TRY(rb_spoke_to_points_u8(ctx, reinterpret_cast<const uint8_t*>(echo), cos_tab, sin_tab, range_res, nullptr, sweep_gain, W,
                                     prm->n_spokes, prm->n_bins, prm->intensity_threshold, prm->point_stride, buf->x, buf->y,
                                     buf->inten, buf->gain, buf->cap, sweep_base, stream_));
    else
        RB_TRY(rb_spoke_to_points(ctx, echo, cos_tab, sin_tab, range_res, nullptr, sweep_gain, W, prm->n_spokes, prm->n_bins,
                                  prm->intensity_threshold, prm->point_stride, buf->x, buf->y, buf->inten, buf->gain, buf->cap,
                                  sweep_base, stream_));
    RB_TRY(rb_frame_offsets(ctx, sweep_base, F, prm->gains_per_frame, buf->frame_off, stream_));
    if (buf->cap > 0) RB_TRY(rb_bounds_devn(ctx, buf->x, buf->y, sweep_base + W, buf->cap, d_bounds, stream));
    // read-back 1: frame offsets (point count, frames built) + bounds
    // pinned staging of the context: [frame offsets | bounds | grid edges (x, y) | frame ids]. Everything the host uploads
    // below goes through it, so no copy waits for the stream the way a copy from pageable memory does; a region is
    // rewritten only after a stream sync.
    const size_t off_bytes = sizeof(int64_t) * (size_t)(F + 1);
    const size_t edge_cap = (size_t)(buf->max_edges > 0 ? buf->max_edges : 0);
    const size_t edges_at = off_bytes + 64, ids_at = edges_at + sizeof(double) * 2 * edge_cap;
    const size_t pinned_need = ids_at + sizeof(float) * (size_t)F + 64;
    if (ctx->pinned_cap < pinned_need) {
        RB_CUDA(cudaStreamSynchronize(stream));
        if (ctx->pinned) RB_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr; ctx->pinned_cap = 0;
        RB_CUDA(cudaMallocHost(&ctx->pinned, pinned_need + 4096));
        ctx->pinned_cap = pinned_need + 4096;
    }
    int64_t* h_off = (int64_t*)ctx->pinned;
    float* h_bounds = (float*)((unsigned char*)ctx->pinned + off_bytes);
    double* h_xe = (double*)((unsigned char*)ctx->pinned + edges_at);
    double* h_ye = h_xe + edge_cap;
    float* h_ids = (float*)((unsigned char*)ctx->pinned + ids_at);
    RB_CUDA(cudaMemcpyAsync(h_off, buf->frame_off, off_bytes, cudaMemcpyDeviceToHost, stream));
    if (buf->cap > 0) RB_CUDA(cudaMemcpyAsync(h_bounds, d_bounds, sizeof(float) * 4, cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    const int64_t n_raw = h_off[F];
    res->n_raw = n_raw;
    res->n_points = n_raw;
    int built = 0;
    for (int64_t f = 0; f < F; ++f) built += h_off[f + 1] > h_off[f];
    res->frames_built = built;
    if (n_raw > buf->cap) {
        rb_set_error("rb_detect_block: %lld points need a capacity of at least that (cap = %lld)", (long long)n_raw, (long long)buf->cap);
        return RB_ERR_CAPACITY;
    }
    if (n_raw > 0) memcpy(res->bounds, h_bounds, sizeof(float) * 4);
    float fmin_id = 0.f, fmax_id = 0.f;
    bool ids_integer = true;
    for (int64_t f = 0; f < F; ++f) {
        const float v = frame_ids[f];
        if (f == 0 || v < fmin_id) fmin_id = v;
        if (f == 0 || v > fmax_id) fmax_id = v;
        ids_integer = ids_integer && v == rintf(v) && fabsf(v) < 8388608.f;
    }

    // ---- a4-a6: land / stationary persistence filter ----------------------------------------------------------
    const float *px = buf->x, *py = buf->y, *pz = buf->inten;
    const int64_t* p_off = buf->frame_off;
    int64_t n_pts = n_raw;
    if (prm->land_filter && n_raw > 0 && built > prm->land_min_frames) {
        const int64_t nxe = rb_arange_edges(res->bounds[0], res->bounds[1], prm->land_resolution, h_xe, buf->max_edges);
        const int64_t nye = rb_arange_edges(res->bounds[2], res->bounds[3], prm->land_resolution, h_ye, buf->max_edges);
        res->n_x_edges = (int32_t)nxe;
        res->n_y_edges = (int32_t)nye;
        const int64_t cells = (nxe - 1) * (nye - 1);
        if (nxe > buf->max_edges || nye > buf->max_edges || cells > buf->max_cells || !buf->x_edges || !buf->y_edges) {
            rb_set_error("rb_detect_block: land grid of %lld x %lld edges exceeds max_edges %d / max_cells %lld", (long long)nxe,
                         (long long)nye, buf->max_edges, (long long)buf->max_cells);
            return RB_ERR_CAPACITY;
        }
        RB_REQUIRE(nxe >= 2 && nye >= 2, "degenerate land grid");
        memcpy(buf->x_edges, h_xe, sizeof(double) * (size_t)nxe);          // the caller's copy (an output)
        memcpy(buf->y_edges, h_ye, sizeof(double) * (size_t)nye);
        RB_REQUIRE(buf->count && buf->isum && buf->land && buf->fx && buf->fy && buf->finten && buf->fgain, "NULL land buffers");
        void* e_v;
        RB_TRY(rb_scratch_get(ctx, RB_S_PIPE_EDGES, sizeof(double) * (size_t)(nxe + nye), &e_v));
        double* d_xe = (double*)e_v;
        double* d_ye = d_xe + nxe;
        RB_CUDA(cudaMemcpyAsync(d_xe, h_xe, sizeof(double) * (size_t)nxe, cudaMemcpyHostToDevice, stream));
        RB_CUDA(cudaMemcpyAsync(d_ye, h_ye, sizeof(double) * (size_t)nye, cudaMemcpyHostToDevice, stream));
        // the fast accumulation is exact for integer-valued intensities (what a radar delivers) and says so itself; if
        // the flag comes back with read-back 2, the land stage is repeated with the ordered accumulation (any input)
        int32_t* d_inexact;
        RB_TRY(rb_land_inexact_flag(ctx, &d_inexact, stream));
        for (int pass = 0; pass < 2; ++pass) {
            RB_CUDA(cudaMemsetAsync(buf->count, 0, sizeof(int32_t) * (size_t)cells, stream));
            RB_CUDA(cudaMemsetAsync(buf->isum, 0, sizeof(double) * (size_t)cells, stream));
            if (pass == 0)
                RB_TRY(rb_land_accumulate(ctx, buf->x, buf->y, buf->inten, n_raw, d_xe, (int)nxe, d_ye, (int)nye, buf->count, buf->isum, stream_));
            else
                RB_TRY(rb_land_accumulate_ordered(ctx, buf->x, buf->y, buf->inten, n_raw, d_xe, (int)nxe, d_ye, (int)nye, buf->count, buf->isum, stream_));
            RB_TRY(rb_land_cells(ctx, buf->count, buf->isum, cells, built, prm->land_persistence, prm->land_min_intensity, buf->land, stream_));
            RB_TRY(rb_land_filter(ctx, buf->x, buf->y, buf->inten, buf->gain, n_raw, buf->frame_off, F, d_xe, (int)nxe, d_ye, (int)nye,
                                  buf->land, buf->fx, buf->fy, buf->finten, buf->fgain, buf->f_frame_off, nullptr, stream_));
            // read-back 2: how many points are left (+ the "inexact" flag of the fast accumulation)
            RB_CUDA(cudaMemcpyAsync(h_off, buf->f_frame_off + F, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
            RB_CUDA(cudaMemcpyAsync(h_off + 1, d_inexact, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            RB_CUDA(cudaMemsetAsync(d_inexact, 0, sizeof(int32_t), stream));
            RB_CUDA(cudaStreamSynchronize(stream));
            if (pass == 1 || *(const int32_t*)(h_off + 1) == 0) break;
            res->land_ordered = 1;
        }
        n_pts = h_off[0];
        res->land_applied = 1;
        res->filtered_is_raw = 0;
        px = buf->fx; py = buf->fy; pz = buf->finten;
        p_off = buf->f_frame_off;
    }
    res->n_points = n_pts;

    // ---- a7: ST-DBSCAN, time = frame id -----------------------------------------------------------------------------
    if (prm->cluster && n_pts > 0) {
        RB_REQUIRE(buf->labels, "labels is NULL");
        void* t_v;
        RB_TRY(rb_scratch_get(ctx, RB_S_PIPE_TIMES, sizeof(float) * (size_t)(n_pts + F), &t_v));
        float* d_times = (float*)t_v;
        float* d_ids = d_times + n_pts;
        memcpy(h_ids, frame_ids, sizeof(float) * (size_t)F);
        RB_CUDA(cudaMemcpyAsync(d_ids, h_ids, sizeof(float) * (size_t)F, cudaMemcpyHostToDevice, stream));
        RB_TRY(rb_expand_frame_times(ctx, p_off, d_ids, F, n_pts, d_times, stream_));
        rb_stdbscan_hint hint;
        hint.lo[0] = res->bounds[0]; hint.hi[0] = res->bounds[1];      // filtered points lie inside the raw bounds
        hint.lo[1] = res->bounds[2]; hint.hi[1] = res->bounds[3];
        hint.lo[2] = hint.hi[2] = 0.f;
        hint.lo[3] = fmin_id; hint.hi[3] = fmax_id;
        hint.times_integer = ids_integer;
        // 3-D (z = intensity): the intensity range is not known on the host, so the plan measures its own bounds (one more sync)
        RB_TRY(rb_stdbscan_enqueue(ctx, px, py, prm->cluster_3d ? pz : nullptr, 1, d_times, n_pts, prm->eps_space, prm->eps_time,
                                   prm->min_samples, buf->labels, nullptr, prm->cluster_3d ? nullptr : &hint, stream_));
        int64_t ncl = 0;
        RB_TRY(rb_stdbscan_fetch_stats(ctx, &ncl, stream_));             // read-back 3 (final)
        res->n_clusters = ncl;
    } else {
        RB_CUDA(cudaStreamSynchronize(stream));
    }
    if (!ctx->retired.empty()) rb_trim(ctx);               // scratch outgrown during this block: nothing of ours is in flight now
    return RB_OK;
}
