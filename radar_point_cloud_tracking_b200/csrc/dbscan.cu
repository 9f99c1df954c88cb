// Temporal ST-DBSCAN (reference 4_temporal_object_tracker.py:443-506, twins 3_stdbscan_point_clouds.py:
// 101-136 and radar_pipeline/processors/clustering.py:49-115; native precedent
// radar-pipeline-rs/src/processors/clustering.rs:209-325).
//
// Result contract (SURVEY.md section 8 N4 — the reference's labels are canonical):
//   neighbour(p,q) <=> fl64 sum_d (double(p_d)-double(q_d))^2 <= eps^2  and  |t_p-t_q| <= eps_t (fl32)
//   core(p) <=> |N(p)| >= min_samples (self included)
//   cluster id = rank of the component's smallest core index; border = smallest id among its core
//   neighbours; noise = -1.
//
// Pipeline (all on the caller's stream):
//   bounds -> [host picks the grid] -> cell ids + counting sort into a dense (t, z, y, x) cell table
//   -> neighbour count with early exit -> core flags -> lock-free min-root union-find over
//   core-core edges -> component min original index -> rank (scan) -> labels -> border pass.
// Cells are at least eps wide (and time bins at least eps_t wide, or exactly one frame when the
// times are integers), so the 3^D x (2*tr+1) block of cells around a point holds all its
// neighbours; the exact predicates are always evaluated, the grid is only a candidate filter.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int DB_THREADS = 256;

struct DbGrid {
    double lo[3];
    double inv_cell;
    double tmin;
    double inv_wt;
    int n[3];          // nx, ny, nz
    int nt;
    int tr;            // time-bin search radius
    int dim;
};

struct DbPoints {      // strided view of the caller's coordinates
    const float* x; const float* y; const float* z;
    int64_t stride;
    const float* t;
};

// ---- bounds: order-preserving int encoding of floats, block reduce, then atomics ------------------
__device__ __forceinline__ int f2ord(float f) {
    int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__host__ __device__ __forceinline__ float ord2f(int o) {
    int b = o ^ ((o >> 31) & 0x7fffffff);
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// out[0..3] = min x,y,z,t ; out[4..7] = max ; out[8] = 1 if some time is not an integer
__global__ void __launch_bounds__(DB_THREADS) db_bounds_kernel(DbPoints p, int dim, int64_t n, int* __restrict__ out) {
    int mn[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
    int mx[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
    int nonint = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v[4];
        v[0] = p.x[i * p.stride];
        v[1] = dim > 1 ? p.y[i * p.stride] : 0.f;
        v[2] = dim > 2 ? p.z[i * p.stride] : 0.f;
        v[3] = p.t[i];
        nonint |= !(v[3] == rintf(v[3]) && fabsf(v[3]) < 8388608.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int o = f2ord(v[k]);
            mn[k] = min(mn[k], o);
            mx[k] = max(mx[k], o);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mn[k] = __reduce_min_sync(0xffffffffu, mn[k]);
        mx[k] = __reduce_max_sync(0xffffffffu, mx[k]);
    }
    nonint = __any_sync(0xffffffffu, nonint);
    if (rb_lane() == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { atomicMin(out + k, mn[k]); atomicMax(out + 4 + k, mx[k]); }
        if (nonint) atomicOr(out + 8, 1);
    }
}

__global__ void db_bounds_init(int* out) {
    if (threadIdx.x < 4) out[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 8) out[threadIdx.x] = INT_MIN;
    else if (threadIdx.x < 16) out[threadIdx.x] = 0;
}

// ---- grid cell of a point ---------------------------------------------------------------------------
__device__ __forceinline__ int axis_cell(float v, double lo, double inv, int n) {
    int c = (int)floor(((double)v - lo) * inv);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__device__ __forceinline__ int cell_index(const DbGrid& g, float x, float y, float z, float t) {
    int cx = axis_cell(x, g.lo[0], g.inv_cell, g.n[0]);
    int cy = g.dim > 1 ? axis_cell(y, g.lo[1], g.inv_cell, g.n[1]) : 0;
    int cz = g.dim > 2 ? axis_cell(z, g.lo[2], g.inv_cell, g.n[2]) : 0;
    int tb = axis_cell(t, g.tmin, g.inv_wt, g.nt);
    return ((tb * g.n[2] + cz) * g.n[1] + cy) * g.n[0] + cx;
}

// cell id per point + slot of the point inside its cell (the atomic's return value)
__global__ void __launch_bounds__(DB_THREADS) db_cell_kernel(DbPoints p, DbGrid g, int64_t n, int* __restrict__ cell_id,
                                                            int* __restrict__ slot, int* __restrict__ cell_count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = p.x[i * p.stride];
    float y = g.dim > 1 ? p.y[i * p.stride] : 0.f;
    float z = g.dim > 2 ? p.z[i * p.stride] : 0.f;
    int c = cell_index(g, x, y, z, p.t[i]);
    cell_id[i] = c;
    slot[i] = atomicAdd(cell_count + c, 1);
}

__global__ void __launch_bounds__(DB_THREADS) db_scatter_kernel(DbPoints p, int dim, int64_t n, const int* __restrict__ cell_id,
                                                               const int* __restrict__ slot, const int* __restrict__ cell_start,
                                                               int* __restrict__ sidx, int* __restrict__ scell,
                                                               float* __restrict__ sx, float* __restrict__ sy,
                                                               float* __restrict__ sz, float* __restrict__ st) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_id[i];
    int pos = cell_start[c] + slot[i];
    sidx[pos] = (int)i;
    scell[pos] = c;
    sx[pos] = p.x[i * p.stride];
    if (dim > 1) sy[pos] = p.y[i * p.stride];
    if (dim > 2) sz[pos] = p.z[i * p.stride];
    st[pos] = p.t[i];
}

// ---- neighbourhood walk ------------------------------------------------------------------------------
struct Sorted {
    const float* x; const float* y; const float* z; const float* t;
    const int* cell; const int* cell_start;
};

template <int DIM>
struct Pt { float x, y, z, t; };

template <int DIM>
__device__ __forceinline__ Pt<DIM> load_pt(const Sorted& s, int i) {
    Pt<DIM> p;
    p.x = s.x[i];
    p.y = DIM > 1 ? s.y[i] : 0.f;
    p.z = DIM > 2 ? s.z[i] : 0.f;
    p.t = s.t[i];
    return p;
}

template <int DIM>
__device__ __forceinline__ bool is_neighbour(const Pt<DIM>& a, const Pt<DIM>& b, double eps2, float eps_t) {
    float dt = __fsub_rn(b.t, a.t);                       // T4:486, float32
    if (!(fabsf(dt) <= eps_t)) return false;
    double d = (double)a.x - (double)b.x;                 // sklearn rdist: d += tmp*tmp, float64, no FMA
    double acc = __dmul_rn(d, d);
    if (DIM > 1) { d = (double)a.y - (double)b.y; acc = __dadd_rn(acc, __dmul_rn(d, d)); }
    if (DIM > 2) { d = (double)a.z - (double)b.z; acc = __dadd_rn(acc, __dmul_rn(d, d)); }
    return acc <= eps2;
}

// Calls f(q) for every sorted index q in the cells around `cell`; f returns false to stop.
template <int DIM, typename F>
__device__ __forceinline__ void for_each_candidate(const DbGrid& g, const int* __restrict__ cell_start, int cell, F&& f) {
    int cx = cell % g.n[0];
    int rest = cell / g.n[0];
    int cy = rest % g.n[1];
    rest /= g.n[1];
    int cz = rest % g.n[2];
    int tb = rest / g.n[2];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.n[0] - 1);
    const int y0 = DIM > 1 ? max(cy - 1, 0) : 0, y1 = DIM > 1 ? min(cy + 1, g.n[1] - 1) : 0;
    const int z0 = DIM > 2 ? max(cz - 1, 0) : 0, z1 = DIM > 2 ? min(cz + 1, g.n[2] - 1) : 0;
    const int t0 = max(tb - g.tr, 0), t1 = min(tb + g.tr, g.nt - 1);
    for (int tt = t0; tt <= t1; ++tt)
        for (int zz = z0; zz <= z1; ++zz)
            for (int yy = y0; yy <= y1; ++yy) {
                int row = ((tt * g.n[2] + zz) * g.n[1] + yy) * g.n[0];
                int b = cell_start[row + x0], e = cell_start[row + x1 + 1];
                for (int q = b; q < e; ++q)
                    if (!f(q)) return;
            }
}

__device__ __forceinline__ void add_counter(unsigned long long* ctr, unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (rb_lane() == 0 && v) atomicAdd(ctr, v);
}

// core[p]: 1 = core, 0 = not core, 2 = not core and alone (no neighbour but itself)
template <int DIM>
__global__ void __launch_bounds__(DB_THREADS) db_count_kernel(Sorted s, DbGrid g, int n, double eps2, float eps_t,
                                                             int min_samples, uint8_t* __restrict__ core,
                                                             int* __restrict__ parent, unsigned long long* __restrict__ ctr) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long tests = 0;
    if (p < n) {
        Pt<DIM> a = load_pt<DIM>(s, p);
        int cnt = 0;
        for_each_candidate<DIM>(g, s.cell_start, s.cell[p], [&](int q) {
            ++tests;
            cnt += is_neighbour<DIM>(a, load_pt<DIM>(s, q), eps2, eps_t);
            return cnt < min_samples || cnt < 2;          // keep going until core is certain
        });
        core[p] = cnt >= min_samples ? 1 : (cnt <= 1 ? 2 : 0);
        parent[p] = p;
    }
    add_counter(ctr, tests);
}

__device__ __forceinline__ int uf_find(int* parent, int a) {
    int cur = a;
    while (true) {
        int p = rb_ld_relaxed_s32(parent + cur);
        if (p == cur) return cur;
        int gp = rb_ld_relaxed_s32(parent + p);
        if (gp != p) parent[cur] = gp;                    // path halving; any ancestor is a valid parent
        cur = p;
    }
}

// roots only ever move to SMALLER indices, so the structure stays acyclic under races
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        if (atomicCAS(parent + a, a, b) == a) return;
    }
}

template <int DIM>
__global__ void __launch_bounds__(DB_THREADS) db_union_kernel(Sorted s, DbGrid g, int n, double eps2, float eps_t,
                                                             const uint8_t* __restrict__ core, int* __restrict__ parent,
                                                             unsigned long long* __restrict__ ctr) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long tests = 0;
    if (p < n && core[p] == 1) {
        Pt<DIM> a = load_pt<DIM>(s, p);
        for_each_candidate<DIM>(g, s.cell_start, s.cell[p], [&](int q) {
            if (q < p && core[q] == 1) {                  // each core-core edge once
                ++tests;
                if (is_neighbour<DIM>(a, load_pt<DIM>(s, q), eps2, eps_t)) uf_union(parent, p, q);
            }
            return true;
        });
    }
    add_counter(ctr, tests);
}

__global__ void __launch_bounds__(DB_THREADS) db_minorig_kernel(int n, const uint8_t* __restrict__ core, int* __restrict__ parent,
                                                               const int* __restrict__ sidx, int* __restrict__ minorig) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || core[p] != 1) return;
    int r = uf_find(parent, p);
    parent[p] = r;
    atomicMin(minorig + r, sidx[p]);
}

__global__ void __launch_bounds__(DB_THREADS) db_rootflag_kernel(int n, const uint8_t* __restrict__ core,
                                                                const int* __restrict__ parent, const int* __restrict__ minorig,
                                                                int* __restrict__ flags) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || core[p] != 1 || parent[p] != p) return;
    flags[minorig[p]] = 1;
}

__global__ void __launch_bounds__(DB_THREADS) db_label_kernel(int n, const uint8_t* __restrict__ core, const int* __restrict__ parent,
                                                             const int* __restrict__ minorig, const int* __restrict__ rank,
                                                             const int* __restrict__ sidx, int* __restrict__ slabel,
                                                             int32_t* __restrict__ labels, uint8_t* __restrict__ core_out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int lab = -1;
    if (core[p] == 1) lab = rank[minorig[parent[p]]];
    slabel[p] = lab;
    int o = sidx[p];
    labels[o] = lab;
    if (core_out) core_out[o] = core[p] == 1;
}

template <int DIM>
__global__ void __launch_bounds__(DB_THREADS) db_border_kernel(Sorted s, DbGrid g, int n, double eps2, float eps_t,
                                                              const uint8_t* __restrict__ core, const int* __restrict__ slabel,
                                                              const int* __restrict__ sidx, int32_t* __restrict__ labels,
                                                              unsigned long long* __restrict__ ctr) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long tests = 0;
    if (p < n && core[p] == 0) {
        Pt<DIM> a = load_pt<DIM>(s, p);
        int best = INT_MAX;
        for_each_candidate<DIM>(g, s.cell_start, s.cell[p], [&](int q) {
            if (core[q] == 1) {
                int lq = slabel[q];
                if (lq < best) {                          // only a smaller id can change the answer
                    ++tests;
                    if (is_neighbour<DIM>(a, load_pt<DIM>(s, q), eps2, eps_t)) best = lq;
                }
            }
            return best != 0;                             // id 0 cannot be beaten
        });
        if (best != INT_MAX) labels[sidx[p]] = best;
    }
    add_counter(ctr, tests);
}

template <typename T>
int scratch(rb_ctx* ctx, rb_slot slot, size_t count, T** out) {
    void* p;
    int rc = rb_scratch_get(ctx, slot, sizeof(T) * count, &p);
    *out = (T*)p;
    return rc;
}

// Host side: choose cell size / time binning so that the dense cell table stays affordable.
int choose_grid(const float mn[4], const float mx[4], bool times_integer, int dim, int64_t n, double eps_space,
                float eps_time, DbGrid* g, double* cell_out, double* wt_out) {
    double ext[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) { g->lo[k] = 0; g->n[k] = 1; }
    for (int k = 0; k < dim; ++k) {
        g->lo[k] = mn[k];
        ext[k] = (double)mx[k] - (double)mn[k];
        if (!(ext[k] >= 0) || !isfinite(ext[k])) { rb_set_error("rb_stdbscan: non-finite coordinates"); return RB_ERR_ARG; }
    }
    double text = (double)mx[3] - (double)mn[3];
    if (!(text >= 0) || !isfinite(text)) { rb_set_error("rb_stdbscan: non-finite times"); return RB_ERR_ARG; }
    double max_ext = fmax(ext[0], fmax(ext[1], ext[2]));
    double cell = eps_space > 0 ? eps_space * (1.0 + 1e-7) : (max_ext > 0 ? max_ext / 1024.0 : 1.0);
    if (!(cell > 0) || !isfinite(cell)) cell = 1.0;
    // time bins
    double wt; int tr;
    bool unit_bins = false;
    double et = (double)eps_time;
    if (!(et >= 0)) et = 0;                                 // negative eps_time: nothing matches anyway
    if (times_integer) {
        wt = 1.0; unit_bins = true;
        tr = (int)fmin(floor(et), 1e6);
    } else if (et > 0) {
        wt = et * (1.0 + 1e-6); tr = 1;
    } else {
        wt = text > 0 ? text / 256.0 : 1.0; tr = 0;         // equal times always share a bin
    }
    const double budget = fmin(fmax(16.0 * (double)n, 1048576.0), 134217728.0);
    for (int iter = 0; iter < 200; ++iter) {
        double total = 1;
        for (int k = 0; k < dim; ++k) { double c = floor(ext[k] / cell) + 1; total *= c; }
        double nt = floor(text / wt) + 1;
        total *= nt;
        if (total <= budget) {
            for (int k = 0; k < dim; ++k) g->n[k] = (int)(floor(ext[k] / cell) + 1);
            g->nt = (int)nt;
            break;
        }
        // coarsen whichever axis family has more cells per search step
        double per_axis = 1;
        for (int k = 0; k < dim; ++k) per_axis = fmax(per_axis, floor(ext[k] / cell) + 1);
        if (nt / (2.0 * tr + 1.0) > per_axis / 3.0 && nt > 1) {
            wt *= 2.0;
            if (unit_bins) { unit_bins = false; }
            tr = (int)fmin(floor(et / wt) + 1, 1e6);
        } else {
            cell *= 2.0;
        }
        if (iter == 199) { rb_set_error("rb_stdbscan: could not fit a cell table"); return RB_ERR_ARG; }
    }
    g->inv_cell = 1.0 / cell;
    g->tmin = mn[3];
    g->inv_wt = 1.0 / wt;
    g->tr = tr;
    g->dim = dim;
    *cell_out = cell;
    *wt_out = wt;
    return RB_OK;
}

template <int DIM>
int run_dbscan(rb_ctx* ctx, const DbPoints& pts, int64_t n64, double eps_space, float eps_time, int min_samples,
               int32_t* labels, uint8_t* core_out, int64_t* n_clusters, cudaStream_t stream) {
    const int n = (int)n64;
    const unsigned blocks = (unsigned)rb_div_up(n, DB_THREADS);

    // 1. bounds -> host
    int* d_bounds;
    RB_TRY(scratch(ctx, RB_S_MISC, 64, &d_bounds));
    unsigned long long* d_ctr = (unsigned long long*)(d_bounds + 16);      // 3 counters + n_clusters slot
    db_bounds_init<<<1, 32, 0, stream>>>(d_bounds);
    RB_LAUNCH_CHECK(ctx);
    RB_CUDA(cudaMemsetAsync(d_ctr, 0, sizeof(unsigned long long) * 4, stream));
    int bblocks = (int)(blocks < (unsigned)ctx->sm_count * 8 ? blocks : (unsigned)ctx->sm_count * 8);
    db_bounds_kernel<<<bblocks, DB_THREADS, 0, stream>>>(pts, DIM, n64, d_bounds);
    RB_LAUNCH_CHECK(ctx);
    int* h = (int*)ctx->pinned;
    RB_CUDA(cudaMemcpyAsync(h, d_bounds, sizeof(int) * 9, cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    float mn[4], mx[4];
    for (int k = 0; k < 4; ++k) { mn[k] = ord2f(h[k]); mx[k] = ord2f(h[4 + k]); }
    bool times_integer = h[8] == 0;

    // 2. grid
    DbGrid g;
    double cell, wt;
    RB_TRY(choose_grid(mn, mx, times_integer, DIM, n64, eps_space, eps_time, &g, &cell, &wt));
    const int64_t n_cells = (int64_t)g.n[0] * g.n[1] * g.n[2] * g.nt;

    // 3. counting sort into the cell table
    int *cell_id, *slot, *cell_start, *sidx, *scell, *parent, *minorig, *flags, *rank, *slabel;
    float *sx, *sy, *sz, *st;
    uint8_t* core;
    RB_TRY(scratch(ctx, RB_S_CELL_ID, (size_t)n, &cell_id));
    RB_TRY(scratch(ctx, RB_S_CELL_FILL, (size_t)n, &slot));
    RB_TRY(scratch(ctx, RB_S_CELL_START, (size_t)n_cells + 1, &cell_start));
    RB_TRY(scratch(ctx, RB_S_SORT_IDX, (size_t)n * 2, &sidx));
    scell = sidx + n;
    RB_TRY(scratch(ctx, RB_S_SX, (size_t)n, &sx));
    RB_TRY(scratch(ctx, RB_S_SY, (size_t)n, &sy));
    RB_TRY(scratch(ctx, RB_S_SZ, (size_t)n, &sz));
    RB_TRY(scratch(ctx, RB_S_ST, (size_t)n, &st));
    RB_TRY(scratch(ctx, RB_S_CORE, (size_t)n, &core));
    RB_TRY(scratch(ctx, RB_S_PARENT, (size_t)n, &parent));
    RB_TRY(scratch(ctx, RB_S_MINORIG, (size_t)n, &minorig));
    RB_TRY(scratch(ctx, RB_S_FLAGS, (size_t)n, &flags));
    RB_TRY(scratch(ctx, RB_S_RANK, (size_t)n, &rank));
    RB_TRY(scratch(ctx, RB_S_SLABEL, (size_t)n, &slabel));

    RB_CUDA(cudaMemsetAsync(cell_start, 0, sizeof(int) * ((size_t)n_cells + 1), stream));
    db_cell_kernel<<<blocks, DB_THREADS, 0, stream>>>(pts, g, n64, cell_id, slot, cell_start);
    RB_LAUNCH_CHECK(ctx);
    RB_TRY(rb_exclusive_scan_i32(ctx, cell_start, cell_start, n_cells + 1, nullptr, stream));
    db_scatter_kernel<<<blocks, DB_THREADS, 0, stream>>>(pts, DIM, n64, cell_id, slot, cell_start, sidx, scell, sx, sy, sz, st);
    RB_LAUNCH_CHECK(ctx);

    // 4. neighbour count -> core flags
    Sorted s{sx, sy, sz, st, scell, cell_start};
    const double eps2 = eps_space * eps_space;
    db_count_kernel<DIM><<<blocks, DB_THREADS, 0, stream>>>(s, g, n, eps2, eps_time, min_samples, core, parent, d_ctr + 0);
    RB_LAUNCH_CHECK(ctx);

    // 5. union-find over core-core edges, canonical numbering
    db_union_kernel<DIM><<<blocks, DB_THREADS, 0, stream>>>(s, g, n, eps2, eps_time, core, parent, d_ctr + 1);
    RB_LAUNCH_CHECK(ctx);
    RB_CUDA(cudaMemsetAsync(minorig, 0x7f, sizeof(int) * (size_t)n, stream));
    RB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)n, stream));
    db_minorig_kernel<<<blocks, DB_THREADS, 0, stream>>>(n, core, parent, sidx, minorig);
    RB_LAUNCH_CHECK(ctx);
    db_rootflag_kernel<<<blocks, DB_THREADS, 0, stream>>>(n, core, parent, minorig, flags);
    RB_LAUNCH_CHECK(ctx);
    int* d_total = (int*)(d_ctr + 3);
    RB_TRY(rb_exclusive_scan_i32(ctx, flags, rank, n, d_total, stream));
    db_label_kernel<<<blocks, DB_THREADS, 0, stream>>>(n, core, parent, minorig, rank, sidx, slabel, labels, core_out);
    RB_LAUNCH_CHECK(ctx);

    // 6. border points
    db_border_kernel<DIM><<<blocks, DB_THREADS, 0, stream>>>(s, g, n, eps2, eps_time, core, slabel, sidx, labels, d_ctr + 2);
    RB_LAUNCH_CHECK(ctx);

    // 7. stats + cluster count to the host
    unsigned long long* hc = (unsigned long long*)ctx->pinned;
    RB_CUDA(cudaMemcpyAsync(hc, d_ctr, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaStreamSynchronize(stream));
    rb_dbscan_stats& stt = ctx->last_stats;
    memset(&stt, 0, sizeof stt);
    stt.n_points = n;
    stt.n_cells = n_cells;
    stt.n_clusters = (int64_t)(int)(hc[3] & 0xffffffffu);
    stt.n_core = -1;
    stt.pair_tests_count = (int64_t)hc[0];
    stt.pair_tests_union = (int64_t)hc[1];
    stt.pair_tests_border = (int64_t)hc[2];
    stt.cell_size = cell;
    stt.time_bin = wt;
    stt.dims[0] = g.n[0]; stt.dims[1] = g.n[1]; stt.dims[2] = g.n[2]; stt.dims[3] = g.nt;
    stt.time_radius = g.tr;
    if (n_clusters) *n_clusters = stt.n_clusters;
    return RB_OK;
}

}  // namespace

extern "C" int rb_stdbscan(rb_ctx* ctx, const float* x, const float* y, const float* z, int64_t stride,
                           const float* times, int64_t n, double eps_space, float eps_time, int min_samples,
                           int32_t* labels, uint8_t* core, int64_t* n_clusters, void* stream_) {
    RB_REQUIRE(ctx, "ctx is NULL");
    RB_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) - 1, "point count out of range");
    if (n_clusters) *n_clusters = 0;
    if (n == 0) return RB_OK;
    RB_REQUIRE(x && times && labels, "NULL argument");
    RB_REQUIRE(stride >= 1, "stride must be >= 1");
    RB_REQUIRE(!(z && !y), "z without y");
    RB_REQUIRE(eps_space >= 0 && isfinite(eps_space), "eps_space must be finite and >= 0");
    cudaStream_t stream = (cudaStream_t)stream_;
    DbPoints pts{x, y, z, stride, times};
    if (z) return run_dbscan<3>(ctx, pts, n, eps_space, eps_time, min_samples, labels, core, n_clusters, stream);
    if (y) return run_dbscan<2>(ctx, pts, n, eps_space, eps_time, min_samples, labels, core, n_clusters, stream);
    return run_dbscan<1>(ctx, pts, n, eps_space, eps_time, min_samples, labels, core, n_clusters, stream);
}

extern "C" int rb_stdbscan_last_stats(rb_ctx* ctx, rb_dbscan_stats* out) {
    RB_REQUIRE(ctx && out, "NULL argument");
    *out = ctx->last_stats;
    return RB_OK;
}
