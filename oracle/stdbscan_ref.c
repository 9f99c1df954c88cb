/*
 * CPU oracle (plain C) for ST-DBSCAN labels — TEST INFRASTRUCTURE ONLY.
 *
 * Restates the RESULT of the reference's sequential ST-DBSCAN
 * (PointCloudWork/4_temporal_object_tracker.py:469-506, twins
 * 3_stdbscan_point_clouds.py:101-136 and radar_pipeline/processors/clustering.py:49-115)
 * in its order-free form (note N4):
 *   neighbour(p,q)  <=>  sum_d (double(p_d)-double(q_d))^2 <= eps*eps   (scikit-learn BallTree,
 *                        float64, inclusive, self included; T4:474-475)
 *                    and |t_p - t_q| <= eps_time in float32               (T4:485-486)
 *   core(p)         <=>  |N(p)| >= min_samples                            (T4:488)
 *   clusters         =   connected components of core points; id = rank of the component's
 *                        smallest core index (discovery order of the loop at T4:479,506)
 *   border(p)        =   smallest cluster id among p's core neighbours    (T4:503-504)
 * It exists so that parity cases too large for the Python oracle (its neighbour lists and
 * per-element loops) still have an independent CPU answer. tests/test_oracle_golden.py pins it
 * against labels produced by the unmodified reference (tests/golden/).
 *
 * Candidate search: a uniform grid with cells at least eps wide in space and eps_time wide in time (a neighbour is
 * never more than one cell away on any axis); every candidate is decided by the exact predicates above. The count and
 * border passes are independent per point and run under OpenMP when available (full-size parity blocks of millions of
 * points finish in seconds); so does the union pass, over a lock-free min-root union-find, skipping pairs that already
 * share a root.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/build.py). No fast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int64_t key; int64_t idx; } keyed_t;

static int cmp_keyed(const void* a, const void* b) {
    const keyed_t* x = (const keyed_t*)a; const keyed_t* y = (const keyed_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

typedef struct {
    const float* xyz; const float* t; int dim; int64_t n;
    double eps2; float eps_t;
    double lo[3]; double cell; int64_t nc[3];
    double tlo; double tcell; int64_t nt;     /* time bins at least eps_time wide: neighbours lie within +-1 bin */
    keyed_t* sorted;      /* points sorted by cell key */
    float* sxyz; float* st; /* coordinates and times in sorted order (contiguous candidate scans) */
    int64_t* ukeys; int64_t* ustart; int64_t nu;
} grid_t;

static int64_t find_cell(const grid_t* g, int64_t key) {
    int64_t lo = 0, hi = g->nu - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) >> 1;
        if (g->ukeys[mid] == key) return mid;
        if (g->ukeys[mid] < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* p = original index of the query point, s = SORTED position of the candidate */
static inline int is_neighbour(const grid_t* g, int64_t p, int64_t s) {
    float dt = g->st[s] - g->t[p];
    if (fabsf(dt) > g->eps_t) return 0;
    if (!(fabsf(dt) <= g->eps_t)) return 0;          /* NaN times never match */
    double d = 0.0;
    for (int k = 0; k < g->dim; ++k) {
        double tmp = (double)g->xyz[p * g->dim + k] - (double)g->sxyz[s * g->dim + k];
        d += tmp * tmp;
    }
    return d <= g->eps2;
}

typedef int (*visit_fn)(int64_t p, int64_t q, void* ctx);      /* nonzero = stop: the answer is already certain */
typedef int (*pre_fn)(int64_t p, int64_t q, void* ctx);       /* optional: 0 = skip this candidate untested */

static inline int64_t time_bin(const grid_t* g, int64_t p) {
    double t = (double)g->t[p];
    if (!(t == t)) return 0;                         /* NaN times never match anything anyway */
    return (int64_t)floor((t - g->tlo) / g->tcell);
}

static void for_each_neighbour_pre(const grid_t* g, int64_t p, pre_fn pre, visit_fn fn, void* ctx) {
    int64_t c[3] = {0, 0, 0};
    for (int k = 0; k < g->dim; ++k)
        c[k] = (int64_t)floor(((double)g->xyz[p * g->dim + k] - g->lo[k]) / g->cell);
    const int64_t ct = time_bin(g, p);
    int64_t z0 = g->dim > 2 ? c[2] - 1 : 0, z1 = g->dim > 2 ? c[2] + 1 : 0;
    int64_t y0 = g->dim > 1 ? c[1] - 1 : 0, y1 = g->dim > 1 ? c[1] + 1 : 0;
    for (int64_t cx = c[0] - 1; cx <= c[0] + 1; ++cx) {
        if (cx < 0 || cx >= g->nc[0]) continue;
        for (int64_t cy = y0; cy <= y1; ++cy) {
            if (cy < 0 || cy >= g->nc[1]) continue;
            for (int64_t cz = z0; cz <= z1; ++cz) {
                if (cz < 0 || cz >= g->nc[2]) continue;
                for (int64_t tb = ct - 1; tb <= ct + 1; ++tb) {
                    if (tb < 0 || tb >= g->nt) continue;
                    int64_t u = find_cell(g, ((cx * g->nc[1] + cy) * g->nc[2] + cz) * g->nt + tb);
                    if (u < 0) continue;
                    for (int64_t s = g->ustart[u]; s < g->ustart[u + 1]; ++s) {
                        int64_t q = g->sorted[s].idx;
                        if (pre && !pre(p, q, ctx)) continue;
                        if (is_neighbour(g, p, s) && fn(p, q, ctx)) return;
                    }
                }
            }
        }
    }
}

static void for_each_neighbour(const grid_t* g, int64_t p, visit_fn fn, void* ctx) {
    for_each_neighbour_pre(g, p, 0, fn, ctx);
}

typedef struct { int64_t* count; int64_t enough; } count_ctx;     /* core is certain once min_samples neighbours were seen */
static int visit_count(int64_t p, int64_t q, void* v) { (void)q; count_ctx* c = (count_ctx*)v; return ++c->count[p] >= c->enough; }

typedef struct { int64_t* parent; const unsigned char* core; } union_ctx;
static int64_t uf_find(int64_t* parent, int64_t a) {
    while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
    return a;
}
/* Lock-free union-find for the parallel union pass (the design of the reference's native tier,
 * radar-pipeline-rs/src/processors/clustering.rs:33-108, with the link direction that makes the result order-free):
 * a root is only ever hooked under a SMALLER index, so the structure stays acyclic under races and the final root
 * of every set is its smallest member - which thread links first changes nothing. */
static int64_t uf_find_mt(int64_t* parent, int64_t a) {
    for (;;) {
        int64_t p = __atomic_load_n(parent + a, __ATOMIC_RELAXED);
        if (p == a) return a;
        int64_t gp = __atomic_load_n(parent + p, __ATOMIC_RELAXED);
        if (gp != p) __atomic_store_n(parent + a, gp, __ATOMIC_RELAXED);      /* path halving: any ancestor is a valid parent */
        a = p;
    }
}
static int pre_union(int64_t p, int64_t q, void* v) {          /* core-core edges once (q < p), not yet in one set */
    union_ctx* u = (union_ctx*)v;
    if (q >= p || !u->core[q]) return 0;
    return uf_find_mt(u->parent, p) != uf_find_mt(u->parent, q);
}
static int visit_union(int64_t p, int64_t q, void* v) {
    union_ctx* u = (union_ctx*)v;
    if (!u->core[q]) return 0;
    for (;;) {
        int64_t a = uf_find_mt(u->parent, p), b = uf_find_mt(u->parent, q);
        if (a == b) return 0;
        if (a < b) { int64_t t = a; a = b; b = t; }                 /* hook the larger root under the smaller */
        int64_t expect = a;
        if (__atomic_compare_exchange_n(u->parent + a, &expect, b, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return 0;
    }
}

typedef struct { const int* labels; const unsigned char* core; int best; } border_ctx;
static int visit_border(int64_t p, int64_t q, void* v) {
    (void)p;
    border_ctx* b = (border_ctx*)v;
    if (b->core[q] && (b->best < 0 || b->labels[q] < b->best)) b->best = b->labels[q];
    return 0;
}

/* the "paper" variant (PointCloudWorkF/stdbscan_denoising_pipeline.py:264-369): distinct int32(times) among the
 * neighbours, and the border rule of its FIFO expansion (see numpy_oracle.st_dbscan_wf_canonical) */
typedef struct { const float* t; int32_t* vals; int n; int64_t count; int64_t enough; int enough_frames; } frames_ctx;
static int visit_frames(int64_t p, int64_t q, void* v) {
    (void)p;
    frames_ctx* f = (frames_ctx*)v;
    f->count++;
    int32_t ti = (int32_t)f->t[q];                       /* numpy astype(int32): truncation toward zero */
    int seen = 0;
    for (int i = 0; i < f->n; ++i) if (f->vals[i] == ti) { seen = 1; break; }
    if (!seen && f->n < 4096) f->vals[f->n++] = ti;
    return f->count >= f->enough && f->n >= f->enough_frames;
}
typedef struct { const int* labels; const unsigned char* core; int64_t* parent; int64_t self; int best; } border_wf_ctx;
static int visit_border_wf(int64_t p, int64_t q, void* v) {
    (void)p;
    border_wf_ctx* b = (border_wf_ctx*)v;
    if (!b->core[q]) return 0;
    int64_t start = b->parent[q];                        /* parents are flattened; min-root union: the root is the cluster's start point */
    if (!(start < b->self || q == start)) return 0;
    if (b->best < 0 || b->labels[q] < b->best) b->best = b->labels[q];
    return 0;
}

static int64_t oracle_stdbscan_impl(const float* xyz, int dim, const float* times, int64_t n,
                                    double eps_space, float eps_time, int min_samples, int min_frames, int wf_border,
                                    int* labels, unsigned char* core_out);

/* Returns number of clusters, or -1 on allocation failure / bad arguments. */
int64_t oracle_stdbscan(const float* xyz, int dim, const float* times, int64_t n,
                        double eps_space, float eps_time, int min_samples,
                        int* labels, unsigned char* core_out) {
    return oracle_stdbscan_impl(xyz, dim, times, n, eps_space, eps_time, min_samples, 0, 0, labels, core_out);
}

int64_t oracle_stdbscan_wf(const float* xyz, int dim, const float* times, int64_t n,
                           double eps_space, float eps_time, int min_samples, int min_frames,
                           int* labels, unsigned char* core_out) {
    return oracle_stdbscan_impl(xyz, dim, times, n, eps_space, eps_time, min_samples, min_frames, 1, labels, core_out);
}

static int64_t oracle_stdbscan_impl(const float* xyz, int dim, const float* times, int64_t n,
                                    double eps_space, float eps_time, int min_samples, int min_frames, int wf_border,
                                    int* labels, unsigned char* core_out) {
    if (dim < 1 || dim > 3 || n < 0) return -1;
    for (int64_t i = 0; i < n; ++i) labels[i] = -1;
    if (core_out) memset(core_out, 0, (size_t)n);
    if (n == 0) return 0;

    grid_t g; memset(&g, 0, sizeof g);
    g.xyz = xyz; g.t = times; g.dim = dim; g.n = n;
    g.eps2 = eps_space * eps_space; g.eps_t = eps_time;
    double hi[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) { g.lo[k] = 0; g.nc[k] = 1; }
    for (int k = 0; k < dim; ++k) {
        g.lo[k] = hi[k] = xyz[k];
        for (int64_t i = 1; i < n; ++i) {
            double v = xyz[i * dim + k];
            if (v < g.lo[k]) g.lo[k] = v;
            if (v > hi[k]) hi[k] = v;
        }
    }
    /* time bins: |fl32(t_q - t_p)| <= eps_t implies |t_q - t_p| < eps_t * (1 + 1e-6), so a neighbour is never more than
     * one bin away; with eps_t <= 0 only equal times match and any positive width will do */
    double thi = 0.0; int have_t = 0;
    g.tlo = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double v = (double)times[i];
        if (!(v == v)) continue;
        if (!have_t || v < g.tlo) g.tlo = v;
        if (!have_t || v > thi) thi = v;
        have_t = 1;
    }
    g.tcell = eps_time > 0 ? (double)eps_time * (1.0 + 1e-6) : 1.0;
    g.cell = eps_space > 0 ? eps_space * (1.0 + 1e-9) : 1.0;
    for (;;) {                                   /* coarsen until the key fits comfortably into 62 bits */
        double prod = 1.0;
        for (int k = 0; k < dim; ++k) {
            g.nc[k] = (int64_t)floor((hi[k] - g.lo[k]) / g.cell) + 1;
            prod *= (double)g.nc[k];
        }
        g.nt = (int64_t)floor((thi - g.tlo) / g.tcell) + 1;
        prod *= (double)g.nt;
        if (prod < 4.0e18) break;
        if ((double)g.nt > 1048576.0) g.tcell *= 2.0; else g.cell *= 2.0;
    }

    g.sorted = (keyed_t*)malloc(sizeof(keyed_t) * (size_t)n);
    g.ukeys = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    g.ustart = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t* count = (int64_t*)calloc((size_t)n, sizeof(int64_t));
    int64_t* parent = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    unsigned char* core = (unsigned char*)calloc((size_t)n, 1);
    if (!g.sorted || !g.ukeys || !g.ustart || !count || !parent || !core) return -1;

    for (int64_t i = 0; i < n; ++i) {
        int64_t c[3] = {0, 0, 0};
        for (int k = 0; k < dim; ++k)
            c[k] = (int64_t)floor(((double)xyz[i * dim + k] - g.lo[k]) / g.cell);
        g.sorted[i].key = ((c[0] * g.nc[1] + c[1]) * g.nc[2] + c[2]) * g.nt + time_bin(&g, i);
        g.sorted[i].idx = i;
    }
    qsort(g.sorted, (size_t)n, sizeof(keyed_t), cmp_keyed);
    g.nu = 0;
    for (int64_t s = 0; s < n; ++s) {
        if (s == 0 || g.sorted[s].key != g.sorted[s - 1].key) {
            g.ukeys[g.nu] = g.sorted[s].key; g.ustart[g.nu] = s; g.nu++;
        }
    }
    g.ustart[g.nu] = n;
    g.sxyz = (float*)malloc(sizeof(float) * (size_t)n * (size_t)dim);
    g.st = (float*)malloc(sizeof(float) * (size_t)n);
    if (!g.sxyz || !g.st) return -1;
    for (int64_t s = 0; s < n; ++s) {
        const int64_t i = g.sorted[s].idx;
        for (int k = 0; k < dim; ++k) g.sxyz[s * dim + k] = xyz[i * dim + k];
        g.st[s] = times[i];
    }

    if (wf_border) {
        int failed = 0;
#pragma omp parallel
        {
            int32_t* vals = (int32_t*)malloc(sizeof(int32_t) * 4096);
            if (!vals) {
#pragma omp atomic write
                failed = 1;
            }
#pragma omp barrier
            if (!failed) {
#pragma omp for schedule(dynamic, 256)
                for (int64_t p = 0; p < n; ++p) {
                    frames_ctx fc = {times, vals, 0, 0, min_samples, min_frames};
                    for_each_neighbour(&g, p, visit_frames, &fc);
                    core[p] = fc.count >= min_samples && fc.n >= min_frames;
                    parent[p] = p;
                }
            }
            free(vals);
        }
        if (failed) return -1;
    } else {
        count_ctx cc = {count, min_samples > 1 ? min_samples : 1};
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t p = 0; p < n; ++p) for_each_neighbour(&g, p, visit_count, &cc);      /* writes count[p] only */
        for (int64_t p = 0; p < n; ++p) { core[p] = count[p] >= min_samples; parent[p] = p; }
    }

    union_ctx uc = {parent, core};
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < n; ++p) if (core[p]) for_each_neighbour_pre(&g, p, pre_union, visit_union, &uc);

    int64_t n_clusters = 0;
    int* root_id = (int*)malloc(sizeof(int) * (size_t)n);
    if (!root_id) return -1;
    for (int64_t p = 0; p < n; ++p)
        root_id[p] = (core[p] && uf_find(parent, p) == p) ? (int)n_clusters++ : -1;
    for (int64_t p = 0; p < n; ++p) if (core[p]) labels[p] = root_id[uf_find(parent, p)];
    for (int64_t p = 0; p < n; ++p) parent[p] = uf_find(parent, p);   /* flatten: the border pass only reads parents */
    /* border points write their own label and read labels of core points only: independent per point */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t p = 0; p < n; ++p) {
        if (core[p]) continue;
        if (wf_border) {
            border_wf_ctx bw = {labels, core, parent, p, -1};
            for_each_neighbour(&g, p, visit_border_wf, &bw);
            labels[p] = bw.best;
        } else {
            border_ctx bc = {labels, core, -1};
            for_each_neighbour(&g, p, visit_border, &bc);
            labels[p] = bc.best;
        }
    }
    if (core_out) memcpy(core_out, core, (size_t)n);

    free(root_id); free(core); free(parent); free(count);
    free(g.ustart); free(g.ukeys); free(g.sorted); free(g.sxyz); free(g.st);
    return n_clusters;
}
