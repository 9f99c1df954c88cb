// Spoke-to-point: threshold + ordered stream compaction + stride + polar->Cartesian + gain concat.
//
// Replaces load_radar_csv's numeric part (reference 4_temporal_object_tracker.py:200-232) and the
// per-gain concatenation of build_frame (:322-344) for a batch of W = frames x gains sweeps.
//
// Three launches, no inter-CTA waiting anywhere (a single-pass decoupled look-back was built and
// measured first: its prefix latency of ~9 us per 64 KiB tile capped it at 1.4 TB/s, see
// profiles/r01_spoke_v2_tma_lookback_trace.txt):
//
//   1. spoke_mask_tma_kernel streams echo[W][S][E] ONCE at HBM speed: one persistent CTA per SM, a producer warp
//                           fills a ring of 64 KiB shared-memory stages with 1-D bulk async copies (TMA), consumer
//                           warps turn them into 1 bit per cell (survivor mask, cell order) and one survivor
//                           count per 4096-cell tile. The HBM-bound kernel: 4 B (float32) or 1 B (uint8 echoes)
//                           read + 1/8 B written per cell. spoke_mask_kernel is the register-staged fallback for
//                           sweeps that are not a multiple of 16 bytes.
//   2. spoke_offsets_kernel in-sweep exclusive prefix of the tile counts (one block per sweep) and,
//                           by the last block to finish, the output base of every sweep:
//                           base[w] = sum_{w'<w} ceil(M_w' / stride)   (x[mask][::stride], T4:222-230)
//   3. spoke_emit_kernel    one warp per group of 4 tiles: popcount-scan of the mask words gives the
//                           in-sweep rank at the start of every 128-cell entry; the KEPT ranks
//                           (multiples of the stride) of the group are dealt to the lanes round-robin,
//                           each lane finds its survivor by a binary search over the entry ranks and
//                           a select-in-word, gathers the echo value, and writes x, y, intensity and
//                           gain. Consecutive lanes write consecutive output slots (coalesced), and
//                           the work is proportional to the points kept, not to the cells.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_TILE = 4096;                       // cells per tile (16 KiB of echo), one warp per tile
constexpr int SK_WORDS = SK_TILE / 32;              // 128 mask words per tile
constexpr int SK_BATCH = 8;                         // 128-bit loads in flight per lane
constexpr int SK_BATCH_CELLS = SK_BATCH * 128;      // 1024 cells per batch and warp
constexpr int SK_BATCHES = SK_TILE / SK_BATCH_CELLS;
constexpr int SK_GROUP = 4;                         // tiles per warp in the emit kernel
constexpr int SK_ENTRIES = SK_GROUP * 32;           // 128-cell entries per group
constexpr int SK_CTAS_PER_SM = 4;

// unsigned division by a launch-invariant divisor (magic multiply), exact for all 32-bit n
struct FastDiv { uint32_t mul, shift, d; };
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f; f.d = d; f.mul = 0; f.shift = 0;
    if (d <= 1) return f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                                   // ceil(log2 d)
    f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
    f.shift = l;
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
    if (f.d <= 1) return n;
    uint32_t t = __umulhi(f.mul, n);
    return (t + ((n - t) >> 1)) >> (f.shift - 1);
}

struct SpokeGeom {
    int64_t total_tiles;
    int sweep_cells;                     // S*E
    int tiles_per_sweep;
    int n_bins;
    int n_spokes;
};

// ---- 1. mask + count ------------------------------------------------------------------------------
// Lane l of a warp owns cells 4l..4l+3 of each 128-cell chunk (one 128-bit load). The 4-bit survivor
// nibbles of 8 chunks are OR-reduced across each group of 8 lanes with a transposing butterfly
// (7 shuffles for 8 chunks): afterwards lane l holds the finished 32-bit word of chunk (l & 7),
// lane group (l >> 3), i.e. word (l & 7) * 4 + (l >> 3) of the batch - one 128-byte store per batch.
__device__ __forceinline__ unsigned set_gt(float a, float b) {            // 0xffffffff if a > b (false for NaN)
    unsigned r;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}

struct LaneBits { unsigned b0, b1, b2, b3; };                             // (1 << i) << 4 * (lane & 7)
__device__ __forceinline__ LaneBits lane_bits(unsigned lane) {
    const unsigned sh = 4u * (lane & 7u);
    return LaneBits{1u << sh, 2u << sh, 4u << sh, 8u << sh};
}
__device__ __forceinline__ unsigned nibble(const float4& v, float thr, const LaneBits& lb) {
    return (set_gt(v.x, thr) & lb.b0) | (set_gt(v.y, thr) & lb.b1) | (set_gt(v.z, thr) & lb.b2) | (set_gt(v.w, thr) & lb.b3);
}

// Echo element types: float32 (the reference's in-memory type, T4:206) and uint8 (what the radar delivers,
// 0..255, PIPELINE_DOCUMENTATION.txt:47 - a quarter of the bytes to move). For uint8 the float comparison
// (float)e > thr is the integer comparison e >= lb with lb = 0 for thr < 0, floor(thr) + 1 otherwise (lb > 255 or a
// NaN threshold: nothing survives), evaluated four cells at a time with the SIMD byte compare.
struct ThrArg {
    float thr;              // float32 path
    unsigned lb4;           // uint8 path: lower bound replicated into the four bytes
    int none;               // uint8 path: nothing can survive
};
inline ThrArg make_thr(float thr) {
    ThrArg a; a.thr = thr; a.lb4 = 0; a.none = 0;
    if (!(thr == thr)) { a.none = 1; return a; }                      // NaN: every comparison is false
    if (thr < 0.f) return a;                                          // every byte >= 0 passes
    const double lb = floor((double)thr) + 1.0;
    if (lb > 255.0) { a.none = 1; return a; }
    a.lb4 = (unsigned)lb * 0x01010101u;
    return a;
}
template <typename T> struct Elem;
template <> struct Elem<float>   { static constexpr int ALIGN_CELLS = 4;  static constexpr int WARPS = 8; };    // 16-byte bulk copies;
template <> struct Elem<uint8_t> { static constexpr int ALIGN_CELLS = 16; static constexpr int WARPS = 16; };   // 4x the cells per stage: twice the consumer warps

__device__ __forceinline__ unsigned u8_bits4(unsigned word, const ThrArg& t) {              // survivor flags of 4 cells in bits 0..3
    const unsigned r = t.none ? 0u : __vcmpgeu4(word, t.lb4);
    return ((r & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ unsigned nibble_u8(unsigned word, const ThrArg& t, unsigned sh) {
    const unsigned r = t.none ? 0u : __vcmpgeu4(word, t.lb4);          // 0xff in every byte that survives
    return (((r & 0x01010101u) * 0x01020408u) >> 24) << sh;           // gather the four flags into bits 0..3
}
// survivor nibble of lane's 4 cells of chunk k of a 1024-cell batch held in shared memory
template <typename T>
__device__ __forceinline__ unsigned batch_nibble(const unsigned char* __restrict__ batch, int k, unsigned lane, const ThrArg& t,
                                                 const LaneBits& lb) {
    if (sizeof(T) == 4) return nibble(reinterpret_cast<const float4*>(batch)[k * 32 + lane], t.thr, lb);
    return nibble_u8(reinterpret_cast<const unsigned*>(batch)[k * 32 + lane], t, 4u * (lane & 7u));
}

// transposing OR-butterfly over the 8 lanes of a group: 8 -> 4 -> 2 -> 1 registers
__device__ __forceinline__ unsigned butterfly8(const unsigned (&nib)[8], unsigned lane) {
    unsigned a4[4], a2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool hi = lane & 1u;
        unsigned keep = hi ? nib[2 * i + 1] : nib[2 * i];
        unsigned give = hi ? nib[2 * i] : nib[2 * i + 1];
        a4[i] = keep | __shfl_xor_sync(0xffffffffu, give, 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool hi = lane & 2u;
        unsigned keep = hi ? a4[2 * i + 1] : a4[2 * i];
        unsigned give = hi ? a4[2 * i] : a4[2 * i + 1];
        a2[i] = keep | __shfl_xor_sync(0xffffffffu, give, 2);
    }
    const bool hi = lane & 4u;
    unsigned keep = hi ? a2[1] : a2[0];
    unsigned give = hi ? a2[0] : a2[1];
    return keep | __shfl_xor_sync(0xffffffffu, give, 4);
}

template <typename T>
struct TileRef { const T* src; int valid; };                              // first cell + cells inside the sweep
template <typename T>
__device__ __forceinline__ TileRef<T> tile_ref(const T* echo, const SpokeGeom& g, long long tile) {
    const int w = (int)(tile / g.tiles_per_sweep);
    const int cell0 = (int)(tile - (long long)w * g.tiles_per_sweep) * SK_TILE;
    return TileRef<T>{echo + (int64_t)w * g.sweep_cells + cell0, min(SK_TILE, g.sweep_cells - cell0)};
}

// ---- 1a. TMA-staged variant (the default) --------------------------------------------------------------
// One persistent CTA per SM: a producer thread streams stages of MT_TILES tiles (64 KiB) into a ring of
// MT_STAGES shared-memory buffers with 1-D bulk async copies (cp.async.bulk ... mbarrier::complete_tx, one
// per stage when its tiles are full), so the bytes in flight per SM (up to 192 KiB) are held by the TMA engine and shared memory,
// not by registers. MT_WARPS consumer warps turn each stage into mask words, one 1024-cell batch (= one
// 128-byte line of mask) per warp at a time, and add their survivor counts to per-tile shared counters
// that the producer flushes to global memory when it recycles the slot. Stages are handed out by one
// atomic ticket per CTA and stage (a ticket per WARP and tile serialises on the atomic unit: measured
// 4.8 instead of 7.0 TB/s, tools/mask_bench.cu).
// The ring is a template parameter pair (stage bytes x stages), chosen with rb_set_option "spoke_ring":
//   2 (default) 32 KiB x 3 =  98 KB: 6.6 TB/s alone, and small enough for the CTAs of OTHER kernels to be resident on the SM
//                                    beside it - the latency-bound clustering kernels of the block before run under this
//                                    HBM-bound one (measured: rings up to ~100 KB co-reside, 131 KB and more do not)
//   0           64 KiB x 3 = 197 KB: the most bytes in flight, 7.0 TB/s alone, owns the SM (round 1's kernel)
//   1           32 KiB x 4 = 131 KB, 3: 48 KiB x 2 = 98 KB: the sweep's other useful points (profiles/r02_ring_sweep.txt)
template <typename T> constexpr int mt_threads() { return Elem<T>::WARPS * 32 + 32; }    // consumer warps + the producer warp

template <int TILES>
struct MtMeta {
    long long first_tile;                 // global id of the stage's first tile; < 0: no more work
    int n_tiles;
    int valid[TILES];                     // cells of each tile inside its sweep
    unsigned cnt[TILES];                  // survivors, accumulated by the consumers
};
template <int TILES, int STAGE_BYTES, int STAGES>
struct __align__(128) MtSmem {
    unsigned char ring[STAGES][STAGE_BYTES];
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
    MtMeta<TILES> meta[STAGES];
};
template <typename T, int STAGE_BYTES> __host__ __device__ constexpr int mt_tiles() { return STAGE_BYTES / (SK_TILE * (int)sizeof(T)); }
template <int STAGE_BYTES> __host__ __device__ constexpr int mt_chunk_stages() {      // stages per work chunk (~256 KiB)
    return (256 * 1024) / STAGE_BYTES > 0 ? (256 * 1024) / STAGE_BYTES : 1;
}
template <typename T, int STAGE_BYTES, int STAGES> constexpr int mt_smem_bytes() {
    return (int)sizeof(MtSmem<mt_tiles<T, STAGE_BYTES>(), STAGE_BYTES, STAGES>) + 128;
}

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
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni MT_DONE;\n"
        "bra.uni MT_WAIT;\n"
        "MT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk async copy (TMA engine, 1-D), completion counted in bytes on an mbarrier. The echo stream is
// read exactly once: with `policy` != 0 (an L2 evict-first policy) its lines are the first to leave L2, so the working
// set of kernels running beside this one (bucket tables, sorted points of the previous block) stays resident.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar, unsigned long long policy) {
    if (policy)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <typename T, int MT_STAGE_BYTES, int MT_STAGES>
__global__ void __launch_bounds__(mt_threads<T>(), 1)
spoke_mask_tma_kernel(const T* __restrict__ echo, const SpokeGeom g, const ThrArg threshold,
                      uint32_t* __restrict__ mask, uint32_t* __restrict__ tile_count, unsigned* __restrict__ ticket, const int l2_hint) {
    constexpr int MT_TILES = mt_tiles<T, MT_STAGE_BYTES>();
    constexpr int MT_WARPS = Elem<T>::WARPS;
    constexpr int MT_STAGE_BATCHES = MT_TILES * SK_BATCHES;
    constexpr int MT_CHUNK = mt_chunk_stages<MT_STAGE_BYTES>();
    static_assert(MT_TILES >= 1 && MT_TILES * SK_TILE * sizeof(T) == MT_STAGE_BYTES, "a stage is a whole number of tiles");
    using Smem = MtSmem<MT_TILES, MT_STAGE_BYTES, MT_STAGES>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const unsigned lane = rb_lane();
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < MT_STAGES; ++i) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], MT_WARPS);
#pragma unroll
            for (int t = 0; t < MT_TILES; ++t) sm.meta[i].cnt[t] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const long long n_stages = (g.total_tiles + MT_TILES - 1) / MT_TILES;

    if (warp == MT_WARPS) {
        // ================= producer warp: lane t looks after tile t of the stage =================
        static_assert(MT_TILES <= 32, "one lane per tile");
        auto flush = [&](int s) {                                          // counts of the stage that just left slot s
            MtMeta<MT_TILES>& m = sm.meta[s];
            if ((int)lane < m.n_tiles) {
                tile_count[m.first_tile + lane] = m.cnt[lane];
                m.cnt[lane] = 0;
            }
            __syncwarp();
        };
        const unsigned long long policy = l2_hint ? l2_evict_first_policy() : 0ull;
        // Work is handed out in CHUNKS of MT_CHUNK stages (256 KiB), the first one statically (chunk = blockIdx.x), the
        // following ones by an atomic ticket. A ticket per STAGE serialises on the counter's address at ~3.4 ns each:
        // with 32 KiB stages that alone is 2.7 ms per 25.8 GB block (measured: 5.5 instead of 7.0 TB/s); per chunk it
        // is 0.3 ms spread over the whole kernel, whatever the stage size.
        long long chunk = blockIdx.x;
        int it = 0;
        auto recycle = [&](int s) {                                        // slot s must be empty again before it is refilled
            if (it >= MT_STAGES) {
                mbar_wait(&sm.empty[s], (uint32_t)(it / MT_STAGES - 1) & 1u);
                flush(s);
            }
        };
        while (true) {
            // the ticket of the NEXT chunk goes out first: its round trip hides behind this chunk's work
            unsigned tk = 0;
            if (lane == 0) tk = atomicAdd(ticket, 1u);
            const long long st0 = chunk * MT_CHUNK;
            if (st0 >= n_stages) {                                         // end marker for the consumers
                const int s = it % MT_STAGES;
                recycle(s);
                if (lane == 0) {
                    sm.meta[s].first_tile = -1;
                    sm.meta[s].n_tiles = 0;
                    mbar_arrive(&sm.full[s]);
                }
                break;
            }
            const long long st1 = min(st0 + (long long)MT_CHUNK, n_stages);
            // lane t follows tile t of the stage through the chunk: sweep and tile-in-sweep by ONE division per chunk, then
            // increments (a 64-bit division per stage on the producer's critical path costs more than the copy's issue)
            long long my_tile = st0 * MT_TILES + lane;
            int my_w = (int)(my_tile / g.tiles_per_sweep);
            int my_tis = (int)(my_tile - (long long)my_w * g.tiles_per_sweep);
            for (long long st = st0; st < st1; ++st) {
                const int s = it % MT_STAGES;
                recycle(s);
                MtMeta<MT_TILES>& m = sm.meta[s];
                const long long t0 = st * MT_TILES;
                const int nt = (int)min((long long)MT_TILES, g.total_tiles - t0);
                int v = 0;
                const T* src = nullptr;
                if ((int)lane < nt) {
                    const int cell0 = my_tis * SK_TILE;
                    v = min(SK_TILE, g.sweep_cells - cell0);
                    src = echo + (int64_t)my_w * g.sweep_cells + cell0;
                }
                my_tile += MT_TILES;
                my_tis += MT_TILES;
                while (my_tis >= g.tiles_per_sweep) { my_tis -= g.tiles_per_sweep; ++my_w; }
                if ((int)lane < MT_TILES) m.valid[lane] = v;
                if (lane == 0) { m.first_tile = t0; m.n_tiles = nt; }
                const uint32_t bytes = __reduce_add_sync(0xffffffffu, (uint32_t)v * (uint32_t)sizeof(T));
                // full tiles are contiguous in memory (also across a sweep boundary) and in the ring
                const bool all_full = __all_sync(0xffffffffu, (int)lane >= nt || v == SK_TILE);
                if (lane == 0) mbar_expect_tx(&sm.full[s], bytes);         // releases the meta data to the consumers
                __syncwarp();
                if (all_full) {                                            // the common case: ONE bulk copy per stage
                    if (lane == 0) bulk_g2s(&sm.ring[s][0], src, bytes, &sm.full[s], policy);
                } else if (v > 0) {                                        // ragged stage: every lane copies its own tile
                    bulk_g2s(&sm.ring[s][(size_t)lane * SK_TILE * sizeof(T)], src, (uint32_t)v * (uint32_t)sizeof(T), &sm.full[s], policy);
                }
                ++it;
            }
            chunk = (long long)gridDim.x + (long long)__shfl_sync(0xffffffffu, tk, 0);
        }
        // stages still in flight: fills it-1 .. it-(MT_STAGES-1)
        for (int j = max(0, it - (MT_STAGES - 1)); j < it; ++j) {
            const int s = j % MT_STAGES;
            mbar_wait(&sm.empty[s], (uint32_t)(j / MT_STAGES) & 1u);
            flush(s);
        }
        return;
    }

    // ================= consumers =================
    const LaneBits lb = lane_bits(lane);
    const unsigned word_slot = (lane & 7u) * 4u + (lane >> 3);
    for (int it = 0;; ++it) {
        const int s = it % MT_STAGES;
        mbar_wait(&sm.full[s], (uint32_t)(it / MT_STAGES) & 1u);
        MtMeta<MT_TILES>& m = sm.meta[s];
        const long long t0 = m.first_tile;
        if (t0 < 0) break;
        const int nt = m.n_tiles;
#pragma unroll
        for (int b = warp; b < MT_STAGE_BATCHES; b += MT_WARPS) {
            const int t = b / SK_BATCHES;                                  // tile of the batch inside the stage
            if (t >= nt) break;
            const int valid = m.valid[t] - (b % SK_BATCHES) * SK_BATCH_CELLS;   // cells of this batch inside the sweep
            const unsigned char* __restrict__ batch = sm.ring[s] + (size_t)b * SK_BATCH_CELLS * sizeof(T);
            uint32_t* __restrict__ mw = mask + (t0 + t) * SK_WORDS + (b % SK_BATCHES) * 32;
            unsigned word;
            if (sizeof(T) == 1 && valid >= SK_BATCH_CELLS) {
                // uint8, full batch (1 KiB): two 128-bit loads per lane - 16 cells of each 512-cell half - give 16 survivor
                // bits per half; lanes 2w and 2w+1 together hold mask word w of a half, so ONE shuffle per half finishes
                // the words: even lanes keep the first half's, odd lanes the second half's, one 128-byte store.
                const uint4* __restrict__ q = reinterpret_cast<const uint4*>(batch);
                const uint4 h0 = q[lane], h1 = q[32 + lane];
                unsigned b0 = u8_bits4(h0.x, threshold) | (u8_bits4(h0.y, threshold) << 4) | (u8_bits4(h0.z, threshold) << 8) |
                              (u8_bits4(h0.w, threshold) << 12);
                unsigned b1 = u8_bits4(h1.x, threshold) | (u8_bits4(h1.y, threshold) << 4) | (u8_bits4(h1.z, threshold) << 8) |
                              (u8_bits4(h1.w, threshold) << 12);
                const unsigned sh16 = (lane & 1u) * 16u;
                b0 <<= sh16; b1 <<= sh16;
                b0 |= __shfl_xor_sync(0xffffffffu, b0, 1);
                b1 |= __shfl_xor_sync(0xffffffffu, b1, 1);
                word = (lane & 1u) ? b1 : b0;
                mw[(lane & 1u) * 16u + (lane >> 1)] = word;
            } else {
                unsigned nib[SK_BATCH];
                if (valid >= SK_BATCH_CELLS) {
#pragma unroll
                    for (int k = 0; k < SK_BATCH; ++k) nib[k] = batch_nibble<T>(batch, k, lane, threshold, lb);
                } else {
#pragma unroll
                    for (int k = 0; k < SK_BATCH; ++k) {
                        const int c = k * 128 + (int)lane * 4;              // valid is a multiple of 4 here
                        nib[k] = c < valid ? batch_nibble<T>(batch, k, lane, threshold, lb) : 0u;
                    }
                }
                word = butterfly8(nib, lane);
                mw[word_slot] = word;
            }
            const unsigned c = __reduce_add_sync(0xffffffffu, __popc(word));
            if (lane == 0 && c) atomicAdd(&m.cnt[t], c);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
}

// ---- 1b. register-staged variant (any shape; 128-bit loads when VEC) -------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(SK_THREADS, SK_CTAS_PER_SM)
spoke_mask_kernel(const T* __restrict__ echo, const SpokeGeom g, const float threshold,
                  uint32_t* __restrict__ mask, uint32_t* __restrict__ tile_count) {
    static_assert(!VEC || sizeof(T) == 4, "128-bit path is float32 only");
    const unsigned lane = rb_lane();
    const LaneBits lb = lane_bits(lane);
    const unsigned sh = 4u * (lane & 7u);
    const unsigned word_slot = (lane & 7u) * 4u + (lane >> 3);
    // static round-robin over the tiles: a per-warp ticket would serialise on the atomic unit (see above)
    const long long n_warps = (long long)gridDim.x * SK_WARPS;
    for (long long tile = (long long)blockIdx.x * SK_WARPS + (threadIdx.x >> 5); tile < g.total_tiles; tile += n_warps) {
        const TileRef<T> tr = tile_ref<T>(echo, g, tile);
        const int valid = tr.valid;
        const T* __restrict__ src = tr.src;
        uint32_t* __restrict__ mw = mask + tile * SK_WORDS;
        unsigned cnt = 0;
#pragma unroll 1
        for (int b = 0; b < SK_BATCHES; ++b) {
            const int c0 = b * SK_BATCH_CELLS + (int)lane * 4;
            unsigned nib[SK_BATCH];
            if (VEC) {
                const float* __restrict__ fsrc = reinterpret_cast<const float*>(src);
                if (valid == SK_TILE) {
                    float4 v[SK_BATCH];
#pragma unroll
                    for (int k = 0; k < SK_BATCH; ++k) v[k] = rb_ld_stream4(fsrc + c0 + k * 128);
#pragma unroll
                    for (int k = 0; k < SK_BATCH; ++k) nib[k] = nibble(v[k], threshold, lb);
                } else {
#pragma unroll
                    for (int k = 0; k < SK_BATCH; ++k) {
                        const int c = c0 + k * 128;
                        nib[k] = c < valid ? nibble(rb_ld_stream4(fsrc + c), threshold, lb) : 0u;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < SK_BATCH; ++k) {
                    const int c = c0 + k * 128;
                    unsigned m = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (c + i < valid) m |= (unsigned)((float)__ldg(src + c + i) > threshold) << i;
                    nib[k] = m << sh;
                }
            }
            const unsigned word = butterfly8(nib, lane);
            cnt += __popc(word);
            mw[b * 32 + word_slot] = word;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) tile_count[tile] = cnt;
    }
}

// ---- 2. offsets -------------------------------------------------------------------------------------
constexpr int SO_THREADS = 512;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp, T* total) {
    const unsigned lane = rb_lane(), warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    T before = 0, all = 0;
#pragma unroll
    for (int i = 0; i < SO_THREADS / 32; ++i) {
        T c = s_warp[i];
        if (i < (int)warp) before += c;
        all += c;
    }
    __syncthreads();
    *total = all;
    return before + incl - v;
}

__global__ void __launch_bounds__(SO_THREADS)
spoke_offsets_kernel(const uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_prefix, int tiles_per_sweep,
                     int64_t n_sweeps, int stride, uint32_t* __restrict__ sweep_total, int64_t* __restrict__ sweep_base,
                     unsigned* __restrict__ done, unsigned* __restrict__ ticket) {
    __shared__ uint32_t s_warp[SO_THREADS / 32];
    __shared__ unsigned long long s_warp64[SO_THREADS / 32];
    __shared__ bool s_last;
    const int64_t w = blockIdx.x;
    const uint32_t* __restrict__ cnt = tile_count + w * tiles_per_sweep;
    uint32_t* __restrict__ pfx = tile_prefix + w * tiles_per_sweep;
    uint32_t carry = 0;
    for (int t0 = 0; t0 < tiles_per_sweep; t0 += SO_THREADS) {
        const int t = t0 + (int)threadIdx.x;
        uint32_t v = t < tiles_per_sweep ? cnt[t] : 0u, total;
        uint32_t e = block_exclusive_scan<uint32_t>(v, s_warp, &total);
        if (t < tiles_per_sweep) pfx[t] = carry + e;
        carry += total;
    }
    if (threadIdx.x == 0) {
        sweep_total[w] = carry;
        __threadfence();
        s_last = atomicAdd(done, 1u) == (unsigned)gridDim.x - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last block: output base of every sweep (int64), each sweep contributes ceil(M / stride) points
    unsigned long long base = 0;
    for (int64_t w0 = 0; w0 < n_sweeps; w0 += SO_THREADS) {
        const int64_t ww = w0 + threadIdx.x;
        const unsigned long long m = ww < n_sweeps ? *(volatile uint32_t*)(sweep_total + ww) : 0u;
        unsigned long long kept = (m + (unsigned)stride - 1u) / (unsigned)stride, total;
        unsigned long long e = block_exclusive_scan<unsigned long long>(kept, s_warp64, &total);
        if (ww < n_sweeps) sweep_base[ww] = (int64_t)(base + e);
        base += total;
    }
    if (threadIdx.x == 0) {
        sweep_base[n_sweeps] = (int64_t)base;
        *done = 0;                        // self-cleaning: ready for the next launch
        *ticket = 0;
    }
}

// ---- 3. emit ------------------------------------------------------------------------------------------
template <typename T>
struct EmitArgs {
    const T* echo;
    const float* cos_tab;
    const float* sin_tab;
    const float* range_res;
    const float* ranges;                 // optional [W][S][E] explicit range per cell (NULL: range_res*j)
    const int32_t* sweep_gain;
    const uint32_t* mask;
    const uint32_t* tile_count;
    const uint32_t* tile_prefix;
    const int64_t* sweep_base;
    float* x;
    float* y;
    float* inten;
    int32_t* gain;
    int64_t cap;
    int64_t total_groups;
    int groups_per_sweep;
    SpokeGeom g;
    int stride;
    FastDiv div_stride, div_bins;
};

__device__ __forceinline__ int select_bit(uint32_t word, uint32_t o) {   // position of the o-th (0-based) set bit
    int pos = 0;
    uint32_t c;
    c = __popc(word & 0xffffu); if (o >= c) { o -= c; pos += 16; word >>= 16; }
    c = __popc(word & 0xffu);   if (o >= c) { o -= c; pos += 8;  word >>= 8; }
    c = __popc(word & 0xfu);    if (o >= c) { o -= c; pos += 4;  word >>= 4; }
    c = __popc(word & 0x3u);    if (o >= c) { o -= c; pos += 2;  word >>= 2; }
    c = word & 1u;              if (o >= c) { pos += 1; }
    return pos;
}

template <typename T>
__global__ void __launch_bounds__(SK_THREADS) spoke_emit_kernel(const EmitArgs<T> a) {
    __shared__ uint4 s_mask[SK_WARPS][SK_ENTRIES];
    __shared__ uint32_t s_rank[SK_WARPS][SK_ENTRIES];
    const unsigned lane = rb_lane();
    const int warp = threadIdx.x >> 5;
    const int64_t gidx = (int64_t)blockIdx.x * SK_WARPS + warp;
    if (gidx >= a.total_groups) return;
    const int w = (int)(gidx / a.groups_per_sweep);
    const int tile0 = (int)(gidx - (int64_t)w * a.groups_per_sweep) * SK_GROUP;
    const int ntile = min(SK_GROUP, a.g.tiles_per_sweep - tile0);
    const int64_t tile_base = (int64_t)w * a.g.tiles_per_sweep + tile0;

    uint32_t my_cnt = 0, my_pfx = 0;
    if ((int)lane < ntile) { my_cnt = a.tile_count[tile_base + lane]; my_pfx = a.tile_prefix[tile_base + lane]; }
    const uint32_t r_begin = __shfl_sync(0xffffffffu, my_pfx, 0);
    const uint32_t r_end = __shfl_sync(0xffffffffu, my_pfx + my_cnt, ntile - 1);
    // kept ranks of the sweep are q * stride; this group owns q in [q0, q1)
    const uint32_t s = (uint32_t)a.stride;
    const uint32_t q0 = fastdiv(r_begin + s - 1u, a.div_stride), q1 = fastdiv(r_end + s - 1u, a.div_stride);
    if (q0 >= q1) return;

    const uint4* __restrict__ mask4 = reinterpret_cast<const uint4*>(a.mask) + tile_base * 32;
    uint4 m[SK_GROUP];
#pragma unroll
    for (int t = 0; t < SK_GROUP; ++t) {
        const uint32_t c = __shfl_sync(0xffffffffu, my_cnt, t);
        m[t] = (t < ntile && c) ? mask4[t * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int t = 0; t < SK_GROUP; ++t) {
        const uint32_t c = __popc(m[t].x) + __popc(m[t].y) + __popc(m[t].z) + __popc(m[t].w);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        const uint32_t pfx = __shfl_sync(0xffffffffu, my_pfx, t);
        s_mask[warp][t * 32 + lane] = m[t];
        s_rank[warp][t * 32 + lane] = t < ntile ? pfx + incl - c : 0xffffffffu;   // rank at the entry's first cell
    }
    __syncwarp();

    const long long out_base = a.sweep_base[w];
    const int gain_label = a.sweep_gain[w];
    const int64_t tab_off = (int64_t)w * a.g.n_spokes;
    const int64_t cell_off = (int64_t)w * a.g.sweep_cells;
    const uint32_t* __restrict__ rk = s_rank[warp];
    for (uint32_t q = q0 + lane; q < q1; q += 32) {
        const uint32_t r = q * s;
        int e = 0;                                    // largest entry with rank[e] <= r (rank[0] <= r always)
#pragma unroll
        for (int step = SK_ENTRIES / 2; step >= 1; step >>= 1)
            if (rk[e + step] <= r) e += step;
        uint32_t o = r - rk[e];
        const uint4 mm = s_mask[warp][e];
        uint32_t word = mm.x;
        int wi = 0;
        uint32_t c = __popc(mm.x);
        if (o >= c) { o -= c; word = mm.y; wi = 1; c = __popc(mm.y);
            if (o >= c) { o -= c; word = mm.z; wi = 2; c = __popc(mm.z);
                if (o >= c) { o -= c; word = mm.w; wi = 3; } } }
        const int cell = (tile0 + (e >> 5)) * SK_TILE + (e & 31) * 128 + wi * 32 + select_bit(word, o);
        const long long pos = out_base + q;
        if (pos < a.cap) {
            const int sp = (int)fastdiv((uint32_t)cell, a.div_bins);
            const int j = cell - sp * a.g.n_bins;
            const float val = (float)__ldg(a.echo + cell_off + cell);
            const float cs = __ldg(a.cos_tab + tab_off + sp), sn = __ldg(a.sin_tab + tab_off + sp);
            const float rng = a.ranges ? __ldg(a.ranges + cell_off + cell)
                                       : __fmul_rn(__ldg(a.range_res + tab_off + sp), (float)j);   // T4:214
            a.x[pos] = __fmul_rn(rng, cs);                  // T4:217
            a.y[pos] = __fmul_rn(rng, sn);                  // T4:218
            a.inten[pos] = val;
            a.gain[pos] = gain_label;
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
    RB_CUDA(rb_launch(ctx, polar_grid_kernel, dim3(blocks), dim3(256), 0, stream, ranges, cos_tab, sin_tab, total, n_cols, x, y));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}

// launches the TMA-staged mask kernel with the ring <STAGE_BYTES x STAGES>; the dynamic shared-memory limit of each
// instance is raised once per context (= per device), in the context's own flags
template <typename T, int STAGE_BYTES, int STAGES>
int launch_mask_tma(rb_ctx* ctx, const T* echo, const SpokeGeom& g, const ThrArg& thr, uint32_t* mask, uint32_t* tile_count,
                    unsigned* ticket, cudaStream_t stream) {
    constexpr int smem = mt_smem_bytes<T, STAGE_BYTES, STAGES>();
    constexpr int ring_id = STAGE_BYTES == 64 * 1024 ? 0 : STAGE_BYTES == 48 * 1024 ? 3 : (STAGES == 4 ? 1 : 2);
    const unsigned bit = 1u << (ring_id * 2 + (sizeof(T) == 1 ? 1 : 0));
    auto kernel = spoke_mask_tma_kernel<T, STAGE_BYTES, STAGES>;
    if (!(ctx->attr_spoke_mask & bit)) {
        RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ctx->attr_spoke_mask |= bit;
    }
    const int64_t want_blocks = rb_div_up(g.total_tiles, (int64_t)mt_tiles<T, STAGE_BYTES>() * mt_chunk_stages<STAGE_BYTES>());
    const unsigned blocks = (unsigned)(want_blocks < ctx->sm_count ? want_blocks : ctx->sm_count);
    RB_CUDA(rb_launch_prio(ctx, true, kernel, dim3(blocks), dim3(mt_threads<T>()), (size_t)smem, stream, echo, g, thr, mask, tile_count, ticket,
                      ctx->opt_spoke_l2_hint));
    return RB_OK;
}

// The mask gate. Blocks in flight run on different streams; left alone, the HBM-bound mask kernels of two blocks start
// together, share the bandwidth (each twice as slow) and leave a stretch afterwards in which NO mask kernel runs and only
// the latency-bound clustering kernels keep the GPU busy (measured: 2.3 ms of every 6.4 ms block). With the gate every
// mask kernel waits for the previous one - of any context of the device - so they run one after the other at full
// bandwidth, and the tail of block k runs beside the mask kernel of block k+1 instead of beside its own twin.
#include <mutex>
static std::mutex g_gate_mu;
static cudaEvent_t g_gate_ev[64];

struct MaskGate {
    rb_ctx* ctx;
    cudaStream_t stream;
    bool held = false;
    int enter() {
        if (!ctx->opt_mask_gate || ctx->device < 0 || ctx->device >= 64) return RB_OK;
        g_gate_mu.lock();
        held = true;
        cudaEvent_t& ev = g_gate_ev[ctx->device];
        if (!ev) RB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        else RB_CUDA(cudaStreamWaitEvent(stream, ev, 0));
        return RB_OK;
    }
    void leave() {                       // right after the mask kernel's launch: what follows on the stream is not gated
        if (!held) return;
        held = false;
        cudaEventRecord(g_gate_ev[ctx->device], stream);
        g_gate_mu.unlock();
    }
    ~MaskGate() { leave(); }
};

template <typename T>
int spoke_to_points_impl(rb_ctx* ctx, const T* echo, const float* cos_tab, const float* sin_tab,
                         const float* range_res, const float* ranges, const int32_t* sweep_gain, int64_t n_sweeps,
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
    RB_REQUIRE(n_sweeps < (int64_t)1 << 31, "too many sweeps in one batch; split the batch");

    SpokeGeom g;
    g.sweep_cells = (int)cells;
    g.tiles_per_sweep = (int)rb_div_up(cells, SK_TILE);
    g.total_tiles = (int64_t)g.tiles_per_sweep * n_sweeps;
    g.n_bins = n_bins;
    g.n_spokes = n_spokes;

    // scratch: mask words (1 bit per cell), tile counts, tile prefixes, sweep totals, {done, ticket}
    const size_t mask_bytes = sizeof(uint32_t) * (size_t)g.total_tiles * SK_WORDS;
    const size_t tile_bytes = (sizeof(uint32_t) * (size_t)g.total_tiles + 255) & ~size_t(255);
    const size_t sweep_bytes = (sizeof(uint32_t) * (size_t)n_sweeps + 255) & ~size_t(255);
    void *mask_v, *tiles_v, *flags_v;
    RB_TRY(rb_scratch_get(ctx, RB_S_SPOKE_MASK, mask_bytes, &mask_v));
    RB_TRY(rb_scratch_get(ctx, RB_S_TILE_STATUS, 2 * tile_bytes + sweep_bytes, &tiles_v));
    rb_scratch& fslot = ctx->slots[RB_S_SWEEP_FLAGS];
    const void* flags_before = fslot.ptr;
    RB_TRY(rb_scratch_get(ctx, RB_S_SWEEP_FLAGS, 256, &flags_v));
    if (fslot.ptr != flags_before) RB_CUDA(cudaMemsetAsync(flags_v, 0, 256, stream));   // counters clean themselves afterwards
    uint32_t* mask = (uint32_t*)mask_v;
    uint32_t* tile_count = (uint32_t*)tiles_v;
    uint32_t* tile_prefix = (uint32_t*)((unsigned char*)tiles_v + tile_bytes);
    uint32_t* sweep_total = (uint32_t*)((unsigned char*)tiles_v + 2 * tile_bytes);
    unsigned* done = (unsigned*)flags_v;
    unsigned* ticket = done + 32;                    // its own 128-byte line

    const bool prof = ctx->opt_spoke_profile != 0;
    if (prof && !ctx->spoke_ev[0])
        for (int i = 0; i < 4; ++i) RB_CUDA(cudaEventCreate(&ctx->spoke_ev[i]));
    if (prof) RB_CUDA(cudaEventRecord(ctx->spoke_ev[0], stream));

    // the TMA-staged kernel moves whole tiles with 16-byte bulk copies: the sweeps must be a multiple of 16 bytes
    const bool vec = (cells % Elem<T>::ALIGN_CELLS == 0) && (((uintptr_t)echo & 15u) == 0);
    // 0 = auto (TMA-staged when the shape allows 16-byte bulk copies), 1 = register-staged, 2 = require TMA
    const int variant = ctx->opt_spoke_mask_variant;
    RB_REQUIRE(variant != 2 || vec, "TMA-staged mask kernel needs sweeps of a multiple of 16 bytes and a 16-byte aligned echo pointer");
    MaskGate gate{ctx, stream};
    RB_TRY(gate.enter());
    if (vec && variant != 1) {
        const ThrArg thr = make_thr(threshold);
        switch (ctx->opt_spoke_ring) {
            case 0: RB_TRY((launch_mask_tma<T, 64 * 1024, 3>(ctx, echo, g, thr, mask, tile_count, ticket, stream))); break;
            case 1: RB_TRY((launch_mask_tma<T, 32 * 1024, 4>(ctx, echo, g, thr, mask, tile_count, ticket, stream))); break;
            case 3: RB_TRY((launch_mask_tma<T, 48 * 1024, 2>(ctx, echo, g, thr, mask, tile_count, ticket, stream))); break;
            default: RB_TRY((launch_mask_tma<T, 32 * 1024, 3>(ctx, echo, g, thr, mask, tile_count, ticket, stream))); break;
        }
        ctx->spoke_last_variant = 2;
    } else {
        const int64_t want_blocks = rb_div_up(g.total_tiles, SK_WARPS);
        const int64_t max_blocks = (int64_t)ctx->sm_count * SK_CTAS_PER_SM;
        const unsigned blocks = (unsigned)(want_blocks < max_blocks ? want_blocks : max_blocks);
        if (vec && sizeof(T) == 4) RB_CUDA(rb_launch(ctx, spoke_mask_kernel<T, sizeof(T) == 4>, dim3(blocks), dim3(SK_THREADS), 0, stream, echo, g, threshold, mask, tile_count));
        else RB_CUDA(rb_launch(ctx, spoke_mask_kernel<T, false>, dim3(blocks), dim3(SK_THREADS), 0, stream, echo, g, threshold, mask, tile_count));
        ctx->spoke_last_variant = 1;
    }
    gate.leave();
    RB_LAUNCH_CHECK(ctx);
    if (prof) RB_CUDA(cudaEventRecord(ctx->spoke_ev[1], stream));

    RB_CUDA(rb_launch(ctx, spoke_offsets_kernel, dim3((unsigned)n_sweeps), dim3(SO_THREADS), 0, stream, tile_count, tile_prefix, g.tiles_per_sweep, n_sweeps,
                                                                       stride, sweep_total, sweep_base, done, ticket));
    RB_LAUNCH_CHECK(ctx);
    if (prof) RB_CUDA(cudaEventRecord(ctx->spoke_ev[2], stream));

    if (cap > 0) {
        EmitArgs<T> a;
        a.echo = echo; a.cos_tab = cos_tab; a.sin_tab = sin_tab; a.range_res = range_res; a.ranges = ranges;
        a.sweep_gain = sweep_gain;
        a.mask = mask; a.tile_count = tile_count; a.tile_prefix = tile_prefix; a.sweep_base = sweep_base;
        a.x = x; a.y = y; a.inten = inten; a.gain = gain;
        a.cap = cap;
        a.groups_per_sweep = (int)rb_div_up(g.tiles_per_sweep, SK_GROUP);
        a.total_groups = (int64_t)a.groups_per_sweep * n_sweeps;
        a.g = g;
        a.stride = stride;
        a.div_stride = make_fastdiv((uint32_t)stride);
        a.div_bins = make_fastdiv((uint32_t)n_bins);
        const int64_t eblocks = rb_div_up(a.total_groups, SK_WARPS);
        RB_REQUIRE(eblocks < (int64_t)1 << 31, "too many tiles in one batch; split the batch");
        RB_CUDA(rb_launch(ctx, spoke_emit_kernel<T>, dim3((unsigned)eblocks), dim3(SK_THREADS), 0, stream, a));
        RB_LAUNCH_CHECK(ctx);
    }
    if (prof) RB_CUDA(cudaEventRecord(ctx->spoke_ev[3], stream));
    return RB_OK;
}

extern "C" int rb_spoke_to_points(rb_ctx* ctx, const float* echo, const float* cos_tab, const float* sin_tab,
                                  const float* range_res, const float* ranges, const int32_t* sweep_gain,
                                  int64_t n_sweeps, int n_spokes, int n_bins, float threshold, int stride, float* x, float* y,
                                  float* inten, int32_t* gain, int64_t cap, int64_t* sweep_base, void* stream_) {
    return spoke_to_points_impl<float>(ctx, echo, cos_tab, sin_tab, range_res, ranges, sweep_gain, n_sweeps, n_spokes, n_bins,
                                       threshold, stride, x, y, inten, gain, cap, sweep_base, stream_);
}

extern "C" int rb_spoke_to_points_u8(rb_ctx* ctx, const uint8_t* echo, const float* cos_tab, const float* sin_tab,
                                     const float* range_res, const float* ranges, const int32_t* sweep_gain,
                                     int64_t n_sweeps, int n_spokes, int n_bins, float threshold, int stride, float* x, float* y,
                                     float* inten, int32_t* gain, int64_t cap, int64_t* sweep_base, void* stream_) {
    return spoke_to_points_impl<uint8_t>(ctx, echo, cos_tab, sin_tab, range_res, ranges, sweep_gain, n_sweeps, n_spokes, n_bins,
                                         threshold, stride, x, y, inten, gain, cap, sweep_base, stream_);
}

extern "C" int rb_frame_offsets(rb_ctx* ctx, const int64_t* sweep_base, int64_t n_frames, int gains_per_frame,
                                int64_t* frame_off, void* stream_) {
    RB_REQUIRE(ctx && sweep_base && frame_off, "NULL argument");
    RB_REQUIRE(n_frames >= 0 && gains_per_frame >= 1, "bad sizes");
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned blocks = (unsigned)rb_div_up(n_frames + 1, 256);
    RB_CUDA(rb_launch(ctx, frame_offsets_kernel, dim3(blocks), dim3(256), 0, stream, sweep_base, n_frames, gains_per_frame, frame_off));
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
    RB_CUDA(rb_launch(ctx, expand_times_kernel, dim3(blocks), dim3(256), 0, stream, frame_off, frame_ids, n_frames, times));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}
