// Ingest: one radar sweep CSV parsed on the GPU (SURVEY.md section 8 f, rank 2; reference 4_temporal_object_tracker.py:
// 184-209 - pd.read_csv(path, header=None, names=[Status, Scale, Range, Gain, Angle, Echo_0..Echo_{E-1}], skiprows=1) is
// ~70 % of the reference's per-sweep load time [SURVEY, probed]).
//
// What the device does: finds the lines, and turns the E echo columns of every data line - 99.5 % of the bytes - into
// uint8, which is what rb_spoke_to_points_u8 consumes (the parsed echoes never exist as float32 on the host and never
// cross PCIe a second time). What stays on the host: the five leading fields of each line (Scale may be a decimal; the
// caller runs the reference's own parser on just those few bytes per line, from the offsets this call returns), file
// discovery and the error message.
//
// The grammar accepted here is deliberately narrow: every echo field is empty (pandas: NaN -> fillna(0) -> 0, T4:206)
// or 1..3 decimal digits with a value <= 255; every data line has exactly E + 5 fields; no blank lines. Anything else
// only sets a status bit - the caller then parses that file with the reference's parser, so exotic input keeps the
// reference's exact behaviour (including its quirks) instead of an imitation of it.
#include "common.cuh"

namespace {

constexpr int CSV_THREADS = 256;
constexpr int CSV_CHUNK = CSV_THREADS * 16;          // bytes per block in the newline passes

// exclusive prefix of a per-thread count over the block; returns the block total through *total
__device__ __forceinline__ int block_exclusive(int v, int* s_warp, int* total) {
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < CSV_THREADS / 32; ++w) {
        const int t = s_warp[w];
        if (w < (int)wid) base += t;
        sum += t;
    }
    __syncthreads();
    *total = sum;
    return base + incl - v;
}

// pass 1 (FILL = false): newlines per chunk. pass 2 (FILL = true): their byte offsets, in order.
template <bool FILL>
__global__ void __launch_bounds__(CSV_THREADS) csv_newline_kernel(const uint8_t* __restrict__ text, int64_t n_bytes,
                                                                 int32_t* __restrict__ chunk_count,
                                                                 const int32_t* __restrict__ chunk_base, int32_t* __restrict__ nl,
                                                                 int64_t nl_cap, int32_t* __restrict__ info) {
    __shared__ int s_warp[CSV_THREADS / 32];
    const int64_t first = (int64_t)blockIdx.x * CSV_CHUNK + (int64_t)threadIdx.x * 16;
    unsigned hits = 0;
    bool lone_cr = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (first + j >= n_bytes) break;
        const uint8_t ch = text[first + j];
        if (ch == '\n') hits |= 1u << j;
        // a carriage return that is not part of "\r\n": the reference's parser ends a line there, this grammar does not
        if (!FILL && ch == '\r' && (first + j + 1 >= n_bytes || text[first + j + 1] != '\n')) lone_cr = true;
    }
    if (!FILL && lone_cr) atomicOr(info + 1, RB_CSV_NOT_INTEGER);
    int total;
    const int rank = block_exclusive(__popc(hits), s_warp, &total);
    if (!FILL) {
        if (threadIdx.x == 0) chunk_count[blockIdx.x] = total;
        return;
    }
    int64_t slot = (int64_t)chunk_base[blockIdx.x] + rank;
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (hits >> j & 1) { if (slot < nl_cap) nl[slot] = (int32_t)(first + j); ++slot; }
}

// One block per data row (row r = line r + 1 of the file).
__global__ void __launch_bounds__(CSV_THREADS) csv_parse_rows_kernel(const uint8_t* __restrict__ text, int64_t n_bytes,
                                                                    const int32_t* __restrict__ nl, const int32_t* __restrict__ n_nl_dev,
                                                                    int n_echo, int64_t max_rows, uint8_t* __restrict__ echo,
                                                                    int32_t* __restrict__ row_start, int32_t* __restrict__ prefix_end,
                                                                    int32_t* __restrict__ info) {
    __shared__ int s_warp[CSV_THREADS / 32];
    const int n_nl = *n_nl_dev;
    // lines = newline-terminated pieces, plus an unterminated last one
    const bool open_tail = n_bytes > 0 && text[n_bytes - 1] != '\n';
    const int n_lines = n_nl + (open_tail ? 1 : 0);
    const int64_t r = blockIdx.x;
    if (r == 0 && threadIdx.x == 0) info[0] = n_lines;
    if (n_nl == 0 || r + 1 >= n_lines) return;                       // (no header line, or) no such row
    if (r >= max_rows) { if (threadIdx.x == 0) atomicOr(info + 1, RB_CSV_CAPACITY); return; }
    const int start = nl[r] + 1;
    int end = (r + 1 < n_nl) ? nl[r + 1] : (int)n_bytes;
    if (end > start && text[end - 1] == '\r') --end;                 // \r\n line ends
    if (threadIdx.x == 0) { row_start[r] = start; prefix_end[r] = end; }
    if (end <= start) { if (threadIdx.x == 0) atomicOr(info + 1, RB_CSV_BLANK_LINE); return; }
    uint8_t* __restrict__ out = echo + r * (int64_t)n_echo;
    int commas_before = 0;                                           // in the chunks already done
    unsigned bad = 0;
    for (int c0 = start; c0 < end; c0 += CSV_THREADS) {
        const int i = c0 + (int)threadIdx.x;
        const bool in = i < end;
        const uint8_t ch = in ? text[i] : 0;
        const bool comma = in && ch == ',';
        if (in && ch == '"') bad |= RB_CSV_NOT_INTEGER;              // quoted fields: leave the file to the reference's parser
        int total;
        const int before = commas_before + block_exclusive(comma ? 1 : 0, s_warp, &total);     // commas left of char i
        if (comma && before == 4) prefix_end[r] = i;                 // the fifth comma ends the leading fields
        const bool field_start = in && (i == start || text[i - 1] == ',');
        if (field_start && before >= 5) {
            const int col = before - 5;
            // the field: [i, first comma or end of line)
            int v = 0, len = 0;
            bool ok = true;
            for (int k = i; k < end; ++k) {
                const uint8_t d = text[k];
                if (d == ',') break;
                if (d < '0' || d > '9' || len == 3) { ok = false; break; }
                v = v * 10 + (d - '0');
                ++len;
            }
            if (!ok) bad |= RB_CSV_NOT_INTEGER;
            else if (v > 255) bad |= RB_CSV_OUT_OF_RANGE;
            else if (col >= n_echo) bad |= RB_CSV_RAGGED;
            else out[col] = (uint8_t)v;                              // empty field: 0 (NaN -> fillna(0))
        }
        commas_before += total;
    }
    if (commas_before != n_echo + 4) bad |= RB_CSV_RAGGED;           // exactly 5 + E fields per line
    if (bad) atomicOr(info + 1, (int)bad);
}

}  // namespace

extern "C" int rb_csv_parse_sweep(rb_ctx* ctx, const uint8_t* text, int64_t n_bytes, int n_echo_columns, int64_t max_rows,
                                  uint8_t* echo, int32_t* row_start, int32_t* prefix_end, int32_t* info, void* stream_) {
    RB_REQUIRE(ctx && info, "NULL argument");
    RB_REQUIRE(n_bytes >= 0 && n_bytes < ((int64_t)1 << 31) && n_echo_columns > 0 && max_rows >= 0, "bad sizes");
    cudaStream_t stream = (cudaStream_t)stream_;
    RB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t) * 2, stream));
    if (n_bytes == 0) return RB_OK;
    RB_REQUIRE(text && (max_rows == 0 || (echo && row_start && prefix_end)), "NULL buffers");
    const int64_t chunks = rb_div_up(n_bytes, CSV_CHUNK);
    const int64_t nl_cap = max_rows + 2;                             // header + rows (+ one to notice an overflow)
    void* scratch;
    RB_TRY(rb_scratch_get(ctx, RB_S_CSV, sizeof(int32_t) * (size_t)(2 * (chunks + 1) + nl_cap + 1), &scratch));
    int32_t* chunk_count = (int32_t*)scratch;
    int32_t* chunk_base = chunk_count + chunks + 1;
    int32_t* nl = chunk_base + chunks + 1;
    RB_CUDA(rb_launch(ctx, csv_newline_kernel<false>, dim3((unsigned)chunks), dim3(CSV_THREADS), 0, stream, text, n_bytes, chunk_count, nullptr, nullptr, 0, info));
    RB_LAUNCH_CHECK(ctx);
    RB_TRY(rb_exclusive_scan_i32(ctx, chunk_count, chunk_base, chunks, chunk_base + chunks, stream));      // [chunks] = total
    RB_CUDA(rb_launch(ctx, csv_newline_kernel<true>, dim3((unsigned)chunks), dim3(CSV_THREADS), 0, stream, text, n_bytes, nullptr, chunk_base, nl, nl_cap, info));
    RB_LAUNCH_CHECK(ctx);
    if (max_rows > 0) RB_CUDA(cudaMemsetAsync(echo, 0, (size_t)max_rows * (size_t)n_echo_columns, stream));
    // one block per possible row; a file with more lines than max_rows + 1 reports RB_CSV_CAPACITY
    RB_CUDA(rb_launch(ctx, csv_parse_rows_kernel, dim3((unsigned)(max_rows + 1)), dim3(CSV_THREADS), 0, stream, text, n_bytes, nl, chunk_base + chunks, n_echo_columns,
                                                                              max_rows, echo, row_start, prefix_end, info));
    RB_LAUNCH_CHECK(ctx);
    return RB_OK;
}
