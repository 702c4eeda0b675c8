// Column reordering + dense/residual classification + device format build, all on the GPU.
//
// Replaces (bit-exact, integer only):
//   colReordering_cpu + analysisDescendingOrderColSegment  (src/colReordering.cu:274-404, 244-271)
//     -- a host/OpenMP routine in the reference that allocates an N-wide counter per panel
//   RPHM::RPHM                                              (src/BSMR.cpp:83-265)
//     -- host hash maps + 12 H2D copies in the reference
//
// Formulation (sort based, O(nnz) memory, no per-panel N-wide arrays):
//   A. enumerate every nnz of the reordered rows as key = (panel | column | row-in-panel),
//      payload = CSR position; one radix sort puts the entries of a (panel, column) pair
//      next to each other with the panel rows ascending
//   B. run-length encode (panel, column) -> the non-empty columns of every panel in ascending
//      column order together with their nnz counts (1..16)
//   C. stable radix sort of the runs by (panel, 16 - count): count descending, ties keep the
//      ascending column order -- exactly thrust's stable sort_by_key with greater<> (:333-336)
//   D/E. per panel: pad to a multiple of 16 with sentinel column N, count 0 (:338-343); a
//      16-column block is dense iff its nnz >= ceil(delta*256) (:244-261); counts are sorted so the
//      dense blocks form a prefix; everything after it (sentinels included) is residual
//   F. three exclusive scans (:360-378)
//   G. scatter the column lists
//   H. device format: per dense tile (<= 128 columns = 8 reference blocks of one panel) a
//      [16 rows][128 cols] table of CSR positions (NULL where S has no entry), per residual nnz
//      (CSR position, B column, A row) in the reference's order (panel, residual column
//      order, row in panel).
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace bsmr {
namespace {

constexpr int kThreads = 256;

inline int grid_for(uint64_t n, int per_cta, int sm_count) {
    uint64_t g = (n + per_cta - 1) / per_cta;
    const uint64_t cap = (uint64_t)sm_count * 16;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (int)g;
}

__global__ void fill_u32_kernel(uint32_t* p, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// A0: nnz of every reordered row
__global__ void row_lengths_kernel(const uint32_t* __restrict__ rows, uint32_t R, const uint32_t* __restrict__ row_offsets,
                                   uint32_t* __restrict__ len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = rows[i];
        len[i] = row_offsets[r + 1] - row_offsets[r];
    }
}

// A1: key = panel << (cbits + 4) | col << 4 | rel, payload = CSR position (one warp per reordered row)
__global__ void make_keys_kernel(const uint32_t* __restrict__ rows, uint32_t R, const uint32_t* __restrict__ row_offsets,
                                 const uint32_t* __restrict__ col_indices, const uint32_t* __restrict__ start,
                                 int cbits, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < R; i += stride) {
        const uint32_t r = rows[i];
        const uint32_t b = row_offsets[r], e = row_offsets[r + 1];
        const uint64_t hi = ((uint64_t)(i / kPanel)) << (cbits + 4);
        const uint64_t rel = i % kPanel;
        const uint32_t dst = start[i];
        for (uint32_t k = b + lane; k < e; k += 32) {
            keys[dst + (k - b)] = hi | ((uint64_t)col_indices[k] << 4) | rel;
            vals[dst + (k - b)] = k;
        }
    }
}

struct ShiftRel {
    __host__ __device__ uint64_t operator()(uint64_t k) const { return k >> 4; }
};

// C0: key2 = panel << 4 | (16 - count); runs per panel counted with one atomic per (warp, distinct panel)
__global__ void run_keys_kernel(const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ counts, uint32_t num_runs,
                                int cbits, uint32_t* __restrict__ key2, uint32_t* __restrict__ iota,
                                uint32_t* __restrict__ runs_per_panel) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n32 = ((uint64_t)num_runs + 31) & ~31ull;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n32; u += (uint64_t)gridDim.x * blockDim.x) {
        const bool valid = u < num_runs;
        const uint32_t panel = valid ? (uint32_t)(ukeys[u] >> cbits) : 0xFFFFFFFFu;
        if (valid) {
            key2[u] = (panel << 4) | (16u - counts[u]);
            iota[u] = (uint32_t)u;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, panel);
        if (valid && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(runs_per_panel + panel, (uint32_t)__popc(peers));
    }
}

// E: classify the 16-column blocks of one panel (one CTA per panel; a warp pass covers two blocks)
__global__ void __launch_bounds__(256)
classify_kernel(uint32_t panels, const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ order,
                const uint32_t* __restrict__ counts, uint32_t threshold,
                uint32_t* __restrict__ n_dense, uint32_t* __restrict__ n_sparse,
                uint32_t* __restrict__ n_sparse_data, uint32_t* __restrict__ n_dense_data,
                uint32_t* __restrict__ n_tiles) {
    __shared__ uint32_t red[3][8];
    const uint32_t lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    for (uint32_t p = blockIdx.x; p < panels; p += gridDim.x) {
        const uint32_t s0 = run_start[p], s1 = run_start[p + 1];
        const uint32_t U = s1 - s0;
        const uint32_t padded = (U + kBlockCols - 1) / kBlockCols * kBlockCols;
        uint32_t dense_cols = 0, dense_data = 0, total = 0;
        // lanes 0-15 -> block b, lanes 16-31 -> block b+1
        for (uint32_t base = wv * 32; base < padded; base += 256) {
            const uint32_t k = base + lane;
            uint32_t s = (k < U) ? counts[order[s0 + k]] : 0u;
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1) s += __shfl_xor_sync(0xffffffffu, s, w);
            const uint32_t s_lo = __shfl_sync(0xffffffffu, s, 0), s_hi = __shfl_sync(0xffffffffu, s, 16);
            total += s_lo + s_hi;
            if (s_lo >= threshold) { dense_cols += kBlockCols; dense_data += s_lo; }
            if (base + 16 < padded && s_hi >= threshold) { dense_cols += kBlockCols; dense_data += s_hi; }
        }
        if (lane == 0) { red[0][wv] = dense_cols; red[1][wv] = dense_data; red[2][wv] = total; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t dc = 0, dd = 0, tt = 0;
            for (int i = 0; i < 8; ++i) { dc += red[0][i]; dd += red[1][i]; tt += red[2][i]; }
            n_dense[p] = dc;
            n_sparse[p] = padded - dc;
            n_dense_data[p] = dd;
            n_sparse_data[p] = tt - dd;
            n_tiles[p] = (dc + kTileCols - 1) / kTileCols;
        }
        __syncthreads();
    }
}

// G: column lists + per-run residual nnz (0 for dense runs) + inverse of `order`
__global__ void scatter_cols_kernel(uint32_t num_runs, const uint32_t* __restrict__ key2_sorted, const uint32_t* __restrict__ order,
                                    const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ counts,
                                    const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ n_dense,
                                    const uint32_t* __restrict__ d_off, const uint32_t* __restrict__ s_off,
                                    uint32_t* __restrict__ dense_cols, uint32_t* __restrict__ sparse_cols,
                                    uint32_t* __restrict__ sparse_cnt, uint32_t* __restrict__ rank_of_run) {
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < num_runs; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = key2_sorted[s] >> 4;
        const uint32_t u = order[s];
        const uint32_t k = (uint32_t)s - run_start[p];
        const uint32_t col = (uint32_t)(ukeys[u] & 0xffffffffull) ;
        rank_of_run[u] = (uint32_t)s;
        if (k < n_dense[p]) {
            dense_cols[d_off[p] + k] = col;
            sparse_cnt[s] = 0;
        } else {
            sparse_cols[s_off[p] + (k - n_dense[p])] = col;
            sparse_cnt[s] = counts[u];
        }
    }
}

// H0: tile metadata (thread per panel)
__global__ void tile_meta_kernel(uint32_t panels, const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ n_dense,
                                 const uint32_t* __restrict__ d_off, uint32_t* __restrict__ tile_panel,
                                 uint32_t* __restrict__ tile_col_begin, uint32_t* __restrict__ tile_ncols,
                                 uint4* __restrict__ tile_meta) {
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < panels; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t nd = n_dense[p];
        const uint32_t t0 = tile_base[p];
        for (uint32_t c = 0, t = t0; c < nd; c += kTileCols, ++t) {
            tile_panel[t] = (uint32_t)p;
            tile_col_begin[t] = d_off[p] + c;
            tile_ncols[t] = nd - c < kTileCols ? nd - c : kTileCols;
            tile_meta[t] = make_uint4((uint32_t)p, d_off[p] + c, nd - c < kTileCols ? nd - c : kTileCols, 0u);
        }
    }
}

// H1: place every nnz (thread per run; a run has <= 16 entries, rows ascending)
__global__ void place_entries_kernel(uint32_t num_runs, const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ counts,
                                     const uint32_t* __restrict__ run_off, const uint32_t* __restrict__ rank_of_run,
                                     const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ n_dense,
                                     const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ res_start,
                                     const uint64_t* __restrict__ keys_sorted, const uint32_t* __restrict__ vals_sorted,
                                     const uint32_t* __restrict__ rows, uint32_t R, int cbits,
                                     uint32_t* __restrict__ scatter, uint32_t* __restrict__ res_out,
                                     uint32_t* __restrict__ res_col, uint32_t* __restrict__ res_row,
                                     uint8_t* __restrict__ res_rel, const uint32_t* __restrict__ start,
                                     const uint32_t* __restrict__ row_offsets, uint32_t* __restrict__ res_flag) {
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < num_runs; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(ukeys[u] >> cbits);
        const uint32_t col = (uint32_t)(ukeys[u] & 0xffffffffull);
        const uint32_t s = rank_of_run[u];
        const uint32_t k = s - run_start[p];
        const uint32_t e0 = run_off[u], cnt = counts[u];
        if (k < n_dense[p]) {
            const uint32_t t = tile_base[p] + k / kTileCols;
            const uint32_t c = k % kTileCols;
            for (uint32_t j = 0; j < cnt; ++j) {
                const uint32_t rel = (uint32_t)(keys_sorted[e0 + j] & 15ull);
                scatter[((size_t)t * kPanel + rel) * kTileCols + c] = vals_sorted[e0 + j];
            }
        } else {
            const uint32_t pos0 = res_start[s];
            for (uint32_t j = 0; j < cnt; ++j) {
                const uint32_t rel = (uint32_t)(keys_sorted[e0 + j] & 15ull);
                const uint32_t ri = p * kPanel + rel;
                res_out[pos0 + j] = vals_sorted[e0 + j];
                res_col[pos0 + j] = col;
                const uint32_t row = rows[ri < R ? ri : R - 1];
                res_row[pos0 + j] = row;
                res_rel[pos0 + j] = (uint8_t)rel;
                // position of this nnz in the reordered-row enumeration (make_keys_kernel's index)
                res_flag[start[ri] + (vals_sorted[e0 + j] - row_offsets[row])] = 1u;
            }
        }
    }
}

// H2: the residual entries once more, ROW-sorted, for the residual kernel (one warp per reordered row)
__global__ void compact_rows_kernel(const uint32_t* __restrict__ rows, uint32_t R, const uint32_t* __restrict__ row_offsets,
                                    const uint32_t* __restrict__ col_indices, const uint32_t* __restrict__ start,
                                    const uint32_t* __restrict__ res_flag, const uint32_t* __restrict__ res_pos,
                                    uint32_t* __restrict__ rr_row, uint32_t* __restrict__ rr_col, uint32_t* __restrict__ rr_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < R; i += stride) {
        const uint32_t r = rows[i];
        const uint32_t b = row_offsets[r], e = row_offsets[r + 1];
        const uint32_t q0 = start[i];
        for (uint32_t k = b + lane; k < e; k += 32) {
            const uint32_t q = q0 + (k - b);
            if (res_flag[q]) {
                const uint32_t pos = res_pos[q];
                rr_row[pos] = r;
                rr_col[pos] = col_indices[k];
                rr_out[pos] = k;
            }
        }
    }
}

int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

template <typename T>
int d2h(std::vector<T>& dst, const T* src, size_t n, cudaStream_t s) {
    dst.resize(n);
    if (n) BSMR_CUDA_OK(cudaMemcpyAsync(dst.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    return BSMR_OK;
}

}  // namespace

int col_reorder_and_format(bsmr_plan* plan, float delta) {
    bsmr_ctx* ctx = plan->ctx;
    cudaStream_t st = ctx->stream;
    const uint32_t R = (uint32_t)plan->h_reordered_rows.size();
    const uint32_t panels = plan->num_row_panels;
    const uint32_t N = plan->N;
    const int sm = ctx->sm_count;
    Workspace* ws = &ctx->ws;
    ws->reset();
    // numNonZeroThreshold = (UIN)ceil(delta * BLOCK_SIZE)   (src/colReordering.cu:246)
    const uint32_t threshold = static_cast<uint32_t>(std::ceil(delta * static_cast<float>(kPanel * kBlockCols)));

    cudaEvent_t e0, e1, e2;
    BSMR_CUDA_OK(cudaEventCreate(&e0));
    BSMR_CUDA_OK(cudaEventCreate(&e1));
    BSMR_CUDA_OK(cudaEventCreate(&e2));
    struct EvGuard { cudaEvent_t a, b, c; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); } } guard{e0, e1, e2};
    BSMR_CUDA_OK(cudaEventRecord(e0, st));

    // per-panel outputs (+1 for the scans)
    TmpBuf<uint32_t> n_dense(ws), n_sparse(ws), n_sparse_data(ws), n_dense_data(ws), n_tiles(ws), d_off(ws), s_off(ws), sv_off(ws), tile_base(ws), run_start(ws), runs_per_panel(ws);
    const size_t P1 = static_cast<size_t>(panels) + 1;
    BSMR_TRY(n_dense.alloc(P1)); BSMR_TRY(n_sparse.alloc(P1)); BSMR_TRY(n_sparse_data.alloc(P1));
    BSMR_TRY(n_dense_data.alloc(P1)); BSMR_TRY(n_tiles.alloc(P1)); BSMR_TRY(d_off.alloc(P1)); BSMR_TRY(s_off.alloc(P1));
    BSMR_TRY(sv_off.alloc(P1)); BSMR_TRY(tile_base.alloc(P1)); BSMR_TRY(run_start.alloc(P1)); BSMR_TRY(runs_per_panel.alloc(P1));
    for (TmpBuf<uint32_t>* b : {&n_dense, &n_sparse, &n_sparse_data, &n_dense_data, &n_tiles, &runs_per_panel})
        BSMR_CUDA_OK(cudaMemsetAsync(b->ptr, 0, b->bytes(), st));

    // ---- A: keys ------------------------------------------------------------------------
    TmpBuf<uint32_t> len(ws), start(ws);
    BSMR_TRY(len.alloc(static_cast<size_t>(R) + 1));
    BSMR_TRY(start.alloc(static_cast<size_t>(R) + 1));
    BSMR_CUDA_OK(cudaMemsetAsync(len.ptr, 0, len.bytes(), st));
    TmpBuf<uint8_t> temp(ws);
    auto ensure_temp = [&](size_t bytes) -> int {
        if (bytes > temp.count) return temp.alloc(bytes + bytes / 8 + 256);
        return BSMR_OK;
    };
    uint32_t total = 0;
    if (R) {
        row_lengths_kernel<<<grid_for(R, kThreads, sm), kThreads, 0, st>>>(plan->reordered_rows.ptr, R, plan->row_offsets.ptr, len.ptr);
        ctx->launches++;
        size_t tb = 0;
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len.ptr, start.ptr, static_cast<size_t>(R) + 1, st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, len.ptr, start.ptr, static_cast<size_t>(R) + 1, st));
        ctx->launches++;
        BSMR_CUDA_OK(cudaMemcpyAsync(&total, start.ptr + R, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        BSMR_CUDA_OK(cudaStreamSynchronize(st));
    }

    const int cbits = 32;                                  // column field width inside the run key
    const int pbits = bits_for(panels ? panels - 1 : 0);
    TmpBuf<uint64_t> keys_a(ws), keys_b(ws), ukeys(ws);
    TmpBuf<uint32_t> vals_a(ws), vals_b(ws), counts(ws), num_runs_d(ws), key2_a(ws), key2_b(ws), ord_a(ws), ord_b(ws), run_off(ws), sparse_cnt(ws), res_start(ws), rank_of_run(ws);
    uint32_t num_runs = 0;
    const uint64_t* keys_sorted = nullptr;
    const uint32_t* vals_sorted = nullptr;
    const uint32_t* key2_sorted = nullptr;
    const uint32_t* order = nullptr;
    if (total) {
        BSMR_TRY(keys_a.alloc(total)); BSMR_TRY(keys_b.alloc(total));
        BSMR_TRY(vals_a.alloc(total)); BSMR_TRY(vals_b.alloc(total));
        make_keys_kernel<<<grid_for((uint64_t)R * 32, kThreads, sm), kThreads, 0, st>>>(
            plan->reordered_rows.ptr, R, plan->row_offsets.ptr, plan->col_indices.ptr, start.ptr, cbits, keys_a.ptr, vals_a.ptr);
        ctx->launches++;
        // sort by (panel, column, row-in-panel); only the populated bit range is sorted
        const int col_bits = bits_for(N ? N - 1 : 0);
        // key layout: [panel : pbits][col : 32][rel : 4]; unused high column bits are zero, so sorting
        // bits [0, 4+col_bits) and [36, 36+pbits) is enough -- done as two stable passes (LSD order)
        size_t tb = 0, tb2 = 0;
        cub::DoubleBuffer<uint64_t> dk(keys_a.ptr, keys_b.ptr);
        cub::DoubleBuffer<uint32_t> dv(vals_a.ptr, vals_b.ptr);
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, static_cast<int64_t>(total), 0, 4 + col_bits, st));
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, static_cast<int64_t>(total), 4 + cbits, 4 + cbits + pbits, st));
        BSMR_TRY(ensure_temp(std::max(tb, tb2)));
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb, dk, dv, static_cast<int64_t>(total), 0, 4 + col_bits, st));
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb2, dk, dv, static_cast<int64_t>(total), 4 + cbits, 4 + cbits + pbits, st));
        ctx->launches += 2;
        keys_sorted = dk.Current();
        vals_sorted = dv.Current();

        // ---- B: runs of (panel, column) ---------------------------------------------------
        BSMR_TRY(ukeys.alloc(total)); BSMR_TRY(counts.alloc(total)); BSMR_TRY(num_runs_d.alloc(1));
        auto shifted = thrust::make_transform_iterator(keys_sorted, ShiftRel());
        BSMR_CUDA_OK(cub::DeviceRunLengthEncode::Encode(nullptr, tb, shifted, ukeys.ptr, counts.ptr, num_runs_d.ptr, static_cast<int64_t>(total), st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceRunLengthEncode::Encode(temp.ptr, tb, shifted, ukeys.ptr, counts.ptr, num_runs_d.ptr, static_cast<int64_t>(total), st));
        ctx->launches++;
        BSMR_CUDA_OK(cudaMemcpyAsync(&num_runs, num_runs_d.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        BSMR_CUDA_OK(cudaStreamSynchronize(st));

        // ---- C: stable sort of the runs by (panel, 16 - count) ----------------------------
        BSMR_TRY(key2_a.alloc(num_runs)); BSMR_TRY(key2_b.alloc(num_runs)); BSMR_TRY(ord_a.alloc(num_runs)); BSMR_TRY(ord_b.alloc(num_runs));
        run_keys_kernel<<<grid_for(num_runs, kThreads, sm), kThreads, 0, st>>>(ukeys.ptr, counts.ptr, num_runs, cbits, key2_a.ptr, ord_a.ptr, runs_per_panel.ptr);
        ctx->launches++;
        cub::DoubleBuffer<uint32_t> dk2(key2_a.ptr, key2_b.ptr), dv2(ord_a.ptr, ord_b.ptr);
        const int end_bit = std::min(32, 4 + pbits);
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk2, dv2, static_cast<int64_t>(num_runs), 0, end_bit, st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb, dk2, dv2, static_cast<int64_t>(num_runs), 0, end_bit, st));
        ctx->launches++;
        key2_sorted = dk2.Current();
        order = dv2.Current();
    }
    if (panels > (1u << 28)) {
        set_error("too many row panels (%u) for the 32-bit run key", panels);
        return BSMR_ERR_UNSUPPORTED;
    }

    // ---- D: first run of every panel ----------------------------------------------------------
    {
        size_t tb = 0;
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, runs_per_panel.ptr, run_start.ptr, P1, st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, runs_per_panel.ptr, run_start.ptr, P1, st));
        ctx->launches++;
    }
    // ---- E: dense / residual split per panel -------------------------------------------------
    if (panels) {
        classify_kernel<<<(panels < (uint32_t)sm * 8 ? panels : (uint32_t)sm * 8), 256, 0, st>>>(
            panels, run_start.ptr, order, counts.ptr, threshold, n_dense.ptr, n_sparse.ptr, n_sparse_data.ptr, n_dense_data.ptr, n_tiles.ptr);
        ctx->launches++;
    }
    // ---- F: scans ----------------------------------------------------------------------------
    {
        size_t tb = 0;
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, n_dense.ptr, d_off.ptr, P1, st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, n_dense.ptr, d_off.ptr, P1, st));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, n_sparse.ptr, s_off.ptr, P1, st));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, n_sparse_data.ptr, sv_off.ptr, P1, st));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, n_tiles.ptr, tile_base.ptr, P1, st));
        ctx->launches += 4;
    }
    BSMR_TRY(d2h(plan->h_dense_col_offsets, d_off.ptr, P1, st));
    BSMR_TRY(d2h(plan->h_sparse_col_offsets, s_off.ptr, P1, st));
    BSMR_TRY(d2h(plan->h_sparse_value_offsets, sv_off.ptr, P1, st));
    std::vector<uint32_t> h_tile_base, h_n_dense_data, h_n_sparse_data;
    BSMR_TRY(d2h(h_tile_base, tile_base.ptr, P1, st));
    BSMR_TRY(d2h(h_n_dense_data, n_dense_data.ptr, P1, st));
    BSMR_TRY(d2h(h_n_sparse_data, n_sparse_data.ptr, P1, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));
    const uint32_t total_dense_cols = plan->h_dense_col_offsets[panels];
    const uint32_t total_sparse_cols = plan->h_sparse_col_offsets[panels];
    const uint32_t total_res = plan->h_sparse_value_offsets[panels];
    const uint32_t total_tiles = h_tile_base[panels];

    // ---- G: column lists ---------------------------------------------------------------------
    TmpBuf<uint32_t> sparse_cols(ws);
    BSMR_TRY(plan->dense_cols.alloc(total_dense_cols));
    BSMR_TRY(sparse_cols.alloc(total_sparse_cols));
    if (total_dense_cols) fill_u32_kernel<<<grid_for(total_dense_cols, kThreads, sm), kThreads, 0, st>>>(plan->dense_cols.ptr, total_dense_cols, N);
    if (total_sparse_cols) fill_u32_kernel<<<grid_for(total_sparse_cols, kThreads, sm), kThreads, 0, st>>>(sparse_cols.ptr, total_sparse_cols, N);
    ctx->launches += 2;
    if (num_runs) {
        BSMR_TRY(sparse_cnt.alloc(static_cast<size_t>(num_runs) + 1));
        BSMR_TRY(res_start.alloc(static_cast<size_t>(num_runs) + 1));
        BSMR_TRY(rank_of_run.alloc(num_runs));
        BSMR_TRY(run_off.alloc(static_cast<size_t>(num_runs) + 1));
        BSMR_CUDA_OK(cudaMemsetAsync(sparse_cnt.ptr, 0, sparse_cnt.bytes(), st));
        scatter_cols_kernel<<<grid_for(num_runs, kThreads, sm), kThreads, 0, st>>>(
            num_runs, key2_sorted, order, ukeys.ptr, counts.ptr, run_start.ptr, n_dense.ptr, d_off.ptr, s_off.ptr,
            plan->dense_cols.ptr, sparse_cols.ptr, sparse_cnt.ptr, rank_of_run.ptr);
        ctx->launches++;
    }
    BSMR_TRY(d2h(plan->h_dense_cols, plan->dense_cols.ptr, total_dense_cols, st));
    BSMR_TRY(d2h(plan->h_sparse_cols, sparse_cols.ptr, total_sparse_cols, st));
    BSMR_CUDA_OK(cudaEventRecord(e1, st));

    // ---- H: device format --------------------------------------------------------------------
    plan->num_res = total_res;
    plan->num_tiles = total_tiles;
    plan->num_dense_blocks = total_dense_cols / kBlockCols;
    BSMR_TRY(plan->res_out.alloc(total_res)); BSMR_TRY(plan->res_col.alloc(total_res));
    BSMR_TRY(plan->res_row.alloc(total_res)); BSMR_TRY(plan->res_rel.alloc(total_res));
    BSMR_TRY(plan->rr_row.alloc(total_res)); BSMR_TRY(plan->rr_col.alloc(total_res)); BSMR_TRY(plan->rr_out.alloc(total_res));
    TmpBuf<uint32_t> res_flag(ws), res_pos(ws);
    BSMR_TRY(res_flag.alloc(static_cast<size_t>(total) + 1));
    BSMR_TRY(res_pos.alloc(static_cast<size_t>(total) + 1));
    BSMR_CUDA_OK(cudaMemsetAsync(res_flag.ptr, 0, res_flag.bytes(), st));
    BSMR_TRY(plan->tile_panel.alloc(total_tiles)); BSMR_TRY(plan->tile_col_begin.alloc(total_tiles));
    BSMR_TRY(plan->tile_ncols.alloc(total_tiles));
    BSMR_TRY(plan->tile_meta.alloc(total_tiles));
    BSMR_TRY(plan->tile_scatter.alloc(static_cast<size_t>(total_tiles) * kPanel * kTileCols));
    if (total_tiles) {
        const uint64_t n = static_cast<uint64_t>(total_tiles) * kPanel * kTileCols;
        fill_u32_kernel<<<grid_for(n, kThreads, sm), kThreads, 0, st>>>(plan->tile_scatter.ptr, n, kNull);
        tile_meta_kernel<<<grid_for(panels, kThreads, sm), kThreads, 0, st>>>(panels, tile_base.ptr, n_dense.ptr, d_off.ptr,
                                                                              plan->tile_panel.ptr, plan->tile_col_begin.ptr, plan->tile_ncols.ptr, plan->tile_meta.ptr);
        ctx->launches += 2;
    }
    if (num_runs) {
        size_t tb = 0;
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, sparse_cnt.ptr, res_start.ptr, static_cast<size_t>(num_runs) + 1, st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, sparse_cnt.ptr, res_start.ptr, static_cast<size_t>(num_runs) + 1, st));
        // run_off: first sorted entry of every run (runs in ascending (panel, column) order = RLE order)
        BSMR_CUDA_OK(cudaMemsetAsync(run_off.ptr + num_runs, 0, sizeof(uint32_t), st));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, counts.ptr, run_off.ptr, static_cast<size_t>(num_runs), st));
        ctx->launches += 2;
        place_entries_kernel<<<grid_for(num_runs, kThreads, sm), kThreads, 0, st>>>(
            num_runs, ukeys.ptr, counts.ptr, run_off.ptr, rank_of_run.ptr, run_start.ptr, n_dense.ptr, tile_base.ptr, res_start.ptr,
            keys_sorted, vals_sorted, plan->reordered_rows.ptr, R, cbits, plan->tile_scatter.ptr, plan->res_out.ptr,
            plan->res_col.ptr, plan->res_row.ptr, plan->res_rel.ptr, start.ptr, plan->row_offsets.ptr, res_flag.ptr);
        ctx->launches++;
        if (total_res) {
            BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, res_flag.ptr, res_pos.ptr, static_cast<size_t>(total) + 1, st));
            BSMR_TRY(ensure_temp(tb));
            BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, res_flag.ptr, res_pos.ptr, static_cast<size_t>(total) + 1, st));
            compact_rows_kernel<<<grid_for((uint64_t)R * 32, kThreads, sm), kThreads, 0, st>>>(
                plan->reordered_rows.ptr, R, plan->row_offsets.ptr, plan->col_indices.ptr, start.ptr, res_flag.ptr, res_pos.ptr,
                plan->rr_row.ptr, plan->rr_col.ptr, plan->rr_out.ptr);
            ctx->launches += 2;
        }
    }
    BSMR_TRY(d2h(plan->h_tile_panel, plan->tile_panel.ptr, total_tiles, st));
    BSMR_CUDA_OK(cudaEventRecord(e2, st));
    BSMR_CUDA_OK(cudaEventSynchronize(e2));
    BSMR_CUDA_OK(cudaGetLastError());
    BSMR_CUDA_OK(cudaEventElapsedTime(&plan->col_ms, e0, e1));
    BSMR_CUDA_OK(cudaEventElapsedTime(&plan->format_ms, e1, e2));

    plan->h_panel_nnz_prefix.assign(P1, 0);
    uint64_t dense_total = 0;
    for (uint32_t p = 0; p < panels; ++p) {
        plan->h_panel_nnz_prefix[p + 1] = plan->h_panel_nnz_prefix[p] + h_n_dense_data[p] + h_n_sparse_data[p];
        dense_total += h_n_dense_data[p];
    }
    plan->num_dense_values = dense_total;
    return BSMR_OK;
}

}  // namespace bsmr
