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
#include <cstdlib>

#include "common.cuh"

namespace bsmr {
namespace {

constexpr int kThreads = 256;
constexpr uint32_t kWideRows = BSMR_WIDE_GROUP_ROWS;
constexpr uint32_t kWideCols = BSMR_WIDE_TILE_COLS;
constexpr uint32_t kWQ = kWideRows / 32;                       // 32-row quarters per tile (sub-group x TMEM lane quarter)
constexpr uint32_t kWW = kWideCols / 32;                       // mask words per (tile, row) = 32-column chunks of a tile
constexpr uint32_t kWH = kWideCols >= 128 ? kWideCols / 128 : 1;   // run starts per (tile, row): one per 128 columns

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
            // a (panel, column) run holds at most 16 entries on a valid pattern (bsmr_plan_create rejects repeated
            // coordinates, set_row_order repeated rows); clamped so that a broken invariant cannot reach the panel bits
            key2[u] = (panel << 4) | (16u - min(counts[u], 16u));
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


// ---- wide row-group format (csrc/wide_tc.cu) ------------------------------------------------------------------
// W0: (panel, column) runs -> key (group = panel / 8, column), value = nnz of the run
__global__ void group_keys_kernel(const uint64_t* __restrict__ ukeys, const uint32_t* __restrict__ counts, uint32_t num_runs,
                                  uint64_t* __restrict__ gkeys, uint32_t* __restrict__ gvals) {
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < num_runs; u += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = ukeys[u];
        gkeys[u] = (((k >> 32) / (kWideRows / kPanel)) << 32) | (k & 0xffffffffull);
        gvals[u] = counts[u];
    }
}

struct GroupOf {
    __host__ __device__ uint32_t operator()(uint64_t k) const { return (uint32_t)(k >> 32); }
};

// W1: the distinct columns of every wide group, ascending (one CTA per wide group)
__global__ void gather_wide_cols_kernel(uint32_t num_wide, const uint64_t* __restrict__ dkeys, const uint32_t* __restrict__ seg_begin,
                                        const uint32_t* __restrict__ col_off, const uint32_t* __restrict__ ncols,
                                        uint32_t* __restrict__ w_cols) {
    for (uint32_t wi = blockIdx.x; wi < num_wide; wi += gridDim.x) {
        const uint32_t s0 = seg_begin[wi], d0 = col_off[wi], n = ncols[wi];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) w_cols[d0 + i] = (uint32_t)(dkeys[s0 + i] & 0xffffffffull);
    }
}

// W2: masks and run starts (one warp per row of a wide group).  Also verifies that the row is sorted by column:
// the contiguous-run epilogue of the wide kernel depends on it.
__global__ void fill_wide_kernel(uint32_t num_wide, const uint32_t* __restrict__ wg_group, const uint32_t* __restrict__ col_off,
                                 const uint32_t* __restrict__ ncols, const uint32_t* __restrict__ tile_off,
                                 const uint32_t* __restrict__ rows, uint32_t R, const uint32_t* __restrict__ row_offsets,
                                 const uint32_t* __restrict__ col_indices, const uint32_t* __restrict__ w_cols,
                                 uint32_t* __restrict__ mask, uint32_t* __restrict__ base, uint32_t* __restrict__ unsorted) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w = warp; w < (uint64_t)num_wide * kWideRows; w += stride) {
        const uint32_t wi = (uint32_t)(w / kWideRows), r = (uint32_t)(w % kWideRows);
        const uint64_t i = (uint64_t)wg_group[wi] * kWideRows + r;
        if (i >= R) continue;
        const uint32_t row = rows[i];
        const uint32_t b = row_offsets[row], e = row_offsets[row + 1];
        const uint32_t* cl = w_cols + col_off[wi];
        const uint32_t n = ncols[wi], t0 = tile_off[wi];
        for (uint32_t k = b + lane; k < e; k += 32) {
            const uint32_t c = col_indices[k];
            if (k > b && col_indices[k - 1] >= c) atomicExch(unsorted, 1u);
            uint32_t lo = 0, hi = n;                      // lower_bound: the column is present by construction
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (cl[mid] < c) lo = mid + 1; else hi = mid;
            }
            const uint32_t t = t0 + lo / kWideCols, bit = lo % kWideCols;
            atomicOr(mask + ((size_t)t * kWW + (bit >> 5)) * kWideRows + r, 1u << (bit & 31));
            atomicMin(base + ((size_t)t * kWH + (bit >> 7)) * kWideRows + r, k);
        }
    }
}

// The wide kernel computes the TRANSPOSED tile (lanes = the tile's 128 columns, TMEM columns = the group's 256 rows).  An
// epilogue warp owns 32 tile columns (TMEM lane quarter j) and one half h of the group's rows.  What it needs per row
// of the tile is one pair {mask of the row's nnz among the warp's 32 columns, CSR position of the first of them}: a
// row's entries inside a column quarter are consecutive CSR positions in ascending column order, so lane c stores
// its accumulator to P[base + popc(mask & ((1 << c) - 1))] when bit c is set.  Unit u = (tile, j, h): 128 pairs = 1 KB;
// the units one warp touches while it walks a range of tiles are contiguous: index ((j * 2 + h) * #tiles + tile).
constexpr uint32_t kWideStageRowWords = kWideStagePitchWords;
constexpr uint32_t kWUnitRows = kWideRows / 2;                  // 128
constexpr uint32_t kWUnitsPerTile = kWW * 2;

// W2b: the row-meta pairs; one thread per (tile, column quarter, row)
__global__ void wide_rowmeta_kernel(uint32_t wtiles, const uint32_t* __restrict__ mask, const uint32_t* __restrict__ base,
                                    uint2* __restrict__ meta) {
    const uint64_t total = (uint64_t)wtiles * kWW * kWideRows;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i % kWideRows), j = (uint32_t)(i / kWideRows) % kWW, t = (uint32_t)(i / ((uint64_t)kWideRows * kWW));
        const uint32_t m = mask[((size_t)t * kWW + j) * kWideRows + row];
        uint32_t k = 0;
        if (m) {
            k = base[((size_t)t * kWH + (j >> 2)) * kWideRows + row];
            for (uint32_t jj = j & ~3u; jj < j; ++jj) k += __popc(mask[((size_t)t * kWW + jj) * kWideRows + row]);
        }
        const uint32_t h = row / kWUnitRows;
        meta[(((size_t)(j * 2 + h) * wtiles + t) * kWUnitRows) + row % kWUnitRows] = make_uint2(m, k);
    }
}

// ---- list form of the epilogue's work (sparse tiles) ----
// An epilogue warp owns 32 tile columns (TMEM lane quarter j) and one half h of the group's rows, and handles them as
// kWUSB sub-blocks of 32 columns x 16 rows.  Unit u = (tile * kWW + j) * 2 + h; sub-block sbi = u * kWUSB + s.
// The work lists are laid out by (j, h, tile): everything one epilogue warp consumes while it walks a range of tiles is
// one contiguous stream of 8-byte slots, which it pages through shared memory with bulk copies.  A unit's list is
//   4 slots of header: 8 words, word s < kWUSB = number of entries in sub-blocks 0..s of the unit
//   the entries of sub-block 0, 1, ... in (row, column) order: {byte offset in the staging image, CSR position}
//   padding to a multiple of 8 slots (so that every unit, and every page of the kernel, starts 64-byte aligned)
constexpr uint32_t kWSbRows = kWideSbRows;
constexpr uint32_t kWSbPerQ = 32 / kWSbRows;                    // sub-blocks per 32-row quarter
constexpr uint32_t kWUSB = (kWideRows / 2) / kWSbRows;          // sub-blocks per unit: 8
constexpr uint32_t kWHeaderSlots = 4;
__host__ __device__ inline uint64_t wide_unit_stream_index(uint64_t u, uint32_t wtiles) {
    return (u % kWUnitsPerTile) * wtiles + u / kWUnitsPerTile;
}

// W2b: nnz of every sub-block; one warp per (tile, column quarter j, 32-row quarter rq) = two sub-blocks, lane = row
__global__ void wide_subblock_count_kernel(uint32_t num_q, const uint32_t* __restrict__ mask, uint32_t* __restrict__ sb_cnt) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t q = warp; q < num_q; q += stride) {
        const uint32_t rq = (uint32_t)(q % kWQ), j = (uint32_t)(q / kWQ) % kWW, t = (uint32_t)(q / (kWQ * kWW));
        uint32_t c = __popc(mask[((size_t)t * kWW + j) * kWideRows + rq * 32 + lane]);
#pragma unroll
        for (int w = kWSbRows / 2; w >= 1; w >>= 1) c += __shfl_xor_sync(0xffffffffu, c, w);
        if ((lane & (kWSbRows - 1)) == 0) sb_cnt[q * kWSbPerQ + lane / kWSbRows] = c;   // sub-block ((t * kWW + j) * kWQ + rq) * kWSbPerQ + part
    }
}

// W2b': slots of every unit (header + entries, padded to a multiple of 8), in stream order
__global__ void wide_unit_totals_kernel(uint32_t num_units, uint32_t wtiles, const uint32_t* __restrict__ sb_cnt, uint32_t* __restrict__ u_tot) {
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < num_units; u += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t s = 0;
        for (uint32_t r = 0; r < kWUSB; ++r) s += sb_cnt[u * kWUSB + r];
        u_tot[wide_unit_stream_index(u, wtiles)] = (kWHeaderSlots + s + 7u) & ~7u;
    }
}
// the unit headers: word s = entries in sub-blocks 0..s
__global__ void wide_unit_header_kernel(uint32_t num_units, uint32_t wtiles, const uint32_t* __restrict__ sb_cnt,
                                        const uint32_t* __restrict__ unit_start, uint2* __restrict__ slots) {
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < num_units; u += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t* hdr = reinterpret_cast<uint32_t*>(slots + unit_start[wide_unit_stream_index(u, wtiles)]);
        uint32_t o = 0;
        for (uint32_t r = 0; r < kWUSB; ++r) {
            o += sb_cnt[u * kWUSB + r];
            hdr[r] = o;
        }
    }
}

// W2c: the epilogue's work list: per sub-block its entries in (row, column) order as
// (byte offset of the element inside the epilogue's padded staging image [column][row], CSR position).  Everything
// follows from the masks and the run starts: a row's entries inside a tile are consecutive CSR positions in ascending
// column order.  One warp per (tile, column quarter, 32-row quarter) = two sub-blocks.
__global__ void wide_subblock_fill_kernel(uint32_t num_q, uint32_t wtiles, const uint32_t* __restrict__ mask,
                                          const uint32_t* __restrict__ base, const uint32_t* __restrict__ sb_cnt,
                                          const uint32_t* __restrict__ unit_start, uint2* __restrict__ slots) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t q = warp; q < num_q; q += stride) {
        const uint32_t rq = (uint32_t)(q % kWQ), j = (uint32_t)(q / kWQ) % kWW, t = (uint32_t)(q / (kWQ * kWW));
        const uint32_t row = rq * 32 + lane;
        uint32_t m = mask[((size_t)t * kWW + j) * kWideRows + row];
        const uint32_t cnt = __popc(m);
        uint32_t incl = cnt;                          // inclusive scan over the rows of the lane's sub-block
#pragma unroll
        for (int w = 1; w < (int)kWSbRows; w <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, w, kWSbRows);
            if ((lane & (kWSbRows - 1)) >= (uint32_t)w) incl += o;
        }
        if (cnt == 0) continue;
        uint32_t k = base[((size_t)t * kWH + (j >> 2)) * kWideRows + row];
        for (uint32_t jj = j & ~3u; jj < j; ++jj) k += __popc(mask[((size_t)t * kWW + jj) * kWideRows + row]);
        const uint64_t sbi = q * kWSbPerQ + lane / kWSbRows;     // sub-block of this lane
        const uint64_t u = sbi / kWUSB;
        uint32_t e = unit_start[wide_unit_stream_index(u, wtiles)] + kWHeaderSlots + incl - cnt;
        for (uint64_t sp = u * kWUSB; sp < sbi; ++sp) e += sb_cnt[sp];
        while (m) {
            const uint32_t b = __ffs(m) - 1;
            m &= m - 1;
            slots[e] = make_uint2((b * kWideStageRowWords + (lane & (kWSbRows - 1))) * 4u, k);
            ++e;
            ++k;
        }
    }
}

// W2d: which parts of a wide tile's column list are runs of consecutive columns (meta.w): bit q = the columns of
// 32-column quarter q, bit 4 = all of the tile's columns.  The wide kernel fetches such a part with ONE tiled TMA load
// (32 or 128 rows of the [N x K] tensor) instead of one gather4 per 4 columns.
__global__ void wide_tile_flags_kernel(uint32_t wtiles, uint4* __restrict__ meta, const uint32_t* __restrict__ cols) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t t = warp; t < wtiles; t += stride) {
        const uint4 m = meta[t];
        const uint32_t* c = cols + m.y;
        uint32_t flags = 0;
        bool all = m.z > 0;
        for (uint32_t q = 0; q < kWW; ++q) {
            const uint32_t i = q * 32 + lane;
            const bool ok = i >= m.z || i == 0 || c[i] == c[i - 1] + 1;            // consecutive with its predecessor
            const bool okq = i >= m.z || lane == 0 || c[i] == c[i - 1] + 1;        // ... inside the quarter
            all = all && __all_sync(0xffffffffu, ok);
            if (q * 32 < m.z && __all_sync(0xffffffffu, okq)) flags |= 1u << q;
        }
        if (all) flags |= 16u;
        if (lane == 0) meta[t].w = flags;
    }
}

// W3: first row-sorted residual entry of every row group
__global__ void group_res_offsets_kernel(uint32_t groups, uint32_t R, const uint32_t* __restrict__ start,
                                         const uint32_t* __restrict__ res_pos, uint32_t* __restrict__ out) {
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= groups; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = g * kWideRows < R ? g * kWideRows : R;
        out[g] = res_pos[start[i]];
    }
}

// W4: the residual entries of the groups that stay on the BSMR path, compacted (one CTA per group)
__global__ void copy_group_ranges_kernel(uint32_t groups, const uint8_t* __restrict__ wide, const uint32_t* __restrict__ src_off,
                                         const uint32_t* __restrict__ dst_off, const uint32_t* __restrict__ a0,
                                         const uint32_t* __restrict__ a1, const uint32_t* __restrict__ a2,
                                         uint32_t* __restrict__ b0, uint32_t* __restrict__ b1, uint32_t* __restrict__ b2) {
    for (uint32_t g = blockIdx.x; g < groups; g += gridDim.x) {
        if (wide[g]) continue;
        const uint32_t s = src_off[g], n = src_off[g + 1] - s, d = dst_off[g];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            b0[d + i] = a0[s + i];
            b1[d + i] = a1[s + i];
            b2[d + i] = a2[s + i];
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


namespace {

template <typename T>
int h2d(T* dst, const std::vector<T>& src, cudaStream_t s) {
    if (!src.empty()) BSMR_CUDA_OK(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return BSMR_OK;
}

float wide_ratio_of(const bsmr_plan* plan) {
    if (plan->wide_ratio >= 0.f) return plan->wide_ratio;
    return 5.0f;
}

// Everything the SDDMM kernels need to treat some row groups "wide" (see wide_tc.cu) and the others the BSMR way.
// ukeys/counts: the (panel << 32 | column) runs of step B with their nnz; start: first enumerated nnz of every
// reordered row; res_pos: exclusive scan of the residual flags over that enumeration (nullptr when there is no residual).
int build_wide_format(bsmr_plan* plan, Workspace* ws, const uint64_t* ukeys, const uint32_t* counts, uint32_t num_runs,
                      const uint32_t* start, const uint32_t* res_pos, const std::vector<uint32_t>& h_n_dense_data) {
    bsmr_ctx* ctx = plan->ctx;
    cudaStream_t st = ctx->stream;
    const int sm = ctx->sm_count;
    const uint32_t R = (uint32_t)plan->h_reordered_rows.size();
    const uint32_t G = (R + kWideRows - 1) / kWideRows;
    const uint32_t panels_per_group = kWideRows / kPanel;
    plan->num_groups = G;
    plan->num_wide_groups = plan->num_wide_tiles = 0;
    plan->num_wide_values = 0;
    plan->h_group_wide.assign(G, 0);
    plan->h_wt_group_off.assign(static_cast<size_t>(G) + 1, 0);
    plan->num_tiles2 = plan->num_tiles;
    plan->num_res2 = plan->num_res;
    plan->num_block_values2 = plan->num_dense_values;
    plan->h_tile2_panel.clear();
    plan->h_rr2_group_off.clear();
    const float ratio = wide_ratio_of(plan);
    if (ratio <= 0.f || num_runs == 0 || G == 0 || plan->N == 0) return BSMR_OK;

    TmpBuf<uint8_t> temp(ws);
    auto ensure_temp = [&](size_t bytes) -> int {
        if (bytes > temp.count) return temp.alloc(bytes + bytes / 8 + 256);
        return BSMR_OK;
    };
    // ---- distinct (group, column) pairs with their nnz ------------------------------------------------------
    TmpBuf<uint64_t> gk_a(ws), gk_b(ws), dkeys(ws);
    TmpBuf<uint32_t> gv_a(ws), gv_b(ws), dcnt(ws), num_d_dev(ws), gid(ws), gnnz(ws), gncols(ws), num_g_dev(ws);
    BSMR_TRY(gk_a.alloc(num_runs)); BSMR_TRY(gk_b.alloc(num_runs)); BSMR_TRY(gv_a.alloc(num_runs)); BSMR_TRY(gv_b.alloc(num_runs));
    group_keys_kernel<<<grid_for(num_runs, kThreads, sm), kThreads, 0, st>>>(ukeys, counts, num_runs, gk_a.ptr, gv_a.ptr);
    ctx->launches++;
    const int col_bits = bits_for(plan->N - 1), gbits = bits_for(G - 1);
    size_t tb = 0, tb2 = 0;
    cub::DoubleBuffer<uint64_t> dk(gk_a.ptr, gk_b.ptr);
    cub::DoubleBuffer<uint32_t> dv(gv_a.ptr, gv_b.ptr);
    BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, static_cast<int64_t>(num_runs), 0, col_bits, st));
    BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, static_cast<int64_t>(num_runs), 32, 32 + gbits, st));
    BSMR_TRY(ensure_temp(std::max(tb, tb2)));
    BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb, dk, dv, static_cast<int64_t>(num_runs), 0, col_bits, st));
    BSMR_CUDA_OK(cub::DeviceRadixSort::SortPairs(temp.ptr, tb2, dk, dv, static_cast<int64_t>(num_runs), 32, 32 + gbits, st));
    ctx->launches += 2;
    BSMR_TRY(dkeys.alloc(num_runs)); BSMR_TRY(dcnt.alloc(num_runs)); BSMR_TRY(num_d_dev.alloc(1));
    BSMR_CUDA_OK(cub::DeviceReduce::ReduceByKey(nullptr, tb, dk.Current(), dkeys.ptr, dv.Current(), dcnt.ptr, num_d_dev.ptr, cub::Sum(),
                                                static_cast<int64_t>(num_runs), st));
    BSMR_TRY(ensure_temp(tb));
    BSMR_CUDA_OK(cub::DeviceReduce::ReduceByKey(temp.ptr, tb, dk.Current(), dkeys.ptr, dv.Current(), dcnt.ptr, num_d_dev.ptr, cub::Sum(),
                                                static_cast<int64_t>(num_runs), st));
    ctx->launches++;
    uint32_t num_d = 0;
    BSMR_CUDA_OK(cudaMemcpyAsync(&num_d, num_d_dev.ptr, 4, cudaMemcpyDeviceToHost, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));
    // ---- per group: distinct columns and nnz ----------------------------------------------------------------
    BSMR_TRY(gid.alloc(static_cast<size_t>(G) + 1)); BSMR_TRY(gnnz.alloc(static_cast<size_t>(G) + 1));
    BSMR_TRY(gncols.alloc(static_cast<size_t>(G) + 1)); BSMR_TRY(num_g_dev.alloc(1));
    auto group_it = thrust::make_transform_iterator(static_cast<const uint64_t*>(dkeys.ptr), GroupOf());
    BSMR_CUDA_OK(cub::DeviceReduce::ReduceByKey(nullptr, tb, group_it, gid.ptr, dcnt.ptr, gnnz.ptr, num_g_dev.ptr, cub::Sum(),
                                                static_cast<int64_t>(num_d), st));
    BSMR_CUDA_OK(cub::DeviceRunLengthEncode::Encode(nullptr, tb2, group_it, gid.ptr, gncols.ptr, num_g_dev.ptr, static_cast<int64_t>(num_d), st));
    BSMR_TRY(ensure_temp(std::max(tb, tb2)));
    BSMR_CUDA_OK(cub::DeviceReduce::ReduceByKey(temp.ptr, tb, group_it, gid.ptr, dcnt.ptr, gnnz.ptr, num_g_dev.ptr, cub::Sum(),
                                                static_cast<int64_t>(num_d), st));
    BSMR_CUDA_OK(cub::DeviceRunLengthEncode::Encode(temp.ptr, tb2, group_it, gid.ptr, gncols.ptr, num_g_dev.ptr, static_cast<int64_t>(num_d), st));
    ctx->launches += 2;
    uint32_t num_g = 0;
    BSMR_CUDA_OK(cudaMemcpyAsync(&num_g, num_g_dev.ptr, 4, cudaMemcpyDeviceToHost, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<uint32_t> h_gid, h_gnnz, h_gncols;
    BSMR_TRY(d2h(h_gid, gid.ptr, num_g, st)); BSMR_TRY(d2h(h_gnnz, gnnz.ptr, num_g, st)); BSMR_TRY(d2h(h_gncols, gncols.ptr, num_g, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));

    // ---- which groups go wide: nnz >= ratio * (256 * tiles + 128) ------------------------------------------
    // (a tile moves (256 + its share of the 128 resident rows) K-vectors through L2 whatever its fill, the BSMR kernels
    //  at most one K-vector per nnz)
    std::vector<uint32_t> wg_group, wg_seg, wg_ncols, wg_col_off, wg_tile_off;
    std::vector<uint4> h_meta;
    uint32_t seg = 0, col_total = 0;
    uint64_t wide_values = 0;
    // Second tier: when the groups that pass the bar already hold >= 90 % of the nnz, the few that remain would cost two
    // more kernel launches (dense-block + residual, ~10 us each however little they do) next to the wide kernel, which
    // owns every SM while it runs; they go wide too if they pass 40 % of the bar.
    auto bar = [&](uint32_t i) {
        const uint32_t tiles = (h_gncols[i] + kWideCols - 1) / kWideCols;
        return static_cast<double>(ratio) * (static_cast<double>(kWideCols) * tiles + kWideRows);
    };
    uint64_t nnz_all = 0, nnz_pass = 0;
    for (uint32_t i = 0; i < num_g; ++i) {
        nnz_all += h_gnnz[i];
        if (h_gid[i] < G && static_cast<double>(h_gnnz[i]) >= bar(i)) nnz_pass += h_gnnz[i];
    }
    const double tier = (nnz_all > 0 && static_cast<double>(nnz_pass) >= 0.9 * static_cast<double>(nnz_all)) ? 0.4 : 1.0;
    for (uint32_t i = 0; i < num_g; ++i) {
        const uint32_t g = h_gid[i], nc = h_gncols[i];
        if (g < G && static_cast<double>(h_gnnz[i]) >= tier * bar(i)) {
            plan->h_group_wide[g] = 1;
            col_total = (col_total + 3u) & ~3u;   // the kernel reads a tile's column ids four at a time (16-byte loads)
            wg_group.push_back(g); wg_seg.push_back(seg); wg_ncols.push_back(nc); wg_col_off.push_back(col_total);
            wg_tile_off.push_back(static_cast<uint32_t>(h_meta.size()));
            for (uint32_t c = 0; c < nc; c += kWideCols)
                h_meta.push_back(make_uint4(g, col_total + c, nc - c < kWideCols ? nc - c : kWideCols, 0u));
            col_total += nc;
            wide_values += h_gnnz[i];
        }
        seg += nc;
    }
    const uint32_t num_wide = static_cast<uint32_t>(wg_group.size());
    if (num_wide == 0) return BSMR_OK;
    const uint32_t wtiles = static_cast<uint32_t>(h_meta.size());

    TmpBuf<uint32_t> d_wg_group(ws), d_wg_seg(ws), d_wg_ncols(ws), d_wg_col_off(ws), d_wg_tile_off(ws), d_unsorted(ws);
    BSMR_TRY(d_wg_group.alloc(num_wide)); BSMR_TRY(d_wg_seg.alloc(num_wide)); BSMR_TRY(d_wg_ncols.alloc(num_wide));
    BSMR_TRY(d_wg_col_off.alloc(num_wide)); BSMR_TRY(d_wg_tile_off.alloc(num_wide)); BSMR_TRY(d_unsorted.alloc(1));
    BSMR_TRY(h2d(d_wg_group.ptr, wg_group, st)); BSMR_TRY(h2d(d_wg_seg.ptr, wg_seg, st)); BSMR_TRY(h2d(d_wg_ncols.ptr, wg_ncols, st));
    BSMR_TRY(h2d(d_wg_col_off.ptr, wg_col_off, st)); BSMR_TRY(h2d(d_wg_tile_off.ptr, wg_tile_off, st));
    BSMR_CUDA_OK(cudaMemsetAsync(d_unsorted.ptr, 0, 4, st));
    BSMR_TRY(plan->wt_meta.alloc(wtiles)); BSMR_TRY(plan->w_cols.alloc(static_cast<size_t>(col_total) + 4));
    BSMR_CUDA_OK(cudaMemsetAsync(plan->w_cols.ptr, 0xFF, plan->w_cols.bytes(), st));
    BSMR_TRY(plan->w_mask.alloc(static_cast<size_t>(wtiles) * kWW * kWideRows));
    BSMR_TRY(plan->w_base.alloc(static_cast<size_t>(wtiles) * kWH * kWideRows));
    BSMR_TRY(h2d(plan->wt_meta.ptr, h_meta, st));
    BSMR_CUDA_OK(cudaMemsetAsync(plan->w_mask.ptr, 0, plan->w_mask.bytes(), st));
    BSMR_CUDA_OK(cudaMemsetAsync(plan->w_base.ptr, 0xFF, plan->w_base.bytes(), st));
    gather_wide_cols_kernel<<<(num_wide < (uint32_t)sm * 8 ? num_wide : (uint32_t)sm * 8), 256, 0, st>>>(
        num_wide, dkeys.ptr, d_wg_seg.ptr, d_wg_col_off.ptr, d_wg_ncols.ptr, plan->w_cols.ptr);
    wide_tile_flags_kernel<<<grid_for((uint64_t)wtiles * 32, kThreads, sm), kThreads, 0, st>>>(wtiles, plan->wt_meta.ptr, plan->w_cols.ptr);
    ctx->launches++;
    fill_wide_kernel<<<grid_for((uint64_t)num_wide * kWideRows * 32, kThreads, sm), kThreads, 0, st>>>(
        num_wide, d_wg_group.ptr, d_wg_col_off.ptr, d_wg_ncols.ptr, d_wg_tile_off.ptr, plan->reordered_rows.ptr, R,
        plan->row_offsets.ptr, plan->col_indices.ptr, plan->w_cols.ptr, plan->w_mask.ptr, plan->w_base.ptr, d_unsorted.ptr);
    ctx->launches += 2;
    // Two forms of the epilogue's work (wide_tc.cu): per-entry lists for sparse tiles (instruction count follows the nnz;
    // 2.5 us per 128 x 256 tile at 4 % fill against 3.2 us for the mask form), row masks for dense ones (cost per tile
    // independent of the fill, stores of a row coalesce: 30 % fill runs 1.6x faster than through lists).
    const double fill = static_cast<double>(wide_values) / (static_cast<double>(wtiles) * kWideCols * kWideRows);
    plan->wide_mask_epilogue = plan->wide_epilogue_form ? plan->wide_epilogue_form == BSMR_WIDE_EPILOGUE_MASK : fill >= 0.08;   // bsmr_plan_set_wide_epilogue
    if (plan->wide_mask_epilogue) {
        // row-meta pairs (8 units of 128 pairs per tile)
        BSMR_TRY(plan->w_entries.alloc(static_cast<size_t>(wtiles) * kWUnitsPerTile * kWUnitRows));
        wide_rowmeta_kernel<<<grid_for((uint64_t)wtiles * kWW * kWideRows, kThreads, sm), kThreads, 0, st>>>(
            wtiles, plan->w_mask.ptr, plan->w_base.ptr, plan->w_entries.ptr);
    } else {
        // the epilogue's work lists (64 sub-blocks of 32 columns x 16 rows per tile, 8 units per tile)
        const uint32_t num_q = wtiles * kWQ * kWW;                 // (tile, column quarter, 32-row quarter)
        const uint32_t num_units = wtiles * kWUnitsPerTile;
        TmpBuf<uint32_t> sb_cnt(ws), u_tot(ws);
        BSMR_TRY(sb_cnt.alloc(static_cast<size_t>(num_q) * kWSbPerQ));
        BSMR_TRY(u_tot.alloc(static_cast<size_t>(num_units) + 1));
        // w_sb_off[stream index of the unit] = first slot of the unit's list (+1 entry: the end); a unit takes 4 header
        // slots + its entries, padded to 8
        BSMR_TRY(plan->w_sb_off.alloc(static_cast<size_t>(num_units) + 1));
        const size_t list_cap = static_cast<size_t>(wide_values) + (kWHeaderSlots + 8u) * num_units + 64;
        BSMR_TRY(plan->w_entries.alloc(list_cap));
        BSMR_CUDA_OK(cudaMemsetAsync(plan->w_entries.ptr, 0, plan->w_entries.bytes(), st));
        BSMR_CUDA_OK(cudaMemsetAsync(u_tot.ptr + num_units, 0, 4, st));
        wide_subblock_count_kernel<<<grid_for((uint64_t)num_q * 32, kThreads, sm), kThreads, 0, st>>>(num_q, plan->w_mask.ptr, sb_cnt.ptr);
        wide_unit_totals_kernel<<<grid_for(num_units, kThreads, sm), kThreads, 0, st>>>(num_units, wtiles, sb_cnt.ptr, u_tot.ptr);
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, u_tot.ptr, plan->w_sb_off.ptr, static_cast<size_t>(num_units) + 1, st));
        BSMR_TRY(ensure_temp(tb));
        BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp.ptr, tb, u_tot.ptr, plan->w_sb_off.ptr, static_cast<size_t>(num_units) + 1, st));
        wide_unit_header_kernel<<<grid_for(num_units, kThreads, sm), kThreads, 0, st>>>(num_units, wtiles, sb_cnt.ptr, plan->w_sb_off.ptr, plan->w_entries.ptr);
        wide_subblock_fill_kernel<<<grid_for((uint64_t)num_q * 32, kThreads, sm), kThreads, 0, st>>>(
            num_q, wtiles, plan->w_mask.ptr, plan->w_base.ptr, sb_cnt.ptr, plan->w_sb_off.ptr, plan->w_entries.ptr);
    }
    ctx->launches += 2;
    ctx->launches += 3;
    uint32_t unsorted = 0;
    BSMR_CUDA_OK(cudaMemcpyAsync(&unsorted, d_unsorted.ptr, 4, cudaMemcpyDeviceToHost, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));   // also keeps the host vectors above alive until their copies are done
    if (unsorted) {
        // a CSR row that is not sorted by column breaks the contiguous-run epilogue: stay on the BSMR kernels
        plan->h_group_wide.assign(G, 0);
        return BSMR_OK;
    }

    // ---- the BSMR-path work lists without the wide groups ---------------------------------------------------
    std::vector<uint32_t> h_rr_goff(static_cast<size_t>(G) + 1, 0);
    if (plan->num_res) {
        TmpBuf<uint32_t> d_goff(ws);
        BSMR_TRY(d_goff.alloc(static_cast<size_t>(G) + 1));
        group_res_offsets_kernel<<<grid_for(static_cast<uint64_t>(G) + 1, kThreads, sm), kThreads, 0, st>>>(G, R, start, res_pos, d_goff.ptr);
        ctx->launches++;
        BSMR_TRY(d2h(h_rr_goff, d_goff.ptr, static_cast<size_t>(G) + 1, st));
        BSMR_CUDA_OK(cudaStreamSynchronize(st));
    }
    std::vector<uint32_t> h_dst(static_cast<size_t>(G) + 1, 0);
    plan->h_rr2_group_off.assign(static_cast<size_t>(G) + 1, 0);
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t n = plan->h_group_wide[g] ? 0u : h_rr_goff[g + 1] - h_rr_goff[g];
        h_dst[g + 1] = h_dst[g] + n;
        plan->h_rr2_group_off[g + 1] = h_dst[g + 1];
    }
    plan->num_res2 = h_dst[G];
    BSMR_TRY(plan->rr2_row.alloc(plan->num_res2)); BSMR_TRY(plan->rr2_col.alloc(plan->num_res2)); BSMR_TRY(plan->rr2_out.alloc(plan->num_res2));
    if (plan->num_res2) {
        TmpBuf<uint32_t> d_src(ws), d_dst(ws);
        TmpBuf<uint8_t> d_wide(ws);
        BSMR_TRY(d_src.alloc(static_cast<size_t>(G) + 1)); BSMR_TRY(d_dst.alloc(static_cast<size_t>(G) + 1)); BSMR_TRY(d_wide.alloc(G));
        BSMR_TRY(h2d(d_src.ptr, h_rr_goff, st)); BSMR_TRY(h2d(d_dst.ptr, h_dst, st)); BSMR_TRY(h2d(d_wide.ptr, plan->h_group_wide, st));
        copy_group_ranges_kernel<<<(G < (uint32_t)sm * 8 ? G : (uint32_t)sm * 8), 256, 0, st>>>(
            G, d_wide.ptr, d_src.ptr, d_dst.ptr, plan->rr_row.ptr, plan->rr_col.ptr, plan->rr_out.ptr,
            plan->rr2_row.ptr, plan->rr2_col.ptr, plan->rr2_out.ptr);
        ctx->launches++;
        BSMR_CUDA_OK(cudaStreamSynchronize(st));
    }
    std::vector<uint32_t> h_list;
    uint64_t block_values = 0;
    for (uint32_t t = 0; t < plan->num_tiles; ++t) {
        const uint32_t pnl = plan->h_tile_panel[t];
        if (!plan->h_group_wide[pnl / panels_per_group]) {
            h_list.push_back(t);
            plan->h_tile2_panel.push_back(pnl);
        }
    }
    for (uint32_t q = 0; q < plan->num_row_panels; ++q)
        if (!plan->h_group_wide[q / panels_per_group]) block_values += h_n_dense_data[q];
    plan->num_tiles2 = static_cast<uint32_t>(h_list.size());
    plan->num_block_values2 = block_values;
    BSMR_TRY(plan->tile_list2.alloc(h_list.size()));
    BSMR_TRY(h2d(plan->tile_list2.ptr, h_list, st));
    BSMR_CUDA_OK(cudaStreamSynchronize(st));

    for (uint32_t i = 0, g = 0; g <= G; ++g) {      // first wide tile at or after every group
        while (i < num_wide && wg_group[i] < g) ++i;
        plan->h_wt_group_off[g] = i < num_wide ? wg_tile_off[i] : wtiles;
    }
    plan->num_wide_groups = num_wide;
    plan->num_wide_tiles = wtiles;
    plan->num_wide_values = wide_values;
    plan->h_wt_group.resize(wtiles);
    for (uint32_t t = 0; t < wtiles; ++t) plan->h_wt_group[t] = h_meta[t].x;
    return wide_partition(plan, 0, wtiles);
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
    BSMR_CUDA_OK(cudaStreamSynchronize(st));
    plan->num_dense_values = 0;
    for (uint32_t p = 0; p < panels; ++p) plan->num_dense_values += h_n_dense_data[p];
    // ---- W: wide row groups (our own execution plan on top of the BSMR split) ------------------------------
    cudaEvent_t ew;
    BSMR_CUDA_OK(cudaEventCreate(&ew));
    BSMR_CUDA_OK(cudaEventRecord(ew, st));
    const int ws_status = build_wide_format(plan, ws, ukeys.ptr, counts.ptr, num_runs, start.ptr, total_res ? res_pos.ptr : nullptr, h_n_dense_data);
    if (ws_status != BSMR_OK) { cudaEventDestroy(ew); return ws_status; }
    BSMR_CUDA_OK(cudaEventRecord(e2, st));
    BSMR_CUDA_OK(cudaEventSynchronize(e2));
    cudaEventElapsedTime(&plan->wide_ms, ew, e2);
    cudaEventDestroy(ew);
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
