// Shared internals of libbsmr_b200.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "bsmr_b200.h"

namespace bsmr {

constexpr uint32_t kPanel = BSMR_ROW_PANEL_SIZE;    // rows per row panel
constexpr uint32_t kBlockCols = BSMR_BLOCK_COL_SIZE;  // columns per reference dense block
constexpr uint32_t kTileCols = 128;                   // dense columns per tcgen05 tile (UMMA M)
constexpr uint32_t kNull = BSMR_NULL_VALUE;
// wide kernel epilogue, list form (wide_tc.cu) <-> its work-list builder (colreorder.cu): a sub-block is 32 tile columns x
// kWideSbRows group rows; the staging image [column][row] has a pitch of kWideSbRows + 4 words (conflict-free STS.128)
constexpr uint32_t kWideSbRows = 32;
constexpr uint32_t kWideStagePitchWords = kWideSbRows + 4;


// ---- error plumbing: every ABI function returns a status and records a message ---------
void set_error(const char* fmt, ...);
uint32_t* kernel_error_flag();   // device pointer to a host-mapped word: wait code of a timed-out mbarrier (0 = none)
const char* get_error();

#define BSMR_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::bsmr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                              cudaGetErrorString(_e));                                  \
            return _e == cudaErrorMemoryAllocation ? BSMR_ERR_OUT_OF_MEMORY : BSMR_ERR_CUDA; \
        }                                                                               \
    } while (0)

#define BSMR_TRY(expr)                      \
    do {                                    \
        int _s = (expr);                    \
        if (_s != BSMR_OK) return _s;       \
    } while (0)

// ---- RAII device buffer (replaces the reference's dev::vector, include/devVector.cuh) ---
template <typename T>
struct DevBuf {
    T* ptr = nullptr;
    size_t count = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
    // make room for n elements (grow-only: a smaller request keeps the block); contents undefined
    size_t capacity = 0;
    int alloc(size_t n) {
        if (ptr && n <= capacity) {
            count = n;
            return BSMR_OK;
        }
        release();
        capacity = 0;
        if (n == 0) return BSMR_OK;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T));
        if (e != cudaSuccess) {
            ptr = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
            (void)cudaGetLastError();
            return BSMR_ERR_OUT_OF_MEMORY;
        }
        count = capacity = n;
        return BSMR_OK;
    }
    size_t bytes() const { return count * sizeof(T); }
};

// ---- scratch arena for the reorder passes ---------------------------------------------------
// cudaMalloc / cudaFree cost 0.1-1 ms each; the reorder passes need ~30 temporaries.  The arena is a
// bump allocator over one device block that is reset at the start of a pass; if a pass outgrows it the
// overflow is served by extra blocks and the main block is regrown to the high-water mark at the next
// reset, so that from the second call on a pass performs no device allocation at all.
struct Workspace {
    char* base = nullptr;
    size_t cap = 0, off = 0, high = 0;
    std::vector<void*> overflow;
    ~Workspace() { release(); }
    void release() {
        for (void* p : overflow) cudaFree(p);
        overflow.clear();
        if (base) cudaFree(base);
        base = nullptr;
        cap = off = 0;
    }
    // start of a pass (the stream must be idle with respect to the previous pass' scratch)
    void reset() {
        if (!overflow.empty() || high > cap) {
            for (void* p : overflow) cudaFree(p);
            overflow.clear();
            if (base) cudaFree(base);
            base = nullptr;
            cap = 0;
            const size_t want = high + high / 4 + (1u << 20);
            if (cudaMalloc(reinterpret_cast<void**>(&base), want) == cudaSuccess) cap = want; else { base = nullptr; (void)cudaGetLastError(); }
        }
        off = 0;
        high = 0;
    }
    void* get(size_t bytes) {
        bytes = (bytes + 255) & ~static_cast<size_t>(255);
        if (bytes == 0) bytes = 256;
        high += bytes;
        if (base && off + bytes <= cap) {
            void* p = base + off;
            off += bytes;
            return p;
        }
        void* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            set_error("scratch allocation of %zu bytes failed", bytes);
            return nullptr;
        }
        overflow.push_back(p);
        return p;
    }
};

// DevBuf-shaped handle on arena memory (never freed individually)
template <typename T>
struct TmpBuf {
    T* ptr = nullptr;
    size_t count = 0;
    Workspace* ws;
    explicit TmpBuf(Workspace* w) : ws(w) {}
    int alloc(size_t n) {
        count = n;
        ptr = static_cast<T*>(ws->get(n * sizeof(T)));
        return ptr ? BSMR_OK : BSMR_ERR_OUT_OF_MEMORY;
    }
    size_t bytes() const { return count * sizeof(T); }
};

}  // namespace bsmr

// ---- the two opaque ABI objects ----------------------------------------------------------
struct bsmr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    std::string device_name;
    uint64_t launches = 0;  // kernels of this library launched through this context
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // the dense-block kernel runs on a side stream so that it overlaps the residual kernel
    // (the reference also uses one stream per kernel, src/sddmmKernel.cu:2555-2559)
    cudaStream_t side_stream = nullptr, side_stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    // copy engines' streams of the pipelined host-data path (created on first use)
    cudaStream_t copy_in_stream = nullptr, copy_out_stream = nullptr;
    // tensor-map encoder resolved at runtime (no link-time dependency on libcuda)
    void* encode_tiled = nullptr;
    bsmr::Workspace ws;     // scratch of the reorder passes
    // per-context (= per-device) one-time state: cudaFuncSetAttribute is a property of a function ON a device, and
    // whether the host's copies in opposite directions may overlap belongs to the device's PCIe path
    bool attr_dense = false, attr_wide = false;
    int duplex = -1;        // -1 unknown, 0 half-duplex copy order, 1 two copy streams
    // multi-GPU data plane (csrc/comm.cu): NCCL communicator of the ranks that share a sharded plan
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    cudaStream_t comm_stream = nullptr;          // the gather of P runs here, behind the kernels of the chunk it carries
    cudaEvent_t comm_ev[9] = {};
};

struct bsmr_plan {
    bsmr_ctx* ctx = nullptr;
    uint32_t M = 0, N = 0, nnz = 0;

    // CSR pattern on the device
    bsmr::DevBuf<uint32_t> row_offsets, col_indices;

    // ---- BSMR object (host copies are what the accessors return) ----
    bool have_rows = false, have_cols = false, have_format = false;
    std::vector<uint32_t> h_reordered_rows;
    bsmr::DevBuf<uint32_t> reordered_rows;
    uint32_t num_row_panels = 0;
    int num_clusters = 1, num_clusters_true = 1;
    uint32_t block_size = 0;
    std::vector<uint32_t> h_dispersions, h_cluster_ids;
    float row_ms = 0.f, col_ms = 0.f, format_ms = 0.f, cluster_ms = 0.f;

    std::vector<uint32_t> h_dense_cols, h_dense_col_offsets, h_sparse_cols, h_sparse_col_offsets,
        h_sparse_value_offsets;
    bsmr::DevBuf<uint32_t> dense_cols;  // per panel, multiple of 16, sentinel N possible

    // ---- device format (our RPHM) ----
    // residual entries ordered (panel, residual column order, row in panel): exactly the
    // reference's sparseValues / sparseRelativeRows / sparseColIndices, with the A row made absolute
    bsmr::DevBuf<uint32_t> res_out;   // CSR index (P position)
    bsmr::DevBuf<uint32_t> res_col;   // B column
    bsmr::DevBuf<uint32_t> res_row;   // A row (absolute)
    bsmr::DevBuf<uint8_t> res_rel;    // row inside the panel (for the RPHM accessor)
    uint64_t num_res = 0;
    // the same entries ROW-sorted (reordered row, then CSR position): what the residual kernel walks, so
    // that an A row stays in registers across its entries.  Panel p still owns the contiguous range
    // [sparse_value_offsets[p], sparse_value_offsets[p+1]).
    bsmr::DevBuf<uint32_t> rr_row, rr_col, rr_out;
    // dense tiles: up to 128 dense columns (8 reference blocks) of one panel
    bsmr::DevBuf<uint32_t> tile_panel;     // panel id per tile
    bsmr::DevBuf<uint32_t> tile_col_begin; // offset into dense_cols
    bsmr::DevBuf<uint32_t> tile_ncols;     // 16..128
    bsmr::DevBuf<uint32_t> tile_scatter;   // [tile][16 rows][128 cols] CSR index or NULL
    bsmr::DevBuf<uint4> tile_meta;         // {panel, col_begin, ncols, 0}: what the dense kernel reads per tile
    std::vector<uint32_t> h_tile_panel;
    uint32_t num_tiles = 0;
    uint32_t num_dense_blocks = 0;
    uint64_t num_dense_values = 0;
    std::vector<uint64_t> h_panel_nnz_prefix;  // nnz (dense + residual) before each panel

    // ---- wide row groups (128 reordered rows = 8 panels each; csrc/wide_tc.cu) ----
    float wide_ratio = -1.f;                  // < 0: default / environment
    uint32_t num_groups = 0, num_wide_groups = 0, num_wide_tiles = 0;
    uint64_t num_wide_values = 0;
    float wide_ms = 0.f;
    std::vector<uint8_t> h_group_wide;        // per row group
    std::vector<uint32_t> h_wt_group_off;     // per row group (+1): first wide tile
    bsmr::DevBuf<uint4> wt_meta;              // {group, first column (offset into w_cols), #columns, 0} per wide tile
    bsmr::DevBuf<uint32_t> w_cols;            // distinct columns of the wide groups, ascending inside a group
    bsmr::DevBuf<uint32_t> w_mask;            // [tile][8][128]
    bsmr::DevBuf<uint32_t> w_base;            // [tile][2][128]
    int wide_epilogue_form = 0;               // BSMR_WIDE_EPILOGUE_AUTO / _LIST / _MASK (bsmr_plan_set_wide_epilogue)
    bool wide_mask_epilogue = false;          // which form of the epilogue's work the format holds (colreorder.cu: by tile fill)
    bsmr::DevBuf<uint2> w_entries;            // mask form: row-meta pairs [(column quarter * 2 + row half) * #tiles + tile][128 rows]:
                                              //   {mask of the row's nnz among the quarter's 32 columns, CSR position of the first}
                                              // list form: per unit 4 header slots (cumulative sub-block counts), then the entries
                                              //   {byte offset inside the epilogue's staging image, CSR position}, padded to 8 slots
    bsmr::DevBuf<uint32_t> w_sb_off;          // list form: [(column quarter * 2 + row half) * #tiles + tile] (+1): first slot of the unit
    std::vector<uint32_t> h_wt_group;         // row group of every wide tile
    bsmr::DevBuf<uint32_t> w_cta_begin;       // CTA -> first tile, for tiles [w_part_begin, w_part_end) (wide_partition)
    bsmr::DevBuf<uint4> w_cta_rec;            // per CTA two words: {first tile, end tile, group of the first tile, its first column id},
                                              // {the first tile's wt_meta}: everything the kernel needs to issue its first loads
    uint32_t w_part_begin = 0, w_part_end = 0, w_grid = 0;
    // dense-block tiles and residual entries of the groups that are NOT wide (what runs next to the wide kernel)
    bsmr::DevBuf<uint32_t> tile_list2;        // tile ids (ascending)
    std::vector<uint32_t> h_tile2_panel;      // panel of tile_list2[i]
    uint32_t num_tiles2 = 0;
    bsmr::DevBuf<uint32_t> rr2_row, rr2_col, rr2_out;
    uint64_t num_res2 = 0;
    std::vector<uint64_t> h_rr2_group_off;    // per row group (+1): first entry of the group in rr2
    uint64_t num_block_values2 = 0;           // nnz of the dense-block tiles outside the wide groups

    // ---- execution plan chosen per K on the first default SDDMM call (capi.cu: bsmr_sddmm) ----
    std::map<uint32_t, uint32_t> auto_flags;

    // ---- identity ("no reorder") residual list, built lazily ----
    bsmr::DevBuf<uint32_t> csr_row_of_nnz;
    // ---- the whole pattern as one row-sorted list in reordered-row order, built lazily (capi.cu: ensure_flat_list) ----
    bsmr::DevBuf<uint32_t> flat_row, flat_col, flat_out;
    bool have_flat = false;

    // ---- hub columns the residual kernel keeps in L2 when B does not fit (residual.cu: hot_columns) ----
    // Defaults from the sweep in profiles/r01h_l2_policy_sweep.md: the policy pays only when the gather is DRAM-bound
    // (B = 4.3 / 8.6 GB: -5 / -7 %); at B = 537 MB (L2 hit rate 70 % without it) it costs 10 %.
    uint32_t l2_hot_budget_mb = 64;           // bytes of hub-column K-vectors marked evict_last; 0 = policy off
    uint32_t l2_hot_min_b_mb = 2048;          // smaller B: plain kernel
    uint32_t l2_cold_first = 1;               // 1: the other columns carry evict_first instead of the default priority
    bsmr::DevBuf<uint32_t> col_degree, col_hot;
    uint32_t hot_K = 0, hot_budget_mb = 0, hot_threshold = 0, hot_elem = 0;
    uint64_t hot_count = 0;

    // ---- shard (multi-GPU) ----
    uint32_t shard_rank = 0, shard_world = 1;
    std::vector<uint32_t> h_shard_bounds;                 // world + 1 panel boundaries (every rank knows every range)
    double tile_work = 3000.0;                            // nnz-equivalents of one wide tile in the shard balance
    bsmr::DevBuf<float> shard_p, shard_slice;             // comm.cu: the rank's P in CSR positions / its contiguous slice (root: all slices)
    uint32_t shard_first_panel = 0, shard_end_panel = 0;
    uint64_t shard_res_begin = 0, shard_res_end = 0;
    uint32_t shard_tile_begin = 0, shard_tile_end = 0;
    uint32_t shard_wt_begin = 0, shard_wt_end = 0;        // wide tiles
    uint32_t shard_tile2_begin = 0, shard_tile2_end = 0;  // positions in tile_list2
    uint64_t shard_res2_begin = 0, shard_res2_end = 0;    // positions in rr2
    bool sharded = false;

    // scratch for the host-data overload
    bsmr::DevBuf<float> dA, dB, dP;
    // pipelined host-data calls (bsmr_sddmm_host_submit / _wait): two slots of device buffers; the copy-in of call
    // i + 1 (copy-in stream), the kernels of call i (context stream) and the copy-out of call i - 1 (copy-out stream)
    // run concurrently, ordered by the slots' events
    struct HostSlot {
        bsmr::DevBuf<float> dA, dB, dP;
        cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
        bool in_flight = false;
        float* hP = nullptr;          // where the slot's result goes (its copy-out may still be waiting to be queued)
        bool d2h_queued = true;
    };
    static constexpr int kHostSlots = 2;
    HostSlot host_slots[kHostSlots];
    uint64_t host_submits = 0;
    ~bsmr_plan() {
        for (HostSlot& s : host_slots) {
            if (s.h2d_done) cudaEventDestroy(s.h2d_done);
            if (s.compute_done) cudaEventDestroy(s.compute_done);
            if (s.d2h_done) cudaEventDestroy(s.d2h_done);
        }
    }
};

namespace bsmr {

// ---- kernels' host launchers (defined in the .cu files) ---------------------------------
struct ResidualArgs {
    uint32_t K = 0;
    const float* A = nullptr;
    const void* B = nullptr;        // fp32, or fp16 when b_half
    bool b_half = false;
    float* P = nullptr;
    const uint32_t *row = nullptr, *col = nullptr, *out = nullptr;   // out == nullptr: P index = list position
    uint64_t begin = 0, end = 0;
    const uint32_t* col_hot = nullptr;
    uint32_t cold_first = 0;
    uint32_t batch = 1;             // batch elements (gridDim.y); strides in elements of A / B / P
    size_t stride_a = 0, stride_b = 0, stride_p = 0;
    cudaStream_t stream = nullptr;
};
int launch_residual(bsmr_ctx* ctx, const ResidualArgs& args);
int launch_f32_to_f16(bsmr_ctx* ctx, const float* src, void* dst, size_t n, cudaStream_t stream);
int launch_residual(bsmr_ctx* ctx, uint32_t K, const float* dA, const float* dB, float* dP,
                    const uint32_t* res_row, const uint32_t* res_col, const uint32_t* res_out,
                    uint64_t begin, uint64_t end, const uint32_t* col_hot = nullptr, uint32_t cold_first = 0);
int hot_columns(bsmr_plan* plan, uint32_t K, const uint32_t** bitmap, uint32_t* cold_first, uint32_t b_elem_bytes = 4);
int ensure_flat_list(bsmr_plan* plan);
void apply_panel_range(bsmr_plan* plan, uint32_t first_panel, uint32_t end_panel);
int launch_expand_rows(bsmr_ctx* ctx, uint32_t M, uint32_t nnz, const uint32_t* row_offsets, uint32_t* row_of_nnz);

int col_reorder_and_format(bsmr_plan* plan, float delta);
int row_reorder(bsmr_plan* plan, float alpha, uint32_t block_size, uint32_t flags);
// tile_list != nullptr: positions [tile_begin, tile_end) of that list of tile ids; else the tile ids themselves
int launch_dense(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP,
                 uint32_t tile_begin, uint32_t tile_end, const uint32_t* tile_list, cudaStream_t stream, uint32_t batch = 1);
bool dense_supports(uint32_t K, const float* dA, const float* dB);
bool wide_supports(uint32_t K, const float* dA, const float* dB);
int wide_partition(bsmr_plan* plan, uint32_t tile_begin, uint32_t tile_end);
int launch_wide(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP,
                uint32_t tile_begin, uint32_t tile_end, cudaStream_t stream, uint32_t batch = 1);
int evaluate_reordering(bsmr_plan* plan, float delta, bsmr_reorder_stats* stats);

}  // namespace bsmr
