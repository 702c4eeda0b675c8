// C ABI of libbsmr_b200.so (declared in include/bsmr_b200.h).
// Host-side orchestration only; kernels live in residual.cu / dense_tc.cu / colreorder.cu /
// rowreorder.cu.  Nothing here falls back to the CPU: without a device every call fails.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cub/cub.cuh>

#include "common.cuh"

namespace bsmr {

static thread_local char g_error[512] = "";

// Error code of the tcgen05 kernels' bounded mbarrier waits: one word of mapped pinned host memory, so that the code
// written just before __trap() can still be read after the context has gone into its sticky error state.
static uint32_t* g_flag_host = nullptr;
static uint32_t* g_flag_dev = nullptr;
uint32_t* kernel_error_flag() {
    if (!g_flag_dev) {
        void* h = nullptr;
        if (cudaHostAlloc(&h, sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        g_flag_host = static_cast<uint32_t*>(h);
        *g_flag_host = 0;
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        g_flag_dev = static_cast<uint32_t*>(d);
    }
    return g_flag_dev;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    if (g_flag_host && *g_flag_host != 0 && n >= 0 && n < (int)sizeof(g_error) - 64)
        snprintf(g_error + n, sizeof(g_error) - n, " [kernel wait code %u: a pipeline barrier timed out]", *g_flag_host);
}
const char* get_error() { return g_error; }

}  // namespace bsmr

using namespace bsmr;

extern "C" {

const char* bsmr_version(void) { return "bsmr_b200 0.1 (sm_100a)"; }
const char* bsmr_last_error(void) { return get_error(); }

const char* bsmr_status_string(int status) {
    switch (status) {
        case BSMR_OK: return "ok";
        case BSMR_ERR_INVALID_ARGUMENT: return "invalid argument";
        case BSMR_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (this library has no CPU path)";
        case BSMR_ERR_CUDA: return "CUDA error";
        case BSMR_ERR_OUT_OF_MEMORY: return "out of device memory";
        case BSMR_ERR_BAD_STATE: return "call order violated";
        case BSMR_ERR_UNSUPPORTED: return "unsupported";
    }
    return "unknown status";
}

int bsmr_ctx_create(int device, void* cuda_stream, bsmr_ctx** out) {
    if (!out) {
        set_error("bsmr_ctx_create: out is NULL");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device visible (%s); libbsmr_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return BSMR_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (0..%d)", device, count - 1);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    cudaDeviceProp prop{};
    BSMR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return BSMR_ERR_NO_DEVICE;
    }
    BSMR_CUDA_OK(cudaSetDevice(device));
    bsmr_ctx* ctx = new bsmr_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->device_name = prop.name;
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            set_error("cudaStreamCreate failed");
            return BSMR_ERR_CUDA;
        }
        ctx->owns_stream = true;
    }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->side_stream2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming);
    // cuTensorMapEncodeTiled through the runtime: no link-time dependency on libcuda.so
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
        ctx->encode_tiled = fn;
    } else {
        (void)cudaGetLastError();
    }
    *out = ctx;
    return BSMR_OK;
}

int bsmr_ctx_destroy(bsmr_ctx* ctx) {
    if (!ctx) return BSMR_OK;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm) bsmr_ctx_comm_destroy(ctx);       // communicator, its stream and events (comm.cu)
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->side_stream2) cudaStreamDestroy(ctx->side_stream2);
    if (ctx->copy_in_stream) cudaStreamDestroy(ctx->copy_in_stream);
    if (ctx->copy_out_stream) cudaStreamDestroy(ctx->copy_out_stream);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return BSMR_OK;
}

int bsmr_ctx_synchronize(bsmr_ctx* ctx) {
    if (!ctx) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BSMR_OK;
}

int bsmr_ctx_device_name(bsmr_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap == 0) return BSMR_ERR_INVALID_ARGUMENT;
    snprintf(buf, cap, "%s", ctx->device_name.c_str());
    return BSMR_OK;
}

int bsmr_ctx_launch_count(bsmr_ctx* ctx, uint64_t* count) {
    if (!ctx || !count) return BSMR_ERR_INVALID_ARGUMENT;
    *count = ctx->launches;
    return BSMR_OK;
}

int bsmr_calculate_block_size(bsmr_ctx* ctx, uint32_t M, uint32_t N, uint64_t free_mem_bytes, uint32_t* block_size) {
    if (!ctx || !block_size) return BSMR_ERR_INVALID_ARGUMENT;
    if (free_mem_bytes == 0) {
        size_t free_b = 0, total_b = 0;
        BSMR_CUDA_OK(cudaSetDevice(ctx->device));
        BSMR_CUDA_OK(cudaMemGetInfo(&free_b, &total_b));
        free_mem_bytes = free_b;
    }
    // src/rowReordering.cu:1014-1024 -- integer numerators, float divisors, ceil
    const float gm = static_cast<float>(static_cast<uint64_t>(M) * M * 4u) / static_cast<float>(free_mem_bytes / 2);
    const float sm = static_cast<float>(static_cast<uint64_t>(N) * 4u) / static_cast<float>(49152u / 2);
    const uint32_t a = static_cast<uint32_t>(std::ceil(gm));
    const uint32_t b = static_cast<uint32_t>(std::ceil(sm));
    const uint32_t bs = std::max(a, b);
    *block_size = bs > 16 ? bs : 16;
    return BSMR_OK;
}

// ------------------------------------------------------------------------------- plan
}  // extern "C"
namespace {
// The checks the reference's loaders make before a matrix reaches any kernel (src/Matrix.cpp:442-465: entry count,
// bounds, duplicate coordinates), for patterns that arrive through the C ABI instead of a loader.  One warp per row:
// bit 0 = offsets not monotone / beyond nnz, bit 1 = column >= N, bit 2 = a row whose columns are not strictly
// ascending (then duplicates have to be looked for by sorting, below).
__global__ void validate_csr_kernel(uint32_t M, uint32_t N, uint32_t nnz, const uint32_t* __restrict__ ro, const uint32_t* __restrict__ ci,
                                    uint32_t* __restrict__ flags) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t f = 0;
    if (warp == 0 && lane == 0 && ro[0] != 0) f |= 1u;
    for (uint64_t r = warp; r < M; r += stride) {
        const uint32_t b = ro[r], e = ro[r + 1];
        if (b > e || e > nnz) { f |= 1u; continue; }
        for (uint32_t k = b + lane; k < e; k += 32) {
            const uint32_t c = ci[k];
            if (c >= N) f |= 2u;
            if (k > b && ci[k - 1] >= c) f |= 4u;
        }
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if (lane == 0 && f) atomicOr(flags, f);
}
__global__ void row_col_keys_kernel(uint32_t M, const uint32_t* __restrict__ ro, const uint32_t* __restrict__ ci, uint64_t* __restrict__ keys) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = warp; r < M; r += stride)
        for (uint32_t k = ro[r] + lane; k < ro[r + 1]; k += 32) keys[k] = (r << 32) | ci[k];
}
__global__ void adjacent_equal_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ flags) {
    uint32_t f = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (keys[i] == keys[i - 1]) f = 8u;
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}
int validate_pattern(bsmr_plan* p) {
    bsmr_ctx* ctx = p->ctx;
    DevBuf<uint32_t> d_flags;
    BSMR_TRY(d_flags.alloc(1));
    BSMR_CUDA_OK(cudaMemsetAsync(d_flags.ptr, 0, 4, ctx->stream));
    const int grid = ctx->sm_count * 8;
    validate_csr_kernel<<<grid, 256, 0, ctx->stream>>>(p->M, p->N, p->nnz, p->row_offsets.ptr, p->col_indices.ptr, d_flags.ptr);
    ctx->launches++;
    uint32_t f = 0;
    BSMR_CUDA_OK(cudaMemcpyAsync(&f, d_flags.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if ((f & 3u) == 0 && (f & 4u) && p->nnz > 1) {
        // some row is not in ascending column order (the .mtx loader keeps file order: src/Matrix.cpp:467-470), so a
        // repeated coordinate need not be adjacent: sort (row, column) keys and compare neighbours
        DevBuf<uint64_t> ka, kb;
        DevBuf<uint8_t> tmp;
        BSMR_TRY(ka.alloc(p->nnz));
        BSMR_TRY(kb.alloc(p->nnz));
        row_col_keys_kernel<<<grid, 256, 0, ctx->stream>>>(p->M, p->row_offsets.ptr, p->col_indices.ptr, ka.ptr);
        cub::DoubleBuffer<uint64_t> dk(ka.ptr, kb.ptr);
        size_t tb = 0;
        int end_bit = 33;
        while (end_bit < 64 && ((uint64_t)p->M >> (end_bit - 32)) != 0) ++end_bit;
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, static_cast<int64_t>(p->nnz), 0, end_bit, ctx->stream));
        BSMR_TRY(tmp.alloc(tb + 256));
        BSMR_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp.ptr, tb, dk, static_cast<int64_t>(p->nnz), 0, end_bit, ctx->stream));
        adjacent_equal_kernel<<<grid, 256, 0, ctx->stream>>>(dk.Current(), p->nnz, d_flags.ptr);
        ctx->launches += 3;
        BSMR_CUDA_OK(cudaMemcpyAsync(&f, d_flags.ptr, 4, cudaMemcpyDeviceToHost, ctx->stream));
        BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    BSMR_CUDA_OK(cudaGetLastError());
    if (f & 1u) { set_error("bsmr_plan_create: row_offsets is not a non-decreasing sequence from 0 to nnz"); return BSMR_ERR_INVALID_ARGUMENT; }
    if (f & 2u) { set_error("bsmr_plan_create: a column index is >= N = %u (the reference's loaders reject this: src/Matrix.cpp:452)", p->N); return BSMR_ERR_INVALID_ARGUMENT; }
    if (f & 8u) { set_error("bsmr_plan_create: a (row, column) coordinate appears twice (the reference's loaders reject this: src/Matrix.cpp:457)"); return BSMR_ERR_INVALID_ARGUMENT; }
    return BSMR_OK;
}
}  // namespace
extern "C" {

int bsmr_plan_create(bsmr_ctx* ctx, uint32_t M, uint32_t N, uint32_t nnz, const uint32_t* row_offsets,
                     const uint32_t* col_indices, int on_device, bsmr_plan** out) {
    if (!ctx || !out || !row_offsets || (nnz && !col_indices)) {
        set_error("bsmr_plan_create: NULL argument");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    bsmr_plan* p = new bsmr_plan();
    p->ctx = ctx;
    p->M = M;
    p->N = N;
    p->nnz = nnz;
    int s = p->row_offsets.alloc(static_cast<size_t>(M) + 1);
    if (s == BSMR_OK) s = p->col_indices.alloc(nnz);
    if (s != BSMR_OK) {
        delete p;
        return s;
    }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    cudaError_t e = cudaMemcpyAsync(p->row_offsets.ptr, row_offsets, p->row_offsets.bytes(), kind, ctx->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(p->col_indices.ptr, col_indices, p->col_indices.bytes(), kind, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("bsmr_plan_create: copy of the CSR pattern failed: %s", cudaGetErrorString(e));
        delete p;
        return BSMR_ERR_CUDA;
    }
    // consistency check the reference performs in its loaders (src/Matrix.cpp:442-465): offsets must end at nnz
    uint32_t last = 0;
    e = cudaMemcpy(&last, p->row_offsets.ptr + M, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || last != nnz) {
        set_error("bsmr_plan_create: row_offsets[M] = %u but nnz = %u", last, nnz);
        delete p;
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    // ... and the rest of the loaders' checks: offsets monotone, columns in range, no coordinate twice.  Anything else
    // would reach the kernels as out-of-bounds reads of B or as a (panel, column) run longer than 16.
    {
        const int vs = validate_pattern(p);
        if (vs != BSMR_OK) {
            delete p;
            return vs;
        }
    }
    *out = p;
    return BSMR_OK;
}

static int queue_copy_out(bsmr_plan* plan, bsmr_plan::HostSlot& s);

int bsmr_plan_destroy(bsmr_plan* plan) {
    if (!plan) return BSMR_OK;
    cudaSetDevice(plan->ctx->device);
    for (bsmr_plan::HostSlot& s : plan->host_slots)       // pipelined host-data calls still in flight use the plan's buffers
        if (s.in_flight) {
            queue_copy_out(plan, s);
            cudaEventSynchronize(s.d2h_done);
        }
    delete plan;
    return BSMR_OK;
}

static void reset_shard(bsmr_plan* p) {
    p->sharded = false;
    p->shard_rank = 0;
    p->shard_world = 1;
    p->h_shard_bounds.assign({0u, p->num_row_panels});
    p->shard_first_panel = 0;
    p->shard_end_panel = p->num_row_panels;
    p->shard_res_begin = 0;
    p->shard_res_end = p->num_res;
    p->shard_tile_begin = 0;
    p->shard_tile_end = p->num_tiles;
    p->shard_wt_begin = 0;
    p->shard_wt_end = p->num_wide_tiles;
    p->shard_tile2_begin = 0;
    p->shard_tile2_end = p->num_tiles2;
    p->shard_res2_begin = 0;
    p->shard_res2_end = p->num_res2;
}

int bsmr_plan_set_tile_work(bsmr_plan* plan, float nnz_equivalents_per_tile) {
    if (!plan || !(nnz_equivalents_per_tile >= 0.f)) return BSMR_ERR_INVALID_ARGUMENT;
    plan->tile_work = nnz_equivalents_per_tile;
    return BSMR_OK;
}

int bsmr_plan_set_wide_epilogue(bsmr_plan* plan, int form) {
    if (!plan || form < BSMR_WIDE_EPILOGUE_AUTO || form > BSMR_WIDE_EPILOGUE_MASK) return BSMR_ERR_INVALID_ARGUMENT;
    plan->wide_epilogue_form = form;
    return BSMR_OK;
}

int bsmr_ctx_set_host_copy_duplex(bsmr_ctx* ctx, int mode) {
    if (!ctx || mode < -1 || mode > 1) return BSMR_ERR_INVALID_ARGUMENT;
    ctx->duplex = mode;
    return BSMR_OK;
}

int bsmr_plan_set_wide_ratio(bsmr_plan* plan, float ratio) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    plan->wide_ratio = ratio;
    return BSMR_OK;
}

int bsmr_plan_set_l2_policy(bsmr_plan* plan, uint32_t hot_budget_mb, uint32_t min_b_mb, uint32_t cold_first) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    plan->l2_hot_budget_mb = hot_budget_mb;
    plan->l2_hot_min_b_mb = min_b_mb;
    plan->l2_cold_first = cold_first ? 1u : 0u;
    plan->auto_flags.clear();               // the per-K execution plan was measured under the old policy
    return BSMR_OK;
}

int bsmr_plan_row_reorder(bsmr_plan* plan, float alpha, uint32_t block_size, uint32_t flags) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    plan->have_cols = plan->have_format = plan->have_flat = false;
    BSMR_TRY(row_reorder(plan, alpha, block_size, flags));
    plan->have_rows = true;
    return BSMR_OK;
}

int bsmr_plan_set_row_order(bsmr_plan* plan, const uint32_t* reordered_rows, uint32_t count) {
    if (!plan || (count && !reordered_rows)) return BSMR_ERR_INVALID_ARGUMENT;
    if (count > plan->M) {
        set_error("bsmr_plan_set_row_order: %u rows given but the matrix has %u", count, plan->M);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    {
        std::vector<uint8_t> seen(plan->M, 0);
        for (uint32_t i = 0; i < count; ++i) {
            if (reordered_rows[i] >= plan->M) {
                set_error("bsmr_plan_set_row_order: row %u out of range", reordered_rows[i]);
                return BSMR_ERR_INVALID_ARGUMENT;
            }
            if (seen[reordered_rows[i]]++) {
                set_error("bsmr_plan_set_row_order: row %u is listed twice", reordered_rows[i]);
                return BSMR_ERR_INVALID_ARGUMENT;
            }
        }
    }
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    plan->h_reordered_rows.assign(reordered_rows, reordered_rows + count);
    BSMR_TRY(plan->reordered_rows.alloc(count));
    if (count) {
        BSMR_CUDA_OK(cudaMemcpyAsync(plan->reordered_rows.ptr, reordered_rows, count * sizeof(uint32_t),
                                     cudaMemcpyHostToDevice, plan->ctx->stream));
        BSMR_CUDA_OK(cudaStreamSynchronize(plan->ctx->stream));
    }
    // numRowPanels_ = ceil(size / ROW_PANEL_SIZE)   (src/BSMR.cpp:57)
    plan->num_row_panels = (count + kPanel - 1) / kPanel;
    plan->have_rows = true;
    plan->have_cols = plan->have_format = plan->have_flat = false;
    return BSMR_OK;
}

int bsmr_plan_col_reorder(bsmr_plan* plan, float delta) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    if (!plan->have_rows) {
        set_error("bsmr_plan_col_reorder: no row order yet (call bsmr_plan_row_reorder or bsmr_plan_set_row_order)");
        return BSMR_ERR_BAD_STATE;
    }
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    BSMR_TRY(col_reorder_and_format(plan, delta));
    plan->have_cols = plan->have_format = true;
    plan->auto_flags.clear();
    reset_shard(plan);
    return BSMR_OK;
}

int bsmr_plan_reorder(bsmr_plan* plan, float alpha, float delta, uint32_t block_size, uint32_t flags) {
    BSMR_TRY(bsmr_plan_row_reorder(plan, alpha, block_size, flags));
    return bsmr_plan_col_reorder(plan, delta);
}

// ---- reorder cache ---------------------------------------------------
// The row order is the expensive part of BSMR (the clustering chain: 20 ms on nips, seconds on 100 k rows) and it
// depends only on the sparsity pattern, alpha, the block size and the reduction mode.  The reference recomputes it on
// every run (src/sddmm.cu:10-39; sddmm_testMode only reuses it across deltas, :70-89).  Here it can be written to a
// file keyed by a fingerprint of the pattern and read back into a plan of the same pattern; the column reorder and the
// device format (about a millisecond) are then rebuilt from it.
}  // extern "C"
namespace {
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// order-independent fingerprint: sum over i of mix(seed, i, a[i])
__global__ void fingerprint_kernel(const uint32_t* __restrict__ a, uint64_t n, unsigned long long seed, unsigned long long* __restrict__ out) {
    unsigned long long h = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        h += mix64(seed + i * 0x9E3779B97F4A7C15ull + a[i]);
    for (int w = 16; w >= 1; w >>= 1) h += __shfl_xor_sync(0xffffffffu, h, w);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}
struct ReorderFileHeader {
    char magic[8];                 // "BSMRRO01"
    uint32_t M, N, nnz, block_size;
    uint64_t fingerprint;
    float alpha;
    uint32_t flags;
    int32_t num_clusters, num_clusters_true;
    uint32_t num_rows, reserved;
};
int pattern_fingerprint(bsmr_plan* p, uint64_t* out) {
    bsmr_ctx* ctx = p->ctx;
    DevBuf<unsigned long long> acc;
    BSMR_TRY(acc.alloc(1));
    BSMR_CUDA_OK(cudaMemsetAsync(acc.ptr, 0, 8, ctx->stream));
    const int grid = ctx->sm_count * 4;
    fingerprint_kernel<<<grid, 256, 0, ctx->stream>>>(p->row_offsets.ptr, static_cast<uint64_t>(p->M) + 1, 0x1234567ull, acc.ptr);
    if (p->nnz) fingerprint_kernel<<<grid, 256, 0, ctx->stream>>>(p->col_indices.ptr, p->nnz, 0x89abcdefull, acc.ptr);
    ctx->launches += 2;
    unsigned long long h = 0;
    BSMR_CUDA_OK(cudaMemcpyAsync(&h, acc.ptr, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    *out = h ^ (static_cast<uint64_t>(p->M) << 40) ^ (static_cast<uint64_t>(p->N) << 20) ^ p->nnz;
    return BSMR_OK;
}
}  // namespace
extern "C" {

int bsmr_plan_fingerprint(bsmr_plan* plan, uint64_t* fingerprint) {
    if (!plan || !fingerprint) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    return pattern_fingerprint(plan, fingerprint);
}

int bsmr_plan_save_row_order(bsmr_plan* plan, const char* path, float alpha, uint32_t flags) {
    if (!plan || !path) return BSMR_ERR_INVALID_ARGUMENT;
    if (!plan->have_rows) {
        set_error("bsmr_plan_save_row_order: the plan has no row order yet");
        return BSMR_ERR_BAD_STATE;
    }
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    ReorderFileHeader h{};
    std::memcpy(h.magic, "BSMRRO01", 8);
    h.M = plan->M; h.N = plan->N; h.nnz = plan->nnz; h.block_size = plan->block_size;
    BSMR_TRY(pattern_fingerprint(plan, &h.fingerprint));
    h.alpha = alpha; h.flags = flags;
    h.num_clusters = plan->num_clusters; h.num_clusters_true = plan->num_clusters_true;
    h.num_rows = static_cast<uint32_t>(plan->h_reordered_rows.size());
    FILE* f = std::fopen(path, "wb");
    if (!f) {
        set_error("bsmr_plan_save_row_order: cannot open %s for writing", path);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 &&
                    (h.num_rows == 0 || std::fwrite(plan->h_reordered_rows.data(), 4, h.num_rows, f) == h.num_rows);
    if (std::fclose(f) != 0 || !ok) {
        set_error("bsmr_plan_save_row_order: short write to %s", path);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    return BSMR_OK;
}

int bsmr_plan_load_row_order(bsmr_plan* plan, const char* path, float alpha, uint32_t flags) {
    if (!plan || !path) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        set_error("bsmr_plan_load_row_order: cannot open %s", path);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    ReorderFileHeader h{};
    std::vector<uint32_t> rows;
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "BSMRRO01", 8) == 0 && h.num_rows <= plan->M;
    if (ok) {
        rows.resize(h.num_rows);
        ok = h.num_rows == 0 || std::fread(rows.data(), 4, h.num_rows, f) == h.num_rows;
    }
    std::fclose(f);
    if (!ok) {
        set_error("bsmr_plan_load_row_order: %s is not a row-order file of this library (or is truncated)", path);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    uint64_t fp = 0;
    BSMR_TRY(pattern_fingerprint(plan, &fp));
    if (h.M != plan->M || h.N != plan->N || h.nnz != plan->nnz || h.fingerprint != fp) {
        set_error("bsmr_plan_load_row_order: %s was computed for another sparsity pattern (%u x %u, nnz %u, fingerprint %016llx; "
                  "this plan: %u x %u, nnz %u, %016llx)", path, h.M, h.N, h.nnz, (unsigned long long)h.fingerprint, plan->M, plan->N,
                  plan->nnz, (unsigned long long)fp);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (h.alpha != alpha || h.flags != flags) {
        set_error("bsmr_plan_load_row_order: %s holds the order for alpha = %g, flags = %u, not alpha = %g, flags = %u", path, h.alpha,
                  h.flags, alpha, flags);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    BSMR_TRY(bsmr_plan_set_row_order(plan, rows.data(), h.num_rows));
    plan->block_size = h.block_size;
    plan->num_clusters = h.num_clusters;
    plan->num_clusters_true = h.num_clusters_true;
    return BSMR_OK;
}

// ---- accessors -----------------------------------------------------------------------
static int rphm_reference_layout(bsmr_plan* p, int which, std::vector<uint32_t>& out);

static int fetch_vector(bsmr_plan* p, int which, std::vector<uint32_t>& tmp, const std::vector<uint32_t>** ref) {
    *ref = nullptr;
    switch (which) {
        case BSMR_VEC_REORDERED_ROWS:
            if (!p->have_rows) break;
            *ref = &p->h_reordered_rows;
            return BSMR_OK;
        case BSMR_VEC_DISPERSIONS: *ref = &p->h_dispersions; return BSMR_OK;
        case BSMR_VEC_CLUSTER_IDS: *ref = &p->h_cluster_ids; return BSMR_OK;
        case BSMR_VEC_GROUP_WIDE:
            if (!p->have_format) break;
            tmp.assign(p->h_group_wide.begin(), p->h_group_wide.end());
            *ref = &tmp;
            return BSMR_OK;
        case BSMR_VEC_DENSE_COLS: if (!p->have_cols) break; *ref = &p->h_dense_cols; return BSMR_OK;
        case BSMR_VEC_DENSE_COL_OFFSETS: if (!p->have_cols) break; *ref = &p->h_dense_col_offsets; return BSMR_OK;
        case BSMR_VEC_SPARSE_COLS: if (!p->have_cols) break; *ref = &p->h_sparse_cols; return BSMR_OK;
        case BSMR_VEC_SPARSE_COL_OFFSETS: if (!p->have_cols) break; *ref = &p->h_sparse_col_offsets; return BSMR_OK;
        case BSMR_VEC_SPARSE_VALUE_OFFSETS: if (!p->have_cols) break; *ref = &p->h_sparse_value_offsets; return BSMR_OK;
        case BSMR_VEC_BLOCK_OFFSETS:
        case BSMR_VEC_BLOCK_VALUES:
        case BSMR_VEC_SPARSE_VALUES:
        case BSMR_VEC_SPARSE_RELATIVE_ROWS:
        case BSMR_VEC_SPARSE_COL_INDICES:
            if (!p->have_format) break;
            BSMR_TRY(rphm_reference_layout(p, which, tmp));
            *ref = &tmp;
            return BSMR_OK;
        default:
            set_error("unknown vector id %d", which);
            return BSMR_ERR_INVALID_ARGUMENT;
    }
    set_error("vector %d is not available yet (reorder not run)", which);
    return BSMR_ERR_BAD_STATE;
}

int bsmr_plan_vector_size(bsmr_plan* plan, int which, uint64_t* size) {
    if (!plan || !size) return BSMR_ERR_INVALID_ARGUMENT;
    std::vector<uint32_t> tmp;
    const std::vector<uint32_t>* v = nullptr;
    // sizes of the on-demand RPHM vectors are known without materialising them
    if (plan->have_format) {
        switch (which) {
            case BSMR_VEC_BLOCK_OFFSETS: *size = static_cast<uint64_t>(plan->num_row_panels) + 1; return BSMR_OK;
            case BSMR_VEC_BLOCK_VALUES: *size = static_cast<uint64_t>(plan->num_dense_blocks) * 256u; return BSMR_OK;
            case BSMR_VEC_SPARSE_VALUES:
            case BSMR_VEC_SPARSE_RELATIVE_ROWS:
            case BSMR_VEC_SPARSE_COL_INDICES: *size = plan->num_res; return BSMR_OK;
            default: break;
        }
    }
    BSMR_TRY(fetch_vector(plan, which, tmp, &v));
    *size = v->size();
    return BSMR_OK;
}

int bsmr_plan_vector_copy(bsmr_plan* plan, int which, uint32_t* host_out, uint64_t capacity) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    std::vector<uint32_t> tmp;
    const std::vector<uint32_t>* v = nullptr;
    BSMR_TRY(fetch_vector(plan, which, tmp, &v));
    if (capacity < v->size()) {
        set_error("bsmr_plan_vector_copy: capacity %llu < size %zu", (unsigned long long)capacity, v->size());
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!v->empty()) {
        if (!host_out) return BSMR_ERR_INVALID_ARGUMENT;
        std::memcpy(host_out, v->data(), v->size() * sizeof(uint32_t));
    }
    return BSMR_OK;
}

// The RPHM accessors in the reference's own layout (src/BSMR.cpp:125-219), derived from our
// device format: blockValues[(blockOffsets[p]+cb)*256 + r*16 + c] and the residual triplets.
static int rphm_reference_layout(bsmr_plan* p, int which, std::vector<uint32_t>& out) {
    bsmr_ctx* ctx = p->ctx;
    const uint32_t panels = p->num_row_panels;
    if (which == BSMR_VEC_BLOCK_OFFSETS) {
        out.assign(static_cast<size_t>(panels) + 1, 0);
        for (uint32_t q = 0; q < panels; ++q)
            out[q + 1] = out[q] + (p->h_dense_col_offsets[q + 1] - p->h_dense_col_offsets[q] + kBlockCols - 1) / kBlockCols;
        return BSMR_OK;
    }
    if (which == BSMR_VEC_BLOCK_VALUES) {
        std::vector<uint32_t> scatter(static_cast<size_t>(p->num_tiles) * kPanel * kTileCols);
        std::vector<uint32_t> tcb(p->num_tiles), tnc(p->num_tiles);
        if (p->num_tiles) {
            BSMR_CUDA_OK(cudaMemcpyAsync(scatter.data(), p->tile_scatter.ptr, scatter.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            BSMR_CUDA_OK(cudaMemcpyAsync(tcb.data(), p->tile_col_begin.ptr, tcb.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            BSMR_CUDA_OK(cudaMemcpyAsync(tnc.data(), p->tile_ncols.ptr, tnc.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        }
        out.assign(static_cast<size_t>(p->num_dense_blocks) * 256u, kNull);
        // dense_cols offsets are multiples of 16, so global block id = column offset / 16
        for (uint32_t t = 0; t < p->num_tiles; ++t) {
            for (uint32_t c = 0; c < tnc[t]; ++c) {
                const size_t blk = (static_cast<size_t>(tcb[t]) + c) / kBlockCols;
                for (uint32_t r = 0; r < kPanel; ++r) {
                    out[blk * 256u + r * kBlockCols + (c % kBlockCols)] =
                        scatter[(static_cast<size_t>(t) * kPanel + r) * kTileCols + c];
                }
            }
        }
        return BSMR_OK;
    }
    out.assign(p->num_res, 0);
    if (p->num_res == 0) return BSMR_OK;
    if (which == BSMR_VEC_SPARSE_VALUES) {
        BSMR_CUDA_OK(cudaMemcpyAsync(out.data(), p->res_out.ptr, p->num_res * 4, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (which == BSMR_VEC_SPARSE_COL_INDICES) {
        BSMR_CUDA_OK(cudaMemcpyAsync(out.data(), p->res_col.ptr, p->num_res * 4, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        std::vector<uint8_t> rel(p->num_res);
        BSMR_CUDA_OK(cudaMemcpyAsync(rel.data(), p->res_rel.ptr, p->num_res, cudaMemcpyDeviceToHost, ctx->stream));
        BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < rel.size(); ++i) out[i] = rel[i];
        return BSMR_OK;
    }
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return BSMR_OK;
}

int bsmr_plan_get_info(bsmr_plan* plan, bsmr_plan_info* info) {
    if (!plan || !info) return BSMR_ERR_INVALID_ARGUMENT;
    std::memset(info, 0, sizeof(*info));
    info->M = plan->M;
    info->N = plan->N;
    info->nnz = plan->nnz;
    info->num_row_panels = plan->num_row_panels;
    info->num_clusters = plan->num_clusters;
    info->num_clusters_true = plan->num_clusters_true;
    info->block_size = plan->block_size;
    info->num_dense_blocks = plan->num_dense_blocks;
    info->num_dense_tiles = plan->num_tiles;
    info->num_dense_values = plan->num_dense_values;
    info->num_sparse_values = plan->num_res;
    info->row_reordering_ms = plan->row_ms;
    info->col_reordering_ms = plan->col_ms;
    info->format_build_ms = plan->format_ms;
    info->cluster_kernel_ms = plan->cluster_ms;
    info->num_row_groups = plan->num_groups;
    info->num_wide_groups = plan->num_wide_groups;
    info->num_wide_tiles = plan->num_wide_tiles;
    const bool wide = plan->num_wide_tiles != 0;
    info->num_block_tiles = wide ? plan->num_tiles2 : plan->num_tiles;
    info->num_wide_values = plan->num_wide_values;
    info->num_block_values = wide ? plan->num_block_values2 : plan->num_dense_values;
    info->num_residual_values = wide ? plan->num_res2 : plan->num_res;
    info->wide_format_ms = plan->wide_ms;
    return BSMR_OK;
}

}  // extern "C"
namespace bsmr {
// The index ranges of the kernels' work lists that belong to the reordered row panels [b, e): residual entries, dense
// tiles, and -- group aligned -- wide tiles.  (The wide kernel's CTA partition is NOT recomputed here: bsmr_plan_set_shard
// does that; a caller that narrows the range temporarily, like the chunked sharded pass in comm.cu, does so only on plans
// without wide tiles.)
void apply_panel_range(bsmr_plan* plan, uint32_t b, uint32_t e) {
    const uint32_t panels = plan->num_row_panels;
    const uint32_t ppg = BSMR_WIDE_GROUP_ROWS / kPanel;
    plan->shard_first_panel = b;
    plan->shard_end_panel = e;
    plan->shard_res_begin = plan->h_sparse_value_offsets.empty() ? 0 : plan->h_sparse_value_offsets[b];
    plan->shard_res_end = plan->h_sparse_value_offsets.empty() ? 0 : plan->h_sparse_value_offsets[e];
    // tiles are ordered by panel
    const std::vector<uint32_t>& tp = plan->h_tile_panel;
    plan->shard_tile_begin = static_cast<uint32_t>(std::lower_bound(tp.begin(), tp.end(), b) - tp.begin());
    plan->shard_tile_end = static_cast<uint32_t>(std::lower_bound(tp.begin(), tp.end(), e) - tp.begin());
    if (plan->num_wide_tiles) {
        const uint32_t gb = (b + ppg - 1) / ppg, ge = e >= panels ? plan->num_groups : e / ppg;   // b, e are group aligned here
        plan->shard_wt_begin = plan->h_wt_group_off[gb];
        plan->shard_wt_end = plan->h_wt_group_off[ge];
        plan->shard_res2_begin = plan->h_rr2_group_off[gb];
        plan->shard_res2_end = plan->h_rr2_group_off[ge];
        const std::vector<uint32_t>& tp2 = plan->h_tile2_panel;
        plan->shard_tile2_begin = static_cast<uint32_t>(std::lower_bound(tp2.begin(), tp2.end(), b) - tp2.begin());
        plan->shard_tile2_end = static_cast<uint32_t>(std::lower_bound(tp2.begin(), tp2.end(), e) - tp2.begin());
    } else {
        plan->shard_wt_begin = plan->shard_wt_end = 0;
    }
}
}  // namespace bsmr
extern "C" {

int bsmr_plan_set_shard(bsmr_plan* plan, uint32_t rank, uint32_t world, uint32_t* first_panel, uint32_t* end_panel,
                        uint64_t* shard_nnz) {
    if (!plan || world == 0 || rank >= world) {
        set_error("bsmr_plan_set_shard: bad rank/world %u/%u", rank, world);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!plan->have_format) {
        set_error("bsmr_plan_set_shard: reorder first");
        return BSMR_ERR_BAD_STATE;
    }
    // contiguous ranges of reordered row panels, boundaries where the work prefix crosses rank * total / world.
    // Work of a panel = its nnz, plus, inside a wide row group, its share of the group's tiles at `tile_work`
    // nnz-equivalents per tile: the wide kernel's time follows the tile count far more than the nnz.  tile_work is a
    // property of the plan: the default (3000) was fitted by hand on an 8-way sharded stack of nips blocks (clustering
    // puts the dense rows together, so row groups differ 5x in nnz: balanced on nnz alone the ranks took 20 to 33 us per
    // step); bsmr_plan_fit_tile_work measures it for this plan and K from the kernels' own times.
    const std::vector<uint64_t>& nnz_pre = plan->h_panel_nnz_prefix;  // size panels + 1
    const uint32_t panels = plan->num_row_panels;
    const uint32_t ppg = BSMR_WIDE_GROUP_ROWS / kPanel;
    std::vector<uint64_t> work_pre;
    if (plan->num_wide_tiles && !nnz_pre.empty()) {
        const double tile_work = plan->tile_work;
        work_pre.assign(nnz_pre.size(), 0);
        for (uint32_t q = 0; q < panels; ++q) {
            const uint32_t g = q / ppg;
            const uint64_t tiles = g + 1 < plan->h_wt_group_off.size() && plan->h_group_wide[g] ? plan->h_wt_group_off[g + 1] - plan->h_wt_group_off[g] : 0;
            const uint32_t in_group = std::min(ppg, panels - g * ppg);
            work_pre[q + 1] = work_pre[q] + (nnz_pre[q + 1] - nnz_pre[q]) + static_cast<uint64_t>(tile_work * static_cast<double>(tiles) / in_group);
        }
    }
    const std::vector<uint64_t>& pre = work_pre.empty() ? nnz_pre : work_pre;
    const uint64_t total = pre.empty() ? 0 : pre[panels];
    auto boundary = [&](uint32_t r) -> uint32_t {
        if (r == 0) return 0;
        if (r >= world) return panels;
        const uint64_t target = total / world * r + (total % world) * r / world;
        uint32_t q = static_cast<uint32_t>(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
        if (q > panels) q = panels;
        if (plan->num_wide_tiles) {
            // a wide row group (16 panels) is one unit of work: move the boundary to the nearest group boundary
            const uint32_t snapped = (q + ppg / 2) / ppg * ppg;
            q = snapped > panels ? panels : snapped;
        }
        return q;
    };
    // every rank computes every boundary (the sharded data plane needs the other ranks' ranges: comm.cu)
    plan->h_shard_bounds.assign(static_cast<size_t>(world) + 1, 0);
    for (uint32_t r = 0; r <= world; ++r) plan->h_shard_bounds[r] = boundary(r);
    for (uint32_t r = 1; r <= world; ++r) plan->h_shard_bounds[r] = std::max(plan->h_shard_bounds[r], plan->h_shard_bounds[r - 1]);
    uint32_t b = plan->h_shard_bounds[rank], e = plan->h_shard_bounds[rank + 1];
    plan->shard_rank = rank;
    plan->shard_world = world;
    if (e < b) e = b;
    plan->sharded = world > 1;
    plan->auto_flags.clear();   // the execution plan is chosen again for the shard
    apply_panel_range(plan, b, e);
    if (plan->num_wide_tiles) {
        BSMR_TRY(wide_partition(plan, plan->shard_wt_begin, plan->shard_wt_end));
    }
    if (first_panel) *first_panel = b;
    if (end_panel) *end_panel = e;
    if (shard_nnz) *shard_nnz = nnz_pre.empty() ? 0 : nnz_pre[e] - nnz_pre[b];
    return BSMR_OK;
}

// ------------------------------------------------------------------------------ SDDMM
static int ensure_identity_rows(bsmr_plan* p) {
    if (p->csr_row_of_nnz.ptr || p->nnz == 0) return BSMR_OK;
    BSMR_TRY(p->csr_row_of_nnz.alloc(p->nnz));
    return launch_expand_rows(p->ctx, p->M, p->nnz, p->row_offsets.ptr, p->csr_row_of_nnz.ptr);
}

// The whole pattern as ONE row-sorted list in reordered-row order: entry e = (A row, B column, CSR position), rows in
// the order of reorderedRows, a row's entries in CSR order.  Panel q owns [h_panel_nnz_prefix[q], h_panel_nnz_prefix[q+1]),
// so a shard (a range of reordered row panels) is one contiguous range of it.  Built on first use; it is what the
// residual kernel walks when an operand has no tensor-core path (K % 4 != 0, unaligned A / B, fp16 B), and the
// order in which the sharded data plane (comm.cu) packs a rank's slice of P.
}  // extern "C"
namespace {
__global__ void reordered_row_len_kernel(uint32_t n, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ ro, uint32_t* __restrict__ len) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = rows[i];
        len[i] = ro[r + 1] - ro[r];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}
__global__ void flat_expand_kernel(uint32_t n, const uint32_t* __restrict__ rows, const uint32_t* __restrict__ ro, const uint32_t* __restrict__ ci,
                                   const uint32_t* __restrict__ off, uint32_t* __restrict__ frow, uint32_t* __restrict__ fcol,
                                   uint32_t* __restrict__ fout) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < n; i += stride) {
        const uint32_t r = rows[i], b = ro[r], e = ro[r + 1], o = off[i];
        for (uint32_t k = b + lane; k < e; k += 32) {
            frow[o + (k - b)] = r;
            fcol[o + (k - b)] = ci[k];
            fout[o + (k - b)] = k;
        }
    }
}
}  // namespace
namespace bsmr {
int ensure_flat_list(bsmr_plan* p) {
    if (p->have_flat || p->nnz == 0) return BSMR_OK;
    bsmr_ctx* ctx = p->ctx;
    const uint32_t n = static_cast<uint32_t>(p->h_reordered_rows.size());
    DevBuf<uint32_t> len, off;
    DevBuf<uint8_t> tmp;
    BSMR_TRY(len.alloc(static_cast<size_t>(n) + 1));
    BSMR_TRY(off.alloc(static_cast<size_t>(n) + 1));
    BSMR_TRY(p->flat_row.alloc(p->nnz));
    BSMR_TRY(p->flat_col.alloc(p->nnz));
    BSMR_TRY(p->flat_out.alloc(p->nnz));
    const int grid = ctx->sm_count * 8;
    reordered_row_len_kernel<<<grid, 256, 0, ctx->stream>>>(n, p->reordered_rows.ptr, p->row_offsets.ptr, len.ptr);
    size_t tb = 0;
    BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, len.ptr, off.ptr, static_cast<size_t>(n) + 1, ctx->stream));
    BSMR_TRY(tmp.alloc(tb + 256));
    BSMR_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp.ptr, tb, len.ptr, off.ptr, static_cast<size_t>(n) + 1, ctx->stream));
    flat_expand_kernel<<<grid, 256, 0, ctx->stream>>>(n, p->reordered_rows.ptr, p->row_offsets.ptr, p->col_indices.ptr, off.ptr,
                                                      p->flat_row.ptr, p->flat_col.ptr, p->flat_out.ptr);
    ctx->launches += 3;
    BSMR_CUDA_OK(cudaGetLastError());
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));      // the scratch above goes out of scope
    p->have_flat = true;
    return BSMR_OK;
}
}  // namespace bsmr
extern "C" {

// One SDDMM pass over the (sharded) plan for `batch` (A, B, P) triples on the same pattern (batch element b at
// A + b*M*K, B + b*N*K, P + b*nnz: sddmm_gpu_batch's strides, src/sddmmKernel.cu:2764-2848; batch = 1: sddmm_gpu).
// The reference runs its two kernels on separate streams (:2555-2559) and folds the batch into gridDim.z; here up to
// three kernels over disjoint sets of nnz run on three streams of the context, each launched ONCE for the whole batch
// (wide: the persistent CTAs walk their tile range once per element, metadata reused; dense: work item = (element,
// tile); residual: gridDim.y = element).
// b_half: B holds fp16 (bsmr_sddmm_f16b); only the CUDA-core kernel reads that, so every nnz goes through it.
static int run_pass(bsmr_plan* p, uint32_t K, const float* dA, const void* dBv, bool b_half, float* dP, uint32_t flags, uint32_t batch) {
    bsmr_ctx* ctx = p->ctx;
    const float* dB = static_cast<const float*>(dBv);
    // B far larger than L2: the residual kernel asks L2 to keep the hub columns (nullptr otherwise; residual.cu)
    const uint32_t* hot = nullptr;
    uint32_t cold_first = 0;
    BSMR_TRY(hot_columns(p, K, &hot, &cold_first, b_half ? 2 : 4));
    ResidualArgs ra{};
    ra.K = K; ra.A = dA; ra.B = dBv; ra.b_half = b_half; ra.P = dP; ra.col_hot = hot; ra.cold_first = cold_first;
    ra.batch = batch; ra.stride_a = (size_t)p->M * K; ra.stride_b = (size_t)p->N * K; ra.stride_p = p->nnz; ra.stream = ctx->stream;
    if (flags & BSMR_SDDMM_NO_REORDER) {
        // CSR order: A row from the expanded row list, B column = CSR column, P index = position
        ra.row = p->csr_row_of_nnz.ptr; ra.col = p->col_indices.ptr; ra.out = nullptr; ra.begin = 0; ra.end = p->nnz;
        return launch_residual(ctx, ra);
    }
    if ((flags & BSMR_SDDMM_RESIDUAL_ONLY) && p->num_tiles != 0) {
        set_error("BSMR_SDDMM_RESIDUAL_ONLY needs a plan whose column reorder ran with delta > 1 (no dense tiles); "
                  "use BSMR_SDDMM_NO_REORDER for the CSR-order path");
        return BSMR_ERR_BAD_STATE;
    }
    // Three kernels over disjoint sets of nnz: the wide row-group kernel (side stream 2), the dense-block kernel
    // (side stream; launched before the residual kernel so that its CTAs take their shared memory / TMEM slots) and
    // the residual kernel (main stream), joined on the main stream.  Without wide groups -- or when the caller asks
    // for the reference's split, or K does not fit the wide kernel -- the full BSMR lists are used.
    const bool wide = p->num_wide_tiles != 0 && !(flags & BSMR_SDDMM_NO_WIDE) && !b_half && wide_supports(K, dA, dB);
    const uint32_t wt_b = wide ? p->shard_wt_begin : 0, wt_e = wide ? p->shard_wt_end : 0;
    const uint32_t tl_b = wide ? p->shard_tile2_begin : p->shard_tile_begin, tl_e = wide ? p->shard_tile2_end : p->shard_tile_end;
    const uint64_t rs_b = wide ? p->shard_res2_begin : p->shard_res_begin, rs_e = wide ? p->shard_res2_end : p->shard_res_end;
    const uint32_t* tile_list = wide ? p->tile_list2.ptr : nullptr;
    ra.row = wide ? p->rr2_row.ptr : p->rr_row.ptr;
    ra.col = wide ? p->rr2_col.ptr : p->rr_col.ptr;
    ra.out = wide ? p->rr2_out.ptr : p->rr_out.ptr;
    ra.begin = rs_b; ra.end = rs_e;
    const bool do_wide = wt_e > wt_b, do_dense = tl_e > tl_b, do_res = rs_e > rs_b;
    if (b_half || (do_dense && !dense_supports(K, dA, dB))) {
        // no tensor-core path for these operands (fp16 B, K % 4 != 0, unaligned pointers): every nnz of the shard through
        // the CUDA-core kernel, rows in reordered order (the TMA row stride must be a multiple of 16 bytes)
        BSMR_TRY(ensure_flat_list(p));
        ra.row = p->flat_row.ptr; ra.col = p->flat_col.ptr; ra.out = p->flat_out.ptr;
        ra.begin = p->h_panel_nnz_prefix.empty() ? 0 : p->h_panel_nnz_prefix[p->shard_first_panel];
        ra.end = p->h_panel_nnz_prefix.empty() ? 0 : p->h_panel_nnz_prefix[p->shard_end_panel];
        return launch_residual(ctx, ra);
    }
    const int kinds = (int)do_wide + (int)do_dense + (int)do_res;
    if (kinds <= 1) {
        if (do_wide) return launch_wide(p, K, dA, dB, dP, wt_b, wt_e, ctx->stream, batch);
        if (do_dense) return launch_dense(p, K, dA, dB, dP, tl_b, tl_e, tile_list, ctx->stream, batch);
        return launch_residual(ctx, ra);
    }
    BSMR_CUDA_OK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    if (do_wide) {
        BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->side_stream2, ctx->ev_fork, 0));
        BSMR_TRY(launch_wide(p, K, dA, dB, dP, wt_b, wt_e, ctx->side_stream2, batch));
        BSMR_CUDA_OK(cudaEventRecord(ctx->ev_join2, ctx->side_stream2));
    }
    if (do_dense && do_res) {
        BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        BSMR_TRY(launch_dense(p, K, dA, dB, dP, tl_b, tl_e, tile_list, ctx->side_stream, batch));
        BSMR_CUDA_OK(cudaEventRecord(ctx->ev_join, ctx->side_stream));
        BSMR_TRY(launch_residual(ctx, ra));
        BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    } else if (do_dense) {
        BSMR_TRY(launch_dense(p, K, dA, dB, dP, tl_b, tl_e, tile_list, ctx->stream, batch));
    } else if (do_res) {
        BSMR_TRY(launch_residual(ctx, ra));
    }
    if (do_wide) BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));
    return BSMR_OK;
}
static int run_once(bsmr_plan* p, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t flags, uint32_t batch = 1) {
    return run_pass(p, K, dA, dB, false, dP, flags, batch);
}

int bsmr_sddmm(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, int iterations, uint32_t flags,
               float* ms_per_iteration) {
    if (!plan || !dA || !dB || (plan->nnz && !dP) || K == 0) {
        set_error("bsmr_sddmm: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!(flags & BSMR_SDDMM_NO_REORDER) && !plan->have_format) {
        set_error("bsmr_sddmm: the plan has no reorder/format yet (call bsmr_plan_reorder)");
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    if (ms_per_iteration) *ms_per_iteration = 0.f;
    if (plan->nnz == 0) return BSMR_OK;
    // Execution plan of a default call: the three-kernel plan, unless the caller asked for a measured choice for this K
    // (bsmr_plan_autotune) or installed one (bsmr_plan_set_execution_choice).  Nothing is measured, synchronised or
    // written here: the call stays asynchronous and stream-capturable, and gives the same bits on every run and rank.
    if (flags == BSMR_SDDMM_DEFAULT && plan->have_format) {
        auto it = plan->auto_flags.find(K);
        if (it != plan->auto_flags.end()) flags = it->second;
    }
    flags &= ~BSMR_SDDMM_THREE_KERNEL;     // = the default plan, whatever choice is installed
    if (flags & BSMR_SDDMM_NO_REORDER) BSMR_TRY(ensure_identity_rows(plan));
    if (iterations <= 0) iterations = 1;
    if (ms_per_iteration) BSMR_CUDA_OK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int it = 0; it < iterations; ++it) BSMR_TRY(run_once(plan, K, dA, dB, dP, flags));
    if (ms_per_iteration) {
        BSMR_CUDA_OK(cudaEventRecord(ctx->ev1, ctx->stream));
        BSMR_CUDA_OK(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        BSMR_CUDA_OK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        *ms_per_iteration = ms / static_cast<float>(iterations);
    }
    return BSMR_OK;
}

int bsmr_plan_execution_choice(bsmr_plan* plan, uint32_t K, uint32_t* flags) {
    if (!plan || !flags) return BSMR_ERR_INVALID_ARGUMENT;
    auto it = plan->auto_flags.find(K);
    *flags = it == plan->auto_flags.end() ? BSMR_SDDMM_DEFAULT : it->second;
    return BSMR_OK;
}

int bsmr_plan_set_execution_choice(bsmr_plan* plan, uint32_t K, uint32_t flags) {
    if (!plan || K == 0) return BSMR_ERR_INVALID_ARGUMENT;
    if (flags != BSMR_SDDMM_DEFAULT && flags != BSMR_SDDMM_NO_WIDE && flags != BSMR_SDDMM_NO_REORDER && flags != BSMR_SDDMM_THREE_KERNEL) {
        set_error("bsmr_plan_set_execution_choice: flags %u is not one of DEFAULT / THREE_KERNEL / NO_WIDE / NO_REORDER", flags);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if ((flags & BSMR_SDDMM_NO_REORDER) && plan->sharded) {
        set_error("bsmr_plan_set_execution_choice: the CSR-order kernel is not shard-aware");
        return BSMR_ERR_BAD_STATE;
    }
    if (flags == BSMR_SDDMM_DEFAULT || flags == BSMR_SDDMM_THREE_KERNEL) plan->auto_flags.erase(K);
    else plan->auto_flags[K] = flags;
    return BSMR_OK;
}

// Explicit measurement of the execution plans for one K (a tool, not part of the default call: it synchronises, writes
// P four times per candidate and its outcome depends on timing).  Candidates: the three-kernel plan, the BSMR split
// alone and -- unsharded -- the CSR-order residual kernel; one warm-up pass and the best of three timed ones each, on
// private events.  The winner is installed for default calls with this K until the next column reorder / set_shard.
// Numerics differ between the candidates (TF32 tensor-core tiles ~1.5e-4 relative, fp32 residual ~1e-6): callers that
// need identical bits across ranks agree on one choice with bsmr_plan_set_execution_choice.
int bsmr_plan_autotune(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t* chosen_flags) {
    if (!plan || !dA || !dB || (plan->nnz && !dP) || K == 0) {
        set_error("bsmr_plan_autotune: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!plan->have_format) {
        set_error("bsmr_plan_autotune: the plan has no reorder/format yet (call bsmr_plan_reorder)");
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    uint32_t best = BSMR_SDDMM_DEFAULT;
    if (plan->nnz) {
        std::vector<uint32_t> cand{BSMR_SDDMM_DEFAULT};
        if (plan->num_wide_tiles != 0 && wide_supports(K, dA, dB)) cand.push_back(BSMR_SDDMM_NO_WIDE);
        if (!plan->sharded) cand.push_back(BSMR_SDDMM_NO_REORDER);
        cudaEvent_t t0, t1;
        BSMR_CUDA_OK(cudaEventCreate(&t0));
        BSMR_CUDA_OK(cudaEventCreate(&t1));
        struct Guard { cudaEvent_t a, b; ~Guard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{t0, t1};
        float best_ms = 0.f;
        for (uint32_t f : cand) {
            if (f & BSMR_SDDMM_NO_REORDER) BSMR_TRY(ensure_identity_rows(plan));
            float ms = 0.f;
            for (int rep = 0; rep < 4; ++rep) {
                float t = 0.f;
                BSMR_CUDA_OK(cudaEventRecord(t0, ctx->stream));
                BSMR_TRY(run_once(plan, K, dA, dB, dP, f));
                BSMR_CUDA_OK(cudaEventRecord(t1, ctx->stream));
                BSMR_CUDA_OK(cudaEventSynchronize(t1));
                BSMR_CUDA_OK(cudaEventElapsedTime(&t, t0, t1));
                if (rep == 1 || (rep > 1 && t < ms)) ms = t;
            }
            if (f == cand[0] || ms < best_ms) { best_ms = ms; best = f; }
        }
    }
    if (best == BSMR_SDDMM_DEFAULT) plan->auto_flags.erase(K);
    else plan->auto_flags[K] = best;
    if (chosen_flags) *chosen_flags = best;
    return BSMR_OK;
}

int bsmr_sddmm_profile3(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t flags,
                        float* wide_ms, float* dense_ms, float* residual_ms) {
    if (!plan || !dA || !dB || (plan->nnz && !dP) || K == 0) return BSMR_ERR_INVALID_ARGUMENT;
    if (!plan->have_format || (flags & BSMR_SDDMM_NO_REORDER)) {
        set_error("bsmr_sddmm_profile: needs a reordered plan");
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_plan* p = plan;
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    const bool wide = p->num_wide_tiles != 0 && !(flags & BSMR_SDDMM_NO_WIDE) && wide_supports(K, dA, dB);
    const uint32_t tl_b = wide ? p->shard_tile2_begin : p->shard_tile_begin, tl_e = wide ? p->shard_tile2_end : p->shard_tile_end;
    const uint64_t rs_b = wide ? p->shard_res2_begin : p->shard_res_begin, rs_e = wide ? p->shard_res2_end : p->shard_res_end;
    const uint32_t* hot = nullptr;           // the residual kernel's L2 policy, as in run_once (built before the clock starts)
    uint32_t cold_first = 0;
    BSMR_TRY(hot_columns(p, K, &hot, &cold_first));
    cudaEvent_t mid0, mid;
    BSMR_CUDA_OK(cudaEventCreate(&mid0));
    BSMR_CUDA_OK(cudaEventCreate(&mid));
    int s = BSMR_OK;
    cudaEventRecord(ctx->ev0, ctx->stream);
    if (wide) s = launch_wide(p, K, dA, dB, dP, p->shard_wt_begin, p->shard_wt_end, ctx->stream);
    cudaEventRecord(mid0, ctx->stream);
    if (s == BSMR_OK) s = launch_dense(p, K, dA, dB, dP, tl_b, tl_e, wide ? p->tile_list2.ptr : nullptr, ctx->stream);
    cudaEventRecord(mid, ctx->stream);
    if (s == BSMR_OK)
        s = launch_residual(ctx, K, dA, dB, dP, wide ? p->rr2_row.ptr : p->rr_row.ptr, wide ? p->rr2_col.ptr : p->rr_col.ptr,
                            wide ? p->rr2_out.ptr : p->rr_out.ptr, rs_b, rs_e, hot, cold_first);
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(ctx->ev1);
    float w = 0.f, a = 0.f, b = 0.f;
    if (e == cudaSuccess) {
        cudaEventElapsedTime(&w, ctx->ev0, mid0);
        cudaEventElapsedTime(&a, mid0, mid);
        cudaEventElapsedTime(&b, mid, ctx->ev1);
    }
    cudaEventDestroy(mid0);
    cudaEventDestroy(mid);
    if (s == BSMR_OK && e != cudaSuccess) {
        set_error("bsmr_sddmm_profile: %s", cudaGetErrorString(e));
        s = BSMR_ERR_CUDA;
    }
    if (wide_ms) *wide_ms = w;
    if (dense_ms) *dense_ms = a;
    if (residual_ms) *residual_ms = b;
    return s;
}

// Fit of the shard balance's tile weight for this plan and K from the kernels' own times (unsharded plan): one wide
// tile costs wide_ms / #tiles, one nnz outside the wide groups (dense_ms + residual_ms) / nnz there.  Timing-dependent:
// in a multi-rank run one rank measures and every rank installs the same value (bsmr_plan_set_tile_work) BEFORE
// bsmr_plan_set_shard, or the ranks' ranges would not partition the panels.
int bsmr_plan_fit_tile_work(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, float* nnz_equivalents_per_tile) {
    if (!plan || !nnz_equivalents_per_tile) return BSMR_ERR_INVALID_ARGUMENT;
    *nnz_equivalents_per_tile = static_cast<float>(plan->tile_work);
    if (plan->sharded) {
        set_error("bsmr_plan_fit_tile_work: measure on the unsharded plan (before bsmr_plan_set_shard)");
        return BSMR_ERR_BAD_STATE;
    }
    const uint64_t outside = plan->num_block_values2 + plan->num_res2;
    if (plan->num_wide_tiles == 0 || outside == 0) return BSMR_OK;      // nothing to weigh against each other
    float best_w = 0.f, best_o = 0.f;
    for (int rep = 0; rep < 4; ++rep) {
        float w = 0.f, a = 0.f, b = 0.f;
        BSMR_TRY(bsmr_sddmm_profile3(plan, K, dA, dB, dP, BSMR_SDDMM_THREE_KERNEL, &w, &a, &b));
        if (rep == 1 || (rep > 1 && w < best_w)) best_w = w;
        if (rep == 1 || (rep > 1 && a + b < best_o)) best_o = a + b;
    }
    if (best_w > 0.f && best_o > 0.f) {
        const double per_tile = best_w / plan->num_wide_tiles, per_nnz = best_o / static_cast<double>(outside);
        *nnz_equivalents_per_tile = static_cast<float>(per_tile / per_nnz);
    }
    return BSMR_OK;
}

int bsmr_sddmm_profile(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t flags,
                       float* dense_ms, float* residual_ms) {
    float w = 0.f, a = 0.f;
    const int s = bsmr_sddmm_profile3(plan, K, dA, dB, dP, flags, &w, &a, residual_ms);
    if (dense_ms) *dense_ms = w + a;
    return s;
}

int bsmr_sddmm_host(bsmr_plan* plan, uint32_t K, const float* hA, const float* hB, float* hP, int iterations,
                    uint32_t flags, float* ms_per_iteration, float* total_ms) {
    if (!plan || !hA || !hB || (plan->nnz && !hP) || K == 0) {
        set_error("bsmr_sddmm_host: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    const size_t na = static_cast<size_t>(plan->M) * K, nb = static_cast<size_t>(plan->N) * K;
    BSMR_TRY(plan->dA.alloc(na));
    BSMR_TRY(plan->dB.alloc(nb));
    BSMR_TRY(plan->dP.alloc(plan->nnz));
    cudaEvent_t t0, t1;
    BSMR_CUDA_OK(cudaEventCreate(&t0));
    BSMR_CUDA_OK(cudaEventCreate(&t1));
    BSMR_CUDA_OK(cudaEventRecord(t0, ctx->stream));
    BSMR_CUDA_OK(cudaMemcpyAsync(plan->dA.ptr, hA, na * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BSMR_CUDA_OK(cudaMemcpyAsync(plan->dB.ptr, hB, nb * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    // dev::vector<float> matrixP_dev(nnz, 0)   (src/sddmmKernel.cu:2525)
    if (plan->nnz) BSMR_CUDA_OK(cudaMemsetAsync(plan->dP.ptr, 0, plan->dP.bytes(), ctx->stream));
    float ms = 0.f;
    int s = bsmr_sddmm(plan, K, plan->dA.ptr, plan->dB.ptr, plan->dP.ptr, iterations, flags, &ms);
    if (s == BSMR_OK && plan->nnz) {
        cudaError_t e = cudaMemcpyAsync(hP, plan->dP.ptr, plan->dP.bytes(), cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) {
            set_error("D2H copy of P failed: %s", cudaGetErrorString(e));
            s = BSMR_ERR_CUDA;
        }
    }
    cudaEventRecord(t1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(t1);
    float tot = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&tot, t0, t1);
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    if (s == BSMR_OK && e != cudaSuccess) {
        set_error("bsmr_sddmm_host: %s", cudaGetErrorString(e));
        s = BSMR_ERR_CUDA;
    }
    if (ms_per_iteration) *ms_per_iteration = ms;
    if (total_ms) *total_ms = tot;
    return s;
}

// ---- pipelined host-data SDDMM --------------------------------------------------------------------------------
// A call = H2D A,B -> zero P -> kernels -> D2H P, exactly what bsmr_sddmm_host does, but asynchronous: successive calls
// alternate between two slots of device buffers, the kernels run on the context's stream and the copies on a copy
// stream, ordered by the slots' events; the host blocks only in _wait.
// Copy order, half-duplex mode.  The copy-out of call i is queued BEHIND the copy-in of call i + 1 (it is held back until
// the next submit or until somebody waits for it): on one in-order copy stream that gives H2D(i+1) | kernels(i) in
// parallel, then D2H(i), and never two copies in opposite directions at once.  Duplex mode: two copy streams, the
// copy-out queued at once.  Which of the two a host gets is measured (probe_duplex below).
static int probe_duplex(bsmr_ctx* ctx);
static int ensure_host_pipeline(bsmr_plan* plan) {
    bsmr_ctx* ctx = plan->ctx;
    if (!ctx->copy_in_stream) BSMR_CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_in_stream, cudaStreamNonBlocking));
    if (!ctx->copy_out_stream) BSMR_CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_out_stream, cudaStreamNonBlocking));
    for (bsmr_plan::HostSlot& s : plan->host_slots) {
        if (!s.h2d_done) BSMR_CUDA_OK(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        if (!s.compute_done) BSMR_CUDA_OK(cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming));
        if (!s.d2h_done) BSMR_CUDA_OK(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
    }
    return probe_duplex(ctx);
}

}  // extern "C"
namespace {
int copy_async(bsmr_ctx*, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream) {
    if (bytes == 0) return BSMR_OK;
    BSMR_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, kind, stream));
    return BSMR_OK;
}
}  // namespace
extern "C" {

// Whether copies in opposite directions may run at the same time.  On some hosts of this pool a 6.4 MB H2D and a 3 MB D2H
// in flight together drop to ~12 GB/s each (53 GB/s alone), on others they overlap perfectly (140 instead of 200 us per
// nips step).  Decided once per context (= per device) by timing 2 MB each way, back to back and concurrently (about a millisecond);
// bsmr_ctx_set_host_copy_duplex overrides.
static int probe_duplex(bsmr_ctx* ctx) {
    if (ctx->duplex >= 0) return BSMR_OK;
    const size_t bytes = 2u << 20;
    void *h = nullptr, *d = nullptr;
    if (cudaHostAlloc(&h, 2 * bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&d, 2 * bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        if (h) cudaFreeHost(h);
        ctx->duplex = 0;
        return BSMR_OK;
    }
    char* hb = static_cast<char*>(h);
    char* db = static_cast<char*>(d);
    auto run = [&](bool duplex) -> double {
        double best = 1e30;
        for (int rep = 0; rep < 4; ++rep) {
            cudaStreamSynchronize(ctx->copy_in_stream);
            cudaStreamSynchronize(ctx->copy_out_stream);
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < 3; ++i) {
                cudaMemcpyAsync(db, hb, bytes, cudaMemcpyHostToDevice, ctx->copy_in_stream);
                cudaMemcpyAsync(hb + bytes, db + bytes, bytes, cudaMemcpyDeviceToHost, duplex ? ctx->copy_out_stream : ctx->copy_in_stream);
            }
            cudaStreamSynchronize(ctx->copy_in_stream);
            cudaStreamSynchronize(ctx->copy_out_stream);
            const double t = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
            if (rep > 0 && t < best) best = t;
        }
        return best;
    };
    const double t_seq = run(false), t_dup = run(true);
    ctx->duplex = t_dup < 0.8 * t_seq ? 1 : 0;
    cudaFree(d);
    cudaFreeHost(h);
    (void)cudaGetLastError();
    return BSMR_OK;
}
static bool host_pipe_duplex(const bsmr_ctx* ctx) { return ctx->duplex == 1; }

// queue the copy-out of a slot whose kernels have been issued
static int queue_copy_out(bsmr_plan* plan, bsmr_plan::HostSlot& s) {
    if (s.d2h_queued) return BSMR_OK;
    bsmr_ctx* ctx = plan->ctx;
    cudaStream_t cs = host_pipe_duplex(ctx) ? ctx->copy_out_stream : ctx->copy_in_stream;
    BSMR_CUDA_OK(cudaStreamWaitEvent(cs, s.compute_done, 0));
    BSMR_TRY(copy_async(ctx, s.hP, s.dP.ptr, s.dP.bytes(), cudaMemcpyDeviceToHost, cs));
    BSMR_CUDA_OK(cudaEventRecord(s.d2h_done, cs));
    s.d2h_queued = true;
    return BSMR_OK;
}

int bsmr_sddmm_host_submit(bsmr_plan* plan, uint32_t K, const float* hA, const float* hB, float* hP, uint32_t flags,
                           uint64_t* ticket) {
    if (!plan || !hA || !hB || (plan->nnz && !hP) || K == 0) {
        set_error("bsmr_sddmm_host_submit: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!(flags & BSMR_SDDMM_NO_REORDER) && !plan->have_format) {
        set_error("bsmr_sddmm_host_submit: the plan has no reorder/format yet (call bsmr_plan_reorder)");
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    BSMR_TRY(ensure_host_pipeline(plan));
    const uint64_t id = plan->host_submits;
    bsmr_plan::HostSlot& s = plan->host_slots[id % bsmr_plan::kHostSlots];
    bsmr_plan::HostSlot& prev = plan->host_slots[(id + bsmr_plan::kHostSlots - 1) % bsmr_plan::kHostSlots];
    const size_t na = static_cast<size_t>(plan->M) * K, nb = static_cast<size_t>(plan->N) * K;
    if (s.in_flight) BSMR_TRY(queue_copy_out(plan, s));       // (two slots: already queued by the previous submit)
    if (s.in_flight && (na > s.dA.capacity || nb > s.dB.capacity || plan->nnz > s.dP.capacity))
        BSMR_CUDA_OK(cudaEventSynchronize(s.d2h_done));      // the buffers are about to be reallocated
    BSMR_TRY(s.dA.alloc(na));
    BSMR_TRY(s.dB.alloc(nb));
    BSMR_TRY(s.dP.alloc(plan->nnz));
    // copy-in: the slot's previous kernels must be done with dA / dB
    if (s.in_flight) BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->copy_in_stream, s.compute_done, 0));
    BSMR_TRY(copy_async(ctx, s.dA.ptr, hA, na * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_in_stream));
    BSMR_TRY(copy_async(ctx, s.dB.ptr, hB, nb * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_in_stream));
    BSMR_CUDA_OK(cudaEventRecord(s.h2d_done, ctx->copy_in_stream));
    // the copy-out of the previous call goes behind this copy-in
    if (&prev != &s && prev.in_flight) BSMR_TRY(queue_copy_out(plan, prev));
    // kernels: after the copy-in, and after the slot's previous copy-out has read dP
    BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->stream, s.h2d_done, 0));
    if (s.in_flight) BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->stream, s.d2h_done, 0));
    if (plan->nnz) BSMR_CUDA_OK(cudaMemsetAsync(s.dP.ptr, 0, s.dP.bytes(), ctx->stream));   // src/sddmmKernel.cu:2525
    BSMR_TRY(bsmr_sddmm(plan, K, s.dA.ptr, s.dB.ptr, s.dP.ptr, 1, flags, nullptr));
    BSMR_CUDA_OK(cudaEventRecord(s.compute_done, ctx->stream));
    s.hP = hP;
    s.d2h_queued = false;
    s.in_flight = true;
    if (host_pipe_duplex(ctx)) BSMR_TRY(queue_copy_out(plan, s));   // two copy streams: nothing to hold back
    plan->host_submits = id + 1;
    if (ticket) *ticket = id;
    return BSMR_OK;
}

int bsmr_sddmm_host_wait(bsmr_plan* plan, uint64_t ticket) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    if (ticket == BSMR_TICKET_ALL) {
        for (bsmr_plan::HostSlot& s : plan->host_slots)
            if (s.in_flight) {
                BSMR_TRY(queue_copy_out(plan, s));
                BSMR_CUDA_OK(cudaEventSynchronize(s.d2h_done));
            }
        return BSMR_OK;
    }
    if (ticket >= plan->host_submits) {
        set_error("bsmr_sddmm_host_wait: ticket %llu was never issued", (unsigned long long)ticket);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    // a ticket older than the slot's latest call completed before that call's kernels started
    bsmr_plan::HostSlot& s = plan->host_slots[ticket % bsmr_plan::kHostSlots];
    if (s.in_flight) {
        BSMR_TRY(queue_copy_out(plan, s));
        BSMR_CUDA_OK(cudaEventSynchronize(s.d2h_done));
    }
    return BSMR_OK;
}

// sddmm_gpu_batch (include/sddmmKernel.cuh:41-47, src/sddmmKernel.cu:2764-2848): numBatch (A, B, P) triples on one
// pattern, batch b at A + b*M*K, B + b*N*K, P + b*nnz.  The reference folds the batch into gridDim.z of its two
// kernels; here every kernel of the plan is launched ONCE for the whole batch (run_once): the wide kernel's persistent
// CTAs walk their tile range once per element with the tile metadata in place, the dense-block kernel's work items are
// (element, tile) pairs, the residual kernel takes the element from gridDim.y -- one launch + one tail per kernel
// instead of one per element.  The stacked operands are addressed through one tensor map with batch*M / batch*N rows:
// batches beyond the 31-bit TMA coordinate range are cut into sub-batches.
int bsmr_sddmm_batch(bsmr_plan* plan, uint32_t num_batch, uint32_t K, const float* dA, const float* dB, float* dP,
                     uint32_t flags, float* total_ms) {
    if (!plan || !dA || !dB || (plan->nnz && !dP) || K == 0) {
        set_error("bsmr_sddmm_batch: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!(flags & BSMR_SDDMM_NO_REORDER) && !plan->have_format) {
        set_error("bsmr_sddmm_batch: the plan has no reorder/format yet (call bsmr_plan_reorder)");
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    if (total_ms) *total_ms = 0.f;
    if (plan->nnz == 0 || num_batch == 0) return BSMR_OK;
    if (flags == BSMR_SDDMM_DEFAULT && plan->have_format) {
        auto it = plan->auto_flags.find(K);
        if (it != plan->auto_flags.end()) flags = it->second;
    }
    flags &= ~BSMR_SDDMM_THREE_KERNEL;
    if (flags & BSMR_SDDMM_NO_REORDER) BSMR_TRY(ensure_identity_rows(plan));
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (total_ms) {
        BSMR_CUDA_OK(cudaEventCreate(&t0));
        BSMR_CUDA_OK(cudaEventCreate(&t1));
        BSMR_CUDA_OK(cudaEventRecord(t0, ctx->stream));
    }
    const size_t sa = static_cast<size_t>(plan->M) * K, sb = static_cast<size_t>(plan->N) * K;
    const uint64_t big = std::max<uint64_t>(std::max(plan->M, plan->N), 1);
    const uint32_t chunk = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(num_batch, 0x7FFFFFFFull / big)));
    int s = BSMR_OK;
    for (uint32_t b = 0; b < num_batch && s == BSMR_OK; b += chunk)
        s = run_once(plan, K, dA + b * sa, dB + b * sb, dP + static_cast<size_t>(b) * plan->nnz, flags, std::min(chunk, num_batch - b));
    if (total_ms) {
        cudaError_t e = cudaEventRecord(t1, ctx->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(t1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(total_ms, t0, t1);
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
        if (s == BSMR_OK && e != cudaSuccess) {
            set_error("bsmr_sddmm_batch: %s", cudaGetErrorString(e));
            s = BSMR_ERR_CUDA;
        }
    }
    return s;
}

// batchedMatrixTranspose (include/sddmmKernel.cuh:49-51, src/sddmmKernel.cu:2486-2515, 2852-2869): every batch element
// is a height x width row-major matrix, written back as width x height; elements are width*height apart.  (The
// reference takes int strides; 64-bit here.)  This is how a caller turns row-major B matrices [K x N] into the
// column-major operand the SDDMM kernels read.
}  // extern "C"
namespace {
__global__ void __launch_bounds__(256) batched_transpose_kernel(uint32_t width, uint32_t height, uint32_t batches, const float* __restrict__ in,
                                                                 float* __restrict__ out) {
    __shared__ float tile[32][33];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8 threads, 4 rows each
    const uint32_t tw = (width + 31) / 32, th = (height + 31) / 32;
    const uint64_t per = (uint64_t)tw * th, total = per * batches;
    const size_t mat = (size_t)width * height;
    for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const uint32_t b = (uint32_t)(w / per);
        const uint32_t r = (uint32_t)(w - (uint64_t)b * per);
        const uint32_t by = r / tw, bx = r - by * tw;
        const float* src = in + b * mat;
        float* dst = out + b * mat;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t x = bx * 32 + tx, y = by * 32 + ty + 8 * j;
            if (x < width && y < height) tile[ty + 8 * j][tx] = src[(size_t)y * width + x];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t x = by * 32 + tx, y = bx * 32 + ty + 8 * j;        // output: row y (an input column), column x
            if (x < height && y < width) dst[(size_t)y * height + x] = tile[tx][ty + 8 * j];
        }
        __syncthreads();
    }
}
}  // namespace
extern "C" {

int bsmr_batched_transpose(bsmr_ctx* ctx, uint32_t width, uint32_t height, uint32_t num_batches, const float* d_input, float* d_output) {
    if (!ctx || (!d_input || !d_output)) {
        set_error("bsmr_batched_transpose: NULL argument");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (width == 0 || height == 0 || num_batches == 0) return BSMR_OK;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    const uint64_t tiles = (uint64_t)((width + 31) / 32) * ((height + 31) / 32) * num_batches;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    batched_transpose_kernel<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, ctx->stream>>>(width, height, num_batches, d_input, d_output);
    ctx->launches++;
    BSMR_CUDA_OK(cudaGetLastError());
    return BSMR_OK;
}

// ---- fp16 storage of B (the reference sketches half operands in include/TensorCoreConfig.cuh:22-56) ----
int bsmr_convert_f32_to_f16(bsmr_ctx* ctx, const float* d_src, void* d_dst, uint64_t count) {
    if (!ctx || (count && (!d_src || !d_dst))) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    return launch_f32_to_f16(ctx, d_src, d_dst, count, ctx->stream);
}

int bsmr_sddmm_f16b(bsmr_plan* plan, uint32_t K, const float* dA, const void* dB_f16, float* dP, int iterations, uint32_t flags,
                    float* ms_per_iteration) {
    if (!plan || !dA || !dB_f16 || (plan->nnz && !dP) || K == 0) {
        set_error("bsmr_sddmm_f16b: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (flags != BSMR_SDDMM_DEFAULT && flags != BSMR_SDDMM_NO_REORDER) {
        set_error("bsmr_sddmm_f16b: flags must be BSMR_SDDMM_DEFAULT (reordered row order) or BSMR_SDDMM_NO_REORDER (CSR order)");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!(flags & BSMR_SDDMM_NO_REORDER) && !plan->have_format) {
        set_error("bsmr_sddmm_f16b: the plan has no reorder/format yet (call bsmr_plan_reorder)");
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    if (ms_per_iteration) *ms_per_iteration = 0.f;
    if (plan->nnz == 0) return BSMR_OK;
    if (flags & BSMR_SDDMM_NO_REORDER) BSMR_TRY(ensure_identity_rows(plan));
    else BSMR_TRY(ensure_flat_list(plan));
    if (iterations <= 0) iterations = 1;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (ms_per_iteration) {
        BSMR_CUDA_OK(cudaEventCreate(&t0));
        BSMR_CUDA_OK(cudaEventCreate(&t1));
        BSMR_CUDA_OK(cudaEventRecord(t0, ctx->stream));
    }
    int s = BSMR_OK;
    for (int it = 0; it < iterations && s == BSMR_OK; ++it) s = run_pass(plan, K, dA, dB_f16, true, dP, flags, 1);
    if (ms_per_iteration) {
        float ms = 0.f;
        cudaError_t e = cudaEventRecord(t1, ctx->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(t1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, t0, t1);
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
        if (s == BSMR_OK && e != cudaSuccess) {
            set_error("bsmr_sddmm_f16b: %s", cudaGetErrorString(e));
            s = BSMR_ERR_CUDA;
        }
        *ms_per_iteration = ms / static_cast<float>(iterations);
    }
    return s;
}

// The batch with host data: one pipelined host-data call per batch element (copy-in of element b + 1 and copy-out of
// element b - 1 overlap the kernels of element b); returns when every P has landed.
int bsmr_sddmm_host_batch(bsmr_plan* plan, uint32_t num_batch, uint32_t K, const float* hA, const float* hB, float* hP,
                          uint32_t flags, float* total_ms) {
    if (!plan || !hA || !hB || (plan->nnz && !hP) || K == 0) {
        set_error("bsmr_sddmm_host_batch: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    const auto t0 = std::chrono::steady_clock::now();
    const size_t sa = static_cast<size_t>(plan->M) * K, sb = static_cast<size_t>(plan->N) * K;
    for (uint32_t b = 0; b < num_batch; ++b)
        BSMR_TRY(bsmr_sddmm_host_submit(plan, K, hA + b * sa, hB + b * sb, hP + static_cast<size_t>(b) * plan->nnz, flags, nullptr));
    BSMR_TRY(bsmr_sddmm_host_wait(plan, BSMR_TICKET_ALL));
    if (total_ms) *total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return BSMR_OK;
}

}  // extern "C"

namespace bsmr {
// evaluationReordering + calculateNumDenseBlocksAndAverageDensityInOriginalMatrix
// (src/BSMR.cpp:826-930, 955-994).  Host-side statistics for the Logger; float sums are taken in
// the reference's order (panel, then block) so the printed averages agree to the last digit.
int evaluate_reordering(bsmr_plan* p, float delta, bsmr_reorder_stats* st) {
    std::memset(st, 0, sizeof(*st));
    bsmr_ctx* ctx = p->ctx;
    const uint32_t panels = p->num_row_panels;
    // nnz of every dense 16x16 block, in (panel, block) order, from the scatter tables
    std::vector<uint32_t> bv;
    BSMR_TRY(rphm_reference_layout(p, BSMR_VEC_BLOCK_VALUES, bv));
    int num_dense_blocks = 0, dense_tb = 0, sparse_tb = 0;
    float total_density = 0.0f;
    size_t blk = 0;
    for (uint32_t q = 0; q < panels; ++q) {
        const uint32_t nblk = (p->h_dense_col_offsets[q + 1] - p->h_dense_col_offsets[q] + kBlockCols - 1) / kBlockCols;
        dense_tb += static_cast<int>((nblk + 3) / 4);                                   // 4 blocks per reference CTA
        const uint32_t sd = p->h_sparse_value_offsets[q + 1] - p->h_sparse_value_offsets[q];
        sparse_tb += static_cast<int>((sd + 127) / 128);                                // 128 nnz per reference CTA
        for (uint32_t b = 0; b < nblk; ++b, ++blk) {
            uint32_t n = 0;
            for (uint32_t i = 0; i < 256; ++i) n += bv[blk * 256 + i] != kNull;
            if (n > 0) {
                const float density = static_cast<float>(n) / 256.0f;
                total_density += density;
                if (density >= delta) ++num_dense_blocks;
            }
        }
    }
    st->num_dense_blocks = num_dense_blocks;
    st->average_density = total_density / num_dense_blocks > 0 ? total_density / num_dense_blocks : 0.0f;
    st->num_dense_thread_blocks = dense_tb;
    st->num_sparse_thread_blocks = sparse_tb;
    st->num_sparse_data = static_cast<int32_t>(p->num_res);
    st->num_dense_data = static_cast<int32_t>(p->nnz - p->num_res);

    // original matrix: 16 x 16 grid blocks in (row panel, column block) order
    std::vector<uint32_t> ro(static_cast<size_t>(p->M) + 1), ci(p->nnz);
    BSMR_CUDA_OK(cudaMemcpyAsync(ro.data(), p->row_offsets.ptr, ro.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (p->nnz) BSMR_CUDA_OK(cudaMemcpyAsync(ci.data(), p->col_indices.ptr, ci.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    const uint32_t rp = (p->M + kPanel - 1) / kPanel, cb = (p->N + kBlockCols - 1) / kBlockCols;
    std::vector<uint32_t> cnt(cb, 0), touched;
    int orig_blocks = 0;
    float orig_total = 0.0f;
    for (uint32_t q = 0; q < rp; ++q) {
        const uint32_t r0 = q * kPanel, r1 = std::min(r0 + kPanel, p->M);
        touched.clear();
        for (uint32_t r = r0; r < r1; ++r)
            for (uint32_t k = ro[r]; k < ro[r + 1]; ++k) {
                const uint32_t b = ci[k] / kBlockCols;
                if (cnt[b]++ == 0) touched.push_back(b);
            }
        std::sort(touched.begin(), touched.end());
        for (uint32_t b : touched) {
            const uint32_t c0 = b * kBlockCols, c1 = std::min(c0 + kBlockCols, p->N);
            const float block_size = static_cast<float>((r1 - r0) * (c1 - c0));
            const float density = static_cast<float>(cnt[b]) / block_size;
            if (density >= delta) {
                orig_total += density;
                ++orig_blocks;
            }
            cnt[b] = 0;
        }
    }
    st->original_num_dense_blocks = orig_blocks;
    st->original_average_density = orig_blocks > 0 ? orig_total / orig_blocks : 0.0f;
    return BSMR_OK;
}
}  // namespace bsmr

extern "C" {

int bsmr_plan_evaluate(bsmr_plan* plan, float delta, bsmr_reorder_stats* stats) {
    if (!plan || !stats) return BSMR_ERR_INVALID_ARGUMENT;
    if (!plan->have_format) {
        set_error("bsmr_plan_evaluate: reorder first");
        return BSMR_ERR_BAD_STATE;
    }
    BSMR_CUDA_OK(cudaSetDevice(plan->ctx->device));
    return evaluate_reordering(plan, delta, stats);
}

}  // extern "C"
