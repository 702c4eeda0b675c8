// Multi-GPU data plane of the sharded SDDMM: NCCL over NVLink 5 / NVSwitch, one process per GPU.
//
// No counterpart in the reference (single process, device 0 only: include/Logger.hpp:23-25).  The path shards by
// contiguous ranges of REORDERED row panels balanced on work (bsmr_plan_set_shard); output entries are independent, so
// the only exchanges are the two the problem has:
//   before the kernels   B (K x N, read by every rank): every rank uploads a slice of it over its own PCIe link and a grouped
//                        ncclBroadcast per rank (an all-gather-v) replicates it -- world x the host->device rate of a root
//                        upload + broadcast; A follows the panels: a rank uploads only the rows of its shard, and the
//                        slices of B are sized so that A rows + B columns per rank come out equal
//   after the kernels    P: the kernels write CSR positions; a rank's entries are ONE contiguous range of the pattern
//                        in reordered-row order (ensure_flat_list), so a pack kernel makes its slice contiguous, the
//                        slices go to the root with grouped ncclSend / ncclRecv (a gather-v: 4 * nnz bytes in total,
//                        not an all-reduce of nnz-length arrays), and one kernel on the root un-permutes to CSR order
// The row order is global (clustering sees every row) and is the cost centre: it is computed on one rank and broadcast
// (bsmr_plan_bcast_row_order); the column reorder + format build (integer only, deterministic) is recomputed on every
// rank concurrently, which is cheaper than shipping the format.
// NCCL is resolved at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already
// loaded), so libbsmr_b200.so has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace bsmr {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        auto sym = [&](const char* n) { return dlsym(a.handle, n); };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
        a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Broadcast && a.AllGather && a.Send && a.Recv && a.GroupStart &&
               a.GroupEnd && a.GetErrorString;
        return a;
    }();
    return api;
}

#define BSMR_NCCL_OK(expr)                                                                         \
    do {                                                                                           \
        ncclResult_t _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                                   \
            ::bsmr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(_r)); \
            return BSMR_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

int need_nccl() {
    if (!nccl().ok) {
        set_error("NCCL is not available (dlopen of libnccl.so.2 failed: %s)", dlerror() ? dlerror() : "symbols missing");
        return BSMR_ERR_UNSUPPORTED;
    }
    return BSMR_OK;
}
int need_comm(const bsmr_ctx* ctx, const char* who) {
    if (!ctx->nccl_comm) {
        set_error("%s: the context has no communicator (call bsmr_ctx_comm_init on every rank first)", who);
        return BSMR_ERR_BAD_STATE;
    }
    return BSMR_OK;
}
ncclComm_t comm_of(const bsmr_ctx* ctx) { return static_cast<ncclComm_t>(ctx->nccl_comm); }

// rows of A that a shard reads, straight from (mapped, pinned) host memory into their place in the device copy of A:
// one warp per row, 16-byte pieces; the reads are PCIe bursts of a whole K-vector
__global__ void upload_rows_kernel(const float* __restrict__ host_a, float* __restrict__ dev_a, const uint32_t* __restrict__ rows,
                                   uint32_t n, uint32_t K) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t stride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t k4 = K / 4;
    for (uint64_t i = warp; i < n; i += stride) {
        const size_t off = (size_t)rows[i] * K;
        const float4* src = reinterpret_cast<const float4*>(host_a + off);
        float4* dst = reinterpret_cast<float4*>(dev_a + off);
        for (uint32_t k = lane; k < k4; k += 32) dst[k] = src[k];
        for (uint32_t k = k4 * 4 + lane; k < K; k += 32) dev_a[off + k] = host_a[off + k];
    }
}

// slice[e - begin] = P[flat_out[e]]  (a rank's entries made contiguous, in reordered-row order)
__global__ void pack_p_kernel(const float* __restrict__ P, const uint32_t* __restrict__ flat_out, uint64_t begin, uint64_t end,
                              float* __restrict__ slice) {
    for (uint64_t e = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < end; e += (uint64_t)gridDim.x * blockDim.x)
        slice[e - begin] = P[__ldg(flat_out + e)];
}
// P[flat_out[e]] = assembled[e]  (root: back to CSR order)
__global__ void unpack_p_kernel(const float* __restrict__ assembled, const uint32_t* __restrict__ flat_out, uint64_t begin, uint64_t end,
                                float* __restrict__ P) {
    for (uint64_t e = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < end; e += (uint64_t)gridDim.x * blockDim.x)
        P[__ldg(flat_out + e)] = assembled[e];
}

const float* mapped_host_alias(const float* host) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? static_cast<const float*>(a.devicePointer) : nullptr;
}

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

constexpr int kGatherChunks = 4;

// chunk boundaries (panels) of rank r's range: equal nnz, the same on every rank
std::vector<uint32_t> chunk_bounds(const bsmr_plan* plan, int r, int chunks) {
    const std::vector<uint64_t>& pre = plan->h_panel_nnz_prefix;
    const uint32_t b = plan->h_shard_bounds[r], e = plan->h_shard_bounds[r + 1];
    std::vector<uint32_t> cb(static_cast<size_t>(chunks) + 1, b);
    for (int c = 1; c < chunks; ++c) {
        const uint64_t target = pre[b] + (pre[e] - pre[b]) * static_cast<uint64_t>(c) / chunks;
        uint32_t q = static_cast<uint32_t>(std::lower_bound(pre.begin() + b, pre.begin() + e, target) - pre.begin());
        cb[c] = std::max(cb[c - 1], std::min(q, e));
    }
    cb[chunks] = e;
    return cb;
}

// kernels of the shard + pack + gather-v to root + un-permute; P of the whole pattern lands in dP_root (root only).
// With more than one rank (and no wide row groups, whose CTA partition belongs to the whole shard) the shard is cut into
// kGatherChunks chunks of equal nnz: the slice of chunk c is packed and sent on the communication stream while the
// kernels of chunk c + 1 run, so that only the last chunk's transfer is exposed (measured at N = 8 on configs[4]: the
// un-overlapped gather was 2.15 ms of a 6.2 ms step).
int assemble_sharded(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP_root, uint32_t flags, int root,
                     bsmr_shard_times* times, cudaEvent_t* stamps /* 5 events or nullptr */) {
    bsmr_ctx* ctx = plan->ctx;
    const int rank = ctx->comm_rank, world = ctx->comm_world;
    if (!plan->have_format || plan->shard_world != (uint32_t)world || plan->shard_rank != (uint32_t)rank) {
        set_error("sharded SDDMM: call bsmr_plan_set_shard(rank = %d, world = %d) with the communicator's rank / size first", rank, world);
        return BSMR_ERR_BAD_STATE;
    }
    if (rank == root && plan->nnz && !dP_root) {
        set_error("sharded SDDMM: the root needs an output buffer");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    BSMR_TRY(ensure_flat_list(plan));
    const std::vector<uint64_t>& pre = plan->h_panel_nnz_prefix;
    const std::vector<uint32_t>& bounds = plan->h_shard_bounds;     // world + 1 panel boundaries
    const uint32_t pb = bounds[rank], pe = bounds[rank + 1];
    const uint64_t e0 = pre[pb], e1 = pre[pe];
    BSMR_TRY(plan->shard_p.alloc(plan->nnz));                        // this rank's P in CSR positions (only its entries are defined)
    BSMR_TRY(plan->shard_slice.alloc(rank == root ? plan->nnz : (size_t)(e1 - e0)));
    // the root's own slice is packed straight into its place in the assembled array; elsewhere the buffer holds the shard only
    const uint64_t slice_base = rank == root ? 0 : e0;             // entry e lives at shard_slice[e - slice_base]
    const int grid = ctx->sm_count * 8;
    // measured on configs[4]: at 2 ranks the transfer is 8 % of the step and chunking costs more than it hides (15.4 -> 17.6 ms:
    // four launches with four tails, NCCL's copy kernels next to the gather kernel); from 4 ranks on the transfer is a third
    const int chunks = (world >= 4 && plan->num_wide_tiles == 0) ? kGatherChunks : 1;
    NcclApi& n = nccl();
    cudaStream_t cs = ctx->comm_stream;
    if (stamps) BSMR_CUDA_OK(cudaEventRecord(stamps[0], ctx->stream));
    BSMR_CUDA_OK(cudaEventRecord(ctx->comm_ev[8], ctx->stream));
    BSMR_CUDA_OK(cudaStreamWaitEvent(cs, ctx->comm_ev[8], 0));      // the buffers' previous readers (last call's un-permute) are done
    std::vector<std::vector<uint32_t>> cb(world);
    for (int r = 0; r < world; ++r) cb[r] = chunk_bounds(plan, r, chunks);
    struct RangeGuard {       // whatever happens below, the plan keeps the rank's whole shard
        bsmr_plan* plan; uint32_t b, e; bool on;
        ~RangeGuard() { if (on) apply_panel_range(plan, b, e); }
    } range_guard{plan, pb, pe, chunks > 1};
    int status = BSMR_OK;
    for (int c = 0; c < chunks && status == BSMR_OK; ++c) {
        const uint32_t qb = cb[rank][c], qe = cb[rank][c + 1];
        if (chunks > 1) apply_panel_range(plan, qb, qe);
        if (qe > qb) status = bsmr_sddmm(plan, K, dA, dB, plan->shard_p.ptr, 1, flags, nullptr);
        if (status != BSMR_OK) break;
        BSMR_CUDA_OK(cudaEventRecord(ctx->comm_ev[c], ctx->stream));
        if (stamps && c + 1 == chunks) BSMR_CUDA_OK(cudaEventRecord(stamps[1], ctx->stream));
        BSMR_CUDA_OK(cudaStreamWaitEvent(cs, ctx->comm_ev[c], 0));
        const uint64_t s0 = pre[qb], s1 = pre[qe];
        if (s1 > s0) {
            pack_p_kernel<<<grid, 256, 0, cs>>>(plan->shard_p.ptr, plan->flat_out.ptr, s0, s1, plan->shard_slice.ptr + (s0 - slice_base));
            ctx->launches++;
        }
        if (world > 1) {
            ncclResult_t nr = n.GroupStart();
            if (rank == root) {
                for (int r = 0; r < world && nr == ncclSuccess; ++r) {
                    const uint64_t b = pre[cb[r][c]], e = pre[cb[r][c + 1]];
                    if (r != root && e > b) nr = n.Recv(plan->shard_slice.ptr + b, e - b, ncclFloat32, r, comm_of(ctx), cs);
                }
            } else if (s1 > s0) {
                nr = n.Send(plan->shard_slice.ptr + (s0 - slice_base), s1 - s0, ncclFloat32, root, comm_of(ctx), cs);
            }
            const ncclResult_t ne = n.GroupEnd();
            if (nr != ncclSuccess || ne != ncclSuccess) {
                set_error("sharded SDDMM: NCCL send / recv of chunk %d failed: %s", c, n.GetErrorString(nr != ncclSuccess ? nr : ne));
                status = BSMR_ERR_CUDA;
            }
        }
    }
    BSMR_TRY(status);
    BSMR_CUDA_OK(cudaEventRecord(ctx->comm_ev[7], cs));
    BSMR_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[7], 0));      // join: every slice has arrived / left
    if (stamps) {
        BSMR_CUDA_OK(cudaEventRecord(stamps[2], ctx->stream));        // (pack runs on the communication stream: no time of its own here)
        BSMR_CUDA_OK(cudaEventRecord(stamps[3], ctx->stream));
    }
    if (rank == root && plan->nnz) {
        unpack_p_kernel<<<grid, 256, 0, ctx->stream>>>(plan->shard_slice.ptr, plan->flat_out.ptr, 0, plan->nnz, dP_root);
        ctx->launches++;
    }
    if (stamps) BSMR_CUDA_OK(cudaEventRecord(stamps[4], ctx->stream));
    BSMR_CUDA_OK(cudaGetLastError());
    if (times) {
        times->shard_nnz = e1 - e0;
        times->gather_p_bytes = rank == root ? (plan->nnz - (e1 - e0)) * 4ull : (e1 - e0) * 4ull;
    }
    return BSMR_OK;
}

}  // namespace
}  // namespace bsmr

using namespace bsmr;

extern "C" {

int bsmr_comm_unique_id(void* id_out) {
    if (!id_out) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_TRY(need_nccl());
    ncclUniqueId id;
    BSMR_NCCL_OK(nccl().GetUniqueId(&id));
    static_assert(sizeof(id) == BSMR_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id_out, &id, sizeof(id));
    return BSMR_OK;
}

int bsmr_ctx_comm_init(bsmr_ctx* ctx, const void* unique_id, int rank, int world) {
    if (!ctx || !unique_id || world < 1 || rank < 0 || rank >= world) {
        set_error("bsmr_ctx_comm_init: bad arguments (rank %d, world %d)", rank, world);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    BSMR_TRY(need_nccl());
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) BSMR_TRY(bsmr_ctx_comm_destroy(ctx));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    BSMR_NCCL_OK(nccl().CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    if (!ctx->comm_stream) {
        BSMR_CUDA_OK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->comm_ev) BSMR_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    return BSMR_OK;
}

int bsmr_ctx_comm_destroy(bsmr_ctx* ctx) {
    if (!ctx) return BSMR_ERR_INVALID_ARGUMENT;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->comm_stream) {
            cudaStreamSynchronize(ctx->comm_stream);
            for (auto& e : ctx->comm_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
            cudaStreamDestroy(ctx->comm_stream);
            ctx->comm_stream = nullptr;
        }
        nccl().CommDestroy(comm_of(ctx));
        ctx->nccl_comm = nullptr;
        ctx->comm_rank = 0;
        ctx->comm_world = 1;
    }
    return BSMR_OK;
}

int bsmr_ctx_comm_bcast(bsmr_ctx* ctx, void* device_ptr, uint64_t bytes, int root) {
    if (!ctx || (bytes && !device_ptr)) return BSMR_ERR_INVALID_ARGUMENT;
    BSMR_TRY(need_comm(ctx, "bsmr_ctx_comm_bcast"));
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    if (bytes) BSMR_NCCL_OK(nccl().Broadcast(device_ptr, device_ptr, bytes, ncclUint8, root, comm_of(ctx), ctx->stream));
    return BSMR_OK;
}

// The row order is computed on `root` (bsmr_plan_row_reorder there; the other ranks skip it) and installed on every
// rank like bsmr_plan_set_row_order; each rank then runs bsmr_plan_col_reorder(delta) itself.
int bsmr_plan_bcast_row_order(bsmr_plan* plan, int root) {
    if (!plan) return BSMR_ERR_INVALID_ARGUMENT;
    bsmr_ctx* ctx = plan->ctx;
    BSMR_TRY(need_comm(ctx, "bsmr_plan_bcast_row_order"));
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    const bool is_root = ctx->comm_rank == root;
    if (is_root && !plan->have_rows) {
        set_error("bsmr_plan_bcast_row_order: the root has no row order yet");
        return BSMR_ERR_BAD_STATE;
    }
    // header: {count, num_clusters, num_clusters_true, block_size}
    DevBuf<uint32_t> hdr, rows;
    BSMR_TRY(hdr.alloc(4));
    uint32_t h[4] = {0, 0, 0, 0};
    if (is_root) {
        h[0] = static_cast<uint32_t>(plan->h_reordered_rows.size());
        h[1] = static_cast<uint32_t>(plan->num_clusters);
        h[2] = static_cast<uint32_t>(plan->num_clusters_true);
        h[3] = plan->block_size;
        BSMR_CUDA_OK(cudaMemcpyAsync(hdr.ptr, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    }
    BSMR_NCCL_OK(nccl().Broadcast(hdr.ptr, hdr.ptr, sizeof(h), ncclUint8, root, comm_of(ctx), ctx->stream));
    BSMR_CUDA_OK(cudaMemcpyAsync(h, hdr.ptr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (h[0] > plan->M) {
        set_error("bsmr_plan_bcast_row_order: the root announced %u rows for a matrix with %u", h[0], plan->M);
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    BSMR_TRY(rows.alloc(h[0] ? h[0] : 1));
    if (is_root && h[0]) BSMR_CUDA_OK(cudaMemcpyAsync(rows.ptr, plan->reordered_rows.ptr, (size_t)h[0] * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    if (h[0]) BSMR_NCCL_OK(nccl().Broadcast(rows.ptr, rows.ptr, (size_t)h[0] * 4, ncclUint8, root, comm_of(ctx), ctx->stream));
    if (!is_root) {
        std::vector<uint32_t> host_rows(h[0]);
        if (h[0]) BSMR_CUDA_OK(cudaMemcpyAsync(host_rows.data(), rows.ptr, (size_t)h[0] * 4, cudaMemcpyDeviceToHost, ctx->stream));
        BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        BSMR_TRY(bsmr_plan_set_row_order(plan, host_rows.data(), h[0]));
        plan->num_clusters = static_cast<int>(h[1]);
        plan->num_clusters_true = static_cast<int>(h[2]);
        plan->block_size = h[3];
    } else {
        BSMR_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }
    return BSMR_OK;
}

int bsmr_sddmm_sharded(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP_root, uint32_t flags, int root,
                       bsmr_shard_times* times) {
    if (!plan || !dA || !dB || K == 0) {
        set_error("bsmr_sddmm_sharded: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_TRY(need_comm(ctx, "bsmr_sddmm_sharded"));
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    if (times) std::memset(times, 0, sizeof(*times));
    if (!times) return assemble_sharded(plan, K, dA, dB, dP_root, flags, root, nullptr, nullptr);
    cudaEvent_t ev[5];
    for (auto& e : ev) BSMR_CUDA_OK(cudaEventCreate(&e));
    int s = assemble_sharded(plan, K, dA, dB, dP_root, flags, root, times, ev);
    if (s == BSMR_OK) {
        cudaError_t e = cudaEventSynchronize(ev[4]);
        if (e != cudaSuccess) {
            set_error("bsmr_sddmm_sharded: %s", cudaGetErrorString(e));
            s = BSMR_ERR_CUDA;
        } else {
            cudaEventElapsedTime(&times->kernel_ms, ev[0], ev[1]);
            cudaEventElapsedTime(&times->gather_p_ms, ev[1], ev[2]);      // what of pack + transfer is NOT hidden behind the kernels
            times->pack_ms = 0.f;
            cudaEventElapsedTime(&times->unpermute_ms, ev[3], ev[4]);
            cudaEventElapsedTime(&times->total_ms, ev[0], ev[4]);
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return s;
}

int bsmr_sddmm_sharded_host(bsmr_plan* plan, uint32_t K, const float* hA, const float* hB, float* hP, uint32_t flags, int root,
                            bsmr_shard_times* times) {
    if (!plan || !hA || !hB || K == 0) {
        set_error("bsmr_sddmm_sharded_host: NULL pointer or K == 0");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    bsmr_ctx* ctx = plan->ctx;
    BSMR_TRY(need_comm(ctx, "bsmr_sddmm_sharded_host"));
    BSMR_CUDA_OK(cudaSetDevice(ctx->device));
    const int rank = ctx->comm_rank, world = ctx->comm_world;
    if (rank == root && plan->nnz && !hP) {
        set_error("bsmr_sddmm_sharded_host: the root needs an output buffer");
        return BSMR_ERR_INVALID_ARGUMENT;
    }
    if (!plan->have_format || plan->shard_world != (uint32_t)world || plan->shard_rank != (uint32_t)rank) {
        set_error("bsmr_sddmm_sharded_host: call bsmr_plan_set_shard(rank = %d, world = %d) first", rank, world);
        return BSMR_ERR_BAD_STATE;
    }
    bsmr_shard_times local{};
    bsmr_shard_times* t = times ? times : &local;
    std::memset(t, 0, sizeof(*t));
    const size_t na = (size_t)plan->M * K;
    // How B is split for the upload: rank r uploads columns [bcol[r], bcol[r+1]) and the slices are exchanged with a grouped
    // ncclBroadcast per rank (an all-gather-v).  The slices are NOT equal: nnz-balanced shards differ a lot in their number
    // of rows (on a power-law graph the first shard holds the hub rows: few rows, many nnz), and what a rank's PCIe link
    // has to carry is its A rows PLUS its slice of B -- so B goes preferentially to the ranks with few A rows (water
    // filling on rows of K floats; every rank computes the same split from the shard bounds it already knows).
    std::vector<uint32_t> bcol(static_cast<size_t>(world) + 1, 0);
    {
        const uint32_t nrows = static_cast<uint32_t>(plan->h_reordered_rows.size());
        std::vector<uint64_t> a_rows(world);
        for (int r = 0; r < world; ++r) {
            const uint64_t lo = std::min<uint64_t>((uint64_t)plan->h_shard_bounds[r] * kPanel, nrows);
            const uint64_t hi = std::min<uint64_t>((uint64_t)plan->h_shard_bounds[r + 1] * kPanel, nrows);
            a_rows[r] = hi - lo;
        }
        // level L such that sum_r max(0, L - a_rows[r]) = N  (bisection on integers)
        uint64_t lo = 0, hi = (uint64_t)plan->N + *std::max_element(a_rows.begin(), a_rows.end());
        while (lo < hi) {
            const uint64_t mid = (lo + hi) / 2;
            uint64_t got = 0;
            for (int r = 0; r < world; ++r) got += mid > a_rows[r] ? mid - a_rows[r] : 0;
            if (got >= plan->N) hi = mid; else lo = mid + 1;
        }
        uint64_t left = plan->N;
        for (int r = 0; r < world; ++r) {
            uint64_t take = lo > a_rows[r] ? lo - a_rows[r] : 0;
            if (take > left || r + 1 == world) take = left;
            bcol[r + 1] = bcol[r] + static_cast<uint32_t>(take);
            left -= take;
        }
    }
    const size_t nb_all = (size_t)plan->N * K;
    BSMR_TRY(plan->dA.alloc(na));
    BSMR_TRY(plan->dB.alloc(nb_all));
    BSMR_TRY(plan->dP.alloc(rank == root ? plan->nnz : 1));
    cudaEvent_t ev[9];
    for (auto& e : ev) BSMR_CUDA_OK(cudaEventCreate(&e));
    struct Guard { cudaEvent_t* e; ~Guard() { for (int i = 0; i < 9; ++i) cudaEventDestroy(e[i]); } } guard{ev};
    cudaStream_t st = ctx->stream;
    BSMR_CUDA_OK(cudaEventRecord(ev[0], st));
    // ---- A: only the rows of this rank's panels ----
    const uint32_t r0 = plan->shard_first_panel * kPanel;
    const uint32_t r1 = std::min<uint32_t>(plan->shard_end_panel * kPanel, static_cast<uint32_t>(plan->h_reordered_rows.size()));
    const float* hA_dev = (K % 4 == 0) ? mapped_host_alias(hA) : nullptr;
    if (r1 > r0) {
        if (hA_dev) {
            upload_rows_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(hA_dev, plan->dA.ptr, plan->reordered_rows.ptr + r0, r1 - r0, K);
            ctx->launches++;
            t->h2d_bytes += (uint64_t)(r1 - r0) * K * 4;
        } else {
            // pageable / unmapped host memory: the rows cannot be gathered over PCIe by a kernel, the whole matrix goes up
            BSMR_CUDA_OK(cudaMemcpyAsync(plan->dA.ptr, hA, na * sizeof(float), cudaMemcpyHostToDevice, st));
            t->h2d_bytes += na * 4;
        }
    }
    BSMR_CUDA_OK(cudaEventRecord(ev[1], st));
    // ---- B: this rank's slice, then the all-gather-v ----
    const uint32_t c0 = bcol[rank], c1 = bcol[rank + 1];
    if (c1 > c0) {
        BSMR_CUDA_OK(cudaMemcpyAsync(plan->dB.ptr + (size_t)c0 * K, hB + (size_t)c0 * K, (size_t)(c1 - c0) * K * sizeof(float), cudaMemcpyHostToDevice, st));
        t->h2d_bytes += (uint64_t)(c1 - c0) * K * 4;
    }
    BSMR_CUDA_OK(cudaEventRecord(ev[2], st));
    if (world > 1) {
        BSMR_NCCL_OK(nccl().GroupStart());
        for (int r = 0; r < world; ++r)
            if (bcol[r + 1] > bcol[r]) {
                float* slice = plan->dB.ptr + (size_t)bcol[r] * K;
                BSMR_NCCL_OK(nccl().Broadcast(slice, slice, (size_t)(bcol[r + 1] - bcol[r]) * K, ncclFloat32, r, comm_of(ctx), st));
            }
        BSMR_NCCL_OK(nccl().GroupEnd());
    }
    t->allgather_b_bytes = world > 1 ? (uint64_t)(plan->N - (c1 - c0)) * K * 4 : 0;
    BSMR_CUDA_OK(cudaEventRecord(ev[3], st));
    // ---- kernels, pack, gather-v, un-permute ----
    BSMR_TRY(assemble_sharded(plan, K, plan->dA.ptr, plan->dB.ptr, plan->dP.ptr, flags, root, t, ev + 3));   // stamps 3..7
    // ---- P: root copies the assembled result out ----
    if (rank == root && plan->nnz) {
        BSMR_CUDA_OK(cudaMemcpyAsync(hP, plan->dP.ptr, (size_t)plan->nnz * sizeof(float), cudaMemcpyDeviceToHost, st));
        t->d2h_bytes = (uint64_t)plan->nnz * 4;
    }
    BSMR_CUDA_OK(cudaEventRecord(ev[8], st));
    BSMR_CUDA_OK(cudaEventSynchronize(ev[8]));
    cudaEventElapsedTime(&t->h2d_a_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&t->h2d_b_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&t->allgather_b_ms, ev[2], ev[3]);
    cudaEventElapsedTime(&t->kernel_ms, ev[3], ev[4]);
    cudaEventElapsedTime(&t->gather_p_ms, ev[4], ev[5]);                  // exposed part of pack + transfer
    t->pack_ms = 0.f;
    cudaEventElapsedTime(&t->unpermute_ms, ev[6], ev[7]);
    cudaEventElapsedTime(&t->d2h_ms, ev[7], ev[8]);
    cudaEventElapsedTime(&t->total_ms, ev[0], ev[8]);
    return BSMR_OK;
}

}  // extern "C"
