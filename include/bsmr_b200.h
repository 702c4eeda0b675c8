/* bsmr_b200.h -- C ABI of libbsmr_b200.so, the B200 (sm_100a) implementation of the
 * BSMR-SDDMM hot path:  P[idx] = sum_k A[row,k] * B[k,col] at the CSR positions of S,
 * with S reordered by BSMR (row clustering alpha + per-panel column reordering delta) into
 * dense 16-row x 16-column tensor-core blocks plus a sparse residual.
 *
 * The reference (CX9898/BSMR-SDDMM) has no FFI layer; its boundary is a C++ host API.
 * Every entry point below names the reference interface it replaces (file:line relative to
 * the reference tree).  The header-only C++ mirror of that API (the .hpp files in bsmr-sddmm_b200/host:
 * Matrix, sparseMatrix::CSR, Options, Logger, BSMR, RPHM, sddmm(), sddmm_gpu()) is a thin
 * layer over exactly these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary
 *   - every function returns a bsmr_status (0 = OK) and never throws; bsmr_last_error()
 *     gives the text for the most recent failure on the calling thread
 *   - indices are uint32_t ("UIN", include/TensorCoreConfig.cuh:10), values are fp32
 *   - A is M x K row-major (ld = K), B is K x N column-major (ld = K)  (src/main.cu:25-29)
 *   - one context = one device + one stream; one caller thread per context.  A process may hold contexts on several
 *     devices: everything that is a property of a device (kernel attributes, the copy-order probe, the NCCL
 *     communicator) lives in the context
 *   - there is NO CPU fallback: without a usable sm_100 device every compute call fails
 *     with BSMR_ERR_NO_DEVICE
 */
#ifndef BSMR_B200_H
#define BSMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSMR_ROW_PANEL_SIZE 16u        /* ROW_PANEL_SIZE, include/BSMR.hpp:8  */
#define BSMR_BLOCK_COL_SIZE 16u        /* BLOCK_COL_SIZE, include/BSMR.hpp:9  */
#define BSMR_NULL_VALUE 0xFFFFFFFFu    /* NULL_VALUE, include/TensorCoreConfig.cuh:12 */
/* wide row groups (no counterpart in the reference): 256 consecutive reordered rows (16 panels, two
 * 128-row tcgen05 operands resident in shared memory) whose distinct columns stream past them 128 at
 * a time; see csrc/wide_tc.cu                                                                       */
#define BSMR_WIDE_GROUP_ROWS 256u
#define BSMR_WIDE_TILE_COLS  128u

typedef enum {
    BSMR_OK = 0,
    BSMR_ERR_INVALID_ARGUMENT = 1,
    BSMR_ERR_NO_DEVICE = 2,       /* no CUDA device / not sm_100: the product has no CPU path */
    BSMR_ERR_CUDA = 3,            /* a CUDA runtime / driver call or kernel launch failed */
    BSMR_ERR_OUT_OF_MEMORY = 4,
    BSMR_ERR_BAD_STATE = 5,       /* e.g. SDDMM before the column reorder was run */
    BSMR_ERR_UNSUPPORTED = 6
} bsmr_status;

typedef struct bsmr_ctx bsmr_ctx;     /* device + stream + scratch                       */
typedef struct bsmr_plan bsmr_plan;   /* one sparsity pattern: BSMR object + RPHM format */

const char* bsmr_version(void);
const char* bsmr_last_error(void);
const char* bsmr_status_string(int status);

/* ---- context ------------------------------------------------------------------------
 * Replaces the reference's implicit "device 0 + default stream + cudaDeviceSynchronize
 * everywhere" (include/Logger.hpp:23-25, src/sddmmKernel.cu:2555-2559).
 * `cuda_stream` is a cudaStream_t (may be NULL = a stream owned by the context).        */
int bsmr_ctx_create(int device, void* cuda_stream, bsmr_ctx** out);
int bsmr_ctx_destroy(bsmr_ctx* ctx);
int bsmr_ctx_synchronize(bsmr_ctx* ctx);
int bsmr_ctx_device_name(bsmr_ctx* ctx, char* buf, size_t cap);     /* Logger::gpu_ */

/* calculateBlockSize (src/rowReordering.cu:1009-1025):
 * max(16, ceil(M*M*4 / (free/2)), ceil(N*4 / 24576)).  free_mem_bytes == 0 -> query the
 * device like the reference does (cudaMemGetInfo).                                       */
int bsmr_calculate_block_size(bsmr_ctx* ctx, uint32_t M, uint32_t N, uint64_t free_mem_bytes,
                              uint32_t* block_size);

/* ---- plan = BSMR object (include/BSMR.hpp:21-63, src/BSMR.cpp:16-81) -----------------
 * The CSR pattern is copied to the device (values of S are never used: the reference does
 * not multiply by them, src/host.cpp:62-73).  `on_device` != 0 -> the two arrays are device
 * pointers on ctx's device.                                                              */
/* The pattern is validated on the device like the reference's loaders validate a file (src/Matrix.cpp:442-465):
 * row_offsets non-decreasing from 0 to nnz, every column < N, no (row, column) coordinate twice; anything else ->
 * BSMR_ERR_INVALID_ARGUMENT (columns inside a row need not be sorted: the .mtx loader keeps file order).            */
int bsmr_plan_create(bsmr_ctx* ctx, uint32_t M, uint32_t N, uint32_t nnz,
                     const uint32_t* row_offsets, const uint32_t* col_indices, int on_device,
                     bsmr_plan** out);
int bsmr_plan_destroy(bsmr_plan* plan);

/* flags for bsmr_plan_row_reorder */
#define BSMR_ROW_REFERENCE_COMPAT 0u  /* reproduce the reference's lossy block reduction
                                         (include/cudaUtil.cuh:27-45) -> bit-identical rows */
#define BSMR_ROW_EXACT_REDUCE     1u  /* sum every warp (mathematically intended Jaccard)  */
#define BSMR_ROW_IDENTITY         2u  /* noReorderRow (src/rowReordering.cu:15-46): keep the
                                         original order, only strip empty rows              */
/* or-ed in: which step the clustering kernel runs.  Default: chosen from the number of column blocks (thread-prune on
 * graph-shaped inputs, the warp-per-candidate step with its per-warp scratch where rows are long).  Same permutation
 * either way (tests run both on every case).                                                                      */
#define BSMR_ROW_THREAD_PRUNE_ON  4u
#define BSMR_ROW_THREAD_PRUNE_OFF 8u
/* or-ed in: the stage kernel (one CTA per run of 32 consecutive clusters: a row is tested against 32 representatives at
 * once and forwarded once) or the cluster-per-CTA kernel.  Default: the stage kernel on graph-shaped inputs (>= 2^15
 * non-empty rows of at most 128 nnz on average).  Same permutation either way.                                    */
#define BSMR_ROW_STAGE_ON         16u
#define BSMR_ROW_STAGE_OFF        32u

/* BSMR::rowReordering -> bsa_rowReordering_gpu (src/BSMR.cpp:27-50,
 * src/rowReordering.cu:1027-1095): dispersion scores, stable sort, BSA clustering with
 * threshold alpha, stable sort by cluster, empty rows stripped.
 * block_size == 0 -> calculateBlockSize() like BSMR::rowReordering does.                 */
int bsmr_plan_row_reorder(bsmr_plan* plan, float alpha, uint32_t block_size, uint32_t flags);

/* BSMR::colReordering(delta, matrix, reorderedRows) with a caller-supplied row order
 * (src/BSMR.cpp:52-59; host pointer).  Row ids out of range or listed twice are rejected.  */
int bsmr_plan_set_row_order(bsmr_plan* plan, const uint32_t* reordered_rows, uint32_t count);

/* BSMR::colReordering -> colReordering_cpu semantics, computed on the GPU
 * (src/BSMR.cpp:52-81, src/colReordering.cu:274-404, 244-271), followed by the device
 * format build that replaces RPHM::RPHM (src/BSMR.cpp:83-265).                           */
int bsmr_plan_col_reorder(bsmr_plan* plan, float delta);

/* One call = BSMR::BSMR(alpha, delta, matrix) + RPHM(matrix, bsmr).                       */
int bsmr_plan_reorder(bsmr_plan* plan, float alpha, float delta, uint32_t block_size, uint32_t flags);

/* Accessors of BSMR (include/BSMR.hpp:40-50).                                            */
typedef enum {
    BSMR_VEC_REORDERED_ROWS = 0,
    BSMR_VEC_DENSE_COLS = 1,
    BSMR_VEC_DENSE_COL_OFFSETS = 2,
    BSMR_VEC_SPARSE_COLS = 3,
    BSMR_VEC_SPARSE_COL_OFFSETS = 4,
    BSMR_VEC_SPARSE_VALUE_OFFSETS = 5,
    /* RPHM accessors (include/BSMR.hpp:91-103), reference layout, produced on demand */
    BSMR_VEC_BLOCK_OFFSETS = 6,
    BSMR_VEC_BLOCK_VALUES = 7,
    BSMR_VEC_SPARSE_VALUES = 8,
    BSMR_VEC_SPARSE_RELATIVE_ROWS = 9,
    BSMR_VEC_SPARSE_COL_INDICES = 10,
    /* diagnostics of the row reorder */
    BSMR_VEC_DISPERSIONS = 11,
    BSMR_VEC_CLUSTER_IDS = 12,     /* cluster id per ORIGINAL row (0 = empty row) */
    /* execution plan: 1 per row group (256 reordered rows) that runs through the wide kernel */
    BSMR_VEC_GROUP_WIDE = 13
} bsmr_vector_id;
int bsmr_plan_vector_size(bsmr_plan* plan, int which, uint64_t* size);
int bsmr_plan_vector_copy(bsmr_plan* plan, int which, uint32_t* host_out, uint64_t capacity);

typedef struct {
    uint32_t M, N, nnz;
    uint32_t num_row_panels;      /* BSMR::numRowPanels()                                 */
    int32_t  num_clusters;        /* BSMR::numClusters(), the reference's value (quirk kept) */
    int32_t  num_clusters_true;   /* distinct clusters of non-empty rows                  */
    uint32_t block_size;          /* clustering column-block size actually used           */
    uint32_t num_dense_blocks;    /* 16x16 blocks in the dense part                       */
    uint32_t num_dense_tiles;     /* tcgen05 work items (<= 8 blocks each)                */
    uint64_t num_dense_values;    /* nnz computed by the dense-block kernel               */
    uint64_t num_sparse_values;   /* nnz computed by the residual kernel                  */
    float row_reordering_ms;      /* BSMR::rowReorderingTime()                            */
    float col_reordering_ms;      /* BSMR::colReorderingTime()                            */
    float format_build_ms;        /* RPHM::time()                                         */
    float cluster_kernel_ms;      /* part of row_reordering_ms spent in the clustering kernel */
    /* execution plan of the SDDMM kernels (internal layout, not part of the reference's BSMR object):
     * row groups that are dense enough at 128-row scale run through the wide tcgen05 kernel, the
     * remaining groups through the dense-block + residual kernels                                 */
    uint32_t num_row_groups;      /* ceil(#reordered rows / 256)                               */
    uint32_t num_wide_groups;     /* row groups on the wide path                               */
    uint32_t num_wide_tiles;      /* 256 x <=128 work items (two 128 x 128 tcgen05 accumulators) of the wide path           */
    uint32_t num_block_tiles;     /* dense-block tiles left outside the wide groups            */
    uint64_t num_wide_values;     /* nnz computed by the wide kernel                           */
    uint64_t num_block_values;    /* nnz computed by the dense-block kernel (outside wide groups) */
    uint64_t num_residual_values; /* nnz computed by the residual kernel (outside wide groups) */
    float wide_format_ms;         /* part of format_build_ms spent on the wide-group format    */
} bsmr_plan_info;
int bsmr_plan_get_info(bsmr_plan* plan, bsmr_plan_info* info);

/* ---- reorder cache (no counterpart: the reference recomputes the row order on every run, src/sddmm.cu:10-39) ----
 * The row order depends only on the sparsity pattern, alpha, the block size and the reduction mode.  _save writes
 * it (with numClusters and a 64-bit fingerprint of the pattern) to a file; _load checks the fingerprint, alpha and
 * flags against the plan and installs the order like bsmr_plan_set_row_order; the caller then runs
 * bsmr_plan_col_reorder(delta), which rebuilds the column vectors and the device format (about a millisecond).
 * alpha / flags are the values the order was (or would be) computed with by bsmr_plan_row_reorder.           */
int bsmr_plan_fingerprint(bsmr_plan* plan, uint64_t* fingerprint);
int bsmr_plan_save_row_order(bsmr_plan* plan, const char* path, float alpha, uint32_t flags);
int bsmr_plan_load_row_order(bsmr_plan* plan, const char* path, float alpha, uint32_t flags);

/* ---- multi-GPU sharding (no counterpart in the reference, which is single-GPU) --------
 * Restrict the plan to the rank-th of world contiguous ranges of reordered row panels, balanced on
 * the work of the panels: their nnz, plus inside a wide row group the group's tiles (3000 nnz-equivalents
 * each: the wide kernel's time follows the tile count); boundaries fall on row-group boundaries when the
 * plan has wide groups.  SDDMM calls then compute (and write) only the nnz of that range.
 * first_panel / last_panel (exclusive) / shard_nnz are outputs (may be NULL).            */
int bsmr_plan_set_shard(bsmr_plan* plan, uint32_t rank, uint32_t world,
                        uint32_t* first_panel, uint32_t* end_panel, uint64_t* shard_nnz);
/* The tile weight of that balance (nnz-equivalents of one wide tile; default 3000, fitted by hand on stacked nips
 * blocks).  _fit measures it for this plan and K on the unsharded plan from the kernels' own times (wide kernel per
 * tile against dense-block + residual kernels per nnz; synchronises, writes P); every rank of a multi-rank run must
 * install the SAME value with _set before bsmr_plan_set_shard, or the ranks' ranges would not partition the panels. */
int bsmr_plan_fit_tile_work(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP,
                            float* nnz_equivalents_per_tile);
int bsmr_plan_set_tile_work(bsmr_plan* plan, float nnz_equivalents_per_tile);

/* ---- multi-GPU data plane: NCCL over NVLink / NVSwitch, one process per GPU (no counterpart in the reference) ----
 * The communicator belongs to the context; collectives are queued on the context's stream.  NCCL is resolved at run
 * time (dlopen of libnccl.so.2), single-GPU users never load it.
 *   rank 0:      bsmr_comm_unique_id(id)  -> ship the 128 bytes to the other ranks (a file, MPI, torch.distributed ...)
 *   every rank:  bsmr_ctx_comm_init(ctx, id, rank, world)
 *   reorder:     bsmr_plan_row_reorder on the root only (clustering is global and the cost centre), then
 *                bsmr_plan_bcast_row_order(plan, root) on every rank, then bsmr_plan_col_reorder(delta) on every rank
 *                (integer-only and deterministic: recomputing it concurrently is cheaper than shipping the format)
 *   shard:       bsmr_plan_set_shard(plan, rank, world, ...)
 *   SDDMM:       bsmr_sddmm_sharded (A, B resident on every rank) or bsmr_sddmm_sharded_host (host buffers)          */
#define BSMR_NCCL_UNIQUE_ID_BYTES 128
int bsmr_comm_unique_id(void* id_out);
int bsmr_ctx_comm_init(bsmr_ctx* ctx, const void* unique_id, int rank, int world);
int bsmr_ctx_comm_destroy(bsmr_ctx* ctx);
int bsmr_ctx_comm_bcast(bsmr_ctx* ctx, void* device_ptr, uint64_t bytes, int root);   /* ncclBroadcast, in place */
int bsmr_plan_bcast_row_order(bsmr_plan* plan, int root);

typedef struct {
    float h2d_a_ms;        /* host path: upload of the shard's A rows                                   */
    float h2d_b_ms;        /* host path: upload of this rank's slice of B                               */
    float allgather_b_ms;  /* host path: the all-gather-v (grouped ncclBroadcast) that replicates B     */
    float kernel_ms;       /* SDDMM kernels of the shard                                                */
    float pack_ms;         /* 0: packing runs on the communication stream, inside gather_p_ms           */
    float gather_p_ms;     /* pack + ncclSend / ncclRecv of the slices to the root that is NOT hidden   */
                           /* behind the kernels (the shard is cut into chunks; chunk c travels while   */
                           /* chunk c + 1 computes)                                                     */
    float unpermute_ms;    /* root: back to CSR order                                                   */
    float d2h_ms;          /* host path, root: P to the host                                            */
    float total_ms;
    uint64_t shard_nnz;
    uint64_t h2d_bytes, d2h_bytes;       /* this rank */
    uint64_t allgather_b_bytes;          /* received by this rank */
    uint64_t gather_p_bytes;             /* sent (non-root) or received (root) by this rank */
} bsmr_shard_times;

/* One SDDMM over a sharded plan with the result assembled on `root`: every rank runs the kernels of its range of
 * reordered row panels (dA: at least the rows of the shard valid; dB: complete) in a few chunks, packs the entries of a
 * finished chunk into a contiguous slice and sends it to the root while the next chunk computes (a gather-v: 4 * nnz
 * bytes over NVLink in total, on a communication stream of the context), and the root un-permutes them into
 * dP_root (CSR order, length nnz; ignored on the other ranks, may be NULL there).  Asynchronous on the context's stream
 * when times == NULL.                                                                                              */
int bsmr_sddmm_sharded(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP_root, uint32_t flags, int root,
                       bsmr_shard_times* times);
/* The same from host buffers (the sharded form of the host-data sddmm_gpu overload, src/sddmmKernel.cu:2518-2538):
 * hA, hB valid on every rank, hP on the root.  Every rank uploads only the A rows of its shard (straight out of the
 * host buffer when that is pinned + mapped memory and K % 4 == 0, else all of A) and a slice of B sized so that every
 * rank's PCIe link carries the same number of K-vectors (A rows + B columns); the slices are exchanged with a grouped
 * ncclBroadcast per rank (an all-gather-v); then as above, and the root copies P out.  Synchronous; times is filled.  */
int bsmr_sddmm_sharded_host(bsmr_plan* plan, uint32_t K, const float* hA, const float* hB, float* hP, uint32_t flags, int root,
                            bsmr_shard_times* times);

/* ---- SDDMM ---------------------------------------------------------------------------
 * sddmm_gpu(M, N, K, dA, dB, rphm, dP, logger) (include/sddmmKernel.cuh:25-30,
 * src/sddmmKernel.cu:2540-2665 and sddmm_gpu_k32 :2667-2762): device pointers, caller-owned,
 * P in CSR order (length nnz).  Runs `iterations` back-to-back iterations (dense-block kernel
 * and residual kernel per iteration) and returns the average ms per iteration like
 * Logger::sddmmTime_.  iterations <= 0 -> 1.  ms_per_iteration may be NULL (then nothing is
 * synchronised and the call is fully asynchronous on the context's stream).               */
#define BSMR_SDDMM_DEFAULT        0u
#define BSMR_SDDMM_RESIDUAL_ONLY  1u   /* every nnz through the CUDA-core kernel (delta > 1) */
#define BSMR_SDDMM_NO_REORDER     2u   /* ignore the plan's reorder: CSR order, residual kernel */
#define BSMR_SDDMM_NO_WIDE        4u   /* dense-block + residual kernels only, exactly the reference's split */
#define BSMR_SDDMM_THREE_KERNEL   8u   /* the three-kernel plan (wide groups + BSMR split), whatever choice is installed for K */
/* A default call runs the three-kernel plan: the wide row-group kernel on the groups that qualify, the dense-block and
 * residual kernels on the rest.  Operands without a tensor-core path (K % 4 != 0, A / B not 16-byte aligned) go through
 * the CUDA-core kernel entirely, rows in reordered order.  Nothing is measured or synchronised behind the caller's
 * back: with ms_per_iteration == NULL the call is asynchronous and stream-capturable.                               */
int bsmr_sddmm(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP,
               int iterations, uint32_t flags, float* ms_per_iteration);

/* sddmm_gpu(matrixA, matrixB, rphm, matrixP, logger) host-data overload
 * (include/sddmmKernel.cuh:19-23, src/sddmmKernel.cu:2518-2538): H2D A and B, zero P,
 * compute, D2H P.  The copies are inside the call; ms_per_iteration covers kernels only,
 * total_ms (may be NULL) covers copies + kernels.                                        */
int bsmr_sddmm_host(bsmr_plan* plan, uint32_t K, const float* hA, const float* hB, float* hP,
                    int iterations, uint32_t flags, float* ms_per_iteration, float* total_ms);

/* Measured choice of the execution plan for one K -- an explicit tool, not part of bsmr_sddmm: _autotune times the
 * three-kernel plan, the BSMR split alone (= BSMR_SDDMM_NO_WIDE) and, on an unsharded plan, the CSR-order residual
 * kernel (= BSMR_SDDMM_NO_REORDER) on the caller's operands (one warm-up pass + best of three each; synchronises,
 * writes P) and installs the winner for default calls with this K until the next column reorder / set_shard.  The
 * candidates differ in numerics (TF32 tiles ~1.5e-4 relative, fp32 residual ~1e-6) and the outcome depends on timing:
 * ranks that must agree install one choice with _set_execution_choice.  _execution_choice returns what a default call
 * with this K runs (BSMR_SDDMM_DEFAULT = the three-kernel plan).                                                    */
int bsmr_plan_autotune(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP, uint32_t* chosen_flags);
int bsmr_plan_set_execution_choice(bsmr_plan* plan, uint32_t K, uint32_t flags);
int bsmr_plan_execution_choice(bsmr_plan* plan, uint32_t K, uint32_t* flags);

/* Pipelined form of the host-data overload: the same H2D A,B -> zero P -> kernels -> D2H P, but the call returns as
 * soon as the work is queued.  Successive calls alternate between two slots of device buffers; the kernels run on the
 * context's stream, the copies on copy streams, so the copy-in of call i+1 overlaps the kernels of call i.  Whether
 * copies in opposite directions may overlap is measured once per process (on some hosts of this pool that costs 4x in
 * rate, on others it is free): if not, the copy-out of a call is queued behind the copy-in of the next one (or by
 * _wait) on one copy stream; bsmr_ctx_set_host_copy_duplex(ctx, 0 / 1; -1 = measure) overrides.  hA / hB must stay valid and hP must not be read until
 * bsmr_sddmm_host_wait(plan, ticket) returns (BSMR_TICKET_ALL: every call submitted so far); pinned host memory is
 * needed for the copies to be asynchronous.  One caller thread per plan, as everywhere in this ABI.          */
#define BSMR_TICKET_ALL 0xFFFFFFFFFFFFFFFFull
int bsmr_sddmm_host_submit(bsmr_plan* plan, uint32_t K, const float* hA, const float* hB, float* hP,
                           uint32_t flags, uint64_t* ticket);
int bsmr_sddmm_host_wait(bsmr_plan* plan, uint64_t ticket);
int bsmr_ctx_set_host_copy_duplex(bsmr_ctx* ctx, int mode);

/* sddmm_gpu_batch(numBatch, M, N, K, nnz, dA, dB, rphm, dP, time) (include/sddmmKernel.cuh:41-47,
 * src/sddmmKernel.cu:2764-2848): num_batch (A, B, P) triples on the plan's pattern, device pointers, batch b at
 * dA + b*M*K (row-major M x K), dB + b*N*K (column-major K x N) and dP + b*nnz (CSR order).  Every kernel of the plan
 * is launched once for the whole batch (the reference folds the batch into gridDim.z).  total_ms (may be NULL:
 * then the call is asynchronous on the context's stream) = Logger-style total time of the batch.           */
int bsmr_sddmm_batch(bsmr_plan* plan, uint32_t num_batch, uint32_t K, const float* dA, const float* dB, float* dP,
                     uint32_t flags, float* total_ms);
/* The batch with host buffers (same strides): one pipelined host-data call per batch element; returns when
 * every P has landed.  total_ms (may be NULL) = wall time of the call.                                      */
int bsmr_sddmm_host_batch(bsmr_plan* plan, uint32_t num_batch, uint32_t K, const float* hA, const float* hB, float* hP,
                          uint32_t flags, float* total_ms);

/* batchedMatrixTranspose(width, height, numBatches, d_input, d_output) (include/sddmmKernel.cuh:49-51,
 * src/sddmmKernel.cu:2486-2515, 2852-2869): every batch element is a height x width row-major matrix, written back as
 * width x height; elements are width*height apart.  Queued on the context's stream.                          */
int bsmr_batched_transpose(bsmr_ctx* ctx, uint32_t width, uint32_t height, uint32_t num_batches,
                           const float* d_input, float* d_output);

/* ---- fp16 storage of B (the reference sketches half operands in include/TensorCoreConfig.cuh:22-56) ----
 * On graph-shaped patterns the time goes into one K-vector of B per nnz; stored as fp16 (11 significant bits, what the
 * TF32 tensor-core operands keep as well) the vector is half the bytes at every level of the gather.  A stays fp32,
 * products and sums are fp32.  Tolerance of this entry point: the reference's checkData (|d| < 1e-5 or relative
 * < 1e-3; measured ~2e-4).  Every nnz goes through the CUDA-core kernel, rows in reordered order (flags =
 * BSMR_SDDMM_DEFAULT, shard-aware) or CSR order (BSMR_SDDMM_NO_REORDER).
 * _convert: count fp32 values -> fp16, round to nearest even, on the context's stream.                        */
int bsmr_convert_f32_to_f16(bsmr_ctx* ctx, const float* d_src, void* d_dst_f16, uint64_t count);
int bsmr_sddmm_f16b(bsmr_plan* plan, uint32_t K, const float* dA, const void* dB_f16, float* dP,
                    int iterations, uint32_t flags, float* ms_per_iteration);

/* One pass with the two kernels timed separately (CUDA events on the context's stream):
 * what bench.py's roofline block reports per kernel.  Either output may be NULL.          */
int bsmr_sddmm_profile(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP,
                       uint32_t flags, float* dense_ms, float* residual_ms);
/* The same with the wide row-group kernel timed on its own (dense_ms above = wide + dense-block). */
int bsmr_sddmm_profile3(bsmr_plan* plan, uint32_t K, const float* dA, const float* dB, float* dP,
                        uint32_t flags, float* wide_ms, float* dense_ms, float* residual_ms);

/* Wide-path policy of a plan, applied at the next column reorder: a row group goes wide when
 * nnz(group) >= ratio * (128 * tiles(group) + 256).  ratio <= 0 disables the wide path; the default
 * is 5.0, i.e. ~2 % fill of a 256 x 128 tile: below that the residual kernel's one K-vector
 * of B per nnz is cheaper than streaming the tile (measured on R-MAT hub groups).                          */
int bsmr_plan_set_wide_ratio(bsmr_plan* plan, float ratio);
/* Form of the wide kernel's epilogue, applied at the next column reorder: per-entry lists (instruction count follows
 * the nnz), row masks with predicated stores (cost per tile independent of the fill), or chosen by the fill of the
 * wide tiles (>= 8 % -> masks).  Both forms compute the same bits.                                             */
#define BSMR_WIDE_EPILOGUE_AUTO 0
#define BSMR_WIDE_EPILOGUE_LIST 1
#define BSMR_WIDE_EPILOGUE_MASK 2
int bsmr_plan_set_wide_epilogue(bsmr_plan* plan, int form);

/* L2 policy of the residual kernel for operands that do not fit in L2 (no counterpart in the reference, whose sparse
 * kernels read B through __ldg with the default policy, src/sddmmKernel.cu:2043-2060).  When B (N * K * 4 bytes) is
 * larger than min_b_mb MiB, the highest-degree columns whose K-vectors fit hot_budget_mb MiB are loaded with the L2
 * evict_last priority, the A rows, the index lists and P (touched once per pass) with evict_first, and the other
 * columns with evict_first (cold_first != 0) or the default priority.  hot_budget_mb = 0 switches the policy off.
 * Defaults: 64, 2048, 1 (measured on R-MAT graphs, K = 256: -5 % at B = 4.3 GB, -7 % at 8.6 GB; +10 % at B = 537 MB,
 * hence the 2 GiB floor).  Results are bit-identical either way; takes effect at the next SDDMM call.        */
int bsmr_plan_set_l2_policy(bsmr_plan* plan, uint32_t hot_budget_mb, uint32_t min_b_mb, uint32_t cold_first);

/* Number of kernels of this library launched on the context so far (bench.py's gpu_launches). */
int bsmr_ctx_launch_count(bsmr_ctx* ctx, uint64_t* count);

/* evaluationReordering (src/BSMR.cpp:826-930): density statistics for the Logger.        */
typedef struct {
    int32_t num_dense_blocks;      /* Logger::numDenseBlock_      */
    float   average_density;       /* Logger::averageDensity_     */
    int32_t num_dense_thread_blocks;
    int32_t num_sparse_thread_blocks;
    int32_t num_dense_data;        /* Logger::numDenseData_       */
    int32_t num_sparse_data;       /* Logger::numSparseData_      */
    int32_t original_num_dense_blocks;
    float   original_average_density;
} bsmr_reorder_stats;
int bsmr_plan_evaluate(bsmr_plan* plan, float delta, bsmr_reorder_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* BSMR_B200_H */
