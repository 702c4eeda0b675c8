// sddmm_gpu entry points with the reference's signatures and file name (include/sddmmKernel.cuh:19-51,
// src/sddmmKernel.cu:2518-2869).  They forward to bsmr_sddmm / bsmr_sddmm_host / bsmr_sddmm_batch /
// bsmr_batched_transpose.  Plain C++ (no CUDA headers needed): the .cuh suffix is the reference's.
#pragma once

#include "BSMR.hpp"
#include "Logger.hpp"
#include "Matrix.hpp"
#include "bsmr_b200.h"

namespace bsmr_host {
// Work sizes for the Logger's launch-shape fields (the kernels are persistent grids: the tile / chunk counts are what
// the reference's gridDim would be).  Tensor-core work of the plan = wide tiles (17 warps per CTA) + dense-block tiles
// outside the wide groups (13 warps); the residual kernel runs 256 threads over 32-entry chunks.
inline void launch_shape(const bsmr_plan_info& info, Logger& logger) {
    const bool wide = info.num_wide_tiles != 0;
    logger.gridDim_dense_.x = wide ? info.num_wide_tiles + info.num_block_tiles : info.num_dense_tiles;
    logger.blockDim_dense_.x = wide ? 17 * 32 : 13 * 32;
    logger.gridDim_sparse_.x = static_cast<unsigned>(((wide ? info.num_residual_values : info.num_sparse_values) + 255) / 256);
    logger.blockDim_sparse_.x = 256;
}
}  // namespace bsmr_host

// Device pointers, caller-owned, P in CSR order (length nnz).  Runs logger.numITER_ iterations and stores
// the average per-iteration time in logger.sddmmTime_ like the reference.
inline void sddmm_gpu(UIN M, UIN N, UIN K, const float* matrixA, const float* matrixB, const RPHM& rphm, float* matrixP, Logger& logger) {
    (void)M;
    (void)N;
    float ms = 0.0f;
    if (!rphm.plan()) {
        fprintf(stderr, "sddmm_gpu: RPHM has no device format\n");
        return;
    }
    bsmr_host::ok(bsmr_sddmm(rphm.plan(), K, matrixA, matrixB, matrixP, logger.numITER_, BSMR_SDDMM_DEFAULT, &ms), "sddmm_gpu");
    logger.sddmmTime_ = ms;
    bsmr_plan_info info{};
    bsmr_plan_get_info(rphm.plan(), &info);
    bsmr_host::launch_shape(info, logger);
}

// K <= 32 variant of the reference; one code path here.
inline void sddmm_gpu_k32(UIN M, UIN N, UIN K, const float* matrixA, const float* matrixB, const RPHM& rphm, float* matrixP, Logger& logger) {
    sddmm_gpu(M, N, K, matrixA, matrixB, rphm, matrixP, logger);
}

// Host data: uploads A and B, zeroes P, computes, downloads P (src/sddmmKernel.cu:2518-2538).
inline void sddmm_gpu(const Matrix<float>& matrixA, const Matrix<float>& matrixB, const RPHM& rphm, sparseMatrix::CSR<float>& matrixP,
                      Logger& logger) {
    if (!rphm.plan()) {
        fprintf(stderr, "sddmm_gpu: RPHM has no device format\n");
        return;
    }
    if (matrixA.storageOrder() != row_major || matrixB.storageOrder() != col_major) {
        fprintf(stderr, "sddmm_gpu: A must be row-major and B column-major (src/main.cu:25-29)\n");
        return;
    }
    float ms = 0.0f, total = 0.0f;
    bsmr_host::ok(bsmr_sddmm_host(rphm.plan(), matrixA.col(), matrixA.data(), matrixB.data(), matrixP.setValues().data(), logger.numITER_,
                                  BSMR_SDDMM_DEFAULT, &ms, &total),
                  "sddmm_gpu");
    logger.sddmmTime_ = ms;
    bsmr_plan_info info{};
    bsmr_plan_get_info(rphm.plan(), &info);
    bsmr_host::launch_shape(info, logger);
}

// numBatch (A, B, P) triples on one pattern, device pointers strided by M*K, N*K and nnz; `time` = total ms of the batch
// (include/sddmmKernel.cuh:41-47, src/sddmmKernel.cu:2764-2848).
inline void sddmm_gpu_batch(const UIN numBatch, const UIN M, const UIN N, const UIN K, const UIN nnz, const float* matrixA,
                            const float* matrixB, const RPHM& rphm, float* matrixP, float& time) {
    (void)M;
    (void)N;
    (void)nnz;
    time = 0.0f;
    if (!rphm.plan()) {
        fprintf(stderr, "sddmm_gpu_batch: RPHM has no device format\n");
        return;
    }
    bsmr_host::ok(bsmr_sddmm_batch(rphm.plan(), numBatch, K, matrixA, matrixB, matrixP, BSMR_SDDMM_DEFAULT, &time), "sddmm_gpu_batch");
}

// width x height transposes of numBatches matrices that lie width*height apart (include/sddmmKernel.cuh:49-51,
// src/sddmmKernel.cu:2852-2869).
inline void batchedMatrixTranspose(const UIN width, const UIN height, const UIN numBatches, const float* d_input, float* d_output) {
    bsmr_ctx* ctx = bsmr_host::context();
    if (!ctx) return;
    bsmr_host::ok(bsmr_batched_transpose(ctx, width, height, numBatches, d_input, d_output), "batchedMatrixTranspose");
}
