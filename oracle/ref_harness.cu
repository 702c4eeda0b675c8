// TEST INFRASTRUCTURE ONLY -- never linked into, or called by, the product library.
//
// C-ABI harness over the UNMODIFIED reference sources.  It is compiled together with
// /root/reference/src/*.{cpp,cu} (where they lie; nothing is copied) into
// oracle/_ref/libbsmr_ref.so by oracle/Makefile.  Only tests/, bench.py's reference /
// cpu_baseline arms and __graft_entry__.smoke() may load the resulting library.
//
// What it exposes (every function is a thin call into a reference entry point):
//   ref_load_matrix_file      -> sparseMatrix::CSR<float>::initializeFromMatrixFile (src/Matrix.cpp:280)
//   ref_make_data             -> Matrix<float>::makeData                            (src/Matrix.cpp:117)
//   ref_sddmm_cpu             -> sddmm_cpu<float> (CSR)                             (src/host.cpp:44)
//   ref_check_data            -> checkOneData<float>                                (include/checkData.hpp:21)
//   ref_col_reordering_cpu    -> colReordering_cpu                                  (src/colReordering.cu:274)
//   ref_calculate_block_size  -> calculateBlockSize            [GPU]                (src/rowReordering.cu:1009)
//   ref_row_reordering_gpu    -> bsa_rowReordering_gpu         [GPU]                (src/rowReordering.cu:1027)
//   ref_bsmr_sddmm_gpu        -> colReordering + RPHM + sddmm_gpu  [GPU]            (src/BSMR.cpp:52,83; src/sddmmKernel.cu:2518)
//   ref_cusparse_sddmm        -> cusparseSDDMM, same call sequence as the reference's comparator
//                                (baselines/cuSPARSE_SDDMM/include/cuSparseSDDMM.cuh:21-110) [GPU]
// Variable-length results are parked in a result slot table and copied out with
// ref_result_size / ref_result_copy.
#include <omp.h>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>

#include <cusparse.h>

#include "Matrix.hpp"
#include "BSMR.hpp"
#include "host.hpp"
#include "checkData.hpp"
#include "sddmm.hpp"
#include "sddmmKernel.cuh"
#include "Logger.hpp"

namespace {
enum Slot {
    SLOT_ROW_OFFSETS = 0,
    SLOT_COL_INDICES = 1,
    SLOT_REORDERED_ROWS = 2,
    SLOT_DENSE_COLS = 3,
    SLOT_DENSE_COL_OFFSETS = 4,
    SLOT_SPARSE_COLS = 5,
    SLOT_SPARSE_COL_OFFSETS = 6,
    SLOT_SPARSE_VALUE_OFFSETS = 7,
    NUM_SLOTS = 8
};
std::vector<UIN> g_slots[NUM_SLOTS];
std::vector<float> g_values;

sparseMatrix::CSR<float> makeCsr(UIN M, UIN N, UIN nnz, const UIN* rowOffsets, const UIN* colIndices){
    std::vector<float> values(nnz, 1.0f);
    return sparseMatrix::CSR<float>(M, N, nnz, rowOffsets, colIndices, values.data());
}

void parkBsmr(const BSMR& bsmr){
    g_slots[SLOT_REORDERED_ROWS] = bsmr.reorderedRows();
    g_slots[SLOT_DENSE_COLS] = bsmr.denseCols();
    g_slots[SLOT_DENSE_COL_OFFSETS] = bsmr.denseColOffsets();
    g_slots[SLOT_SPARSE_COLS] = bsmr.sparseCols();
    g_slots[SLOT_SPARSE_COL_OFFSETS] = bsmr.sparseColOffsets();
    g_slots[SLOT_SPARSE_VALUE_OFFSETS] = bsmr.sparseValueOffsets();
}
} // namespace

extern "C" {

uint64_t ref_result_size(int slot){
    if (slot < 0 || slot >= NUM_SLOTS) return 0;
    return g_slots[slot].size();
}

void ref_result_copy(int slot, uint32_t* out){
    if (slot < 0 || slot >= NUM_SLOTS) return;
    std::memcpy(out, g_slots[slot].data(), g_slots[slot].size() * sizeof(UIN));
}

uint64_t ref_values_size(){ return g_values.size(); }

void ref_values_copy(float* out){ std::memcpy(out, g_values.data(), g_values.size() * sizeof(float)); }

// Returns 1 on success.  CSR arrays land in slots 0/1 and the value array.
int ref_load_matrix_file(const char* path, uint32_t* rows, uint32_t* cols, uint32_t* nnz){
    sparseMatrix::CSR<float> csr;
    if (!csr.initializeFromMatrixFile(path)) return 0;
    *rows = csr.row();
    *cols = csr.col();
    *nnz = csr.nnz();
    g_slots[SLOT_ROW_OFFSETS] = csr.rowOffsets();
    g_slots[SLOT_COL_INDICES] = csr.colIndices();
    g_values = csr.values();
    return 1;
}

// The reference shares one std::mt19937 across an `omp parallel for`; with one thread
// the stream is the deterministic mt19937(5489) sequence.
void ref_make_data(uint32_t rows, uint32_t cols, int colMajor, int numThreads, float* out){
    const int saved = omp_get_max_threads();
    omp_set_num_threads(numThreads > 0 ? numThreads : 1);
    Matrix<float> m(rows, cols, colMajor ? MatrixStorageOrder::col_major : MatrixStorageOrder::row_major);
    m.makeData();
    std::memcpy(out, m.data(), sizeof(float) * static_cast<size_t>(rows) * cols);
    omp_set_num_threads(saved);
}

// A: M x K row-major, B: K x N col-major (the reference's main.cu layout).
void ref_sddmm_cpu(uint32_t M, uint32_t N, uint32_t K, uint32_t nnz,
                   const float* A, const float* B,
                   const uint32_t* rowOffsets, const uint32_t* colIndices,
                   int numThreads, float* P){
    const int saved = omp_get_max_threads();
    if (numThreads > 0) omp_set_num_threads(numThreads);
    Matrix<float> matrixA(M, K, MatrixStorageOrder::row_major, A);
    Matrix<float> matrixB(K, N, MatrixStorageOrder::col_major, B);
    sparseMatrix::CSR<float> S = makeCsr(M, N, nnz, rowOffsets, colIndices);
    sparseMatrix::CSR<float> Pm(S);
    sddmm_cpu(matrixA, matrixB, S, Pm);
    std::memcpy(P, Pm.values().data(), sizeof(float) * nnz);
    omp_set_num_threads(saved);
}

int ref_omp_max_threads(){ return omp_get_max_threads(); }

// Number of elements failing the reference tolerance (include/checkData.hpp:21-30).
uint64_t ref_check_data(uint64_t n, const float* a, const float* b){
    uint64_t errors = 0;
    for (uint64_t i = 0; i < n; ++i){
        if (!checkOneData<float>(a[i], b[i])) ++errors;
    }
    return errors;
}

void ref_col_reordering_cpu(uint32_t M, uint32_t N, uint32_t nnz,
                            const uint32_t* rowOffsets, const uint32_t* colIndices,
                            const uint32_t* reorderedRows, uint32_t numReorderedRows,
                            float delta){
    sparseMatrix::CSR<float> S = makeCsr(M, N, nnz, rowOffsets, colIndices);
    std::vector<UIN> rows(reorderedRows, reorderedRows + numReorderedRows);
    const UIN numRowPanels = std::ceil(static_cast<float>(rows.size()) / ROW_PANEL_SIZE);
    float time = 0.0f;
    colReordering_cpu(S, numRowPanels, rows, delta,
                      g_slots[SLOT_DENSE_COLS], g_slots[SLOT_DENSE_COL_OFFSETS],
                      g_slots[SLOT_SPARSE_COLS], g_slots[SLOT_SPARSE_COL_OFFSETS],
                      g_slots[SLOT_SPARSE_VALUE_OFFSETS], time);
    g_slots[SLOT_REORDERED_ROWS] = rows;
}

// ---------------------------------------------------------------- GPU-only entry points
uint32_t ref_calculate_block_size(uint32_t M, uint32_t N){
    std::vector<UIN> ro(M + 1, 0), ci;
    sparseMatrix::CSR<float> S(M, N, 0, ro, ci);
    return calculateBlockSize(S);
}

// Row permutation lands in SLOT_REORDERED_ROWS.  Returns the reference's numClusters.
int ref_row_reordering_gpu(uint32_t M, uint32_t N, uint32_t nnz,
                           const uint32_t* rowOffsets, const uint32_t* colIndices,
                           float alpha, uint32_t blockSize, float* timeMs){
    sparseMatrix::CSR<float> S = makeCsr(M, N, nnz, rowOffsets, colIndices);
    int numClusters = 0;
    float t = 0.0f;
    g_slots[SLOT_REORDERED_ROWS] = bsa_rowReordering_gpu(S, alpha, blockSize, numClusters, t);
    if (timeMs) *timeMs = t;
    return numClusters;
}

// Whole reference pipeline with a pinned clustering block size:
// bsa_rowReordering_gpu -> BSMR::colReordering -> RPHM -> sddmm_gpu (host-data overload).
// times[0]=row reorder ms, [1]=col reorder ms, [2]=sddmm ms per iteration (reference's own timers).
int ref_bsmr_sddmm_gpu(uint32_t M, uint32_t N, uint32_t nnz, uint32_t K,
                       const uint32_t* rowOffsets, const uint32_t* colIndices,
                       const float* A, const float* B,
                       float alpha, float delta, uint32_t blockSize, int numIterations,
                       float* P, float* times){
    sparseMatrix::CSR<float> S = makeCsr(M, N, nnz, rowOffsets, colIndices);
    int numClusters = 0;
    float rowTime = 0.0f;
    std::vector<UIN> rows = bsa_rowReordering_gpu(S, alpha, blockSize, numClusters, rowTime);
    BSMR bsmr;
    bsmr.colReordering(delta, S, rows, 1);
    parkBsmr(bsmr);

    RPHM rphm(S, bsmr);
    Matrix<float> matrixA(M, K, MatrixStorageOrder::row_major, A);
    Matrix<float> matrixB(K, N, MatrixStorageOrder::col_major, B);
    sparseMatrix::CSR<float> Pm(S);
    Logger logger;
    logger.numITER_ = numIterations;
    logger.alpha_ = alpha;
    logger.delta_ = delta;
    sddmm_gpu(matrixA, matrixB, rphm, Pm, logger);
    cudaDeviceSynchronize();
    std::memcpy(P, Pm.values().data(), sizeof(float) * nnz);
    if (times){
        times[0] = rowTime;
        times[1] = bsmr.colReorderingTime();
        times[2] = logger.sddmmTime_;
    }
    return numClusters;
}

// cusparseSDDMM with alpha=1, beta=0, ALG_DEFAULT, after preprocess; returns ms per call.
float ref_cusparse_sddmm(uint32_t M, uint32_t N, uint32_t nnz, uint32_t K,
                         const uint32_t* rowOffsets, const uint32_t* colIndices,
                         const float* A, const float* B, int numIterations, float* P){
    float *dA, *dB, *dP;
    int *dRo, *dCi;
    cudaMalloc(&dA, sizeof(float) * static_cast<size_t>(M) * K);
    cudaMalloc(&dB, sizeof(float) * static_cast<size_t>(N) * K);
    cudaMalloc(&dP, sizeof(float) * nnz);
    cudaMalloc(&dRo, sizeof(int) * (M + 1));
    cudaMalloc(&dCi, sizeof(int) * nnz);
    cudaMemcpy(dA, A, sizeof(float) * static_cast<size_t>(M) * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B, sizeof(float) * static_cast<size_t>(N) * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dRo, rowOffsets, sizeof(int) * (M + 1), cudaMemcpyHostToDevice);
    cudaMemcpy(dCi, colIndices, sizeof(int) * nnz, cudaMemcpyHostToDevice);
    cudaMemset(dP, 0, sizeof(float) * nnz);

    cusparseHandle_t handle;
    cusparseCreate(&handle);
    cusparseDnMatDescr_t matA, matB;
    cusparseSpMatDescr_t matP;
    cusparseCreateDnMat(&matA, M, K, K, dA, CUDA_R_32F, CUSPARSE_ORDER_ROW);
    cusparseCreateDnMat(&matB, K, N, K, dB, CUDA_R_32F, CUSPARSE_ORDER_COL);
    cusparseCreateCsr(&matP, M, N, nnz, dRo, dCi, dP, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                      CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F);
    const float one = 1.0f, zero = 0.0f;
    size_t bufferSize = 0;
    cusparseSDDMM_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                             &one, matA, matB, &zero, matP, CUDA_R_32F, CUSPARSE_SDDMM_ALG_DEFAULT, &bufferSize);
    void* dBuffer = nullptr;
    cudaMalloc(&dBuffer, bufferSize ? bufferSize : 4);
    cusparseSDDMM_preprocess(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                             &one, matA, matB, &zero, matP, CUDA_R_32F, CUSPARSE_SDDMM_ALG_DEFAULT, dBuffer);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // one untimed warm-up call, then the timed loop
    cusparseSDDMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                  &one, matA, matB, &zero, matP, CUDA_R_32F, CUSPARSE_SDDMM_ALG_DEFAULT, dBuffer);
    cudaEventRecord(e0);
    for (int i = 0; i < numIterations; ++i){
        cusparseSDDMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                      &one, matA, matB, &zero, matP, CUDA_R_32F, CUSPARSE_SDDMM_ALG_DEFAULT, dBuffer);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(P, dP, sizeof(float) * nnz, cudaMemcpyDeviceToHost);
    cusparseDestroySpMat(matP);
    cusparseDestroyDnMat(matA);
    cusparseDestroyDnMat(matB);
    cusparseDestroy(handle);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dBuffer);
    cudaFree(dA); cudaFree(dB); cudaFree(dP); cudaFree(dRo); cudaFree(dCi);
    return ms / (numIterations > 0 ? numIterations : 1);
}

} // extern "C"
