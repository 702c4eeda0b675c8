// Run facts + "[key : value]" log, key-compatible with the reference (include/Logger.hpp:13-187) so that
// scripts/analyze_results.cpp keeps parsing our logs.  dim3 is replaced by a plain triple.
#pragma once

#include <cmath>
#include <iomanip>
#include <iostream>
#include <string>

#include "Matrix.hpp"
#include "Options.hpp"

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};

struct Logger {
    Logger() {
#ifdef NDEBUG
        buildType_ = "Release";
#else
        buildType_ = "Debug";
#endif
        // 128 gathered B columns x 16 panel rows x 8 (tf32) per tcgen05.mma; the reference prints its wmma shape here
        wmma_m_ = 128;
        wmma_n_ = 16;
        wmma_k_ = 8;
        matrixA_type_ = matrixB_type_ = matrixC_type_ = "f";
    }
    void getInformation(const Options& o) {
        inputFile_ = o.inputFile();
        K_ = o.K();
        numITER_ = o.numIterations();
        alpha_ = o.similarityThresholdAlpha();
        delta_ = o.blockDensityThresholdDelta();
    }
    void getInformation(const sparseMatrix::DataBase& m) {
        M_ = m.row();
        N_ = m.col();
        NNZ_ = m.nnz();
        sparsity_ = m.getSparsity();
    }
    template <typename T>
    void getInformation(const Matrix<T>& A, const Matrix<T>& B) {
        K_ = A.col();
        matrixA_storageOrder_ = A.storageOrder() == row_major ? "row_major" : "col_major";
        matrixB_storageOrder_ = B.storageOrder() == row_major ? "row_major" : "col_major";
    }

    void printLogInformation(std::ostream& out = std::cout) const {
        out << "[File : " << inputFile_ << "]\n";
        out << "[Build type : " << buildType_ << "]\n";
        out << "[Device : " << gpu_ << "]\n";
        out << "[WMMA_M : " << wmma_m_ << "], [WMMA_N : " << wmma_n_ << "], [WMMA_K : " << wmma_k_ << "]\n";
        out << "[K : " << K_ << "], [M : " << M_ << "], [N : " << N_ << "], [NNZ : " << NNZ_ << "], ";
        out << "[sparsity : " << std::fixed << std::setprecision(2) << (std::floor(sparsity_ * 10000) / 100.0) << "%]\n";
        out << "[matrixA type : " << matrixA_type_ << "]\n[matrixB type : " << matrixB_type_ << "]\n[matrixC type : " << matrixC_type_ << "]\n";
        out << "[matrixA storageOrder : " << matrixA_storageOrder_ << "]\n[matrixB storageOrder : " << matrixB_storageOrder_ << "]\n";
        out << "[Num iterations : " << numITER_ << "]\n";
        out << "[NumRowPanel : " << numRowPanels_ << "]\n";
        out << "[original_numDenseBlock : " << originalNumDenseBlock_ << "]\n";
        out << "[original_averageDensity : " << originalAverageDensity_ << "]\n";
        out << "[bsmr_alpha : " << alpha_ << "]\n[bsmr_delta : " << delta_ << "]\n";
        out << "[bsmr_numClusters : " << numClusters_ << "]\n";
        out << "[bsmr_numDenseBlock : " << numDenseBlock_ << "]\n";
        out << "[bsmr_averageDensity : " << averageDensity_ << "]\n";
        out << "[bsmr_rowReordering : " << rowReorderingTime_ << "]\n";
        out << "[bsmr_colReordering : " << colReorderingTime_ << "]\n";
        out << "[bsmr_reordering : " << reorderingTime_ << "]\n";
        out << "[gridDim_dense : " << gridDim_dense_.x << ", " << gridDim_dense_.y << ", " << gridDim_dense_.z << "]\n";
        out << "[blockDim_dense : " << blockDim_dense_.x << ", " << blockDim_dense_.y << ", " << blockDim_dense_.z << "]\n";
        out << "[gridDim_sparse : " << gridDim_sparse_.x << ", " << gridDim_sparse_.y << ", " << gridDim_sparse_.z << "]\n";
        out << "[blockDim_sparse : " << blockDim_sparse_.x << ", " << blockDim_sparse_.y << ", " << blockDim_sparse_.z << "]\n";
        out << "[bsmr_numDenseThreadBlocks : " << numDenseThreadBlocks_ << "]\n";
        out << "[bsmr_numSparseThreadBlocks : " << numSparseThreadBlocks_ << "]\n";
        out << "[bsmr_threadBlockRatio : " << std::fixed << std::setprecision(2)
            << static_cast<float>(numDenseThreadBlocks_) / numSparseThreadBlocks_ << "]\n";
        out << "[bsmr_numDenseData : " << numDenseData_ << "]\n";
        out << "[bsmr_numSparseData : " << numSparseData_ << "]\n";
        out << "[bsmr_dataRatio: " << std::fixed << std::setprecision(2) << static_cast<float>(numDenseData_) / numSparseData_ << "]\n";
        const size_t flops = 2 * NNZ_ * K_;                       // include/Logger.hpp:178
        out << "[bsmr_gflops : " << (flops / (sddmmTime_ * 1e6)) << "]\n";
        out << "[bsmr_sddmm : " << sddmmTime_ << "]\n";
        out << "[bsmr_formatBuild : " << formatBuildTime_ << "]\n";   // RPHM build time: measured but never printed by the reference
        if (errorRate_ > 0) out << "[checkResults : NO PASS Error rate : " << std::fixed << std::setprecision(2) << errorRate_ << "%]\n";
    }

    std::string inputFile_, checkData_, gpu_, buildType_;
    float errorRate_ = 0.0f;
    size_t wmma_m_ = 0, wmma_n_ = 0, wmma_k_ = 0;
    std::string matrixA_type_, matrixB_type_, matrixC_type_, matrixA_storageOrder_, matrixB_storageOrder_;
    size_t M_ = 0, N_ = 0, K_ = 0, NNZ_ = 0;
    float sparsity_ = 0.0f;
    Dim3 gridDim_dense_, gridDim_sparse_, blockDim_dense_, blockDim_sparse_;
    int numRowPanels_ = 0, numDenseBlock_ = 0;
    float averageDensity_ = 0.0f;
    int originalNumDenseBlock_ = 0;
    float originalAverageDensity_ = 0.0f;
    int numDenseThreadBlocks_ = 0, numSparseThreadBlocks_ = 0, numDenseData_ = 0, numSparseData_ = 0;
    int numITER_ = 10;
    float alpha_ = 0.3f, delta_ = 0.3f;
    int numClusters_ = 1;
    float sddmmTime_ = 0.0f, rowReorderingTime_ = 0.0f, colReorderingTime_ = 0.0f, reorderingTime_ = 0.0f, formatBuildTime_ = 0.0f;
};
