// sddmm(), sddmm_testMode(), checkSddmm() with the reference's signatures (include/sddmm.hpp:8-21,
// src/sddmm.cu:10-118).
#pragma once

#include <fstream>

#include "BSMR.hpp"
#include "Logger.hpp"
#include "Matrix.hpp"
#include "Options.hpp"
#include "checkData.hpp"
#include "host.hpp"
#include "sddmmKernel.hpp"

inline void fillDeviceName(Logger& logger) {
    char name[256] = "";
    bsmr_ctx* ctx = bsmr_host::context();
    if (ctx && bsmr_ctx_device_name(ctx, name, sizeof(name)) == BSMR_OK) logger.gpu_ = name;
}

// BSMR reorder -> RPHM -> sddmm_gpu -> evaluationReordering (src/sddmm.cu:10-39).  P carries S's pattern in
// and the result values out.
inline void sddmm(const Options& options, const Matrix<float>& matrixA, const Matrix<float>& matrixB, sparseMatrix::CSR<float>& matrixP,
                  Logger& logger) {
    BSMR bsmr;
    bsmr.setBlockSize(options.blockSize());
    bsmr.rowReordering(options.similarityThresholdAlpha(), matrixP, 1);
    bsmr.colReordering(options.blockDensityThresholdDelta(), matrixP, std::vector<UIN>(), 1);
    logger.rowReorderingTime_ = bsmr.rowReorderingTime();
    logger.colReorderingTime_ = bsmr.colReorderingTime();
    logger.reorderingTime_ = bsmr.reorderingTime();
    logger.formatBuildTime_ = bsmr.formatBuildTime();
    logger.numRowPanels_ = bsmr.numRowPanels();
    logger.numClusters_ = bsmr.numClusters();
    fillDeviceName(logger);

    RPHM rphm(matrixP, bsmr);
    sddmm_gpu(matrixA, matrixB, rphm, matrixP, logger);
    evaluationReordering(matrixP, bsmr, logger);
#ifdef VALIDATE
    check_rphm(matrixP, bsmr, rphm, options.blockDensityThresholdDelta());
    checkSddmm(matrixA, matrixB, matrixP, matrixP);
#endif
}

// GPU result vs the host computation, reference tolerance (src/sddmm.cu:41-59)
inline bool checkSddmm(const Matrix<float>& matrixA, const Matrix<float>& matrixB, const sparseMatrix::CSR<float>& matrixS,
                       const sparseMatrix::CSR<float>& matrixP) {
    sparseMatrix::CSR<float> cpu(matrixS);
    sddmm_cpu(matrixA, matrixB, matrixS, cpu);
    printf("check cpu sddmm and BSMR sddmm: \n");
    size_t numError = 0;
    if (!checkData(cpu.values(), matrixP.values(), numError)) {
        printf("[checkData : NO PASS Error rate : %2.2f%%]\n", static_cast<float>(numError) / static_cast<float>(matrixP.values().size()) * 100);
        return false;
    }
    return true;
}

// alpha x delta x K sweep, one appended log record per configuration (src/sddmm.cu:62-118)
inline void sddmm_testMode(const Options& options, sparseMatrix::CSR<float>& matrixP) {
    const std::vector<float> alphas = {0.1f, 0.3f, 0.5f, 0.7f, 0.9f};
    const std::vector<float> deltas = {0.0f, 0.1f, 0.3f, 0.5f, 0.7f, 0.9f, 1.1f};
    const std::vector<UIN> Ks = {32, 64, 128, 256};
    BSMR bsmr;
    bsmr.setBlockSize(options.blockSize());
    for (const float alpha : alphas) {
        bsmr.rowReordering(alpha, matrixP);
        for (const float delta : deltas) {
            for (const UIN k : Ks) {
                Matrix<float> matrixA(matrixP.row(), k, row_major);
                matrixA.makeData();
                Matrix<float> matrixB(k, matrixP.col(), col_major);
                matrixB.makeData();
                Logger logger;
                logger.getInformation(options);
                logger.getInformation(matrixP);
                logger.getInformation(matrixA, matrixB);
                logger.alpha_ = alpha;
                logger.delta_ = delta;
                fillDeviceName(logger);
                bsmr.colReordering(delta, matrixP);
                logger.rowReorderingTime_ = bsmr.rowReorderingTime();
                logger.colReorderingTime_ = bsmr.colReorderingTime();
                logger.reorderingTime_ = bsmr.reorderingTime();
                logger.formatBuildTime_ = bsmr.formatBuildTime();
                logger.numRowPanels_ = bsmr.numRowPanels();
                logger.numClusters_ = bsmr.numClusters();
                RPHM rphm(matrixP, bsmr);
                sddmm_gpu(matrixA, matrixB, rphm, matrixP, logger);
                evaluationReordering(matrixP, bsmr, logger);
                const std::string logFile = options.outputLogDirectory() + "BSMR_k_" + util::to_trimmed_string(k) + "_a_" +
                                            util::to_trimmed_string(alpha) + "_d_" + util::to_trimmed_string(delta) + ".log";
                std::ofstream fout(logFile, std::ios::app);
                if (fout.fail()) {
                    fprintf(stderr, "Error, failed to open log file: %s\n", logFile.c_str());
                    return;
                }
                fout << "\n---New data---\n";
                logger.printLogInformation(fout);
            }
        }
    }
}
