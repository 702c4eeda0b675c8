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
#include "sddmmKernel.cuh"

namespace bsmr_host {

// name of the device the context runs on, for the log's "GPU" key
inline void noteDevice(Logger& log) {
    char name[256] = "";
    bsmr_ctx* ctx = context();
    if (ctx && bsmr_ctx_device_name(ctx, name, sizeof(name)) == BSMR_OK) log.gpu_ = name;
}

// what a reorder contributes to a log record
inline void noteReorder(Logger& log, const BSMR& order) {
    log.rowReorderingTime_ = order.rowReorderingTime();
    log.colReorderingTime_ = order.colReorderingTime();
    log.reorderingTime_ = order.reorderingTime();
    log.formatBuildTime_ = order.formatBuildTime();
    log.numRowPanels_ = order.numRowPanels();
    log.numClusters_ = order.numClusters();
    noteDevice(log);
}

// device format + SDDMM + density statistics for an order that is already in place; the result lands in `pattern`
inline void computeAndEvaluate(const Matrix<float>& lhs, const Matrix<float>& rhs, const BSMR& order, sparseMatrix::CSR<float>& pattern,
                               Logger& log) {
    const RPHM format(pattern, order);
    sddmm_gpu(lhs, rhs, format, pattern, log);
    evaluationReordering(pattern, order, log);
}

// VALIDATE builds: how many of the post-SDDMM self-checks failed so far (the CLI turns it into its exit code; the
// reference only prints, src/sddmm.cu:35-38)
inline int& validationFailures() {
    static int failures = 0;
    return failures;
}

// file name of one record of the test-mode sweep (src/sddmm.cu:107-110: BSMR_k_<K>_a_<alpha>_d_<delta>.log)
inline std::string sweepLogName(const Options& opts, UIN k, float alpha, float delta) {
    return opts.outputLogDirectory() + "BSMR_k_" + util::to_trimmed_string(k) + "_a_" + util::to_trimmed_string(alpha) + "_d_" +
           util::to_trimmed_string(delta) + ".log";
}

}  // namespace bsmr_host

inline bool checkSddmm(const Matrix<float>& matrixA, const Matrix<float>& matrixB, const sparseMatrix::CSR<float>& matrixS,
                       const sparseMatrix::CSR<float>& matrixP);

// The reference's entry point (include/sddmm.hpp:8-12, src/sddmm.cu:10-39): BSMR reorder with the options' alpha and delta,
// device format, SDDMM, density statistics.  P carries S's pattern in and the result values out.
inline void sddmm(const Options& options, const Matrix<float>& matrixA, const Matrix<float>& matrixB, sparseMatrix::CSR<float>& matrixP,
                  Logger& logger) {
    BSMR order;
    order.setBlockSize(options.blockSize());
    order.rowReordering(options.similarityThresholdAlpha(), matrixP, 1);
    order.colReordering(options.blockDensityThresholdDelta(), matrixP, std::vector<UIN>(), 1);
    bsmr_host::noteReorder(logger, order);
    bsmr_host::computeAndEvaluate(matrixA, matrixB, order, matrixP, logger);
#ifdef VALIDATE
    if (!check_rphm(matrixP, order, RPHM(matrixP, order), options.blockDensityThresholdDelta())) ++bsmr_host::validationFailures();
    if (!checkSddmm(matrixA, matrixB, matrixP, matrixP)) ++bsmr_host::validationFailures();
#endif
}

// GPU result against the host computation under the reference's tolerance (include/sddmm.hpp:18-21, src/sddmm.cu:41-59;
// same two lines of output)
inline bool checkSddmm(const Matrix<float>& matrixA, const Matrix<float>& matrixB, const sparseMatrix::CSR<float>& matrixS,
                       const sparseMatrix::CSR<float>& matrixP) {
    sparseMatrix::CSR<float> expected(matrixS);
    sddmm_cpu(matrixA, matrixB, matrixS, expected);
    printf("check cpu sddmm and BSMR sddmm: \n");
    size_t wrong = 0;
    const bool pass = checkData(expected.values(), matrixP.values(), wrong);
    if (!pass) printf("[checkData : NO PASS Error rate : %2.2f%%]\n", 100.0f * static_cast<float>(wrong) / static_cast<float>(matrixP.values().size()));
    return pass;
}

// `-t 1` (include/sddmm.hpp:14-16, src/sddmm.cu:62-118): the alpha x delta x K grid the reference's published logs come
// from, one record appended per point to <logdir>/BSMR_k_*_a_*_d_*.log.  The row order is computed once per alpha and
// shared by all deltas and Ks, the column order once per (alpha, delta).
inline void sddmm_testMode(const Options& options, sparseMatrix::CSR<float>& matrixP) {
    static const float kAlphas[] = {0.1f, 0.3f, 0.5f, 0.7f, 0.9f};
    static const float kDeltas[] = {0.0f, 0.1f, 0.3f, 0.5f, 0.7f, 0.9f, 1.1f};
    static const UIN kDepths[] = {32, 64, 128, 256};
    BSMR order;
    order.setBlockSize(options.blockSize());
    for (const float alpha : kAlphas) {
        order.rowReordering(alpha, matrixP);
        for (const float delta : kDeltas) {
            for (const UIN depth : kDepths) {
                Matrix<float> lhs(matrixP.row(), depth, row_major), rhs(depth, matrixP.col(), col_major);
                lhs.makeData();
                rhs.makeData();
                Logger record;
                record.getInformation(options);
                record.getInformation(matrixP);
                record.getInformation(lhs, rhs);
                record.alpha_ = alpha;
                record.delta_ = delta;
                order.colReordering(delta, matrixP);        // per point, like the reference (its time is part of the record)
                bsmr_host::noteReorder(record, order);
                bsmr_host::computeAndEvaluate(lhs, rhs, order, matrixP, record);
                const std::string path = bsmr_host::sweepLogName(options, depth, alpha, delta);
                std::ofstream out(path, std::ios::app);
                if (!out) {
                    fprintf(stderr, "Error, failed to open log file: %s\n", path.c_str());
                    return;
                }
                out << "\n---New data---\n";
                record.printLogInformation(out);
            }
        }
    }
}
