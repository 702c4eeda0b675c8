// BSMR reorder object and RPHM device format with the reference's class surface
// (include/BSMR.hpp:21-159, src/BSMR.cpp:16-265), as thin owners of a bsmr_plan behind the C ABI.
// All reordering and the format build run on the GPU inside libbsmr_b200.so.
#pragma once

#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "Logger.hpp"
#include "Matrix.hpp"
#include "bsmr_b200.h"

namespace bsmr_host {

// One context per process, created on first use on device 0 and on the LEGACY DEFAULT STREAM (cudaStreamLegacy = 0x1),
// which is what the reference implicitly uses everywhere (include/Logger.hpp:23-25; kernels launched without a stream
// argument): device pointers a caller prepared on the default stream are ordered before our kernels and our results
// before whatever the caller launches next, exactly as with the reference's code.
inline bsmr_ctx* context(int device = 0) {
    static bsmr_ctx* ctx = nullptr;
    if (!ctx) {
        if (bsmr_ctx_create(device, reinterpret_cast<void*>(0x1), &ctx) != BSMR_OK) {
            fprintf(stderr, "bsmr_b200: %s\n", bsmr_last_error());
            return nullptr;
        }
    }
    return ctx;
}

inline bool ok(int status, const char* what) {
    if (status == BSMR_OK) return true;
    fprintf(stderr, "bsmr_b200: %s failed: %s (%s)\n", what, bsmr_last_error(), bsmr_status_string(status));   // the reference prints and carries on
    return false;
}

struct PlanHandle {
    bsmr_plan* plan = nullptr;
    ~PlanHandle() {
        if (plan) bsmr_plan_destroy(plan);
    }
};

inline std::vector<UIN> fetch(const std::shared_ptr<PlanHandle>& h, int which) {
    std::vector<UIN> v;
    if (!h || !h->plan) return v;
    uint64_t n = 0;
    if (!ok(bsmr_plan_vector_size(h->plan, which, &n), "bsmr_plan_vector_size")) return v;
    v.resize(n);
    ok(bsmr_plan_vector_copy(h->plan, which, v.data(), n), "bsmr_plan_vector_copy");
    return v;
}

}  // namespace bsmr_host

class RPHM;

class BSMR {
public:
    BSMR() = default;
    BSMR(const float similarityThreshold, const float blockDensityThreshold, const sparseMatrix::CSR<float>& matrix,
         const int numIterations = 1) {
        rowReordering(similarityThreshold, matrix, numIterations);
        colReordering(blockDensityThreshold, matrix, std::vector<UIN>(), numIterations);
    }

    // bsa_rowReordering_gpu with block size from calculateBlockSize (src/BSMR.cpp:27-50).  setBlockSize() pins it.
    void rowReordering(const float similarityThreshold, const sparseMatrix::CSR<float>& matrix, const int numIterations = 1) {
        if (!ensurePlan(matrix)) return;
        float total = 0.0f;
        for (int it = 0; it < std::max(1, numIterations); ++it) {
            if (!bsmr_host::ok(bsmr_plan_row_reorder(handle_->plan, similarityThreshold, blockSize_, rowFlags_), "row reorder")) return;
            bsmr_plan_info info{};
            bsmr_plan_get_info(handle_->plan, &info);
            total += info.row_reordering_ms;
        }
        rowReorderingTime_ = total / std::max(1, numIterations);
        refreshRows();
    }

    // colReordering_cpu semantics on the GPU + format build (src/BSMR.cpp:52-81)
    void colReordering(const float blockDensityThreshold, const sparseMatrix::CSR<float>& matrix,
                       const std::vector<UIN>& reorderedRows = std::vector<UIN>(), const int numIterations = 1) {
        if (!ensurePlan(matrix)) return;
        if (!reorderedRows.empty()) {
            if (!bsmr_host::ok(bsmr_plan_set_row_order(handle_->plan, reorderedRows.data(), static_cast<uint32_t>(reorderedRows.size())),
                               "set row order"))
                return;
            refreshRows();
        }
        float total = 0.0f, fmt = 0.0f;
        for (int it = 0; it < std::max(1, numIterations); ++it) {
            if (!bsmr_host::ok(bsmr_plan_col_reorder(handle_->plan, blockDensityThreshold), "col reorder")) return;
            bsmr_plan_info info{};
            bsmr_plan_get_info(handle_->plan, &info);
            total += info.col_reordering_ms;
            fmt += info.format_build_ms;
        }
        colReorderingTime_ = total / std::max(1, numIterations);
        formatBuildTime_ = fmt / std::max(1, numIterations);
        denseCols_ = bsmr_host::fetch(handle_, BSMR_VEC_DENSE_COLS);
        denseColOffsets_ = bsmr_host::fetch(handle_, BSMR_VEC_DENSE_COL_OFFSETS);
        sparseCols_ = bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_COLS);
        sparseColOffsets_ = bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_COL_OFFSETS);
        sparseValueOffsets_ = bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_VALUE_OFFSETS);
    }

    int numRowPanels() const { return numRowPanels_; }
    const std::vector<UIN>& reorderedRows() const { return reorderedRows_; }
    const std::vector<UIN>& denseCols() const { return denseCols_; }
    const std::vector<UIN>& denseColOffsets() const { return denseColOffsets_; }
    const std::vector<UIN>& sparseCols() const { return sparseCols_; }
    const std::vector<UIN>& sparseColOffsets() const { return sparseColOffsets_; }
    const std::vector<UIN>& sparseValueOffsets() const { return sparseValueOffsets_; }
    int numClusters() const { return numClusters_; }
    float rowReorderingTime() const { return rowReorderingTime_; }
    float colReorderingTime() const { return colReorderingTime_; }
    float reorderingTime() const { return rowReorderingTime_ + colReorderingTime_; }

    // additions
    // Reorder cache (no counterpart in the reference, which reclusters on every run): the row order of this pattern for
    // (similarityThreshold, row flags), written to / read from a file keyed by a fingerprint of the pattern.
    // loadRowOrder() replaces rowReordering(); colReordering() then rebuilds the column vectors and the device format.
    bool saveRowOrder(const std::string& path, const float similarityThreshold) const {
        return handle_ && bsmr_host::ok(bsmr_plan_save_row_order(handle_->plan, path.c_str(), similarityThreshold, rowFlags_), "save row order");
    }
    bool loadRowOrder(const std::string& path, const float similarityThreshold, const sparseMatrix::CSR<float>& matrix) {
        if (!ensurePlan(matrix)) return false;
        if (!bsmr_host::ok(bsmr_plan_load_row_order(handle_->plan, path.c_str(), similarityThreshold, rowFlags_), "load row order")) return false;
        rowReorderingTime_ = 0.0f;
        refreshRows();
        return true;
    }
    void setBlockSize(UIN blockSize) { blockSize_ = blockSize; }           // pin the clustering block size (0 = calculateBlockSize)
    void setRowFlags(uint32_t flags) { rowFlags_ = flags; }               // BSMR_ROW_REFERENCE_COMPAT / EXACT_REDUCE / IDENTITY
    float formatBuildTime() const { return formatBuildTime_; }
    const std::shared_ptr<bsmr_host::PlanHandle>& handle() const { return handle_; }

private:
    bool ensurePlan(const sparseMatrix::CSR<float>& m) {
        if (handle_ && handle_->plan) return true;
        bsmr_ctx* ctx = bsmr_host::context();
        if (!ctx) return false;
        auto h = std::make_shared<bsmr_host::PlanHandle>();
        if (!bsmr_host::ok(bsmr_plan_create(ctx, m.row(), m.col(), m.nnz(), m.rowOffsets().data(), m.colIndices().data(), 0, &h->plan),
                           "bsmr_plan_create"))
            return false;
        handle_ = h;
        return true;
    }
    void refreshRows() {
        reorderedRows_ = bsmr_host::fetch(handle_, BSMR_VEC_REORDERED_ROWS);
        bsmr_plan_info info{};
        bsmr_plan_get_info(handle_->plan, &info);
        numClusters_ = info.num_clusters;
        numRowPanels_ = static_cast<int>(std::ceil(static_cast<float>(reorderedRows_.size()) / ROW_PANEL_SIZE));
    }

    std::shared_ptr<bsmr_host::PlanHandle> handle_;
    int numRowPanels_ = 0;
    std::vector<UIN> reorderedRows_, denseCols_, denseColOffsets_, sparseCols_, sparseColOffsets_, sparseValueOffsets_;
    int numClusters_ = 1;
    float rowReorderingTime_ = 0.0f, colReorderingTime_ = 0.0f, formatBuildTime_ = 0.0f;
    UIN blockSize_ = 0;
    uint32_t rowFlags_ = BSMR_ROW_REFERENCE_COMPAT;
};

// The reference's RPHM copies 12 host-built arrays to the device (src/BSMR.cpp:83-265).  Here the device
// format already lives inside the plan (built on the GPU by BSMR::colReordering); RPHM shares that plan and
// materialises the reference-layout arrays only when an accessor asks for them.
class RPHM {
public:
    RPHM() = default;
    RPHM(const sparseMatrix::CSR<float>& matrix, const BSMR& bsmr) : handle_(bsmr.handle()) {
        (void)matrix;
        numRowPanels_ = static_cast<UIN>(bsmr.numRowPanels());
        time_ = bsmr.formatBuildTime();
        // per-CTA work-list sizes of the reference kernels, kept for the Logger
        for (UIN p = 0; p + 1 < bsmr.denseColOffsets().size(); ++p) {
            const UIN blocks = (bsmr.denseColOffsets()[p + 1] - bsmr.denseColOffsets()[p] + BLOCK_COL_SIZE - 1) / BLOCK_COL_SIZE;
            maxNumDenseColBlocksInRowPanel_ = std::max(maxNumDenseColBlocksInRowPanel_, blocks);
            numDenseThreadBlocks_ += (blocks + 3) / 4;
            const UIN sd = bsmr.sparseValueOffsets()[p + 1] - bsmr.sparseValueOffsets()[p];
            const UIN tb = (sd + 127) / 128;
            maxNumSparseColBlocksInRowPanel_ = std::max(maxNumSparseColBlocksInRowPanel_, tb);
            numSparseThreadBlocks_ += tb;
        }
    }

    UIN numRowPanels() const { return numRowPanels_; }
    UIN maxNumDenseColBlocksInRowPanel() const { return maxNumDenseColBlocksInRowPanel_; }
    UIN maxNumSparseColBlocksInRowPanel() const { return maxNumSparseColBlocksInRowPanel_; }
    UIN numDenseThreadBlocks() const { return numDenseThreadBlocks_; }
    UIN numSparseThreadBlocks() const { return numSparseThreadBlocks_; }
    // host copies, reference layout
    std::vector<UIN> reorderedRows() const { return bsmr_host::fetch(handle_, BSMR_VEC_REORDERED_ROWS); }
    std::vector<UIN> denseCols() const { return bsmr_host::fetch(handle_, BSMR_VEC_DENSE_COLS); }
    std::vector<UIN> blockValues() const { return bsmr_host::fetch(handle_, BSMR_VEC_BLOCK_VALUES); }
    std::vector<UIN> blockOffsets() const { return bsmr_host::fetch(handle_, BSMR_VEC_BLOCK_OFFSETS); }
    std::vector<UIN> sparseValueOffsets() const { return bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_VALUE_OFFSETS); }
    std::vector<UIN> sparseValues() const { return bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_VALUES); }
    std::vector<UIN> sparseRelativeRows() const { return bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_RELATIVE_ROWS); }
    std::vector<UIN> sparseColIndices() const { return bsmr_host::fetch(handle_, BSMR_VEC_SPARSE_COL_INDICES); }
    float time() const { return time_; }
    UIN getNumDenseBlocks() const {
        bsmr_plan_info info{};
        if (handle_ && handle_->plan) bsmr_plan_get_info(handle_->plan, &info);
        return info.num_dense_blocks;
    }
    UIN getNumSparseBlocks() const {
        bsmr_plan_info info{};
        if (handle_ && handle_->plan) bsmr_plan_get_info(handle_->plan, &info);
        return static_cast<UIN>(info.num_sparse_values / 128);
    }
    bsmr_plan* plan() const { return handle_ ? handle_->plan : nullptr; }

private:
    std::shared_ptr<bsmr_host::PlanHandle> handle_;
    UIN numRowPanels_ = 0, maxNumDenseColBlocksInRowPanel_ = 0, maxNumSparseColBlocksInRowPanel_ = 0;
    UIN numDenseThreadBlocks_ = 0, numSparseThreadBlocks_ = 0;
    float time_ = 0.0f;
};

// calculateBlockSize (src/rowReordering.cu:1009-1025)
inline UIN calculateBlockSize(const sparseMatrix::CSR<float>& matrix) {
    uint32_t bs = 16;
    bsmr_ctx* ctx = bsmr_host::context();
    if (ctx) bsmr_host::ok(bsmr_calculate_block_size(ctx, matrix.row(), matrix.col(), 0, &bs), "calculateBlockSize");
    return bs;
}

// bsa_rowReordering_gpu (src/rowReordering.cu:1027-1095) as a free function, like the reference declares it
inline std::vector<UIN> bsa_rowReordering_gpu(const sparseMatrix::CSR<float>& matrix, const float alpha, const UIN block_size,
                                              int& num_clusters, float& reordering_time) {
    BSMR b;
    b.setBlockSize(block_size);
    b.rowReordering(alpha, matrix, 1);
    num_clusters = b.numClusters();
    reordering_time = b.rowReorderingTime();
    return b.reorderedRows();
}

// evaluationReordering (src/BSMR.cpp:826-930)
inline void evaluationReordering(const sparseMatrix::CSR<float>& matrix, const BSMR& bsmr, Logger& logger) {
    (void)matrix;
    if (!bsmr.handle() || !bsmr.handle()->plan) return;
    bsmr_reorder_stats st{};
    if (!bsmr_host::ok(bsmr_plan_evaluate(bsmr.handle()->plan, logger.delta_, &st), "evaluationReordering")) return;
    logger.numDenseBlock_ = st.num_dense_blocks;
    logger.averageDensity_ = st.average_density;
    logger.numDenseThreadBlocks_ = st.num_dense_thread_blocks;
    logger.numSparseThreadBlocks_ = st.num_sparse_thread_blocks;
    logger.originalNumDenseBlock_ = st.original_num_dense_blocks;
    logger.originalAverageDensity_ = st.original_average_density;
    logger.numSparseData_ = st.num_sparse_data;
    logger.numDenseData_ = st.num_dense_data;
}

// check_rphm (src/BSMR.cpp:444-824, 932-953): structural invariants of the reorder and the format.
inline bool check_rphm(const sparseMatrix::CSR<float>& matrix, const BSMR& bsmr, const RPHM& rphm, const float delta) {
    bool good = true;
    // rows: no duplicates, no empty rows, nothing missing
    std::vector<char> seen(matrix.row(), 0);
    for (UIN r : bsmr.reorderedRows()) {
        if (r >= matrix.row() || seen[r] || matrix.rowOffsets()[r + 1] == matrix.rowOffsets()[r]) good = false;
        if (r < matrix.row()) seen[r] = 1;
    }
    for (UIN r = 0; r < matrix.row(); ++r)
        if (!seen[r] && matrix.rowOffsets()[r + 1] != matrix.rowOffsets()[r]) good = false;
    if (!good) std::cerr << "Error! The row reordering is incorrect!" << std::endl;
    // every nnz exactly once in blockValues U sparseValues
    std::vector<char> hit(matrix.nnz(), 0);
    bool cover = true;
    for (UIN v : rphm.blockValues())
        if (v != NULL_VALUE) {
            if (v >= matrix.nnz() || hit[v]) cover = false; else hit[v] = 1;
        }
    for (UIN v : rphm.sparseValues()) {
        if (v >= matrix.nnz() || hit[v]) cover = false; else hit[v] = 1;
    }
    for (char h : hit) cover = cover && h;
    if (!cover) std::cerr << "Error! The rphm is incorrect!" << std::endl;
    // dense blocks meet the threshold
    const UIN thr = static_cast<UIN>(std::ceil(delta * BLOCK_SIZE));
    const std::vector<UIN> bv = rphm.blockValues();
    bool dense_ok = true;
    for (size_t b = 0; b * 256 < bv.size(); ++b) {
        UIN n = 0;
        for (size_t i = 0; i < 256; ++i) n += bv[b * 256 + i] != NULL_VALUE;
        if (n < thr) dense_ok = false;
    }
    if (!dense_ok) std::cerr << "Error! The col reordering is incorrect!" << std::endl;
    return good && cover && dense_ok;
}
