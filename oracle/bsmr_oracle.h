/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the BSMR-SDDMM hot path (reference: CX9898/BSMR-SDDMM, files cited per
 * function in bsmr_oracle.c).  It is the parity checker for the CUDA product in
 * bsmr-sddmm_b200/csrc; the product never links, loads or calls it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arms may use it.
 *
 * Parity pinning: every function here is checked against the reference's own code compiled
 * unmodified into oracle/_ref/libbsmr_ref.so (tests/test_oracle_vs_ref.py, CPU parts in the
 * build container; GPU parts -- row clustering -- on the B200 box) and against the committed
 * fixtures in tests/golden/ that were generated from that library.
 */
#ifndef BSMR_ORACLE_H
#define BSMR_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_ROW_PANEL_SIZE 16u
#define ORACLE_BLOCK_COL_SIZE 16u
#define ORACLE_NULL_VALUE 0xFFFFFFFFu

/* ---- a1: .mtx loader (src/Matrix.cpp:399-480, 236-250) ------------------------------- */
typedef struct {
    uint32_t rows, cols, nnz;
    uint32_t* row_offsets; /* rows + 1 */
    uint32_t* col_indices; /* nnz */
    float* values;         /* nnz */
} oracle_csr;
/* returns 1 on success, 0 on any of the reference's rejection conditions */
int oracle_load_mtx(const char* path, oracle_csr* out);
void oracle_free_csr(oracle_csr* csr);

/* ---- a2: Matrix<float>::makeData, single-thread stream (src/Matrix.cpp:117-138) ------ */
void oracle_make_data(size_t n, float* out);

/* ---- a14: sddmm_cpu + checkData (src/host.cpp:44-76, include/checkData.hpp:21-30) ---- */
void oracle_sddmm_cpu(uint32_t M, uint32_t N, uint32_t K, const float* A, const float* B,
                      const uint32_t* row_offsets, const uint32_t* col_indices, int num_threads,
                      float* P);
int oracle_check_one(float a, float b);
uint64_t oracle_check_data(uint64_t n, const float* a, const float* b);

/* ---- a3: calculateBlockSize (src/rowReordering.cu:1009-1025) ------------------------- */
uint32_t oracle_calculate_block_size(uint32_t M, uint32_t N, uint64_t free_mem_bytes);

/* ---- a4: calculateDispersion (src/rowReordering.cu:49-93) ---------------------------- */
/* encodings: M x nb (nb = ceil(N / block_size)), dispersions: M.  */
void oracle_dispersion(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                       uint32_t block_size, uint32_t* encodings, uint32_t* dispersions);

/* ---- a6 helper: clustering CTA size (src/rowReordering.cu:911-920) ------------------- */
uint32_t oracle_clustering_blockdim(uint32_t nb);

/* ---- a6 helper: one similarity evaluation exactly as the device computes it ----------- */
/* (src/rowReordering.cu:235-293 with include/cudaUtil.cuh:13-45); exact_reduce != 0 sums all
 * warps (mathematically intended tree), 0 reproduces the reference's lossy smem tree.      */
float oracle_similarity(const uint32_t* enc_rep, const uint32_t* enc_cmp, uint32_t nb, uint32_t blockdim,
                        int exact_reduce);

/* ---- a4-a6: bsa_rowReordering_gpu (src/rowReordering.cu:1027-1095, 893-1007, 325-432) - */
/* perm_out: capacity M; *num_out = rows kept (empty rows stripped).
 * *clusters_compat = the value the reference stores in numClusters (quirky index, :996);
 * *clusters_true   = number of distinct clusters of non-empty rows.
 * cluster_ids_out (optional, M entries, indexed by position in dispersion order).        */
void oracle_row_reordering(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                           float alpha, uint32_t block_size, int exact_reduce,
                           uint32_t* perm_out, uint32_t* num_out,
                           int* clusters_compat, int* clusters_true);

/* The same permutation computed on sparse encodings with the rows filed per column block, so that it
 * finishes on 10^6-row inputs (see the comment at the definition): the checker for the product at graph
 * scale, itself checked against oracle_row_reordering on every small case.  use_filter = 0 files every row
 * under all of its blocks.  stats (optional, 4 words): evaluations, joins, filed entries, encoding runs.   */
void oracle_row_reordering_indexed(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                                   float alpha, uint32_t block_size, int exact_reduce, int use_filter,
                                   uint32_t* perm_out, uint32_t* num_out, int* clusters_compat, int* clusters_true,
                                   uint32_t* cluster_ids_by_pos, uint64_t* stats);

/* ---- a8: colReordering_cpu (src/colReordering.cu:274-404, 244-271) -------------------- */
typedef struct {
    uint32_t num_row_panels;
    uint32_t* dense_cols;           size_t n_dense_cols;
    uint32_t* dense_col_offsets;    /* panels + 1 */
    uint32_t* sparse_cols;          size_t n_sparse_cols;
    uint32_t* sparse_col_offsets;   /* panels + 1 */
    uint32_t* sparse_value_offsets; /* panels + 1 */
} oracle_colreorder;
void oracle_col_reordering(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                           const uint32_t* reordered_rows, uint32_t num_reordered_rows, float delta,
                           oracle_colreorder* out);
void oracle_free_colreorder(oracle_colreorder* r);

/* ---- a9: RPHM (src/BSMR.cpp:83-265): which CSR index lands in which dense block slot /
 * residual slot.  block_values: numDenseBlocks*256 entries (CSR index or NULL_VALUE),
 * residual triplets ordered exactly like the reference (panel, residual column order, row). */
typedef struct {
    uint32_t num_row_panels;
    uint32_t* block_offsets;  /* panels + 1 */
    uint32_t* block_values;   size_t n_block_values;
    uint32_t* sparse_values;  /* CSR index, n_sparse */
    uint32_t* sparse_relative_rows;
    uint32_t* sparse_col_indices;
    size_t n_sparse;
} oracle_rphm;
void oracle_build_rphm(uint32_t M, uint32_t N, const uint32_t* row_offsets, const uint32_t* col_indices,
                       const uint32_t* reordered_rows, uint32_t num_reordered_rows,
                       const oracle_colreorder* cr, oracle_rphm* out);
void oracle_free_rphm(oracle_rphm* r);

#ifdef __cplusplus
}
#endif
#endif
