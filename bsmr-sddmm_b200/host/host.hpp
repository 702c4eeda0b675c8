// VALIDATION ONLY.  sddmm_cpu mirrors the reference's host check (include/host.hpp, src/host.cpp:44-76):
// P[idx] = sum_k A[row,k] * B[k,col] in fp32, k ascending, no multiplication by S's value.  It exists so
// that checkSddmm() keeps its reference meaning (GPU result vs a host computation); sddmm(), sddmm_gpu()
// and the BSMR/RPHM objects never call it -- the compute path has no CPU fallback.
#pragma once

#include "Matrix.hpp"

template <typename T>
void sddmm_cpu(const Matrix<T>& A, const Matrix<T>& B, const sparseMatrix::CSR<T>& S, sparseMatrix::CSR<T>& P) {
    if (A.col() != B.row() || A.row() != P.row() || B.col() != P.col()) {
        std::cerr << "The storage of the three matrices does not match" << std::endl;
        return;
    }
    const UIN K = A.col();
#pragma omp parallel for
    for (long long row = 0; row < static_cast<long long>(S.row()); ++row) {
        for (UIN idx = S.rowOffsets()[row]; idx < S.rowOffsets()[row + 1]; ++idx) {
            const UIN col = S.colIndices()[idx];
            T val = 0;
            for (UIN k = 0; k < K; ++k)
                val += A.getOneValueForMultiplication(left_multiplication, static_cast<UIN>(row), col, k) *
                       B.getOneValueForMultiplication(right_multiplication, static_cast<UIN>(row), col, k);
            P.setValues()[idx] = val;
        }
    }
}
