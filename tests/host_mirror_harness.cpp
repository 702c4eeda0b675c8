// TEST HARNESS for the host mirror (bsmr-sddmm_b200/host/*.hpp): plain g++, no CUDA, no C-ABI library.
// It drives the reference-named host classes the way src/main.cu does and dumps what they produce so that
// tests/test_host_mirror.py can compare it with the oracle and with the reference library.
//   load <file> <out.bin>            CSR<float>::initializeFromMatrixFile -> i32 ok, u32 M N nnz, offsets, cols, values
//   makedata <rows> <cols> <out.bin> Matrix<float>::makeData (row-major) -> rows*cols floats
//   options <argv...>                Options getters, one "key=value" per line
//   cpu <file> <K> <out.bin>         sddmm_cpu on makeData() operands -> nnz floats (OMP_NUM_THREADS=1: mt19937 stream)
//   check <a.bin> <b.bin>            checkData on two float files -> "errors=<n>"
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Matrix.hpp"
#include "Options.hpp"
#include "checkData.hpp"
#include "host.hpp"

namespace {

template <typename T>
void put(FILE* f, const std::vector<T>& v) {
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
}

std::vector<float> slurp(const char* path) {
    std::vector<float> v;
    FILE* f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize(static_cast<size_t>(n) / sizeof(float));
    if (fread(v.data(), sizeof(float), v.size(), f) != v.size()) v.clear();
    fclose(f);
    return v;
}

int cmd_load(const char* file, const char* out) {
    sparseMatrix::CSR<float> csr;
    const int32_t ok = csr.initializeFromMatrixFile(file) ? 1 : 0;
    FILE* f = fopen(out, "wb");
    if (!f) return 2;
    fwrite(&ok, sizeof(ok), 1, f);
    if (ok) {
        const uint32_t hdr[3] = {csr.row(), csr.col(), csr.nnz()};
        fwrite(hdr, sizeof(uint32_t), 3, f);
        put(f, csr.rowOffsets());
        put(f, csr.colIndices());
        put(f, csr.values());
    }
    fclose(f);
    return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc < 2) return 2;
    const std::string cmd = argv[1];
    if (cmd == "load" && argc == 4) return cmd_load(argv[2], argv[3]);
    if (cmd == "makedata" && argc == 5) {
        Matrix<float> m(static_cast<UIN>(std::stoul(argv[2])), static_cast<UIN>(std::stoul(argv[3])), MatrixStorageOrder::row_major);
        m.makeData();
        FILE* f = fopen(argv[4], "wb");
        if (!f) return 2;
        put(f, m.values());
        fclose(f);
        return 0;
    }
    if (cmd == "options") {
        const Options o(argc - 1, argv + 1);           // argv[1] ("options") plays the program name
        printf("inputFile=%s\nK=%zu\nalpha=%.6f\ndelta=%.6f\ntestMode=%d\nlogDirectory=%s\n", o.inputFile().c_str(), o.K(),
               o.similarityThresholdAlpha(), o.blockDensityThresholdDelta(), o.testMode() ? 1 : 0, o.outputLogDirectory().c_str());
        return 0;
    }
    if (cmd == "cpu" && argc == 5) {
        sparseMatrix::CSR<float> S;
        if (!S.initializeFromMatrixFile(argv[2])) return 3;
        const UIN K = static_cast<UIN>(std::stoul(argv[3]));
        Matrix<float> A(S.row(), K, MatrixStorageOrder::row_major), B(K, S.col(), MatrixStorageOrder::col_major);
        A.makeData();
        B.makeData();
        sparseMatrix::CSR<float> P(S);
        sddmm_cpu(A, B, S, P);
        FILE* f = fopen(argv[4], "wb");
        if (!f) return 2;
        put(f, P.values());
        fclose(f);
        return 0;
    }
    if (cmd == "check" && argc == 4) {
        const std::vector<float> a = slurp(argv[2]), b = slurp(argv[3]);
        size_t errors = 0;
        checkData(a, b, errors);
        printf("errors=%zu\n", errors);
        return 0;
    }
    return 2;
}
