// Command-line front end of the B200 path: the flags, exit codes and log keys of the reference's binary
// (`-f matrix -k K -a alpha -d delta [-t 1 -l logdir/]`, src/main.cu), driven through the host mirror in this directory.
//   make VALIDATE=1 additionally runs check_rphm + checkSddmm after the SDDMM (the reference's compile-time self-check,
//   src/sddmm.cu:7,35-38).
#include <cstdio>

#include "sddmm.hpp"

namespace {

// the dense operands of one run: A is M x K row-major, B is K x N column-major, both from makeData()
struct Operands {
    Matrix<float> lhs, rhs;
    Operands(size_t rows, size_t cols, size_t depth)
        : lhs(rows, depth, MatrixStorageOrder::row_major), rhs(depth, cols, MatrixStorageOrder::col_major) {
        lhs.makeData();
        rhs.makeData();
    }
};

// one SDDMM on the pattern: reorder, compute, print the log
int run_single(const Options& opts, const sparseMatrix::CSR<float>& pattern) {
    Operands ab(pattern.row(), pattern.col(), opts.K());
    Logger log;
    log.getInformation(opts);
    log.getInformation(pattern);
    log.getInformation(ab.lhs, ab.rhs);
    sparseMatrix::CSR<float> result(pattern);          // carries the pattern in, the values out
    sddmm(opts, ab.lhs, ab.rhs, result, log);
    log.printLogInformation();
    return bsmr_host::validationFailures() ? 1 : 0;     // non-zero only in VALIDATE builds
}

}  // namespace

int main(int argc, char* argv[]) {
    const Options opts(argc, argv);
    sparseMatrix::CSR<float> pattern;
    if (!pattern.initializeFromMatrixFile(opts.inputFile())) {
        fprintf(stderr, "Error, matrix S initialize failed.\n");
        return -1;
    }
    if (!bsmr_host::context()) {                        // no CPU path: without a usable device there is nothing to run
        fprintf(stderr, "Error, no usable CUDA device for libbsmr_b200.\n");
        return -1;
    }
    if (opts.testMode()) {                              // the sweep that produced the reference's published logs
        sddmm_testMode(opts, pattern);
        return 0;
    }
    return run_single(opts, pattern);
}
