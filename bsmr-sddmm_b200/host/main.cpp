// BSMR-sddmm command line, same flags / flow / log keys as the reference's src/main.cu:6-42.
//   BSMR-sddmm -f matrix.mtx -k 128 -a 0.3 -d 0.3 [-t 1 -l logdir/] [-b blockSize]
// Define VALIDATE (make VALIDATE=1) to run check_rphm + checkSddmm after the SDDMM like the reference's
// compile-time self-check (src/sddmm.cu:7,35-38).
#include "sddmm.hpp"

int main(int argc, char* argv[]) {
    Options options(argc, argv);

    sparseMatrix::CSR<float> matrixS;
    if (!matrixS.initializeFromMatrixFile(options.inputFile())) {
        fprintf(stderr, "Error, matrix S initialize failed.\n");
        return -1;
    }
    if (options.testMode()) {
        sddmm_testMode(options, matrixS);
        return 0;
    }
    const size_t K = options.K();
    Matrix<float> matrixA(matrixS.row(), K, MatrixStorageOrder::row_major);
    matrixA.makeData();
    Matrix<float> matrixB(K, matrixS.col(), MatrixStorageOrder::col_major);
    matrixB.makeData();

    Logger logger;
    logger.getInformation(options);
    logger.getInformation(matrixS);
    logger.getInformation(matrixA, matrixB);

    sparseMatrix::CSR<float> matrixP(matrixS);
    sddmm(options, matrixA, matrixB, matrixP, logger);
    logger.printLogInformation();
    return 0;
}
