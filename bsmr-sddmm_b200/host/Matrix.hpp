// Host-side mirror of the reference data model (include/Matrix.hpp, src/Matrix.cpp): same class
// names, members and loader behaviour, re-implemented header-only on top of the standard library.
// Nothing here touches CUDA; device work goes through the C ABI (include/bsmr_b200.h).
#pragma once

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

using UIN = uint32_t;                              // include/TensorCoreConfig.cuh:10
constexpr UIN MAX_UIN = 0xFFFFFFFFu;
constexpr UIN NULL_VALUE = MAX_UIN;                // include/TensorCoreConfig.cuh:12
constexpr UIN ROW_PANEL_SIZE = 16;                 // include/BSMR.hpp:8
constexpr UIN BLOCK_COL_SIZE = 16;                 // include/BSMR.hpp:9
constexpr UIN BLOCK_SIZE = ROW_PANEL_SIZE * BLOCK_COL_SIZE;

enum MatrixStorageOrder { row_major, col_major };
enum MatrixMultiplicationOrder { left_multiplication, right_multiplication };

namespace util {
// words are separated by ' ', '\t' or '\r' (include/util.hpp:176-190)
inline std::string iterateOneWordFromLine(const std::string& line, int& pos) {
    const int n = static_cast<int>(line.size());
    const int begin = pos;
    while (pos < n && line[pos] != ' ' && line[pos] != '\t' && line[pos] != '\r') ++pos;
    const int end = pos;
    while (pos < n && (line[pos] == ' ' || line[pos] == '\t' || line[pos] == '\r')) ++pos;
    return end > begin ? line.substr(begin, end - begin) : std::string();
}
inline std::string getFileSuffix(const std::string& name) {
    const size_t dot = name.find_last_of('.');
    return dot == std::string::npos ? std::string() : name.substr(dot);
}
inline std::string getParentFolderPath(const std::string& path) {
    const size_t pos = path.find_last_of("/\\");
    return pos == std::string::npos ? std::string() : path.substr(0, pos + 1);
}
inline std::string getFileName(const std::string& path) {
    const size_t pos = path.find_last_of("/\\");
    return pos == std::string::npos ? path : path.substr(pos + 1);
}
template <typename T>
inline std::string to_trimmed_string(T value, int precision = 6) {
    std::ostringstream os;
    os.setf(std::ios::fixed);
    os.precision(precision);
    os << value;
    std::string s = os.str();
    if (s.find('.') != std::string::npos) {
        s.erase(s.find_last_not_of('0') + 1);
        if (!s.empty() && s.back() == '.') s.pop_back();
    }
    return s;
}
}  // namespace util

// Dense matrix with an explicit storage order (include/Matrix.hpp:39-164).
template <typename T>
class Matrix {
public:
    Matrix() = delete;
    Matrix(UIN row, UIN col, MatrixStorageOrder order)
        : row_(row), col_(col), storageOrder_(order), leadingDimension_(order == row_major ? col : row),
          values_(static_cast<size_t>(row) * col) {}
    Matrix(UIN row, UIN col, MatrixStorageOrder order, const std::vector<T>& values)
        : row_(row), col_(col), storageOrder_(order), leadingDimension_(order == row_major ? col : row), values_(values) {
        if (static_cast<size_t>(row) * col != values.size()) std::cout << "Warning! Matrix initialization mismatch" << std::endl;
    }
    Matrix(UIN row, UIN col, MatrixStorageOrder order, const T* values)
        : row_(row), col_(col), storageOrder_(order), leadingDimension_(order == row_major ? col : row),
          values_(values, values + static_cast<size_t>(row) * col) {}

    // makeData (src/Matrix.cpp:117-138): uniform [0, 2) from a default-seeded std::mt19937.  The reference
    // shares the engine across an OpenMP loop (a data race); this is its single-thread stream, which is
    // what the reference produces with OMP_NUM_THREADS=1.
    void makeData() { makeData(row_, col_); }
    void makeData(UIN numRow, UIN numCol) {
        row_ = numRow;
        col_ = numCol;
        leadingDimension_ = storageOrder_ == row_major ? numCol : numRow;
        values_.resize(static_cast<size_t>(numRow) * numCol);
        std::mt19937 generator;
        if constexpr (std::is_floating_point<T>::value) {
            std::uniform_real_distribution<T> dist(static_cast<T>(0), static_cast<T>(2));
            for (auto& v : values_) v = dist(generator);
        } else {
            std::uniform_int_distribution<T> dist(static_cast<T>(0), static_cast<T>(2));
            for (auto& v : values_) v = dist(generator);
        }
    }

    T getOneValue(UIN row, UIN col) const {
        return storageOrder_ == row_major ? values_[static_cast<size_t>(row) * leadingDimension_ + col]
                                          : values_[static_cast<size_t>(col) * leadingDimension_ + row];
    }
    // src/Matrix.cpp:198-222
    T getOneValueForMultiplication(MatrixMultiplicationOrder order, UIN rowC, UIN colC, UIN k) const {
        return order == left_multiplication ? getOneValue(rowC, k) : getOneValue(k, colC);
    }
    void changeStorageOrder() {
        std::vector<T> t(values_.size());
        const MatrixStorageOrder n = storageOrder_ == row_major ? col_major : row_major;
        const UIN ld = n == row_major ? col_ : row_;
        for (UIN r = 0; r < row_; ++r)
            for (UIN c = 0; c < col_; ++c)
                t[n == row_major ? static_cast<size_t>(r) * ld + c : static_cast<size_t>(c) * ld + r] = getOneValue(r, c);
        values_.swap(t);
        storageOrder_ = n;
        leadingDimension_ = ld;
    }
    UIN rowOfValueIndex(UIN idx) const { return storageOrder_ == row_major ? idx / leadingDimension_ : idx % leadingDimension_; }
    UIN colOfValueIndex(UIN idx) const { return storageOrder_ == row_major ? idx % leadingDimension_ : idx / leadingDimension_; }
    std::vector<T> getRowVector(UIN row) const {
        std::vector<T> v(col_);
        for (UIN c = 0; c < col_; ++c) v[c] = getOneValue(row, c);
        return v;
    }
    std::vector<T> getColVector(UIN col) const {
        std::vector<T> v(row_);
        for (UIN r = 0; r < row_; ++r) v[r] = getOneValue(r, col);
        return v;
    }
    bool initializeValue(const std::vector<T>& src) {
        if (src.size() != values_.size()) return false;
        values_ = src;
        return true;
    }
    void print() const {
        for (const auto& v : values_) std::cout << v << " ";
        std::cout << std::endl;
    }

    UIN size() const { return static_cast<UIN>(values_.size()); }
    MatrixStorageOrder storageOrder() const { return storageOrder_; }
    UIN leadingDimension() const { return leadingDimension_; }
    UIN row() const { return row_; }
    UIN col() const { return col_; }
    const std::vector<T>& values() const { return values_; }
    const T* data() const { return values_.data(); }
    const T& operator[](UIN idx) const { return values_[idx]; }
    T& operator[](UIN idx) { return values_[idx]; }

private:
    UIN row_, col_;
    MatrixStorageOrder storageOrder_ = row_major;
    UIN leadingDimension_;
    std::vector<T> values_;
};

template <typename T>
inline std::ostream& operator<<(std::ostream& os, const Matrix<T>& m) {
    return os << " [row : " << m.row() << ", col : " << m.col() << "]";
}

namespace sparseMatrix {

class DataBase {
public:
    UIN row() const { return row_; }
    UIN col() const { return col_; }
    UIN nnz() const { return nnz_; }
    float getSparsity() const {
        const uint64_t total = static_cast<uint64_t>(row_) * col_;
        return total == 0 ? 0.0f : 1.0f - static_cast<float>(nnz_) / static_cast<float>(total);
    }

protected:
    UIN row_ = 0, col_ = 0, nnz_ = 0;
};

template <typename T>
class COO;

namespace detail {
// rowOffsets from sorted row indices (getCsrRowOffsets, src/Matrix.cpp:236-250)
inline std::vector<UIN> rowOffsetsFromSortedRows(UIN rows, const std::vector<UIN>& rowIndices) {
    std::vector<UIN> off(static_cast<size_t>(rows) + 1, 0);
    for (UIN r : rowIndices) ++off[static_cast<size_t>(r) + 1];
    for (UIN r = 0; r < rows; ++r) off[r + 1] += off[r];
    return off;
}
// stable sort by row only: the order inside a row stays the file order (src/Matrix.cpp:467-470).
// Counting sort over the row index: O(nnz + rows), stable by construction (the reference's std::stable_sort over
// tuples is O(nnz log nnz) with a large constant -- minutes on the 2.5e8-nnz graphs).  max_row_hint = number of rows
// when the caller knows it (all indices already validated to be below it), 0 = take it from the data.
template <typename T>
inline void stableSortByRow(std::vector<UIN>& rows, std::vector<UIN>& cols, std::vector<T>& vals, UIN max_row_hint = 0) {
    const size_t n = rows.size();
    if (n == 0) return;
    size_t buckets = max_row_hint;
    if (buckets == 0) buckets = static_cast<size_t>(*std::max_element(rows.begin(), rows.end())) + 1;
    std::vector<size_t> start(buckets + 1, 0);
    for (size_t i = 0; i < n; ++i) ++start[static_cast<size_t>(rows[i]) + 1];
    for (size_t b = 0; b < buckets; ++b) start[b + 1] += start[b];
    std::vector<UIN> r2(n), c2(n);
    std::vector<T> v2(n);
    for (size_t i = 0; i < n; ++i) {
        const size_t dst = start[rows[i]]++;
        r2[dst] = rows[i];
        c2[dst] = cols[i];
        v2[dst] = vals[i];
    }
    rows.swap(r2);
    cols.swap(c2);
    vals.swap(v2);
}
// One word of plain decimal digits (at most 9, so that it is inside std::stoi's range) ending at a separator or at the
// end of the line, parsed in place; `pos` then skips the separators like util::iterateOneWordFromLine.  Anything else
// (sign, blanks first, letters, more digits) returns false and leaves `pos` alone: the caller takes the std::stoi path,
// which accepts or throws exactly as the reference does.
inline bool fastUnsignedWord(const char* s, int& pos, UIN& out) {
    int p = pos, digits = 0;
    UIN value = 0;
    while (s[p] >= '0' && s[p] <= '9' && digits < 10) {
        value = value * 10 + static_cast<UIN>(s[p] - '0');
        ++p;
        ++digits;
    }
    if (digits == 0 || digits > 9) return false;
    if (s[p] != '\0' && s[p] != ' ' && s[p] != '\t' && s[p] != '\r') return false;
    while (s[p] == ' ' || s[p] == '\t' || s[p] == '\r') ++p;
    pos = p;
    out = value;
    return true;
}
// "a b [v]" of one line (src/Matrix.cpp:421-438 reads three words with std::stoi / std::stoi / std::stod).  The common
// case -- two digit words and a number strtod understands -- is parsed in place without a substring per word.
template <typename T>
inline bool readThree(const std::string& line, UIN& a, UIN& b, T& v) {
    if (line.empty()) return false;
    const char* s = line.c_str();
    int pos = 0;
    UIN fa = 0, fb = 0;
    if (fastUnsignedWord(s, pos, fa) && fastUnsignedWord(s, pos, fb)) {
        if (s[pos] == '\0') {
            a = fa;
            b = fb;
            v = static_cast<T>(0);
            return true;
        }
        char* end = nullptr;
        errno = 0;
        const double d = std::strtod(s + pos, &end);
        if (end != s + pos) {                       // std::stod would have converted the same prefix
            a = fa;
            b = fb;
            v = errno == ERANGE ? static_cast<T>(0) : static_cast<T>(d);
            return true;
        }
    }
    pos = 0;
    a = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
    b = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
    const std::string w = util::iterateOneWordFromLine(line, pos);
    try {
        v = w.empty() ? static_cast<T>(0) : static_cast<T>(std::stod(w));
    } catch (const std::out_of_range&) {
        v = static_cast<T>(0);
    }
    return true;
}
}  // namespace detail

// include/Matrix.hpp:196-296
template <typename T>
class CSR : public DataBase {
public:
    CSR() = default;
    CSR(UIN row, UIN col, UIN nnz, const std::vector<UIN>& rowOffsets, const std::vector<UIN>& colIndices,
        const std::vector<T>& values)
        : rowOffsets_(rowOffsets), colIndices_(colIndices), values_(values) { set(row, col, nnz); }
    CSR(UIN row, UIN col, UIN nnz, const UIN* rowOffsets, const UIN* colIndices, const T* values)
        : rowOffsets_(rowOffsets, rowOffsets + row + 1), colIndices_(colIndices, colIndices + nnz), values_(values, values + nnz) {
        set(row, col, nnz);
    }
    CSR(UIN row, UIN col, UIN nnz, const int* rowOffsets, const int* colIndices, const T* values)
        : rowOffsets_(rowOffsets, rowOffsets + row + 1), colIndices_(colIndices, colIndices + nnz), values_(values, values + nnz) {
        set(row, col, nnz);
    }
    CSR(UIN row, UIN col, UIN nnz, const std::vector<UIN>& rowOffsets, const std::vector<UIN>& colIndices)
        : rowOffsets_(rowOffsets), colIndices_(colIndices), values_(nnz, 0) { set(row, col, nnz); }

    // dispatch on the suffix (src/Matrix.cpp:280-294)
    bool initializeFromMatrixFile(const std::string& file) {
        const std::string suffix = util::getFileSuffix(file);
        if (suffix == ".mtx" || suffix == ".mmio") return initializeFromMtxFile(file);
        if (suffix == ".smtx") return initializeFromSmtxFile(file);
        if (suffix == ".txt") return initializeFromGraphDataset(file);
        std::cerr << "Error, file format is not supported : " << file << std::endl;
        return false;
    }

    // MatrixMarket coordinate text (src/Matrix.cpp:399-480): '%' comments, "rows cols nnz", then 1-based
    // "row col [value]"; rejects too many / too few entries, out-of-range, duplicates, nnz <= 1.
    bool initializeFromMtxFile(const std::string& file) {
        std::ifstream in(file);
        if (!in.is_open()) {
            std::cerr << "Error, file cannot be opened : " << file << std::endl;
            return false;
        }
        std::cout << "sparseMatrix::CSR initialize from file : " << file << std::endl;
        std::string line;
        while (std::getline(in, line) && !line.empty() && line[0] == '%') {}
        UIN r = NULL_VALUE, c = NULL_VALUE, n = NULL_VALUE;
        if (!detail::readThree(line, r, c, n) || r == NULL_VALUE || c == NULL_VALUE || n == NULL_VALUE) {
            std::cerr << "Error, file " << file << " format is incorrect!" << std::endl;
            return false;
        }
        set(r, c, n);
        std::vector<UIN> rows(nnz_), cols(nnz_);
        std::vector<T> vals(nnz_);
        UIN idx = 0;
        while (std::getline(in, line)) {
            UIN rr = NULL_VALUE, cc = NULL_VALUE;
            T v{};
            if (!detail::readThree(line, rr, cc, v)) continue;
            if (idx >= nnz_) {
                std::cerr << "Error, file " << file << " too many elements, exceeding the number nnz!" << std::endl;
                return false;
            }
            rows[idx] = rr - 1;
            cols[idx] = cc - 1;
            vals[idx] = v;
            ++idx;
        }
        if (idx < nnz_) {
            std::cerr << "Error, file " << file << " elements is not enough!" << std::endl;
            return false;
        }
        if (!validateCoordinates(file, rows, cols)) return false;
        if (nnz_ <= 1) {
            std::cerr << "Warning, file " << file << " nnz is 1, this is not a valid matrix!" << std::endl;
            return false;
        }
        detail::stableSortByRow(rows, cols, vals, row_);
        rowOffsets_ = detail::rowOffsetsFromSortedRows(row_, rows);
        colIndices_.swap(cols);
        values_.swap(vals);
        return true;
    }

    // DLMC .smtx (src/Matrix.cpp:296-371): "rows, cols, nnz" header (comma or blank separated), one line of
    // row offsets, one line of column indices; values become 1.
    bool initializeFromSmtxFile(const std::string& file) {
        std::ifstream in(file);
        if (!in.is_open()) {
            std::cerr << "Error, file cannot be opened : " << file << std::endl;
            return false;
        }
        std::cout << "sparseMatrix::CSR initialize From file : " << file << std::endl;
        std::string line;
        while (std::getline(in, line) && !line.empty() && line[0] == '%') {}
        int pos = 0;
        const UIN r = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
        const UIN c = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
        const UIN n = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
        set(r, c, n);
        if (nnz_ == 0) {
            std::cerr << "Error, file " << file << " nnz is 0!" << std::endl;
            return false;
        }
        rowOffsets_.assign(static_cast<size_t>(row_) + 1, 0);
        colIndices_.assign(nnz_, 0);
        values_.assign(nnz_, static_cast<T>(1));
        // one line of row offsets, one line of column indices (millions of words on the DLMC masks): digits in place,
        // std::stoi for anything else
        const auto readWords = [&line, &pos](std::vector<UIN>& dst) {
            const char* text = line.c_str();
            for (auto& x : dst)
                if (!detail::fastUnsignedWord(text, pos, x)) x = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
        };
        std::getline(in, line);
        pos = 0;
        readWords(rowOffsets_);
        std::getline(in, line);
        pos = 0;
        readWords(colIndices_);
        std::vector<UIN> sorted_row;                 // duplicate columns inside a row (an unordered_set per row in the reference)
        for (UIN row = 0; row < row_; ++row) {
            sorted_row.assign(colIndices_.begin() + rowOffsets_[row], colIndices_.begin() + rowOffsets_[row + 1]);
            std::sort(sorted_row.begin(), sorted_row.end());
            if (std::adjacent_find(sorted_row.begin(), sorted_row.end()) != sorted_row.end()) {
                std::cerr << "Error, matrix has duplicate data!" << std::endl;
                return false;
            }
        }
        return true;
    }

    // SNAP edge list (src/Matrix.cpp:482-585): "# Nodes: n Edges: m" in the comment header, node ids are
    // renumbered in order of first appearance.
    bool initializeFromGraphDataset(const std::string& file) {
        std::ifstream in(file);
        if (!in.is_open()) {
            std::cerr << "Error, file cannot be opened : " << file << std::endl;
            return false;
        }
        std::cout << "sparseMatrix::CSR initialize From file : " << file << std::endl;
        std::string line;
        while (std::getline(in, line) && !line.empty() && line[0] == '#') {
            const size_t np = line.find("Nodes: "), ep = line.find("Edges: ");
            if (np != std::string::npos) {
                int pos = static_cast<int>(np) + 7;
                row_ = col_ = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
            }
            if (ep != std::string::npos) {
                int pos = static_cast<int>(ep) + 7;
                nnz_ = static_cast<UIN>(std::stoi(util::iterateOneWordFromLine(line, pos)));
            }
        }
        if (!row_ || !col_ || !nnz_) {
            std::cerr << "Error, file " << file << " row or col or nnz not initialized!" << std::endl;
            return false;
        }
        std::vector<UIN> rows(nnz_), cols(nnz_);
        std::vector<T> vals(nnz_, 0);
        std::unordered_map<UIN, UIN> ids;
        UIN idx = 0;
        do {
            UIN a = 0, b = 0;
            T v{};
            if (!detail::readThree(line, a, b, v)) continue;
            const UIN ia = ids.emplace(a, static_cast<UIN>(ids.size())).first->second;
            const UIN ib = ids.emplace(b, static_cast<UIN>(ids.size())).first->second;
            if (idx >= nnz_) {
                std::cerr << "Error, file " << file << " too many elements, exceeding the number nnz!" << std::endl;
                return false;
            }
            rows[idx] = ia;
            cols[idx] = ib;
            vals[idx] = v;
            ++idx;
        } while (std::getline(in, line));
        if (idx < nnz_) {
            std::cerr << "Error, file " << file << " elements is not enough!" << std::endl;
            return false;
        }
        if (!validateCoordinates(file, rows, cols)) return false;
        detail::stableSortByRow(rows, cols, vals, row_);
        rowOffsets_ = detail::rowOffsetsFromSortedRows(row_, rows);
        colIndices_.swap(cols);
        values_.swap(vals);
        return true;
    }

    // src/Matrix.cpp:587-600: MatrixMarket text, 1-based
    bool outputToMarketMatrixFile(const std::string& fileName) const {
        std::ofstream out(fileName + ".mtx");
        if (!out.is_open()) return false;
        out << "%%MatrixMarket matrix coordinate real general\n" << row_ << " " << col_ << " " << nnz_ << "\n";
        for (UIN r = 0; r < row_; ++r)
            for (UIN k = rowOffsets_[r]; k < rowOffsets_[r + 1]; ++k) out << r + 1 << " " << colIndices_[k] + 1 << " " << values_[k] << "\n";
        return true;
    }
    bool outputToMarketMatrixFile() const {
        return outputToMarketMatrixFile("matrix_" + std::to_string(row_) + "_" + std::to_string(col_) + "_" + std::to_string(nnz_));
    }

    const std::vector<UIN>& rowOffsets() const { return rowOffsets_; }
    const std::vector<UIN>& colIndices() const { return colIndices_; }
    const std::vector<T>& values() const { return values_; }
    std::vector<T>& setValues() { return values_; }

private:
    void set(UIN r, UIN c, UIN n) {
        row_ = r;
        col_ = c;
        nnz_ = n;
    }
    // bounds + duplicate check of the loaders (src/Matrix.cpp:442-465: a std::set of pairs filled entry by entry).  A clean
    // file is recognised with one pass and one sort of 64-bit keys; only a file that has a problem takes the entry-by-
    // entry walk, so that the message names the first offending entry exactly as the reference does.
    bool validateCoordinates(const std::string& file, const std::vector<UIN>& rows, const std::vector<UIN>& cols) const {
        bool clean = true;
        for (size_t i = 0; i < rows.size() && clean; ++i) clean = rows[i] < row_ && cols[i] < col_;
        if (clean) {
            std::vector<uint64_t> keys(rows.size());
            for (size_t i = 0; i < rows.size(); ++i) keys[i] = (static_cast<uint64_t>(rows[i]) << 32) | cols[i];
            std::sort(keys.begin(), keys.end());
            clean = std::adjacent_find(keys.begin(), keys.end()) == keys.end();
        }
        if (clean) return true;
        std::set<std::pair<UIN, UIN>> seen;
        for (size_t i = 0; i < rows.size(); ++i) {
            if (rows[i] >= row_ || cols[i] >= col_) {
                std::cerr << "Error, file " << file << " row or col is too big!" << std::endl;
                return false;
            }
            if (!seen.emplace(rows[i], cols[i]).second) {
                std::cerr << "Error, matrix has duplicate data!" << std::endl;
                return false;
            }
        }
        return true;
    }
    std::vector<UIN> rowOffsets_, colIndices_;
    std::vector<T> values_;
};

// include/Matrix.hpp:298-370 (the parts a user of the hot path needs)
template <typename T>
class COO : public DataBase {
public:
    COO() = default;
    COO(UIN row, UIN col, UIN nnz, const std::vector<UIN>& rowIndices, const std::vector<UIN>& colIndices,
        const std::vector<T>& values)
        : rowIndices_(rowIndices), colIndices_(colIndices), values_(values) {
        row_ = row;
        col_ = col;
        nnz_ = nnz;
    }
    explicit COO(const CSR<T>& csr) : colIndices_(csr.colIndices()), values_(csr.values()) {
        row_ = csr.row();
        col_ = csr.col();
        nnz_ = csr.nnz();
        rowIndices_.resize(nnz_);
        for (UIN r = 0; r < row_; ++r)
            for (UIN k = csr.rowOffsets()[r]; k < csr.rowOffsets()[r + 1]; ++k) rowIndices_[k] = r;
    }
    CSR<T> getCsrData() const {
        std::vector<UIN> rows(rowIndices_), cols(colIndices_);
        std::vector<T> vals(values_);
        detail::stableSortByRow(rows, cols, vals);
        return CSR<T>(row_, col_, nnz_, detail::rowOffsetsFromSortedRows(row_, rows), cols, vals);
    }
    const std::vector<UIN>& rowIndices() const { return rowIndices_; }
    const std::vector<UIN>& colIndices() const { return colIndices_; }
    const std::vector<T>& values() const { return values_; }
    std::vector<T>& setValues() { return values_; }
    std::tuple<UIN, UIN, T> operator[](UIN idx) const { return std::make_tuple(rowIndices_[idx], colIndices_[idx], values_[idx]); }

private:
    std::vector<UIN> rowIndices_, colIndices_;
    std::vector<T> values_;
};

}  // namespace sparseMatrix
