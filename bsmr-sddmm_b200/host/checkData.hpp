// Result check with the reference's tolerance (include/checkData.hpp:14-79):
// pass iff |a-b| < 1e-5  or  |a-b| / max(|a|, |b|, 1e-3) < 1e-3.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

const float ERROR_THRESHOLD_EPSILON = 1e-3;

template <typename T>
inline bool checkOneData(const T a, const T b) {
    return a == b;
}
template <>
inline bool checkOneData<float>(const float a, const float b) {
    const float d = std::fabs(a - b);
    if (d < 1e-5f) return true;
    const float m = std::max(std::max(std::fabs(a), std::fabs(b)), ERROR_THRESHOLD_EPSILON);
    return d / m < ERROR_THRESHOLD_EPSILON;
}
template <>
inline bool checkOneData<double>(const double a, const double b) {
    const double d = std::fabs(a - b);
    if (d < 1e-5) return true;
    const double m = std::max(std::max(std::fabs(a), std::fabs(b)), static_cast<double>(ERROR_THRESHOLD_EPSILON));
    return d / m < ERROR_THRESHOLD_EPSILON;
}

template <typename T>
inline bool checkData(const size_t n, const T* a, const T* b, size_t& numError) {
    printf("|---------------------------check data---------------------------|\n");
    printf("| Data size : %zu\n| Error threshold epsilon : %f\n| Checking results...\n", n, ERROR_THRESHOLD_EPSILON);
    size_t errors = 0;
    for (size_t i = 0; i < n; ++i) {
        if (checkOneData(a[i], b[i])) continue;
        if (++errors < 10)
            printf("| Error : idx = %zu, data1 = %f, data2 = %f, difference = %f\n", i, static_cast<float>(a[i]),
                   static_cast<float>(b[i]), static_cast<float>(a[i] - b[i]));
    }
    numError = errors;
    if (errors)
        printf("| No Pass! Inconsistent data! %zu errors! Error rate : %2.2f%%\n", errors, 100.0f * errors / static_cast<float>(n));
    else
        printf("| Pass! Result validates successfully.\n");
    printf("|----------------------------------------------------------------|\n");
    return errors == 0;
}
template <typename T>
inline bool checkData(const std::vector<T>& a, const std::vector<T>& b, size_t& numError) {
    if (a.size() != b.size()) {
        numError = std::max(a.size(), b.size());
        return false;
    }
    return checkData(a.size(), a.data(), b.data(), numError);
}
template <typename T>
inline bool checkData(const std::vector<T>& a, const std::vector<T>& b) {
    size_t e = 0;
    return checkData(a, b, e);
}
