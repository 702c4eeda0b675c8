// Command-line options, same flags and defaults as the reference (include/Options.hpp:13-124):
//   -f file  -k K  -a alpha  -d delta  -t testMode  -l logDirectory ; with no flags: "<file> <K>".
// Additions (ignored by the reference): -b clustering block size (0 = calculateBlockSize),
// -i timed iterations, -g number of GPUs (reserved for the multi-GPU driver).
#pragma once

#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>

#include "Matrix.hpp"

class Options {
public:
    Options(const int argc, const char* const argv[]) {
        programPath_ = util::getParentFolderPath(argv[0]);
        programName_ = util::getFileName(argv[0]);
        std::unordered_map<std::string, std::string> seen;
        for (int i = 1; i < argc; ++i) {
            if (argv[i][0] != '-') continue;
            const std::string opt = argv[i];
            if (seen.count(opt)) {
                std::cerr << "Option " << opt << "is duplicated." << std::endl;
                continue;
            }
            if (i + 1 >= argc) {
                std::cerr << "Option " << opt << "requires an argument." << std::endl;
                continue;
            }
            seen[opt] = argv[i + 1];
        }
        for (const auto& kv : seen) parse(kv.first, kv.second);
        if (seen.empty() && argc > 1) {           // positional fallback (include/Options.hpp:119-123)
            inputFile_ = argv[1];
            if (argc > 2) {
                try {
                    K_ = std::stoi(argv[2]);
                } catch (const std::exception& e) {
                    std::cerr << "Invalid argument: " << e.what() << std::endl;
                }
            }
        }
    }

    std::string programPath() const { return programPath_; }
    std::string programName() const { return programName_; }
    std::string inputFile() const { return inputFile_; }
    size_t K() const { return K_; }
    int numIterations() const { return numIterations_; }
    float similarityThresholdAlpha() const { return similarityThresholdAlpha_; }
    float blockDensityThresholdDelta() const { return blockDensityThresholdDelta_; }
    bool testMode() const { return testMode_; }
    std::string outputLogDirectory() const { return outputLogDirectory_; }
    unsigned blockSize() const { return blockSize_; }
    int numGpus() const { return numGpus_; }

private:
    void parse(const std::string& o, const std::string& v) {
        try {
            if (o == "-F" || o == "-f") inputFile_ = v;
            if (o == "-K" || o == "-k") K_ = std::stoi(v);
            if (o == "-A" || o == "-a") similarityThresholdAlpha_ = std::stof(v);
            if (o == "-D" || o == "-d") blockDensityThresholdDelta_ = std::stof(v);
            if (o == "-T" || o == "-t") testMode_ = std::stoi(v);
            if (o == "-L" || o == "-l") outputLogDirectory_ = v;
            if (o == "-B" || o == "-b") blockSize_ = static_cast<unsigned>(std::stoi(v));
            if (o == "-I" || o == "-i") numIterations_ = std::stoi(v);
            if (o == "-G" || o == "-g") numGpus_ = std::stoi(v);
        } catch (const std::invalid_argument& e) {
            std::cerr << "Invalid argument: " << e.what() << std::endl;
        } catch (const std::out_of_range& e) {
            std::cerr << "Out of range: " << e.what() << std::endl;
        }
    }
    std::string programPath_, programName_, inputFile_, outputLogDirectory_;
    size_t K_ = 32;
    int numIterations_ = 10;
    float similarityThresholdAlpha_ = 0.3f;
    float blockDensityThresholdDelta_ = 0.3f;
    bool testMode_ = false;
    unsigned blockSize_ = 0;
    int numGpus_ = 1;
};
