// TEST INFRASTRUCTURE (oracle side).  Driver around the reference's OWN disparity clean-up functions -- matrix_dilate_zero,
// matrix_erode_zero, clean_and_convert_disparity (src/wass_stereo/wass_stereo.cpp:617-733) -- which oracle/build_ref.sh cuts
// out of the reference source AT BUILD TIME into oracle/_ref/filters.inc (nothing of it lives in this repo) and compiles
// against the header shim in oracle/shim/.  It pins oracle/pipeline.py's restatement of those three functions
// (tests/golden/make_filters_golden.py -> tests/golden/filters_golden.npz -> tests/test_oracle_vs_reference_filters.py).
//
//   filters_ref dilate <in.f32> <rows> <cols> <out.f32>
//   filters_ref erode  <in.f32> <rows> <cols> <out.f32>
//   filters_ref clean  <in.s16> <rows> <cols> <mindisp> <numdisp> <offset> <scale> <out.f32>
#include <opencv2/opencv.hpp>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "filters.inc"

static bool slurp(const char* fn, void* dst, size_t n)
{
    std::ifstream f(fn, std::ios::binary);
    return (bool)f.read((char*)dst, (std::streamsize)n);
}
static bool dump(const char* fn, const void* src, size_t n)
{
    std::ofstream f(fn, std::ios::binary);
    return (bool)f.write((const char*)src, (std::streamsize)n);
}

int main(int argc, char** argv)
{
    if (argc < 6) return 64;
    const std::string mode = argv[1];
    const int rows = atoi(argv[3]), cols = atoi(argv[4]);
    if (mode == "dilate" || mode == "erode") {
        cv::Mat src(rows, cols, CV_32FC1), out;
        if (!slurp(argv[2], src.data, (size_t)rows * cols * 4)) return 1;
        if (mode == "dilate") matrix_dilate_zero<float>(src, out); else matrix_erode_zero<float>(src, out);
        return dump(argv[5], out.data, (size_t)rows * cols * 4) ? 0 : 1;
    }
    if (mode == "clean" && argc == 10) {
        cv::Mat src(rows, cols, CV_16SC1);
        if (!slurp(argv[2], src.data, (size_t)rows * cols * 2)) return 1;
        cv::Mat out = clean_and_convert_disparity(src, atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atof(argv[8]));
        return dump(argv[9], out.data, (size_t)rows * cols * 4) ? 0 : 1;
    }
    return 64;
}
