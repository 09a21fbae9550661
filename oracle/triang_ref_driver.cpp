// TEST INFRASTRUCTURE ONLY.  Driver around the REFERENCE's own per-pixel triangulation: `size_t triangulate( StereoMatchEnv& )`
// with StereoMatchEnv::unrectify (src/wass_stereo/wass_stereo.cpp:299-324, 1039-1386), cut out of the reference source at
// build time by oracle/cut_triangulate.awk (nothing of it lives in this repository; the cut is deleted after the compile),
// plus the reference's PovMesh.cpp and triangulate.hpp from their own paths, against the header shim in oracle/shim/.
// It pins oracle/pipeline.py's triangulate() -- the gates (disparity > 1, rectified-column range, image border, bounding box,
// mask images, burned areas, minimum angle, distance limits), the float32 / float64 mix of the pixel coordinates, both
// un-rectification branches -- to the reference itself
// (tests/golden/make_triang_golden.py -> tests/golden/triang_golden.npz -> tests/test_oracle_vs_reference_triangulate.py).
//
//   triang_ref <case.bin> <config> <out.bin>
//     case.bin  int32 W, H (original images), RW, RH (rectified), roiL[4], roiR[4]; float64 K0[9], K1[9], R[9], T[3], R1[9],
//               R2[9], P1[12], P2[12], HLi[9], HRi[9], disparity_compensation, cam_distance; uint8 left[W*H], right[W*H],
//               left_rect[RW*RH], right_rect[RW*RH]; float32 disparity[RW*RH]
//     config    incfg file (TRIANG_MIN_ANGLE, TRIANG_BBOX_*, LEFT_MASK_IMAGE / RIGHT_MASK_IMAGE as .pgm in the workdir =
//               the directory of case.bin, DISCARD_BURNED_AREAS, DENSE_SCALE, USE_CUSTOM_STEREORECTIFY)
//     out.bin   int64 n; then the roiR[2] x roiR[3] grid: uint8 valid[], float64 xyz[][3], uint8 grey[]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <opencv2/opencv.hpp>
#include <boost/filesystem.hpp>
#include <boost/shared_ptr.hpp>
#include <boost/cstdint.hpp>
#include "log.hpp"
#include "incfg.hpp"
#include "hires_timer.h"

#define private public
#include "PovMesh.cpp"          // the reference's file, from -I<reference>/src/wass_stereo
#undef private
#include "triangulate.hpp"      // -I<reference>/src/wass_lib

#include "keys_ws.inc"          // the INCFG_REQUIRE declarations of wass_stereo.cpp (cut at build time)
#include "triang_env.inc"       // the cut described above

template <typename T> static bool rd(std::ifstream& f, T* p, size_t n) { return (bool)f.read((char*)p, (std::streamsize)(n * sizeof(T))); }
static cv::Mat mat64(std::ifstream& f, int r, int c) { cv::Mat m(r, c, CV_64FC1); rd(f, (double*)m.data, (size_t)r * c); return m; }
static cv::Mat img8(std::ifstream& f, int r, int c) { cv::Mat m(r, c, CV_8UC1); rd(f, m.data, (size_t)r * c); return m; }

int main(int argc, char** argv)
{
    if (argc != 4) return 64;
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) return 1;
    {
        std::ifstream cfg(argv[2]);
        incfg::ConfigOptions::instance().load(cfg);
    }
    int32_t hd[12];
    rd(f, hd, 12);
    const int W = hd[0], H = hd[1], RW = hd[2], RH = hd[3];
    StereoMatchEnv env;
    env.workdir = std::filesystem::path(argv[1]).parent_path();
    env.roi_comb_left = cv::Rect(hd[4], hd[5], hd[6], hd[7]);
    env.roi_comb_right = cv::Rect(hd[8], hd[9], hd[10], hd[11]);
    env.intrinsics_left = mat64(f, 3, 3);
    env.intrinsics_right = mat64(f, 3, 3);
    env.R = mat64(f, 3, 3);
    env.T = mat64(f, 3, 1);
    env.rec_R1 = mat64(f, 3, 3);
    env.rec_R2 = mat64(f, 3, 3);
    env.rec_P1 = mat64(f, 3, 4);
    env.rec_P2 = mat64(f, 3, 4);
    rd(f, env.HLi.val, 9);
    rd(f, env.HRi.val, 9);
    double sc[2];
    rd(f, sc, 2);
    env.disparity_compensation = sc[0];
    env.cam_distance = sc[1];
    env.left = img8(f, H, W);
    env.right = img8(f, H, W);
    env.left_rectified = img8(f, RH, RW);
    env.right_rectified = img8(f, RH, RW);
    env.disparity = cv::Mat(RH, RW, CV_32FC1);
    if (!rd(f, (float*)env.disparity.data, (size_t)RW * RH)) return 1;
    env.P0 = cv::Mat::zeros(3, 4, CV_64FC1);      // only feed the dbg_P0 / dbg_P1 debug images, which are never written
    env.P1 = cv::Mat::zeros(3, 4, CV_64FC1);
    for (int i = 0; i < 3; ++i) { env.P0.at<double>(i, i) = 1; env.P1.at<double>(i, i) = 1; }

    const long long n = (long long)triangulate(env);

    std::ofstream o(argv[3], std::ios::binary);
    o.write((const char*)&n, 8);
    const size_t np = env.mesh->pImpl->size();
    std::vector<unsigned char> valid(np), grey(np);
    std::vector<double> xyz(np * 3);
    for (size_t i = 0; i < np; ++i) {
        const auto& p = env.mesh->pImpl->PTc(i);
        valid[i] = p.valid ? 1 : 0;
        xyz[3 * i] = p.p3d[0]; xyz[3 * i + 1] = p.p3d[1]; xyz[3 * i + 2] = p.p3d[2];
        grey[i] = p.valid ? p.color[0] : 0;
    }
    o.write((const char*)valid.data(), np);
    o.write((const char*)xyz.data(), np * 24);
    o.write((const char*)grey.data(), np);
    return o ? 0 : 1;
}
