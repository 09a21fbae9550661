// File formats at the workdir boundary of wass_stereo: OpenCV FileStorage XML
// matrices in, 8-bit PNG in/out, "%.16e" text matrices out.  No OpenCV / libpng: zlib only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace wasshost {

struct Mat {            // row-major double matrix
    int rows = 0, cols = 0;
    std::vector<double> v;
    double& at(int r, int c) { return v[(size_t)r * cols + c]; }
    double at(int r, int c) const { return v[(size_t)r * cols + c]; }
    bool empty() const { return v.empty(); }
};

struct Image8 {         // 8-bit grey
    int rows = 0, cols = 0;
    std::vector<uint8_t> px;
    bool empty() const { return px.empty(); }
};

// WASS::load_matrix (src/include/utils.hpp:31-66): first top-level node of an OpenCV XML FileStorage
bool load_matrix_xml(const std::string& path, Mat& out, std::string* err);
// WASS::save_matrix_txt<double> (src/include/utils.hpp:69-92): "%.16e", one space, rows by '\n', no trailing newline
bool save_matrix_txt(const std::string& path, const Mat& m);
// cv::imread(..., IMREAD_GRAYSCALE) for PNG: 1/2/4/8/16-bit grey, grey+alpha, RGB(A) 8/16-bit, palette, plain or Adam7-interlaced -> 8-bit grey
bool read_png_gray(const std::string& path, Image8& out, std::string* err);
bool write_png_gray(const std::string& path, const Image8& img);
bool write_file(const std::string& path, const void* data, size_t n);
// cv::imwrite(name.jpg, img): baseline JPEG (jpeg.cpp); px = rows x cols x channels (1 grey, 3 RGB), quality as OpenCV's default 95
bool write_jpeg(const std::string& path, const unsigned char* px, int rows, int cols, int channels, int quality = 95);

}  // namespace wasshost
