// Test helper (compiled by tests/test_host_io.py): the host executable's own PNG reader / JPEG writer on files.
//   io_probe png <in.png> <out.raw>      -> "rows cols" on stdout, 8-bit grey pixels in out.raw
//   io_probe jpeg <in.raw> rows cols channels <out.jpg>
#include "io.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>

int main(int argc, char** argv)
{
    using namespace wasshost;
    if (argc == 4 && !strcmp(argv[1], "png")) {
        Image8 im; std::string err;
        if (!read_png_gray(argv[2], im, &err)) { std::cerr << err << std::endl; return 1; }
        std::cout << im.rows << " " << im.cols << std::endl;
        return write_file(argv[3], im.px.data(), im.px.size()) ? 0 : 1;
    }
    if (argc == 7 && !strcmp(argv[1], "jpeg")) {
        std::ifstream f(argv[2], std::ios::binary);
        std::vector<unsigned char> px((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const int rows = atoi(argv[3]), cols = atoi(argv[4]), ch = atoi(argv[5]);
        if (px.size() != (size_t)rows * cols * ch) return 2;
        return write_jpeg(argv[6], px.data(), rows, cols, ch) ? 0 : 1;
    }
    return 64;
}
