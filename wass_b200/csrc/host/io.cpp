#include "io.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace wasshost {

static bool slurp(const std::string& path, std::string& out)
{
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.is_open()) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

bool write_file(const std::string& path, const void* data, size_t n)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(data, 1, n, f) == n;
    return fclose(f) == 0 && ok;
}

// ---- OpenCV FileStorage XML: <opencv_storage><name type_id="opencv-matrix"><rows>..<cols>..<dt>d</dt><data>..</data>
static bool tag_text(const std::string& s, size_t from, const char* tag, std::string& text, size_t* end)
{
    const std::string open = std::string("<") + tag + ">", close = std::string("</") + tag + ">";
    const size_t a = s.find(open, from);
    if (a == std::string::npos) return false;
    const size_t b = s.find(close, a);
    if (b == std::string::npos) return false;
    text = s.substr(a + open.size(), b - a - open.size());
    if (end) *end = b + close.size();
    return true;
}

bool load_matrix_xml(const std::string& path, Mat& out, std::string* err)
{
    std::string s;
    if (!slurp(path, s)) { if (err) *err = "Unable to load " + path; return false; }
    const size_t root = s.find("<opencv_storage>");
    if (root == std::string::npos) { if (err) *err = path + ": not an OpenCV XML storage"; return false; }
    const size_t node = s.find("type_id=\"opencv-matrix\"", root);
    if (node == std::string::npos) { if (err) *err = path + ": first node is not a matrix"; return false; }
    std::string rows, cols, dt, data;
    size_t pos = node;
    if (!tag_text(s, pos, "rows", rows, nullptr) || !tag_text(s, pos, "cols", cols, nullptr) || !tag_text(s, pos, "dt", dt, nullptr) ||
        !tag_text(s, pos, "data", data, nullptr)) {
        if (err) *err = path + ": malformed matrix node";
        return false;
    }
    out.rows = atoi(rows.c_str());
    out.cols = atoi(cols.c_str());
    if (out.rows <= 0 || out.cols <= 0) { if (err) *err = path + ": bad matrix size"; return false; }
    out.v.clear();
    std::stringstream ds(data);
    std::string tok;
    while (ds >> tok) {
        // OpenCV writes ".Inf" / "-.Inf" / ".Nan" for non-finite values
        if (tok == ".Inf") out.v.push_back(INFINITY);
        else if (tok == "-.Inf") out.v.push_back(-INFINITY);
        else if (tok == ".Nan") out.v.push_back(NAN);
        else out.v.push_back(strtod(tok.c_str(), nullptr));
    }
    if ((int)out.v.size() != out.rows * out.cols) { if (err) *err = path + ": element count does not match rows*cols"; return false; }
    return true;
}

bool save_matrix_txt(const std::string& path, const Mat& m)
{
    std::ofstream ofs(path.c_str());
    if (ofs.fail()) return false;
    ofs.precision(16);
    ofs << std::scientific;
    for (int i = 0; i < m.rows; ++i) {
        for (int j = 0; j < m.cols; ++j) {
            ofs << m.at(i, j);
            if (j != m.cols - 1) ofs << " ";
        }
        if (i != m.rows - 1) ofs << std::endl;
    }
    ofs.close();
    return true;
}

// ---- PNG ------------------------------------------------------------------------------------------
static uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// RGB -> grey as cv::imread(IMREAD_GRAYSCALE) gets it for PNG files: not cv::cvtColor but libpng's own
// png_set_rgb_to_gray(1, 0.299, 0.587) (modules/imgcodecs/src/grfmt_png.cpp): 15-bit coefficients 9797 / 19234 / 3737,
// applied to the samples at their own bit depth (pinned against cv2 in tests/test_host_io.py).
// libpng truncates the 8-bit result ("the historical approach") and rounds the 16-bit one.
static inline int png_rgb_to_gray(int r, int g, int b, unsigned round = 0)
{
    if (r == g && g == b) return r;
    return (int)(((unsigned)r * 9797u + (unsigned)g * 19234u + (unsigned)b * 3737u + round) >> 15);
}

bool read_png_gray(const std::string& path, Image8& out, std::string* err)
{
    std::string s;
    if (!slurp(path, s)) { if (err) *err = "unable to open " + path; return false; }
    const unsigned char* d = (const unsigned char*)s.data();
    static const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (s.size() < 33 || memcmp(d, sig, 8)) { if (err) *err = path + ": not a PNG file"; return false; }
    size_t pos = 8;
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte;
    while (pos + 12 <= s.size()) {
        const uint32_t len = be32(d + pos);
        const std::string type((const char*)d + pos + 4, 4);
        const unsigned char* body = d + pos + 8;
        if (pos + 12 + len > s.size()) break;
        if (type == "IHDR") { w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
        else if (type == "PLTE") plte.assign(body, body + len);
        else if (type == "IDAT") idat.insert(idat.end(), body, body + len);
        else if (type == "IEND") break;
        pos += 12 + len;
    }
    const bool sub_byte = depth == 1 || depth == 2 || depth == 4;       // legal for grey and palette images only
    if (w <= 0 || h <= 0 || interlace > 1 || (depth != 8 && depth != 16 && !(sub_byte && (ctype == 0 || ctype == 3)))) {
        if (err) *err = path + ": unsupported PNG (need 1/2/4/8/16 bits per sample, no or Adam7 interlace)";
        return false;
    }
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch || (ctype == 3 && depth == 16)) { if (err) *err = path + ": unsupported PNG colour type"; return false; }
    const int bpp = sub_byte ? 1 : ch * depth / 8;                       // filter distance in bytes
    auto row_bytes = [&](int pw) { return sub_byte ? ((size_t)pw * depth + 7) / 8 : (size_t)pw * bpp; };
    // the image is one pass, or the seven passes of Adam7: reduced images on the grids (x0 + i dx, y0 + j dy), each with its
    // own scanlines and filter history
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole[1] = {{0, 0, 1, 1}};
    const Pass* passes = interlace ? adam7 : whole;
    const int npass = interlace ? 7 : 1;
    size_t total = 0;
    for (int k = 0; k < npass; ++k) {
        const int pw = (w - passes[k].x0 + passes[k].dx - 1) / passes[k].dx, ph = (h - passes[k].y0 + passes[k].dy - 1) / passes[k].dy;
        if (pw > 0 && ph > 0) total += (row_bytes(pw) + 1) * ph;
    }
    std::vector<unsigned char> raw(total);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) {
        if (err) *err = path + ": zlib inflate failed";
        return false;
    }
    // 8-bit grey of pixel x of an unfiltered scanline.  Sub-byte samples are packed MSB first; grey levels are scaled to
    // 8 bits by bit replication (libpng's expansion, which cv::imread relies on), palette indices go through PLTE;
    // 16-bit samples are big-endian and the 8-bit result keeps the high byte (libpng's strip_16).
    auto pal = [&](int v) {
        const int k = v * 3;
        if (k + 2 >= (int)plte.size()) return 0;
        return png_rgb_to_gray(plte[k], plte[k + 1], plte[k + 2]);
    };
    auto grey_at = [&](const unsigned char* row, int x) -> int {
        if (sub_byte) {
            const int per = 8 / depth, mask = (1 << depth) - 1;
            const int v = (row[x / per] >> ((per - 1 - x % per) * depth)) & mask;
            return ctype == 0 ? v * (255 / mask) : pal(v);
        }
        const unsigned char* p = row + (size_t)x * bpp;
        if (ctype == 0 || ctype == 4) return p[0];
        if (ctype == 3) return pal(p[0]);
        if (depth == 8) return png_rgb_to_gray(p[0], p[1], p[2]);
        return png_rgb_to_gray(p[0] << 8 | p[1], p[2] << 8 | p[3], p[4] << 8 | p[5], 16384u) >> 8;
    };
    out.rows = h; out.cols = w; out.px.assign((size_t)w * h, 0);
    const unsigned char* in = raw.data();
    std::vector<unsigned char> cur, up;
    for (int k = 0; k < npass; ++k) {
        const Pass& ps = passes[k];
        const int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = row_bytes(pw);
        cur.assign(stride, 0); up.assign(stride, 0);
        for (int y = 0; y < ph; ++y) {
            const int ft = in[0];
            for (size_t x = 0; x < stride; ++x) {
                const int a = x >= (size_t)bpp ? cur[x - bpp] : 0, bb = up[x], c = x >= (size_t)bpp ? up[x - bpp] : 0;
                int v = in[1 + x];
                switch (ft) {
                    case 1: v += a; break;
                    case 2: v += bb; break;
                    case 3: v += (a + bb) / 2; break;
                    case 4: { const int q = a + bb - c, pa = abs(q - a), pb = abs(q - bb), pc = abs(q - c); v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? bb : c); break; }
                    default: break;
                }
                cur[x] = (unsigned char)v;
            }
            uint8_t* orow = &out.px[(size_t)(ps.y0 + y * ps.dy) * w];
            for (int x = 0; x < pw; ++x) orow[ps.x0 + x * ps.dx] = (uint8_t)grey_at(cur.data(), x);
            in += stride + 1;
            cur.swap(up);          // this row becomes the prior one (the first row of a pass sees zeros)
        }
    }
    return true;
}

static void put_chunk(std::vector<unsigned char>& o, const char* type, const unsigned char* data, size_t n)
{
    const uint32_t len = (uint32_t)n;
    unsigned char hdr[8] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len,
                            (unsigned char)type[0], (unsigned char)type[1], (unsigned char)type[2], (unsigned char)type[3]};
    o.insert(o.end(), hdr, hdr + 8);
    if (n) o.insert(o.end(), data, data + n);
    uLong crc = crc32(0L, hdr + 4, 4);
    if (n) crc = crc32(crc, data, (uInt)n);
    const unsigned char c[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    o.insert(o.end(), c, c + 4);
}

bool write_png_gray(const std::string& path, const Image8& img)
{
    if (img.empty()) return false;
    std::vector<unsigned char> raw((size_t)(img.cols + 1) * img.rows);
    for (int y = 0; y < img.rows; ++y) {
        raw[(size_t)(img.cols + 1) * y] = 0;
        memcpy(&raw[(size_t)(img.cols + 1) * y + 1], &img.px[(size_t)img.cols * y], img.cols);
    }
    uLongf clen = compressBound(raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 3) != Z_OK) return false;
    std::vector<unsigned char> o = {137, 80, 78, 71, 13, 10, 26, 10};
    unsigned char ihdr[13] = {(unsigned char)(img.cols >> 24), (unsigned char)(img.cols >> 16), (unsigned char)(img.cols >> 8), (unsigned char)img.cols,
                              (unsigned char)(img.rows >> 24), (unsigned char)(img.rows >> 16), (unsigned char)(img.rows >> 8), (unsigned char)img.rows,
                              8, 0, 0, 0, 0};
    put_chunk(o, "IHDR", ihdr, 13);
    put_chunk(o, "IDAT", comp.data(), clen);
    put_chunk(o, "IEND", nullptr, 0);
    return write_file(path, o.data(), o.size());
}

}  // namespace wasshost
