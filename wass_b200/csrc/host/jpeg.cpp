// Baseline JPEG writer (ITU-T T.81 sequential DCT, Huffman tables of Annex K, JFIF header) for the diagnostic images
// wass_stereo leaves in a workdir (stereo.jpg, stereo_input.jpg, disparity_*.jpg, graph_components.jpg, ...;
// src/wass_stereo/wass_stereo.cpp:833,854,1001,1017,1925, PovMesh.cpp:984).  The reference writes them through
// cv::imwrite (libjpeg, quality 95); there is no libjpeg in this image, and their content is informative only.
// Grey images have one component, colour images three (YCbCr, no chroma subsampling); all components use the two
// luminance Huffman tables -- legal, and the few percent of file size it costs do not matter here.
#include "io.hpp"

#include <cmath>
#include <cstring>

namespace wasshost {
namespace {

const unsigned char ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                                  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const unsigned char QLUM[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                                18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const unsigned char QCHR[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const unsigned char DC_BITS[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const unsigned char DC_VALS[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const unsigned char AC_BITS[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const unsigned char AC_VALS[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
    0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
    0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

struct Huff { unsigned short code[256]; unsigned char len[256]; };
void build(const unsigned char* bits, const unsigned char* vals, Huff& h)
{
    memset(&h, 0, sizeof h);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l - 1]; ++i) { h.code[vals[k]] = (unsigned short)code++; h.len[vals[k]] = (unsigned char)l; ++k; }
        code <<= 1;
    }
}

struct BitWriter {
    std::vector<unsigned char>& o; unsigned acc = 0; int n = 0;
    explicit BitWriter(std::vector<unsigned char>& out) : o(out) {}
    void put(unsigned code, int len)
    {
        acc = (acc << len) | (code & ((1u << len) - 1)); n += len;
        while (n >= 8) { const unsigned char b = (unsigned char)(acc >> (n - 8)); o.push_back(b); if (b == 0xFF) o.push_back(0); n -= 8; }
    }
    void flush() { if (n) put(0x7F, 8 - n); }
};

// 8-point forward DCT after Arai, Agui and Nakajima (5 multiplications; the outputs carry the factors AAN[k], which go into
// the quantisation divisors), rows then columns, in place.
const float AAN[8] = {1.0f, 1.387039845f, 1.306562965f, 1.175875602f, 1.0f, 0.785694958f, 0.541196100f, 0.275899379f};
inline void fdct8(float* d, int st)
{
    const float t0 = d[0] + d[7 * st], t7 = d[0] - d[7 * st], t1 = d[st] + d[6 * st], t6 = d[st] - d[6 * st];
    const float t2 = d[2 * st] + d[5 * st], t5 = d[2 * st] - d[5 * st], t3 = d[3 * st] + d[4 * st], t4 = d[3 * st] - d[4 * st];
    const float e0 = t0 + t3, e3 = t0 - t3, e1 = t1 + t2, e2 = t1 - t2;
    d[0] = e0 + e1; d[4 * st] = e0 - e1;
    const float z1 = (e2 + e3) * 0.707106781f;
    d[2 * st] = e3 + z1; d[6 * st] = e3 - z1;
    const float o0 = t4 + t5, o1 = t5 + t6, o2 = t6 + t7;
    const float z5 = (o0 - o2) * 0.382683433f, z2 = 0.541196100f * o0 + z5, z4 = 1.306562965f * o2 + z5, z3 = o1 * 0.707106781f;
    const float z11 = t7 + z3, z13 = t7 - z3;
    d[5 * st] = z13 + z2; d[3 * st] = z13 - z2; d[st] = z11 + z4; d[7 * st] = z11 - z4;
}
inline void fdct8x8(float* b)
{
    for (int y = 0; y < 8; ++y) fdct8(b + 8 * y, 1);
    for (int x = 0; x < 8; ++x) fdct8(b + x, 8);
}

void put16(std::vector<unsigned char>& o, int v) { o.push_back((unsigned char)(v >> 8)); o.push_back((unsigned char)v); }

}  // namespace

// px: rows x cols x channels (1 = grey, 3 = R,G,B interleaved), 8 bit.
bool write_jpeg(const std::string& path, const unsigned char* px, int rows, int cols, int channels, int quality)
{
    if (!px || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3) || rows > 65535 || cols > 65535) return false;
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    unsigned char q[2][64];
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) { int v = ((t ? QCHR[i] : QLUM[i]) * scale + 50) / 100; q[t][i] = (unsigned char)(v < 1 ? 1 : (v > 255 ? 255 : v)); }
    float rq[2][64];                          // 1 / (quantiser * AAN scale of the coefficient), in zig-zag order
    for (int t = 0; t < 2; ++t)
        for (int i = 0; i < 64; ++i) { const int n = ZIGZAG[i]; rq[t][i] = 1.f / ((float)q[t][n] * AAN[n >> 3] * AAN[n & 7] * 8.f); }
    Huff dc, ac;
    build(DC_BITS, DC_VALS, dc);
    build(AC_BITS, AC_VALS, ac);
    std::vector<unsigned char> o;
    o.reserve((size_t)rows * cols * channels / 4 + 1024);
    const unsigned char soi_app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    o.insert(o.end(), soi_app0, soi_app0 + sizeof soi_app0);
    for (int t = 0; t < (channels == 3 ? 2 : 1); ++t) {
        o.push_back(0xFF); o.push_back(0xDB); put16(o, 67); o.push_back((unsigned char)t);
        for (int i = 0; i < 64; ++i) o.push_back(q[t][ZIGZAG[i]]);
    }
    o.push_back(0xFF); o.push_back(0xC0); put16(o, 8 + 3 * channels); o.push_back(8); put16(o, rows); put16(o, cols); o.push_back((unsigned char)channels);
    for (int c = 0; c < channels; ++c) { o.push_back((unsigned char)(c + 1)); o.push_back(0x11); o.push_back((unsigned char)(c ? 1 : 0)); }
    o.push_back(0xFF); o.push_back(0xC4); put16(o, 2 + 1 + 16 + 12); o.push_back(0x00); o.insert(o.end(), DC_BITS, DC_BITS + 16); o.insert(o.end(), DC_VALS, DC_VALS + 12);
    o.push_back(0xFF); o.push_back(0xC4); put16(o, 2 + 1 + 16 + 162); o.push_back(0x10); o.insert(o.end(), AC_BITS, AC_BITS + 16); o.insert(o.end(), AC_VALS, AC_VALS + 162);
    o.push_back(0xFF); o.push_back(0xDA); put16(o, 6 + 2 * channels); o.push_back((unsigned char)channels);
    for (int c = 0; c < channels; ++c) { o.push_back((unsigned char)(c + 1)); o.push_back(0x00); }
    o.push_back(0); o.push_back(63); o.push_back(0);
    BitWriter bw(o);
    int pred[3] = {0, 0, 0};
    float blk[3][64];
    for (int by = 0; by < rows; by += 8)
        for (int bx = 0; bx < cols; bx += 8) {
            if (channels == 1 && by + 8 <= rows && bx + 8 <= cols) {
                for (int y = 0; y < 8; ++y) {
                    const unsigned char* p = px + (size_t)(by + y) * cols + bx;
                    for (int x = 0; x < 8; ++x) blk[0][y * 8 + x] = (float)p[x] - 128.f;
                }
            } else
            for (int y = 0; y < 8; ++y)
                for (int x = 0; x < 8; ++x) {
                    const int yy = by + y < rows ? by + y : rows - 1, xx = bx + x < cols ? bx + x : cols - 1;
                    const unsigned char* p = px + ((size_t)yy * cols + xx) * channels;
                    if (channels == 1) {
                        blk[0][y * 8 + x] = (float)p[0] - 128.f;
                    } else {
                        const float r = p[0], g = p[1], b = p[2];
                        blk[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.f;
                        blk[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                        blk[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                    }
                }
            for (int c = 0; c < channels; ++c) {
                fdct8x8(blk[c]);
                int zz[64];
                const float* r = rq[c ? 1 : 0];
                for (int i = 0; i < 64; ++i) { const float v = blk[c][ZIGZAG[i]] * r[i]; zz[i] = (int)(v + (v < 0 ? -0.5f : 0.5f)); }
                int diff = zz[0] - pred[c];
                pred[c] = zz[0];
                int a = diff < 0 ? -diff : diff, nb = 0;
                while (a) { ++nb; a >>= 1; }
                bw.put(dc.code[nb], dc.len[nb]);
                if (nb) bw.put((unsigned)(diff < 0 ? diff - 1 : diff), nb);
                int run = 0;
                for (int i = 1; i < 64; ++i) {
                    if (zz[i] == 0) { ++run; continue; }
                    while (run > 15) { bw.put(ac.code[0xF0], ac.len[0xF0]); run -= 16; }
                    int v = zz[i], m = v < 0 ? -v : v, s = 0;
                    while (m) { ++s; m >>= 1; }
                    bw.put(ac.code[(run << 4) | s], ac.len[(run << 4) | s]);
                    bw.put((unsigned)(v < 0 ? v - 1 : v), s);
                    run = 0;
                }
                if (run) bw.put(ac.code[0], ac.len[0]);
            }
        }
    bw.flush();
    o.push_back(0xFF); o.push_back(0xD9);
    return write_file(path, o.data(), o.size());
}

}  // namespace wasshost
