// Stages of wass_stereo after the matcher, as sm_100a kernels (compiled with -fmad=false so that the
// fp32/fp64 operation order of the reference is reproduced literally):
//   disparity clean-up      src/wass_stereo/wass_stereo.cpp:617-733, 853-928
//   triangulation           src/wass_stereo/wass_stereo.cpp:299-324, 1039-1386; src/wass_lib/triangulate.hpp:26-72
//   PovMesh                 src/wass_stereo/PovMesh.cpp (zgap percentile, biggest component, RANSAC, crop, refine, export)
#include "geom.cuh"

#include <cub/cub.cuh>
#include <cfloat>
#include <climits>

namespace wsg {

// ------------------------------------------------------------------------------------------------
// zero padding of the two crops (wass_stereo.cpp:820-831): img1 = padded RIGHT, img2 = padded LEFT
// ------------------------------------------------------------------------------------------------
__global__ void pad_kernel(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right, size_t stride, int rows,
                           int cols, int ndisp, int off, int comp, uint8_t* __restrict__ img1, uint8_t* __restrict__ img2, int wp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wp) return;
    const int xr = x - ndisp;
    const int xl = x - (ndisp + off - comp);
    img1[(size_t)y * wp + x] = (xr >= 0 && xr < cols) ? right[(size_t)y * stride + xr] : 0;
    img2[(size_t)y * wp + x] = (xl >= 0 && xl < cols) ? left[(size_t)y * stride + xl] : 0;
}
void launch_pad_images(const uint8_t* left, const uint8_t* right, size_t stride, int rows, int cols, int ndisp,
                       int off, int comp, uint8_t* img1, uint8_t* img2, int wp, cudaStream_t st)
{
    dim3 b(256), g((wp + 255) / 256, rows);
    pad_kernel<<<g, b, 0, st>>>(left, right, stride, rows, cols, ndisp, off, comp, img1, img2, wp);
}

// ------------------------------------------------------------------------------------------------
// clean_and_convert_disparity (wass_stereo.cpp:714-733) on the column range [x0, x0+width)
// ------------------------------------------------------------------------------------------------
__global__ void clean_convert_kernel(const int16_t* __restrict__ disp16, int rows, int cols_full, int x0, int width,
                                     int mindisp, int ndisp, int disp_offset, double scale, float* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= width) return;
    float dval = __fdiv_rn((float)disp16[(size_t)i * cols_full + x0 + j], 16.0f);
    float r = 0.f;
    if (!(dval <= (float)mindisp || dval > (float)ndisp)) {
        dval = __fadd_rn(dval, (float)disp_offset);
        r = (float)__dmul_rn((double)dval, scale);
    }
    out[(size_t)i * width + j] = r;
}
void launch_clean_convert(const int16_t* disp16, int rows, int cols_full, int x0, int width, int mindisp, int ndisp,
                          int disp_offset, double scale, float* out, cudaStream_t st)
{
    dim3 b(256), g((width + 255) / 256, rows);
    clean_convert_kernel<<<g, b, 0, st>>>(disp16, rows, cols_full, x0, width, mindisp, ndisp, disp_offset, scale, out);
}

// matrix_dilate_zero<float> (wass_stereo.cpp:617-662) with its one-column output shift
__global__ void dilate_zero_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (k >= cols) return;
    const size_t o = (size_t)i * cols + k;
    float v = src[o];
    if (i >= 1 && i <= rows - 2 && k <= cols - 3 && v == 0.f) {
        const float* t = src + (size_t)(i - 1) * cols;
        const float* b = src + (size_t)(i + 1) * cols;
        const float* c = src + (size_t)i * cols;
        // reference order: tm1, tp1, t, bm1, bp1, b, cm1, cp1 (columns relative to the centre k+1)
        const float nb[8] = {t[k], t[k + 2], t[k + 1], b[k], b[k + 2], b[k + 1], c[k], c[k + 2]};
        float avg = 0.f; int num = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (nb[q] > 0.f) { avg = __fadd_rn(avg, nb[q]); ++num; }
        if (num > 1) v = __fdiv_rn(avg, (float)num);
    }
    dst[o] = v;
}
void launch_dilate_zero(const float* src, float* dst, int rows, int cols, cudaStream_t st)
{
    dim3 b(256), g((cols + 255) / 256, rows);
    dilate_zero_kernel<<<g, b, 0, st>>>(src, dst, rows, cols);
}

// matrix_erode_zero<float> (wass_stereo.cpp:665-711)
__global__ void erode_zero_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= cols) return;
    const size_t o = (size_t)i * cols + j;
    float v = src[o];
    if (i == 0 || i == rows - 1 || j == 0 || j == cols - 1) {
        v = 0.f;
    } else {
        const float* t = src + (size_t)(i - 1) * cols + j;
        const float* b = src + (size_t)(i + 1) * cols + j;
        const float* c = src + (size_t)i * cols + j;
        if (t[0] == 0.f || t[-1] == 0.f || t[1] == 0.f || b[0] == 0.f || b[-1] == 0.f || b[1] == 0.f || c[-1] == 0.f || c[1] == 0.f)
            v = 0.f;
    }
    dst[o] = v;
}
void launch_erode_zero(const float* src, float* dst, int rows, int cols, cudaStream_t st)
{
    dim3 b(256), g((cols + 255) / 256, rows);
    erode_zero_kernel<<<g, b, 0, st>>>(src, dst, rows, cols);
}
// wass_stereo.cpp:903-928 at DENSE_SCALE==1: zero the pixels that one more erosion would zero == one more erosion
void launch_mask_by_eroded(const float* src, float* dst, int rows, int cols, cudaStream_t st)
{
    launch_erode_zero(src, dst, rows, cols, st);
}

__global__ void paste_kernel(const float* __restrict__ roi, int rh, int rw, float* __restrict__ full, int rows, int cols, int x0, int y0)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int u = x - x0, v = y - y0;
    full[(size_t)y * cols + x] = (u >= 0 && u < rw && v >= 0 && v < rh) ? roi[(size_t)v * rw + u] : 0.f;
}
void launch_paste_roi(const float* roi, int rh, int rw, float* full, int rows, int cols, int x0, int y0, cudaStream_t st)
{
    dim3 b(256), g((cols + 255) / 256, rows);
    paste_kernel<<<g, b, 0, st>>>(roi, rh, rw, full, rows, cols, x0, y0);
}

// ------------------------------------------------------------------------------------------------
// triangulation: one thread per ROI pixel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unrectify_dev(double u, double v, const double* K, const double* Rr, double fx, double fy,
                                              double cx, double cy, double& ox, double& oy)
{
    // wass_stereo.cpp:313-322 : xyw = Rrect^T * ((u-cx)/fx, (v-cy)/fy, 1); project with the original intrinsics
    const double x = (u - cx) / fx, y = (v - cy) / fy;
    double a = Rr[0] * x + Rr[3] * y + Rr[6] * 1.0;
    double b = Rr[1] * x + Rr[4] * y + Rr[7] * 1.0;
    const double c = Rr[2] * x + Rr[5] * y + Rr[8] * 1.0;
    a /= c; b /= c;
    ox = a * K[0] + K[2];
    oy = b * K[4] + K[5];
}

__global__ void triangulate_kernel(const float* __restrict__ disparity, const uint8_t* __restrict__ left,
                                   const uint8_t* __restrict__ right, const uint8_t* __restrict__ lmask,
                                   const uint8_t* __restrict__ rmask, CalibDev c, MeshView m, unsigned long long* counter)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= m.w) return;
    const int xr = c.rrx + u, yr = c.rry + v;
    const float dv = disparity[(size_t)yr * c.rect_cols + xr];
    if (!(dv > 1.0f)) return;
    // wass_stereo.cpp:1180-1181 : int - float in float, then + double, then cast to float
    float xl = (float)((double)((float)(xr - c.rrx + c.rlx) - dv) + c.comp_over_scale);
    const float yl = (float)yr;
    if (xl < 0.f || xl >= (float)c.rect_cols) return;
    double pix, piy, qix, qiy;
    if (c.use_h) {
        // wass_stereo.cpp:301-305 : uvr = H^-1 * (u, v, 1) (cv::Matx product: ((0 + a0 u) + a1 v) + a2), then / uvr[2]
        auto hom = [](const double* Hi, double uu, double vv, double& ox, double& oy) {
            const double a = Hi[0] * uu + Hi[1] * vv + Hi[2] * 1.0, b = Hi[3] * uu + Hi[4] * vv + Hi[5] * 1.0;
            const double w = Hi[6] * uu + Hi[7] * vv + Hi[8] * 1.0;
            ox = a / w; oy = b / w;
        };
        hom(c.HLi, (double)xl, (double)yl, pix, piy);
        hom(c.HRi, (double)xr, (double)yr, qix, qiy);
    } else {
        unrectify_dev((double)xl, (double)yl, c.K0, c.R1, c.P1fx, c.P1fy, c.P1cx, c.P1cy, pix, piy);
        unrectify_dev((double)xr, (double)yr, c.K1, c.R2, c.P2fx, c.P2fy, c.P2cx, c.P2cy, qix, qiy);
    }
    if (pix < 1 || pix >= c.left_cols - 1 || piy < 1 || piy >= c.left_rows - 1 || qix < 1 || qix >= c.right_cols - 1 ||
        qiy < 1 || qiy >= c.right_rows - 1)
        return;
    const double px = (pix - c.K0[2]) / c.K0[0], py = (piy - c.K0[5]) / c.K0[4];
    const double qx = (qix - c.K1[2]) / c.K1[0], qy = (qiy - c.K1[5]) / c.K1[4];
    if (pix <= c.bbox_l || piy <= c.bbox_t || pix >= c.bbox_r || piy >= c.bbox_b) return;
    {
        const size_t li = (size_t)(int)piy * c.left_cols + (int)pix;
        const size_t ri = (size_t)(int)qiy * c.right_cols + (int)qix;
        if (c.has_lmask && lmask[li] == 0) return;
        if (c.has_rmask && rmask[ri] == 0) return;
        if (c.discard_burned && (left[li] > 254 || right[ri] > 254)) return;
    }
    if (c.min_angle > 0) {
        const double n1 = sqrt(px * px + py * py + 1.0 * 1.0);
        const double s1 = n1 != 0 ? 1.0 / n1 : 0.0;
        const double b0 = c.R[0] * qx + c.R[1] * qy + c.R[2] * 1.0 + c.T[0];
        const double b1 = c.R[3] * qx + c.R[4] * qy + c.R[5] * 1.0 + c.T[1];
        const double b2 = c.R[6] * qx + c.R[7] * qy + c.R[8] * 1.0 + c.T[2];
        const double n2 = sqrt(b0 * b0 + b1 * b1 + b2 * b2);
        const double s2 = n2 != 0 ? 1.0 / n2 : 0.0;
        const double dot = (px * s1) * (b0 * s2) + (py * s1) * (b1 * s2) + (1.0 * s1) * (b2 * s2);
        const double ang = fabs(acos(dot) * 57.29577951);
        if (ang < c.min_angle) return;
    }
    // triangulate.hpp:26-72
    const double* R = c.R;
    double Af[12], Bf[4];
    Af[0] = -1.0; Af[1] = 0.0; Af[2] = px;
    Af[3] = 0.0; Af[4] = -1.0; Af[5] = py;
    Af[6] = qx * R[6] - R[0]; Af[7] = qx * R[7] - R[1]; Af[8] = qx * R[8] - R[2];
    Af[9] = qy * R[6] - R[3]; Af[10] = qy * R[7] - R[4]; Af[11] = qy * R[8] - R[5];
    Bf[0] = 0.0; Bf[1] = 0.0; Bf[2] = c.T[0] - c.T[2] * qx; Bf[3] = c.T[1] - c.T[2] * qy;
    double A[9], b[3];
    A[0] = Af[0] * Af[0] + Af[3] * Af[3] + Af[6] * Af[6] + Af[9] * Af[9];
    A[1] = Af[0] * Af[1] + Af[3] * Af[4] + Af[10] * Af[9] + Af[6] * Af[7];
    A[2] = Af[0] * Af[2] + Af[3] * Af[5] + Af[11] * Af[9] + Af[6] * Af[8];
    A[3] = A[1];
    A[4] = Af[1] * Af[1] + Af[10] * Af[10] + Af[4] * Af[4] + Af[7] * Af[7];
    A[5] = Af[10] * Af[11] + Af[1] * Af[2] + Af[4] * Af[5] + Af[7] * Af[8];
    A[6] = A[2]; A[7] = A[5];
    A[8] = Af[11] * Af[11] + Af[2] * Af[2] + Af[5] * Af[5] + Af[8] * Af[8];
    b[0] = Af[0] * Bf[0] + Af[3] * Bf[1] + Af[6] * Bf[2] + Af[9] * Bf[3];
    b[1] = Af[1] * Bf[0] + Af[10] * Bf[3] + Af[4] * Bf[1] + Af[7] * Bf[2];
    b[2] = Af[2] * Bf[0] + Af[11] * Bf[3] + Af[5] * Bf[1] + Af[8] * Bf[2];
    // cv::solve(DECOMP_LU) on 3x3 == closed form with the determinant (OpenCV fast path)
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
    double X0 = 0, X1 = 0, X2 = 0;
    if (det != 0.0) {
        const double d = 1.0 / det;
        X0 = d * (b[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (b[1] * A[8] - A[5] * b[2]) + A[2] * (b[1] * A[7] - A[4] * b[2]));
        X1 = d * (A[0] * (b[1] * A[8] - A[5] * b[2]) - b[0] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * b[2] - b[1] * A[6]));
        X2 = d * (A[0] * (A[4] * b[2] - b[1] * A[7]) - A[1] * (A[3] * b[2] - b[1] * A[6]) + b[0] * (A[3] * A[7] - A[4] * A[6]));
    }
    const double dist = sqrt(X0 * X0 + X1 * X1 + X2 * X2);
    if (dist < c.cam_distance / 10.0 || X2 < 1.0) return;
    if (dist > c.cam_distance * 200.0 || X2 > 1e30) return;
    const size_t o = (size_t)v * m.w + u;
    m.valid[o] = 1; m.X[o] = X0; m.Y[o] = X1; m.Z[o] = X2;
    m.color[o] = right[(size_t)(int)qiy * c.right_cols + (int)qix];
    atomicAdd(counter, 1ull);
}
void launch_triangulate(const float* disparity, const uint8_t* left, const uint8_t* right, const uint8_t* lmask,
                        const uint8_t* rmask, const CalibDev& c, MeshView m, unsigned long long* counter, cudaStream_t st)
{
    dim3 b(128), g((m.w + 127) / 128, m.h);
    triangulate_kernel<<<g, b, 0, st>>>(disparity, left, right, lmask, rmask, c, m, counter);
}

__global__ void count_valid_kernel(MeshView m, unsigned long long* counter)
{
    const size_t n = (size_t)m.w * m.h;
    unsigned long long loc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) loc += m.valid[i];
    for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
    if ((threadIdx.x & 31) == 0 && loc) atomicAdd(counter, loc);
}
void launch_count_valid(const MeshView& m, unsigned long long* counter, cudaStream_t st)
{
    count_valid_kernel<<<296, 256, 0, st>>>(m, counter);
}

// ------------------------------------------------------------------------------------------------
// compute_zgap_percentile (PovMesh.cpp:888-926): exact order statistic by sorting all gaps
// ------------------------------------------------------------------------------------------------
__global__ void zgap_kernel(MeshView m, double* gaps, unsigned long long* count)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= m.w) return;
    const size_t o = (size_t)i * m.w + j;
    double g0 = DBL_MAX, g1 = DBL_MAX, g2 = DBL_MAX;
    int n = 0;
    if (i >= 1 && j >= 1 && j <= m.w - 2 && m.valid[o]) {
        const double z = m.Z[o];
        const size_t up = o - m.w;
        if (m.valid[up - 1]) { g0 = fabs(z - m.Z[up - 1]); ++n; }
        if (m.valid[up]) { g1 = fabs(z - m.Z[up]); ++n; }
        if (m.valid[up + 1]) { g2 = fabs(z - m.Z[up + 1]); ++n; }
    }
    gaps[3 * o] = g0; gaps[3 * o + 1] = g1; gaps[3 * o + 2] = g2;
    // one atomic per warp, not per point (5 M atomics on one address cost 3 ms)
    n = __reduce_add_sync(__activemask(), n);
    if (n && (threadIdx.x & 31) == __ffs(__activemask()) - 1) atomicAdd(count, (unsigned long long)n);
}
// Exact k-th smallest of the gaps by radix select on the bit patterns (non-negative doubles order like unsigned 64-bit
// integers): six passes of 11 bits (the last one 9), each a privatised shared-memory histogram of the keys that match the
// prefix found so far, then one small kernel that walks the 2048 bins.  Reads the 3*W*H keys six times (~0.1 ms each)
// instead of sorting them (the sort was 4 ms of a 25 ms frame).
struct RsState { unsigned long long k, mask, val, count; int bad; };

__global__ void rs_init_kernel(RsState* s, const unsigned long long* count, double percentile)
{
    const unsigned long long n = *count;
    s->count = n; s->mask = 0; s->val = 0; s->bad = 0;
    if (n == 0) { s->bad = 1; s->k = 0; return; }
    const unsigned long long idx = (unsigned long long)floor(percentile / 100.0 * (double)n);   // PovMesh.cpp:924
    if (idx >= n) { s->bad = 1; s->k = 0; return; }     // the reference reads past the end here (percentile >= 100)
    s->k = idx;
}
__global__ void __launch_bounds__(256) rs_hist_kernel(const unsigned long long* __restrict__ keys, size_t n, const RsState* __restrict__ s,
                                                      int shift, unsigned digit_mask, unsigned* __restrict__ hist)
{
    __shared__ unsigned sh[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sh[i] = 0;
    __syncthreads();
    const unsigned long long mask = s->mask, val = s->val;
    // the high digits of the gaps are nearly all equal: aggregate equal bins inside the warp before touching shared memory
    const size_t nround = (n + 255) / 256 * 256;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nround; i += (size_t)gridDim.x * 256) {
        unsigned bin = 0xFFFFFFFFu;
        if (i < n) {
            const unsigned long long k = keys[i];
            if ((k & mask) == val) bin = (unsigned)(k >> shift) & digit_mask;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 0xFFFFFFFFu && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[bin], (unsigned)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += 256)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void rs_pick_kernel(RsState* s, unsigned* hist, int shift, int bits)
{
    // one warp: find the bin that holds rank k
    __shared__ unsigned long long cum[33];
    const int lane = threadIdx.x, nb = 1 << bits, per = nb / 32;
    unsigned long long mine = 0;
    for (int i = 0; i < per; ++i) mine += hist[lane * per + i];
    cum[lane + 1] = mine;
    __syncwarp();
    if (lane == 0) {
        cum[0] = 0;
        for (int i = 1; i <= 32; ++i) cum[i] += cum[i - 1];
        unsigned long long k = s->k;
        int seg = 0;
        while (seg < 31 && cum[seg + 1] <= k) ++seg;
        unsigned long long before = cum[seg];
        int b = seg * per;
        while (b < seg * per + per - 1 && before + hist[b] <= k) { before += hist[b]; ++b; }
        s->k = k - before;
        s->val |= (unsigned long long)b << shift;
        s->mask |= (unsigned long long)(nb - 1) << shift;
    }
    __syncwarp();
    for (int i = lane; i < 2048; i += 32) hist[i] = 0;     // ready for the next pass
}
__global__ void rs_finish_kernel(const RsState* s, double* out)
{
    *out = s->bad ? __longlong_as_double(0x7ff8000000000000LL) : __longlong_as_double((long long)s->val);
}

size_t zgap_scratch_bytes(int w, int h)
{
    const size_t n = (size_t)3 * w * h;
    return n * sizeof(double) + 2048 * sizeof(unsigned) + 512;
}
int mesh_zgap_percentile(const MeshView& m, double percentile, void* scratch, size_t scratch_bytes, double* out_host, cudaStream_t st)
{
    const size_t n = (size_t)3 * m.w * m.h;
    if (scratch_bytes < zgap_scratch_bytes(m.w, m.h)) return -1;
    double* a = (double*)scratch;
    unsigned long long* cnt = (unsigned long long*)(a + n);          // [0] count, then the select state, then the histogram
    RsState* state = (RsState*)(cnt + 2);
    double* d_out = (double*)(cnt + 10);
    unsigned* hist = (unsigned*)(cnt + 16);
    cudaMemsetAsync(cnt, 0, 128 + 2048 * sizeof(unsigned), st);
    dim3 blk(256), g((m.w + 255) / 256, m.h);
    zgap_kernel<<<g, blk, 0, st>>>(m, a, cnt);
    rs_init_kernel<<<1, 1, 0, st>>>(state, cnt, percentile);
    const int shifts[6] = {53, 42, 31, 20, 9, 0}, bits[6] = {11, 11, 11, 11, 11, 9};
    for (int p = 0; p < 6; ++p) {
        rs_hist_kernel<<<592, 256, 0, st>>>((const unsigned long long*)a, n, state, shifts[p], (1u << bits[p]) - 1u, hist);
        rs_pick_kernel<<<1, 32, 0, st>>>(state, hist, shifts[p], bits[p]);
    }
    rs_finish_kernel<<<1, 1, 0, st>>>(state, d_out);
    cudaMemcpyAsync(out_host, d_out, 8, cudaMemcpyDeviceToHost, st);
    return cudaStreamSynchronize(st) == cudaSuccess ? 0 : -1;
}

// ------------------------------------------------------------------------------------------------
// cluster_biggest_connected_component (PovMesh.cpp:929-987): union-find on 4-connected edges with
// |dz| < zgap; label = smallest column-major index of the component, so ties between equally big
// components resolve to the one the reference's column-major rescan finds first.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(int* L, int x)
{
    // path halving; the shortcut is installed with a compare-and-swap so that it can never undo a concurrent union
    int p = L[x];
    while (p != x) {
        const int g = L[p];
        if (g != p) atomicCAS(&L[x], p, g);
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b)
{
    while (true) {
        a = uf_find(L, a); b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;
    }
}
// Labels are ROW-MAJOR point indices (coalesced with the mesh arrays; the first version indexed them column-major to get
// the reference's tie rule for free and paid for it with one 32-byte sector per label access).  The tie rule -- among
// components of equal size the reference keeps the one its column-major seed scan meets first, i.e. the one owning the
// smallest column-major index -- is restored by a tie-break pass that only does work when two roots share the maximum.
// Two levels: (1) every CCL_TW x CCL_TH tile is labelled on its own in shared memory (union-find on local indices; the root
// of a tile component is its smallest local index, i.e. its smallest row-major global index, which is what the global
// labels hold); (2) only the edges that cross tile borders go through the global union-find.  The first version ran every
// edge of the 5 M point mesh through global atomics (0.97 ms) although nearly all points end up in one component.
static constexpr int CCL_TW = 64, CCL_TH = 16;
__global__ void __launch_bounds__(256) ccl_local_kernel(MeshView m, int* L, double zgap)
{
    __shared__ int sl[CCL_TW * CCL_TH];
    __shared__ double sz[CCL_TW * CCL_TH];
    const int u0 = blockIdx.x * CCL_TW, v0 = blockIdx.y * CCL_TH;
    for (int i = threadIdx.x; i < CCL_TW * CCL_TH; i += blockDim.x) {
        const int u = u0 + i % CCL_TW, v = v0 + i / CCL_TW;
        const bool in = u < m.w && v < m.h && m.valid[(size_t)v * m.w + u];
        sl[i] = in ? i : INT_MAX;
        sz[i] = in ? m.Z[(size_t)v * m.w + u] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CCL_TW * CCL_TH; i += blockDim.x) {
        if (sl[i] == INT_MAX) continue;               // (a valid element's label never becomes INT_MAX)
        const int x = i % CCL_TW, y = i / CCL_TW;
        const double z = sz[i];
        if (x + 1 < CCL_TW && sl[i + 1] != INT_MAX && fabs(z - sz[i + 1]) < zgap) uf_union(sl, i, i + 1);
        if (y + 1 < CCL_TH && sl[i + CCL_TW] != INT_MAX && fabs(z - sz[i + CCL_TW]) < zgap) uf_union(sl, i, i + CCL_TW);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CCL_TW * CCL_TH; i += blockDim.x) {
        const int u = u0 + i % CCL_TW, v = v0 + i / CCL_TW;
        if (u >= m.w || v >= m.h) continue;
        int lab = INT_MAX;
        if (sl[i] != INT_MAX) {
            const int r = uf_find(sl, i);
            lab = (v0 + r / CCL_TW) * m.w + u0 + r % CCL_TW;
        }
        L[(size_t)v * m.w + u] = lab;
    }
}
__global__ void ccl_merge_kernel(MeshView m, int* L, double zgap)      // the edges between tiles
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= m.w) return;
    const bool right = u % CCL_TW == CCL_TW - 1 && u + 1 < m.w, down = v % CCL_TH == CCL_TH - 1 && v + 1 < m.h;
    if (!right && !down) return;
    const size_t o = (size_t)v * m.w + u;
    if (!m.valid[o]) return;
    const double z = m.Z[o];
    if (right && m.valid[o + 1] && fabs(z - m.Z[o + 1]) < zgap) uf_union(L, (int)o, (int)o + 1);
    if (down && m.valid[o + m.w] && fabs(z - m.Z[o + m.w]) < zgap) uf_union(L, (int)o, (int)o + m.w);
}
__global__ void ccl_flatten_count_kernel(MeshView m, int* L, unsigned* cnt)
{
    const size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (o >= (size_t)m.w * m.h || L[o] == INT_MAX) return;
    const int r = uf_find(L, (int)o);
    L[o] = r;
    // nearly every point belongs to one component: count equal roots inside the warp first
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, r);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[r], (unsigned)__popc(peers));
}
// best[0] = max count; best[1] = number of roots with that count; best[2] = smallest such root (row-major);
// best[3] = min over the tied components of (smallest column-major member index << 32 | root), filled only on a tie
__global__ void ccl_max_kernel(const unsigned* cnt, int n, unsigned long long* best)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned c = i < n ? cnt[i] : 0;
    const unsigned wmax = __reduce_max_sync(0xffffffffu, c);
    if (wmax && c == wmax && (threadIdx.x & 31) == __ffs(__ballot_sync(0xffffffffu, c == wmax)) - 1) atomicMax(best, (unsigned long long)wmax);
}
__global__ void ccl_ties_kernel(const unsigned* cnt, int n, unsigned long long* best)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || cnt[i] == 0 || cnt[i] != (unsigned)best[0]) return;
    atomicAdd(best + 1, 1ull);
    atomicMin(best + 2, (unsigned long long)i);
}
__global__ void ccl_tiebreak_kernel(MeshView m, const int* L, const unsigned* cnt, unsigned long long* best)
{
    if (best[1] <= 1) return;                        // the usual case: one biggest component, nothing to decide
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= m.w) return;
    const size_t o = (size_t)v * m.w + u;
    const int r = L[o];
    if (r == INT_MAX || cnt[r] != (unsigned)best[0]) return;
    atomicMin(best + 3, ((unsigned long long)((unsigned)u * (unsigned)m.h + (unsigned)v) << 32) | (unsigned)r);
}
__global__ void ccl_extract_kernel(MeshView m, const int* L, const unsigned long long* best)
{
    const size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (o >= (size_t)m.w * m.h) return;
    const int lab = best[1] <= 1 ? (int)best[2] : (int)(best[3] & 0xFFFFFFFFull);
    if (m.valid[o] && L[o] != lab) m.valid[o] = 0;
}
int mesh_biggest_component(MeshView m, double zgap, int* labels, unsigned long long* scratch, unsigned long long* n_left_host, cudaStream_t st)
{
    const int n = m.w * m.h;
    unsigned long long* best = scratch;                // 4 words, then the per-root counts
    unsigned* cnt = (unsigned*)(scratch + 4);
    cudaMemsetAsync(scratch, 0, 32 + (size_t)n * sizeof(unsigned), st);
    cudaMemsetAsync(scratch + 2, 0xff, 16, st);        // the two minima start at the maximum
    dim3 b(128), g((m.w + 127) / 128, m.h);
    const int nb = (n + 255) / 256;
    ccl_local_kernel<<<dim3((m.w + CCL_TW - 1) / CCL_TW, (m.h + CCL_TH - 1) / CCL_TH), 256, 0, st>>>(m, labels, zgap);
    ccl_merge_kernel<<<g, b, 0, st>>>(m, labels, zgap);
    ccl_flatten_count_kernel<<<nb, 256, 0, st>>>(m, labels, cnt);
    ccl_max_kernel<<<nb, 256, 0, st>>>(cnt, n, best);
    ccl_ties_kernel<<<nb, 256, 0, st>>>(cnt, n, best);
    ccl_tiebreak_kernel<<<g, b, 0, st>>>(m, labels, cnt, best);
    ccl_extract_kernel<<<nb, 256, 0, st>>>(m, labels, best);
    unsigned long long cntmax = 0;
    cudaMemcpyAsync(&cntmax, best, 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    *n_left_host = cntmax;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ransac_find_plane (PovMesh.cpp:665-777): host draws the pixel triples with libc rand(); the device
// builds all hypotheses and scores every one against every point in a single pass over the mesh.
// ------------------------------------------------------------------------------------------------
__global__ void ransac_planes_kernel(MeshView m, const int* __restrict__ triples, int n, double* planes, int* ok)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int* t = triples + 6 * r;
    const size_t i1 = (size_t)t[1] * m.w + t[0], i2 = (size_t)t[3] * m.w + t[2], i3 = (size_t)t[5] * m.w + t[4];
    if (!m.valid[i1] || !m.valid[i2] || !m.valid[i3]) {      // (a plane no point is close to: the scorer never looks at ok[])
        ok[r] = 0;
        planes[4 * r] = 0; planes[4 * r + 1] = 0; planes[4 * r + 2] = 0; planes[4 * r + 3] = (double)INFINITY;
        return;
    }
    const double ax = m.X[i2] - m.X[i1], ay = m.Y[i2] - m.Y[i1], az = m.Z[i2] - m.Z[i1];
    const double bx = m.X[i3] - m.X[i1], by = m.Y[i3] - m.Y[i1], bz = m.Z[i3] - m.Z[i1];
    double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);   // cv::Vec / double multiplies by the reciprocal
    nx *= s; ny *= s; nz *= s;
    if (nz < 0) { nx *= -1.0; ny *= -1.0; nz *= -1.0; }
    planes[4 * r] = nx; planes[4 * r + 1] = ny; planes[4 * r + 2] = nz;
    planes[4 * r + 3] = -(nx * m.X[i1] + ny * m.Y[i1] + nz * m.Z[i1]);
    ok[r] = 1;
}
void launch_ransac_planes(const MeshView& m, const int* triples, int n, double* planes, int* ok, cudaStream_t st)
{
    ransac_planes_kernel<<<(n + 127) / 128, 128, 0, st>>>(m, triples, n, planes, ok);
}

static constexpr int RHMAX = 768;    // hypotheses per pass over the points: planes as double4 + float4 and counters in shared memory (40 KB)
static constexpr int RPTS = 8;       // points a thread keeps in registers (as floats) while it walks the planes
// Inlier counts of all hypotheses (PovMesh.cpp:735-752: |n.p + d| < thr over ALL slots, fp64, in the reference's operation
// order; the file is compiled without FMA contraction).  Scoring 400 planes against 5 M points in fp64 is bound by the FP64
// pipe (seven operations per point and plane: 1.5 ms), so the verdict is SCREENED in fp32 first: |dist32 - dist64| is bounded
// by eps = 2^-20 (|x| + |y| + |z| + max|d| + thr + 1) (four roundings and four conversions of at most 2^-24 relative each, on
// terms bounded by that sum since |n| = 1 -- a 3x margin), so dist32 < thr - eps is an inlier and dist32 >= thr + eps is
// not, exactly as in fp64; only the points inside the band (a few in 10^4) and NaNs take the fp64 expression, on values
// re-read from memory.  A thread holds RPTS points in registers and walks ALL planes, four per iteration (broadcast
// shared-memory loads); the counts of two planes share a register (16 bits each: a warp's sum is at most 256), so one
// REDUX.SUM folds the warp for two planes, and lane 0 issues the shared atomics.  The points are read once.
__device__ __forceinline__ unsigned ransac_exact4(const MeshView& m, const double* sp4, size_t base, int lane, size_t stride, size_t npts,
                                                  const float (&hi)[RPTS], double thr)
{
    // exact counts of the RPTS points of this thread for FOUR consecutive planes, one byte each
    unsigned out = 0;
    for (int q = 0; q < 4; ++q) {
        const double2 ab = *reinterpret_cast<const double2*>(sp4 + 4 * q), cd = *reinterpret_cast<const double2*>(sp4 + 4 * q + 2);
        unsigned c = 0;
#pragma unroll
        for (int k = 0; k < RPTS; ++k) {
            if (!(hi[k] > 0.0f)) continue;                // invalid slot
            const size_t j = base + lane + k * stride;
            c += (unsigned)(fabs(ab.x * m.X[j] + ab.y * m.Y[j] + cd.x * m.Z[j] + cd.y) < thr);
        }
        out |= c << (8 * q);
    }
    return out;
}
__global__ void __launch_bounds__(256) ransac_count_kernel(MeshView m, const double* __restrict__ planes, const int* __restrict__ ok,
                                                           int n, double thr, unsigned long long* counts)
{
    __shared__ __align__(16) double sp[RHMAX * 4];
    __shared__ __align__(16) float sf[RHMAX * 4];       // a, b, c, d in fp32
    __shared__ unsigned sc[RHMAX];
    __shared__ int s_dmax;                              // max |d| of the pass (float bits)
    const size_t npts = (size_t)m.w * m.h;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const float EPSK = 9.5367431640625e-7f;             // 2^-20
    const float thr32 = (float)thr;
    for (int h0 = 0; h0 < n; h0 += RHMAX) {
        const int nh = min(RHMAX, n - h0), nh4 = (nh + 3) & ~3;
        __syncthreads();
        if (threadIdx.x == 0) s_dmax = 0;
        __syncthreads();
        // (planes beyond nh in the last group of four: d = +inf, no point is ever close to them)
        for (int i = threadIdx.x; i < nh4 * 4; i += blockDim.x) {
            const double v = i < nh * 4 ? planes[4 * h0 + i] : ((i & 3) == 3 ? (double)INFINITY : 0.0);
            sp[i] = v; sf[i] = (float)v;
        }
        for (int i = threadIdx.x; i < nh4; i += blockDim.x) sc[i] = 0;
        for (int i = threadIdx.x; i < nh; i += blockDim.x) {
            const float ad = fabsf((float)planes[4 * (h0 + i) + 3]);
            if (ad == ad && ad < 3e38f) atomicMax(&s_dmax, __float_as_int(ad));      // (NaN / Inf planes go the fp64 way by themselves)
        }
        __syncthreads();
        const float eh = EPSK * (__int_as_float(s_dmax) + thr32 + 1.0f);
        // (the loop bounds are the same for all lanes of a warp: warp-wide reductions inside)
        for (size_t base = blockIdx.x * (size_t)blockDim.x + (threadIdx.x & ~31); base < npts; base += RPTS * stride) {
            float xf[RPTS], yf[RPTS], zf[RPTS], lo[RPTS], hi[RPTS];
            bool any = false;
#pragma unroll
            for (int k = 0; k < RPTS; ++k) {
                const size_t j = base + lane + k * stride;
                const bool v = j < npts && m.valid[j];
                xf[k] = v ? (float)m.X[j] : 0.f; yf[k] = v ? (float)m.Y[j] : 0.f; zf[k] = v ? (float)m.Z[j] : 0.f;
                const float eps = EPSK * (fabsf(xf[k]) + fabsf(yf[k]) + fabsf(zf[k])) + eh;
                // |dist32| < lo: inlier for certain; >= hi: certainly not; an invalid slot is neither, whatever the plane
                lo[k] = v ? thr32 - eps : -1.0f;
                hi[k] = v ? thr32 + eps : -1.0f;
                any |= v;
            }
            if (!__any_sync(0xffffffffu, any)) continue;
            for (int hh = 0; hh < nh4; hh += 4) {
                unsigned c01 = 0, c23 = 0, h01 = 0, h23 = 0;      // counts below lo / below hi, planes (hh, hh+1) and (hh+2, hh+3)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 pf = *reinterpret_cast<const float4*>(sf + 4 * (hh + q));
                    const unsigned one = (q & 1) ? 0x10000u : 1u;
                    unsigned cl = 0, ch = 0;
#pragma unroll
                    for (int k = 0; k < RPTS; ++k) {
                        const float d32 = fabsf(fmaf(pf.x, xf[k], fmaf(pf.y, yf[k], fmaf(pf.z, zf[k], pf.w))));
                        cl += (d32 < lo[k]) ? one : 0u;
                        ch += !(d32 >= hi[k]) ? one : 0u;         // (NaN counts here: it takes the fp64 way)
                    }
                    if (q < 2) { c01 += cl; h01 += ch; } else { c23 += cl; h23 += ch; }
                }
                if (c01 != h01 || c23 != h23) {          // a point inside the band (or a NaN): the reference's own expression decides
                    const unsigned e = ransac_exact4(m, sp + 4 * hh, base, lane, stride, npts, hi, thr);
                    c01 = (e & 0xFFu) | ((e & 0xFF00u) << 8);
                    c23 = ((e >> 16) & 0xFFu) | ((e >> 24) << 16);
                }
                c01 = __reduce_add_sync(0xffffffffu, c01);
                c23 = __reduce_add_sync(0xffffffffu, c23);
                if (lane == 0 && (c01 | c23)) {
                    if (c01 & 0xFFFFu) atomicAdd(&sc[hh], c01 & 0xFFFFu);
                    if (c01 >> 16) atomicAdd(&sc[hh + 1], c01 >> 16);
                    if (c23 & 0xFFFFu) atomicAdd(&sc[hh + 2], c23 & 0xFFFFu);
                    if (c23 >> 16) atomicAdd(&sc[hh + 3], c23 >> 16);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nh; i += blockDim.x)
            if (sc[i] && ok[h0 + i]) atomicAdd(&counts[h0 + i], (unsigned long long)sc[i]);
    }
}
void launch_ransac_count(const MeshView& m, const double* planes, const int* ok, int n, double thr, unsigned long long* counts, cudaStream_t st)
{
    cudaMemsetAsync(counts, 0, (size_t)n * 8, st);
    ransac_count_kernel<<<148 * 4, 256, 0, st>>>(m, planes, ok, n, thr, counts);
}

// crop_plane (PovMesh.cpp:780-815)
__global__ void crop_plane_kernel(MeshView m, double a, double b, double c, double d, double thr, unsigned long long* counter)
{
    const size_t n = (size_t)m.w * m.h;
    unsigned long long loc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (m.valid[i]) {
            const double dist = fabs(a * m.X[i] + b * m.Y[i] + c * m.Z[i] + d);
            if (dist < thr) ++loc; else m.valid[i] = 0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
    if ((threadIdx.x & 31) == 0 && loc) atomicAdd(counter, loc);
}
void launch_crop_plane(MeshView m, double a, double b, double c, double d, double thr, unsigned long long* counter, cudaStream_t st)
{
    cudaMemsetAsync(counter, 0, 8, st);
    crop_plane_kernel<<<296, 256, 0, st>>>(m, a, b, c, d, thr, counter);
}

// ------------------------------------------------------------------------------------------------
// refine_plane (PovMesh.cpp:581-660): weighted centroid, then weighted scatter matrix; per-block
// partial sums in a fixed order, summed on the host (deterministic).
// ------------------------------------------------------------------------------------------------
static constexpr int RBLK = 296;
int refine_blocks(const MeshView&) { return RBLK; }

__device__ __forceinline__ bool refine_inlier(const MeshView& m, const RefineArgs& a, size_t i, double& x, double& y, double& z, double& w)
{
    if (!m.valid[i]) return false;
    const int v = (int)(i / m.w), u = (int)(i % m.w);
    if (u < a.umin || u > a.umax || v < a.vmin || v > a.vmax) return false;
    x = m.X[i]; y = m.Y[i]; z = m.Z[i];
    const double dist = sqrt(x * x + y * y + z * z);
    if (!(x > a.xmin && x < a.xmax && y > a.ymin && y < a.ymax && dist < a.maxdist)) return false;
    w = a.weight_by_distance ? dist : 1.0;
    return true;
}
template <int NV>
__device__ __forceinline__ void block_sum_store(double (&acc)[NV], double* out)
{
    __shared__ double sh[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int wv = 0; wv < 8; ++wv) s += sh[threadIdx.x][wv];
        out[blockIdx.x * NV + threadIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) refine1_kernel(MeshView m, RefineArgs a, double* partial)
{
    double acc[5] = {0, 0, 0, 0, 0};   // wsum, sum w*x, w*y, w*z, count
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double x, y, z, w;
        if (refine_inlier(m, a, i, x, y, z, w)) { acc[0] += w; acc[1] += x * w; acc[2] += y * w; acc[3] += z * w; acc[4] += 1.0; }
    }
    block_sum_store<5>(acc, partial);
}
__global__ void __launch_bounds__(256) refine2_kernel(MeshView m, RefineArgs a, double cx, double cy, double cz, double* partial)
{
    double acc[6] = {0, 0, 0, 0, 0, 0};   // xx xy xz yy yz zz
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double x, y, z, w;
        if (refine_inlier(m, a, i, x, y, z, w)) {
            x -= cx; y -= cy; z -= cz;
            acc[0] += w * x * x; acc[1] += w * x * y; acc[2] += w * x * z; acc[3] += w * y * y; acc[4] += w * y * z; acc[5] += w * z * z;
        }
    }
    block_sum_store<6>(acc, partial);
}
void launch_refine_pass1(const MeshView& m, const RefineArgs& a, double* partial, cudaStream_t st)
{
    refine1_kernel<<<RBLK, 256, 0, st>>>(m, a, partial);
}
void launch_refine_pass2(const MeshView& m, const RefineArgs& a, double cx, double cy, double cz, double* partial, cudaStream_t st)
{
    refine2_kernel<<<RBLK, 256, 0, st>>>(m, a, cx, cy, cz, partial);
}

// ------------------------------------------------------------------------------------------------
// save_as_xyz_compressed (PovMesh.cpp:377-460): limits in the plane frame, then quantise + compact
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void to_plane(const double* R, const double* T, double x, double y, double z, double& a, double& b, double& c)
{
    a = R[0] * x + R[1] * y + R[2] * z + T[0];
    b = R[3] * x + R[4] * y + R[5] * z + T[1];
    c = R[6] * x + R[7] * y + R[8] * z + T[2];
}
__device__ __forceinline__ unsigned long long dkey(double v)
{   // order-preserving map double -> u64
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k)
{
    const unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
}
__global__ void plane_minmax_kernel(MeshView m, const double* __restrict__ RT, unsigned long long* mm)
{
    double R[9], T[3];
    for (int i = 0; i < 9; ++i) R[i] = RT[i];
    for (int i = 0; i < 3; ++i) T[i] = RT[9 + i];
    unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!m.valid[i]) continue;
        double p[3];
        to_plane(R, T, m.X[i], m.Y[i], m.Z[i], p[0], p[1], p[2]);
        for (int k = 0; k < 3; ++k) { const unsigned long long key = dkey(p[k]); lo[k] = min(lo[k], key); hi[k] = max(hi[k], key); }
    }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&mm[k], lo[k]); atomicMax(&mm[3 + k], hi[k]); }
    }
}
__global__ void minmax_decode_kernel(const unsigned long long* mm, double* out)
{
    if (threadIdx.x < 6) out[threadIdx.x] = dunkey(mm[threadIdx.x]);
}
void launch_plane_minmax(const MeshView& m, const double* RT12_dev_and_out, const double* /*unused*/, double* minmax6, cudaStream_t st)
{
    // minmax6 doubles as scratch: first 6 x u64 keys, decoded in place afterwards
    unsigned long long* mm = (unsigned long long*)minmax6;
    const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0, 0, 0};
    cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, st);
    plane_minmax_kernel<<<296, 256, 0, st>>>(m, RT12_dev_and_out, mm);
    minmax_decode_kernel<<<1, 32, 0, st>>>(mm, minmax6);
}

__global__ void flags_kernel(MeshView m, unsigned* flags)
{
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) flags[i] = m.valid[i] ? 1u : 0u;
}
__global__ void quantise_scatter_kernel(MeshView m, const double* __restrict__ RT, const double* __restrict__ mins, const double* __restrict__ scl,
                                        const unsigned* __restrict__ pos, uint16_t* __restrict__ out)
{
    double R[9], T[3];
    for (int i = 0; i < 9; ++i) R[i] = RT[i];
    for (int i = 0; i < 3; ++i) T[i] = RT[9 + i];
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!m.valid[i]) continue;
        double p[3];
        to_plane(R, T, m.X[i], m.Y[i], m.Z[i], p[0], p[1], p[2]);
        uint16_t* o = out + (size_t)pos[i] * 3;
        for (int k = 0; k < 3; ++k) o[k] = (uint16_t)(unsigned)__double2uint_rz((p[k] - mins[k]) * scl[k]);
    }
}
__global__ void xyz_scatter_kernel(MeshView m, const unsigned* __restrict__ pos, float* __restrict__ out)
{
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!m.valid[i]) continue;
        float* o = out + (size_t)pos[i] * 3;
        o[0] = (float)m.X[i]; o[1] = (float)m.Y[i]; o[2] = (float)m.Z[i];
    }
}
size_t compact_cub_bytes(int n)
{
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const unsigned*)nullptr, (unsigned*)nullptr, n);
    return tmp;
}
static int scan_positions(const MeshView& m, unsigned* scan_tmp, void* cub_tmp, size_t cub_bytes, unsigned long long* n_host, cudaStream_t st)
{
    const int n = m.w * m.h;
    unsigned* flags = scan_tmp;
    unsigned* pos = scan_tmp + n;
    flags_kernel<<<296, 256, 0, st>>>(m, flags);
    cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flags, pos, n, st);
    unsigned lastp = 0, lastf = 0;
    cudaMemcpyAsync(&lastp, pos + n - 1, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&lastf, flags + n - 1, 4, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    *n_host = (unsigned long long)lastp + lastf;
    return 0;
}
int mesh_compact_quantise(const MeshView& m, const double* RT, const double* /*T3*/, const double* min3, const double* scale3,
                          uint16_t* out, unsigned* scan_tmp, void* cub_tmp, size_t cub_bytes, unsigned long long* n_host, cudaStream_t st)
{
    if (scan_positions(m, scan_tmp, cub_tmp, cub_bytes, n_host, st)) return -1;
    quantise_scatter_kernel<<<296, 256, 0, st>>>(m, RT, min3, scale3, scan_tmp + (size_t)m.w * m.h, out);
    return 0;
}
// every `every`-th refinement inlier (refine_inlier above) in grid scan order, as doubles: what main() writes to
// plane_refinement_inliers.xyz (wass_stereo.cpp:2077-2085), selected on the device instead of downloading the mesh
__global__ void refine_flags_kernel(MeshView m, RefineArgs a, unsigned* flags)
{
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double x, y, z, w;
        flags[i] = refine_inlier(m, a, i, x, y, z, w) ? 1u : 0u;
    }
}
__global__ void refine_sample_scatter_kernel(MeshView m, const unsigned* __restrict__ flags, const unsigned* __restrict__ pos, unsigned every,
                                             double* __restrict__ out)
{
    const size_t n = (size_t)m.w * m.h;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!flags[i] || pos[i] % every) continue;
        double* o = out + (size_t)(pos[i] / every) * 3;
        o[0] = m.X[i]; o[1] = m.Y[i]; o[2] = m.Z[i];
    }
}
int mesh_refine_sample(const MeshView& m, const RefineArgs& a, unsigned every, double* out, unsigned* scan_tmp, void* cub_tmp, size_t cub_bytes,
                       unsigned long long* n_inliers, cudaStream_t st)
{
    const int n = m.w * m.h;
    unsigned* flags = scan_tmp;
    unsigned* pos = scan_tmp + n;
    refine_flags_kernel<<<296, 256, 0, st>>>(m, a, flags);
    cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flags, pos, n, st);
    unsigned lastp = 0, lastf = 0;
    cudaMemcpyAsync(&lastp, pos + n - 1, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&lastf, flags + n - 1, 4, cudaMemcpyDeviceToHost, st);
    refine_sample_scatter_kernel<<<296, 256, 0, st>>>(m, flags, pos, every, out);
    if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
    *n_inliers = (unsigned long long)lastp + lastf;
    return 0;
}
int mesh_compact_xyz(const MeshView& m, float* out, unsigned* scan_tmp, void* cub_tmp, size_t cub_bytes,
                     unsigned long long* n_host, cudaStream_t st)
{
    if (scan_positions(m, scan_tmp, cub_tmp, cub_bytes, n_host, st)) return -1;
    xyz_scatter_kernel<<<296, 256, 0, st>>>(m, scan_tmp + (size_t)m.w * m.h, out);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Optional refinement of the ROI disparity (wass_stereo.cpp:941-986), both off at the reference defaults:
//  * MEDIAN_FILTER_WSIZE >= 3: cv::medianBlur on float32 (3 or 5; replicate border)
//  * DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD > 0: zero where the squared Sobel gradient magnitude exceeds the
//    threshold, then keep only the biggest 8-connected component of the non-zero pixels
//    (cv::connectedComponentsWithStats; ties go to the smallest label, and cv2's labels are ordered by the first 2x2 block
//    of a component in block-raster order -- checked against cv2 in tests/test_oracle_pipeline.py)
// Float operation order of cv::Sobel (3x3, scale 1): gx = ((d[y-1] + d[y+1]) + 2 d[y]) with d = p[x+1] - p[x-1];
// gy = s[y+1] - s[y-1] with s = ((p[x-1] + p[x+1]) + 2 p[x]); border BORDER_REFLECT_101.  No FMA (file is -fmad=false).
// ------------------------------------------------------------------------------------------------
template <int KS>
__global__ void median_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    constexpr int N = KS * KS, R = KS / 2;
    float v[N];
#pragma unroll
    for (int j = 0; j < KS; ++j) {
        const int yy = min(max(y + j - R, 0), rows - 1);
#pragma unroll
        for (int i = 0; i < KS; ++i) v[j * KS + i] = src[(size_t)yy * cols + min(max(x + i - R, 0), cols - 1)];
    }
    // partial selection sort up to the median (N <= 25)
#pragma unroll
    for (int a = 0; a <= N / 2; ++a) {
#pragma unroll
        for (int b = a + 1; b < N; ++b) {
            const float lo = fminf(v[a], v[b]), hi = fmaxf(v[a], v[b]);
            v[a] = lo; v[b] = hi;
        }
    }
    dst[(size_t)y * cols + x] = v[N / 2];
}

void launch_median_f32(const float* src, float* dst, int rows, int cols, int ksize, cudaStream_t st)
{
    dim3 b(128), g((cols + 127) / 128, rows);
    if (ksize == 3) median_f32_kernel<3><<<g, b, 0, st>>>(src, dst, rows, cols);
    else            median_f32_kernel<5><<<g, b, 0, st>>>(src, dst, rows, cols);
}

__device__ __forceinline__ int refl101(int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); }

__global__ void gradient_mask_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, float thr)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int xm = refl101(x - 1, cols), xp = refl101(x + 1, cols);
    float d[3], sm[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float* r = src + (size_t)refl101(y + j - 1, rows) * cols;
        d[j] = r[xp] - r[xm];
        sm[j] = (r[xm] + r[xp]) + 2.0f * r[x];
    }
    const float gx = (d[0] + d[2]) + 2.0f * d[1];
    const float gy = sm[2] - sm[0];
    const float g2 = gx * gx + gy * gy;
    const float c = src[(size_t)y * cols + x];
    dst[(size_t)y * cols + x] = g2 > thr ? 0.0f : c;
}

void launch_gradient_mask(const float* src, float* dst, int rows, int cols, float thr, cudaStream_t st)
{
    dim3 b(128), g((cols + 127) / 128, rows);
    gradient_mask_kernel<<<g, b, 0, st>>>(src, dst, rows, cols, thr);
}

// 8-connected components of the non-zero pixels; L: rows*cols ints; cnt / key: rows*cols unsigned each
__global__ void cc8_init_kernel(const float* __restrict__ d, int* L, unsigned* cnt, unsigned* key, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    L[i] = d[i] != 0.0f ? i : -1;
    cnt[i] = 0;
    key[i] = 0xFFFFFFFFu;
}
__global__ void cc8_merge_kernel(const float* __restrict__ d, int* L, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int i = y * cols + x;
    if (d[i] == 0.0f) return;
    if (x + 1 < cols && d[i + 1] != 0.0f) uf_union(L, i, i + 1);
    if (y + 1 < rows) {
        if (d[i + cols] != 0.0f) uf_union(L, i, i + cols);
        if (x + 1 < cols && d[i + cols + 1] != 0.0f) uf_union(L, i, i + cols + 1);
        if (x > 0 && d[i + cols - 1] != 0.0f) uf_union(L, i, i + cols - 1);
    }
}
__global__ void cc8_count_kernel(int* L, unsigned* cnt, unsigned* key, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int i = y * cols + x;
    if (L[i] < 0) return;
    const int r = uf_find(L, i);
    L[i] = r;
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, r);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[r], (unsigned)__popc(peers));
    atomicMin(&key[r], (unsigned)((y >> 1) * ((cols + 1) >> 1) + (x >> 1)));   // order of cv2's labels
}
__global__ void cc8_best_kernel(const unsigned* cnt, const unsigned* key, int n, unsigned long long* best, int* best_root)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || cnt[i] == 0) return;
    atomicMax(best, ((unsigned long long)cnt[i] << 32) | (unsigned long long)(0xFFFFFFFFu - key[i]));
}
__global__ void cc8_apply_kernel(float* d, const int* __restrict__ L, const unsigned* __restrict__ cnt, const unsigned* __restrict__ key,
                                 int n, const unsigned long long* best)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || L[i] < 0) return;
    const int r = L[i];
    const unsigned long long mine = ((unsigned long long)cnt[r] << 32) | (unsigned long long)(0xFFFFFFFFu - key[r]);
    if (mine != *best) d[i] = 0.0f;
}

void launch_keep_biggest_cc8(float* d, int rows, int cols, int* labels, unsigned* cnt, unsigned* key, unsigned long long* best,
                             cudaStream_t st)
{
    const int n = rows * cols, nb = (n + 255) / 256;
    dim3 b(128), g((cols + 127) / 128, rows);
    cudaMemsetAsync(best, 0, 8, st);
    cc8_init_kernel<<<nb, 256, 0, st>>>(d, labels, cnt, key, n);
    cc8_merge_kernel<<<g, b, 0, st>>>(d, labels, rows, cols);
    cc8_count_kernel<<<g, b, 0, st>>>(labels, cnt, key, rows, cols);
    cc8_best_kernel<<<nb, 256, 0, st>>>(cnt, key, n, best, nullptr);
    cc8_apply_kernel<<<nb, 256, 0, st>>>(d, labels, cnt, key, n, best);
}

// ------------------------------------------------------------------------------------------------
// Consumer side of mesh_cam.xyzC: load_camera_mesh + align_on_sea_plane (gridding/wassgridsurface/wass_utils.py:22-35,
// 38-68), the step right after the hot path.  params (doubles): [0..2] scale, [3..5] min,
// [6..14] Rinv, [15..17] Tinv, [18..26] R of the plane to align on, [27..29] T, [30] baseline.
// Operation order of the reference: p = u16/scale + min;  c = Rinv@p + Tinv;  a = R@c + T;  a.z = -a.z;  a *= baseline.
// ------------------------------------------------------------------------------------------------
__global__ void xyzc_decode_align_kernel(const uint16_t* __restrict__ q, size_t n, const double* __restrict__ P,
                                         double* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3], c[3], a[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = (double)(float)q[i * 3 + k] / P[k] + P[3 + k];
#pragma unroll
    for (int r = 0; r < 3; ++r) c[r] = (P[6 + 3 * r] * p[0] + P[6 + 3 * r + 1] * p[1] + P[6 + 3 * r + 2] * p[2]) + P[15 + r];
#pragma unroll
    for (int r = 0; r < 3; ++r) a[r] = (P[18 + 3 * r] * c[0] + P[18 + 3 * r + 1] * c[1] + P[18 + 3 * r + 2] * c[2]) + P[27 + r];
    a[2] = -a[2];
    const double b = P[30];
    out[i] = a[0] * b;
    out[n + i] = a[1] * b;
    out[2 * n + i] = a[2] * b;
}

void launch_xyzc_decode_align(const uint16_t* q, size_t n, const double* d_params, double* out, cudaStream_t st)
{
    if (n == 0) return;
    xyzc_decode_align_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q, n, d_params, out);
}

}  // namespace wsg
