// cv::resize as sgbm_dense_stereo uses it when DENSE_SCALE != 1 (src/wass_stereo/wass_stereo.cpp:788-797 on the 8-bit crops,
// :903-904 on the float disparity).  The arithmetic is OpenCV's own resize code (resizeGeneric_ with the cubic H/V passes,
// resizeNN), i.e. what the reference's IPP-free conda-forge libopencv executes; oracle: oracle/pipeline.py resize_*,
// pinned against cv2 with IPP off in tests/test_resize.py.  Every float operation is written with explicit _rn
// intrinsics: the order and the absence of fused multiply-adds are part of the result.
#include "geom.cuh"

namespace wsg {

// One record per destination coordinate: 4 replicate-clamped source indices, 4 float32 coefficients and their 11-bit
// fixed-point versions (saturate_cast<short>(c * 2048)).
__global__ void cubic_taps_kernel(int dn, int sn, double scale, int4* __restrict__ idx, float4* __restrict__ coef, int4* __restrict__ icoef)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= dn) return;
    // fx = (float)((dx + 0.5) * scale_x - 0.5); sx = cvFloor(fx); fx -= sx
    const float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
    const float fl = floorf(f);
    const int s = (int)fl;
    const float x = __fsub_rn(f, fl);
    const float A = -0.75f;
    const float xp = __fadd_rn(x, 1.f);
    float c0 = __fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp), 5.f * A), xp), 8.f * A);
    c0 = __fsub_rn(__fmul_rn(c0, xp), 4.f * A);
    const float c1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, x), A + 3.f), x), x), 1.f);
    const float xm = __fsub_rn(1.f, x);
    const float c2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, xm), A + 3.f), xm), xm), 1.f);
    const float c3 = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c0), c1), c2);
    auto cl = [sn](int v) { return min(max(v, 0), sn - 1); };
    auto fix = [](float c) { return min(max(__float2int_rn(__fmul_rn(c, 2048.f)), -32768), 32767); };
    idx[d] = make_int4(cl(s - 1), cl(s), cl(s + 1), cl(s + 2));
    coef[d] = make_float4(c0, c1, c2, c3);
    icoef[d] = make_int4(fix(c0), fix(c1), fix(c2), fix(c3));
}

__global__ void resize_cubic_u8_kernel(const uint8_t* __restrict__ src, size_t sstride, int dw, int dh,
                                       const int4* __restrict__ xi, const int4* __restrict__ xa,
                                       const int4* __restrict__ yi, const int4* __restrict__ yb,
                                       uint8_t* __restrict__ dst, size_t dstride)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int4 ix = xi[x], a = xa[x], iy = yi[y], b = yb[y];
    auto hrow = [&](int r) {
        const uint8_t* p = src + (size_t)r * sstride;
        return (int)p[ix.x] * a.x + (int)p[ix.y] * a.y + (int)p[ix.z] * a.z + (int)p[ix.w] * a.w;
    };
    const int h0 = hrow(iy.x), h1 = hrow(iy.y), h2 = hrow(iy.z), h3 = hrow(iy.w);
    int v;
    if (x < (dw & ~7)) {
        // the 8-lane SIMD loop: float32 with beta * 2^-22, nested multiply-adds (not fused), round half to even
        const float sc = 1.f / (2048.f * 2048.f);
        float t = __fmul_rn((float)h3, __fmul_rn((float)b.w, sc));
        t = __fadd_rn(__fmul_rn((float)h2, __fmul_rn((float)b.z, sc)), t);
        t = __fadd_rn(__fmul_rn((float)h1, __fmul_rn((float)b.y, sc)), t);
        t = __fadd_rn(__fmul_rn((float)h0, __fmul_rn((float)b.x, sc)), t);
        v = __float2int_rn(t);
    } else {
        v = (h0 * b.x + h1 * b.y + h2 * b.z + h3 * b.w + (1 << 21)) >> 22;      // the scalar tail: FixedPtCast<int,uchar,22>
    }
    dst[(size_t)y * dstride + x] = (uint8_t)min(max(v, 0), 255);
}

__global__ void resize_cubic_f32_kernel(const float* __restrict__ src, int sw, int dw, int dh,
                                        const int4* __restrict__ xi, const float4* __restrict__ xa,
                                        const int4* __restrict__ yi, const float4* __restrict__ yb, float* __restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int4 ix = xi[x], iy = yi[y];
    const float4 a = xa[x], b = yb[y];
    auto hrow = [&](int r) {
        const float* p = src + (size_t)r * sw;
        float t = __fmul_rn(p[ix.x], a.x);
        t = __fadd_rn(t, __fmul_rn(p[ix.y], a.y));
        t = __fadd_rn(t, __fmul_rn(p[ix.z], a.z));
        return __fadd_rn(t, __fmul_rn(p[ix.w], a.w));
    };
    const float h0 = hrow(iy.x), h1 = hrow(iy.y), h2 = hrow(iy.z), h3 = hrow(iy.w);
    float t;
    if (x < (dw & ~3)) {         // 4-lane SIMD loop: nested
        t = __fmul_rn(h3, b.w);
        t = __fadd_rn(__fmul_rn(h2, b.z), t);
        t = __fadd_rn(__fmul_rn(h1, b.y), t);
        t = __fadd_rn(__fmul_rn(h0, b.x), t);
    } else {                     // scalar tail: left to right
        t = __fmul_rn(h0, b.x);
        t = __fadd_rn(t, __fmul_rn(h1, b.y));
        t = __fadd_rn(t, __fmul_rn(h2, b.z));
        t = __fadd_rn(t, __fmul_rn(h3, b.w));
    }
    dst[(size_t)y * dw + x] = t;
}

__global__ void resize_nn_f32_kernel(const float* __restrict__ src, int sw, int sh, int dw, int dh, double ifx, double ify, float* __restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int sx = min((int)floor(__dmul_rn((double)x, ifx)), sw - 1);
    const int sy = min((int)floor(__dmul_rn((double)y, ify)), sh - 1);
    dst[(size_t)y * dw + x] = src[(size_t)sy * sw + sx];
}

// out = cub where nn_eroded != 0 else 0 (wass_stereo.cpp:913-928)
__global__ void mask_where_zero_kernel(const float* __restrict__ cub, const float* __restrict__ nn_eroded, size_t n, float* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = nn_eroded[i] == 0.f ? 0.f : cub[i];
}

// tab: room for resize_tab_bytes(dw, dh) bytes
size_t resize_tab_bytes(int dw, int dh) { return ((size_t)dw + dh) * 48 + 64; }

static void make_taps(int dw, int sw, double inv_x, int dh, int sh, double inv_y, void* tab, int4*& xi, float4*& xc, int4*& xa,
                      int4*& yi, float4*& yc, int4*& yb, cudaStream_t st)
{
    char* p = (char*)tab;
    xi = (int4*)p; p += (size_t)dw * 16;
    xc = (float4*)p; p += (size_t)dw * 16;
    xa = (int4*)p; p += (size_t)dw * 16;
    yi = (int4*)p; p += (size_t)dh * 16;
    yc = (float4*)p; p += (size_t)dh * 16;
    yb = (int4*)p;
    cubic_taps_kernel<<<(dw + 255) / 256, 256, 0, st>>>(dw, sw, 1.0 / inv_x, xi, xc, xa);
    cubic_taps_kernel<<<(dh + 255) / 256, 256, 0, st>>>(dh, sh, 1.0 / inv_y, yi, yc, yb);
}

void launch_resize_cubic_u8(const uint8_t* src, size_t sstride, int sw, int sh, double fx, double fy, uint8_t* dst, size_t dstride,
                            int dw, int dh, void* tab, cudaStream_t st)
{
    int4 *xi, *xa, *yi, *yb; float4 *xc, *yc;
    make_taps(dw, sw, fx, dh, sh, fy, tab, xi, xc, xa, yi, yc, yb, st);
    dim3 b(128), g((dw + 127) / 128, dh);
    resize_cubic_u8_kernel<<<g, b, 0, st>>>(src, sstride, dw, dh, xi, xa, yi, yb, dst, dstride);
}

void launch_resize_cubic_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, void* tab, cudaStream_t st)
{
    int4 *xi, *xa, *yi, *yb; float4 *xc, *yc;
    make_taps(dw, sw, (double)dw / sw, dh, sh, (double)dh / sh, tab, xi, xc, xa, yi, yc, yb, st);
    dim3 b(128), g((dw + 127) / 128, dh);
    resize_cubic_f32_kernel<<<g, b, 0, st>>>(src, sw, dw, dh, xi, xc, yi, yc, dst);
}

void launch_resize_nn_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, cudaStream_t st)
{
    dim3 b(128), g((dw + 127) / 128, dh);
    resize_nn_f32_kernel<<<g, b, 0, st>>>(src, sw, sh, dw, dh, 1.0 / ((double)dw / sw), 1.0 / ((double)dh / sh), dst);
}

void launch_mask_where_zero(const float* cub, const float* nn_eroded, size_t n, float* out, cudaStream_t st)
{
    mask_where_zero_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cub, nn_eroded, n, out);
}

}  // namespace wsg
