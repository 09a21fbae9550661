// Rectification stage of wass_stereo (src/wass_stereo/wass_stereo.cpp:447-613, cv::stereoRectify branch):
//   wsg_stereo_rectify   host arithmetic of cv::stereoRectify(K0,0,K1,0,size,R,T,...,flags=0,alpha=1,size)
//   wsg_rectify_image    cv::initUndistortRectifyMap(K,0,Rrect,P,size,CV_32FC1) + cv::remap(INTER_CUBIC)
// The published OpenCV algorithm (Bouguet's method, fixed-point bicubic remap with 1/32-pixel
// phases) is restated here; tests/test_rectify.py pins both functions against cv2.
#include "handle.cuh"
#include <vector>

#include <cfloat>
#include <cmath>

namespace {

void mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
void mat3_t(const double* A, double* T) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * j + i]; }
void mat3_vec(const double* A, const double* v, double* o) { for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2]; }

void rodrigues_v2m(const double* om, double* R)
{
    const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    if (theta < DBL_EPSILON) { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0); return; }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
    const double x = om[0] * it, y = om[1] * it, z = om[2] * it;
    R[0] = c + c1 * x * x; R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y; R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}
void rodrigues_m2v(const double* R, double* om)
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    const double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { om[0] = om[1] = om[2] = 0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1. : 1.);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        const double k = theta / sqrt(rx * rx + ry * ry + rz * rz);
        om[0] = rx * k; om[1] = ry * k; om[2] = rz * k;
        return;
    }
    const double vth = 1 / (2 * s) * theta;
    om[0] = rx * vth; om[1] = ry * vth; om[2] = rz * vth;
}

struct RectF { float x, y, w, h; };

// icvGetRectangles with zero distortion: 9x9 grid -> normalise with K, rotate with R, project with P
void get_rectangles(const double* K, const double* R, const double* P /*3x4*/, int W, int H, RectF& inner, RectF& outer)
{
    const int N = 9;
    float iX0 = -FLT_MAX, iX1 = FLT_MAX, iY0 = -FLT_MAX, iY1 = FLT_MAX, oX0 = FLT_MAX, oX1 = -FLT_MAX, oY0 = FLT_MAX, oY1 = -FLT_MAX;
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const float px = (float)x * (W - 1) / (N - 1), py = (float)y * (H - 1) / (N - 1);
            const double xn = ((double)px - K[2]) / K[0], yn = ((double)py - K[5]) / K[4];
            const double X = R[0] * xn + R[1] * yn + R[2], Y = R[3] * xn + R[4] * yn + R[5], Wz = R[6] * xn + R[7] * yn + R[8];
            const double xr = X / Wz, yr = Y / Wz;
            const float qx = (float)(xr * P[0] + P[2]), qy = (float)(yr * P[5] + P[6]);
            oX0 = fminf(oX0, qx); oX1 = fmaxf(oX1, qx); oY0 = fminf(oY0, qy); oY1 = fmaxf(oY1, qy);
            if (x == 0) iX0 = fmaxf(iX0, qx);
            if (x == N - 1) iX1 = fminf(iX1, qx);
            if (y == 0) iY0 = fmaxf(iY0, qy);
            if (y == N - 1) iY1 = fminf(iY1, qy);
        }
    inner = {iX0, iY0, iX1 - iX0, iY1 - iY0};
    outer = {oX0, oY0, oX1 - oX0, oY1 - oY0};
}

// ---- bicubic fixed-point tables of cv::remap (INTER_BITS = 5, coefficient scale 2^15) -----------------
const int TABSZ = 32, COEF_BITS = 15, COEF_SCALE = 1 << COEF_BITS;

inline short sat_short_round(float v)
{
    const long r = lrintf(v);     // round half to even, like cvRound
    return (short)(r > 32767 ? 32767 : r < -32768 ? -32768 : r);
}

void build_cubic_table(short* itab /*[32*32][16]*/)
{
    float tab1[TABSZ * 4];
    const float A = -0.75f;
    for (int i = 0; i < TABSZ; ++i) {
        const float x = i * (1.f / TABSZ);
        float* c = tab1 + 4 * i;
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
    }
    for (int i = 0; i < TABSZ; ++i)
        for (int j = 0; j < TABSZ; ++j) {
            short* t = itab + (i * TABSZ + j) * 16;
            int isum = 0;
            for (int k1 = 0; k1 < 4; ++k1) {
                const float vy = tab1[i * 4 + k1];
                for (int k2 = 0; k2 < 4; ++k2) {
                    const float v = vy * tab1[j * 4 + k2];
                    isum += t[k1 * 4 + k2] = sat_short_round(v * COEF_SCALE);
                }
            }
            if (isum != COEF_SCALE) {
                const int diff = isum - COEF_SCALE;
                int Mk1 = 2, Mk2 = 2, mk1 = 2, mk2 = 2;
                for (int k1 = 2; k1 < 4; ++k1)
                    for (int k2 = 2; k2 < 4; ++k2) {
                        if (t[k1 * 4 + k2] < t[mk1 * 4 + mk2]) { mk1 = k1; mk2 = k2; }
                        else if (t[k1 * 4 + k2] > t[Mk1 * 4 + Mk2]) { Mk1 = k1; Mk2 = k2; }
                    }
                if (diff < 0) t[Mk1 * 4 + Mk2] = (short)(t[Mk1 * 4 + Mk2] - diff);
                else t[mk1 * 4 + mk2] = (short)(t[mk1 * 4 + mk2] - diff);
            }
        }
}

__global__ void rectify_remap_kernel(const uint8_t* __restrict__ src, int rows, int cols, size_t stride, const double* __restrict__ ir /*9*/,
                                     double fx, double fy, double cx, double cy, const short* __restrict__ wtab, uint8_t* __restrict__ dst)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= cols) return;
    // initUndistortRectifyMap, zero distortion: [x y w] = (P_3x3 * R)^-1 [u v 1]
    const double X = ir[0] * u + ir[1] * v + ir[2], Y = ir[3] * u + ir[4] * v + ir[5], Wz = ir[6] * u + ir[7] * v + ir[8];
    const double w = 1. / Wz, x = X * w, y = Y * w;
    const float mx = (float)(x * fx + cx), my = (float)(y * fy + cy);
    // cv::remap: maps -> 1/32-pixel fixed point
    const int sxf = __float2int_rn(mx * TABSZ), syf = __float2int_rn(my * TABSZ);
    int sx = sxf >> 5, sy = syf >> 5;
    sx = min(max(sx, -32768), 32767); sy = min(max(sy, -32768), 32767);
    const short* wt = wtab + ((syf & 31) * TABSZ + (sxf & 31)) * 16;
    sx -= 1; sy -= 1;
    int out = 0;
    if (!(sx >= cols || sx + 4 <= 0 || sy >= rows || sy + 4 <= 0)) {
        int sum = 0;
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            const int yy = sy + k1;
            if (yy < 0 || yy >= rows) continue;
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
                const int xx = sx + k2;
                if (xx < 0 || xx >= cols) continue;
                sum += (int)src[(size_t)yy * stride + xx] * wt[k1 * 4 + k2];
            }
        }
        out = (sum + (1 << (COEF_BITS - 1))) >> COEF_BITS;
        out = min(max(out, 0), 255);
    }
    dst[(size_t)v * cols + u] = (uint8_t)out;
}

bool inv3(const double* m, double* o)
{
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0) return false;
    const double d = 1. / det;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * d; o[1] = (m[2] * m[7] - m[1] * m[8]) * d; o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * d; o[4] = (m[0] * m[8] - m[2] * m[6]) * d; o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * d; o[7] = (m[1] * m[6] - m[0] * m[7]) * d; o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    return true;
}

}  // namespace

extern "C" {

int wsg_stereo_rectify(const double K0[9], const double K1[9], const double R[9], const double T[3], int width, int height,
                       double R1[9], double R2[9], double P1[12], double P2[12], int roi1[4], int roi2[4])
{
    if (!K0 || !K1 || !R || !T || !R1 || !R2 || !P1 || !P2 || width <= 0 || height <= 0) return WSG_ERR_INVALID_ARG;
    double om[3], r_r[9], t[3], uu[3] = {0, 0, 0}, ww[3], wR[9], tmp[9];
    rodrigues_m2v(R, om);
    for (int i = 0; i < 3; ++i) om[i] *= -0.5;          // each camera rotates half way
    rodrigues_v2m(om, r_r);
    mat3_vec(r_r, T, t);
    const int idx = fabs(t[0]) > fabs(t[1]) ? 0 : 1;
    const double c = t[idx], nt = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    uu[idx] = c > 0 ? 1 : -1;
    ww[0] = t[1] * uu[2] - t[2] * uu[1]; ww[1] = t[2] * uu[0] - t[0] * uu[2]; ww[2] = t[0] * uu[1] - t[1] * uu[0];
    const double nw = sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
    if (nw > 0.0) { const double k = acos(fabs(c) / nt) / nw; for (int i = 0; i < 3; ++i) ww[i] *= k; }
    rodrigues_v2m(ww, wR);
    mat3_t(r_r, tmp);
    mat3_mul(wR, tmp, R1);          // R1 = wR * r_r^T
    mat3_mul(wR, r_r, R2);          // R2 = wR * r_r
    mat3_vec(R2, T, t);
    // new focal length: mean of the two focal lengths on the axis perpendicular to the baseline
    // (what cv2 4.13 does with zero distortion; pinned in tests/test_rectify.py)
    const double fc_new0 = (K0[idx == 0 ? 4 : 0] + K1[idx == 0 ? 4 : 0]) * 0.5;
    double fc_new = fc_new0;
    double ccn[2][2];
    for (int k = 0; k < 2; ++k) {
        const double* A = k == 0 ? K0 : K1;
        const double* Rk = k == 0 ? R1 : R2;
        double ax = 0, ay = 0;
        for (int i = 0; i < 4; ++i) {
            const float px = (float)((i % 2) * (width - 1)), py = (float)((i < 2 ? 0 : 1) * (height - 1));
            const float xn = (float)(((double)px - A[2]) / A[0]), yn = (float)(((double)py - A[5]) / A[4]);
            const double X = Rk[0] * xn + Rk[1] * yn + Rk[2], Y = Rk[3] * xn + Rk[4] * yn + Rk[5], Z = Rk[6] * xn + Rk[7] * yn + Rk[8];
            const double z = Z ? 1. / Z : 1.;
            ax += (float)(X * z * fc_new); ay += (float)(Y * z * fc_new);
        }
        ccn[k][0] = (width - 1) / 2. - ax / 4;
        ccn[k][1] = (height - 1) / 2. - ay / 4;
    }
    // flags = 0 (no CALIB_ZERO_DISPARITY): only the coordinate perpendicular to the baseline is shared
    if (idx == 0) ccn[0][1] = ccn[1][1] = (ccn[0][1] + ccn[1][1]) * 0.5;
    else ccn[0][0] = ccn[1][0] = (ccn[0][0] + ccn[1][0]) * 0.5;
    auto fillP = [&](double* P, int k, double f) {
        for (int i = 0; i < 12; ++i) P[i] = 0;
        P[0] = f; P[5] = f; P[2] = ccn[k][0]; P[6] = ccn[k][1]; P[10] = 1;
    };
    fillP(P1, 0, fc_new); fillP(P2, 1, fc_new);
    P2[idx * 4 + 3] = t[idx] * fc_new;
    // alpha = 1 : scale so that all source pixels are retained
    RectF in1, out1, in2, out2;
    get_rectangles(K0, R1, P1, width, height, in1, out1);
    get_rectangles(K1, R2, P2, width, height, in2, out2);
    const double cx1_0 = ccn[0][0], cy1_0 = ccn[0][1], cx2_0 = ccn[1][0], cy2_0 = ccn[1][1];
    const double cx1 = cx1_0, cy1 = cy1_0, cx2 = cx2_0, cy2 = cy2_0;     // newImgSize == imageSize
    const double alpha = 1.0;
    double s0 = fmax(fmax(fmax(cx1 / (cx1_0 - in1.x), cy1 / (cy1_0 - in1.y)), (width - 1 - cx1) / (in1.x + in1.w - cx1_0)),
                     (height - 1 - cy1) / (in1.y + in1.h - cy1_0));
    s0 = fmax(fmax(fmax(fmax(cx2 / (cx2_0 - in2.x), cy2 / (cy2_0 - in2.y)), (width - 1 - cx2) / (in2.x + in2.w - cx2_0)),
                   (height - 1 - cy2) / (in2.y + in2.h - cy2_0)), s0);
    double s1 = fmin(fmin(fmin(cx1 / (cx1_0 - out1.x), cy1 / (cy1_0 - out1.y)), (width - 1 - cx1) / (out1.x + out1.w - cx1_0)),
                     (height - 1 - cy1) / (out1.y + out1.h - cy1_0));
    s1 = fmin(fmin(fmin(fmin(cx2 / (cx2_0 - out2.x), cy2 / (cy2_0 - out2.y)), (width - 1 - cx2) / (out2.x + out2.w - cx2_0)),
                   (height - 1 - cy2) / (out2.y + out2.h - cy2_0)), s1);
    const double s = s0 * (1 - alpha) + s1 * alpha;
    fc_new *= s;
    fillP(P1, 0, fc_new); fillP(P2, 1, fc_new);
    P2[idx * 4 + 3] = t[idx] * fc_new0 * s;
    auto roi = [&](const RectF& in, double cx0, double cy0, double cx, double cy, int* r) {
        int x = (int)ceil((in.x - cx0) * s + cx), y = (int)ceil((in.y - cy0) * s + cy);
        int w = (int)floor(in.w * s), h = (int)floor(in.h * s);
        // intersect with the image rectangle
        const int x0 = x > 0 ? x : 0, y0 = y > 0 ? y : 0;
        const int x1 = x + w < width ? x + w : width, y1 = y + h < height ? y + h : height;
        if (x1 <= x0 || y1 <= y0) { r[0] = r[1] = r[2] = r[3] = 0; return; }
        r[0] = x0; r[1] = y0; r[2] = x1 - x0; r[3] = y1 - y0;
    };
    if (roi1) roi(in1, cx1_0, cy1_0, cx1, cy1, roi1);
    if (roi2) roi(in2, cx2_0, cy2_0, cx2, cy2, roi2);
    return WSG_OK;
}

// ------------------------------------------------------------------------------------------------
// cv::undistort as wass_prepare calls it (src/wass_prepare/wass_prepare.cpp:268, 474-497):
// stripes of max(1, 4096/cols) rows, per stripe initUndistortRectifyMap(K, dist, I, K') with K'(1,2) = cy - y0 into 1/32-pixel
// fixed-point maps (CV_16SC2), then remap(INTER_LINEAR, BORDER_CONSTANT 0).  One thread per image row walks its columns
// with the same running sums as the reference (_x += ir[0] ...), so the quantised map matches bit for bit; the bilinear
// weights (32-fx)(32-fy)*32 ... are exact integers (only 32768 saturates to 32767, which cannot change a u8 result).
// ------------------------------------------------------------------------------------------------
struct UndistortArgs {
    double k[8];        // k1 k2 p1 p2 k3 k4 k5 k6
    double fx, fy, u0, v0;
    int rows, cols, stripe;
};

__global__ void undistort_map_kernel(UndistortArgs a, const double* __restrict__ ir_per_stripe, int2* __restrict__ map)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= a.rows) return;
    const double* ir = ir_per_stripe + (size_t)(y / a.stripe) * 9;
    const int i = y % a.stripe;
    double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
    const double k1 = a.k[0], k2 = a.k[1], p1 = a.k[2], p2 = a.k[3], k3 = a.k[4], k4 = a.k[5], k5 = a.k[6], k6 = a.k[7];
    for (int j = 0; j < a.cols; ++j, _x += ir[0], _y += ir[3], _w += ir[6]) {
        const double w = 1. / _w, x = _x * w, yy = _y * w;
        const double x2 = x * x, y2 = yy * yy;
        const double r2 = x2 + y2, _2xy = 2 * x * yy;
        const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
        const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2));
        const double yd = (yy * kr + p1 * (r2 + 2 * y2) + p2 * _2xy);
        const double u = a.fx * xd + a.u0, v = a.fy * yd + a.v0;
        // saturate_cast<int>(u*INTER_TAB_SIZE) == cvRound (round half to even)
        const double us = u * 32, vs = v * 32;
        const int iu = us >= 2147483647. ? 2147483647 : (us <= -2147483648. ? (int)-2147483648LL : __double2int_rn(us));
        const int iv = vs >= 2147483647. ? 2147483647 : (vs <= -2147483648. ? (int)-2147483648LL : __double2int_rn(vs));
        map[(size_t)y * a.cols + j] = make_int2(iu, iv);
    }
}

__global__ void undistort_remap_kernel(const uint8_t* __restrict__ src, int rows, int cols, size_t stride, const int2* __restrict__ map,
                                       uint8_t* __restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int2 m = map[(size_t)y * cols + x];
    // map1 = (short)(iu >> 5), (short)(iv >> 5): the narrowing cast wraps like the reference's
    const int sx = (short)(m.x >> 5), sy = (short)(m.y >> 5);
    const int fx = m.x & 31, fy = m.y & 31;
    int w[4] = {(32 - fx) * (32 - fy) * 32, fx * (32 - fy) * 32, (32 - fx) * fy * 32, fx * fy * 32};
    if (w[0] > 32767) w[0] = 32767;
    auto px = [&](int yy, int xx) -> int { return (xx >= 0 && xx < cols && yy >= 0 && yy < rows) ? (int)src[(size_t)yy * stride + xx] : 0; };
    int out = 0;
    if (!(sx >= cols || sx + 1 < 0 || sy >= rows || sy + 1 < 0)) {
        const int sum = px(sy, sx) * w[0] + px(sy, sx + 1) * w[1] + px(sy + 1, sx) * w[2] + px(sy + 1, sx + 1) * w[3];
        out = (sum + (1 << 14)) >> 15;
        out = min(max(out, 0), 255);
    }
    dst[(size_t)y * cols + x] = (uint8_t)out;
}

// cv::invert of a 3x3 double matrix (the closed form OpenCV uses for n <= 3 with DECOMP_LU)
static bool inv3_cv(const double* S, double* D)
{
    const double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
    if (d == 0.) return false;
    const double id = 1. / d;
    double t[9];
    t[0] = (S[4] * S[8] - S[5] * S[7]) * id; t[1] = (S[2] * S[7] - S[1] * S[8]) * id; t[2] = (S[1] * S[5] - S[2] * S[4]) * id;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * id; t[4] = (S[0] * S[8] - S[2] * S[6]) * id; t[5] = (S[2] * S[3] - S[0] * S[5]) * id;
    t[6] = (S[3] * S[7] - S[4] * S[6]) * id; t[7] = (S[1] * S[6] - S[0] * S[7]) * id; t[8] = (S[0] * S[4] - S[1] * S[3]) * id;
    for (int i = 0; i < 9; ++i) D[i] = t[i];
    return true;
}

// ------------------------------------------------------------------------------------------------
// stereoRectifyUndistorted (src/wass_stereo/stereorectify.cpp:57-244): host arithmetic in cv::Matx order
// ------------------------------------------------------------------------------------------------
static void mm3(const double* A, const double* B, double* C)      // cv::Matx product: s = 0; s += a(i,k) * b(k,j)
{
    double t[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s_ = 0;
            for (int k = 0; k < 3; ++k) s_ += A[i * 3 + k] * B[k * 3 + j];
            t[i * 3 + j] = s_;
        }
    memcpy(C, t, sizeof t);
}
static double det3(const double* a)                                // cv::determinant(Matx33d)
{
    return a[0] * (a[4] * a[8] - a[7] * a[5]) - a[1] * (a[3] * a[8] - a[6] * a[5]) + a[2] * (a[3] * a[7] - a[6] * a[4]);
}
static void scale3(double* a, double f) { for (int i = 0; i < 9; ++i) a[i] *= f; }

struct HFunctional {           // stereorectify.cpp:70-137
    double K0i[9], K1i[9], Ri[9], Rplane[9], H0[9], H1[9];
    bool init(const double* K0, const double* K1, const double* R, const double* ep1)
    {
        if (!inv3_cv(K0, K0i) || !inv3_cv(K1, K1i)) return false;
        memcpy(Ri, R, 72);
        const double n = sqrt(ep1[0] * ep1[0] + ep1[1] * ep1[1] + ep1[2] * ep1[2]);
        if (!(n > 0)) return false;
        const double Rv[3] = {ep1[0] / n, ep1[1] / n, ep1[2] / n};
        double N[3] = {Rv[1] * 0 - Rv[2] * 1, Rv[2] * 0 - Rv[0] * 0, Rv[0] * 1 - Rv[1] * 0};      // Rv x (0,1,0)
        const double nn = sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
        if (!(nn > 0)) return false;
        for (double& v : N) v /= nn;
        const double Rk[3] = {Rv[1] * N[2] - Rv[2] * N[1], Rv[2] * N[0] - Rv[0] * N[2], Rv[0] * N[1] - Rv[1] * N[0]};
        const double rp[9] = {Rv[0], Rv[1], Rv[2], Rk[0], Rk[1], Rk[2], N[0], N[1], N[2]};
        memcpy(Rplane, rp, 72);
        return true;
    }
    double calc(double x)
    {
        // cv::Rodrigues of (x/180*3.14, 0, 0) -- 3.14, not pi, as in the reference
        const double a = x / 180 * 3.14, th = fabs(a);
        double Radd[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (th >= DBL_EPSILON) {
            const double c = cos(th), s_ = sin(th), c1 = 1. - c, rx = a * (1. / th);
            Radd[0] = c + c1 * rx * rx; Radd[4] = c; Radd[5] = -s_ * rx; Radd[7] = s_ * rx; Radd[8] = c;
        }
        double RP[9];
        mm3(Radd, Rplane, RP);
        mm3(RP, K0i, H0);
        double RPR[9];
        mm3(RP, Ri, RPR);
        mm3(RPR, K1i, H1);
        scale3(H0, 1.0 / H0[8]);
        scale3(H1, 1.0 / H1[8]);
        const double v1 = H0[6] * H0[6] + H0[7] * H0[7], v2 = H1[6] * H1[6] + H1[7] * H1[7];
        scale3(H0, 1.0 / cbrt(det3(H0)));       // det(H) = 1 for numerical stability
        scale3(H1, 1.0 / cbrt(det3(H1)));
        return std::max(v1, v2);
    }
};

// Nelder-Mead over (x0, x1) exactly as the reference sets the solver up: start (0,0), initial step (-0.5,-0.5), stop when
// the simplex has collapsed (function range below 1e-40) or after 500000 evaluations; only x0 enters the functional.
static double minimise_angle(HFunctional& f)
{
    double p[3][2] = {{0, 0}, {-0.5, 0}, {0, -0.5}}, y[3];
    for (int i = 0; i < 3; ++i) y[i] = f.calc(p[i][0]);
    int nfunk = 3;
    auto try_move = [&](int ihi, double fac, double& ynew, double pnew[2]) {
        const double fac1 = (1.0 - fac) / 2, fac2 = fac1 - fac;
        for (int j = 0; j < 2; ++j) pnew[j] = (p[0][j] + p[1][j] + p[2][j]) * fac1 - p[ihi][j] * fac2;
        ynew = f.calc(pnew[0]);
        ++nfunk;
    };
    while (nfunk < 500000) {
        int ilo = 0, ihi = 0, inhi = 0;
        for (int i = 1; i < 3; ++i) { if (y[i] < y[ilo]) ilo = i; if (y[i] > y[ihi]) ihi = i; }
        inhi = ilo;
        for (int i = 0; i < 3; ++i) if (i != ihi && y[i] >= y[inhi]) inhi = i;
        double ext = 0;
        for (int i = 0; i < 3; ++i) ext = std::max(ext, fabs(p[i][0] - p[ilo][0]));
        if (fabs(y[ihi] - y[ilo]) < 1e-40 && ext < 1e-13) break;
        double yt, pt[2];
        try_move(ihi, -1.0, yt, pt);                                   // reflection
        if (yt < y[ihi]) { y[ihi] = yt; p[ihi][0] = pt[0]; p[ihi][1] = pt[1]; }
        if (yt <= y[ilo]) {
            double y2, p2[2];
            try_move(ihi, 2.0, y2, p2);                                // expansion
            if (y2 < y[ihi]) { y[ihi] = y2; p[ihi][0] = p2[0]; p[ihi][1] = p2[1]; }
        } else if (yt >= y[inhi]) {
            const double ysave = y[ihi];
            double y2, p2[2];
            try_move(ihi, 0.5, y2, p2);                                // contraction
            if (y2 < y[ihi]) { y[ihi] = y2; p[ihi][0] = p2[0]; p[ihi][1] = p2[1]; }
            if (y2 >= ysave) {                                         // shrink towards the best vertex
                for (int i = 0; i < 3; ++i)
                    if (i != ilo) {
                        for (int j = 0; j < 2; ++j) p[i][j] = 0.5 * (p[i][j] + p[ilo][j]);
                        y[i] = f.calc(p[i][0]);
                        ++nfunk;
                    }
            }
        }
    }
    int ilo = 0;
    for (int i = 1; i < 3; ++i) if (y[i] < y[ilo]) ilo = i;
    return p[ilo][0];
}

// cv::warpPerspective's coordinate map (INTER_LINEAR): M = H^-1, 32-row x 128-column blocks (X0 from the block origin,
// then + M0 * x1), 32/W scaling, clamp to the int range, round half to even, saturate_cast<short> of the integer part.
__global__ void warp_map_kernel(int rows, int cols, int bw, const double* __restrict__ Mg, int2* __restrict__ map)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = Mg[i];
    const int xb = x / bw * bw, x1 = x - xb;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], (double)xb), __dmul_rn(M[1], (double)y)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], (double)xb), __dmul_rn(M[4], (double)y)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], (double)xb), __dmul_rn(M[7], (double)y)), M[8]);
    double W = __dadd_rn(W0, __dmul_rn(M[6], (double)x1));
    W = W != 0. ? __ddiv_rn(32.0, W) : 0.;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], (double)x1)), W)));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], (double)x1)), W)));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    auto sat = [](int v) { const int i = min(max(v >> 5, -32768), 32767); return (i << 5) | (v & 31); };
    map[(size_t)y * cols + x] = make_int2(sat(X), sat(Y));
}

// cv::undistort on device buffers: d_src -> d_dst (both rows x cols, tight)
static int undistort_device(wsg_handle* h, const uint8_t* d_src, uint8_t* d_dst, int rows, int cols, const double K[9],
                            const double* dist, int ndist)
{
    if (ndist != 0 && ndist != 4 && ndist != 5 && ndist != 8) { h->err = "distortion vector must have 0, 4, 5 or 8 coefficients"; return WSG_ERR_INVALID_ARG; }
    UndistortArgs a;
    for (int i = 0; i < 8; ++i) a.k[i] = i < ndist ? dist[i] : 0.;
    a.fx = K[0]; a.fy = K[4]; a.u0 = K[2]; a.v0 = K[5];
    a.rows = rows; a.cols = cols;
    a.stripe = std::min(std::max(1, (1 << 12) / std::max(cols, 1)), rows);
    const int nstripes = (rows + a.stripe - 1) / a.stripe;
    std::vector<double> ir((size_t)nstripes * 9);
    for (int sidx = 0; sidx < nstripes; ++sidx) {
        double Ar[9] = {K[0], K[1], K[2], K[3], K[4], K[5] - (double)(sidx * a.stripe), K[6], K[7], K[8]};
        if (!inv3_cv(Ar, &ir[(size_t)sidx * 9])) { h->err = "singular camera matrix"; return WSG_ERR_INVALID_ARG; }
    }
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->m_scratch, n * sizeof(int2) + ir.size() * 8 + 64))) return rc;
    int2* d_map = (int2*)h->m_scratch.p;
    double* d_ir = (double*)((char*)h->m_scratch.p + n * sizeof(int2));
    CK(h, cudaMemcpyAsync(d_ir, ir.data(), ir.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));        // `ir` is pageable host memory leaving scope
    undistort_map_kernel<<<(rows + 63) / 64, 64, 0, h->stream>>>(a, d_ir, d_map);
    dim3 b(128), g((cols + 127) / 128, rows);
    undistort_remap_kernel<<<g, b, 0, h->stream>>>(d_src, rows, cols, cols, d_map, d_dst);
    CK(h, cudaGetLastError());
    return WSG_OK;
}

// ------------------------------------------------------------------------------------------------
// cv::CLAHE::apply on CV_8UC1 (src/wass_prepare/wass_prepare.cpp:257-262, 458-462, 479-483).  OpenCV's algorithm: the image is
// padded (BORDER_REFLECT_101, right and bottom) to a multiple of the tile grid -- by a whole extra `tiles` when only one of
// the two sizes divides --, one 256-bin histogram per tile is clipped at clipLimit*area/256 with the excess redistributed
// (equal batch + one count on every `256/residual`-th bin), its running sum times 255/area rounded to the tile's LUT, and
// every pixel is the bilinear blend (float32, products then sums, round half to even) of the four nearest tiles' LUTs.
// Oracle: oracle/pipeline.py clahe_u8, bit-exact vs cv2.createCLAHE in tests/test_prepare.py.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ src, int rows, int cols, int tw, int th,
                                                        int clip, float lut_scale, uint8_t* __restrict__ luts)
{
    __shared__ int hist[256];
    __shared__ int scan[256];
    __shared__ int red[8];
    const int t = threadIdx.x;
    hist[t] = 0;
    __syncthreads();
    const int x0 = blockIdx.x * tw, y0 = blockIdx.y * th;
    for (int i = t; i < tw * th; i += 256) {
        const int y = reflect101(y0 + i / tw, rows), x = reflect101(x0 + i % tw, cols);
        atomicAdd(&hist[src[(size_t)y * cols + x]], 1);
    }
    __syncthreads();
    int hv = hist[t];
    if (clip > 0) {
        int over = max(hv - clip, 0);
        hv = min(hv, clip);
        for (int o = 16; o; o >>= 1) over += __shfl_xor_sync(0xffffffffu, over, o);
        if ((t & 31) == 0) red[t >> 5] = over;
        __syncthreads();
        int clipped = 0;
        for (int k = 0; k < 8; ++k) clipped += red[k];
        const int batch = clipped / 256, residual = clipped - batch * 256;
        hv += batch;
        if (residual) {
            const int step = max(256 / residual, 1);
            if (t % step == 0 && t / step < residual) ++hv;
        }
    }
    scan[t] = hv;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {          // inclusive running sum
        const int v = t >= o ? scan[t - o] : 0;
        __syncthreads();
        scan[t] += v;
        __syncthreads();
    }
    const int v = __float2int_rn(__fmul_rn((float)scan[t], lut_scale));
    luts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 256 + t] = (uint8_t)min(max(v, 0), 255);
}

__global__ void clahe_apply_kernel(const uint8_t* __restrict__ src, int rows, int cols, int tiles_x, int tiles_y, float inv_tw,
                                   float inv_th, const uint8_t* __restrict__ luts, uint8_t* __restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f), tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    const float fx = floorf(txf), fy = floorf(tyf);
    const float xa = __fsub_rn(txf, fx), ya = __fsub_rn(tyf, fy);
    const float xa1 = __fsub_rn(1.f, xa), ya1 = __fsub_rn(1.f, ya);
    const int tx1 = max((int)fx, 0), tx2 = min((int)fx + 1, tiles_x - 1);
    const int ty1 = max((int)fy, 0), ty2 = min((int)fy + 1, tiles_y - 1);
    const int v = src[(size_t)y * cols + x];
    const uint8_t* l1 = luts + (size_t)ty1 * tiles_x * 256 + v;
    const uint8_t* l2 = luts + (size_t)ty2 * tiles_x * 256 + v;
    const float top = __fadd_rn(__fmul_rn((float)l1[tx1 * 256], xa1), __fmul_rn((float)l1[tx2 * 256], xa));
    const float bot = __fadd_rn(__fmul_rn((float)l2[tx1 * 256], xa1), __fmul_rn((float)l2[tx2 * 256], xa));
    const int r = __float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
    dst[(size_t)y * cols + x] = (uint8_t)min(max(r, 0), 255);
}

// d_src -> d_dst (rows x cols, tight)
static int clahe_device(wsg_handle* h, const uint8_t* d_src, uint8_t* d_dst, int rows, int cols, double clip_limit, int tiles)
{
    if (tiles < 1) { h->err = "CLAHE tile grid size must be positive"; return WSG_ERR_INVALID_ARG; }
    int erows = rows, ecols = cols;
    if (cols % tiles != 0 || rows % tiles != 0) { erows = rows + tiles - rows % tiles; ecols = cols + tiles - cols % tiles; }
    const int tw = ecols / tiles, th = erows / tiles, area = tw * th;
    int clip = 0;
    if (clip_limit > 0.0) clip = std::max((int)(clip_limit * area / 256), 1);
    int rc;
    if ((rc = ensure(h, h->rs_tab, (size_t)tiles * tiles * 256))) return rc;
    clahe_lut_kernel<<<dim3(tiles, tiles), 256, 0, h->stream>>>(d_src, rows, cols, tw, th, clip, 255.f / (float)area, (uint8_t*)h->rs_tab.p);
    dim3 b(128), g((cols + 127) / 128, rows);
    clahe_apply_kernel<<<g, b, 0, h->stream>>>(d_src, rows, cols, tiles, tiles, 1.f / (float)tw, 1.f / (float)th,
                                               (const uint8_t*)h->rs_tab.p, d_dst);
    CK(h, cudaGetLastError());
    return WSG_OK;
}

int wsg_stereo_rectify_custom(const double K0[9], const double K1[9], const double R[9], const double T[3], double rot_angle,
                              int width, int height, double H0[9], double H1[9], int roi[4], double* best_angle)
{
    if (!K0 || !K1 || !R || !T || !H0 || !H1 || !roi || width <= 0 || height <= 0) return WSG_ERR_INVALID_ARG;
    HFunctional hf;
    if (!hf.init(K0, K1, R, T)) return WSG_ERR_INVALID_ARG;
    double angle = rot_angle;
    if (rot_angle == 0) angle = minimise_angle(hf);
    hf.calc(angle);
    if (best_angle) *best_angle = angle;
    memcpy(H0, hf.H0, 72); memcpy(H1, hf.H1, 72);
    // the image corners through both homographies (stereorectify.cpp:168-190)
    const double px[4] = {0, (double)width, (double)width, 0}, py[4] = {0, 0, (double)height, (double)height};
    auto corners = [&](const double* H, double* cx, double* cy) {
        for (int i = 0; i < 4; ++i) {
            const double a = H[0] * px[i] + H[1] * py[i] + H[2] * 1., b = H[3] * px[i] + H[4] * py[i] + H[5] * 1.;
            const double w = H[6] * px[i] + H[7] * py[i] + H[8] * 1.;
            cx[i] = a / w; cy[i] = b / w;
        }
    };
    struct RectD { double x, y, w, h; };
    auto rect_of = [](const double* cx, const double* cy) {     // cv::Rect_<double>(pt1, pt2)
        const double ax = std::min(cx[0], cx[3]), ay = std::min(cy[0], cy[1]), bx = std::max(cx[1], cx[2]), by = std::max(cy[2], cy[3]);
        RectD r; r.x = std::min(ax, bx); r.y = std::min(ay, by); r.w = std::max(ax, bx) - r.x; r.h = std::max(ay, by) - r.y;
        return r;
    };
    double c0x[4], c0y[4], c1x[4], c1y[4];
    corners(H0, c0x, c0y); corners(H1, c1x, c1y);
    const RectD r0 = rect_of(c0x, c0y), r1 = rect_of(c1x, c1y);
    const double top = std::min(r0.y, r1.y), bottom = std::max(r0.y + r0.h, r1.y + r1.h);
    auto finish = [&](double* H, const RectD& r) {
        const double Tr[9] = {1, 0, -r.x, 0, 1, -top, 0, 0, 1};
        const double Sc[9] = {width / r.w, 0, 0, 0, height / (bottom - top), 0, 0, 0, 1};
        double ST[9];
        mm3(Sc, Tr, ST);
        mm3(ST, H, H);
        scale3(H, 1.0 / cbrt(det3(H)));
    };
    finish(H0, r0); finish(H1, r1);
    // the ROI: 4th and 5th of the eight sorted corner coordinates (stereorectify.cpp:215-243), truncated to int as cv::Rect
    corners(H0, c0x, c0y); corners(H1, c1x, c1y);
    double xs[8], ys[8];
    for (int i = 0; i < 4; ++i) { xs[2 * i] = c0x[i]; ys[2 * i] = c0y[i]; xs[2 * i + 1] = c1x[i]; ys[2 * i + 1] = c1y[i]; }
    std::sort(xs, xs + 8); std::sort(ys, ys + 8);
    for (int i = 0; i < 8; ++i) if (!std::isfinite(xs[i]) || !std::isfinite(ys[i])) return WSG_ERR_INVALID_ARG;
    roi[0] = (int)xs[3]; roi[1] = (int)ys[3];
    roi[2] = (int)(xs[4] - roi[0]); roi[3] = (int)(ys[4] - roi[1]);
    return WSG_OK;
}

int wsg_warp_perspective(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, const double H[9], uint8_t* out)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!img || !H || !out || rows <= 0 || cols <= 0 || stride < (size_t)cols) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    double M[9];
    if (!inv3_cv(H, M)) { h->err = "singular homography"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->im_left, n))) return rc;
    if ((rc = ensure(h, h->im_right, n))) return rc;
    if ((rc = ensure(h, h->m_scratch, n * sizeof(int2) + 128))) return rc;
    int2* d_map = (int2*)h->m_scratch.p;
    double* d_M = (double*)((char*)h->m_scratch.p + n * sizeof(int2));
    CK(h, cudaMemcpyAsync(d_M, M, 72, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));        // M is on the stack
    CK(h, cudaMemcpy2DAsync(h->im_left.p, cols, img, stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    // OpenCV's block width: bh0 = min(32, rows); bw0 = min(4096 / bh0, cols)
    const int bh0 = std::min(32, rows), bw0 = std::min(64 * 64 / bh0, cols);
    dim3 b(128), g((cols + 127) / 128, rows);
    warp_map_kernel<<<g, b, 0, h->stream>>>(rows, cols, bw0, d_M, d_map);
    undistort_remap_kernel<<<g, b, 0, h->stream>>>((const uint8_t*)h->im_left.p, rows, cols, cols, d_map, (uint8_t*)h->im_right.p);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(out, h->im_right.p, n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

int wsg_prepare_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, int clahe_tiles, double clahe_clip,
                      const double K[9], const double* dist, int ndist, uint8_t* out)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!img || !out || rows <= 0 || cols <= 0 || stride < (size_t)cols || ndist < 0 || (ndist && !dist) || (!K && ndist >= 0 && clahe_tiles <= 0)) {
        h->err = "bad argument"; return WSG_ERR_INVALID_ARG;
    }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->im_left, n))) return rc;
    if ((rc = ensure(h, h->im_right, n))) return rc;
    uint8_t *a = (uint8_t*)h->im_left.p, *b = (uint8_t*)h->im_right.p;
    CK(h, cudaMemcpy2DAsync(a, cols, img, stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    if (clahe_tiles > 0) {                       // wass_prepare.cpp:257-262
        if ((rc = clahe_device(h, a, b, rows, cols, clahe_clip, clahe_tiles))) return rc;
        std::swap(a, b);
    }
    if (K) {                                     // wass_prepare.cpp:268
        if ((rc = undistort_device(h, a, b, rows, cols, K, dist, ndist))) return rc;
        std::swap(a, b);
    }
    CK(h, cudaMemcpyAsync(out, a, n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

int wsg_clahe_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, double clip_limit, int tiles, uint8_t* out)
{
    if (h && tiles < 1) { h->err = "CLAHE tile grid size must be positive"; return WSG_ERR_INVALID_ARG; }
    return wsg_prepare_image(h, img, rows, cols, stride, tiles, clip_limit, nullptr, nullptr, 0, out);
}

int wsg_undistort_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, const double K[9], const double* dist,
                        int ndist, uint8_t* out)
{
    if (h && !K) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    return wsg_prepare_image(h, img, rows, cols, stride, 0, 0.0, K, dist, ndist, out);
}

int wsg_rectify_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, const double K[9], const double Rrect[9],
                      const double P[12], uint8_t* out)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!img || !K || !Rrect || !P || !out || rows <= 0 || cols <= 0 || stride < (size_t)cols) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    double A[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]}, AR[9], iR[9];
    mat3_mul(A, Rrect, AR);
    if (!inv3(AR, iR)) { h->err = "singular rectification"; return WSG_ERR_INVALID_ARG; }
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->im_left, n))) return rc;
    if ((rc = ensure(h, h->im_right, n))) return rc;
    if ((rc = ensure(h, h->m_small, 1 << 16))) return rc;
    static short itab[TABSZ * TABSZ * 16];
    static bool have = false;
    if (!have) { build_cubic_table(itab); have = true; }
    char* small = (char*)h->m_small.p;
    double* d_ir = (double*)(small + 1024);
    short* d_tab = (short*)(small + 2048);
    static_assert(2048 + sizeof(itab) <= (1 << 16), "table fits the scratch buffer");
    CK(h, cudaMemcpyAsync(d_ir, iR, 72, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(d_tab, itab, sizeof(itab), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpy2DAsync(h->im_left.p, cols, img, stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    dim3 b(128), g((cols + 127) / 128, rows);
    rectify_remap_kernel<<<g, b, 0, h->stream>>>((const uint8_t*)h->im_left.p, rows, cols, cols, d_ir, K[0], K[4], K[2], K[5], d_tab,
                                                 (uint8_t*)h->im_right.p);
    CK(h, cudaMemcpyAsync(out, h->im_right.p, n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    return WSG_OK;
}

}  // extern "C"
