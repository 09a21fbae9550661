/*
 * oracle/sgbm_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the arithmetic of cv::StereoSGBM::compute, the
 * third-party routine the reference calls for its dense matcher
 * (reference call sites: src/wass_stereo/wass_stereo.cpp:775-782 create+setters,
 * :837 compute(right_image,left_image,disparity)).  OpenCV is not vendored in
 * the reference tree (pinned libopencv==4.5.5 in meta.yaml:12-13,20-21); the
 * algorithm restated here is the published StereoSGBM algorithm as specified
 * in SURVEY.md Appendix A.  Parity is PINNED: tests/test_oracle_golden.py
 * checks this file bit-for-bit against outputs of the real cv2.StereoSGBM
 * (opencv-python-headless 4.13.0) stored under tests/golden/ by
 * tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.
 *
 * Verified domain: max(C)+P2 <= 32767; the function returns
 * the observed max(C) so the caller can check.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int minDisparity;
    int numDisparities;
    int blockSize;
    int P1;
    int P2;
    int disp12MaxDiff;
    int preFilterCap;
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int mode; /* 0 = MODE_SGBM (5 paths), 1 = MODE_HH (8 paths) */
} sgbm_oracle_params;

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

/* A.1 prefilter + A.2 Birchfield-Tomasi half-pixel bounds.
 * out planes (each rows*cols int16): p[ch], lo[ch], hi[ch], ch=0 (clipped x-Sobel), 1 (raw) */
static void prefilter(const uint8_t* img, int rows, int cols, size_t stride, int ftzero,
                      int16_t* p[2], int16_t* lo[2], int16_t* hi[2])
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++) {
        const uint8_t* r0 = img + (size_t)y * stride;
        const uint8_t* rm = img + (size_t)(y > 0 ? y - 1 : 0) * stride;
        const uint8_t* rp = img + (size_t)(y < rows - 1 ? y + 1 : rows - 1) * stride;
        int16_t* p0 = p[0] + (size_t)y * cols;
        int16_t* p1 = p[1] + (size_t)y * cols;
        for (int x = 0; x < cols; x++) {
            if (x == 0 || x == cols - 1) { p0[x] = (int16_t)ftzero; p1[x] = (int16_t)ftzero; continue; }
            int g = 2 * (r0[x + 1] - r0[x - 1]) + (rm[x + 1] - rm[x - 1]) + (rp[x + 1] - rp[x - 1]);
            p0[x] = (int16_t)(clampi(g, -ftzero, ftzero) + ftzero);
            p1[x] = r0[x];
        }
        for (int ch = 0; ch < 2; ch++) {
            const int16_t* q = p[ch] + (size_t)y * cols;
            int16_t* l = lo[ch] + (size_t)y * cols;
            int16_t* h = hi[ch] + (size_t)y * cols;
            for (int x = 0; x < cols; x++) {
                int c = q[x];
                int a = x > 0 ? (c + q[x - 1]) / 2 : c;
                int b = x < cols - 1 ? (c + q[x + 1]) / 2 : c;
                l[x] = (int16_t)imin(c, imin(a, b));
                h[x] = (int16_t)imax(c, imax(a, b));
            }
        }
    }
}

/* one aggregation step (A.4).  Lp: predecessor L (D values) or NULL for out-of-image, mp its min. */
static inline int agg_step(const int16_t* Lp, int mp, const int16_t* C, int16_t* L, int16_t* S,
                           int D, int P1, int P2)
{
    int mnew = 32767;
    if (!Lp) {
        /* predecessor outside the image: L=0, m=0  =>  t=0, L=C */
        for (int d = 0; d < D; d++) {
            int v = C[d];
            L[d] = (int16_t)v;
            S[d] = (int16_t)sat16(S[d] + v);
            mnew = imin(mnew, v);
        }
        return mnew;
    }
    int delta = sat16(mp + P2);
    for (int d = 0; d < D; d++) {
        int lm = d > 0 ? Lp[d - 1] : 32767;
        int lp = d < D - 1 ? Lp[d + 1] : 32767;
        int t = imin(imin(Lp[d], sat16(lm + P1)), imin(sat16(lp + P1), delta));
        int v = sat16(sat16(t - mp) + C[d]);
        L[d] = (int16_t)v;            /* NB: L may alias Lp only if caller double-buffers */
        S[d] = (int16_t)sat16(S[d] + v);
        mnew = imin(mnew, v);
    }
    return mnew;
}

/* aggregate one direction with predecessor offset (px,py) over the whole volume, adding into S */
static void aggregate_dir(const int16_t* C, int16_t* S, int H, int W1, int D, int px, int py, int P1, int P2)
{
    const size_t rowsz = (size_t)W1 * D;
    if (py == 0) {
        /* horizontal: rows independent */
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            int16_t* bufA = (int16_t*)malloc(sizeof(int16_t) * D * 2);
            int16_t* bufB = bufA + D;
            int16_t* prev = NULL; int mp = 0;
            int x0 = px < 0 ? 0 : W1 - 1, x1 = px < 0 ? W1 : -1, sx = px < 0 ? 1 : -1;
            int16_t* cur = bufA;
            for (int x = x0; x != x1; x += sx) {
                size_t o = (size_t)y * rowsz + (size_t)x * D;
                mp = agg_step(prev, mp, C + o, cur, S + o, D, P1, P2);
                prev = cur; cur = (cur == bufA) ? bufB : bufA;
            }
            free(bufA);
        }
        return;
    }
    /* vertical / diagonal: row-sequential, all x of a row independent */
    int16_t* Lrow[2]; int* mrow[2];
    Lrow[0] = (int16_t*)malloc(sizeof(int16_t) * rowsz); Lrow[1] = (int16_t*)malloc(sizeof(int16_t) * rowsz);
    mrow[0] = (int*)malloc(sizeof(int) * W1); mrow[1] = (int*)malloc(sizeof(int) * W1);
    int y0 = py < 0 ? 0 : H - 1, y1 = py < 0 ? H : -1, sy = py < 0 ? 1 : -1;
    int cur = 0;
    for (int y = y0; y != y1; y += sy) {
        int prevrow_ok = (y != y0);
        int16_t* Lc = Lrow[cur]; int16_t* Lpv = Lrow[cur ^ 1];
        int* mc = mrow[cur]; int* mpv = mrow[cur ^ 1];
#pragma omp parallel for schedule(static)
        for (int x = 0; x < W1; x++) {
            int xp = x + px;
            const int16_t* Lp = (prevrow_ok && xp >= 0 && xp < W1) ? Lpv + (size_t)xp * D : NULL;
            int mp = Lp ? mpv[xp] : 0;
            size_t o = (size_t)y * rowsz + (size_t)x * D;
            mc[x] = agg_step(Lp, mp, C + o, Lc + (size_t)x * D, S + o, D, P1, P2);
        }
        cur ^= 1;
    }
    free(Lrow[0]); free(Lrow[1]); free(mrow[0]); free(mrow[1]);
}

static void median3x3_s16(const int16_t* src, int16_t* dst, int rows, int cols)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; y++) {
        for (int x = 0; x < cols; x++) {
            int16_t v[9]; int n = 0;
            for (int j = -1; j <= 1; j++) {
                int yy = clampi(y + j, 0, rows - 1);
                for (int i = -1; i <= 1; i++) {
                    int xx = clampi(x + i, 0, cols - 1);
                    v[n++] = src[(size_t)yy * cols + xx];
                }
            }
            for (int a = 1; a < 9; a++) { int16_t k = v[a]; int b = a - 1; while (b >= 0 && v[b] > k) { v[b + 1] = v[b]; b--; } v[b + 1] = k; }
            dst[(size_t)y * cols + x] = v[4];
        }
    }
}

/* cv::filterSpeckles restatement (4-connected flood fill, |diff|<=maxDiff), used only when
 * speckleWindowSize>0 (off at WASS defaults: wass_stereo.cpp:759). */
static void filter_speckles(int16_t* img, int rows, int cols, int newVal, int maxSpeckleSize, int maxDiff)
{
    size_t npix = (size_t)rows * cols;
    int* labels = (int*)calloc(npix, sizeof(int));
    int* stack = (int*)malloc(npix * sizeof(int));
    uint8_t* rtype = (uint8_t*)calloc(npix + 1, 1);
    int curlabel = 0;
    for (int i = 0; i < rows; i++) {
        for (int j = 0; j < cols; j++) {
            size_t idx = (size_t)i * cols + j;
            if (img[idx] == newVal) continue;
            if (labels[idx]) { if (rtype[labels[idx]]) img[idx] = (int16_t)newVal; continue; }
            int sp = 0; stack[sp++] = (int)idx; curlabel++; int count = 0; labels[idx] = curlabel;
            while (sp > 0) {
                int p = stack[--sp]; count++;
                int py = p / cols, pxx = p % cols; int dp = img[p];
                const int dy[4] = {1, -1, 0, 0}, dx[4] = {0, 0, 1, -1};
                for (int k = 0; k < 4; k++) {
                    int yy = py + dy[k], xx = pxx + dx[k];
                    if (yy < 0 || yy >= rows || xx < 0 || xx >= cols) continue;
                    int q = yy * cols + xx;
                    if (!labels[q] && img[q] != newVal && abs(dp - img[q]) <= maxDiff) { labels[q] = curlabel; stack[sp++] = q; }
                }
            }
            if (count <= maxSpeckleSize) { rtype[curlabel] = 1; img[idx] = (int16_t)newVal; }
            else rtype[curlabel] = 0;
        }
    }
    free(labels); free(stack); free(rtype);
}

/*
 * Returns max(C) (>=0) on success, negative on bad arguments.
 * disp: rows*cols int16 (x16 fixed point), final output (after median / speckle).
 * C_out, S_out: optional H*W1*D int16 volumes (layout [y][x-minX1][d-minD]); raw_out: optional
 * rows*cols disparity BEFORE the 3x3 median.
 */
int sgbm_oracle_compute(const uint8_t* img1, const uint8_t* img2, int rows, int cols, size_t stride,
                        const sgbm_oracle_params* prm, int16_t* disp,
                        int16_t* C_out, int16_t* S_out, int16_t* raw_out)
{
    if (!img1 || !img2 || !prm || !disp || rows <= 0 || cols <= 0) return -1;
    const int minD = prm->minDisparity, D = prm->numDisparities;
    if (D <= 0 || D % 16) return -2;
    const int maxD = minD + D;
    const int SW2 = prm->blockSize > 0 ? prm->blockSize / 2 : 1, SH2 = SW2;
    const int ftzero = imax(prm->preFilterCap, 15) | 1;
    const int P1 = prm->P1 > 0 ? prm->P1 : 2;
    const int P2 = imax(prm->P2 > 0 ? prm->P2 : 5, P1 + 1);
    const int uniq = prm->uniquenessRatio >= 0 ? prm->uniquenessRatio : 10;
    const int d12 = prm->disp12MaxDiff > 0 ? prm->disp12MaxDiff : 1;
    const int minX1 = imax(maxD, 0), maxX1 = cols + imin(minD, 0);
    const int W1 = maxX1 - minX1;
    const int INVALID = (minD - 1) * 16;
    const int H = rows, W = cols;
    const size_t npix = (size_t)rows * cols;

    int16_t* raw = (int16_t*)malloc(npix * sizeof(int16_t));
    for (size_t i = 0; i < npix; i++) raw[i] = (int16_t)INVALID;
    int maxC = 0;

    if (W1 > 0) {
        int16_t *p1[2], *lo1[2], *hi1[2], *p2[2], *lo2[2], *hi2[2];
        for (int ch = 0; ch < 2; ch++) {
            p1[ch] = (int16_t*)malloc(npix * 2); lo1[ch] = (int16_t*)malloc(npix * 2); hi1[ch] = (int16_t*)malloc(npix * 2);
            p2[ch] = (int16_t*)malloc(npix * 2); lo2[ch] = (int16_t*)malloc(npix * 2); hi2[ch] = (int16_t*)malloc(npix * 2);
        }
        prefilter(img1, rows, cols, stride, ftzero, p1, lo1, hi1);
        prefilter(img2, rows, cols, stride, ftzero, p2, lo2, hi2);

        const size_t rowsz = (size_t)W1 * D, vol = rowsz * H;
        int16_t* HS = (int16_t*)malloc(vol * 2);
        int16_t* C = (int16_t*)malloc(vol * 2);
        int16_t* S = (int16_t*)calloc(vol, 2);

        /* A.2 pixel cost + horizontal half of the A.3 box (replicate clamp in W1 space) */
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            uint8_t* pd = (uint8_t*)malloc(rowsz);
            size_t ro = (size_t)y * W;
            for (int xh = 0; xh < W1; xh++) {
                int x = xh + minX1;
                for (int dd = 0; dd < D; dd++) {
                    int xp = x - (dd + minD);
                    int acc = 0;
                    for (int ch = 0; ch < 2; ch++) {
                        int u = p1[ch][ro + x], ul = lo1[ch][ro + x], uh = hi1[ch][ro + x];
                        int v = p2[ch][ro + xp], vl = lo2[ch][ro + xp], vh = hi2[ch][ro + xp];
                        int c0 = imax(0, imax(u - vh, vl - u));
                        int c1 = imax(0, imax(v - uh, ul - v));
                        int c = imin(c0, c1);
                        acc += ch == 0 ? c : (c >> 2);
                    }
                    pd[(size_t)xh * D + dd] = (uint8_t)acc;
                }
            }
            int16_t* hs = HS + (size_t)y * rowsz;
            for (int xh = 0; xh < W1; xh++)
                for (int dd = 0; dd < D; dd++) {
                    int s = 0;
                    for (int i = -SW2; i <= SW2; i++) s += pd[(size_t)clampi(xh + i, 0, W1 - 1) * D + dd];
                    hs[(size_t)xh * D + dd] = (int16_t)s;
                }
            free(pd);
        }
        /* vertical half of the box (replicate clamp in y) */
#pragma omp parallel for schedule(static) reduction(max : maxC)
        for (int y = 0; y < H; y++) {
            int16_t* c = C + (size_t)y * rowsz;
            for (size_t k = 0; k < rowsz; k++) {
                int s = 0;
                for (int j = -SH2; j <= SH2; j++) s += HS[(size_t)clampi(y + j, 0, H - 1) * rowsz + k];
                c[k] = (int16_t)s;
                if (s > maxC) maxC = s;
            }
        }
        free(HS);
        for (int ch = 0; ch < 2; ch++) { free(p1[ch]); free(lo1[ch]); free(hi1[ch]); free(p2[ch]); free(lo2[ch]); free(hi2[ch]); }

        /* A.4 path aggregation: predecessor offsets (px,py) */
        static const int dirs[8][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
        int ndirs = prm->mode == 1 ? 8 : 5;
        for (int r = 0; r < ndirs; r++) aggregate_dir(C, S, H, W1, D, dirs[r][0], dirs[r][1], P1, P2);

        /* A.5 WTA / uniqueness / sub-pixel / disp2, A.6 LR check */
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            int16_t* d1 = raw + (size_t)y * W;
            int* disp2 = (int*)malloc(sizeof(int) * W * 2);
            int* disp2cost = disp2 + W;
            for (int x = 0; x < W; x++) { disp2[x] = INVALID; disp2cost[x] = 32767; } /* scaled INVALID, as cv2 does (verified) */
            for (int xh = W1 - 1; xh >= 0; xh--) {
                const int16_t* Sp = S + (size_t)y * rowsz + (size_t)xh * D;
                int best = 0, minS = Sp[0];
                for (int d = 1; d < D; d++) if (Sp[d] < minS) { minS = Sp[d]; best = d; }
                int d;
                for (d = 0; d < D; d++)
                    if (Sp[d] * (100 - uniq) < minS * 100 && abs(best - d) > 1) break;
                if (d < D) continue;
                int x2 = xh + minX1 - best - minD;
                if (disp2cost[x2] > minS) { disp2cost[x2] = minS; disp2[x2] = best + minD; }
                int dd;
                if (best > 0 && best < D - 1) {
                    int den = imax(Sp[best - 1] + Sp[best + 1] - 2 * Sp[best], 1);
                    dd = best * 16 + ((Sp[best - 1] - Sp[best + 1]) * 16 + den) / (2 * den); /* C '/' truncates toward zero */
                } else dd = best * 16;
                d1[xh + minX1] = (int16_t)(dd + minD * 16);
            }
            for (int x = minX1; x < maxX1; x++) {
                int dv = d1[x];
                if (dv == INVALID) continue;
                int a = dv >> 4, b = (dv + 15) >> 4;
                int xa = x - a, xb = x - b;
                if (0 <= xa && xa < W && disp2[xa] >= minD && abs(disp2[xa] - a) > d12 &&
                    0 <= xb && xb < W && disp2[xb] >= minD && abs(disp2[xb] - b) > d12)
                    d1[x] = (int16_t)INVALID;
            }
            free(disp2);
        }
        if (C_out) memcpy(C_out, C, vol * 2);
        if (S_out) memcpy(S_out, S, vol * 2);
        free(C); free(S);
    }
    if (raw_out) memcpy(raw_out, raw, npix * 2);
    median3x3_s16(raw, disp, rows, cols);
    free(raw);
    if (prm->speckleWindowSize > 0)
        filter_speckles(disp, rows, cols, INVALID, prm->speckleWindowSize, 16 * prm->speckleRange);
    return maxC;
}
