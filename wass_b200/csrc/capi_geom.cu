// C ABI, part 2: the dense-stereo stage as a whole, triangulation and the PovMesh operations.
#include "handle.cuh"

#include <cmath>

namespace {

MeshView mesh_view(wsg_handle* h)
{
    MeshView m;
    m.w = h->mesh_w; m.h = h->mesh_h;
    m.valid = (uint8_t*)h->m_valid.p; m.X = (double*)h->m_X.p; m.Y = (double*)h->m_Y.p; m.Z = (double*)h->m_Z.p;
    m.color = (uint8_t*)h->m_color.p;
    return m;
}

int alloc_mesh(wsg_handle* h, int w, int hh)
{
    const size_t n = (size_t)w * hh;
    int rc;
    if ((rc = ensure(h, h->m_valid, n))) return rc;
    if ((rc = ensure(h, h->m_color, n))) return rc;
    if ((rc = ensure(h, h->m_X, n * 8))) return rc;
    if ((rc = ensure(h, h->m_Y, n * 8))) return rc;
    if ((rc = ensure(h, h->m_Z, n * 8))) return rc;
    if ((rc = ensure(h, h->m_small, 1 << 16))) return rc;
    h->mesh_w = w; h->mesh_h = hh;
    return WSG_OK;
}

int need_mesh(wsg_handle* h)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!h->have_mesh) { h->err = "no mesh: call wsg_triangulate or wsg_mesh_upload first"; return WSG_ERR_STATE; }
    if (cudaSetDevice(h->device) != cudaSuccess) return WSG_ERR_CUDA;
    return WSG_OK;
}

// saturate_cast<int>(n * factor) of cv::resize's dsize (round half to even)
int resized_len(int n, double factor) { return (int)nearbyint((double)n * factor); }

// device part of wass_stereo.cpp:853-928 on the ROI disparity (int16 x16, `cols_full` per row, ROI starting at x0, size
// rows x width = the matcher's (resized) input); the result, out_rows x out_cols (roi_comb_right's size), ends up in h->fa
int postprocess_device(wsg_handle* h, const int16_t* d_disp16, int rows, int cols_full, int x0, int width, int mindisp,
                       int ndisp, int disp_offset, double dense_scale, int dilate, int erode, int out_rows, int out_cols)
{
    const bool same = out_rows == rows && out_cols == width;
    if (dense_scale == 1.0 && !same) { h->err = "output size differs from the input size at DENSE_SCALE == 1"; return WSG_ERR_INVALID_ARG; }
    const size_t n = (size_t)rows * width, no = (size_t)out_rows * out_cols, nmax = std::max(n, no);
    int rc;
    if ((rc = ensure(h, h->fa, nmax * 4))) return rc;
    if ((rc = ensure(h, h->fb, nmax * 4))) return rc;
    float* a = (float*)h->fa.p;
    float* b = (float*)h->fb.p;
    StageTimer t(h, WSG_STAGE_POSTFILTER, 2 + std::max(dilate, 0) + std::max(erode, 0) + (same ? 0 : 5));
    launch_clean_convert(d_disp16, rows, cols_full, x0, width, mindisp, ndisp, disp_offset, 1.0 / dense_scale, a, h->stream);
    for (int s = 0; s < dilate; ++s) { launch_dilate_zero(a, b, rows, width, h->stream); std::swap(a, b); }
    for (int s = 0; s < erode; ++s) { launch_erode_zero(a, b, rows, width, h->stream); std::swap(a, b); }
    if (same) {
        launch_mask_by_eroded(a, b, rows, width, h->stream);
        std::swap(a, b);
    } else {
        // nearest-neighbour and bicubic enlargements; the eroded nearest-neighbour map masks the bicubic one (:903-928)
        if ((rc = ensure(h, h->fc, no * 4))) return rc;
        if ((rc = ensure(h, h->rs_tab, resize_tab_bytes(out_cols, out_rows)))) return rc;
        float* c = (float*)h->fc.p;
        launch_resize_nn_f32(a, width, rows, b, out_cols, out_rows, h->stream);             // b = nn
        launch_resize_cubic_f32(a, width, rows, c, out_cols, out_rows, h->rs_tab.p, h->stream);   // c = cubic
        launch_erode_zero(b, a, out_rows, out_cols, h->stream);                            // a = eroded nn
        launch_mask_where_zero(c, a, no, b, h->stream);                                    // b = result
        a = b;
    }
    if (a != (float*)h->fa.p) std::swap(h->fa, h->fb);   // result always ends up in h->fa
    CK(h, cudaGetLastError());
    h->dense_rows = out_rows; h->dense_cols = out_cols; h->have_dense = true;
    return WSG_OK;
}

// device part of wass_stereo.cpp:941-986 on the float ROI disparity held in h->fa (result back in h->fa)
int refine_device(wsg_handle* h, int rows, int cols, int median_wsize, int bc_threshold)
{
    if (median_wsize < 3 && bc_threshold <= 0) return WSG_OK;
    if (median_wsize >= 3 && median_wsize != 3 && median_wsize != 5) {
        h->err = "MEDIAN_FILTER_WSIZE must be 3 or 5 (cv::medianBlur on float32 images accepts nothing else)";
        return WSG_ERR_INVALID_ARG;
    }
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->fb, n * 4))) return rc;
    StageTimer t(h, WSG_STAGE_POSTFILTER, (median_wsize >= 3 ? 1 : 0) + (bc_threshold > 0 ? 6 : 0));
    if (median_wsize >= 3) {
        launch_median_f32((const float*)h->fa.p, (float*)h->fb.p, rows, cols, median_wsize, h->stream);
        std::swap(h->fa, h->fb);
    }
    if (bc_threshold > 0) {
        if ((rc = ensure(h, h->m_labels, n * 12 + 64))) return rc;      // labels, counts, keys, best
        launch_gradient_mask((const float*)h->fa.p, (float*)h->fb.p, rows, cols, (float)bc_threshold, h->stream);
        std::swap(h->fa, h->fb);
        int* L = (int*)h->m_labels.p;
        unsigned* cnt = (unsigned*)(L + n);
        unsigned* key = cnt + n;
        unsigned long long* best = (unsigned long long*)(((uintptr_t)(key + n) + 15) & ~(uintptr_t)15);
        launch_keep_biggest_cc8((float*)h->fa.p, rows, cols, L, cnt, key, best, h->stream);
    }
    CK(h, cudaGetLastError());
    return WSG_OK;
}

void jacobi_eigen3(double A[3][3], double V[3][3], double w[3])
{
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = i == j;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

}  // namespace

extern "C" {

void wsg_dense_params_default(wsg_dense_params* p)
{
    if (!p) return;
    // defaults of wass_stereo.cpp:742-761
    p->MIN_DISPARITY = 1; p->MAX_DISPARITY = 640; p->WINSIZE = 13; p->DENSE_SCALE = 1.0; p->DISPARITY_OFFSET = 0;
    p->DISP_DILATE_STEPS = 1; p->DISP_EROSION_STEPS = 2; p->DENSE_P1_MULT = 2; p->DENSE_P2_MULT = 64;
    p->DENSE_UNIQUENESS_RATIO = 1; p->DENSE_DISP12MAXDIFF = -1; p->DENSE_PREFILTER_CAP = 60; p->DENSE_SPECKLE_RANGE = 16;
    p->DENSE_SPECKLE_WINDOW_SIZE = -70; p->mode = WSG_MODE_SGBM;
    p->MEDIAN_FILTER_WSIZE = 0; p->DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD = 0;
}

void wsg_tri_params_default(wsg_tri_params* p)
{
    if (!p) return;
    // defaults of wass_stereo.cpp:1030-1037
    p->TRIANG_MIN_ANGLE = 20.0; p->TRIANG_BBOX_TOP = p->TRIANG_BBOX_LEFT = p->TRIANG_BBOX_RIGHT = p->TRIANG_BBOX_BOTTOM = -1.0;
    p->DISCARD_BURNED_AREAS = 1; p->disparity_compensation = 0; p->DENSE_SCALE = 1.0; p->cam_distance = 1.0;
}

void wsg_refine_params_default(wsg_refine_params* p)
{
    if (!p) return;
    // defaults of wass_stereo.cpp:62-66 and PovMesh.cpp:577-579
    p->PLANE_REFINE_XMIN = -9999; p->PLANE_REFINE_XMAX = 9999; p->PLANE_REFINE_YMIN = -9999; p->PLANE_REFINE_YMAX = 9999;
    p->PLANE_REFINEMENT_MAX_DISTANCE = 70.0; p->PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE = 1; p->PLANE_USE_CENTRAL_THIRD_ONLY = 0;
}

void wsg_dense_scaled_size(int rows, int cols, double dense_scale, int* rows_s, int* cols_s)
{
    // wass_stereo.cpp:788-797: x only when enlarging, both axes when shrinking
    if (rows_s) *rows_s = dense_scale < 1.0 ? resized_len(rows, dense_scale) : rows;
    if (cols_s) *cols_s = dense_scale != 1.0 ? resized_len(cols, dense_scale) : cols;
}

// The dense stage on n pairs of crops of one size (n == 1: wsg_dense_stereo).  Per frame: upload, resize, pad; then ONE
// batched matcher run (run_sgbm_batch: the sweeps walk all frames in one launch each); then per frame the clean-up filters.
// The float ROI disparity of frame f is left in h->fbatch at f * rows * cols; the last one also in h->fa.
static int dense_stereo_batch(wsg_handle* h, int n, const uint8_t* const* left_crops, const uint8_t* const* right_crops, int rows, int cols,
                              size_t stride, const wsg_dense_params* p, float* const* disp_roi, int16_t* disp16_roi)
{
    const double scale = p->DENSE_SCALE;
    if (!(scale > 0.0) || !std::isfinite(scale)) { h->err = "DENSE_SCALE must be positive"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    int srows, scols;                      // size of the matcher's input
    wsg_dense_scaled_size(rows, cols, scale, &srows, &scols);
    if (srows < 1 || scols < 1) { h->err = "DENSE_SCALE leaves no pixels"; return WSG_ERR_INVALID_ARG; }
    const int N = p->MAX_DISPARITY;
    const int off = std::max(p->DISPARITY_OFFSET, 0), comp = std::max(-p->DISPARITY_OFFSET, 0);
    if (comp > N + off) { h->err = "DISPARITY_OFFSET too negative"; return WSG_ERR_INVALID_ARG; }
    const int wp = scols + N + off;
    wsg_sgbm_params sp;
    sp.minDisparity = p->MIN_DISPARITY; sp.numDisparities = N; sp.blockSize = p->WINSIZE;
    sp.P1 = p->DENSE_P1_MULT * p->WINSIZE * p->WINSIZE; sp.P2 = p->DENSE_P2_MULT * p->WINSIZE * p->WINSIZE;
    sp.disp12MaxDiff = p->DENSE_DISP12MAXDIFF; sp.preFilterCap = p->DENSE_PREFILTER_CAP; sp.uniquenessRatio = p->DENSE_UNIQUENESS_RATIO;
    sp.speckleWindowSize = p->DENSE_SPECKLE_WINDOW_SIZE; sp.speckleRange = p->DENSE_SPECKLE_RANGE; sp.mode = p->mode;
    SgbmPlan pl{};
    int rc = wsg_make_plan(h, srows, wp, &sp, pl);
    if (rc) return rc;
    h->plan = pl;
    const size_t ncrop = (size_t)rows * cols, npad = (size_t)srows * wp;
    if ((rc = ensure(h, h->crop_l, ncrop * n))) return rc;
    if ((rc = ensure(h, h->crop_r, ncrop * n))) return rc;
    if ((rc = ensure(h, h->img1, npad * n))) return rc;
    if ((rc = ensure(h, h->img2, npad * n))) return rc;
    if ((rc = ensure(h, h->disp, npad * 2 * n))) return rc;
    if (n > 1 && (rc = ensure(h, h->fbatch, ncrop * 4 * n))) return rc;
    std::vector<const uint8_t*> a(n), b(n);
    std::vector<int16_t*> d(n);
    for (int f = 0; f < n; ++f) {
        uint8_t* cl = (uint8_t*)h->crop_l.p + f * ncrop;
        uint8_t* cr = (uint8_t*)h->crop_r.p + f * ncrop;
        CK(h, cudaMemcpy2DAsync(cl, cols, left_crops[f], stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpy2DAsync(cr, cols, right_crops[f], stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
        const uint8_t *dl = cl, *dr = cr;
        if (scale != 1.0) {
            // cv::resize(..., INTER_CUBIC) of both crops (wass_stereo.cpp:788-797)
            const size_t ns = (size_t)srows * scols;
            if ((rc = ensure(h, h->rs_l, ns))) return rc;
            if ((rc = ensure(h, h->rs_r, ns))) return rc;
            if ((rc = ensure(h, h->rs_tab, resize_tab_bytes(scols, srows)))) return rc;
            const double fy = scale < 1.0 ? scale : 1.0;
            StageTimer t(h, WSG_STAGE_POSTFILTER, 6);
            launch_resize_cubic_u8(dl, cols, cols, rows, scale, fy, (uint8_t*)h->rs_l.p, scols, scols, srows, h->rs_tab.p, h->stream);
            launch_resize_cubic_u8(dr, cols, cols, rows, scale, fy, (uint8_t*)h->rs_r.p, scols, scols, srows, h->rs_tab.p, h->stream);
            dl = (const uint8_t*)h->rs_l.p; dr = (const uint8_t*)h->rs_r.p;
        }
        uint8_t* p1 = (uint8_t*)h->img1.p + f * npad;
        uint8_t* p2 = (uint8_t*)h->img2.p + f * npad;
        launch_pad_images(dl, dr, scols, srows, scols, N, off, comp, p1, p2, wp, h->stream);
        a[f] = p1; b[f] = p2; d[f] = (int16_t*)h->disp.p + f * npad;
    }
    rc = wsg_run_sgbm_batch(h, n, a.data(), b.data(), wp, d.data());
    if (rc) return rc;
    for (int f = 0; f < n; ++f) {
        rc = postprocess_device(h, d[f], srows, wp, N, scols, p->MIN_DISPARITY, N, off, scale,
                                p->DISP_DILATE_STEPS, p->DISP_EROSION_STEPS, rows, cols);
        if (rc) return rc;
        if ((rc = refine_device(h, rows, cols, p->MEDIAN_FILTER_WSIZE, p->DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD))) return rc;
        if (n > 1) CK(h, cudaMemcpyAsync((float*)h->fbatch.p + f * ncrop, h->fa.p, ncrop * 4, cudaMemcpyDeviceToDevice, h->stream));
        if (disp_roi && disp_roi[f]) CK(h, cudaMemcpyAsync(disp_roi[f], h->fa.p, ncrop * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    if (disp16_roi)
        CK(h, cudaMemcpy2DAsync(disp16_roi, (size_t)scols * 2, (const int16_t*)h->disp.p + N, (size_t)wp * 2, (size_t)scols * 2, srows,
                                cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->dense_batch = n; h->dense_batch_rows = rows; h->dense_batch_cols = cols;
    return wsg_check_sweep(h);
}

int wsg_dense_stereo(wsg_handle* h, const uint8_t* left_crop, const uint8_t* right_crop, int rows, int cols, size_t stride,
                     const wsg_dense_params* p, float* disp_roi, int16_t* disp16_roi)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!left_crop || !right_crop || !p || rows <= 0 || cols <= 0 || stride < (size_t)cols) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    return dense_stereo_batch(h, 1, &left_crop, &right_crop, rows, cols, stride, p, disp_roi ? &disp_roi : nullptr, disp16_roi);
}

int wsg_dense_stereo_batch(wsg_handle* h, int n, const uint8_t* const* left_crops, const uint8_t* const* right_crops, int rows, int cols,
                           size_t stride, const wsg_dense_params* p, float* const* disp_roi)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (n <= 0 || n > WSG_MAX_BATCH || !left_crops || !right_crops || !p || rows <= 0 || cols <= 0 || stride < (size_t)cols) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    for (int f = 0; f < n; ++f)
        if (!left_crops[f] || !right_crops[f]) { h->err = "null frame pointer"; return WSG_ERR_INVALID_ARG; }
    return dense_stereo_batch(h, n, left_crops, right_crops, rows, cols, stride, p, disp_roi, nullptr);
}

int wsg_dense_select(wsg_handle* h, int frame)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!h->have_dense || frame < 0 || frame >= h->dense_batch) { h->err = "no such frame in the last dense batch"; return WSG_ERR_STATE; }
    if (h->dense_batch == 1) return WSG_OK;
    CK(h, cudaSetDevice(h->device));
    const size_t ncrop = (size_t)h->dense_batch_rows * h->dense_batch_cols;
    CK(h, cudaMemcpyAsync(h->fa.p, (const float*)h->fbatch.p + frame * ncrop, ncrop * 4, cudaMemcpyDeviceToDevice, h->stream));
    return WSG_OK;
}

int wsg_disparity_postprocess_resized(wsg_handle* h, const int16_t* disp16_roi, int rows, int cols, int minDisparity,
                                      int numDisparities, int disparityOffset, double denseScale, int dilateSteps,
                                      int erosionSteps, float* disp_roi, int out_rows, int out_cols)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!disp16_roi || !disp_roi || rows <= 0 || cols <= 0 || out_rows <= 0 || out_cols <= 0 || !(denseScale > 0.0)) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->disp, n * 2))) return rc;
    CK(h, cudaMemcpyAsync(h->disp.p, disp16_roi, n * 2, cudaMemcpyHostToDevice, h->stream));
    rc = postprocess_device(h, (const int16_t*)h->disp.p, rows, cols, 0, cols, minDisparity, numDisparities,
                            std::max(disparityOffset, 0), denseScale, dilateSteps, erosionSteps, out_rows, out_cols);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(disp_roi, h->fa.p, (size_t)out_rows * out_cols * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

int wsg_disparity_postprocess(wsg_handle* h, const int16_t* disp16_roi, int rows, int cols, int minDisparity, int numDisparities,
                              int disparityOffset, double denseScale, int dilateSteps, int erosionSteps, float* disp_roi)
{
    if (h && denseScale != 1.0) { h->err = "DENSE_SCALE != 1 changes the output size: use wsg_disparity_postprocess_resized"; return WSG_ERR_INVALID_ARG; }
    return wsg_disparity_postprocess_resized(h, disp16_roi, rows, cols, minDisparity, numDisparities, disparityOffset, denseScale,
                                             dilateSteps, erosionSteps, disp_roi, rows, cols);
}

int wsg_resize_u8_cubic(wsg_handle* h, const uint8_t* src, int rows, int cols, size_t stride, double fx, double fy,
                        uint8_t* dst, int dst_rows, int dst_cols)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!src || !dst || rows <= 0 || cols <= 0 || stride < (size_t)cols || !(fx > 0.0) || !(fy > 0.0)) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    if (dst_rows != resized_len(rows, fy) || dst_cols != resized_len(cols, fx) || dst_rows < 1 || dst_cols < 1) {
        h->err = "dst size must be round(rows*fy) x round(cols*fx)"; return WSG_ERR_INVALID_ARG;
    }
    CK(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure(h, h->crop_l, (size_t)rows * cols))) return rc;
    if ((rc = ensure(h, h->rs_l, (size_t)dst_rows * dst_cols))) return rc;
    if ((rc = ensure(h, h->rs_tab, resize_tab_bytes(dst_cols, dst_rows)))) return rc;
    CK(h, cudaMemcpy2DAsync(h->crop_l.p, cols, src, stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    launch_resize_cubic_u8((const uint8_t*)h->crop_l.p, cols, cols, rows, fx, fy, (uint8_t*)h->rs_l.p, dst_cols, dst_cols, dst_rows,
                           h->rs_tab.p, h->stream);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(dst, h->rs_l.p, (size_t)dst_rows * dst_cols, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

int wsg_resize_f32(wsg_handle* h, const float* src, int rows, int cols, float* dst, int dst_rows, int dst_cols, int interpolation)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!src || !dst || rows <= 0 || cols <= 0 || dst_rows <= 0 || dst_cols <= 0 ||
        (interpolation != WSG_INTER_NEAREST && interpolation != WSG_INTER_CUBIC)) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)rows * cols, no = (size_t)dst_rows * dst_cols;
    int rc;
    if ((rc = ensure(h, h->fb, n * 4))) return rc;
    if ((rc = ensure(h, h->fc, no * 4))) return rc;
    if ((rc = ensure(h, h->rs_tab, resize_tab_bytes(dst_cols, dst_rows)))) return rc;
    CK(h, cudaMemcpyAsync(h->fb.p, src, n * 4, cudaMemcpyHostToDevice, h->stream));
    if (interpolation == WSG_INTER_CUBIC)
        launch_resize_cubic_f32((const float*)h->fb.p, cols, rows, (float*)h->fc.p, dst_cols, dst_rows, h->rs_tab.p, h->stream);
    else
        launch_resize_nn_f32((const float*)h->fb.p, cols, rows, (float*)h->fc.p, dst_cols, dst_rows, h->stream);
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(dst, h->fc.p, no * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

int wsg_disparity_refine(wsg_handle* h, float* disp_roi, int rows, int cols, int median_wsize, int bc_threshold)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!disp_roi || rows <= 0 || cols <= 0) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)rows * cols;
    int rc;
    if ((rc = ensure(h, h->fa, n * 4))) return rc;
    CK(h, cudaMemcpyAsync(h->fa.p, disp_roi, n * 4, cudaMemcpyHostToDevice, h->stream));
    if ((rc = refine_device(h, rows, cols, median_wsize, bc_threshold))) return rc;
    CK(h, cudaMemcpyAsync(disp_roi, h->fa.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->have_dense = false;      // h->fa no longer holds the output of wsg_dense_stereo
    return WSG_OK;
}

static int triangulate_common(wsg_handle* h, const float* d_disp_full, const uint8_t* left, const uint8_t* right,
                              const uint8_t* left_mask, const uint8_t* right_mask, const wsg_calib* c, const wsg_tri_params* p,
                              unsigned long long* n_points)
{
    CalibDev cd;
    memcpy(cd.K0, c->K0, 72); memcpy(cd.K1, c->K1, 72); memcpy(cd.R, c->R, 72); memcpy(cd.T, c->T, 24);
    memcpy(cd.R1, c->R1, 72); memcpy(cd.R2, c->R2, 72);
    cd.P1fx = c->P1[0]; cd.P1fy = c->P1[5]; cd.P1cx = c->P1[2]; cd.P1cy = c->P1[6];
    cd.P2fx = c->P2[0]; cd.P2fy = c->P2[5]; cd.P2cx = c->P2[2]; cd.P2cy = c->P2[6];
    cd.rlx = c->roi_left[0]; cd.rly = c->roi_left[1]; cd.rlw = c->roi_left[2]; cd.rlh = c->roi_left[3];
    cd.rrx = c->roi_right[0]; cd.rry = c->roi_right[1]; cd.rrw = c->roi_right[2]; cd.rrh = c->roi_right[3];
    cd.left_cols = c->left_cols; cd.left_rows = c->left_rows; cd.right_cols = c->right_cols; cd.right_rows = c->right_rows;
    cd.rect_cols = c->rect_cols; cd.rect_rows = c->rect_rows;
    cd.min_angle = p->TRIANG_MIN_ANGLE;
    if (p->TRIANG_BBOX_TOP >= 0 && p->TRIANG_BBOX_LEFT >= 0 && p->TRIANG_BBOX_BOTTOM >= 0 && p->TRIANG_BBOX_RIGHT >= 0) {
        cd.bbox_l = p->TRIANG_BBOX_LEFT; cd.bbox_t = p->TRIANG_BBOX_TOP; cd.bbox_r = p->TRIANG_BBOX_RIGHT; cd.bbox_b = p->TRIANG_BBOX_BOTTOM;
    } else {
        cd.bbox_l = 0; cd.bbox_t = 0; cd.bbox_r = c->left_cols; cd.bbox_b = c->left_rows;
    }
    cd.discard_burned = p->DISCARD_BURNED_AREAS; cd.has_lmask = left_mask != nullptr; cd.has_rmask = right_mask != nullptr;
    cd.comp_over_scale = (double)p->disparity_compensation / p->DENSE_SCALE;
    cd.cam_distance = p->cam_distance;
    cd.use_h = c->use_homographies;
    memcpy(cd.HLi, c->HLi, 72); memcpy(cd.HRi, c->HRi, 72);
    if (cd.rrx < 0 || cd.rry < 0 || cd.rrw <= 0 || cd.rrh <= 0 || cd.rrx + cd.rrw > cd.rect_cols || cd.rry + cd.rrh > cd.rect_rows) {
        h->err = "roi_right outside the rectified image"; return WSG_ERR_INVALID_ARG;
    }
    const size_t nl = (size_t)c->left_cols * c->left_rows, nr = (size_t)c->right_cols * c->right_rows;
    int rc;
    if ((rc = ensure(h, h->im_left, nl))) return rc;
    if ((rc = ensure(h, h->im_right, nr))) return rc;
    CK(h, cudaMemcpyAsync(h->im_left.p, left, nl, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpyAsync(h->im_right.p, right, nr, cudaMemcpyHostToDevice, h->stream));
    if (left_mask) { if ((rc = ensure(h, h->mask_l, nl))) return rc; CK(h, cudaMemcpyAsync(h->mask_l.p, left_mask, nl, cudaMemcpyHostToDevice, h->stream)); }
    if (right_mask) { if ((rc = ensure(h, h->mask_r, nr))) return rc; CK(h, cudaMemcpyAsync(h->mask_r.p, right_mask, nr, cudaMemcpyHostToDevice, h->stream)); }
    if ((rc = alloc_mesh(h, cd.rrw, cd.rrh))) return rc;
    const size_t n = (size_t)cd.rrw * cd.rrh;
    MeshView m = mesh_view(h);
    CK(h, cudaMemsetAsync(m.valid, 0, n, h->stream));
    CK(h, cudaMemsetAsync(m.color, 0, n, h->stream));
    CK(h, cudaMemsetAsync(m.X, 0, n * 8, h->stream));
    CK(h, cudaMemsetAsync(m.Y, 0, n * 8, h->stream));
    CK(h, cudaMemsetAsync(m.Z, 0, n * 8, h->stream));
    unsigned long long* counter = (unsigned long long*)h->m_small.p;
    CK(h, cudaMemsetAsync(counter, 0, 8, h->stream));
    {
        StageTimer t(h, WSG_STAGE_TRIANGULATE, 1);
        launch_triangulate(d_disp_full, (const uint8_t*)h->im_left.p, (const uint8_t*)h->im_right.p,
                           left_mask ? (const uint8_t*)h->mask_l.p : nullptr, right_mask ? (const uint8_t*)h->mask_r.p : nullptr,
                           cd, m, counter, h->stream);
    }
    unsigned long long cnt = 0;
    CK(h, cudaMemcpyAsync(&cnt, counter, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    h->have_mesh = true;
    if (n_points) *n_points = cnt;
    return WSG_OK;
}

int wsg_triangulate(wsg_handle* h, const float* disparity, const uint8_t* left, const uint8_t* right, const uint8_t* left_mask,
                    const uint8_t* right_mask, const wsg_calib* calib, const wsg_tri_params* p, unsigned long long* n_points)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!disparity || !left || !right || !calib || !p) { h->err = "null argument"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)calib->rect_cols * calib->rect_rows;
    int rc;
    if ((rc = ensure(h, h->dispfull, n * 4))) return rc;
    CK(h, cudaMemcpyAsync(h->dispfull.p, disparity, n * 4, cudaMemcpyHostToDevice, h->stream));
    return triangulate_common(h, (const float*)h->dispfull.p, left, right, left_mask, right_mask, calib, p, n_points);
}

int wsg_triangulate_from_dense(wsg_handle* h, const uint8_t* left, const uint8_t* right, const uint8_t* left_mask,
                               const uint8_t* right_mask, const wsg_calib* calib, const wsg_tri_params* p, unsigned long long* n_points)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!left || !right || !calib || !p) { h->err = "null argument"; return WSG_ERR_INVALID_ARG; }
    if (!h->have_dense) { h->err = "no dense disparity on the device: call wsg_dense_stereo first"; return WSG_ERR_STATE; }
    if (h->dense_cols != calib->roi_right[2] || h->dense_rows != calib->roi_right[3]) {
        h->err = "dense disparity size does not match roi_right"; return WSG_ERR_INVALID_ARG;
    }
    CK(h, cudaSetDevice(h->device));
    const size_t n = (size_t)calib->rect_cols * calib->rect_rows;
    int rc;
    if ((rc = ensure(h, h->dispfull, n * 4))) return rc;
    // env.disparity = zeros(full); copy ROI (wass_stereo.cpp:990-991)
    launch_paste_roi((const float*)h->fa.p, h->dense_rows, h->dense_cols, (float*)h->dispfull.p, calib->rect_rows, calib->rect_cols,
                     calib->roi_right[0], calib->roi_right[1], h->stream);
    return triangulate_common(h, (const float*)h->dispfull.p, left, right, left_mask, right_mask, calib, p, n_points);
}

int wsg_mesh_upload(wsg_handle* h, int width, int height, const uint8_t* valid, const double* xyz, const uint8_t* grey)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (width <= 0 || height <= 0 || !valid || !xyz) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    int rc = alloc_mesh(h, width, height);
    if (rc) return rc;
    const size_t n = (size_t)width * height;
    std::vector<double> plane(n);
    MeshView m = mesh_view(h);
    double* dst[3] = {m.X, m.Y, m.Z};
    for (int k = 0; k < 3; ++k) {
        for (size_t i = 0; i < n; ++i) plane[i] = xyz[3 * i + k];
        CK(h, cudaMemcpyAsync(dst[k], plane.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    CK(h, cudaMemcpyAsync(m.valid, valid, n, cudaMemcpyHostToDevice, h->stream));
    if (grey) CK(h, cudaMemcpyAsync(m.color, grey, n, cudaMemcpyHostToDevice, h->stream));
    else CK(h, cudaMemsetAsync(m.color, 0, n, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->have_mesh = true;
    return WSG_OK;
}

int wsg_mesh_download(wsg_handle* h, uint8_t* valid, double* xyz, uint8_t* grey)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    const size_t n = (size_t)h->mesh_w * h->mesh_h;
    MeshView m = mesh_view(h);
    if (valid) CK(h, cudaMemcpyAsync(valid, m.valid, n, cudaMemcpyDeviceToHost, h->stream));
    if (grey) CK(h, cudaMemcpyAsync(grey, m.color, n, cudaMemcpyDeviceToHost, h->stream));
    if (xyz) {
        std::vector<double> plane(n);
        const double* src[3] = {m.X, m.Y, m.Z};
        for (int k = 0; k < 3; ++k) {
            CK(h, cudaMemcpyAsync(plane.data(), src[k], n * 8, cudaMemcpyDeviceToHost, h->stream));
            CK(h, cudaStreamSynchronize(h->stream));
            for (size_t i = 0; i < n; ++i) xyz[3 * i + k] = plane[i];
        }
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

int wsg_mesh_size(wsg_handle* h, int* width, int* height, unsigned long long* n_valid)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (width) *width = h->mesh_w;
    if (height) *height = h->mesh_h;
    if (n_valid) {
        unsigned long long* counter = (unsigned long long*)h->m_small.p;
        CK(h, cudaMemsetAsync(counter, 0, 8, h->stream));
        launch_count_valid(mesh_view(h), counter, h->stream);
        CK(h, cudaMemcpyAsync(n_valid, counter, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    return WSG_OK;
}

int wsg_mesh_zgap_percentile(wsg_handle* h, double percentile, double* zgap)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!zgap) return WSG_ERR_INVALID_ARG;
    const size_t bytes = zgap_scratch_bytes(h->mesh_w, h->mesh_h);
    if ((rc = ensure(h, h->m_scratch, bytes))) return rc;
    StageTimer t(h, WSG_STAGE_MESH, 2);
    if (mesh_zgap_percentile(mesh_view(h), percentile, h->m_scratch.p, bytes, zgap, h->stream)) { h->err = "zgap percentile failed"; return WSG_ERR_CUDA; }
    CK(h, cudaGetLastError());
    return WSG_OK;
}

int wsg_mesh_biggest_component(wsg_handle* h, double zgap, unsigned long long* n_left)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    const size_t n = (size_t)h->mesh_w * h->mesh_h;
    if ((rc = ensure(h, h->m_labels, n * 4))) return rc;
    if ((rc = ensure(h, h->m_scratch, 64 + n * 4))) return rc;
    unsigned long long left = 0;
    StageTimer t(h, WSG_STAGE_MESH, 7);
    if (mesh_biggest_component(mesh_view(h), zgap, (int*)h->m_labels.p, (unsigned long long*)h->m_scratch.p, &left, h->stream)) {
        h->err = "connected components failed"; return WSG_ERR_CUDA;
    }
    CK(h, cudaGetLastError());
    if (n_left) *n_left = left;
    return WSG_OK;
}

int wsg_ransac_draw(int width, int height, int rounds, int32_t* triples)
{
    if (width <= 0 || height <= 0 || rounds < 0 || !triples) return WSG_ERR_INVALID_ARG;
    const double mindist = height * 0.01;
    for (int r = 0; r < rounds;) {
        // cv::Vec2i p(rand()%iW, rand()%iH): the two calls are constructor arguments, which GCC -- the compiler of every
        // Linux build of the reference -- evaluates right to left: v is drawn BEFORE u.  Pinned against the reference's own
        // PovMesh.cpp built with GCC (oracle/_ref/povmesh_ref, tests/golden/povmesh_golden.npz).
        int c[6];
        for (int k = 0; k < 3; ++k) { c[2 * k + 1] = rand() % height; c[2 * k] = rand() % width; }
        auto dist = [&](int a, int b) { const double dx = c[2 * a] - c[2 * b], dy = c[2 * a + 1] - c[2 * b + 1]; return sqrt(dx * dx + dy * dy); };
        if (dist(0, 1) < mindist || dist(1, 2) < mindist || dist(0, 2) < mindist) continue;   // "round--; continue" of the reference
        for (int k = 0; k < 6; ++k) triples[6 * r + k] = c[k];
        ++r;
    }
    return WSG_OK;
}

int wsg_mesh_ransac_plane(wsg_handle* h, const int32_t* triples, int n, double threshold, double plane[4], int* ok,
                          unsigned long long* best_inliers)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!triples || n <= 0 || !plane || !ok) { h->err = "bad argument"; return WSG_ERR_INVALID_ARG; }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 6; ++k) {
            const int v = triples[6 * i + k];
            if (v < 0 || v >= ((k & 1) ? h->mesh_h : h->mesh_w)) { h->err = "triple outside the mesh"; return WSG_ERR_INVALID_ARG; }
        }
    const size_t bytes = (size_t)n * (6 * 4 + 4 * 8 + 4 + 8) + 256;
    if ((rc = ensure(h, h->m_scratch, bytes))) return rc;
    char* base = (char*)h->m_scratch.p;
    double* planes = (double*)base;
    unsigned long long* counts = (unsigned long long*)(base + (size_t)n * 32);
    int* d_tr = (int*)(base + (size_t)n * 40);
    int* d_ok = (int*)(base + (size_t)n * 64);
    CK(h, cudaMemcpyAsync(d_tr, triples, (size_t)n * 24, cudaMemcpyHostToDevice, h->stream));
    {
        StageTimer t(h, WSG_STAGE_MESH, 2);
        launch_ransac_planes(mesh_view(h), d_tr, n, planes, d_ok, h->stream);
        launch_ransac_count(mesh_view(h), planes, d_ok, n, threshold, counts, h->stream);
    }
    std::vector<double> hp((size_t)n * 4);
    std::vector<unsigned long long> hc(n);
    std::vector<int> hok(n);
    CK(h, cudaMemcpyAsync(hp.data(), planes, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(hc.data(), counts, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(hok.data(), d_ok, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    unsigned long long best = 0;
    plane[0] = plane[1] = plane[2] = plane[3] = 0;
    for (int i = 0; i < n; ++i)
        if (hok[i] && hc[i] > best) { best = hc[i]; memcpy(plane, &hp[4 * (size_t)i], 32); }   // strict >: first best wins
    *ok = !(best < (unsigned long long)((size_t)h->mesh_w * h->mesh_h / 10));
    if (best_inliers) *best_inliers = best;
    return WSG_OK;
}

int wsg_mesh_crop_plane(wsg_handle* h, const double plane[4], double threshold, unsigned long long* n_left)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!plane) return WSG_ERR_INVALID_ARG;
    unsigned long long* counter = (unsigned long long*)h->m_small.p;
    {
        StageTimer t(h, WSG_STAGE_MESH, 1);
        launch_crop_plane(mesh_view(h), plane[0], plane[1], plane[2], plane[3], threshold, counter, h->stream);
    }
    unsigned long long k = 0;
    CK(h, cudaMemcpyAsync(&k, counter, 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    if (n_left) *n_left = k;
    return WSG_OK;
}

static RefineArgs refine_args(const wsg_refine_params* p, int W, int H)
{
    RefineArgs a;
    a.xmin = p->PLANE_REFINE_XMIN; a.xmax = p->PLANE_REFINE_XMAX; a.ymin = p->PLANE_REFINE_YMIN; a.ymax = p->PLANE_REFINE_YMAX;
    a.maxdist = p->PLANE_REFINEMENT_MAX_DISTANCE; a.weight_by_distance = p->PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE;
    const bool ct = p->PLANE_USE_CENTRAL_THIRD_ONLY != 0;
    a.umin = ct ? W / 4 : 0; a.umax = ct ? W * 3 / 4 : W - 1; a.vmin = ct ? H / 4 : 0; a.vmax = ct ? H * 2 / 3 : H - 1;
    return a;
}

int wsg_mesh_refine_inliers(wsg_handle* h, const wsg_refine_params* p, int every, double* xyz, size_t capacity_points, unsigned long long* n_points,
                            unsigned long long* n_inliers)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!p || !xyz || !n_points || every < 1) return WSG_ERR_INVALID_ARG;
    const int n = h->mesh_w * h->mesh_h;
    const size_t cub_bytes = compact_cub_bytes(n), nmax = ((size_t)n + every - 1) / every;
    if ((rc = ensure(h, h->m_scratch, (size_t)n * 8 + cub_bytes + 512))) return rc;
    if ((rc = ensure(h, h->m_out, nmax * 24 + 16))) return rc;
    char* base = (char*)h->m_scratch.p;
    unsigned long long nin = 0;
    StageTimer t(h, WSG_STAGE_MESH, 3);
    if (mesh_refine_sample(mesh_view(h), refine_args(p, h->mesh_w, h->mesh_h), (unsigned)every, (double*)h->m_out.p, (unsigned*)(base + 256),
                           base + 256 + (size_t)n * 8, cub_bytes, &nin, h->stream)) {
        h->err = "inlier sampling failed"; return WSG_ERR_CUDA;
    }
    const size_t np = (size_t)((nin + every - 1) / every);
    *n_points = np;
    if (n_inliers) *n_inliers = nin;
    if (capacity_points < np) { h->err = "destination too small"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaMemcpyAsync(xyz, h->m_out.p, np * 24, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    return WSG_OK;
}

int wsg_mesh_refine_plane(wsg_handle* h, const wsg_refine_params* p, double plane[4], unsigned long long* n_inliers)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!p || !plane) return WSG_ERR_INVALID_ARG;
    const int W = h->mesh_w, H = h->mesh_h;
    const RefineArgs a = refine_args(p, W, H);
    MeshView m = mesh_view(h);
    const int nb = refine_blocks(m);
    if ((rc = ensure(h, h->m_scratch, (size_t)nb * 6 * 8))) return rc;
    double* part = (double*)h->m_scratch.p;
    std::vector<double> hp((size_t)nb * 6);
    StageTimer t(h, WSG_STAGE_MESH, 2);
    launch_refine_pass1(m, a, part, h->stream);
    CK(h, cudaMemcpyAsync(hp.data(), part, (size_t)nb * 5 * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    double s[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < nb; ++b) for (int k = 0; k < 5; ++k) s[k] += hp[(size_t)b * 5 + k];
    if (n_inliers) *n_inliers = (unsigned long long)s[4];
    const double cx = s[1] / s[0], cy = s[2] / s[0], cz = s[3] / s[0];
    launch_refine_pass2(m, a, cx, cy, cz, part, h->stream);
    CK(h, cudaMemcpyAsync(hp.data(), part, (size_t)nb * 6 * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    double q[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nb; ++b) for (int k = 0; k < 6; ++k) q[k] += hp[(size_t)b * 6 + k];
    double A[3][3] = {{q[0], q[1], q[2]}, {q[1], q[3], q[4]}, {q[2], q[4], q[5]}}, V[3][3], w[3];
    jacobi_eigen3(A, V, w);
    int mi = 0;   // smallest eigenvalue == last singular vector of the symmetric PSD scatter matrix (cv::SVD vt row 2)
    for (int i = 1; i < 3; ++i) if (w[i] < w[mi]) mi = i;
    double n0 = V[0][mi], n1 = V[1][mi], n2 = V[2][mi];
    const double nn = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
    n0 /= nn; n1 /= nn; n2 /= nn;
    if (n2 < 0) { n0 = -n0; n1 = -n1; n2 = -n2; }
    plane[0] = n0; plane[1] = n1; plane[2] = n2; plane[3] = -(n0 * cx + n1 * cy + n2 * cz);
    return WSG_OK;
}

void wsg_rt_from_plane(const double pl[4], double R[9], double T[3], double Rinv[9], double Tinv[3])
{
    const double a = pl[0], b = pl[1], c = pl[2], d = pl[3];
    const double q = (1 - c) / (a * a + b * b);
    R[0] = 1 - a * a * q; R[1] = -a * b * q; R[2] = -a;
    R[3] = -a * b * q; R[4] = 1 - b * b * q; R[5] = -b;
    R[6] = a; R[7] = b; R[8] = c;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rinv[3 * i + j] = R[3 * j + i];
    T[0] = 0; T[1] = 0; T[2] = d;
    for (int i = 0; i < 3; ++i) Tinv[i] = Rinv[3 * i] * (-T[0]) + Rinv[3 * i + 1] * (-T[1]) + Rinv[3 * i + 2] * (-T[2]);
}

int wsg_mesh_export_xyzc(wsg_handle* h, const double plane[4], void* dst, size_t capacity, size_t* nbytes)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!plane || !dst || !nbytes) return WSG_ERR_INVALID_ARG;
    const int n = h->mesh_w * h->mesh_h;
    double RT[12], Rinv[9], Tinv[3];
    wsg_rt_from_plane(plane, RT, RT + 9, Rinv, Tinv);
    const size_t cub_bytes = compact_cub_bytes(n);
    if ((rc = ensure(h, h->m_scratch, (size_t)n * 8 + cub_bytes + 512))) return rc;
    if ((rc = ensure(h, h->m_out, (size_t)n * 6 + 16))) return rc;
    char* base = (char*)h->m_scratch.p;
    double* d_RT = (double*)base;                 // 12 doubles
    double* d_mm = (double*)(base + 128);         // 6 doubles (min xyz, max xyz)
    double* d_ms = (double*)(base + 192);         // 3 mins + 3 scales
    unsigned* scan_tmp = (unsigned*)(base + 256);
    void* cub_tmp = base + 256 + (size_t)n * 8;
    MeshView m = mesh_view(h);
    StageTimer t(h, WSG_STAGE_MESH, 5);
    CK(h, cudaMemcpyAsync(d_RT, RT, 96, cudaMemcpyHostToDevice, h->stream));
    launch_plane_minmax(m, d_RT, nullptr, d_mm, h->stream);
    double mm[6];
    CK(h, cudaMemcpyAsync(mm, d_mm, 48, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    const double MV = 65535.0;
    double ms[6] = {mm[0], mm[1], mm[2], MV / (mm[3] - mm[0]), MV / (mm[4] - mm[1]), MV / (mm[5] - mm[2])};
    CK(h, cudaMemcpyAsync(d_ms, ms, 48, cudaMemcpyHostToDevice, h->stream));
    unsigned long long np = 0;
    if (mesh_compact_quantise(m, d_RT, nullptr, d_ms, d_ms + 3, (uint16_t*)h->m_out.p, scan_tmp, cub_tmp, cub_bytes, &np, h->stream)) {
        h->err = "compaction failed"; return WSG_ERR_CUDA;
    }
    const size_t total = 148 + (size_t)np * 6;
    *nbytes = total;
    if (capacity < total) { h->err = "destination too small"; return WSG_ERR_INVALID_ARG; }
    char* o = (char*)dst;
    const uint32_t n32 = (uint32_t)np;
    memcpy(o, &n32, 4);
    memcpy(o + 4, ms + 3, 24);      // xscale, yscale, zscale
    memcpy(o + 28, ms, 24);         // minx, miny, minz
    memcpy(o + 52, Rinv, 72);
    memcpy(o + 124, Tinv, 24);
    CK(h, cudaMemcpyAsync(o + 148, h->m_out.p, (size_t)np * 6, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    return WSG_OK;
}

int wsg_mesh_export_xyzbin(wsg_handle* h, void* dst, size_t capacity, size_t* nbytes)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!dst || !nbytes) return WSG_ERR_INVALID_ARG;
    const int n = h->mesh_w * h->mesh_h;
    const size_t cub_bytes = compact_cub_bytes(n);
    if ((rc = ensure(h, h->m_scratch, (size_t)n * 8 + cub_bytes + 512))) return rc;
    if ((rc = ensure(h, h->m_out, (size_t)n * 12 + 16))) return rc;
    char* base = (char*)h->m_scratch.p;
    unsigned long long np = 0;
    StageTimer t(h, WSG_STAGE_MESH, 3);
    if (mesh_compact_xyz(mesh_view(h), (float*)h->m_out.p, (unsigned*)(base + 256), base + 256 + (size_t)n * 8, cub_bytes, &np, h->stream)) {
        h->err = "compaction failed"; return WSG_ERR_CUDA;
    }
    const size_t total = 4 + (size_t)np * 12;
    *nbytes = total;
    if (capacity < total) { h->err = "destination too small"; return WSG_ERR_INVALID_ARG; }
    const uint32_t n32 = (uint32_t)np;
    memcpy(dst, &n32, 4);
    CK(h, cudaMemcpyAsync((char*)dst + 4, h->m_out.p, (size_t)np * 12, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    return WSG_OK;
}

void wsg_plane_mean_accumulate(double acc[5], const double plane[4])
{
    if (std::isnan(plane[0]) || std::isnan(plane[1]) || std::isnan(plane[2]) || std::isnan(plane[3])) return;
    for (int k = 0; k < 4; ++k) acc[k] += plane[k];
    acc[4] += 1.0;
}

void wsg_plane_mean_finish(const double acc[5], double mean[4])
{
    for (int k = 0; k < 4; ++k) mean[k] = acc[4] > 0 ? acc[k] / acc[4] : std::nan("");
}


// ---- consumer side: mesh_cam.xyzC -> points aligned on the sea plane ---------------------------------------------------
static int decode_align_device(wsg_handle* h, const uint16_t* d_q, size_t n, const double scale[3], const double mins[3],
                               const double Rinv[9], const double Tinv[3], const double align_plane[4], double baseline,
                               double* out_host)
{
    double P[31], RT[12], Ri[9], Ti[3];
    for (int k = 0; k < 3; ++k) { P[k] = scale[k]; P[3 + k] = mins[k]; P[15 + k] = Tinv[k]; }
    for (int k = 0; k < 9; ++k) P[6 + k] = Rinv[k];
    wsg_rt_from_plane(align_plane, RT, RT + 9, Ri, Ti);           // compute_sea_plane_RT (wass_utils.py:38-48)
    for (int k = 0; k < 9; ++k) P[18 + k] = RT[k];
    for (int k = 0; k < 3; ++k) P[27 + k] = RT[9 + k];
    P[30] = baseline;
    int rc;
    if ((rc = ensure(h, h->m_small, 4096))) return rc;
    if ((rc = ensure(h, h->m_labels, n * 3 * sizeof(double) + 16))) return rc;     // reused as the output staging buffer
    CK(h, cudaMemcpyAsync(h->m_small.p, P, sizeof(P), cudaMemcpyHostToDevice, h->stream));
    StageTimer t(h, WSG_STAGE_MESH, 1);
    launch_xyzc_decode_align(d_q, n, (const double*)h->m_small.p, (double*)h->m_labels.p, h->stream);
    CK(h, cudaMemcpyAsync(out_host, h->m_labels.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaGetLastError());
    return WSG_OK;
}

int wsg_xyzc_decode_align(wsg_handle* h, const void* xyzc, size_t nbytes, const double align_plane[4], double baseline,
                          double* out_xyz, size_t capacity_points, size_t* n_points)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!xyzc || !align_plane || !n_points || nbytes < 148) { h->err = "bad argument or truncated header"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    const char* b = (const char*)xyzc;
    uint32_t n32; memcpy(&n32, b, 4);
    double lim[6], Rinv[9], Tinv[3];
    memcpy(lim, b + 4, 48); memcpy(Rinv, b + 52, 72); memcpy(Tinv, b + 124, 24);
    const size_t n = n32;
    *n_points = n;
    if (nbytes < 148 + n * 6) { h->err = "truncated point data"; return WSG_ERR_INVALID_ARG; }
    if (!out_xyz || capacity_points < n) { h->err = "destination too small"; return WSG_ERR_INVALID_ARG; }
    if (n == 0) return WSG_OK;
    int rc;
    if ((rc = ensure(h, h->m_out, n * 6 + 16))) return rc;
    CK(h, cudaMemcpyAsync(h->m_out.p, b + 148, n * 6, cudaMemcpyHostToDevice, h->stream));
    return decode_align_device(h, (const uint16_t*)h->m_out.p, n, lim, lim + 3, Rinv, Tinv, align_plane, baseline, out_xyz);
}

int wsg_mesh_aligned_points(wsg_handle* h, const double plane[4], const double align_plane[4], double baseline,
                            double* out_xyz, size_t capacity_points, size_t* n_points)
{
    int rc = need_mesh(h);
    if (rc) return rc;
    if (!plane || !align_plane || !n_points) return WSG_ERR_INVALID_ARG;
    // quantise exactly as save_as_xyz_compressed would (the consumer sees the u16 grid, not the fp64 mesh) ...
    std::vector<char> hdr(148);
    size_t nb = 0;
    const int n = h->mesh_w * h->mesh_h;
    std::vector<char> file((size_t)148 + (size_t)n * 6);
    if ((rc = wsg_mesh_export_xyzc(h, plane, file.data(), file.size(), &nb))) return rc;
    uint32_t n32; memcpy(&n32, file.data(), 4);
    *n_points = n32;
    if (!out_xyz || capacity_points < n32) { h->err = "destination too small"; return WSG_ERR_INVALID_ARG; }
    if (n32 == 0) return WSG_OK;
    double lim[6], Rinv[9], Tinv[3];
    memcpy(lim, file.data() + 4, 48); memcpy(Rinv, file.data() + 52, 72); memcpy(Tinv, file.data() + 124, 24);
    // ... and decode + align the copy that is still on the device (h->m_out)
    return decode_align_device(h, (const uint16_t*)h->m_out.p, n32, lim, lim + 3, Rinv, Tinv, align_plane, baseline, out_xyz);
}

}  // extern "C"
