// Internal declarations for the stages after the matcher: disparity clean-up, triangulation, PovMesh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace wsg {

// Device-resident SoA replacement of PovMesh's 40-byte AoS grid (src/wass_stereo/PovMesh.h:33-51).
struct MeshView {
    int w, h;
    uint8_t* valid;   // [h][w]
    double* X;        // [h][w]
    double* Y;
    double* Z;
    uint8_t* color;   // [h][w] grey (the reference stores r=g=b)
};

struct CalibDev {
    double K0[9], K1[9], R[9], T[3], R1[9], R2[9];
    double P1fx, P1fy, P1cx, P1cy, P2fx, P2fy, P2cx, P2cy;   // newintr entries used by unrectify
    int rlx, rly, rlw, rlh, rrx, rry, rrw, rrh;
    int left_cols, left_rows, right_cols, right_rows, rect_cols, rect_rows;
    double min_angle, bbox_l, bbox_t, bbox_r, bbox_b;
    int discard_burned, has_lmask, has_rmask;
    double comp_over_scale;   // disparity_compensation / DENSE_SCALE
    double cam_distance;
    int use_h;                // USE_CUSTOM_STEREORECTIFY: unrectify through HLi / HRi
    double HLi[9], HRi[9];
};

// disparity clean-up
void launch_pad_images(const uint8_t* left, const uint8_t* right, size_t stride, int rows, int cols, int ndisp,
                       int off, int comp, uint8_t* img1, uint8_t* img2, int wp, cudaStream_t st);
void launch_clean_convert(const int16_t* disp16, int rows, int cols_full, int x0, int width, int mindisp, int ndisp,
                          int disp_offset, double scale, float* out, cudaStream_t st);
void launch_dilate_zero(const float* src, float* dst, int rows, int cols, cudaStream_t st);
void launch_erode_zero(const float* src, float* dst, int rows, int cols, cudaStream_t st);
// resize_kernels.cu: cv::resize of the DENSE_SCALE != 1 path (wass_stereo.cpp:788-797, 903-904)
size_t resize_tab_bytes(int dw, int dh);
void launch_resize_cubic_u8(const uint8_t* src, size_t sstride, int sw, int sh, double fx, double fy, uint8_t* dst, size_t dstride,
                            int dw, int dh, void* tab, cudaStream_t st);
void launch_resize_cubic_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, void* tab, cudaStream_t st);
void launch_resize_nn_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, cudaStream_t st);
void launch_mask_where_zero(const float* cub, const float* nn_eroded, size_t n, float* out, cudaStream_t st);
void launch_mask_by_eroded(const float* src, float* dst, int rows, int cols, cudaStream_t st);
void launch_paste_roi(const float* roi, int rh, int rw, float* full, int rows, int cols, int x0, int y0, cudaStream_t st);

// triangulation
void launch_triangulate(const float* disparity, const uint8_t* left, const uint8_t* right, const uint8_t* lmask,
                        const uint8_t* rmask, const CalibDev& c, MeshView m, unsigned long long* counter, cudaStream_t st);

// PovMesh
size_t zgap_scratch_bytes(int w, int h);
int mesh_zgap_percentile(const MeshView& m, double percentile, void* scratch, size_t scratch_bytes, double* out_host, cudaStream_t st);
int mesh_biggest_component(MeshView m, double zgap, int* labels, unsigned long long* scratch, unsigned long long* n_left_host, cudaStream_t st);
void launch_ransac_planes(const MeshView& m, const int* triples, int n, double* planes, int* ok, cudaStream_t st);
void launch_ransac_count(const MeshView& m, const double* planes, const int* ok, int n, double thr, unsigned long long* counts, cudaStream_t st);
void launch_crop_plane(MeshView m, double a, double b, double c, double d, double thr, unsigned long long* counter, cudaStream_t st);
struct RefineArgs { double xmin, xmax, ymin, ymax, maxdist; int weight_by_distance, umin, umax, vmin, vmax; };
int refine_blocks(const MeshView& m);
void launch_refine_pass1(const MeshView& m, const RefineArgs& a, double* partial /*[blocks][5]*/, cudaStream_t st);
void launch_refine_pass2(const MeshView& m, const RefineArgs& a, double cx, double cy, double cz, double* partial /*[blocks][6]*/, cudaStream_t st);
void launch_plane_minmax(const MeshView& m, const double* R9, const double* T3, double* minmax6, cudaStream_t st);
int mesh_compact_quantise(const MeshView& m, const double* R9, const double* T3, const double* min3, const double* scale3,
                          uint16_t* out, unsigned* scan_tmp, void* cub_tmp, size_t cub_bytes, unsigned long long* n_host, cudaStream_t st);
size_t compact_cub_bytes(int n);
int mesh_refine_sample(const MeshView& m, const RefineArgs& a, unsigned every, double* out /*ceil(n/every) x 3*/, unsigned* scan_tmp, void* cub_tmp,
                       size_t cub_bytes, unsigned long long* n_inliers, cudaStream_t st);
int mesh_compact_xyz(const MeshView& m, float* out /*n x 3*/, unsigned* scan_tmp, void* cub_tmp, size_t cub_bytes,
                     unsigned long long* n_host, cudaStream_t st);
void launch_count_valid(const MeshView& m, unsigned long long* counter, cudaStream_t st);

// optional refinement of the ROI disparity (wass_stereo.cpp:941-986)
void launch_median_f32(const float* src, float* dst, int rows, int cols, int ksize, cudaStream_t st);
void launch_gradient_mask(const float* src, float* dst, int rows, int cols, float thr, cudaStream_t st);
void launch_keep_biggest_cc8(float* d, int rows, int cols, int* labels, unsigned* cnt, unsigned* key, unsigned long long* best,
                             cudaStream_t st);

// consumer side of mesh_cam.xyzC (gridding/wassgridsurface/wass_utils.py:22-68): u16 -> plane frame -> camera frame ->
// aligned on the (mean) sea plane, z flipped, scaled by the baseline.  q: n x (x,y,z) u16 on the device; M: 24 doubles
// {inv scale[3], min[3], Rinv[9], Tinv[3], R[9] of the mean plane ... see capi}; out: 3 x n doubles (row-major 3 rows).
void launch_xyzc_decode_align(const uint16_t* q, size_t n, const double* d_params, double* out, cudaStream_t st);

}  // namespace wsg