/*
 * wassgpu.h -- C ABI of the B200-native wass_stereo hot path (libwassgpu.so).
 *
 * The reference (fbergama/wass) has no plugin/FFI layer: its stage boundary is the wass_stereo
 * process, and inside it the seams below are plain C++ calls.  Each entry point names the
 * reference seam it replaces (paths relative to the reference root).  A C++ host (the drop-in
 * wass_stereo executable, wass_b200/csrc/host/) and ctypes (tests, bench.py) bind these symbols;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 (WSG_OK) or a negative WSG_ERR_* code and never throws;
 * the caller owns host buffers; the library owns device buffers inside the opaque handle;
 * one handle = one CUDA device + one stream; a handle is not thread-safe, distinct handles are.
 * There is no CPU fallback: without a CUDA device wsg_create fails with WSG_ERR_CUDA.
 */
#ifndef WASSGPU_H_
#define WASSGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSG_OK 0
#define WSG_ERR_INVALID_ARG (-1)   /* null pointer, bad size, unsupported parameter value */
#define WSG_ERR_CUDA (-2)          /* CUDA runtime error; see wsg_last_error */
#define WSG_ERR_NOMEM (-3)
#define WSG_ERR_TOO_SMALL (-4)     /* image narrower than maxDisparity + blockSize/2 (cv2 raises here) */
#define WSG_ERR_STATE (-5)         /* call order violated (e.g. debug read-back before a compute) */

typedef struct wsg_handle wsg_handle;

/* ---- lifetime ------------------------------------------------------------------------------- */
int wsg_create(int device, wsg_handle** out);
void wsg_destroy(wsg_handle* h);
/* Run all subsequent work of this handle on `cuda_stream` (a cudaStream_t; NULL = handle's own). */
int wsg_set_stream(wsg_handle* h, void* cuda_stream);
/* Block until the handle's stream is idle. */
int wsg_synchronize(wsg_handle* h);
const char* wsg_last_error(const wsg_handle* h);
const char* wsg_version(void);

/* ---- dense matcher -------------------------------------------------------------------------- */
/* Field-for-field the parameter set of cv::StereoSGBM as wass_stereo fills it
 * (src/wass_stereo/wass_stereo.cpp:772-782).  mode: 0 = MODE_SGBM (5 paths, the reference's
 * default), 1 = MODE_HH (8 paths). */
typedef struct wsg_sgbm_params {
    int minDisparity;
    int numDisparities;   /* multiple of 16, <= 1280 */
    int blockSize;        /* odd, <= 25 */
    int P1;
    int P2;
    int disp12MaxDiff;
    int preFilterCap;
    int uniquenessRatio;
    int speckleWindowSize;
    int speckleRange;
    int mode;
} wsg_sgbm_params;

#define WSG_MODE_SGBM 0
#define WSG_MODE_HH 1

/* Replaces dense_stereo->compute(right_image,left_image,disparity), wass_stereo.cpp:837.
 * img1/img2: 8-bit grey, rows x cols, `stride` bytes per row, HOST memory.
 * disp16: rows x cols int16 (dense, cols per row), fixed point x16, invalid = (minDisparity-1)*16.
 * Includes host->device and device->host copies; returns after the result is in disp16. */
int wsg_sgbm_compute(wsg_handle* h, const uint8_t* img1, const uint8_t* img2, int rows, int cols,
                     size_t stride, const wsg_sgbm_params* p, int16_t* disp16);

/* Same computation on DEVICE pointers, asynchronous on the handle's stream (no copies, no sync). */
int wsg_sgbm_compute_device(wsg_handle* h, const uint8_t* d_img1, const uint8_t* d_img2, int rows, int cols,
                            size_t stride, const wsg_sgbm_params* p, int16_t* d_disp16);

/* Batch forms: `n` frames of equal geometry and parameters in ONE call -- the form a sequence driver uses (the reference
 * runs one wass_stereo process per frame, cli/wasscli/wasscli.py:326-346; nothing couples the frames).  The cost volumes of
 * all frames are built first, then ONE launch per aggregation sweep walks the row bands of all frames interleaved, so the
 * wavefront of one frame fills while another drains.  Results are identical to n single calls.
 * Device memory: n * 2 * volume_bytes (wsg_sgbm_stats).
 *   wsg_sgbm_compute_batch         HOST buffers: img1[f], img2[f] (rows x cols, `stride`), disp16[f] (rows x cols dense).
 *   wsg_sgbm_compute_batch_device  DEVICE buffers: frame f at d_img + f * frame_stride bytes, d_disp16 + f * rows * cols;
 *                                  asynchronous on the handle's stream. */
#define WSG_MAX_BATCH 64
int wsg_sgbm_compute_batch(wsg_handle* h, int n, const uint8_t* const* img1, const uint8_t* const* img2, int rows, int cols,
                           size_t stride, const wsg_sgbm_params* p, int16_t* const* disp16);
int wsg_sgbm_compute_batch_device(wsg_handle* h, int n, const uint8_t* d_img1, const uint8_t* d_img2, size_t frame_stride,
                                  int rows, int cols, size_t stride, const wsg_sgbm_params* p, int16_t* d_disp16);

/* Asynchronous batches, for a driver that overlaps its own work (PNG decode, disk) and the PCIe copies with the GPU:
 * wsg_sgbm_batch_submit enqueues the host->device copies, the matcher and the device->host copies of one batch and
 * returns at once; wsg_sgbm_batch_wait blocks until that batch's disparities are in the caller's buffers.  Two batches can
 * be in flight (slot 0 and 1): the copies of one run beside the kernels of the other, on their own streams.  The host
 * buffers must stay valid until the wait, and should be pinned (otherwise the copies are staged and synchronous). */
int wsg_sgbm_batch_submit(wsg_handle* h, int slot, int n, const uint8_t* const* img1, const uint8_t* const* img2, int rows, int cols,
                          size_t stride, const wsg_sgbm_params* p, int16_t* const* disp16);
int wsg_sgbm_batch_wait(wsg_handle* h, int slot);

/* Statistics of the last wsg_sgbm_compute* on this handle (synchronises the stream). */
typedef struct wsg_sgbm_stats {
    int max_cost;            /* max over the cost volume C */
    int out_of_domain;       /* 1 if max_cost + P2 > 32767: outside the range in which cv2 is reproduced bit-exactly */
    int kernel_launches;     /* CUDA kernels launched by the call */
    int width1;              /* W1: matched columns */
    int d_padded;            /* disparity slots per pixel in the HBM volumes */
    int agg_impl;            /* WSG_AGG_* actually used by the call */
    long long volume_bytes;  /* bytes of one int16 volume (C or S) */
} wsg_sgbm_stats;
int wsg_sgbm_get_stats(wsg_handle* h, wsg_sgbm_stats* out);

/* Test hook: copies the cost volume C and the aggregated volume S of the last compute to host,
 * in logical layout [rows][W1][numDisparities] int16.  Either pointer may be NULL. */
int wsg_sgbm_debug_volumes(wsg_handle* h, int16_t* C_host, int16_t* S_host);

/* Which device implementation of the path aggregation (A.4) + winner-take-all (A.5) the next computes use.
 * All three produce identical results; the choice only moves HBM traffic:
 *   WSG_AGG_PER_DIRECTION  one launch per path direction (8 or 5), separate WTA kernel          (23V moved)
 *   WSG_AGG_SWEEPS         fused 4-direction wavefront sweeps, S written out, separate WTA      ( 6V moved)
 *   WSG_AGG_SWEEPS_WTA     fused sweeps, WTA inside the last sweep, S never written             ( 4V moved)
 * WSG_AGG_SWEEPS_WTA is the default; the other two are the cross-checks the parity tests compare it with at full size.
 * The fused forms need numDisparities <= 512; above that the per-direction form is used regardless. */
#define WSG_AGG_PER_DIRECTION 0
#define WSG_AGG_SWEEPS 1
#define WSG_AGG_SWEEPS_WTA 2
int wsg_sgbm_set_impl(wsg_handle* h, int impl);
/* Caps the number of SMs the fused sweeps of this handle occupy (0 = all, the default): a sweep is a persistent launch of
 * one worker per SM; a cap leaves the other SMs to kernels of other streams. */
int wsg_sgbm_set_sweep_workers(wsg_handle* h, int max_sms);

/* ---- dense stereo stage as a whole ------------------------------------------------------------ */
/* Replaces sgbm_dense_stereo(env), wass_stereo.cpp:764-1020, between the two rectified crops and the
 * float disparity of the ROI: P1/P2 from WINSIZE, zero padding by numDisparities(+offset) columns,
 * compute(right,left), crop, clean_and_convert_disparity, DISP_DILATE_STEPS x matrix_dilate_zero,
 * DISP_EROSION_STEPS x matrix_erode_zero, nearest-neighbour mask (one more erosion).
 * Field names are the reference's configuration keys (wass_stereo.cpp:742-761). */
typedef struct wsg_dense_params {
    int MIN_DISPARITY;
    int MAX_DISPARITY;            /* = numberOfDisparities (wass_stereo.cpp:768) */
    int WINSIZE;
    double DENSE_SCALE;           /* != 1: cv::resize INTER_CUBIC before the matcher, NEAREST + CUBIC after (OpenCV's own arithmetic, IPP-free) */
    int DISPARITY_OFFSET;
    int DISP_DILATE_STEPS;
    int DISP_EROSION_STEPS;
    int DENSE_P1_MULT;
    int DENSE_P2_MULT;
    int DENSE_UNIQUENESS_RATIO;
    int DENSE_DISP12MAXDIFF;
    int DENSE_PREFILTER_CAP;
    int DENSE_SPECKLE_RANGE;
    int DENSE_SPECKLE_WINDOW_SIZE;
    int mode;                     /* WSG_MODE_SGBM (reference default) or WSG_MODE_HH */
    int MEDIAN_FILTER_WSIZE;      /* 0 = off (default); 3 or 5: cv::medianBlur on the float disparity (wass_stereo.cpp:941-945) */
    int DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD;   /* 0 = off (default); > 0: wass_stereo.cpp:947-986 */
} wsg_dense_params;
void wsg_dense_params_default(wsg_dense_params* p);

/* left_crop/right_crop: rows x cols uint8 HOST images (the rectified ROI crops, wass_stereo.cpp:607-608).
 * disp_roi: rows x cols float32 HOST, 0 = invalid; may be NULL when the caller goes on with wsg_triangulate_from_dense (the
 * disparity stays on the device).  disp16_roi (optional, may be NULL): the raw
 * matcher output cropped to the ROI, int16 x16. */
int wsg_dense_stereo(wsg_handle* h, const uint8_t* left_crop, const uint8_t* right_crop, int rows, int cols,
                     size_t stride, const wsg_dense_params* p, float* disp_roi, int16_t* disp16_roi);
/* Batch form (the form a sequence driver uses): n pairs of crops of one size; ONE batched matcher run underneath
 * (wsg_sgbm_compute_batch).  The float ROI disparity of every frame stays on the device; disp_roi may be NULL or hold
 * NULLs.  wsg_dense_select(h, f) makes frame f the current dense result, i.e. what wsg_triangulate_from_dense reads. */
int wsg_dense_stereo_batch(wsg_handle* h, int n, const uint8_t* const* left_crops, const uint8_t* const* right_crops, int rows,
                           int cols, size_t stride, const wsg_dense_params* p, float* const* disp_roi);
int wsg_dense_select(wsg_handle* h, int frame);


/* Size of the matcher's input for a rows x cols crop (wass_stereo.cpp:788-797): x is scaled when DENSE_SCALE > 1, both
 * axes when < 1; lengths are cv::resize's saturate_cast<int>(n * scale).  With DENSE_SCALE != 1 wsg_dense_stereo's
 * disp16_roi has this size; disp_roi always has the crop's. */
void wsg_dense_scaled_size(int rows, int cols, double dense_scale, int* rows_s, int* cols_s);

/* Replaces wass_stereo.cpp:853-928 alone: int16 x16 ROI disparity -> cleaned float32 disparity. HOST pointers.
 * denseScale must be 1 here (the output has the input's size). */
int wsg_disparity_postprocess(wsg_handle* h, const int16_t* disp16_roi, int rows, int cols, int minDisparity,
                              int numDisparities, int disparityOffset, double denseScale, int dilateSteps,
                              int erosionSteps, float* disp_roi);
/* The same with DENSE_SCALE != 1: disp16_roi is rows x cols (the resized matcher input's size), disp_roi is
 * out_rows x out_cols (roi_comb_right's size, wass_stereo.cpp:903-904): values are multiplied by 1/denseScale, filtered,
 * enlarged with INTER_NEAREST and INTER_CUBIC, and the bicubic map is zeroed where the eroded nearest map is zero. */
int wsg_disparity_postprocess_resized(wsg_handle* h, const int16_t* disp16_roi, int rows, int cols, int minDisparity,
                                      int numDisparities, int disparityOffset, double denseScale, int dilateSteps,
                                      int erosionSteps, float* disp_roi, int out_rows, int out_cols);

/* cv::resize as the path uses it (OpenCV's own resize code; the Intel IPP path some OpenCV builds substitute differs by
 * +-1 on a few % of 8-bit pixels -- DESIGN.md section 2).  HOST pointers.
 * wsg_resize_u8_cubic == cv::resize(src, dst, Size(), fx, fy, INTER_CUBIC) on CV_8UC1 (wass_stereo.cpp:790-795);
 *   dst must be round(rows*fy) x round(cols*fx).
 * wsg_resize_f32 == cv::resize(src, dst, Size(dst_cols,dst_rows), 0, 0, interpolation) on CV_32FC1 (:903-904). */
#define WSG_INTER_NEAREST 0
#define WSG_INTER_CUBIC 2
int wsg_resize_u8_cubic(wsg_handle* h, const uint8_t* src, int rows, int cols, size_t stride, double fx, double fy,
                        uint8_t* dst, int dst_rows, int dst_cols);
int wsg_resize_f32(wsg_handle* h, const float* src, int rows, int cols, float* dst, int dst_rows, int dst_cols,
                   int interpolation);

/* Replaces wass_stereo.cpp:941-986 alone (both steps off at the reference defaults): optional cv::medianBlur (3 or 5) of
 * the float ROI disparity, then -- if bc_threshold > 0 -- zero where the squared Sobel gradient magnitude exceeds it and
 * keep only the biggest 8-connected component of the non-zero pixels.  disp_roi: HOST, rows x cols float32, in place. */
int wsg_disparity_refine(wsg_handle* h, float* disp_roi, int rows, int cols, int median_wsize, int bc_threshold);

/* ---- triangulation + PovMesh ---------------------------------------------------------------------- */
/* Calibration after load_data()+rectify() (wass_stereo.cpp:337-613), all row-major doubles.
 * K0/K1: intrinsics of the LEFT/RIGHT camera after any left-right swap; R,T: pose of right w.r.t. left
 * (|T| already rescaled to cam_distance); R1,R2,P1,P2: outputs of cv::stereoRectify. */
typedef struct wsg_calib {
    double K0[9], K1[9], R[9], T[3];
    double R1[9], R2[9], P1[12], P2[12];
    int roi_left[4], roi_right[4];        /* x,y,width,height = roi_comb_left / roi_comb_right */
    int left_cols, left_rows, right_cols, right_rows;   /* original (undistorted) images */
    int rect_cols, rect_rows;             /* rectified images */
    /* USE_CUSTOM_STEREORECTIFY (wass_stereo.cpp:301-305): when use_homographies != 0, rectified pixels go back to the
     * original images through HLi / HRi (the inverses of wsg_stereo_rectify_custom's H0 / H1 for the LEFT / RIGHT image)
     * and R1,R2,P1,P2 are not read. */
    int use_homographies;
    double HLi[9], HRi[9];
} wsg_calib;

typedef struct wsg_tri_params {
    double TRIANG_MIN_ANGLE;              /* degrees; <= 0 disables */
    double TRIANG_BBOX_TOP, TRIANG_BBOX_LEFT, TRIANG_BBOX_RIGHT, TRIANG_BBOX_BOTTOM;  /* all >= 0 to enable */
    int DISCARD_BURNED_AREAS;
    int disparity_compensation;           /* max(-DISPARITY_OFFSET,0) */
    double DENSE_SCALE;
    double cam_distance;                  /* 1.0 in wass_stereo (wass_stereo.cpp:1878) */
} wsg_tri_params;
void wsg_tri_params_default(wsg_tri_params* p);

/* Replaces triangulate(env), wass_stereo.cpp:1039-1386.  disparity: rect_rows x rect_cols float32 (HOST);
 * left/right: the original 8-bit images (HOST); left_mask/right_mask: optional 0/1 masks of the same
 * sizes (NULL = none).  Creates the handle's device-resident mesh (roi_right.width x roi_right.height
 * point grid, the PovMesh of the reference) and returns the number of triangulated points. */
int wsg_triangulate(wsg_handle* h, const float* disparity, const uint8_t* left, const uint8_t* right,
                    const uint8_t* left_mask, const uint8_t* right_mask, const wsg_calib* calib,
                    const wsg_tri_params* p, unsigned long long* n_points);

/* Same, but takes the ROI disparity still on the device from the last wsg_dense_stereo call
 * (paste into the full rectified frame, wass_stereo.cpp:990-991, happens on the device). */
int wsg_triangulate_from_dense(wsg_handle* h, const uint8_t* left, const uint8_t* right, const uint8_t* left_mask,
                               const uint8_t* right_mask, const wsg_calib* calib, const wsg_tri_params* p,
                               unsigned long long* n_points);

/* Test / tooling access to the mesh (PovMesh grid): upload replaces the handle's mesh. xyz: h*w*3 doubles. */
int wsg_mesh_upload(wsg_handle* h, int width, int height, const uint8_t* valid, const double* xyz, const uint8_t* grey);
int wsg_mesh_download(wsg_handle* h, uint8_t* valid, double* xyz, uint8_t* grey);
int wsg_mesh_size(wsg_handle* h, int* width, int* height, unsigned long long* n_valid);

/* PovMesh::compute_zgap_percentile, PovMesh.cpp:888-926 */
int wsg_mesh_zgap_percentile(wsg_handle* h, double percentile, double* zgap);
/* PovMesh::cluster_biggest_connected_component, PovMesh.cpp:929-987 */
int wsg_mesh_biggest_component(wsg_handle* h, double zgap, unsigned long long* n_left);
/* PovMesh::ransac_find_plane, PovMesh.cpp:665-777.  triples: n x 6 int32 pixel coordinates
 * (u1,v1,u2,v2,u3,v3) drawn by the caller with libc rand() (see wsg_ransac_draw); *ok = 0 when the
 * best hypothesis has fewer than width*height/10 inliers. */
int wsg_mesh_ransac_plane(wsg_handle* h, const int32_t* triples, int n, double threshold, double plane[4], int* ok,
                          unsigned long long* best_inliers);
/* The draw loop of PovMesh.cpp:678-692: consumes libc rand(); rounds whose points are closer than 0.01*height are
 * redrawn.  Per point v is drawn before u, points 1,2,3 in order: the order a GCC build of the reference produces
 * (constructor arguments, right to left) -- pinned against the reference's own code with a fixed RANDOM_SEED. Host only. */
int wsg_ransac_draw(int width, int height, int rounds, int32_t* triples);
/* PovMesh::crop_plane, PovMesh.cpp:780-815 */
int wsg_mesh_crop_plane(wsg_handle* h, const double plane[4], double threshold, unsigned long long* n_left);
/* PovMesh::refine_plane, PovMesh.cpp:581-660 */
typedef struct wsg_refine_params {
    double PLANE_REFINE_XMIN, PLANE_REFINE_XMAX, PLANE_REFINE_YMIN, PLANE_REFINE_YMAX;
    double PLANE_REFINEMENT_MAX_DISTANCE;
    int PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE;
    int PLANE_USE_CENTRAL_THIRD_ONLY;
} wsg_refine_params;
void wsg_refine_params_default(wsg_refine_params* p);
int wsg_mesh_refine_plane(wsg_handle* h, const wsg_refine_params* p, double plane[4], unsigned long long* n_inliers);
/* The points main() dumps to plane_refinement_inliers.xyz before the refinement (wass_stereo.cpp:2077-2085): every
 * `every`-th (10 there) point that passes refine_plane's inlier test (PovMesh.cpp:596-618), in grid scan order, x y z as
 * doubles.  Selected on the device (flags + scan + scatter): ~1 MB comes back instead of the whole 120 MB mesh.
 * n_points = points written; n_inliers (optional) = all inliers. */
int wsg_mesh_refine_inliers(wsg_handle* h, const wsg_refine_params* p, int every, double* xyz, size_t capacity_points,
                            unsigned long long* n_points, unsigned long long* n_inliers);
/* PovMesh::RT_from_plane, PovMesh.cpp:1044-1074 (host arithmetic) */
void wsg_rt_from_plane(const double plane[4], double R[9], double T[3], double Rinv[9], double Tinv[3]);
/* PovMesh::save_as_xyz_compressed, PovMesh.cpp:377-460: the exact bytes of mesh_cam.xyzC into dst. */
int wsg_mesh_export_xyzc(wsg_handle* h, const double plane[4], void* dst, size_t capacity, size_t* nbytes);
/* PovMesh::save_as_xyz_binary, PovMesh.cpp:346-375: the exact bytes of mesh_cam.xyzbin. */
int wsg_mesh_export_xyzbin(wsg_handle* h, void* dst, size_t capacity, size_t* nbytes);

/* Replaces cv::undistort(img, out, K, dist) of the stage two steps up (src/wass_prepare/wass_prepare.cpp:268; SURVEY section 8f
 * rank 4): 8-bit grey HOST image in and out, dist = (k1 k2 p1 p2 [k3 [k4 k5 k6]]), bilinear in OpenCV's 1/32-pixel fixed
 * point, constant (0) border.  The polarimetric demosaic of wass_prepare is not built. */
int wsg_undistort_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, const double K[9], const double* dist,
                        int ndist, uint8_t* out);
/* Replaces stereoRectifyUndistorted (src/wass_stereo/stereorectify.cpp:57-244; USE_CUSTOM_STEREORECTIFY, SURVEY section 8f
 * rank 3), host arithmetic: rectifying homographies H0 (image of camera K0) and H1 (K1) for the pose R,T taking points of
 * camera 1 into camera 0 (wass_stereo.cpp:502 passes Rinv, Tinv), and the common ROI (x,y,width,height).  rot_angle != 0
 * (degrees, RECTIFY_ANGLE) is used as given; 0 minimises max(v0, v1), the perspective components of the two homographies,
 * over the rotation about the baseline with a Nelder-Mead simplex (the reference uses cv::DownhillSolver with the same
 * start, initial step and tolerance; that solver is not available to pin the iterates, so the optimum -- not the path --
 * is what is reproduced: parity unpinned).  best_angle (optional) receives the angle used. */
int wsg_stereo_rectify_custom(const double K0[9], const double K1[9], const double R[9], const double T[3], double rot_angle,
                              int width, int height, double H0[9], double H1[9], int roi[4], double* best_angle);
/* Replaces cv::warpPerspective(img, out, H, img.size()) (wass_stereo.cpp:515-516: INTER_LINEAR, constant 0 border): OpenCV's
 * blocked double-precision coordinates, 1/32-pixel fixed-point bilinear weights -- bit-exact vs cv2.  HOST pointers. */
int wsg_warp_perspective(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, const double H[9], uint8_t* out);
/* Replaces clahe->apply(img, dst) of wass_prepare (src/wass_prepare/wass_prepare.cpp:257-262; cv::createCLAHE(clip_limit,
 * Size(tiles, tiles)) at :458-462 and :479-483, keys CAMx_CLAHE_CLIPLIMIT / CAMx_CLAHE_TILEGRIDSIZE): 8-bit grey HOST image
 * in and out, bit-exact vs cv2.createCLAHE. */
int wsg_clahe_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, double clip_limit, int tiles, uint8_t* out);
/* process_image() of wass_prepare without the polarimetric branch (wass_prepare.cpp:88-275): optional CLAHE (clahe_tiles > 0)
 * then cv::undistort (K != NULL), chained on the device with one upload and one download. */
int wsg_prepare_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, int clahe_tiles, double clahe_clip,
                      const double K[9], const double* dist, int ndist, uint8_t* out);

/* ---- consumer side of the mesh (the step right after the hot path) -------------------------------- */
/* Replaces load_camera_mesh + align_on_sea_plane (gridding/wassgridsurface/wass_utils.py:22-35 and 38-68, called at
 * gridding/wassgridsurface/wassgridsurface.py:86-87 and 316-318): decodes the bytes of a mesh_cam.xyzC file (HOST
 * pointer) on the device -- u16 / scale + min, Rinv @ p + Tinv -- then rotates/translates onto `align_plane` (normally
 * the mean plane of the sequence, wassgridsurface.py:677-678), flips z and multiplies by `baseline`.
 * out_xyz: HOST, 3 rows of *n_points doubles (the reference's 3xN array, row-major); capacity_points >= the file's n. */
int wsg_xyzc_decode_align(wsg_handle* h, const void* xyzc, size_t nbytes, const double align_plane[4], double baseline,
                          double* out_xyz, size_t capacity_points, size_t* n_points);
/* Same result without the file round trip: quantises the device-resident mesh exactly as wsg_mesh_export_xyzc(plane)
 * would and decodes/aligns that copy on the device. */
int wsg_mesh_aligned_points(wsg_handle* h, const double plane[4], const double align_plane[4], double baseline,
                            double* out_xyz, size_t capacity_points, size_t* n_points);

/* ---- rectification (the step before the hot path) ---------------------- */
/* cv::stereoRectify(K0, 0, K1, 0, size, R, T, R1, R2, P1, P2, Q, flags=0, alpha=1.0, size, &roi1, &roi2)
 * exactly as wass_stereo.cpp:541 calls it (zero distortion).  Host arithmetic only. */
int wsg_stereo_rectify(const double K0[9], const double K1[9], const double R[9], const double T[3], int width, int height,
                       double R1[9], double R2[9], double P1[12], double P2[12], int roi1[4], int roi2[4]);
/* cv::initUndistortRectifyMap(K, 0, Rrect, P, size, CV_32FC1) + cv::remap(img, INTER_CUBIC), wass_stereo.cpp:600-604.
 * img/out: rows x cols uint8 HOST buffers (out is dense). */
int wsg_rectify_image(wsg_handle* h, const uint8_t* img, int rows, int cols, size_t stride, const double K[9],
                      const double Rrect[9], const double P[12], uint8_t* out);

/* NaN-aware accumulation of per-frame planes for the sequence mean (what wassgridsurface does with
 * planes.txt: np.nanmean, gridding/wassgridsurface/wassgridsurface.py:672-678).  acc[5] = sums of a,b,c,d
 * and the count; all-reduce acc (sum) across ranks, then wsg_plane_mean_finish. Host only. */
void wsg_plane_mean_accumulate(double acc[5], const double plane[4]);
void wsg_plane_mean_finish(const double acc[5], double mean[4]);

/* The same reduction across the ranks of a multi-GPU run (one process per GPU, frames sharded over ranks): ONE
 * ncclAllReduce(sum) over the 5 doubles of `acc` on the handle's stream, over NVLink / NVSwitch; mean[4] is the
 * sequence mean plane on every rank (NaN when no frame had a plane), *frames the number of planes in it.
 * nccl_comm is an ncclComm_t the caller created (its own ncclCommInitRank, or wsg_nccl_comm_create).  NCCL is bound at
 * run time (dlopen "libnccl.so.2"): libwassgpu.so does not link it.
 * wsg_plane_allgather: every rank's n_local per-frame planes in rank order (all: nranks x n_local x 4), for an
 * ORDERED planes.txt (the reference appends in completion order, cli/wasscli/wasscli.py:343). */
#define WSG_NCCL_UNIQUE_ID_BYTES 128
int wsg_nccl_unique_id(unsigned char id[WSG_NCCL_UNIQUE_ID_BYTES]);           /* rank 0; hand the bytes to the other ranks */
int wsg_nccl_comm_create(int device, int nranks, int rank, const unsigned char id[WSG_NCCL_UNIQUE_ID_BYTES], void** comm);
void wsg_nccl_comm_destroy(void* comm);
int wsg_plane_allreduce(wsg_handle* h, void* nccl_comm, const double acc[5], double mean[4], long long* frames);
int wsg_plane_allgather(wsg_handle* h, void* nccl_comm, int nranks, const double* planes, int n_local, double* all);
const char* wsg_collective_last_error(void);                                  /* for the handle-less calls above */

/* Per-stage device timing (CUDA events on the handle's stream).  enable!=0 turns recording on.
 * Stage ids: see WSG_STAGE_*.  ms[i] receives the accumulated milliseconds of stage i and
 * launches[i] the number of kernel launches since the last wsg_profile_reset. */
#define WSG_STAGE_PREFILTER 0
#define WSG_STAGE_COST 1
#define WSG_STAGE_AGGREGATE 2
#define WSG_STAGE_WTA 3
#define WSG_STAGE_MEDIAN 4
#define WSG_STAGE_POSTFILTER 5
#define WSG_STAGE_TRIANGULATE 6
#define WSG_STAGE_MESH 7
#define WSG_NUM_STAGES 8
int wsg_profile_enable(wsg_handle* h, int enable);
int wsg_profile_reset(wsg_handle* h);
int wsg_profile_get(wsg_handle* h, float* ms, int* launches, int n);

#ifdef __cplusplus
}
#endif
#endif /* WASSGPU_H_ */
