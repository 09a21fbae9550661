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

/* Statistics of the last wsg_sgbm_compute* on this handle (synchronises the stream). */
typedef struct wsg_sgbm_stats {
    int max_cost;            /* max over the cost volume C */
    int out_of_domain;       /* 1 if max_cost + P2 > 32767: outside the range in which cv2 is reproduced bit-exactly */
    int kernel_launches;     /* CUDA kernels launched by the call */
    int width1;              /* W1: matched columns */
    int d_padded;            /* disparity slots per pixel in the HBM volumes */
    long long volume_bytes;  /* bytes of one int16 volume (C or S) */
} wsg_sgbm_stats;
int wsg_sgbm_get_stats(wsg_handle* h, wsg_sgbm_stats* out);

/* Test hook: copies the cost volume C and the aggregated volume S of the last compute to host,
 * in logical layout [rows][W1][numDisparities] int16.  Either pointer may be NULL. */
int wsg_sgbm_debug_volumes(wsg_handle* h, int16_t* C_host, int16_t* S_host);

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
