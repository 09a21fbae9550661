// Internal declarations shared by the SGBM kernels and the C ABI (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wsg {

// Derived matcher parameters + HBM layout of the C / S volumes.
struct SgbmPlan {
    int H, W;                 // image rows / cols (cols = padded width Wp in wass_stereo)
    int minD, D, maxD;
    int SW2, SH2;
    int ftzero;
    int P1, P2;
    int uniq, d12;
    int minX1, maxX1, W1;     // matched column range in image space, W1 = maxX1-minX1
    int INVALID;              // (minD-1)*16
    int mode;                 // 0 = 5 paths, 1 = 8 paths
    int speckleWindow, speckleMaxDiff;   // cv::filterSpeckles after the median when speckleWindow > 0 (maxDiff = 16*speckleRange)
    // volume layout: int16 [H][W1][Dp]; a pixel's Dp slots are NL*K vectors of 8 disparities (interleaved: vec_pos).
    // memory vector slot s = k*NL + l holds logical vector j = l*K + k (disparities 8j..8j+7), so
    // that lane l of a pixel group owns K*8 consecutive disparities and each of its K loads is
    // one fully coalesced 16-byte access across the group.
    int NL, K, Dp;
};

__host__ __device__ inline int vec_slot(int j, int NL, int K) { return (j % K) * NL + (j / K); }
// Inside a 16-byte vector the 8 disparities are interleaved so that register i (i = 0..3) holds (a+i, a+4+i) in its
// (low, high) half: the d-1 / d+1 neighbours of three of the four registers are then simply the adjacent registers, and
// only two byte-permutes per vector remain in the path recurrence (instead of five with consecutive pairs).
// vec_pos(i): int16 position of disparity a+i inside its vector.
__host__ __device__ inline int vec_pos(int i) { return i < 4 ? 2 * i : 2 * (i - 4) + 1; }

// kernels (sgbm_kernels.cu)
void launch_prefilter(const uint8_t* img, size_t stride, uint2* pre, const SgbmPlan& p, cudaStream_t st);
void launch_cost(const uint2* pre1, const uint2* pre2, int16_t* C, int* maxC, const SgbmPlan& p, cudaStream_t st, int* launches);
// dir: 0..7 = predecessor offsets (-1,0) (-1,-1) (0,-1) (1,-1) (1,0) (-1,1) (0,1) (1,1)
void launch_aggregate_dir(const int16_t* C, int16_t* S, int dir, bool first, const SgbmPlan& p, cudaStream_t st);
void launch_wta(const int16_t* S, int16_t* raw, const SgbmPlan& p, cudaStream_t st);
void launch_median3(const int16_t* src, int16_t* dst, int rows, int cols, cudaStream_t st);
// cv::filterSpeckles on the final x16 disparity (in place); labels / counts: rows*cols ints each
void launch_filter_speckles(int16_t* disp, int rows, int cols, int invalid, int maxSpeckleSize, int maxDiff, int* labels,
                            unsigned* counts, cudaStream_t st);
int cost_smem_bytes(const SgbmPlan& p);
// wide-tile cost kernel (cost_kernels.cu): windows up to 17; launch_cost picks it unless WSG_COST_IMPL=0
bool cost_wide_supported(const SgbmPlan& p);
void launch_cost_wide(const uint2* pre1, const uint2* pre2, int16_t* C, int* maxC, const SgbmPlan& p, cudaStream_t st);

// fused wavefront sweeps (sweep_kernels.cu)
struct SweepScratch {
    void* boundary;             // band-to-band state hand-off: nframes * sweep_boundary_bytes()
    int* ticket;                // zeroed hand-out counter of THIS launch
    int max_workers = 0;        // cap on the SMs a sweep occupies (0 = all)
    int num_sms;
    const int* maxC;            // device: max over the cost volume of frame f at maxC[f * maxC_stride]
    int maxC_stride = 1;
    int* err;                   // raised if a bounded wait overran
    int* dbg;                   // optional [3 * tickets]: {SM id, start ns, end ns} per band (WSG_SWEEP_DEBUG=1), else null
    int epoch;                  // 1..3, changes with every 4-direction sweep that uses `boundary`
    int rows = 0;               // rows per band of this batch (sweep_rows_per_band)
    int nframes = 1;            // frames of the batch: C / S volumes `volume_stride_bytes` apart, keys / d1 H*W apart
    size_t volume_stride_bytes = 0;
    unsigned long long* keys;   // [nframes][H][W] right-view map as packed keys (fused WTA)
    int16_t* d1;                // [nframes][H][W] left-view disparity before the LR check (fused WTA)
};
bool sweep_supported(const SgbmPlan& p);
int sweep_rows_per_band(const SgbmPlan& p, int nframes, int workers);
size_t sweep_boundary_bytes(const SgbmPlan& p, int rows);      // per frame
size_t sweep_volume_pad_bytes();                     // the C / S volumes must be readable this far beyond either end
// One launch over the bands of all `sc.nframes` frames.  flip 0: directions r0..r3 (top->bottom); flip 1: r4..r7
// (bottom->top).  mode 0: S = sum L (write only); 1: S += sum L;  2: S += sum L, then winner-take-all into
// sc.keys / sc.d1 (S is not written).  ndir 4: all four directions of the sweep;  1: the horizontal one only (fifth
// path of MODE_SGBM).
void launch_sweep(const int16_t* C, int16_t* S, int flip, int mode, int ndir, const SgbmPlan& p, const SweepScratch& sc,
                  cudaStream_t st);
void launch_wta_reset(const SweepScratch& sc, const SgbmPlan& p, cudaStream_t st);
void launch_lrcheck(const SweepScratch& sc, int frame, int16_t* raw, const SgbmPlan& p, cudaStream_t st);

}  // namespace wsg
