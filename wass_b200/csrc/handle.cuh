// Internal: the handle object shared by the C ABI translation units.
#pragma once
#include "../../include/wassgpu.h"
#include "sgbm.cuh"
#include "geom.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace wsg;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct wsg_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    // SGBM arena
    DevBuf pre1, pre2, C, S, raw, img1, img2, disp, scalars;
    DevBuf bnd, keys, d1, dbg;               // fused sweeps: band hand-off buffer, right-view keys, left-view map
    int agg_impl = WSG_AGG_SWEEPS_WTA;
    int num_sms = 0;
    int sweep_workers = 0;              // cap on the SMs a sweep occupies (0 = all)
    int sweep_epoch = 0;                // 1..3 after the first sweep
    int bnd_H = 0, bnd_W1 = 0, bnd_K = 0, bnd_n = 0;    // geometry / batch size the hand-off buffer was last used with
    int bnd_rows = 0, sweep_rows = 0;   // rows per band the hand-off buffer was last used with / of the last batch
    int batch_n = 1;                    // frames of the last dense-matcher call
    // asynchronous batches (wsg_sgbm_batch_submit / _wait): two slots, copies on their own streams
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    DevBuf a_img1[2], a_img2[2], a_disp[2];
    bool async_busy[2] = {false, false};
    int* async_flag = nullptr;          // pinned: the sweep error flag of each slot's batch
    SgbmPlan plan{};
    bool have_plan = false;
    wsg_sgbm_stats stats{};
    // dense stage / geometry arena
    DevBuf crop_l, crop_r, rs_l, rs_r, rs_tab, fa, fb, fc, fbatch, dispfull, im_left, im_right, mask_l, mask_r;
    int dense_batch = 1, dense_batch_rows = 0, dense_batch_cols = 0;     // frames / ROI size of the last dense batch (fbatch)
    int dense_rows = 0, dense_cols = 0;     // size of the ROI disparity held in `fa` after wsg_dense_stereo
    bool have_dense = false;
    DevBuf m_valid, m_X, m_Y, m_Z, m_color, m_labels, m_scratch, m_small, m_out;
    int mesh_w = 0, mesh_h = 0;
    bool have_mesh = false;
    // profiling
    bool prof = false;
    float stage_ms[WSG_NUM_STAGES] = {0};
    int stage_launches[WSG_NUM_STAGES] = {0};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::vector<cudaEvent_t> ev_pool;
};

#define CK(h, call)                                                                           \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
            return WSG_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

inline int ensure(wsg_handle* h, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return WSG_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? WSG_ERR_NOMEM : WSG_ERR_CUDA;
    }
    b.cap = bytes;
    return WSG_OK;
}

struct StageTimer {
    wsg_handle* h; int stage; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(wsg_handle* h)
    {
        if (!h->ev_pool.empty()) { cudaEvent_t e = h->ev_pool.back(); h->ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    StageTimer(wsg_handle* h_, int s, int launches) : h(h_), stage(s)
    {
        h->stage_launches[s] += launches;
        if (h->prof) { a = get(h); b = get(h); cudaEventRecord(a, h->stream); }
    }
    ~StageTimer()
    {
        if (h->prof) { cudaEventRecord(b, h->stream); h->pending.push_back({stage, {a, b}}); }
    }
};

inline void drain_profile(wsg_handle* h)
{
    for (auto& pe : h->pending) {
        cudaEventSynchronize(pe.second.second);
        float ms = 0;
        cudaEventElapsedTime(&ms, pe.second.first, pe.second.second);
        h->stage_ms[pe.first] += ms;
        h->ev_pool.push_back(pe.second.first);
        h->ev_pool.push_back(pe.second.second);
    }
    h->pending.clear();
}


// internal helpers implemented in capi.cu (C linkage only because they live inside its extern "C" block)
extern "C" int wsg_make_plan(wsg_handle* h, int rows, int cols, const wsg_sgbm_params* p, SgbmPlan& pl);
extern "C" int wsg_check_sweep(wsg_handle* h);
extern "C" int wsg_run_sgbm(wsg_handle* h, const uint8_t* d_img1, const uint8_t* d_img2, size_t stride, int16_t* d_disp);
extern "C" int wsg_run_sgbm_batch(wsg_handle* h, int n, const uint8_t* const* d_img1, const uint8_t* const* d_img2, size_t stride,
                                  int16_t* const* d_disp);
