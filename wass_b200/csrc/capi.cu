// C ABI of libwassgpu.so (declared in include/wassgpu.h).  Thin: argument checks, a grow-only
// device arena per handle, kernel launches on the handle's stream.  No CPU fallback anywhere.
#include "handle.cuh"

extern "C" {

const char* wsg_version(void) { return "wassgpu 0.1.0 (sm_100a)"; }

int wsg_create(int device, wsg_handle** out)
{
    if (!out) return WSG_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return WSG_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return WSG_ERR_CUDA;
    wsg_handle* h = new wsg_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return WSG_ERR_CUDA; }
    h->stream = h->own_stream;
    if (const char* e = getenv("WSG_AGG_IMPL")) h->agg_impl = std::min(std::max(atoi(e), 0), 4);
    *out = h;
    return WSG_OK;
}

void wsg_destroy(wsg_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    drain_profile(h);
    for (DevBuf* b : {&h->pre1, &h->pre2, &h->C, &h->S, &h->raw, &h->img1, &h->img2, &h->disp, &h->scalars, &h->bnd, &h->keys,
                      &h->d1, &h->dbg, &h->crop_l, &h->crop_r, &h->rs_l, &h->rs_r, &h->rs_tab, &h->fc,
                      &h->fa, &h->fb, &h->dispfull, &h->im_left, &h->im_right, &h->mask_l, &h->mask_r, &h->m_valid, &h->m_X, &h->m_Y,
                      &h->m_Z, &h->m_color, &h->m_labels, &h->m_scratch, &h->m_small, &h->m_out})
        if (b->p) cudaFree(b->p);
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(h->own_stream);
    delete h;
}

int wsg_set_stream(wsg_handle* h, void* cuda_stream)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return WSG_OK;
}

int wsg_synchronize(wsg_handle* h)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    CK(h, cudaStreamSynchronize(h->stream));
    return WSG_OK;
}

const char* wsg_last_error(const wsg_handle* h) { return h ? h->err.c_str() : "null handle"; }

int wsg_make_plan(wsg_handle* h, int rows, int cols, const wsg_sgbm_params* p, SgbmPlan& pl)
{
    if (!p || rows <= 0 || cols <= 0) { h->err = "bad image size or null params"; return WSG_ERR_INVALID_ARG; }
    if (p->numDisparities <= 0 || p->numDisparities % 16 || p->numDisparities > 1280) {
        h->err = "numDisparities must be a positive multiple of 16, <= 1280"; return WSG_ERR_INVALID_ARG;
    }
    if (p->blockSize > 25 || (p->blockSize > 0 && p->blockSize % 2 == 0)) {
        h->err = "blockSize must be odd and <= 25"; return WSG_ERR_INVALID_ARG;
    }
    if (p->mode != WSG_MODE_SGBM && p->mode != WSG_MODE_HH) { h->err = "mode must be 0 (SGBM) or 1 (HH)"; return WSG_ERR_INVALID_ARG; }
    pl.H = rows; pl.W = cols;
    pl.minD = p->minDisparity; pl.D = p->numDisparities; pl.maxD = pl.minD + pl.D;
    pl.SW2 = pl.SH2 = p->blockSize > 0 ? p->blockSize / 2 : 1;
    pl.ftzero = std::max(p->preFilterCap, 15) | 1;
    pl.P1 = p->P1 > 0 ? p->P1 : 2;
    pl.P2 = std::max(p->P2 > 0 ? p->P2 : 5, pl.P1 + 1);
    if (pl.P2 > 32767 || pl.ftzero > 127) { h->err = "P2 must be <= 32767 and preFilterCap <= 127"; return WSG_ERR_INVALID_ARG; }
    pl.uniq = p->uniquenessRatio >= 0 ? p->uniquenessRatio : 10;
    pl.d12 = p->disp12MaxDiff > 0 ? p->disp12MaxDiff : 1;
    pl.minX1 = std::max(pl.maxD, 0);
    pl.maxX1 = cols + std::min(pl.minD, 0);
    pl.W1 = pl.maxX1 - pl.minX1;
    pl.INVALID = (pl.minD - 1) * 16;
    pl.mode = p->mode;
    pl.speckleWindow = p->speckleWindowSize > 0 ? p->speckleWindowSize : 0;
    pl.speckleMaxDiff = 16 * p->speckleRange;
    // cv2 raises for images this narrow (stereosgbm.cpp:511); mirror it as an error code
    if (cols - (pl.minD + pl.D) <= pl.SW2 || pl.W1 <= 0) { h->err = "image too narrow for minDisparity+numDisparities and blockSize"; return WSG_ERR_TOO_SMALL; }
    const int NV = pl.D / 8;
    if (h->agg_impl != WSG_AGG_PER_DIRECTION && NV <= 64) { pl.NL = 32; pl.K = (NV + 31) / 32; }   // one warp per pixel
    else if (NV <= 8) { pl.NL = 8; pl.K = 1; }
    else if (NV <= 16) { pl.NL = 16; pl.K = 1; }
    else { pl.NL = 32; pl.K = (NV + 31) / 32; }
    // tuning override (same results, different lane mapping): WSG_AGG_LANES=8|16 for D=256
    if (const char* e = getenv("WSG_AGG_LANES")) {
        const int nl = atoi(e);
        if (h->agg_impl == WSG_AGG_PER_DIRECTION && NV == 32 && (nl == 8 || nl == 16)) { pl.NL = nl; pl.K = 32 / nl; }
    }
    pl.Dp = pl.NL * pl.K * 8;
    return WSG_OK;
}

// The fused sweeps hand states between CTAs with bounded waits; an overrun (never seen on a healthy device) raises a
// flag instead of hanging.  Synchronises the stream.
int wsg_check_sweep(wsg_handle* h)
{
    if (!h->scalars.p || h->stats.agg_impl == WSG_AGG_PER_DIRECTION) return WSG_OK;
    int flag = 0;
    CK(h, cudaMemcpyAsync(&flag, (int*)h->scalars.p + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->dbg.p && getenv("WSG_SWEEP_DEBUG")) {
        const size_t nb = (h->plan.H + 7) / 8;
        std::vector<int> d(16384);
        cudaMemcpy(d.data(), h->dbg.p, d.size() * sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[wsg] last sweep: band sm start_us end_us\n");
        for (size_t i = 0; i < nb && i < 2048; ++i)
            fprintf(stderr, "[wsg] %zu %d %.1f %.1f\n", i, d[i], (d[4096 + 2 * i] - d[4096]) * 1e-3, (d[4096 + 2 * i + 1] - d[4096]) * 1e-3);
    }
    if (flag) { h->err = "fused aggregation sweep: hand-off wait overran (code " + std::to_string(flag) + ")"; return WSG_ERR_CUDA; }
    return WSG_OK;
}

int wsg_run_sgbm(wsg_handle* h, const uint8_t* d_img1, const uint8_t* d_img2, size_t stride, int16_t* d_disp)
{
    const SgbmPlan& pl = h->plan;
    const size_t npix = (size_t)pl.H * pl.W;
    const size_t vol = (size_t)pl.H * pl.W1 * pl.Dp * sizeof(int16_t);
    int rc;
    if ((rc = ensure(h, h->pre1, npix * sizeof(uint2)))) return rc;
    if ((rc = ensure(h, h->pre2, npix * sizeof(uint2)))) return rc;
    if ((rc = ensure(h, h->C, vol))) return rc;
    if ((rc = ensure(h, h->S, vol))) return rc;
    if ((rc = ensure(h, h->raw, npix * sizeof(int16_t)))) return rc;
    const size_t scal_bytes = (16 + 2 * WSG_SWEEP_TICKET_INTS) * sizeof(int);   // [0] max C, [1] sweep error, [16..] hand-out counters
    if ((rc = ensure(h, h->scalars, scal_bytes))) return rc;
    int launches = 0;
    CK(h, cudaMemsetAsync(h->scalars.p, 0, scal_bytes, h->stream));
    {
        StageTimer t(h, WSG_STAGE_PREFILTER, 2);
        launch_prefilter(d_img1, stride, (uint2*)h->pre1.p, pl, h->stream);
        launch_prefilter(d_img2, stride, (uint2*)h->pre2.p, pl, h->stream);
        launches += 2;
    }
    {
        StageTimer t(h, WSG_STAGE_COST, 1);
        launch_cost((const uint2*)h->pre1.p, (const uint2*)h->pre2.p, (int16_t*)h->C.p, (int*)h->scalars.p, pl, h->stream, &launches);
    }
    const int impl = (h->agg_impl != WSG_AGG_PER_DIRECTION && sweep_supported(pl)) ? h->agg_impl : WSG_AGG_PER_DIRECTION;
    if (impl == WSG_AGG_PER_DIRECTION) {
        {
            const int ndirs = pl.mode == WSG_MODE_HH ? 8 : 5;
            StageTimer t(h, WSG_STAGE_AGGREGATE, ndirs);
            for (int r = 0; r < ndirs; ++r)
                launch_aggregate_dir((const int16_t*)h->C.p, (int16_t*)h->S.p, r, r == 0, pl, h->stream);
            launches += ndirs;
        }
        {
            StageTimer t(h, WSG_STAGE_WTA, 1);
            launch_wta((const int16_t*)h->S.p, (int16_t*)h->raw.p, pl, h->stream);
            launches += 1;
        }
    } else {
        // fused wavefront sweeps (sweep_kernels.cu).  scalars: [0] max C, [1] error flag, [2..] band tickets
        const size_t bbytes = sweep_boundary_bytes(pl);
        const bool grown = bbytes > h->bnd.cap;
        if ((rc = ensure(h, h->bnd, bbytes))) return rc;
        const int bnd_nd = impl == WSG_AGG_SWEEPS3_WTA ? 3 : 4;
        if (grown || h->bnd_H != pl.H || h->bnd_W1 != pl.W1 || h->bnd_K != pl.K || h->bnd_nd != bnd_nd) {
            // epoch tags only tell "this sweep" from "the previous one" for slots that are rewritten every sweep: a change
            // of geometry, or of the set of states a sweep hands down (the 3-direction sweeps leave one third of every slot
            // untouched), would let data of an older sweep with the same 2-bit epoch pass for current
            CK(h, cudaMemsetAsync(h->bnd.p, 0, h->bnd.cap, h->stream));
            h->bnd_H = pl.H; h->bnd_W1 = pl.W1; h->bnd_K = pl.K; h->bnd_nd = bnd_nd;
        }
        const bool fused_wta = (impl == WSG_AGG_SWEEPS_WTA || impl == WSG_AGG_SWEEPS3_WTA || impl == WSG_AGG_SWEEPS2W_WTA) && pl.uniq < 100;   // the in-sweep WTA needs 100-uniq > 0
        if (fused_wta) {
            if ((rc = ensure(h, h->keys, npix * sizeof(unsigned long long)))) return rc;
            if ((rc = ensure(h, h->d1, npix * sizeof(int16_t)))) return rc;
        }
        SweepScratch sc;
        sc.boundary = h->bnd.p;
        sc.two_warps = impl == WSG_AGG_SWEEPS2W_WTA;
        sc.max_workers = h->sweep_workers;
        sc.maxC = (const int*)h->scalars.p;
        sc.err = (int*)h->scalars.p + 1;
        sc.dbg = nullptr;
        if (getenv("WSG_SWEEP_DEBUG")) {
            if ((rc = ensure(h, h->dbg, 16384 * sizeof(int)))) return rc;
            sc.dbg = (int*)h->dbg.p;
        }
        sc.keys = (unsigned long long*)h->keys.p;
        sc.d1 = (int16_t*)h->d1.p;
        int* tickets = (int*)h->scalars.p + 16;
        if (h->num_sms == 0) CK(h, cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
        sc.num_sms = h->num_sms;
        auto next_epoch = [&]() { h->sweep_epoch = h->sweep_epoch % 3 + 1; return h->sweep_epoch; };
        const int last_mode = fused_wta ? 2 : 1;
        // WSG_AGG_SWEEPS3_WTA: the sweeps leave out the (x+1,y-1)-type path (directions 3 and 5), which forces a skew of two
        // columns per row on the wavefront; those two run as independent per-direction launches (HBM-bound) in between.
        const bool split = impl == WSG_AGG_SWEEPS3_WTA;
        const int nd = split ? 3 : 4;
        {
            const int nl = 2 + (fused_wta ? 1 : 0) + (split ? (pl.mode == WSG_MODE_HH ? 2 : 1) : 0);
            StageTimer t(h, WSG_STAGE_AGGREGATE, nl);
            if (fused_wta) launch_wta_reset(sc, pl, h->stream);
            sc.ticket = tickets + 0; sc.epoch = next_epoch();
            launch_sweep((const int16_t*)h->C.p, (int16_t*)h->S.p, 0, 0, nd, pl, sc, h->stream);
            if (split) {
                launch_aggregate_dir((const int16_t*)h->C.p, (int16_t*)h->S.p, 3, false, pl, h->stream);
                if (pl.mode == WSG_MODE_HH) launch_aggregate_dir((const int16_t*)h->C.p, (int16_t*)h->S.p, 5, false, pl, h->stream);
            }
            sc.ticket = tickets + WSG_SWEEP_TICKET_INTS;
            if (pl.mode == WSG_MODE_HH) {
                sc.epoch = next_epoch();
                launch_sweep((const int16_t*)h->C.p, (int16_t*)h->S.p, 1, last_mode, nd, pl, sc, h->stream);
            } else {
                launch_sweep((const int16_t*)h->C.p, (int16_t*)h->S.p, 1, last_mode, 1, pl, sc, h->stream);
            }
            launches += nl;
        }
        {
            StageTimer t(h, WSG_STAGE_WTA, 1);
            if (fused_wta) launch_lrcheck(sc, (int16_t*)h->raw.p, pl, h->stream);
            else launch_wta((const int16_t*)h->S.p, (int16_t*)h->raw.p, pl, h->stream);
            launches += 1;
        }
    }
    h->stats.agg_impl = impl;
    {
        const int nsp = pl.speckleWindow > 0 ? 4 : 0;
        if (nsp && (rc = ensure(h, h->keys, npix * sizeof(unsigned long long)))) return rc;   // free again after the LR check
        StageTimer t(h, WSG_STAGE_MEDIAN, 1 + nsp);
        launch_median3((const int16_t*)h->raw.p, d_disp, pl.H, pl.W, h->stream);
        if (nsp)
            launch_filter_speckles(d_disp, pl.H, pl.W, pl.INVALID, pl.speckleWindow, pl.speckleMaxDiff, (int*)h->keys.p,
                                   (unsigned*)h->keys.p + npix, h->stream);
        launches += 1 + nsp;
    }
    CK(h, cudaGetLastError());
    h->stats.kernel_launches = launches;
    h->stats.width1 = pl.W1;
    h->stats.d_padded = pl.Dp;
    h->stats.volume_bytes = (long long)vol;
    h->have_plan = true;
    return WSG_OK;
}

int wsg_sgbm_compute_device(wsg_handle* h, const uint8_t* d_img1, const uint8_t* d_img2, int rows, int cols,
                            size_t stride, const wsg_sgbm_params* p, int16_t* d_disp16)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!d_img1 || !d_img2 || !d_disp16 || stride < (size_t)std::max(cols, 0)) { h->err = "null pointer or stride < cols"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    SgbmPlan pl{};
    int rc = wsg_make_plan(h, rows, cols, p, pl);
    if (rc) return rc;
    h->plan = pl;
    h->stats.out_of_domain = 0; h->stats.max_cost = 0;
    return wsg_run_sgbm(h, d_img1, d_img2, stride, d_disp16);
}

int wsg_sgbm_compute(wsg_handle* h, const uint8_t* img1, const uint8_t* img2, int rows, int cols, size_t stride,
                     const wsg_sgbm_params* p, int16_t* disp16)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!img1 || !img2 || !disp16 || stride < (size_t)std::max(cols, 0)) { h->err = "null pointer or stride < cols"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    SgbmPlan pl{};
    int rc = wsg_make_plan(h, rows, cols, p, pl);
    if (rc) return rc;
    h->plan = pl;
    const size_t npix = (size_t)rows * cols;
    if ((rc = ensure(h, h->img1, npix))) return rc;
    if ((rc = ensure(h, h->img2, npix))) return rc;
    if ((rc = ensure(h, h->disp, npix * sizeof(int16_t)))) return rc;
    CK(h, cudaMemcpy2DAsync(h->img1.p, cols, img1, stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemcpy2DAsync(h->img2.p, cols, img2, stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    rc = wsg_run_sgbm(h, (const uint8_t*)h->img1.p, (const uint8_t*)h->img2.p, cols, (int16_t*)h->disp.p);
    if (rc) return rc;
    CK(h, cudaMemcpyAsync(disp16, h->disp.p, npix * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return wsg_check_sweep(h);
}

int wsg_sgbm_get_stats(wsg_handle* h, wsg_sgbm_stats* out)
{
    if (!h || !out) return WSG_ERR_INVALID_ARG;
    if (!h->have_plan) { h->err = "no compute yet"; return WSG_ERR_STATE; }
    int maxc = 0;
    CK(h, cudaMemcpyAsync(&maxc, h->scalars.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    h->stats.max_cost = maxc;
    h->stats.out_of_domain = (maxc + h->plan.P2 > 32767) ? 1 : 0;
    *out = h->stats;
    return wsg_check_sweep(h);
}

int wsg_sgbm_debug_volumes(wsg_handle* h, int16_t* C_host, int16_t* S_host)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (!h->have_plan) { h->err = "no compute yet"; return WSG_ERR_STATE; }
    const SgbmPlan& pl = h->plan;
    if (S_host && (h->stats.agg_impl >= WSG_AGG_SWEEPS_WTA)) {
        h->err = "S is never materialised by WSG_AGG_SWEEPS_WTA; select WSG_AGG_SWEEPS or WSG_AGG_PER_DIRECTION first";
        return WSG_ERR_STATE;
    }
    const size_t npx = (size_t)pl.H * pl.W1;
    std::vector<int16_t> tmp(npx * pl.Dp);
    for (int which = 0; which < 2; ++which) {
        int16_t* dst = which ? S_host : C_host;
        if (!dst) continue;
        CK(h, cudaMemcpyAsync(tmp.data(), which ? h->S.p : h->C.p, tmp.size() * 2, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        for (size_t px = 0; px < npx; ++px)
            for (int j = 0; j < pl.D / 8; ++j)
                for (int i = 0; i < 8; ++i)     // undo the in-vector interleave (vec_pos)
                    dst[px * pl.D + (size_t)j * 8 + i] = tmp[px * pl.Dp + (size_t)vec_slot(j, pl.NL, pl.K) * 8 + vec_pos(i)];
    }
    return WSG_OK;
}

int wsg_sgbm_set_sweep_workers(wsg_handle* h, int max_sms)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (max_sms < 0) { h->err = "max_sms must be >= 0"; return WSG_ERR_INVALID_ARG; }
    h->sweep_workers = max_sms;
    return WSG_OK;
}

int wsg_sgbm_set_impl(wsg_handle* h, int impl)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (impl < WSG_AGG_PER_DIRECTION || impl > WSG_AGG_SWEEPS2W_WTA) { h->err = "unknown aggregation implementation"; return WSG_ERR_INVALID_ARG; }
    h->agg_impl = impl;
    return WSG_OK;
}

int wsg_profile_enable(wsg_handle* h, int enable)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    h->prof = enable != 0;
    return WSG_OK;
}

int wsg_profile_reset(wsg_handle* h)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    cudaStreamSynchronize(h->stream);
    drain_profile(h);
    for (int i = 0; i < WSG_NUM_STAGES; ++i) { h->stage_ms[i] = 0; h->stage_launches[i] = 0; }
    return WSG_OK;
}

int wsg_profile_get(wsg_handle* h, float* ms, int* launches, int n)
{
    if (!h || n < 0) return WSG_ERR_INVALID_ARG;
    CK(h, cudaStreamSynchronize(h->stream));
    drain_profile(h);
    for (int i = 0; i < n && i < WSG_NUM_STAGES; ++i) {
        if (ms) ms[i] = h->stage_ms[i];
        if (launches) launches[i] = h->stage_launches[i];
    }
    return WSG_OK;
}

}  // extern "C"
