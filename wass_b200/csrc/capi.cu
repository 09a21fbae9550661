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
    if (const char* e = getenv("WSG_AGG_IMPL")) h->agg_impl = std::min(std::max(atoi(e), 0), 2);
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
                      &h->d1, &h->dbg, &h->crop_l, &h->crop_r, &h->rs_l, &h->rs_r, &h->rs_tab, &h->fc, &h->fbatch,
                      &h->fa, &h->fb, &h->dispfull, &h->im_left, &h->im_right, &h->mask_l, &h->mask_r, &h->m_valid, &h->m_X, &h->m_Y,
                      &h->m_Z, &h->m_color, &h->m_labels, &h->m_scratch, &h->m_small, &h->m_out})
        if (b->p) cudaFree(b->p);
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    if (h->h2d_stream) {
        cudaStreamSynchronize(h->h2d_stream); cudaStreamSynchronize(h->d2h_stream);
        cudaStreamDestroy(h->h2d_stream); cudaStreamDestroy(h->d2h_stream);
        for (int k = 0; k < 2; ++k) { cudaEventDestroy(h->ev_h2d[k]); cudaEventDestroy(h->ev_done[k]); cudaEventDestroy(h->ev_d2h[k]); }
        cudaFreeHost(h->async_flag);
        for (int k = 0; k < 2; ++k) for (DevBuf* b : {&h->a_img1[k], &h->a_img2[k], &h->a_disp[k]}) if (b->p) cudaFree(b->p);
    }
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

// Per-handle scalars (ints): [1] sweep error flag, [16] / [32] band tickets of the two sweeps, [SCAL_MAXC + f] max over the
// cost volume of frame f of the batch.
static constexpr int SCAL_ERR = 1, SCAL_TICKET0 = 16, SCAL_TICKET1 = 32, SCAL_MAXC = 48;

// The fused sweeps hand states between CTAs with bounded waits; an overrun (never seen on a healthy device) raises a
// flag instead of hanging.  Synchronises the stream.
int wsg_check_sweep(wsg_handle* h)
{
    if (!h->scalars.p || h->stats.agg_impl == WSG_AGG_PER_DIRECTION) return WSG_OK;
    int flag = 0;
    CK(h, cudaMemcpyAsync(&flag, (int*)h->scalars.p + SCAL_ERR, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->dbg.p && getenv("WSG_SWEEP_DEBUG")) {
        const int R = h->sweep_rows;
        const size_t nt = (size_t)((h->plan.H + R - 1) / R) * h->batch_n;
        std::vector<int> d(3 * nt);
        cudaMemcpy(d.data(), h->dbg.p, d.size() * sizeof(int), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[wsg] last sweep: ticket frame band sm start_us end_us\n");
        for (size_t i = 0; i < nt; ++i)
            fprintf(stderr, "[wsg] %zu %zu %zu %d %.1f %.1f\n", i, i % h->batch_n, i / h->batch_n, d[3 * i],
                    (d[3 * i + 1] - d[1]) * 1e-3, (d[3 * i + 2] - d[1]) * 1e-3);
    }
    if (flag) { h->err = "fused aggregation sweep: hand-off wait overran (code " + std::to_string(flag) + ")"; return WSG_ERR_CUDA; }
    return WSG_OK;
}

// The dense matcher on `n` frames of the planned geometry, device pointers, asynchronous on the handle's stream.
// Prefilter + cost volume frame by frame, then ONE launch per sweep over the bands of all frames (sweep_kernels.cu),
// then LR check + median frame by frame.
int wsg_run_sgbm_batch(wsg_handle* h, int n, const uint8_t* const* d_img1, const uint8_t* const* d_img2, size_t stride,
                       int16_t* const* d_disp)
{
    const SgbmPlan& pl = h->plan;
    const size_t npix = (size_t)pl.H * pl.W;
    const size_t vol = (size_t)pl.H * pl.W1 * pl.Dp * sizeof(int16_t);
    const size_t pad = sweep_volume_pad_bytes();
    int rc;
    if ((rc = ensure(h, h->pre1, npix * sizeof(uint2)))) return rc;
    if ((rc = ensure(h, h->pre2, npix * sizeof(uint2)))) return rc;
    if ((rc = ensure(h, h->C, vol * n + 2 * pad))) return rc;
    if ((rc = ensure(h, h->S, vol * n + 2 * pad))) return rc;
    if ((rc = ensure(h, h->raw, npix * sizeof(int16_t)))) return rc;
    const size_t scal_bytes = (size_t)(SCAL_MAXC + n) * sizeof(int);
    if ((rc = ensure(h, h->scalars, std::max(scal_bytes, (size_t)4096)))) return rc;
    int16_t* Cv = (int16_t*)((char*)h->C.p + pad);
    int16_t* Sv = (int16_t*)((char*)h->S.p + pad);
    int* scal = (int*)h->scalars.p;
    int launches = 0;
    CK(h, cudaMemsetAsync(h->scalars.p, 0, scal_bytes, h->stream));
    for (int f = 0; f < n; ++f) {
        {
            StageTimer t(h, WSG_STAGE_PREFILTER, 2);
            launch_prefilter(d_img1[f], stride, (uint2*)h->pre1.p, pl, h->stream);
            launch_prefilter(d_img2[f], stride, (uint2*)h->pre2.p, pl, h->stream);
            launches += 2;
        }
        {
            StageTimer t(h, WSG_STAGE_COST, 1);
            launch_cost((const uint2*)h->pre1.p, (const uint2*)h->pre2.p, Cv + (size_t)f * (vol / 2), scal + SCAL_MAXC + f, pl,
                        h->stream, &launches);
        }
    }
    const int impl = (h->agg_impl != WSG_AGG_PER_DIRECTION && sweep_supported(pl)) ? h->agg_impl : WSG_AGG_PER_DIRECTION;
    if (impl == WSG_AGG_PER_DIRECTION) {
        for (int f = 0; f < n; ++f) {
            int16_t* Cf = Cv + (size_t)f * (vol / 2);
            {
                const int ndirs = pl.mode == WSG_MODE_HH ? 8 : 5;
                StageTimer t(h, WSG_STAGE_AGGREGATE, ndirs);
                for (int r = 0; r < ndirs; ++r) launch_aggregate_dir(Cf, Sv, r, r == 0, pl, h->stream);
                launches += ndirs;
            }
            {
                StageTimer t(h, WSG_STAGE_WTA, 1);
                launch_wta(Sv, (int16_t*)h->raw.p, pl, h->stream);
                launches += 1;
            }
            {
                StageTimer t(h, WSG_STAGE_MEDIAN, 1);
                launch_median3((const int16_t*)h->raw.p, d_disp[f], pl.H, pl.W, h->stream);
                launches += 1;
            }
        }
    } else {
        // fused wavefront sweeps over the whole batch (sweep_kernels.cu)
        if (h->num_sms == 0) CK(h, cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
        const int workers = h->sweep_workers > 0 ? std::min(h->sweep_workers, h->num_sms) : h->num_sms;
        const int rows = sweep_rows_per_band(pl, n, workers);
        h->sweep_rows = rows;
        const size_t bbytes = sweep_boundary_bytes(pl, rows) * n;
        const bool grown = bbytes > h->bnd.cap;
        if ((rc = ensure(h, h->bnd, bbytes))) return rc;
        if (grown || h->bnd_H != pl.H || h->bnd_W1 != pl.W1 || h->bnd_K != pl.K || h->bnd_n != n || h->bnd_rows != rows) {
            // epoch tags only tell "this sweep" from "the previous one" for slots that are rewritten every sweep: a change
            // of geometry or batch size would let data of an older sweep with the same 2-bit epoch pass for current
            CK(h, cudaMemsetAsync(h->bnd.p, 0, h->bnd.cap, h->stream));
            h->bnd_H = pl.H; h->bnd_W1 = pl.W1; h->bnd_K = pl.K; h->bnd_n = n; h->bnd_rows = rows;
        }
        const bool fused_wta = impl == WSG_AGG_SWEEPS_WTA && pl.uniq < 99;    // the in-sweep WTA needs 100-uniq >= 2 (its multiply-high division)
        if (fused_wta) {
            if ((rc = ensure(h, h->keys, npix * n * sizeof(unsigned long long)))) return rc;
            if ((rc = ensure(h, h->d1, npix * n * sizeof(int16_t)))) return rc;
        }
        SweepScratch sc;
        sc.boundary = h->bnd.p;
        sc.max_workers = h->sweep_workers;
        sc.maxC = scal + SCAL_MAXC; sc.maxC_stride = 1;
        sc.err = scal + SCAL_ERR;
        sc.dbg = nullptr;
        sc.nframes = n; sc.volume_stride_bytes = vol; sc.rows = rows;
        if (getenv("WSG_SWEEP_DEBUG")) {
            const int R = rows;
            if ((rc = ensure(h, h->dbg, (size_t)3 * ((pl.H + R - 1) / R) * n * sizeof(int)))) return rc;
            sc.dbg = (int*)h->dbg.p;
        }
        sc.keys = (unsigned long long*)h->keys.p;
        sc.d1 = (int16_t*)h->d1.p;
        sc.num_sms = h->num_sms;
        auto next_epoch = [&]() { h->sweep_epoch = h->sweep_epoch % 3 + 1; return h->sweep_epoch; };
        const int last_mode = fused_wta ? 2 : 1;
        {
            const int nl = 2 + (fused_wta ? 1 : 0);
            StageTimer t(h, WSG_STAGE_AGGREGATE, nl);
            if (fused_wta) launch_wta_reset(sc, pl, h->stream);
            sc.ticket = scal + SCAL_TICKET0; sc.epoch = next_epoch();
            launch_sweep(Cv, Sv, 0, 0, 4, pl, sc, h->stream);
            sc.ticket = scal + SCAL_TICKET1;
            if (pl.mode == WSG_MODE_HH) {
                sc.epoch = next_epoch();
                launch_sweep(Cv, Sv, 1, last_mode, 4, pl, sc, h->stream);
            } else {
                launch_sweep(Cv, Sv, 1, last_mode, 1, pl, sc, h->stream);
            }
            launches += nl;
        }
        for (int f = 0; f < n; ++f) {
            {
                StageTimer t(h, WSG_STAGE_WTA, 1);
                if (fused_wta) launch_lrcheck(sc, f, (int16_t*)h->raw.p, pl, h->stream);
                else launch_wta(Sv + (size_t)f * (vol / 2), (int16_t*)h->raw.p, pl, h->stream);
                launches += 1;
            }
            {
                StageTimer t(h, WSG_STAGE_MEDIAN, 1);
                launch_median3((const int16_t*)h->raw.p, d_disp[f], pl.H, pl.W, h->stream);
                launches += 1;
            }
        }
    }
    h->stats.agg_impl = impl;
    if (pl.speckleWindow > 0) {
        if ((rc = ensure(h, h->keys, npix * sizeof(unsigned long long)))) return rc;   // free again after the LR check
        for (int f = 0; f < n; ++f) {
            StageTimer t(h, WSG_STAGE_MEDIAN, 4);
            launch_filter_speckles(d_disp[f], pl.H, pl.W, pl.INVALID, pl.speckleWindow, pl.speckleMaxDiff, (int*)h->keys.p,
                                   (unsigned*)h->keys.p + npix, h->stream);
            launches += 4;
        }
    }
    CK(h, cudaGetLastError());
    h->batch_n = n;
    h->stats.kernel_launches = launches;
    h->stats.width1 = pl.W1;
    h->stats.d_padded = pl.Dp;
    h->stats.volume_bytes = (long long)vol;
    h->have_plan = true;
    return WSG_OK;
}

int wsg_run_sgbm(wsg_handle* h, const uint8_t* d_img1, const uint8_t* d_img2, size_t stride, int16_t* d_disp)
{
    return wsg_run_sgbm_batch(h, 1, &d_img1, &d_img2, stride, &d_disp);
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
    return wsg_sgbm_compute_batch(h, 1, &img1, &img2, rows, cols, stride, p, &disp16);
}

int wsg_sgbm_compute_batch_device(wsg_handle* h, int n, const uint8_t* d_img1, const uint8_t* d_img2, size_t frame_stride,
                                  int rows, int cols, size_t stride, const wsg_sgbm_params* p, int16_t* d_disp16)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (n <= 0 || n > WSG_MAX_BATCH) { h->err = "batch size must be 1.." + std::to_string(WSG_MAX_BATCH); return WSG_ERR_INVALID_ARG; }
    if (!d_img1 || !d_img2 || !d_disp16 || stride < (size_t)std::max(cols, 0) || frame_stride < stride * (size_t)std::max(rows, 0)) {
        h->err = "null pointer, stride < cols or frame_stride < rows*stride"; return WSG_ERR_INVALID_ARG;
    }
    CK(h, cudaSetDevice(h->device));
    SgbmPlan pl{};
    int rc = wsg_make_plan(h, rows, cols, p, pl);
    if (rc) return rc;
    h->plan = pl;
    h->stats.out_of_domain = 0; h->stats.max_cost = 0;
    std::vector<const uint8_t*> a(n), b(n);
    std::vector<int16_t*> d(n);
    for (int f = 0; f < n; ++f) {
        a[f] = d_img1 + (size_t)f * frame_stride; b[f] = d_img2 + (size_t)f * frame_stride;
        d[f] = d_disp16 + (size_t)f * rows * cols;
    }
    return wsg_run_sgbm_batch(h, n, a.data(), b.data(), stride, d.data());
}

int wsg_sgbm_compute_batch(wsg_handle* h, int n, const uint8_t* const* img1, const uint8_t* const* img2, int rows, int cols,
                           size_t stride, const wsg_sgbm_params* p, int16_t* const* disp16)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (n <= 0 || n > WSG_MAX_BATCH) { h->err = "batch size must be 1.." + std::to_string(WSG_MAX_BATCH); return WSG_ERR_INVALID_ARG; }
    if (!img1 || !img2 || !disp16 || stride < (size_t)std::max(cols, 0)) { h->err = "null pointer or stride < cols"; return WSG_ERR_INVALID_ARG; }
    for (int f = 0; f < n; ++f)
        if (!img1[f] || !img2[f] || !disp16[f]) { h->err = "null frame pointer"; return WSG_ERR_INVALID_ARG; }
    CK(h, cudaSetDevice(h->device));
    SgbmPlan pl{};
    int rc = wsg_make_plan(h, rows, cols, p, pl);
    if (rc) return rc;
    h->plan = pl;
    const size_t npix = (size_t)rows * cols;
    if ((rc = ensure(h, h->img1, npix * n))) return rc;
    if ((rc = ensure(h, h->img2, npix * n))) return rc;
    if ((rc = ensure(h, h->disp, npix * n * sizeof(int16_t)))) return rc;
    std::vector<const uint8_t*> a(n), b(n);
    std::vector<int16_t*> d(n);
    for (int f = 0; f < n; ++f) {
        a[f] = (const uint8_t*)h->img1.p + f * npix; b[f] = (const uint8_t*)h->img2.p + f * npix;
        d[f] = (int16_t*)h->disp.p + f * npix;
        CK(h, cudaMemcpy2DAsync((void*)a[f], cols, img1[f], stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpy2DAsync((void*)b[f], cols, img2[f], stride, cols, rows, cudaMemcpyHostToDevice, h->stream));
    }
    rc = wsg_run_sgbm_batch(h, n, a.data(), b.data(), cols, d.data());
    if (rc) return rc;
    for (int f = 0; f < n; ++f)
        CK(h, cudaMemcpyAsync(disp16[f], d[f], npix * sizeof(int16_t), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return wsg_check_sweep(h);
}

// ---- asynchronous batches: submit enqueues H2D + matcher + D2H and returns, wait blocks for the result ---------------------
static int async_init(wsg_handle* h)
{
    if (h->h2d_stream) return WSG_OK;
    CK(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CK(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    for (int s = 0; s < 2; ++s) {
        CK(h, cudaEventCreateWithFlags(&h->ev_h2d[s], cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&h->ev_done[s], cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&h->ev_d2h[s], cudaEventDisableTiming));
    }
    CK(h, cudaHostAlloc((void**)&h->async_flag, 2 * sizeof(int), cudaHostAllocDefault));
    return WSG_OK;
}

int wsg_sgbm_batch_submit(wsg_handle* h, int slot, int n, const uint8_t* const* img1, const uint8_t* const* img2, int rows, int cols,
                          size_t stride, const wsg_sgbm_params* p, int16_t* const* disp16)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (slot < 0 || slot > 1 || n <= 0 || n > WSG_MAX_BATCH) { h->err = "slot must be 0 or 1, batch size 1.." + std::to_string(WSG_MAX_BATCH); return WSG_ERR_INVALID_ARG; }
    if (!img1 || !img2 || !disp16 || stride < (size_t)std::max(cols, 0)) { h->err = "null pointer or stride < cols"; return WSG_ERR_INVALID_ARG; }
    for (int f = 0; f < n; ++f)
        if (!img1[f] || !img2[f] || !disp16[f]) { h->err = "null frame pointer"; return WSG_ERR_INVALID_ARG; }
    if (h->async_busy[slot]) { h->err = "slot still in flight: wsg_sgbm_batch_wait first"; return WSG_ERR_STATE; }
    CK(h, cudaSetDevice(h->device));
    int rc = async_init(h);
    if (rc) return rc;
    SgbmPlan pl{};
    if ((rc = wsg_make_plan(h, rows, cols, p, pl))) return rc;
    h->plan = pl;
    const size_t npix = (size_t)rows * cols;
    if ((rc = ensure(h, h->a_img1[slot], npix * n))) return rc;
    if ((rc = ensure(h, h->a_img2[slot], npix * n))) return rc;
    if ((rc = ensure(h, h->a_disp[slot], npix * n * sizeof(int16_t)))) return rc;
    std::vector<const uint8_t*> a(n), b(n);
    std::vector<int16_t*> d(n);
    for (int f = 0; f < n; ++f) {
        a[f] = (const uint8_t*)h->a_img1[slot].p + f * npix; b[f] = (const uint8_t*)h->a_img2[slot].p + f * npix;
        d[f] = (int16_t*)h->a_disp[slot].p + f * npix;
        CK(h, cudaMemcpy2DAsync((void*)a[f], cols, img1[f], stride, cols, rows, cudaMemcpyHostToDevice, h->h2d_stream));
        CK(h, cudaMemcpy2DAsync((void*)b[f], cols, img2[f], stride, cols, rows, cudaMemcpyHostToDevice, h->h2d_stream));
    }
    CK(h, cudaEventRecord(h->ev_h2d[slot], h->h2d_stream));
    CK(h, cudaStreamWaitEvent(h->stream, h->ev_h2d[slot], 0));
    if ((rc = wsg_run_sgbm_batch(h, n, a.data(), b.data(), cols, d.data()))) return rc;
    // the sweep error flag of this batch: read on the compute stream, before the next batch resets the scalars
    CK(h, cudaMemcpyAsync(h->async_flag + slot, (int*)h->scalars.p + SCAL_ERR, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaEventRecord(h->ev_done[slot], h->stream));
    CK(h, cudaStreamWaitEvent(h->d2h_stream, h->ev_done[slot], 0));
    for (int f = 0; f < n; ++f)
        CK(h, cudaMemcpyAsync(disp16[f], d[f], npix * sizeof(int16_t), cudaMemcpyDeviceToHost, h->d2h_stream));
    CK(h, cudaEventRecord(h->ev_d2h[slot], h->d2h_stream));
    h->async_busy[slot] = true;
    return WSG_OK;
}

int wsg_sgbm_batch_wait(wsg_handle* h, int slot)
{
    if (!h) return WSG_ERR_INVALID_ARG;
    if (slot < 0 || slot > 1 || !h->async_busy[slot]) { h->err = "nothing in flight in this slot"; return WSG_ERR_STATE; }
    CK(h, cudaEventSynchronize(h->ev_d2h[slot]));
    h->async_busy[slot] = false;
    if (h->stats.agg_impl != WSG_AGG_PER_DIRECTION && h->async_flag[slot]) {
        h->err = "fused aggregation sweep: hand-off wait overran (code " + std::to_string(h->async_flag[slot]) + ")";
        return WSG_ERR_CUDA;
    }
    return WSG_OK;
}

int wsg_sgbm_get_stats(wsg_handle* h, wsg_sgbm_stats* out)
{
    if (!h || !out) return WSG_ERR_INVALID_ARG;
    if (!h->have_plan) { h->err = "no compute yet"; return WSG_ERR_STATE; }
    std::vector<int> mc(std::max(h->batch_n, 1), 0);
    CK(h, cudaMemcpyAsync(mc.data(), (int*)h->scalars.p + SCAL_MAXC, mc.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    const int maxc = *std::max_element(mc.begin(), mc.end());      // over the frames of the last batch
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
        CK(h, cudaMemcpyAsync(tmp.data(), (char*)(which ? h->S.p : h->C.p) + sweep_volume_pad_bytes(), tmp.size() * 2, cudaMemcpyDeviceToHost, h->stream));   // frame 0
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
    if (impl < WSG_AGG_PER_DIRECTION || impl > WSG_AGG_SWEEPS_WTA) { h->err = "unknown aggregation implementation"; return WSG_ERR_INVALID_ARG; }
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
