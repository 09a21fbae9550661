// wass_stereo -- drop-in replacement of the reference's stage-4 executable
// (src/wass_stereo/wass_stereo.cpp:1799-2149): same argv, exit codes, progress protocol, configuration
// surface and workdir files; all per-pixel work runs on the GPU through the C ABI of libwassgpu.so.
#include "../../../include/wassgpu.h"
#include "config.hpp"
#include "io.hpp"

#include <sys/stat.h>
#include <sys/time.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

using namespace wasshost;

namespace {

// ---- logging: the NaiveLogger of src/include/log.hpp:78-168 (stdout + log file, "<scope> [sev  ] " prefix)
std::string g_scope;
std::ofstream* g_logfile = nullptr;
struct Log {
    explicit Log(const char* sev) { put(g_scope + " [" + sev + "] "); }
    ~Log() { std::cout << std::endl; if (g_logfile) (*g_logfile) << std::endl; }
    template <typename T> Log& operator<<(const T& v) { std::cout << v; if (g_logfile) (*g_logfile) << v; return *this; }
    static void put(const std::string& s) { std::cout << s; if (g_logfile) (*g_logfile) << s; }
};
#define LOGI Log("info ")
#define LOGE Log("error")
#define LOG_SCOPE(n) (g_scope = (n))

struct Timer {   // cvlab::HiresTimer (src/wass_lib/hires_timer.cpp): cumulative seconds + named events
    double t0 = 0, tstop = -1;
    std::vector<std::pair<double, std::string>> evts;
    static double now() { timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + tv.tv_usec * 1e-6; }
    void start() { t0 = now(); }
    void stop() { tstop = now(); }
    double elapsed() const { return (tstop >= 0 ? tstop : now()) - t0; }
    void mark(const char* name) { evts.push_back({elapsed(), name}); }
};

bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

Mat eye3() { Mat m; m.rows = m.cols = 3; m.v = {1, 0, 0, 0, 1, 0, 0, 0, 1}; return m; }
Mat zeros(int r, int c) { Mat m; m.rows = r; m.cols = c; m.v.assign((size_t)r * c, 0.0); return m; }
Mat mul(const Mat& a, const Mat& b)
{
    Mat o = zeros(a.rows, b.cols);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < b.cols; ++j) { double s = 0; for (int k = 0; k < a.cols; ++k) s += a.at(i, k) * b.at(k, j); o.at(i, j) = s; }
    return o;
}
Mat transpose(const Mat& a) { Mat o = zeros(a.cols, a.rows); for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) o.at(j, i) = a.at(i, j); return o; }
Mat stack_matrices(const Mat& R, const Mat& T)   // wass_stereo.cpp:181-194
{
    Mat o = zeros(3, 4);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) o.at(i, j) = R.at(i, j); o.at(i, 3) = T.at(i, 0); }
    return o;
}
void invert_RT(Mat& R, Mat& T)   // wass_stereo.cpp:197-201
{
    R = transpose(R);
    Mat t = mul(R, T);
    for (auto& x : t.v) x = -x;
    T = t;
}

struct Env {
    Timer timer;
    std::string workdir;
    Mat K0, K1, intr_left, intr_right, R, T, Rinv, Tinv, P0, P1, Rpose0, Tpose0, Rpose1, Tpose1;
    Image8 left, right, left_rect, right_rect, left_crop, right_crop;
    int left_index = 0, right_index = 1;
    double cam_distance = 1.0;
    double R1[9], R2[9], P1r[12], P2r[12];
    bool custom_rectify = false;                 // USE_CUSTOM_STEREORECTIFY: homographies instead of R1,R2,P1,P2
    double HL[9], HR[9], HLi[9], HRi[9];
    int roi_left[4], roi_right[4];
    int disparity_compensation = 0;
    const Image8* pre_left = nullptr;            // batch mode: images already decoded by the prefetch thread
    const Image8* pre_right = nullptr;
    void computeP() { P0 = mul(K0, stack_matrices(Rpose0, Tpose0)); P1 = mul(K1, stack_matrices(Rpose1, Tpose1)); }
    void swapLeftRight()   // wass_stereo.cpp:262-297
    {
        std::swap(left_index, right_index);
        std::swap(left, right);
        std::swap(intr_left, intr_right);
        std::swap(R, Rinv);
        std::swap(T, Tinv);
        std::swap(Rpose0, Rpose1);
        std::swap(Tpose0, Tpose1);
        invert_RT(Rpose0, Tpose0);
        invert_RT(Rpose1, Tpose1);
        computeP();
    }
};

std::string path(const Env& e, const char* f) { return e.workdir + "/" + f; }

int save_configuration(const Config& cfg, const std::string& filename)   // wass_stereo.cpp:1776-1795
{
    LOG_SCOPE("wass_stereo");
    LOGI << "Writing " << filename;
    std::ofstream ofs(filename.c_str());
    if (!ofs.is_open()) { LOGE << "Unable to open " << filename << " for write"; return -1; }
    ofs << cfg.to_config_string();
    ofs.close();
    LOGI << "Done!";
    return 0;
}

// bicubic (A=-0.75) resize for the *_s.png previews (wass_stereo.cpp:401-418); content is informative only
Image8 resize_cubic(const Image8& src, int nw, int nh)
{
    Image8 o; o.rows = nh; o.cols = nw; o.px.resize((size_t)nw * nh);
    auto wgt = [](float x, float* c) {
        const float A = -0.75f;
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
    };
    const double sx = (double)src.cols / nw, sy = (double)src.rows / nh;
    for (int y = 0; y < nh; ++y) {
        const double fy = (y + 0.5) * sy - 0.5; const int iy = (int)floor(fy); float cy[4]; wgt((float)(fy - iy), cy);
        for (int x = 0; x < nw; ++x) {
            const double fx = (x + 0.5) * sx - 0.5; const int ix = (int)floor(fx); float cx[4]; wgt((float)(fx - ix), cx);
            float s = 0;
            for (int a = 0; a < 4; ++a) {
                const int yy = std::min(std::max(iy - 1 + a, 0), src.rows - 1);
                for (int b = 0; b < 4; ++b) { const int xx = std::min(std::max(ix - 1 + b, 0), src.cols - 1); s += cy[a] * cx[b] * src.px[(size_t)yy * src.cols + xx]; }
            }
            o.px[(size_t)y * nw + x] = (uint8_t)std::min(std::max((int)lrintf(s), 0), 255);
        }
    }
    return o;
}

// WASS_HOST_PROFILE=1: wall time of the host stages of the main thread, summed over the workdirs of the process, on stderr
struct HostProf {
    struct Acc { std::map<std::string, double> t; std::vector<std::string> order; bool on = getenv("WASS_HOST_PROFILE") != nullptr; };
    static Acc& acc() { static Acc a; return a; }
    const char* name; std::chrono::steady_clock::time_point t0;
    explicit HostProf(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~HostProf()
    {
        Acc& a = acc();
        if (!a.on) return;
        if (!a.t.count(name)) a.order.push_back(name);
        a.t[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    static void report()
    {
        Acc& a = acc();
        if (!a.on) return;
        for (const auto& n : a.order) std::cerr << "[host] " << std::setw(28) << std::left << n << a.t[n] << " s" << std::endl;
    }
};

// Background file writers: the JPEG encoder (one thread per image) and the inlier dump run beside the device work;
// wait() before a workdir is reported as done.
struct JpegJobs {
    std::vector<std::thread> jobs;
    template <class F> void run(F&& f) { jobs.emplace_back(std::forward<F>(f)); }       // any other background file writer
    void wait() { for (auto& t : jobs) if (t.joinable()) t.join(); jobs.clear(); }
    ~JpegJobs() { wait(); }
} g_jpeg;
bool g_batch_mode = false;

bool load_data(Env& env, const Config& cfg)   // wass_stereo.cpp:337-442
{
    LOG_SCOPE("load_data");
    std::string err;
    if (!load_matrix_xml(path(env, "ext_R.xml"), env.R, &err)) LOGE << err;
    if (env.R.rows != 3 || env.R.cols != 3) { LOGE << "invalid extrinsic rotation matrix (ext_R.xml)"; return false; }
    if (!load_matrix_xml(path(env, "ext_T.xml"), env.T, &err)) LOGE << err;
    if (env.T.cols != 1 || env.T.rows != 3) { LOGE << "invalid extrinsic translation vector (ext_T.xml)"; return false; }
    env.Rinv = env.R; env.Tinv = env.T;
    invert_RT(env.Rinv, env.Tinv);
    const double cur = sqrt(env.T.at(0, 0) * env.T.at(0, 0) + env.T.at(1, 0) * env.T.at(1, 0) + env.T.at(2, 0) * env.T.at(2, 0));
    for (int i = 0; i < 3; ++i) {
        env.T.at(i, 0) = env.T.at(i, 0) / cur * env.cam_distance;
        env.Tinv.at(i, 0) = env.Tinv.at(i, 0) / cur * env.cam_distance;
    }
    env.Rpose0 = eye3(); env.Tpose0 = zeros(3, 1); env.Rpose1 = env.R; env.Tpose1 = env.T;
    if (!load_matrix_xml(path(env, "intrinsics_00000000.xml"), env.K0, &err)) { LOGE << err; return false; }
    if (!load_matrix_xml(path(env, "intrinsics_00000001.xml"), env.K1, &err)) { LOGE << err; return false; }
    if (env.K0.rows != 3 || env.K0.cols != 3 || env.K1.rows != 3 || env.K1.cols != 3) { LOGE << "invalid intrinsics"; return false; }
    env.intr_left = env.K0; env.intr_right = env.K1;
    env.computeP();
    if (env.pre_left) env.left = *env.pre_left;
    else if (!read_png_gray(path(env, "undistorted/00000000.png"), env.left, &err)) { LOGE << "unable to load input images"; LOGE << err; return false; }
    env.left_index = 0;
    LOGI << "image 0 loaded, Size: " << env.left.cols << "x" << env.left.rows;
    if (env.pre_right) env.right = *env.pre_right;
    else if (!read_png_gray(path(env, "undistorted/00000001.png"), env.right, &err)) { LOGE << "unable to load input images"; LOGE << err; return false; }
    env.right_index = 1;
    LOGI << "image 1 loaded, Size: " << env.right.cols << "x" << env.right.rows;
    if (env.left.cols != env.right.cols || env.left.rows != env.right.rows) { LOGE << "left and right images differ in size"; return false; }
    const double sis = cfg.getd("SAVE_INPUT_SCALE");
    if (sis < 1.0) {
        const size_t nw = (size_t)(env.left.cols * sis), nh = (size_t)(env.left.rows * sis);
        const double scale = (double)nw / (double)env.left.cols;
        LOGI << "original size: " << env.left.cols << "x" << env.left.rows;
        LOGI << "  scaled size: " << nw << "x" << nh;
        LOGI << "        scale: " << scale;
        if (nw > 0 && nh > 0) {      // (resize + PNG deflate: ~20 ms per image, off the main thread)
            g_jpeg.run([fn = path(env, "00000000_s.png"), img = env.left, nw, nh] { write_png_gray(fn, resize_cubic(img, (int)nw, (int)nh)); });
            g_jpeg.run([fn = path(env, "00000001_s.png"), img = env.right, nw, nh] { write_png_gray(fn, resize_cubic(img, (int)nw, (int)nh)); });
        }
        Mat k0 = env.intr_left, k1 = env.intr_right;
        for (auto& x : k0.v) x *= scale;
        for (auto& x : k1.v) x *= scale;
        k0.at(2, 2) = 1; k1.at(2, 2) = 1;
        save_matrix_txt(path(env, "K0_small.txt"), k0);
        save_matrix_txt(path(env, "K1_small.txt"), k1);
        std::ofstream ofs(path(env, "scale.txt").c_str());
        ofs.precision(16); ofs << std::scientific << scale; ofs.close();
    }
    return true;
}

bool rectify(Env& env, const Config& cfg, wsg_handle* h)   // wass_stereo.cpp:447-613
{
    LOG_SCOPE("rectify");
    LOGI << "rectifying...";
    bool auto_swap = true, do_swap = false;
    if (fabs(env.T.at(1, 0)) > fabs(env.T.at(0, 0))) { LOGE << "Vertical stereo not supported"; return false; }
    LOGI << "Detected stereo setup:";
    if (env.T.at(0, 0) > 0) LOGI << "CAM1 (L) ---------  CAM0 (R)"; else LOGI << "CAM0 (L) ---------  CAM1 (R)";
    if (cfg.getb("DISABLE_AUTO_LEFT_RIGHT")) {
        auto_swap = false;
        do_swap = cfg.getb("SWAP_LEFT_RIGHT");
        LOGI << "auto left-right detection disabled. Swap left-right? " << (do_swap ? "YES" : "NO");
        if (do_swap) { LOGI << "swapping left-right images as requested"; env.swapLeftRight(); }
    } else if (env.T.at(0, 0) < 0) {
        LOGI << "auto-swapping left-right images" << "\n";
        env.swapLeftRight();
    }
    const int W = env.left.cols, H = env.left.rows;
    auto crop = [&](const Image8& src, const int* r, Image8& dst) {
        dst.rows = r[3]; dst.cols = r[2]; dst.px.resize((size_t)r[2] * r[3]);
        for (int y = 0; y < r[3]; ++y) memcpy(&dst.px[(size_t)y * r[2]], &src.px[(size_t)(r[1] + y) * src.cols + r[0]], r[2]);
    };
    if (cfg.getb("USE_CUSTOM_STEREORECTIFY")) {          // wass_stereo.cpp:496-529
        const double baseline_rot = cfg.getd("RECTIFY_ANGLE");
        LOGI << "Using WASS custom stereorectify, baseline angle delta=" << baseline_rot;
        int roi[4];
        double best = 0;
        if (baseline_rot == 0) std::cout << "Optimizing best rectifying plane... ";
        if (wsg_stereo_rectify_custom(env.intr_left.v.data(), env.intr_right.v.data(), env.Rinv.v.data(), env.Tinv.v.data(), baseline_rot,
                                      W, H, env.HL, env.HR, roi, &best) != WSG_OK) { LOGE << "custom stereorectify failed"; return false; }
        if (baseline_rot == 0) std::cout << "DONE" << std::endl << "Best angle: " << best << " deg." << std::endl;
        auto inv3 = [](const double* S, double* D) {      // cv::Matx33d::inv()
            const double d = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
            const double id = 1. / d;
            D[0] = (S[4] * S[8] - S[5] * S[7]) * id; D[1] = (S[2] * S[7] - S[1] * S[8]) * id; D[2] = (S[1] * S[5] - S[2] * S[4]) * id;
            D[3] = (S[5] * S[6] - S[3] * S[8]) * id; D[4] = (S[0] * S[8] - S[2] * S[6]) * id; D[5] = (S[2] * S[3] - S[0] * S[5]) * id;
            D[6] = (S[3] * S[7] - S[4] * S[6]) * id; D[7] = (S[1] * S[6] - S[0] * S[7]) * id; D[8] = (S[0] * S[4] - S[1] * S[3]) * id;
        };
        inv3(env.HL, env.HLi); inv3(env.HR, env.HRi);
        env.custom_rectify = true;
        Mat hl, hr; hl.rows = hl.cols = hr.rows = hr.cols = 3; hl.v.assign(env.HL, env.HL + 9); hr.v.assign(env.HR, env.HR + 9);
        save_matrix_txt(path(env, env.left_index == 0 ? "H0_rect.txt" : "H1_rect.txt"), hl);
        save_matrix_txt(path(env, env.left_index == 0 ? "H1_rect.txt" : "H0_rect.txt"), hr);
        env.left_rect.rows = env.right_rect.rows = H; env.left_rect.cols = env.right_rect.cols = W;
        env.left_rect.px.resize((size_t)W * H); env.right_rect.px.resize((size_t)W * H);
        if (wsg_warp_perspective(h, env.left.px.data(), H, W, W, env.HL, env.left_rect.px.data()) != WSG_OK ||
            wsg_warp_perspective(h, env.right.px.data(), H, W, W, env.HR, env.right_rect.px.data()) != WSG_OK) {
            LOGE << "warpPerspective failed: " << wsg_last_error(h); return false;
        }
        if (cfg.getb("DISABLE_RECTIFY_ROI")) { roi[0] = 0; roi[1] = 0; roi[2] = W; roi[3] = H; }
        if (roi[0] < 0 || roi[1] < 0 || roi[2] <= 0 || roi[3] <= 0 || roi[0] + roi[2] > W || roi[1] + roi[3] > H) {
            LOGE << "rectification ROI outside the image"; return false;      // (cv::Mat::operator() throws in the reference)
        }
        memcpy(env.roi_left, roi, 16); memcpy(env.roi_right, roi, 16);
        crop(env.left_rect, env.roi_left, env.left_crop);
        crop(env.right_rect, env.roi_right, env.right_crop);
        return true;
    }
    LOGI << "Rectifying via cv::stereoRectify";
    int roi_l[4], roi_r[4];
    bool ok = false;
    int guard = 0;
    do {
        if (wsg_stereo_rectify(env.intr_left.v.data(), env.intr_right.v.data(), env.R.v.data(), env.T.v.data(), W, H, env.R1, env.R2,
                               env.P1r, env.P2r, roi_l, roi_r) != WSG_OK) { LOGE << "stereoRectify failed"; return false; }
        if (fabs(env.P2r[3]) < fabs(env.P2r[7])) { LOGE << "vertical stereo not supported"; return false; }
        if (roi_l[2] == 0 || roi_r[2] == 0 || roi_l[3] == 0 || roi_r[3] == 0) { LOGE << "the epipole lies inside the image plane"; return false; }
        if (auto_swap) {
            if (env.P2r[3] < 0) { LOGI << "auto-swapping left-right images" << "\n"; env.swapLeftRight(); }
            else ok = true;
        } else if (do_swap) {
            LOGI << "swapping left-right images as requested"; env.swapLeftRight(); do_swap = false;
        } else ok = true;
    } while (!ok && ++guard < 4);
    if (!ok) { LOGE << "rectification did not converge"; return false; }
    const int ymin = std::max(roi_l[1], roi_r[1]);
    const int ymax = std::min(roi_l[1] + roi_l[3], roi_r[1] + roi_r[3]);
    env.roi_left[0] = roi_l[0]; env.roi_left[1] = ymin; env.roi_left[2] = roi_l[2]; env.roi_left[3] = ymax - ymin;
    env.roi_right[0] = roi_r[0]; env.roi_right[1] = ymin; env.roi_right[2] = roi_r[2]; env.roi_right[3] = ymax - ymin;
    if (env.roi_left[2] > env.roi_right[2]) env.roi_left[2] = env.roi_right[2]; else env.roi_right[2] = env.roi_left[2];
    if (env.roi_left[3] <= 0 || env.roi_left[2] <= 0) { LOGE << "empty rectification ROI"; return false; }
    env.left_rect.rows = env.right_rect.rows = H; env.left_rect.cols = env.right_rect.cols = W;
    env.left_rect.px.resize((size_t)W * H); env.right_rect.px.resize((size_t)W * H);
    if (wsg_rectify_image(h, env.left.px.data(), H, W, W, env.intr_left.v.data(), env.R1, env.P1r, env.left_rect.px.data()) != WSG_OK ||
        wsg_rectify_image(h, env.right.px.data(), H, W, W, env.intr_right.v.data(), env.R2, env.P2r, env.right_rect.px.data()) != WSG_OK) {
        LOGE << "remap failed: " << wsg_last_error(h); return false;
    }
    crop(env.left_rect, env.roi_left, env.left_crop);
    crop(env.right_rect, env.roi_right, env.right_crop);
    LOGI << "rectification map generated. Size: " << env.left_crop.cols << "x" << env.left_crop.rows;
    return true;
}

void save_poses(const Env& env)   // wass_stereo.cpp:1888-1894, 1902-1908
{
    save_matrix_txt(path(env, "P0cam.txt"), env.P0);
    save_matrix_txt(path(env, "P1cam.txt"), env.P1);
    save_matrix_txt(path(env, "Cam0_poseR.txt"), env.Rpose0);
    save_matrix_txt(path(env, "Cam0_poseT.txt"), env.Tpose0);
    save_matrix_txt(path(env, "Cam1_poseR.txt"), env.Rpose1);
    save_matrix_txt(path(env, "Cam1_poseT.txt"), env.Tpose1);
}

bool save_ply(wsg_handle* h, const std::string& filename)   // PovMesh::save_as_ply_points, PovMesh.cpp:463-517
{
    int w = 0, hh = 0; unsigned long long n = 0;
    if (wsg_mesh_size(h, &w, &hh, &n) != WSG_OK) return false;
    const size_t np = (size_t)w * hh;
    std::vector<uint8_t> valid(np), grey(np);
    std::vector<double> xyz(np * 3);
    if (wsg_mesh_download(h, valid.data(), xyz.data(), grey.data()) != WSG_OK) return false;
    std::ofstream ofs(filename.c_str(), std::ios::binary);
    if (ofs.fail()) return false;
    ofs << "ply\nformat binary_little_endian 1.0\nelement vertex " << n << "\nproperty float x\nproperty float y\nproperty float z\n"
        << "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n";
    std::vector<char> buf(n * 15);
    char* p = buf.data();
    for (size_t i = 0; i < np; ++i)
        if (valid[i]) {
            for (int k = 0; k < 3; ++k) { const float f = (float)xyz[3 * i + k]; memcpy(p, &f, 4); p += 4; }
            p[0] = p[1] = p[2] = (char)grey[i]; p += 3;
        }
    ofs.write(buf.data(), (std::streamsize)buf.size());
    return !ofs.fail();
}

void show_time_stats(const Timer& t)   // src/wass_stereo/render.hpp:175-191
{
    LOGI << "+----------------------------+-------------------+";
    LOGI << "|   Task                     |   Time (seconds)  |";
    LOGI << "+----------------------------+-------------------+";
    double last = 0;
    for (const auto& e : t.evts) {
        std::stringstream ss; ss << "| " << std::setw(25) << e.second << "  |" << std::setw(18) << (e.first - last) << " |";
        LOGI << ss.str();
        last = e.first;
    }
    LOGI << "+----------------------------+-------------------+";
    std::stringstream ss; ss << "| " << std::setw(25) << "TOTAL" << "  |" << std::setw(18) << t.elapsed() << " |";
    LOGI << ss.str();
    LOGI << "+----------------------------+-------------------+";
}

// ---- diagnostic images (the reference writes them unconditionally through cv::imwrite; content is informative only) ------

struct Rgb { int rows = 0, cols = 0; std::vector<uint8_t> px; };
Rgb gray2rgb(const Image8& g)
{
    Rgb o; o.rows = g.rows; o.cols = g.cols; o.px.resize((size_t)g.rows * g.cols * 3);
    for (size_t i = 0; i < g.px.size(); ++i) o.px[3 * i] = o.px[3 * i + 1] = o.px[3 * i + 2] = g.px[i];
    return o;
}
void rect_red(Rgb& im, int x0, int y0, int w, int h, int xoff = 0, int t = 3)   // cv::rectangle(..., CV_RGB(255,0,0), 3)
{
    auto put = [&](int x, int y) {
        if (x < 0 || y < 0 || x >= im.cols || y >= im.rows) return;
        uint8_t* p = &im.px[((size_t)y * im.cols + x) * 3]; p[0] = 255; p[1] = 0; p[2] = 0;
    };
    for (int k = -(t / 2); k <= t / 2; ++k) {
        for (int x = x0; x < x0 + w; ++x) { put(x + xoff, y0 + k); put(x + xoff, y0 + h - 1 + k); }
        for (int y = y0; y < y0 + h; ++y) { put(x0 + k + xoff, y); put(x0 + w - 1 + k + xoff, y); }
    }
}
Rgb half_size(const Rgb& s)      // cv::resize(.., 0.5, 0.5, INTER_LINEAR) up to rounding
{
    auto half_len = [](int n) { const int k = n / 2; return (n & 1) ? k + (k & 1) : k; };    // cvRound(n * 0.5): halves go to the even neighbour
    Rgb o; o.rows = std::max(half_len(s.rows), 1); o.cols = std::max(half_len(s.cols), 1); o.px.resize((size_t)o.rows * o.cols * 3);
    for (int y = 0; y < o.rows; ++y)
        for (int x = 0; x < o.cols; ++x)
            for (int c = 0; c < 3; ++c) {
                const int y0 = std::min(2 * y, s.rows - 1), y1 = std::min(2 * y + 1, s.rows - 1);
                const int x0 = std::min(2 * x, s.cols - 1), x1 = std::min(2 * x + 1, s.cols - 1);
                const size_t r0 = (size_t)y0 * s.cols, r1 = (size_t)y1 * s.cols;
                o.px[((size_t)y * o.cols + x) * 3 + c] = (uint8_t)((s.px[(r0 + x0) * 3 + c] + s.px[(r0 + x1) * 3 + c] +
                                                                    s.px[(r1 + x0) * 3 + c] + s.px[(r1 + x1) * 3 + c] + 2) >> 2);
            }
    return o;
}
void save_stereo_jpg(const Env& env)           // wass_stereo.cpp:1911-1926
{
    const int H = env.left_rect.rows, W = env.left_rect.cols;
    Rgb o; o.rows = H; o.cols = 2 * W; o.px.resize((size_t)H * 2 * W * 3);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < 2 * W; ++x) {
            const uint8_t g = x < W ? env.left_rect.px[(size_t)y * W + x] : env.right_rect.px[(size_t)y * W + x - W];
            uint8_t* p = &o.px[((size_t)y * 2 * W + x) * 3]; p[0] = p[1] = p[2] = g;
        }
    rect_red(o, env.roi_left[0], env.roi_left[1], env.roi_left[2], env.roi_left[3]);
    rect_red(o, env.roi_right[0], env.roi_right[1], env.roi_right[2], env.roi_right[3], W);
    for (int y = 0; y < H; y += 20)
        for (int x = 0; x < 2 * W; ++x) { uint8_t* p = &o.px[((size_t)y * 2 * W + x) * 3]; p[0] = 255; p[1] = 0; p[2] = 0; }
    write_jpeg(path(env, "stereo.jpg"), o.px.data(), o.rows, o.cols, 3);
}
void save_disparity_float_jpg(const std::string& fn, const float* d, int rows, int cols)   // render.hpp:97-136
{
    float mn = (float)(cols + 1), mx = 0.f;
    for (size_t i = 0; i < (size_t)rows * cols; ++i) { mn = std::min(mn, d[i]); mx = std::max(mx, d[i]); }
    std::vector<uint8_t> g((size_t)rows * cols);
    const float sc = mx > mn ? 255.f / (mx - mn) : 0.f;
    for (size_t i = 0; i < g.size(); ++i) g[i] = (uint8_t)((d[i] - mn) * sc);
    write_jpeg(fn, g.data(), rows, cols, 1);
}
// stereo_input.jpg, disparity_stereo_ouput.jpg, disparity_final_scaled.jpg, disparity_coverage.jpg (wass_stereo.cpp:833, 854,
// 1001-1017) from the final float disparity of the ROI.  (The reference renders disparity_stereo_ouput.jpg before its
// dilate / erode passes; that intermediate never leaves the device here, so both renderings show the final map.)
void save_dense_jpgs(const Env& env, const wsg_dense_params& dp, const float* disp_roi)
{
    const int rh = env.right_crop.rows, rw = env.right_crop.cols;
    if (dp.DENSE_SCALE == 1.0) {
        const int N = dp.MAX_DISPARITY, off = std::max(dp.DISPARITY_OFFSET, 0), comp = std::max(-dp.DISPARITY_OFFSET, 0), wp = rw + N + off;
        std::vector<uint8_t> g((size_t)2 * rh * wp, 0);
        for (int y = 0; y < rh; ++y) {
            memcpy(&g[(size_t)y * wp + N + off - comp], &env.left_crop.px[(size_t)y * rw], rw);
            memcpy(&g[(size_t)(rh + y) * wp + N], &env.right_crop.px[(size_t)y * rw], rw);
        }
        write_jpeg(path(env, "stereo_input.jpg"), g.data(), 2 * rh, wp, 1);
    }
    save_disparity_float_jpg(path(env, "disparity_stereo_ouput.jpg"), disp_roi, rh, rw);
    const int H = env.right_rect.rows, W = env.right_rect.cols;
    std::vector<float> full((size_t)H * W, 0.f);
    for (int y = 0; y < rh; ++y) memcpy(&full[(size_t)(env.roi_right[1] + y) * W + env.roi_right[0]], disp_roi + (size_t)y * rw, (size_t)rw * 4);
    save_disparity_float_jpg(path(env, "disparity_final_scaled.jpg"), full.data(), H, W);
    Rgb cov = gray2rgb(env.right_rect);
    for (size_t i = 0; i < full.size(); ++i) if (full[i] > 1.f) cov.px[3 * i + 1] = 100;
    rect_red(cov, env.roi_right[0], env.roi_right[1], env.roi_right[2], env.roi_right[3]);
    Rgb h2 = half_size(cov);
    write_jpeg(path(env, "disparity_coverage.jpg"), h2.px.data(), h2.rows, h2.cols, 3);
}
// graph_components.jpg (PovMesh.cpp:222-252, 982-984): the biggest component in green, at half size.  The reference gives
// every other component its own palette colour; the component labels stay on the device here, so all of them are drawn red.
void save_components_jpg(const Env& env, const std::vector<uint8_t>& before, const std::vector<uint8_t>& after, int w, int h)
{
    Rgb im; im.rows = h; im.cols = w; im.px.assign((size_t)w * h * 3, 0);
    for (size_t i = 0; i < before.size(); ++i)
        if (after[i]) im.px[3 * i + 1] = 255; else if (before[i]) im.px[3 * i] = 255;
    Rgb h2 = half_size(im);
    write_jpeg(path(env, "graph_components.jpg"), h2.px.data(), h2.rows, h2.cols, 3);
}

#define WSG_CHECK(call) do { if ((call) != WSG_OK) throw std::runtime_error(std::string(#call) + ": " + wsg_last_error(h)); } while (0)

}  // namespace

// wsg_dense_params from the configuration (wass_stereo.cpp:742-761, 772-782)
static wsg_dense_params make_dense_params(const Config& cfg)
{
    wsg_dense_params dp;
    wsg_dense_params_default(&dp);
    dp.MIN_DISPARITY = cfg.geti("MIN_DISPARITY"); dp.MAX_DISPARITY = cfg.geti("MAX_DISPARITY"); dp.WINSIZE = cfg.geti("WINSIZE");
    dp.DENSE_SCALE = cfg.getd("DENSE_SCALE"); dp.DISPARITY_OFFSET = cfg.geti("DISPARITY_OFFSET");
    dp.DISP_DILATE_STEPS = cfg.geti("DISP_DILATE_STEPS"); dp.DISP_EROSION_STEPS = cfg.geti("DISP_EROSION_STEPS");
    dp.DENSE_P1_MULT = cfg.geti("DENSE_P1_MULT"); dp.DENSE_P2_MULT = cfg.geti("DENSE_P2_MULT");
    dp.DENSE_UNIQUENESS_RATIO = cfg.geti("DENSE_UNIQUENESS_RATIO"); dp.DENSE_DISP12MAXDIFF = cfg.geti("DENSE_DISP12MAXDIFF");
    dp.DENSE_PREFILTER_CAP = cfg.geti("DENSE_PREFILTER_CAP"); dp.DENSE_SPECKLE_RANGE = cfg.geti("DENSE_SPECKLE_RANGE");
    dp.DENSE_SPECKLE_WINDOW_SIZE = cfg.geti("DENSE_SPECKLE_WINDOW_SIZE");
    dp.mode = cfg.getb("SGM_FULL_8PATH") ? WSG_MODE_HH : WSG_MODE_SGBM;
    dp.MEDIAN_FILTER_WSIZE = cfg.geti("MEDIAN_FILTER_WSIZE");
    dp.DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD = cfg.geti("DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD");
    return dp;
}

// Everything up to the dense matcher: load_data, poses, rectify, poses (wass_stereo.cpp:1877-1926).
// Returns 0, or -1 on failure.
static int stage_load_rectify(Env& env, const Config& cfg, wsg_handle* h)
{
    LOG_SCOPE("wass_stereo");
    LOGI << "Reconstructing \"" << env.workdir << "\"";
    env.timer.start();
    env.cam_distance = 1.0;
    { HostProf hp("load_data"); if (!load_data(env, cfg)) return -1; }
    env.timer.mark("Data load");
    std::cout << "[P|10|100]" << std::endl;
    save_poses(env);
    { HostProf hp("rectify"); if (!rectify(env, cfg, h)) return -1; }   // the reference ignores this return value and runs on undefined state
    env.timer.mark("Rectification");
    std::cout << "[P|20|100]" << std::endl;
    LOG_SCOPE("wass_stereo");
    save_poses(env);
    if (cfg.getb("SAVE_DEBUG_IMAGES")) g_jpeg.run([&env] { save_stereo_jpg(env); });      // (reads env: joined before env goes away)
    return 0;
}

static void log_dense_begin(Env& env, const wsg_dense_params& dp)
{
    LOG_SCOPE("sgbm_dense_stereo");
    if (dp.MEDIAN_FILTER_WSIZE >= 3) LOGI << "applying median filter (window size " << dp.MEDIAN_FILTER_WSIZE << " px.)";
    if (dp.DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD > 0) LOGI << "extracting the biggest connected component from the disparity map";
    LOGI << "Disparity offset: " << dp.DISPARITY_OFFSET << " px";
    env.disparity_compensation = dp.DISPARITY_OFFSET > 0 ? 0 : -dp.DISPARITY_OFFSET;
    LOGI << "computing dense disparity map... (may take a while)";
}

static void log_dense_end(Env& env, wsg_handle* h)
{
    LOG_SCOPE("sgbm_dense_stereo");
    wsg_sgbm_stats st;
    if (wsg_sgbm_get_stats(h, &st) == WSG_OK && st.out_of_domain)
        LOGE << "matching cost range exceeds int16 headroom (max cost " << st.max_cost << "): results may deviate from OpenCV";
    LOGI << "dense stereo completed successfully";
    env.timer.mark("Dense Stereo");
    std::cout << "[P|40|100]" << std::endl;
}

// Everything after the dense matcher, on the disparity the handle holds: triangulation, outlier removal, plane, export
// (wass_stereo.cpp:1981-2141).  plane_out: the four numbers of plane.txt (NaN on a soft RANSAC failure).  Returns 0 / -1.
static int stage_after_dense(Env& env, const Config& cfg, wsg_handle* h, const wsg_dense_params& dp, double plane_out[4],
                             const float* disp_roi = nullptr)
{
    for (int i = 0; i < 4; ++i) plane_out[i] = std::nan("");
    const bool dbg_images = cfg.getb("SAVE_DEBUG_IMAGES");
    if (dbg_images && disp_roi) {
        const size_t nroi = (size_t)env.right_crop.rows * env.right_crop.cols;
        g_jpeg.run([&env, dp, d = std::vector<float>(disp_roi, disp_roi + nroi)] { save_dense_jpgs(env, dp, d.data()); });
    }
    HostProf hp_all("after dense (total)");
    try {
        // ---- triangulation (wass_stereo.cpp:1039-1386)
        LOG_SCOPE("triangulate");
        wsg_calib cal;
        memset(&cal, 0, sizeof cal);
        cal.use_homographies = env.custom_rectify ? 1 : 0;
        memcpy(cal.HLi, env.HLi, 72); memcpy(cal.HRi, env.HRi, 72);
        memcpy(cal.K0, env.intr_left.v.data(), 72); memcpy(cal.K1, env.intr_right.v.data(), 72);
        memcpy(cal.R, env.R.v.data(), 72); memcpy(cal.T, env.T.v.data(), 24);
        memcpy(cal.R1, env.R1, 72); memcpy(cal.R2, env.R2, 72); memcpy(cal.P1, env.P1r, 96); memcpy(cal.P2, env.P2r, 96);
        memcpy(cal.roi_left, env.roi_left, 16); memcpy(cal.roi_right, env.roi_right, 16);
        cal.left_cols = env.left.cols; cal.left_rows = env.left.rows; cal.right_cols = env.right.cols; cal.right_rows = env.right.rows;
        cal.rect_cols = env.right_rect.cols; cal.rect_rows = env.right_rect.rows;
        wsg_tri_params tp;
        wsg_tri_params_default(&tp);
        tp.TRIANG_MIN_ANGLE = cfg.getd("TRIANG_MIN_ANGLE");
        tp.TRIANG_BBOX_TOP = cfg.getd("TRIANG_BBOX_TOP"); tp.TRIANG_BBOX_LEFT = cfg.getd("TRIANG_BBOX_LEFT");
        tp.TRIANG_BBOX_RIGHT = cfg.getd("TRIANG_BBOX_RIGHT"); tp.TRIANG_BBOX_BOTTOM = cfg.getd("TRIANG_BBOX_BOTTOM");
        tp.DISCARD_BURNED_AREAS = cfg.getb("DISCARD_BURNED_AREAS"); tp.disparity_compensation = env.disparity_compensation;
        tp.DENSE_SCALE = dp.DENSE_SCALE; tp.cam_distance = env.cam_distance;
        auto load_mask = [&](const char* key, const Image8& ref, std::vector<uint8_t>& m) -> const uint8_t* {
            if (cfg.gets(key) == "none") return nullptr;
            const std::string fn = env.workdir + "/" + cfg.gets(key);
            LOGI << "Loading " << fn << " as " << (std::string(key) == "LEFT_MASK_IMAGE" ? "left" : "right") << " camera mask";
            Image8 im; std::string err;
            if (!read_png_gray(fn, im, &err) || im.cols != ref.cols || im.rows != ref.rows) { LOGE << "not found or invalid image."; return nullptr; }
            m.resize(im.px.size());
            for (size_t i = 0; i < m.size(); ++i) m[i] = im.px[i] > 0 ? 1 : 0;   // cv::threshold(aux, mask, 0.5, 1, THRESH_BINARY)
            return m.data();
        };
        std::vector<uint8_t> lm, rm;
        const uint8_t* lmp = load_mask("LEFT_MASK_IMAGE", env.left, lm);
        const uint8_t* rmp = load_mask("RIGHT_MASK_IMAGE", env.right, rm);
        LOGI << "triangulating disparity map";
        unsigned long long n_pts = 0;
        { HostProf hp("triangulate"); WSG_CHECK(wsg_triangulate_from_dense(h, env.left.px.data(), env.right.px.data(), lmp, rmp, &cal, &tp, &n_pts)); }
        LOGI << "... 100%";
        LOGI << n_pts << " valid points found";
        env.timer.mark("Triangulation");
        std::cout << "[P|60|100]" << std::endl;
        LOG_SCOPE("wass_stereo");
        if ((long long)n_pts < cfg.geti("MIN_TRIANGULATED_POINTS")) { LOGE << "Too few points triangulated, aborting"; return -1; }

        // ---- outlier removal (wass_stereo.cpp:2046-2060)
        double zgap = 0;
        { HostProf hp("zgap"); WSG_CHECK(wsg_mesh_zgap_percentile(h, cfg.getd("ZGAP_PERCENTILE"), &zgap)); }
        env.timer.mark("Z-gap stats");
        unsigned long long nleft = 0;
        LOG_SCOPE("cluster");
        LOGI << "extracting connected-components";
        std::vector<uint8_t> valid_before, valid_after;
        int cw = 0, ch = 0;
        if (dbg_images) {
            WSG_CHECK(wsg_mesh_size(h, &cw, &ch, nullptr));
            valid_before.resize((size_t)cw * ch);
            WSG_CHECK(wsg_mesh_download(h, valid_before.data(), nullptr, nullptr));
        }
        { HostProf hp("component"); WSG_CHECK(wsg_mesh_biggest_component(h, zgap, &nleft)); }
        if (dbg_images) {
            valid_after.resize(valid_before.size());
            WSG_CHECK(wsg_mesh_download(h, valid_after.data(), nullptr, nullptr));
            g_jpeg.run([&env, b = std::move(valid_before), a2 = std::move(valid_after), cw, ch] { save_components_jpg(env, b, a2, cw, ch); });
        }
        LOGI << "biggest component size: " << nleft << " (px)";
        env.timer.mark("Outlier removal");
        std::cout << "[P|80|100]" << std::endl;
        LOG_SCOPE("wass_stereo");
        if (cfg.getb("SAVE_FULL_MESH") && !save_ply(h, path(env, "mesh_full.ply"))) { LOGE << "unable to save mesh data."; return -1; }

        // ---- plane (wass_stereo.cpp:2062-2107)
        LOGI << "estimating best fitting plane...";
        const int rounds = cfg.geti("PLANE_RANSAC_ROUNDS");
        std::vector<int32_t> triples((size_t)std::max(rounds, 1) * 6);
        int mw = 0, mh = 0;
        WSG_CHECK(wsg_mesh_size(h, &mw, &mh, nullptr));
        wsg_ransac_draw(mw, mh, rounds, triples.data());
        double plane[4] = {0, 0, 0, 0};
        int ok = 0; unsigned long long best = 0;
        { HostProf hp("ransac"); if (rounds > 0) WSG_CHECK(wsg_mesh_ransac_plane(h, triples.data(), rounds, cfg.getd("PLANE_RANSAC_THRESHOLD"), plane, &ok, &best)); }
        LOG_SCOPE("ransac_find_plane");
        LOGI << rounds << " ransac rounds, " << best << " best inliers";
        LOGI << "ransac plane coeffs: " << plane[0] << " " << plane[1] << " " << plane[2] << " " << plane[3];
        LOG_SCOPE("wass_stereo");
        if (ok) {
            env.timer.mark("Plane fitting");
            std::cout << "[P|90|100]" << std::endl;
            LOGI << "refining plane";
            unsigned long long k = 0;
            WSG_CHECK(wsg_mesh_crop_plane(h, plane, cfg.getd("PLANE_RANSAC_THRESHOLD"), &k));
            LOGI << "number of points after plane cropping: " << k;
            wsg_refine_params rp;
            wsg_refine_params_default(&rp);
            rp.PLANE_REFINE_XMIN = cfg.getd("PLANE_REFINE_XMIN"); rp.PLANE_REFINE_XMAX = cfg.getd("PLANE_REFINE_XMAX");
            rp.PLANE_REFINE_YMIN = cfg.getd("PLANE_REFINE_YMIN"); rp.PLANE_REFINE_YMAX = cfg.getd("PLANE_REFINE_YMAX");
            rp.PLANE_REFINEMENT_MAX_DISTANCE = cfg.getd("PLANE_REFINEMENT_MAX_DISTANCE");
            rp.PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE = cfg.getb("PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE");
            rp.PLANE_USE_CENTRAL_THIRD_ONLY = cfg.getb("PLANE_USE_CENTRAL_THIRD_ONLY");
            {   // plane_refinement_inliers.xyz: every 10th refinement inlier, in grid scan order (wass_stereo.cpp:2077-2085),
                // picked on the device.  Same bytes as `ofs << x << " " << y << " " << z << std::endl` (%g is the stream's
                // default format) without one flush per line: ~46 000 lines at the benchmark size.
                std::vector<double> pts(((size_t)mw * mh + 9) / 10 * 3);
                unsigned long long npts = 0;
                { HostProf hp("refine inliers"); WSG_CHECK(wsg_mesh_refine_inliers(h, &rp, 10, pts.data(), pts.size() / 3, &npts, nullptr)); }
                // (glibc spends ~0.7 us per %g: 0.1 s per frame, more than the whole device side -- formatted and written on
                // a worker thread, joined with the JPEG encoders before the workdir is reported done)
                g_jpeg.run([fn = path(env, "plane_refinement_inliers.xyz"), pts = std::move(pts), npts] {
                    std::string txt;
                    txt.reserve((size_t)npts * 40);
                    char line[128];
                    for (size_t i = 0; i < (size_t)npts; ++i)
                        txt.append(line, (size_t)snprintf(line, sizeof line, "%g %g %g\n", pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
                    write_file(fn, txt.data(), txt.size());
                });
            }
            unsigned long long nin = 0;
            { HostProf hp("refine"); WSG_CHECK(wsg_mesh_refine_plane(h, &rp, plane, &nin)); }
            LOG_SCOPE("refine_plane");
            LOGI << "refinement inliers (after cropping): " << nin;
            LOGI << "estimated plane coeffs: " << plane[0] << " " << plane[1] << " " << plane[2] << " " << plane[3];
            LOG_SCOPE("wass_stereo");
            WSG_CHECK(wsg_mesh_crop_plane(h, plane, cfg.getd("PLANE_MAX_DISTANCE"), &k));
            LOGI << "number of points after plane cropping: " << k;
            env.timer.mark("Plane refinement");
            std::ofstream ofs(path(env, "plane.txt").c_str());
            ofs << std::setprecision(20);
            for (int i = 0; i < 4; ++i) ofs << plane[i] << std::endl;
        } else {
            LOGE << "ransac failed. I'll continue anyway but plane data won't be available!";
            std::ofstream ofs(path(env, "plane.txt").c_str());
            ofs << "nan nan nan nan" << std::endl;
        }

        // ---- export (wass_stereo.cpp:2110-2135)
        LOGI << "Exporting point cloud data";
        if (cfg.getb("SAVE_AS_PLY") && !save_ply(h, path(env, "mesh.ply"))) { LOGE << "unable to save mesh data."; return -1; }
        {
            int w2 = 0, h2 = 0; unsigned long long nv = 0;
            WSG_CHECK(wsg_mesh_size(h, &w2, &h2, &nv));
            HostProf hp_exp("export (total)");
            static std::vector<char> buf;            // reused across the workdirs of a batch (zero-filling 56 MB per frame was 20 ms)
            if (buf.size() < 256 + (size_t)nv * 12) buf.resize(256 + (size_t)nv * 12);
            size_t nb = 0;
            if (cfg.getb("SAVE_COMPRESSED")) {
                LOG_SCOPE("save_as_xyz_compressed");
                LOGI << "saving mesh as compressed xyz file...";
                { HostProf hp("export xyzC (device + copy)"); WSG_CHECK(wsg_mesh_export_xyzc(h, plane, buf.data(), buf.size(), &nb)); }
                { HostProf hp("write mesh_cam.xyzC"); if (!write_file(path(env, "mesh_cam.xyzC"), buf.data(), nb)) { LOGE << "unable to save mesh data"; return -1; } }
                LOGI << "total data size: " << (double)nb / 1e6 << " MB";
            } else {
                WSG_CHECK(wsg_mesh_export_xyzbin(h, buf.data(), buf.size(), &nb));
                if (!write_file(path(env, "mesh_cam.xyzbin"), buf.data(), nb)) { LOGE << "unable to save mesh data"; return -1; }
            }
        }
        if (ok) for (int i = 0; i < 4; ++i) plane_out[i] = plane[i];
        LOG_SCOPE("wass_stereo");
        if (!g_batch_mode) g_jpeg.wait();          // (batch mode: joined once per batch, the encoders run beside the next frames)
        env.timer.stop();
        std::cout << "[P|100|100]" << std::endl;
        show_time_stats(env.timer);
        LOGI << "All done.";
    } catch (std::runtime_error& e) {
        LOGE << e.what();
        return -1;
    }
    return 0;
}

static void print_usage()
{
    std::cout << "Usage:" << std::endl;
    std::cout << "wass_stereo [--genconfig] <config_file> <workdir> [--measure] [--rectify-only]" << std::endl << std::endl;
}

// ---- batch mode (an extension; the reference runs one process per frame, cli/wasscli/wasscli.py:326-346) ------------------
//   wass_stereo --batch [--batch-size B] [--planes-out FILE] [--ranks N --rank R --nccl-id-file F]
//               <config_file> (<workdir>... | --workdirs-from FILE)
// One process, one warm device arena: B frames at a time go through ONE batched matcher run (wsg_dense_stereo_batch), the
// PNGs of the next B frames are decoded by a second thread meanwhile.  Every workdir gets exactly the files and the log
// of a single-frame run.  With --ranks N this process takes the workdirs i = R (mod N); the mean plane over ALL ranks'
// frames comes from one NCCL all-reduce (wsg_plane_allreduce), the per-frame planes from an all-gather, and rank 0 writes
// them in frame order to --planes-out (the reference's planes.txt is in completion order, wasscli.py:343).
struct Prefetched { Image8 img0, img1; bool ok0 = false, ok1 = false; };

static int run_batch(int argc, char* argv[])
{
    g_batch_mode = true;
    int B = 8, ranks = 1, rank = 0;
    std::string planes_out, id_file, cfg_file, list_file;
    std::vector<std::string> all;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) throw std::runtime_error("missing value after " + a); return argv[++i]; };
        try {
            if (a == "--batch-size") B = std::max(1, atoi(next().c_str()));
            else if (a == "--planes-out") planes_out = next();
            else if (a == "--ranks") ranks = std::max(1, atoi(next().c_str()));
            else if (a == "--rank") rank = atoi(next().c_str());
            else if (a == "--nccl-id-file") id_file = next();
            else if (a == "--workdirs-from") list_file = next();
            else if (cfg_file.empty()) cfg_file = a;
            else all.push_back(a);
        } catch (std::runtime_error& e) { std::cerr << e.what() << std::endl; return -1; }
    }
    if (!list_file.empty()) {
        std::ifstream f(list_file.c_str());
        if (!f.is_open()) { std::cerr << "cannot open " << list_file << std::endl; return -1; }
        for (std::string l; std::getline(f, l);) { while (!l.empty() && (l.back() == '\r' || l.back() == ' ')) l.pop_back(); if (!l.empty()) all.push_back(l); }
    }
    if (cfg_file.empty() || all.empty() || rank < 0 || rank >= ranks || (ranks > 1 && id_file.empty())) {
        std::cerr << "Invalid arguments (batch mode: --batch [--batch-size B] [--planes-out F] [--ranks N --rank R --nccl-id-file F] "
                     "<config_file> <workdir>... | --workdirs-from FILE)" << std::endl;
        return -1;
    }
    Config cfg;
    {
        std::ifstream ifs(cfg_file.c_str());
        if (!ifs.is_open()) { std::cerr << "Unable to load " << cfg_file << std::endl; return -1; }
        try { cfg.load(ifs); } catch (std::runtime_error& er) { std::cerr << er.what() << std::endl; return -1; }
    }
    int device = rank;
    if (const char* e = getenv("WASS_GPU_DEVICE")) device = atoi(e);
    wsg_handle* h = nullptr;
    if (wsg_create(device, &h) != WSG_OK) { std::cerr << "no usable CUDA device " << device << " (this build has no CPU path)" << std::endl; return -1; }
    std::vector<int> mine;
    for (int i = rank; i < (int)all.size(); i += ranks) mine.push_back(i);
    const wsg_dense_params dp = make_dense_params(cfg);
    std::vector<double> planes(mine.size() * 4, std::nan(""));
    int failed = 0;
    const double t_begin = Timer::now();

    auto prefetch = [&](size_t g0, std::vector<Prefetched>* out) {
        out->clear();
        for (size_t j = g0; j < std::min(g0 + (size_t)B, mine.size()); ++j) {
            Prefetched p; std::string err;
            p.ok0 = read_png_gray(all[mine[j]] + "/undistorted/00000000.png", p.img0, &err);
            p.ok1 = read_png_gray(all[mine[j]] + "/undistorted/00000001.png", p.img1, &err);
            out->push_back(std::move(p));
        }
    };
    std::vector<Prefetched> cur, nxt;
    prefetch(0, &cur);
    for (size_t g0 = 0; g0 < mine.size(); g0 += B) {
        const size_t g1 = std::min(g0 + (size_t)B, mine.size());
        std::thread loader;
        if (g1 < mine.size()) loader = std::thread(prefetch, g1, &nxt);
        std::vector<Env> envs(g1 - g0);
        struct JoinJobs { ~JoinJobs() { g_jpeg.wait(); } } join_before_envs_go;      // background jobs read the Envs
        std::vector<std::ofstream*> logs(g1 - g0, nullptr);
        std::vector<int> state(g1 - g0, 0);                      // 0 ready for the matcher, -1 failed
        for (size_t j = g0; j < g1; ++j) {
            Env& env = envs[j - g0];
            env.workdir = all[mine[j]];
            if (!exists(env.workdir)) { std::cerr << "\"" << env.workdir << "\" does not exists" << std::endl; state[j - g0] = -1; continue; }
            logs[j - g0] = new std::ofstream(path(env, "wass_stereo_log.txt").c_str());
            g_logfile = logs[j - g0];
            LOG_SCOPE("wass_stereo");
            LOGI << "Loading configuration file " << cfg_file;
            if (save_configuration(cfg, path(env, "stereo_config.txt")) != 0) LOGE << "Unable to save stereo configuration file";
            if (cur[j - g0].ok0 && cur[j - g0].ok1) { env.pre_left = &cur[j - g0].img0; env.pre_right = &cur[j - g0].img1; }
            { HostProf hp("load + rectify (total)"); if (stage_load_rectify(env, cfg, h) != 0) state[j - g0] = -1; }
            env.pre_left = env.pre_right = nullptr;
        }
        // frames whose crops have the same size go through the matcher together (one sequence = one size in practice)
        std::vector<char> done(g1 - g0, 0);
        for (size_t a0 = 0; a0 < g1 - g0; ++a0) {
            if (state[a0] != 0 || done[a0]) continue;
            std::vector<size_t> grp;
            for (size_t b = a0; b < g1 - g0; ++b)
                if (state[b] == 0 && !done[b] && envs[b].right_crop.rows == envs[a0].right_crop.rows && envs[b].right_crop.cols == envs[a0].right_crop.cols)
                    grp.push_back(b);
            std::vector<const uint8_t*> lc, rc;
            for (size_t b : grp) { g_logfile = logs[b]; log_dense_begin(envs[b], dp); lc.push_back(envs[b].left_crop.px.data()); rc.push_back(envs[b].right_crop.px.data()); }
            const bool dbg_images = cfg.getb("SAVE_DEBUG_IMAGES");
            std::vector<std::vector<float>> droi(dbg_images ? grp.size() : 0);
            std::vector<float*> dptr;
            for (auto& v : droi) { v.resize((size_t)envs[a0].right_crop.rows * envs[a0].right_crop.cols); dptr.push_back(v.data()); }
            int rcode;
            {
                HostProf hp_dense("dense batch");
                rcode = wsg_dense_stereo_batch(h, (int)grp.size(), lc.data(), rc.data(), envs[a0].right_crop.rows, envs[a0].right_crop.cols,
                                               envs[a0].right_crop.cols, &dp, dbg_images ? dptr.data() : nullptr);
            }
            for (size_t k = 0; k < grp.size(); ++k) {
                const size_t b = grp[k];
                done[b] = 1;
                g_logfile = logs[b];
                if (rcode != WSG_OK) { LOG_SCOPE("sgbm_dense_stereo"); LOGE << "wsg_dense_stereo_batch: " << wsg_last_error(h); state[b] = -1; continue; }
                log_dense_end(envs[b], h);
                // the per-frame seed of a single-frame run: RANDOM_SEED (or the clock) at process start, then one rand() stream
                // per process.  One process per SEQUENCE would couple the frames' draws, so every frame re-seeds.
                if (cfg.geti("RANDOM_SEED") == -1) srand((unsigned)time(0) + (unsigned)mine[g0 + b]); else srand(cfg.geti("RANDOM_SEED"));
                if (wsg_dense_select(h, (int)k) != WSG_OK ||
                    stage_after_dense(envs[b], cfg, h, dp, &planes[(g0 + b) * 4], dbg_images ? droi[k].data() : nullptr) != 0) state[b] = -1;
            }
        }
        for (size_t b = 0; b < g1 - g0; ++b) {
            if (state[b] != 0) { ++failed; std::cout << "[batch] FAILED " << envs[b].workdir << std::endl; }
            if (logs[b]) { logs[b]->close(); delete logs[b]; }
        }
        g_logfile = nullptr;
        g_jpeg.wait();
        if (loader.joinable()) loader.join();
        cur.swap(nxt);
    }
    // ---- the sequence's mean plane (np.nanmean over planes.txt, wassgridsurface.py:672-678)
    double acc[5] = {0, 0, 0, 0, 0}, mean[4];
    for (size_t j = 0; j < mine.size(); ++j) wsg_plane_mean_accumulate(acc, &planes[j * 4]);
    long long nfr = (long long)acc[4];
    std::vector<double> every(all.size() * 4, std::nan(""));
    if (ranks > 1) {
        unsigned char id[WSG_NCCL_UNIQUE_ID_BYTES];
        if (rank == 0) {
            if (wsg_nccl_unique_id(id) != WSG_OK) { std::cerr << wsg_collective_last_error() << std::endl; return -1; }
            write_file(id_file + ".tmp", (const char*)id, sizeof id);
            rename((id_file + ".tmp").c_str(), id_file.c_str());
        } else {
            for (int tries = 0;; ++tries) {
                std::ifstream f(id_file.c_str(), std::ios::binary);
                if (f.is_open() && f.read((char*)id, sizeof id)) break;
                if (tries > 6000) { std::cerr << "no NCCL id in " << id_file << std::endl; return -1; }
                usleep(10000);
            }
        }
        void* comm = nullptr;
        if (wsg_nccl_comm_create(device, ranks, rank, id, &comm) != WSG_OK) { std::cerr << wsg_collective_last_error() << std::endl; return -1; }
        if (wsg_plane_allreduce(h, comm, acc, mean, &nfr) != WSG_OK) { std::cerr << wsg_last_error(h) << std::endl; return -1; }
        const int per = (int)((all.size() + ranks - 1) / ranks);
        std::vector<double> padded((size_t)per * 4, std::nan("")), gathered((size_t)per * 4 * ranks);
        std::copy(planes.begin(), planes.end(), padded.begin());
        if (wsg_plane_allgather(h, comm, ranks, padded.data(), per, gathered.data()) != WSG_OK) { std::cerr << wsg_last_error(h) << std::endl; return -1; }
        for (int r = 0; r < ranks; ++r)
            for (int j = 0; r + j * ranks < (int)all.size(); ++j)
                for (int k = 0; k < 4; ++k) every[(size_t)(r + j * ranks) * 4 + k] = gathered[((size_t)r * per + j) * 4 + k];
        wsg_nccl_comm_destroy(comm);
        if (rank == 0) remove(id_file.c_str());
    } else {
        wsg_plane_mean_finish(acc, mean);
        every = planes;
    }
    const double dt = Timer::now() - t_begin;
    std::cout << std::setprecision(17);
    std::cout << "[batch] rank " << rank << "/" << ranks << ": " << mine.size() << " workdirs in " << std::setprecision(4) << dt << " s ("
              << (mine.empty() ? 0.0 : dt / mine.size() * 1e3) << " ms per workdir), " << failed << " failed" << std::endl;
    if (rank == 0) {
        std::cout << std::setprecision(17) << "[batch] mean plane over " << nfr << " frames: " << mean[0] << " " << mean[1] << " " << mean[2] << " " << mean[3] << std::endl;
        if (!planes_out.empty()) {
            std::ofstream ofs(planes_out.c_str());
            ofs << std::setprecision(17);
            for (size_t i = 0; i < all.size(); ++i) ofs << every[i * 4] << " " << every[i * 4 + 1] << " " << every[i * 4 + 2] << " " << every[i * 4 + 3] << "\n";
        }
    }
    wsg_destroy(h);
    HostProf::report();
    return failed ? -1 : 0;
}

int main(int argc, char* argv[])
{
    std::cout << "wass_stereo  v. " << "1.7_b200-0.2" << std::endl;
    std::cout << "----------------------------------------------" << std::endl;
    std::cout << " [Release] " << wsg_version() << ", OpenCV none" << std::endl << std::endl;
    Config cfg;
    if (argc == 1) {
        print_usage();
        std::cout << "Not enough arguments, aborting." << std::endl;
        return 0;
    }
    if (argc > 1 && std::string("--genconfig") == argv[1]) return save_configuration(cfg, "stereo_config.txt");
    if (argc > 1 && std::string("--batch") == argv[1]) return run_batch(argc, argv);
    if (argc != 3 && argc != 4) { std::cerr << "Invalid arguments" << std::endl; return -1; }
    Env env;
    struct JoinJobs { ~JoinJobs() { g_jpeg.wait(); } } join_before_env_goes;          // background jobs read env
    env.workdir = argv[2];
    if (!exists(env.workdir)) { std::cerr << "\"" << env.workdir << "\" does not exists, aborting." << std::endl; return -1; }
    g_logfile = new std::ofstream(path(env, "wass_stereo_log.txt").c_str());
    LOG_SCOPE("wass_stereo");
    {
        LOGI << "Loading configuration file " << argv[1];
        std::ifstream ifs(argv[1]);
        if (!ifs.is_open()) { LOGE << "Unable to load " << argv[1]; return -1; }
        try { cfg.load(ifs); } catch (std::runtime_error& er) { LOGE << er.what(); return -1; }
        if (save_configuration(cfg, path(env, "stereo_config.txt")) != 0) LOGE << "Unable to save stereo configuration file";
    }
    if (cfg.geti("RANDOM_SEED") == -1) srand((unsigned)time(0));
    else { srand(cfg.geti("RANDOM_SEED")); LOGI << "random seed set to: " << cfg.geti("RANDOM_SEED"); }

    wsg_handle* h = nullptr;
    int device = 0;
    if (const char* e = getenv("WASS_GPU_DEVICE")) device = atoi(e);
    if (wsg_create(device, &h) != WSG_OK) { LOGE << "no usable CUDA device " << device << " (this build has no CPU path)"; return -1; }
    if (stage_load_rectify(env, cfg, h) != 0) return -1;
    if (argc == 4 && std::string("--rectify-only") == argv[3]) { LOGI << "All done."; return 0; }
    if (argc == 4 && std::string("--measure") == argv[3]) { LOGE << "--measure needs the interactive HighGUI point picker; not available in this build"; return -1; }
    // ---- dense stereo (wass_stereo.cpp:764-1020)
    const wsg_dense_params dp = make_dense_params(cfg);
    log_dense_begin(env, dp);
    std::vector<float> disp_roi;
    if (cfg.getb("SAVE_DEBUG_IMAGES")) disp_roi.resize((size_t)env.right_crop.rows * env.right_crop.cols);
    if (wsg_dense_stereo(h, env.left_crop.px.data(), env.right_crop.px.data(), env.right_crop.rows, env.right_crop.cols,
                         env.right_crop.cols, &dp, disp_roi.empty() ? nullptr : disp_roi.data(), nullptr) != WSG_OK) {
        LOGE << "wsg_dense_stereo: " << wsg_last_error(h);
        return -1;
    }
    log_dense_end(env, h);
    double plane[4];
    if (stage_after_dense(env, cfg, h, dp, plane, disp_roi.empty() ? nullptr : disp_roi.data()) != 0) return -1;
    wsg_destroy(h);
    return 0;
}
