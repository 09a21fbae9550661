// TEST INFRASTRUCTURE ONLY.  Driver around the REFERENCE's own mesh stage: src/wass_stereo/PovMesh.cpp and
// src/wass_lib/triangulate.hpp, UNMODIFIED, compiled where they lie under /root/reference by oracle/build_ref.sh into
// oracle/_ref/povmesh_ref.  The OpenCV / Boost vocabulary they use comes from oracle/shim/ (this image has neither library's
// C++ headers); the configuration library is the reference's own ext/incfg.  Nothing of the reference is copied into this
// repository: PovMesh.cpp is #included from its own path so that the driver can read the mesh grid back (its state lives in
// a file-local struct), and `private` is opened up for that one translation unit after the standard headers are in.
//
//   povmesh_ref mesh <in.bin> <outdir> <seed> <rounds> <ransac_thr> <zgap_pct> <plane_max_dist> <xmin> <xmax> <ymin> <ymax> [config]
//       runs the mesh part of main() (src/wass_stereo/wass_stereo.cpp:2046-2135) on a point grid:
//       compute_zgap_percentile -> cluster_biggest_connected_component -> srand(seed), ransac_find_plane ->
//       [crop_plane(thr) -> refine_plane -> crop_plane(plane_max_dist)] -> save_as_xyz_compressed / _binary / ply_points
//       in.bin:  int32 W, H; uint8 valid[W*H]; float64 xyz[W*H*3]; uint8 grey[W*H]            (grid order v*W+u)
//       outdir:  result.txt (key value lines, %.17g), mask_component.u8 / mask_crop1.u8 / mask_final.u8 (W*H bytes),
//                mesh_cam.xyzC, mesh_cam.xyzbin, mesh.ply written by the reference's own writers
//       config:  optional incfg file (PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE, PLANE_USE_CENTRAL_THIRD_ONLY, ...)
//   povmesh_ref tri <in.bin> <out.bin>
//       in.bin: int32 n; n x { float64 p[2], q[2], R[9], T[3] };  out.bin: n x float64 xyz[3] of
//       triangulate(p, q, R, T)  (src/wass_lib/triangulate.hpp:26-72)
//   povmesh_ref rt <a> <b> <c> <d>    prints R, T, Rinv, Tinv of PovMesh::RT_from_plane (PovMesh.cpp:1044-1069), %.17g
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <opencv2/opencv.hpp>
#include <boost/filesystem.hpp>
#include <boost/cstdint.hpp>
#include "log.hpp"
#include "incfg.hpp"

#define private public
#include "PovMesh.cpp"          // the reference's file, from -I<reference>/src/wass_stereo
#undef private
#include "triangulate.hpp"      // -I<reference>/src/wass_lib

static void dump_mask(const PovMesh& m, const std::string& path)
{
    std::vector<unsigned char> v(m.pImpl->size());
    for (size_t i = 0; i < v.size(); ++i) v[i] = m.pImpl->PTc(i).valid ? 1 : 0;
    std::ofstream(path, std::ios::binary).write((const char*)v.data(), v.size());
}

static int run_mesh(int argc, char** argv)
{
    if (argc < 13) return 2;
    const std::string in = argv[2], outdir = argv[3];
    const int seed = atoi(argv[4]);
    const size_t rounds = (size_t)atol(argv[5]);
    const double thr = atof(argv[6]), zpct = atof(argv[7]), maxdist = atof(argv[8]);
    const double xmin = atof(argv[9]), xmax = atof(argv[10]), ymin = atof(argv[11]), ymax = atof(argv[12]);
    if (argc > 13) {
        std::ifstream ifs(argv[13]);
        incfg::ConfigOptions::instance().load(ifs);
    }
    std::ifstream f(in, std::ios::binary);
    int32_t W = 0, H = 0;
    f.read((char*)&W, 4); f.read((char*)&H, 4);
    std::vector<unsigned char> valid((size_t)W * H), grey((size_t)W * H);
    std::vector<double> xyz((size_t)W * H * 3);
    f.read((char*)valid.data(), valid.size());
    f.read((char*)xyz.data(), xyz.size() * 8);
    f.read((char*)grey.data(), grey.size());
    if (!f) { std::cerr << "short input" << std::endl; return 3; }

    PovMesh mesh(W, H);
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            const size_t i = (size_t)v * W + u;
            if (valid[i]) mesh.set_point(u, v, cv::Vec3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), grey[i], grey[i], grey[i]);
        }
    std::ofstream res(outdir + "/result.txt");
    res.precision(17);
    res << "n_points " << mesh.pImpl->num_valid_points() << "\n";
    const double zgap = mesh.compute_zgap_percentile(zpct);
    res << "zgap " << zgap << "\n";
    mesh.cluster_biggest_connected_component(boost::filesystem::path(outdir), zgap);
    dump_mask(mesh, outdir + "/mask_component.u8");
    res << "n_component " << mesh.pImpl->num_valid_points() << "\n";
    srand(seed);
    const bool ok = mesh.ransac_find_plane(rounds, thr);
    std::vector<double> pl = mesh.get_plane_params();
    res << "ransac_ok " << (ok ? 1 : 0) << "\n";
    res << "plane_ransac " << pl[0] << " " << pl[1] << " " << pl[2] << " " << pl[3] << "\n";
    if (ok) {
        mesh.crop_plane(thr);
        dump_mask(mesh, outdir + "/mask_crop1.u8");
        std::vector<cv::Vec3d> inliers;
        mesh.refine_plane(xmin, xmax, ymin, ymax, &inliers);
        res << "n_refine_inliers " << inliers.size() << "\n";
        pl = mesh.get_plane_params();
        res << "plane_refined " << pl[0] << " " << pl[1] << " " << pl[2] << " " << pl[3] << "\n";
        mesh.crop_plane(maxdist);
    }
    dump_mask(mesh, outdir + "/mask_final.u8");
    res << "n_final " << mesh.pImpl->num_valid_points() << "\n";
    cv::Matx33d R, Rinv; cv::Vec3d T, Tinv;
    mesh.RT_from_plane(R, T, Rinv, Tinv);
    res << "Rinv";
    for (int i = 0; i < 9; ++i) res << " " << Rinv.val[i];
    res << "\nTinv " << Tinv[0] << " " << Tinv[1] << " " << Tinv[2] << "\n";
    const bool s1 = mesh.save_as_xyz_compressed(outdir + "/mesh_cam.xyzC");
    const bool s2 = mesh.save_as_xyz_binary(outdir + "/mesh_cam.xyzbin");
    const bool s3 = mesh.save_as_ply_points(outdir + "/mesh.ply");
    res << "saved " << s1 << " " << s2 << " " << s3 << "\n";
    return 0;
}

static int run_tri(int argc, char** argv)
{
    if (argc < 4) return 2;
    std::ifstream f(argv[2], std::ios::binary);
    int32_t n = 0;
    f.read((char*)&n, 4);
    std::vector<double> in((size_t)n * 16), out((size_t)n * 3);
    f.read((char*)in.data(), in.size() * 8);
    if (!f) return 3;
    for (int i = 0; i < n; ++i) {
        double* d = &in[(size_t)i * 16];
        cv::Mat R(3, 3, CV_64FC1, d + 4), T(3, 1, CV_64FC1, d + 13);
        const cv::Vec3d p = triangulate(cv::Vec2d(d[0], d[1]), cv::Vec2d(d[2], d[3]), R, T);
        out[3 * i] = p[0]; out[3 * i + 1] = p[1]; out[3 * i + 2] = p[2];
    }
    std::ofstream(argv[3], std::ios::binary).write((const char*)out.data(), out.size() * 8);
    return 0;
}

static int run_rt(int argc, char** argv)
{
    if (argc < 6) return 2;
    cv::Matx33d R, Rinv; cv::Vec3d T, Tinv;
    PovMesh::RT_from_plane(atof(argv[2]), atof(argv[3]), atof(argv[4]), atof(argv[5]), R, T, Rinv, Tinv);
    printf("R"); for (int i = 0; i < 9; ++i) printf(" %.17g", R.val[i]);
    printf("\nT %.17g %.17g %.17g\nRinv", T[0], T[1], T[2]);
    for (int i = 0; i < 9; ++i) printf(" %.17g", Rinv.val[i]);
    printf("\nTinv %.17g %.17g %.17g\n", Tinv[0], Tinv[1], Tinv[2]);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "mesh")) return run_mesh(argc, argv);
    if (!strcmp(argv[1], "tri")) return run_tri(argc, argv);
    if (!strcmp(argv[1], "rt")) return run_rt(argc, argv);
    return 2;
}
