# Cuts, at build time, the parts of the reference's src/wass_stereo/wass_stereo.cpp that its per-pixel triangulation
# needs -- and nothing else of the 2000-line file, which as a whole needs the real OpenCV / HighGUI:
#   A  the compile-time feature switches (ON / OFF / ENABLED, PLOT_3D_REPROJECTION ..., FIRSTROW)       :93-118
#   B  template struct PointCloud (its nested Point is used as a local)                                 :124-159
#   C  stack_matrices, invert_RT, struct StereoMatchEnv (with unrectify)                                :180-333
#   D  size_t triangulate( StereoMatchEnv& env )                                                        :1039-1386
# The output goes to oracle/_ref/ for the duration of the compile only (oracle/build_ref.sh removes it).
                                     { sub(/\r$/, "") }      # (the reference file has CRLF line ends)
/^#define ON  2-/                    { a = 1 }
/^using namespace nanoflann;/        { a = 0 }
a                                    { print; next }
hold != ""                           { if ($0 ~ /^struct PointCloud/) { b = 1; print hold } hold = "" }
/^template <typename T>$/            { if (!b) { hold = $0; next } }
b                                    { print; if ($0 ~ /^};/) b = 0; next }
/^cv::Mat stack_matrices/            { c = 1 }
c                                    { print; if ($0 ~ /shared_ptr< PovMesh > mesh;/) cend = 1; else if (cend && $0 ~ /^};/) { c = 0; cend = 0 } next }
/^size_t triangulate\( StereoMatchEnv& env \)/ { d = 1 }
d                                    { print; if ($0 ~ /^}/) d = 0; next }
