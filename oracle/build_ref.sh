#!/bin/bash
# Builds the checkers that are the reference's OWN code (TEST INFRASTRUCTURE): incfg_ref, povmesh_ref, filters_ref, triang_ref.
# Sources are compiled where they lie under $REF -- whole files where they are self-contained (ext/incfg, PovMesh.cpp,
# triangulate.hpp, hires_timer.cpp), cut ranges of wass_stereo.cpp where the file as a whole needs the real OpenCV (the cuts
# exist only for the duration of the compile) -- against the header shim in oracle/shim/; binaries go to oracle/_ref/ only
# (git-ignored, shipped to the GPU box by gpurun).
#
# oracle/_ref/incfg_ref: the reference's OWN configuration parser (ext/incfg, two files, no dependencies) with the option
# set of its wass_stereo, as a checker for the drop-in executable's config surface.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -f "$REF/ext/incfg/incfg.cpp" ] || { echo "no reference tree at $REF: keeping the prebuilt oracle/_ref (if any)"; exit 0; }
mkdir -p "$HERE/_ref"
# the option declarations of the default build: WASS_ENABLE_OPTFLOW is never defined (its add_definitions line is commented
# out in src/wass_stereo/CMakeLists.txt:7), so the declarations inside that #ifdef are skipped
awk '/^#ifdef WASS_ENABLE_OPTFLOW/ {skip=1} /^#endif/ {if (skip) {skip=0; next}} !skip && /^INCFG_REQUIRE/' \
    "$REF/src/wass_stereo/wass_stereo.cpp" "$REF/src/wass_stereo/PovMesh.cpp" > "$HERE/_ref/keys.inc"
g++ -O1 -std=c++14 -I"$REF/ext/incfg" -I"$HERE/_ref" "$HERE/incfg_ref_driver.cpp" "$REF/ext/incfg/incfg.cpp" -o "$HERE/_ref/incfg_ref"
rm -f "$HERE/_ref/keys.inc"         # (cut reference text exists only for the duration of the compile)
echo "$HERE/_ref/incfg_ref"

# oracle/_ref/povmesh_ref: the reference's OWN mesh stage -- src/wass_stereo/PovMesh.cpp (z-gap percentile, biggest
# connected component, RANSAC plane, plane refinement, crops, the .xyzC / .xyzbin / PLY writers) and
# src/wass_lib/triangulate.hpp -- unmodified, compiled where they lie, against the header shim in oracle/shim/ (the image
# has no OpenCV C++ / Boost headers) and the reference's own ext/incfg.  It pins oracle/pipeline.py
# (tests/golden/make_povmesh_golden.py -> tests/golden/povmesh_golden.npz -> tests/test_oracle_pipeline.py).
g++ -O1 -std=c++17 -w -I"$HERE/shim" -I"$REF/src/include" -I"$REF/ext/incfg" -I"$REF/src/wass_stereo" -I"$REF/src/wass_lib" \
    "$HERE/povmesh_ref_driver.cpp" "$REF/ext/incfg/incfg.cpp" -o "$HERE/_ref/povmesh_ref"
echo "$HERE/_ref/povmesh_ref"

# oracle/_ref/filters_ref: the reference's OWN disparity clean-up functions (matrix_dilate_zero, matrix_erode_zero,
# clean_and_convert_disparity: src/wass_stereo/wass_stereo.cpp:617-733).  wass_stereo.cpp as a whole needs the real OpenCV,
# but these three only touch cv::Mat: they are cut out of the reference source here, at build time, into
# oracle/_ref/filters.inc (first "template <typename Mat_T>" up to the DENSE STEREO banner) and compiled against the shim.
awk '/^template <typename Mat_T>/ {on=1} /^\/\*{20,}/ {if (on) exit} on' "$REF/src/wass_stereo/wass_stereo.cpp" > "$HERE/_ref/filters.inc"
grep -q "clean_and_convert_disparity" "$HERE/_ref/filters.inc"
g++ -O1 -std=c++17 -w -I"$HERE/shim" -I"$HERE/_ref" "$HERE/filters_ref_driver.cpp" -o "$HERE/_ref/filters_ref"
rm -f "$HERE/_ref/filters.inc"      # no reference source text is kept, not even under the git-ignored _ref/
echo "$HERE/_ref/filters_ref"

# oracle/_ref/triang_ref: the reference's OWN per-pixel triangulation -- size_t triangulate( StereoMatchEnv& ) with
# StereoMatchEnv::unrectify (src/wass_stereo/wass_stereo.cpp:299-324, 1039-1386) -- cut out of the reference source by
# oracle/cut_triangulate.awk for the duration of the compile, with its PovMesh.cpp / triangulate.hpp / hires_timer / incfg
# from their own paths, against the header shim.
awk '{ sub(/\r$/, "") } /^#ifdef WASS_ENABLE_OPTFLOW/ {skip=1} /^#endif/ {if (skip) {skip=0; next}} !skip && /^INCFG_REQUIRE/' \
    "$REF/src/wass_stereo/wass_stereo.cpp" > "$HERE/_ref/keys_ws.inc"
awk -f "$HERE/cut_triangulate.awk" "$REF/src/wass_stereo/wass_stereo.cpp" > "$HERE/_ref/triang_env.inc"
grep -q "size_t triangulate( StereoMatchEnv& env )" "$HERE/_ref/triang_env.inc"
g++ -O1 -std=c++17 -w -I"$HERE/shim" -I"$HERE/_ref" -I"$REF/src/include" -I"$REF/ext/incfg" -I"$REF/src/wass_stereo" -I"$REF/src/wass_lib" \
    "$HERE/triang_ref_driver.cpp" "$REF/ext/incfg/incfg.cpp" "$REF/src/wass_lib/hires_timer.cpp" -o "$HERE/_ref/triang_ref"
rm -f "$HERE/_ref/triang_env.inc" "$HERE/_ref/keys_ws.inc"
echo "$HERE/_ref/triang_ref"
