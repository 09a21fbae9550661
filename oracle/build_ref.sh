#!/bin/bash
# Builds oracle/_ref/incfg_ref: the reference's OWN configuration parser (ext/incfg, two files, no dependencies) with the
# option set of its wass_stereo, as a checker for the drop-in executable's config surface.  Sources are
# compiled where they lie under $REF; outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun).
# The rest of the reference's wass_stereo needs OpenCV C++ and Boost headers and cannot be built here.
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
echo "$HERE/_ref/incfg_ref"

# oracle/_ref/povmesh_ref: the reference's OWN mesh stage -- src/wass_stereo/PovMesh.cpp (z-gap percentile, biggest
# connected component, RANSAC plane, plane refinement, crops, the .xyzC / .xyzbin / PLY writers) and
# src/wass_lib/triangulate.hpp -- unmodified, compiled where they lie, against the header shim in oracle/shim/ (the image
# has no OpenCV C++ / Boost headers) and the reference's own ext/incfg.  It pins oracle/pipeline.py
# (tests/golden/make_povmesh_golden.py -> tests/golden/povmesh_golden.npz -> tests/test_oracle_pipeline.py).
g++ -O1 -std=c++17 -w -I"$HERE/shim" -I"$REF/src/include" -I"$REF/ext/incfg" -I"$REF/src/wass_stereo" -I"$REF/src/wass_lib" \
    "$HERE/povmesh_ref_driver.cpp" "$REF/ext/incfg/incfg.cpp" -o "$HERE/_ref/povmesh_ref"
echo "$HERE/_ref/povmesh_ref"
