// oracle/shim/boost/filesystem.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/shim/opencv2/opencv.hpp).
// The reference's PovMesh.h takes a boost::filesystem::path in one signature and concatenates a file name to it
// (src/wass_stereo/PovMesh.cpp:929, 984); std::filesystem provides the same operations.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem {
using path = std::filesystem::path;
inline bool exists(const path& p) { return std::filesystem::exists(p); }
inline bool create_directories(const path& p) { return std::filesystem::create_directories(p); }
} }
