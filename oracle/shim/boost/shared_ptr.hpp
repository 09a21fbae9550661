// oracle/shim/boost/shared_ptr.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/shim/opencv2/opencv.hpp).
// StereoMatchEnv (src/wass_stereo/wass_stereo.cpp:333) holds its PovMesh in a boost::shared_ptr; std::shared_ptr has the
// same reset() / operator-> the reference uses.
#pragma once
#include <memory>
namespace boost { template <typename T> using shared_ptr = std::shared_ptr<T>; }
