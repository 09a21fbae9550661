// oracle/shim/boost/cstdint.hpp -- TEST INFRASTRUCTURE ONLY: boost::uint32_t / uint16_t as used by PovMesh.cpp.
#pragma once
#include <cstdint>
namespace boost { using std::uint8_t; using std::uint16_t; using std::uint32_t; using std::uint64_t; using std::int32_t; }
