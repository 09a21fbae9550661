// TEST INFRASTRUCTURE ONLY.  Driver around the REFERENCE's own configuration library (ext/incfg/incfg.{hpp,cpp}, compiled
// in place from /root/reference by oracle/build_ref.sh into oracle/_ref/incfg_ref) with the option set of the reference's
// wass_stereo.  keys.inc is GENERATED into oracle/_ref/ by build_ref.sh from the INCFG_REQUIRE lines of
// src/wass_stereo/{wass_stereo,PovMesh}.cpp -- nothing of the reference is copied into this repository.
//
//   incfg_ref --genconfig            prints ConfigOptions::to_config_string() of the defaults (what `wass_stereo --genconfig`
//                                    writes, wass_stereo.cpp:1776-1795)
//   incfg_ref <config_file>          loads the file as wass_stereo.cpp:1836-1846 does; exit 0 and the resulting
//                                    to_config_string() on success, exit 255 and "ERROR: <what>" on a load error
#include "incfg.hpp"
#include <cstring>
#include <fstream>
#include <iostream>

#include "keys.inc"

int main(int argc, char* argv[])
{
    if (argc == 2 && std::strcmp(argv[1], "--genconfig") == 0) {
        std::cout << incfg::ConfigOptions::instance().to_config_string();
        return 0;
    }
    if (argc != 2) return 2;
    try {
        std::ifstream ifs(argv[1]);
        if (!ifs.is_open()) { std::cout << "ERROR: cannot open" << std::endl; return 255; }
        incfg::ConfigOptions::instance().load(ifs);
    } catch (std::runtime_error& er) {
        std::cout << "ERROR: " << er.what() << std::endl;
        return 255;
    }
    std::cout << incfg::ConfigOptions::instance().to_config_string();
    return 0;
}
