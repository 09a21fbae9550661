#include "config.hpp"

#include <algorithm>
#include <istream>
#include <sstream>

namespace wasshost {

// incfg keeps a sticky flag (ext/incfg/incfg.hpp:528-535): an option stays "default" only while every assignment repeats
// the value it already has -- "WINSIZE=11" followed by "WINSIZE=13" (the default) is NOT written back commented out
bool Config::Option::is_default() const { return is_def; }

std::string Config::Option::value_str() const
{
    std::stringstream ss;
    switch (type) {
        case INT: ss << i; break;
        case DOUBLE: ss << d; break;                       // default stream precision, as incfg's to_string_helper
        case BOOL: ss << (b ? "true" : "false"); break;
        default: ss << '"' << s << '"'; break;
    }
    return ss.str();
}

void Config::Option::parse(const std::string& v)
{
    if (type == BOOL) {
        if (v != "true" && v != "false") throw ConfigError("Unable to parse " + v + " to \"true\" or \"false\"");
        const bool nb = v == "true";
        is_def = is_def && nb == b;
        b = nb;
        return;
    }
    if (type == STRING) {
        const std::string ns = (v.length() >= 2 && v.front() == '"' && v.back() == '"') ? v.substr(1, v.length() - 2) : v;
        is_def = is_def && ns == s;
        s = ns;
        return;
    }
    std::stringstream ss(v);
    if (type == INT) { int x; ss >> x; if (ss.fail()) throw ConfigError("Unable to parse " + v + " to its defined type"); is_def = is_def && x == i; i = x; }
    else { double x; ss >> x; if (ss.fail()) throw ConfigError("Unable to parse " + v + " to its defined type"); is_def = is_def && x == d; d = x; }
}

void Config::addi(const char* k, int v, const char* d) { Option o; o.type = INT; o.desc = d; o.i = o.i0 = v; options_[k] = o; }
void Config::addd(const char* k, double v, const char* d) { Option o; o.type = DOUBLE; o.desc = d; o.d = o.d0 = v; options_[k] = o; }
void Config::addb(const char* k, bool v, const char* d) { Option o; o.type = BOOL; o.desc = d; o.b = o.b0 = v; options_[k] = o; }
void Config::adds(const char* k, const char* v, const char* d) { Option o; o.type = STRING; o.desc = d; o.s = o.s0 = v; options_[k] = o; }

const Config::Option& Config::opt(const char* k) const
{
    auto it = options_.find(k);
    if (it == options_.end()) throw ConfigError(std::string("internal: unknown key ") + k);
    return it->second;
}

Config::Config()
{
    // src/wass_stereo/wass_stereo.cpp:52-74
    addi("RANDOM_SEED", -1, "Random seed for ransac. -1 to use system timer");
    addi("MIN_TRIANGULATED_POINTS", 100, "Minimum number of triangulated point to proceed with plane estimation");
    addd("SAVE_INPUT_SCALE", 0.3, "Save a scaled version of input images (Set 1 to skip or a value <1 to specify scale ratio)");
    addd("ZGAP_PERCENTILE", 99.0, "Z-gap percentile for outlier filtering");
    addb("DISABLE_AUTO_LEFT_RIGHT", false, "Disable automatic left-right detection");
    addb("SWAP_LEFT_RIGHT", false, "Swaps left-right images (only valid if DISABLE_AUTO_LEFT_RIGHT is set)");
    addb("SAVE_FULL_MESH", false, "Save 3D point cloud before plane outlier removal");
    addi("PLANE_RANSAC_ROUNDS", 400, "number of RANSAC rounds for plane estimation");
    addd("PLANE_RANSAC_THRESHOLD", 1.0, "RANSAC inlier threshold");
    addd("PLANE_REFINE_XMIN", -9999, "Minimum point x-coordinate for plane refinement");
    addd("PLANE_REFINE_XMAX", 9999, "Maximum point x-coordinate for plane refinement");
    addd("PLANE_REFINE_YMIN", -9999, "Minimum point y-coordinate for plane refinement");
    addd("PLANE_REFINE_YMAX", 9999, "Maximum point y-coordinate for plane refinement");
    addd("PLANE_MAX_DISTANCE", 1.5, "Maximum point-plane distance allowed for the reconstructed point-cloud");
    addb("SAVE_AS_PLY", false, "Save final reconstructed point cloud also in PLY format");
    addb("SAVE_COMPRESSED", true, "Save in 16-bit compressed format");
    addb("USE_CUSTOM_STEREORECTIFY", false, "Use built-in stereorectify algorithm instead of the one provided by OpenCV");
    addb("DISABLE_RECTIFY_ROI", false, "Disable automatic ROI computation during stereo rectification (only enabled if USE_CUSTOM_STEREORECTIFY=true)");
    addd("RECTIFY_ANGLE", 0.0, "Additional rotation to apply around the baseline (only enabled if USE_CUSTOM_STEREORECTIFY=true");
    // wass_stereo.cpp:742-761
    addi("MIN_DISPARITY", 1, "Minimum disparity allowed (in px)");
    addi("MAX_DISPARITY", 640, "Maximum disparity allowed");
    addi("WINSIZE", 13, "Stereo match window size");
    addd("DENSE_SCALE", 1.0, "Image resize along epipolar lines before dense stereo");
    addi("DISPARITY_OFFSET", 0, "Offset in pixel to be applied. Positive: move right image to the right. Negative: move right image to the left");
    addi("DISP_DILATE_STEPS", 1, "Number of dilate steps to be applied to the disparity map");
    addi("DISP_EROSION_STEPS", 2, "Number of erosion steps to be applied to the disparity map");
    addi("MEDIAN_FILTER_WSIZE", 0, "Disparity median filter window size (0 to disable)");
    addi("DENSE_P1_MULT", 2, "SGBM P1 parameter");
    addi("DENSE_P2_MULT", 64, "SGBM P2 parameter");
    addi("DENSE_UNIQUENESS_RATIO", 1, "SGBM Uniqueness ratio");
    addi("DENSE_DISP12MAXDIFF", -1, "SGBM Disp12MaxDiff");
    addi("DENSE_PREFILTER_CAP", 60, "SGBM PreFilterCap");
    addi("DENSE_SPECKLE_RANGE", 16, "SGBM SpeckleRange");
    addi("DENSE_SPECKLE_WINDOW_SIZE", -70, "SGBM SpeckleWindowSize");
    addi("DENSE_DISPARITY_BIGGEST_COMPONENT_THRESHOLD", 0, "Maximum squared gradient magnitude threshold for biggest connected component extraction (0 to disable)");
    // wass_stereo.cpp:1030-1037
    addd("TRIANG_MIN_ANGLE", 20.0, "Minimum ray angle for triangulation (in degrees)");
    addd("TRIANG_BBOX_TOP", -1.0, "Triangulation bounding box top coordinate in px wrt. the left image (-1 to disable)");
    addd("TRIANG_BBOX_LEFT", -1.0, "Triangulation bounding box left coordinate in px wrt. the left image (-1 to disable)");
    addd("TRIANG_BBOX_RIGHT", -1.0, "Triangulation bounding box right coordinate in px wrt. the left image (-1 to disable)");
    addd("TRIANG_BBOX_BOTTOM", -1.0, "Triangulation bounding box bottom coordinate in px wrt. the left image (-1 to disable)");
    adds("LEFT_MASK_IMAGE", "none", "Filename of a (BW) left camera mask image. Note: File path is relative to current workdir. Use \"none\" for no mask");
    adds("RIGHT_MASK_IMAGE", "none", "Filename of a (BW) right camera mask image. Note: File path is relative to current workdir. Use \"none\" for no mask");
    addb("DISCARD_BURNED_AREAS", true, "Discard white pixels (value>254)");
    // PovMesh.cpp:577-579
    addb("PLANE_WEIGHT_PROPORTIONAL_TO_DISTANCE", true, "use point to camera distance as weight during LLS plane fitting");
    addb("PLANE_USE_CENTRAL_THIRD_ONLY", false, "use only the central third of the image to estimate the mean sea plane");
    addd("PLANE_REFINEMENT_MAX_DISTANCE", 70.0, "max point distance for plane refinement");
    // extension of this implementation (absent key == reference behaviour)
    addb("SGM_FULL_8PATH", false, "B200 extension: use the 8-path (MODE_HH) aggregation instead of the reference's 5-path MODE_SGBM");
    addb("SAVE_DEBUG_IMAGES", true, "B200 extension: write the diagnostic JPEGs (stereo.jpg, disparity_*.jpg, graph_components.jpg) as the reference always does");
}

void Config::load(std::istream& is)
{
    if (is.fail()) throw ConfigError("IO Error.");
    unsigned linenum = 0;
    std::string buff;
    while (std::getline(is, buff)) {
        ++linenum;
        if (buff.empty() || buff[0] == '#' || buff[0] == '\n' || buff[0] == '\r') continue;
        // remove all spaces before the first and after the last double quote (incfg.cpp:104-116)
        size_t end_idx = buff.find_first_of('"');
        if (end_idx == std::string::npos) end_idx = buff.length();
        buff.erase(std::remove(buff.begin(), buff.begin() + end_idx, ' '), buff.begin() + end_idx);
        size_t start_idx = buff.find_last_of('"');
        if (start_idx == std::string::npos) start_idx = 0;
        buff.erase(std::remove(buff.begin() + start_idx, buff.end(), ' '), buff.end());
        buff.erase(std::remove(buff.begin() + start_idx, buff.end(), '\r'), buff.end());
        const size_t eq = buff.find_first_of('=');
        if (eq == std::string::npos || eq == 0) {
            std::stringstream err;
            err << "Parse error at line " << linenum - 1 << ": No key found (<key> = <value> expected)";
            throw ConfigError(err.str());
        }
        const std::string key = buff.substr(0, eq), value = buff.substr(eq + 1);
        auto it = options_.find(key);
        if (it == options_.end()) throw ConfigError("Unexpected key: " + key);
        try {
            it->second.parse(value);
        } catch (ConfigError& e) {
            std::stringstream err;
            // incfg prints its 0-based counter minus one here, unsigned (incfg.cpp:139): "Line 4294967295" for the first line
            err << "Config file error for key <" << key << "> (Line " << (unsigned)(linenum - 2) << "): " << e.what();
            throw ConfigError(err.str());
        }
    }
}

std::string Config::to_config_string() const
{
    std::stringstream ss;
    for (const auto& kv : options_) {
        if (!kv.second.desc.empty()) ss << "# " << kv.second.desc << std::endl << "# " << std::endl;
        ss << (kv.second.is_default() ? "#" : "") << kv.first << "=" << kv.second.value_str() << std::endl << std::endl;
    }
    return ss.str();
}

}  // namespace wasshost
