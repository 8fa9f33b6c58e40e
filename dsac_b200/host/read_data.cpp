// read_data.cpp -- see read_data.h.  Follows core/read_data.cpp:69-133 step by step; arithmetic on a float 4x4 like the
// reference's cv::Mat_<float> (products and the inverse are accumulated in double and rounded to float once).
#include "read_data.h"

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

namespace {
std::vector<std::string> split(const std::string& s) {   // util.cpp:43-52: whitespace-separated tokens
    std::istringstream iss(s);
    std::vector<std::string> out;
    std::string tok;
    while (iss >> tok) out.push_back(tok);
    return out;
}
}  // namespace

namespace jp {
bool readData(const std::string infoFile, jp::info_t& info) {
    std::ifstream file(infoFile);
    if (!file.is_open()) {
        info.visible = false;
        return false;
    }
    float trans[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    std::string line;
    for (int i = 0; i < 3; i++) {
        std::getline(file, line);
        std::vector<std::string> tokens = split(line);
        for (int j = 0; j < 4; j++) trans[i][j] = (float)std::atof(j < (int)tokens.size() ? tokens[j].c_str() : "0");
    }
    std::ifstream transFile("translation.txt");   // per-scene offset that centres the scene (read_data.cpp:91-105)
    if (transFile.is_open()) {
        std::getline(transFile, line);
        std::vector<std::string> tokens = split(line);
        for (int j = 0; j < 3 && j < (int)tokens.size(); j++) trans[j][3] = (float)((double)trans[j][3] - std::atof(tokens[j].c_str()));
    } else {
        std::cout << "WARNING! Cannot open translation.txt" << std::endl;
    }
    // 7-Scenes camera frame -> the reference's: columns 1 and 2 change sign (read_data.cpp:107-111), then invert (:113)
    cvlite::Matd M(4, 4);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) M(i, j) = (double)(float)((j == 1 || j == 2) ? -trans[i][j] : trans[i][j]);
    cvlite::Matd Mi = cvlite::inv(M);
    info.rotation = cvlite::Mat_<float>(3, 3);
    for (int y = 0; y < 3; y++)
        for (int x = 0; x < 3; x++) info.rotation(y, x) = (float)Mi(y, x);
    for (int x = 0; x < 3; x++) info.center[x] = (float)Mi(x, 3);
    info.visible = true;
    return true;
}
}  // namespace jp
