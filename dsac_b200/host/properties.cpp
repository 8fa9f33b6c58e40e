// properties.cpp -- see properties.h.  Defaults are those of the reference's constructor
// (/root/reference/core/properties.cpp:39-83).
#include "properties.h"

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

GlobalProperties* GlobalProperties::instance = nullptr;

GlobalProperties::GlobalProperties() {
    pP.randomDraw = true;
    pP.ransacIterations = 256;
    pP.ransacRefinementIterations = 8;
    pP.ransacBatchSize = 100;
    pP.ransacSubSample = 0.01f;
    pP.ransacInlierThreshold2D = 10;
    pP.ransacInlierThreshold3D = 100;
    dP.rawData = true;
    dP.focalLength = 525;
    dP.xShift = dP.yShift = 0;
    dP.secondaryFocalLength = 585;
    dP.rawXShift = dP.rawYShift = 0;
    dP.imageWidth = 640;
    dP.imageHeight = 480;
    dP.objScript = "train_obj.lua";
    dP.scoreScript = "train_score.lua";
    dP.objModel = "obj_model_init.net";
    dP.scoreModel = "score_model_init.net";
    dP.config = "default";
    eP.alpha = 0.1; eP.beta = 0.5; eP.seed = 1305; eP.streams = 1; eP.gpus = 1;
    eP.frames = 1000; eP.batch = 128; eP.trajectory = 1; eP.inlierRatio = 0.5; eP.noise = 25.0;
}

GlobalProperties* GlobalProperties::getInstance() {
    if (!instance) instance = new GlobalProperties();
    return instance;
}

bool GlobalProperties::readArguments(std::vector<std::string> argv) {
    struct Opt { const char* flag; const char* what; int kind; void* dst; };   // kind 0 int, 1 float, 2 bool, 3 string, 4 double, 5 unsigned
    const Opt opts[] = {
        {"-iw", "image width", 0, &dP.imageWidth}, {"-ih", "image height", 0, &dP.imageHeight},
        {"-fl", "focal length", 1, &dP.focalLength}, {"-xs", "x shift", 1, &dP.xShift}, {"-ys", "y shift", 1, &dP.yShift},
        {"-rd", "raw data (rescale rgb)", 2, &dP.rawData}, {"-sfl", "secondary focal length", 1, &dP.secondaryFocalLength},
        {"-rxs", "raw x shift", 1, &dP.rawXShift}, {"-rys", "raw y shift", 1, &dP.rawYShift},
        {"-rdraw", "random draw", 2, &pP.randomDraw}, {"-oscript", "object script", 3, &dP.objScript},
        {"-sscript", "score script", 3, &dP.scoreScript}, {"-omodel", "object model", 3, &dP.objModel},
        {"-smodel", "score model", 3, &dP.scoreModel}, {"-rT2D", "ransac inlier threshold 2D", 1, &pP.ransacInlierThreshold2D},
        {"-rT3D", "ransac inlier threshold 3D", 1, &pP.ransacInlierThreshold3D},
        {"-rRI", "ransac refinement iterations", 0, &pP.ransacRefinementIterations},
        {"-rI", "ransac iterations", 0, &pP.ransacIterations}, {"-rB", "ransac batch size", 0, &pP.ransacBatchSize},
        {"-rSS", "ransac refinement gradient sub sampling", 1, &pP.ransacSubSample},
        {"-alpha", "soft inlier alpha", 4, &eP.alpha}, {"-beta", "soft inlier beta", 4, &eP.beta},
        {"-seed", "sampler seed", 5, &eP.seed}, {"-streams", "sampler streams", 0, &eP.streams}, {"-gpus", "gpus", 0, &eP.gpus},
        {"-frames", "synthetic frames", 0, &eP.frames}, {"-batch", "frames per call", 0, &eP.batch},
        {"-traj", "trajectory mode", 0, &eP.trajectory}, {"-rho", "inlier ratio", 4, &eP.inlierRatio},
        {"-sigma", "inlier noise mm", 4, &eP.noise},
    };
    const int argc = (int)argv.size();
    for (int i = 0; i < argc; i++) {
        const std::string& s = argv[i];
        bool known = false;
        for (const Opt& o : opts) {
            if (s != o.flag) continue;
            known = true;
            if (++i >= argc) { std::cout << "missing value for " << s << "\n"; return false; }
            const char* v = argv[i].c_str();
            switch (o.kind) {
                case 0: *(int*)o.dst = std::atoi(v); break;
                case 1: *(float*)o.dst = (float)std::atof(v); break;
                case 2: *(bool*)o.dst = std::atoi(v) != 0; break;
                case 3: *(std::string*)o.dst = v; break;
                case 4: *(double*)o.dst = std::atof(v); break;
                case 5: *(unsigned*)o.dst = (unsigned)std::strtoul(v, nullptr, 10); break;
            }
            std::cout << o.what << ": " << v << "\n";
            break;
        }
        if (!known) { std::cout << "unkown argument: " << s << "\n"; return false; }
    }
    return true;   // (the reference falls off the end here: SURVEY.md quirk Q11)
}

void GlobalProperties::parseCmdLine(int argc, const char* argv[]) {
    std::vector<std::string> v;
    for (int i = 1; i < argc; i++) v.push_back(argv[i]);
    readArguments(v);
}

void GlobalProperties::parseConfig() {
    std::string configFile = dP.config + ".config";
    std::cout << "Parsing config file: " << configFile << std::endl;
    std::ifstream file(configFile);
    if (!file.is_open()) return;
    std::vector<std::string> v;
    std::string line;
    while (std::getline(file, line)) {
        if (line.empty() || line[0] == '#') continue;
        std::istringstream is(line);
        std::string key, val;
        if (!(is >> key >> val)) continue;
        v.push_back("-" + key);
        v.push_back(val);
    }
    readArguments(v);
}

cvlite::Mat_<float> GlobalProperties::getCamMat() {
    float centerX = dP.imageWidth / 2 + dP.xShift, centerY = dP.imageHeight / 2 + dP.yShift, f = dP.focalLength;
    cvlite::Mat_<float> m = cvlite::Mat_<float>::zeros(3, 3);
    m(0, 0) = f; m(1, 1) = f; m(2, 2) = 1.f; m(0, 2) = centerX; m(1, 2) = centerY;
    return m;
}

dsac_config GlobalProperties::engineConfig(int maxFrames, int device) const {
    dsac_config c;
    dsac_default_config(&c);
    c.focal = dP.focalLength;
    c.cx = dP.imageWidth / 2 + dP.xShift;
    c.cy = dP.imageHeight / 2 + dP.yShift;
    c.n_hyps = pP.ransacIterations;
    c.thr2d = (int)pP.ransacInlierThreshold2D;     // int truncation as in test_ransac_softam.cpp:51
    c.inlier_count = pP.ransacBatchSize;
    c.ref_steps = pP.ransacRefinementIterations;
    c.sub_sample = pP.ransacSubSample;
    c.alpha = eP.alpha; c.beta = eP.beta; c.seed = eP.seed; c.n_streams = eP.streams;
    c.max_frames = maxFrames;
    c.device = device;
    return c;
}
