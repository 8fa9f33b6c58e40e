// properties.h -- GlobalProperties with the surface of /root/reference/core/properties.h:40-141:
// same parameter structs, same flag abbreviations (-rI, -rRI, -rB, -rSS, -rT2D, ...), config file
// first then command line.  Engine-specific flags are added at the end.
#pragma once
#include <string>
#include <vector>

#include "../../include/dsac_b200.h"
#include "types.h"

struct PoseParameters {
    bool randomDraw;                 // -rdraw
    int ransacIterations;            // -rI   hypotheses per frame
    int ransacRefinementIterations;  // -rRI
    int ransacBatchSize;             // -rB   max inliers per refinement step
    float ransacSubSample;           // -rSS  ratio of pixels with refinement gradients
    float ransacInlierThreshold2D;   // -rT2D px
    float ransacInlierThreshold3D;   // -rT3D mm
};

struct DatasetParameters {
    bool rawData;                    // -rd
    float focalLength;               // -fl
    float xShift, yShift;            // -xs -ys
    float secondaryFocalLength;      // -sfl
    float rawXShift, rawYShift;      // -rxs -rys
    int imageWidth, imageHeight;     // -iw -ih
    std::string objScript, scoreScript;   // -oscript -sscript (kept for CLI compatibility; no Lua here)
    std::string objModel, scoreModel;     // -omodel -smodel
    std::string config;
};

struct EngineParameters {            // not in the reference
    double alpha, beta;              // -alpha -beta  soft-inlier score
    unsigned seed;                   // -seed
    int streams;                     // -streams      sampler streams per frame
    int gpus;                        // -gpus
    int frames;                      // -frames       synthetic frames to process
    int batch;                       // -batch        frames per engine call
    int trajectory;                  // -traj         1 = 7Scenes-like camera path
    double inlierRatio, noise;       // -rho -sigma   synthetic data
};

class GlobalProperties {
protected:
    GlobalProperties();
public:
    PoseParameters pP;
    DatasetParameters dP;
    EngineParameters eP;
    static GlobalProperties* getInstance();
    cvlite::Mat_<float> getCamMat();                       // properties.cpp:308-323
    void parseCmdLine(int argc, const char* argv[]);       // properties.cpp:270-275
    void parseConfig();                                    // properties.cpp:277-306
    bool readArguments(std::vector<std::string> argv);     // properties.cpp:97-268
    dsac_config engineConfig(int maxFrames, int device = 0) const;
private:
    static GlobalProperties* instance;
};
