// cnn_softam.h -- processImage with the shape of the reference's entry point
// (/root/reference/core/cnn_softam.h:960-988), backed by the CUDA engine through the C ABI.
//
// Differences from the reference signature, all forced by the scope of this repository:
//   * the image + coordinate CNN (imgBGR, stateRGB, patches) are replaced by the scene-coordinate grid
//     itself (estObj is an INPUT) and its sampling grid;
//   * the score CNN's lua_State is replaced by the engine handle (the score seam is
//     dsac_set_score_hook, include/dsac_b200.h);
//   * cv:: types are the cvlite stand-ins of types.h.
// Every by-reference output of the reference keeps its name, meaning and layout.
#pragma once
#include <vector>

#include "../../include/dsac_b200.h"
#include "Hypothesis.h"
#include "maxloss.h"
#include "properties.h"

#define CNN_OBJ_PATCHSIZE 40   // lua_calls.h:33
#define CNN_RGB_PATCHSIZE 42   // lua_calls.h:30

// stochasticSubSample (cnn_softam.h:283-309) for frame `frameIndex` of the engine's stream contract
cvlite::Mat_<cvlite::Point2i> stochasticSubSample(int width, int height, unsigned seed);

// Batched form: one engine call for n frames (what the drivers use for throughput).
struct FrameResult {
    double loss = 0, sfEntropy = 0, tErr = 0, rotErr = 0;
    bool correct = false;
    std::vector<jp::cv_trans_t> hyps;
    jp::cv_trans_t refAvgHyp, avgHyp;
    std::vector<std::vector<cvlite::Point2f>> imgPts;
    std::vector<std::vector<cvlite::Point3f>> objPts;
    std::vector<std::vector<int>> imgIdx;
    std::vector<double> sfScores;
    std::vector<std::vector<cvlite::Point2i>> sampledPoints;
    cvlite::Mat_<int> inlierMap;
    std::vector<std::vector<int>> pixelIdxs;
    unsigned status = 0;
};

// Runs the forward pass for n frames and unpacks the outputs.  coords: [n][1600][3] int16,
// pix: [n][1600][2] int32, gtJp: [n][12] or nullptr.  Returns 0 or a DSAC_ERR_* code.
int processImages(dsac_engine* engine, int n, long long frame0, const short* coords, const int* pix,
                  const double* gtJp, std::vector<FrameResult>& results);

// Single-frame form with the reference's parameter list (cnn_softam.h:960-988).
int processImage(dsac_engine* engine, long long frameIndex, const Hypothesis& poseGT, int objHyps, int ptCount,
                 const cvlite::Mat_<float>& camMat, int inlierThreshold2D, int inlierCount, int refSteps, double& loss,
                 double& sfEntropy, bool& correct, std::vector<jp::cv_trans_t>& hyps, jp::cv_trans_t& refAvgHyp,
                 jp::cv_trans_t& avgHyp, std::vector<std::vector<cvlite::Point2f>>& imgPts,
                 std::vector<std::vector<cvlite::Point3f>>& objPts, std::vector<std::vector<int>>& imgIdx,
                 std::vector<double>& sfScores, const jp::img_coord_t& estObj,
                 const cvlite::Mat_<cvlite::Point2i>& sampling, std::vector<std::vector<cvlite::Point2i>>& sampledPoints,
                 cvlite::Mat_<int>& inlierMap, std::vector<std::vector<int>>& pixelIdxs, double& tErr, double& rotErr);
