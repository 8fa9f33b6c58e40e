// maxloss.h -- getInvHyp / maxLoss with the surface of /root/reference/core/maxloss.h:39-79.
// (dLossMax runs on the device inside dsac_backward; see dsac_b200/csrc/backward.cuh.)
#pragma once
#include <algorithm>

#include "Hypothesis.h"

#define MAXLOSS 10000000.0

inline Hypothesis getInvHyp(const Hypothesis& hyp) {
    cvlite::Matd trans = hyp.getTransformation();
    trans = cvlite::inv(trans);
    return Hypothesis(trans);
}

inline double maxLoss(const Hypothesis& h1, const Hypothesis& h2) {
    Hypothesis invH1 = getInvHyp(h1), invH2 = getInvHyp(h2);
    double rotErr = invH1.calcAngularDistance(invH2);
    double tErr = cvlite::norm(invH1.getTranslation() - invH2.getTranslation());
    return std::min(std::max(rotErr, tErr / 10), MAXLOSS);
}
