// Hypothesis.h -- 6-DoF pose value type with the public surface of the reference's class
// (/root/reference/core/Hypothesis.h:45-243), implemented from scratch on cvlite types.
#pragma once
#include <vector>

#include "types.h"

class Hypothesis {
public:
    Hypothesis();                                                        // identity
    Hypothesis(cvlite::Matd rot, cvlite::Point3d trans);                 // Hypothesis.cpp:38
    Hypothesis(jp::info_t info);                                         // GT pose, centre in m -> mm (Hypothesis.cpp:45)
    explicit Hypothesis(cvlite::Matd transform4x4);                      // Hypothesis.cpp:60
    explicit Hypothesis(std::vector<std::pair<cvlite::Point3d, cvlite::Point3d>> points);  // Kabsch (Hypothesis.cpp:76)
    explicit Hypothesis(std::vector<double> rodVecAndTrans);             // Hypothesis.cpp:81

    void refine(std::vector<std::pair<cvlite::Point3d, cvlite::Point3d>> points);
    void refine(cvlite::Matd& coV, cvlite::Point3d pointsA, cvlite::Point3d pointsB);

    cvlite::Point3d getTranslation() const;
    cvlite::Matd getRotation() const;
    cvlite::Matd getInvRotation() const;
    cvlite::Vec3d getRodriguesVector() const;
    std::vector<double> getRodVecAndTrans() const;
    void setRotation(cvlite::Matd rot);
    void setTranslation(cvlite::Point3d trans);
    cvlite::Matd getTransformation() const;
    Hypothesis getInv();
    Hypothesis operator*(const Hypothesis& other) const;
    Hypothesis operator/(const Hypothesis& other) const;
    cvlite::Point3d transform(cvlite::Point3d p, bool isNormal = false);
    cvlite::Point3d invTransform(cvlite::Point3d p);
    double calcAngularDistance(const Hypothesis& h) const;

    static std::pair<cvlite::Matd, cvlite::Point3d> calcRigidBodyTransform(cvlite::Matd& coV, cvlite::Point3d pointsA,
                                                                          cvlite::Point3d pointsB);

private:
    cvlite::Matd rotation, invRotation;
    cvlite::Point3d translation;
    std::vector<std::pair<cvlite::Point3d, cvlite::Point3d>> points;
    static std::pair<cvlite::Matd, cvlite::Point3d> calcRigidBodyTransform(
        std::vector<std::pair<cvlite::Point3d, cvlite::Point3d>> points);
};
